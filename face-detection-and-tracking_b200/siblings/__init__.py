"""Sibling decode + NMS implementations of the reference on the same kernels:
FaceBoxes' DataEncoder (FACEBOX/encoderl.py) and MTCNN's nms helpers (MTCNN/mtcnn/core/utils.py, core/nms.py)."""
from . import faceboxes, mtcnn  # noqa: F401
