"""fdt_b200 -- B200-native (sm_100a) SSD box pipeline: PriorBox, Detect, MultiBoxLoss, IoU tracker.

Drop-in for the `layers` package, `utils.calc_performance.calculate_iou` and the tracker loop of
limacv/Face-detection-and-tracking.  Compute runs in hand-written CUDA kernels behind the C ABI of
`csrc/libfdt_b200.so` (include/fdt_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import data, layers, synth, tracker  # noqa: E402,F401
from .layers import Detect, MultiBoxLoss, PriorBoxLayer  # noqa: E402,F401
