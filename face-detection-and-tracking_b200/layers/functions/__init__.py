from .detection import Detect, PendingDetections
from .prior_box import PriorBoxLayer
from .heads import heads_to_loc_conf

__all__ = ['Detect', 'PendingDetections', 'PriorBoxLayer', 'heads_to_loc_conf']
