from .detection import Detect
from .prior_box import PriorBoxLayer
from .heads import heads_to_loc_conf

__all__ = ['Detect', 'PriorBoxLayer', 'heads_to_loc_conf']
