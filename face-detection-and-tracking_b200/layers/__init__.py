"""Drop-in for the reference's `layers` package (layers/__init__.py:1-2)."""
from .functions import *  # noqa: F401,F403
from .modules import *  # noqa: F401,F403
from . import box_utils  # noqa: F401
