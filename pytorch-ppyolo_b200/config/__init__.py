"""Config entry points, same names as the reference's ``config`` package (config/__init__.py:10-16)."""
from . import get_model, ppyolo_2x, ppyolo_r18vd  # noqa: F401
from .get_model import *  # noqa: F401,F403
from .ppyolo_2x import *  # noqa: F401,F403
from .ppyolo_r18vd import *  # noqa: F401,F403
