"""Drop-in `models` package: same import surface as the reference's models/__init__.py for the hot path."""
from .modules import *          # noqa: F401,F403
from .casmvsnet import *        # noqa: F401,F403
from .casmvsnet import CascadeMVSNet, CascadeMVSNet_eval, DepthNet, DepthNet_eval  # noqa: F401
