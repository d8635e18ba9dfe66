"""Drop-in `models` package: same import surface as the reference's models/__init__.py for the hot path."""
from .modules import *            # noqa: F401,F403
from .render_utils import *       # noqa: F401,F403
from .renderer import *           # noqa: F401,F403
from .render_models import *      # noqa: F401,F403
from .casmvsnet import *          # noqa: F401,F403
from .casmvsnet import CascadeMVSNet, CascadeMVSNet_eval, DepthNet, DepthNet_eval  # noqa: F401
from .render_consist_net import Rendering_Consistency_Net  # noqa: F401
