"""Drop-in `losses` package: the hot-path entry (inverse_warping) runs on librcmvs; the remaining loss terms of the
reference (SSIM, smoothness, top-k photometric, augmentation, smooth-L1: losses/unsup_loss.py, modules.py, aug_loss.py,
sl1loss.py) are small 2-D torch expressions and stay in the reference tree ("next")."""
from .homography import inverse_warping  # noqa: F401

# Merge with a same-named package further down sys.path (the reference tree): `losses.unsup_loss`, `losses.aug_loss`,
# `losses.sl1loss`, `losses.modules` keep resolving there, and their `from .homography import *` picks up this one.
import pkgutil as _pkgutil
__path__ = _pkgutil.extend_path(__path__, __name__)
