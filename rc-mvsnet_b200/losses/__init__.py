"""Drop-in `losses` package (reference losses/): inverse_warping (homography.py), the unsupervised photometric loss
stack (unsup_loss.py — SSIM, gradient smooth-L1, top-1 over views and depth smoothness fused into one value+gradient
call per stage), the augmentation-consistency loss (aug_loss.py) and SL1Loss (sl1loss.py), all on librcmvs kernels."""
from .homography import inverse_warping  # noqa: F401
from .unsup_loss import *  # noqa: F401,F403
from .aug_loss import random_image_mask, AugLossMultiStage, masked_loss  # noqa: F401
from .sl1loss import SL1Loss  # noqa: F401
