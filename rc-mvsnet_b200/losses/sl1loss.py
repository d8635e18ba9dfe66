"""Drop-in for the reference's losses/sl1loss.py: SL1Loss = 0.5 * masked mean smooth-L1 (:4-13), one fused
value+gradient call on the CUDA path."""
import torch.nn as nn

from .aug_loss import masked_loss

__all__ = ["SL1Loss"]


class SL1Loss(nn.Module):
    def forward(self, depth_pred, depth_gt, mask=None):
        p = depth_pred.reshape(1, 1, -1)
        g = depth_gt.reshape(1, 1, -1)
        if mask is None:
            return masked_loss(p, g, None, mask_kind=2, mode=0, scale=0.5)
        return masked_loss(p, g, mask.reshape(1, 1, -1).float(), mask_kind=1, mode=0, scale=0.5)
