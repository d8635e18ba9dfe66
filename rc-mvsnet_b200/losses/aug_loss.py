"""Drop-in for the reference's losses/aug_loss.py: random_image_mask (:7-26, host RNG + one masked multiply) and
AugLossMultiStage (:29-66): per stage a masked smooth-L1 between the stage depth and the nearest-down-sampled pseudo
depth — one fused value+gradient call per stage (csrc/unsup_loss.cu masked_loss kernels), no boolean-index gathers."""
import numpy as np
import torch
import torch.nn as nn

from rcmvs import ops

__all__ = ["random_image_mask", "AugLossMultiStage"]


def random_image_mask(img, filter_size):
    """img [B,3,H,W] -> (img with a random fh x fw rectangle zeroed, filter_mask [B,3,H,W]); same RNG call order as the
    reference (np.random.randint for x then y)."""
    fh, fw = filter_size
    _, _, h, w = img.size()
    if fh == h and fw == w:
        return img, None
    x = np.random.randint(0, w - fw)
    y = np.random.randint(0, h - fh)
    filter_mask = torch.ones_like(img)
    filter_mask[:, :, y:y + fh, x:x + fw] = 0.0
    return img * filter_mask, filter_mask


class _MaskedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, mask_kind, mode, scale, step):
        loss, grad = ops.masked_loss_fwd_bwd(pred.detach().float().contiguous(), target, mask, mask_kind, mode, scale, step,
                                             pred.requires_grad)
        ctx.grad = grad
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None, None, None, None


def masked_loss(pred, target, mask=None, mask_kind=0, mode=0, scale=1.0, step=1):
    """Differentiable (w.r.t. pred) scale * masked mean of smooth-L1 (mode 0) or squared error (mode 1)."""
    if mask is not None:
        mask = mask.float().contiguous()
    return _MaskedLoss.apply(pred, target.detach().float().contiguous(), mask, mask_kind, mode, scale, step)


class AugLossMultiStage(nn.Module):
    def forward(self, inputs, pseudo_depth, mask_ms, filter_mask, **kwargs):
        weights = kwargs.get("dlossw", None)
        total = torch.zeros((), dtype=torch.float32, device=pseudo_depth.device)
        scalars = {}
        for key in [k for k in inputs.keys() if "stage" in k]:
            idx = int(key.replace("stage", "")) - 1
            step = {0: 4, 1: 2}.get(idx, 1)
            # F.interpolate(nearest, 1/step) of pseudo depth and filter mask == reading them at stride `step` (:44-52)
            loss = masked_loss(inputs[key]["depth"], pseudo_depth, filter_mask, mask_kind=1, mode=0, step=step)
            total = total + (weights[idx] if weights is not None else 1.0) * loss
            scalars["aug_loss_stage%d" % (idx + 1)] = loss
        return total, scalars
