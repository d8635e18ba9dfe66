"""Drop-in for the reference's losses/unsup_loss.py (UnSupLoss :9-94 and its _06 / _07 / _no_smooth variants,
UnsupLossMultiStage* :350-480): same class names, forward signatures, attributes (`reconstr_loss`, `ssim_loss`,
`smooth_loss`, `unsup_loss`) and returned scalars.  One stage = ONE C call (csrc/unsup_loss.cu) that produces the value
and the gradient w.r.t. depth together; backward() only scales the stored gradient."""
import torch
import torch.nn as nn

from rcmvs import ops

__all__ = ["UnSupLoss", "UnSupLoss_no_smooth", "UnSupLoss_07", "UnSupLoss_06", "UnsupLossMultiStage",
           "UnsupLossMultiStage_06", "UnsupLossMultiStage_07", "UnsupLossMultiStage_no_smooth"]

_STEP = {0: 4, 1: 2, 2: 1}


class _StageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, imgs_nhwc, cams, smooth_weight):
        need = depth.requires_grad
        terms, grad = ops.unsup_loss_fwd_bwd(imgs_nhwc, cams, depth.detach().float().contiguous(), smooth_weight, need)
        ctx.grad = grad
        ctx.mark_non_differentiable(terms)
        return terms[0].clone(), terms

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        return (ctx.grad * g_total if ctx.grad is not None else None), None, None, None


class UnSupLoss(nn.Module):
    smooth_weight = 0.18          # reference :91; the variants below differ only here (:185 none, :279 0.19, :373 0.16)

    def forward(self, imgs, cams, depth, stage_idx):
        """imgs [B,V,3,H,W] (full resolution), cams [B,V,2,4,4] (this stage's scale), depth [B,h,w], stage_idx 0/1/2."""
        assert imgs.shape[1] == cams.shape[1], "Different number of images and projection matrices"
        B, V, _, H, W = imgs.shape
        step = _STEP.get(stage_idx, 1)
        small = ops.images_subsample_nhwc(imgs.reshape(B * V, 3, H, W).float().contiguous(), step)
        small = small.view(B, V, H // step, W // step, 3)
        total, terms = _StageLoss.apply(depth, small, cams.float().contiguous(), self.smooth_weight)
        self.reconstr_loss, self.ssim_loss, self.smooth_loss = terms[1], terms[2], terms[3]
        self.unsup_loss = total
        return total


class UnSupLoss_no_smooth(UnSupLoss):
    smooth_weight = 0.0

    def forward(self, imgs, cams, depth, stage_idx):
        total = super().forward(imgs, cams, depth, stage_idx)
        del self.smooth_loss          # the reference variant never sets it (:117)
        return total


class UnSupLoss_07(UnSupLoss):
    smooth_weight = 0.19


class UnSupLoss_06(UnSupLoss):
    smooth_weight = 0.16


class UnsupLossMultiStage(nn.Module):
    loss_cls = UnSupLoss

    def __init__(self):
        super().__init__()
        self.unsup_loss = self.loss_cls()

    def forward(self, inputs, imgs, cams, **kwargs):
        weights = kwargs.get("dlossw", None)
        total = torch.zeros((), dtype=torch.float32, device=imgs.device)
        scalars = {}
        for key in [k for k in inputs.keys() if "stage" in k]:
            idx = int(key.replace("stage", "")) - 1
            loss = self.unsup_loss(imgs, cams[key], inputs[key]["depth"], idx)
            total = total + (weights[idx] if weights is not None else 1.0) * loss
            scalars["depth_loss_stage%d" % (idx + 1)] = loss
            scalars["reconstr_loss_stage%d" % (idx + 1)] = self.unsup_loss.reconstr_loss
            scalars["ssim_loss_stage%d" % (idx + 1)] = self.unsup_loss.ssim_loss
            if hasattr(self.unsup_loss, "smooth_loss"):
                scalars["smooth_loss_stage%d" % (idx + 1)] = self.unsup_loss.smooth_loss
        return total, scalars


class UnsupLossMultiStage_06(UnsupLossMultiStage):
    loss_cls = UnSupLoss_06


class UnsupLossMultiStage_07(UnsupLossMultiStage):
    loss_cls = UnSupLoss_07


class UnsupLossMultiStage_no_smooth(UnsupLossMultiStage):
    loss_cls = UnSupLoss_no_smooth
