"""Drop-in for the reference's losses/homography.py: inverse_warping(img, left_cam, right_cam, depth) with the same
signature and return values (losses/homography.py:6-62), one CUDA kernel each for forward and for the gradient w.r.t.
depth (the quantity the unsupervised loss differentiates)."""
import torch

from rcmvs import ops


class _InverseWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, left_cam, right_cam, depth):
        if any(ctx.needs_input_grad[:3]):
            # the reference's expression is differentiable w.r.t. img and the cameras too; no caller uses that (the loss
            # warps input images), so only the depth gradient has a kernel: refuse instead of returning zeros
            raise NotImplementedError("rcmvs: inverse_warping is differentiable w.r.t. depth only")
        warped, mask = ops.inverse_warp_fwd(img, left_cam, right_cam, depth)
        ctx.save_for_backward(img, left_cam, right_cam, depth)
        ctx.mark_non_differentiable(mask)
        return warped, mask

    @staticmethod
    def backward(ctx, grad_warped, _grad_mask):
        img, left_cam, right_cam, depth = ctx.saved_tensors
        gd = ops.inverse_warp_bwd(img, left_cam, right_cam, depth, grad_warped.contiguous())
        return None, None, None, gd


def inverse_warping(img, left_cam, right_cam, depth):
    """img [B,h,w,C] (NHWC), left_cam/right_cam [B,2,4,4], depth [B,h,w] -> (warped [B,h,w,C], mask [B,h,w,1])."""
    B, h, w, _ = img.shape
    return _InverseWarp.apply(img.float().contiguous(), left_cam.float().contiguous(), right_cam.float().contiguous(),
                              depth.float().reshape(B, h, w).contiguous())
