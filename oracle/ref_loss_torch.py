"""CPU oracle for row f2 (unsupervised photometric loss stack).  TEST INFRASTRUCTURE ONLY.

Differentiable torch-CPU restatement of the reference's UnSupLoss.forward for one stage (losses/unsup_loss.py:9-94):
inverse_warping (losses/homography.py:6-200, incl. its quirks: source projected with K_left, weights from CLAMPED
corners, asymmetric y mask), compute_reconstr_loss(simple=False) (losses/modules.py:72-82), SSIM (:6-41),
depth_smoothness (:56-69) and the top-1-over-views selection (:78-90).  torch autograd of this restatement is the
checker for the CUDA gradient w.r.t. depth.  Pinned by tests/test_oracle_loss_golden.py against
tests/golden/unsup_loss_small.npz (value and gradient produced by the unmodified reference).
"""
import torch
import torch.nn.functional as F


def inverse_warping(img, left_cam, right_cam, depth):
    """img [B,h,w,C]; cams [B,2,4,4]; depth [B,h,w] -> warped [B,h,w,C], mask [B,h,w,1]."""
    B, h, w, C = img.shape
    Rl, Rr = left_cam[:, 0, :3, :3], right_cam[:, 0, :3, :3]
    tl, tr = left_cam[:, 0, :3, 3:4], right_cam[:, 0, :3, 3:4]
    Kl = left_cam[:, 1, :3, :3]
    R_rel = Rr @ Rl.transpose(1, 2)
    t_rel = tr - R_rel @ tl
    xs = (torch.linspace(-1.0, 1.0, w) + 1.0) * 0.5 * (w - 1)
    ys = (torch.linspace(-1.0, 1.0, h) + 1.0) * 0.5 * (h - 1)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    grid = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.ones(h * w)])[None].expand(B, -1, -1)
    cam = (torch.inverse(Kl) @ grid) * depth.reshape(B, 1, h * w)
    P = Kl @ torch.cat([R_rel, t_rel], 2)                                      # [B,3,4]: K_LEFT on purpose (:55-57)
    pc = P[:, :, :3] @ cam + P[:, :, 3:4]
    px = pc[:, 0] / (pc[:, 2] + 1e-10)
    py = pc[:, 1] / (pc[:, 2] + 1e-10)
    x = ((px / (w - 1) * 2.0 - 1.0) + 1.0) * (w - 1.0) / 2.0
    y = ((py / (h - 1) * 2.0 - 1.0) + 1.0) * (h - 1.0) / 2.0
    x, y = x.reshape(-1), y.reshape(-1)
    x0 = torch.floor(x).int()
    y0 = torch.floor(y).int()
    x1, y1 = x0 + 1, y0 + 1
    mask = ((x0 >= 0) & (x1 <= w - 1) & (y0 >= 0) & (y0 <= h - 1)).float()
    x0, x1 = x0.clamp(0, w - 1), x1.clamp(0, w - 1)
    y0, y1 = y0.clamp(0, h - 1), y1.clamp(0, h - 1)
    base = (torch.arange(B) * (h * w)).reshape(-1, 1).repeat(1, h * w).reshape(-1)
    flat = img.reshape(-1, C)
    pa, pb = flat[base + y0.long() * w + x0.long()], flat[base + y1.long() * w + x0.long()]
    pc_, pd = flat[base + y0.long() * w + x1.long()], flat[base + y1.long() * w + x1.long()]
    dx, dy = x1.float() - x, y1.float() - y
    out = (dx * dy)[:, None] * pa + (dx * (1 - dy))[:, None] * pb + ((1 - dx) * dy)[:, None] * pc_ + \
          ((1 - dx) * (1 - dy))[:, None] * pd
    return out.reshape(B, h, w, C), mask.reshape(B, h, w, 1)


def reconstr_loss(warped, ref, mask):
    """compute_reconstr_loss(simple=False): 0.5 * photo + 0.5 * gradient smooth-L1, all mean-reduced."""
    a, b = warped * mask, ref * mask
    photo = F.smooth_l1_loss(a, b, reduction="mean")
    gx = F.smooth_l1_loss(a[:, :, 1:] - a[:, :, :-1], b[:, :, 1:] - b[:, :, :-1], reduction="mean")
    gy = F.smooth_l1_loss(a[:, 1:] - a[:, :-1], b[:, 1:] - b[:, :-1], reduction="mean")
    return 0.5 * photo + 0.5 * (gx + gy)


def ssim(x, y, mask):
    """losses/modules.py SSIM.forward: 3x3 valid average pools, masked, clamp((1 - SSIM)/2, 0, 1); NHWC in/out."""
    x, y, mask = (t.permute(0, 3, 1, 2) for t in (x, y, mask))
    pool = lambda t: F.avg_pool2d(t, 3, 1)
    mu_x, mu_y = pool(x), pool(y)
    sx, sy, sxy = pool(x * x) - mu_x ** 2, pool(y * y) - mu_y ** 2, pool(x * y) - mu_x * mu_y
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    n = (2 * mu_x * mu_y + c1) * (2 * sxy + c2)
    d = (mu_x ** 2 + mu_y ** 2 + c1) * (sx + sy + c2)
    return (pool(mask) * torch.clamp((1 - n / d) / 2, 0, 1)).permute(0, 2, 3, 1)


def depth_smoothness(depth, img):
    """depth [B,h,w,1], img [B,h,w,3] (losses/modules.py:56-69, lambda = 1)."""
    ddx, ddy = depth[:, :, :-1] - depth[:, :, 1:], depth[:, :-1] - depth[:, 1:]
    idx, idy = img[:, :, :-1] - img[:, :, 1:], img[:, :-1] - img[:, 1:]
    wx = torch.exp(-idx.abs().mean(3, keepdim=True))
    wy = torch.exp(-idy.abs().mean(3, keepdim=True))
    return (ddx * wx).abs().mean() + (ddy * wy).abs().mean()


def unsup_loss_stage(imgs, cams, depth, stage_idx, smooth_weight=0.18):
    """imgs [B,V,3,H,W] full resolution; cams [B,V,2,4,4] at the stage's scale; depth [B,h,w]; stage_idx 0/1/2.
    Returns (total, reconstr, ssim, smooth) like UnSupLoss.forward + its attributes."""
    V = imgs.shape[1]
    step = {0: 4, 1: 2, 2: 1}[stage_idx]
    small = [imgs[:, v, :, ::step, ::step].permute(0, 2, 3, 1) for v in range(V)]     # nearest interpolate == strided pick
    ref = small[0]
    ssim_loss, losses = 0.0, []
    for v in range(1, V):
        warped, mask = inverse_warping(small[v], cams[:, 0], cams[:, v], depth)
        losses.append(reconstr_loss(warped, ref, mask) + 1e4 * (1 - mask))
        if v < 3:
            ssim_loss = ssim_loss + ssim(ref, warped, mask).mean()
    smooth = depth_smoothness(depth.unsqueeze(-1), ref)
    vol = torch.stack(losses).permute(1, 2, 3, 4, 0)                              # [B,h,w,1,V-1]
    top = -torch.topk(-vol, k=1, sorted=False)[0]
    top = top * (top < 1e4).float()
    reconstr = top.sum(-1).mean()
    return 12 * reconstr + 6 * ssim_loss + smooth_weight * smooth, reconstr, ssim_loss, smooth
