"""CPU oracle (numpy) for the rendering-consistency branch a7-a13.  TEST INFRASTRUCTURE ONLY.

Restates reference models/render_models.py (Neural_Volume_Net :736-760, CostReg :690-734, Embedder :15-49,
Renderer_ours :143-220), models/render_utils.py (get_rays_mvs :86-108, get_ndc_coordinate :112-146, build_rays_norm
:149-243, build_color_volume :247-279, index_point_feature :304-330), models/renderer.py (raw2alpha :18-26,
raw2outputs :65-93, rendering :168-195) and losses/homography.py (:6-200) in plain numpy.
Pinned against tests/golden/render_small.npz / inverse_warp_small.npz (made by the unmodified reference,
tests/golden/make_golden_render.py) in tests/test_oracle_golden.py.  Product code never imports this module.
"""
import numpy as np

from . import ref_np as O

f32 = np.float32


# ---- a7 Neural_Volume_Net -----------------------------------------------------------------------------------------
def resize_depth_linear(vol, D1):
    """[C,D0,h,w] -> [C,D1,h,w]: F.interpolate(trilinear, align_corners=True) with unchanged h,w (identity there)."""
    D0 = vol.shape[1]
    src = (np.arange(D1, dtype=f32) * f32((D0 - 1) / (D1 - 1))).astype(f32)
    i0 = np.minimum(np.floor(src).astype(np.int64), D0 - 1)
    i1 = np.minimum(i0 + 1, D0 - 1)
    l1 = (src - i0.astype(f32)).astype(f32)[None, :, None, None]
    return ((f32(1) - l1) * vol[:, i0] + l1 * vol[:, i1]).astype(f32)


def _cb(x, P, name, stride=1):
    y = O.conv3d(x, P[name + ".conv.weight"], stride=stride)
    return O.bn_eval(y, P[name + ".bn.weight"], P[name + ".bn.bias"], P[name + ".bn.running_mean"],
                     P[name + ".bn.running_var"])


def _db(x, P, name):
    y = O.conv_transpose3d(x, P[name + ".0.weight"])
    return O.bn_eval(y, P[name + ".1.weight"], P[name + ".1.bias"], P[name + ".1.running_mean"],
                     P[name + ".1.running_var"])


def neural_volume_net(vol, P, D=128):
    """vol [41,D0,h,w]; P keyed without 'MVSNet.cost_reg_2.' -> [8,128,h,w].  conv -> BN, no ReLU (:685-686)."""
    x = resize_depth_linear(vol, D)
    c0 = _cb(x, P, "conv0")
    c2 = _cb(_cb(c0, P, "conv1", 2), P, "conv2")
    c4 = _cb(_cb(c2, P, "conv3", 2), P, "conv4")
    y = _cb(_cb(c4, P, "conv5", 2), P, "conv6")
    y = c4 + _db(y, P, "conv7")
    y = c2 + _db(y, P, "conv9")
    return c0 + _db(y, P, "conv11")


# ---- a8 rays --------------------------------------------------------------------------------------------------------
def rays_from_pixels(ys, xs, K, c2w):
    """get_rays_mvs: dirs = ((x-cx)/fx, (y-cy)/fy, 1) @ c2w[:3,:3]^T (unnormalised, z=1 in the camera frame)."""
    dirs = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs)], -1).astype(f32)
    return c2w[:3, 3].astype(f32), (dirs @ c2w[:3, :3].T.astype(f32)).astype(f32)


def ndc_coordinates(pts, w2c, K, W, H, near, far):
    """get_ndc_coordinate with intrinsics: pixel/(W-1,H-1), (z-near)/(far-near).  pts [...,3]."""
    p = (pts.reshape(-1, 3) @ w2c[:3, :3].T.astype(f32) + w2c[:3, 3].astype(f32)).astype(f32)
    q = (p @ K.T.astype(f32)).astype(f32)
    out = np.empty_like(q)
    out[:, 0] = q[:, 0] / q[:, 2] / f32(W - 1)
    out[:, 1] = q[:, 1] / q[:, 2] / f32(H - 1)
    out[:, 2] = (q[:, 2] - f32(near)) / f32(far - near)
    return out.reshape(pts.shape)


def stratified_uniform(near, far, n_samples, t_rand):
    """render_utils.py:212-224: linspace, mid-point bins, lower + (upper-lower)*t_rand."""
    t = np.linspace(0.0, 1.0, n_samples, dtype=f32)[None]
    z = (f32(near) * (f32(1) - t) + f32(far) * t).astype(f32)
    mids = f32(0.5) * (z[:, 1:] + z[:, :-1])
    upper = np.concatenate([mids, z[:, -1:]], -1)
    lower = np.concatenate([z[:, :1], mids], -1)
    return (lower + (upper - lower) * t_rand).astype(f32)


# ---- a9 point features ------------------------------------------------------------------------------------------------
def sample_volume_trilinear(vol, ndc):
    """index_point_feature: grid_sample 5-D, bilinear(=trilinear), zeros padding, align_corners=True.
    vol [C,D,h,w]; ndc [...,3] in [0,1] (x,y,z) -> [...,C]."""
    C, D, h, w = vol.shape
    g = ndc.reshape(-1, 3).astype(f32)
    ix = ((g[:, 0] * f32(2) - f32(1) + f32(1)) / f32(2) * f32(w - 1)).astype(f32)
    iy = ((g[:, 1] * f32(2) - f32(1) + f32(1)) / f32(2) * f32(h - 1)).astype(f32)
    iz = ((g[:, 2] * f32(2) - f32(1) + f32(1)) / f32(2) * f32(D - 1)).astype(f32)
    x0, y0, z0 = np.floor(ix), np.floor(iy), np.floor(iz)
    out = np.zeros((g.shape[0], C), f32)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy, zz = x0 + dx, y0 + dy, z0 + dz
                wgt = (f32(1) - np.abs(ix - xx)) * (f32(1) - np.abs(iy - yy)) * (f32(1) - np.abs(iz - zz))
                ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h) & (zz >= 0) & (zz < D)
                xi = np.clip(xx, 0, w - 1).astype(np.int64)
                yi = np.clip(yy, 0, h - 1).astype(np.int64)
                zi = np.clip(zz, 0, D - 1).astype(np.int64)
                out += np.where(ok, wgt, f32(0)).astype(f32)[:, None] * vol[:, zi, yi, xi].T
    return out.reshape(ndc.shape[:-1] + (C,))


def sample_image_border(img, gx, gy):
    """grid_sample 4-D bilinear, padding_mode='border', align_corners=True.  img [C,H,W]; gx,gy normalised [-1,1]."""
    C, H, W = img.shape
    ix = np.clip(((gx + f32(1)) / f32(2) * f32(W - 1)).astype(f32), f32(0), f32(W - 1))
    iy = np.clip(((gy + f32(1)) / f32(2) * f32(H - 1)).astype(f32), f32(0), f32(H - 1))
    x0, y0 = np.floor(ix), np.floor(iy)
    out = np.zeros((C, ix.size), f32)
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            wgt = (f32(1) - np.abs(ix - xx)) * (f32(1) - np.abs(iy - yy))
            ok = (xx <= W - 1) & (yy <= H - 1)
            xi = np.clip(xx, 0, W - 1).astype(np.int64)
            yi = np.clip(yy, 0, H - 1).astype(np.int64)
            out += np.where(ok, wgt, f32(0)).astype(f32)[None] * img[:, yi, xi]
    return out


def color_features(pts, w2cs, Ks, imgs3):
    """build_color_volume(with_mask=True): for i in 0..2 project with pose i, sample image imgs3[i] (the reference
    passes imgs[:, -3:] = the three SOURCE images, paired with poses 0..2: render_consist_net.py:74), append mask."""
    N = pts.reshape(-1, 3).shape[0]
    _, H, W = imgs3[0].shape
    out = np.empty((N, 12), f32)
    for i in range(3):
        ndc = ndc_coordinates(pts.reshape(-1, 3), w2cs[i], Ks[i], W, H, 2.0, 6.0)    # defaults near=2, far=6 (unused z)
        gx = (ndc[:, 0] * f32(2) - f32(1)).astype(f32)
        gy = (ndc[:, 1] * f32(2) - f32(1)).astype(f32)
        out[:, 4 * i:4 * i + 3] = sample_image_border(imgs3[i], gx, gy).T
        out[:, 4 * i + 3] = ((gx > -1) & (gx < 1) & (gy > -1) & (gy < 1)).astype(f32)
    return out.reshape(pts.shape[:-1] + (12,))


# ---- a10 MLP -------------------------------------------------------------------------------------------------------------
def embed(x, multires=10):
    freqs = (2.0 ** np.arange(multires)).astype(f32)
    scaled = (x[..., None, :] * freqs[:, None]).reshape(x.shape[:-1] + (-1,)).astype(f32)
    return np.concatenate([x, np.sin(scaled), np.cos(scaled)], -1).astype(f32)


def render_mlp(ndc, feat, dirs, P):
    """Renderer_ours.forward (use_viewdirs=True).  ndc [N,3], feat [N,20], dirs [N,3]; P keyed 'network_fn.nerf.'."""
    def lin(x, name):
        return (x @ P[name + ".weight"].T + P[name + ".bias"]).astype(f32)
    pts = embed(ndc)
    bias = lin(feat, "pts_bias")
    h = pts
    for i in range(6):
        h = np.maximum(lin(h, "pts_linears.%d" % i) * bias, 0)
        if i == 4:
            h = np.concatenate([pts, h], -1)
    alpha = np.maximum(lin(h, "alpha_linear"), 0)
    h2 = np.maximum(lin(np.concatenate([lin(h, "feature_linear"), dirs], -1), "views_linears.0"), 0)
    rgb = 1.0 / (1.0 + np.exp(-lin(h2, "rgb_linear")))
    return np.concatenate([rgb, alpha], -1).astype(f32)


# ---- a11 compositing -------------------------------------------------------------------------------------------------------
def raw2outputs(raw, z_vals):
    """alpha = 1-exp(-sigma) (dists unused, renderer.py:18-26); T = exclusive cumprod(1-alpha+1e-10)."""
    sigma = raw[..., 3]
    alpha = (f32(1) - np.exp(-sigma)).astype(f32)
    T = np.cumprod(np.concatenate([np.ones_like(alpha[:, :1]), f32(1) - alpha + f32(1e-10)], -1), -1)[:, :-1]
    w = (alpha * T).astype(f32)
    rgb = (w[..., None] * raw[..., :3]).sum(-2)
    depth = (w * z_vals).sum(-1)
    return rgb.astype(f32), depth.astype(f32), w, alpha


# ---- a13 inverse_warping ---------------------------------------------------------------------------------------------------
def inverse_warping(img, left_cam, right_cam, depth):
    """img [h,w,C]; cams [2,4,4]; depth [h,w] -> warped [h,w,C], mask [h,w].  losses/homography.py:6-200: the source
    is projected with K_LEFT (:55-57), weights come from clamped corners (:191-194), asymmetric y mask (:155)."""
    h, w, C = img.shape
    Rl, Rr = left_cam[0, :3, :3].astype(f32), right_cam[0, :3, :3].astype(f32)
    tl, tr = left_cam[0, :3, 3:4].astype(f32), right_cam[0, :3, 3:4].astype(f32)
    Kl = left_cam[1, :3, :3].astype(f32)
    Kinv = np.linalg.inv(Kl).astype(f32)
    R_rel = (Rr @ Rl.T).astype(f32)
    t_rel = (tr - R_rel @ tl).astype(f32)
    xs = ((np.linspace(-1, 1, w, dtype=f32) + f32(1)) * f32(0.5) * f32(w - 1)).astype(f32)
    ys = ((np.linspace(-1, 1, h, dtype=f32) + f32(1)) * f32(0.5) * f32(h - 1)).astype(f32)
    gx, gy = np.meshgrid(xs, ys)
    grid = np.stack([gx.ravel(), gy.ravel(), np.ones(h * w, f32)])
    cam = (Kinv @ grid).astype(f32) * depth.reshape(1, -1).astype(f32)
    P = (Kl @ np.concatenate([R_rel, t_rel], 1)).astype(f32)                 # [3,4]
    pc = (P[:, :3] @ cam + P[:, 3:4]).astype(f32)
    px = pc[0] / (pc[2] + f32(1e-10))
    py = pc[1] / (pc[2] + f32(1e-10))
    xn = (px / f32(w - 1) * f32(2) - f32(1)).astype(f32)
    yn = (py / f32(h - 1) * f32(2) - f32(1)).astype(f32)
    x = ((xn + f32(1)) * f32(w - 1) / f32(2)).astype(f32)
    y = ((yn + f32(1)) * f32(h - 1) / f32(2)).astype(f32)
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    mask = ((x0 >= 0) & (x1 <= w - 1) & (y0 >= 0) & (y0 <= h - 1)).astype(f32)
    x0c, x1c = np.clip(x0, 0, w - 1), np.clip(x1, 0, w - 1)
    y0c, y1c = np.clip(y0, 0, h - 1), np.clip(y1, 0, h - 1)
    flat = img.reshape(-1, C)
    a, b = flat[y0c * w + x0c], flat[y1c * w + x0c]
    c, d = flat[y0c * w + x1c], flat[y1c * w + x1c]
    dx = (x1c.astype(f32) - x).astype(f32)
    dy = (y1c.astype(f32) - y).astype(f32)
    wa, wb = dx * dy, dx * (f32(1) - dy)
    wc, wd = (f32(1) - dx) * dy, (f32(1) - dx) * (f32(1) - dy)
    out = wa[:, None] * a + wb[:, None] * b + wc[:, None] * c + wd[:, None] * d
    return out.reshape(h, w, C).astype(f32), mask.reshape(h, w)
