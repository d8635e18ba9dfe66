"""CPU oracle (numpy) for the RC-MVSNet depth-inference hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy float32, the algorithms of the reference's hot path so the CUDA kernels can be
checked without the reference being importable (it does not exist on the GPU box).  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import it; the product path never does.

Pinned: every function below is checked against golden vectors produced by running the *unmodified* reference
(`tests/golden/make_golden.py`, which imports /root/reference in the authoring container) — see
tests/test_oracle_golden.py.  The reference has no tests/KATs of its own, so the goldens are the pin.

Third-party arithmetic: the reference calls torch (pinned torch==1.10.1 in its requirements.txt:54; goldens made
with torch 2.11 CPU).  grid_sample(bilinear, zeros, align_corners=True), conv3d, conv_transpose3d, batch_norm(eval),
softmax and interpolate(bilinear/trilinear, align_corners=False) are restated here from their published definitions.

Each function cites the reference file:line it follows (paths relative to the reference repo root).
"""
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------------------------------
# a1  homo_warping                                                         reference models/modules.py:304-339
# --------------------------------------------------------------------------------------------------------------
def stage_projection(proj):
    """[...,2,4,4] (extrinsic, intrinsic) -> [...,4,4] K·[R|t] with last row of the extrinsic kept.
    reference models/casmvsnet.py:78-81 (fp32 matmul)."""
    proj = np.asarray(proj, f32)
    out = proj[..., 0, :, :].copy()
    out[..., :3, :4] = np.matmul(proj[..., 1, :3, :3], proj[..., 0, :3, :4]).astype(f32)
    return out


def relative_projection(src_proj, ref_proj):
    """rot [3,3], trans [3] of src_proj @ inverse(ref_proj), fp32 as the reference (modules.py:315-317)."""
    proj = np.matmul(np.asarray(src_proj, f32), np.linalg.inv(np.asarray(ref_proj, f32)).astype(f32)).astype(f32)
    return proj[:3, :3].copy(), proj[:3, 3].copy()


def warp_coords(rot, trans, depth, h, w):
    """Pixel-space sample coordinates (px,py) [D,h,w] in the reference's fp32 operation order, including the
    normalise (/((W-1)/2) - 1, modules.py:330-331) and grid_sample's un-normalise ((g+1)/2*(W-1))."""
    ys, xs = np.meshgrid(np.arange(h, dtype=f32), np.arange(w, dtype=f32), indexing="ij")
    xyz = np.stack([xs.ravel(), ys.ravel(), np.ones(h * w, f32)])            # [3,h*w]
    rot_xyz = np.matmul(rot.astype(f32), xyz).astype(f32)                      # modules.py:325
    D = depth.shape[0]
    dep = depth.reshape(D, -1).astype(f32)                                     # [D,h*w]
    p = rot_xyz[:, None, :] * dep[None] + trans.astype(f32)[:, None, None]     # modules.py:326-328
    with np.errstate(divide="ignore", invalid="ignore"):
        px = p[0] / p[2]
        py = p[1] / p[2]                                                       # modules.py:329
        gx = px / f32((w - 1) / 2) - f32(1)
        gy = py / f32((h - 1) / 2) - f32(1)
        ix = (gx + f32(1)) / f32(2) * f32(w - 1)                               # ATen grid_sampler_unnormalize
        iy = (gy + f32(1)) / f32(2) * f32(h - 1)
    return ix.reshape(D, h, w).astype(f32), iy.reshape(D, h, w).astype(f32)


def bilinear_zeros(src, ix, iy):
    """grid_sample(mode=bilinear, padding_mode=zeros, align_corners=True) on pixel coords.
    src [C,H,W]; ix,iy [...]; each of the 4 taps is individually zeroed when out of bounds; NaN coords give 0
    contribution exactly as ATen (comparisons with NaN are false -> tap skipped)."""
    C, H, W = src.shape
    shp = ix.shape
    ix = ix.ravel()
    iy = iy.ravel()
    with np.errstate(invalid="ignore"):
        x0f = np.floor(ix)
        y0f = np.floor(iy)
        wx1 = (ix - x0f).astype(f32)
        wy1 = (iy - y0f).astype(f32)
        wx0 = (f32(1) - wx1).astype(f32)
        wy0 = (f32(1) - wy1).astype(f32)
    out = np.zeros((C, ix.size), f32)
    finite = np.isfinite(ix) & np.isfinite(iy) & (np.abs(ix) < 1e9) & (np.abs(iy) < 1e9)
    x0 = np.where(finite, x0f, -10).astype(np.int64)
    y0 = np.where(finite, y0f, -10).astype(np.int64)
    # ATen accumulates nw, ne, sw, se in this order (GridSamplerKernel / grid_sampler_2d_kernel)
    for (dx, dy, wgt) in ((0, 0, wx0 * wy0), (1, 0, wx1 * wy0), (0, 1, wx0 * wy1), (1, 1, wx1 * wy1)):
        xx = x0 + dx
        yy = y0 + dy
        ok = finite & (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        xi = np.clip(xx, 0, W - 1)
        yi = np.clip(yy, 0, H - 1)
        v = src[:, yi, xi]                                                     # [C,N]
        out += np.where(ok, wgt, f32(0)).astype(f32)[None] * v
    return out.reshape((C,) + shp)


def homo_warp(src_fea, src_proj, ref_proj, depth_values):
    """src_fea [C,h,w]; src_proj, ref_proj [4,4]; depth_values [D] or [D,h,w] -> [C,D,h,w]."""
    C, h, w = src_fea.shape
    depth_values = np.asarray(depth_values, f32)
    if depth_values.ndim == 1:
        depth_values = np.broadcast_to(depth_values[:, None, None], (depth_values.shape[0], h, w))
    rot, trans = relative_projection(src_proj, ref_proj)
    ix, iy = warp_coords(rot, trans, depth_values, h, w)
    return bilinear_zeros(np.asarray(src_fea, f32), ix, iy)


# --------------------------------------------------------------------------------------------------------------
# a2  cost-volume variance                                          reference models/casmvsnet.py:65-100,254-288
# --------------------------------------------------------------------------------------------------------------
def costvolume_variance(features, projs, depth_values, with_no_ref=False):
    """features: list of V arrays [C,h,w] (ref first); projs: [V,4,4] K·[R|t]; depth_values [D,h,w].
    Returns var [C,D,h,w] (= Σw²/V − (Σw/V)², casmvsnet.py:99) and, if asked, var_no_ref which sums only the
    warped source volumes but still divides by V (casmvsnet.py:100, training-mode semantics :87-90)."""
    V = len(features)
    D = depth_values.shape[0]
    ref = np.asarray(features[0], f32)
    vsum = np.repeat(ref[:, None], D, axis=1).astype(f32)
    vsq = (vsum ** 2).astype(f32)
    nsum = np.zeros_like(vsum)
    nsq = np.zeros_like(vsum)
    for v in range(1, V):
        wv = homo_warp(features[v], projs[v], projs[0], depth_values)
        vsum = vsum + wv
        vsq = vsq + wv ** 2
        nsum = nsum + wv
        nsq = nsq + wv ** 2
    var = (vsq / f32(V) - (vsum / f32(V)) ** 2).astype(f32)
    if with_no_ref:
        return var, (nsq / f32(V) - (nsum / f32(V)) ** 2).astype(f32)
    return var


# --------------------------------------------------------------------------------------------------------------
# a3  CostRegNet                                                       reference models/modules.py:118-210,470-501
# --------------------------------------------------------------------------------------------------------------
def conv3d(x, w, stride=1, pad=1):
    """x [Cin,D,H,W], w [Cout,Cin,3,3,3] -> [Cout,Do,Ho,Wo] (cross-correlation, zero padding)."""
    Cin, D, H, W = x.shape
    Cout, _, kD, kH, kW = w.shape
    xp = np.zeros((Cin, D + 2 * pad, H + 2 * pad, W + 2 * pad), np.float64)
    xp[:, pad:pad + D, pad:pad + H, pad:pad + W] = x
    Do = (D + 2 * pad - kD) // stride + 1
    Ho = (H + 2 * pad - kH) // stride + 1
    Wo = (W + 2 * pad - kW) // stride + 1
    out = np.zeros((Cout, Do, Ho, Wo), np.float64)
    for kd in range(kD):
        for kh in range(kH):
            for kw in range(kW):
                patch = xp[:, kd:kd + stride * Do:stride, kh:kh + stride * Ho:stride, kw:kw + stride * Wo:stride]
                out += np.einsum("oc,cdhw->odhw", w[:, :, kd, kh, kw].astype(np.float64), patch)
    return out.astype(f32)


def conv_transpose3d(x, w, stride=2, pad=1, outpad=1):
    """x [Cin,D,H,W], w [Cin,Cout,3,3,3] (torch ConvTranspose3d layout) -> [Cout, 2D, 2H, 2W]:
    out[o, s*i+k-pad] += x[c,i] * w[c,o,k]."""
    Cin, D, H, W = x.shape
    _, Cout, kD, kH, kW = w.shape
    Do = (D - 1) * stride - 2 * pad + kD + outpad
    Ho = (H - 1) * stride - 2 * pad + kH + outpad
    Wo = (W - 1) * stride - 2 * pad + kW + outpad
    full = np.zeros((Cout, (D - 1) * stride + kD + outpad, (H - 1) * stride + kH + outpad,
                     (W - 1) * stride + kW + outpad), np.float64)
    for kd in range(kD):
        for kh in range(kH):
            for kw in range(kW):
                contrib = np.einsum("co,cdhw->odhw", w[:, :, kd, kh, kw].astype(np.float64), x.astype(np.float64))
                full[:, kd:kd + stride * D:stride, kh:kh + stride * H:stride, kw:kw + stride * W:stride] += contrib
    return full[:, pad:pad + Do, pad:pad + Ho, pad:pad + Wo].astype(f32)


def bn_eval(x, weight, bias, mean, var, eps=1e-5):
    """BatchNorm in eval mode on [C,...]: (x-mean)/sqrt(var+eps)*weight+bias (modules.py:143,151-152)."""
    shp = (-1,) + (1,) * (x.ndim - 1)
    inv = (f32(1) / np.sqrt(var.astype(f32) + f32(eps))).astype(f32)
    return ((x - mean.reshape(shp)) * inv.reshape(shp) * weight.reshape(shp) + bias.reshape(shp)).astype(f32)


def _cbr(x, P, name, transposed=False, stride=1):
    w = P[name + ".conv.weight"]
    y = conv_transpose3d(x, w) if transposed else conv3d(x, w, stride=stride)
    y = bn_eval(y, P[name + ".bn.weight"], P[name + ".bn.bias"], P[name + ".bn.running_mean"],
                P[name + ".bn.running_var"])
    return np.maximum(y, f32(0))


def costregnet(x, P):
    """x [Cin,D,h,w]; P: dict of this CostRegNet's parameters keyed like the state_dict without the
    'cost_regularization.N.' prefix.  Returns [D,h,w] (the squeezed `prob` conv output).  modules.py:492-501:
    skip adds happen after ReLU."""
    conv0 = _cbr(x, P, "conv0")
    conv2 = _cbr(_cbr(conv0, P, "conv1", stride=2), P, "conv2")
    conv4 = _cbr(_cbr(conv2, P, "conv3", stride=2), P, "conv4")
    y = _cbr(_cbr(conv4, P, "conv5", stride=2), P, "conv6")
    y = conv4 + _cbr(y, P, "conv7", transposed=True)
    y = conv2 + _cbr(y, P, "conv9", transposed=True)
    y = conv0 + _cbr(y, P, "conv11", transposed=True)
    return conv3d(y, P["prob.weight"])[0]


# --------------------------------------------------------------------------------------------------------------
# a4  softmax + soft-argmin + photometric confidence        reference models/casmvsnet.py:111-122, modules.py:519-525
# --------------------------------------------------------------------------------------------------------------
def softargmin_conf(cost, depth_values):
    """cost [D,h,w], depth_values [D,h,w] -> depth [h,w], conf [h,w], idx [h,w] int64.
    p = softmax_D(cost); depth = Σ p·d; idx = clamp(trunc(Σ p·i)); conf = p[idx-1]+p[idx]+p[idx+1]+p[idx+2]
    with zero padding (pad (1,2) + 4*avg_pool(4), casmvsnet.py:117-122)."""
    D = cost.shape[0]
    m = cost.max(axis=0, keepdims=True)
    e = np.exp((cost - m).astype(f32)).astype(f32)
    s = np.zeros(cost.shape[1:], f32)
    for d in range(D):
        s = (s + e[d]).astype(f32)
    p = (e / s[None]).astype(f32)
    depth = np.zeros_like(s)
    fidx = np.zeros_like(s)
    for d in range(D):                                      # ascending-d fp32 accumulation
        depth = (depth + p[d] * depth_values[d].astype(f32)).astype(f32)
        fidx = (fidx + p[d] * f32(d)).astype(f32)
    idx = np.clip(np.trunc(fidx).astype(np.int64), 0, D - 1)
    pp = np.zeros((D + 3,) + s.shape, f32)
    pp[1:D + 1] = p
    sum4 = f32(4) * ((pp[0:D] + pp[1:D + 1] + pp[2:D + 2] + pp[3:D + 3]) * f32(0.25)).astype(f32)
    conf = np.take_along_axis(sum4, idx[None], axis=0)[0]
    return depth, conf.astype(f32), idx, p


# --------------------------------------------------------------------------------------------------------------
# a5  depth hypotheses                          reference models/modules.py:549-588, models/casmvsnet.py:191-217
# --------------------------------------------------------------------------------------------------------------
def _src_index(dst, scale, in_size):
    """ATen area_pixel_compute_source_index(align_corners=False): max(scale*(dst+0.5)-0.5, 0); returns i0,i1,l1."""
    src = np.maximum(f32(scale) * (dst.astype(f32) + f32(0.5)) - f32(0.5), f32(0)).astype(f32)
    i0 = np.floor(src).astype(np.int64)
    i0 = np.minimum(i0, in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    return i0, i1, l1


def resize_bilinear(x, out_h, out_w):
    """F.interpolate(mode='bilinear', align_corners=False) on [...,h,w] (scale = in/out as torch computes it)."""
    h, w = x.shape[-2:]
    y0, y1, ly = _src_index(np.arange(out_h), h / out_h, h)
    x0, x1, lx = _src_index(np.arange(out_w), w / out_w, w)
    ly0 = (f32(1) - ly)[:, None]
    lx0 = (f32(1) - lx)[None, :]
    ly1 = ly[:, None]
    lx1 = lx[None, :]
    a = x[..., y0[:, None], x0[None, :]]
    b = x[..., y0[:, None], x1[None, :]]
    c = x[..., y1[:, None], x0[None, :]]
    d = x[..., y1[:, None], x1[None, :]]
    return (ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d)).astype(f32)


def depth_hypotheses_stage1(depth_values_1d, ndepth, h, w):
    """[Dv] -> [ndepth,h,w]: linspace between first/last value (modules.py:574-582); the trilinear resize of a
    constant-in-xy volume with unchanged D is the identity (casmvsnet.py:212-217)."""
    dmin = f32(depth_values_1d[0])
    dmax = f32(depth_values_1d[-1])
    interval = f32((dmax - dmin) / f32(ndepth - 1))
    d = (dmin + np.arange(ndepth, dtype=f32) * interval).astype(f32)
    return np.broadcast_to(d[:, None, None], (ndepth, h, w)).copy()


def depth_hypotheses_next(prev_depth, ndepth, interval_pixel, H, W, h, w):
    """prev_depth [hp,wp] -> [ndepth,h,w].  Upsample to full res (casmvsnet.py:196-198), per-pixel range
    (modules.py:553-564), then trilinear resize to the stage resolution (casmvsnet.py:212-217; D unchanged so it
    is a per-plane bilinear resize)."""
    cur = resize_bilinear(np.asarray(prev_depth, f32), H, W)
    half = f32(ndepth / 2 * interval_pixel)
    cmin = (cur - half).astype(f32)
    cmax = (cur + half).astype(f32)
    new_int = ((cmax - cmin) / f32(ndepth - 1)).astype(f32)
    samples = (cmin[None] + np.arange(ndepth, dtype=f32)[:, None, None] * new_int[None]).astype(f32)
    if (h, w) == (H, W):
        return samples
    return resize_bilinear(samples, h, w)


# --------------------------------------------------------------------------------------------------------------
# f1  FeatureNet (fpn)  — needed to drive the whole forward              reference models/modules.py:363-464
# --------------------------------------------------------------------------------------------------------------
def conv2d(x, w, bias=None, stride=1, pad=1):
    Cin, H, W = x.shape
    Cout, _, kH, kW = w.shape
    xp = np.zeros((Cin, H + 2 * pad, W + 2 * pad), np.float64)
    xp[:, pad:pad + H, pad:pad + W] = x
    Ho = (H + 2 * pad - kH) // stride + 1
    Wo = (W + 2 * pad - kW) // stride + 1
    out = np.zeros((Cout, Ho, Wo), np.float64)
    for kh in range(kH):
        for kw in range(kW):
            out += np.einsum("oc,chw->ohw", w[:, :, kh, kw].astype(np.float64),
                             xp[:, kh:kh + stride * Ho:stride, kw:kw + stride * Wo:stride])
    if bias is not None:
        out += bias.reshape(-1, 1, 1)
    return out.astype(f32)


def _c2(x, P, name, stride=1, pad=1):
    y = conv2d(x, P[name + ".conv.weight"], stride=stride, pad=pad)
    y = bn_eval(y, P[name + ".bn.weight"], P[name + ".bn.bias"], P[name + ".bn.running_mean"],
                P[name + ".bn.running_var"])
    return np.maximum(y, f32(0))


def featurenet(img, P):
    """img [3,H,W]; P keyed without the 'feature.' prefix -> dict stage1..3 ([32,H/4,W/4],[16,H/2,W/2],[8,H,W])."""
    c0 = _c2(_c2(img, P, "conv0.0"), P, "conv0.1")
    c1 = _c2(_c2(_c2(c0, P, "conv1.0", stride=2, pad=2), P, "conv1.1"), P, "conv1.2")
    c2 = _c2(_c2(_c2(c1, P, "conv2.0", stride=2, pad=2), P, "conv2.1"), P, "conv2.2")
    out = {"stage1": conv2d(c2, P["out1.weight"], pad=0)}
    intra = np.repeat(np.repeat(c2, 2, axis=1), 2, axis=2) + conv2d(c1, P["inner1.weight"], P["inner1.bias"], pad=0)
    out["stage2"] = conv2d(intra, P["out2.weight"])
    intra = np.repeat(np.repeat(intra, 2, axis=1), 2, axis=2) + conv2d(c0, P["inner2.weight"], P["inner2.bias"], pad=0)
    out["stage3"] = conv2d(intra, P["out3.weight"])
    return out


def sub_params(P, prefix):
    return {k[len(prefix):]: np.asarray(v) for k, v in P.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------------------------------
# a6  CascadeMVSNet_eval.forward                                      reference models/casmvsnet.py:356-417
# --------------------------------------------------------------------------------------------------------------
def cascade_forward_eval(imgs, proj_matrices, depth_values, P, ndepths=(48, 32, 8), ratios=(4, 2, 1),
                         features=None, keep=False):
    """imgs [V,3,H,W]; proj_matrices {stageK: [V,2,4,4]}; depth_values [Dv]; P full state-dict as numpy.
    Returns {stageK: {depth, photometric_confidence, (variance, cost, depth_values, idx)}}."""
    V, _, H, W = imgs.shape
    dmin = float(depth_values[0])
    dmax = float(depth_values[-1])
    depth_interval = (dmax - dmin) / depth_values.shape[0]                     # casmvsnet.py:357-359
    if features is None:
        fp = sub_params(P, "feature.")
        features = [featurenet(imgs[v], fp) for v in range(V)]
    out = {}
    depth = None
    for s in range(len(ndepths)):
        key = "stage%d" % (s + 1)
        scale = (4, 2, 1)[s]
        h, w = H // scale, W // scale
        D = ndepths[s]
        if depth is None:
            hyp = depth_hypotheses_stage1(depth_values, D, h, w)
        else:
            hyp = depth_hypotheses_next(depth, D, ratios[s] * depth_interval, H, W, h, w)
        projs = stage_projection(proj_matrices[key])                           # [V,4,4]
        feats = [f[key] for f in features]
        var = costvolume_variance(feats, projs, hyp)
        cost = costregnet(var, sub_params(P, "cost_regularization.%d." % s))
        depth, conf, idx, prob = softargmin_conf(cost, hyp)
        out[key] = {"depth": depth, "photometric_confidence": conf}
        if keep:
            out[key].update(variance=var, cost=cost, depth_values=hyp, idx=idx)
    return out
