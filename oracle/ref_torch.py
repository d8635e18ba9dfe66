"""CPU oracle, multi-threaded port (torch CPU tensors).  TEST / BASELINE INFRASTRUCTURE ONLY.

Same algorithms as oracle/ref_np.py (which is pinned to the reference's golden vectors) but vectorised with torch
CPU ops so it finishes at BASELINE sizes and can use every host core: it is the `cpu_baseline` leg of bench.py and
the `--impl reference` arm (the reference is Python and cannot travel to the GPU box), and the
checker for mid-size parity tests.  The dense convolutions call torch's own CPU conv2d/conv3d/conv_transpose3d —
exactly the third-party arithmetic the reference itself calls (reference models/modules.py:140,187); the warp,
variance, hypotheses and regression are restated explicitly (gather-based bilinear, no grid_sample).

Pinned by tests/test_oracle_golden.py::test_torch_port_matches_numpy_oracle_and_golden.
Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def stage_projection(proj):
    """[V,2,4,4] -> [V,4,4]  (reference models/casmvsnet.py:78-81)."""
    proj = _t(proj).float()
    out = proj[:, 0].clone()
    out[:, :3, :4] = torch.matmul(proj[:, 1, :3, :3], proj[:, 0, :3, :4])
    return out


def homo_warp(src, src_proj, ref_proj, depth):
    """src [C,h,w]; depth [D,h,w] -> [C,D,h,w]  (reference models/modules.py:304-339)."""
    C, h, w = src.shape
    D = depth.shape[0]
    proj = src_proj @ torch.linalg.inv(ref_proj)
    rot, trans = proj[:3, :3], proj[:3, 3]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing="ij")
    xyz = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w)])
    rot_xyz = rot @ xyz                                                   # [3,hw]
    p = rot_xyz[:, None, :] * depth.reshape(1, D, -1) + trans[:, None, None]
    px = (p[0] / p[2]).reshape(-1)
    py = (p[1] / p[2]).reshape(-1)
    finite = torch.isfinite(px) & torch.isfinite(py) & (px.abs() < 1e8) & (py.abs() < 1e8)
    px = torch.where(finite, px, torch.full_like(px, -10.0))
    py = torch.where(finite, py, torch.full_like(py, -10.0))
    x0f, y0f = torch.floor(px), torch.floor(py)
    ax, ay = px - x0f, py - y0f
    x0, y0 = x0f.long(), y0f.long()
    flat = src.reshape(C, -1)
    out = torch.zeros((C, px.numel()), dtype=torch.float32)
    for dx, dy, wgt in ((0, 0, (1 - ax) * (1 - ay)), (1, 0, ax * (1 - ay)), (0, 1, (1 - ax) * ay), (1, 1, ax * ay)):
        xx, yy = x0 + dx, y0 + dy
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        idx = yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)
        out += flat[:, idx] * torch.where(ok, wgt, torch.zeros_like(wgt))[None]
    return out.view(C, D, h, w)


def homo_warp_gs(src, src_proj, ref_proj, depth):
    """Same function through the library call the reference itself makes (models/modules.py:315-337): the sample grid is
    normalised to [-1,1] with (w-1)/2, (h-1)/2 and handed to F.grid_sample(bilinear, zeros, align_corners=True).  This is
    the variant the baseline arms time (`bench.py --impl reference`, cpu_baseline, stock_torch_gpu): it is what the
    reference executes, multi-threaded on CPU and ATen's grid_sampler_2d kernel on a GPU.  Runs on src's device."""
    C, h, w = src.shape
    D = depth.shape[0]
    dev = src.device
    proj = src_proj @ torch.inverse(ref_proj)
    rot, trans = proj[:3, :3], proj[:3, 3:4]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev),
                            torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    xyz = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev)])
    p = (rot @ xyz)[:, None, :] * depth.reshape(1, D, -1) + trans.view(3, 1, 1)
    xy = p[:2] / p[2:3]
    grid = torch.stack([xy[0] / ((w - 1) / 2) - 1, xy[1] / ((h - 1) / 2) - 1], dim=-1)         # [D, h*w, 2]
    out = F.grid_sample(src[None], grid.view(1, D * h, w, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(C, D, h, w)


def costvolume_variance(features, projs, depth, grid_sample=False):
    """features: list of V [C,h,w]; projs [V,4,4]; depth [D,h,w]  (reference models/casmvsnet.py:254-288)."""
    V, D = len(features), depth.shape[0]
    ref = features[0][:, None].expand(-1, D, -1, -1)
    vsum, vsq = ref.clone(), ref ** 2
    warp = homo_warp_gs if grid_sample else homo_warp
    for v in range(1, V):
        wv = warp(features[v], projs[v], projs[0], depth)
        vsum += wv
        vsq += wv ** 2
    return vsq / V - (vsum / V) ** 2


def _bn(x, P, name):
    shp = (1, -1) + (1,) * (x.dim() - 2)
    s = P[name + ".weight"] / torch.sqrt(P[name + ".running_var"] + 1e-5)
    return x * s.view(shp) + (P[name + ".bias"] - P[name + ".running_mean"] * s).view(shp)


def costregnet(x, P):
    """x [1,C,D,h,w] -> [D,h,w]  (reference models/modules.py:492-501)."""
    def cbr(y, n, stride=1):
        return F.relu(_bn(F.conv3d(y, P[n + ".conv.weight"], stride=stride, padding=1), P, n + ".bn"))

    def dbr(y, n):
        return F.relu(_bn(F.conv_transpose3d(y, P[n + ".conv.weight"], stride=2, padding=1, output_padding=1), P,
                          n + ".bn"))
    c0 = cbr(x, "conv0")
    c2 = cbr(cbr(c0, "conv1", 2), "conv2")
    c4 = cbr(cbr(c2, "conv3", 2), "conv4")
    y = cbr(cbr(c4, "conv5", 2), "conv6")
    y = c4 + dbr(y, "conv7")
    y = c2 + dbr(y, "conv9")
    y = c0 + dbr(y, "conv11")
    return F.conv3d(y, P["prob.weight"], padding=1)[0, 0]


def softargmin_conf(cost, depth):
    """cost, depth [D,h,w] -> depth [h,w], conf [h,w]  (reference models/casmvsnet.py:299-309)."""
    D = cost.shape[0]
    p = torch.softmax(cost, dim=0)
    d = (p * depth).sum(0)
    idx = (p * torch.arange(D, dtype=torch.float32, device=p.device).view(D, 1, 1)).sum(0).long().clamp(0, D - 1)
    pp = F.pad(p, (0, 0, 0, 0, 1, 2))
    sum4 = pp[0:D] + pp[1:D + 1] + pp[2:D + 2] + pp[3:D + 3]
    return d, torch.gather(sum4, 0, idx[None])[0]


def _resize(x, h, w):
    return F.interpolate(x[None], (h, w), mode="bilinear", align_corners=False)[0]


def depth_hypotheses_next(prev, D, interval_pixel, H, W, h, w):
    """reference models/modules.py:553-564 + casmvsnet.py:383-404."""
    cur = _resize(prev[None], H, W)[0]
    half = float(D / 2 * interval_pixel)
    cmin, cmax = cur - half, cur + half
    step = (cmax - cmin) / (D - 1)
    s = cmin[None] + torch.arange(D, dtype=torch.float32, device=prev.device).view(D, 1, 1) * step[None]
    return s if (h, w) == (H, W) else _resize(s, h, w)


def featurenet(imgs, P):
    """imgs [V,3,H,W] -> dict of [V,C,h,w]  (reference models/modules.py:425-464, fpn)."""
    def cbr(x, n, stride=1, pad=1):
        return F.relu(_bn(F.conv2d(x, P[n + ".conv.weight"], stride=stride, padding=pad), P, n + ".bn"))
    c0 = cbr(cbr(imgs, "conv0.0"), "conv0.1")
    c1 = cbr(cbr(cbr(c0, "conv1.0", 2, 2), "conv1.1"), "conv1.2")
    c2 = cbr(cbr(cbr(c1, "conv2.0", 2, 2), "conv2.1"), "conv2.2")
    out = {"stage1": F.conv2d(c2, P["out1.weight"])}
    intra = F.interpolate(c2, scale_factor=2, mode="nearest") + F.conv2d(c1, P["inner1.weight"], P["inner1.bias"])
    out["stage2"] = F.conv2d(intra, P["out2.weight"], padding=1)
    intra = F.interpolate(intra, scale_factor=2, mode="nearest") + F.conv2d(c0, P["inner2.weight"], P["inner2.bias"])
    out["stage3"] = F.conv2d(intra, P["out3.weight"], padding=1)
    return out


def sub(P, prefix, device="cpu"):
    return {k[len(prefix):]: _t(v).to(device) for k, v in P.items() if k.startswith(prefix)}


@torch.no_grad()
def cascade_forward_eval(imgs, proj_matrices, depth_values, P, ndepths=(48, 32, 8), ratios=(4, 2, 1), device="cpu",
                         grid_sample=False, as_numpy=True):
    """imgs [V,3,H,W]; proj_matrices {stageK: [V,2,4,4]}; depth_values [Dv]  (reference casmvsnet.py:356-417).
    `device` / `grid_sample`: the baseline arms run the same op sequence on a GPU and through F.grid_sample."""
    imgs, depth_values = _t(imgs).float().to(device), _t(depth_values).float().to(device)
    P = {k: (v if isinstance(v, torch.Tensor) and v.device == torch.device(device) else _t(v).to(device)) for k, v in P.items()}
    V, _, H, W = imgs.shape
    interval = (float(depth_values[-1]) - float(depth_values[0])) / depth_values.shape[0]
    feats = featurenet(imgs, sub(P, "feature.", device))
    out, depth = {}, None
    for s, D in enumerate(ndepths):
        key = "stage%d" % (s + 1)
        scale = (4, 2, 1)[s]
        h, w = H // scale, W // scale
        if depth is None:
            step = (depth_values[-1] - depth_values[0]) / (D - 1)
            hyp = (depth_values[0] + torch.arange(D, dtype=torch.float32, device=imgs.device) * step).view(D, 1, 1).expand(D, h, w)
        else:
            hyp = depth_hypotheses_next(depth, D, ratios[s] * interval, H, W, h, w)
        projs = stage_projection(proj_matrices[key]).to(imgs.device)
        var = costvolume_variance([feats[key][v] for v in range(V)], projs, hyp, grid_sample)
        cost = costregnet(var[None], sub(P, "cost_regularization.%d." % s, device))
        depth, conf = softargmin_conf(cost, hyp)
        out[key] = ({"depth": depth.cpu().numpy(), "photometric_confidence": conf.cpu().numpy()} if as_numpy else
                    {"depth": depth, "photometric_confidence": conf})
    return out
