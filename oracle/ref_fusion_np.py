"""CPU oracle for row f3 (geometric-consistency filter).  TEST INFRASTRUCTURE ONLY — never imported by the product.

Restates the reference's numpy/cv2 algorithm (eval_rcmvsnet_dtu.py:281-336 reproject_with_depth,
check_geometric_consistency; :385-389 of filter_depth) with the same dtype promotions; cv2.remap(INTER_LINEAR) is
restated from OpenCV's published algorithm (coordinates quantised to 1/32 px with round-half-even, 32x32 float weight
table, constant border 0, float left-to-right sum) so the oracle does not need OpenCV.
Pinned by tests/test_oracle_fusion_golden.py against tests/golden/geo_filter_small.npz (made by running the reference's
own functions, tests/golden/make_golden_r2.py).
"""
import numpy as np


def remap_linear(src, x, y):
    """cv2.remap(src, x, y, cv2.INTER_LINEAR) for float32 src [H,W] and float32 maps, BORDER_CONSTANT 0."""
    H, W = src.shape

    def cv_round(v):
        v = (v.astype(np.float32) * np.float32(32.0)).astype(np.float32)
        bad = ~(np.abs(v) < 2147483648.0)
        r = np.rint(np.where(bad, 0, v)).astype(np.int64)           # rint = round half to even
        return np.where(bad, -2 ** 31, r)
    sx, sy = cv_round(x), cv_round(y)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    fx = ((sx & 31).astype(np.float32) / np.float32(32.0)).astype(np.float32)
    fy = ((sy & 31).astype(np.float32) / np.float32(32.0)).astype(np.float32)
    one = np.float32(1.0)
    w = [(one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx]
    out = None
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        xx, yy = ix + dx, iy + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = np.where(ok, src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], np.float32(0)).astype(np.float32)
        term = (v * w[k].astype(np.float32)).astype(np.float32)
        out = term if out is None else (out + term).astype(np.float32)
    return out


def reproject_with_depth(depth_ref, K_ref, E_ref, depth_src, K_src, E_src):
    """eval_rcmvsnet_dtu.py:281-318."""
    H, W = depth_ref.shape
    xr, yr = np.meshgrid(np.arange(W), np.arange(H))
    xr, yr = xr.reshape(-1), yr.reshape(-1)
    ones = np.ones_like(xr)
    p_ref = np.linalg.inv(K_ref) @ (np.vstack((xr, yr, ones)) * depth_ref.reshape(-1))
    p_src = ((E_src @ np.linalg.inv(E_ref)) @ np.vstack((p_ref, ones)))[:3]
    kp = K_src @ p_src
    xy = kp[:2] / kp[2:3]
    xs = xy[0].reshape(H, W).astype(np.float32)
    ys = xy[1].reshape(H, W).astype(np.float32)
    sampled = remap_linear(depth_src, xs, ys)
    p_src = np.linalg.inv(K_src) @ (np.vstack((xy, ones)) * sampled.reshape(-1))
    p_back = ((E_ref @ np.linalg.inv(E_src)) @ np.vstack((p_src, ones)))[:3]
    depth_rep = p_back[2].reshape(H, W).astype(np.float32)
    kb = K_ref @ p_back
    xyb = kb[:2] / kb[2:3]
    return depth_rep, xyb[0].reshape(H, W).astype(np.float32), xyb[1].reshape(H, W).astype(np.float32), xs, ys


def check_geometric_consistency(depth_ref, K_ref, E_ref, depth_src, K_src, E_src, img_dist_thresh, depth_thresh):
    """eval_rcmvsnet_dtu.py:321-336."""
    H, W = depth_ref.shape
    xr, yr = np.meshgrid(np.arange(W), np.arange(H))
    depth_rep, x2, y2, xs, ys = reproject_with_depth(depth_ref, K_ref, E_ref, depth_src, K_src, E_src)
    with np.errstate(all="ignore"):
        dist = np.sqrt((x2 - xr) ** 2 + (y2 - yr) ** 2)
        rel = np.abs(depth_rep - depth_ref) / depth_ref
        mask = np.logical_and(dist < img_dist_thresh, rel < depth_thresh)
    depth_rep[~mask] = 0
    return mask, depth_rep, xs, ys


def filter_view(depth_ref, conf, K_ref, E_ref, depths_src, Ks_src, Es_src, prob_threshold=0.8, num_consistent=3,
                img_dist_thresh=0.5, depth_thresh=0.01):
    """eval_rcmvsnet_dtu.py:366-389 for one reference view."""
    geo_sum, reps = 0, []
    for d, k, e in zip(depths_src, Ks_src, Es_src):
        m, dr, _, _ = check_geometric_consistency(depth_ref, K_ref, E_ref, d, k, e, img_dist_thresh, depth_thresh)
        geo_sum = geo_sum + m.astype(np.int32)
        reps.append(dr)
    avg = (sum(reps) + depth_ref) / (geo_sum + 1)
    return {"geo_mask_sum": geo_sum, "depth_est_averaged": avg.astype(np.float32),
            "final_mask": np.logical_and(conf > prob_threshold, geo_sum >= num_consistent)}
