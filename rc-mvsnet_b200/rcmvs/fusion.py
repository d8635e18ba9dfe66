"""f3 — geometric-consistency filter of the reference's fusion step on the GPU.

`check_geometric_consistency` keeps the signature and return values of the reference function
(eval_rcmvsnet_dtu.py:321-336) for one (ref, src) pair; `filter_view` does what `filter_depth` does per reference view
(:366-389: all sources, geo_mask_sum, averaged depth, final mask) in ONE kernel launch (csrc/geo_filter.cu).
numpy arrays in -> numpy arrays out; CUDA tensors in -> CUDA tensors out (no host round trip).
"""
import numpy as np
import torch

from . import _ffi
from .ops import _chk, _opt, _stream, _on_device

__all__ = ["check_geometric_consistency", "filter_view", "pair_matrices"]


def pair_matrices(K_ref, E_ref, K_src, E_src):
    """The six small matrices of one (ref, src) pair, float32 like the reference computes them (:289, :292, :308, :311):
    [inv(K_ref), E_src@inv(E_ref), K_src, inv(K_src), E_ref@inv(E_src), K_ref] flattened to 68 floats."""
    K_ref, E_ref, K_src, E_src = (np.asarray(a, dtype=np.float32) for a in (K_ref, E_ref, K_src, E_src))
    parts = [np.linalg.inv(K_ref), np.matmul(E_src, np.linalg.inv(E_ref)), K_src, np.linalg.inv(K_src),
             np.matmul(E_ref, np.linalg.inv(E_src)), K_ref]
    return np.concatenate([p.astype(np.float32).ravel() for p in parts])


def _to_dev(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def _host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


@_on_device
def _run(depth_ref, conf, depth_src, mats, img_dist_thresh, depth_thresh, prob_thresh, num_consistent, per_source):
    H, W = depth_ref.shape
    NS, Hs, Ws = depth_src.shape
    dev = depth_ref.device
    out = {"geo_mask_sum": torch.empty((H, W), dtype=torch.int32, device=dev),
           "depth_est_averaged": torch.empty((H, W), dtype=torch.float32, device=dev),
           "final_mask": torch.empty((H, W), dtype=torch.uint8, device=dev)}
    if per_source:
        out["masks"] = torch.empty((NS, H, W), dtype=torch.uint8, device=dev)
        out["depth_reprojected"] = torch.empty((NS, H, W), dtype=torch.float32, device=dev)
        out["xy_src"] = torch.empty((NS, 2, H, W), dtype=torch.float32, device=dev)
    ptr = lambda k: out[k].data_ptr() if k in out else None
    _ffi.call("rcmvs_geo_consistency", _chk(depth_ref, "depth_ref"), _opt(conf, "conf"), _chk(depth_src, "depth_src"),
              _chk(mats, "pair_mats"), NS, H, W, Hs, Ws, float(img_dist_thresh), float(depth_thresh),
              float(prob_thresh), int(num_consistent), ptr("masks"), ptr("depth_reprojected"), ptr("xy_src"),
              out["geo_mask_sum"].data_ptr(), out["depth_est_averaged"].data_ptr(), out["final_mask"].data_ptr(),
              _stream())
    return out


def filter_view(depth_ref, confidence, K_ref, E_ref, depths_src, Ks_src, Es_src, prob_threshold=0.8, num_consistent=3,
                img_dist_thresh=0.5, depth_thresh=0.01, per_source=False, device=None):
    """One reference view against all its sources (reference filter_depth :366-389).

    depth_ref [H,W]; confidence [H,W] or None; depths_src [NS,H,W] (or a list); K*: 3x3, E*: 4x4 (numpy / tensors).
    Returns dict(geo_mask_sum int32, depth_est_averaged float32, final_mask bool/uint8[, masks, depth_reprojected,
    xy_src]) — numpy if the depth maps came as numpy, CUDA tensors otherwise."""
    as_numpy = not isinstance(depth_ref, torch.Tensor)
    if device is None:
        device = depth_ref.device if isinstance(depth_ref, torch.Tensor) else torch.device("cuda", torch.cuda.current_device())
    if isinstance(depths_src, (list, tuple)):
        depths_src = torch.stack([_to_dev(d, device) for d in depths_src])
    mats = np.stack([pair_matrices(_host(K_ref), _host(E_ref), _host(k), _host(e)) for k, e in zip(Ks_src, Es_src)])
    out = _run(_to_dev(depth_ref, device), None if confidence is None else _to_dev(confidence, device),
               _to_dev(depths_src, device), _to_dev(mats, device), img_dist_thresh, depth_thresh, prob_threshold,
               num_consistent, per_source)
    if as_numpy:
        out = {k: v.cpu().numpy() for k, v in out.items()}
        for k in ("final_mask", "masks"):
            if k in out:
                out[k] = out[k].astype(bool)
    return out


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                img_dist_thresh, depth_thresh):
    """Drop-in for eval_rcmvsnet_dtu.py:321-336 -> (mask, depth_reprojected, x2d_src, y2d_src)."""
    src = depth_src[None] if not isinstance(depth_src, torch.Tensor) else depth_src.unsqueeze(0)
    out = filter_view(depth_ref, None, intrinsics_ref, extrinsics_ref, src, [intrinsics_src], [extrinsics_src],
                      img_dist_thresh=img_dist_thresh, depth_thresh=depth_thresh, per_source=True)
    return out["masks"][0], out["depth_reprojected"][0], out["xy_src"][0, 0], out["xy_src"][0, 1]
