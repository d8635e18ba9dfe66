"""Synthetic DTU-shaped inputs shared by tests, bench.py and the golden generator.

Nothing here reads a dataset: cameras follow the DTU geometry the reference's loader produces
(reference datasets/dtu_train.py:305-342: depth_min 425 mm, interval 2.5*1.06, stage intrinsics x1/x2/x4
with the base intrinsics at 1/4 resolution), images are either N(0,1) noise or a photo-consistent
smooth texture painted on the plane z=650 mm of the reference camera and rendered into every view.
"""
import math
import numpy as np

DEPTH_MIN = 425.0
DEPTH_INTERVAL = 2.65
NUM_DEPTH = 192


def depth_values(batch=1, num_depth=NUM_DEPTH, dmin=DEPTH_MIN, interval=DEPTH_INTERVAL):
    """[B,num_depth] float32: dmin + i*interval (reference datasets/dtu_train.py:305-306)."""
    d = (dmin + interval * np.arange(num_depth, dtype=np.float64)).astype(np.float32)
    return np.tile(d[None], (batch, 1))


def cameras(num_views, H, W, batch=1, baseline=50.0, focus=650.0):
    """proj_matrices dict {stage1..3: [B,V,2,4,4] float32}, [:,v,0]=world->cam extrinsic, [:,v,1,:3,:3]=K.

    View v sits at C_v=(baseline*v, -0.1*baseline*v, 0) in the reference camera frame and is rotated about y so
    that its optical axis passes through (0,0,focus): all views overlap on the swept depth range.
    """
    f1 = 361.5 * (W / 640.0)            # stage-1 (1/4 resolution) focal length in pixels
    cx1, cy1 = W / 8.0, H / 8.0
    ext = np.zeros((num_views, 4, 4), np.float64)
    for v in range(num_views):
        C = np.array([baseline * v, -0.1 * baseline * v, 0.0])
        ang = math.atan2(C[0], focus)   # ~0.077*v rad
        c, s = math.cos(ang), math.sin(ang)
        # camera z axis (in world) tilts toward -x so that it looks at (0,0,focus) from C
        R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])   # world -> cam
        ext[v, :3, :3] = R
        ext[v, :3, 3] = -R @ C
        ext[v, 3, 3] = 1.0
    out = {}
    for si, mult in enumerate((1.0, 2.0, 4.0)):
        P = np.zeros((batch, num_views, 2, 4, 4), np.float32)
        K = np.array([[f1 * mult, 0, cx1 * mult], [0, f1 * mult, cy1 * mult], [0, 0, 1.0]])
        for v in range(num_views):
            P[:, v, 0] = ext[v].astype(np.float32)
            P[:, v, 1, :3, :3] = K.astype(np.float32)
        out["stage%d" % (si + 1)] = P
    return out


def noise_images(num_views, H, W, batch=1, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((batch, num_views, 3, H, W), dtype=np.float32)


def textured_images(proj_stage3, H, W, plane_z=650.0, seed=0, nwaves=24):
    """Photo-consistent images: smooth random texture on the plane z=plane_z (ref frame) seen from each view.

    proj_stage3: [B,V,2,4,4] full-resolution cameras. Returns [B,V,3,H,W] float32 roughly in [-2,2].
    """
    rng = np.random.default_rng(seed)
    B, V = proj_stage3.shape[:2]
    amp = rng.uniform(0.3, 1.0, (3, nwaves))
    kx = rng.uniform(-0.25, 0.25, (3, nwaves))
    ky = rng.uniform(-0.25, 0.25, (3, nwaves))
    ph = rng.uniform(0, 2 * math.pi, (3, nwaves))
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    imgs = np.zeros((B, V, 3, H, W), np.float32)
    for b in range(B):
        E0 = proj_stage3[b, 0, 0].astype(np.float64)
        for v in range(V):
            E = proj_stage3[b, v, 0].astype(np.float64)
            K = proj_stage3[b, v, 1, :3, :3].astype(np.float64)
            Kinv = np.linalg.inv(K)
            # ray in view-v camera frame -> reference camera frame
            T = E0 @ np.linalg.inv(E)            # cam_v -> ref cam
            d = np.stack([xs, ys, np.ones_like(xs)], 0).reshape(3, -1)
            d = T[:3, :3] @ (Kinv @ d)
            o = T[:3, 3:4]
            t = (plane_z - o[2]) / d[2]
            X = o[0] + t * d[0]
            Y = o[1] + t * d[1]
            for c in range(3):
                val = (amp[c][:, None] * np.sin(kx[c][:, None] * X[None] + ky[c][:, None] * Y[None] + ph[c][:, None])).sum(0)
                imgs[b, v, c] = (val / math.sqrt(nwaves) * 1.6).reshape(H, W).astype(np.float32)
    return imgs


def mvs_inputs(num_views=5, H=512, W=640, batch=1, kind="textured", seed=0):
    """(imgs [B,V,3,H,W], proj_matrices dict, depth_values [B,192]) as numpy float32."""
    proj = cameras(num_views, H, W, batch)
    if kind == "noise":
        imgs = noise_images(num_views, H, W, batch, seed)
    else:
        imgs = textured_images(proj["stage3"], H, W, seed=seed)
        rng = np.random.default_rng(seed + 1)
        imgs = imgs + 0.05 * rng.standard_normal(imgs.shape, dtype=np.float32)
    return imgs, proj, depth_values(batch)


def stage_projections(proj_stage):
    """K·[R|t] 4x4 per view as the reference builds them (models/casmvsnet.py:78-81): float32 matmul."""
    P = proj_stage  # [B,V,2,4,4]
    out = P[:, :, 0].copy()
    out[:, :, :3, :4] = np.matmul(P[:, :, 1, :3, :3], P[:, :, 0, :3, :4])
    return out


# ---- renderer batch (reference datasets/dtu_train.py:352-364 keys) ---------------------------------------------
IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], np.float32)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], np.float32)


def render_batch(H=64, W=64, seed=0, num_views=4):
    """The NeRF-side entries of a training sample for V=4 views (1 ref + 3 src; the renderer is hard-wired to 3
    sources, reference models/render_models.py:750): imgs (ImageNet-normalised), depths_h, w2cs, c2ws, near_fars,
    intrinsics (full resolution), proj_mats; all with a leading batch dimension of 1, float32 numpy."""
    proj = cameras(num_views, H, W)
    ext = proj["stage3"][0, :, 0]                       # [V,4,4] world->cam
    K = proj["stage3"][0, :, 1, :3, :3]                 # [V,3,3] full-res intrinsics
    tex = textured_images(proj["stage3"], H, W, seed=seed)[0]              # roughly [-2,2]
    rgb01 = np.clip(0.5 + 0.2 * tex, 0.0, 1.0)
    imgs = (rgb01 - IMAGENET_MEAN[None, :, None, None]) / IMAGENET_STD[None, :, None, None]
    rng = np.random.default_rng(seed + 7)
    depths = (650.0 + 20.0 * rng.standard_normal((num_views, H, W))).astype(np.float32)
    c2ws = np.stack([np.linalg.inv(ext[v].astype(np.float64)) for v in range(num_views)]).astype(np.float32)
    near_fars = np.tile(np.array([[425.0, 425.0 + 2.65 * 192]], np.float32), (num_views, 1))
    proj_mats = np.stack([np.eye(4)[:3]] + [(stage_projections(proj["stage3"])[0, v] @
                                              np.linalg.inv(stage_projections(proj["stage3"])[0, 0]))[:3]
                                             for v in range(1, num_views)]).astype(np.float32)
    return {"imgs": imgs[None].astype(np.float32), "depths_h": depths[None], "w2cs": ext[None].astype(np.float32),
            "c2ws": c2ws[None], "near_fars": near_fars[None], "intrinsics": K[None].astype(np.float32),
            "proj_mats": proj_mats[None]}


class RenderArgs:
    """Defaults of train_rcmvsnet.py:74-120 for the NeRF branch."""
    multires, i_embed, pts_dim, dir_dim = 10, 0, 3, 3
    netdepth, netwidth, net_type, netchunk = 6, 128, "v0", 1024
    ckpt, N_samples, N_importance, perturb = None, 128, 0, 1.0
    use_viewdirs, white_bkgd, raw_noise_std = False, False, 0.0
    pad, img_downscale, use_color_volume, use_disp = 0, 1.0, False, False
    multires_views = 4
