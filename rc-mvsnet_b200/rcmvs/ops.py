"""Torch-facing wrappers over the C ABI: raw data_ptr() + the current CUDA stream, caller-allocated outputs.

torch is plumbing here (device memory, streams); every function below enqueues hand-written sm_100a kernels from
librcmvs.so and raises if the tensors are not CUDA fp32 contiguous — there is no eager fallback.
"""
import functools

import torch

from . import _ffi


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Run `fn` with the device of its first tensor argument current (kernels launch on the CURRENT device and
    `_stream()` is that device's current stream): a model living on cuda:1 while cuda:0 is current — DataParallel
    replicas, as the reference's scripts wrap the model (eval_rcmvsnet_dtu.py:189) — must not enqueue on cuda:0."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def _chk(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("rcmvs: %s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("rcmvs: %s must be float32, got %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("rcmvs: %s must be contiguous" % name)
    return t.data_ptr()


def _opt(t, name):
    return None if t is None else _chk(t, name)


# ---- layout ---------------------------------------------------------------------------------------------------
@_on_device
def to_channels_last(x):
    """[N,C,*spatial] contiguous -> [N,*spatial,C] contiguous."""
    N, C = x.shape[:2]
    S = x[0, 0].numel()
    out = torch.empty((N,) + tuple(x.shape[2:]) + (C,), device=x.device, dtype=x.dtype)
    _ffi.call("rcmvs_planar_to_cl", _chk(x, "x"), _chk(out, "out"), N, C, S, _stream())
    return out


@_on_device
def to_planar(x):
    """[N,*spatial,C] contiguous -> [N,C,*spatial] contiguous."""
    N, C = x.shape[0], x.shape[-1]
    S = x[0, ..., 0].numel()
    out = torch.empty((N, C) + tuple(x.shape[1:-1]), device=x.device, dtype=x.dtype)
    _ffi.call("rcmvs_cl_to_planar", _chk(x, "x"), _chk(out, "out"), N, C, S, _stream())
    return out


# ---- projections ----------------------------------------------------------------------------------------------
@_on_device
def relative_projections(proj):
    """proj [B,V,2,4,4] -> rt [B,V-1,12]."""
    B, V = proj.shape[:2]
    rt = torch.empty((B, V - 1, 12), device=proj.device, dtype=torch.float32)
    _ffi.call("rcmvs_relative_projections", _chk(proj, "proj"), _chk(rt, "rt"), B, V, _stream())
    return rt


@_on_device
def relative_projection_pair(src_proj, ref_proj):
    B = src_proj.shape[0]
    rt = torch.empty((B, 12), device=src_proj.device, dtype=torch.float32)
    _ffi.call("rcmvs_relative_projection_pair", _chk(src_proj, "src_proj"), _chk(ref_proj, "ref_proj"),
              _chk(rt, "rt"), B, _stream())
    return rt


# ---- hypotheses -----------------------------------------------------------------------------------------------
@_on_device
def depth_hypotheses_first(depth_values, D):
    B, Dv = depth_values.shape
    planes = torch.empty((B, D), device=depth_values.device, dtype=torch.float32)
    _ffi.call("rcmvs_depth_hypotheses_first", _chk(depth_values, "depth_values"), _chk(planes, "planes"), B, Dv, D,
              _stream())
    return planes


@_on_device
def depth_hypotheses_next(prev_depth, D, interval_pixel, H, W, h, w):
    B, hp, wp = prev_depth.shape
    hyp = torch.empty((B, D, h, w), device=prev_depth.device, dtype=torch.float32)
    _ffi.call("rcmvs_depth_hypotheses_next", _chk(prev_depth, "prev_depth"), _chk(hyp, "hyp"), B, hp, wp, H, W, h, w,
              D, float(interval_pixel), _stream())
    return hyp


@_on_device
def depth_hypotheses_next_dev(prev_depth, depth_values, D, interval_ratio, H, W, h, w):
    """Like depth_hypotheses_next, the pixel interval = ratio * (max - min) / Dv of depth_values[0] taken on the device."""
    B, hp, wp = prev_depth.shape
    hyp = torch.empty((B, D, h, w), device=prev_depth.device, dtype=torch.float32)
    _ffi.call("rcmvs_depth_hypotheses_next_dev", _chk(prev_depth, "prev_depth"), _chk(depth_values, "depth_values"),
              depth_values.shape[1], float(interval_ratio), _chk(hyp, "hyp"), B, hp, wp, H, W, h, w, D, _stream())
    return hyp


# ---- warp / cost volume ---------------------------------------------------------------------------------------
def _depth_mode(depth, B, D_expected=None):
    if depth.dim() == 2:
        return 0
    if depth.dim() == 4:
        return 1
    raise RuntimeError("rcmvs: depth must be [B,D] or [B,D,h,w]")


@_on_device
def homo_warp_fwd(src_fea, rt, depth):
    B, C, h, w = src_fea.shape
    D = depth.shape[1]
    out = torch.empty((B, C, D, h, w), device=src_fea.device, dtype=torch.float32)
    _ffi.call("rcmvs_homo_warp_fwd", _chk(src_fea, "src_fea"), _chk(rt, "rt"), _chk(depth, "depth"),
              _depth_mode(depth, B), _chk(out, "out"), B, C, D, h, w, _stream())
    return out


@_on_device
def homo_warp_bwd(grad_out, rt, depth, h, w):
    B, C, D = grad_out.shape[:3]
    gsrc = torch.zeros((B, C, h, w), device=grad_out.device, dtype=torch.float32)
    _ffi.call("rcmvs_homo_warp_bwd", _chk(grad_out, "grad_out"), _chk(rt, "rt"), _chk(depth, "depth"),
              _depth_mode(depth, B), _chk(gsrc, "grad_src"), B, C, D, h, w, _stream())
    return gsrc


@_on_device
def costvolume_variance(feats_cl, rt, depth, want_noref=False):
    """feats_cl [B,V,h,w,C]; rt [B,V-1,12]; depth [B,D] or [B,D,h,w] -> var [B,D,h,w,C] (+ var_noref)."""
    B, V, h, w, C = feats_cl.shape
    D = depth.shape[1]
    var = torch.empty((B, D, h, w, C), device=feats_cl.device, dtype=torch.float32)
    noref = torch.empty_like(var) if want_noref else None
    _ffi.call("rcmvs_costvolume_var_fwd", _chk(feats_cl, "feats"), _chk(rt, "rt"), _chk(depth, "depth"),
              _depth_mode(depth, B), _chk(var, "var"), _opt(noref, "var_noref"), B, V, C, D, h, w, _stream())
    return (var, noref) if want_noref else var


@_on_device
def costvolume_variance_split(feats_cl, rt, depth):
    """Like costvolume_variance, the variance written as split-bf16 planes [2,B,D,h,w,C] (hi, lo) for conv3d_slab."""
    B, V, h, w, C = feats_cl.shape
    D = depth.shape[1]
    var = torch.empty((2, B, D, h, w, C), device=feats_cl.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_costvolume_var_fwd_split", _chk(feats_cl, "feats"), _chk(rt, "rt"), _chk(depth, "depth"),
              _depth_mode(depth, B), var.data_ptr(), B, V, C, D, h, w, _stream())
    return var


# ---- conv3d ---------------------------------------------------------------------------------------------------
@_on_device
def pack_conv3d_weight(w, scale=None, transposed=False):
    """torch Conv3d [Cout,Cin,3,3,3] / ConvTranspose3d [Cin,Cout,3,3,3] weight (x BN scale) -> [27,Cin,Cout]."""
    if transposed:
        Cin, Cout = w.shape[:2]
    else:
        Cout, Cin = w.shape[:2]
    packed = torch.empty((27, Cin, Cout), device=w.device, dtype=torch.float32)
    _ffi.call("rcmvs_pack_conv3d_weight", _chk(w, "w"), _opt(scale, "scale"), _chk(packed, "packed"), Cout, Cin,
              1 if transposed else 0, _stream())
    return packed


@_on_device
def conv3d(x_cl, packed_w, bias=None, skip=None, stride=1, relu=True):
    """x_cl [B,D,H,W,Cin] -> [B,Do,Ho,Wo,Cout]: conv(k3,p1) + bias, ReLU, then + skip."""
    B, D, H, W, Cin = x_cl.shape
    Cout = packed_w.shape[2]
    Do, Ho, Wo = (D - 1) // stride + 1, (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty((B, Do, Ho, Wo, Cout), device=x_cl.device, dtype=torch.float32)
    nws = _ffi.lib().rcmvs_conv3d_workspace_floats(B, D, H, W, Cin, Cout, stride)
    ws = torch.empty((nws,), device=x_cl.device, dtype=torch.float32) if nws > 0 else None
    _ffi.call("rcmvs_conv3d_fwd_ws", _chk(x_cl, "x"), _chk(packed_w, "w"), _opt(bias, "bias"), _opt(skip, "skip"),
              _chk(y, "y"), _opt(ws, "workspace"), B, D, H, W, Cin, Cout, stride, 1 if relu else 0, _stream())
    if ws is not None:
        _ffi.launches += 1          # conv + split-K reduce
    return y


@_on_device
def conv_transpose3d(x_cl, packed_w, bias=None, skip=None, relu=True):
    """x_cl [B,D,H,W,Cin] -> [B,2D,2H,2W,Cout] (k3, s2, p1, output_padding 1)."""
    B, D, H, W, Cin = x_cl.shape
    Cout = packed_w.shape[2]
    y = torch.empty((B, 2 * D, 2 * H, 2 * W, Cout), device=x_cl.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv_transpose3d_fwd", _chk(x_cl, "x"), _chk(packed_w, "w"), _opt(bias, "bias"),
              _opt(skip, "skip"), _chk(y, "y"), B, D, H, W, Cin, Cout, 1 if relu else 0, _stream())
    return y


# ---- conv3d on the tensor cores (tcgen05 / TMEM) ----------------------------------------------------------------
TC_SUPPORTED = {(8, 8), (16, 8), (32, 8), (8, 1), (16, 16)}      # (Cin, Cout) pairs of conv0 / prob / conv2


@_on_device
def conv3d_tc_pack(w, scale=None, nsplit=3):
    """torch Conv3d weight [Cout,Cin,3,3,3] (x BN scale) -> operand image of the tcgen05 kernel."""
    Cout, Cin = w.shape[:2]
    n = _ffi.lib().rcmvs_conv3d_tc_weight_floats(Cin, Cout, nsplit)
    img = torch.empty((n,), device=w.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv3d_tc_pack", _chk(w, "w"), _opt(scale, "scale"), _chk(img, "img"), Cout, Cin, nsplit,
              _stream())
    return img


@_on_device
def conv3d_tc(x_cl, img, Cout, bias=None, skip=None, nsplit=3, relu=True):
    """x_cl [B,D,H,W,Cin] -> [B,D,H,W,Cout]: stride-1 conv(k3,p1) + bias, ReLU, then + skip, on tcgen05.mma."""
    B, D, H, W, Cin = x_cl.shape
    y = torch.empty((B, D, H, W, Cout), device=x_cl.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv3d_tc_fwd", _chk(x_cl, "x"), _chk(img, "wimg"), _opt(bias, "bias"), _opt(skip, "skip"),
              _chk(y, "y"), B, D, H, W, Cin, Cout, nsplit, 1 if relu else 0, _stream())
    return y


# (Cin, Cout) the models route to the TMA-fed split-bf16 slab kernel.  The kernel also takes (8, 1) = `prob`, measured SLOWER
# there than on the FFMA plane-march kernel (CostRegNet of stage 1: 1.03 ms with it, 0.93 ms with the march kernel: one useful
# accumulator column per operand tile), so the models do not use it.
SLAB_SUPPORTED = {(32, 8), (16, 8), (16, 16), (48, 8), (8, 8)}


def _chk_bf16(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise RuntimeError("rcmvs: %s must be a contiguous CUDA bfloat16 tensor" % name)
    return t.data_ptr()


@_on_device
def split_bf16(x):
    """x [...] fp32 -> [2, ...] bf16: hi = bf16(x), lo = bf16(x - hi)."""
    xs = torch.empty((2,) + tuple(x.shape), device=x.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_split_bf16", _chk(x, "x"), xs.data_ptr(), x.numel(), _stream())
    return xs


@_on_device
def conv3d_slab_pack(w, scale=None):
    """torch Conv3d weight [Cout,Cin,3,3,3] (x BN scale) -> split-bf16 operand image of the TMA slab kernel."""
    Cout, Cin = w.shape[:2]
    n = _ffi.lib().rcmvs_conv3d_slab_weight_elems(Cin, Cout)
    img = torch.empty((n,), device=w.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_conv3d_slab_pack", _chk(w, "w"), _opt(scale, "scale"), img.data_ptr(), Cout, Cin, _stream())
    return img


@_on_device
def conv3d_slab(x_split, img, Cout, bias=None, skip=None, relu=True, want_y=True, want_split=False):
    """x_split [2,B,D,H,W,Cin] bf16 (hi, lo) -> y [B,D,H,W,Cout] fp32 and/or y_split [2,B,D,H,W,Cout] bf16."""
    _, B, D, H, W, Cin = x_split.shape
    if Cout == 1 and want_split:
        raise RuntimeError("rcmvs: the single-channel slab layer (`prob`) writes fp32 only")
    y = torch.empty((B, D, H, W, Cout), device=x_split.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2, B, D, H, W, Cout), device=x_split.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv3d_slab_fwd", _chk_bf16(x_split, "x_split"), _chk_bf16(img, "wimg"), _opt(bias, "bias"),
              _opt(skip, "skip"), _opt(y, "y"), None if ys is None else ys.data_ptr(), B, D, H, W, Cin, Cout,
              1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


SLABO_SUPPORTED = {(32, 32)}      # (Cin, Cout) of the all-taps-as-offsets slab kernel (conv4)


@_on_device
def conv3d_o_slab_pack(w, scale=None):
    Cout, Cin = w.shape[:2]
    img = torch.empty((_ffi.lib().rcmvs_conv3d_o_slab_weight_elems(Cin, Cout),), device=w.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_conv3d_o_slab_pack", _chk(w, "w"), _opt(scale, "scale"), img.data_ptr(), Cout, Cin, _stream())
    return img


@_on_device
def conv3d_o_slab(x_split, img, Cout, bias=None, skip=None, relu=True, want_y=True, want_split=False):
    """x_split [2,B,D,H,W,Cin] bf16 -> 3x3x3 stride-1 conv: y [B,D,H,W,Cout] fp32 and/or y_split (conv4: 32 -> 32)."""
    _, B, D, H, W, Cin = x_split.shape
    y = torch.empty((B, D, H, W, Cout), device=x_split.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2, B, D, H, W, Cout), device=x_split.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv3d_o_slab_fwd", _chk_bf16(x_split, "x_split"), _chk_bf16(img, "wimg"), _opt(bias, "bias"),
              _opt(skip, "skip"), _opt(y, "y"), None if ys is None else ys.data_ptr(), B, D, H, W, Cin, Cout,
              1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


SLAB2_SUPPORTED = {(8, 16), (16, 32)}      # (Cin, Cout) of the stride-2 TMA slab kernel (conv1, conv3)


@_on_device
def conv3d_s2_slab_pack(w, scale=None):
    """torch Conv3d weight [Cout,Cin,3,3,3] (x BN scale) -> operand image of the stride-2 TMA slab kernel."""
    Cout, Cin = w.shape[:2]
    n = _ffi.lib().rcmvs_conv3d_s2_slab_weight_elems(Cin, Cout)
    img = torch.empty((n,), device=w.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_conv3d_s2_slab_pack", _chk(w, "w"), _opt(scale, "scale"), img.data_ptr(), Cout, Cin, _stream())
    return img


@_on_device
def conv3d_s2_slab(x_split, img, Cout, bias=None, relu=True, want_y=True, want_split=False):
    """x_split [2,B,D,H,W,Cin] bf16 (D, H, W even) -> 3x3x3 stride-2 conv: y [B,D/2,H/2,W/2,Cout] fp32 and/or y_split."""
    _, B, D, H, W, Cin = x_split.shape
    shp = (B, D // 2, H // 2, W // 2, Cout)
    y = torch.empty(shp, device=x_split.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2,) + shp, device=x_split.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv3d_s2_slab_fwd", _chk_bf16(x_split, "x_split"), _chk_bf16(img, "wimg"), _opt(bias, "bias"),
              _opt(y, "y"), None if ys is None else ys.data_ptr(), B, D, H, W, Cin, Cout, 1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


SLABT_SUPPORTED = {(16, 8), (32, 16)}     # (Cin, Cout) of the TMA transposed-conv kernel (conv11, conv9)


@_on_device
def conv_transpose3d_slab_pack(w, scale=None):
    """torch ConvTranspose3d weight [Cin,Cout,3,3,3] (x BN scale) -> operand image of the TMA transposed-conv kernel."""
    Cin, Cout = w.shape[:2]
    n = _ffi.lib().rcmvs_conv_transpose3d_slab_weight_elems(Cin, Cout)
    img = torch.empty((n,), device=w.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_conv_transpose3d_slab_pack", _chk(w, "w"), _opt(scale, "scale"), img.data_ptr(), Cout, Cin, _stream())
    return img


@_on_device
def conv_transpose3d_slab(x_split, img, Cout, bias=None, skip=None, relu=True, want_y=True, want_split=False):
    """x_split [2,B,D,H,W,Cin] bf16 -> [B,2D,2H,2W,Cout] (k3, s2, p1, output_padding 1; + bias, ReLU, skip) as fp32 and/or
    split-bf16."""
    _, B, D, H, W, Cin = x_split.shape
    shp = (B, 2 * D, 2 * H, 2 * W, Cout)
    y = torch.empty(shp, device=x_split.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2,) + shp, device=x_split.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv_transpose3d_slab_fwd", _chk_bf16(x_split, "x_split"), _chk_bf16(img, "wimg"), _opt(bias, "bias"),
              _opt(skip, "skip"), _opt(y, "y"), None if ys is None else ys.data_ptr(), B, D, H, W, Cin, Cout,
              1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


def gemm_tc_supported(Cin, Cout):
    """Layers the gather implicit-GEMM kernel takes (the 1/4- and 1/8-resolution levels)."""
    return (Cin in (16, 32, 64) and Cout in (16, 32, 64)) or (Cin == 8 and Cout == 16)


@_on_device
def conv3d_gemm_tc_pack(w, scale=None, transposed=False, nsplit=3):
    """torch Conv3d [Cout,Cin,27] / ConvTranspose3d [Cin,Cout,27] weight (x BN scale) -> gather-GEMM operand image."""
    if transposed:
        Cin, Cout = w.shape[:2]
    else:
        Cout, Cin = w.shape[:2]
    n = _ffi.lib().rcmvs_conv3d_gemm_tc_weight_floats(Cin, Cout, nsplit)
    img = torch.empty((n,), device=w.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv3d_gemm_tc_pack", _chk(w, "w"), _opt(scale, "scale"), _chk(img, "img"), Cout, Cin,
              1 if transposed else 0, nsplit, _stream())
    return img


@_on_device
def conv3d_gemm_tc(x_cl, img, Cout, mode, bias=None, skip=None, nsplit=3, relu=True, x_lo=None, want_y=True,
                   want_split=False):
    """x_cl [B,D,H,W,Cin]; mode 0 conv s1 / 1 conv s2 / 2 transposed conv -> [B,Do,Ho,Wo,Cout] on tcgen05.mma.

    3xTF32 only: `x_lo` given -> (x_cl, x_lo) is the pre-split activation (hi, lo) written by the producing layer;
    `want_split` -> also return the result pre-split for the next gather layer.  Returns y, or (y|None, hi, lo)."""
    B, D, H, W, Cin = x_cl.shape
    if mode == 2:
        Do, Ho, Wo = 2 * D, 2 * H, 2 * W
    elif mode == 1:
        Do, Ho, Wo = (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    else:
        Do, Ho, Wo = D, H, W
    mk = lambda: torch.empty((B, Do, Ho, Wo, Cout), device=x_cl.device, dtype=torch.float32)
    y = mk() if (want_y or not want_split) else None
    hi, lo = (mk(), mk()) if want_split else (None, None)
    _ffi.call("rcmvs_conv3d_gemm_tc_fwd", _chk(x_cl, "x"), _opt(x_lo, "x_lo"), _chk(img, "wimg"), _opt(bias, "bias"),
              _opt(skip, "skip"), _opt(y, "y"), _opt(hi, "y_hi"), _opt(lo, "y_lo"), B, D, H, W, Cin, Cout, mode,
              nsplit, 1 if relu else 0, _stream())
    return (y, hi, lo) if want_split else y


TCT_SUPPORTED = {(16, 8), (32, 16)}      # (Cin, Cout) of conv11 / conv9 (transposed)


@_on_device
def conv_transpose3d_tc_pack(w, scale=None, nsplit=3):
    """torch ConvTranspose3d weight [Cin,Cout,3,3,3] (x BN scale) -> operand image of the tcgen05 convT kernel."""
    Cin, Cout = w.shape[:2]
    n = _ffi.lib().rcmvs_conv_transpose3d_tc_weight_floats(Cin, Cout, nsplit)
    img = torch.empty((n,), device=w.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv_transpose3d_tc_pack", _chk(w, "w"), _opt(scale, "scale"), _chk(img, "img"), Cout, Cin,
              nsplit, _stream())
    return img


@_on_device
def conv_transpose3d_tc(x_cl, img, Cout, bias=None, skip=None, nsplit=3, relu=True):
    """x_cl [B,D,H,W,Cin] -> [B,2D,2H,2W,Cout] on tcgen05.mma (k3, s2, p1, output_padding 1)."""
    B, D, H, W, Cin = x_cl.shape
    y = torch.empty((B, 2 * D, 2 * H, 2 * W, Cout), device=x_cl.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv_transpose3d_tc_fwd", _chk(x_cl, "x"), _chk(img, "wimg"), _opt(bias, "bias"),
              _opt(skip, "skip"), _chk(y, "y"), B, D, H, W, Cin, Cout, nsplit, 1 if relu else 0, _stream())
    return y


# ---- regression -----------------------------------------------------------------------------------------------
@_on_device
def softargmin_conf(cost, planes, want_idx=False, want_prob=False):
    """cost [B,D,h,w]; planes [B,D] or [B,D,h,w] -> depth [B,h,w], conf [B,h,w] (, idx int32, prob)."""
    B, D, h, w = cost.shape
    depth = torch.empty((B, h, w), device=cost.device, dtype=torch.float32)
    conf = torch.empty_like(depth)
    idx = torch.empty((B, h, w), device=cost.device, dtype=torch.int32) if want_idx else None
    prob = torch.empty_like(cost) if want_prob else None
    _ffi.call("rcmvs_softargmin_conf_fwd", _chk(cost, "cost"), _chk(planes, "planes"), _depth_mode(planes, B),
              _chk(depth, "depth"), _chk(conf, "conf"), None if idx is None else idx.data_ptr(),
              _opt(prob, "prob"), B, D, h, w, _stream())
    out = [depth, conf]
    if want_idx:
        out.append(idx)
    if want_prob:
        out.append(prob)
    return tuple(out)


# ---- renderer (a7-a12) ------------------------------------------------------------------------------------------
@_on_device
def resize_depth_cl(x_cl, D1, Cpad):
    """x_cl [B,D0,h,w,C] -> [B,D1,h,w,Cpad] (linear along D, align_corners=True, zero-padded channels)."""
    B, D0, h, w, C = x_cl.shape
    y = torch.empty((B, D1, h, w, Cpad), device=x_cl.device, dtype=torch.float32)
    _ffi.call("rcmvs_resize_depth_cl", _chk(x_cl, "x"), _chk(y, "y"), B, D0, D1, h * w, C, Cpad, _stream())
    return y


@_on_device
def resize_depth_cl_split(x_cl, D1, Cpad):
    """Like resize_depth_cl, written as split-bf16 planes [2,B,D1,h,w,Cpad] for the TMA-fed tensor-core conv0."""
    B, D0, h, w, C = x_cl.shape
    y = torch.empty((2, B, D1, h, w, Cpad), device=x_cl.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_resize_depth_cl_split", _chk(x_cl, "x"), y.data_ptr(), B, D0, D1, h * w, C, Cpad, _stream())
    return y


@_on_device
def gu_sample(ray_depth, near, far, seed, n_gauss, S=128):
    N = ray_depth.shape[0]
    out = torch.empty((N, S), device=ray_depth.device, dtype=torch.float32)
    _ffi.call("rcmvs_gu_sample", _chk(ray_depth, "ray_depth"), float(near), float(far), int(seed) & (2 ** 64 - 1),
              _chk(out, "out"), N, n_gauss, S, _stream())
    return out


@_on_device
def ray_points(ys, xs, K, c2w, w2c_ref, z, near, far, W, H):
    N, S = z.shape
    pts = torch.empty((N, S, 3), device=z.device, dtype=torch.float32)
    ndc = torch.empty_like(pts)
    dirs = torch.empty((N, 3), device=z.device, dtype=torch.float32)
    _ffi.call("rcmvs_ray_points", _chk(ys, "ys"), _chk(xs, "xs"), _chk(K, "K"), _chk(c2w, "c2w"),
              _chk(w2c_ref, "w2c_ref"), _chk(z, "z"), float(near), float(far), W, H, _chk(pts, "pts"),
              _chk(ndc, "ndc"), _chk(dirs, "dirs"), N, S, _stream())
    return pts, ndc, dirs


@_on_device
def point_features(vol_cl, ndc, pts, imgs3, w2cs3, Ks3):
    """vol_cl [D,h,w,8]; ndc/pts [N,S,3]; imgs3 [3,3,H,W]; w2cs3 [3,4,4]; Ks3 [3,3,3] -> [N,S,20]."""
    D, h, w, _ = vol_cl.shape
    H, W = imgs3.shape[-2:]
    N, S = ndc.shape[:2]
    out = torch.empty((N, S, 20), device=ndc.device, dtype=torch.float32)
    _ffi.call("rcmvs_point_features", _chk(vol_cl, "vol"), _chk(ndc, "ndc"), _chk(pts, "pts"), _chk(imgs3, "imgs"),
              _chk(w2cs3, "w2cs"), _chk(Ks3, "Ks"), _chk(out, "out"), N * S, D, h, w, H, W, _stream())
    return out


@_on_device
def render_mlp(ndc, feat, dirs, weight_ptrs):
    """ndc [N,S,3], feat [N,S,20], dirs [N,3]; weight_ptrs: int64 CUDA... a HOST ctypes array of 22 device pointers."""
    N, S = ndc.shape[:2]
    out = torch.empty((N, S, 4), device=ndc.device, dtype=torch.float32)
    _ffi.call("rcmvs_render_mlp", _chk(ndc, "ndc"), _chk(feat, "feat"), _chk(dirs, "dirs"), weight_ptrs,
              _chk(out, "out"), N * S, S, _stream())
    return out


def render_mlp_tc_pack(weight_ptrs, device):
    """Split-bf16 weight images of the tensor-core MLP from the 22 fp32 weight pointers of render_mlp (on `device`)."""
    with torch.cuda.device(device):
        img = torch.empty((_ffi.lib().rcmvs_render_mlp_tc_weight_elems(),), device=device, dtype=torch.bfloat16)
        _ffi.call("rcmvs_render_mlp_tc_pack", weight_ptrs, img.data_ptr(), _stream())
    return img


@_on_device
def render_mlp_tc(ndc, feat, dirs, weight_ptrs, wimg):
    """Same contract as render_mlp, Linear layers on tcgen05 (split-bf16, fp32-equivalent)."""
    N, S = ndc.shape[:2]
    out = torch.empty((N, S, 4), device=ndc.device, dtype=torch.float32)
    ws = torch.empty((_ffi.lib().rcmvs_render_mlp_tc_workspace_bytes(N * S),), device=ndc.device, dtype=torch.uint8)
    _ffi.call("rcmvs_render_mlp_tc", _chk(ndc, "ndc"), _chk(feat, "feat"), _chk(dirs, "dirs"), weight_ptrs, _chk_bf16(wimg, "wimg"),
              ws.data_ptr(), _chk(out, "out"), N * S, S, _stream())
    return out


@_on_device
def composite(raw, z):
    N, S = z.shape
    rgb = torch.empty((N, 3), device=z.device, dtype=torch.float32)
    depth = torch.empty((N,), device=z.device, dtype=torch.float32)
    weights = torch.empty((N, S), device=z.device, dtype=torch.float32)
    alpha = torch.empty_like(weights)
    _ffi.call("rcmvs_composite", _chk(raw, "raw"), _chk(z, "z"), _chk(rgb, "rgb"), _chk(depth, "depth"),
              _chk(weights, "weights"), _chk(alpha, "alpha"), N, S, _stream())
    return rgb, depth, weights, alpha


# ---- a13 ------------------------------------------------------------------------------------------------------------
@_on_device
def inverse_warp_fwd(img, left_cam, right_cam, depth):
    B, h, w, C = img.shape
    warped = torch.empty_like(img)
    mask = torch.empty((B, h, w, 1), device=img.device, dtype=torch.float32)
    _ffi.call("rcmvs_inverse_warp_fwd", _chk(img, "img"), _chk(left_cam, "left_cam"), _chk(right_cam, "right_cam"),
              _chk(depth, "depth"), _chk(warped, "warped"), _chk(mask, "mask"), B, h, w, C, _stream())
    return warped, mask


@_on_device
def inverse_warp_bwd(img, left_cam, right_cam, depth, grad_warped):
    B, h, w, C = img.shape
    gd = torch.empty((B, h, w), device=img.device, dtype=torch.float32)
    _ffi.call("rcmvs_inverse_warp_bwd", _chk(img, "img"), _chk(left_cam, "left_cam"), _chk(right_cam, "right_cam"),
              _chk(depth, "depth"), _chk(grad_warped, "grad_warped"), _chk(gd, "grad_depth"), B, h, w, C, _stream())
    return gd


# ---- f1 FeatureNet ------------------------------------------------------------------------------------------------
@_on_device
def images_to_nhwc8(img):
    N, C, H, W = img.shape
    out = torch.empty((N, H, W, 8), device=img.device, dtype=torch.float32)
    _ffi.call("rcmvs_images_to_nhwc8", _chk(img, "img"), _chk(out, "out"), N, C, H, W, _stream())
    return out


@_on_device
def images_to_nhwc8_split(img):
    """[N,C<=8,H,W] fp32 -> split-bf16 [2,N,H,W,8] (channels C..7 zero)."""
    N, C, H, W = img.shape
    out = torch.empty((2, N, H, W, 8), device=img.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_images_to_nhwc8_split", _chk(img, "img"), out.data_ptr(), N, C, H, W, _stream())
    return out


@_on_device
def pack_conv2d_weight(w, scale=None, cin_pad=None):
    Cout, Cin, K, _ = w.shape
    cin_pad = cin_pad or Cin
    packed = torch.empty((K * K, cin_pad, Cout), device=w.device, dtype=torch.float32)
    _ffi.call("rcmvs_pack_conv2d_weight", _chk(w, "w"), _opt(scale, "scale"), _chk(packed, "packed"), Cout, Cin,
              cin_pad, K, _stream())
    return packed


@_on_device
def conv2d(x_cl, packed_w, bias=None, up=None, stride=1, relu=True, want_y=True, want_split=False):
    """x_cl [N,H,W,Cin] -> [N,Ho,Wo,Cout]: conv(K, pad K/2) + bias, ReLU, then + nearest-2x-upsampled `up`.
    want_split: also (or only) the split-bf16 form [2,N,Ho,Wo,Cout] for the tensor-core layers."""
    N, H, W, Cin = x_cl.shape
    KK, _, Cout = packed_w.shape
    K = {1: 1, 9: 3, 25: 5}[KK]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty((N, Ho, Wo, Cout), device=x_cl.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2, N, Ho, Wo, Cout), device=x_cl.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv2d_fwd_split", _chk(x_cl, "x"), _chk(packed_w, "w"), _opt(bias, "bias"), _opt(up, "up"),
              _opt(y, "y"), None if ys is None else ys.data_ptr(), N, H, W, Cin, Cout, K, stride, 1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


SLAB2D_SUPPORTED = {(8, 8), (16, 16), (32, 32), (32, 16), (32, 8)}      # (Cin, Cout) of the 2-D TMA slab kernel (3x3, stride 1)


@_on_device
def conv2d_slab_pack(w, scale=None):
    Cout, Cin = w.shape[:2]
    img = torch.empty((_ffi.lib().rcmvs_conv2d_slab_weight_elems(Cin, Cout),), device=w.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_conv2d_slab_pack", _chk(w, "w"), _opt(scale, "scale"), img.data_ptr(), Cout, Cin, _stream())
    return img


@_on_device
def conv2d_slab(x_split, img, Cout, bias=None, relu=True, want_y=True, want_split=False):
    """x_split [2,N,H,W,Cin] bf16 -> y [N,H,W,Cout] fp32 and/or y_split [2,N,H,W,Cout]: 3x3 stride-1 conv on tcgen05."""
    _, N, H, W, Cin = x_split.shape
    y = torch.empty((N, H, W, Cout), device=x_split.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2, N, H, W, Cout), device=x_split.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv2d_slab_fwd", _chk_bf16(x_split, "x_split"), _chk_bf16(img, "wimg"), _opt(bias, "bias"), _opt(y, "y"),
              None if ys is None else ys.data_ptr(), N, H, W, Cin, Cout, 1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


S2K5_SUPPORTED = {(8, 16), (16, 32)}      # (Cin, Cout) of the 5x5 stride-2 tensor-core kernel (FeatureNet conv1.0 / conv2.0)


@_on_device
def conv2d_s2k5_slab_pack(w, scale=None):
    Cout, Cin = w.shape[:2]
    img = torch.empty((_ffi.lib().rcmvs_conv2d_s2k5_slab_weight_elems(Cin, Cout),), device=w.device, dtype=torch.bfloat16)
    _ffi.call("rcmvs_conv2d_s2k5_slab_pack", _chk(w, "w"), _opt(scale, "scale"), img.data_ptr(), Cout, Cin, _stream())
    return img


@_on_device
def conv2d_s2k5_slab(x_split, img, Cout, bias=None, relu=True, want_y=True, want_split=False):
    """x_split [2,N,H,W,Cin] bf16 (H, W even) -> 5x5 stride-2 pad-2 conv: y [N,H/2,W/2,Cout] fp32 and/or y_split."""
    _, N, H, W, Cin = x_split.shape
    y = torch.empty((N, H // 2, W // 2, Cout), device=x_split.device, dtype=torch.float32) if want_y else None
    ys = torch.empty((2, N, H // 2, W // 2, Cout), device=x_split.device, dtype=torch.bfloat16) if want_split else None
    _ffi.call("rcmvs_conv2d_s2k5_slab_fwd", _chk_bf16(x_split, "x_split"), _chk_bf16(img, "wimg"), _opt(bias, "bias"),
              _opt(y, "y"), None if ys is None else ys.data_ptr(), N, H, W, Cin, Cout, 1 if relu else 0, _stream())
    if want_y and want_split:
        return y, ys
    return y if want_y else ys


# ---- f2 unsupervised loss stack ---------------------------------------------------------------------------------------
@_on_device
def images_subsample_nhwc(imgs, step):
    """imgs [N,3,H,W] -> [N,H/step,W/step,3]: nearest down-sampling by an integer factor + NCHW -> NHWC."""
    N, C, H, W = imgs.shape
    if C != 3:
        raise RuntimeError("rcmvs: images_subsample_nhwc expects 3 channels")
    out = torch.empty((N, H // step, W // step, 3), device=imgs.device, dtype=torch.float32)
    _ffi.call("rcmvs_images_subsample_nhwc", _chk(imgs, "imgs"), _chk(out, "out"), N, H, W, step, _stream())
    return out


@_on_device
def unsup_loss_fwd_bwd(imgs_nhwc, cams, depth, smooth_weight, want_grad=True):
    """imgs_nhwc [B,V,h,w,3], cams [B,V,2,4,4], depth [B,h,w] -> (terms [4] = total, reconstr, ssim, smooth; grad_depth)."""
    B, V, h, w, _ = imgs_nhwc.shape
    terms = torch.empty((4,), device=depth.device, dtype=torch.float32)
    grad = torch.empty((B, h, w), device=depth.device, dtype=torch.float32) if want_grad else None
    nws = _ffi.lib().rcmvs_unsup_loss_workspace_floats(B, V, h, w)
    ws = torch.empty((nws,), device=depth.device, dtype=torch.float32)
    _ffi.call("rcmvs_unsup_loss_fwd_bwd", _chk(imgs_nhwc, "imgs"), _chk(cams, "cams"), _chk(depth, "depth"),
              float(smooth_weight), _chk(terms, "terms"), _opt(grad, "grad_depth"), _chk(ws, "workspace"), B, V, h, w,
              _stream())
    _ffi.launches += 10 + 4 * (V - 1) - 1 if want_grad else 3 + 2 * (V - 1) - 1
    return terms, grad


@_on_device
def masked_loss_fwd_bwd(pred, target, mask=None, mask_kind=0, mode=0, scale=1.0, step=1, want_grad=True):
    """scale * mean over selected elements of smooth-L1 (mode 0) / squared error (mode 1) of pred [B,h,w] vs target
    [B,h*step,w*step] sampled at stride `step`; mask_kind 0 all / 1 mask [B,(C,)H,W] (first channel) > 0.5 / 2 target > 0.
    Returns (loss [1], grad [B,h,w] or None)."""
    B, h, w = pred.shape
    loss = torch.empty((1,), device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    ws = torch.empty((_ffi.lib().rcmvs_masked_loss_workspace_floats(),), device=pred.device, dtype=torch.float32)
    mstride = 0
    if mask is not None:
        mstride = mask.stride(0)
    _ffi.call("rcmvs_masked_loss_fwd_bwd", _chk(pred, "pred"), _chk(target, "target"), _opt(mask, "mask"), mstride,
              mask_kind, mode, float(scale), _chk(loss, "loss"), _opt(grad, "grad"), _chk(ws, "workspace"), B, h, w, step,
              _stream())
    _ffi.launches += 2 if want_grad else 1
    return loss, grad
