"""torch.autograd.Function wrappers that put the TRAINING path (forward with batch-statistics BatchNorm + backward) of
the reference's building blocks on librcmvs kernels (csrc/train.cu + the forward convolution kernels):

    conv_bn_act3d / conv_bn_act2d   Conv3d / Deconv3d / Conv2d blocks = conv -> BatchNorm -> ReLU (+ skip) (reference
                                    models/modules.py:28-62, 118-210), plain convolutions (`prob`, FPN heads) likewise
    costvolume_variance             fused plane sweep + variance (+ var_no_ref), gradient w.r.t. all V feature maps
    softargmin                      softmax over depth + depth regression, gradient w.r.t. the cost volume
    up2_add / to_channels_last / to_planar   FPN lateral add and the layout changes, with their adjoints

torch is used for what it is here: tensor allocation, the autograd graph and streams.  No cuDNN / cuBLAS / ATen compute
kernel runs inside any of these functions (RCMVS_STRICT=1 keeps the rest of the model honest).
"""
import torch

from . import _ffi, ops
from .ops import _chk, _opt, _stream, _on_device


SYNC_BN = False      # True: batch statistics (forward) and their gradient sums (backward) are all-reduced over the ranks
                     # of the default process group = nn.SyncBatchNorm semantics (train_rcmvsnet.py:524-525)


def _f32c(t):
    return t.detach().float().contiguous()


def _sync_world():
    import torch.distributed as dist
    if SYNC_BN and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_world_size()
    return None, 1


# ---- raw ops -------------------------------------------------------------------------------------------------------
@_on_device
def bn_batch_stats(x2d, running_mean, running_var, momentum, eps):
    """x2d [M,C] -> (mean, invstd); running statistics updated in place like nn.BatchNorm in train mode."""
    M, C = x2d.shape
    stats = torch.zeros((2 * C,), device=x2d.device, dtype=torch.float64)
    _ffi.call("rcmvs_bn_stats", _chk(x2d, "x"), stats.data_ptr(), M, C, _stream())
    dist, world = _sync_world()
    if dist is not None:                                  # SyncBatchNorm: 2·C doubles over NCCL, equal rows per rank
        dist.all_reduce(stats)
    mean = torch.empty((C,), device=x2d.device, dtype=torch.float32)
    invstd = torch.empty_like(mean)
    _ffi.call("rcmvs_bn_finalize", stats.data_ptr(), _chk(mean, "mean"), _chk(invstd, "invstd"),
              _opt(running_mean, "running_mean"), _opt(running_var, "running_var"), M * world, C, float(eps),
              float(momentum), _stream())
    return mean, invstd, stats


@_on_device
def bn_apply(x2d, mean, invstd, gamma, beta, skip, relu):
    M, C = x2d.shape
    z = torch.empty_like(x2d)
    _ffi.call("rcmvs_bn_apply", _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"), _chk(gamma, "gamma"),
              _chk(beta, "beta"), _opt(skip, "skip"), _chk(z, "z"), M, C, 1 if relu else 0, _stream())
    return z


@_on_device
def bn_backward(dz2d, x2d, mean, invstd, gamma, beta, relu, frozen):
    M, C = x2d.shape
    sums = torch.zeros((2 * C,), device=x2d.device, dtype=torch.float64)
    dx = torch.empty_like(x2d)
    dgamma = torch.empty((C,), device=x2d.device, dtype=torch.float32)
    dbeta = torch.empty_like(dgamma)
    dist, world = _sync_world()
    if dist is None or frozen:
        _ffi.call("rcmvs_bn_bwd", _chk(dz2d, "dz"), _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"),
                  _chk(gamma, "gamma"), _chk(beta, "beta"), sums.data_ptr(), _chk(dx, "dx"), _chk(dgamma, "dgamma"),
                  _chk(dbeta, "dbeta"), M, C, 1 if relu else 0, 1 if frozen else 0, _stream())
        _ffi.launches += 1
        return dx, dgamma, dbeta
    # SyncBatchNorm backward: the correction terms use the sums over ALL ranks; dgamma / dbeta stay local (DDP's gradient
    # all-reduce averages them like every other parameter gradient)
    _ffi.call("rcmvs_bn_bwd_reduce", _chk(dz2d, "dz"), _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"),
              _chk(gamma, "gamma"), _chk(beta, "beta"), sums.data_ptr(), M, C, 1 if relu else 0, _stream())
    local = sums.clone()
    dist.all_reduce(sums)
    _ffi.call("rcmvs_bn_bwd_apply", _chk(dz2d, "dz"), _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"),
              _chk(gamma, "gamma"), _chk(beta, "beta"), sums.data_ptr(), _chk(dx, "dx"), None, None, M, M * world, C,
              1 if relu else 0, 0, _stream())
    return dx, local[C:].float(), local[:C].float()


@_on_device
def channel_sums(x2d):
    """Σ over rows of x2d [M,C] (bias gradients)."""
    M, C = x2d.shape
    stats = torch.zeros((2 * C,), device=x2d.device, dtype=torch.float64)
    _ffi.call("rcmvs_bn_stats", _chk(x2d, "x"), stats.data_ptr(), M, C, _stream())
    return stats[:C].float()


@_on_device
def conv_wgrad(S, Bg, k, stride, dim3d):
    """S [B,(Ds,)Hs,Ws,Cs], Bg [B,(Db,)Hb,Wb,Cb] -> G [Cs,Cb,taps] (see include/rcmvs.h)."""
    if dim3d:
        B, Ds, Hs, Ws, Cs = S.shape
        _, Db, Hb, Wb, Cb = Bg.shape
    else:
        B, Hs, Ws, Cs = S.shape
        _, Hb, Wb, Cb = Bg.shape
        Ds = Db = 1
    taps = k ** 3 if dim3d else k * k
    G = torch.zeros((Cs, Cb, taps), device=S.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv_wgrad", _chk(S, "S"), _chk(Bg, "Bg"), _chk(G, "G"), B, Ds, Hs, Ws, Cs, Db, Hb, Wb, Cb, k, stride,
              1 if dim3d else 0, _stream())
    return G


@_on_device
def conv2d_dgrad(dy, w, H, W, stride):
    """dy [N,Ho,Wo,Cout], torch weight w [Cout,Cin,k,k] -> dx [N,H,W,Cin]."""
    N, Ho, Wo, Cout = dy.shape
    Cin, k = w.shape[1], w.shape[2]
    dx = torch.empty((N, H, W, Cin), device=dy.device, dtype=torch.float32)
    ws = torch.empty((k * k * Cout * Cin,), device=dy.device, dtype=torch.float32)
    _ffi.call("rcmvs_conv2d_dgrad", _chk(dy, "dy"), _chk(w, "w"), _chk(ws, "ws"), _chk(dx, "dx"), N, H, W, Cin, Cout, k, stride,
              _stream())
    _ffi.launches += 1
    return dx


@_on_device
def conv3d_dgrad(dy, w, transposed, stride):
    """Data gradient of Conv3d / ConvTranspose3d(k3,p1[,s2,op1]) through the forward kernels with re-packed weights."""
    if transposed:                                   # adjoint of a transposed conv = stride-2 conv, same weight tensor
        return ops.conv3d(dy, ops.pack_conv3d_weight(w, None, False), None, None, stride=2, relu=False)
    if stride == 2:                                  # adjoint of a stride-2 conv = transposed conv, same weight tensor
        return ops.conv_transpose3d(dy, ops.pack_conv3d_weight(w, None, True), None, None, relu=False)
    Cout, Cin = w.shape[:2]
    if Cout == 1:                                    # `prob`: 8 <- 1 channel
        B, D, H, W, _ = dy.shape
        dx = torch.empty((B, D, H, W, Cin), device=dy.device, dtype=torch.float32)
        _ffi.call("rcmvs_conv3d_c1_dgrad", _chk(dy, "dcost"), _chk(w, "w"), _chk(dx, "dx"), B, D, H, W, _stream())
        return dx
    packed = torch.empty((27, Cout, Cin), device=w.device, dtype=torch.float32)
    _ffi.call("rcmvs_pack_conv3d_dgrad_s1", _chk(w, "w"), _chk(packed, "packed"), Cout, Cin, _stream())
    return ops.conv3d(dy, packed, None, None, stride=1, relu=False)


# ---- conv -> BatchNorm -> ReLU (+ skip) blocks ------------------------------------------------------------------------
class _ConvBNAct(torch.autograd.Function):
    """x_cl [B,D,H,W,Cin] (3-D) or [N,H,W,Cin] (2-D); weight in torch layout; bn = None for a plain convolution."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, skip, cfg):
        dim3d, transposed, stride, relu = cfg["dim3d"], cfg["transposed"], cfg["stride"], cfg["relu"]
        xc, w = _f32c(x), _f32c(weight)
        if dim3d:
            packed = ops.pack_conv3d_weight(w, None, transposed)
            y = ops.conv_transpose3d(xc, packed, None, None, relu=False) if transposed else \
                ops.conv3d(xc, packed, None, None, stride=stride, relu=False)
        else:
            cin_pad = xc.shape[-1]
            y = ops.conv2d(xc, ops.pack_conv2d_weight(w, None, cin_pad), _f32c(bias) if bias is not None else None, None,
                           stride=stride, relu=False)
        C = y.shape[-1]
        y2 = y.view(-1, C)
        ctx.cfg, ctx.has_bn = cfg, gamma is not None
        if gamma is not None:
            g, b = _f32c(gamma), _f32c(beta)
            if cfg["batch_stats"]:
                mean, invstd, _ = bn_batch_stats(y2, cfg["running_mean"], cfg["running_var"], cfg["momentum"], cfg["eps"])
            else:
                mean = cfg["running_mean"].detach().float().contiguous()
                invstd = torch.rsqrt(cfg["running_var"].detach().float() + cfg["eps"]).contiguous()
            z = bn_apply(y2, mean, invstd, g, b, None if skip is None else _f32c(skip).view(-1, C), relu).view(y.shape)
            ctx.save_for_backward(xc, w, y, mean, invstd, g, b)
        else:
            z = y
            if skip is not None:
                raise RuntimeError("rcmvs: skip without BatchNorm is not a block of the reference")
            ctx.save_for_backward(xc, w)
        ctx.has_skip = skip is not None
        ctx.has_bias = bias is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        cfg = ctx.cfg
        dim3d, transposed, stride, relu = cfg["dim3d"], cfg["transposed"], cfg["stride"], cfg["relu"]
        dz = dz.contiguous()
        C = dz.shape[-1]
        dgamma = dbeta = dbias = None
        if ctx.has_bn:
            xc, w, y, mean, invstd, g, b = ctx.saved_tensors
            dy2, dgamma, dbeta = bn_backward(dz.view(-1, C), y.view(-1, C), mean, invstd, g, b, relu, not cfg["batch_stats"])
            dy = dy2.view(dz.shape)
        else:
            xc, w = ctx.saved_tensors
            dy = dz
            if ctx.has_bias:
                dbias = channel_sums(dy.view(-1, C))
        dx = None
        if ctx.needs_input_grad[0]:
            if dim3d:
                dx = conv3d_dgrad(dy, w, transposed, stride)
            else:
                dx = conv2d_dgrad(dy, w, xc.shape[1], xc.shape[2], stride)
                if dx.shape[-1] != xc.shape[-1]:          # zero-padded input channels (never differentiated: the image)
                    dx = torch.nn.functional.pad(dx, (0, xc.shape[-1] - dx.shape[-1]))
        k = w.shape[2]
        if dim3d and transposed:
            G = conv_wgrad(xc, dy, 3, 2, True)                      # [Cin_t, Cout_t, 27]
        elif dim3d:
            G = conv_wgrad(dy, xc, 3, stride, True)                 # [Cout, Cin, 27]
        else:
            G = conv_wgrad(dy, xc, k, stride, False)[:, :w.shape[1]]  # [Cout, Cin(_pad -> Cin), k*k]
        dw = G.reshape(w.shape) if G.is_contiguous() else G.contiguous().view(w.shape)
        return dx, dw, dbias, dgamma, dbeta, (dz if ctx.has_skip else None), None


def conv_block(x_cl, conv, bn=None, skip=None, relu=True, training=True):
    """One reference block on own kernels, differentiable.  conv: nn.Conv3d / nn.ConvTranspose3d / nn.Conv2d; bn: its
    BatchNorm or None.  x_cl channels-last."""
    dim3d = conv.weight.dim() == 5
    transposed = isinstance(conv, torch.nn.ConvTranspose3d)
    cfg = {"dim3d": dim3d, "transposed": transposed, "stride": conv.stride[0], "relu": bool(relu) and bn is not None, "batch_stats": bool(training), "momentum": 0.1, "eps": 1e-5,
           "running_mean": None, "running_var": None}
    gamma = beta = None
    if bn is not None:
        cfg["momentum"] = bn.momentum if bn.momentum is not None else 0.1
        cfg["eps"] = bn.eps
        cfg["running_mean"], cfg["running_var"] = bn.running_mean, bn.running_var
        gamma, beta = bn.weight, bn.bias
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return _ConvBNAct.apply(x_cl, conv.weight, conv.bias, gamma, beta, skip, cfg)


# ---- cost volume / regression ------------------------------------------------------------------------------------------
class _CostVolumeVar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats_cl, rt, depth, want_noref):
        f = _f32c(feats_cl)
        res = ops.costvolume_variance(f, rt, depth, want_noref)
        ctx.save_for_backward(f, rt, depth)
        ctx.want_noref = want_noref
        if want_noref:
            return res[0], res[1]
        return res, torch.zeros((), device=f.device)

    @staticmethod
    @_on_device
    def backward(ctx, gvar, gnoref):
        f, rt, depth = ctx.saved_tensors
        B, V, h, w, C = f.shape
        D = depth.shape[1]
        gf = torch.zeros_like(f)
        gvar = gvar.contiguous()
        gn = gnoref.contiguous() if ctx.want_noref else None
        _ffi.call("rcmvs_costvolume_var_bwd", _chk(f, "feats"), _chk(rt, "rt"), _chk(depth, "depth"), ops._depth_mode(depth, B),
                  _chk(gvar, "grad_var"), _opt(gn, "grad_noref"), _chk(gf, "grad_feats"), B, V, C, D, h, w, _stream())
        return gf, None, None, None


def costvolume_variance(feats_cl, rt, depth, want_noref=False):
    var, noref = _CostVolumeVar.apply(feats_cl, rt, depth, want_noref)
    return (var, noref) if want_noref else var


class _SoftArgmin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost, planes):
        c = _f32c(cost)
        depth, conf = ops.softargmin_conf(c, planes)
        ctx.save_for_backward(c, planes)
        ctx.mark_non_differentiable(conf)
        return depth, conf

    @staticmethod
    @_on_device
    def backward(ctx, gdepth, _gconf):
        c, planes = ctx.saved_tensors
        B, D, h, w = c.shape
        gc = torch.empty_like(c)
        _ffi.call("rcmvs_softargmin_bwd", _chk(c, "cost"), _chk(planes, "planes"), ops._depth_mode(planes, B),
                  _chk(gdepth.contiguous(), "grad_depth"), _chk(gc, "grad_cost"), B, D, h, w, _stream())
        return gc, None


def softargmin(cost, planes):
    """cost [B,D,h,w], planes [B,D] / [B,D,h,w] (no gradient: grad_method='detach') -> (depth, photometric confidence)."""
    return _SoftArgmin.apply(cost, planes)


# ---- layout / FPN glue ---------------------------------------------------------------------------------------------------
class _ToCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.to_channels_last(_f32c(x))

    @staticmethod
    def backward(ctx, g):
        return ops.to_planar(g.contiguous())


class _ToPlanar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.to_planar(_f32c(x))

    @staticmethod
    def backward(ctx, g):
        return ops.to_channels_last(g.contiguous())


def to_channels_last(x):
    return _ToCL.apply(x)


def to_planar(x):
    return _ToPlanar.apply(x)


class _Up2Add(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, a, up):
        a, up = _f32c(a), _f32c(up)
        N, H, W, C = a.shape
        out = torch.empty_like(a)
        _ffi.call("rcmvs_up2_add", _chk(a, "a"), _chk(up, "up"), _chk(out, "out"), N, H, W, C, _stream())
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        g = g.contiguous()
        N, H, W, C = g.shape
        gu = torch.empty((N, H // 2, W // 2, C), device=g.device, dtype=torch.float32)
        _ffi.call("rcmvs_sum2x2", _chk(g, "in"), _chk(gu, "out"), N, H // 2, W // 2, C, _stream())
        return g, gu


def up2_add(a, up):
    """a [N,H,W,C] + nearest-2x-upsample(up [N,H/2,W/2,C]) (FPN lateral, modules.py:451,455)."""
    return _Up2Add.apply(a, up)


# ---- optimiser ------------------------------------------------------------------------------------------------------------
@_on_device
def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """In-place fused Adam over flat fp32 buffers (torch.optim.Adam semantics)."""
    _ffi.call("rcmvs_adam_step", _chk(params, "params"), _chk(grads, "grads"), _chk(exp_avg, "exp_avg"),
              _chk(exp_avg_sq, "exp_avg_sq"), params.numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
              float(weight_decay), int(step), float(grad_scale), _stream())
