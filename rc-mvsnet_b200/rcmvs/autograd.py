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


GRAD_IN_PLACE = False  # True (rcmvs.train.TrainStep): weight / gamma / beta gradients of conv_block are accumulated by the
                       # kernels straight into the parameters' existing .grad buffers (autograd receives None for them) — no
                       # zero-fill + AccumulateGrad add per parameter use, ~1200 tiny launches per training step
STATS_ARENA = None     # [buffer of zeroed doubles, cursor] while rcmvs.train.TrainStep runs a step: the 2·C-double workspaces of
                       # the BatchNorm kernels are slices of ONE buffer zeroed once per step instead of ~300 torch.zeros launches
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
    stats = _zeros64(2 * C, x2d.device)
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


def stats_arena_begin(dev, n=1 << 17):
    """Zero the step's workspace arena (allocated once per device) and make _zeros64 hand out slices of it."""
    global STATS_ARENA
    buf = _ARENAS.get(dev)
    if buf is None or buf.numel() < n:
        buf = _ARENAS[dev] = torch.empty((n,), device=dev, dtype=torch.float64)
    buf.zero_()
    STATS_ARENA = [buf, 0]


def stats_arena_end():
    global STATS_ARENA
    STATS_ARENA = None


_ARENAS = {}


def _zeros64(n, dev):
    a = STATS_ARENA
    if a is not None and a[0].device == dev and a[1] + n <= a[0].numel():
        out = a[0][a[1]:a[1] + n]
        a[1] += n
        return out
    return torch.zeros((n,), device=dev, dtype=torch.float64)


@_on_device
def bn_backward(dz2d, x2d, mean, invstd, gamma, beta, relu, frozen, grad_into=None):
    """-> dx, dgamma, dbeta.  grad_into = (gamma.grad, beta.grad): the kernel adds into these and (None, None) comes back."""
    M, C = x2d.shape
    sums = _zeros64(2 * C, x2d.device)
    dx = torch.empty_like(x2d)
    dist, world = _sync_world()
    if dist is None or frozen:
        if grad_into is not None:
            dgamma, dbeta = grad_into
        else:
            dgamma = torch.empty((C,), device=x2d.device, dtype=torch.float32)
            dbeta = torch.empty_like(dgamma)
        _ffi.call("rcmvs_bn_bwd", _chk(dz2d, "dz"), _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"),
                  _chk(gamma, "gamma"), _chk(beta, "beta"), sums.data_ptr(), _chk(dx, "dx"), _chk(dgamma, "dgamma"),
                  _chk(dbeta, "dbeta"), M, C, 1 if relu else 0, (1 if frozen else 0) | (2 if grad_into is not None else 0),
                  _stream())
        _ffi.launches += 1
        return (dx, None, None) if grad_into is not None else (dx, dgamma, dbeta)
    # SyncBatchNorm backward: the correction terms use the sums over ALL ranks; dgamma / dbeta stay local (DDP's gradient
    # all-reduce averages them like every other parameter gradient)
    _ffi.call("rcmvs_bn_bwd_reduce", _chk(dz2d, "dz"), _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"),
              _chk(gamma, "gamma"), _chk(beta, "beta"), sums.data_ptr(), M, C, 1 if relu else 0, _stream())
    local = sums.clone()
    dist.all_reduce(sums)
    _ffi.call("rcmvs_bn_bwd_apply", _chk(dz2d, "dz"), _chk(x2d, "x"), _chk(mean, "mean"), _chk(invstd, "invstd"),
              _chk(gamma, "gamma"), _chk(beta, "beta"), sums.data_ptr(), _chk(dx, "dx"), None, None, M, M * world, C,
              1 if relu else 0, 0, _stream())
    if grad_into is not None:
        grad_into[0].add_(local[C:].float())
        grad_into[1].add_(local[:C].float())
        return dx, None, None
    return dx, local[C:].float(), local[:C].float()


@_on_device
def channel_sums(x2d):
    """Σ over rows of x2d [M,C] (bias gradients)."""
    M, C = x2d.shape
    stats = _zeros64(2 * C, x2d.device)
    _ffi.call("rcmvs_bn_stats", _chk(x2d, "x"), stats.data_ptr(), M, C, _stream())
    return stats[:C].float()


@_on_device
def conv_wgrad(S, Bg, k, stride, dim3d, out=None):
    """S [B,(Ds,)Hs,Ws,Cs], Bg [B,(Db,)Hb,Wb,Cb] -> G [Cs,Cb,taps] (see include/rcmvs.h).  `out`: a contiguous fp32 buffer of
    Cs·Cb·taps elements the kernel ADDS into (its partial sums are atomics on G either way)."""
    if dim3d:
        B, Ds, Hs, Ws, Cs = S.shape
        _, Db, Hb, Wb, Cb = Bg.shape
    else:
        B, Hs, Ws, Cs = S.shape
        _, Hb, Wb, Cb = Bg.shape
        Ds = Db = 1
    taps = k ** 3 if dim3d else k * k
    if out is not None:
        if out.numel() != Cs * Cb * taps or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("rcmvs: conv_wgrad out= must be %d contiguous fp32 elements" % (Cs * Cb * taps))
        G = out
    else:
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
    if Cin not in (1, 8, 16, 32, 64):                # e.g. 48: run with the output channels padded to 64, keep the first Cin
        cpad = 64 if Cin < 64 else (Cin + 63) // 64 * 64
        wide = torch.zeros((27, Cout, cpad), device=w.device, dtype=torch.float32)
        wide[:, :, :Cin] = packed
        return ops.conv3d(dy, wide, None, None, stride=1, relu=False)[..., :Cin].contiguous()
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
        inplace = lambda p: GRAD_IN_PLACE and p is not None and p.grad is not None and p.grad.is_contiguous() and \
            p.grad.dtype == torch.float32
        if ctx.has_bn:
            xc, w, y, mean, invstd, g, b = ctx.saved_tensors
            gp, bp = cfg.get("gamma_param"), cfg.get("beta_param")
            into = (gp.grad, bp.grad) if inplace(gp) and inplace(bp) else None
            dy2, dgamma, dbeta = bn_backward(dz.view(-1, C), y.view(-1, C), mean, invstd, g, b, relu, not cfg["batch_stats"],
                                             grad_into=into)
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
        wp = cfg.get("weight_param")
        # (2-D with zero-padded input channels: G is wider than the weight, autograd takes the slice as before)
        out = wp.grad if inplace(wp) and (dim3d or xc.shape[-1] == w.shape[1]) and wp.grad.numel() == w.numel() else None
        if dim3d and transposed:
            G = conv_wgrad(xc, dy, 3, 2, True, out=out)                      # [Cin_t, Cout_t, 27]
        elif dim3d:
            G = conv_wgrad(dy, xc, 3, stride, True, out=out)                 # [Cout, Cin, 27]
        else:
            G = conv_wgrad(dy, xc, k, stride, False, out=out)
            if out is None:
                G = G[:, :w.shape[1]]                                        # [Cout, Cin(_pad -> Cin), k*k]
        if out is not None:
            dw = None
        else:
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
        cfg["gamma_param"], cfg["beta_param"] = bn.weight, bn.bias
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    weight = conv.weight
    cfg["weight_param"] = conv.weight
    if dim3d and not transposed and x_cl.shape[-1] > weight.shape[1]:
        # zero-padded input channels (Neural_Volume_Net: 41 -> 48): pad the weight alike; autograd slices dW back
        weight = torch.nn.functional.pad(weight, (0, 0, 0, 0, 0, 0, 0, x_cl.shape[-1] - weight.shape[1]))
    return _ConvBNAct.apply(x_cl, weight, conv.bias, gamma, beta, skip, cfg)


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


# ---- rendering-consistency branch (training) ------------------------------------------------------------------------------
@_on_device
def sgemm(A, B, bias, M, N, K, trans_a, trans_b, splits=1, out=None):
    C = out if out is not None else (torch.zeros if splits > 1 else torch.empty)((M, N), device=A.device, dtype=torch.float32)
    _ffi.call("rcmvs_sgemm", _chk(A, "A"), _chk(B, "B"), _opt(bias, "bias"), _chk(C, "C"), M, N, K, A.stride(0), B.stride(0),
              C.stride(0), 1 if trans_a else 0, 1 if trans_b else 0, splits, _stream())
    return C


class _Linear(torch.autograd.Function):
    """y = x W^T + b  (nn.Linear), x [M,K] row-major."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xc, w = _f32c(x), _f32c(weight)
        ctx.save_for_backward(xc, w)
        ctx.has_bias = bias is not None
        M, K = xc.shape
        return sgemm(xc, w, _f32c(bias) if bias is not None else None, M, w.shape[0], K, False, True)

    @staticmethod
    def backward(ctx, g):
        xc, w = ctx.saved_tensors
        g = g.contiguous()
        M, K = xc.shape
        N = w.shape[0]
        dx = sgemm(g, w, None, M, K, N, False, False) if ctx.needs_input_grad[0] else None
        dw = sgemm(g, xc, None, N, K, M, True, False, splits=max(1, min(512, M // 512)))
        db = channel_sums(g) if ctx.has_bias and N % 4 == 0 else (g.sum(0) if ctx.has_bias else None)
        return dx, dw, db


def linear(x, lin):
    return _Linear.apply(x, lin.weight, lin.bias)


class _Act(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, a, b, mode):
        a = _f32c(a)
        b = _f32c(b) if b is not None else None
        h = torch.empty_like(a)
        _ffi.call("rcmvs_act_fwd", _chk(a, "a"), _opt(b, "b"), _chk(h, "h"), a.numel(), mode, _stream())
        ctx.mode = mode
        ctx.save_for_backward(a, b, h) if mode == 0 else ctx.save_for_backward(h)
        return h

    @staticmethod
    @_on_device
    def backward(ctx, g):
        g = g.contiguous()
        if ctx.mode == 0:
            a, b, h = ctx.saved_tensors
            da, db = torch.empty_like(a), torch.empty_like(a)
            _ffi.call("rcmvs_act_bwd", _chk(g, "g"), _chk(a, "a"), _chk(b, "b"), _chk(h, "h"), _chk(da, "da"), _chk(db, "db"),
                      a.numel(), 0, _stream())
            return da, db, None
        (h,) = ctx.saved_tensors
        da = torch.empty_like(h)
        _ffi.call("rcmvs_act_bwd", _chk(g, "g"), None, None, _chk(h, "h"), _chk(da, "da"), None, h.numel(), ctx.mode, _stream())
        return da, None, None


def mul_relu(a, b):
    return _Act.apply(a, b, 0)


def relu(a):
    return _Act.apply(a, None, 1)


def sigmoid(a):
    return _Act.apply(a, None, 2)


@_on_device
def embed(x, ld=63):
    """x [P,3] (no gradient) -> [P,ld] with the 63-d positional embedding in the first columns."""
    x = _f32c(x)
    out = torch.empty((x.shape[0], ld), device=x.device, dtype=torch.float32)
    _ffi.call("rcmvs_embed", _chk(x, "x"), _chk(out, "out"), x.shape[0], ld, _stream())
    return out


class _PointFeatures(torch.autograd.Function):
    """[N,S,20] per-point features (csrc/render.cu point_features) with the gradient w.r.t. the encoded volume."""

    @staticmethod
    def forward(ctx, vol_cl, ndc, pts, imgs3, w2cs3, Ks3):
        v = _f32c(vol_cl)
        ndc = _f32c(ndc)
        ctx.save_for_backward(ndc)
        ctx.vshape = v.shape
        return ops.point_features(v, ndc, _f32c(pts), _f32c(imgs3), _f32c(w2cs3), _f32c(Ks3))

    @staticmethod
    @_on_device
    def backward(ctx, g):
        (ndc,) = ctx.saved_tensors
        g = g.contiguous()
        D, h, w, _ = ctx.vshape
        gv = torch.zeros(ctx.vshape, device=g.device, dtype=torch.float32)
        P = ndc.shape[0] * ndc.shape[1]
        _ffi.call("rcmvs_point_features_bwd", _chk(ndc, "ndc"), _chk(g, "grad_feat"), _chk(gv, "grad_vol"), P, g.shape[-1], D, h, w,
                  _stream())
        return gv, None, None, None, None, None


def point_features(vol_cl, ndc, pts, imgs3, w2cs3, Ks3):
    return _PointFeatures.apply(vol_cl, ndc, pts, imgs3, w2cs3, Ks3)


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z):
        r, zz = _f32c(raw), _f32c(z)
        rgb, depth, weights, alpha = ops.composite(r, zz)
        ctx.save_for_backward(r, zz)
        ctx.mark_non_differentiable(alpha)
        return rgb, depth, weights, alpha

    @staticmethod
    @_on_device
    def backward(ctx, g_rgb, g_depth, g_w, _g_alpha):
        r, zz = ctx.saved_tensors
        N, S = zz.shape
        gr = torch.empty_like(r)
        c = lambda t: None if t is None else _chk(t.contiguous(), "grad")
        _ffi.call("rcmvs_composite_bwd", _chk(r, "raw"), _chk(zz, "z"), c(g_rgb), c(g_depth), c(g_w), _chk(gr, "grad_raw"), N, S,
                  _stream())
        return gr, None


def composite(raw, z):
    return _Composite.apply(raw, z)


class _ResizeDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, D1, Cpad):
        x = _f32c(x_cl)
        ctx.shape, ctx.D1, ctx.Cpad = x.shape, D1, Cpad
        return ops.resize_depth_cl(x, D1, Cpad)

    @staticmethod
    @_on_device
    def backward(ctx, g):
        B, D0, h, w, C = ctx.shape
        g = g.contiguous()
        gx = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        _ffi.call("rcmvs_resize_depth_cl_bwd", _chk(g, "grad_y"), _chk(gx, "grad_x"), B, D0, ctx.D1, h * w, C, ctx.Cpad, _stream())
        return gx, None, None


def resize_depth_cl(x_cl, D1, Cpad):
    return _ResizeDepth.apply(x_cl, D1, Cpad)
