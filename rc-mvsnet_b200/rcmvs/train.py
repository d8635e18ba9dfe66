"""One optimiser step of the reference's unsupervised training loop on own kernels (BASELINE configs[3]).

The reference step (train_rcmvsnet.py:130-259) is: `train_sample` (CascadeMVSNet forward + UnsupLossMultiStage, :330-376)
-> `train_sample_aug` (second forward on masked images + AugLossMultiStage against the detached first depth, :397-445)
-> `train_render_net` (Rendering_Consistency_Net + image / depth losses, :279-325) -> ONE backward (:311) -> Adam step
(:312), with DistributedDataParallel averaging the 1.36 M gradients over the ranks (one reference view per GPU, :567-578) and
SyncBatchNorm sharing batch statistics (:524-525).

`TrainStep` keeps that structure:
  * parameters and gradients live in two FLAT fp32 buffers (the modules' .data / .grad are views), so the data-parallel
    exchange is ONE NCCL all-reduce of 5.4 MB over NVLink and the optimiser is ONE fused Adam launch (csrc/train.cu);
  * forward and backward run on librcmvs kernels through rcmvs.autograd (RCMVS_STRICT=1 raises on any library fallback);
  * SyncBatchNorm semantics = `rcmvs.autograd.SYNC_BN` (2·C doubles per BatchNorm layer over NCCL);
  * the weight-gradient and BatchNorm-backward kernels add into the flat gradient buffer directly
    (`rcmvs.autograd.GRAD_IN_PLACE`): no zero-filled temporary and no AccumulateGrad add per parameter use.
The rendering-consistency branch joins the step when `nerf` is given (its differentiable path is selected by the module).

`graph=True` (the default without the render branch) captures "zero the gradients -> both forwards -> losses ->
backward" — about 2000 kernel launches of 5-400 us — into ONE CUDA graph after three eager steps, and replays it per step:
the sample, the masked image and its filter mask are copied into the graph's static inputs first (the mask rectangle is
drawn on the host per step exactly as before), the gradient all-reduce and the Adam launch (whose step count is a host
scalar) follow the replay eagerly.  With SyncBatchNorm the ~300 NCCL all-reduces of 2·C doubles are captured too (measured on
2 GPUs: 51.1 ms per step replayed, 68.4 ms eager); `close()` drops the graph before the process group goes away.
"""
import os

import torch
import torch.distributed as dist

from . import autograd as A


class FlatParams:
    """Re-homes the parameters of `modules` into one flat buffer (and their gradients into another)."""

    def __init__(self, modules):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty((n,), device=dev, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += k
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:                       # autograd may have replaced .grad: re-attach the views
            k = p.numel()
            view = self.grad[off:off + k].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            off += k

    def all_reduce(self):
        """Sum over ranks (the 1/world factor is folded into the Adam kernel's grad_scale)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad)
            return dist.get_world_size()
        return 1

    def adam(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world=1):
        self.step_count += 1
        A.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, lr, betas, eps, weight_decay,
                    grad_scale=1.0 / world)


class TrainStep:
    GRAPH_WARMUP = 3                                 # eager steps before the capture (packing, kernel attributes, caches)

    def __init__(self, model, nerf=None, lr=1e-4, w_aug=0.01, dlossw=(0.5, 1.0, 2.0), sync_bn=None, graph=None,
                 grad_in_place=True):
        from losses import AugLossMultiStage, UnsupLossMultiStage, random_image_mask
        self.model, self.nerf = model, nerf
        self.unsup, self.aug, self.mask_fn = UnsupLossMultiStage(), AugLossMultiStage(), random_image_mask
        self.lr, self.w_aug, self.dlossw = lr, w_aug, list(dlossw)
        self.grad_in_place = bool(grad_in_place)    # conv / BatchNorm parameter gradients go straight into the flat buffer
        self.flat = FlatParams([model] + ([nerf] if nerf is not None else []))
        if sync_bn is None:
            sync_bn = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        A.SYNC_BN = bool(sync_bn)
        if os.environ.get("RCMVS_TRAIN_GRAPH") == "0":
            graph = False
        if graph is None:
            graph = nerf is None        # (SyncBatchNorm's all-reduces of the statistics are captured with the kernels)
        self.use_graph, self._graph, self._static, self._eager_steps, self._stream = bool(graph), None, None, 0, None

    def __call__(self, sample):
        """sample: imgs [B,V,3,H,W] (normalised), center_imgs (the loss's images), imgs_aug, proj_matrices {stageK},
        depth_values [B,Dv] — the keys the reference's loader produces (datasets/dtu_train.py:344-364)."""
        self.model.train()
        imgs_aug = sample["imgs_aug"].clone()
        ref, fmask = self.mask_fn(imgs_aug[:, 0], (imgs_aug.shape[3] // 3, imgs_aug.shape[4] // 3))
        imgs_aug[:, 0] = ref
        if self.use_graph and self._eager_steps >= self.GRAPH_WARMUP:
            loss, scal = self._replay(sample, imgs_aug, fmask)
        elif self.use_graph:
            # the eager steps before the capture run on the capture's own (non-default) stream: autograd's AccumulateGrad
            # nodes remember the stream they were made on, and the legacy default stream cannot join a capture
            cur = torch.cuda.current_stream(imgs_aug.device)
            if self._stream is None:
                self._stream = torch.cuda.Stream(imgs_aug.device)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                loss, scal = self._forward_backward(sample, imgs_aug, fmask)
            cur.wait_stream(self._stream)
            self._eager_steps += 1
        else:
            loss, scal = self._forward_backward(sample, imgs_aug, fmask)
        world = self.flat.all_reduce()
        self.flat.adam(self.lr, world=world)
        return loss, scal

    def _forward_backward(self, sample, imgs_aug, fmask):
        """Zero the gradients, both forwards, the losses and ONE backward — device work only (no host synchronisation), so
        it can run eagerly or inside a CUDA-graph capture."""
        m = self.model
        self.flat.zero_grad()
        prev, A.GRAD_IN_PLACE = A.GRAD_IN_PLACE, self.grad_in_place
        A.stats_arena_begin(imgs_aug.device)
        try:
            return self._forward_backward_body(m, sample, imgs_aug, fmask)
        finally:
            A.GRAD_IN_PLACE = prev
            A.stats_arena_end()

    def _forward_backward_body(self, m, sample, imgs_aug, fmask):
        out, vol = m(sample["imgs"], sample["proj_matrices"], sample["depth_values"])
        loss_base, scal = self.unsup(out, sample["center_imgs"], sample["proj_matrices"], dlossw=self.dlossw)
        pseudo = out["depth"].detach()
        out_aug, _ = m(imgs_aug, sample["proj_matrices"], sample["depth_values"])
        loss_aug, _ = self.aug(out_aug, pseudo, None, fmask, dlossw=self.dlossw)
        loss = loss_base + self.w_aug * loss_aug
        if self.nerf is not None:
            loss = loss + self._render_losses(vol, pseudo, sample)
        loss.backward()
        return loss.detach(), {"loss_base": loss_base.detach(), "loss_aug": loss_aug.detach(), **{k: v.detach() for k, v in scal.items()}}

    def close(self):
        """Release the captured graph (it holds NCCL work when SyncBatchNorm is on: drop it BEFORE the process group)."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph.reset()
            self._graph, self._out, self._static = None, None, None
            self._eager_steps = 0

    _KEYS = ("imgs", "center_imgs", "depth_values")

    def _replay(self, sample, imgs_aug, fmask):
        if fmask is None:                                # (filter = whole image: the reference passes no mask)
            fmask = torch.ones_like(imgs_aug[:, 0])
        if self._graph is None:
            dev = imgs_aug.device
            st = {k: sample[k].clone() for k in self._KEYS}
            st["proj_matrices"] = {k: v.clone() for k, v in sample["proj_matrices"].items()}
            self._static, self._s_aug, self._s_mask = st, imgs_aug.clone(), fmask.clone()
            own = self._stream or torch.cuda.Stream(dev)
            own.wait_stream(torch.cuda.current_stream(dev))
            own.synchronize()
            from . import _ffi
            n0 = _ffi.launches
            self._graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self._graph, stream=own):
                    self._out = self._forward_backward(st, self._s_aug, self._s_mask)
            except Exception as e:                       # e.g. a collective that cannot be captured on this NCCL build
                import warnings
                warnings.warn("rcmvs.train: CUDA-graph capture of the training step failed (%s: %s); the steps stay eager "
                              "(same kernels, launched one by one)" % (type(e).__name__, e))
                self._graph, self._static, self.use_graph = None, None, False
                _ffi.launches = n0
                return self._forward_backward(sample, imgs_aug, fmask)
            self.graph_kernels = _ffi.launches - n0
            _ffi.launches = n0                           # (captured, not executed: the replay below counts them)
            torch.cuda.current_stream(dev).wait_stream(own)
            self.flat.zero_grad()                        # (re-attach .grad views should the capture have replaced any)
        st = self._static
        for k in self._KEYS:
            st[k].copy_(sample[k], non_blocking=True)
        for k, v in sample["proj_matrices"].items():
            st["proj_matrices"][k].copy_(v, non_blocking=True)
        self._s_aug.copy_(imgs_aug, non_blocking=True)
        self._s_mask.copy_(fmask, non_blocking=True)
        self._graph.replay()
        from . import _ffi
        _ffi.launches += self.graph_kernels
        return self._out

    def _render_losses(self, vol, pseudo, sample):
        from losses import SL1Loss, masked_loss
        self.nerf.train()
        rgb, _, _, depth_pred, _, _, rays_depth, target = self.nerf(vol, pseudo, dict(sample["render"]))
        img_loss = masked_loss(rgb.reshape(1, 1, -1), target.reshape(1, 1, -1), None, mask_kind=0, mode=1)
        return img_loss + SL1Loss()(depth_pred, rays_depth, rays_depth > 0)


def synthetic_sample(dev, views=4, H=512, W=640, seed=0):
    """A training sample with the reference loader's keys, from rcmvs.synth."""
    import numpy as np
    from . import synth
    imgs, proj, dv = synth.mvs_inputs(views, H, W, kind="textured", seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    center = (0.5 + 0.2 * imgs).clip(0, 1).astype(np.float32)
    return {"imgs": t(imgs), "center_imgs": t(center), "imgs_aug": t(imgs * 0.9 + 0.05), "depth_values": t(dv),
            "proj_matrices": {k: t(v) for k, v in proj.items()}}


def bench_train_step(dev, steps=5, views=4, H=512, W=640, ndepths=(48, 32, 8), graph=None):
    """ms per optimiser step (2 forwards + losses + backward + gradient all-reduce + fused Adam) of the MVS part of the
    reference step at the default training shape (4 views 512x640, D=48/32/8, batch 1 per GPU), CUDA-event timed."""
    import os
    import numpy as np
    import models
    from . import weights
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    model = models.CascadeMVSNet(ndepths=list(ndepths))
    weights.load_model(model, os.path.join(root, "tests", "golden", "cas_model_000014.npz"))
    model.to(dev)
    step = TrainStep(model, graph=graph)
    rank = dist.get_rank() if dist.is_initialized() else 0
    sample = synthetic_sample(dev, views, H, W, seed=rank)
    np.random.seed(0)
    for _ in range(TrainStep.GRAPH_WARMUP + 2):      # (eager steps, the capture, one replay)
        loss, _ = step(sample)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist.is_initialized():
        dist.barrier()
    e0.record()
    for _ in range(steps):
        loss, _ = step(sample)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist.is_initialized():
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    world = dist.get_world_size() if dist.is_initialized() else 1
    graphed = step._graph is not None
    step.close()
    return {"workload": "configs[3] (MVS part): CascadeMVSNet.train() 2 forwards (%d views %dx%d, D=%s) + UnsupLossMultiStage + "
                        "AugLossMultiStage + backward + flat-gradient all-reduce (%.2f MB) + fused Adam, batch 1 per GPU"
                        % (views, H, W, "/".join(map(str, ndepths)), step.flat.grad.numel() * 4 / 1e6),
            "ms_per_step": round(ms, 3), "value": round(world * 1e3 / ms, 2), "unit": "ref-views/s (training)",
            "n_gpus": world, "sync_bn": bool(A.SYNC_BN),
            "launch": "cuda-graph replay of forward+backward, eager all-reduce + Adam" if graphed else "eager",
            "loss": float(loss), "finite": bool(torch.isfinite(loss))}
