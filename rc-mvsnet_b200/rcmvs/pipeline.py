"""Host-to-host streaming of reference views through a CascadeMVSNet(_eval) on one GPU.

The reference's eval loop (eval_rcmvsnet_dtu.py:118-160) moves a sample to the GPU, runs the model and pulls depth and
confidence back, one after the other.  `HostPipeline` keeps that contract (host buffers in, host buffers out) but gives
the copies their own CUDA streams and two sets of device input buffers, so the host->device copy of view i+1 and the
device->host copy of view i-1 overlap the kernels of view i.  Inputs must be pinned for the copies to be asynchronous.
"""
import torch

__all__ = ["HostPipeline"]


class HostPipeline:
    def __init__(self, model, device, slots=2, compute_streams=1, graphs=False):
        """`compute_streams` > 1 runs consecutive submits on alternating CUDA streams: two reference views in flight
        fill the SMs during the small 1/4- and 1/8-resolution layers and overlap one view's FFMA-bound FeatureNet with
        the other's tensor-core-bound regularisation.  Results are identical; memory use doubles.
        `graphs`: capture the forward of every slot into a CUDA graph (rcmvs.graph) on first use and replay it — one launch
        per reference view instead of ~60; the slot's device input buffers are the graph's static inputs."""
        slots = max(slots, compute_streams + 1)
        self.model, self.device, self.slots = model, torch.device(device), slots
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.compute = [torch.cuda.Stream(self.device) for _ in range(compute_streams)] if compute_streams > 1 else None
        self._pending = []
        self._dev_in = [None] * slots            # per slot: (imgs, {stage: proj}, depth_values) on the device
        self._consumed = [None] * slots          # event: the forward that read slot s has been enqueued and finished
        self._n = 0
        self._graphs = [None] * slots if graphs else None
        self._copied = [None] * slots           # event: the outputs of slot s have been copied to the host

    def _slot_buffers(self, s, imgs_h, proj_h, dv_h):
        cur = self._dev_in[s]
        if cur is None or cur[0].shape != imgs_h.shape or cur[2].shape != dv_h.shape:
            if self._consumed[s] is not None:
                self._consumed[s].synchronize()     # a forward in flight may still read the buffers being replaced
                self._consumed[s] = None
            mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.device)
            cur = (mk(imgs_h), {k: mk(v) for k, v in proj_h.items()}, mk(dv_h))
            self._dev_in[s] = cur
            if self._graphs is not None:
                self._graphs[s] = None              # captured for the old buffers
        return cur

    def submit(self, imgs_h, proj_h, dv_h, depth_out_h, conf_out_h):
        """Enqueue one batch of reference views; returns immediately.  `depth_out_h` / `conf_out_h` (pinned host
        tensors) are valid after `synchronize()` (or after the returned event completes)."""
        s = self._n % self.slots
        caller = torch.cuda.current_stream(self.device)
        main = self.compute[self._n % len(self.compute)] if self.compute else caller
        self._n += 1
        imgs_d, proj_d, dv_d = self._slot_buffers(s, imgs_h, proj_h, dv_h)
        with torch.cuda.stream(self.h2d):
            if self._consumed[s] is not None:
                self.h2d.wait_event(self._consumed[s])          # the forward two submits ago is done with this slot
            imgs_d.copy_(imgs_h, non_blocking=True)
            for k, v in proj_h.items():
                proj_d[k].copy_(v, non_blocking=True)
            dv_d.copy_(dv_h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.h2d)
        main.wait_event(ready)
        if self._graphs is not None:
            if self._graphs[s] is None:
                from .graph import GraphedForward
                ready.synchronize()                                 # capture needs real inputs for its warm-up runs
                self._graphs[s] = GraphedForward(self.model, imgs_d, proj_d, dv_d, in_place=True)
            if self._copied[s] is not None:
                main.wait_event(self._copied[s])                    # the static outputs of this slot have been read
            with torch.cuda.stream(main):
                out = self._graphs[s].replay()
        else:
            with torch.no_grad(), torch.cuda.stream(main):
                out = self.model(imgs_d, proj_d, dv_d)
            if isinstance(out, tuple):
                out = out[0]
        done = torch.cuda.Event()
        done.record(main)
        self._consumed[s] = done
        depth, conf = out["depth"], out["photometric_confidence"]
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(done)
            depth_out_h.copy_(depth, non_blocking=True)
            conf_out_h.copy_(conf, non_blocking=True)
            if self._graphs is None:
                depth.record_stream(self.d2h)
                conf.record_stream(self.d2h)
            copied = torch.cuda.Event()
            copied.record(self.d2h)
        self._copied[s] = copied
        self._pending.append(copied)
        if len(self._pending) > self.slots:
            self._pending.pop(0)
        return copied

    def join(self):
        """Make the current stream wait for every copy enqueued so far (so an event recorded next covers them)."""
        main = torch.cuda.current_stream(self.device)
        for ev in self._pending:
            main.wait_event(ev)
        self._pending = []

    def synchronize(self):
        self.join()
        torch.cuda.current_stream(self.device).synchronize()
