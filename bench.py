#!/usr/bin/env python
"""bench.py — ref-views/s of the RC-MVSNet depth-inference hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's algorithm (torch-CPU port of the oracle) on the host cores

One step = one CascadeMVSNet_eval forward of one reference view (BASELINE configs[1]: 5 views, 512x640, cascade
D=192/64/8, synthetic DTU-shaped cameras, the reference's shipped weights as a numpy fixture).  Multi-GPU = independent
reference views per rank (weak scaling, no data-path collective).

Timing: after >= 1 s of spin-up the K-step loop is timed REPS (9) times, each repetition bracketed by barrier +
synchronize and timed with CUDA events on the launching stream, max over ranks per repetition; the reported
`ms_per_step` / `value` are the MEDIAN repetition (one cold repetition cannot move the number; all repetitions are
listed under `timing`).  Two distinct input sets alternate inside the loop.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; additionally:
  roofline           the fused plane-sweep + variance kernel of stage 1 (the kernel north_star names; HBM-bound by bytes)
  roofline_dominant  the largest kernel of the step by time share: conv0 of stage 1 on tcgen05 (tensor-bound)
  roofline_all       every hot kernel of the step + the standalone homo_warp of BASELINE configs[0] (homo_warp_cfg1)
  cpu_baseline       the oracle's torch-CPU port (F.grid_sample variant = what the reference executes) on this box's
                     host cores, FULL configuration, bounded number of passes (N=1 only)
  stock_torch_gpu    the same port on THIS GPU (cuDNN conv3d + ATen grid_sampler_2d, cudnn.benchmark, TF32 convs as
                     torch defaults): the kernel-for-kernel bar BASELINE.md §5.4 names
  e2e                same metric through the public API with pinned HOST buffers (rcmvs.pipeline.HostPipeline: every
                     step copies its inputs host->device and its depth/confidence device->host inside the timed region)
  single_stream      the same loops with one reference view at a time (per-view latency); the headline loops keep two
                     reference views in flight per GPU on alternating CUDA streams (--in-flight)
  render_branch / train_step / tanks / batched   configs[2] / configs[3] / configs[4] timings, for reference
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rc-mvsnet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("RCMVS_STRICT", "1")          # a torch-library fallback inside the product path raises

import numpy as np  # noqa: E402
import torch  # noqa: E402

GOLDEN_WEIGHTS = os.path.join(ROOT, "tests", "golden", "cas_model_000014.npz")
V, H, W = 5, 512, 640
NDEPTHS = (192, 64, 8)
METRIC = "ref-views/sec (5-view, 512x640, D=192/64/8 cascade)"
REPS = 9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        hot = [c for c, pw in zip(sm, power) if pw > 0.5 * max(power)] if power else sm
        return {"sm_mhz": float(np.median(hot)) if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """Threads for the CPU arms: the affinity/cgroup limit, then whichever candidate count runs a conv3d fastest
    (on big hosts torch's CPU ops get slower past the physical/NUMA-local core count)."""
    import torch.nn.functional as F
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            avail = max(1, min(avail, int(int(q) / int(per))))
    except Exception:
        pass
    x = torch.randn(1, 32, 24, 64, 80)
    w = torch.randn(8, 32, 3, 3, 3)
    best, best_t = 1, float("inf")
    for n in sorted({avail, 96, 64, 48, 32, 16, 8}):
        if n > avail:
            continue
        torch.set_num_threads(n)
        F.conv3d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv3d(x, w, padding=1)
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = n, t
    torch.set_num_threads(best)
    return best, avail


def ncu_traffic(key):
    """dram bytes per launch from the committed `ncu --set full` capture of the CURRENT kernels (profiles/r02_traffic.json
    is regenerated with the profile whenever the kernel changes), or None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(path))[key]["traffic_bytes"]
    except Exception:
        return None


def algorithmic_bytes_costvolume(Vv, C, D, h, w, per_pixel):
    """4·(V·C·h·w + D·h·w + C·D·h·w) — features once, hypotheses once, variance written once."""
    hyp = D * h * w if per_pixel else D
    return 4 * (Vv * C * h * w + hyp + C * D * h * w)


def conv_flops(cin, cout, vox):
    return 2 * 27 * cin * cout * vox


def time_kernel(fn, iters, stream):
    """CUDA-event time of `fn` alone on `stream` (ms per launch), after one warm-up call."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def kernel_rooflines(model, dev, inputs, pk):
    """Per-kernel timings on the tensors of a real forward (stage shapes of the BASELINE config)."""
    from rcmvs import ops
    from models.modules import as_channels_last_views
    imgs, proj, dv = inputs
    stream = torch.cuda.current_stream()
    res = {}
    with torch.no_grad():
        out = model.feature(imgs.reshape(V, 3, H, W))
        feats = {k: as_channels_last_views(f).view(1, V, *f.shape[2:], f.shape[1]) for k, f in out.items()}
        depth = None
        interval = (float(dv[0, -1]) - float(dv[0, 0])) / dv.shape[1]
        for s, D in enumerate(NDEPTHS):
            key = "stage%d" % (s + 1)
            scale = (4, 2, 1)[s]
            h, w = H // scale, W // scale
            C = feats[key].shape[-1]
            hyp = (ops.depth_hypotheses_first(dv, D) if depth is None else
                   ops.depth_hypotheses_next(depth, D, model.depth_interals_ratio[s] * interval, H, W, h, w))
            rt = ops.relative_projections(proj[key])
            cr = model.cost_regularization[s]
            # the sweep writes the variance in the operand format its consumer wants (split-bf16 planes for the TMA-fed
            # tensor-core conv0, fp32 otherwise): same bytes either way
            sweep = (lambda: ops.costvolume_variance_split(feats[key], rt, hyp)) if cr.wants_split_input() else \
                (lambda: ops.costvolume_variance(feats[key], rt, hyp))
            ms = time_kernel(sweep, 10, stream)
            nbytes = algorithmic_bytes_costvolume(V, C, D, h, w, hyp.dim() == 4)
            gbs = nbytes / (ms * 1e-3) / 1e9
            res["costvolume_var_" + key] = {"bound": "hbm", "ms": round(ms, 4), "achieved": round(gbs, 1),
                                            "peak": pk["hbm_gbs"], "unit": "GB/s",
                                            "frac": round(gbs / pk["hbm_gbs"], 4), "bytes": nbytes}
            var = sweep()
            P = cr.packed()
            ms0 = time_kernel(lambda: cr._run(P["conv0"], var), 5, stream)
            fl = conv_flops(C, 8, D * h * w)
            # conv0 is the dominant CostRegNet layer; tensor roofline = TF32 dense = half the measured bf16 peak
            tf32_peak = pk["bf16_tflops"] / 2
            res["conv0_" + key] = {"bound": "tensor", "ms": round(ms0, 4), "achieved": round(fl / (ms0 * 1e-3) / 1e12, 2),
                                   "peak": tf32_peak, "unit": "TFLOP/s", "frac": round(fl / (ms0 * 1e-3) / 1e12 / tf32_peak, 4),
                                   "hbm_gbs": round(4 * (C + 8) * D * h * w / (ms0 * 1e-3) / 1e9, 1), "flops": fl}
            msr = time_kernel(lambda: cr.forward_cl(var), 3, stream)
            res["costregnet_" + key] = {"ms": round(msr, 4)}
            cost = cr.forward_cl(var)
            mss = time_kernel(lambda: ops.softargmin_conf(cost, hyp), 10, stream)
            res["softargmin_" + key] = {"ms": round(mss, 4)}
            depth, _ = ops.softargmin_conf(cost, hyp)
            del var, cost
        msf = time_kernel(lambda: model.feature(imgs.reshape(V, 3, H, W)), 5, stream)
        res["featurenet"] = {"ms": round(msf, 4)}
        res["engines"] = {"stage%d" % (i + 1): {k: list(v.values())[0] for k, v in cr.engines().items() if v}
                          for i, cr in enumerate(model.cost_regularization)}
        # BASELINE configs[0] / the second half of the metric: the STANDALONE homo_warping kernel, one source view,
        # C=32, 128x160, D=48, per-pixel hypotheses: 4·(C·h·w + D·h·w + C·D·h·w) = 132.4 MB per call
        Cw, hw_, ww_, Dw = 32, 128, 160, 48
        g = torch.Generator(device="cpu").manual_seed(7)
        src = torch.randn(1, Cw, hw_, ww_, generator=g).to(dev)
        rt1 = ops.relative_projections(proj["stage1"])[:, 0].contiguous()
        dpp = (torch.linspace(425, 933.8, Dw).view(1, Dw, 1, 1) + 6 * torch.rand(1, Dw, hw_, ww_, generator=g)).to(dev)
        msw = time_kernel(lambda: ops.homo_warp_fwd(src, rt1, dpp), 20, stream)
        wb = 4 * (Cw * hw_ * ww_ + Dw * hw_ * ww_ + Cw * Dw * hw_ * ww_)
        res["homo_warp_cfg1"] = {"bound": "hbm", "ms": round(msw, 4), "achieved": round(wb / (msw * 1e-3) / 1e9, 1),
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(wb / (msw * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                                 "bytes": wb, "note": "132 MB per call: output stays L2-resident across back-to-back calls"}
    return res


def render_branch_timing(dev, pk):
    """BASELINE configs[2] (full RC-MVSNet forward + render), reported next to the headline: 4 views (the renderer is
    hard-wired to 3 sources), 512x640, training-class CascadeMVSNet 48/32/8 (produces volume_feature_no_ref) followed by
    Rendering_Consistency_Net (Neural_Volume_Net + 1024 rays x 128 samples), forward only, inputs resident."""
    from rcmvs import synth, weights
    import models
    nerf_w = os.path.join(ROOT, "tests", "golden", "nerf_model_000014.npz")
    cas = models.CascadeMVSNet(ndepths=[48, 32, 8])
    weights.load_model(cas, GOLDEN_WEIGHTS)
    cas.to(dev).eval()
    nerf = models.Rendering_Consistency_Net(synth.RenderArgs())
    nerf.load_state_dict(weights.load_npz_state_dict(nerf_w), strict=True)
    nerf.to(dev).eval()
    imgs, proj, dv = synth.mvs_inputs(4, H, W, kind="textured", seed=11)
    batch = {k: torch.from_numpy(v).to(dev) for k, v in synth.render_batch(H, W, seed=3).items()}
    imgs, proj, dv = torch.from_numpy(imgs).to(dev), {k: torch.from_numpy(v).to(dev) for k, v in proj.items()}, \
        torch.from_numpy(dv).to(dev)
    stream = torch.cuda.current_stream()
    with torch.no_grad():
        out, vol = cas(imgs, proj, dv)
        ms_mvs = time_kernel(lambda: cas(imgs, proj, dv), 3, stream)
        ms_render = time_kernel(lambda: nerf(vol, out["depth"], dict(batch)), 3, stream)
        ms_vol = time_kernel(lambda: nerf.MVSNet.forward_cl(vol), 3, stream)
        res = nerf(vol, out["depth"], dict(batch))
        # the same forward with the MLP's Linear layers on tcgen05 (opt-in: RCMVS_MLP_TC=1, csrc/mlp_tc.cu)
        prev = os.environ.get("RCMVS_MLP_TC")
        os.environ["RCMVS_MLP_TC"] = "1"
        try:
            ms_render_tc = time_kernel(lambda: nerf(vol, out["depth"], dict(batch)), 3, stream)
        finally:
            if prev is None:
                os.environ.pop("RCMVS_MLP_TC", None)
            else:
                os.environ["RCMVS_MLP_TC"] = prev
    ok = bool(torch.isfinite(res[0]).all() and torch.isfinite(res[3]).all())
    # Neural_Volume_Net 63 GFLOP + MLP 251 392 FLOP/pt x 131 072 pts = 33.0 GFLOP per render (forward)
    flops = 63.0e9 + 251392 * 131072
    tf = flops / (ms_render * 1e-3) / 1e12
    return {"workload": "configs[2]: CascadeMVSNet (train class, 4 views 512x640, D=48/32/8, volume_feature_no_ref) + "
                        "Rendering_Consistency_Net (Neural_Volume_Net, 1024 rays x 128 samples), forward",
            "mvs_forward_ms": round(ms_mvs, 3), "render_forward_ms": round(ms_render, 3),
            "neural_volume_net_ms": round(ms_vol, 3), "render_forward_mlp_tc_ms": round(ms_render_tc, 3), "finite": ok,
            "roofline": {"bound": "tensor", "achieved": round(tf, 2), "peak": pk["bf16_tflops"] / 2, "unit": "TFLOP/s",
                         "frac": round(tf / (pk["bf16_tflops"] / 2), 4), "flops": flops,
                         "note": "96 GFLOP of dense fp32 contractions per render against the TF32 tensor peak"}}


def tanks_timing(model_cls, dev):
    """BASELINE configs[4] shape on one GPU: 7 views 1056x1920, cascade 48/32/8 (eval_rcmvsnet_tanks.py defaults)."""
    from rcmvs import synth, weights
    m = model_cls(ndepths=[48, 32, 8])
    weights.load_model(m, GOLDEN_WEIGHTS)
    m.to(dev).eval()
    imgs, proj, dv = synth.mvs_inputs(7, 1056, 1920, kind="noise", seed=1)
    a = (torch.from_numpy(imgs).to(dev), {k: torch.from_numpy(v).to(dev) for k, v in proj.items()}, torch.from_numpy(dv).to(dev))
    with torch.no_grad():
        # median of 7 individually timed forwards: a forward allocates ~3 GB of volumes, and an allocator round trip
        # (cudaFree + cudaMalloc after the other sections of this process) inside a 5-iteration mean moved it by 60 %
        ms = float(np.median([time_kernel(lambda: m(*a), 1, torch.cuda.current_stream()) for _ in range(7)]))
    return {"workload": "configs[4] shape: 7 views 1056x1920, D=48/32/8, one GPU", "ms_per_step": round(ms, 3),
            "value": round(1e3 / ms, 2), "unit": "ref-views/s"}


def stock_torch_gpu(dev, inputs_np):
    """The reference's op sequence on stock torch kernels on THIS GPU (oracle/ref_torch.py with F.grid_sample, cuDNN
    conv2d/conv3d/conv_transpose3d, cudnn.benchmark=True as eval_rcmvsnet_dtu.py:24, TF32 convolutions = torch default)."""
    from oracle import ref_torch as T
    imgs, proj, dv = inputs_np
    P = {k: torch.from_numpy(v).to(dev) for k, v in np.load(GOLDEN_WEIGHTS).items()}
    pr = {k: torch.from_numpy(v[0]).to(dev) for k, v in proj.items()}
    im, d = torch.from_numpy(imgs[0]).to(dev), torch.from_numpy(dv[0]).to(dev)
    run = lambda: T.cascade_forward_eval(im, pr, d, P, ndepths=NDEPTHS, device=dev, grid_sample=True, as_numpy=False)
    for _ in range(3):
        run()
    ms = time_kernel(run, 5, torch.cuda.current_stream())
    return {"value": round(1e3 / ms, 2), "unit": "ref-views/s", "ms_per_step": round(ms, 3),
            "what": "oracle/ref_torch.py port on cuda: cuDNN %s convs (allow_tf32=%s, benchmark=True) + ATen grid_sampler_2d, "
                    "eager, same config" % (torch.backends.cudnn.version(), torch.backends.cudnn.allow_tf32)}


def make_inputs(seed):
    from rcmvs import synth
    return synth.mvs_inputs(V, H, W, kind="textured", seed=seed)


def cpu_port_step_fn():
    """One FULL-configuration forward of the reference algorithm on the host (torch-CPU port, F.grid_sample variant)."""
    from oracle import ref_torch as T
    from rcmvs import synth
    P = {k: torch.from_numpy(v) for k, v in np.load(GOLDEN_WEIGHTS).items()}
    imgs, proj, dv = synth.mvs_inputs(V, H, W, kind="textured", seed=0)
    pr = {k: v[0] for k, v in proj.items()}
    return lambda: T.cascade_forward_eval(imgs[0], pr, dv[0], P, ndepths=NDEPTHS, grid_sample=True)


def run_reference_arm(args, rank, world):
    """The reference's CPU algorithm (torch-CPU port of the oracle, all useful host threads), full configuration:
    one step = one whole 5-view 512x640 D=192/64/8 forward."""
    if rank != 0:
        return
    cores, avail = host_threads()
    step = cpu_port_step_fn()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = ("full configuration, one whole ref-view per step (%d views %dx%d, D=%s), %d of %d host threads (fastest "
              "count), torch-CPU port of the oracle incl. F.grid_sample as the reference calls it"
              % (V, H, W, "/".join(map(str, NDEPTHS)), cores, avail))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "ref-views/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CascadeMVSNet_eval forward, %d views %dx%d D=%s" % (V, H, W, "/".join(map(str, NDEPTHS))),
                       "class": "CPU port of the reference algorithm (the reference is Python and cannot travel to the box)"},
            "cpu_baseline": {"value": value, "unit": "ref-views/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "ref-views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rcmvs", choices=["rcmvs", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="reference views per step per GPU")
    ap.add_argument("--reps", type=int, default=REPS, help="timed repetitions of the K-step loop (median reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the configs[2]/[3]/[4] side timings")
    ap.add_argument("--no-stock-torch", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="eager launch sequence instead of CUDA-graph replay")
    ap.add_argument("--in-flight", type=int, default=2, dest="in_flight",
                    help="reference views in flight per GPU (alternating CUDA streams) in the headline loops; the "
                         "one-view-at-a-time numbers are reported under `single_stream`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from rcmvs import _ffi, weights
    rc = _ffi.lib().rcmvs_check_device()
    assert rc == 0, _ffi.lib().rcmvs_last_error()
    from models import CascadeMVSNet_eval
    torch.backends.cudnn.benchmark = True          # as the reference's eval script (eval_rcmvsnet_dtu.py:24)
    model = CascadeMVSNet_eval(ndepths=list(NDEPTHS))
    weights.load_model(model, GOLDEN_WEIGHTS)
    model.to(dev).eval()

    B = args.batch
    NSETS = 2                                       # distinct input sets alternating inside the timed loops
    sets_h, sets_d = [], []
    for s in range(NSETS):
        host = [make_inputs(seed=rank * 131 + s * 17 + b) for b in range(B)]
        imgs_h = torch.from_numpy(np.concatenate([h[0] for h in host])).pin_memory()
        proj_h = {k: torch.from_numpy(np.concatenate([h[1][k] for h in host])).pin_memory() for k in host[0][1]}
        dv_h = torch.from_numpy(np.concatenate([h[2] for h in host])).pin_memory()
        sets_h.append((imgs_h, proj_h, dv_h))
        sets_d.append((imgs_h.to(dev), {k: v.to(dev) for k, v in proj_h.items()}, dv_h.to(dev)))
        if s == 0:
            first_np = host[0]
    imgs_h, proj_h, dv_h = sets_h[0]
    depth_out_h = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    conf_out_h = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    h2d = imgs_h.numel() * 4 + sum(v.numel() for v in proj_h.values()) * 4 + dv_h.numel() * 4
    d2h = 2 * B * H * W * 4
    counter = [0]

    # Headline loops keep `in_flight` reference views in flight per GPU on alternating CUDA streams (a serving process
    # does the same: the 1/4- and 1/8-resolution layers of one view leave SMs idle that the next view's kernels fill).
    nfl = max(1, args.in_flight)
    lanes = [torch.cuda.Stream(dev) for _ in range(nfl)] if nfl > 1 else None

    # CUDA graphs (rcmvs.graph): the ~60 launches of one forward are captured once per lane and replayed — the host then
    # issues one launch per reference view; the inputs of every step are copied into the graph's static buffers first
    # (device-to-device, inside the timed region).  --no-graphs times the eager launch sequence instead.
    use_graphs = not args.no_graphs
    graphs = {}

    def lane_graph(i):
        if i not in graphs:
            from rcmvs.graph import GraphedForward
            graphs[i] = GraphedForward(model, *sets_d[0])
        return graphs[i]

    def step_resident():
        counter[0] += 1
        lane = counter[0] % nfl
        with torch.no_grad():
            if lanes is None:
                if use_graphs:
                    return lane_graph(0)(*sets_d[counter[0] % NSETS])
                return model(*sets_d[counter[0] % NSETS])
            with torch.cuda.stream(lanes[lane]):
                if use_graphs:
                    return lane_graph(lane)(*sets_d[counter[0] % NSETS])
                return model(*sets_d[counter[0] % NSETS])

    def step_single():
        counter[0] += 1
        with torch.no_grad():
            if use_graphs:
                return lane_graph(-1)(*sets_d[counter[0] % NSETS])
            return model(*sets_d[counter[0] % NSETS])

    from rcmvs.pipeline import HostPipeline
    pipe = HostPipeline(model, dev, compute_streams=nfl, graphs=use_graphs)

    def step_e2e():
        # host buffers in, host buffers out, every step: the H2D copy of this step's inputs and the D2H copy of its
        # depth/confidence are inside the timed region; they run on their own streams and overlap the neighbouring
        # steps' kernels (rcmvs/pipeline.py)
        counter[0] += 1
        ih, ph, dh = sets_h[counter[0] % NSETS]
        pipe.submit(ih, ph, dh, depth_out_h, conf_out_h)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_once(fn, steps):
        barrier()
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _ffi.launches
        e0.record(stream)
        if fn is step_resident and lanes is not None:
            for ln in lanes:
                ln.wait_event(e0)             # the lanes start after e0 ...
        for _ in range(steps):
            fn()
        if fn is step_e2e:
            pipe.join()                       # the last step's device->host copies end before e1
        if fn is step_resident and lanes is not None:
            for ln in lanes:
                stream.wait_stream(ln)        # ... and e1 is recorded after every lane has drained
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _ffi.launches - n0

    def timed(fn, steps, reps):
        """REPS repetitions of the K-step loop, each max-over-ranks; (median ms, launches of one repetition, all ms)."""
        all_ms, launches = [], 0
        for _ in range(reps):
            ms, launches = timed_once(fn, steps)
            all_ms.append(ms)
        return float(np.median(all_ms)), launches, [round(m, 3) for m in all_ms]

    for _ in range(args.warmup):
        step_resident()
    for _ in range(args.warmup):
        step_e2e()
    # spin-up: >= 1 s of the real workload so clocks and the allocator are in steady state before anything is timed
    barrier()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 1.0:
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
    # ONE reference view at a time first (single stream: the per-view latency), then the headline loops
    single = None
    if nfl > 1:
        for _ in range(args.warmup):
            step_single()
        ms1, _, _ = timed(step_single, args.steps, max(5, args.reps // 2))
        pipe_n, pipe = pipe, HostPipeline(model, dev, compute_streams=1, graphs=use_graphs)
        for _ in range(args.warmup):
            step_e2e()
        ms1e, _, _ = timed(step_e2e, args.steps, max(5, args.reps // 2))
        pipe = pipe_n
        single = {"value": world * B * args.steps / (ms1 * 1e-3), "ms_per_step": ms1 / args.steps,
                  "e2e_value": world * B * args.steps / (ms1e * 1e-3), "e2e_ms_per_step": ms1e / args.steps,
                  "unit": "ref-views/s", "views_in_flight": 1}
        for _ in range(args.warmup):
            step_resident()
            step_e2e()
        barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches, reps_resident = timed(step_resident, args.steps, args.reps)
    ms_e2e, _, reps_e2e = timed(step_e2e, args.steps, args.reps)
    clocks = sampler.stop() if sampler else None

    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    pk = peaks()
    roof_all, cpu_baseline, stock = None, None, None
    side = rank == 0 and world == 1
    if rank == 0 and not args.no_kernel_rooflines:
        d0 = sets_d[0]
        roof_all = kernel_rooflines(model, dev, (d0[0][:1], {k: v[:1] for k, v in d0[1].items()}, d0[2][:1]), pk)
    render, batched, tanks, train = None, None, None, None
    if side and not args.no_render and not args.no_kernel_rooflines:
        render = render_branch_timing(dev, pk)
        tanks = tanks_timing(CascadeMVSNet_eval, dev)
        from rcmvs import train as rtrain
        train = rtrain.bench_train_step(dev, steps=5)
        # same workload with 4 reference views per launch sequence (better fill of the small 1/4- and 1/8-scale layers)
        b4 = [make_inputs(seed=1000 + i) for i in range(4)]
        i4 = torch.from_numpy(np.concatenate([h[0] for h in b4])).to(dev)
        p4 = {k: torch.from_numpy(np.concatenate([h[1][k] for h in b4])).to(dev) for k in b4[0][1]}
        d4 = torch.from_numpy(np.concatenate([h[2] for h in b4])).to(dev)
        with torch.no_grad():
            ms4 = time_kernel(lambda: model(i4, p4, d4), 5, torch.cuda.current_stream())
        batched = {"batch_per_gpu": 4, "value": round(4 / (ms4 * 1e-3), 1), "unit": "ref-views/s", "ms_per_step": round(ms4, 3)}
        del i4, p4, d4
    if side and not args.no_stock_torch:
        stock = stock_torch_gpu(dev, first_np)
    if side and not args.no_cpu_baseline:
        cores, avail = host_threads()
        step = cpu_port_step_fn()
        step()                                                                     # warm-up
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": 1.0 / dt, "unit": "ref-views/s", "cores": cores, "kind": "port",
                        "sample": "FULL configuration (%d views %dx%d, D=%s), %d passes of %.1f s after 1 warm-up, %d of "
                                  "%d host threads (fastest count), torch-CPU port incl. F.grid_sample as the reference "
                                  "calls it" % (V, H, W, "/".join(map(str, NDEPTHS)), reps, dt, cores, avail)}
    if dist is not None and not args.no_render:
        # configs[3] under torchrun: every rank trains on its own reference view, one NCCL all-reduce of the flat gradient
        # buffer + SyncBatchNorm statistics per step (device-timed, max over ranks inside bench_train_step)
        from rcmvs import train as rtrain
        train = rtrain.bench_train_step(dev, steps=5)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    s1 = (roof_all or {}).get("costvolume_var_stage1")
    roofline = None
    if s1:
        roofline = {"kernel": "fused plane-sweep + variance (stage 1, D=192, 128x160, V=5, C=32)", "bound": "hbm",
                    "achieved": s1["achieved"], "peak": s1["peak"], "unit": "GB/s", "frac": s1["frac"],
                    "peak_source": pk["source"] + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                    "algorithmic_bytes": s1["bytes"], "traffic": ncu_traffic("costvolume_var_stage1")}
    # the single largest kernel of the step by time share (profiles/) is tensor-bound: conv0 of stage 1 on the slab
    # tcgen05 kernel.  FLOPs = 2*27*Cin*Cout*voxels of the fp32 convolution (the 3xTF32 mode issues three TF32 products
    # per fp32 product; they are not counted), peak = dense TF32 = half the measured bf16 peak.
    c0 = (roof_all or {}).get("conv0_stage1")
    roofline_dominant = None
    if c0:
        roofline_dominant = {"kernel": "conv3d tcgen05 slab kernel (conv0 of stage 1: 32->8 channels, 192x128x160, %s)"
                                       % os.environ.get("RCMVS_CONV", "bf16x3"),
                             "bound": "tensor", "achieved": c0["achieved"], "peak": c0["peak"], "unit": "TFLOP/s",
                             "frac": c0["frac"], "ms": c0["ms"], "flops": c0["flops"],
                             "traffic": ncu_traffic("conv0_stage1")}
    line = {
        "metric": METRIC, "value": value, "unit": "ref-views/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CascadeMVSNet_eval forward: %d views, %dx%d, D=%s (BASELINE configs[1]); "
                               "reference pretrained weights (numpy fixture), synthetic DTU-shaped cameras"
                               % (V, H, W, "/".join(map(str, NDEPTHS))),
                   "batch_per_gpu": B, "views_in_flight_per_gpu": nfl,
                   "parallelism": "view-sharded replicas x%d" % world,
                   "cache": "per-step working set %.0f MB of volumes >> 126 MB L2 (no explicit flush needed); %d input "
                            "sets alternate" % (4 * (32 * 192 * 128 * 160 + 16 * 64 * 256 * 320 + 8 * 8 * 512 * 640) / 1e6, NSETS),
                   "conv_precision": os.environ.get("RCMVS_CONV", "bf16x3"),
                   "launch": ("CUDA-graph replay: one graph launch per reference view, inputs copied into the graph's static "
                              "buffers inside the timed region") if use_graphs else "eager: one launch per kernel"},
        "timing": {"stat": "median of %d repetitions of the %d-step loop after >= 1 s spin-up" % (args.reps, args.steps),
                   "ms_per_rep_resident": reps_resident, "ms_per_rep_e2e": reps_e2e},
        "e2e": {"value": e2e_value, "unit": "ref-views/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "views_in_flight": nfl},
        "single_stream": single,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_dominant": roofline_dominant,
        "roofline_all": roof_all,
        "cpu_baseline": cpu_baseline,
        "stock_torch_gpu": stock,
        "vs_reference_note": "the --impl reference arm is a CPU port (GPU-vs-CPU, context only); the like-for-like bar is "
                             "stock_torch_gpu (same GPU, library kernels)",
        "render_branch": render,
        "train_step": train,
        "tanks": tanks,
        "batched": batched,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
