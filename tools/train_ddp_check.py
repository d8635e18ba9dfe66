"""torchrun --nproc-per-node N tools/train_ddp_check.py: a few optimiser steps of rcmvs.train.TrainStep on N GPUs (one
reference view per rank, NCCL all-reduce of the flat gradient buffer, SyncBatchNorm statistics), asserting that every rank
holds bit-identical parameters after each step and that the loss is finite.  Evidence for BASELINE configs[3]."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rc-mvsnet_b200"))
os.environ.setdefault("RCMVS_STRICT", "1")
import models
from rcmvs import train, weights, _ffi

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
H, W = (int(os.environ.get("H", 256)), int(os.environ.get("W", 320)))
model = models.CascadeMVSNet(ndepths=[48, 32, 8])
weights.load_model(model, os.path.join(ROOT, "tests", "golden", "cas_model_000014.npz"))
model.to(dev)
use_graph = os.environ.get("GRAPH", "0") == "1"                   # CUDA-graph replay with the SyncBatchNorm all-reduces captured
step = train.TrainStep(model, lr=1e-4, graph=use_graph)
sample = train.synthetic_sample(dev, 4, H, W, seed=rank)           # a different reference view on every rank
nsteps = int(os.environ.get("STEPS", 3 + (train.TrainStep.GRAPH_WARMUP if use_graph else 0)))
for it in range(nsteps):
    n0 = _ffi.launches
    loss, scal = step(sample)
    chk = torch.stack([step.flat.flat.double().sum(), step.flat.flat.double().abs().sum()])
    all_chk = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    same = all(torch.equal(all_chk[0], c) for c in all_chk)
    losses = [torch.empty_like(loss) for _ in range(world)]
    dist.all_gather(losses, loss)
    if rank == 0:
        print("step %d: losses per rank %s, parameters identical on all %d ranks: %s, rcmvs kernel launches this step: %d"
              % (it, [round(float(l), 4) for l in losses], world, same, _ffi.launches - n0), flush=True)
    assert same and torch.isfinite(loss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for it in range(5):
    loss, _ = step(sample)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print("OK: %d ranks, sync_bn=%s, graph=%s, flat gradient buffer %.2f MB, %.2f ms per step (max over ranks, %dx%d)"
          % (world, train.A.SYNC_BN, step._graph is not None, step.flat.grad.numel() * 4 / 1e6, float(ms), H, W))
step.close()
dist.destroy_process_group()
