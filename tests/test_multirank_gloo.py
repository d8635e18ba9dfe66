"""N>1 path on CPU: world_size-2 gloo processes exercise the view sharding, the max-over-ranks timing reduction and
the result gathering that bench.py / a view-sharded evaluation use on NCCL (no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "rc-mvsnet_b200")]
    from rcmvs import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_units = 7
    mine = shard.shard_units(n_units, rank, world)
    # stand-in for the per-view hot path: a deterministic function of the unit index
    local = {i: torch.full((2, 3), float(i * i)) for i in mine}
    ms = 10.0 + 5.0 * rank                                     # rank 1 is the slow one
    ms_max = shard.max_over_ranks(ms)
    total_units = shard.sum_over_ranks(len(mine))
    allres = shard.gather_results(local, n_units, rank, world)
    dist.barrier()
    if rank == 0:
        torch.save({"mine": mine, "ms_max": ms_max, "total": total_units,
                    "vals": [float(t[0, 0]) for t in allres],
                    "tput": shard.throughput(total_units, ms_max)}, out)
    dist.destroy_process_group()


def test_view_sharding_two_ranks_gloo(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["mine"] == [0, 2, 4, 6]
    assert r["ms_max"] == 15.0                                 # max over ranks, never rank 0's own clock
    assert r["total"] == 7.0
    assert r["vals"] == [float(i * i) for i in range(7)]       # every unit exactly once, in order
    assert abs(r["tput"] - 7 / 0.015) < 1e-6


def test_shard_units_partition():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "rc-mvsnet_b200")]
    from rcmvs import shard
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 49):
            parts = [shard.shard_units(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _train_worker(rank, world, port, out):
    """Host logic of the data-parallel training step (rcmvs.train.FlatParams) on CPU/gloo: parameters and gradients are
    views of two flat buffers, ONE all-reduce sums the gradients of all ranks, the 1/world factor goes to the optimiser."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "rc-mvsnet_b200")]
    from rcmvs.train import FlatParams
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    flat = FlatParams([net])
    n = sum(p.numel() for p in net.parameters())
    assert flat.flat.numel() == n and all(p.data_ptr() >= flat.flat.data_ptr() for p in net.parameters())
    flat.zero_grad()
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()
    local = flat.grad.clone()
    assert torch.equal(torch.cat([p.grad.reshape(-1) for p in net.parameters()]), local)    # .grad are views
    world_seen = flat.all_reduce()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(flat.grad, sum(gathered)) and world_seen == world
    flat.zero_grad()
    ok = ok and float(flat.grad.abs().sum()) == 0.0 and all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters())
    if rank == 0:
        torch.save({"ok": bool(ok)}, out)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks_gloo(tmp_path):
    out = str(tmp_path / "t0.pt")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["ok"]


def test_step_workspace_arena_hands_out_zeroed_slices():
    """rcmvs.autograd's per-step arena (the BatchNorm kernels' 2·C-double workspaces in a training step): slices are zero,
    disjoint, come from ONE buffer that is re-zeroed by the next begin, and the plain torch.zeros path takes over when the
    arena is inactive or exhausted."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "rc-mvsnet_b200")]
    from rcmvs import autograd as A
    dev = torch.device("cpu")
    assert A.STATS_ARENA is None
    plain = A._zeros64(16, dev)
    assert plain.dtype == torch.float64 and float(plain.abs().sum()) == 0.0
    A.stats_arena_begin(dev, n=64)
    try:
        a, b = A._zeros64(16, dev), A._zeros64(32, dev)
        assert a.data_ptr() + 16 * 8 == b.data_ptr() and float(a.abs().sum() + b.abs().sum()) == 0.0
        a.fill_(3.0); b.fill_(5.0)
        big = A._zeros64(32, dev)                      # 16 + 32 + 32 > 64: falls back to its own allocation
        assert big.data_ptr() != b.data_ptr() + 32 * 8 and float(big.abs().sum()) == 0.0
        A.stats_arena_begin(dev, n=64)                 # next step: same buffer, zeroed again, cursor at 0
        c = A._zeros64(16, dev)
        assert c.data_ptr() == a.data_ptr() and float(c.abs().sum()) == 0.0
    finally:
        A.stats_arena_end()
    assert A.STATS_ARENA is None
