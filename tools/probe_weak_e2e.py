"""Weak-scaling end-to-end probe: every rank runs 2^20-point kgr_msm_oneshot calls from pinned host memory at the same time (barrier per step):
torchrun --nproc-per-node N tools/probe_weak_e2e.py [growth percent, 0 = automatic]"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, ".")
import kogarashi_b200 as k
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dist.init_process_group("gloo")
k.init([int(os.environ.get("LOCAL_RANK", 0))])
n = 1 << 20
bases = k.Bases.generate(0, n, seed=3 + rank)
pts = torch.from_numpy(bases.download().view(np.int64)).pin_memory()
sc = np.random.default_rng(rank).integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
scp = torch.from_numpy(sc.view(np.int64)).pin_memory()
for growth in [int(a) for a in sys.argv[1:]] or [0]:
    k.set_param("oneshot_growth", growth)
    for _ in range(4):
        dist.barrier(); k.msm_oneshot_ptr(0, pts.data_ptr(), n, scp.data_ptr(), n)
    ts = []
    for _ in range(8):
        dist.barrier()
        t0 = time.perf_counter(); k.msm_oneshot_ptr(0, pts.data_ptr(), n, scp.data_ptr(), n); ts.append(time.perf_counter() - t0)
    t = torch.tensor([sum(ts) / len(ts)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"N={world} growth={growth or 'auto'}: slowest rank mean {t.item()*1e3:.3f} ms per 2^20-point oneshot = {world * n / t.item() / 1e6:.0f} Mpoints/s aggregate", flush=True)
dist.destroy_process_group()
