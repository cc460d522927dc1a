"""CPU test of the N > 1 host logic: contiguous sharding + gather of one partial point per rank + host-side
combination, on world_size 2 with the gloo backend.  The per-shard MSM is computed by the oracle here
(no GPU in this container); on the GPU box the same helpers are used by bench.py with the CUDA engine."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import same_affine

from oracle import oracle as A


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kogarashi_b200 import sharding
    curve = A.BN254_G1
    pts = A.random_points(curve, n, seed=bytes(range(16)), threads=2)
    sc = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(1, 17)))
    first, count = sharding.shard_range(n, world, rank)
    partial = A.msm(curve, pts[first:first + count], sc[first:first + count], threads=2)
    parts = sharding.gather_partials(partial)
    if rank == 0:
        total = sharding.combine_partials(curve, parts)
        import kogarashi_b200 as k
        q.put((k.to_affine(curve, total), A.to_affine(curve, A.msm(curve, pts, sc, threads=2)), len(parts)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_combine():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 301  # odd: ragged last shard
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, exp, nparts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert nparts == 2 and same_affine(got, exp)


def test_shard_range_covers_everything():
    from kogarashi_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 9, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, count in spans:
                assert first == pos or count == 0
                pos += count
