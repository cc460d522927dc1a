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


def test_bench_scalars_do_not_depend_on_the_sharding():
    """bench.py draws the scalar vector in global blocks: any shard of it is the same numbers (strong scaling needs ONE vector for every N)."""
    import bench
    whole = bench.scalars_range(0, 3 * bench.SC_BLOCK + 17)
    for first, n in ((0, 5), (bench.SC_BLOCK - 3, 10), (bench.SC_BLOCK, bench.SC_BLOCK), (2 * bench.SC_BLOCK + 9, bench.SC_BLOCK + 8)):
        assert (bench.scalars_range(first, n) == whole[first:first + n]).all()
    assert (whole[:, 3] < np.uint64(bench.FR_TOP)).all()


def _strong_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from kogarashi_b200 import sharding
    from oracle import pyref as B
    curve = A.BN254_G1
    first, count = sharding.shard_range(n, world, rank)
    ks = A.bench_scalars(curve, bench.BASE_SEED, first, count)     # the stream kgr_bases_generate_at(first) uses on the device
    pts = A.fixed_base(curve, ks, threads=2)
    sc = bench.scalars_range(first, count)
    partial = A.msm(curve, pts, sc, threads=2)                      # stands in for the rank's GPU MSM
    dot = bench.dot_of_shard(curve, ks, sc)
    parts = sharding.gather_partials(partial)
    dots = sharding.gather_partials(dot)
    if rank == 0:
        import kogarashi_b200 as k
        total = k.to_affine(curve, sharding.combine_partials(curve, parts))
        acc = sum(B.limbs_to_int(d) for d in dots) % B.FR
        exp = bench.expected_from_dot(curve, np.array(B.int_to_limbs(acc), dtype=np.uint64))
        # the same vector in one piece
        ks_all = A.bench_scalars(curve, bench.BASE_SEED, 0, n)
        one = A.to_affine(curve, A.msm(curve, A.fixed_base(curve, ks_all, threads=2), bench.scalars_range(0, n), threads=2))
        q.put((total, exp, one))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_strong_scaling_checksum():
    """The N > 1 leg of bench.py on CPU: shards of ONE seeded vector, partial points gathered and summed, and the all-gathered
    sum k_i s_i giving the same point — which is also the unsharded MSM."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strong_worker, args=(r, 2, port, 515, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, exp, one = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same_affine(total, exp) and same_affine(total, one)
