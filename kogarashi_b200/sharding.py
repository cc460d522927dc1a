"""Multi-GPU sharding helpers for the one-process-per-GPU launch (torchrun).

MSM is linear, so the (points, scalars) vectors are split into contiguous shards, each rank runs an
independent MSM on its GPU, and the per-rank partial results (one 96-byte projective point each)
are added on the host.  There is no data-path collective: the only communication is the gather of
those points (groth16/src/msm.rs:45-47 does the same fold over windows).
The in-process multi-GPU path (kgr_init with several devices) uses the same split in msm.cu.
"""
import numpy as np

from .msm import proj_add


def shard_range(n, world, rank):
    """Contiguous even split: (first, count) of rank's shard of an n-element vector."""
    per = (n + world - 1) // world
    first = min(n, rank * per)
    return first, min(n, (rank + 1) * per) - first


def identity_projective(curve):
    """(0, R, 0): zkstd/src/macros/curve/weierstrass/group.rs:106-110."""
    from .msm import coord_limbs, to_affine
    cl = coord_limbs(curve)
    out = np.zeros(3 * cl, dtype=np.uint64)
    out[cl:2 * cl] = to_affine(curve, out)[cl:2 * cl]  # to_affine of Z = 0 returns (0, R, inf)
    return out


def combine_partials(curve, partials):
    total = identity_projective(curve)
    for p in partials:
        total = proj_add(curve, total, np.ascontiguousarray(p, dtype=np.uint64))
    return total


def gather_partials(partial, device=None):
    """all_gather of one projective point per rank (works on gloo/CPU and nccl/CUDA)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [partial]
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return [p.cpu().numpy().view(np.uint64) for p in parts]
