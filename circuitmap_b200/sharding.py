"""Sharding of independent units (fits, trace blocks) over the GPUs of one box.

The path shards without any data-path collective: simulation sweeps, LOHO-CV folds and demixer trace blocks are
independent (reference: one python process / SLURM task per unit, scripts/run_simulations.py:12-23,
scripts/generate_loho_cv_slurm_scripts.py:108-117).  One process per GPU; a static block partition of the unit
index; results are gathered once at the end with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_units, rank, world_size):
    """Contiguous block [lo, hi) of unit indices owned by `rank`; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(n_units), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_sharded(n_units, unit_fn, group=None):
    """Every rank runs `unit_fn(lo, hi)` on its block and returns a dict of equally-shaped-per-unit torch tensors
    (first dim = hi - lo).  Rank 0 gets the concatenation over ranks in unit order, the others None."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return unit_fn(0, n_units)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_units, rank, world)
    local = unit_fn(lo, hi)
    out = {}
    for key in sorted(local):
        t = local[key].contiguous()
        sizes = [shard_range(n_units, r, world) for r in range(world)]
        if rank == 0:
            bufs = [torch.empty((b - a,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for a, b in sizes]
        else:
            bufs = None
        # ragged first dimensions: point-to-point to rank 0 (works on NCCL and gloo alike); one message per rank
        if rank == 0:
            bufs[0].copy_(t)
            for r in range(1, world):
                if sizes[r][1] > sizes[r][0]:
                    dist.recv(bufs[r], src=r, group=group)
        elif hi > lo:
            dist.send(t, dst=0, group=group)
        if rank == 0:
            out[key] = torch.cat(bufs, dim=0)
    return out if rank == 0 else None
