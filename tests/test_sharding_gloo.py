"""world_size-2 gloo test of the unit sharding + final gather used for fits / trace blocks (CPU, no kernels)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from circuitmap_b200.sharding import shard_range, run_sharded


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 20001):
        for w in (1, 2, 3, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_units, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def unit_fn(lo, hi):                      # stands in for a batch of fits: result depends on the unit index only
        idx = torch.arange(lo, hi, dtype=torch.float64)
        return {"mu": torch.stack([idx, idx * idx], 1), "shape": idx + 0.5}

    out = run_sharded(n_units, unit_fn)
    if rank == 0:
        ret["mu"] = out["mu"].numpy()
        ret["shape"] = out["shape"].numpy()
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process():
    n_units = 11
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_units, ret), nprocs=2, join=True)
    idx = np.arange(n_units, dtype=np.float64)
    assert np.array_equal(ret["mu"], np.stack([idx, idx * idx], 1))
    assert np.array_equal(ret["shape"], idx + 0.5)
