"""Multi-process host logic (world_size 2, gloo, CPU): sharding + final gather reproduce the
single-process result bit for bit.  The per-rank solve is a CPU stand-in (the C oracle) -- the
scheduler does not care what solves its shard."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_blaster_b200 import scenarios as sc
from mpc_blaster_b200.scheduler import ShardedSolve, gather_batch, shard_range


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 1024, 1025):
        for W in (1, 2, 3, 8):
            r = [shard_range(B, g, W) for g in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[g][1] == r[g + 1][0] for g in range(W - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import blaster_oracle as bo, c_oracle as co
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = bo.canonical_problem(10)
    x0, yref = sc.random_setpoints(B, seed=42)
    lo, hi = shard_range(B, rank, world)
    orc = co.BatchRTI(P, hi - lo, nthreads=1)

    def solve_fn(x, y, p):
        u0, X, U, st = orc.solve(x.numpy(), y.numpy())
        return torch.from_numpy(u0), None, None, torch.from_numpy(st)

    sh = ShardedSolve(solve_fn, B)
    assert (sh.lo, sh.hi) == (lo, hi)
    u0, st = sh.solve(torch.from_numpy(x0), torch.from_numpy(yref))
    if rank == 0:
        np.save(os.path.join(out_dir, "u0.npy"), u0.numpy())
        np.save(os.path.join(out_dir, "st.npy"), st.numpy())
    # ragged gather on its own
    t = torch.full((hi - lo, 2), float(rank))
    g = gather_batch(t, B)
    assert g.shape == (B, 2) and float(g[0, 0]) == 0.0 and float(g[-1, 0]) == world - 1
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_process(tmp_path):
    from oracle import blaster_oracle as bo, c_oracle as co
    B = 9  # odd on purpose: ragged shards
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    P = bo.canonical_problem(10)
    x0, yref = sc.random_setpoints(B, seed=42)
    u0, _, _, st = co.BatchRTI(P, B, nthreads=1).solve(x0, yref)
    assert np.array_equal(np.load(tmp_path / "u0.npy"), u0)
    assert np.array_equal(np.load(tmp_path / "st.npy"), st)


def test_shard_treats_one_dimensional_tensors_as_shared_whatever_their_length():
    """A shared yref[ny] with B == ny (23) or a shared p[25] with B == 25 must not be sliced; per-instance tensors lead
    with B and anything else is an error (the solver's own shape rule, BlasterMPC._mode)."""
    import pytest
    import torch
    from mpc_blaster_b200.scheduler import shard
    for B in (23, 25, 17):
        y = torch.arange(23.0)
        p = torch.arange(25.0)
        assert shard(y, B, 1, 2) is y and shard(p, B, 1, 2) is p
        x0 = torch.zeros(B, 17)
        lo = B // 2 + (B % 2)
        assert shard(x0, B, 1, 2).shape == (B - lo, 17)
        assert shard(torch.zeros(B, 21, 23), B, 0, 2).shape == (lo, 21, 23)
    with pytest.raises(ValueError):
        shard(torch.zeros(5, 23), 23, 0, 2)
    assert shard(None, 8, 0, 2) is None
