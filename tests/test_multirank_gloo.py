"""CPU, world_size = 2 over gloo: the host-side logic of the multi-GPU path.

The N > 1 design shards whole data sets across
ranks with `gadfit_b200_partition` and all-reduces ONLY the packed shared
block [J^T J (shared x shared), J^T r (shared), chi2].  Here every rank runs
the oracle on the data sets the partition gives it, the packed buffer is
summed with torch.distributed (gloo), and the result must equal what a single
rank computes for the whole problem.  The same is checked for the point-split
mode used when there are fewer data sets than ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shared_block(w, owned, point_slice=None):
    """Oracle on a subset: packed [G (pg x pg), b_g (pg), chi2 at the start]."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import backends
    s = backends.oracle_solver(2)(w.model)
    for i in owned:
        xs, ys, es = w.slice(i)
        if point_slice is not None:
            a, b = point_slice(len(xs))
            xs, ys, es = xs[a:b], ys[a:b], es[a:b]
        s.addDataset(xs, ys, es)
    for k in w.shared:
        s.setPar(k, float(w.start[0, k]), True)
    for j, i in enumerate(owned):
        for k in w.local:
            s.setPar(k, float(w.start[i, k]), True, j)
    s.settings.iteration_limit = 1
    s.fit(w.lambda0)
    n = int(round(np.sqrt(s.getJTJ().size)))
    pg = len(w.shared)
    jtj = s.getJTJ().reshape(n, n)
    return np.concatenate([jtj[:pg, :pg].ravel(), s.getRightSide()[:pg], [s.trace()[0]["old_chi2"]]])


def _worker(rank, world, port, split_mode, out):
    sys.path.insert(0, ROOT)
    from gadfit_b200 import partition, workloads as W
    dist.init_process_group("gloo", rank=rank, world_size=world,
                            init_method=f"tcp://127.0.0.1:{port}")
    if split_mode:
        w = W.c2(n_sets=1, n_points=4001)
        first, end, split = partition([w.n_points], rank, world)
        assert (first, end, split) == (0, 1, True)
        packed = _shared_block(w, range(first, end),
                               lambda n: (n * rank // world, n * (rank + 1) // world))
    else:
        w = W.c3(n_sets=7, n_points=150)
        first, end, split = partition([w.n_points] * w.n_sets, rank, world)
        assert not split
        packed = _shared_block(w, range(first, end))
    t = torch.from_numpy(packed.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    bounds = torch.tensor([first, end], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, bounds)
    if rank == 0:
        full = _shared_block(w, range(w.n_sets))
        out["reduced"] = t.numpy().tolist()
        out["full"] = full.tolist()
        out["bounds"] = [g.tolist() for g in gathered]
        out["n_sets"] = w.n_sets
    dist.destroy_process_group()


@pytest.mark.parametrize("split_mode", [False, True], ids=["data_set_sharded", "point_split"])
def test_sharded_shared_block_equals_single_rank(split_mode):
    world = 2
    with mp.Manager() as manager:
        out = manager.dict()
        mp.spawn(_worker, args=(world, _free_port(), split_mode, out), nprocs=world, join=True)
        reduced, full = np.array(out["reduced"]), np.array(out["full"])
        bounds, n_sets = out["bounds"], out["n_sets"]
    # every data set has exactly one owner (or, split mode, every rank all sets)
    if not split_mode:
        assert bounds[0][0] == 0 and bounds[-1][1] == n_sets and bounds[0][1] == bounds[1][0]
    # entries span ten orders of magnitude: absolute floor relative to the largest
    np.testing.assert_allclose(reduced, full, rtol=1e-12, atol=1e-13 * np.max(np.abs(full)))
