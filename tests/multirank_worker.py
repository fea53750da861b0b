"""Worker of tests/test_gpu_multirank.py (run under torch.distributed.run, one
rank per GPU): a global fit sharded over the ranks through the public API +
NCCL all-reduce inside the library, compared with the single-process oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import backends  # noqa: E402
from gadfit_b200 import LMsolver, comm_unique_id, partition  # noqa: E402
from gadfit_b200 import workloads as W  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    report = {}
    cases = [("C3", dict(n_sets=37, n_points=400)),      # whole data sets per rank
             ("C4", dict(n_sets=9, n_points=300)),       # + geodesic acceleration
             ("C2", dict(n_sets=1, n_points=50_000))]    # fewer sets than ranks: point split
    for i, (name, kw) in enumerate(cases):
        w = W.WORKLOADS[name](**kw)
        first, end, split = partition([w.n_points] * w.n_sets, rank, world)
        s = LMsolver(w.model)
        s.setDevice(local)
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        s.setComm(rank, world, ids[0])
        # each rank only hands over the data sets it owns (all of them when split)
        w.setup(s, owned=None if split else range(first, end))
        s.fit(w.lambda0)
        got = dict(chi2=s.chi2(), pars=[[s.getParValue(k, j) for k in range(w.n_pars)]
                                        for j in range(w.n_sets)],
                   decisions=[t["decision"] for t in s.trace()], iterations=s.iterations(),
                   launches=s.launches(), comm=s.commStats())
        if rank == 0:
            o = backends.oracle_solver(2)(w.model)
            w.setup(o)
            o.settings.n_threads = 4
            o.fit(w.lambda0)
            want = dict(chi2=o.chi2(), pars=[[o.getParValue(k, j) for k in range(w.n_pars)]
                                             for j in range(w.n_sets)],
                        decisions=[t["decision"] for t in o.trace()], iterations=o.iterations())
            report[name] = dict(got=got, want=want, split=split)
        # all ranks must hold the same parameters after the fit
        t = torch.tensor(got["pars"], dtype=torch.float64)
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), f"rank {rank}: parameters differ from rank 0 ({name})"
        del s
    if rank == 0:
        json.dump(report, open(sys.argv[1], "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
