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
from gadfit_b200 import GadfOptions, LMsolver, comm_unique_id, partition  # noqa: E402
from gadfit_b200 import workloads as W  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    report = {}
    cases = [("C3", dict(n_sets=37, n_points=400), None),    # whole data sets per rank
             ("C4", dict(n_sets=9, n_points=300), None),     # + geodesic acceleration
             ("C2", dict(n_sets=1, n_points=50_000), None),  # fewer sets than ranks: point split
             # the Fortran driver: its reduced scalars are all-reduced too
             ("C3", dict(n_sets=12, n_points=300), dict(nielsen=True, rel_error=1e-7)),
             ("C2", dict(n_sets=1, n_points=20_000), dict(chi2_rel=1e-6, grad_chi2=1e-3))]
    for i, (name, kw, gadf) in enumerate(cases):
        w = W.WORKLOADS[name](**kw)
        first, end, split = partition([w.n_points] * w.n_sets, rank, world)
        s = LMsolver(w.model)
        s.setDevice(local)
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        s.setComm(rank, world, ids[0])
        # each rank only hands over the data sets it owns (all of them when split)
        w.setup(s, owned=None if split else range(first, end))
        if gadf is not None:
            s.gadf = GadfOptions(enabled=True, **gadf)
            s.settings.iteration_limit = 12
        s.fit(w.lambda0)
        got = dict(chi2=s.chi2(), pars=[[s.getParValue(k, j) for k in range(w.n_pars)]
                                        for j in range(w.n_sets)],
                   decisions=[t["decision"] for t in s.trace()], iterations=s.iterations(),
                   launches=s.launches(), comm=s.commStats())
        if rank == 0:
            o = backends.oracle_solver(2)(w.model)
            w.setup(o)
            o.settings.n_threads = 4
            if gadf is not None:
                o.gadf = GadfOptions(enabled=True, **gadf)
                o.settings.iteration_limit = 12
            o.fit(w.lambda0)
            want = dict(chi2=o.chi2(), pars=[[o.getParValue(k, j) for k in range(w.n_pars)]
                                             for j in range(w.n_sets)],
                        decisions=[t["decision"] for t in o.trace()], iterations=o.iterations())
            got["stop"], want["stop"] = s.stopReason(), o.stopReason()
            report[name if gadf is None else name + "_gadf"] = dict(got=got, want=want, split=split)
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
