"""GPU, N = 2 ranks (skipped on a single-GPU box): data sets sharded across
GPUs, only the shared normal block / chi2 all-reduced over NCCL inside the
library.  Must reproduce the single-process oracle: same accept/reject
sequence, parameters and chi2 to 1e-10."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _n_gpus() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_ranks_reproduce_the_single_process_oracle(tmp_path):
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "multirank_worker.py"), str(out)]
    subprocess.run(cmd, check=True, timeout=600)
    report = json.load(open(out))
    for name, r in report.items():
        got, want = r["got"], r["want"]
        assert got["launches"] > 0
        # the per-trial reductions ran over NVLink peer memory, NCCL only for
        # the large / rare ones (csrc/peer_reduce.cuh)
        assert got["comm"]["peer_reduce_active"], got["comm"]
        assert got["comm"]["peer_reductions"] > 0
        n = min(len(got["decisions"]), len(want["decisions"]), 6)
        assert got["decisions"][:n] == want["decisions"][:n], name
        np.testing.assert_allclose(got["chi2"], want["chi2"], rtol=1e-10, err_msg=name)
        gp, wp = np.array(got["pars"]), np.array(want["pars"])
        floor = 1e-9 * np.max(np.abs(wp), axis=0, keepdims=True)
        assert np.all(np.abs(gp - wp) <= 1e-8 * np.abs(wp) + floor), name
        if name.endswith("_gadf"):  # the Fortran driver: same stop criterion on every path
            assert got["stop"] == want["stop"] and got["iterations"] == want["iterations"], name
