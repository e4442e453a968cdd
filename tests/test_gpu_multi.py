"""Multi-GPU parity (SURVEY.md §4 tier iv): GTN_Master on 1 GPU vs. on N ranks over NCCL — the per-generation all-gather of
fitness scores and either the replicated update (every rank regenerates all eps from Philox: theta bit-identical for any rank
count) or the sharded update + all-reduce (fp32 reassociation only).  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped on
a single-GPU box.  The CPU analogue over gloo is tests/test_host_logic.py::test_gtn_master_two_ranks_gloo_matches_single_rank."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["replicated", "allreduce"])
def test_gtn_master_theta_on_n_ranks_nccl_matches_one_gpu(tmp_path, mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    ranks = 4 if n >= 4 else 2
    script = os.path.join(ROOT, "tests", "nccl_gtn_worker.py")
    env = dict(os.environ, PYTHONPATH=ROOT)
    out1 = tmp_path / "single.npy"
    subprocess.check_call([sys.executable, script, "--mode", mode, "--out", str(out1)], env=env, cwd=str(tmp_path), timeout=900)
    out2 = tmp_path / "dist"
    port = 29500 + os.getpid() % 2000
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % ranks, "--master-addr",
                           "127.0.0.1", "--master-port", str(port), script, "--mode", mode, "--out", str(out2)], env=env,
                          cwd=str(tmp_path), timeout=900)
    single = np.load(out1)
    thetas = [np.load(str(out2) + ".rank%d.npy" % r) for r in range(ranks)]
    for t in thetas[1:]:
        assert np.array_equal(t, thetas[0])                    # every rank holds the same theta after 3 generations
    # a lane's result does not depend on the GPU it runs on: the gathered fitness scores equal the single-GPU ones exactly
    assert np.array_equal(np.load(str(out2) + ".scores.rank0.npy"), np.load(str(out1) + ".scores.npy"))
    if mode == "replicated":
        assert np.array_equal(thetas[0], single)               # bit-identical for any rank count
    else:
        assert np.allclose(thetas[0], single, rtol=0, atol=1e-7 * max(1.0, np.abs(single).max()))
