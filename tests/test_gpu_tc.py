"""tcgen05 / TMEM dense-layer GEMM (csrc/le_tc.cuh, 3xTF32) against float64 numpy, through the C ABI (le_tc_gemm).

The three operand forms are the three contractions of one nn.Linear under autograd: forward X W^T (+ bias, activation),
input gradient dZ W, weight gradient dZ^T X — at the shapes of the reference's DuelingDDQN nets
(default_config_cartpole_syn_env.yaml: 60/61-wide layers, B = 193; default_config_acrobot.yaml: 128x128, B = 128) and at the
edges of the DDQN_vary range (B = 597 rows, 171-wide layers).  Tolerance: 3xTF32 keeps fp32-level accuracy — the error is
held to 2e-6 of the scale of the summed terms (fp32 FFMA accumulation itself is ~1e-6 at these depths)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(rng, *shape):
    return (rng.standard_normal(shape) * rng.uniform(0.2, 2.0)).astype(np.float32)


@pytest.mark.parametrize("I,J,L", [(128, 128, 128), (193, 60, 61), (193, 61, 4), (128, 3, 128), (597, 171, 171), (64, 48, 48), (149, 112, 112)])
@pytest.mark.parametrize("form", ["nt", "nn", "tn"])
def test_tc_gemm_matches_float64(form, I, J, L):
    from learning_environments_b200 import ops
    rng = np.random.RandomState(I * 7 + J * 3 + L + len(form))
    A = _rand(rng, I, L) if form != "tn" else _rand(rng, L, I)
    B = _rand(rng, J, L) if form == "nt" else _rand(rng, L, J)
    a64 = (A if form != "tn" else A.T).astype(np.float64)
    b64 = (B.T if form == "nt" else B).astype(np.float64)
    want = a64 @ b64
    scale = (np.abs(a64) @ np.abs(b64)).max()
    C = torch.full((I, J), 7.0, dtype=torch.float32, device="cuda")
    ops.tc_gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), C, form, I, J, L)
    torch.cuda.synchronize()
    got = C.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 2e-6 * scale, (np.abs(got - want).max(), scale)


def test_tc_gemm_bias_activation_and_accumulate():
    from learning_environments_b200 import ops
    rng = np.random.RandomState(5)
    I, J, L = 193, 60, 61
    X, W, b = _rand(rng, I, L) * 0.3, _rand(rng, J, L) * 0.3, _rand(rng, J)
    z = X.astype(np.float64) @ W.astype(np.float64).T + b
    for act, ref in ((0, z), (1, np.tanh(z)), (2, np.maximum(z, 0.01 * z))):
        C = torch.zeros((I, J), dtype=torch.float32, device="cuda")
        ops.tc_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), C, "nt", I, J, L, bias=torch.from_numpy(b).cuda(), act=act, slope=0.01)
        torch.cuda.synchronize()
        assert np.abs(C.cpu().numpy() - ref).max() <= 3e-6 * max(1.0, np.abs(z).max()), act
    # accumulate: C += A B (the dueling advantage stream adds its input gradient to the value stream's)
    C0 = _rand(rng, I, J)
    C = torch.from_numpy(C0.copy()).cuda()
    ops.tc_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), C, "nt", I, J, L, accumulate=True)
    torch.cuda.synchronize()
    want = C0 + X.astype(np.float64) @ W.astype(np.float64).T
    assert np.abs(C.cpu().numpy() - want).max() <= 3e-6 * np.abs(want).max()
