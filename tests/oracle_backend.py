"""Test infrastructure: an oracle-backed stand-in for the CUDA evaluator / NES kernels so that the HOST logic of
GTN_Master (sharding, all-gather, score transform, update order, checkpointing) can be exercised without a GPU,
including world_size-2 gloo runs.  Lane scores come from oracle/le_oracle.c, noise from oracle/philox.py."""
import numpy as np
import torch

from learning_environments_b200 import ops as le_ops
from learning_environments_b200.engine import LaneLayout
from oracle import c_oracle, nes as onese


class OracleEvaluator(LaneLayout):
    def __init__(self, cfg, pop, member_lo=0, member_hi=None, num_grad_evals=1, seed=0, noise_std=0.01, mirrored=True, **kw):
        super().__init__(cfg, pop, member_lo, member_hi, num_grad_evals, seed, noise_std, mirrored)

    def evaluate(self, theta_host, generation):
        theta = np.asarray(theta_host, np.float32).reshape(-1)
        thetas = np.zeros((self.n_members * 3, theta.size), np.float32)
        for m in range(self.n_members):
            eps = onese.noise(self.seed, generation, self.member_lo + m, theta.size, self.noise_std)
            thetas[3 * m], thetas[3 * m + 1], thetas[3 * m + 2] = theta, theta + eps, theta - eps
        res = c_oracle.run_lanes(self.cfg, thetas, self.env_index_host, self.lane_keys(generation), n_threads=4)
        out = np.zeros(self.n_lanes, dtype=le_ops.lane_out_dtype())
        for f in ("n_episodes", "timed_out", "train_steps", "learn_iters", "test_steps", "score"):
            out[f] = res[f]
        return out


def nes_update_numpy(theta, pop, seed, generation, noise_std, weight_decay, coef, sign):
    th = theta.cpu().numpy()
    P = th.size
    eps = np.stack([onese.noise(seed, generation, i, P, noise_std) * np.float32(sign[i].item()) for i in range(pop)])
    new = (th * np.float32(1.0 - weight_decay)).astype(np.float32)
    for i in range(pop):
        new = (new + np.float32(coef[i].item()) * eps[i]).astype(np.float32)
    theta.copy_(torch.from_numpy(new))


def nes_partial_update_numpy(P, lo, hi, seed, generation, noise_std, coef, sign):
    acc = np.zeros(P, np.float32)
    for i in range(lo, hi):
        eps = onese.noise(seed, generation, i, P, noise_std) * np.float32(sign[i].item())
        acc = (acc + np.float32(coef[i].item()) * eps).astype(np.float32)
    return torch.from_numpy(acc)


def patch_master_for_cpu(monkeypatch_or_module):
    """Route GTN_Master's device calls to the oracle-backed numpy versions (CPU tests only)."""
    from learning_environments_b200 import gtn
    setter = monkeypatch_or_module.setattr if hasattr(monkeypatch_or_module, "setattr") else None
    pairs = [(gtn.ops, "nes_update", nes_update_numpy), (gtn.ops, "nes_partial_update", nes_partial_update_numpy),
             (gtn.GTN_Master, "_nes_device", lambda self: torch.device("cpu"))]
    for obj, name, fn in pairs:
        if setter:
            setter(obj, name, fn)
        else:
            setattr(obj, name, fn)
