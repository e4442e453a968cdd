"""PopulationEvaluator — one NES generation's population evaluation on one GPU.

Replaces what `num_workers` GTN_Worker processes do per generation (agents/GTN_worker.py:76-114): for every
member i the three fitness evaluations calc_score(theta), calc_score(theta+eps_i), calc_score(theta-eps_i)
(x num_grad_evals) run as independent lanes of ONE persistent kernel launch; eps_i is regenerated from the
Philox (seed, generation, member) stream on the device and never stored or communicated.
"""
import ctypes as C

import numpy as np
import torch

from . import ops
from ._abi import LaneCfg
from .rng import lane_keys


class LaneLayout(object):
    """Host-side bookkeeping: which lane evaluates which (member, variant, eval) and how lane scores fold back
    into (score_orig, score_add, score_sub) per member (agents/GTN_worker.py:84-104, 234-242)."""

    def __init__(self, cfg, pop, member_lo=0, member_hi=None, num_grad_evals=1, seed=0, noise_std=0.01, mirrored=True):
        self.cfg = cfg
        self.pop = int(pop)
        self.member_lo = int(member_lo)
        self.member_hi = int(pop if member_hi is None else member_hi)
        self.n_members = self.member_hi - self.member_lo
        self.num_grad_evals = int(num_grad_evals)
        self.seed = int(seed)
        self.noise_std = float(noise_std)
        self.mirrored = bool(mirrored)
        # lanes: member-major, then variant (0: theta, 1: +eps, 2: -eps), then grad-eval index.
        # The unperturbed theta is evaluated once per member (score_orig, agents/GTN_worker.py:84).
        self.variants = 3 if mirrored else 2
        self.lanes_per_member = 1 + (self.variants - 1) * self.num_grad_evals
        self.n_lanes = self.n_members * self.lanes_per_member
        self.n_env = self.n_members * 3
        self.P = cfg.env_params()
        self.env_index_host = np.zeros(self.n_lanes, np.int32)
        self.lane_member = np.zeros(self.n_lanes, np.int32)
        self.lane_variant = np.zeros(self.n_lanes, np.int32)
        self.lane_eval = np.zeros(self.n_lanes, np.int32)
        k = 0
        for m in range(self.n_members):
            for v in range(self.variants):
                for e in range(1 if v == 0 else self.num_grad_evals):
                    self.env_index_host[k] = m * 3 + v
                    self.lane_member[k], self.lane_variant[k], self.lane_eval[k] = self.member_lo + m, v, e
                    k += 1
        assert k == self.n_lanes

    def lane_keys(self, generation):
        return lane_keys(self.seed, generation, self.lane_member, self.lane_variant, self.lane_eval)

    def member_scores(self, out, grad_eval_type="mean"):
        """(score_orig[n_members], score_add[n_members], score_sub[n_members]) from the lane scores."""
        if grad_eval_type not in ("mean", "minmax"):
            raise NotImplementedError("Unknown parameter for grad_eval_type: " + str(grad_eval_type))
        red = np.mean if grad_eval_type == "mean" else np.min
        # lanes are member-major: [theta | +eps x num_grad_evals | -eps x num_grad_evals] per member
        E = self.num_grad_evals
        sc = np.asarray(out["score"], np.float64).reshape(self.n_members, self.lanes_per_member)
        orig = sc[:, 0].copy()
        add = red(sc[:, 1:1 + E], axis=1)
        sub = red(sc[:, 1 + E:1 + 2 * E], axis=1) if self.mirrored else np.zeros(self.n_members)
        return orig, add, sub


class PopulationEvaluator(LaneLayout):
    def __init__(self, cfg, pop, member_lo=0, member_hi=None, num_grad_evals=1, seed=0, noise_std=0.01, device="cuda",
                 mirrored=True):
        super().__init__(cfg, pop, member_lo, member_hi, num_grad_evals, seed, noise_std, mirrored)
        self.device = torch.device(device)
        self.bufs = ops.InnerLoopBuffers(cfg, self.n_lanes, self.n_env, self.device)
        self.env_index = torch.from_numpy(self.env_index_host).to(self.device)
        self._theta_dev = torch.empty(self.P, dtype=torch.float32, device=self.device)
        self._keys_host = torch.empty((self.n_lanes, 2), dtype=torch.int32).pin_memory()
        self._keys_dev = torch.empty((self.n_lanes, 2), dtype=torch.int32, device=self.device)
        self._out_host = torch.empty(self.bufs.out.shape, dtype=torch.uint8).pin_memory()

    @property
    def h2d_bytes(self):
        return self.P * 4 + self.n_lanes * 8 + C.sizeof(LaneCfg)

    @property
    def d2h_bytes(self):
        return self.bufs.out.numel()

    def launch(self, theta_host, generation):
        """H2D of theta + lane keys, perturbation, fused kernel — all asynchronous on the current stream."""
        th = theta_host if torch.is_tensor(theta_host) else torch.from_numpy(np.ascontiguousarray(theta_host, np.float32))
        self._theta_dev.copy_(th.reshape(-1), non_blocking=True)
        keys = self.lane_keys(generation)
        self._keys_host.copy_(torch.from_numpy(keys.view(np.int32)))
        self._keys_dev.copy_(self._keys_host, non_blocking=True)
        thetas = ops.nes_perturb(self._theta_dev, self.pop, self.member_lo, self.n_members, self.seed, generation, self.noise_std)
        ops.inner_loop_run(self.bufs, self.cfg, thetas, self.env_index, self._keys_dev)
        self._thetas = thetas  # keep alive until the kernel has run

    def collect(self):
        """D2H of the per-lane results (blocking). Returns the structured array (ops.lane_out_dtype)."""
        self._out_host.copy_(self.bufs.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return np.frombuffer(self._out_host.numpy().tobytes(), dtype=ops.lane_out_dtype())

    def evaluate(self, theta_host, generation):
        self.launch(theta_host, generation)
        return self.collect()
