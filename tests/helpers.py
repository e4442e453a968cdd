"""Shared helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

from learning_environments_b200._abi import LaneCfg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def cfg_from_bytes(arr):
    c = LaneCfg()
    raw = np.ascontiguousarray(arr, np.uint8).tobytes()
    assert len(raw) == C.sizeof(LaneCfg), "golden le_lane_cfg size mismatch: regenerate fixtures"
    C.memmove(C.byref(c), raw, len(raw))
    return c


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0


def sync_prefix(gold_actions, got_actions):
    """Number of leading steps on which two action traces agree."""
    n = min(len(gold_actions), len(got_actions))
    neq = np.nonzero(np.asarray(gold_actions[:n]) != np.asarray(got_actions[:n]))[0]
    return int(neq[0]) if neq.size else n


def assert_params_close(got, want, lr, what=""):
    """Adam moves every parameter by <= lr per step whatever the gradient's magnitude, so ulp-level differences in
    near-zero gradients appear as a small fraction of lr on a few parameters: demand |diff| <= 2% of lr everywhere and
    5e-5 relative (to the parameter scale, floor 1e-2) on >= 99% of the parameters."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    d = np.abs(got - want)
    assert d.max() <= 0.02 * lr, "%s: max |diff| %.3g exceeds 2%% of lr=%g" % (what, d.max(), lr)
    rel = d / np.maximum(np.abs(want), 1e-2)
    assert (rel < 5e-5).mean() >= 0.99, "%s: only %.2f%% of the parameters within 5e-5" % (what, 100 * (rel < 5e-5).mean())
