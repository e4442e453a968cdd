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
