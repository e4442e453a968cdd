"""Host-side Philox4x32-10 (numpy) used only to derive the per-lane 64-bit keys handed to the kernels.

The device kernels own every random stream of the hot path (csrc/le_common.cuh); the host never draws the
random numbers themselves.  Key derivation: one Philox block keyed by (seed, 'LANE') on the counter
(generation, member, variant, eval) — so (seed, generation, member, variant, eval) fixes a lane's whole
trajectory regardless of how lanes are sharded over GPUs.
"""
import numpy as np

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
LANE_TAG = 0x4C414E45  # 'LANE'


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(int(k0) & 0xFFFFFFFF)
    k1 = np.uint64(int(k1) & 0xFFFFFFFF)
    s32 = np.uint64(32)
    for _ in range(rounds):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        hi0, lo0 = p0 >> s32, p0 & _MASK
        hi1, lo1 = p1 >> s32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def lane_keys(seed, generation, member, variant, eval_idx):
    """uint32 [n, 2] lane keys for arrays of (member, variant, eval_idx)."""
    w = philox4x32(np.uint64(generation), np.asarray(member, np.uint64), np.asarray(variant, np.uint64),
                   np.asarray(eval_idx, np.uint64), seed, LANE_TAG)
    return np.ascontiguousarray(w.reshape(-1, 4)[:, :2])
