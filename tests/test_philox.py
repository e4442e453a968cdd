"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the numpy and C restatements."""
import numpy as np

from oracle import c_oracle, philox

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_kat_python_numpy_c():
    for ctr, key, want in KAT:
        assert philox.philox4x32(*ctr, *key) == want
        assert tuple(int(x) for x in philox.philox4x32_np(*[np.uint64(c) for c in ctr], *key).reshape(-1)) == want
        assert c_oracle.philox(*ctr, *key) == want


def test_sample_indices_match_scalar():
    key = (123, 456)
    idx = philox.sample_indices(key, 7, 199, 1000)
    assert idx.shape == (199,) and idx.min() >= 0 and idx.max() < 1000
    for j in (0, 1, 49):
        w = philox.philox4x32(7, j, philox.P_SAMPLE, 0, *key)
        for k in range(4):
            if 4 * j + k < 199:
                assert idx[4 * j + k] == (w[k] * 1000) >> 32


def test_normals_moments_and_determinism():
    z = philox.normals(5, 3, 2, 20000)
    assert abs(float(z.mean())) < 0.03 and abs(float(z.std()) - 1.0) < 0.03
    assert np.array_equal(z[:100], philox.normals(5, 3, 2, 100))
    assert not np.array_equal(z[:100], philox.normals(5, 3, 3, 100))


def test_qinit_matches_c():
    from learning_environments_b200._abi import LaneCfg
    c = LaneCfg()
    c.sd, c.ad, c.q_hidden = 4, 2, 57
    P = c.q_params()
    bounds = np.concatenate([np.full(57 * 4 + 57, 1 / np.sqrt(4.0)), np.full(2 * 57 + 2, 1 / np.sqrt(57.0))])
    a = philox.qnet_init((9, 8), P, bounds)
    b = c_oracle.q_init(c, (9, 8))
    assert np.array_equal(a, b)
    assert np.all(np.abs(a[:57 * 5]) <= 0.5)
