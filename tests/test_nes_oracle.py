"""NES outer-step restatement (oracle/nes.py) vs the reference's GTN_Master (golden) and SURVEY Appendix C."""
import numpy as np
import pytest

from oracle import nes
from tests.helpers import load_golden

APPENDIX_C = {
    0: [0.003148, 1.0, 0.136936, 0.0, 0.737671, 1.0, 0.265477, 0.013641],
    1: [0.142857, 0.857143, 0.428571, 0.0, 0.714286, 1.0, 0.571429, 0.285714],
    2: [-0.338994, 1.0, -0.338994, -0.338994, 0.085995, 0.423327, -0.153346, -0.338994],
    3: [0.0, 1.0, 0.0, 0.0, 0.317394, 0.569323, 0.138647, 0.0],
    4: [0, 1, 0, 0, 0, 0, 0, 0], 5: [0, 1, 0, 0, 0, 0, 0, 0],
    6: [0, 1.0, 0, 0, 0.666667, 1.0, 0.066667, 0],
    7: [0, 0.365854, 0, 0, 0.243902, 0.365854, 0.02439, 0],
}


def test_score_transform_appendix_c():
    sl = [10, 200, 35.5, 9.4, 150, 200, 60, 12]
    for t, want in APPENDIX_C.items():
        got = nes.score_transform(sl, [50] * 8, t)
        assert np.allclose(got, want, atol=1e-6), t


def test_score_transform_golden_all_types():
    g = load_golden("nes_cartpole.npz")
    for li in range(int(g["n_lists"])):
        for t in range(8):
            with np.errstate(all="ignore"):
                got = nes.score_transform(g["scores%d" % li], g["scores_orig%d" % li], t)
            want = g["transform%d_type%d" % (li, t)]
            assert np.array_equal(np.isnan(got), np.isnan(want)), (li, t)
            assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), (li, t)   # fp64, same op order: bit-exact


def test_score_transform_unknown_type():
    with pytest.raises(ValueError):
        nes.score_transform([1, 2], [1, 2], 8)


def test_update_env_golden():
    g = load_golden("nes_cartpole.npz")
    w = g["transform0_type3"]
    for wd, tag in ((0.0, "nowd"), (0.01, "wd")):
        got = nes.update_env(g["theta0"], g["eps"], w, float(g["step_size"]), weight_decay=wd)
        assert np.array_equal(got, g["theta_after_%s" % tag])   # fp32, sequential member order: bit-exact


def test_calc_best_score():
    assert nes.calc_best_score([10.0], [12.0]) == (12.0, -1.0)
    assert nes.calc_best_score([12.0], [12.0]) == (12.0, 1.0)      # ties keep +eps (strict >)
    assert nes.calc_best_score([5.0, 7.0], [3.0, 20.0], "minmax") == (5.0, 1.0)
    assert nes.calc_best_score([5.0], [9.0], mirrored_sampling=False) == (5.0, 1.0)


def test_product_score_transform_matches_reference_golden_and_oracle():
    """learning_environments_b200.nes (host logic of the product) vs golden (reference) and the oracle restatement."""
    from learning_environments_b200 import nes as pnes
    g = load_golden("nes_cartpole.npz")
    for li in range(int(g["n_lists"])):
        for t in range(8):
            with np.errstate(all="ignore"):
                got = np.asarray(pnes.score_transform(g["scores%d" % li], g["scores_orig%d" % li], t), dtype=np.float64)
            want = g["transform%d_type%d" % (li, t)]
            assert np.array_equal(np.isnan(got), np.isnan(want)), (li, t)
            assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), (li, t, got, want)
    rng = np.random.RandomState(5)
    for _ in range(50):
        n = rng.randint(2, 40)
        s = np.round(rng.uniform(0, 200, n), rng.randint(0, 3))     # rounding creates ties
        so = rng.uniform(0, 200, n)
        for t in range(8):
            with np.errstate(all="ignore"):
                a = np.asarray(pnes.score_transform(s, so, t), dtype=np.float64)
                b = nes.score_transform(s, so, t)
            assert np.array_equal(a, b, equal_nan=True), (t, s)
    with pytest.raises(ValueError):
        pnes.score_transform([1.0, 2.0], [1.0, 2.0], 9)
    sb, sg = pnes.best_of_mirror([10.0, 12.0, 5.0], [12.0, 12.0, 1.0])
    assert sb.tolist() == [12.0, 12.0, 5.0] and sg.tolist() == [-1.0, 1.0, 1.0]
