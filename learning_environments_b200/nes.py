"""Host side of the NES outer step: fitness shaping and the mirrored-sampling decision.

These are O(population) scalar computations on fp64 values gathered from the GPUs; the O(P x population)
parts (perturbation, weighted update) are device kernels (csrc/le_api.cu: nes_perturb / nes_update).
Semantics follow agents/GTN_master.py:197-265 (score_transform) and agents/GTN_worker.py:234-254
(calc_best_score), including numpy's default argsort order on ties and left-to-right fp64 summation.
"""
import numpy as np


def _seq_sum(values):
    """Left-to-right fp64 sum (python's builtin sum, as the reference uses) — np.sum's pairwise order differs by ulps."""
    total = 0.0
    for v in values:
        total = total + float(v)
    return total


def _best_only(scores):
    onehot = np.zeros(scores.size)
    onehot[int(np.argmax(scores))] = 1
    return onehot


def score_transform(scores, scores_orig, transform_type):
    """Fitness shaping of one generation. scores: best-of-mirror score per member; scores_orig: score of the
    unperturbed parameters per member. Returns the fp64 weight per member (score_transform_list)."""
    x = np.array(scores, dtype=np.float64)
    x0 = np.asarray(scores_orig, dtype=np.float64)
    n = x.size
    kind = int(transform_type)
    if kind == 0:                                   # min-max to [0, 1]
        lo, hi = x.min(), x.max()
        return (x - lo) / (hi - lo + 1e-9)
    if kind == 1:                                   # rank / (n - 1), ascending
        out = np.empty(n)
        out[np.argsort(x)] = np.arange(n) / (n - 1)
        return out
    if kind in (2, 3):                              # NES utilities (Wierstra et al. 2014), with / without zero mean
        rank = np.empty(n)
        rank[np.argsort(-x)] = np.arange(1, n + 1)
        util = np.maximum(0.0, np.log(n / 2 + 1) - np.log(rank))
        util = util / _seq_sum(util)
        if kind == 2:
            util = util - 1 / n
        return util / util.max()
    if kind == 4:                                   # single best perturbation
        return _best_only(x)
    if kind in (5, 6, 7):                           # only perturbations better than the unperturbed average
        avg = np.mean(x0)
        better = np.where(x > avg + 1e-6, 1, 0)
        if better.sum() == 0:
            return better.astype(np.float64) if kind != 5 else better
        if kind == 5:
            return _best_only(x)
        w = better * (x - avg) / (x.max() - avg + 1e-9)
        return w / (w.max() if kind == 6 else _seq_sum(w))
    raise ValueError("Unknown rank transform type: " + str(transform_type))


def best_of_mirror(score_add, score_sub, mirrored_sampling=True):
    """calc_best_score for already-reduced (mean / min over num_grad_evals) scores, vectorised over members.
    Returns (score_best, sign): sign = -1 where -eps was strictly better (eps is inverted), else +1."""
    add = np.asarray(score_add, dtype=np.float64)
    if not mirrored_sampling:
        return add.copy(), np.ones_like(add)
    sub = np.asarray(score_sub, dtype=np.float64)
    flip = sub > add
    return np.where(flip, sub, add), np.where(flip, -1.0, 1.0)
