"""TEST INFRASTRUCTURE ONLY — numpy restatement of the NES outer step.

  score_transform   agents/GTN_master.py:197-265 (second restatement: experiments/demo_score_transform.py:9-82)
  update_env        agents/GTN_master.py:267-298 (sequential member order, fp32 accumulate)
  calc_best_score   agents/GTN_worker.py:234-254
  noise             agents/GTN_worker.py:156-163 with torch.normal replaced by the P_NOISE Philox stream
"""
import statistics

import numpy as np

from . import philox


def score_transform(score_list, score_orig_list, transform_type):
    scores = np.asarray(score_list, dtype=np.float64).copy()
    scores_orig = np.asarray(score_orig_list, dtype=np.float64)
    t = transform_type
    if t == 0:
        scores = (scores - min(scores)) / (max(scores) - min(scores) + 1e-9)
    elif t == 1:
        s = np.argsort(scores)
        n = len(scores)
        for i in range(n):
            scores[s[i]] = i / (n - 1)
    elif t in (2, 3):
        lmbda = len(scores)
        s = np.argsort(-scores)
        for i in range(lmbda):
            scores[s[i]] = i + 1
        for i in range(lmbda):
            scores[i] = max(0, np.log(lmbda / 2 + 1) - np.log(scores[i]))
        scores = scores / sum(scores)
        if t == 2:
            scores -= 1 / lmbda
        scores /= max(scores)
    elif t == 4:
        tmp = np.zeros(scores.size)
        tmp[np.argmax(scores)] = 1
        scores = tmp
    elif t == 5:
        avg = np.mean(scores_orig)
        idx = np.where(scores > avg + 1e-6, 1, 0)
        if sum(idx) > 0:
            tmp = np.zeros(scores.size)
            tmp[np.argmax(scores)] = 1
            scores = tmp
        else:
            scores = idx
    elif t in (6, 7):
        avg = np.mean(scores_orig)
        idx = np.where(scores > avg + 1e-6, 1, 0)
        if sum(idx) > 0:
            scores = idx * (scores - avg) / (max(scores) - avg + 1e-9)
            if t == 6:
                scores /= max(scores)
            else:
                scores /= sum(scores)
        else:
            scores = idx
    else:
        raise ValueError("Unknown rank transform type: " + str(t))
    return np.asarray(scores, dtype=np.float64)


def update_env(theta, eps_list, score_transform_list, step_size, weight_decay=0.0, nes_step_size=False):
    """theta float32 [P]; eps_list float32 [pop][P] (already signed). Returns new theta (float32)."""
    ss = step_size / len(eps_list) if nes_step_size else step_size
    th = np.asarray(theta, np.float32).copy()
    th = (th * np.float32(1 - weight_decay)).astype(np.float32)
    for eps, w in zip(eps_list, score_transform_list):
        th = (th + np.float32(ss * w) * np.asarray(eps, np.float32)).astype(np.float32)
    return th


def calc_best_score(score_add, score_sub, grad_eval_type="mean", mirrored_sampling=True):
    """Returns (score_best, sign) — sign = -1 when eps is inverted (sub strictly better)."""
    if grad_eval_type == "mean":
        sub, add = statistics.mean(score_sub), statistics.mean(score_add)
    elif grad_eval_type == "minmax":
        sub, add = min(score_sub), min(score_add)
    else:
        raise NotImplementedError("Unknown parameter for grad_eval_type: " + str(grad_eval_type))
    if mirrored_sampling:
        return max(add, sub), (-1.0 if sub > add else 1.0)
    return add, 1.0


def noise(seed, generation, member, P, noise_std):
    return (philox.normals(seed, generation, member, P) * np.float32(noise_std)).astype(np.float32)
