"""Helper process of tests/test_host_logic.py: vary_hp.evaluate_agents on CPU with an oracle-backed launch function,
alone or as one rank of a gloo group (agents block-sharded over the ranks, one all-reduce assembles the table)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from learning_environments_b200 import default_configs, vary_hp  # noqa: E402
from oracle import c_oracle  # noqa: E402

CALLS = []


def oracle_run_group(sub, cfg0, theta, env_index, keys, n_env, device):
    k = len(sub)
    tr = np.zeros((k, max(c.test_episodes for c in sub)), np.float64)
    st, ep = np.zeros(k, np.int64), np.zeros(k, np.int64)
    for j, c in enumerate(sub):
        th = None if theta is None else theta[0 if env_index is None else int(env_index[j])]
        res = c_oracle.run_lane(c, th, (int(keys[j][0]), int(keys[j][1])))
        tr[j, :c.test_episodes] = res["test_rewards"]
        st[j], ep[j] = res["train_steps"], res["n_episodes"]
    CALLS.append(k)
    return tr, st, ep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    vary_hp._run_group = oracle_run_group
    cfg = default_configs.get("cartpole_syn_env")
    rng = np.random.RandomState(0)
    P = 2247
    thetas = (rng.uniform(-1, 1, size=(2, P)) * 0.3).astype(np.float32)
    over = dict(print_rate=10, early_out_num=2, train_episodes=3, init_episodes=1, test_episodes=2, early_out_virtual_diff=0.01,
                batch_size=32)
    r, s, e, cfgs = vary_hp.evaluate_agents(cfg, thetas, agents_num=3, seed=9, overrides=over, device="cpu")
    res = dict(rewards=r, steps=s, episodes=e, lanes_run_here=int(sum(CALLS)), hidden=[c.q_hidden for c in cfgs])
    rank = int(os.environ.get("RANK", "0"))
    with open(a.out + (".rank%d.json" % rank if world > 1 else ""), "w") as f:
        json.dump(res, f)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
