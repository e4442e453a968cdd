"""Tiny shapes of the two persistent loop kernels for `compute-sanitizer --tool racecheck` (2 episodes x <= 24 steps, B = 24):
the fused warp-per-lane kernel (CartPole SE + DDQN, U = 2 and U = 4) and the CTA-per-lane general kernel (DuelingDDQN, FFMA and
tcgen05 instantiations)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from learning_environments_b200 import config, default_configs, ops  # noqa: E402
from learning_environments_b200._abi import ENV_SE  # noqa: E402
from oracle import philox  # noqa: E402


def run(agent, over, n_lanes, tc=False, mw=False):
    os.environ["LE_TC"] = "1" if tc else "0"
    os.environ["LE_MW"] = "1" if mw else "0"
    d = default_configs.get("cartpole_syn_env")
    d["agents"][agent].update(dict(dict(train_episodes=2, test_episodes=2, init_episodes=1, batch_size=24), **over))
    cfg = config.lane_cfg(d, agent, ENV_SE)
    cfg.max_steps = 24
    rng = np.random.RandomState(0)
    theta = (rng.uniform(-1, 1, size=(1, cfg.se_params())) * 0.2).astype(np.float32)
    keys = [philox.lane_key(3, 0, i, 0, 0) for i in range(n_lanes)]
    bufs = ops.InnerLoopBuffers(cfg, n_lanes, 1, "cuda")
    ops.inner_loop_run(bufs, cfg, torch.from_numpy(theta).cuda(), None, ops.keys_tensor(keys, "cuda"))
    torch.cuda.synchronize()
    res = bufs.results()
    print(agent, over, "tc" if tc else "", "mw" if mw else "", "lanes", n_lanes, "steps", int(res["train_steps"].sum()), "learn", int(res["learn_iters"].sum()))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "fused"):
        run("ddqn", {}, 6)
        run("ddqn", dict(hidden_size=100), 3)
    if which in ("all", "mw"):     # multi-warp lanes: named barriers, shared weight records, gradient exchange (3 passes of 32 rows)
        os.environ["LE_MWC"] = "0"
        run("ddqn", dict(batch_size=70), 3, mw=True)
        run("ddqn", dict(hidden_size=100, batch_size=70), 2, mw=True)
    if which in ("all", "mwc"):    # cluster lanes: cluster barriers, weight records / command replicated and gradients exchanged through DSMEM
        os.environ["LE_MWC"] = "1"
        run("ddqn", dict(batch_size=70), 3, mw=True)
        run("ddqn", dict(hidden_size=100, batch_size=70), 2, mw=True)
    if which in ("all", "general"):
        run("duelingddqn", dict(hidden_size=32, feature_dim=32), 2)
    if which in ("all", "tc"):
        run("duelingddqn", dict(hidden_size=64, feature_dim=64, batch_size=64), 2, tc=True)
