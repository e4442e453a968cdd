"""Batched vary_hp evaluation (BASELINE config 4): many DDQN agents with sampled hyper-parameters trained on ONE fixed
Synthetic Environment and tested on the real env.

Mirrors experiments/syn_env_evaluate_cartpole_vary_hp_2.py:25-48 (`train_test_agents`): the overrides
init_episodes=10, train_episodes=1000, test_episodes=10, early_out_num=10, early_out_virtual_diff=0.01; per agent
`select_agent('DDQN_vary')` -> `agent.train(env=SE)` (no test_env: virtual-env plateau early-out) -> `agent.test(real_env)`;
returns (reward_list, train_steps_needed, episodes_needed).  Here every agent is one lane: agents whose sampled Q-net
fits the register kernel set (hidden_layer <= 1, hidden_size <= 128) run in one launch, the others (two hidden layers
or wider) in a second launch of the general kernel; lr / batch_size / hidden_size / hidden_layer travel per lane.
"""
import copy

import numpy as np
import torch

from . import config as le_config
from . import ops
from ._abi import ENV_SE, LaneCfg
from .agents import vary_hyperparameters
from .rng import lane_keys

OVERRIDES = dict(print_rate=10, early_out_num=10, train_episodes=1000, init_episodes=10, test_episodes=10,
                 early_out_virtual_diff=0.01)


def sample_agent_cfgs(config, agents_num, rng, overrides=None, vary=True, env_slopes=None):
    """One le_lane_cfg per agent (DDQN_vary sampling, agents/DDQN_vary.py:26-59)."""
    cfgs = []
    base = copy.deepcopy(config)
    base["agents"]["ddqn"].update(OVERRIDES if overrides is None else overrides)
    for _ in range(agents_num):
        c = copy.deepcopy(base)
        if vary:
            a = vary_hyperparameters(c["agents"]["ddqn"], rng)
            a["hidden_layer"] = max(a["hidden_layer"], 1)
            c["agents"]["ddqn"] = a
        lc = le_config.lane_cfg(c, "ddqn", ENV_SE, use_test_env=False, final_test=True)
        if env_slopes is not None:
            for i in range(3):
                lc.env_slope[i] = env_slopes[i]
        cfgs.append(lc)
    return cfgs


def _max_cfg(cfgs):
    """The configuration that fixes strides / kernel set for a group: elementwise maxima of the shape fields."""
    m = cfgs[0].copy()
    m.q_hidden = max(c.q_hidden for c in cfgs)
    m.q_layers = max(c.q_layers for c in cfgs)
    m.batch_size = max(c.batch_size for c in cfgs)
    return m


def train_test_agents(config, se_theta, agents_num=10, seed=0, overrides=None, vary=True, env_slopes=None, device="cuda"):
    """Returns (reward_list [agents][test_episodes], train_steps_needed [agents], episodes_needed [agents], lane_cfgs)."""
    rng = np.random.RandomState(seed)
    cfgs = sample_agent_cfgs(config, agents_num, rng, overrides, vary, env_slopes)
    keys = lane_keys(seed, 0, np.arange(agents_num), np.zeros(agents_num, int), np.zeros(agents_num, int))
    theta = torch.as_tensor(np.asarray(se_theta, np.float32)).to(device).reshape(1, -1).contiguous()
    rewards = [None] * agents_num
    steps = [0] * agents_num
    episodes = [0] * agents_num
    groups = {True: [], False: []}
    for i, c in enumerate(cfgs):
        groups[c.q_is_register_resident()].append(i)
    for resident, idx in groups.items():
        if not idx:
            continue
        sub = [cfgs[i] for i in idx]
        cfg0 = _max_cfg(sub)
        bufs = ops.InnerLoopBuffers(cfg0, len(idx), 1, device, n_cfg=len(idx))
        ops.inner_loop_run(bufs, sub, theta, None, ops.keys_tensor(keys[idx], device), cfg0=cfg0)
        torch.cuda.current_stream().synchronize()
        out = bufs.results()
        tr = bufs.test_rewards.cpu().numpy()
        for k, i in enumerate(idx):
            rewards[i] = tr[k, :cfgs[i].test_episodes].tolist()
            steps[i] = int(out["train_steps"][k])
            episodes[i] = int(out["n_episodes"][k])
    return rewards, steps, episodes, cfgs
