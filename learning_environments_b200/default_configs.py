"""Hyper-parameters of the BASELINE configurations as nested dicts in the reference's config schema
(same keys as the YAML files the reference loads with yaml.safe_load; values cited from
/root/reference/default_config_cartpole_syn_env.yaml, default_config_acrobot_syn_env.yaml,
default_config_cartpole_reward_env.yaml).  A user's own YAML dict is accepted unchanged everywhere these are.
"""
import copy


def _gtn(**kw):
    base = dict(mode="multi", max_iterations=200, num_threads_per_worker=1, num_workers=16, nes_step_size=False,
                mirrored_sampling=True, num_grad_evals=1, grad_eval_type="mean", weight_decay=0.0, time_mult=3.0,
                time_max=600.0, time_sleep_master=0.2, time_sleep_worker=2.0, score_transform_type=3,
                quit_when_solved=False, synthetic_env_type=0, unsolved_weight=10000.0, agent_name="DDQN")
    base.update(kw)
    return base


_CARTPOLE_SE = {
    "env_name": "CartPole-v0", "device": "cpu", "render_env": False,
    "agents": {
        "gtn": _gtn(noise_std=0.0124, step_size=0.148),
        "ddqn": dict(train_episodes=1000, test_episodes=10, init_episodes=1, batch_size=199, gamma=0.988, lr=0.000304,
                     tau=0.00848, eps_init=0.809, eps_min=0.0371, eps_decay=0.961, rb_size=100000, same_action_num=1,
                     activation_fn="tanh", hidden_size=57, hidden_layer=1, print_rate=10, early_out_num=10,
                     early_out_virtual_diff=0.01),
        "ddqn_vary": dict(vary_hp=True),
        # default_config_cartpole_syn_env.yaml:58-77
        "duelingddqn": dict(train_episodes=1000, test_episodes=10, init_episodes=1, batch_size=193, gamma=0.961, lr=0.0091437,
                            tau=0.07348, eps_init=0.906, eps_min=0.00645, eps_decay=0.8267, rb_size=100000, same_action_num=1,
                            activation_fn="tanh", hidden_size=61, hidden_layer=1, feature_dim=60, print_rate=1, early_out_num=1,
                            early_out_virtual_diff=0.01),
        "duelingddqn_vary": dict(vary_hp=True),
        # default_config_cartpole_syn_env.yaml:79-101
        "td3_discrete_vary": dict(train_episodes=1000, test_episodes=10, init_episodes=1, batch_size=122, gamma=0.9989, lr=0.0017496,
                               tau=0.0724303, policy_delay=1, rb_size=1000000, same_action_num=1, activation_fn="tanh", hidden_size=510,
                               hidden_layer=2, action_std=0.037275, policy_std=0.2225286, policy_std_clip=0.5, print_rate=1,
                               early_out_num=1, early_out_virtual_diff=0.01, gumbel_softmax_temp=2.3076235, gumbel_softmax_hard=True,
                               vary_hp=False),
    },
    "envs": {"CartPole-v0": dict(solved_reward=195.0, max_steps=200, activation_fn="leakyrelu", hidden_size=83,
                                 hidden_layer=1, info_dim=0, reward_env_type=0)},
}

_ACROBOT_SE = {
    "env_name": "Acrobot-v1", "device": "cpu", "render_env": False,
    "agents": {
        "gtn": _gtn(noise_std=0.0114, step_size=0.727),
        "ddqn": dict(train_episodes=1000, test_episodes=10, init_episodes=20, batch_size=149, gamma=0.991, lr=0.00222,
                     tau=0.0209, eps_init=0.904, eps_min=0.0471, eps_decay=0.899, rb_size=100000, same_action_num=1,
                     activation_fn="leakyrelu", hidden_size=112, hidden_layer=1, print_rate=1, early_out_num=10,
                     early_out_virtual_diff=0.01),
        "ddqn_vary": dict(vary_hp=True),
        # default_config_acrobot.yaml:61-80 (the DuelingDDQN section the transfer evaluations use)
        "duelingddqn": dict(train_episodes=1000, test_episodes=10, init_episodes=10, batch_size=128, gamma=0.99, lr=1e-3, tau=0.01,
                            eps_init=1.0, eps_min=0.01, eps_decay=0.9, rb_size=100000, same_action_num=1, activation_fn="relu",
                            hidden_size=128, hidden_layer=2, feature_dim=128, print_rate=1, early_out_num=10,
                            early_out_virtual_diff=0.01),
        "duelingddqn_vary": dict(vary_hp=True),
        # default_config_acrobot_syn_env.yaml:58-80
        "td3_discrete_vary": dict(train_episodes=1000, test_episodes=10, init_episodes=1, batch_size=122, gamma=0.9989, lr=0.0017496,
                               tau=0.0724303, policy_delay=1, rb_size=1000000, same_action_num=1, activation_fn="tanh", hidden_size=510,
                               hidden_layer=2, action_std=0.037275, policy_std=0.2225286, policy_std_clip=0.5, print_rate=1,
                               early_out_num=1, early_out_virtual_diff=0.01, gumbel_softmax_temp=2.3076235, gumbel_softmax_hard=True,
                               vary_hp=False),
    },
    "envs": {"Acrobot-v1": dict(solved_reward=-100.0, max_steps=500, activation_fn="prelu", hidden_size=167,
                                hidden_layer=1, info_dim=0, reward_env_type=0)},
}

_CARTPOLE_RN = {
    "env_name": "CartPole-v0", "device": "cpu", "render_env": False,
    "agents": {
        "gtn": _gtn(max_iterations=50, noise_std=0.1, step_size=0.5, time_max=3600.0, quit_when_solved=True,
                    synthetic_env_type=1, unsolved_weight=100.0),
        "ddqn": dict(train_episodes=100, test_episodes=1, init_episodes=1, batch_size=192, gamma=0.99, lr=0.003, tau=0.01,
                     eps_init=0.8, eps_min=0.03, eps_decay=0.95, rb_size=1000000, same_action_num=1,
                     activation_fn="leakyrelu", hidden_size=64, hidden_layer=1, print_rate=10, early_out_num=10,
                     early_out_virtual_diff=0.02),
        "ddqn_vary": dict(vary_hp=True),
    },
    "envs": {"CartPole-v0": dict(solved_reward=195.0, max_steps=200, activation_fn="prelu", hidden_size=64,
                                 hidden_layer=1, info_dim=0, reward_env_type=2)},
}

_CONFIGS = {"cartpole_syn_env": _CARTPOLE_SE, "acrobot_syn_env": _ACROBOT_SE, "cartpole_reward_env": _CARTPOLE_RN}


def get(name):
    """name in {'cartpole_syn_env', 'acrobot_syn_env', 'cartpole_reward_env'} -> a fresh deep copy."""
    return copy.deepcopy(_CONFIGS[name])


def names():
    return sorted(_CONFIGS)
