"""Agent side of the drop-in API: DDQN / DDQN_vary and select_agent with the reference's surface

    agents/base_agent.py:64-227  train(env, test_env=None, time_remaining=1e9) / test(env, time_remaining=1e9)
                                 -> (reward_list, episode_length_list, replay_buffer)
    agents/DDQN.py:14-120        learn / select_train_action / select_test_action / update_parameters_per_episode /
                                 reset_optimizer, attributes model, model_target, eps, it, full_config
    agents/DDQN_vary.py:26-59    sampled lr / batch_size / hidden_size / hidden_layer
    agents/agent_utils.py:15-66  select_agent(config, agent_name)

train()/test() hand the WHOLE loop (acting, env step, replay append, TD update, per-episode test, early-out) to the
fused persistent kernel as one lane; learn()/select_*_action() are the step-by-step forms on the unit kernels.
The Q-net parameters, the target net and Adam's moments are flat CUDA tensors in torch's layout; `model` /
`model_target` are nn.Modules whose parameters VIEW those tensors, so state_dict() matches the reference's keys.
"""
import copy
import math
import random

import numpy as np
import torch
import torch.nn as nn

from . import config as le_config
from . import ops
from ._abi import ACT_IDS, ENV_REAL, Q_DQN, Q_DUELING, REAL_ENV_IDS, LaneCfg
from .envs import EnvFactory, EnvWrapper, _cuda_device, build_nn_from_config
from .rng import lane_keys
from .utils import ReplayBuffer

#: training env steps per second assumed when a wall-clock `time_remaining` is mapped onto the deterministic
#: step budget of the kernel (the reference's time budget is wall-clock on one CPU core: ~800 steps/s).
STEPS_PER_SECOND = 800.0


class Critic_DQN(nn.Module):
    """models/actor_critic.py:84-91 with parameters that are views into one flat tensor."""

    def __init__(self, state_dim, action_dim, agent_name, config, flat=None):
        super().__init__()
        self.net = build_nn_from_config(input_dim=state_dim, output_dim=action_dim, nn_config=config["agents"][agent_name])
        if flat is not None:
            self.bind(flat)

    def bind(self, flat):
        """Re-point every nn.Linear parameter at a slice of `flat` (torch order: W1, b1, W2, b2, ...)."""
        _bind_linear_params(self, flat)

    def forward(self, state):
        return self.net(state)


def _bind_linear_params(module, flat):
    """Re-point every nn.Linear parameter of `module` (in module order = state_dict order) at a slice of `flat`."""
    off = 0
    for l in module.modules():
        if isinstance(l, nn.Linear):
            for name in ("weight", "bias"):
                p = getattr(l, name)
                n = p.numel()
                setattr(l, name, nn.Parameter(flat[off:off + n].view(p.shape), requires_grad=False))
                off += n
    assert off == flat.numel()


class Critic_DuelingDQN(nn.Module):
    """models/actor_critic.py:94-122: feature stream -> value head + advantage head, q = V + (A - A.mean())."""

    def __init__(self, state_dim, action_dim, agent_name, config, flat=None):
        super().__init__()
        c = config["agents"][agent_name]
        self.feature_stream = build_nn_from_config(input_dim=state_dim, output_dim=c["feature_dim"], nn_config=c)
        heads = copy.copy(c)
        heads["hidden_layer"] = 1
        heads["hidden_size"] = c["feature_dim"]
        self.value_stream = build_nn_from_config(input_dim=c["feature_dim"], output_dim=1, nn_config=heads)
        self.advantage_stream = build_nn_from_config(input_dim=c["feature_dim"], output_dim=action_dim, nn_config=heads)
        if flat is not None:
            self.bind(flat)

    def bind(self, flat):
        _bind_linear_params(self, flat)

    def forward(self, state):
        features = self.feature_stream(state)
        values = self.value_stream(features)
        advantages = self.advantage_stream(features)
        return values + (advantages - advantages.mean())


class BaseAgent(nn.Module):
    def __init__(self, agent_name, env, config):
        super().__init__()
        self.state_dim = env.get_state_dim()
        self.action_dim = env.get_action_dim()
        agent_config = config["agents"][agent_name]
        self.train_episodes = agent_config["train_episodes"]
        self.test_episodes = agent_config["test_episodes"]
        self.init_episodes = agent_config["init_episodes"]
        self.rb_size = agent_config["rb_size"]
        self.same_action_num = agent_config["same_action_num"]
        self.print_rate = agent_config["print_rate"]
        self.early_out_num = agent_config["early_out_num"]
        self.early_out_virtual_diff = agent_config["early_out_virtual_diff"]
        self.render_env = config["render_env"]
        self.device = config["device"]


class DDQN(BaseAgent):
    _AGENT_NAME = "ddqn"
    _CRITIC = Critic_DQN
    _Q_KIND = Q_DQN

    def __init__(self, env, config, icm=False):
        self.agent_name = self._AGENT_NAME
        super().__init__(agent_name=self.agent_name, env=env, config=config)
        if icm:
            raise NotImplementedError("the ICM baseline branch (models/icm_baseline.py) is outside the hot path")
        c = config["agents"][self.agent_name]
        self.full_config = config
        self.batch_size = c["batch_size"]
        self.rb_size = c["rb_size"]
        self.gamma = c["gamma"]
        self.lr = c["lr"]
        self.tau = c["tau"]
        self.eps = c["eps_init"]
        self.eps_init = c["eps_init"]
        self.eps_min = c["eps_min"]
        self.eps_decay = c["eps_decay"]
        self.hidden_size = int(c["hidden_size"])
        self.hidden_layer = max(int(c.get("hidden_layer", 1)), 1)
        self.feature_dim = int(c.get("feature_dim", 0)) if self._Q_KIND == Q_DUELING else 0
        if self.hidden_layer > 3:   # DDQN_vary / DuelingDDQN_vary sample at most yaml hidden_layer + 1 = 3 (agents/DDQN_vary.py:52-57)
            raise NotImplementedError("Q-nets with hidden_layer > 3 are outside the compiled kernel set")
        self._act_id = ACT_IDS[str(c["activation_fn"])]
        self._env_name = config["env_name"]
        dev = _cuda_device()
        # torch-default initialised nets (models/model_utils.py:31,38), then flattened onto the device
        init = self._CRITIC(self.state_dim, self.action_dim, self.agent_name, config)
        from .envs import linear_theta
        flat = linear_theta(init)
        self._theta = flat.to(dev).contiguous()
        self._target = self._theta.clone()
        self.model = self._CRITIC(self.state_dim, self.action_dim, self.agent_name, config, flat=self._theta)
        self.model_target = self._CRITIC(self.state_dim, self.action_dim, self.agent_name, config, flat=self._target)
        self.reset_optimizer()
        self.it = 0
        self.icm = None
        self._seed = random.getrandbits(32)
        self._runs = 0
        self.step_budget = 0   # cap on training env steps per train() call (0: none)

    # ---------------------------------------------------------------------------------------------------
    def _lane_cfg(self, env, test_env, train_episodes, final_test, time_remaining):
        kind, theta, fields = env.kernel_env() if isinstance(env, EnvWrapper) else (ENV_REAL, None, {})
        c = LaneCfg()
        c.sd, c.ad = self.state_dim, self.action_dim
        c.env_kind = kind
        c.real_env = REAL_ENV_IDS[self._env_name]
        c.env_hidden = int(fields.get("env_hidden", 0))
        c.env_act = int(fields.get("env_act", ACT_IDS["identity"]))
        for i, sl in enumerate(fields.get("env_slope", [0.01] * 3)):
            c.env_slope[i] = sl
        c.rn_type = int(fields.get("rn_type", 0))
        c.q_hidden, c.q_act = self.hidden_size, self._act_id
        c.q_kind, c.q_layers, c.q_feature_dim = self._Q_KIND, self.hidden_layer, self.feature_dim
        c.batch_size, c.rb_size = int(self.batch_size), int(self.rb_size)
        c.train_episodes, c.test_episodes, c.init_episodes = int(train_episodes), int(self.test_episodes), int(self.init_episodes)
        c.max_steps = int(env.max_episode_steps())
        c.early_out_num = int(self.early_out_num)
        c.same_action_num = max(int(self.same_action_num), 1)
        c.use_test_env = 1 if test_env is not None else 0
        c.final_test = 1 if final_test else 0
        budget = int(self.step_budget)
        if time_remaining < 1e8:
            t_budget = max(int(time_remaining * STEPS_PER_SECOND), 0)
            budget = t_budget if budget == 0 else min(budget, t_budget)
            budget = max(budget, 1)
        c.step_budget = budget
        c.gamma, c.lr, c.tau = float(self.gamma), float(self.lr), float(self.tau)
        c.eps_init, c.eps_min, c.eps_decay = float(self.eps_init), float(self.eps_min), float(self.eps_decay)
        c.early_out_virtual_diff = float(self.early_out_virtual_diff)
        brk = test_env if test_env is not None else env
        c.solved_reward = float(brk.get_solved_reward())
        c.beta1, c.beta2, c.adam_eps = 0.9, 0.999, 1e-8
        return c, theta

    def _run_lane(self, cfg, theta):
        dev = self._theta.device
        bufs = ops.InnerLoopBuffers(cfg, 1, 1, dev, want_q_final=True)
        key = lane_keys(self._seed, self._runs, [0], [0], [0])
        self._runs += 1
        th = None if theta is None else theta.to(dev).contiguous()[None]
        ops.inner_loop_run(bufs, cfg, th, None, ops.keys_tensor(key, dev), q_init=self._theta[None].contiguous())
        torch.cuda.current_stream().synchronize()
        return bufs, bufs.results()[0]

    def _replay_from_ring(self, bufs, cfg, n_rows):
        """The kernel's HBM replay ring (slot 0) as the reference's ReplayBuffer object."""
        sd = cfg.sd
        rowf = 2 * sd + 4
        rb = ReplayBuffer(state_dim=sd, action_dim=1, device=self.device, max_size=int(cfg.rb_size))
        plan = ops.inner_loop_plan(cfg, 1, 1)
        n = int(min(n_rows, plan["ring_cap"]))
        if n <= 0:
            return rb
        import ctypes as C
        from ._abi import load_library
        off = ops.ring_offset_bytes(cfg, 1, 1)
        ring = bufs.workspace[off:off + n * rowf * 4].view(torch.float32).reshape(n, rowf).cpu()
        tail = (sd % 4) == 0
        off_a = 2 * sd if tail else sd
        off_s2 = sd if tail else sd + 2
        rb._alloc(max(n, 1))
        rb.state[:n] = ring[:, 0:sd]
        rb.action[:n, 0] = ring[:, off_a].contiguous().view(torch.int32).float()   # the ring stores the action's int32 bits
        rb.next_state[:n] = ring[:, off_s2:off_s2 + sd]
        rb.reward[:n, 0] = ring[:, off_a + 1]
        rb.done[:n, 0] = ring[:, 2 * sd + 2]
        rb.size = n
        rb.ptr = n % rb.max_size
        return rb

    def train(self, env, test_env=None, time_remaining=1e9):
        """agents/base_agent.py:64-153 as one lane of the fused kernel. The agent's current online weights are the
        initial weights; the target net starts as a copy and Adam's state starts at zero (a fresh agent per
        calc_score is the reference's usage, agents/GTN_worker.py:190)."""
        env.set_agent_params(same_action_num=self.same_action_num, gamma=self.gamma)
        cfg, theta = self._lane_cfg(env, test_env, self.train_episodes, final_test=False, time_remaining=time_remaining)
        bufs, out = self._run_lane(cfg, theta)
        n = int(out["n_episodes"])
        rewards = bufs.rewards[0, :n].cpu().tolist()
        lengths = bufs.lengths[0, :n].cpu().tolist()
        if int(out["timed_out"]):  # time_is_up (agents/base_agent.py:30-47): pad with the worst reward / longest episode
            print("timeout")
            if len(rewards) == 0:
                rewards.append(-1e9)
            while len(rewards) < self.train_episodes:
                rewards.append(min(rewards))
            if len(lengths) == 0:
                lengths.append(1e9)
            while len(lengths) < self.train_episodes:
                lengths.append(max(lengths))
        self._theta.copy_(bufs.q_final[0])
        # The kernel returns the trained ONLINE net only.  The reference keeps target net / Adam moments / eps on the agent
        # object across calls; a fresh agent per calc_score is its only usage on this path (agents/GTN_worker.py:190), so the
        # step-by-step API continues from a re-synchronised state: target <- online (the Polyak lag of tau per step is not
        # carried over), Adam moments zero, eps advanced to the value after the episodes just run (agents/DDQN.py:112-117).
        if hasattr(self, "_target"):
            self._target.copy_(self._theta)
        if n > 0 and hasattr(self, "eps"):
            self.eps = max(float(self.eps_init) * float(self.eps_decay) ** (n - 1), float(self.eps_min))
        self.it += int(out["learn_iters"])
        self.last_run = dict(out=out, cfg=cfg)
        rb = self._replay_from_ring(bufs, cfg, int(out["train_steps"]))
        env.close()
        return rewards, lengths, rb

    def test(self, env, time_remaining=1e9):
        """agents/base_agent.py:155-227: test_episodes greedy rollouts on the (real) env inside the kernel."""
        if isinstance(env, EnvWrapper) and env.is_virtual_env():
            raise NotImplementedError("test() on a synthetic env is outside the hot path (the reference tests on the real env)")
        env.set_agent_params(same_action_num=self.same_action_num, gamma=self.gamma)
        cfg, _ = self._lane_cfg(env, None, 0, final_test=True, time_remaining=1e9)
        cfg.env_kind = ENV_REAL
        bufs, out = self._run_lane(cfg, None)
        rewards = bufs.test_rewards[0].cpu().tolist()
        lengths = bufs.test_lengths[0].cpu().tolist()
        env.close()
        return rewards, lengths, ReplayBuffer(state_dim=self.state_dim, action_dim=1, device=self.device, max_size=int(1e6))

    # ---- step-by-step API (unit kernels) ------------------------------------------------------------------
    def _unit_cfg(self):
        c = LaneCfg()
        c.sd, c.ad = self.state_dim, self.action_dim
        c.real_env = REAL_ENV_IDS[self._env_name]
        c.q_hidden, c.q_act = self.hidden_size, self._act_id
        c.q_kind, c.q_layers, c.q_feature_dim = self._Q_KIND, self.hidden_layer, self.feature_dim
        c.batch_size = int(self.batch_size)
        c.gamma, c.lr, c.tau = float(self.gamma), float(self.lr), float(self.tau)
        c.beta1, c.beta2, c.adam_eps = 0.9, 0.999, 1e-8
        return c

    def learn(self, replay_buffer, env, episode):
        self.it += 1
        states, actions, next_states, rewards, dones = replay_buffer.sample(self.batch_size)
        rows = torch.cat([states.reshape(self.batch_size, -1).float(), actions.reshape(self.batch_size, 1).float(),
                          next_states.reshape(self.batch_size, -1).float(), rewards.reshape(self.batch_size, 1).float(),
                          dones.reshape(self.batch_size, 1).float()], dim=1)
        dev = self._theta.device
        loss = ops.td_update(self._unit_cfg(), self._theta[None], self._target[None], self._adam_m[None], self._adam_v[None],
                             self._adam_t, rows.to(dev).contiguous()[None])
        return loss[0]

    def _greedy(self, state):
        dev = self._theta.device
        s = torch.as_tensor(state, dtype=torch.float32).reshape(1, -1).to(dev)
        _, am = ops.qnet_forward(self._unit_cfg(), self._theta[None], s)
        return am.to(torch.int64).cpu()

    def select_train_action(self, state, env, episode):
        if random.random() < self.eps:
            return env.get_random_action()
        return self._greedy(state)

    def select_test_action(self, state, env):
        return self._greedy(state)

    def update_parameters_per_episode(self, episode):
        if episode == 0:
            self.eps = self.eps_init
        else:
            self.eps *= self.eps_decay
            self.eps = max(self.eps, self.eps_min)

    def reset_optimizer(self):
        self._adam_m = torch.zeros_like(self._theta)
        self._adam_v = torch.zeros_like(self._theta)
        self._adam_t = torch.zeros(1, dtype=torch.int32, device=self._theta.device)


def vary_hyperparameters(agent_cfg, rng):
    """Sampling distribution of agents/DDQN_vary.py:26-59 (ConfigSpace 0.4.13 semantics): lr log-uniform on
    [lr/3, 3 lr]; batch_size and hidden_size log-uniform integers on [int(x/3), int(3x)]; hidden_layer uniform on
    {L-1, L, L+1}.  `rng`: numpy RandomState. Returns a modified copy."""
    out = dict(agent_cfg)

    def log_int(lo, hi):
        lo_f, hi_f = math.log(lo - 0.49999), math.log(hi + 0.49999)
        return int(min(max(int(round(math.exp(lo_f + (hi_f - lo_f) * rng.random_sample()))), lo), hi))
    lr, bs, hs, hl = agent_cfg["lr"], agent_cfg["batch_size"], agent_cfg["hidden_size"], agent_cfg["hidden_layer"]
    out["lr"] = float(math.exp(math.log(lr / 3) + (math.log(lr * 3) - math.log(lr / 3)) * rng.random_sample()))
    out["batch_size"] = log_int(int(bs / 3), int(bs * 3))
    out["hidden_size"] = log_int(int(hs / 3), int(hs * 3))
    out["hidden_layer"] = int(rng.randint(hl - 1, hl + 2))
    return out


class DuelingDDQN(DDQN):
    """agents/DuelingDDQN.py: the same algorithm on Critic_DuelingDQN (general CTA-per-lane kernel)."""
    _AGENT_NAME = "duelingddqn"
    _CRITIC = Critic_DuelingDQN
    _Q_KIND = Q_DUELING


def _varied_config(config, agent_name, vary_name, rng, vary_feature_dim=False):
    if not config["agents"][vary_name]["vary_hp"]:
        return config
    config_mod = copy.deepcopy(config)
    a = vary_hyperparameters(config_mod["agents"][agent_name], rng)
    if vary_feature_dim:   # agents/DuelingDDQN_vary.py:24-75 also samples feature_dim log-uniformly on [fd/3, 3 fd]
        fd = config_mod["agents"][agent_name]["feature_dim"]
        lo, hi = math.log(int(fd / 3) - 0.49999), math.log(int(fd * 3) + 0.49999)
        a["feature_dim"] = int(min(max(int(round(math.exp(lo + (hi - lo) * rng.random_sample()))), int(fd / 3)), int(fd * 3)))
    # hidden_layer 0 and 1 build the same net (models/model_utils.py:34 loops hidden_layer-1 times)
    a["hidden_layer"] = max(a["hidden_layer"], 1)
    config_mod["agents"][agent_name] = a
    return config_mod


class DDQN_vary(DDQN):
    """agents/DDQN_vary.py: DDQN with sampled lr / batch_size / hidden_size / hidden_layer."""
    _rng = np.random.RandomState()

    def __init__(self, env, config, icm=False):
        config_mod = _varied_config(config, "ddqn", "ddqn_vary", self._rng)
        print("full config: ", config_mod['agents']["ddqn"])
        super().__init__(env=env, config=config_mod, icm=icm)


class DuelingDDQN_vary(DuelingDDQN):
    """agents/DuelingDDQN_vary.py."""
    _rng = np.random.RandomState()

    def __init__(self, env, config, icm=False):
        config_mod = _varied_config(config, "duelingddqn", "duelingddqn_vary", self._rng, vary_feature_dim=True)
        print("full config: ", config_mod['agents']["duelingddqn"])
        super().__init__(env=env, config=config_mod, icm=icm)


class Actor_TD3_discrete(nn.Module):
    """models/actor_critic.py:22-36: MLP * max_action -> F.gumbel_softmax(tau, hard)."""

    def __init__(self, state_dim, action_dim, max_action, agent_name, config):
        super().__init__()
        c = config["agents"][agent_name]
        self.net = build_nn_from_config(input_dim=state_dim, output_dim=action_dim, nn_config=c)
        self.max_action = max_action
        self.gumbel_softmax_temp = c["gumbel_softmax_temp"]
        self.gumbel_softmax_hard = c["gumbel_softmax_hard"]

    def forward(self, state, tau):
        return torch.nn.functional.gumbel_softmax(self.net(state) * self.max_action, tau=tau, hard=self.gumbel_softmax_hard)


class Critic_Q(nn.Module):
    """models/actor_critic.py:69-76."""

    def __init__(self, state_dim, action_dim, agent_name, config):
        super().__init__()
        self.net = build_nn_from_config(input_dim=state_dim + action_dim, output_dim=1, nn_config=config["agents"][agent_name])

    def forward(self, state, action):
        return self.net(torch.cat([state, action], dim=len(state.shape) - 1))


class TD3_discrete_vary(BaseAgent):
    """agents/TD3_discrete_vary.py: TD3 with a Gumbel-softmax actor for discrete action spaces.  train() / test() run as
    one lane of the TD3 kernel (le_td3_run_host); the nets below hold the torch-default initial weights and receive the
    trained actor back."""

    def __init__(self, env, min_action, max_action, config):
        self.agent_name = 'td3_discrete_vary'
        if config["agents"][self.agent_name]["vary_hp"]:
            config = copy.deepcopy(config)
            a = vary_hyperparameters(config["agents"][self.agent_name], _VARY_RNG)
            a["hidden_layer"] = max(a["hidden_layer"], 1)
            config["agents"][self.agent_name] = a
        super().__init__(agent_name=self.agent_name, env=env, config=config)
        c = config["agents"][self.agent_name]
        self.full_config = config
        self.max_action, self.min_action = max_action, min_action
        for k in ("batch_size", "rb_size", "gamma", "tau", "policy_delay", "lr", "action_std", "policy_std", "policy_std_clip"):
            setattr(self, k, c[k])
        if str(c["activation_fn"]) not in ("tanh", "relu", "leakyrelu") or int(c["hidden_layer"]) > 3:
            raise NotImplementedError("TD3 nets with activation %r / hidden_layer %r are outside the compiled kernel set"
                                      % (c["activation_fn"], c["hidden_layer"]))
        _cuda_device()
        self.actor = Actor_TD3_discrete(self.state_dim, self.action_dim, max_action, self.agent_name, config)
        self.actor_target = copy.deepcopy(self.actor)
        self.critic_1 = Critic_Q(self.state_dim, self.action_dim, self.agent_name, config)
        self.critic_2 = Critic_Q(self.state_dim, self.action_dim, self.agent_name, config)
        self.critic_target_1, self.critic_target_2 = copy.deepcopy(self.critic_1), copy.deepcopy(self.critic_2)
        self.total_it = 0
        self.gumbel_temp_anneal_steps = np.linspace(self.actor.gumbel_softmax_temp, self.actor.gumbel_softmax_temp / 20, 2000)
        self.gumbel_temp_annealed = self.gumbel_temp_anneal_steps[0]
        self._env_name = config["env_name"]
        self._cfg_section = c
        self._seed = random.getrandbits(32)
        self._runs = 0
        self.step_budget = 0

    def _td3_cfg(self, env, test_env, train_episodes, final_test, time_remaining):
        from ._abi import Td3Cfg
        from .envs import linear_theta
        kind, theta, fields = env.kernel_env() if isinstance(env, EnvWrapper) else (ENV_REAL, None, {})
        t = Td3Cfg()
        c, a = t.base, self._cfg_section
        c.sd, c.ad, c.env_kind, c.real_env = self.state_dim, self.action_dim, kind, REAL_ENV_IDS[self._env_name]
        c.env_hidden, c.env_act = int(fields.get("env_hidden", 0)), int(fields.get("env_act", ACT_IDS["identity"]))
        for i, sl in enumerate(fields.get("env_slope", [0.01] * 3)):
            c.env_slope[i] = sl
        c.q_hidden, c.q_layers, c.q_act = int(a["hidden_size"]), max(int(a["hidden_layer"]), 1), ACT_IDS[str(a["activation_fn"])]
        c.batch_size, c.rb_size = int(self.batch_size), int(self.rb_size)
        c.train_episodes, c.test_episodes, c.init_episodes = int(train_episodes), int(self.test_episodes), int(self.init_episodes)
        c.max_steps, c.early_out_num = int(env.max_episode_steps()), int(self.early_out_num)
        c.same_action_num = max(int(self.same_action_num), 1)
        c.use_test_env, c.final_test = (1 if test_env is not None else 0), (1 if final_test else 0)
        budget = int(self.step_budget)
        if time_remaining < 1e8:
            tb = max(int(time_remaining * STEPS_PER_SECOND), 1)
            budget = tb if budget == 0 else min(budget, tb)
        c.step_budget = budget
        c.gamma, c.lr, c.tau = float(self.gamma), float(self.lr), float(self.tau)
        c.early_out_virtual_diff = float(self.early_out_virtual_diff)
        c.solved_reward = float((test_env if test_env is not None else env).get_solved_reward())
        c.beta1, c.beta2, c.adam_eps = 0.9, 0.999, 1e-8
        t.policy_delay, t.gumbel_hard = int(self.policy_delay), int(bool(self.actor.gumbel_softmax_hard))
        t.action_std, t.policy_std, t.policy_std_clip = float(self.action_std), float(self.policy_std), float(self.policy_std_clip)
        t.gumbel_temp, t.max_action = float(self.actor.gumbel_softmax_temp), float(self.max_action)
        nets = [linear_theta(m).numpy() for m in (self.actor, self.critic_1, self.critic_2)]
        return t, (None if theta is None else theta.numpy()), nets

    def _run(self, t, theta, nets):
        key = lane_keys(self._seed, self._runs, [0], [0], [0])
        self._runs += 1
        return ops.td3_run_host(t, theta, None, key, nets[0], nets[1], nets[2], device=torch.cuda.current_device())

    def train(self, env, test_env=None, time_remaining=1e9):
        """agents/base_agent.py:64-153 with the TD3 act / learn as one lane of the TD3 kernel (fresh targets / optimizers)."""
        from .envs import set_linear_theta
        env.set_agent_params(same_action_num=self.same_action_num, gamma=self.gamma)
        t, theta, nets = self._td3_cfg(env, test_env, self.train_episodes, False, time_remaining)
        res = self._run(t, theta, nets)
        out = res["out"][0]
        n = int(out["n_episodes"])
        rewards, lengths = res["rewards"][0, :n].tolist(), res["lengths"][0, :n].tolist()
        if int(out["timed_out"]):
            print("timeout")
            rewards = rewards or [-1e9]
            lengths = lengths or [1e9]
            rewards += [min(rewards)] * (self.train_episodes - len(rewards))
            lengths += [max(lengths)] * (self.train_episodes - len(lengths))
        set_linear_theta(self.actor, torch.from_numpy(res["actor_final"][0]))
        self.total_it += int(out["learn_iters"])
        if self.total_it > 0:
            self.gumbel_temp_annealed = self.gumbel_temp_anneal_steps[min(self.total_it, 2000) - 1]
        env.close()
        return rewards, lengths, ReplayBuffer(state_dim=self.state_dim, action_dim=self.action_dim, device=self.device,
                                              max_size=int(self.rb_size))

    def test(self, env, time_remaining=1e9):
        """agents/base_agent.py:155-227 with select_test_action (:164-166) at the current (annealed) Gumbel temperature."""
        if isinstance(env, EnvWrapper) and env.is_virtual_env():
            raise NotImplementedError("test() on a synthetic env is outside the hot path (the reference tests on the real env)")
        env.set_agent_params(same_action_num=self.same_action_num, gamma=self.gamma)
        t, _, nets = self._td3_cfg(env, None, 0, True, 1e9)
        t.base.env_kind = ENV_REAL
        t.gumbel_temp = float(self.gumbel_temp_annealed)
        res = self._run(t, None, nets)
        env.close()
        return res["test_rewards"][0].tolist(), [], ReplayBuffer(state_dim=self.state_dim, action_dim=self.action_dim,
                                                                  device=self.device, max_size=int(1e6))


_VARY_RNG = np.random.RandomState()

_OUTSIDE_HOT_PATH = {"td3", "td3_icm", "td3_vary", "td3_icm_vary", "ppo", "ppo_icm", "duelingddqn_icm",
                     "duelingddqn_icm_vary", "ql", "ql_cb", "sarsa", "sarsa_cb", "ddqn_icm", "ddqn_icm_vary"}


def select_agent(config, agent_name):
    """agents/agent_utils.py:15-66."""
    env_factory = EnvFactory(config)
    dummy_env = env_factory.generate_real_env(print_str='Select Agent: ')
    agent_name = agent_name.lower()
    if agent_name == "ddqn":
        return DDQN(env=dummy_env, config=config)
    if agent_name == "ddqn_vary":
        return DDQN_vary(env=dummy_env, config=config)
    if agent_name == "duelingddqn":
        return DuelingDDQN(env=dummy_env, config=config)
    if agent_name == "duelingddqn_vary":
        return DuelingDDQN_vary(env=dummy_env, config=config)
    if agent_name == "td3_discrete_vary":
        return TD3_discrete_vary(env=dummy_env, config=config, min_action=dummy_env.get_min_action(), max_action=dummy_env.get_max_action())
    if agent_name in _OUTSIDE_HOT_PATH:
        raise NotImplementedError("RL agent %r is outside the B200 hot path (DDQN / DuelingDDQN / TD3_discrete_vary families are built)" % agent_name)
    raise NotImplementedError("Unknownn RL agent")
