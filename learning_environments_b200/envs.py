"""Environment side of the drop-in API: EnvFactory, EnvWrapper, VirtualEnv, RewardEnv and the device-resident
real environments.  Class and method names, argument meaning and error behaviour follow the reference:

    envs/env_factory.py:10-93   EnvFactory(config).generate_real_env / generate_virtual_env / generate_reward_env
    envs/env_wrapper.py:9-167   EnvWrapper.step/reset/get_random_action/get_state_dim/...
    envs/virtual_env.py:7-54    VirtualEnv.step(action, state=None) / reset()
    envs/reward_env.py:7-149    RewardEnv reward types, set_agent_params(gamma)

All arithmetic runs in the CUDA extension (ops.se_forward / ops.rn_reward / ops.real_env_step); there is no CPU
implementation here.  These step-by-step calls exist for API parity (one kernel launch per call); the fast path is
agent.train(env, ...) which hands the whole loop to the fused kernel.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._abi import ACT_IDS, ENV_REAL, ENV_RN, ENV_SE, REAL_ENV_IDS
from .config import ENV_DIMS
from .utils import from_one_hot_encoding, to_one_hot_encoding


# ------------------------------------------------------------------------------------------------------------
# minimal gym.spaces look-alikes (gym 0.17.3 is not a dependency of this package)
class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.np_random = np.random.RandomState()

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box(object):
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.np_random = np.random.RandomState()

    def sample(self):
        return self.np_random.uniform(self.low, self.high).astype(self.low.dtype)


def build_nn_from_config(input_dim, output_dim, nn_config):
    """MLP topology of models/model_utils.py:4-39 (same module order, so state_dict keys match the reference)."""
    hidden_size = int(nn_config['hidden_size'])
    hidden_layer = int(nn_config['hidden_layer'])
    name = nn_config['activation_fn']
    acts = {'prelu': nn.PReLU, 'relu': nn.ReLU, 'leakyrelu': nn.LeakyReLU, 'tanh': nn.Tanh, 'identity': nn.Identity}
    if name not in acts:
        raise ValueError('Unknown activation function: %r' % (name,))
    act_fn = acts[name]()          # ONE shared module instance (a PReLU weight is shared by every site)
    norm = nn.LayerNorm(hidden_size) if nn_config.get("use_layer_norm", False) else nn.Identity()
    modules = [nn.Linear(input_dim, hidden_size), act_fn]
    for _ in range(hidden_layer - 1):
        modules += [nn.Linear(hidden_size, hidden_size), norm, act_fn]
    modules.append(nn.Linear(hidden_size, output_dim))
    return nn.Sequential(*modules)


def linear_theta(module):
    """Parameter vector 'theta' of include/le_b200.h: nn.Linear weights/biases in module order, float32."""
    parts = []
    for l in module.modules():
        if isinstance(l, nn.Linear):
            parts.append(l.weight.detach().reshape(-1))
            if l.bias is not None:
                parts.append(l.bias.detach().reshape(-1))
    return torch.cat(parts).float()


def set_linear_theta(module, theta):
    """Inverse of linear_theta: writes a flat parameter vector back into the module's nn.Linear layers."""
    off = 0
    theta = torch.as_tensor(theta).detach().float().cpu()
    with torch.no_grad():
        for l in module.modules():
            if isinstance(l, nn.Linear):
                n = l.weight.numel()
                l.weight.copy_(theta[off:off + n].reshape(l.weight.shape))
                off += n
                if l.bias is not None:
                    n = l.bias.numel()
                    l.bias.copy_(theta[off:off + n])
                    off += n
    assert off == theta.numel()


def _prelu_slopes(nets):
    out = []
    for net in nets:
        slope = 0.25
        for m in net.modules():
            if isinstance(m, nn.PReLU):
                slope = float(m.weight.detach().reshape(-1)[0])
        out.append(slope)
    return out


def _cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError("learning_environments_b200: a CUDA device is required (there is no CPU implementation of this path)")
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------------------------------
class DeviceRealEnv(object):
    """CartPole-v0 / Acrobot-v1 (gym 0.17.3 classic_control + TimeLimit) with the dynamics in the CUDA extension.
    gym-like surface: reset() -> obs, step(a) -> (obs, reward, done, info), action_space, observation_space."""

    def __init__(self, env_name):
        if env_name not in ENV_DIMS:
            raise NotImplementedError("no device kernel for environment %r (CartPole-v0, Acrobot-v1)" % (env_name,))
        self.env_name = env_name
        self.kind = REAL_ENV_IDS[env_name]
        self.sd, self.ad = ENV_DIMS[env_name]
        self.action_space = Discrete(self.ad)
        if env_name == "CartPole-v0":
            high = np.array([4.8, np.finfo(np.float32).max, 24 * np.pi / 360 * 2, np.finfo(np.float32).max], np.float32)
            self._max_episode_steps = 200
            self._reset_half = 0.05
        else:
            high = np.array([1, 1, 1, 1, 4 * np.pi, 9 * np.pi], np.float32)
            self._max_episode_steps = 500
            self._reset_half = 0.1
        self.observation_space = Box(-high, high)
        self.np_random = np.random.RandomState()
        self._state = None
        self._elapsed = None

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        self.action_space.np_random = np.random.RandomState(seed)
        return [seed]

    def _obs(self, st):
        if self.kind == 0:
            return np.array(st)
        return np.array([np.cos(st[0]), np.sin(st[0]), np.cos(st[1]), np.sin(st[1]), st[2], st[3]])

    def reset(self):
        st = self.np_random.uniform(low=-self._reset_half, high=self._reset_half, size=(4,))
        self._state = torch.from_numpy(st.reshape(1, 4).copy())   # uploaded by the first step()
        self._elapsed = torch.zeros(1, dtype=torch.int32)
        return self._obs(st)

    def step(self, action):
        assert self._state is not None, "Cannot call env.step() before calling reset()"
        if not self._state.is_cuda:
            dev = _cuda_device()
            self._state, self._elapsed = self._state.to(dev), self._elapsed.to(dev)
        a = torch.tensor([int(action)], dtype=torch.int32, device=self._state.device)
        obs, r, d = ops.real_env_step(self.kind, int(self._max_episode_steps), self._state, self._elapsed, a, self.sd)
        st = self._state.cpu().numpy()[0]
        return self._obs(st), float(r.item()), bool(d.item() > 0.5), {}

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path")

    def close(self):
        pass


class VirtualEnv(nn.Module):
    """Synthetic Environment (envs/virtual_env.py): three MLPs over cat(one_hot(action), state)."""

    def __init__(self, kwargs):
        super().__init__()
        self.env_name = str(kwargs["env_name"])
        self.device = str(kwargs["device"])
        self.state_dim = int(kwargs["state_dim"])
        self.action_dim = int(kwargs["action_dim"])
        self.solved_reward = float(kwargs["solved_reward"])
        self._max_episode_steps = int(kwargs["max_steps"])
        self.action_space = kwargs["action_space"]
        self.observation_space = kwargs["observation_space"]
        self.reset_env = kwargs["reset_env"]
        self.kwargs = kwargs
        in_dim = self.state_dim + self.action_dim
        self.state_net = build_nn_from_config(in_dim, self.state_dim, kwargs)
        self.reward_net = build_nn_from_config(in_dim, 1, kwargs)
        self.done_net = build_nn_from_config(in_dim, 1, kwargs)
        self._theta_cache = None
        self.state = self.reset()

    def reset(self):
        self.state = torch.as_tensor(self.reset_env.reset()).clone()
        return self.state

    # -- parameter vector for the kernels -------------------------------------------------------------------
    def lane_cfg_fields(self):
        return dict(env_hidden=int(self.kwargs["hidden_size"]), env_act=ACT_IDS[str(self.kwargs["activation_fn"])],
                    env_slope=_prelu_slopes([self.state_net, self.reward_net, self.done_net]))

    def theta(self):
        return linear_theta(self)

    def _theta_dev(self):
        ver = tuple(p._version for p in self.parameters()) + tuple(id(p) for p in self.parameters())
        if self._theta_cache is None or self._theta_cache[0] != ver:
            self._theta_cache = (ver, self.theta().to(_cuda_device()).contiguous())
        return self._theta_cache[1]

    def _cfg(self):
        from ._abi import LaneCfg
        if int(self.kwargs.get("hidden_layer", 1)) > 1:
            raise NotImplementedError("SE nets with hidden_layer > 1 are outside the compiled kernel set")
        c = LaneCfg()
        c.sd, c.ad, c.env_kind = self.state_dim, self.action_dim, ENV_SE
        c.real_env = REAL_ENV_IDS[self.env_name]
        f = self.lane_cfg_fields()
        c.env_hidden, c.env_act = f["env_hidden"], f["env_act"]
        for i in range(3):
            c.env_slope[i] = f["env_slope"][i]
        return c

    def step(self, action, state=None):
        """action: one-hot [ad] (or [N, ad]); state: None (use self.state) or [N, sd] (envs/virtual_env.py:43-54)."""
        dev = _cuda_device()
        st = self.state if state is None else state
        st = torch.as_tensor(st, dtype=torch.float32)
        single = st.dim() == 1
        st2 = st.reshape(-1, self.state_dim).to(dev).contiguous()
        act = torch.as_tensor(action, dtype=torch.float32).reshape(-1, self.action_dim)
        a_idx = torch.argmax(act, dim=1).to(torch.int32).to(dev).contiguous()
        ns, r, d = ops.se_forward(self._cfg(), self._theta_dev()[None], st2, a_idx, lanes_per_member=st2.shape[0])
        if single:
            next_state, reward, done = ns[0], r[0:1], d[0:1]
        else:
            next_state, reward, done = ns, r[:, None], d[:, None]
        self.state = next_state
        return next_state, reward, done


class RewardEnv(nn.Module):
    """Reward Network around a real env (envs/reward_env.py); reward types 0,1,2,5,6 (state-only)."""

    def __init__(self, real_env, kwargs):
        super().__init__()
        self.env_name = str(kwargs["env_name"])
        self.device = str(kwargs["device"])
        self.state_dim = int(kwargs["state_dim"])
        self.action_dim = int(kwargs["action_dim"])
        self.info_dim = int(kwargs["info_dim"])
        self.solved_reward = float(kwargs["solved_reward"])
        self.reward_env_type = int(kwargs["reward_env_type"])
        self._max_episode_steps = int(kwargs["max_steps"])
        self.action_space = kwargs["action_space"]
        self.observation_space = kwargs["observation_space"]
        self.kwargs = kwargs
        self.real_env = real_env
        self.reward_net = self.build_reward_net(kwargs)
        self._theta_cache = None
        self.state = self.reset()

    def build_reward_net(self, kwargs):
        t = self.reward_env_type
        if t < 100:
            if t == 0:
                input_dim = 1
            elif t in (1, 2, 5, 6):
                input_dim = self.state_dim
            elif t in (3, 4, 7, 8):
                input_dim = self.state_dim + self.info_dim
            else:
                raise NotImplementedError('Unknown reward_env_type: ' + str(t))
            return build_nn_from_config(input_dim=input_dim, output_dim=1, nn_config=kwargs)
        if t in (101, 102):
            return nn.Linear(self.info_dim, 1, bias=False)
        raise NotImplementedError('Unknown reward_env_type: ' + str(t))

    def theta(self):
        return linear_theta(self.reward_net)

    def lane_cfg_fields(self):
        return dict(env_hidden=int(self.kwargs["hidden_size"]), env_act=ACT_IDS[str(self.kwargs["activation_fn"])],
                    env_slope=_prelu_slopes([self.reward_net]) * 3, rn_type=self.reward_env_type)

    def _cfg(self):
        from ._abi import LaneCfg
        c = LaneCfg()
        c.sd, c.ad, c.env_kind = self.state_dim, self.action_dim, ENV_RN
        c.real_env = REAL_ENV_IDS[self.env_name]
        f = self.lane_cfg_fields()
        c.env_hidden, c.env_act, c.rn_type = f["env_hidden"], f["env_act"], f["rn_type"]
        for i in range(3):
            c.env_slope[i] = f["env_slope"][i]
        c.gamma = float(getattr(self, "gamma", 0.0))
        return c

    def step(self, action):
        next_state, reward, done, info = self.real_env.step(action)
        reward_res = self._calc_reward(state=self.state, next_state=next_state, reward=reward, info=info)
        self.state = next_state
        return next_state, reward_res, done, {}

    def _calc_reward(self, state, next_state, reward, info):
        if 'TimeLimit.truncated' in info:
            info.pop('TimeLimit.truncated')
        t = self.reward_env_type
        if t in (3, 4, 7, 8, 101, 102):
            if not info:
                raise ValueError('No info dict provided by environment')
            raise NotImplementedError("info-vector reward types are outside the compiled kernel set")
        if t not in (0, 1, 2, 5, 6):
            raise NotImplementedError('Unknown reward_env_type: ' + str(t))
        dev = _cuda_device()
        if self._theta_cache is None or self._theta_cache[0] != tuple(p._version for p in self.parameters()):
            self._theta_cache = (tuple(p._version for p in self.parameters()), self.theta().to(dev).contiguous())
        s = torch.as_tensor(np.asarray(state), dtype=torch.float32).reshape(1, -1).to(dev)
        s2 = torch.as_tensor(np.asarray(next_state), dtype=torch.float32).reshape(1, -1).to(dev)
        rr = torch.tensor([reward], dtype=torch.float32, device=dev)
        theta = self._theta_cache[1] if t != 0 else torch.zeros(self._cfg().rn_params(), device=dev)
        out = ops.rn_reward(self._cfg(), theta[None], s, s2, rr)
        return out.item()

    def seed(self, seed):
        return self.real_env.seed(seed)

    def render(self):
        return self.real_env.render()

    def reset(self):
        self.state = self.real_env.reset()
        return self.state

    def close(self):
        return self.real_env.close()

    def set_agent_params(self, gamma):
        self.gamma = gamma


class EnvWrapper(nn.Module):
    """Uniform torch-tensor API over real / virtual / reward envs (envs/env_wrapper.py)."""

    def __init__(self, env):
        super().__init__()
        self.env = env
        self.same_action_num = 1

    def step(self, action, state=None):
        if self.is_virtual_env():
            reward_sum = None
            if self.has_discrete_action_space():
                action = to_one_hot_encoding(action, self.get_action_dim())
            for _ in range(self.same_action_num):
                if state is None:
                    state_, reward, done = self.env.step(action=action)
                else:
                    state_, reward, done = self.env.step(action=action, state=state)
                    state = state_
                reward_sum = reward if reward_sum is None else reward_sum + reward
            next_state = state_.to("cpu")
            if self.has_discrete_state_space():
                next_state = from_one_hot_encoding(next_state)
            return next_state, reward_sum.to("cpu"), done.to("cpu")
        action = torch.as_tensor(action).cpu().detach().numpy()
        if self.has_discrete_action_space():
            action = action.astype(int).reshape(-1)[0]
        reward_sum = 0
        for _ in range(self.same_action_num):
            state, reward, done, _ = self.env.step(action)
            reward_sum += reward
            if done:
                break
        next_state_torch = torch.tensor(np.asarray(state), device="cpu", dtype=torch.float32)
        reward_torch = torch.tensor(reward_sum, device="cpu", dtype=torch.float32)
        done_torch = torch.tensor(done, device="cpu", dtype=torch.float32)
        if next_state_torch.dim() == 0:
            next_state_torch = next_state_torch.unsqueeze(0)
        return next_state_torch, reward_torch, done_torch

    def reset(self):
        state = self.env.reset()
        if type(state) == np.ndarray:
            state_torch = torch.from_numpy(state).float().cpu()
        elif torch.is_tensor(state):
            state_torch = state.float().cpu()
        else:
            state_torch = torch.tensor([state], device="cpu", dtype=torch.float32)
        if self.has_discrete_state_space() and self.is_virtual_env():
            return from_one_hot_encoding(state_torch)
        return state_torch

    def get_random_action(self):
        action = self.env.action_space.sample()
        if type(action) == np.ndarray:
            return torch.from_numpy(action)
        return torch.tensor([action], device="cpu", dtype=torch.float32)

    def get_state_dim(self):
        return self.env.observation_space.shape[0] if self.env.observation_space.shape else self.env.observation_space.n

    def get_action_dim(self):
        return self.env.action_space.shape[0] if self.env.action_space.shape else self.env.action_space.n

    def get_max_action(self):
        return 2 if self.env.env_name == 'Pendulum-v0' else 1

    def get_min_action(self):
        return 0 if self.env.env_name == 'CartPole-v0' else -self.get_max_action()

    def has_discrete_action_space(self):
        return isinstance(self.env.action_space, Discrete)

    def has_discrete_state_space(self):
        return isinstance(self.env.observation_space, Discrete)

    def render(self):
        return self.env.render()

    def close(self):
        if not self.is_virtual_env():
            return self.env.close()

    def get_solved_reward(self):
        return self.env.solved_reward

    def max_episode_steps(self):
        return self.env._max_episode_steps

    def can_be_solved(self):
        return self.env.solved_reward < 1e9

    def seed(self, seed):
        if not self.is_virtual_env():
            return self.env.seed(seed)
        print("Setting manuel seed not yet implemented, performance may decrease")
        return 0

    def is_virtual_env(self):
        return isinstance(self.env, VirtualEnv)

    def set_agent_params(self, same_action_num, gamma):
        self.same_action_num = same_action_num
        if hasattr(self.env, "set_agent_params"):
            self.env.set_agent_params(gamma=gamma)

    # -- what the fused kernel needs to know about this env ---------------------------------------------------
    def kernel_env(self):
        """(env_kind, theta or None, cfg fields) for le_inner_loop_run."""
        if isinstance(self.env, VirtualEnv):
            return ENV_SE, self.env.theta(), self.env.lane_cfg_fields()
        if isinstance(self.env, RewardEnv):
            return ENV_RN, self.env.theta(), self.env.lane_cfg_fields()
        return ENV_REAL, None, {}


class EnvFactory(object):
    """envs/env_factory.py: builds real / virtual / reward envs from the nested config dict."""

    def __init__(self, config):
        self.env_name = config["env_name"]
        self.device = config["device"]
        self.env_config = config["envs"][self.env_name]
        dummy_env = self.generate_real_env(print_str='EnvFactory (dummy_env): ')
        self.state_dim = dummy_env.get_state_dim()
        self.action_dim = dummy_env.get_action_dim()
        self.observation_space = dummy_env.env.observation_space
        self.action_space = dummy_env.env.action_space

    def generate_real_env(self, print_str=''):
        kwargs = self._get_default_parameters(virtual_env=False)
        return EnvWrapper(env=self._generate_real_env_with_kwargs(kwargs=kwargs, env_name=self.env_name))

    def generate_virtual_env(self, print_str=''):
        kwargs = self._get_default_parameters(virtual_env=True)
        return EnvWrapper(env=VirtualEnv(kwargs))

    def generate_reward_env(self, print_str=''):
        kwargs = self._get_default_parameters(virtual_env=True)
        real_env = self._generate_real_env_with_kwargs(kwargs=kwargs, env_name=self.env_name)
        return EnvWrapper(env=RewardEnv(real_env=real_env, kwargs=kwargs))

    def _get_default_parameters(self, virtual_env):
        kwargs = {"env_name": self.env_name, "device": self.device}
        if virtual_env:
            kwargs["state_dim"] = self.state_dim
            kwargs["action_dim"] = self.action_dim
            kwargs["observation_space"] = self.observation_space
            kwargs["action_space"] = self.action_space
            kwargs["reset_env"] = self.generate_real_env()
        for key, value in self.env_config.items():
            kwargs[key] = float(value[1]) if isinstance(value, list) else value
        return kwargs

    def _generate_real_env_with_kwargs(self, kwargs, env_name):
        env = DeviceRealEnv(env_name)
        for key, value in kwargs.items():
            if key not in ("action_space", "observation_space"):
                setattr(env, key, value)
        env._max_episode_steps = int(kwargs["max_steps"])
        env.kwargs = kwargs
        return env
