"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference.

Puts the stand-in modules under ``oracle/stubs`` (gym 0.17.3, seaborn, ConfigSpace — all absent from
this image) and ``/root/reference`` on ``sys.path`` so that the reference's
own ``EnvFactory``, ``VirtualEnv``, ``RewardEnv``, ``DDQN``, ``BaseAgent.train/test``, ``GTN_Master`` …
run as they are.  Used by ``oracle/gen_golden.py`` (fixture generation, in the build container only) and by
the ``-m "not gpu"`` tests that are skipped when /root/reference is absent (it does not exist on the GPU box).

Also provides the RNG injection used for lock-step comparison (SURVEY.md §8c last row): every random source
of the reference's inner loop is redirected to the Philox4x32-10 lane streams of ``oracle/philox.py``:

    random.random()                      (agents/DDQN.py:98)            -> P_ACT word 0
    env.action_space.sample()            (envs/env_wrapper.py:88)       -> P_ACT word 1
    np.random.randint(0, size, B)        (utils.py:35)                  -> P_SAMPLE
    real-env reset() uniform draw        (gym classic_control)          -> P_RESET_TRAIN / P_RESET_TEST
    TD3_discrete_vary only (Td3RngInjector):
    Tensor.exponential_() in F.gumbel_softmax (models/actor_critic.py:36) -> P_TD3_EXPO
    torch.randn / torch.randn_like       (agents/TD3_discrete_vary.py:75,162,166) -> P_TD3_NORMAL
"""
import contextlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")


def _reference_root():
    """LE_REFERENCE_ROOT, else /root/reference (build container), else the byte-for-byte copy staged by oracle/make_ref.py
    under the git-ignored oracle/_ref/pyref (the only form in which the reference reaches the GPU box)."""
    env = os.environ.get("LE_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/agents"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref", "pyref")


REFERENCE_ROOT = _reference_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "agents"))


_imported = {}


def import_reference():
    """Returns a namespace dict of the reference modules on the hot path."""
    if _imported:
        return _imported
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    names = ["utils", "models.model_utils", "models.actor_critic", "envs.virtual_env", "envs.reward_env",
             "envs.env_wrapper", "envs.env_factory", "agents.base_agent", "agents.DDQN", "agents.DuelingDDQN",
             "agents.agent_utils", "agents.GTN_base", "agents.GTN_master", "agents.GTN_worker"]
    for n in names:
        _imported[n] = importlib.import_module(n)
    return _imported


def load_reference_yaml(name):
    import yaml
    with open(os.path.join(REFERENCE_ROOT, name), "r") as f:
        return yaml.safe_load(f)


@contextlib.contextmanager
def in_tmp_cwd():
    """GTN_Base.__init__ creates ./results/GTN_sync in the CWD (agents/GTN_base.py:13-17)."""
    import tempfile
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


class LaneRngInjector(object):
    """Redirects the reference's RNG sources for ONE agent/env pair to a Philox lane stream.

    Mirrors the counters of oracle/le_oracle.c (`lane_t`): train_steps, learn_iters, episode, test_calls.
    Use as a context manager around ``agent.train(...)`` / ``agent.test(...)``.
    """

    def __init__(self, key, action_dim, real_env_kind):
        from oracle import philox
        self.px = philox
        self.key = (int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF)
        self.ad = int(action_dim)
        self.real_env_kind = real_env_kind  # 'cartpole' | 'acrobot'
        self.train_steps = 0     # number of select_train_action calls so far
        self.learn_iters = 0     # number of replay samples so far
        self.episode = 0         # training-episode resets so far
        self.test_calls = 0      # test() invocations so far
        self.test_episode = 0
        self.in_test = False
        self._act_words = None
        self.sample_log = []

    # --- sources -----------------------------------------------------------------------------
    def random_random(self):
        w = self.px.philox4x32(self.train_steps, 0, self.px.P_ACT, 0, *self.key)
        self._act_words = w
        self.train_steps += 1
        return float(w[0] >> 8) * (1.0 / 16777216.0)

    def action_sample(self, space):
        w = self._act_words
        return int((int(w[1]) * self.ad) >> 32)

    def randint(self, low, high=None, size=None, dtype=int):
        assert low == 0 and size is not None
        import numpy as np
        idx = self.px.sample_indices(self.key, self.learn_iters, int(size), int(high))
        self.learn_iters += 1
        self.sample_log.append(idx.copy())
        return idx.astype(np.int64)

    def reset_draw(self, env):
        if self.in_test:
            w = self.px.philox4x32(self.test_calls, self.test_episode, self.px.P_RESET_TEST, 0, *self.key)
            self.test_episode += 1
        else:
            w = self.px.philox4x32(self.episode, 0, self.px.P_RESET_TRAIN, 0, *self.key)
            self.episode += 1
        half = 0.05 if self.real_env_kind == "cartpole" else 0.1
        return self.px.uniform_f64(w, -half, half)


@contextlib.contextmanager
def injected_rng(inj, reset_envs=(), action_spaces=()):
    """Patch the reference modules so that agents + envs draw from `inj`.

    reset_envs:    the *unwrapped* gym stand-in env objects whose reset() draw is hooked
    action_spaces: the Discrete space objects whose sample() is hooked (note: a VirtualEnv shares the
                   action_space object of EnvFactory's dummy env, envs/env_factory.py:17-21)
    """
    import random as _random
    import numpy as _np
    mods = import_reference()
    ddqn_mod = mods["agents.DDQN"]
    duel_mod = mods["agents.DuelingDDQN"]
    utils_mod = mods["utils"]

    class _R(object):
        random = staticmethod(inj.random_random)

    class _NPR(object):
        randint = staticmethod(inj.randint)

    class _NP(object):
        random = _NPR

        def __getattr__(self, name):
            return getattr(_np, name)

    saved = (ddqn_mod.random, duel_mod.random, utils_mod.np)
    ddqn_mod.random = _R
    duel_mod.random = _R
    utils_mod.np = _NP()
    for e in reset_envs:
        e.reset_hook = inj.reset_draw
    for sp in action_spaces:
        sp.sample_hook = inj.action_sample

    # BaseAgent.test(): mark the test phase so reset draws come from the test stream
    base_mod = mods["agents.base_agent"]
    orig_test = base_mod.BaseAgent.test

    def test_wrapper(self, env, time_remaining=1e9):
        inj.in_test = True
        inj.test_episode = 0
        try:
            return orig_test(self, env, time_remaining)
        finally:
            inj.in_test = False
            inj.test_calls += 1

    base_mod.BaseAgent.test = test_wrapper
    try:
        yield inj
    finally:
        base_mod.BaseAgent.test = orig_test
        ddqn_mod.random, duel_mod.random, utils_mod.np = saved
        for e in reset_envs:
            e.reset_hook = None
        for sp in action_spaces:
            sp.sample_hook = None


class Td3RngInjector(LaneRngInjector):
    """LaneRngInjector plus the torch draws of TD3_discrete_vary (agents/TD3_discrete_vary.py:73-75,155-166 and the
    exponential_() inside F.gumbel_softmax) on the P_TD3_EXPO / P_TD3_NORMAL streams of oracle/philox.py."""

    def __init__(self, key, action_dim, real_env_kind):
        super().__init__(key, action_dim, real_env_kind)
        self.test_step_in_ep = 0
        self.ctx = None          # ("train" | "test" | "learn", c0, sub); learn: expo call count selects phase 2 / 3
        self.learn_expo_calls = 0

    def action_sample(self, space):      # env.get_random_action() of the init episodes: P_ACT word 1 of this train step
        w = self.px.philox4x32(self.train_steps, 0, self.px.P_ACT, 0, *self.key)
        return int((int(w[1]) * self.ad) >> 32)

    def reset_draw(self, env):
        if self.in_test:
            self.test_step_in_ep = 0
        return super().reset_draw(env)

    def expo(self, n):
        kind, c0, sub = self.ctx
        if kind == "learn":
            phase = 2 + self.learn_expo_calls
            self.learn_expo_calls += 1
        else:
            phase = 0 if kind == "train" else 1
        return self.px.td3_expo(self.key, phase, c0, n, sub)

    def normal(self, n):
        kind, c0, sub = self.ctx
        return self.px.td3_normal(self.key, {"train": 0, "test": 1, "learn": 2}[kind], c0, n, sub)


@contextlib.contextmanager
def injected_rng_td3(inj, agent, reset_envs=(), action_spaces=()):
    """injected_rng for a TD3_discrete_vary agent: additionally patches torch.randn / torch.randn_like /
    Tensor.exponential_ and wraps the agent's select_*_action / learn so that every draw knows its stream position."""
    import torch
    orig = (torch.randn, torch.randn_like, torch.Tensor.exponential_, agent.select_train_action, agent.select_test_action, agent.learn)

    def fake_randn(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        return torch.from_numpy(inj.normal(int(torch.Size(shape).numel())).reshape(shape))

    def fake_randn_like(t, **kw):
        return torch.from_numpy(inj.normal(t.numel()).reshape(tuple(t.shape)))

    def fake_exponential_(self, lambd=1.0, generator=None):
        with torch.no_grad():
            self.copy_(torch.from_numpy(inj.expo(self.numel()).reshape(tuple(self.shape))))
        return self

    def select_train(state, env, episode):
        inj.ctx = ("train", inj.train_steps, 0)
        try:
            return orig[3](state=state, env=env, episode=episode)
        finally:
            inj.train_steps += 1

    def select_test(state, env):
        inj.ctx = ("test", (inj.test_calls << 16) | inj.test_step_in_ep, inj.test_episode - 1)
        try:
            return orig[4](state, env)
        finally:
            inj.test_step_in_ep += 1

    def learn(replay_buffer, env, episode):
        inj.ctx = ("learn", inj.learn_iters, 0)   # replay_buffer.sample() inside increments learn_iters afterwards
        inj.learn_expo_calls = 0
        return orig[5](replay_buffer=replay_buffer, env=env, episode=episode)

    torch.randn, torch.randn_like, torch.Tensor.exponential_ = fake_randn, fake_randn_like, fake_exponential_
    agent.select_train_action, agent.select_test_action, agent.learn = select_train, select_test, learn
    try:
        with injected_rng(inj, reset_envs=reset_envs, action_spaces=action_spaces):
            yield inj
    finally:
        torch.randn, torch.randn_like, torch.Tensor.exponential_ = orig[:3]
        agent.select_train_action, agent.select_test_action, agent.learn = orig[3:]
