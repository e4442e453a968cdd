"""Minimal stand-in for gym==0.17.3 (TEST INFRASTRUCTURE ONLY).

The reference pins gym==0.17.3 (/root/reference/requirements.txt:47) but does not vendor it and it
cannot be installed here (no network).  This package restates the few pieces the reference's hot path
touches (call sites: envs/env_factory.py:83 ``gym.make``, envs/env_wrapper.py:58,73,88,
envs/reward_env.py:62,142): ``Env``, ``make`` for CartPole-v0 / Acrobot-v1, ``spaces.Discrete/Box``,
``wrappers.TimeLimit`` and ``utils.seeding.np_random``.  The dynamics follow the public gym 0.17.3
``classic_control`` sources as restated in SURVEY.md Appendix A ("parity unpinned" w.r.t. upstream gym:
no upstream checkout is available to diff against).

It is injected into ``sys.modules`` by ``oracle/ref_harness.py`` so that the UNMODIFIED reference under
/root/reference can be imported.  Nothing in the product package imports it.
"""
from . import spaces, wrappers, utils  # noqa: F401
from .core import Env  # noqa: F401
from .envs.classic_control import CartPoleEnv, AcrobotEnv
from .wrappers import TimeLimit

_REGISTRY = {
    "CartPole-v0": (CartPoleEnv, 200),
    "Acrobot-v1": (AcrobotEnv, 500),
}


def make(env_name):
    if env_name not in _REGISTRY:
        raise ValueError("gym stand-in knows only %s (asked for %r)" % (sorted(_REGISTRY), env_name))
    cls, max_steps = _REGISTRY[env_name]
    return TimeLimit(cls(), max_episode_steps=max_steps)
