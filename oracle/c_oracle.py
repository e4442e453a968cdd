"""TEST INFRASTRUCTURE ONLY — ctypes/numpy front end of oracle/le_oracle.c (the CPU restatement).

Builds oracle/_ref/lible_oracle.so on demand with oracle/Makefile.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from learning_environments_b200._abi import LaneCfg, LaneOut, Trace

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "lible_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "le_oracle.c")
    stale = (not os.path.isfile(_SO)) or os.path.getmtime(_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "le_oracle.h")),
        os.path.getmtime(os.path.join(_HERE, "..", "include", "le_b200.h")))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.le_oracle_td_update.restype = C.c_float
        _lib.le_oracle_sizeof_cfg.restype = C.c_int
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def philox(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint32 * 4)()
    lib().le_oracle_philox(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), C.c_uint32(k0),
                           C.c_uint32(k1), out)
    return tuple(out)


def se_step(cfg, theta, state, action):
    theta = np.ascontiguousarray(theta, np.float32)
    state = np.ascontiguousarray(state, np.float32)
    ns = np.zeros(cfg.sd, np.float32)
    r = C.c_float()
    d = C.c_float()
    lib().le_oracle_se_step(C.byref(cfg), _p(theta), _p(state), C.c_int(int(action)), _p(ns), C.byref(r), C.byref(d))
    return ns, r.value, d.value


def rn_reward(cfg, theta, s, s2, real_reward):
    theta = np.ascontiguousarray(theta, np.float32)
    s = np.ascontiguousarray(s, np.float32)
    s2 = np.ascontiguousarray(s2, np.float32)
    out = C.c_float()
    rc = lib().le_oracle_rn_reward(C.byref(cfg), _p(theta), _p(s), _p(s2), C.c_float(real_reward), C.byref(out))
    if rc != 0:
        raise ValueError("No info dict provided by environment")  # envs/reward_env.py:92
    return out.value


def q_forward(cfg, q_theta, state):
    q_theta = np.ascontiguousarray(q_theta, np.float32)
    state = np.ascontiguousarray(state, np.float32)
    q = np.zeros(cfg.ad, np.float32)
    a = C.c_int()
    lib().le_oracle_q_forward(C.byref(cfg), _p(q_theta), _p(state), _p(q), C.byref(a))
    return q, a.value


def real_step(real_env, max_steps, state64, elapsed, action, sd):
    st = np.ascontiguousarray(state64, np.float64).copy()
    el = C.c_int(int(elapsed))
    obs = np.zeros(sd, np.float32)
    r = C.c_float()
    d = C.c_float()
    lib().le_oracle_real_step(C.c_int(real_env), C.c_int(max_steps), _p(st), C.byref(el), C.c_int(int(action)), _p(obs),
                              C.byref(r), C.byref(d))
    return st, el.value, obs, r.value, d.value


def td_update(cfg, th, thT, m, v, adam_t, rows):
    """In-place DDQN.learn on explicit rows [B][2sd+3]. Returns (loss, new_adam_t)."""
    for a in (th, thT, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    rows = np.ascontiguousarray(rows, np.float32)
    t = C.c_int32(int(adam_t))
    loss = lib().le_oracle_td_update(C.byref(cfg), _p(th), _p(thT), _p(m), _p(v), C.byref(t), _p(rows),
                                     C.c_int(rows.shape[0]))
    return float(loss), t.value


def q_init(cfg, key):
    th = np.zeros(cfg.q_params(), np.float32)
    lib().le_oracle_q_init(C.byref(cfg), C.c_uint32(key[0]), C.c_uint32(key[1]), _p(th))
    return th


class TraceBuf(object):
    def __init__(self, cap, sd):
        self.cap = cap
        self.action = np.full(cap, -1, np.int32)
        self.explore = np.zeros(cap, np.int32)
        self.next_state = np.zeros((cap, sd), np.float32)
        self.reward = np.zeros(cap, np.float32)
        self.done = np.zeros(cap, np.float32)
        self.loss = np.full(cap, np.nan, np.float32)
        self.qgap = np.full(cap, np.nan, np.float32)

    def struct(self):
        t = Trace()
        t.cap = self.cap
        for n in ("action", "explore", "next_state", "reward", "done", "loss", "qgap"):
            setattr(t, n, getattr(self, n).ctypes.data)
        return t


def run_lane(cfg, env_theta, key, q_init_w=None, trace_cap=0):
    """One full calc_score on the CPU restatement. Returns a dict."""
    env_theta = None if env_theta is None else np.ascontiguousarray(env_theta, np.float32)
    Pq = cfg.q_params()
    qi = None if q_init_w is None else np.ascontiguousarray(q_init_w, np.float32)
    qf = np.zeros(Pq, np.float32)
    out = LaneOut()
    rewards = np.zeros(max(cfg.train_episodes, 1), np.float64)
    lengths = np.zeros(max(cfg.train_episodes, 1), np.int32)
    test_rewards = np.zeros(max(cfg.test_episodes, 1), np.float64)
    tb = TraceBuf(trace_cap, cfg.sd) if trace_cap > 0 else None
    ts = tb.struct() if tb else None
    rc = lib().le_oracle_run_lane(C.byref(cfg), _p(env_theta), C.c_uint32(key[0]), C.c_uint32(key[1]), _p(qi), _p(qf),
                                  C.byref(out), _p(rewards), _p(lengths), _p(test_rewards),
                                  C.byref(ts) if ts is not None else None)
    if rc != 0:
        raise RuntimeError("le_oracle_run_lane failed: %d" % rc)
    n = out.n_episodes
    return dict(n_episodes=n, timed_out=out.timed_out, train_steps=out.train_steps, learn_iters=out.learn_iters,
                test_steps=out.test_steps, score=out.score, rewards=rewards[:n].copy(), lengths=lengths[:n].copy(),
                test_rewards=test_rewards[:cfg.test_episodes].copy(), q_final=qf, trace=tb)


def run_lanes(cfgs, env_theta, env_index, keys, q_init_w=None, n_threads=1):
    """n lanes on n_threads host threads. cfgs: one LaneCfg or a list (per-lane). Returns dict of arrays."""
    if isinstance(cfgs, LaneCfg):
        cfgs = [cfgs]
    n_cfg = len(cfgs)
    arr = (LaneCfg * n_cfg)(*cfgs)
    c0 = cfgs[0]
    keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 2)
    n = keys.shape[0]
    env_theta = np.ascontiguousarray(env_theta, np.float32)
    if env_theta.ndim == 1:
        env_theta = env_theta[None]
    P_env = env_theta.shape[1]
    ei = None if env_index is None else np.ascontiguousarray(env_index, np.int32)
    Pq = c0.q_params()
    qi = None if q_init_w is None else np.ascontiguousarray(q_init_w, np.float32)
    qf = np.zeros((n, Pq), np.float32)
    out = (LaneOut * n)()
    rewards = np.zeros((n, max(c0.train_episodes, 1)), np.float64)
    lengths = np.zeros((n, max(c0.train_episodes, 1)), np.int32)
    test_rewards = np.zeros((n, max(c0.test_episodes, 1)), np.float64)
    rc = lib().le_oracle_run_lanes(arr, C.c_int(n_cfg), _p(env_theta), C.c_int(P_env), _p(ei), _p(keys), _p(qi), _p(qf),
                                   C.c_int(n), out, _p(rewards), _p(lengths), _p(test_rewards), C.c_int(n_threads))
    if rc != 0:
        raise RuntimeError("le_oracle_run_lanes failed: %d" % rc)
    return dict(n_episodes=np.array([o.n_episodes for o in out]), timed_out=np.array([o.timed_out for o in out]),
                train_steps=np.array([o.train_steps for o in out]), learn_iters=np.array([o.learn_iters for o in out]),
                test_steps=np.array([o.test_steps for o in out]), score=np.array([o.score for o in out]),
                rewards=rewards, lengths=lengths, test_rewards=test_rewards, q_final=qf)


def td3_learn(dims, hyper, nets, adam, counters, total_it, rows, policy_noise, expo_target, expo_actor, gumbel_tau):
    """TD3_discrete_vary.learn restatement, in place on the float32 arrays in `nets` (actor, actor_target, critic_1,
    critic_target_1, critic_2, critic_target_2) and `adam` (m_a, v_a, m_c1, v_c1, m_c2, v_c2).
    dims = (sd, ad, H, L, act); hyper = dict(gamma, tau, lr, policy_delay, max_action, policy_std, policy_std_clip, gumbel_hard);
    counters = [t_actor, t_critic] (updated).  Returns (critic_loss, actor_loss or nan)."""
    L = lib()
    L.le_oracle_td3_learn.restype = C.c_float
    sd, ad, H, nl, act = dims
    ta, tc = C.c_int32(counters[0]), C.c_int32(counters[1])
    al = C.c_float()
    rows = np.ascontiguousarray(rows, np.float32)
    args = [np.ascontiguousarray(a, np.float32) for a in (policy_noise, expo_target, expo_actor)]
    for a in list(nets.values()) + list(adam.values()):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    loss = L.le_oracle_td3_learn(
        C.c_int(sd), C.c_int(ad), C.c_int(H), C.c_int(nl), C.c_int(act), C.c_double(hyper["gamma"]), C.c_double(hyper["tau"]),
        C.c_double(hyper["lr"]), C.c_int(hyper["policy_delay"]), C.c_float(hyper["max_action"]), C.c_float(hyper["policy_std"]),
        C.c_float(hyper["policy_std_clip"]), C.c_float(gumbel_tau), C.c_int(hyper["gumbel_hard"]),
        _p(nets["actor"]), _p(nets["actor_target"]), _p(nets["critic_1"]), _p(nets["critic_target_1"]), _p(nets["critic_2"]),
        _p(nets["critic_target_2"]), _p(adam["m_a"]), _p(adam["v_a"]), _p(adam["m_c1"]), _p(adam["v_c1"]), _p(adam["m_c2"]),
        _p(adam["v_c2"]), C.byref(ta), C.byref(tc), C.c_int(total_it), _p(rows), C.c_int(rows.shape[0]), _p(args[0]), _p(args[1]),
        _p(args[2]), C.byref(al))
    counters[0], counters[1] = ta.value, tc.value
    return float(loss), float(al.value)


class Td3Cfg(C.Structure):
    """struct le_oracle_td3_cfg (oracle/le_oracle.h)."""
    _fields_ = [("base", LaneCfg), ("policy_delay", C.c_int32), ("gumbel_hard", C.c_int32), ("action_std", C.c_double),
                ("policy_std", C.c_double), ("policy_std_clip", C.c_double), ("gumbel_temp", C.c_double), ("max_action", C.c_double)]


def td3_cfg(base, agent_cfg, max_action=1.0):
    """base: LaneCfg with the env / loop fields; agent_cfg: the reference's config['agents']['td3_discrete_vary'] dict."""
    t = Td3Cfg()
    C.memmove(C.byref(t.base), C.byref(base), C.sizeof(LaneCfg))
    t.policy_delay, t.gumbel_hard = int(agent_cfg["policy_delay"]), int(bool(agent_cfg["gumbel_softmax_hard"]))
    t.action_std, t.policy_std, t.policy_std_clip = float(agent_cfg["action_std"]), float(agent_cfg["policy_std"]), float(agent_cfg["policy_std_clip"])
    t.gumbel_temp, t.max_action = float(agent_cfg["gumbel_softmax_temp"]), float(max_action)
    return t


def run_lane_td3(tcfg, env_theta, key, actor_init, c1_init, c2_init, trace_cap=0):
    """One TD3_discrete_vary calc_score-style lane (train with per-episode test + final test) on the CPU restatement."""
    cfg = tcfg.base
    env_theta = None if env_theta is None else np.ascontiguousarray(env_theta, np.float32)
    a0, q1, q2 = [np.ascontiguousarray(x, np.float32) for x in (actor_init, c1_init, c2_init)]
    af = np.zeros_like(a0)
    out = LaneOut()
    rewards = np.zeros(max(cfg.train_episodes, 1), np.float64)
    lengths = np.zeros(max(cfg.train_episodes, 1), np.int32)
    test_rewards = np.zeros(max(cfg.test_episodes, 1), np.float64)
    tb = TraceBuf(trace_cap, cfg.sd) if trace_cap > 0 else None
    ts = tb.struct() if tb else None
    rc = lib().le_oracle_run_lane_td3(C.byref(tcfg), _p(env_theta), C.c_uint32(key[0]), C.c_uint32(key[1]), _p(a0), _p(q1), _p(q2), _p(af),
                                      C.byref(out), _p(rewards), _p(lengths), _p(test_rewards), C.byref(ts) if ts is not None else None)
    if rc != 0:
        raise RuntimeError("le_oracle_run_lane_td3 failed: %d" % rc)
    n = out.n_episodes
    return dict(n_episodes=n, timed_out=out.timed_out, train_steps=out.train_steps, learn_iters=out.learn_iters, test_steps=out.test_steps,
                score=out.score, rewards=rewards[:n].copy(), lengths=lengths[:n].copy(),
                test_rewards=test_rewards[:cfg.test_episodes].copy(), actor_final=af, trace=tb)
