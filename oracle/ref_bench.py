"""BASELINE INFRASTRUCTURE — times the UNMODIFIED reference (its own torch CPU path) on the host cores.

Used only by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg).  One worker process per host core with
``torch.set_num_threads(1)`` (experiments/GTN_Worker.py:8, yaml ``num_threads_per_worker: 1``; the process-per-worker layout of
agents/GTN.py:14-44); every task is one ``GTN_Worker.calc_score`` (agents/GTN_worker.py:187-221): a fresh agent from
``select_agent``, ``agent.train(env, test_env=real_env)`` and ``agent.test(real_env)`` — on the SAME bounded lane
configuration and the same synthetic SE / RN parameter vector as the GPU arm.  The reference runs with its own RNG (no
injection): this is a throughput measurement, not a parity test.  The reference tree is read from
``oracle/ref_harness.REFERENCE_ROOT`` (on the GPU box: the copy staged by ``oracle/make_ref.py``).
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_theta_into(module, theta):
    """Inverse of gen_golden.linear_params: nn.Linear weights / biases of `module` in module order <- flat theta."""
    import torch
    import torch.nn as nn
    off = 0
    with torch.no_grad():
        for l in module.modules():
            if isinstance(l, nn.Linear):
                n = l.weight.numel()
                l.weight.copy_(torch.from_numpy(theta[off:off + n].reshape(l.weight.shape).copy()))
                off += n
                if l.bias is not None:
                    n = l.bias.numel()
                    l.bias.copy_(torch.from_numpy(theta[off:off + n].copy()))
                    off += n
    assert off == theta.size, (off, theta.size)


def _worker(args):
    spec, theta, n_tasks, seed, deadline = args
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings("ignore")              # torch 2.x deprecation chatter of the 2020 reference code
    import numpy as np
    import torch
    torch.set_num_threads(1)                       # experiments/GTN_Worker.py:8
    import random
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    from oracle import ref_harness as rh
    mods = rh.import_reference()
    cfg = rh.load_reference_yaml(spec["yaml"])
    agent_name = spec["agent"]
    cfg["agents"][agent_name].update(spec["agent_overrides"])
    cfg["agents"][agent_name]["print_rate"] = 10 ** 9
    cfg["envs"][cfg["env_name"]].update(spec.get("env_overrides", {}))
    fac = mods["envs.env_factory"].EnvFactory(cfg)
    env = fac.generate_virtual_env() if spec["kind"] == "se" else fac.generate_reward_env()
    load_theta_into(env, theta)
    steps, times = 0, []
    devnull = open(os.devnull, "w")
    t_begin = time.perf_counter()
    for _ in range(n_tasks):
        if times and time.perf_counter() > deadline:
            break
        t0 = time.perf_counter()
        old = sys.stdout
        sys.stdout = devnull                        # BaseAgent.train prints per-episode progress
        try:
            # GTN_Worker.calc_score (agents/GTN_worker.py:187-221)
            agent = mods["agents.agent_utils"].select_agent(config=cfg, agent_name=agent_name)
            real_env = fac.generate_real_env()
            _, lengths, _ = agent.train(env=env, test_env=real_env)
            agent.test(env=real_env)
        finally:
            sys.stdout = old
        steps += int(sum(lengths))
        times.append(time.perf_counter() - t0)
    return steps, time.perf_counter() - t_begin, times


def run(spec, theta, n_tasks, n_workers=None, budget_s=60.0):
    """Runs `n_tasks` calc_scores of the unmodified reference over `n_workers` single-thread processes.
    Returns dict(steps, seconds (wall of the slowest worker), tasks, task_seconds (list), workers)."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        raise RuntimeError("reference tree not staged (run python -m oracle.make_ref where /root/reference exists)")
    n_workers = n_workers or os.cpu_count() or 1
    n_workers = max(1, min(n_workers, n_tasks))
    per = [n_tasks // n_workers + (1 if i < n_tasks % n_workers else 0) for i in range(n_workers)]
    deadline = time.perf_counter() + budget_s
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_workers) as pool:
        res = pool.map(_worker, [(spec, theta, per[i], 1000 + i, deadline) for i in range(n_workers)])
    wall = time.perf_counter() - t0
    task_seconds = [t for r in res for t in r[2]]
    return dict(steps=sum(r[0] for r in res), seconds=max(r[1] for r in res), wall_with_spawn=wall, tasks=len(task_seconds),
                task_seconds=task_seconds, workers=n_workers)
