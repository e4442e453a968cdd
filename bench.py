#!/usr/bin/env python
"""bench.py — SE env-steps/s of the NES inner loop (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cartpole_se|acrobot_se|cartpole_rn]

One "step" = one NES generation's population evaluation over one batch of synthetic input: every member's SE is
perturbed from the Philox stream (theta, theta+eps, theta-eps) and every resulting lane runs a complete, bounded
calc_score (agents/GTN_worker.py:187-221: DDQN train() with per-episode test() on the real env, then the final
test()) inside the persistent fused kernel; with N > 1 ranks members are sharded (weak scaling, no data-path
collective) and the generation ends with the all-gather of fitness scores + the NES update on every rank.

value  = training env-steps of all lanes of all ranks / device time (CUDA events, max over ranks), inputs resident in HBM
e2e    = the same through PopulationEvaluator.evaluate() with HOST theta: H2D of theta/keys and D2H of lane results inside
roofline.achieved = F_step (SURVEY.md §8d algorithmic flop per inner-loop step) x steps / fused-kernel time
cpu_baseline      = oracle/le_oracle.c (C restatement of the reference's loop, kind "port") on the host cores
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# workload name -> (default_configs name, env kind, members per GPU, bounded train_episodes)
WORKLOADS = {
    "cartpole_se": dict(cfg="cartpole_syn_env", kind="se", members_per_gpu=1184, train_episodes=10),
    "acrobot_se": dict(cfg="acrobot_syn_env", kind="se", members_per_gpu=1184, train_episodes=3, init_episodes=1),
    "cartpole_rn": dict(cfg="cartpole_reward_env", kind="rn", members_per_gpu=1184, train_episodes=40),
    # BASELINE config 5 (scaling sweep): SE hidden width 1024, 257 lanes per member (theta + 128 x (+eps, -eps) evaluations);
    # 64 members per GPU = population 512 x 256 envs on 8 GPUs
    "sweep_h1024": dict(cfg="cartpole_syn_env", kind="se", members_per_gpu=64, train_episodes=2, env_hidden=1024, grad_evals=128),
    # DuelingDDQN inner loops (general CTA-per-lane kernel): CartPole yaml section, Acrobot section of default_config_acrobot.yaml
    "cartpole_se_dueling": dict(cfg="cartpole_syn_env", kind="se", agent="duelingddqn", members_per_gpu=296, train_episodes=3),
    # 197 members x 3 lanes = 591 lanes = two full waves of the 296 resident CTAs
    "acrobot_se_dueling": dict(cfg="acrobot_syn_env", kind="se", agent="duelingddqn", members_per_gpu=197, train_episodes=2, init_episodes=1),
    # BASELINE config 4 (vary_hp evaluation): 4096 DDQN agents per GPU with per-lane lr / batch_size / hidden_size / hidden_layer
    # on ONE fixed CartPole SE, init_episodes=10, plateau early-out; train_episodes bounded (the evaluator's cap is 1000)
    "vary_hp": dict(cfg="cartpole_syn_env", kind="se", members_per_gpu=4096, train_episodes=30),
}


def build_lane_cfg(workload):
    from learning_environments_b200 import config, default_configs
    from learning_environments_b200._abi import ENV_RN, ENV_SE
    w = WORKLOADS[workload]
    d = default_configs.get(w["cfg"])
    name = w.get("agent", "ddqn")
    agent = d["agents"][name]
    agent["train_episodes"] = w["train_episodes"]
    if "init_episodes" in w:
        agent["init_episodes"] = w["init_episodes"]
    if "env_hidden" in w:
        d["envs"][d["env_name"]]["hidden_size"] = w["env_hidden"]
    if "grad_evals" in w:
        d["agents"]["gtn"]["num_grad_evals"] = w["grad_evals"]
    cfg = config.lane_cfg(d, name, ENV_SE if w["kind"] == "se" else ENV_RN, use_test_env=True, final_test=True)
    return d, cfg


def f_parts(cfg):
    """(flop per env step without learning, flop per TD update): F_env + F_q and 5*B*F_q."""
    from learning_environments_b200._abi import ENV_SE, ENV_RN
    fq = 2 * sum(i * o for i, o in cfg.q_layer_dims())
    if cfg.env_kind == ENV_SE:
        i, h = cfg.sd + cfg.ad, cfg.env_hidden
        fenv = 2 * (3 * i * h + h * (cfg.sd + 2))
    elif cfg.env_kind == ENV_RN:
        fenv = 2 * 2 * (cfg.sd * cfg.env_hidden + cfg.env_hidden)
    else:
        fenv = 0
    return fenv + fq, 5 * cfg.batch_size * fq


def f_step(cfg):
    """Algorithmic flop per inner-loop step: F_env + F_q + 5*B*F_q, F_mlp = sum 2*in*out."""
    from learning_environments_b200._abi import ENV_SE, ENV_RN
    fq = 2 * sum(i * o for i, o in cfg.q_layer_dims())
    if cfg.env_kind == ENV_SE:
        i, h = cfg.sd + cfg.ad, cfg.env_hidden
        fenv = 2 * (3 * i * h + h * (cfg.sd + 2))
    elif cfg.env_kind == ENV_RN:
        fenv = 2 * 2 * (cfg.sd * cfg.env_hidden + cfg.env_hidden)
    else:
        fenv = 0
    return fenv + fq + 5 * cfg.batch_size * fq


def synthetic_theta(cfg, seed=0):
    """torch-default-init-like parameter vector of the SE / RN (U(+-1/sqrt(fan_in)) per layer), seeded."""
    from learning_environments_b200._abi import ENV_SE
    rng = np.random.RandomState(seed)
    parts = []
    if cfg.env_kind == ENV_SE:
        i, h = cfg.sd + cfg.ad, cfg.env_hidden
        for out in (cfg.sd, 1, 1):
            parts += [rng.uniform(-1, 1, h * i) / np.sqrt(i), rng.uniform(-1, 1, h) / np.sqrt(i),
                      rng.uniform(-1, 1, out * h) / np.sqrt(h), rng.uniform(-1, 1, out) / np.sqrt(h)]
    else:
        i, h = cfg.sd, cfg.env_hidden
        parts += [rng.uniform(-1, 1, h * i) / np.sqrt(i), rng.uniform(-1, 1, h) / np.sqrt(i),
                  rng.uniform(-1, 1, h) / np.sqrt(h), rng.uniform(-1, 1, 1) / np.sqrt(h)]
    return np.concatenate(parts).astype(np.float32)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv is not None and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(cfg, theta, target_seconds, n_threads):
    """Times the C restatement (oracle/) on a bounded sample of the same workload: `n` lanes on n_threads threads."""
    from oracle import c_oracle, philox
    c_oracle.build()
    keys1 = np.array([philox.lane_key(99, 0, 0, 0, 0)], np.uint32)
    t0 = time.perf_counter()
    r1 = c_oracle.run_lanes(cfg, theta, None, keys1, n_threads=1)
    t1 = max(time.perf_counter() - t0, 1e-4)
    n = int(max(n_threads, min(4096, n_threads * target_seconds / t1)))
    keys = np.array([philox.lane_key(99, 0, i, 0, 0) for i in range(n)], np.uint32)
    t0 = time.perf_counter()
    r = c_oracle.run_lanes(cfg, theta, None, keys, n_threads=n_threads)
    dt = time.perf_counter() - t0
    steps = int(r["train_steps"].sum())
    return steps / dt, dict(lanes=n, steps=steps, seconds=dt, single_lane_seconds=t1, single_lane_steps=int(r1["train_steps"][0]))


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (C restatement, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    d, cfg = build_lane_cfg(args.workload)
    theta = synthetic_theta(cfg)
    cores = os.cpu_count() or 1
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        v, info = cpu_baseline(cfg, theta, target_seconds=3.0, n_threads=cores)
        if i >= args.warmup:
            vals.append((v, info))
    tot_steps = sum(i["steps"] for _, i in vals)
    tot_s = sum(i["seconds"] for _, i in vals)
    value = tot_steps / tot_s
    line = {
        "impl": "reference", "metric": "se_env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, cfg, 0, None),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d lanes/step of the same lane config (bounded calc_score), C restatement oracle/le_oracle.c, "
                                   "%d pthreads" % (info["lanes"], cores)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(name, cfg, members_per_gpu, plan):
    w = WORKLOADS[name]
    c = {"workload": "%s: NES generation, %s, %s inner loop (B=%d, Q %d->%d->%d), per lane %d train episodes x <=%d steps "
                     "+ per-episode test() of %d real-env episodes + final test()" % (
                         name, w["cfg"], "DuelingDDQN" if cfg.q_kind else "DDQN", cfg.batch_size, cfg.sd, cfg.q_hidden, cfg.ad,
                         cfg.train_episodes, cfg.max_steps,
                         cfg.test_episodes),
         "members_per_gpu": members_per_gpu, "lanes_per_member": 1 + 2 * w.get("grad_evals", 1), "env_hidden": cfg.env_hidden,
         "l2": "256 MiB buffer written between timed steps (L2 flush)"}
    if plan:
        c.update({"resident_warp_slots": plan["slots"], "replay_ring_rows": plan["ring_cap"], "units_per_thread": plan["units"]})
    return c


def run_vary_hp_workload(args):
    """--workload vary_hp: BASELINE config 4.  One step = every agent of this rank trained on the SE (virtual-env plateau rule)
    and tested on the real env: one launch of the register kernel (hidden_layer <= 1, hidden_size <= 128) and one of the
    general kernel (the rest), per-lane le_lane_cfg."""
    import torch
    import torch.distributed as dist
    from learning_environments_b200 import default_configs, ops, vary_hp
    from learning_environments_b200.rng import lane_keys
    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    w = WORKLOADS["vary_hp"]
    d = default_configs.get(w["cfg"])
    n = args.members_per_gpu or w["members_per_gpu"]
    over = dict(vary_hp.OVERRIDES, train_episodes=w["train_episodes"])
    cfgs = vary_hp.sample_agent_cfgs(d, n, np.random.RandomState(1000 + rank), over, True, None)
    theta_host = synthetic_theta(cfgs[0])
    theta_dev = torch.from_numpy(theta_host).to(dev).reshape(1, -1)
    groups = []
    for resident in (True, False):
        idx = np.array([i for i, c in enumerate(cfgs) if c.q_is_register_resident() == resident], int)
        if len(idx):
            sub = [cfgs[i] for i in idx]
            cfg0 = vary_hp._max_cfg(sub)
            groups.append(dict(idx=idx, sub=sub, cfg0=cfg0, bufs=ops.InnerLoopBuffers(cfg0, len(idx), 1, dev, n_cfg=len(idx))))
    total = args.warmup + args.steps
    keys = [torch.from_numpy(lane_keys(4321, g, rank * n + np.arange(n), np.zeros(n, int), np.zeros(n, int)).view(np.int32).copy()).to(dev)
            for g in range(total)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ffma_peak = ops.bench_ffma()

    def step(g):
        for gr in groups:
            ops.inner_loop_run(gr["bufs"], gr["sub"], theta_dev, None, keys[g][torch.from_numpy(gr["idx"]).to(dev)].contiguous(), cfg0=gr["cfg0"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for g in range(args.warmup):
        step(g)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms, steps_done, flop = 0.0, 0, 0.0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(args.warmup + k)
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
        for gr in groups:
            res = gr["bufs"].results()
            steps_done += int(res["train_steps"].sum())
            for c, st, li in zip(gr["sub"], res["train_steps"], res["learn_iters"]):
                a, b = f_parts(c)
                flop += float(st) * a + float(li) * b
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([dev_ms, float(steps_done), flop], dtype=torch.float64, device=dev)
    if world > 1:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, steps_all, flop_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        steps_all, flop_all = float(steps_done), flop
    # end to end through the public API: sampling, H2D of theta / keys / per-lane cfgs, both launches, D2H of the results
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0
    for k in range(args.steps):
        _, st, _, _ = vary_hp.evaluate_agents(d, theta_host.reshape(1, -1), n, seed=1000 + rank, overrides=over, device=dev, shard=False)
        e2e_steps += sum(st[0])
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        a, b = te.clone(), te.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_steps = float(a[0]), float(b[1])
    if rank == 0:
        achieved = flop_all / world / (dev_ms * 1e-3) / 1e12
        n_gen = sum(len(g["idx"]) for g in groups if not g["cfg0"].q_is_register_resident())
        line = {
            "metric": "se_env_steps_per_s", "value": steps_all / (dev_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "vary_hp: %d DDQN_vary agents per GPU on one CartPole SE (lr, batch_size in [66,597], hidden_size in "
                                   "[19,171], hidden_layer in {1,2} per lane), init_episodes=10, <=%d train episodes, plateau early-out, "
                                   "final test() of 10 real-env episodes" % (n, w["train_episodes"]),
                       "agents_per_gpu": n, "general_kernel_lanes": n_gen, "l2": "256 MiB buffer written between timed steps (L2 flush)"},
            "e2e": {"value": e2e_steps / e2e_s, "unit": "env-steps/s",
                    "h2d_bytes_per_step": int(theta_host.nbytes + n * 8 + n * 200), "d2h_bytes_per_step": int(n * (40 + 10 * 8))},
            "gpu_launches": (1 + len(groups)) * args.steps, "clocks": clocks,
            "roofline": {"bound": "fp32_ffma", "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s",
                         "frac": achieved / ffma_peak if ffma_peak else None, "traffic": None,
                         "peak_source": "le_bench_ffma microbenchmark in this run"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cartpole_se", choices=sorted(WORKLOADS))
    ap.add_argument("--members-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lane-override", action="append", default=[], metavar="FIELD=VALUE",
                    help="profiling aid: override an le_lane_cfg field (e.g. max_steps=100); recorded in config.overrides")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "vary_hp":
        return run_vary_hp_workload(args)

    import torch
    import torch.distributed as dist
    from learning_environments_b200 import ops
    from learning_environments_b200.engine import PopulationEvaluator
    from learning_environments_b200.nes import score_transform

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    d, cfg = build_lane_cfg(args.workload)
    for ov in args.lane_override:
        k, v = ov.split("=")
        setattr(cfg, k, type(getattr(cfg, k))(float(v)))
    gtn = d["agents"]["gtn"]
    mpg = args.members_per_gpu or WORKLOADS[args.workload]["members_per_gpu"]
    pop = mpg * world
    ev = PopulationEvaluator(cfg, pop, member_lo=rank * mpg, member_hi=(rank + 1) * mpg, num_grad_evals=gtn["num_grad_evals"],
                             seed=1234, noise_std=gtn["noise_std"], device=dev)
    plan = ops.inner_loop_plan(cfg, ev.n_lanes, ev.n_env)
    theta_host = torch.from_numpy(synthetic_theta(cfg)).pin_memory()
    theta_dev = theta_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    scores_all = torch.zeros((world, mpg, 2), dtype=torch.float64, device=dev)
    coef_dev = torch.zeros(pop, dtype=torch.float32, device=dev)
    sign_dev = torch.ones(pop, dtype=torch.float32, device=dev)
    ffma_peak = ops.bench_ffma()

    def generation(gen, host_path):
        """One NES generation. host_path: theta from pinned HOST memory + results read back (e2e)."""
        if host_path:
            out = ev.evaluate(theta_host, gen)
        else:
            ev._theta_dev.copy_(theta_dev)
            ev._keys_dev.copy_(ev._keys_for(gen))
            thetas = ops.nes_perturb(ev._theta_dev, ev.pop, ev.member_lo, ev.n_members, ev.seed, gen, ev.noise_std)
            ops.inner_loop_run(ev.bufs, cfg, thetas, ev.env_index, ev._keys_dev)
            ev._thetas = thetas
            out = None
        return out

    # lane keys for the device-resident path are precomputed (host-side key derivation is not the timed work)
    from learning_environments_b200.rng import lane_keys
    key_cache = {}

    def keys_for(gen):
        if gen not in key_cache:
            k = lane_keys(ev.seed, gen, ev.lane_member, ev.lane_variant, ev.lane_eval)
            key_cache[gen] = torch.from_numpy(k.view(np.int32).copy()).to(dev)
        return key_cache[gen]
    ev._keys_for = keys_for
    total = args.warmup + args.steps
    for g in range(2 * total + 2):
        keys_for(g)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    for g in range(args.warmup):
        generation(g, False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev_pairs = []
    steps_done = 0
    learn_done = 0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        generation(args.warmup + k, False)
        e1.record()
        ev_pairs.append((e0, e1))
        torch.cuda.synchronize()
        res_k = ev.bufs.results()
        steps_done += int(res_k["train_steps"].sum())
        learn_done += int(res_k["learn_iters"].sum())
    barrier()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev_pairs)
    t = torch.tensor([dev_ms, float(steps_done), float(learn_done)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, steps_all, learn_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        steps_all, learn_all = float(steps_done), float(learn_done)
    value = steps_all / (dev_ms * 1e-3)

    # ---------------- end-to-end timing through the host API (e2e) ----------------
    base = total
    for g in range(args.warmup):
        generation(base + g, True)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0
    for k in range(args.steps):
        out = generation(base + args.warmup + k, True)          # H2D theta/keys, kernels, D2H lane results
        e2e_steps += int(out["train_steps"].sum())
        orig, add, sub = ev.member_scores(out)
        local = torch.from_numpy(np.stack([np.maximum(add, sub), orig], 1)).to(dev)
        if world > 1:
            dist.all_gather_into_tensor(scores_all.view(-1), local.view(-1))   # per-generation all-gather of fitness scores
            sc = scores_all.cpu().numpy().reshape(pop, 2)
        else:
            sc = local.cpu().numpy()
        w = score_transform(sc[:, 0], sc[:, 1], gtn["score_transform_type"])
        coef_dev.copy_(torch.from_numpy((gtn["step_size"] * w).astype(np.float32)))
        # every rank regenerates all eps_i and applies them in member order: bit-identical theta on all ranks
        ops.nes_update(theta_dev.clone(), pop, ev.seed, base + args.warmup + k, ev.noise_std, gtn["weight_decay"], coef_dev, sign_dev)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_steps_all = float(a[0]), float(b[1])
    else:
        e2e_steps_all = float(e2e_steps)
    e2e_value = e2e_steps_all / e2e_s

    if rank == 0:
        F = f_step(cfg)
        f_env_q, f_td = f_parts(cfg)
        # per-GPU TFLOP/s of algorithmic work: every env step costs F_env + F_q, every TD update 5*B*F_q (steps of the
        # init_episodes do not learn); the fused kernel is >= 95% of the timed region (profiles/)
        achieved = (steps_all * f_env_q + learn_all * f_td) / world / (dev_ms * 1e-3) / 1e12
        line = {
            "metric": "se_env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.workload, cfg, mpg, plan), **({"overrides": args.lane_override} if args.lane_override else {})),
            "nes_generations_per_hour": 3600.0 / (e2e_s / max(args.steps, 1)), "nes_population": pop,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": ev.h2d_bytes, "d2h_bytes_per_step": ev.d2h_bytes},
            "gpu_launches": 3 * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "fp32_ffma", "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s",
                         "frac": achieved / ffma_peak if ffma_peak else None,
                         "traffic": _ncu_traffic(steps_all / world / max(args.steps, 1)),
                         "flop_per_env_step": F, "td_updates_per_env_step": learn_all / max(steps_all, 1.0), "peak_source": "le_bench_ffma microbenchmark in this run (FP32 FFMA; "
                                                                "MEASURED_PEAKS.json has no FP32 figure)",
                         "hbm": {"algorithmic_bytes_per_env_step": (cfg.batch_size + 1) * (2 * cfg.sd + 3) * 4,
                                 "achieved_gbs": value / world * (cfg.batch_size + 1) * (2 * cfg.sd + 3) * 4 / 1e9,
                                 "peak_gbs": _measured_hbm()}},
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, info = cpu_baseline(cfg, synthetic_theta(cfg), target_seconds=12.0, n_threads=cores)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d lanes of the same lane config, %d steps in %.1f s on %d pthreads (C restatement "
                                              "oracle/le_oracle.c; the reference's own torch path is ~0.8k steps/s/core, BASELINE.md)"
                                              % (info["lanes"], info["steps"], info["seconds"], cores)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _ncu_traffic(env_steps_per_launch):
    """DRAM bytes per launch of the fused kernel from the committed `ncu --set full` capture (profiles/), scaled to
    this run's env steps per launch; None if no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as f:
            t = json.load(f)
        return t["dram_bytes_per_env_step"] * env_steps_per_launch
    except Exception:
        return None


def _measured_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    sys.exit(main())
