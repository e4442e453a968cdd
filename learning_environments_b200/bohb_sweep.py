"""The reference's third optimisation level as a sweep over GTN_Master runs.

Mirrors `experiments/GTNC_evaluate_cartpole_params.py:16-118` (`ExperimentWrapper`): the configuration space over the NES,
DDQN and SE hyper-parameters (:27-53), the mapping of a sampled configuration onto the yaml config (:55-79) and the objective
(:81-118: three `GTN_Master.run()`s, loss = the total number of generations they needed; an exception scores +inf and is
recorded).  The reference drives it with hpbandster's BOHB (`automl/bohb_optim.py`); hpbandster / ConfigSpace / Pyro4 are not
part of this build, so the space is restated with the same names, bounds, log scales and defaults and the driver is BOHB's
own two model-free ingredients: random sampling (`random_fraction`) and successive halving over the budgets
(`min_budget` 1, `max_budget` 3, `eta` 3 of :17-25).  The reference's objective ignores `budget` (:55, :81-98); here the budget
is the number of GTN repetitions of the evaluation, so the full budget reproduces the reference's three runs.

On this path one NES generation of the yaml's population is one launch of the fused kernel, so every evaluated configuration
uses the whole GPU; with a torch.distributed group the population of each run is sharded over the ranks as usual (gtn.py).
Configurations outside the compiled kernel set (SE hidden_layer 2, PReLU Q-nets) raise inside `compute()` and score +inf with
the error text in `info`, exactly as any failing configuration does in the reference.
"""
import copy
import math
import traceback

import numpy as np

from . import default_configs
from .gtn import GTN_Master

# (name, kind, lower / choices, upper, log, default)  —  experiments/GTNC_evaluate_cartpole_params.py:30-51
SPACE = [
    ("gtn_score_transform_type", "int", 0, 7, False, 7),
    ("gtn_step_size", "float", 0.1, 1.0, True, 0.5),
    ("gtn_mirrored_sampling", "cat", [False, True], None, False, True),
    ("gtn_noise_std", "float", 0.01, 1.0, True, 0.1),
    ("ddqn_init_episodes", "int", 1, 20, True, 10),
    ("ddqn_batch_size", "int", 64, 256, False, 128),
    ("ddqn_gamma", "float", 0.001, 0.1, True, 0.01),
    ("ddqn_lr", "float", 1e-4, 5e-3, True, 1e-3),
    ("ddqn_tau", "float", 0.005, 0.05, True, 0.01),
    ("ddqn_eps_init", "float", 0.8, 1.0, True, 0.9),
    ("ddqn_eps_min", "float", 0.005, 0.05, True, 0.05),
    ("ddqn_eps_decay", "float", 0.01, 0.2, True, 0.1),
    ("ddqn_activation_fn", "cat", ["tanh", "relu", "leakyrelu", "prelu"], None, False, "relu"),
    ("ddqn_hidden_size", "int", 48, 192, True, 128),
    ("ddqn_hidden_layer", "int", 1, 2, False, 2),
    ("cartpole_activation_fn", "cat", ["tanh", "relu", "leakyrelu", "prelu"], None, False, "leakyrelu"),
    ("cartpole_hidden_size", "int", 48, 192, True, 128),
    ("cartpole_hidden_layer", "int", 1, 2, False, 1),
]

# default_config_cartpole.yaml (the file compute() loads, :82-83): the values that differ from default_config_cartpole_syn_env.yaml
_CARTPOLE_YAML_OVERRIDES = {
    "gtn": dict(max_iterations=50, noise_std=0.05, quit_when_solved=True, score_transform_type=7, step_size=1.0, time_max=300.0),
    "ddqn": dict(activation_fn="relu", batch_size=32, eps_decay=0.9, eps_init=1.0, eps_min=0.1, gamma=0.99, hidden_size=64, lr=0.00025,
                 print_rate=100, rb_size=1000000, tau=0.01, test_episodes=1),
    "env": dict(hidden_size=128),
}


def default_cartpole_config():
    """default_config_cartpole.yaml as a config dict (reference schema)."""
    d = default_configs.get("cartpole_syn_env")
    d["agents"]["gtn"].update(_CARTPOLE_YAML_OVERRIDES["gtn"])
    d["agents"]["ddqn"].update(_CARTPOLE_YAML_OVERRIDES["ddqn"])
    d["envs"]["CartPole-v0"].update(_CARTPOLE_YAML_OVERRIDES["env"])
    return d


def default_configuration():
    return {name: default for name, _, _, _, _, default in SPACE}


def sample_configuration(rng):
    """One configuration drawn like ConfigSpace's uniform / log-uniform / integer / categorical hyper-parameters."""
    cso = {}
    for name, kind, lo, hi, log, _ in SPACE:
        if kind == "cat":
            cso[name] = lo[int(rng.randint(len(lo)))]
        elif kind == "float":
            cso[name] = float(math.exp(rng.uniform(math.log(lo), math.log(hi)))) if log else float(rng.uniform(lo, hi))
        else:   # integers: ConfigSpace samples the (log-)uniform float on [lo - 0.5, hi + 0.5) and rounds
            a, b = lo - 0.4999, hi + 0.4999
            v = math.exp(rng.uniform(math.log(a), math.log(b))) if log else rng.uniform(a, b)
            cso[name] = int(min(max(int(round(v)), lo), hi))
    return cso


class ExperimentWrapper(object):
    """experiments/GTNC_evaluate_cartpole_params.py:16-118."""

    def get_bohb_parameters(self):
        return {"min_budget": 1, "max_budget": 3, "eta": 3, "random_fraction": 0.3, "iterations": 10000}

    def get_configspace(self):
        return list(SPACE)

    def get_specific_config(self, cso, default_config, budget):
        config = copy.deepcopy(default_config)
        g, a, e = config["agents"]["gtn"], config["agents"]["ddqn"], config["envs"]["CartPole-v0"]
        g["score_transform_type"] = cso["gtn_score_transform_type"]
        g["step_size"] = cso["gtn_step_size"]
        g["mirrored_sampling"] = cso["gtn_mirrored_sampling"]
        g["noise_std"] = cso["gtn_noise_std"]
        a["init_episodes"] = cso["ddqn_init_episodes"]
        a["batch_size"] = cso["ddqn_batch_size"]
        a["gamma"] = 1 - cso["ddqn_gamma"]
        a["lr"] = cso["ddqn_lr"]
        a["tau"] = cso["ddqn_tau"]
        a["eps_init"] = cso["ddqn_eps_init"]
        a["eps_min"] = cso["ddqn_eps_min"]
        a["eps_decay"] = 1 - cso["ddqn_eps_decay"]
        a["activation_fn"] = cso["ddqn_activation_fn"]
        a["hidden_size"] = cso["ddqn_hidden_size"]
        a["hidden_layer"] = cso["ddqn_hidden_layer"]
        e["activation_fn"] = cso["cartpole_activation_fn"]
        e["hidden_size"] = cso["cartpole_hidden_size"]
        e["hidden_layer"] = cso["cartpole_hidden_layer"]
        return config

    def compute(self, working_dir, bohb_id, config_id, cso, budget, default_config=None, master_cls=GTN_Master, master_kwargs=None,
                repeats=None, **kwargs):
        """loss = total generations of `repeats` GTN runs (the reference: 3; here min(3, budget) unless given), lower is better."""
        config = self.get_specific_config(cso, default_config or default_cartpole_config(), budget)
        n_rep = int(repeats if repeats is not None else max(1, min(3, int(round(budget)))))
        score, score_list, error = 0, [], ""
        try:
            for _ in range(n_rep):
                gtn = master_cls(config, bohb_id=bohb_id, bohb_working_dir=working_dir, **(master_kwargs or {}))
                ret = gtn.run()
                score_list = ret[1]
                score += len(score_list)
        except Exception:   # the reference: bare except, loss = inf, traceback into info
            score = float("inf")
            score_list = []
            error = traceback.format_exc()
        return {"loss": score, "info": {"error": str(error), "config": str(config), "score_list": str(score_list)}}


def run_sweep(n_configs=9, seed=0, working_dir=None, default_config=None, master_cls=GTN_Master, master_kwargs=None, include_default=True,
              verbose=False):
    """One successive-halving bracket over randomly sampled configurations (BOHB's model-free part with the reference's
    parameters: budgets 1 -> 3, eta 3): all `n_configs` at budget 1, the best third again at budget 3.
    Returns the evaluated configurations as a list of dicts (config_id, cso, budget, loss, info), best first."""
    ew = ExperimentWrapper()
    p = ew.get_bohb_parameters()
    rng = np.random.RandomState(seed)
    csos = [default_configuration()] if include_default else []
    while len(csos) < n_configs:
        csos.append(sample_configuration(rng))
    results = []
    budget, alive = float(p["min_budget"]), list(range(len(csos)))
    while True:
        stage = []
        for cid in alive:
            r = ew.compute(working_dir, 0, cid, csos[cid], budget, default_config=default_config, master_cls=master_cls,
                           master_kwargs=master_kwargs)
            stage.append(dict(config_id=cid, cso=csos[cid], budget=budget, loss=r["loss"], info=r["info"]))
            if verbose:
                print("config %d budget %g loss %s %s" % (cid, budget, r["loss"], r["info"]["error"].strip().splitlines()[-1:] or ""))
        results += stage
        if budget >= p["max_budget"] or len(alive) <= 1:
            break
        stage.sort(key=lambda x: x["loss"])
        alive = [s["config_id"] for s in stage[:max(1, len(stage) // p["eta"])] if math.isfinite(s["loss"])]
        if not alive:
            break
        budget = min(budget * p["eta"], float(p["max_budget"]))
    results.sort(key=lambda x: (-x["budget"], x["loss"]))
    return results
