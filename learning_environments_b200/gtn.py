"""NES outer loop ("GTN") with the reference's contract, without the file-IO worker protocol.

    agents/GTN_master.py:16-298  GTN_Master(config, bohb_id=-1, bohb_working_dir=None).run()
                                 -> (mean_score, mean_score_orig_list, model_name); score_transform; update_env;
                                 save_good_model / save_model ({'model': state_dict, 'config': config})
    agents/GTN_worker.py:76-254  per member: score_orig, +eps / -eps evaluations, calc_best_score
    agents/GTN.py:14-56          run_gtn_on_single_pc / run_gtn_on_multiple_pcs

What changed: the `num_workers` worker PROCESSES become lanes of one persistent kernel per GPU
(engine.PopulationEvaluator).  The master no longer ships state_dicts through results/GTN_sync: perturbations are
regenerated on device from the Philox (seed, generation, member) stream.  With torch.distributed initialised
(one process per GPU) members are block-sharded over ranks; per generation the ranks all-gather the
(score, score_orig) pairs — the only data-path collective — and then either

    update_mode="replicated" (default): every rank regenerates ALL eps_i and applies them in member order
                                        (bit-identical theta on every rank for any GPU count, no second collective)
    update_mode="allreduce":            each rank forms sum_{i local} w_i eps_i and the ranks all-reduce (sum) it

Checkpoints keep the reference format so its evaluators (experiments/syn_env_evaluate_*_vary_hp_2.py:12-22) load them.
"""
import os
import random
import statistics
import string
import time

import numpy as np
import torch

from . import config as le_config
from . import nes, ops
from ._abi import ENV_RN, ENV_SE
from .engine import PopulationEvaluator
from .envs import EnvFactory, linear_theta, set_linear_theta


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


class GTN_Base(object):
    """agents/GTN_base.py: kept for the file-name scheme (unused by the in-process transport)."""

    def __init__(self, bohb_id):
        self.bohb_id = bohb_id
        self.sync_dir = str(os.path.join(os.getcwd(), 'results/GTN_sync'))

    def get_input_file_name(self, id):
        return os.path.join(self.sync_dir, str(self.bohb_id) + '_' + str(id) + '_input.pt')

    def get_input_check_file_name(self, id):
        return os.path.join(self.sync_dir, str(self.bohb_id) + '_' + str(id) + '_input_check.pt')

    def get_result_file_name(self, id):
        return os.path.join(self.sync_dir, str(self.bohb_id) + '_' + str(id) + '_result.pt')

    def get_result_check_file_name(self, id):
        return os.path.join(self.sync_dir, str(self.bohb_id) + '_' + str(id) + '_result_check.pt')

    def get_quit_file_name(self):
        return os.path.join(self.sync_dir, 'quit.flag')


class GTN_Master(GTN_Base):
    def __init__(self, config, bohb_id=-1, bohb_working_dir=None, seed=None, update_mode="replicated", step_budget=0,
                 evaluator_cls=PopulationEvaluator, device=None, verbose=True):
        super().__init__(bohb_id)
        self.config = config
        self.device = config["device"]
        self.env_name = config['env_name']
        g = config["agents"]["gtn"]
        self.max_iterations = g["max_iterations"]
        self.agent_name = g["agent_name"]
        self.num_workers = g["num_workers"]
        self.noise_std = g["noise_std"]
        self.step_size = g["step_size"]
        self.nes_step_size = g["nes_step_size"]
        self.mirrored_sampling = g["mirrored_sampling"]
        self.num_grad_evals = g["num_grad_evals"]
        self.grad_eval_type = g["grad_eval_type"]
        self.weight_decay = g["weight_decay"]
        self.score_transform_type = g["score_transform_type"]
        self.time_mult = g["time_mult"]
        self.time_max = g["time_max"]
        self.quit_when_solved = g["quit_when_solved"]
        self.synthetic_env_type = g["synthetic_env_type"]
        self.unsolved_weight = g["unsolved_weight"]
        self.update_mode = update_mode
        self.verbose = verbose
        if update_mode not in ("replicated", "allreduce"):
            raise ValueError("update_mode must be 'replicated' or 'allreduce'")
        if self.agent_name.lower() not in ("ddqn", "duelingddqn"):
            raise NotImplementedError("GTN inner-loop agent %r is outside the B200 hot path (DDQN / DuelingDDQN are built)" % self.agent_name)

        self.time_elapsed_list = [None] * self.num_workers
        self.score_list = [None] * self.num_workers
        self.score_orig_list = [None] * self.num_workers
        self.score_transform_list = [None] * self.num_workers
        self.sign_list = [1.0] * self.num_workers

        self.env_factory = EnvFactory(config)
        if self.synthetic_env_type == 0:
            self.synthetic_env_orig = self.env_factory.generate_virtual_env(print_str='GTN_Base: ')
            env_kind = ENV_SE
        elif self.synthetic_env_type == 1:
            self.synthetic_env_orig = self.env_factory.generate_reward_env(print_str='GTN_Base: ')
            env_kind = ENV_RN
        else:
            raise NotImplementedError("Unknown synthetic_env_type value: " + str(self.synthetic_env_type))
        self.real_env = self.env_factory.generate_real_env()

        if bohb_working_dir:
            self.model_dir = str(os.path.join(bohb_working_dir, 'GTN_models_' + self.env_name))
        else:
            self.model_dir = str(os.path.join(os.getcwd(), "results", 'GTN_models_' + self.env_name))
        self.model_name = self.get_model_file_name(
            self.env_name + '_' + ''.join(random.choices(string.ascii_uppercase + string.digits, k=6)) + '.pt')
        self.best_score = -float('Inf')
        os.makedirs(self.model_dir, exist_ok=True)

        # ---- device side: this rank's shard of the population ------------------------------------------------
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        per = (self.num_workers + self.world - 1) // self.world
        self.member_lo = min(self.rank * per, self.num_workers)
        self.member_hi = min(self.member_lo + per, self.num_workers)
        self.seed = int(seed if seed is not None else random.getrandbits(31))
        if dist and seed is None:   # all ranks must draw the same perturbations
            t = torch.tensor([self.seed], dtype=torch.int64, device=self._coll_device(device))
            dist.broadcast(t, 0)
            self.seed = int(t.item())
        inner = self.agent_name.lower()
        gamma = config["agents"][inner]["gamma"]
        self.lane_cfg = le_config.lane_cfg(config, inner, env_kind, use_test_env=True, final_test=True, step_budget=step_budget,
                                           gamma=gamma)
        slopes = self.synthetic_env_orig.env.lane_cfg_fields()["env_slope"]
        for i in range(3):
            self.lane_cfg.env_slope[i] = slopes[i]
        self._device = device
        self.evaluator = None
        if self.member_hi > self.member_lo:
            self.evaluator = evaluator_cls(self.lane_cfg, self.num_workers, self.member_lo, self.member_hi,
                                           num_grad_evals=self.num_grad_evals, seed=self.seed, noise_std=self.noise_std,
                                           mirrored=self.mirrored_sampling, **({"device": device} if device is not None else {}))
        self.theta = linear_theta(self.synthetic_env_orig.env).cpu().contiguous()
        self.generation = 0
        self._generation_base = 0     # generations of earlier run() calls: a second run() continues the Philox generation counter
        if self.verbose and self.rank == 0:
            print('Starting GTN Master with bohb_id {} ({} members on {} rank(s))'.format(bohb_id, self.num_workers, self.world))

    def _coll_device(self, device=None):
        dist = _dist()
        if dist and dist.get_backend() == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def get_model_file_name(self, file_name):
        return os.path.join(self.model_dir, file_name)

    # ------------------------------------------------------------------------------------------------------
    def evaluate_population(self):
        """One generation of fitness evaluations (what write_worker_inputs/read_worker_results wrap in the
        reference, agents/GTN_master.py:147-195): fills score_list / score_orig_list / sign_list for ALL members."""
        t0 = time.time()
        local = np.zeros((max(self.member_hi - self.member_lo, 0), 3))
        if self.evaluator is not None:
            out = self.evaluator.evaluate(self.theta, self.generation)
            orig, add, sub = self.evaluator.member_scores(out, self.grad_eval_type)
            best, sign = nes.best_of_mirror(add, sub, self.mirrored_sampling)
            local = np.stack([best, orig, sign], axis=1)
        dist = _dist()
        if dist:
            per = (self.num_workers + self.world - 1) // self.world
            dev = self._coll_device()
            mine = torch.zeros((per, 3), dtype=torch.float64, device=dev)
            mine[:local.shape[0]] = torch.from_numpy(local).to(dev)
            gathered = torch.zeros((self.world * per, 3), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(gathered, mine)     # the per-generation all-gather of fitness scores
            allv = gathered.cpu().numpy()[:self.num_workers] if self.world * per == self.num_workers else \
                np.concatenate([gathered.cpu().numpy()[r * per:r * per + max(0, min(per, self.num_workers - r * per))]
                                for r in range(self.world)])
        else:
            allv = local
        self.score_list = [float(v) for v in allv[:, 0]]
        self.score_orig_list = [float(v) for v in allv[:, 1]]
        self.sign_list = [float(v) for v in allv[:, 2]]
        self.time_elapsed_list = [time.time() - t0] * self.num_workers

    def score_transform(self):
        self.score_transform_list = nes.score_transform(self.score_list, self.score_orig_list, self.score_transform_type).tolist()

    def update_env(self):
        """agents/GTN_master.py:267-298 on device: weight decay, then theta += ss * w_i * eps_i for i = 0..pop-1."""
        ss = self.step_size / self.num_workers if self.nes_step_size else self.step_size
        coef = np.array([np.float32(ss * w) for w in self.score_transform_list], dtype=np.float32)
        sign = np.asarray(self.sign_list, dtype=np.float32)
        dev = self._nes_device()
        theta_dev = self.theta.to(dev)
        coef_dev = torch.from_numpy(coef).to(dev)
        sign_dev = torch.from_numpy(sign).to(dev)
        dist = _dist()
        if self.update_mode == "replicated" or not dist:
            ops.nes_update(theta_dev, self.num_workers, self.seed, self.generation, self.noise_std, self.weight_decay, coef_dev, sign_dev)
        else:
            delta = ops.nes_partial_update(theta_dev.numel(), self.member_lo, self.member_hi, self.seed, self.generation,
                                           self.noise_std, coef_dev, sign_dev)
            if dist.get_backend() != "nccl":
                delta = delta.cpu()
            dist.all_reduce(delta)                           # all-reduce (sum) of the weighted update
            theta_dev = (theta_dev * np.float32(1 - self.weight_decay)) + delta.to(dev)
        self.theta = theta_dev.cpu().contiguous()
        set_linear_theta(self.synthetic_env_orig.env, self.theta)

    def _nes_device(self):
        return torch.device("cuda", torch.cuda.current_device())

    # ------------------------------------------------------------------------------------------------------
    def run(self):
        mean_score_orig_list = []
        for it in range(self.max_iterations):
            t1 = time.time()
            self.generation = self._generation_base + it
            self.evaluate_population()
            mean_score = np.mean(self.score_orig_list)
            mean_score_orig_list.append(mean_score)
            solved_flag = self.save_good_model(mean_score)
            if solved_flag and self.quit_when_solved:
                if self.verbose and self.rank == 0:
                    print('ENV SOLVED')
                break
            self.score_transform()
            self.update_env()
            self.print_statistics(it=it, time_elapsed=time.time() - t1)
        self._generation_base = self.generation + 1 if self.max_iterations > 0 else self._generation_base
        if self.verbose and self.rank == 0:
            print('Master quitting')
        if len(mean_score_orig_list) > 0:
            return np.mean(self.score_orig_list), mean_score_orig_list, self.model_name
        return 1e9, mean_score_orig_list, self.model_name

    def save_good_model(self, mean_score):
        if self.synthetic_env_orig.is_virtual_env():
            if mean_score > self.real_env.get_solved_reward() and mean_score > self.best_score:
                self.save_model()
                self.best_score = mean_score
                return True
        else:
            if mean_score > self.best_score:
                self.save_model()
                self.best_score = mean_score
        return False

    def save_model(self):
        if self.rank != 0:
            return
        save_dict = {'model': self.synthetic_env_orig.state_dict(), 'config': self.config}
        save_path = os.path.join(self.model_dir, self.model_name)
        if self.verbose:
            print('save model: ' + str(save_path))
        torch.save(save_dict, save_path)

    def calc_worker_timeout(self):
        """agents/GTN_master.py: `time_max` for the first generation, then mean(worker wall time) * time_mult.  Kept for the
        contract; NOT fed back into the lanes: the figure is the wall time of a CPU worker process, and the lanes of one launch
        finish together in milliseconds, so a budget derived from it would cut every agent off.  A fixed per-agent budget is the
        constructor's `step_budget` (env steps)."""
        if self.time_elapsed_list[0] is None:
            return self.time_max
        return statistics.mean(self.time_elapsed_list) * self.time_mult

    def print_statistics(self, it, time_elapsed):
        if not self.verbose or self.rank != 0:
            return
        print('--------------')
        print('GTN iteration:    ' + str(it))
        print('GTN mstr t_elaps: ' + str(time_elapsed))
        print('GTN avg eval score:   ' + str(statistics.mean(self.score_orig_list)))
        print('--------------')


class GTN_Worker(GTN_Base):
    """agents/GTN_worker.py: in this build a worker is a set of lanes of the master's persistent kernel, so there is
    no worker process to start.  The class keeps calc_score() — one complete fitness evaluation of an environment."""

    def __init__(self, id, bohb_id=-1):
        super().__init__(bohb_id)
        self.id = id

    def run(self):
        print('GTN_Worker %d: workers are lanes of GTN_Master\'s fused kernel in this build; nothing to run.' % self.id)

    def calc_score(self, env, config, time_remaining=1e9):
        from .agents import select_agent
        agent = select_agent(config=config, agent_name=config["agents"]["gtn"]["agent_name"])
        real_env = EnvFactory(config).generate_real_env()
        agent.train(env=env, test_env=real_env, time_remaining=time_remaining)
        reward_list_test, _, _ = agent.test(env=real_env, time_remaining=time_remaining)
        return statistics.mean(reward_list_test)


def run_gtn_on_single_pc(config, bohb_id=-1, **kw):
    """agents/GTN.py:14-44: master + workers on one machine == one GTN_Master on one GPU."""
    return GTN_Master(config, bohb_id=bohb_id, **kw).run()


def run_gtn_on_multiple_pcs(config, bohb_id=-1, **kw):
    """agents/GTN.py:47-56: with torch.distributed initialised (torchrun, one process per GPU) the same call shards
    the population over the ranks."""
    return GTN_Master(config, bohb_id=bohb_id, **kw).run()
