"""Helper process of tests/test_gpu_multi.py: runs GTN_Master on the GPU path (the real PopulationEvaluator and NES kernels),
alone or as one rank of an NCCL group (torch.distributed.run sets RANK / LOCAL_RANK / WORLD_SIZE), and saves theta."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from learning_environments_b200 import default_configs, gtn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="replicated")
    ap.add_argument("--out", required=True)
    ap.add_argument("--members", type=int, default=7)
    ap.add_argument("--generations", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)            # same initial SE on every rank
    cfg = default_configs.get("cartpole_syn_env")
    cfg["agents"]["gtn"].update(num_workers=a.members, max_iterations=a.generations, quit_when_solved=False)   # uneven shards
    cfg["agents"]["ddqn"].update(train_episodes=3, test_episodes=2, init_episodes=1)
    m = gtn.GTN_Master(cfg, seed=123, update_mode=a.mode, device=torch.device("cuda", local), verbose=False)
    m.run()
    th = m.theta.numpy()
    scores = np.asarray(m.score_list, np.float64)
    if world > 1:
        np.save(a.out + ".rank%d.npy" % int(os.environ["RANK"]), th)
        np.save(a.out + ".scores.rank%d.npy" % int(os.environ["RANK"]), scores)
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    else:
        np.save(a.out, th)
        np.save(a.out + ".scores.npy", scores)


if __name__ == "__main__":
    main()
