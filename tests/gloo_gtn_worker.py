"""Helper process of tests/test_host_logic.py: runs GTN_Master on CPU with the oracle backend, alone or as one
rank of a gloo group (torch.distributed.run sets RANK / WORLD_SIZE)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from learning_environments_b200 import default_configs, gtn  # noqa: E402
from tests import oracle_backend  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="replicated")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    oracle_backend.patch_master_for_cpu(None)
    torch.manual_seed(0)            # same initial SE on every rank
    cfg = default_configs.get("cartpole_syn_env")
    cfg["agents"]["gtn"].update(num_workers=5, max_iterations=2)      # 5 members: uneven shards (3 + 2)
    cfg["agents"]["ddqn"].update(train_episodes=2, test_episodes=2, init_episodes=1)
    m = gtn.GTN_Master(cfg, seed=123, update_mode=a.mode, evaluator_cls=oracle_backend.OracleEvaluator, verbose=False)
    m.run()
    th = m.theta.numpy()
    if world > 1:
        np.save(a.out + ".rank%d.npy" % int(os.environ["RANK"]), th)
        import torch.distributed as dist
        dist.destroy_process_group()
    else:
        np.save(a.out, th)


if __name__ == "__main__":
    main()
