"""Times le_td_update (one TD minibatch update per lane) for a general-kernel Q-net: isolates DuelingDDQN.learn.

    python tools/bench_td_general.py [acrobot|cartpole] [n_lanes] [reps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_environments_b200 import config, default_configs, ops  # noqa: E402
from learning_environments_b200._abi import ENV_SE  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "acrobot"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
d = default_configs.get("acrobot_syn_env" if which == "acrobot" else "cartpole_syn_env")
cfg = config.lane_cfg(d, "duelingddqn", ENV_SE)
P = cfg.q_params()
g = torch.Generator(device="cuda").manual_seed(0)
th = (torch.rand((n, P), device="cuda", generator=g) - 0.5) * 0.2
thT = th.clone()
m = torch.zeros_like(th)
v = torch.zeros_like(th)
t = torch.zeros(n, dtype=torch.int32, device="cuda")
rows = torch.rand((n, cfg.batch_size, 2 * cfg.sd + 3), device="cuda", generator=g)
rows[:, :, cfg.sd] = torch.randint(0, cfg.ad, (n, cfg.batch_size), device="cuda").float()
for _ in range(3):
    ops.td_update(cfg, th, thT, m, v, t, rows)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.td_update(cfg, th, thT, m, v, t, rows)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fq = 2 * sum(i * o for i, o in cfg.q_layer_dims())
flop = 5 * cfg.batch_size * fq * n
print("%s dueling: %d lanes, %.3f ms per launch, %.1f us per TD update per lane-slot, %.2f TFLOP/s algorithmic" %
      (which, n, ms, 1e3 * ms / max(n / 296.0, 1.0), flop / ms / 1e9))
