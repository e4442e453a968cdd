import sys; sys.path.insert(0,'.')
import bench
from learning_environments_b200 import ops
d,cfg=bench.build_lane_cfg("cartpole_se_dueling")
print(ops.inner_loop_plan(cfg, 888, 888))
d,cfg=bench.build_lane_cfg("acrobot_se_dueling")
print(ops.inner_loop_plan(cfg, 591, 591))
