"""Minimal K1 driver for ncu: a few launches of the config-2 workload (map check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
flags = int(os.environ.get("FLAGS", "2"))
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
S = 1 << 20
dq = torch.from_numpy(syn.state_batch(S, seed=42)).cuda()
dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
for _ in range(4):
    ctx.check_states_dev(dq, S, flags, dm, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", int(np.unpackbits(dm.cpu().numpy().view(np.uint8)).sum()))
