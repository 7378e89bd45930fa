"""e2e (host pinned buffers -> sqmp_check_states -> host mask) against the pipeline chunk size."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
S = 1 << 20
h_q = torch.from_numpy(syn.state_batch(S, seed=42)).pin_memory()
h_mask = torch.zeros(S // 32, dtype=torch.int32).pin_memory()
out = {}
for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
    ctx.set_option("pipeline_chunk", chunk)
    for _ in range(3):
        ctx.check_states(h_q, 2, out=h_mask)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.check_states(h_q, 2, out=h_mask)
    dt = (time.perf_counter() - t0) / 20
    out[str(chunk)] = round(S / dt / 1e6, 1)
print(json.dumps(out))
