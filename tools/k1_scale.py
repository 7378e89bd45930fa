"""Validity kernel time against the batch size (fixed overhead vs slope)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
Smax = 1 << 24
dq = torch.from_numpy(syn.state_batch(Smax, seed=42)).cuda()
dm = torch.zeros(Smax // 32, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
out = {}
for lg in (16, 18, 19, 20, 21, 22, 24):
    S = 1 << lg
    n_off = Smax // S
    reps = max(8, min(64, (1 << 26) // S))
    for i in range(3):
        ctx.check_states_dev(dq[(i % n_off) * S:], S, 2, dm, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        ctx.check_states_dev(dq[(i % n_off) * S:], S, 2, dm, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out["2^%d" % lg] = {"ms": round(ms, 4), "us_per_2^20": round(ms * 1e3 * (1 << 20) / S, 1)}
print(json.dumps(out))
