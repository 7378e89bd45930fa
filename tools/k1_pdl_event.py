"""Does a cudaEventRecord between two K1F launches cancel their programmatic overlap?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
st = torch.cuda.current_stream().cuda_stream
S = 1 << 20
dq = [torch.from_numpy(syn.state_batch(S, seed=100 + i)).cuda() for i in range(4)]
dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
out = {}
side = torch.cuda.Stream()
ev_done = torch.cuda.Event()
with torch.cuda.stream(side):
    ev_done.record(side)
torch.cuda.synchronize()
for mode in ("plain", "event_between", "wait_between", "wait_every_4", "overlap_off"):
    ctx.set_option("launch_overlap", 0 if mode == "overlap_off" else 1)
    evs = [torch.cuda.Event() for _ in range(64)]
    for i in range(5): ctx.check_states_dev(dq[i % 4], S, 2, dm, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(60):
        ctx.check_states_dev(dq[i % 4], S, 2, dm, st)
        if mode == "event_between": evs[i].record()
        if mode == "wait_between" or (mode == "wait_every_4" and i % 4 == 3):
            torch.cuda.current_stream().wait_event(ev_done)     # an (already completed) event of another stream
    b.record(); torch.cuda.synchronize()
    out[mode] = round(a.elapsed_time(b) / 60 * 1e3, 1)
print(json.dumps(out))
