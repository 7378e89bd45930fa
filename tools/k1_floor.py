"""Cost floor of the validity kernel: same states against an EMPTY grid (stage 1 only: no pending sphere, no exact scan)
and against the config-2 world."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
S = 1 << 20
dq = [torch.from_numpy(syn.state_batch(S, seed=42 + 7919 * i)).cuda() for i in range(4)]
dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
out = {}
w = syn.box_world(g=256, n_boxes=64, seed=1234)
for name in ("world", "empty"):
    bits = w.bits if name == "world" else np.zeros_like(w.bits)
    ctx.set_voxel_grid(bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
    for k in (1, 2):
        ctx.set_option("k1_kernel", k)
        for i in range(5):
            ctx.check_states_dev(dq[i % 4], S, 2, dm, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40):
            ctx.check_states_dev(dq[i % 4], S, 2, dm, st)
        b.record(); torch.cuda.synchronize()
        out["%s_k%d_ms" % (name, k)] = round(a.elapsed_time(b) / 40, 4)
print(json.dumps(out))
