"""Quick timing of the validity kernel on config 2 (map check) for the library named by SQMP_B200_LIB."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
S = 1 << 20
dq = [torch.from_numpy(syn.state_batch(S, seed=42 + 7919 * i)).cuda() for i in range(4)]
dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
out = {"lib": os.environ.get("SQMP_B200_LIB", "default")}
for flags, name in ((2, "map"), (3, "map+self")):
    for k in (1, 2):
        ctx.set_option("k1_kernel", k)
        for i in range(5):
            ctx.check_states_dev(dq[i % 4], S, flags, dm, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40):
            ctx.check_states_dev(dq[i % 4], S, flags, dm, st)
        b.record(); torch.cuda.synchronize()
        out["%s_k%d_ms" % (name, k)] = round(a.elapsed_time(b) / 40, 4)
        ctx.check_states_dev(dq[0], S, flags, dm, st); torch.cuda.synchronize()
        out["%s_k%d_sum" % (name, k)] = int(np.unpackbits(dm.cpu().numpy().view(np.uint8)).sum())
print(json.dumps(out))
