"""Programmatic dependent launch of the validity kernel: back-to-back calls that reuse ONE mask buffer must leave the
result of the last call (serial semantics), with and without overlap; timing of both."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
st = torch.cuda.current_stream().cuda_stream
out = {}
for lg in (12, 16, 18, 20):
    S = 1 << lg
    dq = [torch.from_numpy(syn.state_batch(S, seed=100 + i)).cuda() for i in range(6)]
    dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
    ref = []
    ctx.set_option("launch_overlap", 0)
    for i in range(6):
        ctx.check_states_dev(dq[i], S, 2, dm, st); torch.cuda.synchronize(); ref.append(dm.cpu().numpy().copy())
    ok = True
    for ov in (1, 0):
        ctx.set_option("launch_overlap", ov)
        for trial in range(20):
            order = np.random.default_rng(trial).permutation(6)
            for i in order:
                ctx.check_states_dev(dq[i], S, 2, dm, st)       # same mask buffer, no sync in between
            torch.cuda.synchronize()
            ok = ok and np.array_equal(dm.cpu().numpy(), ref[order[-1]])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(5): ctx.check_states_dev(dq[i % 6], S, 2, dm, st)
        a.record()
        for i in range(60): ctx.check_states_dev(dq[i % 6], S, 2, dm, st)
        b.record(); torch.cuda.synchronize()
        out["2^%d_overlap%d_us" % (lg, ov)] = round(a.elapsed_time(b) / 60 * 1e3, 1)
    out["2^%d_serial_semantics" % lg] = bool(ok)
# edges entry (k_first_invalid reads the mask right after the kernel) and map+self
E = 40000
a_, b_ = syn.edge_batch(E, 42)
da, db = torch.from_numpy(a_).cuda(), torch.from_numpy(b_).cuda()
dm = torch.zeros((E * 21 + 31) // 32, dtype=torch.int32, device="cuda"); df = torch.empty(E, dtype=torch.int32, device="cuda")
res = []
for ov in (0, 1):
    ctx.set_option("launch_overlap", ov)
    for _ in range(3):
        ctx.check_edges_dev(da, db, E, 20, 3, dm, df, st)
    torch.cuda.synchronize(); res.append((dm.cpu().numpy().copy(), df.cpu().numpy().copy()))
out["edges_equal"] = bool(np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]))
print(json.dumps(out))
