import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
w = syn.box_world(g=256, n_boxes=64, seed=1234)
ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
S = 1 << 20
E = S // 21
a, b = syn.edge_batch(E, 42)
ha, hb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
hm = torch.zeros((E * 21 + 31) // 32, dtype=torch.int32).pin_memory()
hf = torch.zeros(E, dtype=torch.int32).pin_memory()
for chunk in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21):
    ctx.set_option("pipeline_chunk", chunk)
    for _ in range(3):
        ctx.check_edges(ha, hb, 20, api.CHECK_MAP, out_mask=hm, out_first=hf)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.check_edges(ha, hb, 20, api.CHECK_MAP, out_mask=hm, out_first=hf)
    dt = (time.perf_counter() - t0) / 20
    print("chunk 2^%d: %.3f ms, %.3g states/s" % (int(np.log2(chunk)), dt * 1e3, E * 21 / dt))
