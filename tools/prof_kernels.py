"""Minimal driver for ncu: a few launches of one kernel family.  KERNEL=k2|k3|k4|k1self  (k1 map: tools/prof_k1.py)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
which = os.environ.get("KERNEL", "k2"); prec = int(os.environ.get("PREC", "64")); reps = int(os.environ.get("REPS", "4"))
ctx = api.default_context(0)
ctx.set_option("fk_precision", prec)
st = torch.cuda.current_stream().cuda_stream
S = 1 << 20
if which in ("k2", "k1self"):
    dq = torch.from_numpy(syn.state_batch(S, seed=42)).cuda()
if which == "k2":
    out = torch.empty((S, 7), dtype=torch.float64, device="cuda")
    for _ in range(reps): ctx.fk_dev(dq, S, out, st)
    torch.cuda.synchronize(); print("ok", float(out[:, :3].abs().sum()))
elif which == "k1self":
    w = syn.box_world(g=256, n_boxes=64, seed=1234)
    ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
    dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
    for _ in range(reps): ctx.check_states_dev(dq, S, api.CHECK_MAP | api.CHECK_SELF, dm, st)
    torch.cuda.synchronize(); print("ok", int(np.unpackbits(dm.cpu().numpy().view(np.uint8)).sum()))
else:
    N = int(os.environ.get("N", str(1 << 20))); Q = int(os.environ.get("Q", str(1 << 14)))
    nodes, cost = syn.random_tree(N, 7, 9)
    ctx.tree_reset(0); ctx.tree_append_batch(0, nodes, cost)
    qs = torch.from_numpy(syn.sample_configs(Q, 8)).cuda()
    if which == "k3":
        ids = torch.empty(Q, dtype=torch.int32, device="cuda")
        for _ in range(reps): ctx.nn_nearest_dev(0, qs, Q, ids, st)
        torch.cuda.synchronize(); print("ok", int(ids.sum()))
    else:
        low = torch.empty((Q, 20), dtype=torch.int32, device="cuda"); high = torch.empty_like(low)
        cnt = torch.empty(Q, dtype=torch.int32, device="cuda"); ov = torch.zeros(1, dtype=torch.int32, device="cuda")
        for _ in range(reps): ctx.nn_radius_dev(0, qs, None, Q, 1.5, 20, low, high, cnt, ov, st)
        torch.cuda.synchronize(); print("ok", int(cnt.sum()))
