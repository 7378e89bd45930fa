"""Randomised cross-check of the two validity kernels: worlds of different resolution, shape, origin and clutter, batches
that stay inside the grid and batches that leave it on every side; masks must be identical bit for bit (map, map+self,
edge entry).  usage (GPU box): python tools/k1_stress.py [n_worlds]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn

n_worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = api.default_context(0)
rng = np.random.default_rng(2026)
bad, total_states, total_coll = [], 0, 0
for wi in range(n_worlds):
    res = float(rng.choice([0.025, 0.04, 0.05, 0.08, 0.1]))
    lo_cells = int(np.ceil(2.0 / res))               # box_world wants at least ~1 m of free floor in x and y
    dims = (int(rng.integers(lo_cells, lo_cells + 160)), int(rng.integers(lo_cells, lo_cells + 160)), int(rng.integers(24, 160)))
    if dims[0] * dims[1] * dims[2] > 6_000_000:
        dims = (dims[0], dims[1], max(24, 6_000_000 // (dims[0] * dims[1])))
    w = syn.box_world(g=64, res=res, n_boxes=int(rng.integers(0, 80)), seed=1000 + wi, dims=dims,
                      z_low=float(rng.uniform(-0.5, 0.0)))
    origin = w.origin + rng.uniform(-0.3, 0.3, size=3) + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    try:
        ctx.set_voxel_grid(w.bits, w.dims, origin, res)
    except api.SqmpError as e:                       # e.g. sphere radius too large for a very fine grid
        print("world %d skipped: %s" % (wi, e)); continue
    half = min(dims[0], dims[1]) * res / 2.0
    for wide in (0.8, 1.5):
        S = 21 * 6000 + int(rng.integers(0, 40))
        q = syn.state_batch(S, seed=5000 + wi, xy_half=max(0.3, half * wide))
        for flags in (api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF):
            ctx.set_option("k1_kernel", 1); a = ctx.check_states(q, flags)
            ctx.set_option("k1_kernel", 2); b = ctx.check_states(q, flags)
            total_states += S; total_coll += int(np.unpackbits(a.view(np.uint8)).sum())
            if not np.array_equal(a, b):
                bad.append({"world": wi, "res": res, "dims": dims, "wide": wide, "flags": flags,
                            "diff_bits": int(np.unpackbits((a ^ b).view(np.uint8)).sum())})
    ea, eb = syn.edge_batch(3000, seed=7000 + wi, xy_half=max(0.3, half * 1.2))
    ctx.set_option("k1_kernel", 1); m1, f1 = ctx.check_edges(ea, eb, 20, api.CHECK_MAP)
    ctx.set_option("k1_kernel", 2); m2, f2 = ctx.check_edges(ea, eb, 20, api.CHECK_MAP)
    if not (np.array_equal(m1, m2) and np.array_equal(f1, f2)):
        bad.append({"world": wi, "edges": True})
ctx.set_option("k1_kernel", 0)
print(json.dumps({"worlds": n_worlds, "states_compared": total_states, "colliding": total_coll, "mismatches": bad}))
sys.exit(1 if bad else 0)
