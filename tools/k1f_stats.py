"""Development aid: where the work of the fast validity kernel's bound hierarchy goes, on the config-2 workload.
Needs the counter build:  tools/build_variant.sh stats -DSQMP_KF_STATS   then   SQMP_B200_LIB=build/variants/libsqmp_stats.so
python tools/k1f_stats.py [world: 256|512]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn

NAMES = ["group root tests (warp)", "open roots (lanes)", "groups entered (warp)", "bound batches (warp)", "open clusters (lanes)",
         "clusters some lane opens (warp)", "spheres pending after stage 1", "stage-2 rounds (warp)", "stage-2 tests",
         "exact-scan items", "queue drains (warp)", "warps", "colliding states", "cluster rounds (warp)"]


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ctx = api.default_context(0)
    w = syn.box_world(g=g, n_boxes=64 if g == 256 else 2048, seed=1234 if g == 256 else 4321)
    ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
    S = 1 << 20
    dq = torch.from_numpy(syn.state_batch(S, seed=42, xy_half=5.0 if g == 256 else 12.0)).cuda()
    dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
    L = api.load_library()
    out = (C.c_ulonglong * 16)()
    L.sqmp_debug_kf_stats(out, 1)
    ctx.check_states_dev(dq, S, api.CHECK_MAP, dm, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    L.sqmp_debug_kf_stats(out, 1)
    warps = out[11]
    for i, n in enumerate(NAMES):
        print("%-42s %12d   per warp %8.2f   per state %8.3f" % (n, out[i], out[i] / warps, out[i] / S))


if __name__ == "__main__":
    main()
