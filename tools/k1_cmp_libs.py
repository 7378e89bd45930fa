"""Bitwise comparison of the validity masks of two builds of the library (A/B of kernel changes) on config 2 and on a
wide batch that leaves the grid.  usage: python tools/k1_cmp_libs.py libA.so libB.so"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 3:
    outs = []
    for lib in sys.argv[1:3]:
        env = dict(os.environ, SQMP_B200_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        if r.returncode:
            print(r.stderr); sys.exit(1)
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    same = all(outs[0][k] == outs[1][k] for k in outs[0])
    diff = {k: (outs[0][k], outs[1][k]) for k in outs[0] if outs[0][k] != outs[1][k]}
    print(json.dumps({"equal": same, "n_keys": len(outs[0]), "diff": diff}))
    sys.exit(0 if same else 2)
sys.path.insert(0, ROOT)
import hashlib
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
res = {}
for g, nb, seed, xy in ((256, 64, 1234, 5.0), (128, 200, 99, 3.0)):
    w = syn.box_world(g=g, n_boxes=nb, seed=seed)
    ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
    for name, S, half in (("in", 1 << 20, xy), ("wide", 1 << 18, xy * 1.6)):
        q = syn.state_batch(S, seed=42, xy_half=half)
        dq = torch.from_numpy(q).cuda(); dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
        for k in (1, 2):
            for flags in (2, 3):
                ctx.set_option("k1_kernel", k)
                ctx.check_states_dev(dq, S, flags, dm, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                m = dm.cpu().numpy()
                res["g%d_%s_k%d_f%d" % (g, name, k, flags)] = hashlib.sha1(m.tobytes()).hexdigest()[:12] + ":%d" % int(np.unpackbits(m.view(np.uint8)).sum())
print(json.dumps(res))
