"""Development aid: time the validity kernel on config 2 (and optionally W512) and check the fast kernel against the
generic one bit for bit.  usage: python tools/k1f_quick.py [nocheck]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn


def t_ms(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ctx = api.default_context(0)
    st = torch.cuda.current_stream().cuda_stream
    S = 1 << 20
    for g, nb, seed, half in ((256, 64, 1234, 5.0), (512, 2048, 4321, 12.0)):
        w = syn.box_world(g=g, n_boxes=nb, seed=seed)
        ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET]), w.res)
        dq = [torch.from_numpy(syn.state_batch(S, seed=42 + i, xy_half=half)).cuda() for i in range(4)]
        dm = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
        ref = None
        if "nocheck" not in sys.argv:
            ctx.set_option("k1_kernel", 1)
            ctx.check_states_dev(dq[0], S, api.CHECK_MAP | api.CHECK_SELF, dm, st)
            torch.cuda.synchronize()
            ref2 = dm.clone()
            ms1 = t_ms(lambda: ctx.check_states_dev(dq[0], S, api.CHECK_MAP, dm, st), n=3, warm=1)
            ref = dm.clone()
            ctx.set_option("k1_kernel", 2)
            ctx.check_states_dev(dq[0], S, api.CHECK_MAP, dm, st)
            torch.cuda.synchronize()
            same = bool(torch.equal(ref, dm))
            ctx.check_states_dev(dq[0], S, api.CHECK_MAP | api.CHECK_SELF, dm, st)
            torch.cuda.synchronize()
            same2 = bool(torch.equal(ref2, dm))
            print("W%d generic K1 %.3f ms; fast == generic: map %s, map+self %s" % (g, ms1, same, same2))
        ctx.set_option("k1_kernel", 0)
        i = [0]

        def step(flags=api.CHECK_MAP):
            ctx.check_states_dev(dq[i[0] % 4], S, flags, dm, st)
            i[0] += 1
        for ov in (0, 1):
            ctx.set_option("launch_overlap", ov)
            print("W%d K1F map      overlap=%d  %.4f ms / 2^20 states" % (g, ov, t_ms(step)))
        ctx.set_option("launch_overlap", 0)
        print("W%d K1F map+self            %.4f ms" % (g, t_ms(lambda: step(api.CHECK_MAP | api.CHECK_SELF), n=5)))
        print("W%d colliding fraction %.3f" % (g, float(np.unpackbits(dm.cpu().numpy().view(np.uint8)).mean())))


if __name__ == "__main__":
    main()
