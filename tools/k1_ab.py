"""A/B of the two validity kernels on the bench workload (config 2): masks must be identical bit for bit; timings with
CUDA events on device-resident inputs rotating over 4 batches (> L2).  Usage (GPU box): python tools/k1_ab.py [--self]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--boxes", type=int, default=64)
    args = ap.parse_args()
    import torch
    from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn

    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = api.default_context(0)
    w = syn.box_world(g=args.grid, n_boxes=args.boxes, seed=1234 if args.grid == 256 else 4321)
    origin = w.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    S = args.states
    xy = 5.0 if args.grid == 256 else 11.0
    batches = [syn.state_batch(S, seed=42 + 7919 * i, xy_half=xy) for i in range(4)]
    # a batch that leaves the grid on every side (uncomfortable / forced lanes)
    wide = syn.state_batch(1 << 16, seed=5, xy_half=xy * 1.6)
    d_q = [torch.from_numpy(b).to(dev) for b in batches]
    d_wide = torch.from_numpy(wide).to(dev)
    words = (S + 31) // 32
    out = {}

    def run(kernel, flags, dq, n):
        ctx.set_option("k1_kernel", kernel)
        m = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
        ctx.check_states_dev(dq, n, flags, m, stream)
        torch.cuda.synchronize()
        return m.cpu().numpy()

    def t_ms(kernel, flags, edges=None):
        ctx.set_option("k1_kernel", kernel)
        m = torch.zeros(words, dtype=torch.int32, device=dev)
        if edges is None:
            fn = lambda i: ctx.check_states_dev(d_q[i % 4], S, flags, m, stream)
        else:
            da, db, E, first = edges
            fn = lambda i: ctx.check_edges_dev(da, db, E, 20, flags, m, first, stream)
        for i in range(5):
            fn(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.reps

    for name, flags in (("map", api.CHECK_MAP), ("map+self", api.CHECK_MAP | api.CHECK_SELF)):
        g = run(1, flags, d_q[0], S)
        f = run(2, flags, d_q[0], S)
        out[name + "_mask_equal"] = bool(np.array_equal(g, f))
        out[name + "_diff_bits"] = int(np.unpackbits((g ^ f).view(np.uint8)).sum())
        out[name + "_colliding"] = int(np.unpackbits(g.view(np.uint8)).sum())
        gw = run(1, flags, d_wide, len(wide))
        fw = run(2, flags, d_wide, len(wide))
        out[name + "_wide_equal"] = bool(np.array_equal(gw, fw))
        out[name + "_wide_colliding"] = int(np.unpackbits(gw.view(np.uint8)).sum())
        out[name + "_generic_ms"] = t_ms(1, flags)
        out[name + "_fast_ms"] = t_ms(2, flags)
    E = S // 21
    a, b = syn.edge_batch(E, 42)
    da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    first = torch.empty(E, dtype=torch.int32, device=dev)
    ctx.set_option("k1_kernel", 1)
    m1 = torch.zeros(words, dtype=torch.int32, device=dev)
    ctx.check_edges_dev(da, db, E, 20, api.CHECK_MAP, m1, first, stream)
    f1 = first.cpu().numpy().copy()
    ctx.set_option("k1_kernel", 2)
    m2 = torch.zeros(words, dtype=torch.int32, device=dev)
    ctx.check_edges_dev(da, db, E, 20, api.CHECK_MAP, m2, first, stream)
    torch.cuda.synchronize()
    out["edges_equal"] = bool(np.array_equal(m1.cpu().numpy()[: E * 21 // 32], m2.cpu().numpy()[: E * 21 // 32]) and
                              np.array_equal(f1, first.cpu().numpy()))
    out["edges_generic_ms"] = t_ms(1, api.CHECK_MAP, (da, db, E, first))
    out["edges_fast_ms"] = t_ms(2, api.CHECK_MAP, (da, db, E, first))
    ctx.set_option("k1_kernel", 1)
    ctx.set_option("fk_precision", 32)
    out["map_generic_fp32_ms"] = t_ms(1, api.CHECK_MAP)
    ctx.set_option("fk_precision", 64)
    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
