"""Quick GPU sanity + timing script used during development (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn

def ev_time(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]

def main():
    torch.cuda.init(); print(torch.cuda.get_device_name(0))
    ctx = api.default_context(0)
    g = int(os.environ.get("G", "256")); nb = 64 if g == 256 else 2048
    world = syn.box_world(g=g, n_boxes=nb, seed=1234)
    origin = world.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    t0=time.time(); ctx.set_voxel_grid(world.bits, world.dims, origin, world.res); print("set_voxel_grid s", time.time()-t0)
    S = 1 << 20
    q = syn.state_batch(S, seed=42)
    dq = torch.from_numpy(q).cuda(); dm = torch.zeros((S+31)//32, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for prec in (64, 32):
        ctx.set_option("fk_precision", prec)
        for flags, name in ((api.CHECK_MAP, "map"), (api.CHECK_MAP|api.CHECK_SELF, "map+self"), (api.CHECK_SELF, "self")):
            mn, md = ev_time(lambda: ctx.check_states_dev(dq, S, flags, dm, st))
            ncol = int(np.unpackbits(dm.cpu().numpy().view(np.uint8)).sum())
            print(f"K1 prec={prec} {name}: min {mn:.3f} ms med {md:.3f} ms -> {S/mn/1e6:.1f} M states/s, colliding {ncol/S:.3f}, HBM frac {S*64.125/mn/1e6/6544.7:.4f}")
        dout = torch.empty((S, 7), dtype=torch.float64, device="cuda")
        mn, md = ev_time(lambda: ctx.fk_dev(dq, S, dout, st))
        print(f"K2 prec={prec}: min {mn:.3f} ms -> {S/mn/1e6:.1f} M states/s, {S*120/mn/1e6:.0f} GB/s, HBM frac {S*120/mn/1e6/6544.7:.3f}")
    ctx.set_option("fk_precision", 64)
    # edges entry
    E = S // 21
    a, b = syn.edge_batch(E, 42)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dfi = torch.empty(E, dtype=torch.int32, device="cuda")
    mn, md = ev_time(lambda: ctx.check_edges_dev(da, db, E, 20, api.CHECK_MAP, dm, dfi, st))
    print(f"K1 edges: min {mn:.3f} ms -> {E*21/mn/1e6:.1f} M states/s")
    # NN
    N = int(os.environ.get("N", str(1 << 20))); Q = int(os.environ.get("Q", str(1 << 14)))
    nodes, cost = syn.random_tree(N, 7, 9)
    ctx.tree_reset(0); ctx.tree_append_batch(0, nodes, cost)
    qs = torch.from_numpy(syn.sample_configs(Q, 8)).cuda()
    ids = torch.empty(Q, dtype=torch.int32, device="cuda")
    mn, md = ev_time(lambda: ctx.nn_nearest_dev(0, qs, Q, ids, st), n=3, warm=1)
    print(f"K3 nearest N={N} Q={Q}: {mn:.2f} ms -> {Q/mn*1e3:.0f} q/s, {N*Q/mn/1e6:.1f} G pairs/s")
    low = torch.empty((Q, 20), dtype=torch.int32, device="cuda"); high = torch.empty_like(low)
    cnt = torch.empty(Q, dtype=torch.int32, device="cuda"); ov = torch.zeros(1, dtype=torch.int32, device="cuda")
    mn, md = ev_time(lambda: ctx.nn_radius_dev(0, qs, None, Q, 1.5, 20, low, high, cnt, ov, st), n=3, warm=1)
    print(f"K4 radius N={N} Q={Q}: {mn:.2f} ms -> {Q/mn*1e3:.0f} q/s, mean hits {cnt.float().mean().item():.1f} max {cnt.max().item()} ov {ov.item()}")
    for Qs in (1, 32):
        ctx2q = qs[:Qs].contiguous(); i2 = ids[:Qs]
        for Ns in (10000,):
            ctx.tree_reset(1); ctx.tree_append_batch(1, nodes[:Ns], cost[:Ns])
            mn, md = ev_time(lambda: ctx.nn_nearest_dev(1, ctx2q, Qs, i2, st), n=10, warm=3)
            print(f"K3 nearest N={Ns} Q={Qs}: {mn*1e3:.1f} us")
    print("launches", ctx.kernel_launches)

if __name__ == "__main__":
    main()
