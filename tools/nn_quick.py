"""Timing + checksum of the NN kernels (N = 2^20 nodes, Q = 2^16 queries) for the library named by SQMP_B200_LIB."""
import os, sys, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from squirrel_motion_planner_b200 import api, synthetic as syn
ctx = api.default_context(0)
ctx.set_option("nn_variant", int(os.environ.get("NN_VARIANT", "0")))
st = torch.cuda.current_stream().cuda_stream
N, Q = 1 << 20, 1 << 16
nodes, cost = syn.random_tree(N, 7, 9)
ctx.tree_reset(0); ctx.tree_append_batch(0, nodes, cost)
d_qs = torch.from_numpy(syn.sample_configs(Q, 8)).cuda()
d_ids = torch.empty(Q, dtype=torch.int32, device="cuda")
d_low = torch.empty((Q, 20), dtype=torch.int32, device="cuda"); d_high = torch.empty_like(d_low)
d_cnt = torch.empty(Q, dtype=torch.int32, device="cuda"); d_ov = torch.zeros(1, dtype=torch.int32, device="cuda")
def t_ms(fn, n=3, warm=1):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
out = {}
out["nearest_ms"] = round(t_ms(lambda: ctx.nn_nearest_dev(0, d_qs, Q, d_ids, st)), 3)
out["radius_ms"] = round(t_ms(lambda: ctx.nn_radius_dev(0, d_qs, None, Q, 1.5, 20, d_low, d_high, d_cnt, d_ov, st)), 3)
out["nearest_q_per_s"] = round(Q / out["nearest_ms"] * 1e3)
out["ids_sha"] = hashlib.sha1(d_ids.cpu().numpy().tobytes()).hexdigest()[:12]
out["radius_sha"] = hashlib.sha1(d_low.cpu().numpy().tobytes() + d_high.cpu().numpy().tobytes() + d_cnt.cpu().numpy().tobytes()).hexdigest()[:12]
out["mean_hits"] = round(float(d_cnt.float().mean()), 1)
print(json.dumps(out))
