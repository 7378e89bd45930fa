import os, sys, hashlib
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from squirrel_motion_planner_b200 import api, synthetic as syn
ctx = api.default_context(0)
N, Q = 5_000_003, 4099
nodes, cost = syn.random_tree(N, 11, 12)
qs = syn.sample_configs(Q, 13)
ctx.tree_reset(0); ctx.tree_append_batch(0, nodes, cost)
res = {}
for v in (3, 4):
    ctx.set_option("nn_variant", v)
    ids = ctx.nn_nearest(0, qs)
    lo, hi, cnt = ctx.nn_radius(0, qs, None, 1.2, 20)
    res[v] = hashlib.sha1(ids.tobytes() + lo.tobytes() + hi.tobytes() + cnt.tobytes()).hexdigest()[:12]
print(res, "SAME" if res[3] == res[4] else "DIFFERENT", "mean hits", float(cnt.mean()))
