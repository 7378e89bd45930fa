"""Small near-neighbour run through the tensor-core scan (variant 4) for compute-sanitizer; checks the ids against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from squirrel_motion_planner_b200 import api, synthetic as syn
ctx = api.default_context(0)
ctx.set_option("nn_variant", 4)
N, Q = 20011, 2500
nodes, cost = syn.random_tree(N, 3, 4)
qs = syn.sample_configs(Q, 5)
ctx.tree_reset(0); ctx.tree_append_batch(0, nodes, cost)
ok = np.array_equal(ctx.nn_nearest(0, qs), oracle.nearest(nodes, qs, nthreads=8))
got = ctx.nn_radius(0, qs, None, 1.5, 20)
ref = oracle.near_topk(nodes, cost, qs, None, 1.5, 20, nthreads=8)
ok = ok and all(np.array_equal(a, b) for a, b in zip(got, ref))
print("nn_small:", "OK" if ok else "MISMATCH")
