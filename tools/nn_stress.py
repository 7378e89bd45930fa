"""Randomised cross-check of the tensor-core near-neighbour scan (variant 4) against the fp32-pre-filter scan (variant 3):
random tree and query sizes (ragged tiles, several chunks), clustered and duplicated nodes, random radii, each call repeated
(a race in the ring / barrier protocol would show as a run-to-run difference).  usage (GPU box): python tools/nn_stress.py [trials]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from squirrel_motion_planner_b200 import api, synthetic as syn

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctx = api.default_context(0)
rng = np.random.default_rng(4242)
bad, pairs = [], 0
for t in range(trials):
    N = int(rng.choice([1, 7, 127, 128, 129, 4095, 4096, 4097, int(rng.integers(1000, 400000))]))
    Q = int(rng.choice([1, 255, 256, 257, 2048, int(rng.integers(1, 6000))]))
    nodes, cost = syn.random_tree(N, seed=100 + t, cost_seed=200 + t)
    if N > 300 and t % 3 == 0:                      # clusters + exact duplicates: ties and near-ties
        c = nodes[rng.integers(0, N, 20)]
        idx = rng.integers(0, N, N // 4)
        nodes[idx] = c[rng.integers(0, 20, len(idx))] + rng.normal(size=(len(idx), 8)) * float(rng.choice([0.0, 1e-12, 1e-6, 1e-3]))
    qs = syn.sample_configs(Q, seed=300 + t)
    if N > 10:
        qs[: min(Q, 10)] = nodes[rng.integers(0, N, min(Q, 10))]
    r = float(rng.choice([0.3, 1.0, 1.5, 2.5]))
    k = int(rng.choice([1, 5, 20]))
    ctx.tree_reset(0)
    ctx.tree_append_batch(0, nodes, cost)
    ref = None
    for variant, reps in ((3, 1), (4, 3)):
        ctx.set_option("nn_variant", variant)
        for _ in range(reps):
            got = (ctx.nn_nearest(0, qs),) + tuple(ctx.nn_radius(0, qs, None, r, k))
            if ref is None:
                ref = got
            elif not all(np.array_equal(a, b) for a, b in zip(ref, got)):
                bad.append({"trial": t, "N": N, "Q": Q, "r": r, "k": k, "variant": variant})
    pairs += N * Q
print(json.dumps({"trials": trials, "pairs": pairs, "mismatches": bad}))
