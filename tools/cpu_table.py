"""CPU rows of BASELINE.md §3: the oracle port timed on this host's cores (bounded samples of the bench workloads).
usage: python tools/cpu_table.py > gpurun_out/cpu_table.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn, sphere_model


def rate(fn, units, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return units / best


def main():
    nt = oracle.max_threads()
    lt, fx = rd.load_fixture()
    orc = oracle.Oracle(lt, [(n, tuple(c), r) for n, c, r in fx["spheres"]])
    out = {"cores": nt}
    q = syn.state_batch(1 << 16, seed=42)
    out["fk_states_per_s_1t"] = rate(lambda: orc.fk(q[:1 << 14]), 1 << 14)
    N = 1 << 20
    nodes, cost = syn.random_tree(N, 7, 9)
    qs = syn.sample_configs(1 << 16, 8)
    out["nearest_queries_per_s_1t"] = rate(lambda: oracle.nearest(nodes, qs[:32], nthreads=1), 32, reps=2)
    out["nearest_queries_per_s_all"] = rate(lambda: oracle.nearest(nodes, qs[:32 * nt], nthreads=nt), 32 * nt, reps=2)
    out["radius_queries_per_s_1t"] = rate(lambda: oracle.near_topk(nodes, cost, qs[:32], None, 1.5, 20, nthreads=1), 32, reps=2)
    out["radius_queries_per_s_all"] = rate(lambda: oracle.near_topk(nodes, cost, qs[:32 * nt], None, 1.5, 20, nthreads=nt), 32 * nt, reps=2)
    out["note"] = "oracle port (kind: port); nearest / radius: full scan of the 2^20-node tree per query, the reference's loop shape"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
