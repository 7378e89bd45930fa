"""Writes the work directory the C++ planner driver (shim/sqmp_plan) reads: world grid, planning queries and — for the
CPU-oracle backend of the test / baseline binary — the oracle's link model.  No arithmetic, only file layout."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squirrel_motion_planner_b200 import robot_desc as rd  # noqa: E402


def write_world(d, world):
    world.bits.astype(np.uint32).tofile(os.path.join(d, "grid_bits.u32"))
    with open(os.path.join(d, "grid_meta.txt"), "w") as f:     # raw origin: the shim applies the -0.02 of CC:87 itself
        f.write("%d %d %d %.17g %.17g %.17g %.17g\n" % (*world.dims, *world.origin, world.res))


def write_queries(d, starts, goals):
    np.ascontiguousarray(np.concatenate([starts, goals], axis=1), dtype=np.float64).tofile(os.path.join(d, "queries.f64"))


def write_oracle_model(d, lt, spheres):
    """oracle_model.bin: the argument list of orc_model_create (oracle/sqmp_oracle.h), raw little-endian."""
    sph_link = np.array([lt.index(n) for n, _, _ in spheres], dtype=np.int32)
    sph_c = np.array([c for _, c, _ in spheres], dtype=np.float64).reshape(-1, 3)
    sph_r = np.array([r for _, _, r in spheres], dtype=np.float64)
    with open(os.path.join(d, "oracle_model.bin"), "wb") as f:
        f.write(np.array([lt.n, len(sph_r)], dtype=np.int32).tobytes())
        for a, t in ((lt.parent, np.int32), (lt.joint_type, np.int32), (lt.qidx, np.int32), (lt.joint_origin, np.float64),
                     (lt.joint_axis, np.float64), (lt.to_parent, np.float64), (lt.frame_to_tip, np.float64),
                     (lt.shaped, np.int32), (lt.self_pair_matrix(), np.uint8), (sph_link, np.int32), (sph_c, np.float64),
                     (sph_r, np.float64)):
            f.write(np.ascontiguousarray(a, dtype=t).tobytes())
        f.write(np.array([lt.index("hand_wrist_link")], dtype=np.int32).tobytes())   # FK chain tip (SRDF:4-6)


def folded_arm():
    return np.array([-0.7, 1.9, 0.0, 1.7, 0.0])       # pose_folded_arm, squirrel_8dof_planner/config/parameters.yaml:37


def half_folded_arm():
    return np.array([-0.17, 1.96, 0.19, 0.65, 0.53])   # folding key-frame 13 (folding_poses_tuw-robotino2.yaml:13)


def config1_queries(n=1):
    """start (-4,-4,0, arm half folded), goal (4,4,pi/2, arm half unfolded).  (The survey's own
    arm poses — pose_folded_arm and the first folding key-frame — rest the hand ON the shell: with a sphere model that
    covers the URDF's front shell box they are in self collision, and init_planner rejects them like BS:351-358
    rejects any invalid end point; tests/test_oracle_golden.py shows the URDF box itself contains the fingers there.)"""
    start = np.concatenate([[-4.0, -4.0, 0.0], half_folded_arm()])
    goal = np.concatenate([[4.0, 4.0, np.pi / 2], [-0.2, 1.2, 0.4, 0.9, 0.3]])
    return np.tile(start, (n, 1)), np.tile(goal, (n, 1))


def config5_queries(check_states, n_queries, seed=5, xy_half=11.0, goal_dist=(1.0, 2.5)):
    """collision-free (start, goal) pairs in W512 by rejection sampling with the checker that is
    handed in (`check_states(q) -> bool[n]`, True = in collision; self + map).  2048 boxes leave ~1 % of uniformly drawn
    base poses free, and free space falls into pockets: a goal is drawn 1 - 2.5 m from its start and kept when it is
    free and the straight base motion between them is free for the start's arm pose — so that most queries are solvable
    within their iteration budget and spend it in the optimisation phase (near-vertex search, choose-parent, rewire),
    which is what the configuration is meant to exercise."""
    from squirrel_motion_planner_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    s0, g0 = config1_queries(1)
    starts, goals = [], []
    for _ in range(1500):   # ~1.5 pairs per round survive in W512 with the covering model
        cand = syn.sample_configs(1 << 15, seed=int(rng.integers(1 << 30)), xy_half=xy_half)
        cand[:, 3:] = s0[0, 3:] + rng.uniform(-0.15, 0.15, size=(len(cand), 5))
        free = cand[~check_states(cand)]
        if len(free) == 0:
            continue
        ang, dist = rng.uniform(-np.pi, np.pi, len(free)), rng.uniform(goal_dist[0], goal_dist[1], len(free))
        g = free.copy()
        g[:, 0] += dist * np.cos(ang)
        g[:, 1] += dist * np.sin(ang)
        g[:, 2] = rng.uniform(-np.pi, np.pi, len(free))
        g[:, 3:] = g0[0, 3:] + rng.uniform(-0.15, 0.15, size=(len(free), 5))
        keep = ~check_states(g) & (np.abs(g[:, 0]) < xy_half) & (np.abs(g[:, 1]) < xy_half)
        free, g = free[keep], g[keep]
        if len(free) == 0:
            continue
        mid = np.repeat(free, 9, axis=0)                      # the straight base motion, sampled every ~12 % of the way
        t = np.tile(np.linspace(0.1, 0.9, 9), len(free))[:, None]
        mid[:, :2] = mid[:, :2] + t * (np.repeat(g, 9, axis=0)[:, :2] - mid[:, :2])
        clear = ~check_states(mid).reshape(-1, 9).any(axis=1)
        starts += list(free[clear])
        goals += list(g[clear])
        if len(starts) >= n_queries:
            break
    if len(starts) < n_queries:
        raise RuntimeError("not enough collision-free query pairs (%d of %d)" % (len(starts), n_queries))
    return np.array(starts[:n_queries]), np.array(goals[:n_queries])
