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
