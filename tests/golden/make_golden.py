#!/usr/bin/env python
"""Golden vectors for the oracle, computed by code paths that do NOT share anything with oracle/sqmp_oracle.c:

  fk_scipy.npz        forward kinematics and link frames straight from the reference URDF with scipy Rotations
                      (needs /root/reference — run in the build container; output is committed)
  collision_numpy.npz sphere-vs-voxel clearances by dense numpy brute force over every occupied voxel
  nn_numpy.npz        nearest / radius results by numpy brute force with the reference's summation order
  interp_numpy.npz    interpolateConfigurations by the numpy formula
  rooms_bt.npz        the reference's room3/4/5.bt octomaps parsed to dense occupancy (counts + bits)
  nn_reference.npz    distances, nearest ids and near-vertex id lists produced by the REFERENCE'S OWN code
                      (DistanceHeuristics::euclidean_joint_space_distance, Node::operator< + std::sort) compiled
                      from /root/reference by oracle/Makefile.ref — needs the build container

    python tests/golden/make_golden.py
"""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial.transform import Rotation as Rot

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn  # noqa: E402

REF = "/root/reference/squirrel_8dof_planner/config/robotino_plan.urdf"


def urdf_fk():
    root = ET.parse(REF).getroot()
    J = {}
    for j in root.findall("joint"):
        o, a = j.find("origin"), j.find("axis")
        f = lambda s: np.array([float(v) for v in s.split()])
        J[j.find("child").get("link")] = dict(
            parent=j.find("parent").get("link"), type=j.get("type"),
            xyz=f(o.get("xyz")) if o is not None and o.get("xyz") else np.zeros(3),
            rpy=f(o.get("rpy")) if o is not None and o.get("rpy") else np.zeros(3),
            axis=f(a.get("xyz")) if a is not None else None)

    def T_of(link, q):
        if link not in J:
            return np.eye(4)
        j = J[link]
        T = np.eye(4)
        T[:3, :3] = Rot.from_euler("xyz", j["rpy"]).as_matrix()
        T[:3, 3] = j["xyz"]
        if link in rd.PLANNING_LINKS:
            v = q[rd.PLANNING_LINKS[link]]
            M = np.eye(4)
            if j["type"] == "prismatic":
                M[:3, 3] = j["axis"] * v
            else:
                M[:3, :3] = Rot.from_rotvec(j["axis"] / np.linalg.norm(j["axis"]) * v).as_matrix()
            T = T @ M
        return T_of(j["parent"], q) @ T

    q = syn.sample_configs(256, seed=101)
    ee = np.empty((len(q), 7))
    frames = {}
    names = ["base_body_link", "head_link", "arm_link2", "arm_link4", "hand_wrist_link", "hand_left_finger_upper_link",
             "kinect_link", "shell_base_link_front", "hokuyo_link"]
    for n in names:
        frames[n] = np.empty((len(q), 12))
    for i in range(len(q)):
        T = T_of("hand_wrist_link", q[i])
        ee[i, :3] = T[:3, 3]
        ee[i, 3:] = Rot.from_matrix(T[:3, :3]).as_quat()   # x y z w, sign fixed by the test
        for n in names:
            T = T_of(n, q[i])
            T[2, 3] += 0.02                                  # CC:201
            c = root.find("link[@name='%s']/collision/origin" % n)
            if c is not None and n not in rd.PLANNING_LINKS:  # CC:329-336
                Tc = np.eye(4)
                Tc[:3, 3] = [float(v) for v in c.get("xyz").split()]
                Tc[:3, :3] = Rot.from_euler("xyz", [float(v) for v in c.get("rpy").split()]).as_matrix()
                T = T @ Tc
            frames[n][i] = np.concatenate([T[:3, :3].reshape(9), T[:3, 3]])
    np.savez_compressed(os.path.join(HERE, "fk_scipy.npz"), q=q, ee=ee, frame_names=np.array(names),
                        **{"frame_" + n: frames[n] for n in names})


def collision_numpy():
    """Clearance of each state's sphere set against a small world, by brute force over ALL occupied voxels."""
    lt, fx = rd.load_fixture()
    spheres = [(n, np.array(c), r) for n, c, r in fx["spheres"]]
    w = syn.box_world(g=48, n_boxes=6, seed=77, dims=(48, 40, 32))
    origin = w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    occ = w.occupancy()
    kk, jj, ii = np.nonzero(occ)
    lo = origin[None, :] + np.stack([ii, jj, kk], 1) * w.res
    hi = lo + w.res
    q = syn.sample_configs(160, seed=202, xy_half=1.0)

    # link frames by plain numpy 4x4 products over the link table (the same constants, independent arithmetic)
    def frames(qv):
        T = [None] * lt.n
        for i in range(lt.n):
            p = lt.parent[i]
            if p < 0:
                T[i] = rd.frame(p=[0, 0, 0.02])
                continue
            if lt.joint_type[i] == rd.JOINT_ROT:
                M = rd.frame(Rot.from_rotvec(lt.joint_axis[i] * qv[lt.qidx[i]]).as_matrix(), lt.joint_origin[i])
            elif lt.joint_type[i] == rd.JOINT_TRANS:
                M = rd.frame(p=lt.joint_origin[i] + lt.joint_axis[i] * qv[lt.qidx[i]])
            else:
                tp = lt.to_parent[i]
                M = rd.frame(tp[:9].reshape(3, 3), tp[9:])
            T[i] = T[p] @ M
        return T
    clr = np.empty(len(q))
    per_link = np.zeros((len(q), lt.n), dtype=bool)
    for s in range(len(q)):
        T = frames(q[s])
        best = np.inf
        for n, c, r in spheres:
            i = lt.index(n)
            cw = T[i][:3, :3] @ c + T[i][:3, 3]
            d = np.maximum(np.maximum(lo - cw, cw - hi), 0.0)
            dist = np.sqrt((d * d).sum(1)).min() - r
            best = min(best, dist)
            if dist <= 0:
                per_link[s, i] = True
        clr[s] = best
    np.savez_compressed(os.path.join(HERE, "collision_numpy.npz"), q=q, clearance=clr, per_link=per_link,
                        bits=w.bits, dims=np.array(w.dims), origin=origin, res=w.res)


def nn_numpy():
    nodes, cost = syn.random_tree(3000, seed=303, cost_seed=304)
    nodes[100] = nodes[50]            # exact duplicates: ties must go to the lower id
    nodes[2000] = nodes[50]
    queries = syn.sample_configs(40, seed=305)
    queries[0] = nodes[50]
    queries[1] = nodes[2000] + 1e-9
    ids = np.empty(len(queries), dtype=np.int32)
    near = []
    for k, x in enumerate(queries):
        s = np.zeros(len(nodes))
        for i in range(8):            # DH:145-149: sequential sum of squares
            d = x[i] - nodes[:, i]
            s = s + d * d
        d = np.sqrt(s)
        best, bid = 10000.0, 0
        for n in range(len(nodes)):
            if d[n] < best:
                best, bid = d[n], n
        ids[k] = bid
        hit = np.nonzero(d < 1.5)[0]
        order = sorted(hit.tolist(), key=lambda n: (cost[n], n))
        near.append(np.array(order, dtype=np.int32))
    np.savez_compressed(os.path.join(HERE, "nn_numpy.npz"), nodes=nodes, cost=cost, queries=queries, nearest=ids,
                        near_count=np.array([len(v) for v in near]), near_flat=np.concatenate(near))


def interp_numpy():
    a, b = syn.edge_batch(64, seed=404)
    out = syn.interpolate_edges(a, b, 20)
    np.savez_compressed(os.path.join(HERE, "interp_numpy.npz"), start=a, end=b, states=out)


def nn_reference():
    """Outputs of the reference itself (oracle/_ref): these PIN the near-neighbour rows of the oracle."""
    import oracle
    if oracle.build_ref() is None:
        print("oracle/_ref not buildable here (no /root/reference): nn_reference.npz left as committed")
        return
    rng = np.random.default_rng(606)
    nodes, cost = syn.random_tree(2500, seed=601, cost_seed=602)
    cost[10] = cost[20] = cost[30]          # cost ties: std::sort is unstable, ids of a tie group are a SET
    nodes[700] = nodes[7]                   # distance ties: serial '<' keeps the lower id
    queries = syn.sample_configs(48, seed=603)
    queries[0] = nodes[7]
    queries[1] = nodes[700] - 1e-12
    a, b = rng.normal(size=(256, 8)) * 3.0, rng.normal(size=(256, 8)) * 3.0
    dist = np.array([oracle.ref_joint_space_distance(x, y) for x, y in zip(a, b)])
    seg_types = np.array([2, 2, 1, 0, 1, 1, 0, 1, 1, 1], dtype=np.int32)   # chain with fixed segments interleaved
    dist3 = np.array([oracle.ref_joint_space_distance_chain(seg_types, x, y) for x, y in zip(a, b)])
    nearest = oracle.ref_nearest(nodes, queries)
    near = [oracle.ref_near_vertices(nodes, cost, x, -1, 1.5) for x in queries]
    near_ex = [oracle.ref_near_vertices(nodes, cost, x, int(v[0]) if len(v) else -1, 1.5) for x, v in zip(queries, near)]
    np.savez_compressed(os.path.join(HERE, "nn_reference.npz"), nodes=nodes, cost=cost, queries=queries, a=a, b=b,
                        dist=dist, seg_types=seg_types, dist3=dist3, nearest=nearest,
                        near_count=np.array([len(v) for v in near]), near_flat=np.concatenate(near),
                        near_ex_count=np.array([len(v) for v in near_ex]), near_ex_flat=np.concatenate(near_ex))


def rooms_bt():
    """The reference's shipped octomaps (squirrel_8dof_planner/config/room{3,4,5}.bt) parsed to dense grids: node /
    leaf / voxel counts + the packed occupancy.  Derived data only; needs the build container."""
    from squirrel_motion_planner_b200 import octomap_io as oio
    base = "/root/reference/squirrel_8dof_planner/config/"
    if not os.path.isdir(base):
        print("no /root/reference: rooms_bt.npz left as committed")
        return
    out = {}
    for name in ("room3", "room4", "room5"):
        res, occ, free, n = oio.parse_bt(open(base + name + ".bt", "rb").read())
        w = oio.to_voxel_world(res, occ)
        out[name + "_counts"] = np.array([n, len(occ), len(free), int((occ[:, 3] ** 3).sum())])
        out[name + "_bits"] = w.bits
        out[name + "_dims"] = np.array(w.dims)
        out[name + "_origin"] = w.origin
        out[name + "_res"] = res
    np.savez_compressed(os.path.join(HERE, "rooms_bt.npz"), **out)


if __name__ == "__main__":
    rooms_bt()
    nn_reference()
    urdf_fk()
    collision_numpy()
    nn_numpy()
    interp_numpy()
    print("golden vectors written to", HERE)
