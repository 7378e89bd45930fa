"""Host-side mirror of the reference's interfaces for the hot path — same names, argument meaning and
error behaviour — implemented on the C ABI of libsqmp_b200.so.

    reference interface (file:line in /root/reference)                                   here
    -------------------------------------------------------------------------------------------------------
    CollisionChecker::isInCollision / setOcTree / setDisabledLinkMapCollisions /          CollisionChecker
        getCollisions                      collision_checker.hpp:76-133
    BiRRTstarPlanner::isEdgeValid (x2) / isConfigValid          birrt_star.cpp:6864-6908   BiRRTstarHotPath
    BiRRTstarPlanner::find_nearest_neighbour_interpolation      birrt_star.cpp:4076-4133   BiRRTstarHotPath
    BiRRTstarPlanner::find_near_vertices_interpolation          birrt_star.cpp:4272-4324   BiRRTstarHotPath
    BiRRTstarPlanner::insertNode (tree mirror)                  birrt_star.cpp:3298-3322   BiRRTstarHotPath
    KDLRobotModel::compute_FK                                   kdl_kuka_model.cpp:278-305 KDLRobotModel
    Node / Rrt_star_tree                                        data_structs.h:29-54       Node / Rrt_star_tree

The C++ planner binds the same entry points through the header-only shim in shim/ (see INTEGRATION.md);
this module is what the parity tests drive, so they read like calls into the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import api, robot_desc as rd
from .synthetic import VoxelWorld


class CollisionChecker:
    """collision_checker.hpp:35-133.  True = in collision; fails closed when not initialised (CC:106-107)."""

    def __init__(self, device: int = 0, link_table=None, spheres=None):
        self.initialized = False
        self.ctx = None
        try:
            if link_table is None:
                link_table, fx = rd.load_fixture()
                spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
            self.link_table = link_table
            self.ctx = api.Context(rd.compile_robot(link_table, spheres), device)
            self.initialized = True
        except (api.SqmpError, ImportError, OSError, ValueError) as e:   # CC:68-72: initialise() failed
            self.init_error = e

    # CC:76-88 — deep copy; None ignored; the z=-0.02 octree offset (CC:87) is applied here
    def setOcTree(self, world: VoxelWorld | None):
        if world is None or not self.initialized:
            return
        origin = np.array(world.origin, dtype=np.float64) + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])
        self.ctx.set_voxel_grid(world.bits, world.dims, origin, world.res)

    # CC:90-102
    def setDisabledLinkMapCollisions(self, linksDisabled):
        if self.initialized:
            self.ctx.set_disabled_links(rd.link_mask(self.link_table, linksDisabled))

    # CC:104-121
    def isInCollision(self, jointPositions, checkSelfCollision: bool, checkMapCollision: bool) -> bool:
        if not self.initialized:
            return True
        if not checkSelfCollision and not checkMapCollision:
            return False
        flags = (api.CHECK_SELF if checkSelfCollision else 0) | (api.CHECK_MAP if checkMapCollision else 0)
        try:
            return bool(self.ctx.check_states(np.asarray(jointPositions, dtype=np.float64)[:8], flags)[0] & 1)
        except api.SqmpError:
            return True   # fail closed

    # CC:123-133 — names of colliding link pairs and of links hitting the map
    def getCollisions(self, jointPositions):
        lm, ph = self.ctx.get_collisions(jointPositions)
        names = self.link_table.names
        pairs = [(names[a], names[b]) for a, b in zip(*np.nonzero(ph[:len(names), :len(names)])) if a < b]   # CC:387-390: i < j
        return pairs, sorted(names[i] for i in np.nonzero(lm[:len(names)])[0])   # std::map order (CC:614)


@dataclass
class Node:                       # data_structs.h:29-45 (fields the hot path reads)
    node_id: int = 0
    parent_id: int = -1
    config: np.ndarray = None
    cost_reach_total: float = 0.0

    def __lt__(self, other):      # DS:41-44
        return self.cost_reach_total < other.cost_reach_total


@dataclass
class Rrt_star_tree:              # data_structs.h:47-54, plus the handle of its device mirror
    name: str = "START"
    nodes: list = field(default_factory=list)
    index: int = 0                # sqmp tree slot


@dataclass
class Edge:                       # data_structs.h:11-19
    joint_trajectory: np.ndarray = None     # (n_states, 8)


class KDLRobotModel:
    """kdl_kuka_model.cpp:278-305."""

    def __init__(self, checker: CollisionChecker):
        self.ctx = checker.ctx

    def compute_FK(self, chain_config):
        q = np.asarray(chain_config, dtype=np.float64).reshape(-1, 8)
        out = self.ctx.fk(q)
        return out[0] if q.shape[0] == 1 else out


class BiRRTstarHotPath:
    """The members of BiRRTstarPlanner that sit on the hot path (birrt_star.h:103-107, 322-325)."""

    def __init__(self, checker: CollisionChecker, near_threshold_interpolation=1.5, num_traj_segments_interp=20,
                 max_near_nodes=20):
        self.collisionChecker = checker
        self.ctx = checker.ctx
        self.m_near_threshold_interpolation = near_threshold_interpolation
        self.m_num_traj_segments_interp = num_traj_segments_interp
        self.m_max_near_nodes = max_near_nodes

    @staticmethod
    def _flags(check_self, check_map):
        return (api.CHECK_SELF if check_self else 0) | (api.CHECK_MAP if check_map else 0)

    # BS:6897-6908
    def isConfigValid(self, config, check_self_collision=True, check_map_collision=True) -> bool:
        return not self.collisionChecker.isInCollision(config, check_self_collision, check_map_collision)

    # BS:6864-6895 — returns (valid, last_valid_node_idx); the bool-only overload is [0] of it
    def isEdgeValid(self, tree_edge: Edge, check_self_collision=True, check_map_collision=True):
        traj = np.ascontiguousarray(tree_edge.joint_trajectory, dtype=np.float64).reshape(-1, 8)
        n = traj.shape[0]
        if not self.collisionChecker.initialized:
            return False, 0
        if n == 0:
            return True, 0
        flags = self._flags(check_self_collision, check_map_collision)
        if flags == 0:
            return True, n - 1
        mask = self.ctx.check_states(traj, flags)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n]
        bad = np.nonzero(bits)[0]
        if bad.size == 0:
            return True, n - 1
        return False, max(int(bad[0]) - 1, 0)

    # speculative batch form of connectNodesInterpolation + isEdgeValid for all candidates of one iteration
    def areEdgesValid(self, q_start, q_end, check_self_collision=True, check_map_collision=True):
        mask, first = self.ctx.check_edges(q_start, q_end, self.m_num_traj_segments_interp,
                                           self._flags(check_self_collision, check_map_collision))
        last_valid = np.where(first < 0, self.m_num_traj_segments_interp, np.maximum(first - 1, 0))
        return first < 0, last_valid

    # BS:3881-3930 sampleJointConfig_Vector over getRandomConf (CL:1120-1188): uniform in the limits, end-effector
    # above the platform; drawn `pool` at a time on the device, sample i a pure function of (seed, i)
    def setSamplingBox(self, q_min, q_max, seed=1, pool=256):
        self._sample_box = (np.asarray(q_min, dtype=np.float64), np.asarray(q_max, dtype=np.float64), int(seed), int(pool))
        self._sample_next, self._sample_pool, self._sample_pos = 0, np.zeros((0, 8)), 0

    def sampleJointConfig_Vector(self):
        lo, hi, seed, pool = self._sample_box
        if self._sample_pos >= len(self._sample_pool):
            self._sample_pool, _ = self.ctx.sample_configs(seed, self._sample_next, pool, lo, hi, 64)
            self._sample_next += pool
            self._sample_pos = 0
        q = self._sample_pool[self._sample_pos]
        self._sample_pos += 1
        return q.tolist()

    # BS:3298-3322 / DS invariant node_id == index
    def insertNode(self, tree: Rrt_star_tree, node: Node):
        nid = self.ctx.tree_append(tree.index, node.config, node.cost_reach_total)
        node.node_id = nid
        tree.nodes.append(node)
        return nid

    def resetTree(self, tree: Rrt_star_tree):
        tree.nodes.clear()
        self.ctx.tree_reset(tree.index)

    # BS:4076-4133
    def find_nearest_neighbour_interpolation(self, tree: Rrt_star_tree, x_rand) -> Node:
        nid = int(self.ctx.nn_nearest(tree.index, np.asarray(x_rand, dtype=np.float64).reshape(1, 8))[0])
        return tree.nodes[nid]

    # BS:4272-4324 — ids in ascending (cost, id) order.  Only the first / last max_near_nodes are ever used by
    # the callers (BS:4618-4623, 5075-5089); `full=True` returns the first k and last k with the total count.
    def find_near_vertices_interpolation(self, tree: Rrt_star_tree, x_new: Node):
        low, high, cnt = self.ctx.nn_radius(tree.index, np.asarray(x_new.config, dtype=np.float64).reshape(1, 8),
                                            np.array([x_new.node_id], dtype=np.int32),
                                            self.m_near_threshold_interpolation, self.m_max_near_nodes)
        n = int(cnt[0])
        lo = [int(v) for v in low[0] if v >= 0]
        hi = [int(v) for v in high[0] if v >= 0]
        return lo, hi, n
