"""A second, independent restatement of one BI2RRT* iteration of the reference (birrt_star.cpp, C-space search,
single_extend_step = true — the configuration planner_parameters_robotino.yaml ships), in Python over the CPU oracle.

TEST INFRASTRUCTURE.  Its purpose is to pin the decisions of the package's C++ front-end (shim/sqmp_birrt.hpp): both are
written from the reference's source, in different languages, with different data structures (the C++ one batches its
collision checks and resolves the order afterwards; this one walks the reference's loops one edge at a time, with its
early exits), and they must produce the same decision trace, node counts and solution costs on the same seeds.

Reference lines restated: run_planner BS:1160-1293, expandTree BS:2224-2256, choose_node_parent_interpolation BS:4618-4735,
insertNode BS:3298-3322, rewireTreeInterpolation BS:5060-5229, recursiveNodeCostUpdate BS:5495-5606,
connectGraphsInterpolation BS:2608-3295, connectNodesInterpolation BS:4380-4440, the two samplers BS:3832-3878 / 3607-3829.
"""
import numpy as np

import oracle
from squirrel_motion_planner_b200 import informed

IS_REV = np.array([0, 0, 1, 1, 1, 1, 1, 1], dtype=np.int32)
FNV_OFFSET, FNV_PRIME, M64 = 1469598103934665603, 1099511628211, (1 << 64) - 1


class Node:
    def __init__(self, node_id, parent_id, config, cost):
        self.node_id, self.parent_id, self.config = node_id, parent_id, np.array(config, dtype=np.float64)
        self.cost = np.array(cost, dtype=np.float64)          # cost_reach: total, revolute, prismatic
        self.children = []                                     # outgoing_edges[].child_node_id, in insertion order


class Tree:
    def __init__(self, name):
        self.name, self.nodes = name, []

    def configs(self):
        return np.array([n.config for n in self.nodes])

    def costs(self):
        return np.array([n.cost[0] for n in self.nodes])


class Planner:
    def __init__(self, orc, flags, lo, hi, seed, near_threshold=1.5, num_points=20, max_near=20, step=0.6):
        self.orc, self.flags, self.lo, self.hi, self.seed = orc, flags, np.array(lo, float), np.array(hi, float), int(seed)
        self.r, self.P, self.k, self.step = near_threshold, num_points, max_near, step
        self.trace = FNV_OFFSET
        self.stats = dict(iterations=0, nn_calls=0, near_calls=0, edges_checked=0, rewires=0, parents_chosen=0, informed_samples=0,
                          solution_rewires=0)

    # ---- bookkeeping -------------------------------------------------------------------------------------------
    def note(self, v):
        for i in range(8):
            self.trace ^= (v >> (8 * i)) & 0xFF
            self.trace = (self.trace * FNV_PRIME) & M64

    def valid(self, q):
        return not self.orc.is_in_collision(q, bool(self.flags & 1), bool(self.flags & 2))

    def edge(self, a, b):
        """connectNodesInterpolation (BS:4380-4440): trajectory, its end configuration and its three-way cost."""
        traj = oracle.interpolate(a, b, self.P)
        return traj, traj[-1].copy(), np.array(oracle.edge_cost(traj, IS_REV))

    def edge_valid(self, traj):
        """isEdgeValid (BS:6878-6895): (valid, last_valid_node_idx)."""
        self.stats["edges_checked"] += 1
        return self.orc.is_edge_valid(traj, bool(self.flags & 1), bool(self.flags & 2))

    def insert(self, tree, config, parent, cost):
        n = Node(len(tree.nodes), parent, config, cost)
        tree.nodes.append(n)
        if parent >= 0:
            tree.nodes[parent].children.append(n.node_id)
        return n.node_id

    def insert_stepwise(self, tree, frm, target):
        """the `while (rand_sample_reached == false)` loops (BS:4655-4700, 2675-2710)."""
        cur = frm
        while True:
            reached, goal = oracle.step_towards(tree.nodes[cur].config, target, IS_REV, self.step)
            _, end, c = self.edge(tree.nodes[cur].config, goal)
            cur = self.insert(tree, end, cur, tree.nodes[cur].cost + c)
            if reached:
                return cur

    def nearest(self, tree, q):
        self.stats["nn_calls"] += 1
        return int(oracle.nearest(tree.configs(), np.asarray(q)[None])[0])

    def near(self, tree, q, exclude):
        """find_near_vertices_interpolation (BS:4272-4324): every near id, sorted by (cost, id)."""
        self.stats["near_calls"] += 1
        return [int(v) for v in oracle.near_vertices(tree.configs(), tree.costs(), np.asarray(q), int(exclude), self.r)]

    # ---- init_planner (BS:335-536) -----------------------------------------------------------------------------
    def init(self, start, goal):
        if not self.valid(start) or not self.valid(goal):
            return False
        self.start_tree, self.goal_tree = Tree("START"), Tree("GOAL")
        self.insert(self.start_tree, start, -1, np.zeros(3))
        self.insert(self.goal_tree, goal, -1, np.zeros(3))
        self.tree_a, self.tree_b = self.start_tree, self.goal_tree
        self.solution = False
        self.best = np.array([10000.0, 10000.0, 10000.0])
        self.connected_tree, self.node_tree_a, self.node_tree_b = None, -1, -1
        self.ellipse = informed.ellipse_init(start, goal, IS_REV)
        self.n_uniform = self.n_informed = 0
        return True

    def sample(self):
        if self.solution:                                       # BS:1168 (m_informed_sampling_active)
            el = informed.ctypes_struct(self.ellipse, float(self.best[1]), float(self.best[2]), (self.lo[0], self.hi[0]), (self.lo[1], self.hi[1]))
            q, _ = self.orc.sample_informed((self.seed + 0x5851f42d4c957f2d) & M64, self.n_informed, 1, self.lo, self.hi, el, 64)
            self.n_informed += 1
            self.stats["informed_samples"] += 1
            return q[0]
        q, _ = self.orc.sample_configs(self.seed, self.n_uniform, 1, self.lo, self.hi, 64)
        self.n_uniform += 1
        return q[0]

    # ---- recursiveNodeCostUpdate (BS:5495-5606) ------------------------------------------------------------------
    def cost_update(self, tree, vid, reduction):
        node = tree.nodes[vid]
        node.cost = node.cost + reduction
        if self.solution:
            if (vid == self.node_tree_b and tree is self.connected_tree) or (vid == self.node_tree_a and tree is not self.connected_tree):
                self.best = self.best + reduction
                self.stats["solution_rewires"] += 1
                self.note(vid | (6 << 40))
        for ch in node.children:
            self.cost_update(tree, ch, reduction)

    # ---- one iteration of the while loop of run_planner (BS:1160-1293) -----------------------------------------
    def iterate(self):
        self.stats["iterations"] += 1
        A, B = self.tree_a, self.tree_b
        x_rand = self.sample()
        nn = self.nearest(A, x_rand)
        self.note(nn)
        # expandTree (BS:2224-2256)
        _, ext = oracle.step_towards(A.nodes[nn].config, x_rand, IS_REV, self.step)
        traj, end, c = self.edge(A.nodes[nn].config, ext)
        extend_nn = self.edge_valid(traj)[0]
        self.note(int(extend_nn))
        if extend_nn:
            x_new, x_new_cost = end, A.nodes[nn].cost + c
        else:
            x_new, x_new_cost = np.array(x_rand), np.array([10000.0, 10000.0, 10000.0])
        x_new_parent, extend_bp = nn, False
        optimize = self.solution
        near_ids = []
        if optimize:                                            # BS:1223-1231
            near_ids = self.near(A, x_new, len(A.nodes))
            self.note(len(near_ids))
            # choose_node_parent_interpolation (BS:4618-4735): first <= max_near vertices, ascending cost; first valid wins
            for v in near_ids[:self.k]:
                if A.nodes[v].cost[0] < x_new_cost[0]:
                    traj, _, c = self.edge(A.nodes[v].config, x_new)
                    if A.nodes[v].cost[0] + c[0] <= x_new_cost[0]:
                        if self.edge_valid(traj)[0]:
                            extend_bp, x_new_parent = True, v
                            self.stats["parents_chosen"] += 1
                            self.note(v | (1 << 40))
                            break
        if not extend_nn and not extend_bp:                      # BS:1233
            self.tree_a, self.tree_b = B, A
            return
        if extend_bp:
            x_new_id = self.insert_stepwise(A, x_new_parent, x_new)
        else:
            x_new_id = self.insert(A, x_new, nn, x_new_cost)
        x_new, x_new_cost, x_new_parent = A.nodes[x_new_id].config.copy(), A.nodes[x_new_id].cost.copy(), A.nodes[x_new_id].parent_id
        if optimize:
            # rewireTreeInterpolation (BS:5060-5229): the last <= max_near near vertices, walked from the end
            for v in near_ids[::-1][:self.k]:
                nv = A.nodes[v]
                if not (x_new_cost[0] < nv.cost[0]):
                    continue
                if v == x_new_parent or nv.parent_id <= 0:
                    continue
                traj, end, c = self.edge(x_new, nv.config)
                gen_cost = x_new_cost + c
                if not (gen_cost[0] < nv.cost[0]):
                    continue
                if not self.edge_valid(traj)[0]:
                    continue
                reduction = gen_cost - nv.cost
                A.nodes[nv.parent_id].children.remove(v)
                nv.parent_id = x_new_id
                nv.config = end
                A.nodes[x_new_id].children.append(v)
                self.cost_update(A, v, reduction)
                self.stats["rewires"] += 1
                self.note(v | (2 << 40))
        # connectGraphsInterpolation (BS:2608-3295)
        xc = self.nearest(B, x_new)
        self.note(xc | (3 << 40))
        cost_solution = self.best.copy()
        connect_from, partial_target = -1, None
        traj, _, c = self.edge(B.nodes[xc].config, x_new)
        sol = B.nodes[xc].cost + c + x_new_cost
        if sol[0] < cost_solution[0]:
            ok, last_valid = self.edge_valid(traj)
            if ok:
                cost_solution, connect_from = sol, xc
            elif last_valid != 0:
                partial_target = traj[last_valid].copy()
        if self.solution:                                       # BS:2758: a cheaper connection through a near vertex of tree B
            for v in self.near(B, x_new, x_new_id)[:self.k]:
                if B.nodes[v].cost[0] < x_new_cost[0]:
                    traj, _, c = self.edge(B.nodes[v].config, x_new)
                    sol = B.nodes[v].cost + c + x_new_cost
                    if sol[0] < cost_solution[0] and self.edge_valid(traj)[0]:
                        cost_solution, connect_from, partial_target = sol, v, None
        if connect_from >= 0:
            reached = self.insert_stepwise(B, connect_from, x_new)
            self.solution, self.best = True, cost_solution
            self.connected_tree, self.node_tree_b, self.node_tree_a = B, reached, x_new_id
            self.note(connect_from | (4 << 40))
        elif partial_target is not None:
            self.insert_stepwise(B, xc, partial_target)
            self.note(xc | (5 << 40))
        self.tree_a, self.tree_b = B, A
