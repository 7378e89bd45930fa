// sqmp_birrt.hpp — batched planner front-end: a host restatement of one BI²RRT* iteration of
// birrt_star.cpp for the configuration the reference ships for this robot (planner_parameters_robotino.yaml:
// C-space search, single_extend_step = true, interpolation edges with 20 segments, near radius 1.5, 20 near nodes,
// extend step 0.6) that hands the hot path to a Backend in BATCHES:
//
//   reference (sequential, early exits)                         here
//   -----------------------------------------------------------------------------------------------------------
//   expandTree: isEdgeValid(e_new)                    BS:2237   batch A: 1 edge
//   choose_node_parent_interpolation: <= 20 x         BS:4618   batch B: all candidate parent edges + all candidate
//       (connectNodesInterpolation + isEdgeValid)                        rewire edges of the iteration in ONE
//   rewireTreeInterpolation: <= 20 x (... isEdgeValid) BS:5077           check_edges call; the reference's order-
//   connectGraphsInterpolation: 1 + <= 20 x            BS:2642  batch C  dependent decisions (first valid parent wins,
//                                                                        costs change while rewiring, the best
//                                                                        connection so far gates the next) are then
//                                                                        resolved on the host in the reference's order
// Validity of an edge never depends on costs, so checking candidates speculatively in one batch and applying the
// cost tests afterwards yields exactly the decisions of the sequential code.
//
// Backend concept (the C ABI calls of include/sqmp_b200.h, or the CPU oracle in the tests):
//   void tree_reset(int slot);  int tree_append(int slot, const double q[8], double cost);
//   void tree_update(int slot, int id, const double q[8], double cost);
//   int  nearest(int slot, const double q[8]);
//   void near(int slot, const double q[8], int exclude_id, double r, int k, std::vector<int>& low, std::vector<int>& high, int& count);
//   bool config_valid(const double q[8]);
//   void check_edges(const std::vector<double>& starts, const std::vector<double>& ends, int num_points, std::vector<int>& first_invalid);
//
// Deliberately not restated (control plane / §8 f4): informed (ellipsoid) sampling, the EE-pose "control" search space,
// motion constraints, visualisation, time budgets (iterations are counted instead so CPU and GPU do identical work).
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

#include "sqmp_informed.hpp"

namespace sqmp_planner {

struct Params {
  double near_threshold = 1.5;        // PY:4
  int num_points = 20;                // PY:5  (21 states per edge)
  int max_near_nodes = 20;            // PY:6
  double extend_step = 0.6;           // PY:15
  bool tree_optimization = true;      // m_tree_optimization_active (BS:167)
  bool informed_sampling = true;      // m_informed_sampling_active (BS:172)
  double edge_cost_weights[8] = {1, 1, 1, 1, 1, 1, 1, 1};   // m_edge_cost_weights (BS:109-113)
  double lo[8] = {-5, -5, -3.14159265358979323846, -1.2, -1.7, -1.8, -2.4, -2.9};   // x, y from the map; arm: URDF limits
  double hi[8] = {5, 5, 3.14159265358979323846, 1.5, 2.6, 1.8, 2.4, 2.9};
  int is_revolute[8] = {0, 0, 1, 1, 1, 1, 1, 1};
  uint64_t seed = 1;
};

struct PNode {
  int id, parent;
  double q[8];
  double cost;                  // cost_reach.total
  double cost_rev, cost_pri;    // cost_reach.revolute / .prismatic
  std::vector<int> children;    // child node ids (the reference keeps outgoing_edges)
};
struct PTree {
  int slot;
  std::vector<PNode> nodes;
};

struct Stats {
  long iterations = 0, nn_calls = 0, near_calls = 0, edge_batches = 0, edges_checked = 0, states_checked = 0,
       rewires = 0, parents_chosen = 0, nodes_start = 0, nodes_goal = 0, informed_samples = 0, solution_rewires = 0;
  bool solved = false;
  double best_cost = 10000.0, best_cost_rev = 10000.0, best_cost_pri = 10000.0;
  uint64_t trace = 1469598103934665603ull;   // FNV-1a over every decision: equal traces <=> equal decision sequences
};

// stepTowardsRandSample, birrt_star.cpp:5712-5868: x_rand is pulled back to one extend step from nn, separately in the
// revolute and in the prismatic subspace; returns true when the sample itself was reached.
inline bool step_towards(const double *nn, double *x_rand, const int *is_revolute, double factor) {
  bool reached = true;
  double d[8], s_rev = 0.0, s_pri = 0.0;
  for (int j = 0; j < 8; ++j) {
    d[j] = x_rand[j] - nn[j];
    const double sq = d[j] * d[j];
    if (!is_revolute[j]) s_pri += sq; else s_rev += sq;
  }
  const double len_rot = std::sqrt(s_rev), len_pri = std::sqrt(s_pri);
  const bool rot_reached = len_rot < 0.001, pri_reached = len_pri < 0.001;
  double ext[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  s_rev = s_pri = 0.0;
  for (int j = 0; j < 8; ++j) {
    if (!is_revolute[j]) {
      if (!pri_reached) {
        d[j] = d[j] / len_pri;
        const double c_step = factor * d[j];
        ext[j] = nn[j] + c_step;
        s_pri += c_step * c_step;
      }
    } else if (!rot_reached) {
      d[j] = d[j] / len_rot;
      const double c_step = factor * d[j];
      ext[j] = nn[j] + c_step;
      s_rev += c_step * c_step;
    }
  }
  const double ext_pri = s_pri == 0.0 ? 1000.0 : std::sqrt(s_pri), ext_rot = s_rev == 0.0 ? 1000.0 : std::sqrt(s_rev);
  if (ext_rot < len_rot) {
    for (int j = 0; j < 8; ++j)
      if (is_revolute[j]) x_rand[j] = ext[j];
    reached = false;
  }
  if (ext_pri < len_pri) {
    for (int j = 0; j < 8; ++j)
      if (!is_revolute[j]) x_rand[j] = ext[j];
    reached = false;
  }
  return reached;
}

// interpolateConfigurations (BS:4461-4478) + compute_edge_cost_interpolation (BS:4162-4242, unit weights): total cost
inline double edge_cost(const double *a, const double *b, int num_points) {
  double step[8], prev[8], cost = 0.0;
  for (int j = 0; j < 8; ++j) {
    step[j] = (b[j] - a[j]) / double(num_points);
    prev[j] = a[j] + 0.0 * step[j];
  }
  for (int k = 1; k <= num_points; ++k) {
    double s = 0.0;
    for (int j = 0; j < 8; ++j) {
      const double cur = a[j] + double(k) * step[j];
      s += (cur - prev[j]) * (cur - prev[j]) * 1.0;
      prev[j] = cur;
    }
    cost += std::sqrt(s);
  }
  return cost;
}
// compute_edge_cost_interpolation (BS:4162-4242) over the interpolated trajectory of BS:4461-4478: per segment the squared
// coordinate differences, weighted for the total (m_edge_cost_weights), split by joint class for the other two; the
// three costs are the sums of the segments' square roots.  out = {total, revolute, prismatic}.
// This is the planner's hottest host function (up to ~40 candidate edges per iteration): the loops run joint-outer,
// segment-inner so that the compiler vectorises over the segments (an AVX2 clone where the CPU has it; no FMA contraction —
// that would change the bits).  Every segment's three sums still receive their terms in joint order and the segments are
// added up in order, so the result is bit for bit that of the reference's nesting (tests/test_planner_frontend.py).
#if defined(__GNUC__) && defined(__x86_64__) && !defined(SQMP_NO_TARGET_CLONES)
#define SQMP_SIMD_CLONES __attribute__((target_clones("avx2", "default"), optimize("O3")))
#else
#define SQMP_SIMD_CLONES
#endif
SQMP_SIMD_CLONES
static void edge_cost3(const double *a, const double *b, int num_points, const double *weights, const int *is_revolute, double *out) {
  enum { MAXP = 256 };
  double tot[MAXP], rot[MAXP], pri[MAXP];
  const int np = num_points < MAXP ? num_points : MAXP - 1;   // (the library limits num_points to 255)
  for (int k = 1; k <= np; ++k) tot[k] = rot[k] = pri[k] = 0.0;
  for (int j = 0; j < 8; ++j) {
    const double aj = a[j], step = (b[j] - a[j]) / double(num_points), w = weights[j];
    double *cls = is_revolute[j] ? rot : pri;
    for (int k = 1; k <= np; ++k) {
      const double cur = aj + double(k) * step, prev = aj + double(k - 1) * step;   // prev = the previous segment's cur
      const double coordinate_dist = (cur - prev) * (cur - prev);
      tot[k] += coordinate_dist * w;
      cls[k] += coordinate_dist;
    }
  }
  double edge_cost_total = 0.0, edge_cost_rev = 0.0, edge_cost_prism = 0.0;
  for (int k = 1; k <= np; ++k) {
    edge_cost_total += std::sqrt(tot[k]);
    edge_cost_rev += std::sqrt(rot[k]);
    edge_cost_prism += std::sqrt(pri[k]);
  }
  out[0] = edge_cost_total;
  out[1] = edge_cost_rev;
  out[2] = edge_cost_prism;
}
// last state of the interpolation = what connectNodesInterpolation stores as x_new.config (BS:4425-4426)
inline void edge_end(const double *a, const double *b, int num_points, double *out) {
  for (int j = 0; j < 8; ++j) out[j] = a[j] + double(num_points) * ((b[j] - a[j]) / double(num_points));
}

struct Cost3 {   // CostsReach (DS:21-27): total / revolute / prismatic
  double t, r, p;
};
inline Cost3 operator+(const Cost3 &a, const Cost3 &b) { return Cost3{a.t + b.t, a.r + b.r, a.p + b.p}; }
inline Cost3 operator-(const Cost3 &a, const Cost3 &b) { return Cost3{a.t - b.t, a.r - b.r, a.p - b.p}; }

template <class Backend>
class BiRRTStar {
 public:
  BiRRTStar(Backend &backend, const Params &p) : be(backend), P(p) {
    start_tree.slot = 0;
    goal_tree.slot = 1;
  }

  // init_planner (BS:335-536): both roots must be valid (BS:351-352); trees hold their root with cost 0; the ellipse of
  // the informed sampler is set up from start and goal (jointConfigEllipseInitialization, BS:519)
  bool init(const double *q_start, const double *q_goal) {
    st = Stats();
    pool.clear();
    pool_pos = 0;
    pool_next = 0;
    ipool.clear();
    ipool_pos = 0;
    inf_next = 0;
    if (!be.config_valid(q_start) || !be.config_valid(q_goal)) return false;
    for (PTree *t : {&start_tree, &goal_tree}) {
      t->nodes.clear();
      be.tree_reset(t->slot);
    }
    insert(start_tree, q_start, -1, Cost3{0.0, 0.0, 0.0});
    insert(goal_tree, q_goal, -1, Cost3{0.0, 0.0, 0.0});
    tree_a = &start_tree;
    tree_b = &goal_tree;
    solution = false;
    best = Cost3{10000.0, 10000.0, 10000.0};   // m_cost_best_solution_path* initial values (BS:132-134)
    conn_slot = node_tree_a = node_tree_b = -1;
    ellipse.init(q_start, q_goal, P.is_revolute);
    return true;
  }

  const Stats &run(long iterations) {
    for (long it = 0; it < iterations; ++it) iterate();
    st.nodes_start = long(start_tree.nodes.size());
    st.nodes_goal = long(goal_tree.nodes.size());
    st.solved = solution;
    st.best_cost = best.t;
    st.best_cost_rev = best.r;
    st.best_cost_pri = best.p;
    return st;
  }
  const Stats &stats() const { return st; }
  const PTree &tree(int slot) const { return slot == 0 ? start_tree : goal_tree; }

 private:
  struct Cand {   // one candidate edge of a batch
    double a[8], b[8];
  };

  void note(uint64_t v) {
    for (int i = 0; i < 8; ++i) {
      st.trace ^= (v >> (8 * i)) & 0xff;
      st.trace *= 1099511628211ull;
    }
  }
  static Cost3 cost_of(const PNode &n) { return Cost3{n.cost, n.cost_rev, n.cost_pri}; }
  Cost3 edge3(const double *a, const double *b) const {
    double c[3];
    edge_cost3(a, b, P.num_points, P.edge_cost_weights, P.is_revolute, c);
    return Cost3{c[0], c[1], c[2]};
  }
  int insert(PTree &t, const double *q, int parent, const Cost3 &cost) {   // insertNode, BS:3298-3322
    PNode n;
    n.id = int(t.nodes.size());
    n.parent = parent;
    for (int j = 0; j < 8; ++j) n.q[j] = q[j];
    n.cost = cost.t;
    n.cost_rev = cost.r;
    n.cost_pri = cost.p;
    t.nodes.push_back(n);
    if (parent >= 0) t.nodes[size_t(parent)].children.push_back(n.id);
    const int id = be.tree_append(t.slot, q, cost.t);
    (void)id;
    return n.id;
  }
  void check(const std::vector<Cand> &c, std::vector<int> &first_invalid) {
    first_invalid.clear();
    if (c.empty()) return;
    std::vector<double> a(c.size() * 8), b(c.size() * 8);
    for (size_t e = 0; e < c.size(); ++e)
      for (int j = 0; j < 8; ++j) {
        a[e * 8 + j] = c[e].a[j];
        b[e * 8 + j] = c[e].b[j];
      }
    be.check_edges(a, b, P.num_points, first_invalid);
    st.edge_batches += 1;
    st.edges_checked += long(c.size());
    st.states_checked += long(c.size()) * (P.num_points + 1);
  }
  static Cand cand(const double *a, const double *b) {
    Cand c;
    for (int j = 0; j < 8; ++j) { c.a[j] = a[j]; c.b[j] = b[j]; }
    return c;
  }
  // recursiveNodeCostUpdate, BS:5495-5606: the whole subtree moves by the cost reduction; when it passes one of the two
  // nodes the solution path is connected through, the best solution cost moves with it (BS:5523-5590)
  void update_subtree_cost(PTree &t, int id, const Cost3 &delta) {
    std::vector<int> todo(1, id);
    while (!todo.empty()) {
      const int n = todo.back();
      todo.pop_back();
      PNode &nd = t.nodes[size_t(n)];
      nd.cost += delta.t;
      nd.cost_rev += delta.r;
      nd.cost_pri += delta.p;
      be.tree_update(t.slot, n, nd.q, nd.cost);
      if (solution && ((n == node_tree_b && t.slot == conn_slot) || (n == node_tree_a && t.slot != conn_slot))) {
        best = best + delta;
        st.solution_rewires += 1;
        note(uint64_t(n) | (6ull << 40));
      }
      for (size_t k = nd.children.size(); k-- > 0;) todo.push_back(nd.children[k]);   // depth first, children in order
    }
  }
  // the "while (rand_sample_reached == false)" loops (BS:4655-4700, 2675-2710): walk from `from` towards `target` in
  // extend steps, inserting the via nodes; returns the id of the node inserted at the target
  int insert_stepwise(PTree &t, int from, const double *target) {
    int cur = from;
    for (int guard = 0; guard < 100000; ++guard) {
      double goal[8];
      for (int j = 0; j < 8; ++j) goal[j] = target[j];
      const bool reached = step_towards(t.nodes[size_t(cur)].q, goal, P.is_revolute, P.extend_step);
      double endq[8];
      edge_end(t.nodes[size_t(cur)].q, goal, P.num_points, endq);
      const Cost3 c = cost_of(t.nodes[size_t(cur)]) + edge3(t.nodes[size_t(cur)].q, goal);
      cur = insert(t, endq, cur, c);
      if (reached) break;
    }
    return cur;
  }

  // ---- sampling: uniform in the limits with the end-effector above the platform (sampleJointConfig_JntArray, BS:3832-3878)
  //      until a solution exists, from the informed ellipse afterwards (BS:1168; sampleJointConfigfromEllipse_JntArray,
  //      BS:3607-3829).  Drawn in pools by the backend (the device sampler); sample i of a stream is a pure function of
  //      (seed, i) and — for the informed stream — of the best solution cost at the time it is drawn: the pool is thrown
  //      away whenever that cost changes, so the sequence does not depend on the pool size or the backend ----
  void draw(double *x_rand) {
    if (P.informed_sampling && solution) {
      if (ipool_pos >= ipool.size() / 8 || ipool_r != best.r || ipool_p != best.p) {
        sqmp_informed el;
        ellipse.fill(el, best.r, best.p, P.lo, P.hi);
        ipool.resize(size_t(INFORMED_POOL) * 8);
        be.sample_informed_batch(P.seed + 0x5851f42d4c957f2dull, inf_next, INFORMED_POOL, P.lo, P.hi, SAMPLE_MAX_TRIES, el, ipool.data());
        ipool_pos = 0;
        ipool_r = best.r;
        ipool_p = best.p;
      }
      for (int j = 0; j < 8; ++j) x_rand[j] = ipool[ipool_pos * 8 + size_t(j)];
      ipool_pos += 1;
      inf_next += 1;
      st.informed_samples += 1;
      return;
    }
    if (pool_pos >= pool.size() / 8) {
      pool.resize(size_t(SAMPLE_POOL) * 8);
      be.sample_batch(P.seed, pool_next, SAMPLE_POOL, P.lo, P.hi, SAMPLE_MAX_TRIES, pool.data());
      pool_next += SAMPLE_POOL;
      pool_pos = 0;
    }
    for (int j = 0; j < 8; ++j) x_rand[j] = pool[pool_pos * 8 + size_t(j)];
    pool_pos += 1;
  }

  void iterate() {
    st.iterations += 1;
    PTree &A = *tree_a, &B = *tree_b;
    double x_rand[8];
    draw(x_rand);
    // ---- nearest neighbour in tree A (BS:1189) ----
    const int nn = be.nearest(A.slot, x_rand);
    st.nn_calls += 1;
    note(uint64_t(nn));
    // ---- expandTree, single extend step (BS:2224-2256) ----
    double ext[8];
    for (int j = 0; j < 8; ++j) ext[j] = x_rand[j];
    step_towards(A.nodes[size_t(nn)].q, ext, P.is_revolute, P.extend_step);
    std::vector<Cand> batch(1, cand(A.nodes[size_t(nn)].q, ext));
    std::vector<int> fi;
    check(batch, fi);                                              // batch A
    const bool extend_nn = fi[0] < 0;
    note(extend_nn);
    double x_new[8];
    Cost3 x_new_cost;
    int x_new_parent = nn;
    if (extend_nn) {
      edge_end(A.nodes[size_t(nn)].q, ext, P.num_points, x_new);
      x_new_cost = cost_of(A.nodes[size_t(nn)]) + edge3(A.nodes[size_t(nn)].q, ext);
    } else {
      for (int j = 0; j < 8; ++j) x_new[j] = x_rand[j];            // x_new = x_rand (BS:2245)
      x_new_cost = Cost3{10000.0, 10000.0, 10000.0};               // BS:1209-1211
    }
    // ---- near vertices + choose parent + rewire candidates: ONE batch (BS:1223-1250) ----
    std::vector<int> low, high;
    int n_near = 0;
    bool extend_bp = false;
    std::vector<int> fi_b;
    size_t n_parent_cands = 0;
    std::vector<int> parent_cand_ids, rewire_cand_ids;
    const bool optimize = P.tree_optimization && solution;
    if (optimize) {
      be.near(A.slot, x_new, int(A.nodes.size()), P.near_threshold, P.max_near_nodes, low, high, n_near);
      st.near_calls += 1;
      note(uint64_t(n_near));
      std::vector<Cand> cb;
      for (int v : low) {                                          // first <= 20, ascending cost (BS:4618-4623)
        if (v < 0) break;
        if (A.nodes[size_t(v)].cost < x_new_cost.t) {              // BS:4626
          const double c = A.nodes[size_t(v)].cost + edge3(A.nodes[size_t(v)].q, x_new).t;
          if (c <= x_new_cost.t) {                                 // BS:4640  ('<=', unlike the rewire test)
            parent_cand_ids.push_back(v);
            cb.push_back(cand(A.nodes[size_t(v)].q, x_new));
          }
        }
      }
      n_parent_cands = cb.size();
      for (int v : high) {                                         // last <= 20, walked from the end (BS:5075-5089)
        if (v < 0) break;
        rewire_cand_ids.push_back(v);
        cb.push_back(cand(x_new, A.nodes[size_t(v)].q));
      }
      check(cb, fi_b);                                             // batch B
      for (size_t i = 0; i < n_parent_cands; ++i)
        if (fi_b[i] < 0) {                                         // first valid candidate wins (BS:4644-4717)
          extend_bp = true;
          x_new_parent = parent_cand_ids[i];
          st.parents_chosen += 1;
          note(uint64_t(x_new_parent) | (1ull << 40));
          break;
        }
    }
    if (!extend_nn && !extend_bp) {                                // BS:1233: nothing to insert, swap trees
      std::swap(tree_a, tree_b);
      return;
    }
    // ---- insertNode (BS:1237); a chosen parent is reached stepwise with via nodes (BS:4655-4700) ----
    int x_new_id;
    if (extend_bp) {
      x_new_id = insert_stepwise(A, x_new_parent, x_new);
    } else {
      x_new_id = insert(A, x_new, nn, x_new_cost);
    }
    {
      const PNode &xn = A.nodes[size_t(x_new_id)];
      for (int j = 0; j < 8; ++j) x_new[j] = xn.q[j];
      x_new_cost = cost_of(xn);
      x_new_parent = xn.parent;
    }
    // ---- rewireTreeInterpolation (BS:5060-5229): validity came with batch B, cost tests in the reference's order ----
    if (optimize) {
      for (size_t i = 0; i < rewire_cand_ids.size(); ++i) {
        const int v = rewire_cand_ids[i];
        PNode &nv = A.nodes[size_t(v)];
        if (!(x_new_cost.t < nv.cost)) continue;                   // BS:5079-5084 (counter over the walked suffix)
        if (v == x_new_parent || nv.parent == 0 || nv.parent < 0) continue;   // BS:5089 (children of the root are kept)
        const Cost3 c = x_new_cost + edge3(x_new, nv.q);           // gen_node.cost_reach (connectNodesInterpolation, BS:4417-4423)
        if (!(c.t < nv.cost)) continue;                            // BS:5107
        if (fi_b[n_parent_cands + i] >= 0) continue;               // BS:5110 isEdgeValid
        const Cost3 reduction = c - cost_of(nv);                   // BS:5117-5121
        std::vector<int> &sib = A.nodes[size_t(nv.parent)].children;
        for (size_t k = 0; k < sib.size(); ++k)
          if (sib[k] == v) { sib.erase(sib.begin() + long(k)); break; }
        nv.parent = x_new_id;
        double moved[8];
        edge_end(x_new, nv.q, P.num_points, moved);                // BS:5180: the vertex takes the config the edge ends in
        for (int j = 0; j < 8; ++j) nv.q[j] = moved[j];
        A.nodes[size_t(x_new_id)].children.push_back(v);
        update_subtree_cost(A, v, reduction);                      // BS:5203 recursiveNodeCostUpdate
        st.rewires += 1;
        note(uint64_t(v) | (2ull << 40));
      }
    }
    // ---- connect to the other tree (BS:1256-1263, connectGraphsInterpolation BS:2608-3295) ----
    const int xc = be.nearest(B.slot, x_new);
    st.nn_calls += 1;
    note(uint64_t(xc) | (3ull << 40));
    std::vector<Cand> cc(1, cand(B.nodes[size_t(xc)].q, x_new));
    std::vector<int> near_b_low, near_b_high, cc_ids;
    int n_near_b = 0;
    if (solution) {                                                // BS:2758
      // find_near_vertices_interpolation(tree_B, x_new) leaves out the node of tree B whose id equals x_new's id in
      // tree A (BS:4295 compares node ids only)
      be.near(B.slot, x_new, x_new_id, P.near_threshold, P.max_near_nodes, near_b_low, near_b_high, n_near_b);
      st.near_calls += 1;
      for (int v : near_b_low) {
        if (v < 0) break;
        if (B.nodes[size_t(v)].cost < x_new_cost.t) {              // BS:2827
          cc_ids.push_back(v);
          cc.push_back(cand(B.nodes[size_t(v)].q, x_new));
        }
      }
    }
    std::vector<int> fi_c;
    check(cc, fi_c);                                               // batch C
    Cost3 cost_solution = best;                                    // BS:2633-2635
    int connect_from = -1;          // node of tree B the connection starts from
    bool partial = false;           // connection blocked: extend B up to the last valid state (BS:2712-2755)
    double partial_target[8];
    {
      const Cost3 sol = cost_of(B.nodes[size_t(xc)]) + edge3(B.nodes[size_t(xc)].q, x_new) + x_new_cost;
      if (sol.t < cost_solution.t) {                               // BS:2648
        if (fi_c[0] < 0) {
          cost_solution = sol;
          connect_from = xc;
        } else {
          const int last_valid = fi_c[0] > 0 ? fi_c[0] - 1 : 0;    // isEdgeValid's last_valid_node_idx (BS:6878-6895)
          if (last_valid != 0) {                                   // BS:2716
            partial = true;
            const double *a = B.nodes[size_t(xc)].q;
            for (int j = 0; j < 8; ++j) partial_target[j] = a[j] + double(last_valid) * ((x_new[j] - a[j]) / double(P.num_points));
          }
        }
      }
    }
    for (size_t i = 0; i < cc_ids.size(); ++i) {                   // BS:2758-2850: a better connection via a near vertex
      const int v = cc_ids[i];
      const Cost3 sol = cost_of(B.nodes[size_t(v)]) + edge3(B.nodes[size_t(v)].q, x_new) + x_new_cost;   // BS:2838-2840
      if (sol.t < cost_solution.t && fi_c[1 + i] < 0) {
        cost_solution = sol;
        connect_from = v;
        partial = false;
      }
    }
    if (connect_from >= 0) {
      const int reached = insert_stepwise(B, connect_from, x_new); // trees are connected at x_new
      solution = true;
      best = cost_solution;                                        // BS:3267-3269
      conn_slot = B.slot;                                          // m_connected_tree_name (BS:3254)
      node_tree_b = reached;                                       // m_node_tree_B: the node of the connected tree at x_new
      node_tree_a = x_new_id;                                      // m_node_tree_A
      note(uint64_t(connect_from) | (4ull << 40));
    } else if (partial) {
      insert_stepwise(B, xc, partial_target);
      note(uint64_t(xc) | (5ull << 40));
    }
    std::swap(tree_a, tree_b);                                     // BS:1281-1293
  }

  Backend &be;
  Params P;
  static const int SAMPLE_POOL = 256, SAMPLE_MAX_TRIES = 64, INFORMED_POOL = 32;
  std::vector<double> pool, ipool;
  size_t pool_pos = 0, ipool_pos = 0;
  long long pool_next = 0, inf_next = 0;
  double ipool_r = -1.0, ipool_p = -1.0;   // best solution cost (revolute / prismatic) the informed pool was drawn with
  sqmp_shim::InformedEllipse ellipse;
  PTree start_tree, goal_tree;
  PTree *tree_a = nullptr, *tree_b = nullptr;
  bool solution = false;
  Cost3 best = Cost3{10000.0, 10000.0, 10000.0};
  int conn_slot = -1, node_tree_a = -1, node_tree_b = -1;
  Stats st;
};

}  // namespace sqmp_planner
