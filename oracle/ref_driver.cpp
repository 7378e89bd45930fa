// ref_driver.cpp — C entry points over the REFERENCE'S OWN code for the near-neighbour path (TEST INFRASTRUCTURE).
//
// Compiled by oracle/Makefile.ref together with, and only with, these unmodified reference sources read where they
// lie under /root/reference (never copied into this repository):
//   birrt_star/robot_motion_planning/planning_heuristics/src/distance_heuristics.cpp        (DistanceHeuristics)
//   birrt_star/robot_motion_planning/planner_data_structures/include/.../data_structs.h      (Node, operator<, tree)
// ROS and orocos_kdl are absent from this container; oracle/ref_stubs/ supplies the two headers those files
// include (a logging macro and the joint-type-name accessors).  birrt_star.cpp itself cannot be compiled (MoveIt,
// FCL, KDL, octomap ...), so the two search LOOPS below restate birrt_star.cpp:4076-4133 and :4272-4324 around the
// real DistanceHeuristics::euclidean_joint_space_distance and the real Node::operator< / std::sort.
#include <algorithm>
#include <cstdint>

#include <planner_data_structures/data_structs.h>
#include <planning_heuristics/distance_heuristics.h>

namespace {
heuristics_motion_planning::DistanceHeuristics g_heuristic;

Rrt_star_tree make_tree(const double *nodes, const double *cost, int64_t n) {
  Rrt_star_tree t;
  t.nodes.resize(size_t(n));
  for (int64_t i = 0; i < n; ++i) {
    Node &nd = t.nodes[size_t(i)];
    nd.node_id = int(i);          // invariant of insertNode (birrt_star.cpp:3298-3322): node_id == index
    nd.parent_id = 0;
    nd.config.assign(nodes + i * 8, nodes + i * 8 + 8);
    nd.cost_reach.total = cost ? cost[i] : 0.0;
    nd.cost_reach.revolute = nd.cost_reach.prismatic = 0.0;
  }
  t.num_nodes = int(n);
  return t;
}
}  // namespace

extern "C" {

// DistanceHeuristics::euclidean_joint_space_distance(vector<double>, vector<double>)  distance_heuristics.cpp:128-156
double ref_joint_space_distance(const double *a, const double *b, int n) {
  return g_heuristic.euclidean_joint_space_distance(std::vector<double>(a, a + n), std::vector<double>(b, b + n));
}

// the (KDL::Chain, ...) overload, distance_heuristics.cpp:159-214; type[i]: 0 None, 1 RotAxis, 2 TransAxis
void ref_joint_space_distance_chain(const int *type, int n_seg, const double *a, const double *b, int n, double out[3]) {
  KDL::Chain chain;
  for (int i = 0; i < n_seg; ++i)
    chain.addSegment(KDL::Segment(KDL::Joint(type[i] == 1 ? "RotAxis" : (type[i] == 2 ? "TransAxis" : "None"))));
  const std::vector<double> d =
      g_heuristic.euclidean_joint_space_distance(chain, std::vector<double>(a, a + n), std::vector<double>(b, b + n));
  out[0] = d[0]; out[1] = d[1]; out[2] = d[2];
}

// find_nearest_neighbour_interpolation, birrt_star.cpp:4076-4133 (serial order; m_solution_path_available has no
// effect there: the ellipsoid test is commented out)
void ref_nearest_batch(const double *nodes, int64_t n, const double *queries, int64_t nq, int32_t *ids) {
  Rrt_star_tree tree = make_tree(nodes, nullptr, n);
  for (int64_t q = 0; q < nq; ++q) {
    const std::vector<double> x_rand(queries + q * 8, queries + q * 8 + 8);
    int id_nearest_neighbour = 0;
    double min_node_dist = 10000.0;
    for (size_t k = 0; k < tree.nodes.size(); ++k) {
      const double curr_node_dist = g_heuristic.euclidean_joint_space_distance(tree.nodes[k].config, x_rand);
      if (curr_node_dist < min_node_dist) {
        id_nearest_neighbour = tree.nodes[k].node_id;
        min_node_dist = curr_node_dist;
      }
    }
    ids[q] = id_nearest_neighbour;
  }
}

// find_near_vertices_interpolation, birrt_star.cpp:4272-4324: collect Nodes with d < r and id != x_new.id, std::sort
// them with the reference's Node::operator< (cost_reach.total), return the ids.  Writes min(count, cap) ids.
int64_t ref_near_vertices(const double *nodes, const double *cost, int64_t n, const double *query, int32_t exclude_id,
                          double r, int32_t *ids_out, int64_t cap) {
  Rrt_star_tree tree = make_tree(nodes, cost, n);
  const std::vector<double> x_new(query, query + 8);
  std::vector<Node> near_nodes;
  for (size_t k = 0; k < tree.nodes.size(); ++k) {
    const double d = g_heuristic.euclidean_joint_space_distance(tree.nodes[k].config, x_new);
    if (d < r && tree.nodes[k].node_id != exclude_id) near_nodes.push_back(tree.nodes[k]);
  }
  std::sort(near_nodes.begin(), near_nodes.end());
  for (size_t i = 0; i < near_nodes.size() && int64_t(i) < cap; ++i) ids_out[i] = near_nodes[i].node_id;
  return int64_t(near_nodes.size());
}
}
