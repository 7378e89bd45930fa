// test_shim.cpp — drives the C++ shim (sqmp_shim.hpp) exactly the way birrt_star.cpp drives the reference classes,
// on inputs written by tests/test_cpp_shim.py, and writes the results back for comparison with the CPU oracle.
//   usage: test_shim <robot.sqmpdesc> <work_dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "sqmp_shim.hpp"

template <class T>
static std::vector<T> read_raw(const std::string &path) {
  std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
  std::vector<T> v;
  if (!f) return v;
  const std::streamsize n = f.tellg();
  f.seekg(0);
  v.resize(size_t(n) / sizeof(T));
  f.read(reinterpret_cast<char *>(v.data()), n);
  return v;
}
template <class T>
static void write_raw(const std::string &path, const std::vector<T> &v) {
  std::ofstream f(path.c_str(), std::ios::binary);
  f.write(reinterpret_cast<const char *>(v.data()), std::streamsize(v.size() * sizeof(T)));
}

int main(int argc, char **argv) {
  if (argc >= 4 && std::string(argv[1]) == "--bt") {   // test_shim --bt in.bt out_prefix [pose_x pose_y floor_d]
    sqmp_shim::VoxelGrid g;
    std::string err;
    const bool floor = argc >= 7;
    if (!sqmp_shim::read_bt(argv[2], g, &err, floor, floor ? atof(argv[4]) : 0.0, floor ? atof(argv[5]) : 0.0,
                            floor ? atof(argv[6]) : 0.0)) {
      std::fprintf(stderr, "read_bt: %s\n", err.c_str());
      return 5;
    }
    write_raw(std::string(argv[3]) + ".bits.u32", g.bits);
    std::ofstream meta((std::string(argv[3]) + ".meta.txt").c_str());
    meta.precision(17);
    meta << g.gx << " " << g.gy << " " << g.gz << " " << g.origin[0] << " " << g.origin[1] << " " << g.origin[2] << " " << g.res << "\n";
    return 0;
  }
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s robot.sqmpdesc work_dir | --bt in.bt out_prefix [pose_x pose_y floor_d]\n", argv[0]);
    return 2;
  }
  const std::string dir = std::string(argv[2]) + "/";
  sqmp_shim::CollisionChecker::setDescriptionPath(argv[1]);
  sqmp_shim::CollisionChecker *collisionChecker = new sqmp_shim::CollisionChecker();   // BS:73, literally
  if (!collisionChecker->isInitialized()) {
    // fail-closed behaviour of an uninitialised checker (CC:106-107)
    std::printf("uninitialised: isInCollision -> %d\n", int(collisionChecker->isInCollision(std::vector<double>(8, 0.0), true, true)));
    return 3;
  }
  sqmp_shim::HotPath planner(collisionChecker);

  // ---- world: before setOcTree the map check reports free (CC:556-557) ----
  const std::vector<double> states = read_raw<double>(dir + "states.f64");
  const size_t S = states.size() / 8;
  std::vector<uint8_t> no_map(S), verdict(S), verdict_map(S);
  for (size_t s = 0; s < S; ++s)
    no_map[s] = collisionChecker->isInCollision(std::vector<double>(states.begin() + s * 8, states.begin() + s * 8 + 8), false, true);
  sqmp_shim::VoxelGrid g;
  {
    std::ifstream meta((dir + "grid_meta.txt").c_str());
    meta >> g.gx >> g.gy >> g.gz >> g.origin[0] >> g.origin[1] >> g.origin[2] >> g.res;
    g.bits = read_raw<uint32_t>(dir + "grid_bits.u32");
  }
  collisionChecker->setVoxelGrid(g);
  for (size_t s = 0; s < S; ++s) {
    const std::vector<double> q(states.begin() + s * 8, states.begin() + s * 8 + 8);
    verdict[s] = collisionChecker->isInCollision(q, true, true);
    verdict_map[s] = !planner.isConfigValid(q, false, true);
  }
  write_raw(dir + "out_no_map.u8", no_map);
  write_raw(dir + "out_verdict.u8", verdict);
  write_raw(dir + "out_verdict_map.u8", verdict_map);

  // ---- edges of 21 consecutive states (BS:6864-6895) ----
  std::vector<int32_t> edge_out;
  for (size_t e = 0; e + 21 <= S; e += 21) {
    Edge edge;
    for (size_t k = 0; k < 21; ++k)
      edge.joint_trajectory.push_back(std::vector<double>(states.begin() + (e + k) * 8, states.begin() + (e + k) * 8 + 8));
    int last_valid_node_idx = -7;
    const bool ok2 = planner.isEdgeValid(edge, last_valid_node_idx, false, true);
    const bool ok1 = planner.isEdgeValid(edge, false, true);
    edge_out.push_back(ok1);
    edge_out.push_back(ok2);
    edge_out.push_back(last_valid_node_idx);
  }
  write_raw(dir + "out_edges.i32", edge_out);

  // ---- disabled links (CC:90-102) ----
  std::vector<std::string> disabled;
  {
    std::ifstream f((dir + "disabled_links.txt").c_str());
    std::string n;
    while (f >> n) disabled.push_back(n);
  }
  collisionChecker->setDisabledLinkMapCollisions(disabled);
  std::vector<uint8_t> verdict_dis(S);
  for (size_t s = 0; s < S; ++s)
    verdict_dis[s] = collisionChecker->isInCollision(std::vector<double>(states.begin() + s * 8, states.begin() + s * 8 + 8), false, true);
  write_raw(dir + "out_verdict_disabled.u8", verdict_dis);
  collisionChecker->setDisabledLinkMapCollisions(std::vector<std::string>());

  // ---- getCollisions of the first colliding state (CC:123-133) ----
  {
    std::ofstream f((dir + "out_collisions.txt").c_str());
    for (size_t s = 0; s < S; ++s) {
      if (!verdict[s]) continue;
      std::vector<std::pair<std::string, std::string> > self;
      std::vector<std::string> map;
      collisionChecker->getCollisions(std::vector<double>(states.begin() + s * 8, states.begin() + s * 8 + 8), self, map);
      f << "state " << s << "\n";
      for (size_t i = 0; i < self.size(); ++i) f << "self " << self[i].first << " " << self[i].second << "\n";
      for (size_t i = 0; i < map.size(); ++i) f << "map " << map[i] << "\n";
      break;
    }
  }

  // ---- tree + near neighbours (BS:3298-3322, 4076-4133, 4272-4324) ----
  const std::vector<double> nodes = read_raw<double>(dir + "tree_nodes.f64"), cost = read_raw<double>(dir + "tree_cost.f64");
  const std::vector<double> queries = read_raw<double>(dir + "queries.f64");
  Rrt_star_tree tree;
  tree.name = "START";
  planner.resetTree(0);
  for (size_t n = 0; n < cost.size(); ++n) {
    Node node;
    node.node_id = int(n);
    node.parent_id = n ? 0 : -1;
    node.config.assign(nodes.begin() + n * 8, nodes.begin() + n * 8 + 8);
    node.cost_reach.total = cost[n];
    tree.nodes.push_back(node);
    if (!planner.insertNode(0, node)) {
      std::fprintf(stderr, "insertNode failed at %zu\n", n);
      return 4;
    }
  }
  tree.num_nodes = int(tree.nodes.size());
  std::vector<int32_t> nn_out, near_out;
  for (size_t k = 0; k * 8 < queries.size(); ++k) {
    const std::vector<double> x(queries.begin() + k * 8, queries.begin() + k * 8 + 8);
    const Node nn = planner.find_nearest_neighbour_interpolation(&tree, 0, x);
    nn_out.push_back(nn.node_id);
    Node x_new = tree.nodes[size_t(nn.node_id)];   // as after insertNode: x_new is in the tree, its id is excluded
    const std::vector<int> near = planner.find_near_vertices_interpolation(&tree, 0, x_new);
    near_out.push_back(int32_t(near.size()));
    near_out.insert(near_out.end(), near.begin(), near.end());
  }
  write_raw(dir + "out_nearest.i32", nn_out);
  write_raw(dir + "out_near.i32", near_out);

  // rewire moves a node (BS:5173-5193): node 3 jumps onto query 0
  {
    Node moved = tree.nodes[3];
    moved.config.assign(queries.begin(), queries.begin() + 8);
    tree.nodes[3] = moved;
    planner.updateNode(0, moved);
    std::vector<int32_t> after(1, planner.find_nearest_neighbour_interpolation(&tree, 0, moved.config).node_id);
    write_raw(dir + "out_nearest_after_update.i32", after);
  }

  // ---- the reference's exact signatures (BH:322-325): no slot, no mirror calls; the host tree is mutated like the
  //      planner mutates it (append = insertNode, move = rewire, cost change = recursiveNodeCostUpdate, clear = reset) ----
  {
    Rrt_star_tree t2;
    t2.name = "GOAL";
    std::vector<int32_t> lit;
    auto query_all = [&]() {
      for (size_t k = 0; k * 8 < queries.size() && k < 24; ++k) {
        const std::vector<double> x(queries.begin() + k * 8, queries.begin() + k * 8 + 8);
        const Node nn = planner.find_nearest_neighbour_interpolation(&t2, x);
        lit.push_back(nn.node_id);
        const std::vector<int> near = planner.find_near_vertices_interpolation(&t2, t2.nodes[size_t(nn.node_id)]);
        lit.push_back(int32_t(near.size()));
        lit.insert(lit.end(), near.begin(), near.end());
      }
    };
    for (size_t n = 0; n < cost.size() / 2; ++n) t2.nodes.push_back(tree.nodes[n]);
    query_all();                                                        // phase 0: first half of the tree
    for (size_t n = cost.size() / 2; n < cost.size(); ++n) t2.nodes.push_back(tree.nodes[n]);
    query_all();                                                        // phase 1: appended nodes
    for (size_t n = 5; n < t2.nodes.size(); n += 7) {                   // phase 2: rewires + cost propagation
      t2.nodes[n].config.assign(queries.begin() + (n % 16) * 8, queries.begin() + (n % 16) * 8 + 8);
      t2.nodes[n].cost_reach.total *= 0.5;
    }
    query_all();
    t2.nodes.resize(40);                                                // phase 3: planner reset, a shorter tree
    query_all();
    write_raw(dir + "out_literal.i32", lit);
    // isConfigValid(JntArray-like) (BS:6901-6908) and compute_edge_cost_interpolation (BS:4162-4242)
    struct JntArrayLike {
      std::vector<double> d;
      unsigned int rows() const { return static_cast<unsigned int>(d.size()); }
      double operator()(unsigned int i) const { return d[i]; }
    };
    std::vector<uint8_t> jv;
    std::vector<double> ec;
    planner.m_edge_cost_weights[1] = 0.5;
    planner.m_edge_cost_weights[4] = 3.0;
    for (size_t s = 0; s + 21 <= S && s < 21 * 40; s += 21) {
      JntArrayLike ja;
      ja.d.assign(states.begin() + s * 8, states.begin() + s * 8 + 8);
      jv.push_back(planner.isConfigValid(ja, true, true) ? 1 : 0);
      std::vector<std::vector<double> > traj;
      for (size_t k = 0; k < 21; ++k) traj.push_back(std::vector<double>(states.begin() + (s + k) * 8, states.begin() + (s + k) * 8 + 8));
      const std::vector<double> c3 = planner.compute_edge_cost_interpolation(traj);
      ec.insert(ec.end(), c3.begin(), c3.end());
    }
    write_raw(dir + "out_jntarray_valid.u8", jv);
    write_raw(dir + "out_edge_cost.f64", ec);
    // informed sampler through the reference-named members
    const double ilo[8] = {-5.0, -4.0, -3.0, -1.2, -1.7, -1.8, -2.4, -2.9}, ihi[8] = {5.0, 6.0, 3.0, 1.5, 2.6, 1.8, 2.4, 2.9};
    planner.setSamplingBox(std::vector<double>(ilo, ilo + 8), std::vector<double>(ihi, ihi + 8), 424242ull);
    planner.jointConfigEllipseInitialization(std::vector<double>(states.begin(), states.begin() + 8),
                                             std::vector<double>(states.begin() + 8 * 500, states.begin() + 8 * 501));
    std::vector<double> inf;
    for (int i = 0; i < 40; ++i) {
      const std::vector<double> q = planner.sampleJointConfigfromEllipse_Vector(9.0, 14.0);
      inf.insert(inf.end(), q.begin(), q.end());
    }
    write_raw(dir + "out_informed.f64", inf);
  }

  // ---- compute_FK (KM:278-305) ----
  std::vector<double> poses;
  for (size_t s = 0; s < S && s < 64; ++s) {
    const std::vector<double> p = planner.compute_FK(std::vector<double>(states.begin() + s * 8, states.begin() + s * 8 + 8));
    poses.insert(poses.end(), p.begin(), p.end());
  }
  write_raw(dir + "out_fk.f64", poses);

  // ---- sampleJointConfig_Vector (BS:3881-3930 over CL:1120-1188): 600 draws = three device batches ----
  {
    const double lo[8] = {-5.0, -4.0, -3.0, -1.2, -1.7, -1.8, -2.4, -2.9}, hi[8] = {5.0, 6.0, 3.0, 1.5, 2.6, 1.8, 2.4, 2.9};
    planner.setSamplingBox(std::vector<double>(lo, lo + 8), std::vector<double>(hi, hi + 8), 20240607ull);
    std::vector<double> samples;
    for (int i = 0; i < 600; ++i) {
      const std::vector<double> q = planner.sampleJointConfig_Vector();
      samples.insert(samples.end(), q.begin(), q.end());
    }
    write_raw(dir + "out_samples.f64", samples);
  }
  delete collisionChecker;
  std::printf("test_shim ok: %zu states, %zu nodes, %zu queries\n", S, cost.size(), queries.size() / 8);
  return 0;
}
