// sqmp_shim.hpp — C++ host side of the drop-in: the reference's collision-check, FK and near-neighbour interfaces,
// same names, argument meaning and error behaviour, served by the C ABI of libsqmp_b200.so (include/sqmp_b200.h).
//
//   reference interface (file:line in /root/reference)                                         here
//   ----------------------------------------------------------------------------------------------------------------
//   CollisionChecker (collision_checker.hpp:35-133): isInCollision, setOcTree,                 sqmp_shim::CollisionChecker
//       setDisabledLinkMapCollisions, getCollisions
//   BiRRTstarPlanner::isEdgeValid x2 / isConfigValid x2 (birrt_star.cpp:6864-6908)             sqmp_shim::HotPath
//   BiRRTstarPlanner::find_nearest_neighbour_interpolation (birrt_star.cpp:4076-4133)          sqmp_shim::HotPath
//   BiRRTstarPlanner::find_near_vertices_interpolation (birrt_star.cpp:4272-4324)              sqmp_shim::HotPath
//   BiRRTstarPlanner::insertNode and the rewire / cost-update sites (3298-3322, 5173-5193,     sqmp_shim::HotPath
//       5518-5520) — kept in step with the device mirror of the tree
//   KDLRobotModel::compute_FK (kdl_kuka_model.cpp:278-305)                                     sqmp_shim::HotPath::compute_FK
//   Edge / Node / CostsReach / Rrt_star_tree (data_structs.h:11-54)                            the reference's own header when
//                                                                                              SQMP_SHIM_USE_REFERENCE_STRUCTS
//                                                                                              is defined, else the mirror below
//
// Header-only, C++11 like the reference (birrt_star_algorithm/CMakeLists.txt).  No CPU fallback: if the library or
// the GPU is missing the checker stays uninitialised and isInCollision answers true (fails closed, CC:106-107).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "sqmp_b200.h"
#include "sqmp_informed.hpp"

#ifdef SQMP_SHIM_USE_REFERENCE_STRUCTS
#include <planner_data_structures/data_structs.h>
#else
// Mirror of data_structs.h:11-54 — only so that the shim compiles without the reference tree (tests).
struct Edge {
  int edge_id;
  int root_node_id;
  int child_node_id;
  std::vector<std::vector<double> > joint_trajectory;
  std::vector<std::vector<double> > ee_trajectory;
  double cost_g;
};
struct CostsReach {
  double revolute;
  double prismatic;
  double total;
};
struct Node {
  int node_id;
  int parent_id;
  std::vector<double> config;
  std::vector<double> ee_pose;
  CostsReach cost_h;
  CostsReach cost_reach;
  std::vector<Edge> outgoing_edges;
  bool operator<(const Node &o) const { return cost_reach.total < o.cost_reach.total; }
};
struct Rrt_star_tree {
  std::string name;
  std::vector<Node> nodes;
  int num_nodes;
  int num_edges;
  int num_rewire_operations;
};
#endif

namespace sqmp_shim {

// ---------------------------------------------------------------------------------------------------------------
// robot description file ("sqmpdesc 2", written by tools/export_robot_desc.py from the reference URDF + SRDF)
// ---------------------------------------------------------------------------------------------------------------
struct RobotDescription {
  sqmp_robot_desc desc;
  std::vector<std::string> link_names;
  std::vector<std::string> octomap_disabled;
  std::vector<int32_t> sph_seg, sph_link;
  std::vector<double> sph_center, sph_radius;

  bool load(const std::string &path, std::string *err) {
    std::ifstream in(path.c_str());
    if (!in) return fail(err, "cannot open " + path);
    std::string tok;
    int version = 0;
    if (!(in >> tok >> version) || tok != "sqmpdesc" || version != 2) return fail(err, "not a sqmpdesc 2 file");
    desc = sqmp_robot_desc();
    int seg = 0, pair_rows = 0, pair_rows_read = 0;
    while (in >> tok) {
      if (tok == "n_seg") {
        in >> desc.n_seg;
        if (desc.n_seg < 1 || desc.n_seg > SQMP_MAX_SEG) return fail(err, "n_seg out of range");
      } else if (tok == "seg") {
        if (seg >= SQMP_MAX_SEG) return fail(err, "too many segments");
        in >> desc.seg_type[seg] >> desc.seg_qidx[seg];
        for (int i = 0; i < 12; ++i) in >> desc.seg_pre[seg][i];
        for (int i = 0; i < 3; ++i) in >> desc.seg_axis[seg][i];
        ++seg;
      } else if (tok == "root_offset") {
        for (int i = 0; i < 3; ++i) in >> desc.root_offset[i];
      } else if (tok == "ee_tip") {
        for (int i = 0; i < 12; ++i) in >> desc.ee_tip[i];
      } else if (tok == "n_links") {
        size_t n = 0;
        in >> n;
        link_names.assign(n, std::string());
      } else if (tok == "link") {
        size_t i = 0;
        std::string name;
        in >> i >> name;
        if (i >= link_names.size()) return fail(err, "link index out of range");
        link_names[i] = name;
      } else if (tok == "octomap_disabled") {
        size_t n = 0;
        in >> n;
        octomap_disabled.resize(n);
        for (size_t i = 0; i < n; ++i) in >> octomap_disabled[i];
      } else if (tok == "n_spheres") {
        in >> desc.n_spheres;
      } else if (tok == "sphere") {
        int32_t sg = 0, lk = 0;
        double c[3], r = 0;
        in >> sg >> lk >> c[0] >> c[1] >> c[2] >> r;
        sph_seg.push_back(sg);
        sph_link.push_back(lk);
        sph_center.insert(sph_center.end(), c, c + 3);
        sph_radius.push_back(r);
      } else if (tok == "n_link_pair_rows") {
        in >> pair_rows;
      } else if (tok == "link_pairs") {   // selfCollisionPairs at link level (CC:353-393): row index, 64-bit mask in hex
        int row = -1;
        std::string hex;
        in >> row >> hex;
        if (row < 0 || row >= SQMP_MAX_LINKS) return fail(err, "link_pairs row out of range");
        desc.link_pair_mask[row] = std::strtoull(hex.c_str(), nullptr, 16);
        ++pair_rows_read;
      } else if (tok == "self_const_collision") {
        in >> desc.self_const_collision;
      } else {
        return fail(err, "unknown key " + tok);
      }
      if (!in) return fail(err, "truncated value for key " + tok);
    }
    if (seg != desc.n_seg || int(sph_radius.size()) != desc.n_spheres || pair_rows_read != pair_rows)
      return fail(err, "counts do not match their declarations");
    desc.sph_seg = sph_seg.data();
    desc.sph_link = sph_link.data();
    desc.sph_center = sph_center.data();
    desc.sph_radius = sph_radius.data();
    return true;
  }

  int link_index(const std::string &name) const {
    for (size_t i = 0; i < link_names.size(); ++i)
      if (link_names[i] == name) return int(i);
    return -1;
  }

 private:
  static bool fail(std::string *err, const std::string &what) {
    if (err) *err = what;
    return false;
  }
};

// Dense occupancy as the library takes it: bit (k*gy+j)*gx+i = voxel origin + [i,i+1]x[j,j+1]x[k,k+1]*res.
struct VoxelGrid {
  std::vector<uint32_t> bits;
  int gx, gy, gz;
  double origin[3];
  double res;
  VoxelGrid() : gx(0), gy(0), gz(0), res(0.0) { origin[0] = origin[1] = origin[2] = 0.0; }
  void resize(int nx, int ny, int nz) {
    gx = nx; gy = ny; gz = nz;
    bits.assign((size_t(nx) * ny * nz + 31) / 32, 0u);
  }
  void set(int i, int j, int k) {
    const size_t b = (size_t(k) * gy + j) * gx + i;
    bits[b >> 5] |= 1u << (b & 31);
  }
};

// Octomap binary tree (.bt file body = octomap_msgs binary map data) -> dense grid, without liboctomap.
// Stream format (octomap OcTree::readBinaryNode): per inner node two bytes, 2 bits per child,
// (bit 2i, bit 2i+1) = (1,0) free leaf, (0,1) occupied leaf, (1,1) inner node, (0,0) unknown; depth-first; child i at
// x = i&1, y = (i>>1)&1, z = (i>>2)&1; 16 levels; key = floor(coord/res) + 32768.  Pruned leaves are expanded; unknown
// and free space do not collide (what FCL does with an OcTree).  floor_distance > 0 adds the planner node's floor
// patch around (pose_x, pose_y) (squirrel_8dof_planner.cpp:886-904).  Same logic as octomap_io.py (tests compare them).
inline bool read_bt(const std::string &path, VoxelGrid &grid, std::string *err = NULL, bool with_floor = false,
                    double pose_x = 0.0, double pose_y = 0.0, double floor_distance = 0.0) {
  std::ifstream in(path.c_str(), std::ios::binary);
  if (!in) { if (err) *err = "cannot open " + path; return false; }
  std::string line;
  double res = 0.0;
  long size = -1;
  bool data_seen = false;
  while (std::getline(in, line)) {
    if (line == "data") { data_seen = true; break; }
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "res") ls >> res;
    else if (key == "size") ls >> size;
    else if (key == "id") { std::string id; ls >> id; if (id != "OcTree") { if (err) *err = "unsupported tree id " + id; return false; } }
  }
  if (!data_seen || !(res > 0.0) || size < 0) { if (err) *err = "not an octomap .bt stream"; return false; }
  std::vector<unsigned char> body((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  struct Leaf { long k[3]; long size; };
  struct Item { long k[3]; int depth; };
  std::vector<Leaf> occ;
  long n_nodes = size > 0 ? 1 : 0;
  size_t off = 0;
  std::vector<Item> stack;
  if (size > 0) { Item root = {{0, 0, 0}, 0}; stack.push_back(root); }
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    if (off + 2 > body.size()) { if (err) *err = "truncated data section"; return false; }
    const unsigned bits = unsigned(body[off]) | (unsigned(body[off + 1]) << 8);
    off += 2;
    const long half = 1L << (16 - it.depth - 1);
    std::vector<Item> inner;
    for (int i = 0; i < 8; ++i) {
      const unsigned b0 = (bits >> (2 * i)) & 1u, b1 = (bits >> (2 * i + 1)) & 1u;
      if (!b0 && !b1) continue;
      ++n_nodes;
      const long cx = it.k[0] + (i & 1) * half, cy = it.k[1] + ((i >> 1) & 1) * half, cz = it.k[2] + ((i >> 2) & 1) * half;
      if (b0 && b1) {
        if (it.depth + 1 >= 16) { if (err) *err = "inner node below the finest level"; return false; }
        Item c = {{cx, cy, cz}, it.depth + 1};
        inner.push_back(c);
      } else if (b1) {
        Leaf l = {{cx, cy, cz}, half};
        occ.push_back(l);
      }
    }
    for (size_t j = inner.size(); j-- > 0;) stack.push_back(inner[j]);   // depth-first, ascending child order
  }
  if (off != body.size() || (size > 0 && n_nodes != size)) { if (err) *err = "tree does not match its header"; return false; }
  if (with_floor && floor_distance > 0.0) {
    const long kx = long(std::floor(pose_x / res)) + 32768, ky = long(std::floor(pose_y / res)) + 32768;
    const long kz = long(std::floor(-0.5)) + 32768, n = long(floor_distance / res);
    for (long x = kx - n; x <= kx + n; ++x)
      for (long y = ky - n; y <= ky + n; ++y) { Leaf l = {{x, y, kz}, 1}; occ.push_back(l); }
  }
  if (occ.empty()) { if (err) *err = "no occupied leaves"; return false; }
  long lo[3] = {occ[0].k[0], occ[0].k[1], occ[0].k[2]}, hi[3] = {lo[0], lo[1], lo[2]};
  for (size_t j = 0; j < occ.size(); ++j)
    for (int a = 0; a < 3; ++a) {
      if (occ[j].k[a] < lo[a]) lo[a] = occ[j].k[a];
      if (occ[j].k[a] + occ[j].size > hi[a]) hi[a] = occ[j].k[a] + occ[j].size;
    }
  for (int a = 0; a < 3; ++a) { lo[a] -= 1; hi[a] += 1; }   // one voxel of margin, like octomap_io.to_voxel_world
  grid.res = res;
  for (int a = 0; a < 3; ++a) grid.origin[a] = double(lo[a] - 32768) * res;
  grid.resize(int(hi[0] - lo[0]), int(hi[1] - lo[1]), int(hi[2] - lo[2]));
  for (size_t j = 0; j < occ.size(); ++j)
    for (long z = occ[j].k[2]; z < occ[j].k[2] + occ[j].size; ++z)
      for (long y = occ[j].k[1]; y < occ[j].k[1] + occ[j].size; ++y)
        for (long x = occ[j].k[0]; x < occ[j].k[0] + occ[j].size; ++x) grid.set(int(x - lo[0]), int(y - lo[1]), int(z - lo[2]));
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// CollisionChecker  (collision_checker.hpp:35-133)
// ---------------------------------------------------------------------------------------------------------------
class CollisionChecker {
 public:
  // CC:66-74 / 156-172: the reference builds its link model from the URDF/SRDF on the parameter server; here the
  // pre-folded description comes from a file (tools/export_robot_desc.py).  Failure leaves the checker uninitialised.
  // CC:66-74 as the reference calls it — `new CollisionChecker()` (BS:73): the description path comes from
  // setDescriptionPath() (the analogue of the robot_description ROS parameter the reference reads, CC:178) or the
  // SQMP_ROBOT_DESC environment variable, the device from SQMP_DEVICE (default 0).
  CollisionChecker() : ctx_(NULL), initialized_(false) {
    const char *env = std::getenv("SQMP_ROBOT_DESC"), *dev = std::getenv("SQMP_DEVICE");
    const std::string path = !description_path().empty() ? description_path() : std::string(env ? env : "");
    if (path.empty()) {
      std::fprintf(stderr, "CollisionChecker: no robot description (CollisionChecker::setDescriptionPath / SQMP_ROBOT_DESC)\n");
      return;
    }
    open(path, dev ? std::atoi(dev) : 0);
  }
  static void setDescriptionPath(const std::string &path) { description_path() = path; }
  explicit CollisionChecker(const std::string &desc_path, int device = 0) : ctx_(NULL), initialized_(false) { open(desc_path, device); }
 private:
  static std::string &description_path() {
    static std::string p;
    return p;
  }
  void open(const std::string &desc_path, int device) {
    std::string err;
    if (!robot_.load(desc_path, &err)) {
      std::fprintf(stderr, "CollisionChecker: %s\n", err.c_str());
      return;
    }
    int code = 0;
    ctx_ = sqmp_create(device, &robot_.desc, &code);
    if (!ctx_) {
      std::fprintf(stderr, "CollisionChecker: sqmp_create failed: %s\n", sqmp_strerror(code));
      return;
    }
    initialized_ = true;
    setDisabledLinkMapCollisions(std::vector<std::string>());
  }
 public:
  ~CollisionChecker() { sqmp_destroy(ctx_); }

  bool isInitialized() const { return initialized_; }
  sqmp_ctx *context() const { return ctx_; }
  const RobotDescription &robot() const { return robot_; }

  // CC:76-88.  Any octomap-like tree: begin_leafs()/end_leafs() iterators with getX/Y/Z(), getSize(), and
  // isNodeOccupied(*it); getResolution(); getMetricMin/Max(x,y,z).  NULL is ignored (CC:78-79); deep copy; the
  // collision object's z = -0.02 offset (CC:87) is folded into the grid origin.
  template <class OcTreeT>
  void setOcTree(const OcTreeT *tree) {
    if (!tree || !initialized_) return;
    VoxelGrid g;
    const double res = tree->getResolution();
    double lo[3], hi[3];
    tree->getMetricMin(lo[0], lo[1], lo[2]);
    tree->getMetricMax(hi[0], hi[1], hi[2]);
    int n[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = res * double((long long)(lo[a] / res) - 1);
      n[a] = int((hi[a] - lo[a]) / res) + 2;
      g.origin[a] = lo[a];
    }
    g.res = res;
    g.resize(n[0], n[1], n[2]);
    for (typename OcTreeT::leaf_iterator it = tree->begin_leafs(), end = tree->end_leafs(); it != end; ++it) {
      if (!tree->isNodeOccupied(*it)) continue;   // unknown and free space do not collide
      const double h = 0.5 * it.getSize();        // a pruned leaf covers (size/res)^3 finest voxels
      const int i0 = int((it.getX() - h - lo[0]) / res + 0.5), i1 = int((it.getX() + h - lo[0]) / res + 0.5);
      const int j0 = int((it.getY() - h - lo[1]) / res + 0.5), j1 = int((it.getY() + h - lo[1]) / res + 0.5);
      const int k0 = int((it.getZ() - h - lo[2]) / res + 0.5), k1 = int((it.getZ() + h - lo[2]) / res + 0.5);
      for (int k = k0; k < k1; ++k)
        for (int j = j0; j < j1; ++j)
          for (int i = i0; i < i1; ++i)
            if (i >= 0 && j >= 0 && k >= 0 && i < g.gx && j < g.gy && k < g.gz) g.set(i, j, k);
    }
    setVoxelGrid(g);
  }
  // A further checker on the same device (one per planner thread) borrows the voxel fields of `other` (CC:76-88 done once).
  bool shareOcTreeWith(const CollisionChecker &other) {
    return initialized_ && other.initialized_ && sqmp_share_voxel_grid(ctx_, other.ctx_) == SQMP_OK;
  }
  // Same, from an occupancy that is already dense (tests, synthetic worlds, .bt ingest).
  void setVoxelGrid(const VoxelGrid &g) {
    if (!initialized_ || g.bits.empty()) return;
    const double origin[3] = {g.origin[0], g.origin[1], g.origin[2] - 0.02};   // CC:87
    if (sqmp_set_voxel_grid(ctx_, g.bits.data(), g.gx, g.gy, g.gz, origin, g.res) != SQMP_OK)
      std::fprintf(stderr, "CollisionChecker::setOcTree: %s\n", sqmp_last_error(ctx_));
  }

  // CC:90-102 (+ the SRDF "octomap" disables, CC:408-416)
  void setDisabledLinkMapCollisions(const std::vector<std::string> &linksDisabled) {
    if (!initialized_) return;
    uint64_t mask = 0;
    for (size_t i = 0; i < linksDisabled.size(); ++i) add_link(mask, linksDisabled[i]);
    for (size_t i = 0; i < robot_.octomap_disabled.size(); ++i) add_link(mask, robot_.octomap_disabled[i]);
    sqmp_set_disabled_links(ctx_, mask);
  }

  // CC:104-121: true = in collision; true when uninitialised; false when both flags are off.
  bool isInCollision(const std::vector<double> &jointPositions, const bool checkSelfCollision,
                     const bool checkMapCollision) {
    if (!initialized_) return true;
    if (!checkSelfCollision && !checkMapCollision) return false;
    if (jointPositions.size() < size_t(SQMP_DOF)) return true;
    uint32_t word = 0;
    if (sqmp_check_states(ctx_, jointPositions.data(), 1, flags(checkSelfCollision, checkMapCollision), &word) != SQMP_OK)
      return true;   // fail closed
    return (word & 1u) != 0;
  }

  // CC:123-133, 594-631: names of colliding link pairs (i < j, sorted) and of links hitting the map (sorted).
  void getCollisions(const std::vector<double> &jointPositions,
                     std::vector<std::pair<std::string, std::string> > &selfCollisions,
                     std::vector<std::string> &mapCollisions) {
    selfCollisions.clear();
    mapCollisions.clear();
    if (!initialized_ || jointPositions.size() < size_t(SQMP_DOF)) return;
    uint8_t link_hit[SQMP_MAX_LINKS];
    std::vector<uint8_t> pair_hit(size_t(SQMP_MAX_LINKS) * SQMP_MAX_LINKS, 0);
    if (sqmp_get_collisions(ctx_, jointPositions.data(), link_hit, pair_hit.data()) != SQMP_OK) return;
    const size_t nl = robot_.link_names.size() < size_t(SQMP_MAX_LINKS) ? robot_.link_names.size() : size_t(SQMP_MAX_LINKS);
    for (size_t a = 0; a < nl; ++a)   // (i < j) in index order, like selfCollisionPairs (CC:387-390)
      for (size_t b = a + 1; b < nl; ++b)
        if (pair_hit[a * SQMP_MAX_LINKS + b]) selfCollisions.push_back(std::make_pair(robot_.link_names[a], robot_.link_names[b]));
    std::map<std::string, bool> names;
    for (size_t i = 0; i < robot_.link_names.size() && i < 64; ++i)
      if (link_hit[i]) names[robot_.link_names[i]] = true;
    for (std::map<std::string, bool>::const_iterator it = names.begin(); it != names.end(); ++it)
      mapCollisions.push_back(it->first);
  }

  static unsigned flags(bool self, bool map) { return (self ? SQMP_CHECK_SELF : 0u) | (map ? SQMP_CHECK_MAP : 0u); }

 private:
  void add_link(uint64_t &mask, const std::string &name) const {
    const int i = robot_.link_index(name);
    if (i >= 0 && i < 64) mask |= uint64_t(1) << i;
  }
  CollisionChecker(const CollisionChecker &);
  CollisionChecker &operator=(const CollisionChecker &);

  RobotDescription robot_;
  sqmp_ctx *ctx_;
  bool initialized_;
};

// ---------------------------------------------------------------------------------------------------------------
// The members of BiRRTstarPlanner on the hot path (birrt_star.h:103-107, 322-325)
// ---------------------------------------------------------------------------------------------------------------
class HotPath {
 public:
  HotPath(CollisionChecker *checker, double near_threshold_interpolation = 1.5 /* PY:4 */, int max_near_nodes = 20 /* PY:6 */)
      : collisionChecker(checker), m_near_threshold_interpolation(near_threshold_interpolation),
        m_max_near_nodes(max_near_nodes) {}

  // BS:6897-6908
  bool isConfigValid(const std::vector<double> &config, bool check_self_collision, bool check_map_collision) {
    return !collisionChecker->isInCollision(config, check_self_collision, check_map_collision);
  }

  // BS:6901-6908: the KDL::JntArray overload — any array type with rows() and operator()(i) (SP:763-839 call sites)
  template <class JntArrayT>
  auto isConfigValid(const JntArrayT &config, bool check_self_collision, bool check_map_collision)
      -> decltype(config.rows(), config(0), bool()) {
    std::vector<double> v(size_t(config.rows()));
    for (size_t i = 0; i < v.size(); ++i) v[i] = config(static_cast<unsigned int>(i));
    return isConfigValid(v, check_self_collision, check_map_collision);
  }

  // BS:4162-4242 compute_edge_cost_interpolation: {total, revolute, prismatic}; the total weights every squared
  // coordinate difference with m_edge_cost_weights (BS:109-113: all 1.0 unless set), the other two do not.
  // m_joint_is_revolute stands for the reference's getTypeName() == "RotAxis" test on the KDL chain (BS:4193, 4203).
  std::vector<double> compute_edge_cost_interpolation(const std::vector<std::vector<double> > &joint_traj) const {
    std::vector<double> cost_list(3, 0.0);
    for (size_t wp = 0; wp + 1 < joint_traj.size(); ++wp) {
      double wp_dist_total = 0.0, wp_dist_rot = 0.0, wp_dist_prism = 0.0;
      for (size_t joint_idx = 0; joint_idx < joint_traj[wp].size() && joint_idx < size_t(SQMP_DOF); ++joint_idx) {
        const double coordinate_dist = (joint_traj[wp + 1][joint_idx] - joint_traj[wp][joint_idx]) *
                                       (joint_traj[wp + 1][joint_idx] - joint_traj[wp][joint_idx]);
        wp_dist_total += coordinate_dist * m_edge_cost_weights[joint_idx];
        if (m_joint_is_revolute[joint_idx]) wp_dist_rot += coordinate_dist;
        else wp_dist_prism += coordinate_dist;
      }
      cost_list[0] += std::sqrt(wp_dist_total);
      cost_list[1] += std::sqrt(wp_dist_rot);
      cost_list[2] += std::sqrt(wp_dist_prism);
    }
    return cost_list;
  }
  double m_edge_cost_weights[SQMP_DOF] = {1, 1, 1, 1, 1, 1, 1, 1};
  int m_joint_is_revolute[SQMP_DOF] = {0, 0, 1, 1, 1, 1, 1, 1};   // x, y prismatic; theta and the five arm joints revolute (URDF:30-66, 615-759)

  // BS:6864-6876
  bool isEdgeValid(const Edge &tree_edge, bool check_self_collision, bool check_map_collision) {
    int unused = 0;
    return isEdgeValid(tree_edge, unused, check_self_collision, check_map_collision);
  }
  // BS:6878-6895: last_valid_node_idx = index of the last state before the first colliding one (0 if none), n-1 if valid.
  // One library call for the whole trajectory instead of one isInCollision per state.
  bool isEdgeValid(const Edge &tree_edge, int &last_valid_node_idx, bool check_self_collision, bool check_map_collision) {
    last_valid_node_idx = 0;
    const size_t n = tree_edge.joint_trajectory.size();
    if (n == 0) return true;
    if (!collisionChecker->isInitialized()) return false;
    if (!check_self_collision && !check_map_collision) {
      last_valid_node_idx = int(n) - 1;
      return true;
    }
    scratch_q_.resize(n * SQMP_DOF);
    for (size_t s = 0; s < n; ++s)
      for (int i = 0; i < SQMP_DOF; ++i) scratch_q_[s * SQMP_DOF + i] = tree_edge.joint_trajectory[s][i];   // BS:6868-6870
    scratch_mask_.assign((n + 31) / 32, 0u);
    if (sqmp_check_states(collisionChecker->context(), scratch_q_.data(), int64_t(n),
                          CollisionChecker::flags(check_self_collision, check_map_collision), scratch_mask_.data()) != SQMP_OK)
      return false;
    for (size_t s = 0; s < n; ++s)
      if ((scratch_mask_[s >> 5] >> (s & 31)) & 1u) {
        last_valid_node_idx = s > 0 ? int(s) - 1 : 0;
        return false;
      }
    last_valid_node_idx = int(n) - 1;
    return true;
  }

  // Batched form for a planner front-end that submits all candidate edges of an iteration at once: edge e is the
  // interpolation of q_start[e] -> q_end[e] over num_points+1 states (BS:4461-4478); first_invalid[e] = -1 if valid.
  bool areEdgesValid(const std::vector<double> &q_start, const std::vector<double> &q_end, int num_points,
                     bool check_self_collision, bool check_map_collision, std::vector<int32_t> &first_invalid) {
    const int64_t E = int64_t(q_start.size() / SQMP_DOF);
    first_invalid.assign(size_t(E), 0);
    if (!collisionChecker->isInitialized() || q_end.size() != q_start.size()) return false;
    scratch_mask_.assign(size_t((E * (num_points + 1) + 31) / 32), 0u);
    return sqmp_check_edges(collisionChecker->context(), q_start.data(), q_end.data(), E, num_points,
                            CollisionChecker::flags(check_self_collision, check_map_collision), scratch_mask_.data(),
                            first_invalid.data()) == SQMP_OK;
  }

  // ---- tree mirror: call next to the reference's own mutation of `tree` ------------------------------------------
  // `slot` selects the device tree (0 = start tree, 1 = goal tree, ...).
  bool resetTree(int slot) { return sqmp_tree_reset(collisionChecker->context(), slot) == SQMP_OK; }
  // BS:3298-3322 insertNode: appends; the id the library returns equals node.node_id (node_id == index invariant).
  bool insertNode(int slot, const Node &node) {
    int32_t id = -1;
    return sqmp_tree_append(collisionChecker->context(), slot, node.config.data(), node.cost_reach.total, &id) == SQMP_OK &&
           id == node.node_id;
  }
  // BS:5173-5193 rewire (config / cost of an existing node change), BS:5518-5520 recursiveNodeCostUpdate
  bool updateNode(int slot, const Node &node) {
    return sqmp_tree_update(collisionChecker->context(), slot, node.node_id, node.config.data(), node.cost_reach.total) == SQMP_OK;
  }

  // BS:4076-4133: returns the Node by value like the reference (id 0 when nothing is closer than 10000.0).
  Node find_nearest_neighbour_interpolation(Rrt_star_tree *tree, int slot, std::vector<double> x_rand) {
    int32_t id = 0;
    if (sqmp_nn_nearest(collisionChecker->context(), slot, x_rand.data(), 1, &id) != SQMP_OK) id = 0;
    return tree->nodes[size_t(id)];
  }

  // BS:4272-4324.  The reference returns every near id sorted by cost; its callers only ever walk the first
  // m_max_near_nodes (BS:4618-4623, 2823) or the last m_max_near_nodes from the end (BS:5075-5089).  The returned
  // vector has the reference's layout for both walks: if count <= 2*max_near it is the full sorted list, otherwise the
  // first max_near ids, then (count - 2*max_near) entries of -1 that no caller touches, then the last max_near ids.
  std::vector<int> find_near_vertices_interpolation(Rrt_star_tree * /*tree*/, int slot, const Node &x_new) {
    const int k = m_max_near_nodes;
    std::vector<int32_t> low(size_t(k), -1), high(size_t(k), -1);
    int32_t count = 0, excl = x_new.node_id;
    std::vector<int> out;
    if (sqmp_nn_radius(collisionChecker->context(), slot, x_new.config.data(), &excl, 1, m_near_threshold_interpolation, k,
                       low.data(), high.data(), &count) != SQMP_OK)
      return out;
    out.assign(size_t(count), -1);
    for (int t = 0; t < k && t < count; ++t) {
      out[size_t(t)] = low[size_t(t)];
      out[size_t(count - 1 - t)] = high[size_t(t)];
    }
    return out;
  }

  // ---- the two search functions with the reference's EXACT signatures (birrt_star.h:322-325): no slot argument and no
  //      mirror calls at insertNode / rewire / recursiveNodeCostUpdate.  The device mirror of `tree` is brought up to
  //      date lazily, right before the search, from a host shadow of what it holds: nodes appended since the last call
  //      go up in one batch, nodes whose config or cost_reach.total changed (rewire, BS:5173-5193; cost propagation,
  //      BS:5518-5520) are rewritten, a tree that shrank (reset_planner, BS:1411-1619) is rebuilt.  The scan is one
  //      pass over the host tree — less work than the reference's own O(N) distance loop it replaces.
  Node find_nearest_neighbour_interpolation(Rrt_star_tree *tree, std::vector<double> x_rand) {
    const int slot = sync_mirror(tree);
    if (slot < 0) return tree->nodes.empty() ? Node() : tree->nodes[0];
    return find_nearest_neighbour_interpolation(tree, slot, x_rand);
  }
  std::vector<int> find_near_vertices_interpolation(Rrt_star_tree *tree, Node x_new) {
    const int slot = sync_mirror(tree);
    if (slot < 0) return std::vector<int>();
    return find_near_vertices_interpolation(tree, slot, x_new);
  }
  // device tree slot of a host tree (assigned on first use), or -1 when the mirror could not be updated
  int sync_mirror(const Rrt_star_tree *tree) {
    Mirror &m = mirrors_[tree];
    sqmp_ctx *c = collisionChecker->context();
    if (m.slot < 0) {   // first use: a slot of its own, counted down from the top (explicit slots count up from 0)
      if (next_slot_ >= SQMP_MAX_TREES) return -1;
      m.slot = SQMP_MAX_TREES - 1 - next_slot_++;
      if (sqmp_tree_reset(c, m.slot) != SQMP_OK) return -1;
    }
    const size_t n = tree->nodes.size();
    if (n < m.cost.size()) {   // the planner was reset: start over
      if (sqmp_tree_reset(c, m.slot) != SQMP_OK) return -1;
      m.q.clear();
      m.cost.clear();
    }
    for (size_t i = 0; i < m.cost.size(); ++i) {   // rewired nodes / propagated costs
      const Node &nd = tree->nodes[i];
      bool same = nd.cost_reach.total == m.cost[i];
      for (int j = 0; same && j < SQMP_DOF; ++j) same = nd.config[size_t(j)] == m.q[i * SQMP_DOF + size_t(j)];
      if (same) continue;
      if (sqmp_tree_update(c, m.slot, int32_t(i), nd.config.data(), nd.cost_reach.total) != SQMP_OK) return -1;
      for (int j = 0; j < SQMP_DOF; ++j) m.q[i * SQMP_DOF + size_t(j)] = nd.config[size_t(j)];
      m.cost[i] = nd.cost_reach.total;
    }
    if (n > m.cost.size()) {   // appended nodes (insertNode, BS:3298-3322), one batch
      const size_t first = m.cost.size();
      m.q.resize(n * SQMP_DOF);
      m.cost.resize(n);
      for (size_t i = first; i < n; ++i) {
        for (int j = 0; j < SQMP_DOF; ++j) m.q[i * SQMP_DOF + size_t(j)] = tree->nodes[i].config[size_t(j)];
        m.cost[i] = tree->nodes[i].cost_reach.total;
      }
      int32_t first_id = -1;
      if (sqmp_tree_append_batch(c, m.slot, &m.q[first * SQMP_DOF], &m.cost[first], int64_t(n - first), &first_id) != SQMP_OK ||
          first_id != int32_t(first)) {
        m.q.resize(first * SQMP_DOF);
        m.cost.resize(first);
        return -1;
      }
    }
    return m.slot;
  }

  // KM:278-305: [x y z qx qy qz qw] of the SRDF chain tip (hand_wrist_link).
  std::vector<double> compute_FK(const std::vector<double> &chain_config) {
    std::vector<double> pose(7, 0.0);
    if (chain_config.size() >= size_t(SQMP_DOF)) sqmp_fk(collisionChecker->context(), chain_config.data(), 1, pose.data());
    return pose;
  }

  // BS:3881-3930 sampleJointConfig_Vector / BS:3832-3878 sampleJointConfig_JntArray over getRandomConf (CL:1120-1188):
  // uniform in [q_min, q_max] (base joints from m_env_size_x / m_env_size_y), end-effector above the platform.
  // Samples are drawn 256 at a time on the device (sqmp_sample_configs) and handed out one by one; sample i of the
  // stream is a pure function of (seed, i) — call setSamplingBox again to restart the stream.
  void setSamplingBox(const std::vector<double> &q_min, const std::vector<double> &q_max, uint64_t seed) {
    for (int j = 0; j < SQMP_DOF; ++j) {
      sample_lo_[j] = j < int(q_min.size()) ? q_min[size_t(j)] : 0.0;
      sample_hi_[j] = j < int(q_max.size()) ? q_max[size_t(j)] : 0.0;
    }
    sample_seed_ = seed;
    sample_next_ = 0;
    sample_pool_.clear();
    sample_pos_ = 0;
  }
  std::vector<double> sampleJointConfig_Vector() {
    if (sample_pos_ >= sample_pool_.size() / SQMP_DOF) {
      sample_pool_.assign(size_t(256) * SQMP_DOF, 0.0);
      if (sqmp_sample_configs(collisionChecker->context(), sample_seed_, sample_next_, 256, sample_lo_, sample_hi_, 64,
                              sample_pool_.data(), NULL) != SQMP_OK)
        return std::vector<double>();
      sample_next_ += 256;
      sample_pos_ = 0;
    }
    const double *q = &sample_pool_[sample_pos_ * SQMP_DOF];
    sample_pos_ += 1;
    return std::vector<double>(q, q + SQMP_DOF);
  }

  // BS:3607-3829 sampleJointConfigfromEllipse (informed sampling once a solution exists, BS:1168): the ellipse is set up
  // from start and goal like jointConfigEllipseInitialization (BS:3472-3604); samples are drawn on the device.
  void jointConfigEllipseInitialization(const std::vector<double> &config_start, const std::vector<double> &config_goal) {
    ellipse_.init(config_start.data(), config_goal.data(), m_joint_is_revolute);
    informed_next_ = 0;
  }
  std::vector<double> sampleJointConfigfromEllipse_Vector(double cost_best_revolute, double cost_best_prismatic) {
    sqmp_informed el;
    ellipse_.fill(el, cost_best_revolute, cost_best_prismatic, sample_lo_, sample_hi_);
    std::vector<double> q(SQMP_DOF, 0.0);
    if (sqmp_sample_informed(collisionChecker->context(), sample_seed_ + 0x5851f42d4c957f2dull, informed_next_, 1, sample_lo_, sample_hi_,
                             64, &el, q.data(), NULL) != SQMP_OK)
      return std::vector<double>();
    informed_next_ += 1;
    return q;
  }

  CollisionChecker *collisionChecker;   // same member name as birrt_star.h
  double m_near_threshold_interpolation;
  int m_max_near_nodes;

 private:
  struct Mirror {   // what the device tree of one host tree holds
    int slot = -1;
    std::vector<double> q, cost;
  };
  std::map<const Rrt_star_tree *, Mirror> mirrors_;
  int next_slot_ = 0;
  InformedEllipse ellipse_;
  int64_t informed_next_ = 0;
  std::vector<double> scratch_q_;
  std::vector<uint32_t> scratch_mask_;
  double sample_lo_[SQMP_DOF] = {0}, sample_hi_[SQMP_DOF] = {0};
  uint64_t sample_seed_ = 1;
  int64_t sample_next_ = 0;
  std::vector<double> sample_pool_;
  size_t sample_pos_ = 0;
};

}  // namespace sqmp_shim
