// capi.cu — the C ABI of libsqmp_b200.so (include/sqmp_b200.h): context, device memory, host<->device staging.
// No CPU fallback anywhere: every entry point either runs the CUDA kernels or returns an error code.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "sqmp_common.cuh"

namespace sqmp {

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct HostSphere {   // metres, in the frame of its rigid group
  double c[3], r;
  int link, grp;
};
struct HostCluster {  // <= 8 spheres of one link that lie close together; the unit of the bound hierarchies
  int grp, link;
  std::vector<int> members;   // sphere indices
  double c[3], R;             // enclosing sphere (metres, group frame)
  double c2[2], R2;           // enclosing vertical cylinder: axis (x, y), radius (yaw-only base group)
  double zlo, zhi;            // z extent of the members (centre -+ radius), group frame
  int slot0, nslots;          // its slots (4 or 8)
};

struct Ctx {
  int device = 0;
  int sm_count = 148;
  RobotDev rb;        // metres (FK kernel)
  RobotDev rb_grid;   // grid units of the current voxel grid (validity kernel); == rb while no grid is set
  GridDev gd;
  // ---- collision model (fixed at sqmp_create) ----
  std::vector<HostSphere> sph;
  std::vector<HostCluster> clusters;        // ordered by group
  std::vector<int> slot_sph;                // slot -> sphere index, -1 for a pad
  int grp_cluster_end[SQMP_MAX_GROUPS + 1];
  double grp_root_c[SQMP_MAX_GROUPS][3], grp_root_R[SQMP_MAX_GROUPS];   // enclosing sphere of a whole group
  double grp_root_c2[SQMP_MAX_GROUPS][2], grp_root_R2[SQMP_MAX_GROUPS];
  double grp_zlo[SQMP_MAX_GROUPS], grp_zhi[SQMP_MAX_GROUPS];
  int n_links = 0;
  uint64_t link_pair_mask[SQMP_MAX_LINKS];
  uint64_t link_disabled = 0;
  int base2d_group = -1;        // group that only translates in x, y and turns about z: tested against the slab fields
  double base2d_z0 = 0.0;       // world z (metres) of that group's frame origin
  // ---- device tables ----
  K1Slot *d_k1slots = nullptr;
  SelfGroupPair *d_self_gp = nullptr;
  int *d_self_lp = nullptr;
  SelfALink *d_self_al = nullptr;
  SelfLink *d_self_links = nullptr;
  SelfCluster *d_self_cl = nullptr;
  SelfLeaf *d_self_leaves = nullptr;
  size_t self_cap[6] = {0, 0, 0, 0, 0, 0};
  SelfDev self;
  bool grid_borrowed = false;   // the grid fields belong to another context (sqmp_share_voxel_grid): never freed here
  uint32_t *d_occ = nullptr;
  uint8_t *d_lb2 = nullptr;     // halo-padded (GridDev), followed by the slab fields
  uint16_t *d_ft = nullptr;
  std::vector<int> slab_k0, slab_k1;   // layer range of every slab field of the current grid
  FastDev fast;                 // tables of the fast validity kernel (validity_fast.cu)
  KfTab kftab;                  // its by-value node tables (host copy; thresholds follow the link-disable mask)
  std::vector<int> kf_leaf_thr, kf_bound_thr;   // thresholds with every link enabled (per slot / per cluster)
  std::vector<int> kf_bound_idx;                // cluster -> index in KfTab::bound
  int slab_of_link[SQMP_MAX_LINKS];             // link -> slab field (base2d group)
  int slab_root = -1;
  int ft_window = 0, slab_window = 0;           // search windows the current fields were built with (cells)
  FastAux *d_fast_aux = nullptr;
  float4 *d_fast_leaf = nullptr;
  double4 *d_fast_sph64 = nullptr;
  int fast_chain_ok = 0;        // description: translation-only base, then rotations only, spheres on the rotating part
  double reach_m = 0.0;         // max distance of a sphere centre from the base reference point (metres)
  int launch_overlap = 0;       // opt-in (include/sqmp_b200.h): K1F grids may start while the previous K1F grid of the stream drains (PDL)
  int64_t k1f_seq = -1;         // value of `launches` right after the last K1F launch, and the stream it went to:
  cudaStream_t k1f_stream = nullptr;   // overlap is requested only when nothing else was launched by this context since
  int k1_kernel = 0;            // 0 auto (fast when its preconditions hold) | 1 generic | 2 fast (error if not possible)
  TreeDev trees[SQMP_MAX_TREES];
  NNScratch nn;
  size_t nn_part_bytes = 0, nn_hits_bytes = 0;
  DevBuf scratch[8];  // grow-only staging buffers for the host-pointer entry points
  // low-latency path of the planner-shaped calls (one query, a few edges): mapped pinned host memory the kernels
  // read their inputs from and write their answers to — no cudaMemcpy on either side of the launch
  unsigned char *pin_h = nullptr, *pin_d = nullptr;
  unsigned char *pin2_h = nullptr, *pin2_d = nullptr;   // mapped pinned staging of the multi-tree NN calls (lazy)
  unsigned *d_ticket = nullptr;     // [0] ticket, [1] hit counter of k_nn_single
  bool tree_writes_pending = false; // unsynchronised single-node writes are in flight on streams[0]
  // single-node tree writes wait here (mapped pinned memory, two halves of TW_CAP records) until a call reads a tree
  unsigned char *tw_h = nullptr, *tw_d = nullptr;
  int tw_half = 0, tw_n = 0;
  cudaEvent_t tw_done[2] = {nullptr, nullptr};   // the flush that last read a half
  int64_t small_call_states = 4096; // validity calls up to this many states take the low-latency path
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int64_t launches = 0;
  int fk_precision = 64;  // 64 | 32 (FK chain arithmetic of sqmp_fk; the validity kernels always walk the chain in fp64)
  int fk_float_ok = 0;
  int chain_specialization = 1;  // 1: use the compile-time unrolled chain when the description matches a known signature
  int nn_variant = 0;     // 0 auto | 1 QT | 2 WQ | 3 QT + fp32 pre-filter | 4 tensor-core pre-filter
  int64_t nn_hit_pool_bytes = int64_t(1) << 30;
  int64_t pipeline_chunk = 1 << 18;  // states per H2D/kernel/D2H pipeline stage
  std::string err;
};

// layout of Ctx::pin_h
constexpr size_t PIN_RESULT = 0;            // int32[2 * 64 + 8]: NN answers
constexpr size_t PIN_MASK = 4096;           // uint32[...]: validity mask of a small call
constexpr size_t PIN_QA = 8192;             // doubles: states / edge start points of a small call
constexpr size_t PIN_BYTES = PIN_QA + 2 * 4096 * SQMP_DOF * sizeof(double);
constexpr size_t PIN_QB = PIN_QA + 4096 * SQMP_DOF * sizeof(double);

static int fail(Ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) {
      c->err += ": ";
      c->err += cudaGetErrorString(e);
    }
  }
  return code;
}
#define CU(call)                                                  \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return fail(c, SQMP_ERR_CUDA, #call, e__); \
  } while (0)

static int ensure(Ctx *c, int slot, size_t bytes, void **out) {
  DevBuf &b = c->scratch[slot];
  if (b.bytes < bytes) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&b.p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) return fail(c, SQMP_ERR_NOMEM, "cudaMalloc(scratch)", e);
    b.bytes = want;
  }
  *out = b.p;
  return SQMP_OK;
}

static int axis_code_of(const double *a) {
  const double eps = 0.0;
  if (a[1] == eps && a[2] == eps && a[0] == 1.0) return AX_PX;
  if (a[1] == eps && a[2] == eps && a[0] == -1.0) return AX_NX;
  if (a[0] == eps && a[2] == eps && a[1] == 1.0) return AX_PY;
  if (a[0] == eps && a[2] == eps && a[1] == -1.0) return AX_NY;
  if (a[0] == eps && a[1] == eps && a[2] == 1.0) return AX_PZ;
  if (a[0] == eps && a[1] == eps && a[2] == -1.0) return AX_NZ;
  return AX_GENERAL;
}
static int frame_kind_of(const double *f) {
  const bool rot_id = f[0] == 1.0 && f[4] == 1.0 && f[8] == 1.0 && f[1] == 0.0 && f[2] == 0.0 && f[3] == 0.0 &&
                      f[5] == 0.0 && f[6] == 0.0 && f[7] == 0.0;
  if (!rot_id) return 2;
  return (f[9] == 0.0 && f[10] == 0.0 && f[11] == 0.0) ? 0 : 1;
}

static void refresh_float_copies(RobotDev &r) {
  for (int s = 0; s < SQMP_MAX_SEG; ++s) {
    for (int i = 0; i < 12; ++i) r.seg_pre_f[s][i] = float(r.seg_pre[s][i]);
    for (int i = 0; i < 3; ++i) r.seg_axis_f[s][i] = float(r.seg_axis[s][i]);
    r.seg_axis_f[s][3] = 0.0f;
  }
  for (int i = 0; i < 12; ++i) r.ee_tip_f[i] = float(r.ee_tip[i]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Bound hierarchy of the collision model (once per sqmp_create, metres): the spheres of every (rigid group, link) are
// split by recursive median bisection into clusters of <= 8; a cluster gets 4 or 8 consecutive slots.  The same
// clusters serve the map check (cluster bound vs lower-bound field) and the self check (cluster bound vs cluster bound).
// ---------------------------------------------------------------------------------------------------------------------
static void enclose(const std::vector<HostSphere> &sph, const std::vector<int> &ids, double (&c)[3], double &R,
                    double (&c2)[2], double &R2, double &zlo, double &zhi) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i : ids)
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::fmin(lo[k], sph[size_t(i)].c[k] - sph[size_t(i)].r);
      hi[k] = std::fmax(hi[k], sph[size_t(i)].c[k] + sph[size_t(i)].r);
    }
  for (int k = 0; k < 3; ++k) c[k] = 0.5 * (lo[k] + hi[k]);   // centre of the bounding box of the spheres
  c2[0] = c[0];
  c2[1] = c[1];
  R = R2 = 0.0;
  for (int i : ids) {
    const HostSphere &s = sph[size_t(i)];
    const double dx = s.c[0] - c[0], dy = s.c[1] - c[1], dz = s.c[2] - c[2];
    R = std::fmax(R, std::sqrt(dx * dx + dy * dy + dz * dz) + s.r);
    R2 = std::fmax(R2, std::sqrt(dx * dx + dy * dy) + s.r);
  }
  zlo = lo[2];
  zhi = hi[2];
}

static void bisect(const std::vector<HostSphere> &sph, std::vector<int> ids, double zweight, std::vector<std::vector<int> > &out) {
  if (ids.size() <= 8) {
    out.push_back(ids);
    return;
  }
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i : ids)
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::fmin(lo[k], sph[size_t(i)].c[k]);
      hi[k] = std::fmax(hi[k], sph[size_t(i)].c[k]);
    }
  int ax = 0;
  double best = -1.0;
  for (int k = 0; k < 3; ++k) {
    const double ext = (hi[k] - lo[k]) * (k == 2 ? zweight : 1.0);
    if (ext > best) { best = ext; ax = k; }
  }
  std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return sph[size_t(a)].c[ax] < sph[size_t(b)].c[ax]; });
  // split so that both halves fill whole clusters as far as possible
  const size_t n = ids.size(), parts = (n + 7) / 8, left = (parts / 2) * 8 < n ? (parts / 2) * 8 : n / 2;
  std::vector<int> a(ids.begin(), ids.begin() + long(left)), b(ids.begin() + long(left), ids.end());
  bisect(sph, a, zweight, out);
  bisect(sph, b, zweight, out);
}

static int build_hierarchy(Ctx *c) {
  c->clusters.clear();
  c->slot_sph.clear();
  int slot = 0;
  for (int g = 0; g < SQMP_MAX_GROUPS; ++g) {
    c->grp_cluster_end[g] = int(c->clusters.size());
    c->rb.grp_slot_end[g] = slot;
    std::vector<int> all;
    for (int l = 0; l < SQMP_MAX_LINKS; ++l) {
      std::vector<int> ids;
      for (int i = 0; i < int(c->sph.size()); ++i)
        if (c->sph[size_t(i)].grp == g && c->sph[size_t(i)].link == l) ids.push_back(i);
      if (ids.empty()) continue;
      all.insert(all.end(), ids.begin(), ids.end());
      std::vector<std::vector<int> > parts;
      bisect(c->sph, ids, g == c->base2d_group ? 0.3 : 1.0, parts);
      for (auto &m : parts) {
        HostCluster cl;
        cl.grp = g;
        cl.link = l;
        cl.members = m;
        enclose(c->sph, m, cl.c, cl.R, cl.c2, cl.R2, cl.zlo, cl.zhi);
        cl.slot0 = slot;
        cl.nslots = m.size() <= 4 ? 4 : 8;
        for (int k = 0; k < cl.nslots; ++k) c->slot_sph.push_back(k < int(m.size()) ? m[size_t(k)] : -1);
        slot += cl.nslots;
        c->clusters.push_back(cl);
      }
    }
    c->grp_root_R[g] = c->grp_root_R2[g] = 0.0;
    if (!all.empty())
      enclose(c->sph, all, c->grp_root_c[g], c->grp_root_R[g], c->grp_root_c2[g], c->grp_root_R2[g], c->grp_zlo[g], c->grp_zhi[g]);
  }
  c->grp_cluster_end[SQMP_MAX_GROUPS] = int(c->clusters.size());
  c->rb.grp_slot_end[SQMP_MAX_GROUPS] = slot;
  c->rb.n_slots = slot;
  if (int(c->clusters.size()) > SQMP_MAX_CLUSTERS || slot > SQMP_MAX_SLOTS) return SQMP_ERR_INVALID;
  return SQMP_OK;
}

template <typename E>
static cudaError_t upload_vec(E *&dptr, size_t &cap, const std::vector<E> &v) {
  if (cap < v.size() + 1) {
    if (dptr) cudaFree(dptr);
    dptr = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&dptr, sizeof(E) * (v.size() + 8));
    if (e != cudaSuccess) return e;
    cap = v.size() + 8;
  }
  return v.empty() ? cudaSuccess : cudaMemcpy(dptr, v.data(), sizeof(E) * v.size(), cudaMemcpyHostToDevice);
}

// Self-collision tables (SelfDev), grid units `inv` = 1/res (1 while no grid is set).
static int build_self_tables(Ctx *c, double inv) {
  RobotDev &g = c->rb_grid;
  auto pad = [&](double r_sum) { return float((r_sum * (1.0 + 1e-4) + 1e-6) * inv); };   // bounds padded against fp32 rounding
  // links = (group, link) with spheres; clusters and leaves in slot order
  std::vector<SelfLeaf> leaves(c->slot_sph.size());
  for (size_t sl = 0; sl < c->slot_sph.size(); ++sl) {
    const int i = c->slot_sph[sl];
    if (i < 0) leaves[sl] = {0.f, 0.f, 0.f, -1.0e6f};   // pad: can never touch anything
    else leaves[sl] = {float(c->sph[size_t(i)].c[0] * inv), float(c->sph[size_t(i)].c[1] * inv), float(c->sph[size_t(i)].c[2] * inv),
                       float(c->sph[size_t(i)].r * inv)};
  }
  // bounding box of a sphere set (group frame, grid units), padded like the radii
  auto box_of = [&](const std::vector<int> &ids, float (&lo)[3], float (&hi)[3]) {
    for (int k = 0; k < 3; ++k) {
      double l = 1e300, h = -1e300;
      for (int i : ids) {
        l = std::fmin(l, c->sph[size_t(i)].c[k] - c->sph[size_t(i)].r);
        h = std::fmax(h, c->sph[size_t(i)].c[k] + c->sph[size_t(i)].r);
      }
      const double m = 1e-4 * std::fmax(std::fabs(l), std::fabs(h)) + 1e-6;
      lo[k] = float((l - m) * inv);
      hi[k] = float((h + m) * inv);
      lo[k] = std::nextafterf(lo[k], -INFINITY);
      hi[k] = std::nextafterf(hi[k], INFINITY);
    }
  };
  std::vector<SelfCluster> cls(c->clusters.size());
  std::vector<SelfLink> links;
  std::vector<int> link_grp;
  for (size_t k = 0; k < c->clusters.size(); ++k) {
    const HostCluster &cl = c->clusters[k];
    cls[k].x = float(cl.c[0] * inv); cls[k].y = float(cl.c[1] * inv); cls[k].z = float(cl.c[2] * inv); cls[k].R = pad(cl.R);
    box_of(cl.members, cls[k].lo, cls[k].hi);
    cls[k].leaf_begin = cl.slot0;
    cls[k].leaf_end = cl.slot0 + int(cl.members.size());
    if (links.empty() || links.back().link != cl.link || link_grp.back() != cl.grp) {
      SelfLink L = {0.f, 0.f, 0.f, 0.f, int(k), int(k), cl.link, 0};
      links.push_back(L);
      link_grp.push_back(cl.grp);
    }
    links.back().cl_end = int(k) + 1;
  }
  for (size_t li = 0; li < links.size(); ++li) {
    std::vector<int> ids;
    for (int k = links[li].cl_begin; k < links[li].cl_end; ++k)
      ids.insert(ids.end(), c->clusters[size_t(k)].members.begin(), c->clusters[size_t(k)].members.end());
    double cc[3], R, c2[2], R2, zl, zh;
    enclose(c->sph, ids, cc, R, c2, R2, zl, zh);
    links[li].x = float(cc[0] * inv); links[li].y = float(cc[1] * inv); links[li].z = float(cc[2] * inv);
    links[li].R = pad(R);
  }
  for (int gi = 0; gi < SQMP_MAX_GROUPS; ++gi) g.grp_fidx[gi] = -1;
  g.n_frames = 0;
  std::vector<SelfGroupPair> gps;
  std::vector<SelfALink> als;
  std::vector<int> lps;
  for (int ga = 0; ga < SQMP_MAX_GROUPS; ++ga)
    for (int gb = ga + 1; gb < SQMP_MAX_GROUPS; ++gb) {
      std::vector<int> ida, idb;
      const int al0 = int(als.size());
      for (size_t a = 0; a < links.size(); ++a) {
        if (link_grp[a] != ga) continue;
        const int lp0 = int(lps.size());
        for (size_t b = 0; b < links.size(); ++b)
          if (link_grp[b] == gb && ((c->link_pair_mask[links[a].link] >> links[b].link) & 1ull)) {
            lps.push_back(int(b));
            for (int k = links[b].cl_begin; k < links[b].cl_end; ++k)
              idb.insert(idb.end(), c->clusters[size_t(k)].members.begin(), c->clusters[size_t(k)].members.end());
          }
        if (int(lps.size()) == lp0) continue;
        for (int k = links[a].cl_begin; k < links[a].cl_end; ++k)
          ida.insert(ida.end(), c->clusters[size_t(k)].members.begin(), c->clusters[size_t(k)].members.end());
        SelfALink al;
        std::vector<int> own;
        for (int k = links[a].cl_begin; k < links[a].cl_end; ++k)
          own.insert(own.end(), c->clusters[size_t(k)].members.begin(), c->clusters[size_t(k)].members.end());
        box_of(own, al.lo, al.hi);
        al.lp_begin = lp0;
        al.lp_end = int(lps.size());
        al.a = int(a);
        al.pad[0] = al.pad[1] = al.pad[2] = 0;
        als.push_back(al);
      }
      if (int(als.size()) == al0) continue;
      if (g.grp_fidx[ga] < 0) g.grp_fidx[ga] = g.n_frames++;
      if (g.grp_fidx[gb] < 0) g.grp_fidx[gb] = g.n_frames++;
      double ca[3], RA, cb[3], RB, c2[2], R2, zl, zh;
      enclose(c->sph, ida, ca, RA, c2, R2, zl, zh);
      enclose(c->sph, idb, cb, RB, c2, R2, zl, zh);
      SelfGroupPair gp;
      gp.fa = g.grp_fidx[ga]; gp.fb = g.grp_fidx[gb];
      gp.al_begin = al0; gp.al_end = int(als.size());
      gp.ax = float(ca[0] * inv); gp.ay = float(ca[1] * inv); gp.az = float(ca[2] * inv); gp.RA = pad(RA);
      gp.bx = float(cb[0] * inv); gp.by = float(cb[1] * inv); gp.bz = float(cb[2] * inv); gp.RB = pad(RB);
      gps.push_back(gp);
    }
  cudaError_t e = upload_vec(c->d_self_gp, c->self_cap[0], gps);
  if (e == cudaSuccess) e = upload_vec(c->d_self_lp, c->self_cap[1], lps);
  if (e == cudaSuccess) e = upload_vec(c->d_self_al, c->self_cap[5], als);
  if (e == cudaSuccess) e = upload_vec(c->d_self_links, c->self_cap[2], links);
  if (e == cudaSuccess) e = upload_vec(c->d_self_cl, c->self_cap[3], cls);
  if (e == cudaSuccess) e = upload_vec(c->d_self_leaves, c->self_cap[4], leaves);
  if (e != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "self-collision tables", e);
  c->self.gpairs = c->d_self_gp; c->self.alinks = c->d_self_al; c->self.lpairs = c->d_self_lp; c->self.links = c->d_self_links;
  c->self.clusters = c->d_self_cl; c->self.leaves = c->d_self_leaves;
  c->self.n_gpairs = int(gps.size()); c->self.n_lpairs = int(lps.size()); c->self.n_links = int(links.size());
  c->rb.n_frames = g.n_frames;
  for (int gi = 0; gi < SQMP_MAX_GROUPS; ++gi) c->rb.grp_fidx[gi] = g.grp_fidx[gi];
  return SQMP_OK;
}

static inline void rec_set_thr(float4 *rec, int idx, int thr) {   // pair records: (x0 x1 y0 y1) (z0 z1 thr0 thr1)
  float4 &r = rec[2 * (idx >> 1) + 1];
  std::memcpy((idx & 1) ? &r.w : &r.z, &thr, 4);
}
static inline void rec_set_xyz(float4 *rec, int idx, float x, float y, float z) {
  float4 &a = rec[2 * (idx >> 1)], &b = rec[2 * (idx >> 1) + 1];
  if (idx & 1) { a.y = x; a.w = y; b.y = z; }
  else { a.x = x; a.z = y; b.x = z; }
}

// Thresholds of the fast kernel's node tables under the current link-disable mask (host only: the tables travel in the
// kernel parameters, nothing is copied or synchronised).  CollisionChecker::setDisabledLinkMapCollisions (CC:90-102).
static void apply_link_mask(Ctx *c) {
  c->rb.link_disabled = c->rb_grid.link_disabled = c->link_disabled;
  if (!c->fast.ok) return;
  KfTab &t = c->kftab;
  for (size_t k = 0; k < c->clusters.size(); ++k) {
    const HostCluster &cl = c->clusters[k];
    const bool off = (c->link_disabled >> cl.link) & 1ull;
    rec_set_thr(t.bound, c->kf_bound_idx[k], off ? -2 : c->kf_bound_thr[k]);   // a cluster belongs to one link: its spheres are only reached through it
  }
}

// Layer ranges of the slab fields for the current grid (yaw-only base group): one slab per link of that group (the
// layers its spheres can touch) + one for the whole group (root bound).  slab_of_link[l] / slab_root index c->slab_k0/k1.
static void plan_slabs(Ctx *c, double oz, double res, int gz) {
  c->slab_k0.clear();
  c->slab_k1.clear();
  for (int l = 0; l < SQMP_MAX_LINKS; ++l) c->slab_of_link[l] = -1;
  c->slab_root = -1;
  const int g = c->base2d_group;
  if (g < 0) return;
  auto range_of = [&](double zlo, double zhi, int &k0, int &k1) {
    k0 = int(std::floor((c->base2d_z0 + zlo - oz) / res - 1e-4));
    k1 = int(std::floor((c->base2d_z0 + zhi - oz) / res + 1e-4));
    if (k0 < 0) k0 = 0;
    if (k1 > gz - 1) k1 = gz - 1;
    if (k1 < k0) { k0 = 0; k1 = -1; }   // out of the grid: an empty slab (nothing to hit)
  };
  auto slab_for = [&](int k0, int k1) {
    for (size_t s = 0; s < c->slab_k0.size(); ++s)
      if (c->slab_k0[s] == k0 && c->slab_k1[s] == k1) return int(s);
    c->slab_k0.push_back(k0);
    c->slab_k1.push_back(k1);
    return int(c->slab_k0.size()) - 1;
  };
  int k0, k1;
  range_of(c->grp_zlo[g], c->grp_zhi[g], k0, k1);
  c->slab_root = slab_for(k0, k1);
  for (int l = 0; l < SQMP_MAX_LINKS; ++l) {
    double zlo = 1e300, zhi = -1e300;
    for (int k = c->grp_cluster_end[g]; k < c->grp_cluster_end[g + 1]; ++k)
      if (c->clusters[size_t(k)].link == l) {
        zlo = std::fmin(zlo, c->clusters[size_t(k)].zlo);
        zhi = std::fmax(zhi, c->clusters[size_t(k)].zhi);
      }
    if (zlo > zhi) continue;
    range_of(zlo, zhi, k0, k1);
    c->slab_of_link[l] = slab_for(k0, k1);
  }
}

// Grid-unit copy of the robot for the validity kernels: lengths / res, root offset measured from the grid origin, the
// slot tables with their distance-field thresholds, the node tables of the fast kernel, the self-collision tables.
// Re-derived whenever the grid changes (CC:76-88).
static int refresh_grid_robot(Ctx *c) {
  const RobotDev &rb = c->rb;
  RobotDev &g = c->rb_grid;
  g = rb;
  const bool has = c->gd.has_grid != 0;
  const double inv = has ? c->gd.inv_res : 1.0;
  const double o[3] = {has ? c->gd.ox : 0.0, has ? c->gd.oy : 0.0, has ? c->gd.oz : 0.0};
  g.len_scale = inv;
  for (int s = 0; s < rb.n_seg; ++s)
    for (int i = 9; i < 12; ++i) g.seg_pre[s][i] = rb.seg_pre[s][i] * inv;
  for (int i = 0; i < 3; ++i) g.root_off[i] = (rb.root_off[i] - o[i]) * inv;
  refresh_float_copies(g);
  refresh_float_copies(c->rb);
  const int n_slots = rb.n_slots;
  // ---- slot table of the generic kernel ----
  {
    std::vector<K1Slot> ks(size_t(n_slots) + 4);
    for (int sl = 0; sl < n_slots; ++sl) {
      int i = c->slot_sph[size_t(sl)];
      const bool pad_slot = i < 0;
      if (pad_slot) {   // a pad repeats the last real sphere of its cluster and never tests positive
        int t = sl;
        while (c->slot_sph[size_t(t)] < 0) --t;
        i = c->slot_sph[size_t(t)];
      }
      const HostSphere &s = c->sph[size_t(i)];
      K1Slot &k = ks[size_t(sl)];
      k.x = s.c[0] * inv; k.y = s.c[1] * inv; k.z = s.c[2] * inv; k.r = s.r * inv;
      const double rc = s.r * inv;   // radius in cells; sqmp_set_voxel_grid guarantees rc <= 13, so rc^2 < 255
      const double t = std::floor(rc * rc + 1e-9);
      k.thr = (!has || pad_slot) ? -2 : int(t > 254.0 ? 254.0 : t);
      const double n = std::ceil(rc);
      k.ncell = int(n > 250.0 ? 250.0 : n);
      k.link = pad_slot ? -1 : s.link;
      k.pad = 0;
    }
    if (!c->d_k1slots && cudaMalloc(&c->d_k1slots, sizeof(K1Slot) * (SQMP_MAX_SLOTS + 4)) != cudaSuccess)
      return fail(c, SQMP_ERR_NOMEM, "cudaMalloc(slot table)");
    if (n_slots > 0) CU(cudaMemcpy(c->d_k1slots, ks.data(), sizeof(K1Slot) * size_t(n_slots), cudaMemcpyHostToDevice));
  }
  // ---- tables of the fast validity kernel.  Its stage 1 runs on an fp32 copy of the node centre; err bounds the
  //      distance between that copy and the fp64 centre (cells): per axis <= 2^-24 * 13 (reach + 1) from rounding the
  //      frame, the sphere offset and three FMAs; doubled here, then sqrt(3) for the Euclidean norm.
  c->fast.ok = 0;
  int n_bounds = 0;
  for (int gi = 0; gi < SQMP_MAX_GROUPS; ++gi) n_bounds += (c->grp_cluster_end[gi + 1] - c->grp_cluster_end[gi] + 7) / 8 * 8;
  if (has && c->fast_chain_ok && c->gd.halo > 0 && n_slots <= KF_MAX_SLOTS && n_bounds <= KF_MAX_BOUNDS) {
    const double reach_c = c->reach_m * inv + 1.0;
    const double err = std::sqrt(3.0) * std::ldexp(26.0 * (reach_c + 1.0), -24) + 1e-6;
    const int W3 = c->ft_window, W2 = c->slab_window;
    auto thr_of = [&](double rg, int window) {   // field > thr proves a node of radius rg free; 255: cannot be proven (always open)
      const double t = std::floor((rg + err) * (rg + err) + 1e-9);
      return (t > 254.0 || t >= double(window) * window) ? 255 : int(t);
    };
    KfTab &t = c->kftab;
    std::memset(&t, 0, sizeof(t));
    t.base2d_group = c->base2d_group;
    c->kf_leaf_thr.assign(size_t(n_slots) + 2, -2);
    c->kf_bound_thr.assign(c->clusters.size(), -2);
    c->kf_bound_idx.assign(c->clusters.size(), 0);
    std::vector<FastAux> fa(size_t(n_slots) + 1);
    std::vector<double4> s64(size_t(n_slots) + 1);
    std::vector<float4> leaf(size_t(KF_MAX_SLOTS), make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = 0; i < KF_MAX_SLOTS; ++i) rec_set_thr(leaf.data(), i, -2);
    for (int i = 0; i < KF_MAX_BOUNDS; ++i) rec_set_thr(t.bound, i, -2);
    int bidx = 0;
    for (int gi = 0; gi < SQMP_MAX_GROUPS; ++gi) {
      t.grp_bound_end[gi] = short(bidx);
      const bool twod = gi == c->base2d_group;
      for (int k = c->grp_cluster_end[gi]; k < c->grp_cluster_end[gi + 1]; ++k, ++bidx) {
        const HostCluster &cl = c->clusters[size_t(k)];
        const float slabf = twod ? float(c->slab_of_link[cl.link]) : 0.0f;
        c->kf_bound_idx[size_t(k)] = bidx;
        if (twod) {
          rec_set_xyz(t.bound, bidx, float(cl.c2[0] * inv), float(cl.c2[1] * inv), slabf);
          c->kf_bound_thr[size_t(k)] = thr_of(cl.R2 * inv, W2);
        } else {
          rec_set_xyz(t.bound, bidx, float(cl.c[0] * inv), float(cl.c[1] * inv), float(cl.c[2] * inv));
          c->kf_bound_thr[size_t(k)] = thr_of(cl.R * inv, W3);
        }
        t.bound_first[bidx] = short(cl.slot0);
        t.bound_count[bidx] = (unsigned char)cl.nslots;
        for (int sl = cl.slot0; sl < cl.slot0 + cl.nslots; ++sl) {
          int i = c->slot_sph[size_t(sl)];
          const bool pad_slot = i < 0;
          if (pad_slot) i = cl.members.back();
          const HostSphere &sp = c->sph[size_t(i)];
          const double rg = sp.r * inv;
          rec_set_xyz(leaf.data(), sl, float(sp.c[0] * inv), float(sp.c[1] * inv), twod ? slabf : float(sp.c[2] * inv));
          c->kf_leaf_thr[size_t(sl)] = pad_slot ? -2 : thr_of(rg, twod ? W2 : W3);
          rec_set_thr(leaf.data(), sl, c->kf_leaf_thr[size_t(sl)]);
          FastAux &a = fa[size_t(sl)];
          a.x = float(sp.c[0] * inv); a.y = float(sp.c[1] * inv); a.z = float(sp.c[2] * inv);
          const double lo = rg > err ? (rg - err) * (rg - err) * (1.0 - 1e-6) : -1.0;
          float lof = float(lo);
          if (double(lof) > lo) lof = std::nextafterf(lof, -INFINITY);
          a.r2_lo = lof;
          a.ncell = int(std::ceil(rg)) + 1;
          a.pad[0] = a.pad[1] = a.pad[2] = 0;
          s64[size_t(sl)] = make_double4(sp.c[0] * inv, sp.c[1] * inv, sp.c[2] * inv, rg);
        }
      }
      while (bidx & 7) ++bidx;   // the kernel tests bounds eight at a time
      // root bound of the group
      int rthr = -2;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (c->grp_cluster_end[gi + 1] > c->grp_cluster_end[gi]) {
        if (twod) {
          rx = float(c->grp_root_c2[gi][0] * inv); ry = float(c->grp_root_c2[gi][1] * inv); rz = float(c->slab_root);
          rthr = thr_of(c->grp_root_R2[gi] * inv, W2);
        } else {
          rx = float(c->grp_root_c[gi][0] * inv); ry = float(c->grp_root_c[gi][1] * inv); rz = float(c->grp_root_c[gi][2] * inv);
          rthr = thr_of(c->grp_root_R[gi] * inv, W3);
        }
      }
      float rt;
      std::memcpy(&rt, &rthr, 4);
      t.root[gi] = make_float4(rx, ry, rz, rt);
    }
    t.grp_bound_end[SQMP_MAX_GROUPS] = short(bidx);
    cudaError_t e = cudaSuccess;
    if (!c->d_fast_aux) e = cudaMalloc(&c->d_fast_aux, sizeof(FastAux) * (KF_MAX_SLOTS + 1));
    if (e == cudaSuccess && !c->d_fast_leaf) e = cudaMalloc(&c->d_fast_leaf, sizeof(float4) * KF_MAX_SLOTS);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_fast_leaf, leaf.data(), sizeof(float4) * KF_MAX_SLOTS, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !c->d_fast_sph64) e = cudaMalloc(&c->d_fast_sph64, sizeof(double4) * (KF_MAX_SLOTS + 1));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_fast_aux, fa.data(), sizeof(FastAux) * size_t(n_slots), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_fast_sph64, s64.data(), sizeof(double4) * size_t(n_slots), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "fast validity tables", e);
    c->fast.leafrec = c->d_fast_leaf;
    c->fast.aux = c->d_fast_aux;
    c->fast.sph64 = c->d_fast_sph64;
    c->fast.n_slots = n_slots;
    c->fast.ok = 1;
  }
  apply_link_mask(c);
  return build_self_tables(c, inv);
}

// |node|^2 bound used by the fp32 pre-filter of K3/K4 (nn.cu); non-finite rows make it +inf (= no filtering)
static void tree_track_norm(TreeDev &t, const double *q, int64_t n) {
  double m = t.rmax2;
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int d = 0; d < SQMP_DOF; ++d) s += q[i * SQMP_DOF + d] * q[i * SQMP_DOF + d];
    if (!(s <= m)) m = (s == s) ? s : INFINITY;
  }
  t.rmax2 = m;
}

// ---- deferred single-node writes: sqmp_tree_append / sqmp_tree_update cost a record in pinned memory instead of a launch;
//      ONE launch on streams[0] writes the queue before anything reads a tree (every sqmp_nn_* entry, the batch writers,
//      reset, growth).  A second write to the same node replaces its record, so the queue never holds two for one target.
static const int TW_CAP = 256;
static int flush_tree_writes(Ctx *c) {
  if (c->tw_n == 0) return SQMP_OK;
  const size_t rec = tree_write_bytes();
  cudaError_t e = launch_tree_write_many(c->tw_d + size_t(c->tw_half) * TW_CAP * rec, c->tw_n, c->streams[0]);
  if (e == cudaSuccess) e = cudaEventRecord(c->tw_done[c->tw_half], c->streams[0]);
  if (e != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "tree write flush", e);
  c->tree_writes_pending = true;
  c->launches += 1;
  c->tw_half ^= 1;
  c->tw_n = 0;
  e = cudaEventSynchronize(c->tw_done[c->tw_half]);   // the half about to be refilled: its flush (two ago) has long finished
  if (e != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "cudaEventSynchronize(tree write half)", e);
  return SQMP_OK;
}
static int queue_tree_write(Ctx *c, const TreeDev &t, int64_t id, const double *q, double cost, bool may_be_queued) {
  const size_t rec = tree_write_bytes();
  if (!c->tw_h) {
    cudaError_t e = cudaHostAlloc((void **)&c->tw_h, 2 * TW_CAP * rec, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&c->tw_d, c->tw_h, 0);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->tw_done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return fail(c, SQMP_ERR_NOMEM, "cudaHostAlloc(tree write queue)", e);
  }
  unsigned char *base = c->tw_h + size_t(c->tw_half) * TW_CAP * rec;
  for (int i = may_be_queued ? c->tw_n - 1 : -1; i >= 0; --i)   // (an appended node is new: nothing to look for)
    if (tree_write_same_target(base + size_t(i) * rec, t, id)) {
      tree_write_fill(base + size_t(i) * rec, t, id, q, cost);
      return SQMP_OK;
    }
  if (c->tw_n == TW_CAP) {
    int rc = flush_tree_writes(c);
    if (rc) return rc;
    base = c->tw_h + size_t(c->tw_half) * TW_CAP * rec;
  }
  tree_write_fill(base + size_t(c->tw_n) * rec, t, id, q, cost);
  c->tw_n += 1;
  return SQMP_OK;
}

static int tree_reserve(Ctx *c, TreeDev &t, int64_t need) {
  if (need <= t.cap) return SQMP_OK;
  {   // queued single-node writes carry the old buffers' addresses
    int rcf = flush_tree_writes(c);
    if (rcf) return rcf;
  }
  // growing copies the old buffers with blocking copies on the legacy stream, which does not order against the
  // non-blocking streams[0]: single-node writes still in flight there (sqmp_tree_append / sqmp_tree_update) must land
  // in the OLD buffers before they are copied
  if (c->streams[0]) {
    cudaError_t es = cudaStreamSynchronize(c->streams[0]);
    if (es != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "cudaStreamSynchronize(tree grow)", es);
  }
  c->tree_writes_pending = false;
  int64_t cap = t.cap > 0 ? t.cap : 4096;
  while (cap < need) cap *= 2;
  // one allocation for the 8 coordinate columns + cost; columns 256-byte aligned
  const int64_t pitch = (cap + 31) / 32 * 32;
  double *blk = nullptr;
  cudaError_t e = cudaMalloc(&blk, size_t(pitch) * (SQMP_DOF + 1) * sizeof(double));
  if (e != cudaSuccess) return fail(c, SQMP_ERR_NOMEM, "cudaMalloc(tree)", e);
  for (int d = 0; d <= SQMP_DOF; ++d) {
    double *dst = blk + int64_t(d) * pitch;
    const double *src = d < SQMP_DOF ? t.q[d] : t.cost;
    if (t.n > 0) {
      e = cudaMemcpy(dst, src, size_t(t.n) * sizeof(double), cudaMemcpyDeviceToDevice);
      if (e != cudaSuccess) {
        cudaFree(blk);
        return fail(c, SQMP_ERR_CUDA, "cudaMemcpy(tree grow)", e);
      }
    }
  }
  // row-major copy + fp32 shadow (float4 + float4 + float per node)
  double *aos = nullptr;
  float4 *fa = nullptr, *fb = nullptr;
  float *fn = nullptr;
  float4 *img = nullptr;
  e = cudaMalloc(&aos, size_t(pitch) * SQMP_DOF * sizeof(double));
  if (e == cudaSuccess && t.n > 0) e = cudaMemcpy(aos, t.aos, size_t(t.n) * SQMP_DOF * sizeof(double), cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&fa, size_t(pitch) * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(&fb, size_t(pitch) * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(&fn, size_t(pitch) * sizeof(float));
  // (cap is a multiple of 128: whole tiles; rows past n are never trusted by the scan, zero keeps them finite)
  const size_t img_bytes = size_t((pitch + 127) / 128) * SQMP_NN_TILE_IMG_BYTES;
  if (e == cudaSuccess) e = cudaMalloc(&img, img_bytes);
  if (e == cudaSuccess) e = cudaMemset(img, 0, img_bytes);
  if (e == cudaSuccess && t.n > 0) {
    e = cudaMemcpy(fa, t.fa, size_t(t.n) * sizeof(float4), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(fb, t.fb, size_t(t.n) * sizeof(float4), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(fn, t.fn, size_t(t.n) * sizeof(float), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(img, t.img, size_t((t.n + 127) / 128) * SQMP_NN_TILE_IMG_BYTES, cudaMemcpyDeviceToDevice);
  }
  if (e != cudaSuccess) {
    cudaFree(blk);
    if (aos) cudaFree(aos);
    if (fa) cudaFree(fa);
    if (fb) cudaFree(fb);
    if (fn) cudaFree(fn);
    if (img) cudaFree(img);
    return fail(c, SQMP_ERR_NOMEM, "cudaMalloc(tree fp32 shadow)", e);
  }
  if (t.q[0]) cudaFree(t.q[0]);
  if (t.aos) cudaFree(t.aos);
  t.aos = aos;
  if (t.fa) cudaFree(t.fa);
  if (t.fb) cudaFree(t.fb);
  if (t.fn) cudaFree(t.fn);
  if (t.img) cudaFree(t.img);
  t.img = img;
  for (int d = 0; d < SQMP_DOF; ++d) t.q[d] = blk + int64_t(d) * pitch;
  t.cost = blk + int64_t(SQMP_DOF) * pitch;
  t.fa = fa; t.fb = fb; t.fn = fn;
  t.cap = cap;
  return SQMP_OK;
}

// smallest double s with sqrt(s) >= r: then  sqrt(x) < r  <=>  x < s   (sqrt is monotone and correctly rounded)
static double sumsq_threshold(double r) {
  if (!(r > 0.0)) return 0.0;
  double s = r * r;
  while (std::sqrt(s) >= r && s > 0.0) s = std::nextafter(s, 0.0);
  while (std::sqrt(s) < r) s = std::nextafter(s, INFINITY);
  return s;
}

static int pick_variant(const Ctx *c, const TreeDev &t, int64_t Q) {
  int v = c->nn_variant >= 1 && c->nn_variant <= 4 ? c->nn_variant : (Q >= 2048 ? 4 : 2);   // many queries: the tensor-core scan; few: a warp per query
  if (v == 4 && !(t.rmax2 < 6.0e4)) v = 3;   // its fp16 operands cannot hold this tree (tc_umma.cuh: TC_MAX_NORM2): fp32 pre-filter
  return v;
}
static int pick_chunk_hint(const Ctx *c, int variant, int64_t N, int64_t Q) {
  if (N <= 0) return 1;
  int64_t chunks;
  if (variant == 4) {
    // one 320-thread CTA per SM: the fewest chunks (each restarts the search for the best) that fill whole waves
    const int64_t per = nn_tc_queries_per_cta(), qtiles = (Q + per - 1) / per;
    const int64_t maxc = (N + 4095) / 4096;
    chunks = 1;
    double best_eff = 0.0;
    for (int64_t k = 1; k <= 16 && k <= maxc; ++k) {
      const int64_t ctas = qtiles * k, waves = (ctas + c->sm_count - 1) / c->sm_count;
      const double eff = double(ctas) / double(waves * c->sm_count);
      if (eff > best_eff + 0.02) { best_eff = eff; chunks = k; }
    }
    const int64_t minc = (N + (int64_t(1) << 21) - 1) >> 21;   // the kernel's ring entries hold node offsets below 2^22
    if (chunks < minc) chunks = minc;
  } else if (variant == 1 || variant == 3) {
    const int64_t per = variant == 3 ? 256 : 128;   // queries per CTA (nn.cu: NN_THREADS * NN_QPT)
    const int64_t qtiles = (Q + per - 1) / per;
    chunks = (int64_t(c->sm_count) * 16 + qtiles - 1) / qtiles;
    const int64_t maxc = (N + 2047) / 2048;
    if (chunks > maxc) chunks = maxc;
  } else {
    chunks = (int64_t(c->sm_count) * 32 + Q - 1) / Q;
    const int64_t maxc = (N + 255) / 256;
    if (chunks > maxc) chunks = maxc;
  }
  if (chunks < 1) chunks = 1;
  if (chunks > 4096) chunks = 4096;
  return int(chunks);
}
static int nn_prepare(Ctx *c, int64_t Q, int chunks_hint, bool radius, int64_t N) {
  const size_t part = size_t(chunks_hint + 1) * size_t(Q);
  if (c->nn_part_bytes < part * 12) {
    if (c->nn.part_d) cudaFree(c->nn.part_d);
    if (c->nn.part_id) cudaFree(c->nn.part_id);
    c->nn.part_d = nullptr;
    c->nn.part_id = nullptr;
    c->nn_part_bytes = 0;
    CU(cudaMalloc(&c->nn.part_d, part * sizeof(double)));
    CU(cudaMalloc(&c->nn.part_id, part * sizeof(int32_t)));
    c->nn_part_bytes = part * 12;
  }
  if (radius) {
    int64_t cap = c->nn_hit_pool_bytes / 4 / (Q > 0 ? Q : 1);
    if (cap > N) cap = N;
    if (cap < 32) cap = 32;
    const size_t bytes = size_t(cap) * size_t(Q) * sizeof(int32_t);
    if (c->nn_hits_bytes < bytes) {
      if (c->nn.hits) cudaFree(c->nn.hits);
      c->nn.hits = nullptr;
      c->nn_hits_bytes = 0;
      CU(cudaMalloc(&c->nn.hits, bytes));
      c->nn_hits_bytes = bytes;
    }
    c->nn.hit_cap = cap;
  }
  return SQMP_OK;
}

}  // namespace sqmp

using namespace sqmp;

struct sqmp_ctx : public sqmp::Ctx {};

#define CHECK_CTX()                        \
  if (!ctx) return SQMP_ERR_INVALID;       \
  Ctx *c = ctx;                            \
  {                                        \
    cudaError_t e0 = cudaSetDevice(c->device); \
    if (e0 != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "cudaSetDevice", e0); \
  }

extern "C" {

const char *sqmp_version(void) { return "libsqmp_b200 r2 (sm_100a)"; }

const char *sqmp_strerror(int code) {
  switch (code) {
    case SQMP_OK: return "ok";
    case SQMP_ERR_INVALID: return "invalid argument";
    case SQMP_ERR_CUDA: return "CUDA error";
    case SQMP_ERR_NOMEM: return "out of device memory";
    case SQMP_ERR_OVERFLOW: return "radius search hit buffer overflow";
    case SQMP_ERR_NODEVICE: return "no usable CUDA device (there is no CPU fallback)";
    default: return "unknown error";
  }
}

const char *sqmp_last_error(const sqmp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

sqmp_ctx *sqmp_create(int device, const sqmp_robot_desc *d, int *err_out) {
  auto bail = [&](int code) -> sqmp_ctx * {
    if (err_out) *err_out = code;
    return nullptr;
  };
  if (!d || d->n_seg < 1 || d->n_seg > SQMP_MAX_SEG || d->n_spheres < 0 || d->n_spheres > SQMP_MAX_SPHERES)
    return bail(SQMP_ERR_INVALID);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return bail(SQMP_ERR_NODEVICE);
  if (cudaSetDevice(device) != cudaSuccess) return bail(SQMP_ERR_NODEVICE);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(SQMP_ERR_NODEVICE);
  if (prop.major != 10) return bail(SQMP_ERR_NODEVICE);  // the library only contains sm_100a SASS

  sqmp_ctx *c = new (std::nothrow) sqmp_ctx();
  if (!c) return bail(SQMP_ERR_NOMEM);
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  std::memset(&c->rb, 0, sizeof(c->rb));
  std::memset(&c->gd, 0, sizeof(c->gd));
  std::memset(c->trees, 0, sizeof(c->trees));
  std::memset(&c->nn, 0, sizeof(c->nn));
  std::memset(&c->self, 0, sizeof(c->self));
  std::memset(&c->fast, 0, sizeof(c->fast));
  std::memset(&c->kftab, 0, sizeof(c->kftab));
  RobotDev &rb = c->rb;
  rb.n_seg = d->n_seg;
  rb.self_const_collision = d->self_const_collision;
  rb.ref_seg = -1;
  rb.trans_mask = 0;
  rb.len_scale = 1.0;
  bool leading = true;
  for (int s = 0; s < d->n_seg; ++s) {
    if ((d->seg_type[s] != SQMP_JOINT_ROT && d->seg_type[s] != SQMP_JOINT_TRANS) || d->seg_qidx[s] < 0 ||
        d->seg_qidx[s] >= SQMP_DOF) {
      delete c;
      return bail(SQMP_ERR_INVALID);
    }
    rb.seg_type[s] = d->seg_type[s];
    rb.seg_qidx[s] = d->seg_qidx[s];
    if (d->seg_type[s] == SQMP_JOINT_TRANS) rb.trans_mask |= 1u << d->seg_qidx[s];
    std::memcpy(rb.seg_pre[s], d->seg_pre[s], sizeof(double) * 12);
    std::memcpy(rb.seg_axis[s], d->seg_axis[s], sizeof(double) * 3);
    rb.seg_axis_code[s] = axis_code_of(d->seg_axis[s]);
    rb.seg_pre_kind[s] = frame_kind_of(d->seg_pre[s]);
    if (leading && d->seg_type[s] == SQMP_JOINT_TRANS && rb.seg_pre_kind[s] <= 1) rb.ref_seg = s;
    else leading = false;
  }
  std::memcpy(rb.root_off, d->root_offset, sizeof(double) * 3);
  std::memcpy(rb.ee_tip, d->ee_tip, sizeof(double) * 12);
  rb.ee_tip_kind = frame_kind_of(d->ee_tip);
  int prev = -1;
  c->fk_float_ok = 1;
  c->sph.resize(size_t(d->n_spheres));
  for (int i = 0; i < d->n_spheres; ++i) {
    const int sg = d->sph_seg[i];
    if (sg < prev || sg < -1 || sg >= d->n_seg || d->sph_link[i] < 0 || d->sph_link[i] >= SQMP_MAX_LINKS ||
        !(d->sph_radius[i] > 0.0)) {
      delete c;
      return bail(SQMP_ERR_INVALID);
    }
    prev = sg;
    HostSphere &h = c->sph[size_t(i)];
    h.c[0] = d->sph_center[3 * i + 0];
    h.c[1] = d->sph_center[3 * i + 1];
    h.c[2] = d->sph_center[3 * i + 2];
    h.r = d->sph_radius[i];
    h.link = d->sph_link[i];
    h.grp = sg + 1;
    if (h.link + 1 > c->n_links) c->n_links = h.link + 1;
    if (sg <= rb.ref_seg) c->fk_float_ok = 0;  // float chain starts after the prismatic base
  }
  for (int l = 0; l < SQMP_MAX_LINKS; ++l) {
    c->link_pair_mask[l] = d->link_pair_mask[l];
    c->slab_of_link[l] = -1;
  }
  for (int a = 0; a < SQMP_MAX_LINKS; ++a)   // symmetric, no self pairs
    for (int b = 0; b < SQMP_MAX_LINKS; ++b)
      if (a == b) c->link_pair_mask[a] &= ~(1ull << a);
      else if ((d->link_pair_mask[a] >> b) & 1ull) c->link_pair_mask[b] |= 1ull << a;
  // fast validity kernel: the base is a translation-only run, everything after it rotates (bounded reach), and every
  // sphere sits on the rotating part.  reach_m bounds |sphere centre - base reference point| over all configurations.
  c->fast_chain_ok = c->fk_float_ok && rb.ref_seg >= 0;
  {
    double cum = 0.0;
    c->reach_m = 0.0;
    for (int s = rb.ref_seg + 1; s < rb.n_seg; ++s) {
      if (rb.seg_type[s] != SQMP_JOINT_ROT) c->fast_chain_ok = 0;
      cum += std::sqrt(rb.seg_pre[s][9] * rb.seg_pre[s][9] + rb.seg_pre[s][10] * rb.seg_pre[s][10] +
                       rb.seg_pre[s][11] * rb.seg_pre[s][11]);
      for (const HostSphere &h : c->sph)
        if (h.grp == s + 1)
          c->reach_m = std::fmax(c->reach_m, cum + std::sqrt(h.c[0] * h.c[0] + h.c[1] * h.c[1] + h.c[2] * h.c[2]) + h.r);
    }
    if (!(c->reach_m < 1.0e6)) c->fast_chain_ok = 0;
  }
  // yaw-only base group: the leading prismatic run moves in x, y only, the first rotating segment turns about +-z with a
  // translation-only pre-frame -> the world z of that group's spheres never changes, so they can be tested against
  // 2-D slab fields that leave out the layers they cannot touch (in particular the floor right under the robot).
  c->base2d_group = -1;
  if (c->fast_chain_ok && rb.ref_seg + 1 < rb.n_seg) {
    const int s0 = rb.ref_seg + 1;
    bool ok2 = (rb.seg_axis_code[s0] == AX_PZ || rb.seg_axis_code[s0] == AX_NZ) && rb.seg_pre_kind[s0] <= 1;
    double z0 = rb.root_off[2];
    for (int s = 0; s <= s0; ++s) {
      z0 += rb.seg_pre[s][11];
      if (s < s0 && rb.seg_axis[s][2] != 0.0) ok2 = false;
    }
    if (ok2) {
      c->base2d_group = s0 + 1;
      c->base2d_z0 = z0;
    }
  }
  if (build_hierarchy(c) != SQMP_OK) {
    delete c;
    return bail(SQMP_ERR_INVALID);
  }
  bool ok = true;
  ok = ok && cudaHostAlloc((void **)&c->pin_h, PIN_BYTES, cudaHostAllocMapped) == cudaSuccess;
  ok = ok && cudaHostGetDevicePointer((void **)&c->pin_d, c->pin_h, 0) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_ticket, 64) == cudaSuccess && cudaMemset(c->d_ticket, 0, 64) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->streams[0], cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->streams[1], cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming) == cudaSuccess;
  if (!ok || refresh_grid_robot(c) != SQMP_OK) {
    sqmp_destroy(c);
    return bail(SQMP_ERR_CUDA);
  }
  if (err_out) *err_out = SQMP_OK;
  return c;
}

void sqmp_destroy(sqmp_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->d_self_gp) cudaFree(ctx->d_self_gp);
  if (ctx->d_self_lp) cudaFree(ctx->d_self_lp);
  if (ctx->d_self_al) cudaFree(ctx->d_self_al);
  if (ctx->d_self_links) cudaFree(ctx->d_self_links);
  if (ctx->d_self_cl) cudaFree(ctx->d_self_cl);
  if (ctx->d_self_leaves) cudaFree(ctx->d_self_leaves);
  if (ctx->d_k1slots) cudaFree(ctx->d_k1slots);
  if (!ctx->grid_borrowed) {
    if (ctx->d_occ) cudaFree(ctx->d_occ);
    if (ctx->d_lb2) cudaFree(ctx->d_lb2);
    if (ctx->d_ft) cudaFree(ctx->d_ft);
  }
  if (ctx->d_fast_aux) cudaFree(ctx->d_fast_aux);
  if (ctx->d_fast_leaf) cudaFree(ctx->d_fast_leaf);
  if (ctx->d_fast_sph64) cudaFree(ctx->d_fast_sph64);
  for (auto &t : ctx->trees) {
    if (t.q[0]) cudaFree(t.q[0]);
    if (t.aos) cudaFree(t.aos);
    if (t.fa) cudaFree(t.fa);
    if (t.fb) cudaFree(t.fb);
    if (t.fn) cudaFree(t.fn);
    if (t.img) cudaFree(t.img);
  }
  if (ctx->nn.part_d) cudaFree(ctx->nn.part_d);
  if (ctx->nn.part_id) cudaFree(ctx->nn.part_id);
  if (ctx->nn.hits) cudaFree(ctx->nn.hits);
  for (auto &b : ctx->scratch)
    if (b.p) cudaFree(b.p);
  if (ctx->pin_h) cudaFreeHost(ctx->pin_h);
  if (ctx->pin2_h) cudaFreeHost(ctx->pin2_h);
  if (ctx->tw_h) cudaFreeHost(ctx->tw_h);
  for (auto &e : ctx->tw_done)
    if (e) cudaEventDestroy(e);
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  for (auto &s : ctx->streams)
    if (s) cudaStreamDestroy(s);
  for (auto &e : ctx->ev)
    if (e) cudaEventDestroy(e);
  delete ctx;
}

int64_t sqmp_kernel_launches(const sqmp_ctx *ctx) { return ctx ? ctx->launches : 0; }

int sqmp_set_option(sqmp_ctx *ctx, const char *name, int64_t value) {
  if (!ctx || !name) return SQMP_ERR_INVALID;
  Ctx *c = ctx;
  const std::string n(name);
  if (n == "fk_precision") {
    if (value != 32 && value != 64) return fail(c, SQMP_ERR_INVALID, "fk_precision must be 32 or 64");
    if (value == 32 && !c->fk_float_ok) return fail(c, SQMP_ERR_INVALID, "float chain needs a translation-only base");
    c->fk_precision = int(value);
  } else if (n == "k1_kernel") {
    if (value < 0 || value > 2) return fail(c, SQMP_ERR_INVALID, "k1_kernel must be 0 (auto), 1 (generic) or 2 (fast)");
    if (value == 2 && !c->fast_chain_ok) return fail(c, SQMP_ERR_INVALID, "fast validity kernel needs a translation-only base followed by rotations");
    c->k1_kernel = int(value);
  } else if (n == "launch_overlap") {
    c->launch_overlap = value != 0;
  } else if (n == "chain_specialization") {
    c->chain_specialization = value != 0;
  } else if (n == "nn_variant") {
    if (value < 0 || value > 4) return fail(c, SQMP_ERR_INVALID, "nn_variant must be 0..4");
    c->nn_variant = int(value);
  } else if (n == "nn_hit_pool_bytes") {
    if (value < 4096) return fail(c, SQMP_ERR_INVALID, "nn_hit_pool_bytes too small");
    c->nn_hit_pool_bytes = value;
  } else if (n == "small_call_states") {
    if (value < 0 || value > 4096) return fail(c, SQMP_ERR_INVALID, "small_call_states must be 0..4096");
    c->small_call_states = value;
  } else if (n == "pipeline_chunk") {
    if (value < 1024) return fail(c, SQMP_ERR_INVALID, "pipeline_chunk too small");
    c->pipeline_chunk = (value + 31) / 32 * 32;
  } else {
    return fail(c, SQMP_ERR_INVALID, "unknown option");
  }
  return SQMP_OK;
}

// ---------------------------------------------------------------------------------------------------- map
int sqmp_set_voxel_grid(sqmp_ctx *ctx, const uint32_t *bits, int gx, int gy, int gz, const double origin[3],
                        double res) {
  CHECK_CTX();
  if (!bits) return SQMP_OK;  // CC:78-79
  if (gx <= 0 || gy <= 0 || gz <= 0 || !origin || !(res > 0.0)) return fail(c, SQMP_ERR_INVALID, "bad grid");
  const int64_t nvox = int64_t(gx) * gy * gz;
  if (nvox >= (int64_t(1) << 32)) return fail(c, SQMP_ERR_INVALID, "grid too large (>= 2^32 voxels)");
  double rmax = 0.0, bmax3 = 0.0, bmax2 = 0.0;
  for (const HostSphere &h : c->sph) rmax = std::fmax(rmax, h.r);
  for (const HostCluster &cl : c->clusters) {
    if (cl.grp == c->base2d_group) bmax2 = std::fmax(bmax2, cl.R2);
    else bmax3 = std::fmax(bmax3, cl.R);
  }
  for (int g = 0; g < SQMP_MAX_GROUPS; ++g) {
    if (g == c->base2d_group) bmax2 = std::fmax(bmax2, c->grp_root_R2[g]);
    else bmax3 = std::fmax(bmax3, c->grp_root_R[g]);
  }
  // the 16-bit feature field stores cell offsets in [-15, 15]: every occupied cell within reach of a sphere
  // (|offset| <= ceil(r/res) + 1 per axis) must be representable
  if (rmax / res > 13.0) return fail(c, SQMP_ERR_INVALID, "sphere radius too large for this resolution (> 13 voxels)");
  // search windows of the field builds: a field value is exact up to window^2 - 1 and "255" beyond it, so a node can
  // only be proven free when its threshold lies below window^2 (refresh_grid_robot marks the others "always open")
  int window = int(std::ceil(std::fmax(rmax, bmax3) / res)) + 1;
  if (window < 3) window = 3;
  if (window > 15) window = 15;
  int window2 = int(std::ceil(std::fmax(rmax, bmax2) / res)) + 1;
  if (window2 < 3) window2 = 3;
  if (window2 > 17) window2 = 17;
  // halo of the distance / feature fields: every cell a sphere centre can fall into while the base is inside the grid
  int halo = 0;
  if (c->fast_chain_ok) {
    const double h = std::ceil(c->reach_m / res) + 2.0;
    const double pn = (gx + 2.0 * h) * (gy + 2.0 * h) * (gz + 2.0 * h + 64.0);
    if (h <= 2048.0 && pn < 4294967296.0) halo = int(h);   // else: no halo, generic kernel only
  }
  plan_slabs(c, origin[2], res, gz);
  const int n_slabs = halo > 0 ? int(c->slab_k0.size()) : 0;
  const int64_t PX = gx + 2 * halo, PY = gy + 2 * halo, PZ = gz + 2 * halo;
  const int64_t pvox = PX * PY * PZ, pcol = PX * PY;
  const int row_words = (gx + 31) / 32 + 1;
  const size_t lin_words = size_t((nvox + 31) / 32) + 1;
  const size_t occ_words = size_t(gy) * gz * row_words;
  const size_t lb_bytes = size_t(pvox) + 16 + size_t(n_slabs) * size_t(pcol);   // lb2, 16 zero bytes, the slab fields
  CU(cudaDeviceSynchronize());
  uint32_t *d_lin = nullptr, *d_occ = nullptr, *d_col = nullptr;
  int8_t *d_a = nullptr;
  uint8_t *d_lb2 = nullptr;
  uint16_t *d_b = nullptr, *d_ft = nullptr;
  auto cleanup = [&]() {
    if (d_lin) cudaFree(d_lin);
    if (d_a) cudaFree(d_a);
    if (d_b) cudaFree(d_b);
    if (d_col) cudaFree(d_col);
  };
  cudaError_t e = cudaMalloc(&d_lin, lin_words * 4);
  if (e == cudaSuccess) e = cudaMemset(d_lin, 0, lin_words * 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_lin, bits, size_t((nvox + 31) / 32) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&d_occ, occ_words * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_a, size_t(pvox));
  if (e == cudaSuccess) e = cudaMalloc(&d_b, size_t(pvox) * 2);
  if (e == cudaSuccess) e = cudaMalloc(&d_lb2, lb_bytes);
  if (e == cudaSuccess) e = cudaMemset(d_lb2 + pvox, 0, 16);   // the zero byte at [pvox] (GridDev::zero_off)
  if (e == cudaSuccess) e = cudaMalloc(&d_ft, size_t(pvox) * 2);
  if (e == cudaSuccess && n_slabs > 0) e = cudaMalloc(&d_col, size_t(n_slabs) * gy * row_words * 4 + 16);
  if (e == cudaSuccess) e = launch_pack_rows(d_lin, gx, gy, gz, row_words, d_occ, nullptr);
  int nl = 1;
  if (e == cudaSuccess) e = launch_ft_build(d_occ, row_words, gx, gy, gz, halo, window, d_a, d_b, d_lb2, d_ft, nullptr, &nl);
  if (e == cudaSuccess && n_slabs > 0)   // d_a is free again after the feature transform: scratch of the slab build
    e = launch_slab_build(d_occ, row_words, gx, gy, gz, halo, window2, n_slabs, c->slab_k0.data(), c->slab_k1.data(), d_col, d_a,
                          d_lb2 + pvox + 16, nullptr, &nl);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cleanup();
  if (e != cudaSuccess) {
    if (d_occ) cudaFree(d_occ);
    if (d_lb2) cudaFree(d_lb2);
    if (d_ft) cudaFree(d_ft);
    return fail(c, e == cudaErrorMemoryAllocation ? SQMP_ERR_NOMEM : SQMP_ERR_CUDA, "set_voxel_grid", e);
  }
  c->launches += nl;
  if (!c->grid_borrowed) {
    if (c->d_occ) cudaFree(c->d_occ);  // previous collision object deleted (CC:84-85)
    if (c->d_lb2) cudaFree(c->d_lb2);
    if (c->d_ft) cudaFree(c->d_ft);
  }
  c->grid_borrowed = false;
  c->d_occ = d_occ;
  c->d_lb2 = d_lb2;
  c->d_ft = d_ft;
  c->ft_window = window;
  c->slab_window = window2;
  GridDev &g = c->gd;
  g.sx = unsigned(PX);
  g.splane = unsigned(PX * PY);
  g.org_off = unsigned((int64_t(halo) * PY + halo) * PX + halo);
  g.halo = halo;
  g.zero_off = unsigned(pvox);
  g.slab_off = unsigned(pvox + 16);
  g.splane2 = unsigned(pcol);
  g.n_slabs = n_slabs;
  g.occ = d_occ; g.lb2_base = d_lb2; g.ft_base = d_ft;
  g.lb2 = d_lb2 + g.org_off; g.ft = d_ft + g.org_off;
  g.gx = gx; g.gy = gy; g.gz = gz; g.row_words = row_words;
  g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
  g.res = res; g.inv_res = 1.0 / res;
  g.has_grid = 1;
  return refresh_grid_robot(c);
}

int sqmp_share_voxel_grid(sqmp_ctx *ctx, const sqmp_ctx *src) {
  CHECK_CTX();
  if (!src || src == ctx) return fail(c, SQMP_ERR_INVALID, "share_voxel_grid: bad source");
  const Ctx *o = src;
  if (o->device != c->device || !o->gd.has_grid) return fail(c, SQMP_ERR_INVALID, "share_voxel_grid: source has no grid on this device");
  if (o->reach_m != c->reach_m || o->fast_chain_ok != c->fast_chain_ok || o->base2d_group != c->base2d_group ||
      o->sph.size() != c->sph.size() || o->clusters.size() != c->clusters.size())
    return fail(c, SQMP_ERR_INVALID, "share_voxel_grid: contexts were created from different robot descriptions");
  CU(cudaDeviceSynchronize());
  if (!c->grid_borrowed) {
    if (c->d_occ) cudaFree(c->d_occ);
    if (c->d_lb2) cudaFree(c->d_lb2);
    if (c->d_ft) cudaFree(c->d_ft);
  }
  c->d_occ = o->d_occ;
  c->d_lb2 = o->d_lb2;
  c->d_ft = o->d_ft;
  c->grid_borrowed = true;
  c->gd = o->gd;
  c->ft_window = o->ft_window;
  c->slab_window = o->slab_window;
  plan_slabs(c, c->gd.oz, c->gd.res, c->gd.gz);   // same description, same grid -> the same slabs the source built
  if (c->slab_k0 != o->slab_k0 || c->slab_k1 != o->slab_k1)
    return fail(c, SQMP_ERR_INVALID, "share_voxel_grid: slab plans differ");
  return refresh_grid_robot(c);
}

// CollisionChecker::setDisabledLinkMapCollisions (CC:90-102): host-side only — the link mask and the thresholds it
// switches travel in the kernel parameters of the next launch; no device work, no synchronisation.
int sqmp_set_disabled_links(sqmp_ctx *ctx, uint64_t link_mask) {
  if (!ctx) return SQMP_ERR_INVALID;
  Ctx *c = ctx;
  c->link_disabled = link_mask;
  apply_link_mask(c);
  return SQMP_OK;
}

// ---------------------------------------------------------------------------------------------------- validity
// The fast kernel decides the same predicate as the generic fp64 kernel (fp32 only where a proven error bound makes it
// safe, fp64 for everything near contact); it needs a map check against a grid with halo-padded fields.
// Programmatic overlap is requested only for a K1F launch that directly follows another K1F launch of this context on
// the same stream: K1F reads its inputs before it waits for its predecessor, so the predecessor must not be their
// producer.  K1F never writes states, the library's own producers (the sampler, ...) break the chain here, and a
// caller who enqueues a foreign producer of the states between two validity calls on one stream must set
// launch_overlap = 0 (include/sqmp_b200.h).
static bool k1f_may_overlap(const Ctx *c, cudaStream_t st) {
  return c->launch_overlap != 0 && c->k1f_seq == c->launches && c->k1f_stream == st;
}
static void k1f_launched(Ctx *c, cudaStream_t st) {
  c->k1f_seq = c->launches;
  c->k1f_stream = st;
}
static bool use_fast_k1(const Ctx *c, unsigned flags) {
  if (c->k1_kernel == 1) return false;
  return c->fast.ok && (flags & SQMP_CHECK_MAP) && c->gd.has_grid;
}
int sqmp_check_states_dev(sqmp_ctx *ctx, const double *d_q, int64_t S, unsigned flags, uint32_t *d_mask,
                          void *stream) {
  CHECK_CTX();
  if (S < 0 || (S > 0 && (!d_q || !d_mask))) return fail(c, SQMP_ERR_INVALID, "check_states_dev: null pointer");
  if (S == 0) return SQMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & (SQMP_CHECK_SELF | SQMP_CHECK_MAP))) {  // CC:109-110
    CU(cudaMemsetAsync(d_mask, 0, size_t((S + 31) / 32) * 4, st));
    return SQMP_OK;
  }
  int nl = 0;
  const bool fast = use_fast_k1(c, flags);
  if (fast)
    CU(launch_check_states_fast(c->rb_grid, c->gd, c->self, c->fast, c->kftab, d_q, nullptr, 0, S, flags, d_mask,
                                c->chain_specialization != 0, k1f_may_overlap(c, st), st, &nl));
  else
    CU(launch_check_states(c->rb_grid, c->gd, c->self, c->d_k1slots, d_q, nullptr, 0, S, flags, d_mask, st, &nl));
  c->launches += nl;
  if (fast) k1f_launched(c, st);
  return SQMP_OK;
}

static int fill_peer_dev(Ctx *c, const sqmp_peer_gather *g, uint32_t step, PeerDev &pd) {
  if (!g || g->n_peers < 1 || g->n_peers > SQMP_MAX_PEERS || g->rank < 0 || g->rank >= g->n_peers || g->word_offset < 0)
    return fail(c, SQMP_ERR_INVALID, "peer gather: bad descriptor");
  std::memset(&pd, 0, sizeof(pd));
  for (int p = 0; p < g->n_peers; ++p) {
    if (!g->mask[p] || !g->flag[p]) return fail(c, SQMP_ERR_INVALID, "peer gather: null peer pointer");
    pd.mask[p] = g->mask[p];
    pd.flag[p] = g->flag[p];
  }
  pd.word_offset = g->word_offset;
  pd.n = g->n_peers;
  pd.rank = g->rank;
  pd.step = step;
  return SQMP_OK;
}

int sqmp_check_states_gather_dev(sqmp_ctx *ctx, const double *d_q, int64_t S, unsigned flags, const sqmp_peer_gather *g,
                                 void *stream) {
  CHECK_CTX();
  if (S <= 0 || !d_q) return fail(c, SQMP_ERR_INVALID, "check_states_gather_dev: bad argument");
  if (!use_fast_k1(c, flags)) return fail(c, SQMP_ERR_INVALID, "check_states_gather_dev: needs the fast validity kernel (map check)");
  PeerDev pd;
  int rc = fill_peer_dev(c, g, 0, pd);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int nl = 0;
  CU(launch_check_states_fast(c->rb_grid, c->gd, c->self, c->fast, c->kftab, d_q, nullptr, 0, S, flags, nullptr,
                              c->chain_specialization != 0, k1f_may_overlap(c, st), st, &nl, &pd));
  c->launches += nl;
  k1f_launched(c, st);
  return SQMP_OK;
}

int sqmp_gather_signal_dev(sqmp_ctx *ctx, const sqmp_peer_gather *g, uint32_t step, void *stream) {
  CHECK_CTX();
  PeerDev pd;
  int rc = fill_peer_dev(c, g, step, pd);
  if (rc) return rc;
  CU(launch_gather_signal(pd, (cudaStream_t)stream));
  // (like the wait kernel) not counted in `launches`: issued on the communication stream, outside the overlap chain
  return SQMP_OK;
}

int sqmp_gather_wait_dev(sqmp_ctx *ctx, const uint32_t *d_flags, int n_peers, uint32_t step, int32_t *d_status, void *stream) {
  CHECK_CTX();
  if (!d_flags || n_peers < 1 || n_peers > SQMP_MAX_PEERS) return fail(c, SQMP_ERR_INVALID, "gather_wait_dev: bad argument");
  CU(launch_gather_wait(d_flags, n_peers, step, d_status, (cudaStream_t)stream));
  // not counted in `launches`: it may run on another stream and must not break the overlap chain of the validity kernel
  return SQMP_OK;
}

int sqmp_check_edges_dev(sqmp_ctx *ctx, const double *d_a, const double *d_b, int64_t E, int num_points,
                         unsigned flags, uint32_t *d_mask, int32_t *d_first, void *stream) {
  CHECK_CTX();
  if (E < 0 || num_points < 1 || num_points > 255 || (E > 0 && (!d_a || !d_b || !d_mask)))
    return fail(c, SQMP_ERR_INVALID, "check_edges_dev: bad argument");
  if (E == 0) return SQMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int P1 = num_points + 1;
  const int64_t S = E * P1;
  if (!(flags & (SQMP_CHECK_SELF | SQMP_CHECK_MAP))) {
    CU(cudaMemsetAsync(d_mask, 0, size_t((S + 31) / 32) * 4, st));
    if (d_first) CU(cudaMemsetAsync(d_first, 0xFF, size_t(E) * 4, st));   // -1: every edge valid
    return SQMP_OK;
  }
  int nl = 0;
  const bool fast = use_fast_k1(c, flags);
  if (fast) {
    // the fast kernel lowers first_invalid[e] itself (atomicMin on the unsigned view): preset to 0xFFFFFFFF = -1
    if (d_first) CU(cudaMemsetAsync(d_first, 0xFF, size_t(E) * 4, st));
    CU(launch_check_states_fast(c->rb_grid, c->gd, c->self, c->fast, c->kftab, d_a, d_b, num_points, S, flags, d_mask,
                                c->chain_specialization != 0, !d_first && k1f_may_overlap(c, st), st, &nl, nullptr, d_first));
  } else {
    CU(launch_check_states(c->rb_grid, c->gd, c->self, c->d_k1slots, d_a, d_b, num_points, S, flags, d_mask, st, &nl));
    if (d_first) {
      CU(launch_first_invalid(d_mask, E, P1, d_first, st));
      nl += 1;
    }
  }
  c->launches += nl;
  if (fast && !d_first) k1f_launched(c, st);
  return SQMP_OK;
}

int sqmp_check_states(sqmp_ctx *ctx, const double *q, int64_t S, unsigned flags, uint32_t *mask_out) {
  CHECK_CTX();
  if (S < 0 || (S > 0 && (!q || !mask_out))) return fail(c, SQMP_ERR_INVALID, "check_states: null pointer");
  if (S == 0) return SQMP_OK;
  if (S <= c->small_call_states) {   // planner-shaped call: inputs and mask through mapped pinned memory, one launch
    std::memcpy(c->pin_h + PIN_QA, q, size_t(S) * SQMP_DOF * sizeof(double));
    int rc = sqmp_check_states_dev(ctx, (const double *)(c->pin_d + PIN_QA), S, flags, (uint32_t *)(c->pin_d + PIN_MASK),
                                   c->streams[0]);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->streams[0]));
    std::memcpy(mask_out, c->pin_h + PIN_MASK, size_t((S + 31) / 32) * 4);
    return SQMP_OK;
  }
  // H2D -> kernel -> D2H in stages over two streams so that copies overlap compute for large batches
  const int64_t chunk = c->pipeline_chunk;
  void *dq[2], *dm[2];
  const int64_t cs = S < chunk ? ((S + 31) / 32 * 32) : chunk;
  for (int b = 0; b < 2; ++b) {
    int rc = ensure(c, b, size_t(cs) * SQMP_DOF * sizeof(double), &dq[b]);
    if (rc) return rc;
    rc = ensure(c, 2 + b, size_t(cs / 32 + 1) * 4, &dm[b]);
    if (rc) return rc;
  }
  int b = 0;
  for (int64_t s0 = 0; s0 < S; s0 += cs, b ^= 1) {
    const int64_t n = S - s0 < cs ? S - s0 : cs;
    cudaStream_t st = c->streams[b];
    CU(cudaMemcpyAsync(dq[b], q + s0 * SQMP_DOF, size_t(n) * SQMP_DOF * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = sqmp_check_states_dev(ctx, (const double *)dq[b], n, flags, (uint32_t *)dm[b], st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(mask_out + s0 / 32, dm[b], size_t((n + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(c->streams[0]));
  CU(cudaStreamSynchronize(c->streams[1]));
  return SQMP_OK;
}

int sqmp_check_edges(sqmp_ctx *ctx, const double *qa, const double *qb, int64_t E, int num_points, unsigned flags,
                     uint32_t *mask_out, int32_t *first_invalid) {
  CHECK_CTX();
  if (E < 0 || num_points < 1 || num_points > 255 || (E > 0 && (!qa || !qb || !mask_out)))
    return fail(c, SQMP_ERR_INVALID, "check_edges: bad argument");
  if (E == 0) return SQMP_OK;
  const int64_t S = E * (num_points + 1);
  if (S <= c->small_call_states) {   // planner-shaped call (<= ~60 candidate edges of one iteration)
    std::memcpy(c->pin_h + PIN_QA, qa, size_t(E) * SQMP_DOF * sizeof(double));
    std::memcpy(c->pin_h + PIN_QB, qb, size_t(E) * SQMP_DOF * sizeof(double));
    int rc = sqmp_check_edges_dev(ctx, (const double *)(c->pin_d + PIN_QA), (const double *)(c->pin_d + PIN_QB), E, num_points,
                                  flags, (uint32_t *)(c->pin_d + PIN_MASK), nullptr, c->streams[0]);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->streams[0]));
    const uint32_t *m = (const uint32_t *)(c->pin_h + PIN_MASK);
    std::memcpy(mask_out, m, size_t((S + 31) / 32) * 4);
    if (first_invalid) {   // isEdgeValid's early exit (BS:6866-6874), from the mask
      const int P1 = num_points + 1;
      for (int64_t e = 0; e < E; ++e) {
        int32_t fi = -1;
        for (int kq = 0; kq < P1; ++kq) {
          const int64_t b = e * P1 + kq;
          if ((m[b >> 5] >> (b & 31)) & 1u) { fi = kq; break; }
        }
        first_invalid[e] = fi;
      }
    }
    return SQMP_OK;
  }
  // H2D -> kernel (+ first_invalid in its epilogue) -> D2H in stages over two streams: the copies of one stage overlap
  // the kernel of the other.  A stage holds a multiple of 32 edges, so its first state starts a mask word.
  const int P1 = num_points + 1;
  // (an edge costs 128 B over PCIe for its num_points + 1 states: the copies are short, so a stage holds twice the states of
  //  the state path's — fewer launches, each with its own ramp-up and tail; tools/e2e_edges_chunk.py)
  int64_t ce = (2 * c->pipeline_chunk / P1) / 32 * 32;
  if (ce < 32) ce = 32;
  if (ce > E) ce = (E + 31) / 32 * 32;
  void *da[2], *db[2], *dm[2], *df[2];
  for (int b = 0; b < 2; ++b) {
    int rc;
    if ((rc = ensure(c, b, size_t(ce) * 2 * SQMP_DOF * 8, &da[b]))) return rc;
    db[b] = (char *)da[b] + size_t(ce) * SQMP_DOF * 8;
    if ((rc = ensure(c, 2 + b, size_t(ce * P1 / 32 + 1) * 4, &dm[b]))) return rc;
    if ((rc = ensure(c, 4 + b, size_t(ce) * 4, &df[b]))) return rc;
  }
  int b = 0;
  for (int64_t e0 = 0; e0 < E; e0 += ce, b ^= 1) {
    const int64_t n = E - e0 < ce ? E - e0 : ce;
    cudaStream_t st = c->streams[b];
    CU(cudaMemcpyAsync(da[b], qa + e0 * SQMP_DOF, size_t(n) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(db[b], qb + e0 * SQMP_DOF, size_t(n) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
    int rc = sqmp_check_edges_dev(ctx, (const double *)da[b], (const double *)db[b], n, num_points, flags, (uint32_t *)dm[b],
                                  first_invalid ? (int32_t *)df[b] : nullptr, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(mask_out + e0 * P1 / 32, dm[b], size_t((n * P1 + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
    if (first_invalid) CU(cudaMemcpyAsync(first_invalid + e0, df[b], size_t(n) * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(c->streams[0]));
  CU(cudaStreamSynchronize(c->streams[1]));
  return SQMP_OK;
}

int sqmp_get_collisions(sqmp_ctx *ctx, const double q[SQMP_DOF], uint8_t *link_map_hit, uint8_t *link_pair_hit) {
  CHECK_CTX();
  if (!q || !link_map_hit) return fail(c, SQMP_ERR_INVALID, "get_collisions: null pointer");
  void *dq, *ds, *dp;
  int rc;
  const size_t LL = size_t(SQMP_MAX_LINKS) * SQMP_MAX_LINKS;
  if ((rc = ensure(c, 0, SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 2, SQMP_MAX_SLOTS, &ds))) return rc;
  if ((rc = ensure(c, 3, LL, &dp))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(dq, q, SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(ds, 0, SQMP_MAX_SLOTS, st));
  CU(cudaMemsetAsync(dp, 0, LL, st));
  CU(launch_get_collisions(c->rb_grid, c->gd, c->self, c->d_k1slots, (const double *)dq, (uint8_t *)ds, (uint8_t *)dp, st));
  c->launches += 1;
  std::vector<uint8_t> sh(SQMP_MAX_SLOTS);
  CU(cudaMemcpyAsync(sh.data(), ds, SQMP_MAX_SLOTS, cudaMemcpyDeviceToHost, st));
  if (link_pair_hit) CU(cudaMemcpyAsync(link_pair_hit, dp, LL, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::memset(link_map_hit, 0, SQMP_MAX_LINKS);
  for (int sl = 0; sl < c->rb_grid.n_slots; ++sl)
    if (sh[size_t(sl)] && c->slot_sph[size_t(sl)] >= 0) link_map_hit[c->sph[size_t(c->slot_sph[size_t(sl)])].link] = 1;
  return SQMP_OK;
}

// ---------------------------------------------------------------------------------------------------- FK
int sqmp_fk_dev(sqmp_ctx *ctx, const double *d_q, int64_t S, double *d_out, void *stream) {
  CHECK_CTX();
  if (S < 0 || (S > 0 && (!d_q || !d_out))) return fail(c, SQMP_ERR_INVALID, "fk_dev: null pointer");
  if (S == 0) return SQMP_OK;
  CU(launch_fk(c->rb, d_q, S, d_out, c->fk_precision, c->chain_specialization != 0, (cudaStream_t)stream));
  c->launches += 1;
  return SQMP_OK;
}

int sqmp_fk(sqmp_ctx *ctx, const double *q, int64_t S, double *out) {
  CHECK_CTX();
  if (S < 0 || (S > 0 && (!q || !out))) return fail(c, SQMP_ERR_INVALID, "fk: null pointer");
  if (S == 0) return SQMP_OK;
  void *dq, *dout;
  int rc;
  if ((rc = ensure(c, 0, size_t(S) * SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 4, size_t(S) * 7 * 8, &dout))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(dq, q, size_t(S) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  rc = sqmp_fk_dev(ctx, (const double *)dq, S, (double *)dout, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, dout, size_t(S) * 7 * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SQMP_OK;
}

// ---------------------------------------------------------------------------------------------------- sampling
int sqmp_sample_configs_dev(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                            const double hi[SQMP_DOF], int max_tries, double *d_q_out, int32_t *d_tries_out, void *stream) {
  CHECK_CTX();
  if (n < 0 || first < 0 || max_tries < 1 || max_tries > 65536 || !lo || !hi || (n > 0 && !d_q_out))
    return fail(c, SQMP_ERR_INVALID, "sample_configs: bad argument");
  for (int j = 0; j < SQMP_DOF; ++j)
    if (!(lo[j] <= hi[j]) || !std::isfinite(lo[j]) || !std::isfinite(hi[j]))
      return fail(c, SQMP_ERR_INVALID, "sample_configs: limits must be finite with lo <= hi");
  if (n == 0) return SQMP_OK;
  CU(launch_sample(c->rb, seed, first, n, lo, hi, max_tries, d_q_out, d_tries_out, c->chain_specialization != 0,
                   (cudaStream_t)stream));
  c->launches += 1;
  return SQMP_OK;
}

int sqmp_sample_configs(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                        const double hi[SQMP_DOF], int max_tries, double *q_out, int32_t *tries_out) {
  CHECK_CTX();
  if (n > 0 && !q_out) return fail(c, SQMP_ERR_INVALID, "sample_configs: null pointer");
  if (n <= 0) return n == 0 ? SQMP_OK : fail(c, SQMP_ERR_INVALID, "sample_configs: bad argument");
  void *dq, *dt;
  int rc;
  if ((rc = ensure(c, 0, size_t(n) * SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 5, size_t(n) * 4, &dt))) return rc;
  cudaStream_t st = c->streams[0];
  rc = sqmp_sample_configs_dev(ctx, seed, first, n, lo, hi, max_tries, (double *)dq, (int32_t *)dt, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(q_out, dq, size_t(n) * SQMP_DOF * 8, cudaMemcpyDeviceToHost, st));
  if (tries_out) CU(cudaMemcpyAsync(tries_out, dt, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SQMP_OK;
}

static bool informed_ok(const sqmp_informed *el) {
  if (!el || el->n_rev < 0 || el->n_pri < 0 || el->n_rev + el->n_pri != SQMP_DOF) return false;
  bool seen[SQMP_DOF] = {false};
  for (int cls = 0; cls < 2; ++cls) {
    const int n = cls ? el->n_rev : el->n_pri;
    const int32_t *idx = cls ? el->rev_idx : el->pri_idx;
    for (int a = 0; a < n; ++a) {
      if (idx[a] < 0 || idx[a] >= SQMP_DOF || seen[idx[a]]) return false;
      seen[idx[a]] = true;
    }
  }
  return true;
}
int sqmp_sample_informed_dev(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                             const double hi[SQMP_DOF], int max_tries, const sqmp_informed *ellipse, double *d_q_out,
                             int32_t *d_tries_out, void *stream) {
  CHECK_CTX();
  if (n < 0 || first < 0 || max_tries < 1 || max_tries > 65536 || !lo || !hi || (n > 0 && !d_q_out) || !informed_ok(ellipse))
    return fail(c, SQMP_ERR_INVALID, "sample_informed: bad argument");
  for (int j = 0; j < SQMP_DOF; ++j)
    if (!(lo[j] <= hi[j]) || !std::isfinite(lo[j]) || !std::isfinite(hi[j]))
      return fail(c, SQMP_ERR_INVALID, "sample_informed: limits must be finite with lo <= hi");
  if (n == 0) return SQMP_OK;
  CU(launch_sample_informed(c->rb, seed, first, n, lo, hi, max_tries, *ellipse, d_q_out, d_tries_out, c->chain_specialization != 0,
                            (cudaStream_t)stream));
  c->launches += 1;
  return SQMP_OK;
}
int sqmp_sample_informed(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                         const double hi[SQMP_DOF], int max_tries, const sqmp_informed *ellipse, double *q_out,
                         int32_t *tries_out) {
  CHECK_CTX();
  if (n > 0 && !q_out) return fail(c, SQMP_ERR_INVALID, "sample_informed: null pointer");
  if (n <= 0) return n == 0 ? SQMP_OK : fail(c, SQMP_ERR_INVALID, "sample_informed: bad argument");
  void *dq, *dt;
  int rc;
  if ((rc = ensure(c, 0, size_t(n) * SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 5, size_t(n) * 4, &dt))) return rc;
  cudaStream_t st = c->streams[0];
  rc = sqmp_sample_informed_dev(ctx, seed, first, n, lo, hi, max_tries, ellipse, (double *)dq, (int32_t *)dt, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(q_out, dq, size_t(n) * SQMP_DOF * 8, cudaMemcpyDeviceToHost, st));
  if (tries_out) CU(cudaMemcpyAsync(tries_out, dt, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SQMP_OK;
}

// ---------------------------------------------------------------------------------------------------- tree
#define CHECK_TREE()                                                                 \
  if (tree < 0 || tree >= SQMP_MAX_TREES) return fail(c, SQMP_ERR_INVALID, "bad tree index"); \
  TreeDev &t = c->trees[tree];

#define FLUSH_TREE_WRITES()                  \
  {                                          \
    int rc_flush_ = flush_tree_writes(c);    \
    if (rc_flush_) return rc_flush_;         \
  }

int sqmp_tree_reset(sqmp_ctx *ctx, int tree) {
  CHECK_CTX();
  CHECK_TREE();
  FLUSH_TREE_WRITES();
  t.n = 0;
  t.rmax2 = 0.0;
  return SQMP_OK;
}
int sqmp_tree_size(sqmp_ctx *ctx, int tree, int64_t *n_out) {
  CHECK_CTX();
  CHECK_TREE();
  if (n_out) *n_out = t.n;
  return SQMP_OK;
}
int sqmp_tree_append(sqmp_ctx *ctx, int tree, const double q[SQMP_DOF], double cost, int32_t *id_out) {
  CHECK_CTX();
  CHECK_TREE();
  if (!q) return fail(c, SQMP_ERR_INVALID, "tree_append: null q");
  if (t.n >= 0x7fffffff) return fail(c, SQMP_ERR_INVALID, "tree full");
  // deferred: the node waits in the write queue (queue_tree_write) until a call reads a tree; the flush is stream-ordered
  // on streams[0] and every reader runs on (or synchronises with) that stream.  Growing the buffers flushes first.
  int rc = tree_reserve(c, t, t.n + 1);   // (synchronises streams[0] itself when it has to grow)
  if (rc) return rc;
  tree_track_norm(t, q, 1);
  rc = queue_tree_write(c, t, t.n, q, cost, false);
  if (rc) return rc;
  if (id_out) *id_out = int32_t(t.n);
  t.n += 1;
  return SQMP_OK;
}
int sqmp_tree_append_batch(sqmp_ctx *ctx, int tree, const double *q, const double *cost, int64_t n,
                           int32_t *first_id_out) {
  CHECK_CTX();
  CHECK_TREE();
  if (n < 0 || (n > 0 && (!q || !cost))) return fail(c, SQMP_ERR_INVALID, "tree_append_batch: null pointer");
  if (t.n + n > 0x7fffffff) return fail(c, SQMP_ERR_INVALID, "tree full");
  if (first_id_out) *first_id_out = int32_t(t.n);
  if (n == 0) return SQMP_OK;
  FLUSH_TREE_WRITES();
  int rc = tree_reserve(c, t, t.n + n);
  if (rc) return rc;
  void *dq, *dc;
  if ((rc = ensure(c, 0, size_t(n) * SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 1, size_t(n) * 8, &dc))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(dq, q, size_t(n) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dc, cost, size_t(n) * 8, cudaMemcpyHostToDevice, st));
  tree_track_norm(t, q, n);
  CU(launch_tree_append(t, (const double *)dq, (const double *)dc, t.n, n, st));
  CU(cudaStreamSynchronize(st));
  c->launches += 1;
  t.n += n;
  return SQMP_OK;
}
int sqmp_tree_update(sqmp_ctx *ctx, int tree, int32_t id, const double q[SQMP_DOF], double cost) {
  CHECK_CTX();
  CHECK_TREE();
  if (!q || id < 0 || id >= t.n) return fail(c, SQMP_ERR_INVALID, "tree_update: bad id");
  tree_track_norm(t, q, 1);
  return queue_tree_write(c, t, id, q, cost, true);   // deferred (see sqmp_tree_append)
}
int sqmp_tree_update_costs(sqmp_ctx *ctx, int tree, const int32_t *ids, const double *cost, int64_t n) {
  CHECK_CTX();
  CHECK_TREE();
  if (n < 0 || (n > 0 && (!ids || !cost))) return fail(c, SQMP_ERR_INVALID, "tree_update_costs: null pointer");
  if (n == 0) return SQMP_OK;
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= t.n) return fail(c, SQMP_ERR_INVALID, "tree_update_costs: bad id");
  FLUSH_TREE_WRITES();
  void *di, *dc;
  int rc;
  if ((rc = ensure(c, 0, size_t(n) * 4, &di))) return rc;
  if ((rc = ensure(c, 1, size_t(n) * 8, &dc))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(di, ids, size_t(n) * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dc, cost, size_t(n) * 8, cudaMemcpyHostToDevice, st));
  CU(launch_tree_update_costs(t, (const int32_t *)di, (const double *)dc, n, st));
  CU(cudaStreamSynchronize(st));
  c->launches += 1;
  return SQMP_OK;
}

// ---------------------------------------------------------------------------------------------------- NN
int sqmp_nn_nearest_dev(sqmp_ctx *ctx, int tree, const double *d_queries, int64_t Q, int32_t *d_id, void *stream) {
  CHECK_CTX();
  FLUSH_TREE_WRITES();
  CHECK_TREE();
  if (Q < 0 || (Q > 0 && (!d_queries || !d_id))) return fail(c, SQMP_ERR_INVALID, "nn_nearest_dev: null pointer");
  if (Q == 0) return SQMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->tree_writes_pending && st != c->streams[0]) {   // single-node writes are ordered on streams[0] only
    CU(cudaStreamSynchronize(c->streams[0]));
    c->tree_writes_pending = false;
  }
  if (t.n == 0) {  // BS:4084: id stays 0 (the reference would then index an empty vector; we report 0)
    CU(cudaMemsetAsync(d_id, 0, size_t(Q) * 4, st));
    return SQMP_OK;
  }
  const int variant = pick_variant(c, t, Q);
  const int hint = pick_chunk_hint(c, variant, t.n, Q);
  int rc = nn_prepare(c, Q, hint, false, t.n);
  if (rc) return rc;
  int nl = 0;
  CU(launch_nn_nearest(t, d_queries, Q, d_id, c->nn, hint, variant, st, &nl));
  c->launches += nl;
  return SQMP_OK;
}

int sqmp_nn_radius_dev(sqmp_ctx *ctx, int tree, const double *d_queries, const int32_t *d_exclude, int64_t Q,
                       double r, int k, int32_t *d_low, int32_t *d_high, int32_t *d_count, int32_t *d_overflow,
                       void *stream) {
  CHECK_CTX();
  FLUSH_TREE_WRITES();
  CHECK_TREE();
  if (Q < 0 || k < 1 || k > 1024 || (Q > 0 && (!d_queries || !d_low || !d_high || !d_count || !d_overflow)))
    return fail(c, SQMP_ERR_INVALID, "nn_radius_dev: bad argument");
  if (Q == 0) return SQMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->tree_writes_pending && st != c->streams[0]) {
    CU(cudaStreamSynchronize(c->streams[0]));
    c->tree_writes_pending = false;
  }
  CU(cudaMemsetAsync(d_overflow, 0, 4, st));
  if (t.n == 0) {
    CU(cudaMemsetAsync(d_count, 0, size_t(Q) * 4, st));
    CU(cudaMemsetAsync(d_low, 0xff, size_t(Q) * k * 4, st));
    CU(cudaMemsetAsync(d_high, 0xff, size_t(Q) * k * 4, st));
    return SQMP_OK;
  }
  const int variant = pick_variant(c, t, Q);
  const int hint = pick_chunk_hint(c, variant, t.n, Q);
  int rc = nn_prepare(c, Q, hint, true, t.n);
  if (rc) return rc;
  int nl = 0;
  CU(launch_nn_radius(t, d_queries, d_exclude, Q, sumsq_threshold(r), r, k, d_low, d_high, d_count, d_overflow, c->nn,
                      hint, variant, st, &nl));
  c->launches += nl;
  return SQMP_OK;
}

int sqmp_nn_nearest(sqmp_ctx *ctx, int tree, const double *queries, int64_t Q, int32_t *id_out) {
  CHECK_CTX();
  FLUSH_TREE_WRITES();
  if (Q < 0 || (Q > 0 && (!queries || !id_out))) return fail(c, SQMP_ERR_INVALID, "nn_nearest: null pointer");
  if (Q == 0) return SQMP_OK;
  if (Q == 1 && c->nn_variant == 0) {   // the planner's call shape: one launch, answer in mapped host memory
    CHECK_TREE();
    if (t.n == 0) { id_out[0] = 0; return SQMP_OK; }
    int rc1 = nn_prepare(c, 1, c->sm_count, false, t.n);
    if (rc1) return rc1;
    CU(launch_nn_single(t, queries, false, -1, 0.0, 0, c->nn, c->d_ticket, (int32_t *)(c->d_ticket + 1),
                        (int32_t *)(c->pin_d + PIN_RESULT), c->sm_count, c->streams[0]));
    c->launches += 1;
    CU(cudaStreamSynchronize(c->streams[0]));
    id_out[0] = ((volatile int32_t *)(c->pin_h + PIN_RESULT))[0];
    return SQMP_OK;
  }
  void *dq, *di;
  int rc;
  if ((rc = ensure(c, 0, size_t(Q) * SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 2, size_t(Q) * 4, &di))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(dq, queries, size_t(Q) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  rc = sqmp_nn_nearest_dev(ctx, tree, (const double *)dq, Q, (int32_t *)di, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(id_out, di, size_t(Q) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SQMP_OK;
}

static int radius_host_range(sqmp_ctx *ctx, int tree, const double *queries, const int32_t *exclude, int64_t q0,
                             int64_t nq, double r, int k, int32_t *low, int32_t *high, int32_t *count, int *overflow) {
  Ctx *c = ctx;
  void *dq, *de, *dl, *dh, *dc, *dv;
  int rc;
  if ((rc = ensure(c, 0, size_t(nq) * SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 1, size_t(nq) * 4, &de))) return rc;
  if ((rc = ensure(c, 2, size_t(nq) * k * 4, &dl))) return rc;
  if ((rc = ensure(c, 3, size_t(nq) * k * 4, &dh))) return rc;
  if ((rc = ensure(c, 4, size_t(nq) * 4, &dc))) return rc;
  if ((rc = ensure(c, 5, 64, &dv))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(dq, queries + q0 * SQMP_DOF, size_t(nq) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  if (exclude) CU(cudaMemcpyAsync(de, exclude + q0, size_t(nq) * 4, cudaMemcpyHostToDevice, st));
  rc = sqmp_nn_radius_dev(ctx, tree, (const double *)dq, exclude ? (const int32_t *)de : nullptr, nq, r, k,
                          (int32_t *)dl, (int32_t *)dh, (int32_t *)dc, (int32_t *)dv, st);
  if (rc) return rc;
  int32_t ov = 0;
  CU(cudaMemcpyAsync(&ov, dv, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *overflow = ov;
  if (ov) return SQMP_OK;
  CU(cudaMemcpyAsync(low + q0 * k, dl, size_t(nq) * k * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(high + q0 * k, dh, size_t(nq) * k * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(count + q0, dc, size_t(nq) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SQMP_OK;
}

int sqmp_nn_radius(sqmp_ctx *ctx, int tree, const double *queries, const int32_t *exclude_id, int64_t Q, double r,
                   int k, int32_t *low_ids, int32_t *high_ids, int32_t *count_out) {
  CHECK_CTX();
  FLUSH_TREE_WRITES();
  if (Q < 0 || k < 1 || k > 1024 || (Q > 0 && (!queries || !low_ids || !high_ids || !count_out)))
    return fail(c, SQMP_ERR_INVALID, "nn_radius: bad argument");
  if (Q == 0) return SQMP_OK;
  if (Q == 1 && k <= 64 && c->nn_variant == 0) {   // the planner's call shape (k_nn_single)
    CHECK_TREE();
    if (t.n == 0) {
      count_out[0] = 0;
      for (int i = 0; i < k; ++i) low_ids[i] = high_ids[i] = -1;
      return SQMP_OK;
    }
    int rc1 = nn_prepare(c, 1, c->sm_count, true, t.n);   // hit buffer = whole tree at Q = 1
    if (rc1) return rc1;
    if (c->nn.hit_cap < t.n) return fail(c, SQMP_ERR_OVERFLOW, "nn_radius: hit pool smaller than the tree");
    CU(launch_nn_single(t, queries, true, exclude_id ? exclude_id[0] : -1, sumsq_threshold(r), k, c->nn, c->d_ticket,
                        (int32_t *)(c->d_ticket + 1), (int32_t *)(c->pin_d + PIN_RESULT), c->sm_count, c->streams[0]));
    c->launches += 1;
    CU(cudaStreamSynchronize(c->streams[0]));
    const volatile int32_t *res = (volatile int32_t *)(c->pin_h + PIN_RESULT);
    count_out[0] = res[0];
    for (int i = 0; i < k; ++i) { low_ids[i] = res[1 + i]; high_ids[i] = res[1 + k + i]; }
    return SQMP_OK;
  }
  // The per-query hit buffer shares a fixed pool; if a batch overflows it, halve the batch (the buffer per
  // query doubles) until it fits — at batch size 1 the buffer holds the whole tree.
  int64_t batch = Q;
  for (int64_t q0 = 0; q0 < Q;) {
    const int64_t nq = Q - q0 < batch ? Q - q0 : batch;
    int ov = 0;
    int rc = radius_host_range(ctx, tree, queries, exclude_id, q0, nq, r, k, low_ids, high_ids, count_out, &ov);
    if (rc) return rc;
    if (ov) {
      if (nq == 1) return fail(c, SQMP_ERR_OVERFLOW, "nn_radius: hit pool smaller than the tree");
      batch = (nq + 1) / 2;
      continue;
    }
    q0 += nq;
  }
  return SQMP_OK;
}

}  // extern "C"

// ---- every query against its own tree (k_nn_multi) -------------------------------------------------------------
constexpr int64_t MULTI_MAX_Q = 1024;   // queries per launch; larger calls are split
constexpr int MULTI_MAX_K = 64;
static int nn_multi_range(Ctx *c, const int32_t *tree, const double *queries, const int32_t *exclude, int64_t nq,
                          bool radius, double r, int k, int32_t *out_ids, int32_t *low, int32_t *high, int32_t *count) {
  const size_t desc_b = nn_multi_desc_bytes();
  const size_t off_q = size_t(MULTI_MAX_Q) * desc_b, off_out = off_q + size_t(MULTI_MAX_Q) * SQMP_DOF * 8;
  const size_t total = off_out + size_t(MULTI_MAX_Q) * (2 * MULTI_MAX_K + 2) * 4;
  if (!c->pin2_h) {
    CU(cudaHostAlloc((void **)&c->pin2_h, total, cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void **)&c->pin2_d, c->pin2_h, 0));
  }
  int64_t nmax = 1;
  for (int64_t i = 0; i < nq; ++i) {
    if (tree[i] < 0 || tree[i] >= SQMP_MAX_TREES) return fail(c, SQMP_ERR_INVALID, "nn_multi: bad tree index");
    const TreeDev &t = c->trees[tree[i]];
    nn_multi_fill_desc(c->pin2_h + size_t(i) * desc_b, t, (radius && exclude) ? exclude[i] : -1);
    if (t.n > nmax) nmax = t.n;
  }
  std::memcpy(c->pin2_h + off_q, queries, size_t(nq) * SQMP_DOF * 8);
  if (radius) {
    int rc = nn_prepare(c, nq, 1, true, nmax);
    if (rc) return rc;
  }
  int32_t *d_out = (int32_t *)(c->pin2_d + off_out);
  CU(launch_nn_multi(c->pin2_d, (const double *)(c->pin2_d + off_q), nq, radius, radius ? sumsq_threshold(r) : 0.0, k,
                     c->nn, d_out, c->streams[0]));
  c->launches += 1;
  CU(cudaStreamSynchronize(c->streams[0]));
  const volatile int32_t *res = (volatile int32_t *)(c->pin2_h + off_out);
  if (!radius) {
    for (int64_t i = 0; i < nq; ++i) out_ids[i] = res[i];
    return SQMP_OK;
  }
  for (int64_t i = 0; i < nq; ++i) {
    const volatile int32_t *o = res + i * (2 * k + 2);
    if (o[1]) {   // this query's hits overflowed its slice of the pool: the single-tree path holds the whole tree
      int rc = sqmp_nn_radius(static_cast<sqmp_ctx *>(c), tree[i], queries + i * SQMP_DOF, exclude ? exclude + i : nullptr, 1,
                              r, k, low + i * k, high + i * k, count + i);
      if (rc) return rc;
      continue;
    }
    count[i] = o[0];
    for (int j = 0; j < k; ++j) {
      low[i * k + j] = o[2 + j];
      high[i * k + j] = o[2 + k + j];
    }
  }
  return SQMP_OK;
}

int sqmp_nn_nearest_multi(sqmp_ctx *ctx, const int32_t *tree, const double *queries, int64_t Q, int32_t *id_out) {
  CHECK_CTX();
  FLUSH_TREE_WRITES();
  if (Q < 0 || (Q > 0 && (!tree || !queries || !id_out))) return fail(c, SQMP_ERR_INVALID, "nn_nearest_multi: null pointer");
  for (int64_t q0 = 0; q0 < Q; q0 += MULTI_MAX_Q) {
    const int64_t nq = Q - q0 < MULTI_MAX_Q ? Q - q0 : MULTI_MAX_Q;
    int rc = nn_multi_range(c, tree + q0, queries + q0 * SQMP_DOF, nullptr, nq, false, 0.0, 0, id_out + q0, nullptr, nullptr, nullptr);
    if (rc) return rc;
  }
  return SQMP_OK;
}

int sqmp_nn_radius_multi(sqmp_ctx *ctx, const int32_t *tree, const double *queries, const int32_t *exclude_id, int64_t Q,
                         double r, int k, int32_t *low_ids, int32_t *high_ids, int32_t *count_out) {
  CHECK_CTX();
  FLUSH_TREE_WRITES();
  if (Q < 0 || k < 1 || k > MULTI_MAX_K || (Q > 0 && (!tree || !queries || !low_ids || !high_ids || !count_out)))
    return fail(c, SQMP_ERR_INVALID, "nn_radius_multi: bad argument");
  for (int64_t q0 = 0; q0 < Q; q0 += MULTI_MAX_Q) {
    const int64_t nq = Q - q0 < MULTI_MAX_Q ? Q - q0 : MULTI_MAX_Q;
    int rc = nn_multi_range(c, tree + q0, queries + q0 * SQMP_DOF, exclude_id ? exclude_id + q0 : nullptr, nq, true, r, k,
                            nullptr, low_ids + q0 * k, high_ids + q0 * k, count_out + q0);
    if (rc) return rc;
  }
  return SQMP_OK;
}
