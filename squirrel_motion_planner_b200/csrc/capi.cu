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

struct Ctx {
  int device = 0;
  int sm_count = 148;
  RobotDev rb;        // metres (FK kernel)
  RobotDev rb_grid;   // grid units of the current voxel grid (validity kernel); == rb while no grid is set
  GridDev gd;
  int sph_link[SQMP_MAX_SPHERES];
  int slot_sph[SQMP_MAX_SLOTS];   // K1 slot -> sphere index, -1 for a pad
  int sph_slot[SQMP_MAX_SPHERES];
  int n_pairs = 0;
  std::vector<SelfPair> h_pairs;   // the description's pairs: a, b = SPHERE indices, rsum2 in metres^2
  SelfHPair *d_pairs = nullptr;   // device: hierarchical order (grid units)
  SelfPair *d_flat = nullptr;     // device: description order (slots, grid units)
  SelfA *d_as = nullptr;
  SelfBlock *d_blocks = nullptr;
  SelfDev self;
  bool grid_borrowed = false;   // the grid fields belong to another context (sqmp_share_voxel_grid): never freed here
  uint32_t *d_occ = nullptr;
  uint8_t *d_lb2 = nullptr;     // halo-padded (GridDev)
  uint16_t *d_ft = nullptr;
  FastDev fast;                 // tables of the fast validity kernel (validity_fast.cu)
  FastPair *d_fast_pairs = nullptr;
  FastAux *d_fast_aux = nullptr;
  double4 *d_fast_sph64 = nullptr;
  int fast_chain_ok = 0;        // description: translation-only base, then rotations only, spheres on the rotating part
  double reach_m = 0.0;         // max distance of a sphere centre from the base reference point (metres)
  int launch_overlap = 1;       // K1F grids may start while the previous K1F grid of the stream drains its last wave (PDL)
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
  int64_t small_call_states = 4096; // validity calls up to this many states take the low-latency path
  cudaStream_t streams[2] = {nullptr, nullptr};
  int64_t launches = 0;
  int fk_precision = 64;  // 64 | 32 (validity + FK chain arithmetic)
  int fk_float_ok = 0;
  int chain_specialization = 1;  // 1: use the compile-time unrolled chain when the description matches a known signature
  int nn_variant = 0;     // 0 auto | 1 QT | 2 WQ
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
  for (int i = 0; i < SQMP_MAX_SLOTS; ++i)
    r.sph_f[i] = make_float4(float(r.sph[i][0]), float(r.sph[i][1]), float(r.sph[i][2]), float(r.sph[i][3]));
}

// Grid-unit copy of the robot for the validity kernel: lengths / res, root offset measured from the grid origin,
// per-sphere distance-field threshold and scan half-extent, self-pair radii in grid units.  Re-derived whenever
// the grid (CC:76-88) or the link-disable mask (CC:90-102) changes.
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
  // slots: every group's spheres padded to a multiple of 4 with copies of its last sphere
  int n_slots = 0;
  for (int s = 0; s <= SQMP_MAX_SEG; ++s) {   // group s-1 (s = 0: the root) owns spheres [end[s], end[s+1])
    g.seg_sph_end[s] = n_slots;
    const int b = rb.seg_sph_end[s], e = rb.seg_sph_end[s + 1];
    for (int i = b; i < e; ++i) {
      c->slot_sph[n_slots] = i;
      c->sph_slot[i] = n_slots++;
    }
    while (e > b && (n_slots & 3)) c->slot_sph[n_slots++] = -(e - 1) - 1;   // pad: -(sphere copied)-1
  }
  g.seg_sph_end[SQMP_MAX_SEG + 1] = n_slots;
  g.n_spheres = n_slots;
  for (int sl = 0; sl < n_slots; ++sl) {
    const bool pad_slot = c->slot_sph[sl] < 0;
    const int i = pad_slot ? -c->slot_sph[sl] - 1 : c->slot_sph[sl];
    if (pad_slot) c->slot_sph[sl] = -1;
    for (int k = 0; k < 4; ++k) g.sph[sl][k] = rb.sph[i][k] * inv;
    const double r = rb.sph[i][3];
    if (!has) {
      g.sph_thr[sl] = -2;
      g.sph_ncell[sl] = 0;
      continue;
    }
    const double rc = r * inv;  // radius in cells; sqmp_set_voxel_grid guarantees rc <= 13, so rc^2 < 255
    const double t = std::floor(rc * rc + 1e-9);
    const bool enabled = !pad_slot && ((rb.sph_enabled[i >> 5] >> (i & 31)) & 1u);
    g.sph_thr[sl] = enabled ? short(t > 254.0 ? 254.0 : t) : short(-2);
    const double n = std::ceil(r / c->gd.res);
    g.sph_ncell[sl] = (unsigned char)(n > 250.0 ? 250.0 : n);
  }
  refresh_float_copies(g);
  refresh_float_copies(c->rb);
  // ---- tables of the fast validity kernel.  Its stage 1 runs on an fp32 copy of the sphere centre; err bounds the
  //      distance between that copy and the fp64 centre (cells): per axis <= 2^-24 * 13 (reach + 1) from rounding the
  //      frame, the sphere offset and three FMAs; doubled here, then sqrt(3) for the Euclidean norm.
  {
    c->fast.ok = 0;
    if (has && c->fast_chain_ok && c->gd.halo > 0) {
      const double reach_c = c->reach_m * inv + 1.0;
      const double err = std::sqrt(3.0) * std::ldexp(26.0 * (reach_c + 1.0), -24) + 1e-6;
      std::vector<FastPair> fp(size_t(n_slots) / 2 + 1);
      std::vector<FastAux> fa(size_t(n_slots) + 1);
      std::vector<double4> s64(size_t(n_slots) + 1);
      for (int sl = 0; sl < n_slots; ++sl) {
        const double rg = g.sph[sl][3];
        FastPair &f = fp[sl >> 1];
        const int h = sl & 1;
        f.x[h] = float(g.sph[sl][0]); f.y[h] = float(g.sph[sl][1]); f.z[h] = float(g.sph[sl][2]);
        const double t = std::floor((rg + err) * (rg + err) + 1e-9);
        f.thr[h] = g.sph_thr[sl] < 0 ? -2 : int(t > 254.0 ? 254.0 : t);
        const double lo = rg > err ? (rg - err) * (rg - err) * (1.0 - 1e-6) : -1.0;
        float lof = float(lo);
        if (double(lof) > lo) lof = std::nextafterf(lof, -INFINITY);
        fa[sl].r2_lo = lof;
        fa[sl].ncell = int(g.sph_ncell[sl]) + 1;
        s64[sl] = make_double4(g.sph[sl][0], g.sph[sl][1], g.sph[sl][2], rg);
      }
      cudaError_t e = cudaSuccess;
      if (!c->d_fast_pairs) e = cudaMalloc(&c->d_fast_pairs, sizeof(FastPair) * (SQMP_MAX_SLOTS / 2 + 1));
      if (e == cudaSuccess && !c->d_fast_aux) e = cudaMalloc(&c->d_fast_aux, sizeof(FastAux) * (SQMP_MAX_SLOTS + 1));
      if (e == cudaSuccess && !c->d_fast_sph64) e = cudaMalloc(&c->d_fast_sph64, sizeof(double4) * (SQMP_MAX_SLOTS + 1));
      if (e == cudaSuccess) e = cudaMemcpy(c->d_fast_pairs, fp.data(), sizeof(FastPair) * (n_slots / 2), cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMemcpy(c->d_fast_aux, fa.data(), sizeof(FastAux) * n_slots, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMemcpy(c->d_fast_sph64, s64.data(), sizeof(double4) * n_slots, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "fast validity tables", e);
      c->fast.pairs = c->d_fast_pairs;
      c->fast.aux = c->d_fast_aux;
      c->fast.sph64 = c->d_fast_sph64;
      c->fast.n_slots = n_slots;
      c->fast.ok = 1;
    }
  }
  // ---- self-collision hierarchy (SelfDev): blocks of group pairs -> A-spheres -> partner lists ----
  {
    int seg_of[SQMP_MAX_SPHERES];
    for (int sgi = 0; sgi <= SQMP_MAX_SEG; ++sgi)
      for (int i = rb.seg_sph_end[sgi]; i < rb.seg_sph_end[sgi + 1]; ++i) seg_of[i] = sgi;   // group index = segment + 1
    for (int gi = 0; gi <= SQMP_MAX_SEG; ++gi) g.grp_fidx[gi] = -1;
    g.n_frames = 0;
    for (int p = 0; p < c->n_pairs; ++p)
      for (int side = 0; side < 2; ++side) {
        const int gi = seg_of[side ? c->h_pairs[p].b : c->h_pairs[p].a];
        if (g.grp_fidx[gi] < 0) g.grp_fidx[gi] = g.n_frames++;
      }
    auto bound = [&](const std::vector<int> &sph, double (&cen)[3]) {   // enclosing sphere of a sphere set (metres)
      cen[0] = cen[1] = cen[2] = 0.0;
      for (int i : sph)
        for (int k = 0; k < 3; ++k) cen[k] += rb.sph[i][k] / double(sph.size());
      double R = 0.0;
      for (int i : sph) {
        double d2 = 0.0;
        for (int k = 0; k < 3; ++k) d2 += (rb.sph[i][k] - cen[k]) * (rb.sph[i][k] - cen[k]);
        R = std::fmax(R, std::sqrt(d2) + rb.sph[i][3]);
      }
      return R;
    };
    auto inflate2 = [&](double r_sum) {   // squared culling radius in grid units, padded against fp32 rounding
      const double v = (r_sum * (1.0 + 1e-4) + 1e-6) * inv;
      return float(v * v * (1.0 + 1e-6));
    };
    std::vector<SelfPair> flat(size_t(c->n_pairs) + 1);
    std::vector<SelfHPair> hier;
    std::vector<SelfA> as;
    std::vector<SelfBlock> blocks;
    for (int p = 0; p < c->n_pairs; ++p) {
      const int a = c->h_pairs[p].a, b = c->h_pairs[p].b;
      const double rs = (rb.sph[a][3] + rb.sph[b][3]) * inv;
      flat[p].a = (unsigned char)c->sph_slot[a];
      flat[p].b = (unsigned char)c->sph_slot[b];
      flat[p].fa = (unsigned char)g.grp_fidx[seg_of[a]];
      flat[p].fb = (unsigned char)g.grp_fidx[seg_of[b]];
      flat[p].rsum2 = float(rs * rs);
    }
    for (int ga = 0; ga <= SQMP_MAX_SEG; ++ga)
      for (int gb = 0; gb <= SQMP_MAX_SEG; ++gb) {
        std::vector<int> members, sa, sb;
        for (int p = 0; p < c->n_pairs; ++p)
          if (seg_of[c->h_pairs[p].a] == ga && seg_of[c->h_pairs[p].b] == gb) members.push_back(p);
        if (members.empty()) continue;
        for (int p : members) {
          const int a = c->h_pairs[p].a, b = c->h_pairs[p].b;
          if (std::find(sa.begin(), sa.end(), a) == sa.end()) sa.push_back(a);
          if (std::find(sb.begin(), sb.end(), b) == sb.end()) sb.push_back(b);
        }
        // The A side is split into spatial clusters of <= ~5 spheres (farthest-point seeds + a few Lloyd steps on the
        // sphere centres): the base group spans the whole robot, so one bounding sphere for it never rules a block
        // out, while a cluster's does.  All clusters of one (group A, group B) share the B-side bound, so the kernel
        // forms that bound in A's frame once and tests every cluster, and every A-sphere, against it as constants.
        std::vector<std::vector<int> > clusters;
        {
          const int n = int(sa.size());
          int k = n <= 6 ? 1 : (n + 4) / 5;
          if (k > 8) k = 8;
          std::vector<int> seed(1, 0), assign(size_t(n), 0);
          auto d2 = [&](int i, const double *q) {
            double v = 0.0;
            for (int t = 0; t < 3; ++t) v += (rb.sph[sa[size_t(i)]][t] - q[t]) * (rb.sph[sa[size_t(i)]][t] - q[t]);
            return v;
          };
          std::vector<std::array<double, 3> > cen;
          cen.push_back({rb.sph[sa[0]][0], rb.sph[sa[0]][1], rb.sph[sa[0]][2]});
          while (int(cen.size()) < k) {   // farthest point from the chosen seeds
            int best = 0;
            double bd = -1.0;
            for (int i = 0; i < n; ++i) {
              double m = 1e300;
              for (auto &q : cen) m = std::fmin(m, d2(i, q.data()));
              if (m > bd) { bd = m; best = i; }
            }
            cen.push_back({rb.sph[sa[size_t(best)]][0], rb.sph[sa[size_t(best)]][1], rb.sph[sa[size_t(best)]][2]});
          }
          for (int it = 0; it < 8; ++it) {
            for (int i = 0; i < n; ++i) {
              int best = 0;
              for (int q = 1; q < k; ++q)
                if (d2(i, cen[size_t(q)].data()) < d2(i, cen[size_t(best)].data())) best = q;
              assign[size_t(i)] = best;
            }
            for (int q = 0; q < k; ++q) {
              double sum[3] = {0, 0, 0};
              int cnt = 0;
              for (int i = 0; i < n; ++i)
                if (assign[size_t(i)] == q) {
                  for (int t = 0; t < 3; ++t) sum[t] += rb.sph[sa[size_t(i)]][t];
                  ++cnt;
                }
              if (cnt)
                for (int t = 0; t < 3; ++t) cen[size_t(q)][size_t(t)] = sum[t] / cnt;
            }
          }
          clusters.assign(size_t(k), std::vector<int>());
          for (int i = 0; i < n; ++i) clusters[size_t(assign[size_t(i)])].push_back(sa[size_t(i)]);
        }
        double cB[3];
        const double RB = bound(sb, cB);
        for (const std::vector<int> &cl : clusters) {
          if (cl.empty()) continue;
          double cA[3];
          const double RA = bound(cl, cA);
          SelfBlock blk;
          blk.fa = g.grp_fidx[ga];
          blk.fb = g.grp_fidx[gb];
          blk.a_begin = int(as.size());
          for (int k = 0; k < 3; ++k) {
            blk.ca[k] = float(cA[k] * inv);
            blk.cb[k] = float(cB[k] * inv);
          }
          blk.thr2 = inflate2(RA + RB);
          blk.pad = 0.0f;
          for (int a : cl) {
            SelfA ea;
            ea.a = c->sph_slot[a];
            ea.pad = 0;
            ea.cx = float(rb.sph[a][0] * inv); ea.cy = float(rb.sph[a][1] * inv); ea.cz = float(rb.sph[a][2] * inv);
            ea.p_begin = int(hier.size());
            for (int p : members)
              if (c->h_pairs[p].a == a) {
                const int b = c->h_pairs[p].b;
                SelfHPair hp;
                hp.bx = float(rb.sph[b][0] * inv); hp.by = float(rb.sph[b][1] * inv); hp.bz = float(rb.sph[b][2] * inv);
                hp.rsum2 = flat[p].rsum2;
                hier.push_back(hp);
              }
            ea.p_end = int(hier.size());
            ea.thr2 = inflate2(rb.sph[a][3] + RB);
            as.push_back(ea);
          }
          blk.a_end = int(as.size());
          blocks.push_back(blk);
        }
      }
    hier.resize(size_t(c->n_pairs) + 4);   // the kernels read partners and A-spheres four at a time (masked)
    const size_t n_as_real = as.size();
    as.resize(n_as_real + 4);
    auto upload = [&](auto *&dptr, const auto &vec, size_t min_count) -> cudaError_t {
      using E = typename std::remove_reference<decltype(vec)>::type::value_type;
      const size_t count = vec.size() > min_count ? vec.size() : min_count;
      if (!dptr) {
        cudaError_t e = cudaMalloc(&dptr, sizeof(E) * (count + 1));
        if (e != cudaSuccess) return e;
      }
      return vec.empty() ? cudaSuccess : cudaMemcpy(dptr, vec.data(), sizeof(E) * vec.size(), cudaMemcpyHostToDevice);
    };
    // sizes depend only on the pair list, which is fixed at sqmp_create: allocated once with that capacity
    cudaError_t e = upload(c->d_pairs, hier, size_t(c->n_pairs) + 4);
    if (e == cudaSuccess) e = upload(c->d_flat, flat, size_t(c->n_pairs) + 1);
    if (e == cudaSuccess) e = upload(c->d_as, as, size_t(c->n_pairs) + 4);
    if (e == cudaSuccess) e = upload(c->d_blocks, blocks, size_t(SQMP_MAX_SEG + 1) * (SQMP_MAX_SEG + 1) * 8);   // <= 8 clusters per group pair
    if (e != cudaSuccess) return fail(c, SQMP_ERR_CUDA, "self-collision tables", e);
    c->self.hpairs = c->d_pairs;
    c->self.flat = c->d_flat;
    c->self.as = c->d_as;
    c->self.blocks = c->d_blocks;
    c->self.n_pairs = c->n_pairs;
    c->self.n_as = int(n_as_real);
    c->self.n_blocks = int(blocks.size());
  }
  return SQMP_OK;
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

static int tree_reserve(Ctx *c, TreeDev &t, int64_t need) {
  if (need <= t.cap) return SQMP_OK;
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
  // fp32 shadow (float4 + float4 + float per node)
  float4 *fa = nullptr, *fb = nullptr;
  float *fn = nullptr;
  e = cudaMalloc(&fa, size_t(pitch) * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(&fb, size_t(pitch) * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(&fn, size_t(pitch) * sizeof(float));
  if (e == cudaSuccess && t.n > 0) {
    e = cudaMemcpy(fa, t.fa, size_t(t.n) * sizeof(float4), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(fb, t.fb, size_t(t.n) * sizeof(float4), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(fn, t.fn, size_t(t.n) * sizeof(float), cudaMemcpyDeviceToDevice);
  }
  if (e != cudaSuccess) {
    cudaFree(blk);
    if (fa) cudaFree(fa);
    if (fb) cudaFree(fb);
    if (fn) cudaFree(fn);
    return fail(c, SQMP_ERR_NOMEM, "cudaMalloc(tree fp32 shadow)", e);
  }
  if (t.q[0]) cudaFree(t.q[0]);
  if (t.fa) cudaFree(t.fa);
  if (t.fb) cudaFree(t.fb);
  if (t.fn) cudaFree(t.fn);
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

static int pick_variant(const Ctx *c, int64_t Q) {
  if (c->nn_variant >= 1 && c->nn_variant <= 3) return c->nn_variant;
  return Q >= 2048 ? 3 : 2;
}
static int pick_chunk_hint(const Ctx *c, int variant, int64_t N, int64_t Q) {
  if (N <= 0) return 1;
  int64_t chunks;
  if (variant == 1 || variant == 3) {
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

const char *sqmp_version(void) { return "libsqmp_b200 r1 (sm_100a)"; }

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
  if (!d || d->n_seg < 1 || d->n_seg > SQMP_MAX_SEG || d->n_spheres < 0 || d->n_spheres > SQMP_MAX_SPHERES ||
      d->n_pairs < 0 || d->n_pairs > SQMP_MAX_PAIRS)
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
  RobotDev &rb = c->rb;
  rb.n_seg = d->n_seg;
  rb.n_spheres = d->n_spheres;
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
  for (int s = 0; s <= SQMP_MAX_SEG + 1; ++s) rb.seg_sph_end[s] = 0;
  c->fk_float_ok = 1;
  for (int i = 0; i < d->n_spheres; ++i) {
    const int sg = d->sph_seg[i];
    if (sg < prev || sg < -1 || sg >= d->n_seg || d->sph_link[i] < 0 || d->sph_link[i] > 63 ||
        !(d->sph_radius[i] > 0.0)) {
      delete c;
      return bail(SQMP_ERR_INVALID);
    }
    prev = sg;
    rb.seg_sph_end[sg + 2] = i + 1;
    rb.sph[i][0] = d->sph_center[3 * i + 0];
    rb.sph[i][1] = d->sph_center[3 * i + 1];
    rb.sph[i][2] = d->sph_center[3 * i + 2];
    rb.sph[i][3] = d->sph_radius[i];
    c->sph_link[i] = d->sph_link[i];
    if (sg <= rb.ref_seg) c->fk_float_ok = 0;  // float chain starts after the prismatic base
  }
  for (int s = 1; s <= SQMP_MAX_SEG + 1; ++s)
    if (rb.seg_sph_end[s] < rb.seg_sph_end[s - 1]) rb.seg_sph_end[s] = rb.seg_sph_end[s - 1];
  for (int w = 0; w < SQMP_MAX_SPHERES / 32; ++w) rb.sph_enabled[w] = 0xffffffffu;
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
      for (int i = rb.seg_sph_end[s + 1]; i < rb.seg_sph_end[s + 2]; ++i)
        c->reach_m = std::fmax(c->reach_m, cum + std::sqrt(rb.sph[i][0] * rb.sph[i][0] + rb.sph[i][1] * rb.sph[i][1] +
                                                            rb.sph[i][2] * rb.sph[i][2]));
    }
    if (!(c->reach_m < 1.0e6)) c->fast_chain_ok = 0;
  }

  c->n_pairs = d->n_pairs;
  c->h_pairs.resize(size_t(d->n_pairs) + 1);
  for (int p = 0; p < d->n_pairs; ++p) {
    const int a = d->pair_a[p], b = d->pair_b[p];
    if (a < 0 || b < 0 || a >= d->n_spheres || b >= d->n_spheres) {
      delete c;
      return bail(SQMP_ERR_INVALID);
    }
    const double rs = d->sph_radius[a] + d->sph_radius[b];
    c->h_pairs[p].a = (unsigned char)a;
    c->h_pairs[p].b = (unsigned char)b;
    c->h_pairs[p].fa = c->h_pairs[p].fb = 0;
    c->h_pairs[p].rsum2 = float(rs * rs);
  }
  bool ok = true;
  ok = ok && cudaHostAlloc((void **)&c->pin_h, PIN_BYTES, cudaHostAllocMapped) == cudaSuccess;
  ok = ok && cudaHostGetDevicePointer((void **)&c->pin_d, c->pin_h, 0) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_ticket, 64) == cudaSuccess && cudaMemset(c->d_ticket, 0, 64) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->streams[0], cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->streams[1], cudaStreamNonBlocking) == cudaSuccess;
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
  if (ctx->d_pairs) cudaFree(ctx->d_pairs);
  if (ctx->d_flat) cudaFree(ctx->d_flat);
  if (ctx->d_as) cudaFree(ctx->d_as);
  if (ctx->d_blocks) cudaFree(ctx->d_blocks);
  if (!ctx->grid_borrowed) {
    if (ctx->d_occ) cudaFree(ctx->d_occ);
    if (ctx->d_lb2) cudaFree(ctx->d_lb2);
    if (ctx->d_ft) cudaFree(ctx->d_ft);
  }
  if (ctx->d_fast_pairs) cudaFree(ctx->d_fast_pairs);
  if (ctx->d_fast_aux) cudaFree(ctx->d_fast_aux);
  if (ctx->d_fast_sph64) cudaFree(ctx->d_fast_sph64);
  for (auto &t : ctx->trees) {
    if (t.q[0]) cudaFree(t.q[0]);
    if (t.fa) cudaFree(t.fa);
    if (t.fb) cudaFree(t.fb);
    if (t.fn) cudaFree(t.fn);
  }
  if (ctx->nn.part_d) cudaFree(ctx->nn.part_d);
  if (ctx->nn.part_id) cudaFree(ctx->nn.part_id);
  if (ctx->nn.hits) cudaFree(ctx->nn.hits);
  for (auto &b : ctx->scratch)
    if (b.p) cudaFree(b.p);
  if (ctx->pin_h) cudaFreeHost(ctx->pin_h);
  if (ctx->pin2_h) cudaFreeHost(ctx->pin2_h);
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  for (auto &s : ctx->streams)
    if (s) cudaStreamDestroy(s);
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
    if (value < 0 || value > 3) return fail(c, SQMP_ERR_INVALID, "nn_variant must be 0..3");
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
  double rmax = 0.0;
  for (int i = 0; i < c->rb.n_spheres; ++i) rmax = std::fmax(rmax, c->rb.sph[i][3]);
  // the 16-bit feature field stores cell offsets in [-15, 15]: every occupied cell within reach of a sphere
  // (|offset| <= ceil(r/res) + 1 per axis) must be representable
  if (rmax / res > 13.0) return fail(c, SQMP_ERR_INVALID, "sphere radius too large for this resolution (> 13 voxels)");
  int window = int(std::ceil(rmax / res)) + 1;
  if (window < 3) window = 3;
  if (window > 15) window = 15;
  // halo of the distance / feature fields: every cell a sphere centre can fall into while the base is inside the grid
  int halo = 0;
  if (c->fast_chain_ok) {
    const double h = std::ceil(c->reach_m / res) + 2.0;
    const double pn = (gx + 2.0 * h) * (gy + 2.0 * h) * (gz + 2.0 * h);
    if (h <= 2048.0 && pn < 4294967296.0) halo = int(h);   // else: no halo, generic kernel only
  }
  const int64_t PX = gx + 2 * halo, PY = gy + 2 * halo, PZ = gz + 2 * halo;
  const int64_t pvox = PX * PY * PZ;
  const int row_words = (gx + 31) / 32 + 1;
  const size_t lin_words = size_t((nvox + 31) / 32) + 1;
  const size_t occ_words = size_t(gy) * gz * row_words;
  CU(cudaDeviceSynchronize());
  uint32_t *d_lin = nullptr, *d_occ = nullptr;
  int8_t *d_a = nullptr;
  uint8_t *d_lb2 = nullptr;
  uint16_t *d_b = nullptr, *d_ft = nullptr;
  auto cleanup = [&]() {
    if (d_lin) cudaFree(d_lin);
    if (d_a) cudaFree(d_a);
    if (d_b) cudaFree(d_b);
  };
  cudaError_t e = cudaMalloc(&d_lin, lin_words * 4);
  if (e == cudaSuccess) e = cudaMemset(d_lin, 0, lin_words * 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_lin, bits, size_t((nvox + 31) / 32) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&d_occ, occ_words * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_a, size_t(pvox));
  if (e == cudaSuccess) e = cudaMalloc(&d_b, size_t(pvox) * 2);
  if (e == cudaSuccess) e = cudaMalloc(&d_lb2, size_t(pvox) + 16);   // + a zero byte at [pvox] (GridDev::zero_off)
  if (e == cudaSuccess) e = cudaMemset(d_lb2 + pvox, 0, 16);
  if (e == cudaSuccess) e = cudaMalloc(&d_ft, size_t(pvox) * 2);
  if (e == cudaSuccess) e = launch_pack_rows(d_lin, gx, gy, gz, row_words, d_occ, nullptr);
  int nl = 1;
  if (e == cudaSuccess) e = launch_ft_build(d_occ, row_words, gx, gy, gz, halo, window, d_a, d_b, d_lb2, d_ft, nullptr, &nl);
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
  GridDev &g = c->gd;
  g.sx = unsigned(PX);
  g.splane = unsigned(PX * PY);
  g.org_off = unsigned((int64_t(halo) * PY + halo) * PX + halo);
  g.halo = halo;
  g.zero_off = unsigned(pvox);
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
  if (o->reach_m != c->reach_m || o->fast_chain_ok != c->fast_chain_ok)
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
  return refresh_grid_robot(c);
}

int sqmp_set_disabled_links(sqmp_ctx *ctx, uint64_t link_mask) {
  CHECK_CTX();
  for (int w = 0; w < SQMP_MAX_SPHERES / 32; ++w) c->rb.sph_enabled[w] = 0;
  for (int i = 0; i < c->rb.n_spheres; ++i)
    if (!((link_mask >> c->sph_link[i]) & 1ull)) c->rb.sph_enabled[i >> 5] |= 1u << (i & 31);
  CU(cudaDeviceSynchronize());
  return refresh_grid_robot(c);
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
  if (c->k1_kernel == 1 || c->fk_precision != 64) return false;
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
    CU(launch_check_states_fast(c->rb_grid, c->gd, c->self, c->fast, d_q, nullptr, 0, S, flags, d_mask,
                                c->chain_specialization != 0, k1f_may_overlap(c, st), st, &nl));
  else
    CU(launch_check_states(c->rb_grid, c->gd, c->self, d_q, nullptr, 0, S, flags, d_mask, c->fk_precision,
                           st, &nl));
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
  CU(launch_check_states_fast(c->rb_grid, c->gd, c->self, c->fast, d_q, nullptr, 0, S, flags, nullptr,
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
  } else {
    int nl = 0;
    const bool fast = use_fast_k1(c, flags);
    if (fast)
      CU(launch_check_states_fast(c->rb_grid, c->gd, c->self, c->fast, d_a, d_b, num_points, S, flags, d_mask,
                                  c->chain_specialization != 0, k1f_may_overlap(c, st), st, &nl));
    else
      CU(launch_check_states(c->rb_grid, c->gd, c->self, d_a, d_b, num_points, S, flags, d_mask,
                             c->fk_precision, st, &nl));
    c->launches += nl;
    if (fast && !d_first) k1f_launched(c, st);
  }
  if (d_first) {
    CU(launch_first_invalid(d_mask, E, P1, d_first, st));
    c->launches += 1;
  }
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
  void *da, *db, *dm, *df;
  int rc;
  if ((rc = ensure(c, 0, size_t(E) * SQMP_DOF * 8, &da))) return rc;
  if ((rc = ensure(c, 1, size_t(E) * SQMP_DOF * 8, &db))) return rc;
  if ((rc = ensure(c, 2, size_t((S + 31) / 32) * 4, &dm))) return rc;
  if ((rc = ensure(c, 3, size_t(E) * 4, &df))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(da, qa, size_t(E) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(db, qb, size_t(E) * SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  rc = sqmp_check_edges_dev(ctx, (const double *)da, (const double *)db, E, num_points, flags, (uint32_t *)dm,
                            first_invalid ? (int32_t *)df : nullptr, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(mask_out, dm, size_t((S + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
  if (first_invalid) CU(cudaMemcpyAsync(first_invalid, df, size_t(E) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SQMP_OK;
}

int sqmp_get_collisions(sqmp_ctx *ctx, const double q[SQMP_DOF], uint8_t *link_map_hit, uint8_t *pair_hit) {
  CHECK_CTX();
  if (!q || !link_map_hit) return fail(c, SQMP_ERR_INVALID, "get_collisions: null pointer");
  void *dq, *ds, *dp;
  int rc;
  if ((rc = ensure(c, 0, SQMP_DOF * 8, &dq))) return rc;
  if ((rc = ensure(c, 2, SQMP_MAX_SLOTS, &ds))) return rc;
  if ((rc = ensure(c, 3, size_t(c->n_pairs) + 1, &dp))) return rc;
  cudaStream_t st = c->streams[0];
  CU(cudaMemcpyAsync(dq, q, SQMP_DOF * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(ds, 0, SQMP_MAX_SLOTS, st));
  CU(cudaMemsetAsync(dp, 0, size_t(c->n_pairs) + 1, st));
  CU(launch_get_collisions(c->rb_grid, c->gd, c->self, (const double *)dq, (uint8_t *)ds, (uint8_t *)dp, st));
  c->launches += 1;
  uint8_t sh[SQMP_MAX_SLOTS];
  CU(cudaMemcpyAsync(sh, ds, SQMP_MAX_SLOTS, cudaMemcpyDeviceToHost, st));
  if (pair_hit && c->n_pairs) CU(cudaMemcpyAsync(pair_hit, dp, size_t(c->n_pairs), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::memset(link_map_hit, 0, 64);
  for (int sl = 0; sl < c->rb_grid.n_spheres; ++sl)
    if (sh[sl] && c->slot_sph[sl] >= 0) link_map_hit[c->sph_link[c->slot_sph[sl]]] = 1;
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

// ---------------------------------------------------------------------------------------------------- tree
#define CHECK_TREE()                                                                 \
  if (tree < 0 || tree >= SQMP_MAX_TREES) return fail(c, SQMP_ERR_INVALID, "bad tree index"); \
  TreeDev &t = c->trees[tree];

int sqmp_tree_reset(sqmp_ctx *ctx, int tree) {
  CHECK_CTX();
  CHECK_TREE();
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
  // stream-ordered, not synchronised: the node travels in the kernel parameters, and every later call that reads
  // the tree runs on (or synchronises with) the same stream.  Growing the buffers synchronises the device itself.
  if (t.n + 1 > t.cap) CU(cudaDeviceSynchronize());
  int rc = tree_reserve(c, t, t.n + 1);
  if (rc) return rc;
  tree_track_norm(t, q, 1);
  CU(launch_tree_set_one(t, t.n, q, cost, c->streams[0]));
  c->tree_writes_pending = true;
  c->launches += 1;
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
  CU(launch_tree_set_one(t, id, q, cost, c->streams[0]));   // stream-ordered (see sqmp_tree_append)
  c->tree_writes_pending = true;
  c->launches += 1;
  return SQMP_OK;
}
int sqmp_tree_update_costs(sqmp_ctx *ctx, int tree, const int32_t *ids, const double *cost, int64_t n) {
  CHECK_CTX();
  CHECK_TREE();
  if (n < 0 || (n > 0 && (!ids || !cost))) return fail(c, SQMP_ERR_INVALID, "tree_update_costs: null pointer");
  if (n == 0) return SQMP_OK;
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= t.n) return fail(c, SQMP_ERR_INVALID, "tree_update_costs: bad id");
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
  const int variant = pick_variant(c, Q);
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
  const int variant = pick_variant(c, Q);
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
