// validity.cu — K1: fused (interpolate ->) forward kinematics -> sphere-vs-voxel -> (self pairs ->) 1 bit/state.
//
// Reference path replaced (file:line in /root/reference, tags as in include/sqmp_b200.h):
//   CollisionChecker::isInCollision CC:104-121, updateTransforms CC:519-539, checkMapCollision CC:554-592,
//   checkSelfCollision CC:541-552; BiRRTstarPlanner::interpolateConfigurations BS:4443-4490, isEdgeValid BS:6864-6895.
// The reference walks 39 links and calls fcl::collide per link per state, serially, with early exits.  Here one
// thread owns one state: it walks the 8 folded chain segments keeping a single frame in registers, transforms the
// spheres of each rigid group as soon as that group's frame exists, and tests every sphere against
//   (1) a byte distance field (conservative "certainly free" filter, one L2 gather), then, only if that fails,
//   (2) the exact test "some occupied voxel AABB within r of the centre" by scanning bit-packed x-rows:
//       per (j,k) row the admissible x-interval is closed-form, so a row costs one (rarely two) word loads.
// The verdicts of a warp are packed by one ballot into one 32-bit word — the mask layout of the C ABI.
#include "fk_device.cuh"

namespace sqmp {

// ------------------------------------------------------------------------------------------------
// exact sphere vs occupied-voxel-boxes test, in GRID UNITS: voxel (i,j,k) is the unit cube [i,i+1]x[j,j+1]x[k,k+1],
// the sphere has centre u and radius rg = r / res.  true <=> dist(u, some occupied unit cube) <= rg.
// Per (j,k) row the admissible x-interval is closed-form, so a row costs one (rarely two) word loads.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool exact_scan(const GridDev &g, double ux, double uy, double uz, double rg) {
  const double r2 = rg * rg;
  const int k0 = max(__double2int_rd(uz - rg), 0), k1 = min(__double2int_rd(uz + rg), g.gz - 1);
  const int j0 = max(__double2int_rd(uy - rg), 0), j1 = min(__double2int_rd(uy + rg), g.gy - 1);
  double kd = double(k0);
  for (int k = k0; k <= k1; ++k, kd += 1.0) {
    const double dz = fmax(fmax(kd - uz, uz - (kd + 1.0)), 0.0);
    const double remz = fma(-dz, dz, r2);
    if (remz < 0.0) continue;
    const uint32_t *slab = g.occ + size_t(k) * g.gy * g.row_words;
    double jd = double(j0);
    for (int j = j0; j <= j1; ++j, jd += 1.0) {
      const double dy = fmax(fmax(jd - uy, uy - (jd + 1.0)), 0.0);
      const double rem = fma(-dy, dy, remz);
      if (rem < 0.0) continue;
      const double w = sqrt(rem);  // half-width of the admissible x interval
      // cube i is within reach  <=>  i + 1 >= ux - w  and  i <= ux + w
      int i0 = max(__double2int_ru(ux - w) - 1, 0);
      const int i1 = min(__double2int_rd(ux + w), g.gx - 1);
      const uint32_t *row = slab + size_t(j) * g.row_words;
      while (i0 <= i1) {  // at most 2*ceil(rg)+2 wide: one word, two when it straddles a word boundary
        const int sh = i0 & 31;
        const int n = min(32 - sh, i1 - i0 + 1);
        const uint32_t bits = (row[i0 >> 5] >> sh) & (n >= 32 ? ~0u : ((1u << n) - 1u));
        if (bits) return true;
        i0 += n;
      }
    }
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 128;
constexpr int K1_QCAP = 64;       // deferred exact-scan items per warp (drained 32 at a time)
constexpr int K1_BATCH = 4;       // spheres whose distance-field gathers are in flight together
constexpr unsigned FULL = 0xffffffffu;

struct WarpQueue {   // shared memory, one per warp
  double x[K1_QCAP], y[K1_QCAP], z[K1_QCAP], r[K1_QCAP];
  int info[K1_QCAP];   // sphere index << 8 | owner lane
};

// MODE 0: verdict bit per state.  MODE 1 (diagnostics, one state): per-sphere / per-pair hits, no early exit.
//
// rb is the GRID-UNIT copy of the robot (translations, sphere centres and radii divided by res, root offset
// measured from the grid origin), so a sphere centre comes out of the chain directly in voxel coordinates.
//
// Divergence control: the cheap filter (one byte gather from the distance field) runs in lockstep for all lanes;
// the few spheres it cannot clear are NOT scanned on the spot (that left ~4 of 32 lanes active) but pushed, by
// ballot compaction, on a per-warp queue in shared memory, and scanned 32 at a time with one item per lane.
template <typename T, bool EDGES, int MODE>
__global__ void __launch_bounds__(K1_THREADS)
k1_check_states(const __grid_constant__ RobotDev rb, const __grid_constant__ GridDev gd,
                const SelfPair *__restrict__ pairs, int n_pairs, const double *__restrict__ q_in,
                const double *__restrict__ q_end, int num_points, int64_t S, unsigned flags,
                uint32_t *__restrict__ mask_out, uint8_t *__restrict__ sphere_hit, uint8_t *__restrict__ pair_hit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpQueue *wq = reinterpret_cast<WarpQueue *>(smem_raw) + (threadIdx.x >> 5);
  float *s_rel = reinterpret_cast<float *>(smem_raw + sizeof(WarpQueue) * (K1_THREADS / 32));  // [3][n][blockDim]
  const int lane = threadIdx.x & 31;
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = s < S;
  const int64_t sl = active ? s : S - 1;  // idle lanes of the last warp shadow the last state (lockstep loops)
  const bool do_map = (flags & SQMP_CHECK_MAP) && gd.has_grid;
  const bool do_self = (flags & SQMP_CHECK_SELF) != 0;
  bool hit = false;
  int qcount = 0;  // warp-uniform

  // ---- the state: either given, or interpolated from the edge end points (BS:4461-4462, 4478; no FMA) ----
  double q[SQMP_DOF];
  if (EDGES) {
    const int P1 = num_points + 1;
    const int64_t e = sl / P1;
    const int kk = int(sl - e * P1);
    const double4 *a4 = reinterpret_cast<const double4 *>(q_in + e * SQMP_DOF);
    const double4 *b4 = reinterpret_cast<const double4 *>(q_end + e * SQMP_DOF);
    const double4 a0 = a4[0], a1 = a4[1], b0 = b4[0], b1 = b4[1];
    const double st[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const double en[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const double np = double(num_points), fk = double(kk);
#pragma unroll
    for (int j = 0; j < SQMP_DOF; ++j)
      q[j] = __dadd_rn(st[j], __dmul_rn(fk, __ddiv_rn(__dsub_rn(en[j], st[j]), np)));
  } else {
    const double4 *p4 = reinterpret_cast<const double4 *>(q_in + sl * SQMP_DOF);
    const double4 v0 = p4[0], v1 = p4[1];
    q[0] = v0.x; q[1] = v0.y; q[2] = v0.z; q[3] = v0.w;
    q[4] = v1.x; q[5] = v1.y; q[6] = v1.z; q[7] = v1.w;
  }
  // prismatic joint values are lengths: into grid units with everything else
#pragma unroll
  for (int j = 0; j < SQMP_DOF; ++j) q[j] = ((rb.trans_mask >> j) & 1) ? q[j] * rb.len_scale : q[j];

  // ---- base reference point = position after the leading prismatic run ----
  double pref[3] = {rb.root_off[0], rb.root_off[1], rb.root_off[2]};
  if (do_self || sizeof(T) == 4) {
    for (int sg = 0; sg <= rb.ref_seg; ++sg) {
      const int qi = rb.seg_qidx[sg];
      double qv = q[0];
#pragma unroll
      for (int j = 1; j < SQMP_DOF; ++j) qv = (qi == j) ? q[j] : qv;
#pragma unroll
      for (int i = 0; i < 3; ++i) pref[i] += rb.seg_pre[sg][9 + i] + rb.seg_axis[sg][i] * qv;
    }
  }

  // ---- chain walk; T = double walks the whole chain, T = float walks the part after the prismatic base
  //      relative to pref (host guarantees the leading run is translation-only) ----
  Frame<T> f;
  double off[3];
  int sg0;
  if (sizeof(T) == 8) {
    frame_set_translation<T>(f, T(rb.root_off[0]), T(rb.root_off[1]), T(rb.root_off[2]));
    off[0] = off[1] = off[2] = 0.0;
    sg0 = 0;
  } else {
    frame_set_translation<T>(f, T(0), T(0), T(0));
    off[0] = pref[0]; off[1] = pref[1]; off[2] = pref[2];
    sg0 = rb.ref_seg + 1;
  }

  // scan the queued items [base, base+n) with one item per lane; route the verdicts back to the owning lanes
  auto drain = [&](int base, int n) {
    __syncwarp();
    const unsigned known = __ballot_sync(FULL, hit);
    bool h = false;
    int owner = 0;
    if (lane < n) {
      const int info = wq->info[base + lane];
      owner = info & 31;
      if (!((known >> owner) & 1u))
        h = exact_scan(gd, wq->x[base + lane], wq->y[base + lane], wq->z[base + lane], wq->r[base + lane]);
    }
    const unsigned hm = __reduce_or_sync(FULL, h ? (1u << owner) : 0u);
    hit |= (hm >> lane) & 1u;
    __syncwarp();
  };

  for (int sg = -1; sg < rb.n_seg; ++sg) {
    if (sg >= sg0) frame_advance<T>(f, rb, sg, q);
    else if (sg >= 0) continue;
    const int b = rb.seg_sph_end[sg + 1], e = rb.seg_sph_end[sg + 2];
    for (int i0 = b; i0 < e; i0 += K1_BATCH) {
      double ux[K1_BATCH], uy[K1_BATCH], uz[K1_BATCH];
      int ev[K1_BATCH];
      bool need[K1_BATCH];
#pragma unroll
      for (int u = 0; u < K1_BATCH; ++u) {  // transform + gather (independent loads, issued back to back)
        const int i = min(i0 + u, e - 1);
        const T lx = T(rb.sph[i][0]), ly = T(rb.sph[i][1]), lz = T(rb.sph[i][2]);
        const T wx = fma(f.r[0], lx, fma(f.r[1], ly, fma(f.r[2], lz, f.p[0])));
        const T wy = fma(f.r[3], lx, fma(f.r[4], ly, fma(f.r[5], lz, f.p[1])));
        const T wz = fma(f.r[6], lx, fma(f.r[7], ly, fma(f.r[8], lz, f.p[2])));
        ux[u] = double(wx) + off[0]; uy[u] = double(wy) + off[1]; uz[u] = double(wz) + off[2];
        if (do_self && i0 + u < e) {
          const int n = rb.n_spheres;
          s_rel[(0 * n + i) * blockDim.x + threadIdx.x] = float(ux[u] - pref[0]);
          s_rel[(1 * n + i) * blockDim.x + threadIdx.x] = float(uy[u] - pref[1]);
          s_rel[(2 * n + i) * blockDim.x + threadIdx.x] = float(uz[u] - pref[2]);
        }
        // saturating conversions: anything absurdly far lands outside every range test below
        const int ix = __double2int_rd(ux[u]), iy = __double2int_rd(uy[u]), iz = __double2int_rd(uz[u]);
        const int pad = rb.sph_ncell[i] + 1;
        const bool near_grid = unsigned(ix + pad) < unsigned(gd.gx + 2 * pad) &&
                               unsigned(iy + pad) < unsigned(gd.gy + 2 * pad) &&
                               unsigned(iz + pad) < unsigned(gd.gz + 2 * pad);
        const bool inside = unsigned(ix) < unsigned(gd.gx) && unsigned(iy) < unsigned(gd.gy) && unsigned(iz) < unsigned(gd.gz);
        const bool enabled = (MODE == 1) || ((rb.sph_enabled[i >> 5] >> (i & 31)) & 1u);
        need[u] = do_map && enabled && near_grid && (i0 + u < e) && (MODE == 1 || !hit);
        // box-to-box distance of this cell to the nearest occupied cell bounds dist(c, occupied) from below
        ev[u] = (need[u] && inside) ? int(gd.edt[(size_t(iz) * gd.gy + iy) * gd.gx + ix]) : 0;
      }
#pragma unroll
      for (int u = 0; u < K1_BATCH; ++u) {  // filter verdict; survivors are queued for the exact scan
        const int i = min(i0 + u, e - 1);
        const bool nd = need[u] && ev[u] <= int(rb.sph_thr8[i]);
        if (MODE == 1) {
          if (nd) sphere_hit[i] = exact_scan(gd, ux[u], uy[u], uz[u], rb.sph[i][3]) ? 1 : 0;
          continue;
        }
        const unsigned m = __ballot_sync(FULL, nd);
        if (m) {
          if (nd) {
            const int pos = qcount + __popc(m & ((1u << lane) - 1u));
            wq->x[pos] = ux[u]; wq->y[pos] = uy[u]; wq->z[pos] = uz[u]; wq->r[pos] = rb.sph[i][3];
            wq->info[pos] = (i << 8) | lane;
          }
          qcount += __popc(m);
          if (qcount >= 32) {
            drain(qcount - 32, 32);
            qcount -= 32;
          }
        }
      }
    }
    // whole warp decided (CC:114-115: map before self, first hit wins) -> stop walking the chain
    if (MODE == 0 && !do_self && qcount == 0 && __all_sync(FULL, hit)) break;
  }
  if (MODE == 0 && qcount > 0) {
    drain(0, qcount);
    qcount = 0;
  }

  // ---- self collision over the enabled cross-group sphere pairs (CC:541-552) ----
  if (do_self && (MODE == 1 || !hit)) {
    if (rb.self_const_collision) hit = true;
    const int n = rb.n_spheres;
    for (int p = 0; p < n_pairs; ++p) {
      const SelfPair pr = pairs[p];
      const float dx = s_rel[(0 * n + pr.a) * blockDim.x + threadIdx.x] - s_rel[(0 * n + pr.b) * blockDim.x + threadIdx.x];
      const float dy = s_rel[(1 * n + pr.a) * blockDim.x + threadIdx.x] - s_rel[(1 * n + pr.b) * blockDim.x + threadIdx.x];
      const float dz = s_rel[(2 * n + pr.a) * blockDim.x + threadIdx.x] - s_rel[(2 * n + pr.b) * blockDim.x + threadIdx.x];
      const bool h = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= pr.rsum2;
      hit |= h;
      if (MODE == 1) pair_hit[p] = h ? 1 : 0;
      if (MODE == 0 && hit) break;
    }
  }

  if (MODE == 0) {
    const unsigned word = __ballot_sync(FULL, hit && active);
    if (lane == 0 && s < ((S + 31) & ~int64_t(31))) mask_out[s >> 5] = word;
  }
}

static size_t k1_smem_bytes(const RobotDev &rb, unsigned flags, int threads) {
  size_t b = sizeof(WarpQueue) * (K1_THREADS / 32);
  if (flags & SQMP_CHECK_SELF) b += size_t(3) * rb.n_spheres * threads * sizeof(float);
  return b;
}

template <typename T>
static cudaError_t launch_k1(const RobotDev &rb, const GridDev &gd, const SelfPair *pairs, int n_pairs,
                             const double *d_q, const double *d_q_end, int num_points, int64_t S, unsigned flags,
                             uint32_t *d_mask, cudaStream_t st) {
  const size_t smem = k1_smem_bytes(rb, flags, K1_THREADS);
  const int64_t blocks = (S + K1_THREADS - 1) / K1_THREADS;
  if (blocks <= 0) return cudaSuccess;
  if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
  if (d_q_end) {
    auto kern = k1_check_states<T, true, 0>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    kern<<<unsigned(blocks), K1_THREADS, smem, st>>>(rb, gd, pairs, n_pairs, d_q, d_q_end, num_points, S, flags,
                                                      d_mask, nullptr, nullptr);
  } else {
    auto kern = k1_check_states<T, false, 0>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    kern<<<unsigned(blocks), K1_THREADS, smem, st>>>(rb, gd, pairs, n_pairs, d_q, nullptr, 0, S, flags, d_mask,
                                                      nullptr, nullptr);
  }
  return cudaGetLastError();
}

cudaError_t launch_check_states(const RobotDev &rb, const GridDev &gd, const SelfPair *pairs, int n_pairs,
                                const double *d_q, const double *d_q_end, int num_points, int64_t S, unsigned flags,
                                uint32_t *d_mask, int fk_precision, cudaStream_t st, int *n_launch) {
  if (n_launch) *n_launch += 1;
  if (fk_precision == 32)
    return launch_k1<float>(rb, gd, pairs, n_pairs, d_q, d_q_end, num_points, S, flags, d_mask, st);
  return launch_k1<double>(rb, gd, pairs, n_pairs, d_q, d_q_end, num_points, S, flags, d_mask, st);
}

cudaError_t launch_get_collisions(const RobotDev &rb, const GridDev &gd, const SelfPair *pairs, int n_pairs,
                                  const double *d_q, uint8_t *d_sphere_hit, uint8_t *d_pair_hit, cudaStream_t st) {
  const size_t smem = k1_smem_bytes(rb, SQMP_CHECK_SELF, K1_THREADS);
  auto kern = k1_check_states<double, false, 1>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  kern<<<1, K1_THREADS, smem, st>>>(rb, gd, pairs, n_pairs, d_q, nullptr, 0, 1, SQMP_CHECK_MAP | SQMP_CHECK_SELF, nullptr,
                                    d_sphere_hit, d_pair_hit);
  return cudaGetLastError();
}

// first colliding state of every edge, from the packed mask (isEdgeValid's early exit, BS:6866-6874)
__global__ void k_first_invalid(const uint32_t *__restrict__ mask, int64_t E, int P1, int32_t *__restrict__ first) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t b0 = e * P1, b1 = b0 + P1;
  int32_t res = -1;
  for (int64_t w = b0 >> 5; w <= (b1 - 1) >> 5 && res < 0; ++w) {
    uint32_t v = mask[w];
    const int64_t base = w << 5;
    if (base < b0) v &= ~0u << int(b0 - base);
    if (base + 32 > b1) v &= (b1 - base >= 32) ? ~0u : ((1u << int(b1 - base)) - 1u);
    if (v) res = int32_t(base + (__ffs(v) - 1) - b0);
  }
  first[e] = res;
}
cudaError_t launch_first_invalid(const uint32_t *d_mask, int64_t E, int P1, int32_t *d_first, cudaStream_t st) {
  if (E <= 0) return cudaSuccess;
  k_first_invalid<<<unsigned((E + 255) / 256), 256, 0, st>>>(d_mask, E, P1, d_first);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// map ingest: row packing + separable capped distance transform (one-off per sqmp_set_voxel_grid, CC:76-88)
// ------------------------------------------------------------------------------------------------
// linear bit index (k*gy+j)*gx+i  ->  rows of row_words words (x fastest), trailing pad words zero
__global__ void k_pack_rows(const uint32_t *__restrict__ lin, int gx, int gy, int gz, int row_words,
                            uint32_t *__restrict__ occ) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;  // one output word
  const int64_t total = int64_t(gz) * gy * row_words;
  if (idx >= total) return;
  const int w = int(idx % row_words);
  const int64_t row = idx / row_words;
  uint32_t out = 0;
  const int i0 = w * 32;
  if (i0 < gx) {
    const int n = min(32, gx - i0);
    const int64_t b = row * gx + i0;  // first linear bit
    const int64_t wi = b >> 5;
    const int sh = int(b & 31);
    uint64_t v = lin[wi];
    if (sh + n > 32) v |= uint64_t(lin[wi + 1]) << 32;
    out = uint32_t(v >> sh);
    if (n < 32) out &= (1u << n) - 1u;
  }
  occ[idx] = out;
}
cudaError_t launch_pack_rows(const uint32_t *d_bits_linear, int gx, int gy, int gz, int row_words, uint32_t *d_occ,
                             cudaStream_t st) {
  const int64_t total = int64_t(gz) * gy * row_words;
  k_pack_rows<<<unsigned((total + 255) / 256), 256, 0, st>>>(d_bits_linear, gx, gy, gz, row_words, d_occ);
  return cudaGetLastError();
}

// Squared box-to-box distance (in cells) from every cell to the nearest occupied cell, separable:
//   cost of an index offset d along one axis is max(|d|-1, 0)^2  (gap between two unit boxes).
// Window W caps the search; a cell with nothing inside the window gets the (still valid) lower bound W.
__global__ void k_edt_x(const uint32_t *__restrict__ occ, int row_words, int gx, int64_t rows, int W,
                        uint8_t *__restrict__ a) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * gx) return;
  const int i = int(idx % gx);
  const uint32_t *row = occ + (idx / gx) * row_words;
  int best = W;
  for (int d = 0; d <= W && best == W; ++d) {
    const int l = i - d, r = i + d;
    const bool hl = l >= 0 && ((row[l >> 5] >> (l & 31)) & 1u);
    const bool hr = r < gx && ((row[r >> 5] >> (r & 31)) & 1u);
    if (hl || hr) best = max(d - 1, 0);
  }
  a[idx] = uint8_t(best);
}
__global__ void k_edt_y(const uint8_t *__restrict__ a, int gx, int gy, int gz, int W, uint16_t *__restrict__ b) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(gx) * gy * gz) return;
  const int i = int(idx % gx), j = int((idx / gx) % gy);
  const int64_t k = idx / (int64_t(gx) * gy);
  int best = W * W;
  for (int d = -W; d <= W; ++d) {
    const int jj = j + d;
    if (jj < 0 || jj >= gy) continue;
    const int g = max(abs(d) - 1, 0);
    const int av = a[(k * gy + jj) * gx + i];
    best = min(best, av * av + g * g);
  }
  b[idx] = uint16_t(min(best, W * W));
}
__global__ void k_edt_z(const uint16_t *__restrict__ b, int gx, int gy, int gz, int W, uint8_t *__restrict__ edt) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t plane = int64_t(gx) * gy;
  if (idx >= plane * gz) return;
  const int k = int(idx / plane);
  const int64_t ij = idx % plane;
  int best = W * W;
  for (int d = -W; d <= W; ++d) {
    const int kk = k + d;
    if (kk < 0 || kk >= gz) continue;
    const int g = max(abs(d) - 1, 0);
    best = min(best, int(b[kk * plane + ij]) + g * g);
  }
  best = min(best, W * W);
  // floor(4 * sqrt(best)) = isqrt(16 * best), exact in integers
  const int t = 16 * best;
  int r = int(sqrtf(float(t)));
  while (r * r > t) --r;
  while ((r + 1) * (r + 1) <= t) ++r;
  edt[idx] = uint8_t(min(r, 255));
}
cudaError_t launch_edt_build(const uint32_t *d_occ, int row_words, int gx, int gy, int gz, int window,
                             uint8_t *d_tmp_a, uint16_t *d_tmp_b, uint8_t *d_edt, cudaStream_t st, int *n_launch) {
  const int64_t n = int64_t(gx) * gy * gz;
  const unsigned blocks = unsigned((n + 255) / 256);
  k_edt_x<<<blocks, 256, 0, st>>>(d_occ, row_words, gx, int64_t(gy) * gz, window, d_tmp_a);
  k_edt_y<<<blocks, 256, 0, st>>>(d_tmp_a, gx, gy, gz, window, d_tmp_b);
  k_edt_z<<<blocks, 256, 0, st>>>(d_tmp_b, gx, gy, gz, window, d_edt);
  if (n_launch) *n_launch += 3;
  return cudaGetLastError();
}

}  // namespace sqmp
