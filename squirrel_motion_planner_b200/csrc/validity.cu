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
// exact sphere vs occupied-voxel-boxes test
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double axis_gap(double c, double lo, double hi) {
  return fmax(fmax(lo - c, c - hi), 0.0);
}

// true  <=> dist(centre, AABB of some occupied voxel) <= r.   thr8 / ncell: per-sphere constants for this grid.
__device__ __forceinline__ bool sphere_hits_map(const GridDev &g, double cx, double cy, double cz, double r,
                                                int thr8, int ncell) {
  const double ux = (cx - g.ox) * g.inv_res, uy = (cy - g.oy) * g.inv_res, uz = (cz - g.oz) * g.inv_res;
  // reject anything that cannot reach the grid (also keeps the int conversions below in range)
  const double reach = double(ncell + 1);
  if (!(ux > -reach && uy > -reach && uz > -reach && ux < g.gx + reach && uy < g.gy + reach && uz < g.gz + reach))
    return false;
  const int ix = __double2int_rd(ux), iy = __double2int_rd(uy), iz = __double2int_rd(uz);
  if (ix >= 0 && iy >= 0 && iz >= 0 && ix < g.gx && iy < g.gy && iz < g.gz) {
    // box-to-box distance of this cell to the nearest occupied cell bounds dist(c, occupied) from below
    const int e = g.edt[(size_t(iz) * g.gy + iy) * g.gx + ix];
    if (e > thr8) return false;
  }
  const double rc = r * g.inv_res, r2 = r * r;
  const int k0 = max(__double2int_rd(uz - rc), 0), k1 = min(__double2int_rd(uz + rc), g.gz - 1);
  const int j0 = max(__double2int_rd(uy - rc), 0), j1 = min(__double2int_rd(uy + rc), g.gy - 1);
  for (int k = k0; k <= k1; ++k) {
    const double dz = axis_gap(cz, fma(double(k), g.res, g.oz), fma(double(k + 1), g.res, g.oz));
    const double remz = r2 - dz * dz;
    if (remz < 0.0) continue;
    const uint32_t *slab = g.occ + size_t(k) * g.gy * g.row_words;
    for (int j = j0; j <= j1; ++j) {
      const double dy = axis_gap(cy, fma(double(j), g.res, g.oy), fma(double(j + 1), g.res, g.oy));
      const double rem = remz - dy * dy;
      if (rem < 0.0) continue;
      const double wc = sqrt(rem) * g.inv_res;  // half-width of the admissible x interval, in cells
      // voxel i is within reach  <=>  ox + (i+1) res >= cx - w  and  ox + i res <= cx + w
      int i0 = __double2int_ru(ux - wc) - 1, i1 = __double2int_rd(ux + wc);
      i0 = max(i0, 0);
      i1 = min(i1, g.gx - 1);
      if (i0 > i1) continue;
      const uint32_t *row = slab + size_t(j) * g.row_words;
      // [i0, i1] is at most 2*ncell+2 wide: one word, two when it straddles a word boundary
      while (i0 <= i1) {
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

// MODE 0: verdict bit per state.  MODE 1 (diagnostics, one state): per-sphere / per-pair hits, no early exit.
template <typename T, bool EDGES, int MODE>
__global__ void __launch_bounds__(K1_THREADS)
k1_check_states(const __grid_constant__ RobotDev rb, const __grid_constant__ GridDev gd,
                const SelfPair *__restrict__ pairs, int n_pairs, const double *__restrict__ q_in,
                const double *__restrict__ q_end, int num_points, int64_t S, unsigned flags,
                uint32_t *__restrict__ mask_out, uint8_t *__restrict__ sphere_hit, uint8_t *__restrict__ pair_hit) {
  extern __shared__ float s_rel[];  // [3][n_spheres][blockDim] sphere centres relative to the base, SELF only
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = s < S;
  const bool do_map = (flags & SQMP_CHECK_MAP) && gd.has_grid;
  const bool do_self = (flags & SQMP_CHECK_SELF) != 0;
  bool hit = false;

  if (active) {
    // ---- the state: either given, or interpolated from the edge end points (BS:4461-4462, 4478; no FMA) ----
    double q[SQMP_DOF];
    if (EDGES) {
      const int P1 = num_points + 1;
      const int64_t e = s / P1;
      const int kk = int(s - e * P1);
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
      const double4 *p4 = reinterpret_cast<const double4 *>(q_in + s * SQMP_DOF);
      const double4 v0 = p4[0], v1 = p4[1];
      q[0] = v0.x; q[1] = v0.y; q[2] = v0.z; q[3] = v0.w;
      q[4] = v1.x; q[5] = v1.y; q[6] = v1.z; q[7] = v1.w;
    }

    // ---- base reference point (world) = position after the leading prismatic run ----
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

    // ---- chain walk; T = double walks the whole chain in world coordinates, T = float walks the part after
    //      the prismatic base relative to pref (host guarantees the leading run is translation-only) ----
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

    for (int sg = -1; sg < rb.n_seg; ++sg) {
      if (sg >= sg0) frame_advance<T>(f, rb, sg, q);
      else if (sg >= 0) continue;
      const int b = rb.seg_sph_end[sg + 1], e = rb.seg_sph_end[sg + 2];
      for (int i = b; i < e; ++i) {
        const T lx = T(rb.sph[i][0]), ly = T(rb.sph[i][1]), lz = T(rb.sph[i][2]);
        const T wx = fma(f.r[0], lx, fma(f.r[1], ly, fma(f.r[2], lz, f.p[0])));
        const T wy = fma(f.r[3], lx, fma(f.r[4], ly, fma(f.r[5], lz, f.p[1])));
        const T wz = fma(f.r[6], lx, fma(f.r[7], ly, fma(f.r[8], lz, f.p[2])));
        const double cx = double(wx) + off[0], cy = double(wy) + off[1], cz = double(wz) + off[2];
        if (do_self) {
          const int n = rb.n_spheres;
          s_rel[(0 * n + i) * blockDim.x + threadIdx.x] = float(cx - pref[0]);
          s_rel[(1 * n + i) * blockDim.x + threadIdx.x] = float(cy - pref[1]);
          s_rel[(2 * n + i) * blockDim.x + threadIdx.x] = float(cz - pref[2]);
        }
        const bool enabled = (MODE == 1) || ((rb.sph_enabled[i >> 5] >> (i & 31)) & 1u);
        if (do_map && enabled && (MODE == 1 || !hit)) {
          const bool h = sphere_hits_map(gd, cx, cy, cz, rb.sph[i][3], rb.sph_thr8[i], rb.sph_ncell[i]);
          hit |= h;
          if (MODE == 1) sphere_hit[i] = h ? 1 : 0;
        }
      }
      if (MODE == 0 && hit) break;  // verdict known (CC:114-115: map before self, first hit wins)
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
  }

  if (MODE == 0) {
    const unsigned word = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && s < ((S + 31) & ~int64_t(31))) mask_out[s >> 5] = word;
  }
}

template <typename T>
static cudaError_t launch_k1(const RobotDev &rb, const GridDev &gd, const SelfPair *pairs, int n_pairs,
                             const double *d_q, const double *d_q_end, int num_points, int64_t S, unsigned flags,
                             uint32_t *d_mask, cudaStream_t st) {
  const size_t smem = (flags & SQMP_CHECK_SELF) ? size_t(3) * rb.n_spheres * K1_THREADS * sizeof(float) : 0;
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
  const size_t smem = size_t(3) * rb.n_spheres * 32 * sizeof(float);
  auto kern = k1_check_states<double, false, 1>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  kern<<<1, 32, smem, st>>>(rb, gd, pairs, n_pairs, d_q, nullptr, 0, 1, SQMP_CHECK_MAP | SQMP_CHECK_SELF, nullptr,
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
