// validity.cu — K1: fused (interpolate ->) forward kinematics -> sphere-vs-voxel -> (self pairs ->) 1 bit/state.
//
// Reference path replaced (file:line in /root/reference, tags as in include/sqmp_b200.h):
//   CollisionChecker::isInCollision CC:104-121, updateTransforms CC:519-539, checkMapCollision CC:554-592,
//   checkSelfCollision CC:541-552; BiRRTstarPlanner::interpolateConfigurations BS:4443-4490, isEdgeValid BS:6864-6895.
// The reference walks 39 links and calls fcl::collide per link per state, serially, with early exits.  Here one
// thread owns one state: it walks the 8 folded chain segments keeping a single frame in registers, transforms the
// spheres of each rigid group as soon as that group's frame exists, and tests every sphere against
//   (1) a byte field: squared box-to-box distance from the centre's cell to the nearest occupied cell (exact lower
//       bound of the point distance: "certainly free", one gather), then, only if that fails,
//   (2) the offset of that nearest occupied cell (16-bit gather): an exact upper bound ("certainly hit"), then
//   (3) the exact test "some occupied voxel AABB within r of the centre" by scanning bit-packed x-rows:
//       per (j,k) row the admissible x-interval is closed-form, so a row costs one (rarely two) word loads.
// The verdicts of a warp are packed by one ballot into one 32-bit word — the mask layout of the C ABI.
#include "k1_scan.cuh"
#include "k1_self.cuh"

namespace sqmp {

// MODE 0: verdict bit per state.  MODE 1 (diagnostics, one state): per-slot / per-link-pair hits, no early exit.
//
// rb is the GRID-UNIT copy of the robot (translations, sphere centres and radii divided by res, root offset
// measured from the grid origin), so a sphere centre comes out of the chain directly in voxel coordinates; the
// whole chain is walked in fp64 in absolute grid coordinates.  The slots (spheres, cluster by cluster, padded to
// multiples of four) come from a table in global memory, read with warp-uniform indices.
//
// Per sphere, in lockstep for all 32 states of a warp:
//   1. centre = R l + p, cell index by the round-to-nearest magic-add (no F2I);
//   2. one byte gather  lb2 = squared box-to-box distance (cells, integer) from the centre's cell to the nearest
//      occupied cell.  Every point of the cell is at least that far from every occupied cube, so
//      lb2 > floor(rg^2) proves the sphere free;
//   3. otherwise one 16-bit gather = offset of that nearest occupied cell O*: dist(centre, O*) is an upper bound of
//      the true distance, dist(centre, O*) <= rg proves a hit;
//   4. what is left (lb <= rg < dist(centre,O*)) is pushed by ballot compaction on a per-warp queue in shared
//      memory and decided 32 items at a time by the exact row scan (drain_queue).
// Steps 3-4 exist once in the code: each lane serialises its own pending spheres of a batch through them.
// This kernel is the reference the fast kernel (validity_fast.cu) is compared with bit for bit; it carries no bound
// hierarchy and no slab fields, so with a covering sphere model it is several times slower.
template <bool EDGES, int MODE, bool SELF>
__global__ void __launch_bounds__(K1_THREADS, 4)
k1_check_states(const __grid_constant__ RobotDev rb, const __grid_constant__ GridDev gd,
                const __grid_constant__ SelfDev self, const K1Slot *__restrict__ slots, const double *__restrict__ q_in,
                const double *__restrict__ q_end, int num_points, int64_t S, unsigned flags,
                uint32_t *__restrict__ mask_out, uint8_t *__restrict__ slot_hit, uint8_t *__restrict__ link_pair_hit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpQueue *wq = reinterpret_cast<WarpQueue *>(smem_raw) + (threadIdx.x >> 5);
  double *q_mine = reinterpret_cast<double *>(smem_raw + sizeof(WarpQueue) * (K1_THREADS / 32)) + threadIdx.x;  // [8][blockDim]
  // SELF: the frame of every rigid group (rotation + translation relative to the base reference, fp32), own column
  float *f_mine = reinterpret_cast<float *>(smem_raw + sizeof(WarpQueue) * (K1_THREADS / 32) +
                                            sizeof(double) * SQMP_DOF * K1_THREADS) + threadIdx.x;  // [n_frames*12][blockDim]
  const int lane = threadIdx.x & 31;
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = s < S;
  const int64_t sl = active ? s : S - 1;  // idle lanes of the last warp shadow the last state (lockstep loops)
  const bool do_map = (flags & SQMP_CHECK_MAP) && gd.has_grid;
  const bool do_self = SELF;   // (flags & SQMP_CHECK_SELF), resolved at launch: the map-only kernel carries no self state
  int qcount = 0;  // warp-uniform

  // ---- the state (given, or interpolated from the edge end points), parked in shared memory (own column): the chain
  //      walk picks q[seg_qidx] with one LDS ----
  double q[SQMP_DOF];
  k1_load_state<EDGES>(q_in, q_end, num_points, sl, q);
  const bool sane = k1_park_state(rb, q, q_mine, K1_THREADS);
  // a non-finite or absurd joint value is "in collision" (fail closed, like an uninitialised checker, CC:106-107)
  bool hit = !sane;

  // ---- base reference point = position after the leading prismatic run ----
  double pref[3] = {rb.root_off[0], rb.root_off[1], rb.root_off[2]};
  if (do_self) {
    for (int sg = 0; sg <= rb.ref_seg; ++sg) {
      const double qv = q_mine[rb.seg_qidx[sg] * K1_THREADS];
#pragma unroll
      for (int i = 0; i < 3; ++i) pref[i] += rb.seg_pre[sg][9 + i] + rb.seg_axis[sg][i] * qv;
    }
  }

  Frame<double> f;
  frame_set_translation<double>(f, rb.root_off[0], rb.root_off[1], rb.root_off[2]);

  const unsigned plane = gd.splane, sx = gd.sx;   // strides of the halo-padded fields (host: padded cells < 2^32)
  bool live = do_map && sane;   // this lane still needs map tests (MODE 0: until its first hit)
  for (int sg = -1; sg < rb.n_seg; ++sg) {
    if (sg >= 0) frame_advance_smem<double>(f, rb, sg, q_mine, K1_THREADS);
    const int b = rb.grp_slot_end[sg + 1], e = rb.grp_slot_end[sg + 2];
    if (do_self) {
      const int fi = rb.grp_fidx[sg + 1];
      if (fi >= 0) {
        float *dst = f_mine + fi * 12 * K1_THREADS;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k * K1_THREADS] = float(f.r[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[(9 + k) * K1_THREADS] = float(f.p[k] - pref[k]);
      }
    }
    if (b == e || !do_map) continue;
    // translation of this group minus half a cell: round-to-nearest of (w - 0.5) is the cell index floor(w)
    const double pm0 = f.p[0] - 0.5, pm1 = f.p[1] - 0.5, pm2 = f.p[2] - 0.5;
    for (int i0 = b; i0 < e; i0 += K1_BATCH) {
      double wm[K1_BATCH][3];   // centre - 0.5
      int ci[K1_BATCH][3];      // cell of the centre
      int lbv[K1_BATCH], thr[K1_BATCH];
      double rgv[K1_BATCH];
#pragma unroll
      for (int u = 0; u < K1_BATCH; ++u) {  // transform + byte gather (independent loads, issued back to back)
        const K1Slot sp = slots[i0 + u];    // clusters are padded to a multiple of K1_BATCH (pads never test positive)
        wm[u][0] = fma(f.r[0], sp.x, fma(f.r[1], sp.y, fma(f.r[2], sp.z, pm0)));
        wm[u][1] = fma(f.r[3], sp.x, fma(f.r[4], sp.y, fma(f.r[5], sp.z, pm1)));
        wm[u][2] = fma(f.r[6], sp.x, fma(f.r[7], sp.y, fma(f.r[8], sp.z, pm2)));
        ci[u][0] = rn_int(wm[u][0]); ci[u][1] = rn_int(wm[u][1]); ci[u][2] = rn_int(wm[u][2]);
        const int ix = ci[u][0], iy = ci[u][1], iz = ci[u][2];
        const bool inside = unsigned(ix) < unsigned(gd.gx) && unsigned(iy) < unsigned(gd.gy) && unsigned(iz) < unsigned(gd.gz);
        // squared box-to-box distance of this cell to the nearest occupied cell: a lower bound of dist(centre, occupied)^2
        lbv[u] = inside ? int(gd.lb2[unsigned(iz) * plane + unsigned(iy) * sx + unsigned(ix)]) : -1;
        rgv[u] = sp.r;
        // thr = floor(rg^2), or -2 for a pad slot / a sphere whose link is disabled (lbv is -1 outside the grid)
        const bool off = sp.link < 0 || (MODE == 0 && ((rb.link_disabled >> (sp.link & 63)) & 1ull));
        thr[u] = off ? -2 : (MODE == 1 ? 1000 : sp.thr);
      }
      unsigned pend = 0;
#pragma unroll
      for (int u = 0; u < K1_BATCH; ++u)
        if (lbv[u] <= thr[u]) pend |= 1u << u;
      if (!live) pend = 0;
      while (__any_sync(FULL, pend != 0)) {
        // every lane with pending spheres takes its first one through the single copy of steps 3-4
        const int u = __ffs(pend) - 1;
        pend &= pend - 1;
        double cx = wm[0][0], cy = wm[0][1], cz = wm[0][2], rg = rgv[0];
        int jx = ci[0][0], jy = ci[0][1], jz = ci[0][2], lb = lbv[0];
#pragma unroll
        for (int v = 1; v < K1_BATCH; ++v) {
          const bool sel = u == v;
          cx = sel ? wm[v][0] : cx; cy = sel ? wm[v][1] : cy; cz = sel ? wm[v][2] : cz;
          jx = sel ? ci[v][0] : jx; jy = sel ? ci[v][1] : jy; jz = sel ? ci[v][2] : jz;
          lb = sel ? lbv[v] : lb;
          rg = sel ? rgv[v] : rg;
        }
        const int i = i0 + max(u, 0);
        bool amb = false;   // neither bound decides: exact scan
        if (u >= 0 && live) {
          if (MODE == 1) {
            slot_hit[i] = exact_scan(gd, cx + 0.5, cy + 0.5, cz + 0.5, rg) ? 1 : 0;
          } else if (lb >= 0) {
            const unsigned v = gd.ft[unsigned(jz) * plane + unsigned(jy) * sx + unsigned(jx)];
            const float fx = float(cx - double(jx)) + 0.5f, fy = float(cy - double(jy)) + 0.5f, fz = float(cz - double(jz)) + 0.5f;
            const float gx_ = axis_gap_off(int(v & 31u) - 15, fx), gy_ = axis_gap_off(int((v >> 5) & 31u) - 15, fy),
                        gz_ = axis_gap_off(int((v >> 10) & 31u) - 15, fz);
            const float ub2 = fmaf(gx_, gx_, fmaf(gy_, gy_, gz_ * gz_));
            const float r2 = float(rg) * float(rg);
            // float slack 1e-5 rg^2 (~2.5e-7 m at contact, far inside the 1e-5 m band): closer calls go to the scan
            if (ub2 <= r2 * 0.99999f) { hit = true; live = false; pend = 0; }
            else amb = true;
          } else {
            const int pad = int(ceil(rg)) + 1;   // centre outside the grid: can the sphere still reach it?
            amb = unsigned(jx + pad) < unsigned(gd.gx + 2 * pad) && unsigned(jy + pad) < unsigned(gd.gy + 2 * pad) &&
                  unsigned(jz + pad) < unsigned(gd.gz + 2 * pad);
          }
        }
        if (MODE == 1) continue;
        const unsigned m = __ballot_sync(FULL, amb);
        if (m) {
          if (amb) {
            const int pos = qcount + __popc(m & ((1u << lane) - 1u));
            wq->x[pos] = cx + 0.5;
            wq->y[pos] = cy + 0.5;
            wq->z[pos] = cz + 0.5;
            wq->r[pos] = rg;
            wq->info[pos] = (i << 8) | lane;
          }
          qcount += __popc(m);
          if (qcount >= 32) {
            __syncwarp();
            const unsigned hm = drain_queue(gd, wq, qcount - 32, 32, __ballot_sync(FULL, hit));
            __syncwarp();
            qcount -= 32;
            if ((hm >> lane) & 1u) { hit = true; live = false; pend = 0; }
          }
        }
      }
    }
    // whole warp decided (CC:114-115: map before self, first hit wins) -> stop walking the chain
    if (MODE == 0 && !do_self && qcount == 0 && __all_sync(FULL, hit)) break;
  }
  if (MODE == 0 && qcount > 0) {
    __syncwarp();
    const unsigned hm = drain_queue(gd, wq, 0, qcount, __ballot_sync(FULL, hit));
    hit |= (hm >> lane) & 1u;
    qcount = 0;
  }

  // ---- self collision over the enabled link pairs on different rigid groups (CC:541-552), nested bounds (SelfDev) ----
  if (do_self && (MODE == 1 || !hit)) {
    if (rb.self_const_collision) hit = true;
    if (MODE == 1) {   // diagnostics: every colliding link pair; the threads of the block share the group pairs
      self_collision<true>(self, f_mine, K1_THREADS, link_pair_hit, threadIdx.x, K1_THREADS);
    } else if (!hit) {
      hit = self_collision<false>(self, f_mine, K1_THREADS, nullptr);
    }
  }

  if (MODE == 0) {
    const unsigned word = __ballot_sync(FULL, hit && active);
    if (lane == 0 && s < ((S + 31) & ~int64_t(31))) mask_out[s >> 5] = word;
  }
}

static size_t k1_smem_bytes(const RobotDev &rb, unsigned flags, int threads) {
  size_t b = sizeof(WarpQueue) * (K1_THREADS / 32) + sizeof(double) * SQMP_DOF * K1_THREADS;
  if (flags & SQMP_CHECK_SELF) b += size_t(12) * rb.n_frames * threads * sizeof(float);
  return b;
}

cudaError_t launch_check_states(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const K1Slot *d_slots,
                                const double *d_q, const double *d_q_end, int num_points, int64_t S, unsigned flags,
                                uint32_t *d_mask, cudaStream_t st, int *n_launch) {
  if (n_launch) *n_launch += 1;
  const size_t smem = k1_smem_bytes(rb, flags, K1_THREADS);
  const int64_t blocks = (S + K1_THREADS - 1) / K1_THREADS;
  if (blocks <= 0) return cudaSuccess;
  if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    kern<<<unsigned(blocks), K1_THREADS, smem, st>>>(rb, gd, self, d_slots, d_q, d_q_end, d_q_end ? num_points : 0, S, flags,
                                                      d_mask, nullptr, nullptr);
  };
  const bool with_self = (flags & SQMP_CHECK_SELF) != 0;
  if (d_q_end) {
    if (with_self) go(k1_check_states<true, 0, true>);
    else go(k1_check_states<true, 0, false>);
  } else {
    if (with_self) go(k1_check_states<false, 0, true>);
    else go(k1_check_states<false, 0, false>);
  }
  return cudaGetLastError();
}

cudaError_t launch_get_collisions(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const K1Slot *d_slots,
                                  const double *d_q, uint8_t *d_slot_hit, uint8_t *d_link_pair_hit, cudaStream_t st) {
  const size_t smem = k1_smem_bytes(rb, SQMP_CHECK_SELF, K1_THREADS);
  auto kern = k1_check_states<false, 1, true>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  kern<<<1, K1_THREADS, smem, st>>>(rb, gd, self, d_slots, d_q, nullptr, 0, 1, SQMP_CHECK_MAP | SQMP_CHECK_SELF, nullptr,
                                    d_slot_hit, d_link_pair_hit);
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

// Feature transform: for every cell the occupied cell O* minimising the squared box-to-box distance
//   bb2 = sum_a max(|o_a|-1, 0)^2   (gap between two unit boxes per axis), ties: smaller centre distance sum_a o_a^2,
// found exactly by three separable passes: along one axis both keys are monotone in |o_a|, so the nearest occupied
// cell of a row (pass x) is the only candidate that row can contribute to any cell of its column (pass y), and the
// best (ox,oy) of a plane is the only candidate the plane can contribute along z.  Window W caps the search.
constexpr int FT_NONE8 = -128;
__device__ __forceinline__ int bbgap2(int d) { const int g = max(abs(d) - 1, 0); return g * g; }

// All three passes run over the HALO-PADDED domain (gx+2H) x (gy+2H) x (gz+2H), occupancy zero outside the grid, so
// that lb2 / ft are defined for every cell a sphere centre can fall into while the base is inside the grid.
__global__ void k_ft_x(const uint32_t *__restrict__ occ, int row_words, int gx, int gy, int gz, int H, int W,
                       int8_t *__restrict__ a) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int PX = gx + 2 * H, PY = gy + 2 * H, PZ = gz + 2 * H;
  if (idx >= int64_t(PX) * PY * PZ) return;
  const int i = int(idx % PX) - H, j = int((idx / PX) % PY) - H, k = int(idx / (int64_t(PX) * PY)) - H;
  int best = FT_NONE8;
  if (unsigned(j) < unsigned(gy) && unsigned(k) < unsigned(gz)) {
    const uint32_t *row = occ + (int64_t(k) * gy + j) * row_words;
    for (int d = 0; d <= W && best == FT_NONE8; ++d) {
      const int l = i - d, r = i + d;
      if (unsigned(r) < unsigned(gx) && ((row[r >> 5] >> (r & 31)) & 1u)) best = d;
      else if (unsigned(l) < unsigned(gx) && ((row[l >> 5] >> (l & 31)) & 1u)) best = -d;
    }
  }
  a[idx] = int8_t(best);
}
__global__ void k_ft_y(const int8_t *__restrict__ a, int PX, int PY, int PZ, int W, uint16_t *__restrict__ b) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(PX) * PY * PZ) return;
  const int i = int(idx % PX), j = int((idx / PX) % PY);
  const int64_t k = idx / (int64_t(PX) * PY);
  int best = 0x7fffffff, bx = 0, by = 0;
  for (int d = -W; d <= W; ++d) {
    const int jj = j + d;
    if (jj < 0 || jj >= PY) continue;
    const int ox = a[(k * PY + jj) * PX + i];
    if (ox == FT_NONE8) continue;
    const int key = (bbgap2(ox) + bbgap2(d)) * 4096 + ox * ox + d * d;
    if (key < best) { best = key; bx = ox; by = d; }
  }
  b[idx] = best == 0x7fffffff ? uint16_t(0xFFFF) : uint16_t((bx + 128) | ((by + 128) << 8));
}
__global__ void k_ft_z(const uint16_t *__restrict__ b, int PX, int PY, int PZ, int W, uint8_t *__restrict__ lb2,
                       uint16_t *__restrict__ ft) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t plane = int64_t(PX) * PY;
  if (idx >= plane * PZ) return;
  const int k = int(idx / plane);
  const int64_t ij = idx % plane;
  int best = 0x7fffffff, bx = 0, by = 0, bz = 0;
  for (int d = -W; d <= W; ++d) {
    const int kk = k + d;
    if (kk < 0 || kk >= PZ) continue;
    const unsigned v = b[kk * plane + ij];
    if (v == 0xFFFFu) continue;
    const int ox = int(v & 255u) - 128, oy = int(v >> 8) - 128;
    const int key = (bbgap2(ox) + bbgap2(oy) + bbgap2(d)) * 4096 + ox * ox + oy * oy + d * d;
    if (key < best) { best = key; bx = ox; by = oy; bz = d; }
  }
  if (best == 0x7fffffff) {
    lb2[idx] = 255;
    ft[idx] = 0xFFFF;
  } else {
    lb2[idx] = uint8_t(min(best >> 12, 255));
    ft[idx] = uint16_t((bx + 15) | ((by + 15) << 5) | ((bz + 15) << 10));
  }
}
// ------------------------------------------------------------------------------------------------
// Slab fields (fast kernel, yaw-only base group): for slab s = layers [k0, k1] and column (i, j) of the halo-padded
// plane, the squared 2-D box-to-box distance (cells, capped at 255) to the nearest column that holds an occupied cell
// in those layers.  A sphere whose z never changes can only touch the layers of its slab, so "slab field > thr" proves
// it free no matter what lies under or above those layers (the floor slab right under the robot, above all).
// ------------------------------------------------------------------------------------------------
struct SlabRanges {
  int k0[64], k1[64];
};
__global__ void k_slab_or(const uint32_t *__restrict__ occ, int row_words, int gy, int n_slabs, SlabRanges rg,
                          uint32_t *__restrict__ col) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;   // one word of one slab's column bitmap
  const int64_t per = int64_t(gy) * row_words;
  if (idx >= per * n_slabs) return;
  const int s = int(idx / per);
  const int64_t jw = idx % per;
  uint32_t v = 0;
  for (int k = rg.k0[s]; k <= rg.k1[s]; ++k) v |= occ[int64_t(k) * per + jw];
  col[idx] = v;
}
__global__ void k_slab_x(const uint32_t *__restrict__ col, int row_words, int gx, int gy, int H, int W, int n_slabs,
                         int8_t *__restrict__ a) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int PX = gx + 2 * H, PY = gy + 2 * H;
  if (idx >= int64_t(PX) * PY * n_slabs) return;
  const int i = int(idx % PX) - H, j = int((idx / PX) % PY) - H, s = int(idx / (int64_t(PX) * PY));
  int best = FT_NONE8;
  if (unsigned(j) < unsigned(gy)) {
    const uint32_t *row = col + (int64_t(s) * gy + j) * row_words;
    for (int d = 0; d <= W && best == FT_NONE8; ++d) {
      const int l = i - d, r = i + d;
      if ((unsigned(r) < unsigned(gx) && ((row[r >> 5] >> (r & 31)) & 1u)) ||
          (unsigned(l) < unsigned(gx) && ((row[l >> 5] >> (l & 31)) & 1u)))
        best = d;
    }
  }
  a[idx] = int8_t(best);
}
__global__ void k_slab_y(const int8_t *__restrict__ a, int PX, int PY, int W, int n_slabs, uint8_t *__restrict__ out) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(PX) * PY * n_slabs) return;
  const int i = int(idx % PX), j = int((idx / PX) % PY);
  const int64_t s = idx / (int64_t(PX) * PY);
  int best = 255;
  for (int d = -W; d <= W; ++d) {
    const int jj = j + d;
    if (jj < 0 || jj >= PY) continue;
    const int ox = a[(s * PY + jj) * PX + i];
    if (ox == FT_NONE8) continue;
    best = min(best, bbgap2(ox) + bbgap2(d));
  }
  out[idx] = uint8_t(best);
}
// d_col: n_slabs * gy * row_words words; d_tmp: >= n_slabs * PX * PY bytes; d_slab: n_slabs * PX * PY bytes
cudaError_t launch_slab_build(const uint32_t *d_occ, int row_words, int gx, int gy, int gz, int halo, int window,
                              int n_slabs, const int *h_k0, const int *h_k1, uint32_t *d_col, int8_t *d_tmp, uint8_t *d_slab,
                              cudaStream_t st, int *n_launch) {
  if (n_slabs <= 0) return cudaSuccess;
  if (n_slabs > 64) return cudaErrorInvalidValue;
  SlabRanges rg;
  for (int s = 0; s < 64; ++s) {
    rg.k0[s] = s < n_slabs ? h_k0[s] : 0;
    rg.k1[s] = s < n_slabs ? h_k1[s] : -1;
    if (rg.k0[s] < 0) rg.k0[s] = 0;
    if (rg.k1[s] > gz - 1) rg.k1[s] = gz - 1;
  }
  const int PX = gx + 2 * halo, PY = gy + 2 * halo;
  const int64_t nw = int64_t(n_slabs) * gy * row_words, n = int64_t(n_slabs) * PX * PY;
  k_slab_or<<<unsigned((nw + 255) / 256), 256, 0, st>>>(d_occ, row_words, gy, n_slabs, rg, d_col);
  k_slab_x<<<unsigned((n + 255) / 256), 256, 0, st>>>(d_col, row_words, gx, gy, halo, window, n_slabs, d_tmp);
  k_slab_y<<<unsigned((n + 255) / 256), 256, 0, st>>>(d_tmp, PX, PY, window, n_slabs, d_slab);
  if (n_launch) *n_launch += 3;
  return cudaGetLastError();
}

// d_tmp_a / d_tmp_b / d_lb2 / d_ft all hold (gx+2H)(gy+2H)(gz+2H) cells
cudaError_t launch_ft_build(const uint32_t *d_occ, int row_words, int gx, int gy, int gz, int halo, int window,
                            int8_t *d_tmp_a, uint16_t *d_tmp_b, uint8_t *d_lb2, uint16_t *d_ft, cudaStream_t st,
                            int *n_launch) {
  const int PX = gx + 2 * halo, PY = gy + 2 * halo, PZ = gz + 2 * halo;
  const int64_t n = int64_t(PX) * PY * PZ;
  const unsigned blocks = unsigned((n + 255) / 256);
  k_ft_x<<<blocks, 256, 0, st>>>(d_occ, row_words, gx, gy, gz, halo, window, d_tmp_a);
  k_ft_y<<<blocks, 256, 0, st>>>(d_tmp_a, PX, PY, PZ, window, d_tmp_b);
  k_ft_z<<<blocks, 256, 0, st>>>(d_tmp_b, PX, PY, PZ, window, d_lb2, d_ft);
  if (n_launch) *n_launch += 3;
  return cudaGetLastError();
}

}  // namespace sqmp
