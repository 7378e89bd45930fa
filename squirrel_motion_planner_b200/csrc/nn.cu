// nn.cu — device tree mirror and K3/K4: exact brute-force nearest / radius + top-k-by-cost search.
//
// Reference path replaced (file:line in /root/reference, tags as in include/sqmp_b200.h):
//   DistanceHeuristics::euclidean_joint_space_distance DH:128-156 (fp64, sum in index order, sqrt),
//   BiRRTstarPlanner::find_nearest_neighbour_interpolation BS:4076-4133 (strict '<', serial order, min0 = 10000, id0 = 0),
//   find_near_vertices_interpolation BS:4272-4324 (d < r, id != x_new.id, sort by cost_reach.total DS:41-44),
//   consumers BS:4618-4623 (first max_near_nodes) and BS:5075-5089 (last max_near_nodes, walked backwards),
//   tree types DS:11-54 and mutation sites BS:3298-3322, 5173-5193, 5518-5520.
// The reference keeps an AoS of std::vector-laden Nodes and deep-copies Nodes to sort them.  Here the tree is a
// device-resident SoA (q[d][n] + cost[n]); distances are evaluated with explicit round-to-nearest mul/add
// intrinsics (never contracted to FMA) in the reference's summation order, so sums of squares are bit-identical
// to the CPU's, sqrt is IEEE (sqrt.rn.f64), and every comparison is made on the same values the reference
// compares.  Ties resolve to the lowest node id, which is what the reference's serial loop does.
//
// Three mappings, all exact:
//   QT ("query per thread", many queries): nodes stream through shared memory in tiles and are broadcast to
//      all lanes; each lane keeps its query and running best in registers; no cross-lane traffic at all.
//   WQ ("warp per query", few queries — the planner's own call shape): lanes stride over the nodes of a chunk,
//      the warp reduces (distance, id) lexicographically with shuffles, radius hits are compacted with a ballot.
//   TC ("tensor-core pre-filter", >= 2048 queries, the default there): the pairwise estimates |n|^2 - 2 q.n of a 256-query
//      tile against the streamed nodes come out of tcgen05.mma over fp16-split operands (tc_umma.cuh); whatever the
//      estimate cannot reject is evaluated exactly as in the other mappings.  See k_nn_tc below.
// All write one partial result per (query, node chunk); a merge kernel folds the chunks in id order.
#include <math_constants.h>

#include "sqmp_common.cuh"
#include "tc_umma.cuh"

namespace sqmp {

// ---------------------------------------------------------------------------------------------------------
// tree mirror
// ---------------------------------------------------------------------------------------------------------
struct TreePtrs {
  double *q[SQMP_DOF];
  double *cost;
  double *aos;
  float4 *fa, *fb;
  float *fn;
  float4 *img;
};
static inline TreePtrs ptrs_of(const TreeDev &t) {
  TreePtrs p;
  for (int d = 0; d < SQMP_DOF; ++d) p.q[d] = t.q[d];
  p.cost = t.cost;
  p.aos = t.aos;
  p.fa = t.fa; p.fb = t.fb; p.fn = t.fn;
  p.img = t.img;
  return p;
}
// fp32 shadow of one node (see TreeDev)
__device__ __forceinline__ void write_shadow(const TreePtrs &t, int64_t i, const double (&v)[SQMP_DOF]) {
  float f[SQMP_DOF];
  double n2 = 0.0;
#pragma unroll
  for (int d = 0; d < SQMP_DOF; ++d) {
    f[d] = __double2float_rn(v[d]);
    n2 += double(f[d]) * double(f[d]);
  }
  t.fa[i] = make_float4(f[0], f[1], f[2], f[3]);
  t.fb[i] = make_float4(f[4], f[5], f[6], f[7]);
  t.fn[i] = __double2float_rn(n2);
  tc_write_node_row(t.img + (i >> 7) * TC_TILE_FLOAT4, int(i & 127), v);   // fp16 operand image (tc_umma.cuh)
}

// rows (AoS, n x 8) -> SoA at [first, first+n)
__global__ void k_tree_append(TreePtrs t, const double *__restrict__ rows, const double *__restrict__ cost,
                              int64_t first, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 *r4 = reinterpret_cast<const double4 *>(rows + i * SQMP_DOF);
  const double4 a = r4[0], b = r4[1];
  t.q[0][first + i] = a.x; t.q[1][first + i] = a.y; t.q[2][first + i] = a.z; t.q[3][first + i] = a.w;
  t.q[4][first + i] = b.x; t.q[5][first + i] = b.y; t.q[6][first + i] = b.z; t.q[7][first + i] = b.w;
  t.cost[first + i] = cost[i];
  double4 *row = reinterpret_cast<double4 *>(t.aos + (first + i) * SQMP_DOF);
  row[0] = a;
  row[1] = b;
  const double v[SQMP_DOF] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  write_shadow(t, first + i, v);
}
struct OneNode {
  double q[SQMP_DOF];
  double cost;
};
__global__ void k_tree_set_one(TreePtrs t, int64_t id, OneNode v) {
  if (threadIdx.x < SQMP_DOF) {
    t.q[threadIdx.x][id] = v.q[threadIdx.x];
    t.aos[id * SQMP_DOF + threadIdx.x] = v.q[threadIdx.x];
  }
  if (threadIdx.x == SQMP_DOF) t.cost[id] = v.cost;
  if (threadIdx.x == SQMP_DOF + 1) write_shadow(t, id, v.q);
}
__global__ void k_tree_update_costs(TreePtrs t, const int32_t *__restrict__ ids, const double *__restrict__ cost,
                                    int64_t n, int64_t size) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && ids[i] >= 0 && ids[i] < size) t.cost[ids[i]] = cost[i];
}
// Deferred single-node writes (sqmp_tree_append / sqmp_tree_update): capi.cu queues them in mapped pinned memory and one
// launch writes the whole queue before the next call that reads a tree.  A record carries its tree's pointers.
struct TreeWrite {
  TreePtrs p;
  double q[SQMP_DOF];
  double cost;
  int64_t id;
};
__global__ void k_tree_write_many(const TreeWrite *__restrict__ w, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TreeWrite e = w[i];   // (mapped host memory: read once)
#pragma unroll
  for (int d = 0; d < SQMP_DOF; ++d) {
    e.p.q[d][e.id] = e.q[d];
    e.p.aos[e.id * SQMP_DOF + d] = e.q[d];
  }
  e.p.cost[e.id] = e.cost;
  write_shadow(e.p, e.id, e.q);
}
size_t tree_write_bytes() { return sizeof(TreeWrite); }
void tree_write_fill(void *dst, const TreeDev &t, int64_t id, const double q[SQMP_DOF], double cost) {
  TreeWrite *w = reinterpret_cast<TreeWrite *>(dst);
  w->p = ptrs_of(t);
  for (int d = 0; d < SQMP_DOF; ++d) w->q[d] = q[d];
  w->cost = cost;
  w->id = id;
}
bool tree_write_same_target(const void *rec, const TreeDev &t, int64_t id) {
  const TreeWrite *w = reinterpret_cast<const TreeWrite *>(rec);
  return w->id == id && w->p.cost == t.cost;
}
cudaError_t launch_tree_write_many(const void *d_records, int n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_tree_write_many<<<unsigned((n + 63) / 64), 64, 0, st>>>(reinterpret_cast<const TreeWrite *>(d_records), n);
  return cudaGetLastError();
}
cudaError_t launch_tree_append(TreeDev &t, const double *d_q_rows, const double *d_cost, int64_t first, int64_t n,
                               cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_tree_append<<<unsigned((n + 255) / 256), 256, 0, st>>>(ptrs_of(t), d_q_rows, d_cost, first, n);
  return cudaGetLastError();
}
cudaError_t launch_tree_set_one(TreeDev &t, int64_t id, const double q[SQMP_DOF], double cost, cudaStream_t st) {
  OneNode v;
  for (int d = 0; d < SQMP_DOF; ++d) v.q[d] = q[d];
  v.cost = cost;
  k_tree_set_one<<<1, 32, 0, st>>>(ptrs_of(t), id, v);
  return cudaGetLastError();
}
cudaError_t launch_tree_update_costs(TreeDev &t, const int32_t *d_ids, const double *d_cost, int64_t n,
                                     cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_tree_update_costs<<<unsigned((n + 255) / 256), 256, 0, st>>>(ptrs_of(t), d_ids, d_cost, n, t.n);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// distance: sum_error_sq += (d*d) for i = 0..7, d = node[i] - query[i]   (DH:145-149; argument order BS:4119)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sumsq8(const double (&node)[SQMP_DOF], const double (&query)[SQMP_DOF]) {
  // conf_start = node, conf_end = query  =>  distance_vec[i] = query[i] - node[i]
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < SQMP_DOF; ++i) {
    const double d = __dsub_rn(query[i], node[i]);
    s = __dadd_rn(s, __dmul_rn(d, d));
  }
  return s;
}

constexpr double NN_INIT_DIST = 10000.0;      // BS:4090
constexpr double NN_INIT_SUMSQ = 100000000.0;  // 10000^2, exactly representable; sqrt(1e8) == 1e4 exactly

struct Best {
  double s;   // sum of squares of the incumbent (a node can only win if its sum is strictly smaller)
  double d;   // sqrt of it: the value the reference compares
  int32_t id;
};
// candidate (s, id) against the incumbent; ids must be offered in increasing order per thread
__device__ __forceinline__ void offer(Best &b, double s, int32_t id) {
  if (s < b.s) {                 // necessary for sqrt(s) < b.d (sqrt is monotone); rare after warm-up
    const double d = sqrt(s);    // IEEE round-to-nearest
    if (d < b.d) { b.s = s; b.d = d; b.id = id; }   // BS:4122 strict '<'
  }
}
// lexicographic (d, id) minimum: equal distances keep the lower id = what the serial reference loop keeps
__device__ __forceinline__ bool better(double d1, int32_t id1, double d2, int32_t id2) {
  return d1 < d2 || (d1 == d2 && id1 < id2);
}

constexpr int NN_THREADS = 128;
#ifndef NN_TILE_DEF
#define NN_TILE_DEF 256
#endif
constexpr int NN_TILE = NN_TILE_DEF;  // nodes per shared-memory tile (QT)

// ---------------------------------------------------------------------------------------------------------
// K3 nearest — QT variant.  grid = (ceil(Q/128), chunks)
// ---------------------------------------------------------------------------------------------------------
template <bool RADIUS>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_qt(TreePtrs t, int64_t N, int64_t chunk_len, const double *__restrict__ queries, int64_t Q,
        double *__restrict__ part_d, int32_t *__restrict__ part_id,
        // radius mode
        const int32_t *__restrict__ exclude, double s_thr, int32_t *__restrict__ count, int32_t *__restrict__ hits,
        int64_t hit_cap) {
  __shared__ double tile[SQMP_DOF][NN_TILE];
  const int64_t qi = int64_t(blockIdx.x) * NN_THREADS + threadIdx.x;
  const bool active = qi < Q;
  double query[SQMP_DOF];
  if (active) {
    const double4 *p4 = reinterpret_cast<const double4 *>(queries + qi * SQMP_DOF);
    const double4 a = p4[0], b = p4[1];
    query[0] = a.x; query[1] = a.y; query[2] = a.z; query[3] = a.w;
    query[4] = b.x; query[5] = b.y; query[6] = b.z; query[7] = b.w;
  } else {
#pragma unroll
    for (int d = 0; d < SQMP_DOF; ++d) query[d] = 0.0;
  }
  const int32_t excl = (RADIUS && active && exclude) ? exclude[qi] : -1;
  Best best = {NN_INIT_SUMSQ, NN_INIT_DIST, 0};
  const int64_t n0 = int64_t(blockIdx.y) * chunk_len, n1 = min(N, n0 + chunk_len);

  for (int64_t base = n0; base < n1; base += NN_TILE) {
    const int len = int(imin64(NN_TILE, n1 - base));
    __syncthreads();
    for (int i = threadIdx.x; i < NN_TILE; i += NN_THREADS) {
#pragma unroll
      for (int d = 0; d < SQMP_DOF; ++d) tile[d][i] = (i < len) ? t.q[d][base + i] : CUDART_INF;
    }
    __syncthreads();
    if (!active) continue;
#pragma unroll 4
    for (int i = 0; i < len; ++i) {
      double node[SQMP_DOF];
#pragma unroll
      for (int d = 0; d < SQMP_DOF; ++d) node[d] = tile[d][i];  // broadcast read
      const double s = sumsq8(node, query);
      const int32_t id = int32_t(base + i);
      if (RADIUS) {
        if (s < s_thr && id != excl) {   // sqrt(s) < r  <=>  s < s_thr (s_thr from the host, exact)
          const int32_t slot = atomicAdd(&count[qi], 1);
          if (slot < hit_cap) hits[qi * hit_cap + slot] = id;
        }
      } else {
        offer(best, s, id);
      }
    }
  }
  if (!RADIUS && active) {
    part_d[int64_t(blockIdx.y) * Q + qi] = best.d;
    part_id[int64_t(blockIdx.y) * Q + qi] = best.id;
  }
}

// ---------------------------------------------------------------------------------------------------------
// K3/K4 — WQ variant.  grid = (Q, ceil(chunks / warps_per_cta)); every warp owns one node chunk of one query
// ---------------------------------------------------------------------------------------------------------
template <bool RADIUS>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_wq(TreePtrs t, int64_t N, int64_t chunk_len, int chunks, const double *__restrict__ queries, int64_t Q,
        double *__restrict__ part_d, int32_t *__restrict__ part_id, const int32_t *__restrict__ exclude,
        double s_thr, int32_t *__restrict__ count, int32_t *__restrict__ hits, int64_t hit_cap) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t qi = blockIdx.x;
  const int chunk = blockIdx.y * (NN_THREADS / 32) + warp;
  if (chunk >= chunks) return;  // whole warp leaves together
  double query[SQMP_DOF];
#pragma unroll
  for (int d = 0; d < SQMP_DOF; ++d) query[d] = __ldg(queries + qi * SQMP_DOF + d);
  const int32_t excl = (RADIUS && exclude) ? exclude[qi] : -1;
  Best best = {NN_INIT_SUMSQ, NN_INIT_DIST, 0};
  const int64_t n0 = int64_t(chunk) * chunk_len, n1 = min(N, n0 + chunk_len);
  for (int64_t base = n0; base < n1; base += 32) {
    const int64_t n = base + lane;
    bool hit = false;
    if (n < n1) {
      double node[SQMP_DOF];
#pragma unroll
      for (int d = 0; d < SQMP_DOF; ++d) node[d] = t.q[d][n];  // coalesced: 32 consecutive doubles per dim
      const double s = sumsq8(node, query);
      if (RADIUS) hit = s < s_thr && int32_t(n) != excl;
      else offer(best, s, int32_t(n));
    }
    if (RADIUS) {
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        int32_t slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&count[qi], __popc(m));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        const int32_t slot = slot0 + __popc(m & ((1u << lane) - 1u));  // ballot compaction
        if (hit && slot < hit_cap) hits[qi * hit_cap + slot] = int32_t(n);
      }
    }
  }
  if (!RADIUS) {
    // warp argmin over (d, id), lexicographic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, best.d, o);
      const int32_t oid = __shfl_xor_sync(0xffffffffu, best.id, o);
      // a lane that never improved on the initial value carries (10000.0, 0): it must not win on id
      // against a real candidate with the same distance -- d == 10000.0 can only be the initial value
      // because real candidates need d < 10000.0 (strict) to be taken at all.
      if (better(od, oid, best.d, best.id)) { best.d = od; best.id = oid; }
    }
    if (lane == 0) {
      part_d[int64_t(chunk) * Q + qi] = best.d;
      part_id[int64_t(chunk) * Q + qi] = best.id;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K3/K4 — QT variant with a SOUND fp32 pre-filter.  grid = (ceil(Q / (128 * NN_QPT)), chunks)
//
// The exact test costs 24 fp64 instructions per (query, node) pair and the fp64 pipe issues at half rate.  Here every
// pair first gets an fp32 estimate of the squared distance in the expanded form
//     s~ = (|q^|^2 + |n^|^2) + sum_i (-2 q^_i) n^_i          (1 FADD + 8 FFMA, q^ = fl32(q), n^ = fl32(n))
// whose error against the reference's fp64 value s is bounded rigorously:
//     |s~ - s| <= 2 u R sqrt(s) + 37 u R^2,   u = 2^-23,  R >= max(|q|, |n|)      (input rounding + 18 fp32 roundings of
//                                                                                  partial sums bounded by 4 R^2)
// A node can replace the incumbent only if s < s_best, and then s~ < s_best + E with E = 4 u R d_best + 64 u R^2
// (twice the bound).  Pairs with s~ >= that threshold are skipped — provably without changing the result; the few
// that pass are evaluated exactly as before (same fp64 operations, same order, same comparisons), so ids stay
// bit-identical to the CPU's.  Radius mode: s~ + E < s_thr is a proven hit, s~ - E >= s_thr a proven miss, only
// the band in between is evaluated in fp64.
// ---------------------------------------------------------------------------------------------------------
#ifndef NN_GROUP_DEF
#define NN_GROUP_DEF 8
#endif
constexpr int NN_GROUP = NN_GROUP_DEF;   // nodes whose estimates are taken before the one (rare) branch
#ifndef NN_QPT_DEF
#define NN_QPT_DEF 2
#endif
constexpr int NN_QPT = NN_QPT_DEF;  // queries per thread (even): the three LDS of a node are shared by all of them
constexpr double NN_U = 1.1920928955078125e-07;   // 2^-23

// The estimates of the thread's TWO queries against one node, on the packed fp32x2 pipe (FFMA2: the node coordinate is
// a broadcast .F32 operand, the two queries' -2q_d sit in one register pair): 8 FFMA2 + 1 FADD2 per node instead of
// 16 FFMA + 2 FADD.  Per lane each operation is the same IEEE fma in the same order as the scalar form, so the
// estimates — and the error bound above — are unchanged.  The 3-register FFMA issues every other cycle per scheduler
// on this part; the scan was at ~89 % of that rate.
static_assert(NN_QPT % 2 == 0, "estimate_s2 packs the queries of a thread in pairs");
__device__ __forceinline__ float2 estimate_s2(const float2 (&m2q)[SQMP_DOF], float2 qq, const float4 a, const float4 b, float nn) {
  float2 s = f32x2_add(qq, nn);
  s = f32x2_fma(a.x, m2q[0], s); s = f32x2_fma(a.y, m2q[1], s); s = f32x2_fma(a.z, m2q[2], s); s = f32x2_fma(a.w, m2q[3], s);
  s = f32x2_fma(b.x, m2q[4], s); s = f32x2_fma(b.y, m2q[5], s); s = f32x2_fma(b.z, m2q[6], s); s = f32x2_fma(b.w, m2q[7], s);
  return s;
}
// __noinline__: inlined, the compiler hoists its sixteen 64-bit address computations above the (rarely taken) branch
// and they become a third of all executed instructions.
__device__ __noinline__ double exact_s(const TreePtrs &t, int64_t n, const double *__restrict__ query) {
  // node and query are 64-byte rows: four 16-byte loads each, all independent (the SoA columns would be eight L2 round trips)
  const double2 *np = reinterpret_cast<const double2 *>(t.aos + n * SQMP_DOF), *qp = reinterpret_cast<const double2 *>(query);
  const double2 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], q0 = __ldg(qp), q1 = __ldg(qp + 1), q2 = __ldg(qp + 2), q3 = __ldg(qp + 3);
  const double node[SQMP_DOF] = {n0.x, n0.y, n1.x, n1.y, n2.x, n2.y, n3.x, n3.y};
  const double q[SQMP_DOF] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
  return sumsq8(node, q);
}

template <bool RADIUS>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_qt_pf(TreePtrs t, int64_t N, int64_t chunk_len, const double *__restrict__ queries, int64_t Q, double rmax2,
           double *__restrict__ part_d, int32_t *__restrict__ part_id,
           // radius mode
           const int32_t *__restrict__ exclude, double s_thr, double r, int32_t *__restrict__ count,
           int32_t *__restrict__ hits, int64_t hit_cap) {
  // per node: q0..q7, |n|^2, pad — one base register, immediates.  Two buffers: the next tile is on its way (cp.async)
  // while this one is scanned
  __shared__ __align__(16) float tiles[2][NN_TILE][12];
  int64_t qi[NN_QPT];
  bool active[NN_QPT];
  float2 m2q[NN_QPT / 2][SQMP_DOF], qq2[NN_QPT / 2];   // pair p: .x = query 2p of the thread, .y = query 2p + 1
  float qq[NN_QPT];
  float thr_lo[NN_QPT];   // nearest: s~ >= thr_lo cannot win.  radius: s~ < thr_lo is a proven hit
  float thr_hi[NN_QPT];   // radius: s~ >= thr_hi is a proven miss
  double eR[NN_QPT], eR2[NN_QPT];   // 4 u R and 64 u R^2 of this query
  Best best[NN_QPT];
  int32_t excl[NN_QPT];
#pragma unroll
  for (int k = 0; k < NN_QPT; ++k) {
    qi[k] = (int64_t(blockIdx.x) * NN_QPT + k) * NN_THREADS + threadIdx.x;
    active[k] = qi[k] < Q;
    const double *qp = queries + (active[k] ? qi[k] : 0) * SQMP_DOF;
    double qn2 = 0.0, qq_d = 0.0;
#pragma unroll
    for (int d = 0; d < SQMP_DOF; ++d) {
      const double v = qp[d];
      const float f = __double2float_rn(v);
      if ((k & 1) == 0) m2q[k / 2][d].x = -2.0f * f;
      else m2q[k / 2][d].y = -2.0f * f;
      qn2 += v * v;
      qq_d += double(f) * double(f);
    }
    qq[k] = __double2float_rn(qq_d);
    if ((k & 1) == 0) qq2[k / 2].x = qq[k];
    else qq2[k / 2].y = qq[k];
    const double R2 = fmax(rmax2, qn2) * 1.000001;
    eR[k] = 4.0 * NN_U * sqrt(R2);
    eR2[k] = 64.0 * NN_U * R2;
    best[k].s = NN_INIT_SUMSQ; best[k].d = NN_INIT_DIST; best[k].id = 0;
    excl[k] = (RADIUS && active[k] && exclude) ? exclude[qi[k]] : -1;
    if (RADIUS) {
      const double E = eR[k] * r + eR2[k];
      // the "proven hit" direction needs s - 2uR sqrt(s) increasing at s_thr, i.e. r >> u R: otherwise exact only
      thr_lo[k] = (r > 1.0e3 * NN_U * sqrt(R2)) ? __double2float_rd(s_thr - E) : -CUDART_INF_F;
      thr_hi[k] = __double2float_ru(s_thr + E);
    } else {
      thr_lo[k] = __double2float_ru(NN_INIT_SUMSQ + eR[k] * NN_INIT_DIST + eR2[k]);
      thr_hi[k] = 0.0f;
    }
    if (!active[k]) { thr_lo[k] = -CUDART_INF_F; thr_hi[k] = -CUDART_INF_F; }
  }
  const int64_t n0 = int64_t(blockIdx.y) * chunk_len, n1 = min(N, n0 + chunk_len);

  auto fetch = [&](int buf, int64_t base) {   // asynchronous copies global -> shared; pad nodes: |n|^2 = +inf, never pass a threshold
    const int len = int(imin64(NN_TILE, n1 - base));
    for (int i = threadIdx.x; i < NN_TILE; i += NN_THREADS) {
      float *dst = &tiles[buf][i][0];
      if (i < len) {
        const unsigned d = unsigned(__cvta_generic_to_shared(dst));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(t.fa + base + i) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16u), "l"(t.fb + base + i) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 32u), "l"(t.fn + base + i) : "memory");
      } else {
        *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(dst + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        dst[8] = CUDART_INF_F;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n0 < n1) fetch(0, n0);
  int buf = 0;
  for (int64_t base = n0; base < n1; base += NN_TILE, buf ^= 1) {
    if (base + NN_TILE < n1) {
      fetch(buf ^ 1, base + NN_TILE);   // (its buffer was released by the barrier at the end of the previous iteration)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    float (*tile)[12] = tiles[buf];
    // whole tile (pad nodes are inert), NN_GROUP nodes at a time: all estimates first, ONE branch per group — the
    // exact path is re-entered per (node, query) only inside that rarely taken branch
#pragma unroll 1
    for (int i0 = 0; i0 < NN_TILE; i0 += NN_GROUP) {
      float se[NN_GROUP][NN_QPT];
      bool any = false;
#pragma unroll
      for (int u = 0; u < NN_GROUP; ++u) {
        const float4 a = *reinterpret_cast<const float4 *>(&tile[i0 + u][0]);   // broadcast reads
        const float4 b = *reinterpret_cast<const float4 *>(&tile[i0 + u][4]);
        const float nn = tile[i0 + u][8];
#pragma unroll
        for (int p = 0; p < NN_QPT / 2; ++p) {
          const float2 s2 = estimate_s2(m2q[p], qq2[p], a, b, nn);
          se[u][2 * p] = s2.x;
          se[u][2 * p + 1] = s2.y;
        }
#pragma unroll
        for (int k = 0; k < NN_QPT; ++k) any = any || (se[u][k] < (RADIUS ? thr_hi[k] : thr_lo[k]));
      }
      if (any) {
#pragma unroll
        for (int u = 0; u < NN_GROUP; ++u) {
#pragma unroll
          for (int k = 0; k < NN_QPT; ++k) {
            const int64_t n = base + i0 + u;
            if (RADIUS) {
              if (se[u][k] < thr_hi[k]) {            // not a proven miss (inactive queries: thr_hi = -inf)
                bool hit = se[u][k] < thr_lo[k];     // proven hit
                if (!hit) hit = exact_s(t, n, queries + qi[k] * SQMP_DOF) < s_thr;   // the band: exact
                if (hit && int32_t(n) != excl[k]) {
                  const int32_t slot = atomicAdd(&count[qi[k]], 1);
                  if (slot < hit_cap) hits[qi[k] * hit_cap + slot] = int32_t(n);
                }
              }
            } else {
              if (se[u][k] < thr_lo[k]) {            // cannot be excluded: exact evaluation, exact comparison
                const double s = exact_s(t, n, queries + qi[k] * SQMP_DOF);
                const double before = best[k].s;
                offer(best[k], s, int32_t(n));
                if (best[k].s != before) thr_lo[k] = __double2float_ru(best[k].s + eR[k] * best[k].d + eR2[k]);
              }
            }
          }
        }
      }
    }
    __syncthreads();   // everyone is done with this buffer: the fetch two tiles ahead may overwrite it
  }
  if (!RADIUS) {
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k)
      if (active[k]) {
        part_d[int64_t(blockIdx.y) * Q + qi[k]] = best[k].d;
        part_id[int64_t(blockIdx.y) * Q + qi[k]] = best[k].id;
      }
  }
}

// ---------------------------------------------------------------------------------------------------------
// TC ("tensor-core pre-filter", many queries): the estimate of every (query, node) pair comes out of tcgen05.mma
// instead of nine FFMA2 per pair.  A CTA keeps TC_QTILES x 128 queries as the A operand in shared memory and streams
// its node chunk through a ring of 6 KB tile images (bulk async copies, one elected thread); one elected thread issues
// two K = 16 fp16 MMAs per (query tile, node tile) into one of four TMEM accumulators (tc_umma.cuh).  The eight scan warps of
// a query tile (one per TMEM lane quarter and column half: a second warp keeps the TMEM read port busy while the first
// files) copy their part of their query's accumulator row to registers (one TMEM lane per thread, 64 columns), hand the
// accumulator back at once, and reduce the estimates with three-input minima; individual columns are only looked at
// when that minimum passes the thread's threshold.  What passes goes to a shared-memory ring and is re-evaluated
// exactly (exact_s) and compared exactly by two dedicated warps, 32 pairs at a time, off the scan's critical path — the
// same contract as QT/PF, so the answers are bit-identical; the band is wider (fp16 pieces instead of a few fp32 ulps) and
// still admits < 0.1 extra exact evaluations per query on the benchmark's tree.
//
// Error bound of the estimate v = D[query][node] against t = |n|^2 - 2 q.n in real arithmetic, R >= |q|, |n|:
//   operands:      q = qh + qm + eq, n = nh + nm + en with |eq| <= 2^-22 |q_d| + 2^-24 (the second piece may be an fp16
//                  subnormal), same for en; the dropped terms qm.nm, eq.n, q.en add up to <= 3 * 2^-22 R^2 + 2^-20 R in 2 q.n
//   |n|^2 pieces:  <= 2^-30 R^2 + 2^-23
//   accumulation:  19 exact products of magnitude <= R^2 summed by the tensor core over two instructions; budgeted
//                  2^-15 R^2 (a per-addition error of two fp32 ulps of the running magnitude 4 R^2, 32 times).
//                  The hardware's accumulation order is not documented, which is why the budget is 60 times the error the
//                  probe (tools/tc_probe_f16.cu) and the parity tests observe (2^-21.8 R^2).
//   total          < 2^-14 R^2 + 2^-19 R + 2^-22 = tc_error_bound(R^2).
// ---------------------------------------------------------------------------------------------------------
#ifndef TC_QTILES_DEF
#define TC_QTILES_DEF 2
#endif
constexpr int TC_QTILES = TC_QTILES_DEF;           // query tiles per CTA: every node tile is multiplied by each
constexpr int TC_QUERIES = TC_QTILES * TC_ROWS;    // 256 queries per CTA
constexpr int TC_STAGES = 8;                       // node tiles in flight
#ifndef TC_HALVES_DEF
#define TC_HALVES_DEF 2
#endif
constexpr int TC_HALVES = TC_HALVES_DEF;           // scan warps per (query tile, TMEM lane quarter): each takes TC_COLS columns of the accumulator
constexpr int TC_COLS = TC_ROWS / TC_HALVES;
constexpr int TC_EPI_WARPS = 4 * TC_QTILES * TC_HALVES;   // one scan warp per (query tile, TMEM lane quarter, column part)
constexpr int TC_EVAL_WARPS = 2;                   // exact evaluation of what the scan could not reject; warp e serves scan warps e, e + 2, ..
constexpr int TC_FIRST_EPI = 2 + TC_EVAL_WARPS;    // warp 0: copies, warp 1: MMA issue + TMEM, then the evaluation warps, then the scan warps
constexpr int TC_THREADS = 32 * (TC_FIRST_EPI + TC_EPI_WARPS);
static_assert(TC_FIRST_EPI % 4 == 0 && TC_EVAL_WARPS == 2, "scan warp k reads TMEM quarter k % 4; evaluation warp e owns the quarters of its parity");
constexpr int TC_TMEM_COLS = 512;                  // all of the SM's tensor memory:
constexpr int TC_BUFS = TC_TMEM_COLS / (TC_QTILES * TC_ROWS);   // accumulator buffers (node tiles in flight) x query tiles x 128 columns
constexpr int TC_RING = 256;                       // surviving pairs waiting for their exact evaluation, per scan warp
#ifndef TC_SEEDS_DEF
#define TC_SEEDS_DEF 32
#endif
#ifndef TC_PATIENCE_DEF
#define TC_PATIENCE_DEF 6
#endif
constexpr int TC_SEEDS = TC_SEEDS_DEF;                       // nodes of the chunk every query is measured against before the scan starts
constexpr int TC_MAX_CHUNK = 1 << 22;              // a ring entry keeps 22 bits of node offset
static_assert(TC_TILE_BYTES == SQMP_NN_TILE_IMG_BYTES, "capi.cu sizes the image with this");
__device__ __forceinline__ double tc_error_bound(double R2) {
  return (R2 * (1.0 / 16384.0) + sqrt(R2) * (1.0 / 524288.0) + 1.0 / 4194304.0) * 1.000001;
}

struct TcShared {
  uint64_t full[TC_STAGES], empty[TC_STAGES];   // node tile landed / its MMAs are done
  uint64_t acc_full[TC_BUFS][TC_QTILES], acc_empty[TC_BUFS][TC_QTILES];   // accumulator (buffer, query tile) written / copied out by its four scan warps
  uint32_t tmem;
  uint32_t tail[TC_EPI_WARPS], head[TC_EPI_WARPS];     // rings (one per scan warp): filled up to tail, consumed up to head
  uint32_t done;                                // scan warps that have finished
  float thr[TC_QUERIES];                        // per query: an accumulator value below this has to be looked at
  unsigned long long d_bits[TC_QUERIES];        // incumbent: distance (bits of a non-negative double order like integers)
  double best_s[TC_QUERIES];                    //            its sum of squares
  double e_minus_qq[TC_QUERIES];                // error bound of the estimate minus |q|^2
  int32_t best_id[TC_QUERIES];
  int32_t excl[TC_QUERIES];
  uint32_t ring[TC_EPI_WARPS][TC_RING];         // node offset in the chunk << 9 | query row << 1 | proven hit
};
constexpr int TC_SMEM_BYTES = (TC_QTILES + TC_STAGES) * TC_TILE_BYTES + int(sizeof(TcShared)) + 16;

__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ float ld_vol(const float *p) { return *reinterpret_cast<const volatile float *>(p); }

#ifdef SQMP_TC_NO_LD   // bring-up aid: the pipeline without the accumulator reads
#define TC_LD(a, v) do { for (int i_ = 0; i_ < 32; ++i_) v[i_] = 1e30f + float(a); } while (0)
#else
#define TC_LD(a, v) tc_tmem_ld32(a, v)
#endif
template <bool RADIUS>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_nn_tc(TreePtrs t, int64_t N, int64_t chunk_len, const double *__restrict__ queries, int64_t Q, double rmax2,
        double *__restrict__ part_d, int32_t *__restrict__ part_id,
        const int32_t *__restrict__ exclude, double s_thr, int32_t *__restrict__ count, int32_t *__restrict__ hits,
        int64_t hit_cap) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *sA = smem, *sB = smem + TC_QTILES * TC_TILE_BYTES;
  TcShared *sh = reinterpret_cast<TcShared *>(sB + TC_STAGES * TC_TILE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n0 = int64_t(blockIdx.y) * chunk_len, n1 = min(N, n0 + chunk_len);
  const int n_tiles = n0 < n1 ? int((n1 - n0 + TC_ROWS - 1) / TC_ROWS) : 0;
  const float4 *img = t.img + (n0 / TC_ROWS) * TC_TILE_FLOAT4;
  const int64_t q_first = int64_t(blockIdx.x) * TC_QUERIES;   // query of row 0

  // ---- set-up: barriers, TMEM, the queries as the A operand, their thresholds
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { tc_mbar_init(&sh->full[i], 1); tc_mbar_init(&sh->empty[i], 1); }
    for (int i = 0; i < TC_BUFS * TC_QTILES; ++i) { tc_mbar_init(&sh->acc_full[0][i], 1); tc_mbar_init(&sh->acc_empty[0][i], 4 * TC_HALVES); }
    for (int i = 0; i < TC_EPI_WARPS; ++i) sh->tail[i] = sh->head[i] = 0u;
    sh->done = 0u;
    tc_fence_mbar_init();
  }
  if (warp == 1) tc_tmem_alloc(&sh->tmem, TC_TMEM_COLS);
  const int quarter = warp & 3;   // a warp may only read TMEM lanes [32 (warp % 4), +32)
  const int qt = ((warp - TC_FIRST_EPI) >> 2) % TC_QTILES, part = (warp - TC_FIRST_EPI) / (4 * TC_QTILES);   // its query tile, its columns
  const int row = qt * TC_ROWS + quarter * 32 + lane;              // this scan thread's query
  float thr_lo = -CUDART_INF_F;                                    // radius: an accumulator value below this is a proven hit
  if (warp >= TC_FIRST_EPI) {
    const bool live = q_first + row < Q;
    const double *qp = queries + (q_first + row) * SQMP_DOF;
    double v[SQMP_DOF], qq = 0.0;
#pragma unroll
    for (int d = 0; d < SQMP_DOF; ++d) {
      v[d] = live ? qp[d] : 0.0;
      qq += v[d] * v[d];
    }
    const bool usable = qq < TC_MAX_NORM2;   // fits fp16; otherwise (or NaN): a zero row, every pair of this query is evaluated exactly
    if (part == 0) tc_write_query_row(reinterpret_cast<float4 *>(sA + qt * TC_TILE_BYTES), quarter * 32 + lane, v, usable);
    double E = tc_error_bound(fmax(rmax2, qq));
    float thr = -CUDART_INF_F;   // inactive: nothing passes
    if (live && !usable) {
      thr = CUDART_INF_F;
      E = CUDART_INF;
    } else if (live) {
      if (RADIUS) {
        thr = __double2float_ru(s_thr + E - qq);      // v >= thr: proven miss
        thr_lo = __double2float_rd(s_thr - E - qq);
      } else {
        // a few nodes of the chunk measured exactly: the scan starts with a threshold that already rejects most nodes
        // (they are not offered as incumbents here — the scan meets them again in node order)
        double smin = NN_INIT_SUMSQ;
        const int64_t len = n1 - n0, step = len / TC_SEEDS;
        if (step > 0 && part == 0)
          for (int k = 0; k < TC_SEEDS; ++k) smin = fmin(smin, exact_s(t, n0 + k * step, qp));
        thr = __double2float_ru(smin + E - qq);
      }
    }
    if (part == 0) {   // the query's shared state has one owner among the threads that scan for it
      sh->thr[row] = thr;
      sh->d_bits[row] = (unsigned long long)__double_as_longlong(NN_INIT_DIST);
      sh->best_s[row] = NN_INIT_SUMSQ;
      sh->best_id[row] = 0;
      sh->e_minus_qq[row] = usable ? E - qq : CUDART_INF;
      sh->excl[row] = (RADIUS && live && exclude) ? exclude[q_first + row] : -1;
    }
    tc_fence_async_smem();   // the A operand was written through the generic proxy; the tensor core reads through the async one
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem;

  if (warp == 0) {
    // ---- copies: one thread keeps TC_STAGES node tiles in flight
    if (lane == 0) {
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % TC_STAGES;
        tc_mbar_wait(&sh->empty[s], ((it / TC_STAGES) & 1) ^ 1);
        tc_mbar_expect_tx(&sh->full[s], TC_TILE_BYTES);
        tc_bulk_g2s(sB + s * TC_TILE_BYTES, img + int64_t(it) * TC_TILE_FLOAT4, TC_TILE_BYTES, &sh->full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issue: one thread
    if (lane == 0) {
      const uint32_t a_base = tc_smem_u32(sA), b_base = tc_smem_u32(sB);
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % TC_STAGES, buf = it % TC_BUFS;
        tc_mbar_wait(&sh->full[s], (it / TC_STAGES) & 1);
        const uint32_t b = b_base + s * TC_TILE_BYTES;
        const uint64_t b_hi_hi = tc_desc(b, 0), b_mid_nrm = tc_desc(b + TC_CHUNK_BYTES, TC_CHUNK_BYTES);
#pragma unroll
        for (int k = 0; k < TC_QTILES; ++k) {
          const uint32_t a = a_base + k * TC_TILE_BYTES, d = tmem + uint32_t((buf * TC_QTILES + k) * TC_ROWS);
          const uint64_t a_hi_mid = tc_desc(a, TC_CHUNK_BYTES), a_hi_one = tc_desc(a, 2 * TC_CHUNK_BYTES);
          tc_mbar_wait(&sh->acc_empty[buf][k], ((it / TC_BUFS) & 1) ^ 1);
          tc_fence_after();
          tc_mma_f16(d, a_hi_mid, b_hi_hi, false);    // [qh | qm] . [-2nh | -2nh]
          tc_mma_f16(d, a_hi_one, b_mid_nrm, true);   // [qh | 1 1 1] . [-2nm | |n|^2 pieces]
          tc_commit(&sh->acc_full[buf][k]);   // this query tile's accumulator may be read
        }
        tc_commit(&sh->empty[s]);             // the node tile's slot may be refilled
      }
    }
    __syncwarp();
  } else if (warp < TC_FIRST_EPI) {
    // ---- exact evaluation of the pairs the scan could not reject, 32 at a time (one pair per lane), off the scan's
    //      critical path: an L2 round trip here does not hold up the accumulator ring.  Warp e serves the rings of scan
    //      warps e and e + 4, so every query's incumbent has exactly one writer.
    const int ev = warp - 2;
    constexpr int SERVED = TC_EPI_WARPS / TC_EVAL_WARPS;
    uint32_t head[SERVED];
#pragma unroll
    for (int k = 0; k < SERVED; ++k) head[k] = 0u;
    int patience = 0;
    for (;;) {
      // a batch is gathered from all its rings; it waits a little for a full one (an L2 round trip costs the same for 3 pairs
      // as for 32), but not long: the scan's thresholds tighten only when pairs are evaluated
      uint32_t av[SERVED], total = 0u;
#pragma unroll
      for (int i = 0; i < SERVED; ++i) {
        av[i] = ld_vol(&sh->tail[ev + i * TC_EVAL_WARPS]) - head[i];
        total += av[i];
      }
      const bool scan_done = ld_vol(&sh->done) == uint32_t(TC_EPI_WARPS);
      if (total == 0u) {
        if (!scan_done) { __nanosleep(200); continue; }
        // every scan warp has published its last tail before `done`: one more look
        bool empty = true;
#pragma unroll
        for (int i = 0; i < SERVED; ++i) empty = empty && ld_vol(&sh->tail[ev + i * TC_EVAL_WARPS]) == head[i];
        if (empty) break;
        continue;
      }
      if (total < 32u && !scan_done && patience < TC_PATIENCE_DEF) { ++patience; __nanosleep(200); continue; }
      patience = 0;
      uint32_t e = 0u;
      int start = 0;
#pragma unroll
      for (int i = 0; i < SERVED; ++i) {
        const int rid = ev + i * TC_EVAL_WARPS;
        const int take = min(int(av[i]), 32 - start);
        if (lane >= start && lane < start + take)
          e = 0x80000000u | *reinterpret_cast<volatile uint32_t *>(&sh->ring[rid][(head[i] + uint32_t(lane - start)) % TC_RING]);
        head[i] += uint32_t(take);
        start += take;
      }
      const bool have = e != 0u;
      __threadfence_block();
      __syncwarp();
      if (lane < SERVED) {   // the slots may be reused
        uint32_t h = 0u;
#pragma unroll
        for (int i = 0; i < SERVED; ++i) h = lane == i ? head[i] : h;
        *reinterpret_cast<volatile uint32_t *>(&sh->head[ev + lane * TC_EVAL_WARPS]) = h;
      }
      const int r = int(e >> 1) & (TC_QUERIES - 1);
      const int64_t n = n0 + int64_t((e >> 9) & (TC_MAX_CHUNK - 1));
      const double *qp = queries + (q_first + r) * SQMP_DOF;
      if (RADIUS) {
        if (have && ((e & 1u) || exact_s(t, n, qp) < s_thr) && int32_t(n) != sh->excl[r]) {
          const int32_t slot = atomicAdd(&count[q_first + r], 1);
          if (slot < hit_cap) hits[(q_first + r) * hit_cap + slot] = int32_t(n);
        }
      } else {
        // lexicographic (distance, id) minimum per query = what the reference's serial loop keeps (strict '<', BS:4122)
        double s = CUDART_INF, d = CUDART_INF;
        if (have) {
          s = exact_s(t, n, qp);
          d = sqrt(s);
        }
        const unsigned long long db = (unsigned long long)__double_as_longlong(d);
        // (several scan warps file for one query, so its pairs do not arrive in node order: an equal distance with a lower
        //  id has to replace the incumbent too; INIT's id 0 only ever meets distances below 10000)
        const unsigned long long before = have ? sh->d_bits[r] : 0ull;
        const bool better_d = have && db <= before && d < NN_INIT_DIST;
        __syncwarp();
        if (better_d) atomicMin(&sh->d_bits[r], db);
        __syncwarp();
        const bool best_d = better_d && sh->d_bits[r] == db;
        const int32_t id_before = best_d ? sh->best_id[r] : 0;
        __syncwarp();
        if (best_d && db < before) sh->best_id[r] = 0x7fffffff;   // a strictly smaller distance: the old id is out
        __syncwarp();
        if (best_d) atomicMin(&sh->best_id[r], int32_t(n));
        __syncwarp();
        if (best_d && sh->best_id[r] == int32_t(n) && (db < before || int32_t(n) < id_before)) {
          sh->best_s[r] = s;
          const float nt = __double2float_ru(s + sh->e_minus_qq[r]);
          if (nt < sh->thr[r]) *reinterpret_cast<volatile float *>(&sh->thr[r]) = nt;
        }
        __syncwarp();
      }
    }
    if (!RADIUS) {
      for (int k = 0; k < SERVED; ++k) {   // the queries of the scan warps it serves
        const int sw = ev + k * TC_EVAL_WARPS;                               // that scan warp: its query tile and TMEM quarter
        if (sw >= 4 * TC_QTILES) continue;                                   // (the other column parts scan for the same queries)
        const int r = (sw >> 2) * TC_ROWS + (sw & 3) * 32 + lane;
        if (q_first + r < Q) {
          part_d[int64_t(blockIdx.y) * Q + q_first + r] = __longlong_as_double((long long)sh->d_bits[r]);
          part_id[int64_t(blockIdx.y) * Q + q_first + r] = sh->best_id[r];
        }
      }
    }
  } else {
    // ---- scan: this thread's query against every node of the chunk; what passes the threshold goes to the warp's ring
    //      (single producer: the tail lives in a register and is published once per 32-column group)
    const int rid = warp - TC_FIRST_EPI;
    uint32_t *ring = sh->ring[rid];
    uint32_t tail = 0u, head_seen = 0u;   // warp-uniform
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it % TC_BUFS;
      const int64_t base = n0 + int64_t(it) * TC_ROWS + part * TC_COLS;
      tc_mbar_wait(&sh->acc_full[buf][qt], (it / TC_BUFS) & 1);
      tc_fence_after();
      // the query's estimates for this warp's columns move to registers in one go and the accumulator is handed back before
      // anything is looked at: the tensor core never waits for the scan, only for this copy
      const uint32_t taddr = tmem + (uint32_t(quarter * 32) << 16) + uint32_t((buf * TC_QTILES + qt) * TC_ROWS + part * TC_COLS);
      constexpr int NG = TC_COLS / 32;
      float v[TC_COLS];
#pragma unroll
      for (int g = 0; g < NG; ++g) TC_LD(taddr + 32 * g, v + 32 * g);
      tc_tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&sh->acc_empty[buf][qt]);
      const float thr = ld_vol(&sh->thr[row]);
#ifdef SQMP_TC_LD_ONLY
      if (v[0] == 123.456f && v[17] == thr) thr_lo = 0.f;
      continue;
#endif
      float ms[4 * NG];   // minimum of every run of eight estimates
#pragma unroll
      for (int k = 0; k < 4 * NG; ++k)
        ms[k] = fminf(tc_min3(tc_min3(v[8 * k], v[8 * k + 1], v[8 * k + 2]), tc_min3(v[8 * k + 3], v[8 * k + 4], v[8 * k + 5]), v[8 * k + 6]),
                      v[8 * k + 7]);
      float mg[NG];       // ... of every 32
      float m = CUDART_INF_F;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        mg[g] = fminf(tc_min3(ms[4 * g], ms[4 * g + 1], ms[4 * g + 2]), ms[4 * g + 3]);
        m = fminf(m, mg[g]);
      }
#ifdef SQMP_TC_MIN_ONLY
      if (m == 123.456f) thr_lo = 0.f;
      continue;
#endif
      if (!__any_sync(0xffffffffu, m < thr)) continue;   // the common case: nothing in this tile for any of the 32 queries
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        unsigned any = __ballot_sync(0xffffffffu, mg[g] < thr);
        if (any == 0u) continue;
        uint32_t cm = 0u, pm = 0u;   // columns that pass / that are proven hits; only the runs some lane needs are looked at
#pragma unroll
        for (int k = 4 * g; k < 4 * g + 4; ++k)
          if (__any_sync(0xffffffffu, ms[k] < thr)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              cm |= v[8 * k + j] < thr ? 1u << (8 * (k & 3) + j) : 0u;
              if (RADIUS) pm |= v[8 * k + j] < thr_lo ? 1u << (8 * (k & 3) + j) : 0u;
            }
          }
        while (any) {   // one owner at a time; lane j files column j of this run of 32
          const int owner = __ffs(any) - 1;
          any &= any - 1u;
          const uint32_t cmo = __shfl_sync(0xffffffffu, cm, owner), pmo = __shfl_sync(0xffffffffu, pm, owner);
          const int64_t n = base + 32 * g + lane;
          const bool mine = ((cmo >> lane) & 1u) && n < n1;
          const unsigned bal = __ballot_sync(0xffffffffu, mine);
          const uint32_t add = uint32_t(__popc(bal));
          if (tail + add - head_seen > uint32_t(TC_RING)) {   // full: let the evaluator see what is there, wait for it
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&sh->tail[rid]) = tail;
            do head_seen = ld_vol(&sh->head[rid]);
            while (tail + add - head_seen > uint32_t(TC_RING));
          }
          if (mine)
            ring[(tail + __popc(bal & lt_mask)) % TC_RING] = uint32_t(n - n0) << 9 | uint32_t(row - lane + owner) << 1 | ((pmo >> lane) & 1u);
          tail += add;
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&sh->tail[rid]) = tail;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicAdd(&sh->done, 1u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_tmem_free(tmem, TC_TMEM_COLS);
  }
}

// fold the per-chunk partial results in chunk (= id) order; strict '<' keeps the lowest id on ties (BS:4122)
__global__ void k_nn_merge(const double *__restrict__ part_d, const int32_t *__restrict__ part_id, int chunks,
                           int64_t Q, int32_t *__restrict__ id_out) {
  const int64_t qi = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (qi >= Q) return;
  double bd = NN_INIT_DIST;
  int32_t bid = 0;
  for (int c = 0; c < chunks; ++c) {
    const double d = part_d[int64_t(c) * Q + qi];
    if (d < bd) { bd = d; bid = part_id[int64_t(c) * Q + qi]; }
  }
  id_out[qi] = bid;
}

// ---------------------------------------------------------------------------------------------------------
// K4 selection: one warp per query picks the k lowest and k highest (cost, id) among the compacted hits
// ---------------------------------------------------------------------------------------------------------
struct Key {
  double c;
  int32_t id;
};
__device__ __forceinline__ bool key_lt(const Key &a, const Key &b) { return a.c < b.c || (a.c == b.c && a.id < b.id); }
__device__ __forceinline__ Key warp_min_key(Key v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Key w;
    w.c = __shfl_xor_sync(0xffffffffu, v.c, o);
    w.id = __shfl_xor_sync(0xffffffffu, v.id, o);
    if (key_lt(w, v)) v = w;
  }
  return v;
}
__device__ __forceinline__ Key warp_max_key(Key v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Key w;
    w.c = __shfl_xor_sync(0xffffffffu, v.c, o);
    w.id = __shfl_xor_sync(0xffffffffu, v.id, o);
    if (key_lt(v, w)) v = w;
  }
  return v;
}

// one warp: the k lowest and the k highest (cost, id) of the m hits h[0..m) -> low[0..k), high[0..k) (-1 padded)
__device__ __forceinline__ void warp_topk(const double *__restrict__ cost, const int32_t *__restrict__ h, int32_t m, int k,
                                          int32_t *__restrict__ low, int32_t *__restrict__ high, int lane) {
  const Key kmax = {CUDART_INF, 0x7fffffff}, kmin = {-CUDART_INF, -1};
  Key last = kmin;
  for (int tsel = 0; tsel < k; ++tsel) {   // t-th lowest = smallest key strictly above the previous pick
    Key mine = kmax;
    for (int i = lane; i < m; i += 32) {
      Key c;
      c.id = h[i];
      c.c = cost[c.id];
      if (key_lt(last, c) && key_lt(c, mine)) mine = c;
    }
    const Key w = warp_min_key(mine);
    const bool found = w.id != kmax.id;
    if (lane == 0) low[tsel] = found ? w.id : -1;
    if (!found) {
      for (int r = tsel + 1 + lane; r < k; r += 32) low[r] = -1;
      break;
    }
    last = w;
  }
  last = kmax;
  for (int tsel = 0; tsel < k; ++tsel) {   // t-th highest = largest key strictly below the previous pick
    Key mine = kmin;
    for (int i = lane; i < m; i += 32) {
      Key c;
      c.id = h[i];
      c.c = cost[c.id];
      if (key_lt(c, last) && key_lt(mine, c)) mine = c;
    }
    const Key w = warp_max_key(mine);
    const bool found = w.id != kmin.id;
    if (lane == 0) high[tsel] = found ? w.id : -1;
    if (!found) {
      for (int r = tsel + 1 + lane; r < k; r += 32) high[r] = -1;
      break;
    }
    last = w;
  }
}

__global__ void __launch_bounds__(NN_THREADS)
k_nn_topk(const double *__restrict__ cost, const int32_t *__restrict__ hits, int64_t hit_cap,
          const int32_t *__restrict__ count, int64_t Q, int k, int32_t *__restrict__ low, int32_t *__restrict__ high,
          int32_t *__restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int64_t qi = int64_t(blockIdx.x) * (NN_THREADS / 32) + (threadIdx.x >> 5);
  if (qi >= Q) return;
  const int32_t total = count[qi];
  if (total > hit_cap) {
    if (lane == 0) atomicExch(overflow, 1);
  }
  warp_topk(cost, hits + qi * hit_cap, int32_t(imin64(total, hit_cap)), k, low + qi * k, high + qi * k, lane);
}

// ---------------------------------------------------------------------------------------------------------
// Many queries, EACH AGAINST ITS OWN TREE (the lock-step multi-query planner, shim/sqmp_multi.hpp: every query of
// a batch owns two small trees).  One warp per query: lanes stride over the nodes (coalesced SoA reads), exact fp64
// distances in the reference's order, warp argmin over (d, id) for the nearest node; for the near vertices the hits
// are ballot-compacted into the query's slice of the hit pool and the same warp selects the k lowest / k highest
// (cost, id).  Inputs and answers live in mapped pinned host memory: one launch + one synchronise per batch.
// ---------------------------------------------------------------------------------------------------------
struct MultiQuery {          // 96 bytes, written by the host for every query of a call
  const double *q[SQMP_DOF];
  const double *cost;
  int64_t n;
  int32_t exclude;
  int32_t pad;
};
template <bool RADIUS>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_multi(const MultiQuery *__restrict__ desc, const double *__restrict__ queries, int64_t Q, double s_thr, int k,
           int32_t *__restrict__ hits, int64_t hit_cap, int32_t *__restrict__ out /* nearest: [Q]; radius: [Q][2k+2] */) {
  const int lane = threadIdx.x & 31;
  const int64_t qi = int64_t(blockIdx.x) * (NN_THREADS / 32) + (threadIdx.x >> 5);
  if (qi >= Q) return;
  const MultiQuery t = desc[qi];
  double query[SQMP_DOF];
#pragma unroll
  for (int d = 0; d < SQMP_DOF; ++d) query[d] = queries[qi * SQMP_DOF + d];
  Best best = {NN_INIT_SUMSQ, NN_INIT_DIST, 0};
  int32_t total = 0;   // warp-uniform
  int32_t *h = hits + qi * hit_cap;
  for (int64_t base = 0; base < t.n; base += 32) {
    const int64_t n = base + lane;
    bool hit = false;
    if (n < t.n) {
      double node[SQMP_DOF];
#pragma unroll
      for (int d = 0; d < SQMP_DOF; ++d) node[d] = t.q[d][n];
      const double s = sumsq8(node, query);
      if (RADIUS) hit = s < s_thr && int32_t(n) != t.exclude;
      else offer(best, s, int32_t(n));
    }
    if (RADIUS) {
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      const int32_t slot = total + __popc(m & ((1u << lane) - 1u));
      if (hit && slot < hit_cap) h[slot] = int32_t(n);
      total += __popc(m);
    }
  }
  if (!RADIUS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, best.d, o);
      const int32_t oid = __shfl_xor_sync(0xffffffffu, best.id, o);
      if (better(od, oid, best.d, best.id)) { best.d = od; best.id = oid; }
    }
    if (lane == 0) out[qi] = best.id;
  } else {
    __syncwarp();   // the hit list was written by this warp: visible to all its lanes
    int32_t *o = out + qi * (2 * k + 2);
    if (lane == 0) {
      o[0] = total;
      o[1] = total > hit_cap ? 1 : 0;   // overflow: the host redoes this query through the single-tree path
    }
    warp_topk(t.cost, h, int32_t(imin64(total, hit_cap)), k, o + 2, o + 2 + k, lane);
  }
}
cudaError_t launch_nn_multi(const void *d_desc, const double *d_queries, int64_t Q, bool radius, double s_thr, int k,
                            NNScratch &sc, int32_t *d_out, cudaStream_t st) {
  if (Q <= 0) return cudaSuccess;
  const unsigned blocks = unsigned((Q + NN_THREADS / 32 - 1) / (NN_THREADS / 32));
  if (radius)
    k_nn_multi<true><<<blocks, NN_THREADS, 0, st>>>((const MultiQuery *)d_desc, d_queries, Q, s_thr, k, sc.hits, sc.hit_cap, d_out);
  else
    k_nn_multi<false><<<blocks, NN_THREADS, 0, st>>>((const MultiQuery *)d_desc, d_queries, Q, 0.0, 0, nullptr, 0, d_out);
  return cudaGetLastError();
}
size_t nn_multi_desc_bytes() { return sizeof(MultiQuery); }
void nn_multi_fill_desc(void *dst, const TreeDev &t, int32_t exclude) {
  MultiQuery m;
  for (int d = 0; d < SQMP_DOF; ++d) m.q[d] = t.q[d];
  m.cost = t.cost;
  m.n = t.n;
  m.exclude = exclude;
  m.pad = 0;
  *reinterpret_cast<MultiQuery *>(dst) = m;
}

// ---------------------------------------------------------------------------------------------------------
// The planner's own call shape: ONE query against a tree of 10^2..10^5 nodes (BS:1189, 1226, 1256 call these once
// per iteration).  One launch does everything: the query arrives in the kernel parameters (no H2D copy), every CTA
// scans a slice exactly (fp64, reference order), the last CTA to finish (atomic ticket) folds the partial results —
// for the radius query it also selects the k lowest / k highest (cost, id) — and writes the answer straight into
// mapped pinned host memory (no D2H copy).  Host cost of a call: one launch + one stream synchronise.
// ---------------------------------------------------------------------------------------------------------
struct OneQuery {
  double q[SQMP_DOF];
};
constexpr int NN1_THREADS = 256;

__device__ __forceinline__ Key block_min_key(Key v, Key *sh) {   // all threads get the block minimum
  v = warp_min_key(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  Key r = sh[0];
  for (int w = 1; w < NN1_THREADS / 32; ++w)
    if (key_lt(sh[w], r)) r = sh[w];
  return r;
}
__device__ __forceinline__ Key block_max_key(Key v, Key *sh) {
  v = warp_max_key(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  Key r = sh[0];
  for (int w = 1; w < NN1_THREADS / 32; ++w)
    if (key_lt(r, sh[w])) r = sh[w];
  return r;
}

template <bool RADIUS>
__global__ void __launch_bounds__(NN1_THREADS)
k_nn_single(TreePtrs t, int64_t N, OneQuery query, int32_t excl, double s_thr, int k, double *__restrict__ part_d,
            int32_t *__restrict__ part_id, unsigned *__restrict__ ticket, int32_t *__restrict__ hit_count,
            int32_t *__restrict__ hits, volatile int32_t *__restrict__ out) {
  __shared__ Key sh[NN1_THREADS / 32];
  __shared__ bool is_last;
  Best best = {NN_INIT_SUMSQ, NN_INIT_DIST, 0};
  for (int64_t n = int64_t(blockIdx.x) * NN1_THREADS + threadIdx.x; n < N; n += int64_t(gridDim.x) * NN1_THREADS) {
    double node[SQMP_DOF];
#pragma unroll
    for (int d = 0; d < SQMP_DOF; ++d) node[d] = t.q[d][n];   // coalesced SoA
    const double s = sumsq8(node, query.q);
    if (RADIUS) {
      if (s < s_thr && int32_t(n) != excl) hits[atomicAdd(hit_count, 1)] = int32_t(n);   // capacity = tree size
    } else {
      offer(best, s, int32_t(n));
    }
  }
  if (!RADIUS) {   // (distance, id) lexicographic minimum of the CTA; Key reuses the (double, id) ordering
    Key v = {best.d, best.id};
    v = block_min_key(v, sh);
    if (threadIdx.x == 0) { part_d[blockIdx.x] = v.c; part_id[blockIdx.x] = v.id; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (!RADIUS) {
    // a CTA that found nothing below 10000.0 carries (10000.0, 0): it cannot beat a real candidate (d < 10000.0)
    Key v = {NN_INIT_DIST, 0};
    for (int b = threadIdx.x; b < int(gridDim.x); b += NN1_THREADS) {
      const Key c = {__ldcg(part_d + b), __ldcg(part_id + b)};
      if (key_lt(c, v)) v = c;
    }
    v = block_min_key(v, sh);
    if (threadIdx.x == 0) out[0] = v.id;
  } else {
    const int32_t total = __ldcg(hit_count);
    const Key kmax = {CUDART_INF, 0x7fffffff}, kmin = {-CUDART_INF, -1};
    Key last = kmin;
    bool done = false;
    for (int tsel = 0; tsel < k; ++tsel) {   // t-th lowest = smallest key strictly above the previous pick
      Key mine = kmax;
      if (!done)
        for (int i = threadIdx.x; i < total; i += NN1_THREADS) {
          Key c;
          c.id = __ldcg(hits + i);
          c.c = t.cost[c.id];
          if (key_lt(last, c) && key_lt(c, mine)) mine = c;
        }
      const Key w = block_min_key(mine, sh);
      const bool found = !done && w.id != kmax.id;
      if (threadIdx.x == 0) out[1 + tsel] = found ? w.id : -1;
      if (!found) done = true;
      last = w;
    }
    last = kmax;
    done = false;
    for (int tsel = 0; tsel < k; ++tsel) {   // t-th highest = largest key strictly below the previous pick
      Key mine = kmin;
      if (!done)
        for (int i = threadIdx.x; i < total; i += NN1_THREADS) {
          Key c;
          c.id = __ldcg(hits + i);
          c.c = t.cost[c.id];
          if (key_lt(c, last) && key_lt(mine, c)) mine = c;
        }
      const Key w = block_max_key(mine, sh);
      const bool found = !done && w.id != kmin.id;
      if (threadIdx.x == 0) out[1 + k + tsel] = found ? w.id : -1;
      if (!found) done = true;
      last = w;
    }
    if (threadIdx.x == 0) {
      out[0] = total;
      *hit_count = 0;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;   // ready for the next call on this stream
}

cudaError_t launch_nn_single(const TreeDev &t, const double *h_query, bool radius, int32_t excl, double s_thr, int k,
                             NNScratch &sc, unsigned *d_ticket, int32_t *d_hit_count, int32_t *d_out_mapped, int sm_count,
                             cudaStream_t st) {
  OneQuery q;
  for (int d = 0; d < SQMP_DOF; ++d) q.q[d] = h_query[d];
  int blocks = int((t.n + 4 * NN1_THREADS - 1) / (4 * NN1_THREADS));   // >= 4 nodes per thread before adding CTAs
  if (blocks > sm_count) blocks = sm_count;
  if (blocks < 1) blocks = 1;
  if (radius)
    k_nn_single<true><<<blocks, NN1_THREADS, 0, st>>>(ptrs_of(t), t.n, q, excl, s_thr, k, sc.part_d, sc.part_id, d_ticket,
                                                      d_hit_count, sc.hits, d_out_mapped);
  else
    k_nn_single<false><<<blocks, NN1_THREADS, 0, st>>>(ptrs_of(t), t.n, q, excl, s_thr, k, sc.part_d, sc.part_id, d_ticket,
                                                       d_hit_count, sc.hits, d_out_mapped);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
// the tensor-core scan needs more dynamic shared memory than the default limit (a per-device attribute: set per launch)
static cudaError_t tc_prepare() {
  cudaError_t e = cudaFuncSetAttribute(k_nn_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_nn_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  return e;
}
int nn_tc_queries_per_cta() { return TC_QUERIES; }
static void pick_chunks(int64_t N, int64_t Q, int variant, int chunks_hint, int &chunks, int64_t &chunk_len) {
  (void)Q;
  chunks = chunks_hint < 1 ? 1 : chunks_hint;
  chunk_len = (N + chunks - 1) / chunks;
  const int64_t gran = variant == 2 ? 32 : NN_TILE;   // variants 1, 3 and 4 stream NN_TILE-node tiles (two TC tiles)
  chunk_len = ((chunk_len + gran - 1) / gran) * gran;
  if (chunk_len < gran) chunk_len = gran;
  chunks = int((N + chunk_len - 1) / chunk_len);
  if (chunks < 1) chunks = 1;
}

cudaError_t launch_nn_nearest(const TreeDev &t, const double *d_queries, int64_t Q, int32_t *d_id, NNScratch &sc,
                              int chunks_hint, int variant, cudaStream_t st, int *n_launch) {
  if (Q <= 0) return cudaSuccess;
  int chunks;
  int64_t chunk_len;
  pick_chunks(t.n, Q, variant, chunks_hint, chunks, chunk_len);
  if (variant == 4) {
    cudaError_t ea = tc_prepare();
    if (ea != cudaSuccess) return ea;
    dim3 grid(unsigned((Q + TC_QTILES * TC_ROWS - 1) / (TC_QTILES * TC_ROWS)), unsigned(chunks));
    k_nn_tc<false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ptrs_of(t), t.n, chunk_len, d_queries, Q, t.rmax2, sc.part_d,
                                                           sc.part_id, nullptr, 0.0, nullptr, nullptr, 0);
  } else if (variant == 3) {
    dim3 grid(unsigned((Q + NN_THREADS * NN_QPT - 1) / (NN_THREADS * NN_QPT)), unsigned(chunks));
    k_nn_qt_pf<false><<<grid, NN_THREADS, 0, st>>>(ptrs_of(t), t.n, chunk_len, d_queries, Q, t.rmax2, sc.part_d,
                                                   sc.part_id, nullptr, 0.0, 0.0, nullptr, nullptr, 0);
  } else if (variant == 2) {
    dim3 grid(unsigned(Q), unsigned((chunks + NN_THREADS / 32 - 1) / (NN_THREADS / 32)));
    k_nn_wq<false><<<grid, NN_THREADS, 0, st>>>(ptrs_of(t), t.n, chunk_len, chunks, d_queries, Q, sc.part_d,
                                                sc.part_id, nullptr, 0.0, nullptr, nullptr, 0);
  } else {
    dim3 grid(unsigned((Q + NN_THREADS - 1) / NN_THREADS), unsigned(chunks));
    k_nn_qt<false><<<grid, NN_THREADS, 0, st>>>(ptrs_of(t), t.n, chunk_len, d_queries, Q, sc.part_d, sc.part_id,
                                                nullptr, 0.0, nullptr, nullptr, 0);
  }
  k_nn_merge<<<unsigned((Q + 255) / 256), 256, 0, st>>>(sc.part_d, sc.part_id, chunks, Q, d_id);
  if (n_launch) *n_launch += 2;
  return cudaGetLastError();
}

cudaError_t launch_nn_radius(const TreeDev &t, const double *d_queries, const int32_t *d_exclude, int64_t Q,
                             double s_thr, double r, int k, int32_t *d_low, int32_t *d_high, int32_t *d_count,
                             int32_t *d_overflow, NNScratch &sc, int chunks_hint, int variant, cudaStream_t st,
                             int *n_launch) {
  if (Q <= 0) return cudaSuccess;
  int chunks;
  int64_t chunk_len;
  pick_chunks(t.n, Q, variant, chunks_hint, chunks, chunk_len);
  cudaError_t e = cudaMemsetAsync(d_count, 0, size_t(Q) * sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  if (variant == 4) {
    cudaError_t ea = tc_prepare();
    if (ea != cudaSuccess) return ea;
    dim3 grid(unsigned((Q + TC_QTILES * TC_ROWS - 1) / (TC_QTILES * TC_ROWS)), unsigned(chunks));
    k_nn_tc<true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ptrs_of(t), t.n, chunk_len, d_queries, Q, t.rmax2, nullptr, nullptr,
                                                          d_exclude, s_thr, d_count, sc.hits, sc.hit_cap);
  } else if (variant == 3) {
    dim3 grid(unsigned((Q + NN_THREADS * NN_QPT - 1) / (NN_THREADS * NN_QPT)), unsigned(chunks));
    k_nn_qt_pf<true><<<grid, NN_THREADS, 0, st>>>(ptrs_of(t), t.n, chunk_len, d_queries, Q, t.rmax2, nullptr, nullptr,
                                                  d_exclude, s_thr, r, d_count, sc.hits, sc.hit_cap);
  } else if (variant == 2) {
    dim3 grid(unsigned(Q), unsigned((chunks + NN_THREADS / 32 - 1) / (NN_THREADS / 32)));
    k_nn_wq<true><<<grid, NN_THREADS, 0, st>>>(ptrs_of(t), t.n, chunk_len, chunks, d_queries, Q, nullptr, nullptr,
                                               d_exclude, s_thr, d_count, sc.hits, sc.hit_cap);
  } else {
    dim3 grid(unsigned((Q + NN_THREADS - 1) / NN_THREADS), unsigned(chunks));
    k_nn_qt<true><<<grid, NN_THREADS, 0, st>>>(ptrs_of(t), t.n, chunk_len, d_queries, Q, nullptr, nullptr, d_exclude,
                                               s_thr, d_count, sc.hits, sc.hit_cap);
  }
  k_nn_topk<<<unsigned((Q + NN_THREADS / 32 - 1) / (NN_THREADS / 32)), NN_THREADS, 0, st>>>(
      t.cost, sc.hits, sc.hit_cap, d_count, Q, k, d_low, d_high, d_overflow);
  if (n_launch) *n_launch += 2;
  return cudaGetLastError();
}

}  // namespace sqmp
