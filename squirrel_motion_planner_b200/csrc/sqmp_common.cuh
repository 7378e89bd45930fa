// sqmp_common.cuh — device-side descriptors shared by the kernels of libsqmp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sqmp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsqmp_b200 is written for sm_100a (B200) only"
#endif

namespace sqmp {

// The validity kernels walk "slots": the spheres of a rigid group are ordered cluster by cluster (a cluster = up to 8
// spheres of ONE link that lie close together, capi.cu: build_hierarchy) and every cluster is padded to a multiple of 4
// slots (pads can never test positive), so the 4- and 8-wide batches need no bounds logic.
constexpr int SQMP_MAX_CLUSTERS = 256;
constexpr int SQMP_MAX_SLOTS = SQMP_MAX_SPHERES + 3 * SQMP_MAX_CLUSTERS;
constexpr int SQMP_MAX_GROUPS = SQMP_MAX_SEG + 1;   // rigid group g = segment g - 1; group 0 hangs on the root

// axis codes: the joint axes of the reference robot are all signed unit axes (URDF:30-66, 615-759);
// a general axis takes the Rodrigues path.
__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

enum AxisCode : int { AX_GENERAL = 0, AX_PX = 1, AX_NX = 2, AX_PY = 3, AX_NY = 4, AX_PZ = 5, AX_NZ = 6 };

// The kinematic chain, passed by value as a __grid_constant__ kernel parameter: every access is warp-uniform, so it
// is served from the constant bank / uniform datapath and never touches the LSU.  (The sphere tables live in global
// memory / in KfTab: a covering model has several hundred spheres.)
struct RobotDev {
  int n_seg;
  int n_slots;                            // grid-unit copy: slots of the generic kernel's table (K1Slot)
  int ref_seg;                            // last segment of the leading prismatic run (base position)
  int self_const_collision;
  unsigned trans_mask;                    // bit j: q[j] drives a prismatic joint (a length)
  double len_scale;                       // 1 (metres) or 1/res (grid units, the copy K1 runs on)
  int seg_type[SQMP_MAX_SEG];
  int seg_qidx[SQMP_MAX_SEG];
  int seg_axis_code[SQMP_MAX_SEG];
  int seg_pre_kind[SQMP_MAX_SEG];         // 0 identity, 1 translation only, 2 general
  int grp_slot_end[SQMP_MAX_GROUPS + 1];  // slots of group g are [end[g], end[g+1])
  int grp_fidx[SQMP_MAX_GROUPS];          // group -> slot of its frame in the self-collision scratch, -1: none
  int n_frames;
  double seg_pre[SQMP_MAX_SEG][12];
  double seg_axis[SQMP_MAX_SEG][3];
  double root_off[3];
  double ee_tip[12];
  int ee_tip_kind;
  // fp32 copies of the constants (made by the host once; saves the F2F conversions in the float chain)
  float seg_pre_f[SQMP_MAX_SEG][12];
  float seg_axis_f[SQMP_MAX_SEG][4];
  float ee_tip_f[12];
  // per-request (sqmp_set_disabled_links, CC:90-102): bit l = link l is skipped by the map check
  uint64_t link_disabled;
};

// One slot of the generic validity kernel (validity.cu), global memory, grid units.
struct K1Slot {
  double x, y, z, r;   // centre in the group frame, radius
  int thr;             // floor(rg^2): "lb2 > thr" proves the sphere free (integers, exact); -2: pad slot
  int ncell;           // ceil(rg)
  int link;            // link index (disable mask, diagnostics); -1: pad slot
  int pad;
};

struct GridDev {
  const uint32_t *occ;   // bit-packed rows along x, row_words words per (j,k) row (>= 1 zero pad word)
  // Feature transform of the occupancy, [gz][gy][gx] each.  For cell C let O* be the occupied cell that minimises the
  // squared box-to-box distance bb2(C,O) = sum_a max(|o_a|-1,0)^2 (o = O - C; ties: smaller centre distance):
  const uint8_t *lb2;    //   min(bb2(C,O*), 255); 255 also when nothing is occupied within the build window
  const uint16_t *ft;    //   (ox+15) | (oy+15)<<5 | (oz+15)<<10 of O*, 0xFFFF when none
  // lb2 and ft live in a domain padded by `halo` cells on every side (halo >= reach of any sphere centre from the
  // base reference + 2): a state whose base lies inside the grid can index them for every sphere with no bounds
  // test.  lb2 / ft point at in-grid cell (0,0,0); cell (i,j,k) is at [k * splane + j * sx + i], i in [-halo, gx+halo).
  // lb2_base / ft_base are the allocation starts = cell (-halo,-halo,-halo); org_off = offset of (0,0,0) from them.
  const uint8_t *lb2_base;
  const uint16_t *ft_base;
  unsigned sx, splane, org_off;
  unsigned zero_off;     // lb2_base[zero_off] == 0: where the fast kernel points the gathers of lanes that must not be pruned
  // Slab fields (fast kernel, yaw-only base group): for a set of z-layers [k0, k1] the 2-D squared box-to-box distance
  // (cells, capped at 255) from column (i, j) to the nearest column that holds an occupied cell in those layers.  They
  // follow lb2 in the same allocation: slab s, column (i, j) at lb2_base[slab_off + s * splane2 + (j + halo) * sx + i + halo].
  unsigned slab_off, splane2;
  int n_slabs;
  int halo;
  int gx, gy, gz, row_words;
  int has_grid;
  double ox, oy, oz, res, inv_res;
};

// Fused mask all-gather (validity_fast.cu): instead of writing its mask words to one local buffer, a rank's kernel
// writes them straight into the gather buffer of EVERY rank of the node through peer-mapped pointers (NVLink stores);
// a one-warp signal kernel ordered after it publishes the step number in every rank's flag array.  n == 0: plain local store.
struct PeerDev {
  uint32_t *mask[SQMP_MAX_PEERS];   // every rank's gather buffer (identical layout), as mapped into this process
  uint32_t *flag[SQMP_MAX_PEERS];   // every rank's flag array [n]: flag[p][rank] = step once this rank's words of `step` have landed
  int64_t word_offset;              // first word of this rank's slot inside a gather buffer
  int n, rank;
  unsigned step;
};

// ---- tables of the fast validity kernel (validity_fast.cu) ----------------------------------------------------------
// A rigid group is a two-level hierarchy: root bound -> cluster bounds -> leaf spheres.  Bounds and leaves are "nodes"
// tested the same way (stage 1: one byte gather from a lower-bound field, `field > thr` proves the node free); a level
// is entered only by the lanes whose node above is not provably free.
// Node records come in pairs laid out for packed fp32x2 arithmetic — two consecutive float4:
//   (x0 x1 y0 y1) (z0 z1 thr0 thr1)       3-D node: centre in its group frame, grid units; thr as int bits
//   (x0 x1 y0 y1) (so0 so1 thr0 thr1)     2-D node (yaw-only base group): so = byte offset of the node's slab field
//                                         from slab field 0 (uint bits) — its z never changes, so its layers are fixed
constexpr int KF_MAX_SLOTS = 1024;
constexpr int KF_MAX_BOUNDS = 128;
struct KfTab {   // by value in the kernel parameters: the lockstep bound loops read it with warp-uniform indices
  float4 bound[KF_MAX_BOUNDS];         // pair records of the cluster bounds (groups padded to a multiple of 8 bounds)
  short bound_first[KF_MAX_BOUNDS];    // first leaf slot of the cluster
  unsigned char bound_count[KF_MAX_BOUNDS];   // its slots (4 or 8)
  float4 root[SQMP_MAX_GROUPS];        // x y (z | slab offset bits) thr bits: bound of the whole group
  short grp_bound_end[SQMP_MAX_GROUPS + 1];   // bounds of group g are [end[g], end[g+1])
  int base2d_group;                    // group tested against the slab fields (-1: none)
};
struct FastAux {       // per leaf slot, global memory: what the rare stages need (divergent indices)
  float x, y, z;       // centre in the group frame, grid units, fp32 (same values as the pair record; z also for 2-D nodes)
  float r2_lo;         // (rg - err)^2 rounded down: ub2 <= r2_lo proves a hit even after the fp32 centre error
  int ncell;           // ceil(rg) + 1 (reach test of a sphere whose state lies outside the comfortable region)
  int pad[3];
};
struct FastDev {
  const float4 *leafrec;     // [n_slots]: pair records of the leaf slots (read per queued (cluster, state) item: divergent indices)
  const FastAux *aux;        // [n_slots]
  const double4 *sph64;      // [n_slots]: x y z r, grid units, fp64 (exact stage)
  int n_slots;
  int ok;                    // the description and the grid satisfy the fast kernel's preconditions
};

// ---- self collision (CC:541-552) -------------------------------------------------------------------------------------
// Nested bounds over the enabled link pairs whose links hang on different rigid groups: group pair -> link of A vs the
// bound of group B -> link pair -> cluster of B vs link A -> cluster pair -> leaf of B vs cluster of A -> leaf pair.  A level is entered only if the level
// above cannot rule it out, so a state far from self contact costs ~one test per link pair.  All centres are fp32 in
// their group frame, grid units; bounds are padded against fp32 rounding (capi.cu: build_self_tables).
struct SelfLeaf {   // 16 bytes
  float x, y, z, r;
};
struct SelfCluster {   // 48 bytes: bounding sphere (its role on the B side), bounding box (its role on the A side), leaf range
  float x, y, z, R;
  float lo[3], hi[3];     // axis-aligned box of its spheres in the group frame, padded like R: flat shapes (the shell box,
                          // the drum) fill a box far better than a sphere, and the A side is tested as constants anyway
  int leaf_begin, leaf_end;
};
struct SelfLink {   // 32 bytes: bounding sphere of a link + its clusters
  float x, y, z, R;
  int cl_begin, cl_end;
  int link, pad;
};
struct SelfALink {   // 48 bytes: a link of group A within one group pair, with the B-side links it is checked against
  float lo[3], hi[3];     // the link's bounding box (A frame), padded against fp32 rounding
  int lp_begin, lp_end;   // its partners in SelfDev::lpairs
  int a;                  // its record in SelfDev::links (clusters)
  int pad[3];
};
struct SelfGroupPair {   // 48 bytes
  int fa, fb;             // frame slots of the two groups (RobotDev::grp_fidx)
  int al_begin, al_end;   // the A-side links that take part (SelfDev::alinks)
  float ax, ay, az, RA;   // bound of the A-side links that take part
  float bx, by, bz, RB;   // ... B side
};
struct SelfDev {
  const SelfGroupPair *gpairs;
  const SelfALink *alinks;
  const int *lpairs;          // B-side link (index into links) of every enabled link pair, grouped by A-side link
  const SelfLink *links;
  const SelfCluster *clusters;
  const SelfLeaf *leaves;
  int n_gpairs, n_lpairs, n_links;
};

// packed fp32x2 (sm_100 FFMA2 / FADD2): two spheres per instruction; a plain float operand is broadcast (R.F32)
__device__ __forceinline__ float2 f32x2_fma(float a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 f32x2_add(float2 a, float b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b));
  return d;
}

// ---- launch helpers (defined in the .cu files) ---------------------------------------------------------
struct Ctx;

cudaError_t launch_check_states(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const K1Slot *d_slots,
                                const double *d_q, const double *d_q_end, int num_points, int64_t S,
                                unsigned flags, uint32_t *d_mask, cudaStream_t st, int *n_launch);
cudaError_t launch_first_invalid(const uint32_t *d_mask, int64_t E, int P1, int32_t *d_first, cudaStream_t st);
cudaError_t launch_get_collisions(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const K1Slot *d_slots,
                                  const double *d_q, uint8_t *d_slot_hit, uint8_t *d_link_pair_hit, cudaStream_t st);
cudaError_t launch_fk(const RobotDev &rb, const double *d_q, int64_t S, double *d_out, int fk_precision,
                      bool allow_specialized, cudaStream_t st);
cudaError_t launch_sample(const RobotDev &rb, uint64_t seed, int64_t first, int64_t n, const double *lo, const double *hi,
                          int max_tries, double *d_q_out, int32_t *d_tries_out, bool allow_specialized, cudaStream_t st);
cudaError_t launch_sample_informed(const RobotDev &rb, uint64_t seed, int64_t first, int64_t n, const double *lo, const double *hi,
                                   int max_tries, const sqmp_informed &el, double *d_q_out, int32_t *d_tries_out,
                                   bool allow_specialized, cudaStream_t st);
cudaError_t launch_slab_build(const uint32_t *d_occ, int row_words, int gx, int gy, int gz, int halo, int window,
                              int n_slabs, const int *h_k0, const int *h_k1, uint32_t *d_col, int8_t *d_tmp, uint8_t *d_slab,
                              cudaStream_t st, int *n_launch);
cudaError_t launch_ft_build(const uint32_t *d_occ, int row_words, int gx, int gy, int gz, int halo, int window,
                            int8_t *d_tmp_a, uint16_t *d_tmp_b, uint8_t *d_lb2, uint16_t *d_ft, cudaStream_t st,
                            int *n_launch);
cudaError_t launch_check_states_fast(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const FastDev &fd,
                                     const KfTab &tab, const double *d_q, const double *d_q_end, int num_points, int64_t S,
                                     unsigned flags, uint32_t *d_mask, bool allow_specialized, bool overlap,
                                     cudaStream_t st, int *n_launch, const PeerDev *peers = nullptr,
                                     int32_t *d_first_invalid = nullptr);
cudaError_t launch_gather_signal(const PeerDev &peer, cudaStream_t st);
cudaError_t launch_gather_wait(const uint32_t *d_flags, int n, unsigned step, int32_t *d_status, cudaStream_t st);
cudaError_t launch_pack_rows(const uint32_t *d_bits_linear, int gx, int gy, int gz, int row_words, uint32_t *d_occ,
                             cudaStream_t st);

constexpr int SQMP_NN_TILE_IMG_BYTES = 6144;
struct TreeDev {
  double *q[SQMP_DOF];   // SoA: q[d][n]
  double *cost;
  double *aos;           // the same nodes row-major [n][8]: one 64-byte line per node for the exact re-evaluation of the
                         // few pairs that survive the fp32 pre-filter (eight SoA columns cost eight L2 round trips)
  // fp32 shadow for the sound pre-filter of K3/K4: fa[n] = (q0..q3), fb[n] = (q4..q7) rounded to float,
  // fn[n] = |float node|^2 (fp64 sum rounded once)
  float4 *fa, *fb;
  float *fn;
  // fp16 operand image of the nodes for the tensor-core scan (tc_umma.cuh): one SQMP_NN_TILE_IMG_BYTES record per 128-node tile
  float4 *img;
  double rmax2;          // upper bound of |node|^2 over every node written since the last reset (host-tracked)
  int64_t n, cap;
};
cudaError_t launch_tree_append(TreeDev &t, const double *d_q_rows, const double *d_cost, int64_t first, int64_t n,
                               cudaStream_t st);
cudaError_t launch_tree_set_one(TreeDev &t, int64_t id, const double q[SQMP_DOF], double cost, cudaStream_t st);
cudaError_t launch_tree_update_costs(TreeDev &t, const int32_t *d_ids, const double *d_cost, int64_t n,
                                     cudaStream_t st);
// deferred single-node writes (nn.cu: k_tree_write_many): records of tree_write_bytes() each, readable by the device
size_t tree_write_bytes();
void tree_write_fill(void *dst, const TreeDev &t, int64_t id, const double q[SQMP_DOF], double cost);
bool tree_write_same_target(const void *rec, const TreeDev &t, int64_t id);
cudaError_t launch_tree_write_many(const void *d_records, int n, cudaStream_t st);

struct NNScratch {
  double *part_d;      // [chunks][Q]
  int32_t *part_id;    // [chunks][Q]
  int32_t *hits;       // [Q][hit_cap]
  int64_t hit_cap;
};
// variant: 1 QT fp64-only, 2 WQ fp64-only, 3 QT with fp32 pre-filter, 4 tensor-core pre-filter (default for many queries)
cudaError_t launch_nn_nearest(const TreeDev &t, const double *d_queries, int64_t Q, int32_t *d_id, NNScratch &sc,
                              int chunks, int variant, cudaStream_t st, int *n_launch);
// one query, one launch, result in mapped pinned host memory (nn.cu: k_nn_single)
cudaError_t launch_nn_single(const TreeDev &t, const double *h_query, bool radius, int32_t excl, double s_thr, int k,
                             NNScratch &sc, unsigned *d_ticket, int32_t *d_hit_count, int32_t *d_out_mapped, int sm_count,
                             cudaStream_t st);
cudaError_t launch_nn_radius(const TreeDev &t, const double *d_queries, const int32_t *d_exclude, int64_t Q,
                             double s_thr, double r, int k, int32_t *d_low, int32_t *d_high, int32_t *d_count,
                             int32_t *d_overflow, NNScratch &sc, int chunks, int variant, cudaStream_t st,
                             int *n_launch);

// every query against its own tree (nn.cu: k_nn_multi); d_desc: one 96-byte record per query (nn_multi_fill_desc)
cudaError_t launch_nn_multi(const void *d_desc, const double *d_queries, int64_t Q, bool radius, double s_thr, int k,
                            NNScratch &sc, int32_t *d_out, cudaStream_t st);
int nn_tc_queries_per_cta();   // nn.cu: queries a CTA of the tensor-core scan handles
size_t nn_multi_desc_bytes();
void nn_multi_fill_desc(void *dst, const TreeDev &t, int32_t exclude);

}  // namespace sqmp
