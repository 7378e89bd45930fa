// sqmp_common.cuh — device-side descriptors shared by the kernels of libsqmp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sqmp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsqmp_b200 is written for sm_100a (B200) only"
#endif

namespace sqmp {

// K1 walks "slots": the spheres of every rigid group padded to a multiple of 4 (pads repeat the last sphere and can
// never test positive), so its 4-wide batches need no bounds logic.
constexpr int SQMP_MAX_SLOTS = SQMP_MAX_SPHERES + 4 * (SQMP_MAX_SEG + 1);

// axis codes: the joint axes of the reference robot are all signed unit axes (URDF:30-66, 615-759);
// a general axis takes the Rodrigues path.
__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

enum AxisCode : int { AX_GENERAL = 0, AX_PX = 1, AX_NX = 2, AX_PY = 3, AX_NY = 4, AX_PZ = 5, AX_NZ = 6 };

// Chain + sphere table, passed by value as a __grid_constant__ kernel parameter: every access is
// warp-uniform, so it is served from the constant bank / uniform datapath and never touches the LSU.
struct RobotDev {
  int n_seg;
  int n_spheres;                          // metres copy: spheres; grid-unit copy (K1): slots, see SQMP_MAX_SLOTS
  int ref_seg;                            // last segment of the leading prismatic run (base position)
  int self_const_collision;
  unsigned trans_mask;                    // bit j: q[j] drives a prismatic joint (a length)
  double len_scale;                       // 1 (metres) or 1/res (grid units, the copy K1 runs on)
  int seg_type[SQMP_MAX_SEG];
  int seg_qidx[SQMP_MAX_SEG];
  int seg_axis_code[SQMP_MAX_SEG];
  int seg_pre_kind[SQMP_MAX_SEG];         // 0 identity, 1 translation only, 2 general
  int seg_sph_end[SQMP_MAX_SEG + 2];      // spheres of segment s (-1 = root) are [end[s+1-1], end[s+1])
  int grp_fidx[SQMP_MAX_SEG + 1];         // group (segment + 1) -> slot of its frame in the self-collision scratch, -1: none
  int n_frames;
  double seg_pre[SQMP_MAX_SEG][12];
  double seg_axis[SQMP_MAX_SEG][3];
  double root_off[3];
  double ee_tip[12];
  int ee_tip_kind;
  double sph[SQMP_MAX_SLOTS][4];        // cx cy cz r in the segment frame
  // fp32 copies of the constants (made by the host once; saves the F2F conversions in the float chain)
  float seg_pre_f[SQMP_MAX_SEG][12];
  float seg_axis_f[SQMP_MAX_SEG][4];
  float ee_tip_f[12];
  float4 sph_f[SQMP_MAX_SLOTS];
  // per-grid derived values (recomputed by sqmp_set_voxel_grid)
  short sph_thr[SQMP_MAX_SLOTS];             // floor((r/res)^2): "lb2 > thr" proves the sphere free (integers, exact);
                                              // -2 while the sphere's link is disabled (sqmp_set_disabled_links)
  unsigned char sph_ncell[SQMP_MAX_SLOTS];  // ceil(r / res)
  // per-request (sqmp_set_disabled_links)
  uint32_t sph_enabled[SQMP_MAX_SPHERES / 32];
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

// Tables of the fast validity kernel (validity_fast.cu), in global memory; rebuilt with the grid-unit robot.
struct FastPair {      // two consecutive slots (slots come in multiples of 4), laid out for packed fp32x2 arithmetic;
  float x[2], y[2], z[2];   // sphere centres in their group frame, grid units, fp32
  int thr[2];          // stage-1 threshold: lb2 > thr proves the sphere free even after the fp32 centre error; -2: never test
};
struct FastAux {       // per slot, rare path
  float r2_lo;         // (rg - err)^2 rounded down: ub2 <= r2_lo proves a hit even after the fp32 centre error
  int ncell;           // ceil(rg) + 1 (reach test of a sphere whose state lies outside the comfortable region)
};
struct FastDev {
  const FastPair *pairs;     // [n_slots / 2], staged in shared memory by every CTA
  const FastAux *aux;        // [n_slots], staged too
  const double4 *sph64;      // [n_slots]: x y z r, grid units, fp64 (exact stage)
  int n_slots;
  int ok;                    // the description and the grid satisfy the fast kernel's preconditions
};

// Self collision (CC:541-552) as a three-level hierarchy over the enabled sphere pairs, all between different rigid
// groups: block = (group A, group B) with one bounding sphere per side; per block the distinct A-spheres, each with
// the bounding sphere of its partners; per A-sphere its partner list.  A level is entered only if the level above
// cannot rule it out, so a state far from self contact costs ~one test per block instead of one per pair.
struct SelfPair {   // 8 bytes
  unsigned char a, b;     // sphere slots
  unsigned char fa, fb;   // frame slots of their groups (RobotDev::grp_fidx)
  float rsum2;            // (ra + rb)^2, grid units
};
struct SelfA {      // 32 bytes: everything one A-test needs in one record (no dependent second load)
  float cx, cy, cz;       // the A-sphere's centre in its group frame, grid units
  float thr2;             // (ra + RB)^2
  int p_begin, p_end;     // its partners in SelfDev::hpairs
  int a;                  // sphere slot
  int pad;
};
struct SelfHPair {  // 16 bytes: a partner of an A-sphere, hierarchical order (block, a, b)
  float bx, by, bz;       // the B-sphere's centre in its group frame, grid units
  float rsum2;            // (ra + rb)^2
};
struct SelfBlock {  // 48 bytes
  int fa, fb;             // frame slots of the two groups
  int a_begin, a_end;     // its A-spheres in SelfDev::as
  float ca[3], cb[3];     // bounding-sphere centres in the group frames, grid units
  float thr2;             // (RA + RB)^2
  float pad;
};
struct SelfDev {
  const SelfHPair *hpairs;  // hierarchical order (block, a, b)
  const SelfA *as;
  const SelfBlock *blocks;
  const SelfPair *flat;     // the description's own pair order (diagnostics: one result per pair)
  int n_pairs, n_as, n_blocks;
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

cudaError_t launch_check_states(const RobotDev &rb, const GridDev &gd, const SelfDev &self,
                                const double *d_q, const double *d_q_end, int num_points, int64_t S,
                                unsigned flags, uint32_t *d_mask, int fk_precision, cudaStream_t st, int *n_launch);
cudaError_t launch_first_invalid(const uint32_t *d_mask, int64_t E, int P1, int32_t *d_first, cudaStream_t st);
cudaError_t launch_get_collisions(const RobotDev &rb, const GridDev &gd, const SelfDev &self,
                                  const double *d_q, uint8_t *d_sphere_hit, uint8_t *d_pair_hit, cudaStream_t st);
cudaError_t launch_fk(const RobotDev &rb, const double *d_q, int64_t S, double *d_out, int fk_precision,
                      bool allow_specialized, cudaStream_t st);
cudaError_t launch_sample(const RobotDev &rb, uint64_t seed, int64_t first, int64_t n, const double *lo, const double *hi,
                          int max_tries, double *d_q_out, int32_t *d_tries_out, bool allow_specialized, cudaStream_t st);
cudaError_t launch_ft_build(const uint32_t *d_occ, int row_words, int gx, int gy, int gz, int halo, int window,
                            int8_t *d_tmp_a, uint16_t *d_tmp_b, uint8_t *d_lb2, uint16_t *d_ft, cudaStream_t st,
                            int *n_launch);
cudaError_t launch_check_states_fast(const RobotDev &rb, const GridDev &gd, const SelfDev &self, const FastDev &fd,
                                     const double *d_q, const double *d_q_end, int num_points, int64_t S,
                                     unsigned flags, uint32_t *d_mask, bool allow_specialized, bool overlap,
                                     cudaStream_t st, int *n_launch, const PeerDev *peers = nullptr);
cudaError_t launch_gather_signal(const PeerDev &peer, cudaStream_t st);
cudaError_t launch_gather_wait(const uint32_t *d_flags, int n, unsigned step, int32_t *d_status, cudaStream_t st);
cudaError_t launch_pack_rows(const uint32_t *d_bits_linear, int gx, int gy, int gz, int row_words, uint32_t *d_occ,
                             cudaStream_t st);

struct TreeDev {
  double *q[SQMP_DOF];   // SoA: q[d][n]
  double *cost;
  // fp32 shadow for the sound pre-filter of K3/K4: fa[n] = (q0..q3), fb[n] = (q4..q7) rounded to float,
  // fn[n] = |float node|^2 (fp64 sum rounded once)
  float4 *fa, *fb;
  float *fn;
  double rmax2;          // upper bound of |node|^2 over every node written since the last reset (host-tracked)
  int64_t n, cap;
};
cudaError_t launch_tree_append(TreeDev &t, const double *d_q_rows, const double *d_cost, int64_t first, int64_t n,
                               cudaStream_t st);
cudaError_t launch_tree_set_one(TreeDev &t, int64_t id, const double q[SQMP_DOF], double cost, cudaStream_t st);
cudaError_t launch_tree_update_costs(TreeDev &t, const int32_t *d_ids, const double *d_cost, int64_t n,
                                     cudaStream_t st);

struct NNScratch {
  double *part_d;      // [chunks][Q]
  int32_t *part_id;    // [chunks][Q]
  int32_t *hits;       // [Q][hit_cap]
  int64_t hit_cap;
};
// variant: 1 QT fp64-only, 2 WQ fp64-only, 3 QT with fp32 pre-filter (default for many queries)
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
size_t nn_multi_desc_bytes();
void nn_multi_fill_desc(void *dst, const TreeDev &t, int32_t exclude);

}  // namespace sqmp
