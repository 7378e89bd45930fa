/*
 * sqmp_b200.h — C ABI of libsqmp_b200.so: the state-validity (FK + collision) and near-neighbour
 * inner loop of squirrel_motion_planner's birrt_star / squirrel_8dof_planner on NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI; its boundary for this path is a set of C++ member functions.  Each entry
 * point below names the reference interface it sits behind (file:line in /root/reference):
 *   CC = birrt_star/robot_motion_planning/birrt_star_algorithm/include/birrt_star_algorithm/collision_checker.hpp
 *   BS = birrt_star/robot_motion_planning/birrt_star_algorithm/src/birrt_star.cpp
 *   BH = .../birrt_star_algorithm/include/birrt_star_algorithm/birrt_star.h
 *   DH = birrt_star/robot_motion_planning/planning_heuristics/src/distance_heuristics.cpp
 *   DS = birrt_star/robot_motion_planning/planner_data_structures/include/planner_data_structures/data_structs.h
 *   KM = birrt_star/robot_motion_control/kuka_motion_control/src/kdl_kuka_model.cpp
 *
 * Conventions: every call is synchronous from the planner thread (the reference's checker is not
 * re-entrant either, CC:141-142); int return 0 = OK, <0 = error (sqmp_strerror); no exception crosses
 * the boundary; the caller owns all host buffers, the library owns all device memory; one context per
 * GPU.  There is no CPU fallback: without a usable CUDA device sqmp_create fails.
 * All joint vectors have SQMP_DOF = 8 doubles: x, y, theta, arm joints 1..5 (BS:6868-6870 hard-codes 8).
 */
#ifndef SQMP_B200_H
#define SQMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQMP_DOF 8
#define SQMP_MAX_SEG 8
#define SQMP_MAX_SPHERES 1024  /* a covering model of the reference's 27 link shapes needs ~330 (sphere_model.py) */
#define SQMP_MAX_LINKS 64
#define SQMP_MAX_PEERS 8       /* GPUs of one node in the fused mask all-gather */
#define SQMP_MAX_TREES 1024   /* two per concurrent planning query (shim/sqmp_multi.hpp) */

#define SQMP_JOINT_ROT 1
#define SQMP_JOINT_TRANS 2

/* flags of the validity calls: the two bools of isInCollision(q, checkSelf, checkMap) (CC:104) */
#define SQMP_CHECK_SELF 1u
#define SQMP_CHECK_MAP 2u

#define SQMP_OK 0
#define SQMP_ERR_INVALID (-1)   /* bad argument */
#define SQMP_ERR_CUDA (-2)      /* CUDA runtime error, see sqmp_last_error */
#define SQMP_ERR_NOMEM (-3)
#define SQMP_ERR_OVERFLOW (-4)  /* radius search: more hits than the per-query buffer even at batch size 1 */
#define SQMP_ERR_NODEVICE (-5)

typedef struct sqmp_ctx sqmp_ctx;

/*
 * Robot description = what CollisionChecker builds at start-up (CC:176-351, 353-393), pre-folded by the
 * host (squirrel_motion_planner_b200/robot_desc.py: compile_robot) into a serial chain of planning
 * segments with rigid sphere groups.  Frames are 12 doubles: R row-major (9) then p (3).
 *   world frame of segment k:  T_k = T_{k-1} * seg_pre[k] * Motion(seg_axis[k], q[seg_qidx[k]])
 *   T_{-1} = Frame(root_offset) for the collision path (CC:199-201), identity for sqmp_fk (KM:278).
 */
typedef struct sqmp_robot_desc {
  int32_t n_seg;
  int32_t seg_type[SQMP_MAX_SEG];      /* SQMP_JOINT_ROT | SQMP_JOINT_TRANS */
  int32_t seg_qidx[SQMP_MAX_SEG];      /* index into the 8-vector */
  double seg_pre[SQMP_MAX_SEG][12];    /* fixed frame before the joint motion */
  double seg_axis[SQMP_MAX_SEG][3];    /* unit axis in the joint frame */
  double root_offset[3];
  double ee_tip[12];                   /* last segment -> FK tip link (SRDF chain tip, hand_wrist_link) */
  int32_t n_spheres;
  const int32_t *sph_seg;              /* [n] segment carrying the sphere, -1 = root; sorted ascending */
  const int32_t *sph_link;             /* [n] link index 0..63 (bit position in the disable mask) */
  const double *sph_center;            /* [n*3] in the segment frame */
  const double *sph_radius;            /* [n] */
  /* selfCollisionPairs (CC:353-393) at link level: bit j of link_pair_mask[i] = links i and j are checked against
   * each other (both shaped, not disabled in the SRDF; symmetric).  The library collides every sphere of link i with
   * every sphere of link j when the two links hang on different segments. */
  uint64_t link_pair_mask[SQMP_MAX_LINKS];
  int32_t self_const_collision;        /* 1: an enabled pair inside one rigid group always collides */
} sqmp_robot_desc;

/* ---- life cycle ------------------------------------------------------------------------------- */
/* CollisionChecker::CollisionChecker()/initialize() (CC:66-74, 156-172).  NULL on failure
 * (*err_out receives the code); the shim then fails closed like CC:106-107. */
sqmp_ctx *sqmp_create(int device, const sqmp_robot_desc *desc, int *err_out);
void sqmp_destroy(sqmp_ctx *ctx);
const char *sqmp_last_error(const sqmp_ctx *ctx);
const char *sqmp_strerror(int code);
/* Build tag + compiled SASS architecture ("sm_100a"). */
const char *sqmp_version(void);

/* ---- map -------------------------------------------------------------------------------------- */
/* CollisionChecker::setOcTree (CC:76-88) on a dense copy of the octomap: bit (k*gy+j)*gx+i of `bits`
 * is voxel (i,j,k) = origin + [i,i+1] x [j,j+1] x [k,k+1] * res; 1 = occupied, unknown = free.  The
 * caller folds the octree's z = -0.02 offset (CC:87) into origin.  NULL bits is ignored (CC:78-79).
 * Deep-copies, replaces the previous grid, and rebuilds the distance field used as a conservative
 * filter.  Before the first call every map check reports "free" (CC:556-557). */
int sqmp_set_voxel_grid(sqmp_ctx *ctx, const uint32_t *bits, int gx, int gy, int gz, const double origin[3],
                        double res);
/* CollisionChecker::setDisabledLinkMapCollisions (CC:90-102): bit i = link i is skipped by the map check. */
/* Second and further contexts on the SAME device (one planner thread each, e.g. shim/sqmp_plan_multi --threads) can use
 * the voxel fields of a context that already holds the grid instead of building their own copy: dst references src's
 * device memory (src keeps ownership and must outlive dst or dst's next sqmp_set_voxel_grid).  Both contexts must
 * have been created from the same robot description (the halo of the fields depends on it). */
int sqmp_share_voxel_grid(sqmp_ctx *dst, const sqmp_ctx *src);
int sqmp_set_disabled_links(sqmp_ctx *ctx, uint64_t link_mask);

/* ---- state validity ---------------------------------------------------------------------------- */
/* CollisionChecker::isInCollision (CC:104-121) over S states; BiRRTstarPlanner::isConfigValid
 * (BS:6897-6908).  q: S*8 row-major.  mask_out: ceil(S/32) words, bit s%32 of word s/32 = 1 ⇒ state s
 * is in collision.  flags == 0 ⇒ all zero (CC:109-110). */
int sqmp_check_states(sqmp_ctx *ctx, const double *q, int64_t S, unsigned flags, uint32_t *mask_out);
/* connectNodesInterpolation/interpolateConfigurations (BS:4380-4490) + isEdgeValid (BS:6864-6895) over E
 * edges: state k of edge e is q_start + k*((q_end-q_start)/num_points), k = 0..num_points (BS:4461-4478).
 * mask_out: bit e*(num_points+1)+k.  first_invalid[e]: first colliding k or -1 (may be NULL);
 * isEdgeValid's last_valid_node_idx is max(first_invalid-1, 0), or num_points when valid. */
int sqmp_check_edges(sqmp_ctx *ctx, const double *q_start, const double *q_end, int64_t E, int num_points,
                     unsigned flags, uint32_t *mask_out, int32_t *first_invalid);
/* CollisionChecker::getCollisions (CC:123-133, 594-631) for one state: link_map_hit[64] (1 = link hits
 * the map; like the reference, ignores the disable mask) and link_pair_hit[64 * 64] (may be NULL): entry [i * 64 + j],
 * i < j, is 1 when the enabled link pair (i, j) collides. */
int sqmp_get_collisions(sqmp_ctx *ctx, const double q[SQMP_DOF], uint8_t *link_map_hit, uint8_t *link_pair_hit);

/* ---- forward kinematics ------------------------------------------------------------------------- */
/* KDLRobotModel::compute_FK (KM:278-305): S*7 doubles x y z qx qy qz qw of the chain tip. */
int sqmp_fk(sqmp_ctx *ctx, const double *q, int64_t S, double *ee_pose_out);

/* ---- sampling ----------------------------------------------------------------- */
/* BiRRTstarPlanner::sampleJointConfig_JntArray (BS:3832-3878) over RobotController::getRandomConf (CL:1120-1188):
 * every joint uniform in [lo, hi] (the base from the environment size, CL:1153-1166), redrawn until the end-effector of
 * compute_FK lies above the platform (0.0 <= ee_z, CL:1182-1184 and BS:3870-3871).  The reference draws from a
 * time-seeded Mersenne twister; here sample i is a pure function of (seed, first + i): try t of sample i uses the
 * outputs 8 * (i * max_tries + t) + j, j = 0..7, of the splitmix64 stream started at `seed`, so a batch is
 * reproducible, order-free and identical on the CPU oracle.  q_out: n*8.  tries_out (may be NULL): tries used, or
 * -max_tries when every try was below the platform (the last try is returned). */
int sqmp_sample_configs(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                        const double hi[SQMP_DOF], int max_tries, double *q_out, int32_t *tries_out);

/* BiRRTstarPlanner::sampleJointConfigfromEllipse_JntArray (BS:3607-3829): the informed sampler the planner switches to
 * once a solution exists (BS:1168).  The ellipse comes from jointConfigEllipseInitialization (BS:3472-3604), per joint
 * class (revolute / prismatic): centre = midpoint of start and goal, rotation C with first column (goal - start) / c_min,
 * and the diagonal L = (c_best / 2, sqrt(c_best^2 - c_min^2) / 2, ...) of BS:3709-3752 (the caller refreshes l_rev /
 * l_pri whenever the best solution cost changes).  One try: a joint vector uniform in [lo, hi] (getRandomConf), split by
 * class, each part normalised and scaled by ONE uniform number in [0, 1); x = C (L ball) + centre, written back prismatic
 * joints first, then revolute ones (BS:3773-3785); redrawn until the end-effector lies above the platform (0.0 <= ee_z) and
 * the first two joints lie strictly inside env_x / env_y (all four zero: unconfined; BS:3800-3806).  Try t of sample i
 * uses outputs 9 * ((first + i) * max_tries + t) + j, j = 0..8 (j = 8: the scale), of the splitmix64 stream of
 * sqmp_sample_configs.  Products and sums are evaluated in index order with round-to-nearest operations, so the device
 * and the CPU oracle produce the same doubles. */
typedef struct sqmp_informed {
  int32_t n_rev, n_pri;              /* joints per class (n_rev + n_pri = 8) */
  int32_t rev_idx[SQMP_DOF], pri_idx[SQMP_DOF];   /* which entries of the joint vector belong to the class, ascending */
  double c_rev[SQMP_DOF * SQMP_DOF], c_pri[SQMP_DOF * SQMP_DOF];   /* rotation C, row-major n x n */
  double center_rev[SQMP_DOF], center_pri[SQMP_DOF];
  double l_rev[2], l_pri[2];         /* L = diag(l[0], l[1], ..., l[1]) */
  double env_x[2], env_y[2];
} sqmp_informed;
int sqmp_sample_informed(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                         const double hi[SQMP_DOF], int max_tries, const sqmp_informed *ellipse, double *q_out,
                         int32_t *tries_out);

/* ---- tree mirror (DS:29-54; mutation sites BS:3298-3322, 5173-5193, 5518-5520) ------------------- */
/* Invariant kept from the reference: node_id == index.  tree in [0, SQMP_MAX_TREES). */
/* sqmp_tree_append / sqmp_tree_update cost a record in a host-side queue, not a launch: the queue reaches the device in
 * ONE launch before the next call that reads a tree (any sqmp_nn_*, the batch writers, reset, growth).  Program order is
 * kept — a later write to the same node replaces the queued one — so callers see exactly the semantics of immediate
 * writes (the reference's push_back / in-place assignment at BS:3298-3322, 5173-5193). */
int sqmp_tree_reset(sqmp_ctx *ctx, int tree);
int sqmp_tree_append(sqmp_ctx *ctx, int tree, const double q[SQMP_DOF], double cost, int32_t *id_out);
int sqmp_tree_append_batch(sqmp_ctx *ctx, int tree, const double *q, const double *cost, int64_t n,
                           int32_t *first_id_out);
int sqmp_tree_update(sqmp_ctx *ctx, int tree, int32_t id, const double q[SQMP_DOF], double cost);
int sqmp_tree_update_costs(sqmp_ctx *ctx, int tree, const int32_t *ids, const double *cost, int64_t n);
int sqmp_tree_size(sqmp_ctx *ctx, int tree, int64_t *n_out);

/* ---- near neighbours ---------------------------------------------------------------------------- */
/* find_nearest_neighbour_interpolation (BS:4076-4133) with euclidean_joint_space_distance (DH:128-156):
 * argmin_n sqrt(sum_i (node_n[i]-query[i])^2), fp64, summed in index order without FMA contraction,
 * strict '<' in serial order (lowest id wins ties), min initialised to 10000.0 and id to 0. */
int sqmp_nn_nearest(sqmp_ctx *ctx, int tree, const double *queries, int64_t Q, int32_t *id_out);
/* find_near_vertices_interpolation (BS:4272-4324): {n : d(n, query) < r and n != exclude_id[q]} ordered by
 * (cost asc, id asc) — DS:41-44 orders by cost; the id tie-break makes std::sort's unstable order
 * deterministic.  low_ids[q*k + t]: t-th lowest (choose-parent / connect walk, BS:4618-4623, 2823);
 * high_ids[q*k + t]: t-th highest (rewire walk from the end, BS:5075-5089); -1 padded.
 * count_out[q]: number of near vertices.  exclude_id may be NULL (no exclusion). */
int sqmp_nn_radius(sqmp_ctx *ctx, int tree, const double *queries, const int32_t *exclude_id, int64_t Q, double r,
                   int k, int32_t *low_ids, int32_t *high_ids, int32_t *count_out);

/* Many queries, each against ITS OWN tree (query i searches tree[i]) — the call shape of the lock-step multi-query
 * planner (shim/sqmp_multi.hpp), where every planning query of a batch owns two small trees.  Same semantics and the
 * same bit-exact results as Q single calls of sqmp_nn_nearest / sqmp_nn_radius; one launch per call. */
int sqmp_nn_nearest_multi(sqmp_ctx *ctx, const int32_t *tree, const double *queries, int64_t Q, int32_t *id_out);
int sqmp_nn_radius_multi(sqmp_ctx *ctx, const int32_t *tree, const double *queries, const int32_t *exclude_id, int64_t Q,
                         double r, int k, int32_t *low_ids, int32_t *high_ids, int32_t *count_out);

/* ---- device-pointer variants (benchmark harness, multi-GPU sharding) ---------------------------- */
/* Same semantics; all pointers are device pointers on the context's GPU; work is enqueued on `stream`
 * (a cudaStream_t, NULL = default stream) and NOT synchronised. */
/* Ordering of the validity *_dev calls: plain stream order by default.  sqmp_set_option(ctx, "launch_overlap", 1) is an
 * opt-in for callers whose input states are complete before the call is enqueued (bench.py, the host-buffer entry
 * points of this library): consecutive validity grids on one stream may then overlap (a grid starts while its
 * predecessor drains; it writes its mask only after the predecessor has completed, so reusing one mask buffer stays
 * safe), but a grid reads its states BEFORE that point — a kernel enqueued immediately before the call on the same
 * stream must therefore not be the producer of those states. */
int sqmp_check_states_dev(sqmp_ctx *ctx, const double *d_q, int64_t S, unsigned flags, uint32_t *d_mask,
                          void *stream);
int sqmp_check_edges_dev(sqmp_ctx *ctx, const double *d_q_start, const double *d_q_end, int64_t E, int num_points,
                         unsigned flags, uint32_t *d_mask, int32_t *d_first_invalid, void *stream);
/* Multi-GPU (states sharded over the ranks of one node, masks gathered on every rank) with the gather
 * fused into the validity kernel: every rank's kernel writes its mask words straight into the gather buffer of every
 * rank through peer-mapped device pointers (CUDA IPC + NVLink stores) — no collective launch in the step.  `mask[p]` /
 * `flag[p]` are rank p's gather buffer and flag array (uint32[n_peers], zero-initialised) as mapped into this process;
 * this rank's words go to mask[p][word_offset ...].  The kernel only stores; sqmp_gather_signal_dev, enqueued AFTER it
 * in stream order (same stream, or another stream behind an event), publishes `step` (use 1, 2, 3, ...) in flag[p][rank]
 * of every rank — the completed validity grid orders its stores before that.  sqmp_gather_wait_dev enqueues a one-warp
 * kernel that returns once this rank's flag array shows `step` for all n_peers ranks (status 0) or ~2 s have passed (1):
 * work enqueued behind it sees the complete gathered mask of the step.
 * REUSE RULE: the protocol has no consumer acknowledgement.  A peer may store the words of a later step into a gather
 * buffer as soon as it has been told to; a buffer may therefore be handed out again (for step i + NBUF) only after EVERY
 * rank's consumer of step i has finished reading it — the caller provides that hand-shake (a host barrier, an
 * all-reduce, or a second flag array the consumers publish and the producers wait on).  bench.py rotates four buffers
 * and never reads them inside its timed loop. */
typedef struct sqmp_peer_gather {
  int32_t n_peers, rank;
  uint32_t *mask[SQMP_MAX_PEERS];
  uint32_t *flag[SQMP_MAX_PEERS];
  int64_t word_offset;
} sqmp_peer_gather;
int sqmp_check_states_gather_dev(sqmp_ctx *ctx, const double *d_q, int64_t S, unsigned flags, const sqmp_peer_gather *g,
                                 void *stream);
int sqmp_gather_signal_dev(sqmp_ctx *ctx, const sqmp_peer_gather *g, uint32_t step, void *stream);
int sqmp_gather_wait_dev(sqmp_ctx *ctx, const uint32_t *d_flags, int n_peers, uint32_t step, int32_t *d_status, void *stream);
int sqmp_sample_configs_dev(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                            const double hi[SQMP_DOF], int max_tries, double *d_q_out, int32_t *d_tries_out, void *stream);
int sqmp_sample_informed_dev(sqmp_ctx *ctx, uint64_t seed, int64_t first, int64_t n, const double lo[SQMP_DOF],
                             const double hi[SQMP_DOF], int max_tries, const sqmp_informed *ellipse, double *d_q_out,
                             int32_t *d_tries_out, void *stream);
int sqmp_fk_dev(sqmp_ctx *ctx, const double *d_q, int64_t S, double *d_ee_pose, void *stream);
int sqmp_nn_nearest_dev(sqmp_ctx *ctx, int tree, const double *d_queries, int64_t Q, int32_t *d_id, void *stream);
/* d_overflow: one int32, set to 1 if some query had more hits than the internal per-query buffer. */
int sqmp_nn_radius_dev(sqmp_ctx *ctx, int tree, const double *d_queries, const int32_t *d_exclude_id, int64_t Q,
                       double r, int k, int32_t *d_low_ids, int32_t *d_high_ids, int32_t *d_count,
                       int32_t *d_overflow, void *stream);

/* Introspection for the bench: number of kernels this context has launched so far. */
int64_t sqmp_kernel_launches(const sqmp_ctx *ctx);
/* Tuning knobs (implementation detail, stable names): "fk_precision" 32|64, "nn_variant" 0=auto | 1 query-per-thread fp64 | 2 warp-per-query |
 * 3 fp32 pre-filter | 4 tensor-core pre-filter (auto for >= 2048 queries while the tree fits fp16; results identical in all) ... */
int sqmp_set_option(sqmp_ctx *ctx, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* SQMP_B200_H */
