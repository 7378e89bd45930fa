/*
 * sqmp_oracle.h — CPU restatement of the state-validity / near-neighbour path of
 * tpatten/squirrel_motion_planner.   TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libsqmp_b200.so and the package
 * squirrel_motion_planner_b200) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no test, golden vector or fixture that pins
 * any result on this path, and its arithmetic lives in orocos_kdl / FCL / octomap
 * (un-vendored, versions unpinned: birrt_star_algorithm/package.xml:19-22), none of
 * which exist in the build container.  KDL's frame algebra is restated from its
 * published algorithm (frames.cpp: Rotation::Rot2, Rotation::GetQuaternion,
 * Joint::pose, Segment::pose); FCL's mesh-vs-octree test is *replaced* by
 * sphere-vs-voxel-AABB on the authored sphere model.
 *
 * File tags: CC = birrt_star_algorithm/include/birrt_star_algorithm/collision_checker.hpp
 *            BS = birrt_star_algorithm/src/birrt_star.cpp
 *            DH = planning_heuristics/src/distance_heuristics.cpp
 *            DS = planner_data_structures/include/planner_data_structures/data_structs.h
 *            KM = kuka_motion_control/src/kdl_kuka_model.cpp
 */
#ifndef SQMP_ORACLE_H
#define SQMP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_DOF 8
#define ORC_JOINT_NONE 0
#define ORC_JOINT_ROT 1
#define ORC_JOINT_TRANS 2
#define ORC_FLAG_SELF 1u
#define ORC_FLAG_MAP 2u

typedef struct orc_model orc_model;

/* Frames are 12 doubles: R row-major (9) then p (3). */
orc_model *orc_model_create(int n_links, const int *parent, const int *joint_type, const int *qidx,
                            const double *joint_origin /*L*3*/, const double *joint_axis /*L*3*/,
                            const double *to_parent /*L*12*/, const double *frame_to_tip /*L*12*/,
                            const int *shaped, const unsigned char *pair_enabled /*L*L*/,
                            int n_spheres, const int *sph_link, const double *sph_center /*n*3, link frame*/,
                            const double *sph_radius);
void orc_model_destroy(orc_model *m);

/* setOcTree (CC:76-88) on a dense copy: bit (k*gy+j)*gx+i of `bits`; voxel (i,j,k) = origin + [i,i+1]*res.
 * The caller has already folded the z=-0.02 octree offset (CC:87) into origin.  Copies. */
void orc_set_grid(orc_model *m, const uint32_t *bits, int gx, int gy, int gz, const double origin[3], double res);
/* setDisabledLinkMapCollisions (CC:90-102): bit i set = link i skipped in the map check. */
void orc_set_disabled_links(orc_model *m, uint64_t mask);

/* updateTransforms (CC:519-539): world frame of every link (root z=+0.02, CC:201). */
void orc_update_transforms(const orc_model *m, const double q[ORC_DOF], double *frames /*L*12*/);

/* isInCollision (CC:104-121) with the reference's early exits.  Returns 1 = in collision. */
int orc_is_in_collision(const orc_model *m, const double q[ORC_DOF], int check_self, int check_map);

/* Same verdicts without early exit, plus the signed minimum clearances (metres):
 *   map : min over spheres of  min(dist(centre, occupied voxel AABB), r+cap) - r
 *   self: min over enabled link pairs / sphere pairs of |ca-cb| - (ra+rb)
 * A check that is switched off reports +INFINITY.  Returns 1 = in collision. */
int orc_state_clearance(const orc_model *m, const double q[ORC_DOF], unsigned flags, double cap,
                        double *clr_map, double *clr_self);

/* Batched: mask bit s (word s/32, bit s%32) = 1 ⇒ state s in collision.  clr_* may be NULL
 * (then the early-exit path runs).  nthreads<=1 ⇒ serial like the reference. */
void orc_check_states(const orc_model *m, const double *q /*S*8*/, int64_t S, unsigned flags, double cap,
                      uint32_t *mask_out, double *clr_map, double *clr_self, int nthreads);

/* getCollisions (CC:123-133, 594-631): per-link map hits and colliding link pairs. */
void orc_get_collisions(const orc_model *m, const double q[ORC_DOF], unsigned char *link_map_hit /*L*/,
                        unsigned char *pair_hit /*L*L*/);

/* interpolateConfigurations (BS:4443-4490): out[(k*8)+j] = start[j] + k*((end[j]-start[j])/num_points). */
void orc_interpolate(const double *start, const double *end, int num_points, double *out);

/* isEdgeValid (BS:6864-6895).  Returns 1 = valid; *last_valid_idx as the reference sets it. */
int orc_is_edge_valid(const orc_model *m, const double *traj /*n*8*/, int n_states, int check_self,
                      int check_map, int *last_valid_idx);

/* connectNodesInterpolation + isEdgeValid over E edges given their end points (BS:4380-4440, 6864-6895):
 * mask bit e*(num_points+1)+k; first_invalid[e] = index of first colliding state or -1. */
void orc_check_edges(const orc_model *m, const double *q_start, const double *q_end, int64_t E, int num_points,
                     unsigned flags, uint32_t *mask_out, int32_t *first_invalid, int nthreads);

/* compute_FK (KM:278-305): KDL chain solver from the root to link `tip` — x y z qx qy qz qw. */
void orc_fk(const orc_model *m, int tip, const double *q /*S*8*/, int64_t S, double *out /*S*7*/);

/* sampleJointConfig_JntArray (BS:3832-3878) over getRandomConf (CL:1120-1188): every joint uniform in [lo, hi],
 * redrawn until 0.0 <= ee_z of compute_FK to link `tip`.  The reference's time-seeded Mersenne twister is replaced by
 * a counter-based stream so a batch is reproducible: try t of sample i uses outputs 8 * ((first + i) * max_tries + t)
 * + j of the splitmix64 stream started at `seed` (k-th output = mix(seed + (k + 1) * 0x9E3779B97F4A7C15)), and
 * q_j = lo_j + (hi_j - lo_j) * (double(out >> 11) * 2^-53) with no fused multiply-add.  tries_out (may be NULL): tries
 * used, or -max_tries when every try was below the platform (the last try is returned). */
void orc_sample_configs(const orc_model *m, int tip, uint64_t seed, int64_t first, int64_t n, const double *lo,
                        const double *hi, int max_tries, double *q_out /*n*8*/, int32_t *tries_out);
/* layout of sqmp_informed (include/sqmp_b200.h) */
typedef struct orc_informed {
  int32_t n_rev, n_pri;
  int32_t rev_idx[ORC_DOF], pri_idx[ORC_DOF];
  double c_rev[ORC_DOF * ORC_DOF], c_pri[ORC_DOF * ORC_DOF];
  double center_rev[ORC_DOF], center_pri[ORC_DOF];
  double l_rev[2], l_pri[2];
  double env_x[2], env_y[2];
} orc_informed;
void orc_sample_informed(const orc_model *m, int tip, uint64_t seed, int64_t first, int64_t n, const double *lo,
                         const double *hi, int max_tries, const orc_informed *el, double *q_out, int32_t *tries_out);

/* euclidean_joint_space_distance (DH:128-156). */
double orc_joint_space_distance(const double *a, const double *b, int n);

/* find_nearest_neighbour_interpolation (BS:4076-4133), serial semantics. nodes: N*8 row-major. */
int32_t orc_nearest(const double *nodes, int64_t N, const double *query);
void orc_nearest_batch(const double *nodes, int64_t N, const double *queries, int64_t Q, int32_t *ids,
                       int nthreads);

/* find_near_vertices_interpolation (BS:4272-4324): ids with d < r and id != exclude, ordered by
 * (cost asc, id asc) — the reference's std::sort is unstable on ties; id makes it deterministic.
 * Writes at most cap ids; returns the full count. */
int64_t orc_near_vertices(const double *nodes, const double *cost, int64_t N, const double *query,
                          int32_t exclude_id, double r, int32_t *ids_out, int64_t cap);
/* Fixed-size consumer view: low = first k (BS:4618-4623), high = last k walked from the end
 * (BS:5075-5089), -1 padded. */
void orc_near_topk_batch(const double *nodes, const double *cost, int64_t N, const double *queries,
                         const int32_t *exclude_id, int64_t Q, double r, int k, int32_t *low_ids,
                         int32_t *high_ids, int32_t *count, int nthreads);

/* compute_edge_cost_interpolation (BS:4162-4242): out = total, revolute, prismatic. */
void orc_edge_cost(const double *traj, int n_states, const int *is_revolute /*8*/, const double *weights /*8*/,
                   double out[3]);

/* stepTowardsRandSample (BS:5712-5868): x_rand updated in place; returns 1 = sample reached. */
int orc_step_towards(const double *nn, double *x_rand, const int *is_revolute, double extend_step_factor);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
