/*
 * sqmp_oracle.c — CPU restatement (plain C, fp64, no FMA contraction: build with -ffp-contract=off).
 * TEST INFRASTRUCTURE ONLY — see sqmp_oracle.h for who may load it and for the "parity unpinned" note.
 * Every function cites the reference lines it follows (tags in sqmp_oracle.h).
 */
#include "sqmp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_LINKS 64

struct orc_model {
  int L;
  int parent[MAX_LINKS], joint_type[MAX_LINKS], qidx[MAX_LINKS], shaped[MAX_LINKS];
  double joint_origin[MAX_LINKS][3], joint_axis[MAX_LINKS][3];
  double to_parent[MAX_LINKS][12], frame_to_tip[MAX_LINKS][12];
  unsigned char pair_enabled[MAX_LINKS][MAX_LINKS];
  int n_spheres;
  int *sph_link;        /* sorted by link in creation order is not required */
  double *sph_center;   /* n*3 */
  double *sph_radius;
  int sph_first[MAX_LINKS], sph_count[MAX_LINKS]; /* spheres regrouped by link */
  int *sph_order;
  /* grid */
  uint32_t *bits;
  int gx, gy, gz;
  double origin[3], res;
  uint64_t disabled_links;
};

/* ---------------------------------------------------------------- KDL frame algebra (frames.inl) */
static void frame_identity(double *F) {
  memset(F, 0, 12 * sizeof(double));
  F[0] = F[4] = F[8] = 1.0;
}
/* KDL::operator*(Frame,Frame) = Frame(l.M*r.M, l.M*r.p + l.p) */
static void frame_mul(const double *A, const double *B, double *C) {
  double T[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
      T[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    T[9 + i] = (A[i * 3 + 0] * B[9] + A[i * 3 + 1] * B[10] + A[i * 3 + 2] * B[11]) + A[9 + i];
  memcpy(C, T, sizeof(T));
}
static void frame_apply(const double *F, const double *v, double *out) {
  for (int i = 0; i < 3; ++i) out[i] = (F[i * 3 + 0] * v[0] + F[i * 3 + 1] * v[1] + F[i * 3 + 2] * v[2]) + F[9 + i];
}
/* KDL::Rotation::Rot2(axis, angle) — Rodrigues, axis already normalised */
static void rot2(const double *v, double angle, double *R) {
  double ct = cos(angle), st = sin(angle), vt = 1.0 - ct;
  double m_vt_0 = vt * v[0], m_vt_1 = vt * v[1], m_vt_2 = vt * v[2];
  double m_st_0 = v[0] * st, m_st_1 = v[1] * st, m_st_2 = v[2] * st;
  double m_vt_0_1 = m_vt_0 * v[1], m_vt_0_2 = m_vt_0 * v[2], m_vt_1_2 = m_vt_1 * v[2];
  R[0] = ct + m_vt_0 * v[0]; R[1] = -m_st_2 + m_vt_0_1; R[2] = m_st_1 + m_vt_0_2;
  R[3] = m_st_2 + m_vt_0_1;  R[4] = ct + m_vt_1 * v[1]; R[5] = -m_st_0 + m_vt_1_2;
  R[6] = -m_st_1 + m_vt_0_2; R[7] = m_st_0 + m_vt_1_2;  R[8] = ct + m_vt_2 * v[2];
}
/* KDL::Joint::pose(q): RotAxis -> Frame(Rot2(axis,q), origin); TransAxis -> Frame(origin + q*axis) */
static void joint_pose(int type, const double *origin, const double *axis, double q, double *F) {
  if (type == ORC_JOINT_ROT) {
    rot2(axis, q, F);
    F[9] = origin[0]; F[10] = origin[1]; F[11] = origin[2];
  } else {
    frame_identity(F);
    F[9] = origin[0] + q * axis[0]; F[10] = origin[1] + q * axis[1]; F[11] = origin[2] + q * axis[2];
  }
}
/* KDL::Rotation::GetQuaternion (frames.cpp) */
static void get_quaternion(const double *R, double *x, double *y, double *z, double *w) {
  double trace = R[0] + R[4] + R[8];
  double epsilon = 1E-12;
  if (trace > epsilon) {
    double s = 0.5 / sqrt(trace + 1.0);
    *w = 0.25 / s;
    *x = (R[7] - R[5]) * s;
    *y = (R[2] - R[6]) * s;
    *z = (R[3] - R[1]) * s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double s = 2.0 * sqrt(1.0 + R[0] - R[4] - R[8]);
    *w = (R[7] - R[5]) / s;
    *x = 0.25 * s;
    *y = (R[1] + R[3]) / s;
    *z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    double s = 2.0 * sqrt(1.0 + R[4] - R[0] - R[8]);
    *w = (R[2] - R[6]) / s;
    *x = (R[1] + R[3]) / s;
    *y = 0.25 * s;
    *z = (R[5] + R[7]) / s;
  } else {
    double s = 2.0 * sqrt(1.0 + R[8] - R[0] - R[4]);
    *w = (R[3] - R[1]) / s;
    *x = (R[2] + R[6]) / s;
    *y = (R[5] + R[7]) / s;
    *z = 0.25 * s;
  }
}

/* ---------------------------------------------------------------- model */
orc_model *orc_model_create(int n_links, const int *parent, const int *joint_type, const int *qidx,
                            const double *joint_origin, const double *joint_axis, const double *to_parent,
                            const double *frame_to_tip, const int *shaped, const unsigned char *pair_enabled,
                            int n_spheres, const int *sph_link, const double *sph_center,
                            const double *sph_radius) {
  if (n_links <= 0 || n_links > MAX_LINKS) return NULL;
  orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
  m->L = n_links;
  for (int i = 0; i < n_links; ++i) {
    m->parent[i] = parent[i]; m->joint_type[i] = joint_type[i]; m->qidx[i] = qidx[i]; m->shaped[i] = shaped[i];
    memcpy(m->joint_origin[i], joint_origin + 3 * i, 3 * sizeof(double));
    memcpy(m->joint_axis[i], joint_axis + 3 * i, 3 * sizeof(double));
    memcpy(m->to_parent[i], to_parent + 12 * i, 12 * sizeof(double));
    memcpy(m->frame_to_tip[i], frame_to_tip + 12 * i, 12 * sizeof(double));
    for (int j = 0; j < n_links; ++j) m->pair_enabled[i][j] = pair_enabled[i * n_links + j];
  }
  m->n_spheres = n_spheres;
  m->sph_link = (int *)malloc(sizeof(int) * (n_spheres + 1));
  m->sph_center = (double *)malloc(sizeof(double) * 3 * (n_spheres + 1));
  m->sph_radius = (double *)malloc(sizeof(double) * (n_spheres + 1));
  m->sph_order = (int *)malloc(sizeof(int) * (n_spheres + 1));
  memcpy(m->sph_link, sph_link, sizeof(int) * n_spheres);
  memcpy(m->sph_center, sph_center, sizeof(double) * 3 * n_spheres);
  memcpy(m->sph_radius, sph_radius, sizeof(double) * n_spheres);
  int pos = 0;
  for (int l = 0; l < n_links; ++l) {
    m->sph_first[l] = pos;
    for (int s = 0; s < n_spheres; ++s)
      if (sph_link[s] == l) m->sph_order[pos++] = s;
    m->sph_count[l] = pos - m->sph_first[l];
  }
  return m;
}

void orc_model_destroy(orc_model *m) {
  if (!m) return;
  free(m->sph_link); free(m->sph_center); free(m->sph_radius); free(m->sph_order); free(m->bits);
  free(m);
}

void orc_set_grid(orc_model *m, const uint32_t *bits, int gx, int gy, int gz, const double origin[3], double res) {
  if (!bits) return; /* CC:78-79: NULL octree ignored */
  free(m->bits);
  size_t nbits = (size_t)gx * gy * gz, nwords = (nbits + 31) / 32;
  m->bits = (uint32_t *)malloc(nwords * sizeof(uint32_t));
  memcpy(m->bits, bits, nwords * sizeof(uint32_t));
  m->gx = gx; m->gy = gy; m->gz = gz; m->res = res;
  memcpy(m->origin, origin, 3 * sizeof(double));
}

void orc_set_disabled_links(orc_model *m, uint64_t mask) { m->disabled_links = mask; }

/* CC:519-539.  links[0].transform = Identity with p.z = 0.02 (CC:199-201). */
void orc_update_transforms(const orc_model *m, const double q[ORC_DOF], double *frames) {
  frame_identity(frames);
  frames[11] = 0.02;
  for (int i = 1; i < m->L; ++i) {
    const double *P = frames + 12 * m->parent[i];
    if (m->joint_type[i] != ORC_JOINT_NONE) {
      double J[12];
      joint_pose(m->joint_type[i], m->joint_origin[i], m->joint_axis[i], q[m->qidx[i]], J); /* CC:528 */
      frame_mul(P, J, frames + 12 * i);
    } else {
      frame_mul(P, m->to_parent[i], frames + 12 * i); /* CC:530 */
    }
  }
}

/* ---------------------------------------------------------------- sphere vs occupied voxel boxes */
static inline int grid_bit(const orc_model *m, int i, int j, int k) {
  size_t lin = ((size_t)k * m->gy + j) * m->gx + i;
  return (m->bits[lin >> 5] >> (lin & 31)) & 1u;
}
static inline double axis_dist(double c, double lo, double hi) {
  double d = 0.0;
  if (c < lo) d = lo - c;
  else if (c > hi) d = c - hi;
  return d;
}
static inline void axis_range(double c, double R, double o, double res, int n, int *a, int *b) {
  double fa = floor((c - R - o) / res) - 1.0, fb = floor((c + R - o) / res) + 1.0;
  if (fa < 0.0) fa = 0.0;
  if (fb > (double)(n - 1)) fb = (double)(n - 1);
  if (fa > (double)n) fa = (double)n;
  if (fb < -1.0) fb = -1.0;
  *a = (int)fa; *b = (int)fb;
}
/* min over occupied voxels (within reach R of c) of dist(c, voxel AABB); returns R if none closer. */
static double sphere_grid_distance(const orc_model *m, const double *c, double R) {
  if (!m->bits) return R;
  int i0, i1, j0, j1, k0, k1;
  axis_range(c[0], R, m->origin[0], m->res, m->gx, &i0, &i1);
  axis_range(c[1], R, m->origin[1], m->res, m->gy, &j0, &j1);
  axis_range(c[2], R, m->origin[2], m->res, m->gz, &k0, &k1);
  double best = R;
  for (int k = k0; k <= k1; ++k) {
    double dz = axis_dist(c[2], m->origin[2] + k * m->res, m->origin[2] + (k + 1) * m->res);
    if (dz > best) continue;
    for (int j = j0; j <= j1; ++j) {
      double dy = axis_dist(c[1], m->origin[1] + j * m->res, m->origin[1] + (j + 1) * m->res);
      if (dy > best) continue;
      for (int i = i0; i <= i1; ++i) {
        if (!grid_bit(m, i, j, k)) continue;
        double dx = axis_dist(c[0], m->origin[0] + i * m->res, m->origin[0] + (i + 1) * m->res);
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        if (d < best) best = d;
      }
    }
  }
  return best;
}
/* verdict only, first hit exits: collide ⇔ dist(c, some occupied voxel AABB) - r <= 0 */
static int sphere_hits_grid(const orc_model *m, const double *c, double r) {
  if (!m->bits) return 0; /* CC:556-557: no octomap object yet */
  int i0, i1, j0, j1, k0, k1;
  axis_range(c[0], r, m->origin[0], m->res, m->gx, &i0, &i1);
  axis_range(c[1], r, m->origin[1], m->res, m->gy, &j0, &j1);
  axis_range(c[2], r, m->origin[2], m->res, m->gz, &k0, &k1);
  for (int k = k0; k <= k1; ++k) {
    double dz = axis_dist(c[2], m->origin[2] + k * m->res, m->origin[2] + (k + 1) * m->res);
    if (dz > r) continue;
    for (int j = j0; j <= j1; ++j) {
      double dy = axis_dist(c[1], m->origin[1] + j * m->res, m->origin[1] + (j + 1) * m->res);
      if (dy > r) continue;
      size_t row = ((size_t)k * m->gy + j) * m->gx;
      for (int i = i0; i <= i1; ++i) {
        size_t lin = row + i;
        uint32_t w = m->bits[lin >> 5] >> (lin & 31);
        if (w == 0) { i += 31 - (int)(lin & 31); continue; } /* rest of this word is empty */
        if (!(w & 1u)) continue;
        double dx = axis_dist(c[0], m->origin[0] + i * m->res, m->origin[0] + (i + 1) * m->res);
        if (sqrt(dx * dx + dy * dy + dz * dz) - r <= 0.0) return 1;
      }
    }
  }
  return 0;
}

/* checkMapCollision (CC:554-592): every shaped link not disabled, first hit wins. */
static int check_map(const orc_model *m, const double *frames) {
  for (int l = 0; l < m->L; ++l) {
    if (!m->shaped[l] || ((m->disabled_links >> l) & 1ull)) continue;
    for (int t = 0; t < m->sph_count[l]; ++t) {
      int s = m->sph_order[m->sph_first[l] + t];
      double c[3];
      frame_apply(frames + 12 * l, m->sph_center + 3 * s, c);
      if (sphere_hits_grid(m, c, m->sph_radius[s])) return 1;
    }
  }
  return 0;
}
static double pair_clearance(const orc_model *m, const double *frames, int a, int b) {
  double best = INFINITY;
  for (int t = 0; t < m->sph_count[a]; ++t) {
    int s = m->sph_order[m->sph_first[a] + t];
    double ca[3];
    frame_apply(frames + 12 * a, m->sph_center + 3 * s, ca);
    for (int u = 0; u < m->sph_count[b]; ++u) {
      int v = m->sph_order[m->sph_first[b] + u];
      double cb[3];
      frame_apply(frames + 12 * b, m->sph_center + 3 * v, cb);
      double dx = ca[0] - cb[0], dy = ca[1] - cb[1], dz = ca[2] - cb[2];
      double d = sqrt(dx * dx + dy * dy + dz * dz) - (m->sph_radius[s] + m->sph_radius[v]);
      if (d < best) best = d;
    }
  }
  return best;
}
/* checkSelfCollision (CC:541-552) over selfCollisionPairs (CC:387-390: i<j, both shaped, enabled). */
static int check_self(const orc_model *m, const double *frames) {
  for (int a = 0; a < m->L; ++a)
    for (int b = a + 1; b < m->L; ++b)
      if (m->pair_enabled[a][b] && m->shaped[a] && m->shaped[b])
        if (pair_clearance(m, frames, a, b) <= 0.0) return 1;
  return 0;
}

/* Fail-closed policy of the new boundary (no counterpart in the reference, which would feed NaN frames to FCL): a
 * joint value that is not finite, or absurd (|q_j| >= 1e7 m / rad), makes the state "in collision" — the spirit of
 * CC:106-107, where a checker that cannot answer says "collision". */
static int state_is_sane(const double q[ORC_DOF]) {
  for (int j = 0; j < ORC_DOF; ++j)
    if (!(fabs(q[j]) < 1.0e7)) return 0;
  return 1;
}

/* isInCollision (CC:104-121).  "not initialised ⇒ true" (CC:106-107) is the NULL model here. */
int orc_is_in_collision(const orc_model *m, const double q[ORC_DOF], int self, int map) {
  if (!m) return 1;
  if (!self && !map) return 0;
  if (!state_is_sane(q)) return 1;
  double frames[MAX_LINKS * 12];
  orc_update_transforms(m, q, frames);
  if (map && check_map(m, frames)) return 1;
  if (self && check_self(m, frames)) return 1;
  return 0;
}

int orc_state_clearance(const orc_model *m, const double q[ORC_DOF], unsigned flags, double cap, double *clr_map,
                        double *clr_self) {
  double frames[MAX_LINKS * 12];
  double cm = INFINITY, cs = INFINITY;
  if ((flags & (ORC_FLAG_MAP | ORC_FLAG_SELF)) && !state_is_sane(q)) {
    if (clr_map) *clr_map = -INFINITY;
    if (clr_self) *clr_self = -INFINITY;
    return 1;
  }
  orc_update_transforms(m, q, frames);
  if ((flags & ORC_FLAG_MAP) && m->bits) {
    for (int l = 0; l < m->L; ++l) {
      if (!m->shaped[l] || ((m->disabled_links >> l) & 1ull)) continue;
      for (int t = 0; t < m->sph_count[l]; ++t) {
        int s = m->sph_order[m->sph_first[l] + t];
        double c[3], r = m->sph_radius[s];
        frame_apply(frames + 12 * l, m->sph_center + 3 * s, c);
        double d = sphere_grid_distance(m, c, r + cap) - r;
        if (d < cm) cm = d;
      }
    }
  }
  if (flags & ORC_FLAG_SELF) {
    for (int a = 0; a < m->L; ++a)
      for (int b = a + 1; b < m->L; ++b)
        if (m->pair_enabled[a][b] && m->shaped[a] && m->shaped[b]) {
          double d = pair_clearance(m, frames, a, b);
          if (d < cs) cs = d;
        }
  }
  if (clr_map) *clr_map = cm;
  if (clr_self) *clr_self = cs;
  return (cm <= 0.0) || (cs <= 0.0);
}

void orc_check_states(const orc_model *m, const double *q, int64_t S, unsigned flags, double cap, uint32_t *mask_out,
                      double *clr_map, double *clr_self, int nthreads) {
  int64_t nwords = (S + 31) / 32;
  int full = (clr_map != NULL) || (clr_self != NULL);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (int64_t w = 0; w < nwords; ++w) {
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
      int64_t s = w * 32 + b;
      if (s >= S) break;
      int hit;
      if (full) {
        double cm, cs;
        hit = orc_state_clearance(m, q + s * ORC_DOF, flags, cap, &cm, &cs);
        if (clr_map) clr_map[s] = cm;
        if (clr_self) clr_self[s] = cs;
      } else {
        hit = orc_is_in_collision(m, q + s * ORC_DOF, (flags & ORC_FLAG_SELF) != 0, (flags & ORC_FLAG_MAP) != 0);
      }
      bits |= (uint32_t)hit << b;
    }
    mask_out[w] = bits;
  }
}

/* getCollisions (CC:123-133): getSelfCollisions (CC:594-608) + getMapCollisions (CC:610-631); note that
 * getMapCollisions ignores the per-request disable flags (it tests every shaped link). */
void orc_get_collisions(const orc_model *m, const double q[ORC_DOF], unsigned char *link_map_hit,
                        unsigned char *pair_hit) {
  double frames[MAX_LINKS * 12];
  orc_update_transforms(m, q, frames);
  memset(link_map_hit, 0, (size_t)m->L);
  memset(pair_hit, 0, (size_t)m->L * m->L);
  for (int a = 0; a < m->L; ++a)
    for (int b = a + 1; b < m->L; ++b)
      if (m->pair_enabled[a][b] && m->shaped[a] && m->shaped[b] && pair_clearance(m, frames, a, b) <= 0.0)
        pair_hit[a * m->L + b] = 1;
  if (!m->bits) return;
  for (int l = 0; l < m->L; ++l) {
    if (!m->shaped[l]) continue;
    for (int t = 0; t < m->sph_count[l]; ++t) {
      int s = m->sph_order[m->sph_first[l] + t];
      double c[3];
      frame_apply(frames + 12 * l, m->sph_center + 3 * s, c);
      if (sphere_hits_grid(m, c, m->sph_radius[s])) { link_map_hit[l] = 1; break; }
    }
  }
}

/* ---------------------------------------------------------------- edges */
/* BS:4455-4463, 4473-4480 */
void orc_interpolate(const double *start, const double *end, int num_points, double *out) {
  double step[ORC_DOF];
  for (int j = 0; j < ORC_DOF; ++j) {
    double c_space_dist = end[j] - start[j];
    step[j] = c_space_dist / (double)num_points;
  }
  for (int inc = 0; inc <= num_points; ++inc)
    for (int j = 0; j < ORC_DOF; ++j) out[inc * ORC_DOF + j] = start[j] + inc * step[j];
}

/* BS:6878-6895 (the int& overload; the bool-only overload BS:6864-6876 is the same walk) */
int orc_is_edge_valid(const orc_model *m, const double *traj, int n_states, int self, int map, int *last_valid_idx) {
  int last = 0;
  for (int edp = 0; edp < n_states; ++edp) {
    if (orc_is_in_collision(m, traj + edp * ORC_DOF, self, map)) {
      if (last_valid_idx) *last_valid_idx = last;
      return 0;
    }
    last = edp;
  }
  if (last_valid_idx) *last_valid_idx = last;
  return 1;
}

void orc_check_edges(const orc_model *m, const double *q_start, const double *q_end, int64_t E, int num_points,
                     unsigned flags, uint32_t *mask_out, int32_t *first_invalid, int nthreads) {
  int P1 = num_points + 1;
  int64_t nbits = E * P1, nwords = (nbits + 31) / 32;
  memset(mask_out, 0, (size_t)nwords * sizeof(uint32_t));
  unsigned char *hits = (unsigned char *)malloc((size_t)nbits);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (int64_t e = 0; e < E; ++e) {
    double traj[ORC_DOF * 256];
    orc_interpolate(q_start + e * ORC_DOF, q_end + e * ORC_DOF, num_points, traj);
    int fi = -1;
    for (int k = 0; k < P1; ++k) {
      int h = orc_is_in_collision(m, traj + k * ORC_DOF, (flags & ORC_FLAG_SELF) != 0, (flags & ORC_FLAG_MAP) != 0);
      hits[e * P1 + k] = (unsigned char)h;
      if (h && fi < 0) fi = k;
    }
    if (first_invalid) first_invalid[e] = fi;
  }
  for (int64_t b = 0; b < nbits; ++b)
    if (hits[b]) mask_out[b >> 5] |= 1u << (b & 31);
  free(hits);
}

/* ---------------------------------------------------------------- compute_FK (KM:278-305) */
void orc_fk(const orc_model *m, int tip, const double *q, int64_t S, double *out) {
  int path[MAX_LINKS], n = 0;
  for (int l = tip; l > 0; l = m->parent[l]) path[n++] = l;
  for (int64_t s = 0; s < S; ++s) {
    const double *qs = q + s * ORC_DOF;
    double T[12];
    frame_identity(T); /* ChainFkSolverPos_recursive::JntToCart: p_out = Identity */
    for (int t = n - 1; t >= 0; --t) {
      int l = path[t];
      if (m->joint_type[l] != ORC_JOINT_NONE) {
        /* Segment::pose(q) = joint.pose(q) * f_tip,  f_tip = joint.pose(0).Inverse() * F */
        double J[12], J0[12], Jinv[12], ftip[12], seg[12];
        joint_pose(m->joint_type[l], m->joint_origin[l], m->joint_axis[l], qs[m->qidx[l]], J);
        joint_pose(m->joint_type[l], m->joint_origin[l], m->joint_axis[l], 0.0, J0);
        /* Frame::Inverse(): (M^T, -M^T p) */
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) Jinv[i * 3 + j] = J0[j * 3 + i];
        for (int i = 0; i < 3; ++i)
          Jinv[9 + i] = -(Jinv[i * 3 + 0] * J0[9] + Jinv[i * 3 + 1] * J0[10] + Jinv[i * 3 + 2] * J0[11]);
        frame_mul(Jinv, m->frame_to_tip[l], ftip);
        frame_mul(J, ftip, seg);
        frame_mul(T, seg, T);
      } else {
        frame_mul(T, m->frame_to_tip[l], T);
      }
    }
    double *o = out + s * 7;
    o[0] = T[9]; o[1] = T[10]; o[2] = T[11];
    get_quaternion(T, &o[3], &o[4], &o[5], &o[6]);
  }
}

/* ---------------------------------------------------------------- sampling (BS:3832-3878, CL:1120-1188) */
static uint64_t splitmix_at(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void orc_sample_configs(const orc_model *m, int tip, uint64_t seed, int64_t first, int64_t n, const double *lo,
                        const double *hi, int max_tries, double *q_out, int32_t *tries_out) {
  for (int64_t i = 0; i < n; ++i) {
    double *q = q_out + i * ORC_DOF, ee[7];
    int used = -max_tries;
    for (int t = 0; t < max_tries; ++t) {   /* while(true) of CL:1130, bounded */
      const uint64_t k0 = ((uint64_t)(first + i) * (uint64_t)max_tries + (uint64_t)t) * ORC_DOF;
      for (int j = 0; j < ORC_DOF; ++j) {   /* uniformReal(q_min, q_max) / (env_size[0], env_size[1]), CL:1141-1166 */
        const double u = (double)(splitmix_at(seed, k0 + (uint64_t)j) >> 11) * (1.0 / 9007199254740992.0);
        const double span = hi[j] - lo[j];
        const double step = span * u;
        q[j] = lo[j] + step;
      }
      orc_fk(m, tip, q, 1, ee);
      if (0.0 <= ee[2]) {   /* CL:1182-1184, BS:3870-3871 */
        used = t + 1;
        break;
      }
    }
    if (tries_out) tries_out[i] = used;
  }
}

/* sampleJointConfigfromEllipse_JntArray (BS:3607-3829) over the counter stream of orc_sample_configs, nine outputs per
 * try (eight joint draws for getRandomConf, BS:3633, + the point scale, BS:3690).  The ellipse (centres, rotations C,
 * diagonal L) is prepared by the caller like jointConfigEllipseInitialization (BS:3472-3604) / BS:3709-3752 do. */
void orc_sample_informed(const orc_model *m, int tip, uint64_t seed, int64_t first, int64_t n, const double *lo,
                         const double *hi, int max_tries, const orc_informed *el, double *q_out, int32_t *tries_out) {
  const int unconfined = el->env_x[0] == 0.0 && el->env_x[1] == 0.0 && el->env_y[0] == 0.0 && el->env_y[1] == 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double *rand_conf_array = q_out + i * ORC_DOF, ee_pose[7];
    int used = -max_tries;
    for (int t = 0; t < max_tries; ++t) {   /* while (collision_free == false || above_platform == false || ...), bounded */
      const uint64_t k0 = ((uint64_t)(first + i) * (uint64_t)max_tries + (uint64_t)t) * 9ull;
      double ball_conf_array[ORC_DOF];
      for (int j = 0; j < ORC_DOF; ++j) {   /* getRandomConf (BS:3633) */
        const double u = (double)(splitmix_at(seed, k0 + (uint64_t)j) >> 11) * (1.0 / 9007199254740992.0);
        const double span = hi[j] - lo[j];
        const double step = span * u;
        ball_conf_array[j] = lo[j] + step;
      }
      const double point_scale = (double)(splitmix_at(seed, k0 + 8ull) >> 11) * (1.0 / 9007199254740992.0);   /* BS:3690 */
      int joint_index = 0;
      for (int cls = 0; cls < 2; ++cls) {   /* BS:3773-3785: prismatic joints first, then revolute ones */
        const int nc = cls == 0 ? el->n_pri : el->n_rev;
        const int32_t *idx = cls == 0 ? el->pri_idx : el->rev_idx;
        const double *C = cls == 0 ? el->c_pri : el->c_rev;
        const double *center = cls == 0 ? el->center_pri : el->center_rev;
        const double *L = cls == 0 ? el->l_pri : el->l_rev;
        double sum_squares = 0.0, y[ORC_DOF];
        for (int a = 0; a < nc; ++a) {   /* BS:3676-3683 */
          const double sq = ball_conf_array[idx[a]] * ball_conf_array[idx[a]];
          sum_squares += sq;
        }
        const double conf_norm = sqrt(sum_squares);
        for (int a = 0; a < nc; ++a) {   /* BS:3697-3704, then the diagonal L of BS:3709-3752 */
          const double unit = ball_conf_array[idx[a]] / conf_norm;
          const double ball = point_scale * unit;
          y[a] = (a == 0 ? L[0] : L[1]) * ball;
        }
        for (int r = 0; r < nc; ++r) {   /* BS:3759-3763: C * L * ball + centre */
          double acc = 0.0;
          for (int a = 0; a < nc; ++a) {
            const double prod = C[r * nc + a] * y[a];
            acc += prod;
          }
          rand_conf_array[joint_index++] = acc + center[r];
        }
      }
      orc_fk(m, tip, rand_conf_array, 1, ee_pose);
      const int above_platform = 0.0 <= ee_pose[2];   /* BS:3792-3795 */
      const int inside_environment = unconfined || (rand_conf_array[0] < el->env_x[1] && rand_conf_array[0] > el->env_x[0] &&
                                                    rand_conf_array[1] < el->env_y[1] && rand_conf_array[1] > el->env_y[0]);   /* BS:3810-3816 */
      if (above_platform && inside_environment) {
        used = t + 1;
        break;
      }
    }
    if (tries_out) tries_out[i] = used;
  }
}

/* ---------------------------------------------------------------- distances and neighbours */
/* DH:128-156 */
double orc_joint_space_distance(const double *a, const double *b, int n) {
  double sum_error_sq = 0.0;
  for (int i = 0; i < n; ++i) {
    double d = b[i] - a[i];
    sum_error_sq += (d * d);
  }
  return sqrt(sum_error_sq);
}

/* BS:4076-4133, serial order: strict '<', min starts at 10000.0, id starts at 0. */
int32_t orc_nearest(const double *nodes, int64_t N, const double *query) {
  int32_t id_nearest_neighbour = 0;
  double min_node_dist = 10000.0;
  for (int64_t n = 0; n < N; ++n) {
    double curr_node_dist = orc_joint_space_distance(nodes + n * ORC_DOF, query, ORC_DOF);
    if (curr_node_dist < min_node_dist) {
      id_nearest_neighbour = (int32_t)n; /* node_id == index (DS invariant) */
      min_node_dist = curr_node_dist;
    }
  }
  return id_nearest_neighbour;
}
void orc_nearest_batch(const double *nodes, int64_t N, const double *queries, int64_t Q, int32_t *ids, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (int64_t i = 0; i < Q; ++i) ids[i] = orc_nearest(nodes, N, queries + i * ORC_DOF);
}

typedef struct { double cost; int32_t id; } near_t;
static int near_cmp(const void *a, const void *b) {
  const near_t *x = (const near_t *)a, *y = (const near_t *)b;
  if (x->cost < y->cost) return -1; /* DS:41-44 Node::operator< */
  if (x->cost > y->cost) return 1;
  return (x->id > y->id) - (x->id < y->id);
}
static near_t *near_collect(const double *nodes, const double *cost, int64_t N, const double *query, int32_t exclude_id,
                            double r, int64_t *count) {
  int64_t cap = 64, n_found = 0;
  near_t *v = (near_t *)malloc(sizeof(near_t) * cap);
  for (int64_t n = 0; n < N; ++n) {
    double joint_space_dist = orc_joint_space_distance(nodes + n * ORC_DOF, query, ORC_DOF);
    if (joint_space_dist < r && exclude_id != (int32_t)n) { /* BS:4295 */
      if (n_found == cap) { cap *= 2; v = (near_t *)realloc(v, sizeof(near_t) * cap); }
      v[n_found].cost = cost[n]; v[n_found].id = (int32_t)n; ++n_found;
    }
  }
  qsort(v, (size_t)n_found, sizeof(near_t), near_cmp); /* BS:4308 */
  *count = n_found;
  return v;
}
int64_t orc_near_vertices(const double *nodes, const double *cost, int64_t N, const double *query, int32_t exclude_id,
                          double r, int32_t *ids_out, int64_t cap) {
  int64_t n_found;
  near_t *v = near_collect(nodes, cost, N, query, exclude_id, r, &n_found);
  for (int64_t i = 0; i < n_found && i < cap; ++i) ids_out[i] = v[i].id;
  free(v);
  return n_found;
}
void orc_near_topk_batch(const double *nodes, const double *cost, int64_t N, const double *queries,
                         const int32_t *exclude_id, int64_t Q, double r, int k, int32_t *low_ids, int32_t *high_ids,
                         int32_t *count, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (int64_t i = 0; i < Q; ++i) {
    int64_t n_found;
    near_t *v = near_collect(nodes, cost, N, queries + i * ORC_DOF, exclude_id ? exclude_id[i] : -1, r, &n_found);
    for (int t = 0; t < k; ++t) {
      low_ids[i * k + t] = t < n_found ? v[t].id : -1;                   /* BS:4618-4623 */
      high_ids[i * k + t] = t < n_found ? v[n_found - 1 - t].id : -1;    /* BS:5075-5089 */
    }
    count[i] = (int32_t)n_found;
    free(v);
  }
}

/* BS:4162-4242 */
void orc_edge_cost(const double *traj, int n_states, const int *is_revolute, const double *weights, double out[3]) {
  double edge_cost_prism = 0.0, edge_cost_rev = 0.0, edge_cost_total = 0.0;
  for (int wp = 0; wp < n_states - 1; ++wp) {
    double wp_dist_total = 0.0, wp_dist_rot = 0.0, wp_dist_prism = 0.0;
    for (int j = 0; j < ORC_DOF; ++j) {
      double coordinate_dist = (traj[(wp + 1) * ORC_DOF + j] - traj[wp * ORC_DOF + j]) *
                               (traj[(wp + 1) * ORC_DOF + j] - traj[wp * ORC_DOF + j]);
      wp_dist_total += coordinate_dist * weights[j];
      if (is_revolute[j]) wp_dist_rot += coordinate_dist;
      else wp_dist_prism += coordinate_dist;
    }
    edge_cost_total += sqrt(wp_dist_total);
    edge_cost_rev += sqrt(wp_dist_rot);
    edge_cost_prism += sqrt(wp_dist_prism);
  }
  out[0] = edge_cost_total; out[1] = edge_cost_rev; out[2] = edge_cost_prism;
}

/* BS:5712-5868 */
int orc_step_towards(const double *nn, double *x_rand, const int *is_revolute, double extend_step_factor) {
  int rand_sample_reached = 1;
  double eucl_dist[ORC_DOF], extend_config[ORC_DOF];
  double sum_squares_rev_joints = 0.0, sum_squares_prism_joints = 0.0;
  for (int j = 0; j < ORC_DOF; ++j) {
    eucl_dist[j] = x_rand[j] - nn[j];
    double dist_sqrt = eucl_dist[j] * eucl_dist[j];
    if (!is_revolute[j]) sum_squares_prism_joints += dist_sqrt;
    else sum_squares_rev_joints += dist_sqrt;
    extend_config[j] = 0.0;
  }
  double vec_length_rot = sqrt(sum_squares_rev_joints), vec_length_prism = sqrt(sum_squares_prism_joints);
  int rot_joint_target_reached = vec_length_rot < 0.001, prism_joint_target_reached = vec_length_prism < 0.001;
  sum_squares_rev_joints = 0.0; sum_squares_prism_joints = 0.0;
  for (int j = 0; j < ORC_DOF; ++j) {
    if (!is_revolute[j]) {
      if (!prism_joint_target_reached) {
        eucl_dist[j] = eucl_dist[j] / vec_length_prism;
        double c_step = extend_step_factor * eucl_dist[j];
        extend_config[j] = nn[j] + c_step;
        sum_squares_prism_joints += c_step * c_step;
      }
    } else {
      if (!rot_joint_target_reached) {
        eucl_dist[j] = eucl_dist[j] / vec_length_rot;
        double c_step = extend_step_factor * eucl_dist[j];
        extend_config[j] = nn[j] + c_step;
        sum_squares_rev_joints += c_step * c_step;
      }
    }
  }
  double extend_vec_length_prism = sum_squares_prism_joints == 0.0 ? 1000.0 : sqrt(sum_squares_prism_joints);
  double extend_vec_length_rot = sum_squares_rev_joints == 0.0 ? 1000.0 : sqrt(sum_squares_rev_joints);
  if (extend_vec_length_rot < vec_length_rot) {
    for (int j = 0; j < ORC_DOF; ++j)
      if (is_revolute[j]) x_rand[j] = extend_config[j];
    rand_sample_reached = 0;
  }
  if (extend_vec_length_prism < vec_length_prism) {
    for (int j = 0; j < ORC_DOF; ++j)
      if (!is_revolute[j]) x_rand[j] = extend_config[j];
    rand_sample_reached = 0;
  }
  return rand_sample_reached;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
