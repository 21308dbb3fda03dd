/*
 * Oracle physics: one MuJoCo-style forward + implicitfast step for one environment.  TEST INFRASTRUCTURE.
 *
 * Restates the pipeline `mjx.forward` / `mjx.step` run at ksim/engine.py:213,257 (source un-vendored:
 * mujoco-mjx, unpinned, ksim/requirements.txt:20).  Stage names below follow MuJoCo's computation docs
 * and the MJX module layout (smooth / passive / collision_primitive / constraint / solver / sensor /
 * forward).  Dense mass matrix + Cholesky as MJX uses for nv < 60; Newton solver with MJX's bracketing
 * line search; pyramidal cones; implicitfast integrator with Euler damping disabled (ksim/task/rl.py:717-748).
 *
 * PARITY UNPINNED: neither mujoco nor mujoco-mjx can be imported or built in this repository's image and the reference holds
 * no physics test vectors, so this file is pinned only by the analytic cases of tests/test_oracle_physics.py (closed forms,
 * friction / contact equilibria, momentum-conservation convergence of both robots).  scripts/gen_mjx_fixtures.py +
 * tests/test_oracle_vs_mjx_fixtures.py pin it against MJX on a machine that has jax + mujoco-mjx.
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

#if defined(ORACLE_F32)
#define SQRT sqrtf
#define SIN sinf
#define COS cosf
#define FABS fabsf
#define POW powf
#define ATAN2 atan2f
#else
#define SQRT sqrt
#define SIN sin
#define COS cos
#define FABS fabs
#define POW pow
#define ATAN2 atan2
#endif

/* ------------------------------------------------------------------------------------------- helpers */

static inline REAL dot3(const REAL* a, const REAL* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(REAL* r, const REAL* a, const REAL* b) {
  REAL x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mat_vec3(REAL* r, const REAL* m, const REAL* v) {
  REAL x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  REAL y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  REAL z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matT_vec3(REAL* r, const REAL* m, const REAL* v) {
  REAL x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  REAL y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  REAL z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}

void o_quat_mul(REAL* r, const REAL* a, const REAL* b) {
  REAL w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  REAL x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  REAL y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  REAL z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}

void o_quat_to_mat(REAL* m, const REAL* q) {
  REAL w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

static void quat_normalize(REAL* q) {
  REAL n = SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < O_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  REAL inv = (REAL)1 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}

static void quat_rot(REAL* r, const REAL* q, const REAL* v) {
  REAL m[9];
  o_quat_to_mat(m, q);
  mat_vec3(r, m, v);
}

static void axis_angle_quat(REAL* q, const REAL* axis, REAL angle) {
  REAL s = SIN(angle * (REAL)0.5);
  q[0] = COS(angle * (REAL)0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* xax.rotate_vector_by_quat [restated from memory of the un-vendored `xax` package]: the quaternion is
 * normalised with eps=1e-6 first.  The reference-owned check is tests/test_terminations.py:66-79
 * (un-normalised quat (4,5,6,7) must give arccos(z) in (0.1, 10) -- only true with normalisation). */
void o_rot_vec_quat(REAL* r, const REAL* v, const REAL* q, int inverse) {
  REAL n = SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + (REAL)1e-6;
  REAL w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  if (inverse) { x = -x; y = -y; z = -z; }
  REAL vx = v[0], vy = v[1], vz = v[2];
  REAL xx = w * w * vx + 2 * y * w * vz - 2 * z * w * vy + x * x * vx + 2 * y * x * vy + 2 * z * x * vz -
            z * z * vx - y * y * vx;
  REAL yy = 2 * x * y * vx + y * y * vy + 2 * z * y * vz + 2 * w * z * vx - z * z * vy + w * w * vy -
            2 * w * x * vz - x * x * vy;
  REAL zz = 2 * x * z * vx + 2 * y * z * vy + z * z * vz - 2 * w * y * vx + w * w * vz + 2 * w * x * vy -
            y * y * vz - x * x * vz;
  r[0] = xx; r[1] = yy; r[2] = zz;
}

/* xax.euler_to_quat([roll, pitch, yaw]) -> wxyz, ZYX composition [restated from memory]. */
void o_euler_to_quat(REAL* q, const REAL* rpy) {
  REAL cr = COS(rpy[0] * (REAL)0.5), sr = SIN(rpy[0] * (REAL)0.5);
  REAL cp = COS(rpy[1] * (REAL)0.5), sp = SIN(rpy[1] * (REAL)0.5);
  REAL cy = COS(rpy[2] * (REAL)0.5), sy = SIN(rpy[2] * (REAL)0.5);
  q[0] = cr * cp * cy + sr * sp * sy;
  q[1] = sr * cp * cy - cr * sp * sy;
  q[2] = cr * sp * cy + sr * cp * sy;
  q[3] = cr * cp * sy - sr * sp * cy;
}

/* spatial algebra; 6-vectors are [rotational(3), translational(3)] as in MuJoCo */
static void mul_inert_vec(REAL* r, const REAL* i, const REAL* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void cross_motion(REAL* r, const REAL* vel, const REAL* v) {
  REAL a[3], b[3], c[3];
  cross3(a, vel, v);
  cross3(b, vel, v + 3);
  cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(REAL* r, const REAL* vel, const REAL* f) {
  REAL a[3], b[3], c[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* ------------------------------------------------------------------------------------- position stage */

static void kinematics(const OModel* m, OData* d) {
  for (int k = 0; k < 3; k++) d->xpos[0][k] = 0;
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  o_quat_to_mat(d->xmat[0], d->xquat[0]);
  for (int k = 0; k < 3; k++) d->xipos[0][k] = 0;
  o_quat_to_mat(d->ximat[0], d->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    int jn = m->body_jntnum[b], ja = m->body_jntadr[b];
    REAL pos[3], quat[4];
    if (jn == 1 && m->jnt_type[ja] == O_JNT_FREE) {
      int qa = m->jnt_qposadr[ja];
      quat_normalize(d->qpos + qa + 3); /* written back, as mj_kinematics does */
      for (int k = 0; k < 3; k++) pos[k] = d->qpos[qa + k];
      for (int k = 0; k < 4; k++) quat[k] = d->qpos[qa + 3 + k];
      for (int k = 0; k < 3; k++) { d->xanchor[ja][k] = pos[k]; d->xaxis[ja][k] = m->jnt_axis[ja][k]; }
    } else {
      REAL t[3];
      mat_vec3(t, d->xmat[p], m->body_pos[b]);
      for (int k = 0; k < 3; k++) pos[k] = d->xpos[p][k] + t[k];
      o_quat_mul(quat, d->xquat[p], m->body_quat[b]);
      for (int j = ja; j < ja + jn; j++) {
        REAL anchor[3], axis[3], qloc[4], qn[4], v[3];
        quat_rot(anchor, quat, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) anchor[k] += pos[k];
        quat_rot(axis, quat, m->jnt_axis[j]);
        for (int k = 0; k < 3; k++) { d->xanchor[j][k] = anchor[k]; d->xaxis[j][k] = axis[k]; }
        int qa = m->jnt_qposadr[j];
        axis_angle_quat(qloc, m->jnt_axis[j], d->qpos[qa] - m->qpos0[qa]);
        o_quat_mul(qn, quat, qloc);
        for (int k = 0; k < 4; k++) quat[k] = qn[k];
        quat_rot(v, quat, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) pos[k] = anchor[k] - v[k];
      }
    }
    quat_normalize(quat);
    for (int k = 0; k < 3; k++) d->xpos[b][k] = pos[k];
    for (int k = 0; k < 4; k++) d->xquat[b][k] = quat[k];
    o_quat_to_mat(d->xmat[b], quat);
    REAL t[3], iq[4];
    mat_vec3(t, d->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; k++) d->xipos[b][k] = pos[k] + t[k];
    o_quat_mul(iq, quat, m->body_iquat[b]);
    o_quat_to_mat(d->ximat[b], iq);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    REAL t[3], q[4];
    mat_vec3(t, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + t[k];
    o_quat_mul(q, d->xquat[b], m->geom_quat[g]);
    o_quat_to_mat(d->geom_xmat[g], q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    REAL t[3], q[4];
    mat_vec3(t, d->xmat[b], m->site_pos[s]);
    for (int k = 0; k < 3; k++) d->site_xpos[s][k] = d->xpos[b][k] + t[k];
    o_quat_mul(q, d->xquat[b], m->site_quat[s]);
    o_quat_to_mat(d->site_xmat[s], q);
  }
}

static void com_pos(const OModel* m, OData* d) {
  int nb = m->nbody;
  for (int b = 0; b < nb; b++)
    for (int k = 0; k < 3; k++) d->subtree_com[b][k] = m->body_mass[b] * d->xipos[b][k];
  for (int b = nb - 1; b > 0; b--)
    for (int k = 0; k < 3; k++) d->subtree_com[m->body_parentid[b]][k] += d->subtree_com[b][k];
  for (int b = 0; b < nb; b++) {
    /* nominal body_subtreemass even when body_mass is randomised (ksim/randomization.py:351-355) */
    REAL sm = m->body_subtreemass[b] > O_MINVAL ? m->body_subtreemass[b] : O_MINVAL;
    for (int k = 0; k < 3; k++) d->subtree_com[b][k] /= sm;
  }
  memset(d->cinert[0], 0, sizeof(d->cinert[0]));
  for (int b = 1; b < nb; b++) {
    const REAL* com = d->subtree_com[m->body_rootid[b]];
    REAL off[3], mass = m->body_mass[b];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[b][k] - com[k];
    /* R diag(I) R^T */
    const REAL* R = d->ximat[b];
    const REAL* I = m->body_inertia[b];
    REAL full[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        full[3 * r + c] = R[3 * r] * I[0] * R[3 * c] + R[3 * r + 1] * I[1] * R[3 * c + 1] + R[3 * r + 2] * I[2] * R[3 * c + 2];
    REAL dd = dot3(off, off);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) full[3 * r + c] += mass * ((r == c ? dd : 0) - off[r] * off[c]);
    REAL* ci = d->cinert[b];
    ci[0] = full[0]; ci[1] = full[4]; ci[2] = full[8]; ci[3] = full[1]; ci[4] = full[2]; ci[5] = full[5];
    ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2]; ci[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    const REAL* com = d->subtree_com[m->body_rootid[b]];
    REAL off[3];
    for (int k = 0; k < 3; k++) off[k] = com[k] - d->xanchor[j][k];
    if (m->jnt_type[j] == O_JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        memset(d->cdof[da + k], 0, 6 * sizeof(REAL));
        d->cdof[da + k][3 + k] = 1;
        REAL ax[3] = {d->xmat[b][k], d->xmat[b][3 + k], d->xmat[b][6 + k]};
        for (int c = 0; c < 3; c++) d->cdof[da + 3 + k][c] = ax[c];
        cross3(d->cdof[da + 3 + k] + 3, ax, off);
      }
    } else {
      for (int c = 0; c < 3; c++) d->cdof[da][c] = d->xaxis[j][c];
      cross3(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

static void crb(const OModel* m, OData* d) {
  int nb = m->nbody, nv = m->nv;
  for (int b = 0; b < nb; b++) memcpy(d->crb[b], d->cinert[b], 10 * sizeof(REAL));
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0)
      for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
  }
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) d->qM[i][j] = 0;
  for (int i = 0; i < nv; i++) {
    REAL buf[6];
    mul_inert_vec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      REAL s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
      d->qM[i][j] = d->qM[j][i] = s;
    }
    d->qM[i][i] += m->dof_armature[i];
  }
}

/* dense Cholesky A = L L^T (lower); returns 0 on success */
static int cholesky(int n, REAL A[][O_MAXNV], REAL L[][O_MAXNV]) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) L[i][j] = 0;
  for (int j = 0; j < n; j++) {
    REAL s = A[j][j];
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (s < O_MINVAL) s = O_MINVAL;
    REAL dj = SQRT(s);
    L[j][j] = dj;
    for (int i = j + 1; i < n; i++) {
      REAL t = A[i][j];
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t / dj;
    }
  }
  return 0;
}
static void chol_solve(int n, REAL L[][O_MAXNV], REAL* x, const REAL* b) {
  REAL y[O_MAXNV];
  for (int i = 0; i < n; i++) {
    REAL s = b[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    REAL s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
}

/* ------------------------------------------------------------------------------------------ collision */

static void make_frame(REAL* frame, const REAL* n) {
  /* rows: normal, tangent1, tangent2 (mjx math.make_frame: Gram-Schmidt against y or z) */
  REAL a[3] = {n[0], n[1], n[2]};
  REAL na = SQRT(dot3(a, a));
  for (int k = 0; k < 3; k++) a[k] /= na;
  REAL y[3] = {0, 1, 0}, z[3] = {0, 0, 1};
  const REAL* ref = (FABS(a[1]) < (REAL)0.5) ? y : z;
  REAL b[3];
  REAL dp = dot3(a, ref);
  for (int k = 0; k < 3; k++) b[k] = ref[k] - a[k] * dp;
  REAL nbm = SQRT(dot3(b, b));
  for (int k = 0; k < 3; k++) b[k] /= nbm;
  REAL c[3];
  cross3(c, a, b);
  for (int k = 0; k < 3; k++) { frame[k] = a[k]; frame[3 + k] = b[k]; frame[6 + k] = c[k]; }
}

static void add_contact(const OModel* m, OData* d, int pair, REAL dist, const REAL* pos, const REAL* n) {
  if (d->ncon >= O_MAXCON) return;
  OContact* c = &d->contact[d->ncon++];
  c->pair = pair;
  c->geom1 = m->pair_geom1[pair];
  c->geom2 = m->pair_geom2[pair];
  c->dim = m->pair_dim[pair];
  c->dist = dist;
  for (int k = 0; k < 3; k++) c->pos[k] = pos[k];
  make_frame(c->frame, n);
  c->includemargin = m->pair_margin[pair] - m->pair_gap[pair];
  for (int k = 0; k < 5; k++) { c->friction[k] = m->pair_friction[pair][k]; c->solimp[k] = m->pair_solimp[pair][k]; }
  c->solref[0] = m->pair_solref[pair][0]; c->solref[1] = m->pair_solref[pair][1];
  c->efc_address = -1;
}

static void plane_sphere(const REAL* ppos, const REAL* pn, const REAL* spos, REAL r, REAL* dist, REAL* pos) {
  REAL dv[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  *dist = dot3(dv, pn) - r;
  for (int k = 0; k < 3; k++) pos[k] = spos[k] - pn[k] * (r + (REAL)0.5 * (*dist));
}

/* MJX keeps every candidate contact of every pair (fixed shapes) -- rows whose dist >= margin are masked,
 * not dropped; `contact.dist` is what ksim's geoms_colliding reads (ksim/utils/mujoco.py:175-217). */
static void collision(const OModel* m, OData* d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    if (t1 == O_GEOM_PLANE && t2 == O_GEOM_SPHERE) {
      REAL n[3] = {d->geom_xmat[g1][2], d->geom_xmat[g1][5], d->geom_xmat[g1][8]};
      REAL dist, pos[3];
      plane_sphere(d->geom_xpos[g1], n, d->geom_xpos[g2], m->geom_size[g2][0], &dist, pos);
      add_contact(m, d, p, dist, pos, n);
    } else if (t1 == O_GEOM_PLANE && t2 == O_GEOM_CAPSULE) {
      /* mjx plane_capsule: two sphere tests at the segment ends; the contact frame's first tangent is
       * aligned with the capsule axis -- the frame only matters for condim>1 tangent directions. */
      REAL n[3] = {d->geom_xmat[g1][2], d->geom_xmat[g1][5], d->geom_xmat[g1][8]};
      REAL axis[3] = {d->geom_xmat[g2][2], d->geom_xmat[g2][5], d->geom_xmat[g2][8]};
      REAL r = m->geom_size[g2][0], half = m->geom_size[g2][1];
      for (int s = -1; s <= 1; s += 2) {
        REAL end[3], dist, pos[3];
        for (int k = 0; k < 3; k++) end[k] = d->geom_xpos[g2][k] + axis[k] * half * (REAL)s;
        plane_sphere(d->geom_xpos[g1], n, end, r, &dist, pos);
        add_contact(m, d, p, dist, pos, n);
        /* align tangent with the capsule axis projected on the plane, when well defined */
        OContact* c = &d->contact[d->ncon - 1];
        REAL dp = dot3(axis, n), b[3];
        for (int k = 0; k < 3; k++) b[k] = axis[k] - n[k] * dp;
        REAL nb = SQRT(dot3(b, b));
        if (nb > (REAL)0.5) {
          for (int k = 0; k < 3; k++) b[k] /= nb;
          REAL cc[3];
          cross3(cc, n, b);
          for (int k = 0; k < 3; k++) { c->frame[3 + k] = b[k]; c->frame[6 + k] = cc[k]; }
        }
      }
    } else if (t1 == O_GEOM_SPHERE && t2 == O_GEOM_SPHERE) {
      REAL dv[3], pos[3], n[3];
      for (int k = 0; k < 3; k++) dv[k] = d->geom_xpos[g2][k] - d->geom_xpos[g1][k];
      REAL len = SQRT(dot3(dv, dv));
      if (len < O_MINVAL) { n[0] = 1; n[1] = n[2] = 0; } else { for (int k = 0; k < 3; k++) n[k] = dv[k] / len; }
      REAL dist = len - m->geom_size[g1][0] - m->geom_size[g2][0];
      for (int k = 0; k < 3; k++) pos[k] = d->geom_xpos[g1][k] + n[k] * (m->geom_size[g1][0] + (REAL)0.5 * dist);
      add_contact(m, d, p, dist, pos, n);
    }
  }
}

/* -------------------------------------------------------------------------------------- constraints */

static void body_jac_point(const OModel* m, const OData* d, int body, const REAL* point, REAL jacp[3][O_MAXNV],
                           REAL jacr[3][O_MAXNV]) {
  for (int r = 0; r < 3; r++)
    for (int i = 0; i < m->nv; i++) { jacp[r][i] = 0; if (jacr) jacr[r][i] = 0; }
  int b = body;
  while (b != 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
  if (b == 0) return;
  const REAL* com = d->subtree_com[m->body_rootid[body]];
  REAL off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]};
  for (int i = m->body_dofadr[b] + m->body_dofnum[b] - 1; i >= 0; i = m->dof_parentid[i]) {
    REAL t[3];
    cross3(t, d->cdof[i], off);
    for (int r = 0; r < 3; r++) {
      jacp[r][i] = d->cdof[i][3 + r] + t[r];
      if (jacr) jacr[r][i] = d->cdof[i][r];
    }
  }
}

static void kbi(const OModel* m, const REAL* solref, const REAL* solimp, REAL pos, REAL* k, REAL* b, REAL* imp) {
  REAL timeconst = solref[0], dampratio = solref[1];
  if (!m->disable_refsafe && solref[0] > 0) {
    REAL lo = 2 * m->timestep;
    if (timeconst < lo) timeconst = lo;
  }
  REAL dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin < O_MINIMP) dmin = O_MINIMP; if (dmin > O_MAXIMP) dmin = O_MAXIMP;
  if (dmax < O_MINIMP) dmax = O_MINIMP; if (dmax > O_MAXIMP) dmax = O_MAXIMP;
  if (width < O_MINVAL) width = O_MINVAL;
  if (mid < O_MINIMP) mid = O_MINIMP; if (mid > O_MAXIMP) mid = O_MAXIMP;
  if (power < 1) power = 1;
  REAL kk = (REAL)1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  REAL bb = (REAL)2 / (dmax * timeconst);
  if (solref[0] <= 0) kk = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) bb = -solref[1] / dmax;
  REAL x = FABS(pos) / width, y;
  if (x < mid) y = ((REAL)1 / POW(mid, power - 1)) * POW(x, power);
  else y = (REAL)1 - ((REAL)1 / POW((REAL)1 - mid, power - 1)) * POW((REAL)1 - x, power);
  if (x > 1) y = 1; /* also guards pow of a negative base */
  REAL im = dmin + y * (dmax - dmin);
  if (im < dmin) im = dmin; if (im > dmax) im = dmax;
  if (x > 1) im = dmax;
  *k = kk; *b = bb; *imp = im;
}

static void add_row(const OModel* m, OData* d, int type, int id, const REAL* J, REAL pos, REAL margin_pos,
                    REAL invweight, const REAL* solref, const REAL* solimp, REAL frictionloss) {
  if (d->nefc >= O_MAXEFC) return;
  int r = d->nefc++;
  int nv = m->nv;
  REAL vel = 0;
  for (int i = 0; i < nv; i++) { d->efc_J[r][i] = J[i]; vel += J[i] * d->qvel[i]; }
  REAL k, b, imp;
  kbi(m, solref, solimp, margin_pos, &k, &b, &imp);
  REAL R = invweight * ((REAL)1 - imp) / imp;
  if (R < O_MINVAL) R = O_MINVAL;
  d->efc_type[r] = type; d->efc_id[r] = id;
  d->efc_pos[r] = pos;
  d->efc_aref[r] = -b * vel - k * imp * margin_pos;
  d->efc_R[r] = R; d->efc_D[r] = (REAL)1 / R;
  d->efc_frictionloss[r] = frictionloss;
  d->efc_force[r] = 0;
}

/* Row order follows MJX make_constraint: (equality,) friction, limit, contact.  MJX instantiates every
 * candidate row and zeroes the inactive ones; zero rows contribute nothing to cost, gradient or Hessian,
 * so only active rows are materialised here. */
static void make_constraint(const OModel* m, OData* d) {
  int nv = m->nv;
  d->nefc = 0;
  REAL J[O_MAXNV];
  for (int i = 0; i < nv; i++) {
    if (m->dof_frictionloss[i] > 0) {
      for (int k = 0; k < nv; k++) J[k] = 0;
      J[i] = 1;
      add_row(m, d, O_EFC_FRICTION, i, J, 0, 0, m->dof_invweight0[i], m->dof_solref[i], m->dof_solimp[i],
              m->dof_frictionloss[i]);
    }
  }
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] != O_JNT_HINGE || !m->jnt_limited[j]) continue;
    REAL q = d->qpos[m->jnt_qposadr[j]];
    REAL dmin = q - m->jnt_range[j][0], dmax = m->jnt_range[j][1] - q;
    REAL pos = (dmin < dmax ? dmin : dmax) - m->jnt_margin[j];
    if (pos < 0) {
      for (int k = 0; k < nv; k++) J[k] = 0;
      J[m->jnt_dofadr[j]] = (dmin < dmax) ? (REAL)1 : (REAL)-1;
      add_row(m, d, O_EFC_LIMIT, j, J, pos, pos, m->dof_invweight0[m->jnt_dofadr[j]], m->jnt_solref[j],
              m->jnt_solimp[j], 0);
    }
  }
  for (int c = 0; c < d->ncon; c++) {
    OContact* con = &d->contact[c];
    con->efc_address = -1;
    REAL pos = con->dist - con->includemargin;
    if (!(pos < 0)) continue;
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    REAL j1[3][O_MAXNV], j2[3][O_MAXNV], diff[3][O_MAXNV];
    body_jac_point(m, d, b1, con->pos, j1, 0);
    body_jac_point(m, d, b2, con->pos, j2, 0);
    /* rotate Jacobian difference into the contact frame */
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < nv; i++) {
        REAL s = 0;
        for (int k = 0; k < 3; k++) s += con->frame[3 * r + k] * (j2[k][i] - j1[k][i]);
        diff[r][i] = s;
      }
    REAL t = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    con->efc_address = d->nefc;
    if (con->dim == 1) {
      add_row(m, d, O_EFC_CONTACT, c, diff[0], pos, pos, t, con->solref, con->solimp, 0);
    } else {
      /* pyramidal, condim 3: rows n +/- mu_k t_k; R shared: 2 mu0^2 (1-imp)/imp (t + mu0^2 t) / impratio */
      REAL mu0 = con->friction[0];
      REAL iw = (t + mu0 * mu0 * t) * 2 * mu0 * mu0 / m->impratio;
      for (int k = 0; k < 2; k++) {
        REAL mu = con->friction[k];
        for (int s = 1; s >= -1; s -= 2) {
          for (int i = 0; i < nv; i++) J[i] = diff[0][i] + (REAL)s * mu * diff[1 + k][i];
          add_row(m, d, O_EFC_CONTACT, c, J, pos, pos, iw, con->solref, con->solimp, 0);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ velocity stage */

static void com_vel(const OModel* m, OData* d) {
  memset(d->cvel[0], 0, 6 * sizeof(REAL));
  for (int b = 1; b < m->nbody; b++) {
    REAL v[6];
    memcpy(v, d->cvel[m->body_parentid[b]], sizeof(v));
    int da = m->body_dofadr[b], dn = m->body_dofnum[b];
    int ja = m->body_jntadr[b];
    if (dn == 6 && m->jnt_type[ja] == O_JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        memset(d->cdof_dot[da + k], 0, 6 * sizeof(REAL));
        for (int c = 0; c < 6; c++) v[c] += d->cdof[da + k][c] * d->qvel[da + k];
      }
      for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot[da + k], v, d->cdof[da + k]);
      for (int k = 3; k < 6; k++)
        for (int c = 0; c < 6; c++) v[c] += d->cdof[da + k][c] * d->qvel[da + k];
    } else {
      for (int k = 0; k < dn; k++) {
        cross_motion(d->cdof_dot[da + k], v, d->cdof[da + k]);
        for (int c = 0; c < 6; c++) v[c] += d->cdof[da + k][c] * d->qvel[da + k];
      }
    }
    memcpy(d->cvel[b], v, sizeof(v));
  }
}

static void passive(const OModel* m, OData* d) {
  for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] != O_JNT_HINGE) continue; /* free-joint stiffness not supported (always 0 here) */
    int qa = m->jnt_qposadr[j];
    d->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
  }
}

/* recursive Newton-Euler; with_acc adds cdof*qacc (rne_postconstraint's cacc) */
static void rne_forward_acc(const OModel* m, OData* d, int with_acc) {
  for (int k = 0; k < 3; k++) { d->cacc[0][k] = 0; d->cacc[0][3 + k] = -m->gravity[k]; }
  for (int b = 1; b < m->nbody; b++) {
    memcpy(d->cacc[b], d->cacc[m->body_parentid[b]], 6 * sizeof(REAL));
    int da = m->body_dofadr[b];
    for (int k = 0; k < m->body_dofnum[b]; k++)
      for (int c = 0; c < 6; c++) {
        d->cacc[b][c] += d->cdof_dot[da + k][c] * d->qvel[da + k];
        if (with_acc) d->cacc[b][c] += d->cdof[da + k][c] * d->qacc[da + k];
      }
  }
}

static void rne(const OModel* m, OData* d) {
  REAL cfrc[O_MAXNB][6];
  rne_forward_acc(m, d, 0);
  memset(cfrc[0], 0, sizeof(cfrc[0]));
  for (int b = 1; b < m->nbody; b++) {
    REAL t1[6], t2[6], t3[6];
    mul_inert_vec(t1, d->cinert[b], d->cacc[b]);
    mul_inert_vec(t2, d->cinert[b], d->cvel[b]);
    cross_force(t3, d->cvel[b], t2);
    for (int c = 0; c < 6; c++) cfrc[b][c] = t1[c] + t3[c];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int c = 0; c < 6; c++) cfrc[p][c] += cfrc[b][c];
  }
  for (int i = 0; i < m->nv; i++) {
    REAL s = 0;
    for (int c = 0; c < 6; c++) s += d->cdof[i][c] * cfrc[m->dof_bodyid[i]][c];
    d->qfrc_bias[i] = s;
  }
}

static void actuation(const OModel* m, OData* d) {
  for (int i = 0; i < m->nv; i++) d->qfrc_actuator[i] = 0;
  for (int u = 0; u < m->nu; u++) {
    REAL c = d->ctrl[u];
    if (m->actuator_ctrllimited[u]) {
      if (c < m->actuator_ctrlrange[u][0]) c = m->actuator_ctrlrange[u][0];
      if (c > m->actuator_ctrlrange[u][1]) c = m->actuator_ctrlrange[u][1];
    }
    REAL f = c; /* motor: gain 1, no bias */
    if (m->actuator_forcelimited[u]) {
      if (f < m->actuator_forcerange[u][0]) f = m->actuator_forcerange[u][0];
      if (f > m->actuator_forcerange[u][1]) f = m->actuator_forcerange[u][1];
    }
    d->actuator_force[u] = f;
    d->qfrc_actuator[m->jnt_dofadr[m->actuator_trnid[u]]] += m->actuator_gear[u] * f;
  }
}

static void acceleration(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  /* xfrc_applied: Cartesian wrench [force(3), torque(3)] at the body com (mj_xfrcAccumulate) */
  for (int b = 1; b < m->nbody; b++) {
    const REAL* f = d->xfrc_applied[b];
    if (f[0] == 0 && f[1] == 0 && f[2] == 0 && f[3] == 0 && f[4] == 0 && f[5] == 0) continue;
    REAL jp[3][O_MAXNV], jr[3][O_MAXNV];
    body_jac_point(m, d, b, d->xipos[b], jp, jr);
    for (int i = 0; i < nv; i++)
      for (int k = 0; k < 3; k++) d->qfrc_smooth[i] += jp[k][i] * f[k] + jr[k][i] * f[3 + k];
  }
  chol_solve(nv, d->qL, d->qacc_smooth, d->qfrc_smooth);
}

/* --------------------------------------------------------------------------------------------- solver */

typedef struct {
  REAL qacc[O_MAXNV], Ma[O_MAXNV], Jaref[O_MAXEFC], grad[O_MAXNV], Mgrad[O_MAXNV], search[O_MAXNV];
  REAL qfrc_constraint[O_MAXNV], efc_force[O_MAXEFC];
  int active[O_MAXEFC];
  REAL gauss, cost, prev_cost;
  int niter;
} SolverCtx;

static void ctx_update_constraint(const OModel* m, const OData* d, SolverCtx* c) {
  int nv = m->nv, ne = d->nefc;
  REAL cost = 0;
  for (int r = 0; r < ne; r++) {
    REAL x = c->Jaref[r], D = d->efc_D[r];
    if (d->efc_type[r] == O_EFC_FRICTION) {
      REAL f = d->efc_frictionloss[r], rf = d->efc_R[r] * f;
      if (x <= -rf) { c->efc_force[r] = f; c->active[r] = 0; cost += -f * ((REAL)0.5 * rf + x); }
      else if (x >= rf) { c->efc_force[r] = -f; c->active[r] = 0; cost += -f * ((REAL)0.5 * rf - x); }
      else { c->efc_force[r] = -D * x; c->active[r] = 1; cost += (REAL)0.5 * D * x * x; }
    } else {
      int act = x < 0;
      c->active[r] = act;
      c->efc_force[r] = act ? -D * x : 0;
      if (act) cost += (REAL)0.5 * D * x * x;
    }
  }
  for (int i = 0; i < nv; i++) {
    REAL s = 0;
    for (int r = 0; r < ne; r++) s += d->efc_J[r][i] * c->efc_force[r];
    c->qfrc_constraint[i] = s;
  }
  REAL gauss = 0;
  for (int i = 0; i < nv; i++) gauss += (c->Ma[i] - d->qfrc_smooth[i]) * (c->qacc[i] - d->qacc_smooth[i]);
  c->gauss = (REAL)0.5 * gauss;
  c->prev_cost = c->cost;
  c->cost = c->gauss + cost;
}

static void ctx_update_gradient(const OModel* m, const OData* d, SolverCtx* c) {
  int nv = m->nv, ne = d->nefc;
  static _Thread_local REAL H[O_MAXNV][O_MAXNV], L[O_MAXNV][O_MAXNV];
  for (int i = 0; i < nv; i++) c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - c->qfrc_constraint[i];
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) H[i][j] = d->qM[i][j];
  for (int r = 0; r < ne; r++) {
    if (!c->active[r]) continue;
    REAL D = d->efc_D[r];
    for (int i = 0; i < nv; i++) {
      REAL ji = d->efc_J[r][i];
      if (ji == 0) continue;
      for (int j = 0; j < nv; j++) H[i][j] += D * ji * d->efc_J[r][j];
    }
  }
  cholesky(nv, H, L);
  chol_solve(nv, L, c->Mgrad, c->grad);
}

static void ctx_create(const OModel* m, const OData* d, SolverCtx* c, const REAL* qacc, int grad) {
  int nv = m->nv, ne = d->nefc;
  for (int i = 0; i < nv; i++) c->qacc[i] = qacc[i];
  for (int i = 0; i < nv; i++) {
    REAL s = 0;
    for (int j = 0; j < nv; j++) s += d->qM[i][j] * qacc[j];
    c->Ma[i] = s;
  }
  for (int r = 0; r < ne; r++) {
    REAL s = 0;
    for (int j = 0; j < nv; j++) s += d->efc_J[r][j] * qacc[j];
    c->Jaref[r] = s - d->efc_aref[r];
  }
  c->cost = INFINITY; c->prev_cost = INFINITY; c->niter = 0;
  ctx_update_constraint(m, d, c);
  if (grad) {
    ctx_update_gradient(m, d, c);
    for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
  }
}

typedef struct { REAL alpha, cost, deriv0, deriv1; } LSPoint;

static LSPoint ls_point(const OData* d, const SolverCtx* c, REAL alpha, const REAL* jv, const REAL* quad_gauss) {
  int ne = d->nefc;
  REAL q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  for (int r = 0; r < ne; r++) {
    REAL ja = c->Jaref[r], D = d->efc_D[r];
    REAL x = ja + alpha * jv[r];
    if (d->efc_type[r] == O_EFC_FRICTION) {
      REAL f = d->efc_frictionloss[r], rf = d->efc_R[r] * f;
      if (x <= -rf) { q0 += f * (-(REAL)0.5 * rf - ja); q1 += -f * jv[r]; }
      else if (x >= rf) { q0 += f * (-(REAL)0.5 * rf + ja); q1 += f * jv[r]; }
      else { q0 += (REAL)0.5 * D * ja * ja; q1 += D * jv[r] * ja; q2 += (REAL)0.5 * D * jv[r] * jv[r]; }
    } else if (x < 0) {
      q0 += (REAL)0.5 * D * ja * ja; q1 += D * jv[r] * ja; q2 += (REAL)0.5 * D * jv[r] * jv[r];
    }
  }
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv0 = 2 * alpha * q2 + q1;
  p.deriv1 = 2 * q2 + (q2 == 0 ? O_MINVAL : 0);
  return p;
}

static int in_bracket(LSPoint x, LSPoint y) {
  return (x.deriv0 < y.deriv0 && y.deriv0 < 0) || (x.deriv0 > y.deriv0 && y.deriv0 > 0);
}

/* MJX solver._linesearch [restated from memory]: Newton-step bracketing with three candidates per
 * iteration (lo Newton step, hi Newton step, midpoint), at most ls_iterations rounds. */
static void linesearch(const OModel* m, const OData* d, SolverCtx* c) {
  int nv = m->nv, ne = d->nefc;
  REAL mv[O_MAXNV], jv[O_MAXEFC];
  REAL snorm = 0;
  for (int i = 0; i < nv; i++) snorm += c->search[i] * c->search[i];
  snorm = SQRT(snorm);
  REAL scale = m->meaninertia * (REAL)(nv > 1 ? nv : 1);
  REAL gtol = m->tolerance * m->ls_tolerance * snorm * scale;
  for (int i = 0; i < nv; i++) {
    REAL s = 0;
    for (int j = 0; j < nv; j++) s += d->qM[i][j] * c->search[j];
    mv[i] = s;
  }
  for (int r = 0; r < ne; r++) {
    REAL s = 0;
    for (int j = 0; j < nv; j++) s += d->efc_J[r][j] * c->search[j];
    jv[r] = s;
  }
  REAL qg[3] = {c->gauss, 0, 0};
  for (int i = 0; i < nv; i++) {
    qg[1] += c->search[i] * (c->Ma[i] - d->qfrc_smooth[i]);
    qg[2] += (REAL)0.5 * c->search[i] * mv[i];
  }
  LSPoint p0 = ls_point(d, c, 0, jv, qg);
  LSPoint lo_in = p0;
  LSPoint lo = ls_point(d, c, lo_in.alpha - lo_in.deriv0 / lo_in.deriv1, jv, qg);
  LSPoint hi;
  if (lo_in.deriv0 < lo.deriv0) { hi = lo; lo = lo_in; } else { hi = lo_in; }
  int swap = 1, it = 0;
  while (1) {
    int done = it >= m->ls_iterations;
    done |= !swap;
    done |= (lo.deriv0 < 0) && (lo.deriv0 > -gtol);
    done |= (hi.deriv0 > 0) && (hi.deriv0 < gtol);
    if (done) break;
    LSPoint lo_next = ls_point(d, c, lo.alpha - lo.deriv0 / lo.deriv1, jv, qg);
    LSPoint hi_next = ls_point(d, c, hi.alpha - hi.deriv0 / hi.deriv1, jv, qg);
    LSPoint mid = ls_point(d, c, (REAL)0.5 * (lo.alpha + hi.alpha), jv, qg);
    int s1 = in_bracket(lo, lo_next); if (s1) lo = lo_next;
    int s2 = in_bracket(lo, mid); if (s2) lo = mid;
    int s3 = in_bracket(lo, hi_next); if (s3) lo = hi_next;
    int s4 = in_bracket(hi, hi_next); if (s4) hi = hi_next;
    int s5 = in_bracket(hi, mid); if (s5) hi = mid;
    int s6 = in_bracket(hi, lo_next); if (s6) hi = lo_next;
    swap = s1 | s2 | s3 | s4 | s5 | s6;
    it++;
  }
  int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  REAL alpha = (lo.cost < hi.cost) ? lo.alpha : hi.alpha;
  if (!improved) alpha = 0;
  for (int i = 0; i < nv; i++) { c->qacc[i] += alpha * c->search[i]; c->Ma[i] += alpha * mv[i]; }
  for (int r = 0; r < ne; r++) c->Jaref[r] += alpha * jv[r];
}

static void solve(const OModel* m, OData* d) {
  int nv = m->nv, ne = d->nefc;
  if (ne == 0) {
    /* MJX forward(): no constraint rows -> qacc = qacc_smooth; qacc_warmstart keeps its old value */
    for (int i = 0; i < nv; i++) { d->qacc[i] = d->qacc_smooth[i]; d->qfrc_constraint[i] = 0; }
    d->solver_niter = 0;
    return;
  }
  static _Thread_local SolverCtx ctx, warm, smth;
  const REAL* start = d->qacc_smooth;
  if (!m->disable_warmstart) {
    ctx_create(m, d, &warm, d->qacc_warmstart, 0);
    ctx_create(m, d, &smth, d->qacc_smooth, 0);
    if (warm.cost < smth.cost) start = d->qacc_warmstart;
  }
  ctx_create(m, d, &ctx, start, 1);
  REAL scale = (REAL)1 / (m->meaninertia * (REAL)(nv > 1 ? nv : 1));
  while (1) {
    if (m->iterations != 1 || ctx.niter > 0) {
      REAL improvement = (ctx.prev_cost - ctx.cost) * scale;
      REAL gn = 0;
      for (int i = 0; i < nv; i++) gn += ctx.grad[i] * ctx.grad[i];
      REAL gradient = SQRT(gn) * scale;
      int done = ctx.niter >= m->iterations;
      done |= improvement < m->tolerance;
      done |= gradient < m->tolerance;
      if (done) break;
    }
    linesearch(m, d, &ctx);
    ctx_update_constraint(m, d, &ctx);
    ctx_update_gradient(m, d, &ctx);
    for (int i = 0; i < nv; i++) ctx.search[i] = -ctx.Mgrad[i];
    ctx.niter++;
    if (m->iterations == 1) break;
  }
  for (int i = 0; i < nv; i++) {
    d->qacc[i] = ctx.qacc[i];
    d->qacc_warmstart[i] = ctx.qacc[i];
    d->qfrc_constraint[i] = ctx.qfrc_constraint[i];
  }
  for (int r = 0; r < ne; r++) d->efc_force[r] = ctx.efc_force[r];
  d->solver_niter = ctx.niter;
}

/* -------------------------------------------------------------------------------------------- sensors */

/* mju_transformSpatial for a motion vector: move reference point old->new, optionally rotate by rot^T */
static void transform_motion(REAL* res, const REAL* vec, const REAL* newpos, const REAL* oldpos, const REAL* rot) {
  REAL dif[3] = {newpos[0] - oldpos[0], newpos[1] - oldpos[1], newpos[2] - oldpos[2]};
  REAL cr[3], lin[3];
  cross3(cr, dif, vec);
  for (int k = 0; k < 3; k++) lin[k] = vec[3 + k] - cr[k];
  if (rot) { matT_vec3(res, rot, vec); matT_vec3(res + 3, rot, lin); }
  else { for (int k = 0; k < 3; k++) { res[k] = vec[k]; res[3 + k] = lin[k]; } }
}

static void object_frame(const OModel* m, const OData* d, int objtype, int objid, int* body, const REAL** pos,
                         const REAL** rot) {
  if (objtype == O_OBJ_SITE) { *body = m->site_bodyid[objid]; *pos = d->site_xpos[objid]; *rot = d->site_xmat[objid]; }
  else if (objtype == O_OBJ_BODY) { *body = objid; *pos = d->xipos[objid]; *rot = d->ximat[objid]; }
  else { *body = objid; *pos = d->xpos[objid]; *rot = d->xmat[objid]; }
}

static void object_velocity(const OModel* m, const OData* d, int objtype, int objid, REAL* res, int local) {
  int body; const REAL *pos, *rot;
  object_frame(m, d, objtype, objid, &body, &pos, &rot);
  transform_motion(res, d->cvel[body], pos, d->subtree_com[m->body_rootid[body]], local ? rot : 0);
}

static void object_acceleration(const OModel* m, const OData* d, int objtype, int objid, REAL* res, int local) {
  int body; const REAL *pos, *rot;
  object_frame(m, d, objtype, objid, &body, &pos, &rot);
  const REAL* com = d->subtree_com[m->body_rootid[body]];
  REAL vel[6], cr[3];
  transform_motion(vel, d->cvel[body], pos, com, local ? rot : 0);
  transform_motion(res, d->cacc[body], pos, com, local ? rot : 0);
  cross3(cr, vel, vel + 3);
  for (int k = 0; k < 3; k++) res[3 + k] += cr[k];
}

static void sensor_pos(const OModel* m, OData* d) {
  for (int s = 0; s < m->nsensor; s++) {
    REAL* out = d->sensordata + m->sensor_adr[s];
    int t = m->sensor_type[s], ot = m->sensor_objtype[s], id = m->sensor_objid[s];
    int body; const REAL *pos, *rot;
    if (t == O_SENS_MAGNETOMETER) {
      /* mj_sensorPos: site_xmat^T * opt.magnetic, MuJoCo's default field (0, -0.5, 0) */
      const REAL* r = d->site_xmat[id];
      const REAL mag[3] = {0, (REAL)-0.5, 0};
      for (int k = 0; k < 3; k++) out[k] = r[k] * mag[0] + r[3 + k] * mag[1] + r[6 + k] * mag[2];
      continue;
    }
    if (t < O_SENS_FRAMEPOS || t > O_SENS_FRAMEZAXIS) continue;
    object_frame(m, d, ot, id, &body, &pos, &rot);
    if (t == O_SENS_FRAMEPOS) { for (int k = 0; k < 3; k++) out[k] = pos[k]; }
    else if (t == O_SENS_FRAMEQUAT) {
      if (ot == O_OBJ_SITE) o_quat_mul(out, d->xquat[body], m->site_quat[id]);
      else if (ot == O_OBJ_BODY) o_quat_mul(out, d->xquat[body], m->body_iquat[id]);
      else for (int k = 0; k < 4; k++) out[k] = d->xquat[body][k];
    } else {
      int col = t - O_SENS_FRAMEXAXIS;
      for (int k = 0; k < 3; k++) out[k] = rot[3 * k + col];
    }
  }
}

static void sensor_vel(const OModel* m, OData* d) {
  for (int s = 0; s < m->nsensor; s++) {
    REAL* out = d->sensordata + m->sensor_adr[s];
    int t = m->sensor_type[s], ot = m->sensor_objtype[s], id = m->sensor_objid[s];
    REAL v[6];
    if (t == O_SENS_VELOCIMETER) { object_velocity(m, d, O_OBJ_SITE, id, v, 1); for (int k = 0; k < 3; k++) out[k] = v[3 + k]; }
    else if (t == O_SENS_GYRO) { object_velocity(m, d, O_OBJ_SITE, id, v, 1); for (int k = 0; k < 3; k++) out[k] = v[k]; }
    else if (t == O_SENS_FRAMELINVEL) { object_velocity(m, d, ot, id, v, 0); for (int k = 0; k < 3; k++) out[k] = v[3 + k]; }
    else if (t == O_SENS_FRAMEANGVEL) { object_velocity(m, d, ot, id, v, 0); for (int k = 0; k < 3; k++) out[k] = v[k]; }
  }
}

/* contact force in the contact frame: normal force = sum of the contact's efc_force rows (pyramid) */
static void contact_force_world(const OData* d, const OContact* c, REAL* fw) {
  REAL fc[3] = {0, 0, 0};
  if (c->efc_address >= 0) {
    int a = c->efc_address;
    if (c->dim == 1) fc[0] = d->efc_force[a];
    else {
      fc[0] = d->efc_force[a] + d->efc_force[a + 1] + d->efc_force[a + 2] + d->efc_force[a + 3];
      fc[1] = (d->efc_force[a] - d->efc_force[a + 1]) * c->friction[0];
      fc[2] = (d->efc_force[a + 2] - d->efc_force[a + 3]) * c->friction[1];
    }
  }
  for (int k = 0; k < 3; k++) fw[k] = c->frame[k] * fc[0] + c->frame[3 + k] * fc[1] + c->frame[6 + k] * fc[2];
}

static void rne_postconstraint(const OModel* m, OData* d) {
  int nb = m->nbody;
  for (int b = 0; b < nb; b++)
    for (int c = 0; c < 6; c++) d->cfrc_ext[b][c] = 0;
  /* xfrc_applied -> com-based [torque, force] */
  for (int b = 1; b < nb; b++) {
    const REAL* f = d->xfrc_applied[b];
    const REAL* com = d->subtree_com[m->body_rootid[b]];
    REAL arm[3] = {d->xipos[b][0] - com[0], d->xipos[b][1] - com[1], d->xipos[b][2] - com[2]}, cr[3];
    cross3(cr, arm, f);
    for (int k = 0; k < 3; k++) { d->cfrc_ext[b][k] = f[3 + k] + cr[k]; d->cfrc_ext[b][3 + k] = f[k]; }
  }
  for (int c = 0; c < d->ncon; c++) {
    const OContact* con = &d->contact[c];
    if (con->efc_address < 0) continue;
    REAL fw[3];
    contact_force_world(d, con, fw);
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    for (int side = 0; side < 2; side++) {
      int b = side ? b2 : b1;
      if (b == 0) continue;
      REAL sgn = side ? (REAL)1 : (REAL)-1;
      const REAL* com = d->subtree_com[m->body_rootid[b]];
      REAL arm[3] = {con->pos[0] - com[0], con->pos[1] - com[1], con->pos[2] - com[2]}, cr[3];
      cross3(cr, arm, fw);
      for (int k = 0; k < 3; k++) { d->cfrc_ext[b][k] += sgn * cr[k]; d->cfrc_ext[b][3 + k] += sgn * fw[k]; }
    }
  }
  rne_forward_acc(m, d, 1);
  for (int c = 0; c < 6; c++) d->cfrc_int[0][c] = 0;
  for (int b = 1; b < nb; b++) {
    REAL t1[6], t2[6], t3[6];
    mul_inert_vec(t1, d->cinert[b], d->cacc[b]);
    mul_inert_vec(t2, d->cinert[b], d->cvel[b]);
    cross_force(t3, d->cvel[b], t2);
    for (int c = 0; c < 6; c++) d->cfrc_int[b][c] = t1[c] + t3[c] - d->cfrc_ext[b][c];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0)
      for (int c = 0; c < 6; c++) d->cfrc_int[p][c] += d->cfrc_int[b][c];
  }
}

static void sensor_acc(const OModel* m, OData* d) {
  rne_postconstraint(m, d);
  for (int s = 0; s < m->nsensor; s++) {
    REAL* out = d->sensordata + m->sensor_adr[s];
    int t = m->sensor_type[s], ot = m->sensor_objtype[s], id = m->sensor_objid[s];
    REAL v[6];
    if (t == O_SENS_ACCEL) { object_acceleration(m, d, O_OBJ_SITE, id, v, 1); for (int k = 0; k < 3; k++) out[k] = v[3 + k]; }
    else if (t == O_SENS_FRAMELINACC) { object_acceleration(m, d, ot, id, v, 0); for (int k = 0; k < 3; k++) out[k] = v[3 + k]; }
    else if (t == O_SENS_FRAMEANGACC) { object_acceleration(m, d, ot, id, v, 0); for (int k = 0; k < 3; k++) out[k] = v[k]; }
    else if (t == O_SENS_FORCE || t == O_SENS_TORQUE) {
      /* interaction wrench between the site's body and its parent, in the site frame (force about site) */
      int b = m->site_bodyid[id];
      const REAL* com = d->subtree_com[m->body_rootid[b]];
      const REAL* w = d->cfrc_int[b]; /* [torque about com, force] */
      REAL dif[3] = {d->site_xpos[id][0] - com[0], d->site_xpos[id][1] - com[1], d->site_xpos[id][2] - com[2]}, cr[3], tq[3];
      cross3(cr, dif, w + 3);
      for (int k = 0; k < 3; k++) tq[k] = w[k] - cr[k];
      if (t == O_SENS_FORCE) matT_vec3(out, d->site_xmat[id], w + 3); else matT_vec3(out, d->site_xmat[id], tq);
    } else if (t == O_SENS_TOUCH) {
      /* sum of normal forces of contacts on the site's body whose point lies inside the site volume */
      int b = m->site_bodyid[id];
      REAL total = 0;
      for (int c = 0; c < d->ncon; c++) {
        const OContact* con = &d->contact[c];
        if (con->efc_address < 0) continue;
        int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
        if (b1 != b && b2 != b) continue;
        REAL rel[3] = {con->pos[0] - d->site_xpos[id][0], con->pos[1] - d->site_xpos[id][1], con->pos[2] - d->site_xpos[id][2]}, loc[3];
        matT_vec3(loc, d->site_xmat[id], rel);
        const REAL* sz = m->site_size[id];
        int inside;
        if (m->site_type[id] == 6) inside = FABS(loc[0]) <= sz[0] && FABS(loc[1]) <= sz[1] && FABS(loc[2]) <= sz[2];
        else inside = dot3(loc, loc) <= sz[0] * sz[0];
        if (!inside) continue;
        REAL nf = 0;
        int a = con->efc_address, nr = con->dim == 1 ? 1 : 4;
        for (int k = 0; k < nr; k++) nf += d->efc_force[a + k];
        total += nf;
      }
      out[0] = total;
    }
  }
}

/* ---------------------------------------------------------------------------------------- top level */

void o_make_data(const OModel* m, OData* d) {
  memset(d, 0, sizeof(*d));
  for (int i = 0; i < m->nq; i++) d->qpos[i] = m->qpos0[i];
}

void o_kinematics_only(const OModel* m, OData* d) {
  kinematics(m, d);
  com_pos(m, d);
  crb(m, d);
}

void o_forward(const OModel* m, OData* d) {
  kinematics(m, d);
  com_pos(m, d);
  crb(m, d);
  cholesky(m->nv, d->qM, d->qL);
  collision(m, d);
  sensor_pos(m, d);
  com_vel(m, d);
  make_constraint(m, d); /* needs qvel only through J*qvel; placed after com_vel for clarity */
  passive(m, d);
  rne(m, d);
  sensor_vel(m, d);
  actuation(m, d);
  acceleration(m, d);
  solve(m, d);
  sensor_acc(m, d);
}

static void advance(const OModel* m, OData* d, const REAL* qacc) {
  REAL dt = m->timestep;
  for (int i = 0; i < m->nv; i++) d->qvel[i] += qacc[i] * dt;
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == O_JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += dt * d->qvel[da + k];
      REAL w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      REAL n = SQRT(dot3(w, w));
      REAL axis[3];
      REAL inv = (REAL)1 / (n + (n == 0 ? O_MINVAL : 0));
      for (int k = 0; k < 3; k++) axis[k] = w[k] * inv;
      REAL qr[4], qn[4];
      axis_angle_quat(qr, axis, dt * n);
      o_quat_mul(qn, d->qpos + qa + 3, qr);
      quat_normalize(qn);
      for (int k = 0; k < 4; k++) d->qpos[qa + 3 + k] = qn[k];
    } else {
      d->qpos[qa] += dt * d->qvel[da];
    }
  }
  d->time += dt;
}

/* implicitfast: (M - dt*dqfrc_smooth/dqvel) qacc = qfrc_smooth + qfrc_constraint; for this feature subset
 * (motors without velocity-dependent gain, joint dampers, RNE derivative dropped) the derivative is
 * -diag(dof_damping). */
void o_step(const OModel* m, OData* d) {
  int nv = m->nv;
  o_forward(m, d);
  static _Thread_local REAL A[O_MAXNV][O_MAXNV], L[O_MAXNV][O_MAXNV];
  REAL rhs[O_MAXNV], qacc[O_MAXNV];
  for (int i = 0; i < nv; i++) {
    for (int j = 0; j < nv; j++) A[i][j] = d->qM[i][j];
    A[i][i] += m->timestep * m->dof_damping[i];
    rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  }
  cholesky(nv, A, L);
  chol_solve(nv, L, qacc, rhs);
  advance(m, d, qacc);
}
