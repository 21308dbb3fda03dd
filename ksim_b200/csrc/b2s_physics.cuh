// b200sim: warp-cooperative rigid-body dynamics for one environment (one warp).
//
// Replaces `mjx.forward` / `mjx.step` at ksim/engine.py:213,257 for the feature subset of SURVEY.md 9.1:
// free + hinge joints, capsule/sphere/plane contacts from explicit or dynamic pairs, joint limits, dof
// friction loss, motors, Newton solver with pyramidal cones, implicitfast integration (Euler damping off).
// Stage by stage it computes what the CPU oracle (oracle/physics.c) computes, in fp32, with lanes mapped to
// bodies / dofs / constraint rows.
#pragma once

#include "b2s_work.cuh"

namespace b2s {

#if B2S_DEVICE
#define LANE_FOR_ALL(i, n) for (int i = lane; i < (((n) + 31) & ~31); i += 32)
DEV int wcompact(bool pred, int& base, LANE_ARG) {
  unsigned mask = __ballot_sync(0xffffffffu, pred);
  int idx = base + __popc(mask & ((1u << lane) - 1u));
  base += __popc(mask);
  return idx;
}
#else
#define LANE_FOR_ALL(i, n) for (int i = 0; i < (n); i++)
DEV int wcompact(bool pred, int& base, LANE_ARG) {
  int idx = base;
  if (pred) base++;
  return idx;
}
#endif

#define LD3(arr, i) v3((arr)[0][i], (arr)[1][i], (arr)[2][i])
#define ST3(arr, i, v) do { (arr)[0][i] = (v).x; (arr)[1][i] = (v).y; (arr)[2][i] = (v).z; } while (0)
#define LDQ(arr, i) Q4{(arr)[0][i], (arr)[1][i], (arr)[2][i], (arr)[3][i]}
#define STQ(arr, i, q) do { (arr)[0][i] = (q).w; (arr)[1][i] = (q).x; (arr)[2][i] = (q).y; (arr)[3][i] = (q).z; } while (0)

template <class A> DEV M9 ldm9(const A& arr, int i) {
  M9 r;
#pragma unroll
  for (int k = 0; k < 9; k++) r.m[k] = arr[k][i];
  return r;
}
template <class A> DEV void stm9(A& arr, int i, const M9& m) {
#pragma unroll
  for (int k = 0; k < 9; k++) arr[k][i] = m.m[k];
}
template <class A> DEV V6 ld6(const A& arr, int i) {
  V6 r;
#pragma unroll
  for (int k = 0; k < 6; k++) r.v[k] = arr[k][i];
  return r;
}
template <class A> DEV void st6(A& arr, int i, const V6& v) {
#pragma unroll
  for (int k = 0; k < 6; k++) arr[k][i] = v.v[k];
}
DEV V3 ldg3(const float* p) { return v3(p[0], p[1], p[2]); }
DEV Q4 ldgq(const float* p) { Q4 q; q.w = p[0]; q.x = p[1]; q.y = p[2]; q.z = p[3]; return q; }

// ------------------------------------------------------------------------------------------ position stage

// Rotates v by the unit quaternion q without forming the matrix: v + 2 w (u x v) + 2 u x (u x v), u = q.xyz
DEV V3 qrotv(Q4 q, V3 v) {
  const V3 u = v3(q.x, q.y, q.z);
  V3 t = cross(u, v);
  t = t * 2.0f;
  return v + t * q.w + cross(u, t);
}

// Forward kinematics in three passes so that only a short compose step sits on the tree's serial path:
//   A  (lanes = joints, then bodies)  half-angle sin/cos of every hinge; each body's transform RELATIVE to its parent
//      (body frame followed by its joints' rotations), its joints' anchors / axes in the parent frame;
//   B  (level by level)               x = parent o local, normalised: one quaternion product + one rotation per level;
//   C  (lanes = bodies / joints)      rotation matrices, inertial frames, world anchors / axes, sites.
// Same composition as mj_kinematics (oracle/physics.c), associated differently.
template <class D>
DEV void kinematics(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY), nj = DIM(NJ, MH_NJNT);
  const int* parentid = MI(m, MH_BODY_PARENTID);
  const int* jntnum = MI(m, MH_BODY_JNTNUM);
  const int* jntadr = MI(m, MH_BODY_JNTADR);
  const int* jtype = MI(m, MH_JNT_TYPE);
  const int* jbody = MI(m, MH_JNT_BODYID);
  const int* qposadr = MI(m, MH_JNT_QPOSADR);
  const int* lstart = MI(m, MH_LEVEL_START);
  const int* lbody = MI(m, MH_LEVEL_BODY);
  const float* body_pos = MF(m, MH_BODY_POS);
  const float* body_quat = MF(m, MH_BODY_QUAT);
  const float* body_iquat = MF(m, MH_BODY_IQUAT);
  const float* jnt_pos = MF(m, MH_JNT_POS);
  const float* jnt_axis = MF(m, MH_JNT_AXIS);
  const float* qpos0 = MF(m, MH_QPOS0);
  SUB_BEGIN();
  // ---- A: hinge half-angle sin / cos (scratch: the solver vectors are free here)
  float* hs = w.s_qacc;
  float* hc = w.s_Ma;
  LANE_FOR(jn, nj) {
    if (jtype[jn] == JNT_HINGE) {
      const int qa = qposadr[jn];
      const float half = (w.qpos[qa] - qpos0[qa]) * 0.5f;
      float sn, cs;
#if B2S_DEVICE
      sincosf(half, &sn, &cs);  // one shared range reduction
#else
      sn = sinf(half); cs = cosf(half);
#endif
      hs[jn] = sn; hc[jn] = cs;
    }
  }
  WSYNC();
  SUB_MARK(w, 24);
  // local transforms and local anchors / axes (into xanchor / xaxis); lane = body
#if B2S_DEVICE
  {
    constexpr unsigned FULL = 0xffffffffu;
    const int b = lane < nb ? lane : 0;
    const int p = parentid[b];
    V3 pos = v3(0, 0, 0);
    Q4 quat; quat.w = 1; quat.x = quat.y = quat.z = 0;
    if (lane < nb && b > 0) {
      const int jn = jntnum[b], ja = jntadr[b];
      if (jn == 1 && jtype[ja] == JNT_FREE) {
        const int qa = qposadr[ja];
        Q4 q; q.w = w.qpos[qa + 3]; q.x = w.qpos[qa + 4]; q.y = w.qpos[qa + 5]; q.z = w.qpos[qa + 6];
        q = qnormalize(q);
        w.qpos[qa + 3] = q.w; w.qpos[qa + 4] = q.x; w.qpos[qa + 5] = q.y; w.qpos[qa + 6] = q.z;
        pos = v3(w.qpos[qa], w.qpos[qa + 1], w.qpos[qa + 2]);
        quat = q;
        ST3(w.xanchor, ja, pos);
        const V3 ax0 = ldg3(jnt_axis + 3 * ja);
        ST3(w.xaxis, ja, ax0);
      } else {
        pos = ldg3(body_pos + 3 * b);
        quat = ldgq(body_quat + 4 * b);
        for (int j = ja; j < ja + jn; j++) {
          const V3 jp = ldg3(jnt_pos + 3 * j), jax = ldg3(jnt_axis + 3 * j);
          const V3 anchor = qrotv(quat, jp) + pos;
          ST3(w.xanchor, j, anchor);
          const V3 axis = qrotv(quat, jax);
          ST3(w.xaxis, j, axis);
          Q4 qloc; qloc.w = hc[j]; qloc.x = jax.x * hs[j]; qloc.y = jax.y * hs[j]; qloc.z = jax.z * hs[j];
          quat = qmul(quat, qloc);
          pos = anchor - qrotv(quat, jp);
        }
      }
    }
    SUB_MARK(w, 25);
    // ---- B: compose down the tree in registers.  Every pass rebuilds x = parent o local from the parent's current
    // frame (fetched by shuffle); after pass k all bodies of depth <= k + 1 are final, so nlevel - 1 passes suffice and
    // no per-level body lists or shared-memory round trips sit on the serial path.
    const V3 lpos = pos;
    const Q4 lquat = quat;
    quat = qnormalize(quat);
    const int nlevel = m.h[MH_NLEVEL];
    for (int l = 1; l < nlevel; l++) {
      Q4 pq; V3 pp;
      pq.w = __shfl_sync(FULL, quat.w, p); pq.x = __shfl_sync(FULL, quat.x, p);
      pq.y = __shfl_sync(FULL, quat.y, p); pq.z = __shfl_sync(FULL, quat.z, p);
      pp.x = __shfl_sync(FULL, pos.x, p); pp.y = __shfl_sync(FULL, pos.y, p); pp.z = __shfl_sync(FULL, pos.z, p);
      if (p > 0) {
        pos = pp + qrotv(pq, lpos);
        quat = qnormalize(qmul(pq, lquat));
      }
    }
    if (lane < nb) {
      ST3(w.xpos, b, pos);
      STQ(w.xquat, b, quat);
    }
    __syncwarp();
  }
#else
  LANE_FOR(b, nb) {
    V3 pos = v3(0, 0, 0);
    Q4 quat; quat.w = 1; quat.x = quat.y = quat.z = 0;
    if (b > 0) {
      const int jn = jntnum[b], ja = jntadr[b];
      if (jn == 1 && jtype[ja] == JNT_FREE) {
        const int qa = qposadr[ja];
        Q4 q; q.w = w.qpos[qa + 3]; q.x = w.qpos[qa + 4]; q.y = w.qpos[qa + 5]; q.z = w.qpos[qa + 6];
        q = qnormalize(q);
        w.qpos[qa + 3] = q.w; w.qpos[qa + 4] = q.x; w.qpos[qa + 5] = q.y; w.qpos[qa + 6] = q.z;
        pos = v3(w.qpos[qa], w.qpos[qa + 1], w.qpos[qa + 2]);
        quat = q;
        ST3(w.xanchor, ja, pos);
        const V3 ax0 = ldg3(jnt_axis + 3 * ja);
        ST3(w.xaxis, ja, ax0);
      } else {
        pos = ldg3(body_pos + 3 * b);
        quat = ldgq(body_quat + 4 * b);
        for (int j = ja; j < ja + jn; j++) {
          const V3 jp = ldg3(jnt_pos + 3 * j), jax = ldg3(jnt_axis + 3 * j);
          const V3 anchor = qrotv(quat, jp) + pos;
          ST3(w.xanchor, j, anchor);
          const V3 axis = qrotv(quat, jax);
          ST3(w.xaxis, j, axis);
          Q4 qloc; qloc.w = hc[j]; qloc.x = jax.x * hs[j]; qloc.y = jax.y * hs[j]; qloc.z = jax.z * hs[j];
          quat = qmul(quat, qloc);
          pos = anchor - qrotv(quat, jp);
        }
      }
    }
    ST3(w.xpos, b, pos);
    STQ(w.xquat, b, quat);
  }
  // ---- B: compose down the tree.  Level 1 (children of the world) is already absolute up to normalisation.
  const int nlevel = m.h[MH_NLEVEL];
  for (int l = 0; l < nlevel; l++) {
    const int s = lstart[l], e = lstart[l + 1];
    LANE_FOR(t, e - s) {
      const int b = lbody[s + t];
      const int p = parentid[b];
      Q4 quat = LDQ(w.xquat, b);
      V3 pos = LD3(w.xpos, b);
      if (p > 0) {
        const Q4 pq = LDQ(w.xquat, p);
        pos = LD3(w.xpos, p) + qrotv(pq, pos);
        quat = qmul(pq, quat);
      }
      quat = qnormalize(quat);
      ST3(w.xpos, b, pos);
      STQ(w.xquat, b, quat);
    }
    WSYNC();
  }
#endif
  SUB_MARK(w, 26);
  // ---- C: everything that only reads the finished frames
  LANE_FOR(jn, nj) {
    const int p = parentid[jbody[jn]];
    if (p > 0) {
      const Q4 pq = LDQ(w.xquat, p);
      const V3 anchor = LD3(w.xpos, p) + qrotv(pq, LD3(w.xanchor, jn)), axis = qrotv(pq, LD3(w.xaxis, jn));
      ST3(w.xanchor, jn, anchor);  // (ST3 evaluates its argument per component: never pass it an in-place expression)
      ST3(w.xaxis, jn, axis);
    }
  }
  LANE_FOR(b, nb) {
    const Q4 quat = LDQ(w.xquat, b);
    const M9 xm = q2mat(quat);
    stm9(w.xmat, b, xm);
    const V3 ip = LD3(w.xpos, b) + mulv(xm, LD3(w.body_ipos, b));
    ST3(w.xipos, b, ip);
    stm9(w.ximat, b, q2mat(qmul(quat, ldgq(body_iquat + 4 * b))));
  }
  WSYNC();
  const int ns = DIM(NS, MH_NSITE);
  const int* site_body = MI(m, MH_SITE_BODYID);
  const float* site_pos = MF(m, MH_SITE_POS);
  const float* site_quat = MF(m, MH_SITE_QUAT);
  LANE_FOR(s, ns) {
    const int b = site_body[s];
    V3 sp; Q4 sq;
    if constexpr (D::SMALL) { sp = ldg3(site_pos + 3 * s); sq = ldgq(site_quat + 4 * s); }
    else { sp = LD3(w.site_pos, s); sq = LDQ(w.site_quat, s); }  // per-env copies (IMUAlignmentRandomizer)
    V3 p = LD3(w.xpos, b) + mulv(ldm9(w.xmat, b), sp);
    ST3(w.site_xpos, s, p);
    stm9(w.site_xmat, s, q2mat(qmul(LDQ(w.xquat, b), sq)));
  }
  WSYNC();
  SUB_MARK(w, 27);
}

// Leaf-to-root accumulation of NC per-body components held as arr[c][body]: after the call arr[c][b] is the sum over
// b's subtree (children of the world excepted -- nothing reads the world's sum).  The schedule (MH_UP_START / MH_UP_BODY)
// goes up the tree one level and one sibling rank at a time, so a pass touches every parent at most once: ~NC * (bodies
// in the pass) independent adds per pass instead of one serial loop over each subtree.
template <int NC, class A>
DEV void subtree_accumulate(const Mdl& m, A& arr, LANE_ARG) {
  const int npass = m.h[MH_NUPPASS];
#if B2S_DEVICE
  // lane = body, the NC components in registers; in pass p a parent pulls its (at most one) child of that pass by
  // shuffle (MH_UP_CHILD): no shared-memory round trip or barrier between the passes
  {
    const int nb = m.h[MH_NBODY];
    const int* uchild = MI(m, MH_UP_CHILD);
    const int b = lane < nb ? lane : 0;
    float v[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) v[c] = arr[c][b];
    for (int p = 0; p < npass; p++) {
      const int ch = lane < nb ? uchild[p * nb + b] : -1;
      const int src = ch >= 0 ? ch : 0;
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const float x = __shfl_sync(0xffffffffu, v[c], src);
        v[c] += ch >= 0 ? x : 0.0f;
      }
    }
    if (lane < nb) {
#pragma unroll
      for (int c = 0; c < NC; c++) arr[c][b] = v[c];
    }
    __syncwarp();
    return;
  }
#endif
  const int* ustart = MI(m, MH_UP_START);
  const int* ubody = MI(m, MH_UP_BODY);
  const int* parentid = MI(m, MH_BODY_PARENTID);
  for (int p = 0; p < npass; p++) {
    const int s = ustart[p], cnt = ustart[p + 1] - s;
    LANE_FOR(idx, NC * cnt) {
      const int t = idx / NC, c = idx - t * NC;
      const int b = ubody[s + t];
      arr[c][parentid[b]] += arr[c][b];
    }
    WSYNC();
  }
}

template <class D>
DEV void com_pos(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY), nj = DIM(NJ, MH_NJNT);
  const int* rootid = MI(m, MH_BODY_ROOTID);
  const int* sub_end = MI(m, MH_BODY_SUBTREE_END);
  const float* subtreemass = MF(m, MH_BODY_SUBTREEMASS);
  // mass-weighted positions into cfrc[0..2] (free at this point of the pipeline)
  LANE_FOR(b, nb) {
    const float ms = w.body_mass[b];
    w.cfrc[0][b] = ms * w.xipos[0][b]; w.cfrc[1][b] = ms * w.xipos[1][b]; w.cfrc[2][b] = ms * w.xipos[2][b];
  }
  WSYNC();
  subtree_accumulate<3>(m, w.cfrc, LANE_PASS);
  LANE_FOR(b, nb) {
    // nominal body_subtreemass even when body_mass is randomised (ksim/randomization.py:351-355)
    const float sm = fmaxf(subtreemass[b], MINVAL);
    w.subtree_com[0][b] = w.cfrc[0][b] / sm; w.subtree_com[1][b] = w.cfrc[1][b] / sm; w.subtree_com[2][b] = w.cfrc[2][b] / sm;
  }
  WSYNC();
  LANE_FOR(b, nb) {
    if (b == 0) {
#pragma unroll
      for (int k = 0; k < 10; k++) w.cinert[k][0] = 0.0f;
    } else {
      V3 com = LD3(w.subtree_com, rootid[b]);
      V3 off = LD3(w.xipos, b) - com;
      const float mass = w.body_mass[b];
      M9 R = ldm9(w.ximat, b);
      const float I0 = w.body_inertia[0][b], I1 = w.body_inertia[1][b], I2 = w.body_inertia[2][b];
      float full[9];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          full[3 * r + c] = R.m[3 * r] * I0 * R.m[3 * c] + R.m[3 * r + 1] * I1 * R.m[3 * c + 1] + R.m[3 * r + 2] * I2 * R.m[3 * c + 2];
      const float dd = dot(off, off);
      const float o[3] = {off.x, off.y, off.z};
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) full[3 * r + c] += mass * ((r == c ? dd : 0.0f) - o[r] * o[c]);
      w.cinert[0][b] = full[0]; w.cinert[1][b] = full[4]; w.cinert[2][b] = full[8];
      w.cinert[3][b] = full[1]; w.cinert[4][b] = full[2]; w.cinert[5][b] = full[5];
      w.cinert[6][b] = mass * off.x; w.cinert[7][b] = mass * off.y; w.cinert[8][b] = mass * off.z;
      w.cinert[9][b] = mass;
    }
  }
  const int* jbody = MI(m, MH_JNT_BODYID);
  const int* jdof = MI(m, MH_JNT_DOFADR);
  const int* jtype = MI(m, MH_JNT_TYPE);
  LANE_FOR(j, nj) {
    const int b = jbody[j], da = jdof[j];
    V3 off = LD3(w.subtree_com, rootid[b]) - LD3(w.xanchor, j);
    if (jtype[j] == JNT_FREE) {
      M9 xm = ldm9(w.xmat, b);
#pragma unroll
      for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int c = 0; c < 6; c++) w.cdof[c][da + k] = (c == 3 + k) ? 1.0f : 0.0f;
        V3 ax = v3(xm.m[k], xm.m[3 + k], xm.m[6 + k]);
        V3 cr = cross(ax, off);
        w.cdof[0][da + 3 + k] = ax.x; w.cdof[1][da + 3 + k] = ax.y; w.cdof[2][da + 3 + k] = ax.z;
        w.cdof[3][da + 3 + k] = cr.x; w.cdof[4][da + 3 + k] = cr.y; w.cdof[5][da + 3 + k] = cr.z;
      }
    } else {
      V3 ax = LD3(w.xaxis, j);
      V3 cr = cross(ax, off);
      w.cdof[0][da] = ax.x; w.cdof[1][da] = ax.y; w.cdof[2][da] = ax.z;
      w.cdof[3][da] = cr.x; w.cdof[4][da] = cr.y; w.cdof[5][da] = cr.z;
    }
  }
  WSYNC();
}

template <class D>
DEV void crb_mass_matrix(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY), nv = DIM(NV, MH_NV);
  const int* sub_end = MI(m, MH_BODY_SUBTREE_END);
  const int* dof_body = MI(m, MH_DOF_BODYID);
  const int* dof_parent = MI(m, MH_DOF_PARENTID);
  // composite inertia: sum of cinert over the (contiguous, depth-first) subtree of each body
  LANE_FOR(idx, 10 * nb) {
    const int b = idx / 10, k = idx - b * 10;
    w.crb[k][b] = w.cinert[k][b];
  }
  WSYNC();
  subtree_accumulate<10>(m, w.crb, LANE_PASS);
  LANE_FOR(i, nv) {
    float ci[10];
    const int b = dof_body[i];
#pragma unroll
    for (int k = 0; k < 10; k++) ci[k] = w.crb[k][b];
    V6 buf = mul_inert_vec(ci, ld6(w.cdof, i));
    // entries outside the (static) ancestor pattern were zeroed once by apply_overrides and are never written
    for (int j = i; j >= 0; j = dof_parent[j]) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < 6; k++) s += w.cdof[k][j] * buf.v[k];
      if (j == i) s += w.armature[i];
      w.M[i][j] = s;
      w.M[j][i] = s;
    }
  }
  WSYNC();
}

#if !B2S_DEVICE
// in-place dense Cholesky of the lower triangle held in W, one phase per column.  The diagonal of W keeps
// A[j][j] (other lanes read it during the phase); the pivots live in Linv_diag.
template <class D>
DEVNI void chol_inplace(Work<D>& w, int nv, LANE_ARG) {
  for (int j = 0; j < nv; j++) {
    LANE_FOR(i, nv) {
      if (i >= j) {
        float t = w.W[i][j], s = w.W[j][j];
        for (int k = 0; k < j; k++) {
          const float ljk = w.W[j][k];
          t -= w.W[i][k] * ljk;
          s -= ljk * ljk;
        }
        if (s < MINVAL) s = MINVAL;
        const float dj = sqrtf(s);
        if (i == j) w.Linv_diag[j] = 1.0f / dj;
        else w.W[i][j] = t / dj;
      }
    }
    WSYNC();
  }
}

// solves (L L^T) x = b with the factor in W; x and b may alias.  Uses s_tmp.
template <class D>
DEVNI void chol_solve(Work<D>& w, float* x, const float* b, int nv, LANE_ARG) {
  float* y = w.s_tmp;
  LANE_FOR(i, nv) y[i] = b[i];
  WSYNC();
  for (int k = 0; k < nv; k++) {
    const float yk = y[k] * w.Linv_diag[k];
    LANE_FOR(i, nv) if (i > k) y[i] -= w.W[i][k] * yk;
    WSYNC();
  }
  LANE_FOR(i, nv) y[i] *= w.Linv_diag[i];
  WSYNC();
  for (int k = nv - 1; k >= 0; k--) {
    const float xk = y[k] * w.Linv_diag[k];
    LANE_FOR(i, nv) if (i < k) y[i] -= w.W[k][i] * xk;
    WSYNC();
  }
  LANE_FOR(i, nv) x[i] = y[i] * w.Linv_diag[i];
  WSYNC();
}
#endif

// FS_NEWTON_DIAG: FS_NEWTON with the single-dof rows' diagonal term already summed per dof in w.s_mv (newton_dof());
// it also leaves 1/diag(L) behind (device: w.s_tmp, host: Linv_diag) so that fs_resolve() can reuse the factor.
enum { FS_MASS = 0, FS_NEWTON = 1, FS_IMPLICIT = 2, FS_NEWTON_DIAG = 3 };

// bit 8 of sync_mask switches the constraint-space Newton path off (tuning / A-B tests)
DEV bool dual_enabled(const Mdl& m) { return !((m.sync_mask >> 8) & 1); }
// bit 10 switches the exact one-pass line search off (the bracketing iteration is used everywhere)
DEV bool exact_ls_enabled(const Mdl& m) { return !((m.sync_mask >> 10) & 1); }
// true: the constraint-space (dual) Newton path handles this sub-step's rows, and acceleration() must provide
// Z = M^-1 J^T.  The register-resident solve (ne <= 8, no friction rows) always wins; the shared-memory dual path is
// only kept for A/B tests (bit 9) and for the vote-synchronised mode (bit 7), which newton_dof() does not take part in.
template <class D>
DEV bool dual_path(const Mdl& m, const Work<D>& w, int ne) {
  if (!(ne <= D::NDUAL && ((m.sync_mask >> 8) & 1) == 0)) return false;
  if (D::NEFC <= D::NDUAL) return true;  // every row count of this model fits (humanoid): newton_dof is compiled out
#if B2S_DEVICE
  if (ne <= 8 && w.scratch_i[0] == 0 && ((m.sync_mask >> 10) & 1) == 0) return true;
#endif
  return ((m.sync_mask >> 9) & 1) || ((m.sync_mask >> 7) & 1);
}
// bit 9 switches the factor-reusing dof-space Newton path (newton_dof) off: the reference-shaped solve_primal runs instead
DEV bool newton_dof_enabled(const Mdl& m) { return !((m.sync_mask >> 9) & 1); }
// bit 11: newton_dof() runs the REFERENCE-SHAPED iteration (B2S_SOLVER_REFERENCE on models whose rows exceed NDUAL): the
// reference's bracketing line search, its iteration cap and its stopping rules, no early finish -- the same sequence of
// iterates as solve_primal up to fp32 rounding, with newton_dof's bookkeeping (factor reuse while the active set stands,
// delta form, per-dof scatters)
DEV bool newton_dof_refshape(const Mdl& m) { return (m.sync_mask >> 11) & 1; }

// Builds A (FS_MASS: M; FS_NEWTON: M + J^T diag(D*active) J; FS_IMPLICIT: M + dt*diag(damping)), factors it and
// solves A x = b.  The three dense factorisations of a physics sub-step all go through here.
//
// Device path: lane i keeps row i of the lower triangle in REGISTERS; column j of the factor needs L[j][k] from
// lane j, which arrives by warp shuffle, so the whole factorisation and the forward substitution run without
// shared memory or barriers.  Only the backward substitution needs columns of L: the rows are spilled to W once
// and read back transposed (conflict-free: lanes read consecutive columns of one row).
//
// FS_MASS with nz > 0 additionally solves M Z_s = J_s^T for the first nz constraint rows (the dual Newton path's
// Z = M^-1 J^T), two right-hand sides at a time so their dependent shuffle chains overlap.
template <class D>
DEV float jac_entry(const Work<D>& w, int s, int nsparse, int i) {
  return s < nsparse ? (w.efc_dof[s] == i ? w.efc_sign[s] : 0.0f) : w.Jd[s - nsparse][i];
}

template <class D>
DEVNI void factor_solve(const Mdl& m, Work<D>& w, int mode, float* x, const float* b, int nz, LANE_ARG) {
  const int nv = DIM(NV, MH_NV);
#if B2S_DEVICE && !defined(B2S_COMPACT_CHOL)
  constexpr int N = D::NV;
  constexpr unsigned FULL = 0xffffffffu;
  const int i = lane;
  const bool live = i < nv;
  // Rows / columns >= nv (and the duplicate row idle lanes load) only ever feed entries that are never read
  // back: column j takes L[j][k] from lane j < nv, and every store / substitution below is guarded.
  int ri = i < N ? i : N - 1;
  // ptxas 12.9 miscompiles the NV = 32 instantiation of this function: it hoists the row address &w.M[lane] into a
  // register the callers are supposed to set, one call site passes something else in it, and the 32 row loads below were
  // merged into 64 / 128-bit ones although a row of the odd stride NVP is only 4-byte aligned ("misaligned address", found
  // by the generic-instantiation GPU test).  An opaque row index keeps the address arithmetic inside the function and
  // volatile keeps the loads scalar; the measured instantiations (NV = 23, 26) are left as they were.
  if constexpr (N % 4 == 0) asm volatile("mov.b32 %0, %0;" : "+r"(ri));
  float r[N];
  float dg = 0.0f;
  DBG_STOP(m, 40);
  if (mode == FS_NEWTON) {
    const int nsparse = w.nsparse;
    if (live)
      for (int rr = 0; rr < nsparse; rr++)
        if (w.efc_active[rr] && w.efc_dof[rr] == i) dg += w.efc_D[rr];
  } else if (mode == FS_NEWTON_DIAG) {
    if (live) dg = w.s_mv[i];
  } else if (mode == FS_IMPLICIT) {
    if (live) dg = m.mf[MF_TIMESTEP] * w.damping[i];
  }
  // The diagonal term goes through shared memory: r[i] += dg would be a dynamic register index (the whole row would
  // move to local memory), and adding it inside the column loop costs three instructions per column.  Lane i bumps
  // M[i][i], every lane loads its row, lane i puts the old value back.
  const float diag0 = live ? w.M[ri][ri] : 0.0f;
  if (mode != FS_MASS) {
    if (live) w.M[ri][ri] = diag0 + dg;
    __syncwarp();
  }
  if constexpr (N % 4 == 0) {
    const volatile float* row = w.M[ri];
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = row[j];
  } else {
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = w.M[ri][j];
  }
  DBG_STOP(m, 41);
  if (mode != FS_MASS) {
    __syncwarp();
    if (live) w.M[ri][ri] = diag0;
  }
  if (mode == FS_NEWTON || mode == FS_NEWTON_DIAG) {
    // dense constraint rows: J^T D J, accumulated on the register rows (after the load)
    const int ne = w.nefc, nsparse = w.nsparse;
    for (int rr = nsparse; rr < ne; rr++) {
      if (!w.efc_active[rr]) continue;
      const float* J = w.Jd[rr - nsparse];
      const float c = live ? w.efc_D[rr] * J[i] : 0.0f;
#pragma unroll
      for (int j = 0; j < N; j++) r[j] = fmaf(c, J[j], r[j]);
    }
  }
  float invd = 1.0f;
#pragma unroll
  for (int j = 0; j < N; j++) {
    float t = r[j];
#pragma unroll
    for (int k = 0; k < j; k++) t = fmaf(-r[k], __shfl_sync(FULL, r[k], j), t);
    float sp = __shfl_sync(FULL, t, j);
    sp = fmaxf(sp, MINVAL);
    // rsqrt + one Newton refinement: no slow-path call (a call here would spill the whole row to local memory)
    float idj = rsqrtf(sp);
    idj = idj * fmaf(-0.5f * sp, idj * idj, 1.5f);
    r[j] = t * idj;  // lane j: sqrt(pivot); lanes below: L[i][j]
    if (i == j) invd = idj;
  }
  DBG_STOP(m, 42);
  float y = live ? b[i] : 0.0f;
  if (mode == FS_NEWTON_DIAG && live) w.s_tmp[i] = invd;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const float yk = __shfl_sync(FULL, y * invd, k);
    if (i > k) y = fmaf(-r[k], yk, y);
  }
  y *= invd;
  DBG_STOP(m, 43);
#pragma unroll
  for (int j = 0; j < N; j++) if (live && j < i) w.W[i][j] = r[j];
  __syncwarp();
  DBG_STOP(m, 44);
  float xv = y;
#pragma unroll
  for (int k = N - 1; k >= 0; k--) {
    const float xk = __shfl_sync(FULL, xv * invd, k);
    if (i < k && k < nv) xv = fmaf(-w.W[k][i], xk, xv);
  }
  DBG_STOP(m, 45);
  if (live) x[i] = xv * invd;
  DBG_STOP(m, 46);
  if (nz > 0) {
    const int nsparse = w.nsparse;
    for (int s0 = 0; s0 < nz; s0 += 2) {
      const bool two = s0 + 1 < nz;
      float y0 = live ? jac_entry(w, s0, nsparse, i) : 0.0f;
      float y1 = (live && two) ? jac_entry(w, s0 + 1, nsparse, i) : 0.0f;
#pragma unroll
      for (int k = 0; k < N; k++) {
        const float a0 = __shfl_sync(FULL, y0 * invd, k), a1 = __shfl_sync(FULL, y1 * invd, k);
        if (i > k) { y0 = fmaf(-r[k], a0, y0); y1 = fmaf(-r[k], a1, y1); }
      }
      y0 *= invd; y1 *= invd;
#pragma unroll
      for (int k = N - 1; k >= 0; k--) {
        const float wk = (i < k && k < nv) ? w.W[k][i] : 0.0f;
        const float a0 = __shfl_sync(FULL, y0 * invd, k), a1 = __shfl_sync(FULL, y1 * invd, k);
        y0 = fmaf(-wk, a0, y0); y1 = fmaf(-wk, a1, y1);
      }
      if (live) { w.Z[s0][i] = y0 * invd; if (two) w.Z[s0 + 1][i] = y1 * invd; }
    }
  }
  __syncwarp();
#else
  LANE_FOR(i, nv) {
    for (int j = 0; j <= i; j++) w.W[i][j] = w.M[i][j];
    float dg = 0.0f;
    if (mode == FS_NEWTON || mode == FS_NEWTON_DIAG) {
      const int ne = w.nefc, nsparse = w.nsparse;
      for (int j = 0; j <= i; j++)
        for (int rr = nsparse; rr < ne; rr++)
          if (w.efc_active[rr]) w.W[i][j] += w.efc_D[rr] * w.Jd[rr - nsparse][i] * w.Jd[rr - nsparse][j];
      if (mode == FS_NEWTON_DIAG) dg = w.s_mv[i];
      else for (int rr = 0; rr < nsparse; rr++) if (w.efc_active[rr] && w.efc_dof[rr] == i) dg += w.efc_D[rr];
    } else if (mode == FS_IMPLICIT) {
      dg = m.mf[MF_TIMESTEP] * w.damping[i];
    }
    w.W[i][i] += dg;
  }
  WSYNC();
  chol_inplace(w, nv, LANE_PASS);
  chol_solve(w, x, b, nv, LANE_PASS);
  for (int s = 0; s < nz; s++) {
    LANE_FOR(i, nv) w.Z[s][i] = jac_entry(w, s, w.nsparse, i);
    WSYNC();
    chol_solve(w, w.Z[s], w.Z[s], nv, LANE_PASS);
  }
#endif
}

// Solves A x = b again with the factor the last factor_solve(FS_NEWTON_DIAG) left behind (strict lower triangle in W,
// 1/diag in s_tmp / Linv_diag): substitutions only.  x and b may alias; b must not be s_tmp.
template <class D>
DEVNI void fs_resolve(Work<D>& w, float* x, const float* b, int nv, LANE_ARG) {
#if B2S_DEVICE && !defined(B2S_COMPACT_CHOL)
  constexpr int N = D::NV;
  constexpr unsigned FULL = 0xffffffffu;
  const int i = lane;
  const bool live = i < nv;
  int ri = i < N ? i : N - 1;
  if constexpr (N % 4 == 0) asm volatile("mov.b32 %0, %0;" : "+r"(ri));  // same precaution as in factor_solve
  float r[N];
#pragma unroll
  for (int j = 0; j < N; j++) r[j] = j < ri ? w.W[ri][j] : 0.0f;
  const float invd = live ? w.s_tmp[i] : 1.0f;
  float y = live ? b[i] : 0.0f;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const float yk = __shfl_sync(FULL, y * invd, k);
    if (i > k) y = fmaf(-r[k], yk, y);
  }
  y *= invd;
  float xv = y;
#pragma unroll
  for (int k = N - 1; k >= 0; k--) {
    const float xk = __shfl_sync(FULL, xv * invd, k);
    if (i < k && k < nv) xv = fmaf(-w.W[k][i], xk, xv);
  }
  if (live) x[i] = xv * invd;
  __syncwarp();
#else
  chol_solve(w, x, b, nv, LANE_PASS);
#endif
}

// ---------------------------------------------------------------------------------------------- collision

DEV void make_frame(V3 n, float* frame) {
  float na = sqrtf(dot(n, n));
  V3 a = n * (1.0f / na);
  V3 ref = (fabsf(a.y) < 0.5f) ? v3(0, 1, 0) : v3(0, 0, 1);
  float dp = dot(a, ref);
  V3 b = ref - a * dp;
  float nbm = sqrtf(dot(b, b));
  b = b * (1.0f / nbm);
  V3 c = cross(a, b);
  frame[0] = a.x; frame[1] = a.y; frame[2] = a.z; frame[3] = b.x; frame[4] = b.y; frame[5] = b.z;
  frame[6] = c.x; frame[7] = c.y; frame[8] = c.z;
}

template <class D>
DEV void put_contact(Work<D>& w, int c, int pair, float dist, V3 pos, const float* frame) {
  w.con_pair[c] = pair;
  w.con_dist[c] = dist;
  ST3(w.con_pos, c, pos);
#pragma unroll
  for (int k = 0; k < 9; k++) w.con_frame[k][c] = frame[k];
  w.con_efc[c] = -1;
}

template <class D>
DEV void collision(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int np = DIM(NPAIR, MH_NPAIR);
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  const int* pg2 = MI(m, MH_PAIR_GEOM2);
  const int* conadr = MI(m, MH_PAIR_CONADR);
  const int* gtype = MI(m, MH_GEOM_TYPE);
  const int* gbody = MI(m, MH_GEOM_BODYID);
  const float* gpos = MF(m, MH_GEOM_POS);
  const float* gquat = MF(m, MH_GEOM_QUAT);
  const float* gsize = MF(m, MH_GEOM_SIZE);
  LANE_FOR(p, np) {
    const int g1 = pg1[p], g2 = pg2[p];
    const int t1 = gtype[g1], t2 = gtype[g2];
    const int b1 = gbody[g1], b2 = gbody[g2];
    // geom_pos / geom_size: per-env copies where the instantiation carries them (CollisionBodyRandomizer)
    V3 gp1, gp2, gs1, gs2;
    if constexpr (D::SMALL) {
      gp1 = ldg3(gpos + 3 * g1); gp2 = ldg3(gpos + 3 * g2); gs1 = ldg3(gsize + 3 * g1); gs2 = ldg3(gsize + 3 * g2);
    } else {
      gp1 = v3(w.pair_gpos[0][0][p], w.pair_gpos[0][1][p], w.pair_gpos[0][2][p]);
      gp2 = v3(w.pair_gpos[1][0][p], w.pair_gpos[1][1][p], w.pair_gpos[1][2][p]);
      gs1 = v3(w.pair_gsize[0][0][p], w.pair_gsize[0][1][p], w.pair_gsize[0][2][p]);
      gs2 = v3(w.pair_gsize[1][0][p], w.pair_gsize[1][1][p], w.pair_gsize[1][2][p]);
    }
    V3 p1 = LD3(w.xpos, b1) + mulv(ldm9(w.xmat, b1), gp1);
    V3 p2 = LD3(w.xpos, b2) + mulv(ldm9(w.xmat, b2), gp2);
    M9 m1 = q2mat(qmul(LDQ(w.xquat, b1), ldgq(gquat + 4 * g1)));
    M9 m2 = q2mat(qmul(LDQ(w.xquat, b2), ldgq(gquat + 4 * g2)));
    const int c0 = conadr[p];
    float frame[9];
    if (t1 == GEOM_PLANE && t2 == GEOM_SPHERE) {
      V3 n = v3(m1.m[2], m1.m[5], m1.m[8]);
      const float r = gs2.x;
      const float dist = dot(p2 - p1, n) - r;
      V3 pos = p2 - n * (r + 0.5f * dist);
      make_frame(n, frame);
      put_contact(w, c0, p, dist, pos, frame);
    } else if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) {
      V3 n = v3(m1.m[2], m1.m[5], m1.m[8]);
      V3 axis = v3(m2.m[2], m2.m[5], m2.m[8]);
      const float r = gs2.x, half = gs2.y;
      make_frame(n, frame);
      // tangent aligned with the capsule axis projected on the plane, when well defined (oracle/physics.c)
      const float dpa = dot(axis, n);
      V3 bt = axis - n * dpa;
      const float nbt = sqrtf(dot(bt, bt));
      if (nbt > 0.5f) {
        bt = bt * (1.0f / nbt);
        V3 cc = cross(n, bt);
        frame[3] = bt.x; frame[4] = bt.y; frame[5] = bt.z; frame[6] = cc.x; frame[7] = cc.y; frame[8] = cc.z;
      }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const float sg = s == 0 ? -1.0f : 1.0f;
        V3 end = p2 + axis * (half * sg);
        const float dist = dot(end - p1, n) - r;
        V3 pos = end - n * (r + 0.5f * dist);
        put_contact(w, c0 + s, p, dist, pos, frame);
      }
    } else {  // sphere-sphere
      V3 dv = p2 - p1;
      const float len = sqrtf(dot(dv, dv));
      V3 n = len < MINVAL ? v3(1, 0, 0) : dv * (1.0f / len);
      const float r1 = gs1.x, r2 = gs2.x;
      const float dist = len - r1 - r2;
      V3 pos = p1 + n * (r1 + 0.5f * dist);
      make_frame(n, frame);
      put_contact(w, c0, p, dist, pos, frame);
    }
  }
  WSYNC();
}

// --------------------------------------------------------------------------------------------- constraints

struct KBI { float k, b, imp; };

DEVNI KBI kbi(const Mdl& m, const float* solref, const float* solimp, float pos) {
  float timeconst = solref[0], dampratio = solref[1];
  if (!m.h[MH_DISABLE_REFSAFE] && solref[0] > 0) {
    const float lo = 2 * m.mf[MF_TIMESTEP];
    if (timeconst < lo) timeconst = lo;
  }
  float dmin = clipf(solimp[0], MINIMP, MAXIMP), dmax = clipf(solimp[1], MINIMP, MAXIMP);
  float width = fmaxf(solimp[2], MINVAL), mid = clipf(solimp[3], MINIMP, MAXIMP), power = fmaxf(solimp[4], 1.0f);
  float kk = 1.0f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  float bb = 2.0f / (dmax * timeconst);
  if (solref[0] <= 0) kk = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) bb = -solref[1] / dmax;
  const float x = fabsf(pos) / width;
  float y;
  if (power == 2.0f) {  // MuJoCo's default solimp power: same value without the pow calls
    if (x < mid) y = (1.0f / mid) * (x * x);
    else { const float u = fmaxf(1.0f - x, 0.0f); y = 1.0f - (1.0f / (1.0f - mid)) * (u * u); }
  } else if (x < mid) y = (1.0f / powf(mid, power - 1)) * powf(x, power);
  else y = 1.0f - (1.0f / powf(1.0f - mid, power - 1)) * powf(fmaxf(1.0f - x, 0.0f), power);
  if (x > 1) y = 1;
  float im = clipf(dmin + y * (dmax - dmin), dmin, dmax);
  if (x > 1) im = dmax;
  KBI r; r.k = kk; r.b = bb; r.imp = im;
  return r;
}

template <class D>
DEV void fill_row(const Mdl& m, Work<D>& w, int r, int type, float vel, float pos, float invweight, const float* solref,
                  const float* solimp, float floss) {
  KBI q = kbi(m, solref, solimp, pos);
  float R = invweight * (1.0f - q.imp) / q.imp;
  if (R < MINVAL) R = MINVAL;
  w.efc_type[r] = type;
  w.efc_aref[r] = -q.b * vel - q.k * q.imp * pos;
  w.efc_R[r] = R;
  w.efc_D[r] = 1.0f / R;
  w.efc_floss[r] = floss;
  w.efc_force[r] = 0.0f;
}

enum { EFC_FRICTION = 1, EFC_LIMIT = 2, EFC_CONTACT = 3 };

// Row order follows the oracle / MJX make_constraint: friction, limit, contact.  Single-dof rows (friction,
// limit) are kept in compact form (dof + sign); contact rows are dense in Jd.
template <class D>
DEV void make_constraint(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), nlimit = DIM(NLIMIT, MH_NLIMIT), ncon = DIM(NCON, MH_NCON);
  const float* dof_invw = MF(m, MH_DOF_INVWEIGHT0);
  const float* dof_solref = MF(m, MH_DOF_SOLREF);
  const float* dof_solimp = MF(m, MH_DOF_SOLIMP);
  int nrow = 0;
  LANE_FOR_ALL(i, nv) {
    const bool on = i < nv && w.frictionloss[i] > 0.0f;
    const int r = wcompact(on, nrow, LANE_PASS);
    if (on) {
      w.efc_dof[r] = i; w.efc_sign[r] = 1.0f;
      fill_row(m, w, r, EFC_FRICTION, w.qvel[i], 0.0f, dof_invw[i], dof_solref + 2 * i, dof_solimp + 5 * i, w.frictionloss[i]);
    }
  }
  IF_LANE0 w.scratch_i[0] = nrow;  // number of friction rows (solve() picks its path on it)
  const int* limit_jnt = MI(m, MH_LIMIT_JNT);
  const int* jqpos = MI(m, MH_JNT_QPOSADR);
  const int* jdof = MI(m, MH_JNT_DOFADR);
  const float* jrange = MF(m, MH_JNT_RANGE);
  const float* jmargin = MF(m, MH_JNT_MARGIN);
  const float* jsolref = MF(m, MH_JNT_SOLREF);
  const float* jsolimp = MF(m, MH_JNT_SOLIMP);
  LANE_FOR_ALL(t, nlimit) {
    bool on = false;
    int j = 0, dof = 0;
    float pos = 0, sign = 1;
    if (t < nlimit) {
      j = limit_jnt[t];
      dof = jdof[j];
      const float q = w.qpos[jqpos[j]];
      const float dmin = q - jrange[2 * j], dmax = jrange[2 * j + 1] - q;
      pos = fminf(dmin, dmax) - jmargin[j];
      sign = dmin < dmax ? 1.0f : -1.0f;
      on = pos < 0.0f;
    }
    const int r = wcompact(on, nrow, LANE_PASS);
    if (on) {
      w.efc_dof[r] = dof; w.efc_sign[r] = sign;
      fill_row(m, w, r, EFC_LIMIT, sign * w.qvel[dof], pos, dof_invw[dof], jsolref + 2 * j, jsolimp + 5 * j, 0.0f);
    }
  }
  const int nsparse = nrow;
  // contacts: every lane walks the (few) candidate contacts to agree on row addresses
  const int* pdim = MI(m, MH_PAIR_DIM);
  const float* pmargin = MF(m, MH_PAIR_MARGIN);
  const float* pgap = MF(m, MH_PAIR_GAP);
  for (int c = 0; c < ncon; c++) {
    const int p = w.con_pair[c];
    const float pos = w.con_dist[c] - (pmargin[p] - pgap[p]);
    int adr = -1;
    if (pos < 0.0f) { adr = nrow; nrow += (pdim[p] == 1) ? 1 : 4; }
    IF_LANE0 w.con_efc[c] = adr;
  }
  IF_LANE0 { w.nefc = nrow; w.nsparse = nsparse; }
  WSYNC();
  if (nrow == nsparse) return;
  // dense Jacobian rows: lanes over dofs
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  const int* pg2 = MI(m, MH_PAIR_GEOM2);
  const int* gbody = MI(m, MH_GEOM_BODYID);
  const int* rootid = MI(m, MH_BODY_ROOTID);
  const int* bdofmask = MI(m, MH_BODY_DOFMASK);
  const float* pfric = MF(m, MH_PAIR_FRICTION);
  LANE_FOR(i, nv) {
    for (int c = 0; c < ncon; c++) {
      const int adr = w.con_efc[c];
      if (adr < 0) continue;
      const int p = w.con_pair[c];
      const int b1 = gbody[pg1[p]], b2 = gbody[pg2[p]];
      V3 cp = LD3(w.con_pos, c);
      V3 diff = v3(0, 0, 0);
      V3 ang = v3(w.cdof[0][i], w.cdof[1][i], w.cdof[2][i]), lin = v3(w.cdof[3][i], w.cdof[4][i], w.cdof[5][i]);
      if (b2 != 0 && ((bdofmask[b2] >> i) & 1)) diff = diff + lin + cross(ang, cp - LD3(w.subtree_com, rootid[b2]));
      if (b1 != 0 && ((bdofmask[b1] >> i) & 1)) diff = diff - (lin + cross(ang, cp - LD3(w.subtree_com, rootid[b1])));
      float d[3];
#pragma unroll
      for (int r = 0; r < 3; r++)
        d[r] = w.con_frame[3 * r][c] * diff.x + w.con_frame[3 * r + 1][c] * diff.y + w.con_frame[3 * r + 2][c] * diff.z;
      const int rd = adr - nsparse;
      if (pdim[p] == 1) {
        w.Jd[rd][i] = d[0];
      } else {
        const float mu0 = pfric[5 * p], mu1 = pfric[5 * p + 1];
        w.Jd[rd + 0][i] = d[0] + mu0 * d[1];
        w.Jd[rd + 1][i] = d[0] - mu0 * d[1];
        w.Jd[rd + 2][i] = d[0] + mu1 * d[2];
        w.Jd[rd + 3][i] = d[0] - mu1 * d[2];
      }
    }
  }
  WSYNC();
  const float* binvw = MF(m, MH_BODY_INVWEIGHT0);
  const float* psolref = MF(m, MH_PAIR_SOLREF);
  const float* psolimp = MF(m, MH_PAIR_SOLIMP);
  const float impratio = m.mf[MF_IMPRATIO];
  LANE_FOR(rd, nrow - nsparse) {
    const int r = nsparse + rd;
    // owning contact of this row
    int c = 0;
    for (int cc = 0; cc < ncon; cc++) if (w.con_efc[cc] >= 0 && w.con_efc[cc] <= r) c = cc;
    const int p = w.con_pair[c];
    float vel = 0.0f;
    for (int i = 0; i < nv; i++) vel += w.Jd[rd][i] * w.qvel[i];
    const int b1 = gbody[pg1[p]], b2 = gbody[pg2[p]];
    const float t = binvw[2 * b1] + binvw[2 * b2];
    const float pos = w.con_dist[c] - (pmargin[p] - pgap[p]);
    float iw = t;
    if (pdim[p] != 1) {
      const float mu0 = pfric[5 * p];
      iw = (t + mu0 * mu0 * t) * 2 * mu0 * mu0 / impratio;
    }
    w.efc_dof[r] = -1; w.efc_sign[r] = 0.0f;
    fill_row(m, w, r, EFC_CONTACT, vel, pos, iw, psolref + 2 * p, psolimp + 5 * p, 0.0f);
  }
  WSYNC();
}

// ------------------------------------------------------------------------------------------ velocity stage

// Spatial velocity of each body and cdof_dot of each dof, by walking the dof chain instead of recursing over
// bodies: cvel[b] = sum of cdof*qvel over the dofs between the root and b (same quantity mj_comVel builds).
template <class D>
DEV void com_vel(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY), nv = DIM(NV, MH_NV);
  const int* dof_parent = MI(m, MH_DOF_PARENTID);
  const int* dof_velparent = MI(m, MH_DOF_VELPARENT);
  const int* body_lastdof = MI(m, MH_BODY_LASTDOF);
  LANE_FOR(b, nb) {
    V6 v;
#pragma unroll
    for (int c = 0; c < 6; c++) v.v[c] = 0.0f;
    for (int j = body_lastdof[b]; j >= 0; j = dof_parent[j]) {
      const float qv = w.qvel[j];
#pragma unroll
      for (int c = 0; c < 6; c++) v.v[c] += w.cdof[c][j] * qv;
    }
    st6(w.cvel, b, v);
  }
  LANE_FOR(i, nv) {
    const int vp = dof_velparent[i];
    V6 r;
    if (vp == -2) {  // translational dof of a free joint
#pragma unroll
      for (int c = 0; c < 6; c++) r.v[c] = 0.0f;
    } else {
      V6 v;
#pragma unroll
      for (int c = 0; c < 6; c++) v.v[c] = 0.0f;
      for (int j = vp; j >= 0; j = dof_parent[j]) {
        const float qv = w.qvel[j];
#pragma unroll
        for (int c = 0; c < 6; c++) v.v[c] += w.cdof[c][j] * qv;
      }
      r = cross_motion(v, ld6(w.cdof, i));
    }
    st6(w.cdof_dot, i, r);
  }
  WSYNC();
}

template <class D>
DEV void passive(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nv = DIM(NV, MH_NV);
  const int* dof_jnt = MI(m, MH_DOF_JNTID);
  const int* jtype = MI(m, MH_JNT_TYPE);
  const int* jqpos = MI(m, MH_JNT_QPOSADR);
  const float* stiff = MF(m, MH_JNT_STIFFNESS);
  const float* spring = MF(m, MH_QPOS_SPRING);
  LANE_FOR(i, nv) {
    float f = -w.damping[i] * w.qvel[i];
    const int j = dof_jnt[i];
    if (jtype[j] == JNT_HINGE) f -= stiff[j] * (w.qpos[jqpos[j]] - spring[jqpos[j]]);
    w.qfrc_passive[i] = f;
  }
}

// cacc[b] = -gravity + sum over the dof chain of cdof_dot*qvel (+ cdof*qacc)
template <class D>
DEVNI void rne_cacc(const Mdl& m, Work<D>& w, bool with_acc, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY);
  const int* dof_parent = MI(m, MH_DOF_PARENTID);
  const int* body_lastdof = MI(m, MH_BODY_LASTDOF);
  const float gx = m.mf[MF_GRAVITY], gy = m.mf[MF_GRAVITY + 1], gz = m.mf[MF_GRAVITY + 2];
  LANE_FOR(b, nb) {
    V6 a;
    a.v[0] = a.v[1] = a.v[2] = 0.0f; a.v[3] = -gx; a.v[4] = -gy; a.v[5] = -gz;
    for (int j = body_lastdof[b]; j >= 0; j = dof_parent[j]) {
      const float qv = w.qvel[j];
#pragma unroll
      for (int c = 0; c < 6; c++) a.v[c] += w.cdof_dot[c][j] * qv;
      if (with_acc) {
        const float qa = w.qacc[j];
#pragma unroll
        for (int c = 0; c < 6; c++) a.v[c] += w.cdof[c][j] * qa;
      }
    }
    st6(w.cacc, b, a);
  }
  WSYNC();
}

template <class D>
DEV V6 body_inertial_force(Work<D>& w, int b) {
  float ci[10];
#pragma unroll
  for (int k = 0; k < 10; k++) ci[k] = w.cinert[k][b];
  V6 v = ld6(w.cvel, b);
  V6 t1 = mul_inert_vec(ci, ld6(w.cacc, b));
  V6 t2 = mul_inert_vec(ci, v);
  V6 t3 = cross_force(v, t2);
  V6 r;
#pragma unroll
  for (int c = 0; c < 6; c++) r.v[c] = t1.v[c] + t3.v[c];
  return r;
}

template <class D>
DEV void rne(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY), nv = DIM(NV, MH_NV);
  const int* sub_end = MI(m, MH_BODY_SUBTREE_END);
  const int* dof_body = MI(m, MH_DOF_BODYID);
  rne_cacc(m, w, false, LANE_PASS);
  LANE_FOR(b, nb) {
    V6 f;
    if (b == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) f.v[c] = 0.0f;
    } else {
      f = body_inertial_force(w, b);
    }
    st6(w.cfrc, b, f);
  }
  WSYNC();
  subtree_accumulate<6>(m, w.cfrc, LANE_PASS);
  LANE_FOR(i, nv) {
    const int b = dof_body[i];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 6; k++) s += w.cdof[k][i] * w.cfrc[k][b];
    w.qfrc_bias[i] = s;
  }
  WSYNC();
}

template <class D>
DEV void actuation(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), nu = DIM(NU, MH_NU);
  const int* trnid = MI(m, MH_ACT_TRNID);
  const int* climited = MI(m, MH_ACT_CTRLLIMITED);
  const int* flimited = MI(m, MH_ACT_FORCELIMITED);
  const int* jdof = MI(m, MH_JNT_DOFADR);
  const float* gear = MF(m, MH_ACT_GEAR);
  const float* crange = MF(m, MH_ACT_CTRLRANGE);
  const float* frange = MF(m, MH_ACT_FORCERANGE);
  LANE_FOR(u, nu) {
    float c = w.ctrl[u];
    if (climited[u]) c = clipf(c, crange[2 * u], crange[2 * u + 1]);
    float f = c;
    if (flimited[u]) f = clipf(f, frange[2 * u], frange[2 * u + 1]);
    w.actuator_force[u] = f;
  }
  WSYNC();
  // MH_DOF_ACT names the one motor of a dof (the usual case); several motors on one dof are gathered
  const int* dof_act = MI(m, MH_DOF_ACT);
  LANE_FOR(i, nv) {
    const int ua = dof_act[i];
    float s = 0.0f;
    if (ua >= 0) s = gear[ua] * w.actuator_force[ua];
    else if (ua == -2) {
      for (int u = 0; u < nu; u++) if (jdof[trnid[u]] == i) s += gear[u] * w.actuator_force[u];
    }
    w.qfrc_actuator[i] = s;
  }
  WSYNC();
}

template <class D>
DEV void acceleration(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nv = DIM(NV, MH_NV);
  const int* rootid = MI(m, MH_BODY_ROOTID);
  const int* bdofmask = MI(m, MH_BODY_DOFMASK);
  const int xb = w.xfrc_body;
  LANE_FOR(i, nv) {
    float s = w.qfrc_passive[i] - w.qfrc_bias[i] + w.qfrc_actuator[i];
    if (xb > 0 && ((bdofmask[xb] >> i) & 1)) {
      const bool nz = w.xfrc[0] != 0 || w.xfrc[1] != 0 || w.xfrc[2] != 0 || w.xfrc[3] != 0 || w.xfrc[4] != 0 || w.xfrc[5] != 0;
      if (nz) {
        V3 ang = v3(w.cdof[0][i], w.cdof[1][i], w.cdof[2][i]), lin = v3(w.cdof[3][i], w.cdof[4][i], w.cdof[5][i]);
        V3 jp = lin + cross(ang, LD3(w.xipos, xb) - LD3(w.subtree_com, rootid[xb]));
        s += jp.x * w.xfrc[0] + ang.x * w.xfrc[3];
        s += jp.y * w.xfrc[1] + ang.y * w.xfrc[4];
        s += jp.z * w.xfrc[2] + ang.z * w.xfrc[5];
      }
    }
    w.qfrc_smooth[i] = s;
  }
  WSYNC();
  // with 1..NDUAL constraint rows the Newton solve runs in constraint space and needs Z = M^-1 J^T (solve())
  const int ne = w.nefc;
  factor_solve(m, w, FS_MASS, w.qacc_smooth, w.qfrc_smooth, dual_path(m, w, ne) ? ne : 0, LANE_PASS);
}

// ------------------------------------------------------------------------------------------------- solver

// J * x for row r
template <class D>
DEV float row_dot(const Work<D>& w, int r, int nsparse, const float* x, int nv) {
  if (r < nsparse) return w.efc_sign[r] * x[w.efc_dof[r]];
  float s = 0.0f;
  const float* J = w.Jd[r - nsparse];
  for (int i = 0; i < nv; i++) s += J[i] * x[i];
  return s;
}

// bit r = efc_active[r] (ne <= 32)
template <class D>
DEV unsigned active_mask(const Work<D>& w, int ne, LANE_ARG) {
#if B2S_DEVICE
  return __ballot_sync(0xffffffffu, lane < ne && w.efc_active[lane < ne ? lane : 0]);
#else
  unsigned mk = 0u;
  for (int r = 0; r < ne && r < 32; r++) mk |= w.efc_active[r] ? (1u << r) : 0u;
  return mk;
#endif
}

// cost of constraint row r at Jaref = x (same pieces as constraint_rows)
template <class D>
DEV float constraint_row_cost(const Work<D>& w, int r, float x) {
  const float Dr = w.efc_D[r];
  if (w.efc_type[r] == EFC_FRICTION) {
    const float f = w.efc_floss[r], rf = w.efc_R[r] * f;
    if (x <= -rf) return -f * (0.5f * rf + x);
    if (x >= rf) return -f * (0.5f * rf - x);
    return 0.5f * Dr * x * x;
  }
  return x < 0.0f ? 0.5f * Dr * x * x : 0.0f;
}

// efc_force / efc_active from efc_Jaref; returns the constraint cost (summed over the warp)
template <class D>
DEV float constraint_rows(Work<D>& w, int ne, LANE_ARG) {
  float cost = 0.0f;
  LANE_FOR(r, ne) {
    const float x = w.efc_Jaref[r], Dr = w.efc_D[r];
    if (w.efc_type[r] == EFC_FRICTION) {
      const float f = w.efc_floss[r], rf = w.efc_R[r] * f;
      if (x <= -rf) { w.efc_force[r] = f; w.efc_active[r] = 0; cost += -f * (0.5f * rf + x); }
      else if (x >= rf) { w.efc_force[r] = -f; w.efc_active[r] = 0; cost += -f * (0.5f * rf - x); }
      else { w.efc_force[r] = -Dr * x; w.efc_active[r] = 1; cost += 0.5f * Dr * x * x; }
    } else {
      const int act = x < 0.0f;
      w.efc_active[r] = act;
      w.efc_force[r] = act ? -Dr * x : 0.0f;
      if (act) cost += 0.5f * Dr * x * x;
    }
  }
  cost = wsum(cost);
  WSYNC();
  return cost;
}

// qfrc_constraint = J^T efc_force
template <class D>
DEV void constraint_force_to_dofs(Work<D>& w, int nv, int ne, int nsparse, LANE_ARG) {
  LANE_FOR(i, nv) {
    float s = 0.0f;
    for (int r = 0; r < nsparse; r++) if (w.efc_dof[r] == i) s += w.efc_sign[r] * w.efc_force[r];
    for (int r = nsparse; r < ne; r++) s += w.Jd[r - nsparse][i] * w.efc_force[r];
    w.qfrc_constraint[i] = s;
  }
  WSYNC();
}

// efc_force / active / cost from Jaref; qfrc_constraint = J^T force; returns total cost (gauss + constraint)
template <class D>
DEVNI float solver_update_constraint(const Mdl& m, Work<D>& w, float* gauss_out, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  const float cost = constraint_rows(w, ne, LANE_PASS);
  float gauss = 0.0f;
  LANE_FOR(i, nv) gauss += (w.s_Ma[i] - w.qfrc_smooth[i]) * (w.s_qacc[i] - w.qacc_smooth[i]);
  gauss = 0.5f * wsum(gauss);
  constraint_force_to_dofs(w, nv, ne, nsparse, LANE_PASS);
  *gauss_out = gauss;
  return gauss + cost;
}

// grad = Ma - qfrc_smooth - qfrc_constraint; Mgrad = H^-1 grad with H = M + J^T diag(D*active) J
template <class D>
DEVNI void solver_update_gradient(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nv = DIM(NV, MH_NV);
  LANE_FOR(i, nv) w.s_grad[i] = w.s_Ma[i] - w.qfrc_smooth[i] - w.qfrc_constraint[i];
  WSYNC();
  factor_solve(m, w, FS_NEWTON, w.s_Mgrad, w.s_grad, 0, LANE_PASS);
}

// initialises s_qacc/s_Ma/Jaref from `qacc0`; returns cost
template <class D>
DEVNI float solver_init(const Mdl& m, Work<D>& w, const float* qacc0, float* gauss, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  LANE_FOR(i, nv) {
    w.s_qacc[i] = qacc0[i];
    float s = 0.0f;
    for (int j = 0; j < nv; j++) s += w.M[i][j] * qacc0[j];
    w.s_Ma[i] = s;
  }
  LANE_FOR(r, ne) w.efc_Jaref[r] = row_dot(w, r, nsparse, qacc0, nv) - w.efc_aref[r];
  WSYNC();
  return solver_update_constraint(m, w, gauss, LANE_PASS);
}

struct LSPoint { float alpha, cost, deriv0, deriv1; };

DEV LSPoint ls_make(float alpha, float q0, float q1, float q2) {
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv0 = 2 * alpha * q2 + q1;
  p.deriv1 = 2 * q2 + (q2 == 0 ? MINVAL : 0.0f);
  return p;
}
DEV bool in_bracket(const LSPoint& x, const LSPoint& y) {
  return (x.deriv0 < y.deriv0 && y.deriv0 < 0) || (x.deriv0 > y.deriv0 && y.deriv0 > 0);
}

// one constraint row's contribution to the quadratic model of the cost at step size alpha
struct LSRow { float ja, jv, D, f, rf; int type; };

template <class D>
DEV LSRow ls_row(const Work<D>& w, int r) {
  LSRow o;
  o.ja = w.efc_Jaref[r]; o.jv = w.efc_jv[r]; o.D = w.efc_D[r]; o.type = w.efc_type[r];
  o.f = w.efc_floss[r]; o.rf = w.efc_R[r] * o.f;
  return o;
}
DEV void ls_accumulate(const LSRow& o, float alpha, float& q0, float& q1, float& q2) {
  const float x = o.ja + alpha * o.jv;
  // quadratic zone (limit / contact rows: x < 0; friction rows: |x| < R f), else a friction row's linear zones;
  // written with selects so the three step sizes of a warp do not diverge
  const bool fr = o.type == EFC_FRICTION;
  const bool quad = fr ? (x > -o.rf && x < o.rf) : (x < 0.0f);
  const float dj = o.D * o.ja;
  const float sg = (x <= -o.rf) ? -1.0f : 1.0f;  // friction, linear zones: force -sg f
  const bool lin = fr && !quad;
  q0 += quad ? 0.5f * dj * o.ja : (lin ? o.f * (-0.5f * o.rf + sg * o.ja) : 0.0f);
  q1 += quad ? dj * o.jv : (lin ? sg * o.f * o.jv : 0.0f);
  q2 += quad ? 0.5f * o.D * o.jv * o.jv : 0.0f;
}

// Evaluates three step sizes in one pass over the constraint rows.
// Device: lanes [8g, 8g+8) evaluate alpha[g] over rows (lane & 7), +8, ...; a 3-step butterfly inside each group of 8
// sums the three quadratic coefficients for all step sizes at once (9 shuffles instead of 45), and the three
// resulting points are handed to every lane with 9 more.  `row0` is row (lane & 7), loaded once per line search.
template <class D>
DEV void ls_eval3(const Work<D>& w, int ne, const LSRow& row0, const float* alpha, const float* qg, LSPoint* out, LANE_ARG) {
#if B2S_DEVICE
  constexpr unsigned FULL = 0xffffffffu;
  const int g = lane >> 3, sub = lane & 7;
  const float a = g == 0 ? alpha[0] : (g == 1 ? alpha[1] : alpha[2]);
  float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f;
  if (sub < ne) ls_accumulate(row0, a, q0, q1, q2);
  for (int r = sub + 8; r < ne; r += 8) ls_accumulate(ls_row(w, r), a, q0, q1, q2);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    q0 += __shfl_xor_sync(FULL, q0, o); q1 += __shfl_xor_sync(FULL, q1, o); q2 += __shfl_xor_sync(FULL, q2, o);
  }
  const LSPoint mine = ls_make(a, q0 + qg[0], q1 + qg[1], q2 + qg[2]);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    out[k].alpha = alpha[k];
    out[k].cost = __shfl_sync(FULL, mine.cost, 8 * k);
    out[k].deriv0 = __shfl_sync(FULL, mine.deriv0, 8 * k);
    out[k].deriv1 = __shfl_sync(FULL, mine.deriv1, 8 * k);
  }
#else
  (void)row0;
  for (int k = 0; k < 3; k++) {
    float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f;
    for (int r = 0; r < ne; r++) ls_accumulate(ls_row(w, r), alpha[k], q0, q1, q2);
    out[k] = ls_make(alpha[k], q0 + qg[0], q1 + qg[1], q2 + qg[2]);
  }
#endif
}

// d/dalpha of one row's cost at step size alpha:  jv * s'(ja + alpha jv)
DEV float ls_row_slope(const LSRow& o, float alpha) {
  const float x = o.ja + alpha * o.jv;
  if (o.type == EFC_FRICTION) return o.jv * (x <= -o.rf ? -o.f : (x >= o.rf ? o.f : o.D * x));
  return x < 0.0f ? o.jv * o.D * x : 0.0f;
}

// EXACT line search.  Along the search direction the cost is convex, C1 and piecewise quadratic: the Gauss term is a
// parabola and every row switches between a quadratic and a linear (or zero) piece where ja + alpha jv crosses 0 (limit /
// contact rows) or +-R f (friction rows).  Its derivative g is therefore continuous, piecewise linear and non-decreasing,
// and the minimiser is g's root: evaluate g at every breakpoint (one lane per breakpoint), take the last breakpoint
// with g <= 0 and the first with g > 0, and solve the linear piece between them in closed form.  One pass instead of
// the reference's up to 2 + 3 x 8 bracketing evaluations, and the step is the exact minimiser rather than an
// 8-iteration approximation of it; the Newton iteration converges to the same (unique) minimum.  Needs two breakpoint
// slots per row: ne <= 16.
template <class D>
DEV float ls_exact(const Work<D>& w, int ne, float qg1, float qg2, LANE_ARG) {
  float g0 = qg1;
#if B2S_DEVICE
  constexpr unsigned FULL = 0xffffffffu;
  const int slot_row = lane >> 1;
  const bool have = slot_row < ne;
  const LSRow mine = ls_row(w, have ? slot_row : 0);
  // this lane's breakpoint: row (lane >> 1), threshold 0 (limit / contact; odd slot unused) or -+R f (friction)
  float thr = 0.0f;
  bool used = have && mine.jv != 0.0f;
  if (mine.type == EFC_FRICTION) thr = (lane & 1) ? mine.rf : -mine.rf;
  else used = used && !(lane & 1);
  float ab = used ? __fdividef(thr - mine.ja, mine.jv) : -1.0f;
  used = used && ab > 0.0f;
  // g at this lane's breakpoint and at 0: sum over the rows (row r lives in lane 2r)
  float gb = fmaf(2.0f * qg2, ab, qg1);
  for (int r = 0; r < ne; r++) {
    LSRow o;
    o.ja = __shfl_sync(FULL, mine.ja, 2 * r); o.jv = __shfl_sync(FULL, mine.jv, 2 * r); o.D = __shfl_sync(FULL, mine.D, 2 * r);
    o.f = __shfl_sync(FULL, mine.f, 2 * r); o.rf = __shfl_sync(FULL, mine.rf, 2 * r); o.type = __shfl_sync(FULL, mine.type, 2 * r);
    gb += ls_row_slope(o, ab);
    g0 += ls_row_slope(o, 0.0f);
  }
  if (!(g0 < 0.0f)) return 0.0f;  // not a descent direction (round-off at convergence)
  // bracket: alpha > 0 orders like its bit pattern
  const unsigned lo_bits = __reduce_max_sync(FULL, (used && gb <= 0.0f) ? __float_as_uint(ab) : 0u);
  const unsigned hi_bits = __reduce_min_sync(FULL, (used && gb > 0.0f) ? __float_as_uint(ab) : 0x7f800000u);
  const float lo = __uint_as_float(lo_bits);
  const bool open = hi_bits == 0x7f800000u;
  const float hi = open ? lo + 1.0f : __uint_as_float(hi_bits);
  const float mid = 0.5f * (lo + hi);
  // the linear piece of g around mid: g(alpha) = G1 + alpha G2
  float G1 = 0.0f, G2 = 0.0f;
  if (have && !(lane & 1)) {
    const float x = mine.ja + mid * mine.jv;
    if (mine.type == EFC_FRICTION) {
      if (x <= -mine.rf) G1 = -mine.f * mine.jv;
      else if (x >= mine.rf) G1 = mine.f * mine.jv;
      else { G1 = mine.D * mine.ja * mine.jv; G2 = mine.D * mine.jv * mine.jv; }
    } else if (x < 0.0f) { G1 = mine.D * mine.ja * mine.jv; G2 = mine.D * mine.jv * mine.jv; }
  }
  G1 = wsum(G1) + qg1;
  G2 = wsum(G2) + 2.0f * qg2;
  float alpha = G2 > 0.0f ? __fdividef(-G1, G2) : lo;
  alpha = fmaxf(alpha, lo);
  if (!open) alpha = fminf(alpha, hi);
  return alpha;
#else
  for (int r = 0; r < ne; r++) g0 += ls_row_slope(ls_row(w, r), 0.0f);
  if (!(g0 < 0.0f)) return 0.0f;
  float lo = 0.0f, hi = INFINITY;
  for (int r = 0; r < ne; r++) {
    const LSRow o = ls_row(w, r);
    if (o.jv == 0.0f) continue;
    for (int k = 0; k < 2; k++) {
      float thr = 0.0f;
      if (o.type == EFC_FRICTION) thr = k ? o.rf : -o.rf;
      else if (k) continue;
      const float ab = (thr - o.ja) / o.jv;
      if (!(ab > 0.0f)) continue;
      float gb = 2.0f * qg2 * ab + qg1;
      for (int q = 0; q < ne; q++) gb += ls_row_slope(ls_row(w, q), ab);
      if (gb <= 0.0f) lo = fmaxf(lo, ab); else hi = fminf(hi, ab);
    }
  }
  const bool open = !(hi < INFINITY);
  const float mid = 0.5f * (lo + (open ? lo + 1.0f : hi));
  float G1 = qg1, G2 = 2.0f * qg2;
  for (int r = 0; r < ne; r++) {
    const LSRow o = ls_row(w, r);
    const float x = o.ja + mid * o.jv;
    if (o.type == EFC_FRICTION) {
      if (x <= -o.rf) G1 += -o.f * o.jv;
      else if (x >= o.rf) G1 += o.f * o.jv;
      else { G1 += o.D * o.ja * o.jv; G2 += o.D * o.jv * o.jv; }
    } else if (x < 0.0f) { G1 += o.D * o.ja * o.jv; G2 += o.D * o.jv * o.jv; }
  }
  float alpha = G2 > 0.0f ? -G1 / G2 : lo;
  alpha = fmaxf(alpha, lo);
  if (!open) alpha = fminf(alpha, hi);
  return alpha;
#endif
}

// The bracketing line search of the Newton solver over efc_Jaref / efc_jv and the quadratic Gauss term qg
// (cost(alpha) = qg0 + alpha qg1 + alpha^2 qg2 + sum of the rows); returns the step size.
// `live` == false (lockstep only): this warp's solve has finished; it only takes part in the CTA votes.
template <class D>
DEV float ls_bracket(const Mdl& m, const Work<D>& w, int nv, int ne, const float* qg, float snorm, bool lockstep, bool live,
                     LANE_ARG) {
  lockstep = false;  // votes per line-search iteration measured slower than per Newton iteration only (profiles/)
  if (!live) {
    if (lockstep) while (CTA_OR(0)) {}
    return 0.0f;
  }
  const float scale = m.mf[MF_MEANINERTIA] * (float)(nv > 1 ? nv : 1);
  const float gtol = m.mf[MF_TOLERANCE] * m.mf[MF_LS_TOLERANCE] * snorm * scale;
#if B2S_DEVICE
  const LSRow row0 = ls_row(w, (lane & 7) < ne ? (lane & 7) : 0);
#else
  const LSRow row0 = ls_row(w, 0);
#endif
  LSPoint pts[3];
  float al[3] = {0.0f, 0.0f, 0.0f};
  ls_eval3(w, ne, row0, al, qg, pts, LANE_PASS);
  const LSPoint p0 = pts[0];
  al[0] = al[1] = al[2] = p0.alpha - fast_div(p0.deriv0, p0.deriv1);
  ls_eval3(w, ne, row0, al, qg, pts, LANE_PASS);
  LSPoint lo = pts[0], hi;
  if (p0.deriv0 < lo.deriv0) { hi = lo; lo = p0; } else { hi = p0; }
  int swap = 1, it = 0;
  const int ls_iterations = m.h[MH_LS_ITERATIONS];
  while (true) {
    bool done = it >= ls_iterations;
    done |= !swap;
    done |= (lo.deriv0 < 0) && (lo.deriv0 > -gtol);
    done |= (hi.deriv0 > 0) && (hi.deriv0 < gtol);
    // lockstep: the CTA's warps iterate together (one instruction-cache fill serves all) until none has work left
    if (lockstep ? !CTA_OR(!done) : done) break;
    if (done) continue;
    al[0] = lo.alpha - fast_div(lo.deriv0, lo.deriv1);
    al[1] = hi.alpha - fast_div(hi.deriv0, hi.deriv1);
    al[2] = 0.5f * (lo.alpha + hi.alpha);
    ls_eval3(w, ne, row0, al, qg, pts, LANE_PASS);
    const LSPoint lo_next = pts[0], hi_next = pts[1], mid = pts[2];
    const int s1 = in_bracket(lo, lo_next); if (s1) lo = lo_next;
    const int s2 = in_bracket(lo, mid); if (s2) lo = mid;
    const int s3 = in_bracket(lo, hi_next); if (s3) lo = hi_next;
    const int s4 = in_bracket(hi, hi_next); if (s4) hi = hi_next;
    const int s5 = in_bracket(hi, mid); if (s5) hi = mid;
    const int s6 = in_bracket(hi, lo_next); if (s6) hi = lo_next;
    swap = s1 | s2 | s3 | s4 | s5 | s6;
    it++;
  }
  const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  float alpha = (lo.cost < hi.cost) ? lo.alpha : hi.alpha;
  if (!improved) alpha = 0.0f;
  return alpha;
}

// primal (dof-space) line search: search direction s_search, state s_qacc / s_Ma / efc_Jaref
template <class D>
DEV void linesearch(const Mdl& m, Work<D>& w, float gauss, bool lockstep, bool live, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  if (!live) {
    ls_bracket(m, w, nv, ne, nullptr, 0.0f, lockstep, false, LANE_PASS);
    return;
  }
  float snorm = 0.0f, qg1 = 0.0f, qg2 = 0.0f;
  LANE_FOR(i, nv) {
    float s = 0.0f;
    for (int j = 0; j < nv; j++) s += w.M[i][j] * w.s_search[j];
    w.s_mv[i] = s;
    const float si = w.s_search[i];
    snorm += si * si;
    qg1 += si * (w.s_Ma[i] - w.qfrc_smooth[i]);
    qg2 += 0.5f * si * s;
  }
  LANE_FOR(r, ne) w.efc_jv[r] = row_dot(w, r, nsparse, w.s_search, nv);
  snorm = sqrtf(wsum(snorm));
  const float qg[3] = {gauss, wsum(qg1), wsum(qg2)};
  WSYNC();
  const float alpha = ls_bracket(m, w, nv, ne, qg, snorm, lockstep, true, LANE_PASS);
  LANE_FOR(i, nv) { w.s_qacc[i] += alpha * w.s_search[i]; w.s_Ma[i] += alpha * w.s_mv[i]; }
  LANE_FOR(r, ne) w.efc_Jaref[r] += alpha * w.efc_jv[r];
  WSYNC();
}

// `lockstep` (bit 7 of sync_mask): every warp of the CTA goes through the same sequence of CTA votes -- one per
// Newton iteration and one per line-search iteration -- so the warps stay on the same code while they iterate.
// solve_follower() is the vote sequence of a warp that has no environment or no constraint rows.
DEV void solve_follower(const Mdl& m) {
  while (CTA_OR(0)) {}
}

// Newton solver in dof space: re-factorises H = M + J^T diag(D active) J every iteration.  Used when the
// environment has more constraint rows than the constraint-space path holds (ne > NDUAL).
template <class D>
DEV void solve_primal(const Mdl& m, Work<D>& w, bool lockstep, LANE_ARG) {
  const int nv = DIM(NV, MH_NV);
  float gauss;
  const float* start = w.qacc_smooth;
  if (!m.h[MH_DISABLE_WARMSTART]) {
    const float cw = solver_init(m, w, w.warm, &gauss, LANE_PASS);
    const float cs = solver_init(m, w, w.qacc_smooth, &gauss, LANE_PASS);
    if (cw < cs) start = w.warm;
  }
  float cost = solver_init(m, w, start, &gauss, LANE_PASS);
  float prev_cost = INFINITY;
  solver_update_gradient(m, w, LANE_PASS);
  LANE_FOR(i, nv) w.s_search[i] = -w.s_Mgrad[i];
  WSYNC();
  const int iterations = m.h[MH_ITERATIONS];
  const float tol = m.mf[MF_TOLERANCE];
  const float scale = 1.0f / (m.mf[MF_MEANINERTIA] * (float)(nv > 1 ? nv : 1));
  int niter = 0;
  bool finished = false;
  while (true) {
    if (!finished && (iterations != 1 || niter > 0)) {
      const float improvement = (prev_cost - cost) * scale;
      float gn = 0.0f;
      LANE_FOR(i, nv) gn += w.s_grad[i] * w.s_grad[i];
      const float gradient = sqrtf(wsum(gn)) * scale;
      finished = niter >= iterations;
      finished |= improvement < tol;
      finished |= gradient < tol;
    }
    if (lockstep ? !CTA_OR(!finished) : finished) break;
    linesearch(m, w, gauss, lockstep, !finished, LANE_PASS);
    if (finished) continue;
    prev_cost = cost;
    cost = solver_update_constraint(m, w, &gauss, LANE_PASS);
    solver_update_gradient(m, w, LANE_PASS);
    LANE_FOR(i, nv) w.s_search[i] = -w.s_Mgrad[i];
    WSYNC();
    niter++;
    if (iterations == 1) finished = true;
  }
  LANE_FOR(i, nv) { w.qacc[i] = w.s_qacc[i]; w.warm[i] = w.s_qacc[i]; }
  WSYNC();
}

// ------------------------------------------------------------------- dof-space Newton with factor reuse (any ne)
//
// The same Newton iteration as solve_primal, organised around what changes between iterations.  With D fixed, the
// Hessian H = M + J^T diag(D active) J depends on the iterate only through the ACTIVE SET, so it is rebuilt and
// refactorised only when a row switched piece; otherwise the factor in W is reused (substitutions only).  The iterate is
// kept as delta = qacc - qacc_smooth with M delta alongside, so grad = M delta - J^T f and the Gauss term is
// 0.5 delta^T M delta (no M qacc products); single-dof rows (friction loss, limits) enter H's diagonal and J^T f by
// per-dof scatter instead of per-lane loops over all rows; the line search is exact (ls_newton).  A full step that
// keeps the active set has landed on the minimiser of its quadratic piece, which for the convex cost is the minimum.
// This is the path of every model whose rows do not fit the register-resident constraint-space solve (the K-Bot:
// 20 friction rows + limits + up to 32 pyramid rows).

// out[i] = sum of val(r) over the single-dof rows r of dof i.  Friction rows [0, nfric) sit on distinct dofs and so
// do limit rows [nfric, nsparse): two conflict-free passes.
template <class D, class F>
DEV void sparse_scatter(Work<D>& w, float* out, int nv, F val, LANE_ARG) {
  const int nfric = w.scratch_i[0], nlim = w.nsparse - nfric;
  LANE_FOR(i, nv) out[i] = 0.0f;
  WSYNC();
  LANE_FOR(r, nfric) out[w.efc_dof[r]] += val(r);
  WSYNC();
  LANE_FOR(k, nlim) out[w.efc_dof[nfric + k]] += val(nfric + k);
  WSYNC();
}

// qfrc_constraint = J^T efc_force (scatter for the single-dof rows, lanes over dofs for the dense ones)
template <class D>
DEV void constraint_force_scatter(Work<D>& w, int nv, int ne, int nsparse, LANE_ARG) {
  sparse_scatter(w, w.qfrc_constraint, nv, [&](int r) { return w.efc_sign[r] * w.efc_force[r]; }, LANE_PASS);
  LANE_FOR(i, nv) {
    float s = w.qfrc_constraint[i];
    for (int r = nsparse; r < ne; r++) s = fmaf(w.Jd[r - nsparse][i], w.efc_force[r], s);
    w.qfrc_constraint[i] = s;
  }
  WSYNC();
}

// constraint_rows() that also reports whether any row switched piece: *hess_changed when an active flag flipped (the
// Hessian changes), *piece_changed also when a friction row jumped between its two linear zones (same Hessian, other
// gradient: a full step across such a jump has not landed on a minimiser)
template <class D>
DEV float constraint_rows_track(Work<D>& w, int ne, int* hess_changed, int* piece_changed, LANE_ARG) {
  float cost = 0.0f;
  int ch = 0, cp = 0;
  LANE_FOR(r, ne) {
    const float x = w.efc_Jaref[r], Dr = w.efc_D[r];
    const int was = w.efc_active[r];
    int act;
    if (w.efc_type[r] == EFC_FRICTION) {
      const float f = w.efc_floss[r], rf = w.efc_R[r] * f, fwas = w.efc_force[r];
      if (x <= -rf) { w.efc_force[r] = f; act = 0; cost += -f * (0.5f * rf + x); cp |= (fwas != f); }
      else if (x >= rf) { w.efc_force[r] = -f; act = 0; cost += -f * (0.5f * rf - x); cp |= (fwas != -f); }
      else { w.efc_force[r] = -Dr * x; act = 1; cost += 0.5f * Dr * x * x; }
    } else {
      act = x < 0.0f;
      w.efc_force[r] = act ? -Dr * x : 0.0f;
      if (act) cost += 0.5f * Dr * x * x;
    }
    w.efc_active[r] = act;
    ch |= (act != was);
  }
  cost = wsum(cost);
  *hess_changed = wany(ch);
  *piece_changed = wany(ch | cp);
  WSYNC();
  return cost;
}

// EXACT line search for any number of rows.  g(alpha) = d cost / d alpha is continuous, piecewise linear and
// non-decreasing (see ls_exact); its root is found by Newton's iteration on g with the slope of the piece the iterate
// is in -- from a point inside the root's piece one step lands exactly on the root -- safeguarded by the bracket
// [lo, hi] of the sign changes seen so far (an iterate leaving it is replaced by the secant point).  Every evaluation is
// one pass with the rows spread over all 32 lanes and two warp sums.  The returned step never overshoots the root's
// piece: if the iteration cap is hit, the last point with g <= 0 is returned, which is a descent step of a convex cost.
template <class D>
DEV float ls_newton(const Work<D>& w, int ne, float qg1, float qg2, LANE_ARG) {
  float lo = 0.0f, hi = INFINITY, glo = 0.0f, ghi = 0.0f;
  float a = 0.0f;
  for (int it = 0; it < 12; it++) {
    float g = 0.0f, h = 0.0f;
    LANE_FOR(r, ne) {
      const LSRow o = ls_row(w, r);
      const float x = fmaf(a, o.jv, o.ja);
      const bool fr = o.type == EFC_FRICTION;
      const bool quad = fr ? (x > -o.rf && x < o.rf) : (x < 0.0f);
      const float dj = o.D * o.jv;
      g += quad ? dj * x : (fr ? (x <= -o.rf ? -o.f : o.f) * o.jv : 0.0f);
      h += quad ? dj * o.jv : 0.0f;
    }
    g = wsum(g) + fmaf(2.0f * qg2, a, qg1);
    h = wsum(h) + 2.0f * qg2;
    if (it == 0) {
      if (!(g < 0.0f)) return 0.0f;  // not a descent direction (round-off at convergence)
      glo = g;
    } else if (g <= 0.0f) { lo = a; glo = g; }
    else { hi = a; ghi = g; }
    float an = h > 0.0f ? a - fast_div(g, h) : (hi < INFINITY ? 0.5f * (lo + hi) : 2.0f * a + 1.0f);
    // converged: the Newton step of this piece no longer moves the iterate (the root's piece has been reached)
    if (fabsf(an - a) <= 1e-6f * fmaxf(fabsf(a), 1e-6f)) return a;
    if (!(an > lo) || !(an < hi)) {
      if (!(hi < INFINITY)) { an = 2.0f * a + 1.0f; }
      else {
        an = lo + (hi - lo) * fast_div(glo, glo - ghi);
        if (!(an > lo) || !(an < hi)) an = 0.5f * (lo + hi);
      }
    }
    a = an;
  }
  // iteration cap: a is the Newton point of the last piece; keep it only if it is inside the bracket's descent side
  return lo > 0.0f ? lo : a;
}

template <class D>
DEV void newton_dof(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  const bool warmstart = !m.h[MH_DISABLE_WARMSTART];
  // ---- start point: costs at qacc_smooth and at the warm start in one pass over the rows
  // (profile build: slots 16..21 = setup / gradient / refactorisation / substitution-only solve / direction + line search /
  //  step + rows; slot 22 counts iterations, slot 23 refactorisations)
  SUB_BEGIN();
  float gw = 0.0f;
  LANE_FOR(i, nv) w.s_search[i] = warmstart ? w.warm[i] - w.qacc_smooth[i] : 0.0f;
  WSYNC();
  LANE_FOR(i, nv) {
    float s = 0.0f;
    for (int j = 0; j < nv; j++) s = fmaf(w.M[i][j], w.s_search[j], s);
    w.s_mv[i] = s;
    gw += w.s_search[i] * s;
  }
  gw = 0.5f * wsum(gw);
  float cs = 0.0f, cw = 0.0f;
  LANE_FOR(r, ne) {
    const float js = row_dot(w, r, nsparse, w.qacc_smooth, nv) - w.efc_aref[r];
    const float jd = row_dot(w, r, nsparse, w.s_search, nv);
    cs += constraint_row_cost(w, r, js);
    cw += constraint_row_cost(w, r, js + jd);
    w.efc_Jaref[r] = js;
    w.efc_jv[r] = jd;
    w.efc_active[r] = -1;  // unknown piece: the first constraint_rows_track() reports a change
  }
  cs = wsum(cs);
  cw = wsum(cw) + gw;
  const bool use_warm = warmstart && cw < cs;
  float gauss = use_warm ? gw : 0.0f;
  // iterate: s_qacc = delta = qacc - qacc_smooth, s_Ma = M delta
  LANE_FOR(i, nv) { w.s_qacc[i] = use_warm ? w.s_search[i] : 0.0f; w.s_Ma[i] = use_warm ? w.s_mv[i] : 0.0f; }
  if (use_warm) { LANE_FOR(r, ne) w.efc_Jaref[r] += w.efc_jv[r]; }
  WSYNC();
  int changed, moved;
  float cost = constraint_rows_track(w, ne, &changed, &moved, LANE_PASS) + gauss;
  constraint_force_scatter(w, nv, ne, nsparse, LANE_PASS);
  float prev_cost = INFINITY;
  SUB_MARK(w, 16);
  const int iterations = m.h[MH_ITERATIONS];
  const float tol = m.mf[MF_TOLERANCE];
  const float scale = 1.0f / (m.mf[MF_MEANINERTIA] * (float)(nv > 1 ? nv : 1));
  bool need_factor = true;
  const bool refshape = newton_dof_refshape(m);
  // The exact line search and the factor reuse make an iteration cheap, and a cold start (after a reset) with many
  // contact rows can need more than the reference's cap to settle its active set: up to 2x `iterations` are allowed, so
  // that what is returned is the converged minimum rather than wherever the cap happened to stop the reference.
#ifndef B2S_NDOF_ITER_NUM
#define B2S_NDOF_ITER_NUM 2   // iteration cap = B2S_NDOF_ITER_NUM / B2S_NDOF_ITER_DEN x the model's `iterations`
#define B2S_NDOF_ITER_DEN 1
#endif
  const int iter_cap = refshape ? iterations : (B2S_NDOF_ITER_NUM * iterations) / B2S_NDOF_ITER_DEN;
  for (int niter = 0; niter < iter_cap; niter++) {
    float gn = 0.0f;
    LANE_FOR(i, nv) { const float g = w.s_Ma[i] - w.qfrc_constraint[i]; w.s_grad[i] = g; gn += g * g; }
    gn = sqrtf(wsum(gn));
    WSYNC();
    SUB_MARK(w, 17);
    if (iterations != 1 || niter > 0) {
      if ((prev_cost - cost) * scale < tol || gn * scale < tol) break;
    }
    PROF_COUNT(w, 22);
    if (need_factor) {
      sparse_scatter(w, w.s_mv, nv, [&](int r) { return w.efc_active[r] ? w.efc_D[r] : 0.0f; }, LANE_PASS);
      factor_solve(m, w, FS_NEWTON_DIAG, w.s_Mgrad, w.s_grad, 0, LANE_PASS);
      SUB_MARK(w, 18);
      PROF_COUNT(w, 23);
    } else {
      fs_resolve(w, w.s_Mgrad, w.s_grad, nv, LANE_PASS);
      SUB_MARK(w, 19);
    }
    LANE_FOR(i, nv) w.s_search[i] = -w.s_Mgrad[i];
    WSYNC();
    float qg1 = 0.0f, qg2 = 0.0f, sn = 0.0f;
    LANE_FOR(i, nv) {
      float s = 0.0f;
      for (int j = 0; j < nv; j++) s = fmaf(w.M[i][j], w.s_search[j], s);
      w.s_mv[i] = s;
      qg1 += w.s_search[i] * w.s_Ma[i];
      qg2 += w.s_search[i] * s;
      sn += w.s_search[i] * w.s_search[i];
    }
    LANE_FOR(r, ne) w.efc_jv[r] = row_dot(w, r, nsparse, w.s_search, nv);
    qg1 = wsum(qg1);
    qg2 = 0.5f * wsum(qg2);
    WSYNC();
    float alpha;
    if (refshape) {
      // the reference's line search (mjx solver._linesearch: bracketing on the derivative, <= ls_iterations rounds, a step
      // is kept only if it lowers the cost); (M qacc - qfrc_smooth) . s == (M delta) . s because M qacc_smooth = qfrc_smooth
      const float qg[3] = {gauss, qg1, qg2};
      alpha = ls_bracket(m, w, nv, ne, qg, sqrtf(wsum(sn)), false, true, LANE_PASS);
    } else {
      alpha = ls_newton(w, ne, qg1, qg2, LANE_PASS);
    }
    SUB_MARK(w, 20);
    if (alpha == 0.0f) break;
    LANE_FOR(i, nv) { w.s_qacc[i] = fmaf(alpha, w.s_search[i], w.s_qacc[i]); w.s_Ma[i] = fmaf(alpha, w.s_mv[i], w.s_Ma[i]); }
    LANE_FOR(r, ne) w.efc_Jaref[r] = fmaf(alpha, w.efc_jv[r], w.efc_Jaref[r]);
    WSYNC();
    gauss += alpha * qg1 + alpha * alpha * qg2;
    prev_cost = cost;
    cost = constraint_rows_track(w, ne, &changed, &moved, LANE_PASS) + gauss;
    constraint_force_scatter(w, nv, ne, nsparse, LANE_PASS);
    need_factor = changed != 0;
    SUB_MARK(w, 21);
    if (!refshape && !moved && fabsf(alpha - 1.0f) < 1e-3f) break;
  }
  LANE_FOR(i, nv) { const float a = w.qacc_smooth[i] + w.s_qacc[i]; w.qacc[i] = a; w.warm[i] = a; }
  WSYNC();
}

// ------------------------------------------------------------------------ constraint-space Newton (ne <= NDUAL)
//
// The same Newton iteration, in the coordinates the iterates actually live in.  With a0 = qacc_smooth, Z = M^-1 J^T
// (from acceleration()), A = J Z and delta0 = warmstart - a0, every iterate is  qacc = a0 + beta delta0 + Z c  with a
// scalar beta and an ne-vector c, because the Newton direction is
//     -H^-1 grad = -delta + Z (f + z - c),   (R_act + A_act,act) z_act = [J delta - A f]_act,  z = 0 off the active set
// (Woodbury on H = M + J_act^T D_act J_act; f = efc_force, R = 1/D).  So instead of re-factorising the nv x nv Hessian
// every iteration, one iteration factors the (at most ne x ne) reduced matrix and works on ne-vectors; the line search
// sees the same efc_Jaref / efc_jv rows and the same quadratic Gauss term as the dof-space solver:
//     J search = -beta j0 + A d,   search^T M delta = -beta^2 m00 + beta j0.(d - c) + c.(A d),
//     search^T M search = beta^2 m00 - 2 beta j0.d + d.(A d),      d = f + z - c, j0 = J delta0, m00 = delta0^T M delta0.

template <class D>
DEV float dual_A(const Work<D>& w, int r, int s) { return s >= r ? w.AS[r][s] : w.AS[s][r]; }

// A (upper triangle of AS), delta0, M delta0, c0 = J a0 - aref, j0 = J delta0; returns m00
template <class D>
DEVNI float dual_setup(const Mdl& m, Work<D>& w, bool warmstart, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  // rows r <= s with a single-dof row among them: A[r][s] is an entry of Z
  LANE_FOR(r, ne) {
    for (int s = r; s < ne; s++) {
      if (s < nsparse) w.AS[r][s] = w.efc_sign[s] * w.Z[r][w.efc_dof[s]];
      else if (r < nsparse) w.AS[r][s] = w.efc_sign[r] * w.Z[s][w.efc_dof[r]];
    }
  }
  LANE_FOR(i, nv) w.du_d0[i] = warmstart ? w.warm[i] - w.qacc_smooth[i] : 0.0f;
  WSYNC();
  // dense x dense entries: dot products over the dofs
  for (int r = nsparse; r < ne; r++) {
    for (int s = r; s < ne; s++) {
      float p = 0.0f;
      LANE_FOR(i, nv) p += w.Jd[r - nsparse][i] * w.Z[s][i];
      p = wsum(p);
      IF_LANE0 w.AS[r][s] = p;
    }
  }
  float m00 = 0.0f;
  if (warmstart) {
    LANE_FOR(i, nv) {
      float s = 0.0f;
      for (int j = 0; j < nv; j++) s += w.M[i][j] * w.du_d0[j];
      w.du_Md0[i] = s;
      m00 += s * w.du_d0[i];
    }
    m00 = wsum(m00);
  } else {
    LANE_FOR(i, nv) w.du_Md0[i] = 0.0f;
  }
  LANE_FOR(r, nsparse) {
    const int dof = w.efc_dof[r];
    w.du_c0[r] = w.efc_sign[r] * w.qacc_smooth[dof] - w.efc_aref[r];
    w.du_j0[r] = w.efc_sign[r] * w.du_d0[dof];
  }
  for (int r = nsparse; r < ne; r++) {
    float p = 0.0f, q = 0.0f;
    LANE_FOR(i, nv) { const float j = w.Jd[r - nsparse][i]; p += j * w.qacc_smooth[i]; q += j * w.du_d0[i]; }
    p = wsum(p); q = wsum(q);
    IF_LANE0 { w.du_c0[r] = p - w.efc_aref[r]; w.du_j0[r] = q; }
  }
  WSYNC();
  return m00;
}

struct DualDir { float qg1, qg2, snorm, gradnorm; };

// Newton direction at the current point (efc_force / efc_active / efc_Jaref up to date): fills du_d and efc_jv and
// returns the line-search coefficients, |search| and |grad| (both in dof space, as the dof-space solver measures them).
#if B2S_DEVICE
// Device fast path for ne <= 8 (what the humanoid sees in practice): lane r < ne owns row r, the rows of A and of the
// reduced system [S | u] sit in registers, and every exchange is a shuffle -- no shared-memory round trips and no
// loops whose trip count the compiler cannot see (each such iteration would cost a full load latency in order).
// S z = u is solved by Gauss-Jordan elimination on the rows (S is SPD: no pivoting), which leaves z in place.
template <class D>
DEV DualDir dual_direction8(const Mdl& m, Work<D>& w, float beta, float m00, bool norms, LANE_ARG) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int K = 8;
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  const bool live = lane < ne;
  const int r = live ? lane : 0;
  float a[K];
#pragma unroll
  for (int s = 0; s < K; s++) a[s] = (live && s < ne) ? dual_A(w, r, s) : 0.0f;
  const float f = live ? w.efc_force[r] : 0.0f, c = live ? w.du_c[r] : 0.0f, j0 = live ? w.du_j0[r] : 0.0f;
  const bool act = live && w.efc_active[r];
  const unsigned am = __ballot_sync(FULL, act);
  // u = J delta - A f (active rows)
  float u = live ? w.efc_Jaref[r] - w.du_c0[r] : 0.0f;
#pragma unroll
  for (int s = 0; s < K; s++) u = fmaf(-a[s], __shfl_sync(FULL, f, s), u);
  // rows of [S | u]: S = R + A on the active set, identity elsewhere
  float row[K];
#pragma unroll
  for (int s = 0; s < K; s++) {
    const bool on = act && ((am >> s) & 1u);
    row[s] = (s == lane) ? (act ? a[s] + w.efc_R[r] : 1.0f) : (on ? a[s] : 0.0f);
  }
  float rhs = act ? u : 0.0f;
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (k < ne) {  // warp-uniform; pivots past the last row would be identity rows
      const float piv = __shfl_sync(FULL, row[k], k);
      const float inv = __fdividef(1.0f, fmaxf(piv, MINVAL));
      const float fac = (lane == k) ? 0.0f : row[k] * inv;
      const float sc = (lane == k) ? inv : 1.0f;
#pragma unroll
      for (int j = k + 1; j < K; j++) row[j] = fmaf(-fac, __shfl_sync(FULL, row[j], k), row[j]) * sc;
      rhs = fmaf(-fac, __shfl_sync(FULL, rhs, k), rhs) * sc;
    }
  }
  const float z = act ? rhs : 0.0f;
  const float d = live ? f + z - c : 0.0f;
  float ad = 0.0f;
#pragma unroll
  for (int s = 0; s < K; s++) ad = fmaf(a[s], __shfl_sync(FULL, d, s), ad);
  if (live) { w.du_d[r] = d; w.efc_jv[r] = ad - beta * j0; }
  float j0c = j0 * c, j0d = j0 * d, cAd = c * ad, dAd = d * ad;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    j0c += __shfl_xor_sync(FULL, j0c, o); j0d += __shfl_xor_sync(FULL, j0d, o);
    cAd += __shfl_xor_sync(FULL, cAd, o); dAd += __shfl_xor_sync(FULL, dAd, o);
  }
  j0c = __shfl_sync(FULL, j0c, 0); j0d = __shfl_sync(FULL, j0d, 0);
  cAd = __shfl_sync(FULL, cAd, 0); dAd = __shfl_sync(FULL, dAd, 0);
  // |search| and |grad| in dof space: search = -beta delta0 + Z d, grad = beta M delta0 + J^T (c - f).  |search| only
  // scales the bracketing line search's tolerance and |grad| only feeds `gradient < tolerance`, which fp32 satisfies
  // solely with an exactly-zero gradient (beta = 0 and c = f on every row): with the exact line search that test is
  // made directly and the two dof-space sums are skipped.
  const float e = c - f;
  float sn = 1.0f, gn;
  if (norms) {
    const bool dof = lane < nv;
    const int i = dof ? lane : 0;
    float sv = dof ? -beta * w.du_d0[i] : 0.0f, gv = dof ? beta * w.du_Md0[i] : 0.0f;
#pragma unroll
    for (int s = 0; s < K; s++) {
      const float ds = __shfl_sync(FULL, d, s), es = __shfl_sync(FULL, e, s);
      if (s < ne && dof) {
        sv = fmaf(w.Z[s][i], ds, sv);
        gv = fmaf(jac_entry(w, s, nsparse, i), es, gv);
      }
    }
    sn = wsum(sv * sv); gn = wsum(gv * gv);
  } else {
    gn = (beta == 0.0f && __all_sync(FULL, e == 0.0f)) ? 0.0f : 1e30f;
  }
  __syncwarp();
  DualDir o;
  o.qg1 = -beta * beta * m00 + beta * (j0d - j0c) + cAd;
  o.qg2 = 0.5f * (beta * beta * m00 - 2.0f * beta * j0d + dAd);
  o.snorm = sqrtf(sn);
  o.gradnorm = sqrtf(gn);
  return o;
}
#endif

template <class D>
DEVNI DualDir dual_direction(const Mdl& m, Work<D>& w, float beta, float m00, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
#if B2S_DEVICE
  if (ne <= 8) return dual_direction8(m, w, beta, m00, !exact_ls_enabled(m), LANE_PASS);
#endif
  // u = J delta - A f on the active rows (into du_d); lower triangle of AS <- active part of A
  LANE_FOR(r, ne) {
    const int ar = w.efc_active[r];
    float u = w.efc_Jaref[r] - w.du_c0[r];
    for (int s = 0; s < ne; s++) u = fmaf(-dual_A(w, r, s), w.efc_force[s], u);
    w.du_d[r] = ar ? u : 0.0f;
    for (int s = 0; s < r; s++) w.AS[r][s] = (ar && w.efc_active[s]) ? w.AS[s][r] : 0.0f;
  }
  WSYNC();
  // Cholesky of S = R_act + A_act,act (identity on inactive rows), column by column; lanes are rows
  for (int j = 0; j < ne; j++) {
    LANE_FOR(i, ne) {
      if (i >= j) {
        float sj = w.efc_active[j] ? w.AS[j][j] + w.efc_R[j] : 1.0f;
        float t = w.AS[i][j];
        for (int k = 0; k < j; k++) {
          const float ljk = w.AS[j][k];
          t = fmaf(-w.AS[i][k], ljk, t);
          sj = fmaf(-ljk, ljk, sj);
        }
        sj = fmaxf(sj, MINVAL);
        const float inv = inv_sqrt(sj);
        if (i == j) w.du_dinv[j] = inv;
        else w.AS[i][j] = t * inv;
      }
    }
    WSYNC();
  }
  // z = S^-1 u, then d = f + z - c
#if B2S_DEVICE
  {
    constexpr unsigned FULL = 0xffffffffu;
    const bool live = lane < ne;
    const int rr = live ? lane : 0;
    const float dinv = live ? w.du_dinv[rr] : 1.0f;
    float y = live ? w.du_d[rr] : 0.0f;
    for (int k = 0; k < ne; k++) {
      const float yk = __shfl_sync(FULL, y * dinv, k);
      if (live && lane > k) y = fmaf(-w.AS[rr][k], yk, y);
    }
    y *= dinv;
    for (int k = ne - 1; k >= 0; k--) {
      const float xk = __shfl_sync(FULL, y * dinv, k);
      if (lane < k) y = fmaf(-w.AS[k][rr], xk, y);
    }
    if (live) w.du_d[rr] = w.efc_force[rr] + y * dinv - w.du_c[rr];
    __syncwarp();
  }
#else
  for (int k = 0; k < ne; k++) {
    w.du_d[k] *= w.du_dinv[k];
    for (int i = k + 1; i < ne; i++) w.du_d[i] -= w.AS[i][k] * w.du_d[k];
  }
  for (int k = ne - 1; k >= 0; k--) {
    w.du_d[k] *= w.du_dinv[k];
    for (int i = 0; i < k; i++) w.du_d[i] -= w.AS[k][i] * w.du_d[k];
  }
  for (int r = 0; r < ne; r++) w.du_d[r] = w.efc_force[r] + w.du_d[r] - w.du_c[r];
#endif
  // J search and the scalar products of the line search
  float j0c = 0.0f, j0d = 0.0f, cAd = 0.0f, dAd = 0.0f;
  LANE_FOR(r, ne) {
    float ad = 0.0f;
    for (int s = 0; s < ne; s++) ad = fmaf(dual_A(w, r, s), w.du_d[s], ad);
    const float j0 = w.du_j0[r], c = w.du_c[r], d = w.du_d[r];
    w.efc_jv[r] = ad - beta * j0;
    j0c += j0 * c; j0d += j0 * d; cAd += c * ad; dAd += d * ad;
  }
  // |search| and |grad| in dof space: search = -beta delta0 + Z d, grad = beta M delta0 + J^T (c - f)
  float sn = 0.0f, gn = 0.0f;
  LANE_FOR(i, nv) {
    float sv = -beta * w.du_d0[i], gv = beta * w.du_Md0[i];
    for (int s = 0; s < ne; s++) sv = fmaf(w.Z[s][i], w.du_d[s], sv);
    for (int r = 0; r < nsparse; r++) if (w.efc_dof[r] == i) gv += w.efc_sign[r] * (w.du_c[r] - w.efc_force[r]);
    for (int r = nsparse; r < ne; r++) gv = fmaf(w.Jd[r - nsparse][i], w.du_c[r] - w.efc_force[r], gv);
    sn += sv * sv; gn += gv * gv;
  }
  j0c = wsum(j0c); j0d = wsum(j0d); cAd = wsum(cAd); dAd = wsum(dAd); sn = wsum(sn); gn = wsum(gn);
  WSYNC();
  DualDir o;
  o.qg1 = -beta * beta * m00 + beta * (j0d - j0c) + cAd;
  o.qg2 = 0.5f * (beta * beta * m00 - 2.0f * beta * j0d + dAd);
  o.snorm = sqrtf(sn);
  o.gradnorm = sqrtf(gn);
  return o;
}

#if B2S_DEVICE
// The constraint-space Newton solve for ne <= 8 rows without friction rows (what the humanoid runs), entirely in
// registers: lane r < ne owns row r -- its row of A, R, D, c0, j0, the iterate (c, Jaref), force and active flag -- and
// every exchange between rows is a shuffle.  No shared-memory traffic, calls or data-dependent loops inside the Newton
// loop; the algorithm is solve_dual's (same direction, exact line search, early finish, polish).
template <class D>
DEVNI void solve_dual8(const Mdl& m, Work<D>& w, bool lockstep, LANE_ARG) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int K = 8;
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  const bool warmstart = !m.h[MH_DISABLE_WARMSTART];
  const float m00 = dual_setup(m, w, warmstart, LANE_PASS);
  const bool live = lane < ne;
  const int r = live ? lane : 0;
  float a[K];
#pragma unroll
  for (int s = 0; s < K; s++) a[s] = (live && s < ne) ? dual_A(w, r, s) : 0.0f;
  const float Rr = live ? w.efc_R[r] : 1.0f, Dr = live ? w.efc_D[r] : 0.0f;
  const float c0 = live ? w.du_c0[r] : 0.0f, j0 = live ? w.du_j0[r] : 0.0f;
  // sum over the rows (lanes 0..7), result in every lane
#define B2S_SUM8(v) do { v += __shfl_xor_sync(FULL, v, 4); v += __shfl_xor_sync(FULL, v, 2); v += __shfl_xor_sync(FULL, v, 1); \
                         v = __shfl_sync(FULL, v, 0); } while (0)
  float beta = 0.0f, gauss = 0.0f;
  if (warmstart) {
    const float xw = c0 + j0;
    float cw = (live && xw < 0.0f) ? 0.5f * Dr * xw * xw : 0.0f, cs = (live && c0 < 0.0f) ? 0.5f * Dr * c0 * c0 : 0.0f;
    B2S_SUM8(cw); B2S_SUM8(cs);
    if (cw + 0.5f * m00 < cs) { beta = 1.0f; gauss = 0.5f * m00; }
  }
  float ja = c0 + beta * j0, c = 0.0f;
  bool act = live && ja < 0.0f;
  float f = act ? -Dr * ja : 0.0f;
  float rc = act ? 0.5f * Dr * ja * ja : 0.0f;
  B2S_SUM8(rc);
  float cost = rc + gauss, prev_cost = INFINITY;
  const int iterations = m.h[MH_ITERATIONS];
  const float tol = m.mf[MF_TOLERANCE];
  const float scale = 1.0f / (m.mf[MF_MEANINERTIA] * (float)(nv > 1 ? nv : 1));
  int niter = 0;
  bool finished = false;
  float d = 0.0f, jv = 0.0f, qg1 = 0.0f, qg2 = 0.0f;
  bool grad_zero = false;
  bool dirty = true;  // the point moved since the direction was last computed
  while (true) {
    // ---- Newton direction at the current point: d (rows), jv = J search, qg1 / qg2 (Gauss term along the search)
    if (dirty) {
      dirty = false;
      float u = live ? ja - c0 : 0.0f;
#pragma unroll
      for (int s = 0; s < K; s++) u = fmaf(-a[s], __shfl_sync(FULL, f, s), u);
      const unsigned am = __ballot_sync(FULL, act);
      float row[K];
#pragma unroll
      for (int s = 0; s < K; s++) {
        const bool on = act && ((am >> s) & 1u);
        row[s] = (s == lane) ? (act ? a[s] + Rr : 1.0f) : (on ? a[s] : 0.0f);
      }
      float rhs = act ? u : 0.0f;
#pragma unroll
      for (int k = 0; k < K; k++) {
        if (k < ne) {  // warp-uniform; pivots past the last row would be identity rows
          const float piv = __shfl_sync(FULL, row[k], k);
          const float inv = __fdividef(1.0f, fmaxf(piv, MINVAL));
          const float fac = (lane == k) ? 0.0f : row[k] * inv;
          const float sc = (lane == k) ? inv : 1.0f;
#pragma unroll
          for (int j = k + 1; j < K; j++) row[j] = fmaf(-fac, __shfl_sync(FULL, row[j], k), row[j]) * sc;
          rhs = fmaf(-fac, __shfl_sync(FULL, rhs, k), rhs) * sc;
        }
      }
      d = live ? f + (act ? rhs : 0.0f) - c : 0.0f;
      float ad = 0.0f;
#pragma unroll
      for (int s = 0; s < K; s++) ad = fmaf(a[s], __shfl_sync(FULL, d, s), ad);
      jv = ad - beta * j0;
      float j0c = j0 * c, j0d = j0 * d, cAd = c * ad, dAd = d * ad;
      B2S_SUM8(j0c); B2S_SUM8(j0d); B2S_SUM8(cAd); B2S_SUM8(dAd);
      qg1 = -beta * beta * m00 + beta * (j0d - j0c) + cAd;
      qg2 = 0.5f * (beta * beta * m00 - 2.0f * beta * j0d + dAd);
      grad_zero = beta == 0.0f && __all_sync(FULL, c - f == 0.0f);
    }
    // ---- convergence (ksim / MJX semantics) and the CTA vote
    if (!finished && (iterations != 1 || niter > 0)) {
      const float improvement = (prev_cost - cost) * scale;
      finished = niter >= iterations;
      finished |= improvement < tol;
      finished |= grad_zero;
    }
    if (lockstep ? !CTA_OR(!finished) : finished) break;
    if (finished) continue;
    // ---- exact line search: root of the piecewise-linear derivative g (ls_exact; one breakpoint per row here)
    float alpha = 0.0f;
    {
      const bool used0 = live && jv != 0.0f;
      const float ab = used0 ? __fdividef(-ja, jv) : -1.0f;
      const bool used = used0 && ab > 0.0f;
      float gb = fmaf(2.0f * qg2, ab, qg1), g0 = qg1;
#pragma unroll
      for (int s = 0; s < K; s++) {
        const float sja = __shfl_sync(FULL, ja, s), sjv = __shfl_sync(FULL, jv, s), sD = __shfl_sync(FULL, Dr, s);
        const float xb = fmaf(ab, sjv, sja);
        gb += xb < 0.0f ? sjv * sD * xb : 0.0f;
        g0 += sja < 0.0f ? sjv * sD * sja : 0.0f;
      }
      if (g0 < 0.0f) {
        const unsigned lo_bits = __reduce_max_sync(FULL, (used && gb <= 0.0f) ? __float_as_uint(ab) : 0u);
        const unsigned hi_bits = __reduce_min_sync(FULL, (used && gb > 0.0f) ? __float_as_uint(ab) : 0x7f800000u);
        const float lo = __uint_as_float(lo_bits);
        const bool open = hi_bits == 0x7f800000u;
        const float hi = open ? lo + 1.0f : __uint_as_float(hi_bits);
        const float xm = fmaf(0.5f * (lo + hi), jv, ja);
        const bool on = live && xm < 0.0f;
        float G1 = on ? Dr * ja * jv : 0.0f, G2 = on ? Dr * jv * jv : 0.0f;
        B2S_SUM8(G1); B2S_SUM8(G2);
        G1 += qg1; G2 += 2.0f * qg2;
        alpha = G2 > 0.0f ? __fdividef(-G1, G2) : lo;
        alpha = fmaxf(alpha, lo);
        if (!open) alpha = fminf(alpha, hi);
      }
    }
    // ---- step, new forces / active set / cost
    dirty = true;
    c = fmaf(alpha, d, c);
    ja = fmaf(alpha, jv, ja);
    gauss = gauss + alpha * qg1 + alpha * alpha * qg2;
    beta -= alpha * beta;
    prev_cost = cost;
    const unsigned act_before = __ballot_sync(FULL, act);
    act = live && ja < 0.0f;
    f = act ? -Dr * ja : 0.0f;
    rc = act ? 0.5f * Dr * ja * ja : 0.0f;
    B2S_SUM8(rc);
    cost = rc + gauss;
    niter++;
    if (iterations == 1) finished = true;
    // a full step that left the active set alone is on the minimiser; the polish below is the confirming iteration
    if (fabsf(alpha - 1.0f) < 1e-3f && act_before == __ballot_sync(FULL, act)) finished = true;
  }
  // ---- polish: one full step along the direction at the final point if it keeps the active set (see solve_dual)
  {
    const float c2 = c + d, ja2 = ja + jv;
    const bool act2 = live && ja2 < 0.0f;
    if (__ballot_sync(FULL, act2) == __ballot_sync(FULL, act)) {
      c = c2; ja = ja2; beta = 0.0f;
      f = act2 ? -Dr * ja2 : 0.0f;
    }
  }
#undef B2S_SUM8
  if (live) { w.du_c[r] = c; w.efc_force[r] = f; w.efc_active[r] = act ? 1 : 0; w.efc_Jaref[r] = ja; }
  __syncwarp();
  // back to dof space: qacc = a0 + beta delta0 + Z c, qfrc_constraint = J^T f
  LANE_FOR(i, nv) {
    float acc = w.qacc_smooth[i] + beta * w.du_d0[i];
    for (int s = 0; s < ne; s++) acc = fmaf(w.Z[s][i], w.du_c[s], acc);
    w.qacc[i] = acc; w.warm[i] = acc;
  }
  constraint_force_to_dofs(w, nv, ne, nsparse, LANE_PASS);
}
#endif

template <class D>
DEV void solve_dual(const Mdl& m, Work<D>& w, bool lockstep, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc, nsparse = w.nsparse;
  const bool warmstart = !m.h[MH_DISABLE_WARMSTART];
  const bool exact_ls = ne <= 16 && exact_ls_enabled(m);
  SUB_BEGIN();
  const float m00 = dual_setup(m, w, warmstart, LANE_PASS);
  SUB_MARK(w, 16);
  float beta = 0.0f, gauss = 0.0f;
  if (warmstart) {
    // cost at the warm start and at qacc_smooth, both in one pass over the rows
    float cw = 0.0f, cs = 0.0f;
    LANE_FOR(r, ne) {
      cw += constraint_row_cost(w, r, w.du_c0[r] + w.du_j0[r]);
      cs += constraint_row_cost(w, r, w.du_c0[r]);
    }
    cw = wsum(cw) + 0.5f * m00;
    cs = wsum(cs);
    if (cw < cs) { beta = 1.0f; gauss = 0.5f * m00; }
  }
  LANE_FOR(r, ne) { w.efc_Jaref[r] = w.du_c0[r] + beta * w.du_j0[r]; w.du_c[r] = 0.0f; }
  WSYNC();
  float cost = constraint_rows(w, ne, LANE_PASS) + gauss;
  float prev_cost = INFINITY;
  SUB_MARK(w, 17);
  DualDir dir = dual_direction(m, w, beta, m00, LANE_PASS);
  SUB_MARK(w, 18);
  const int iterations = m.h[MH_ITERATIONS];
  const float tol = m.mf[MF_TOLERANCE];
  const float scale = 1.0f / (m.mf[MF_MEANINERTIA] * (float)(nv > 1 ? nv : 1));
  int niter = 0;
  bool finished = false;
  while (true) {
    if (!finished && (iterations != 1 || niter > 0)) {
      const float improvement = (prev_cost - cost) * scale;
      finished = niter >= iterations;
      finished |= improvement < tol;
      finished |= dir.gradnorm * scale < tol;
    }
    if (lockstep ? !CTA_OR(!finished) : finished) break;
    const float qg[3] = {gauss, dir.qg1, dir.qg2};
    SUB_MARK(w, 19);
    // exact one-pass line search when two breakpoint slots per row fit a warp, else the bracketing iteration
    float alpha;
    if (exact_ls) alpha = finished ? 0.0f : ls_exact(w, ne, dir.qg1, dir.qg2, LANE_PASS);
    else alpha = ls_bracket(m, w, nv, ne, qg, dir.snorm, lockstep, !finished, LANE_PASS);
    SUB_MARK(w, 20);
    if (finished) continue;
    LANE_FOR(r, ne) { w.du_c[r] += alpha * w.du_d[r]; w.efc_Jaref[r] += alpha * w.efc_jv[r]; }
    WSYNC();
    gauss = gauss + alpha * dir.qg1 + alpha * alpha * dir.qg2;
    beta -= alpha * beta;
    prev_cost = cost;
    const unsigned act_before = active_mask(w, ne, LANE_PASS);
    cost = constraint_rows(w, ne, LANE_PASS) + gauss;
    SUB_MARK(w, 21);
    dir = dual_direction(m, w, beta, m00, LANE_PASS);
    SUB_MARK(w, 18);
    niter++;
    if (iterations == 1) finished = true;
    // A full Newton step (exact line search returned ~1) that left the active set alone has landed on the minimiser
    // of its quadratic piece, which for this convex cost is the minimum.  The reference would spend one more
    // iteration confirming it (a step of ~0 and an improvement below tolerance); the polish step below already is that
    // confirmation, taken along the direction just computed at the new point.
    if (exact_ls && ne <= 32 && fabsf(alpha - 1.0f) < 1e-3f && act_before == active_mask(w, ne, LANE_PASS)) finished = true;
  }
  // Polish: the loop above accepts a step only if it lowers the cost, and in fp32 the cost (O(1e5) with stiff rows)
  // stops resolving improvements while the iterate is still ~sqrt(eps) away from the minimiser.  `dir` is the Newton
  // direction at the final point; if the full step leaves the active set unchanged it is the exact minimiser of the
  // quadratic piece, so it is taken.  This brings the forces to the accuracy of the dof-space solver (tests/).
  {
    int changed = 0;
    LANE_FOR(r, ne) {
      w.du_c0[r] = w.efc_force[r];               // c0 / j0 / dinv are dead from here on: backups of f, active, c
      w.du_j0[r] = (float)w.efc_active[r];
      w.du_dinv[r] = w.du_c[r];
      w.du_c[r] += w.du_d[r];
      w.efc_Jaref[r] += w.efc_jv[r];
    }
    WSYNC();
    constraint_rows(w, ne, LANE_PASS);
    LANE_FOR(r, ne) changed |= (w.efc_active[r] != (int)w.du_j0[r]);
    changed = wany(changed);
    if (changed) {
      LANE_FOR(r, ne) { w.du_c[r] = w.du_dinv[r]; w.efc_force[r] = w.du_c0[r]; w.efc_active[r] = (int)w.du_j0[r]; }
    } else {
      beta = 0.0f;
    }
    WSYNC();
  }
  // back to dof space: qacc = a0 + beta delta0 + Z c, qfrc_constraint = J^T f
  LANE_FOR(i, nv) {
    float a = w.qacc_smooth[i] + beta * w.du_d0[i];
    for (int s = 0; s < ne; s++) a = fmaf(w.Z[s][i], w.du_c[s], a);
    w.qacc[i] = a; w.warm[i] = a;
  }
  constraint_force_to_dofs(w, nv, ne, nsparse, LANE_PASS);
}

template <class D>
DEV void solve(const Mdl& m, Work<D>& w, bool lockstep, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), ne = w.nefc;
  if (ne == 0) {
    // no constraint rows: qacc = qacc_smooth, qacc_warmstart keeps its old value (MJX forward())
    LANE_FOR(i, nv) { w.qacc[i] = w.qacc_smooth[i]; w.qfrc_constraint[i] = 0.0f; }
    WSYNC();
    if (lockstep) solve_follower(m);
    PROF_COUNT(w, 28);
    return;
  }
  // profile build: slots 28..31 count solves by path (none / register dual / shared dual / primal); 13, 14, 22 hold
  // the cycles of the primal, shared-memory dual and register dual paths
  SUB_BEGIN();
  if (dual_path(m, w, ne)) {
#if B2S_DEVICE
    if (ne <= 8 && w.scratch_i[0] == 0 && exact_ls_enabled(m)) {
      solve_dual8(m, w, lockstep, LANE_PASS);
      SUB_MARK(w, 22);
      PROF_COUNT(w, 29);
      return;
    }
#endif
    solve_dual(m, w, lockstep, LANE_PASS);
    SUB_MARK(w, 14);
    PROF_COUNT(w, 30);
  } else {
    if constexpr (D::NEFC > D::NDUAL) {
      if (!lockstep && newton_dof_enabled(m)) newton_dof(m, w, LANE_PASS);
      else solve_primal(m, w, lockstep, LANE_PASS);
    } else {
      solve_primal(m, w, lockstep, LANE_PASS);
    }
    SUB_MARK(w, 13);
    PROF_COUNT(w, 31);
  }
}

// ------------------------------------------------------------------------------------------------ sensors

template <class D>
DEV void object_frame(const Mdl& m, const Work<D>& w, int objtype, int objid, int* body, V3* pos, M9* rot) {
  if (objtype == OBJ_SITE) { *body = MI(m, MH_SITE_BODYID)[objid]; *pos = LD3(w.site_xpos, objid); *rot = ldm9(w.site_xmat, objid); }
  else if (objtype == OBJ_BODY) { *body = objid; *pos = LD3(w.xipos, objid); *rot = ldm9(w.ximat, objid); }
  else { *body = objid; *pos = LD3(w.xpos, objid); *rot = ldm9(w.xmat, objid); }
}

// mju_transformSpatial for a motion vector
DEV V6 transform_motion(const V6& vec, V3 newpos, V3 oldpos, const M9* rot) {
  V3 dif = newpos - oldpos;
  V3 ang = v3(vec.v[0], vec.v[1], vec.v[2]);
  V3 lin = v3(vec.v[3], vec.v[4], vec.v[5]) - cross(dif, ang);
  if (rot) { ang = mulTv(*rot, ang); lin = mulTv(*rot, lin); }
  V6 r;
  r.v[0] = ang.x; r.v[1] = ang.y; r.v[2] = ang.z; r.v[3] = lin.x; r.v[4] = lin.y; r.v[5] = lin.z;
  return r;
}

template <class D>
DEV void sensors_pos_vel(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int ns = m.h[MH_NSENSOR];
  const int* stype = MI(m, MH_SENSOR_TYPE);
  const int* sobjt = MI(m, MH_SENSOR_OBJTYPE);
  const int* sobj = MI(m, MH_SENSOR_OBJID);
  const int* sadr = MI(m, MH_SENSOR_ADR);
  const int* rootid = MI(m, MH_BODY_ROOTID);
  LANE_FOR(s, ns) {
    const int t = stype[s], ot = sobjt[s], id = sobj[s];
    float* out = w.sensordata + sadr[s];
    int body; V3 pos; M9 rot;
    if (t >= SENS_FRAMEPOS && t <= SENS_FRAMEZAXIS) {
      object_frame(m, w, ot, id, &body, &pos, &rot);
      if (t == SENS_FRAMEPOS) { out[0] = pos.x; out[1] = pos.y; out[2] = pos.z; }
      else if (t == SENS_FRAMEQUAT) {
        Q4 q = LDQ(w.xquat, body);
        if (ot == OBJ_SITE) {
          if constexpr (D::SMALL) q = qmul(q, ldgq(MF(m, MH_SITE_QUAT) + 4 * id)); else q = qmul(q, LDQ(w.site_quat, id));
        }
        else if (ot == OBJ_BODY) q = qmul(q, ldgq(MF(m, MH_BODY_IQUAT) + 4 * id));
        out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
      } else {
        const int col = t - SENS_FRAMEXAXIS;
        out[0] = rot.m[col]; out[1] = rot.m[3 + col]; out[2] = rot.m[6 + col];
      }
    } else if (t == SENS_MAGNETOMETER) {
      // site frame^T * opt.magnetic; the MJCF compiler only accepts MuJoCo's default field (0, -0.5, 0)
      const M9 r = ldm9(w.site_xmat, id);
      const V3 o = mulTv(r, v3(0.0f, -0.5f, 0.0f));
      out[0] = o.x; out[1] = o.y; out[2] = o.z;
    } else if (t == SENS_VELOCIMETER || t == SENS_GYRO || t == SENS_FRAMELINVEL || t == SENS_FRAMEANGVEL) {
      const bool local = (t == SENS_VELOCIMETER || t == SENS_GYRO);
      object_frame(m, w, local ? OBJ_SITE : ot, id, &body, &pos, &rot);
      V6 v = transform_motion(ld6(w.cvel, body), pos, LD3(w.subtree_com, rootid[body]), local ? &rot : nullptr);
      const int o = (t == SENS_VELOCIMETER || t == SENS_FRAMELINVEL) ? 3 : 0;
      out[0] = v.v[o]; out[1] = v.v[o + 1]; out[2] = v.v[o + 2];
    }
  }
  WSYNC();
}

template <class D>
DEV V3 contact_force_world(const Mdl& m, const Work<D>& w, int c) {
  const int a = w.con_efc[c];
  const int p = w.con_pair[c];
  float f0 = 0, f1 = 0, f2 = 0;
  if (a >= 0) {
    if (MI(m, MH_PAIR_DIM)[p] == 1) f0 = w.efc_force[a];
    else {
      const float* fr = MF(m, MH_PAIR_FRICTION) + 5 * p;
      f0 = w.efc_force[a] + w.efc_force[a + 1] + w.efc_force[a + 2] + w.efc_force[a + 3];
      f1 = (w.efc_force[a] - w.efc_force[a + 1]) * fr[0];
      f2 = (w.efc_force[a + 2] - w.efc_force[a + 3]) * fr[1];
    }
  }
  return v3(w.con_frame[0][c] * f0 + w.con_frame[3][c] * f1 + w.con_frame[6][c] * f2,
            w.con_frame[1][c] * f0 + w.con_frame[4][c] * f1 + w.con_frame[7][c] * f2,
            w.con_frame[2][c] * f0 + w.con_frame[5][c] * f1 + w.con_frame[8][c] * f2);
}

// rne_postconstraint + acceleration-stage sensors
template <class D>
DEV void sensors_acc(const Mdl& m, Work<D>& w, LANE_ARG) {
  const int nb = DIM(NB, MH_NBODY), ncon = DIM(NCON, MH_NCON), ns = m.h[MH_NSENSOR];
  const int* rootid = MI(m, MH_BODY_ROOTID);
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  const int* pg2 = MI(m, MH_PAIR_GEOM2);
  const int* gbody = MI(m, MH_GEOM_BODYID);
  const int* sub_end = MI(m, MH_BODY_SUBTREE_END);
  rne_cacc(m, w, true, LANE_PASS);
  LANE_FOR(b, nb) {
    V6 e;
#pragma unroll
    for (int c = 0; c < 6; c++) e.v[c] = 0.0f;
    if (b > 0) {
      V3 com = LD3(w.subtree_com, rootid[b]);
      if (b == w.xfrc_body) {
        V3 f = v3(w.xfrc[0], w.xfrc[1], w.xfrc[2]);
        V3 cr = cross(LD3(w.xipos, b) - com, f);
        e.v[0] = w.xfrc[3] + cr.x; e.v[1] = w.xfrc[4] + cr.y; e.v[2] = w.xfrc[5] + cr.z;
        e.v[3] = f.x; e.v[4] = f.y; e.v[5] = f.z;
      }
      for (int c = 0; c < ncon; c++) {
        if (w.con_efc[c] < 0) continue;
        const int p = w.con_pair[c];
        const int b1 = gbody[pg1[p]], b2 = gbody[pg2[p]];
        if (b1 != b && b2 != b) continue;
        V3 fw = contact_force_world(m, w, c);
        V3 cr = cross(LD3(w.con_pos, c) - com, fw);
        const float sg = (b2 == b) ? 1.0f : -1.0f;
        e.v[0] += sg * cr.x; e.v[1] += sg * cr.y; e.v[2] += sg * cr.z;
        e.v[3] += sg * fw.x; e.v[4] += sg * fw.y; e.v[5] += sg * fw.z;
      }
    }
    st6(w.cfrc_ext, b, e);
    V6 f;
    if (b == 0) {
#pragma unroll
      for (int c = 0; c < 6; c++) f.v[c] = 0.0f;
    } else {
      f = body_inertial_force(w, b);
#pragma unroll
      for (int c = 0; c < 6; c++) f.v[c] -= e.v[c];
    }
    st6(w.cfrc, b, f);
  }
  WSYNC();
  const int* stype = MI(m, MH_SENSOR_TYPE);
  const int* sobjt = MI(m, MH_SENSOR_OBJTYPE);
  const int* sobj = MI(m, MH_SENSOR_OBJID);
  const int* sadr = MI(m, MH_SENSOR_ADR);
  LANE_FOR(s, ns) {
    const int t = stype[s], ot = sobjt[s], id = sobj[s];
    float* out = w.sensordata + sadr[s];
    if (t == SENS_ACCEL || t == SENS_FRAMELINACC || t == SENS_FRAMEANGACC) {
      const bool local = (t == SENS_ACCEL);
      int body; V3 pos; M9 rot;
      object_frame(m, w, local ? OBJ_SITE : ot, id, &body, &pos, &rot);
      V3 com = LD3(w.subtree_com, rootid[body]);
      V6 vel = transform_motion(ld6(w.cvel, body), pos, com, local ? &rot : nullptr);
      V6 acc = transform_motion(ld6(w.cacc, body), pos, com, local ? &rot : nullptr);
      V3 cr = cross(v3(vel.v[0], vel.v[1], vel.v[2]), v3(vel.v[3], vel.v[4], vel.v[5]));
      acc.v[3] += cr.x; acc.v[4] += cr.y; acc.v[5] += cr.z;
      const int o = (t == SENS_FRAMEANGACC) ? 0 : 3;
      out[0] = acc.v[o]; out[1] = acc.v[o + 1]; out[2] = acc.v[o + 2];
    } else if (t == SENS_FORCE || t == SENS_TORQUE) {
      // interaction wrench between the site's body and its parent: subtree sum of cfrc (world excluded)
      const int b = MI(m, MH_SITE_BODYID)[id];
      V6 wv;
#pragma unroll
      for (int c = 0; c < 6; c++) wv.v[c] = 0.0f;
      for (int cb = b; cb < sub_end[b]; cb++)
#pragma unroll
        for (int c = 0; c < 6; c++) wv.v[c] += w.cfrc[c][cb];
      V3 com = LD3(w.subtree_com, rootid[b]);
      V3 dif = LD3(w.site_xpos, id) - com;
      V3 frc = v3(wv.v[3], wv.v[4], wv.v[5]);
      V3 tq = v3(wv.v[0], wv.v[1], wv.v[2]) - cross(dif, frc);
      M9 rot = ldm9(w.site_xmat, id);
      V3 o = mulTv(rot, t == SENS_FORCE ? frc : tq);
      out[0] = o.x; out[1] = o.y; out[2] = o.z;
    } else if (t == SENS_TOUCH) {
      const int b = MI(m, MH_SITE_BODYID)[id];
      const float* sz = MF(m, MH_SITE_SIZE) + 3 * id;
      const int site_type = MI(m, MH_SITE_TYPE)[id];
      M9 rot = ldm9(w.site_xmat, id);
      V3 sp = LD3(w.site_xpos, id);
      float total = 0.0f;
      for (int c = 0; c < ncon; c++) {
        const int a = w.con_efc[c];
        if (a < 0) continue;
        const int p = w.con_pair[c];
        const int b1 = gbody[pg1[p]], b2 = gbody[pg2[p]];
        if (b1 != b && b2 != b) continue;
        V3 loc = mulTv(rot, LD3(w.con_pos, c) - sp);
        bool inside;
        if (site_type == 6) inside = fabsf(loc.x) <= sz[0] && fabsf(loc.y) <= sz[1] && fabsf(loc.z) <= sz[2];
        else inside = dot(loc, loc) <= sz[0] * sz[0];
        if (!inside) continue;
        const int nr = MI(m, MH_PAIR_DIM)[p] == 1 ? 1 : 4;
        for (int k = 0; k < nr; k++) total += w.efc_force[a + k];
      }
      out[0] = total;
    }
  }
  WSYNC();
}

// --------------------------------------------------------------------------------------------- top level

// `sensors`: also evaluate sensordata / cfrc_ext.  They are pure outputs (recomputed from scratch by every
// forward pass and read only by observations), so the control step asks for them on its last sub-step only.
template <class D>
DEVNI void forward(const Mdl& m, Work<D>& w, bool sensors, bool lockstep, LANE_ARG) {
  const int nv = DIM(NV, MH_NV);
  // `lockstep` (sub-steps of a control step, never the divergent reset path): CTA barriers between stages keep
  // the CTA's warps on the same code, so one instruction-cache fill serves all of them.
#define B2S_STAGE_SYNC(k) do { if (lockstep && ((m.sync_mask >> (k)) & 1)) CTA_SYNC(); } while (0)
  STAGE_BEGIN();
  DBG_STOP(m, 20);
  kinematics(m, w, LANE_PASS);
  DBG_STOP(m, 21);
  B2S_STAGE_SYNC(0);
  STAGE_MARK(w, 0);
  com_pos(m, w, LANE_PASS);
  DBG_STOP(m, 22);
  crb_mass_matrix(m, w, LANE_PASS);
  DBG_STOP(m, 23);
  B2S_STAGE_SYNC(1);
  STAGE_MARK(w, 1);
  collision(m, w, LANE_PASS);
  DBG_STOP(m, 24);
  com_vel(m, w, LANE_PASS);
  DBG_STOP(m, 25);
  make_constraint(m, w, LANE_PASS);
  DBG_STOP(m, 26);
  B2S_STAGE_SYNC(2);
  STAGE_MARK(w, 2);
  passive(m, w, LANE_PASS);
  DBG_STOP(m, 27);
  rne(m, w, LANE_PASS);
  DBG_STOP(m, 28);
  if (sensors) sensors_pos_vel(m, w, LANE_PASS);
  actuation(m, w, LANE_PASS);
  DBG_STOP(m, 29);
  B2S_STAGE_SYNC(3);
  STAGE_MARK(w, 3);
  acceleration(m, w, LANE_PASS);
  DBG_STOP(m, 30);
  B2S_STAGE_SYNC(4);
  STAGE_MARK(w, 4);
  solve(m, w, lockstep && ((m.sync_mask >> 7) & 1), LANE_PASS);
  STAGE_MARK(w, 5);
  DBG_STOP(m, 31);
  B2S_STAGE_SYNC(5);
  STAGE_MARK(w, 6);
#undef B2S_STAGE_SYNC
  if (sensors) sensors_acc(m, w, LANE_PASS);
  STAGE_MARK(w, 7);
}

// the CTA barriers / votes of one lockstep forward(), for a warp that has no environment (tail of the last CTA)
DEV void forward_follower(const Mdl& m) {
#pragma unroll
  for (int k = 0; k < 5; k++) if ((m.sync_mask >> k) & 1) CTA_SYNC();
  if ((m.sync_mask >> 7) & 1) solve_follower(m);
  if ((m.sync_mask >> 5) & 1) CTA_SYNC();
}

// implicitfast with Euler damping disabled: (M + dt*diag(damping)) qacc = qfrc_smooth + qfrc_constraint
template <class D>
DEV void physics_step(const Mdl& m, Work<D>& w, bool sensors, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), nj = DIM(NJ, MH_NJNT);
  const float dt = m.mf[MF_TIMESTEP];
  forward(m, w, sensors, true, LANE_PASS);
  STAGE_BEGIN();
  LANE_FOR(i, nv) w.s_grad[i] = w.qfrc_smooth[i] + w.qfrc_constraint[i];
  WSYNC();
  factor_solve(m, w, FS_IMPLICIT, w.s_Mgrad, w.s_grad, 0, LANE_PASS);
  LANE_FOR(i, nv) w.qvel[i] += w.s_Mgrad[i] * dt;
  WSYNC();
  const int* jtype = MI(m, MH_JNT_TYPE);
  const int* jqpos = MI(m, MH_JNT_QPOSADR);
  const int* jdof = MI(m, MH_JNT_DOFADR);
  LANE_FOR(j, nj) {
    const int qa = jqpos[j], da = jdof[j];
    if (jtype[j] == JNT_FREE) {
#pragma unroll
      for (int k = 0; k < 3; k++) w.qpos[qa + k] += dt * w.qvel[da + k];
      V3 wv = v3(w.qvel[da + 3], w.qvel[da + 4], w.qvel[da + 5]);
      const float n = sqrtf(dot(wv, wv));
      const float inv = 1.0f / (n + (n == 0 ? MINVAL : 0.0f));
      Q4 qr = axis_angle(wv * inv, dt * n);
      Q4 q; q.w = w.qpos[qa + 3]; q.x = w.qpos[qa + 4]; q.y = w.qpos[qa + 5]; q.z = w.qpos[qa + 6];
      q = qnormalize(qmul(q, qr));
      w.qpos[qa + 3] = q.w; w.qpos[qa + 4] = q.x; w.qpos[qa + 5] = q.y; w.qpos[qa + 6] = q.z;
    } else {
      w.qpos[qa] += dt * w.qvel[da];
    }
  }
  IF_LANE0 w.time += dt;
  WSYNC();
  STAGE_MARK(w, 8);
}

}  // namespace b2s
