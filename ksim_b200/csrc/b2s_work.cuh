// b200sim: compile-time size traits and the per-warp shared-memory workspace.
#pragma once

#include "b2s_common.cuh"

namespace b2s {

constexpr int dims_min(int a, int b) { return a < b ? a : b; }
// largest n with n rows of odd stride (n | 1) fitting in `cap` floats
constexpr int dims_asq(int n, int cap) { return n * (n | 1) <= cap ? n : dims_asq(n - 1, cap); }

// Size traits.  A kernel instantiation serves every model whose sizes are <= the traits' (exact match is
// fastest because lane loops collapse to a single predicated pass).
template <int NQ_, int NV_, int NB_, int NJ_, int NU_, int NS_, int NSD_, int NPAIR_, int NCON_, int NDENSE_, int NLIMIT_,
          int NFRIC_, bool EXACT_, bool SMALL_ = EXACT_>
struct Dims {
  static constexpr bool EXACT = EXACT_;       // model sizes equal the traits: lane loops get compile-time bounds
  static constexpr bool SMALL = SMALL_;       // task budgets of the walking example, no per-env pair geoms (fills the SM with 14 workspaces)
  static constexpr int NQ = NQ_, NV = NV_, NB = NB_, NJ = NJ_, NU = NU_, NS = NS_, NSD = NSD_;
  static constexpr int NPAIR = NPAIR_, NCON = NCON_, NDENSE = NDENSE_, NLIMIT = NLIMIT_, NFRIC = NFRIC_;
  // action size: PositionVelocityActuator doubles it; the exact-fit instantiation serves nu-sized actions only
  static constexpr int NA = EXACT_ ? NU_ : 2 * NU_;
  static constexpr int NVP = NV_ | 1;         // odd row stride: conflict-free column access
  static constexpr int NSPARSE = NFRIC_ + NLIMIT_;
  static constexpr int NEFC = NFRIC_ + NLIMIT_ + NDENSE_;
  // task-state budgets: the SMALL instantiation is sized for the walking task (2 event floats, 8 command floats)
  // so that the model AND the task program fit in shared memory next to 14 workspaces
  static constexpr int EVS = SMALL_ ? 4 : 16;  // event state floats
  static constexpr int ACS = 2 * NU_ + 3;     // PositionActuators state
  static constexpr int CMDS = SMALL_ ? 8 : 24;  // command floats
  // observation carry floats: two projected-gravity observations (14) in the exact-fit instantiations; the generic one also
  // has room for the delayed joint position / velocity rings (delay_steps x joints each)
  static constexpr int OCS = EXACT_ ? 16 : 16 + 4 * NU_;
  static constexpr int NOBS = SMALL_ ? 16 : 32;  // observation plugins (bias keys)
  // Dual (constraint-space) Newton path: at most NDUAL rows.  Z = M^-1 J^T lives in the storage of
  // crb/cacc/cfrc/cfrc_ext (28*NB floats, dead during the solve); A = J Z and the factor of the reduced Hessian
  // share the storage of W.  Rows beyond NDUAL fall back to the primal (dof-space) Newton path.
  static constexpr int NDUAL = dims_min(dims_min(NEFC, 32), dims_min((28 * NB_) / (NV_ | 1), dims_asq(NV_, NV_ * (NV_ | 1))));
  static constexpr int NDS = NDUAL | 1;       // odd row stride of AS
};

// examples/data/scene.mjcf (walking.py): nq 24 nv 23 nbody 17 njnt 18 nu 17 nsite 3 nsensordata 54,
// 2 sphere-plane pairs condim 1, 17 limited hinges, no frictionloss.
using DimsHumanoid = Dims<24, 23, 17, 18, 17, 3, 54, 2, 2, 2, 17, 0, true>;
// examples/kbot/robot/kbot-headless/robot.mjcf: nq 27 nv 26 nbody 24 nu 20, 4 capsule-plane pairs condim 3
// (8 contacts x 4 pyramid rows), 20 limited hinges, 20 frictionloss dofs.
// Exact-fit as well (compile-time loop bounds), with the larger task budgets and per-env pair geoms of its task.
using DimsKbot = Dims<27, 26, 24, 21, 20, 4, 56, 4, 8, 32, 20, 20, true, false>;
// anything else that fits one dof / body per lane
using DimsGeneric = Dims<33, 32, 32, 32, 32, 8, 96, 8, 16, 64, 32, 32, false>;

// model dimension: compile-time when the traits match the model exactly
#define DIM(X, H) (D::EXACT ? (int)D::X : m.h[H])

// geom_pos / geom_size of the two geoms of every contact pair, per environment (CollisionBodyRandomizer,
// randomization.py:380-442).  The humanoid (SMALL) instantiation has no shared memory to spare (its workspaces, the model and the
// task program fill the SM) and its task has no such override: it reads the shared model instead and carries nothing.
template <class D, bool SMALL>
struct PairGeom {
  float pair_gpos[2][3][D::NPAIR], pair_gsize[2][3][D::NPAIR];
  float site_pos[3][D::NS], site_quat[4][D::NS];  // per-env site frames (IMUAlignmentRandomizer, randomization.py:245-283)
};
template <class D>
struct PairGeom<D, true> {};

template <class D>
struct Work : PairGeom<D, D::SMALL> {
  // ---- persistent per-env state (mirrors the HBM state block) ----
  float qpos[D::NQ], qvel[D::NV], qacc[D::NV], warm[D::NV], ctrl[D::NU];
  float last_ctrl[D::NA], zero_off[D::NU], event[D::EVS], act_state[D::ACS], cmd[D::CMDS], obs_carry[D::OCS];
  float time, latency, level, nonfinite;
  uint32_t keys[2 + 2 * D::NOBS];
  // ---- per-env model overrides (ksim/randomization.py); filled from the base model then the rand block ----
  float armature[D::NV], damping[D::NV], frictionloss[D::NV];
  float body_mass[D::NB], body_inertia[3][D::NB], body_ipos[3][D::NB];
  // ---- position stage ----
  float xpos[3][D::NB], xquat[4][D::NB], xmat[9][D::NB], xipos[3][D::NB], ximat[9][D::NB];
  float subtree_com[3][D::NB], cinert[10][D::NB];
  float xanchor[3][D::NJ], xaxis[3][D::NJ];
  float cdof[6][D::NV], cdof_dot[6][D::NV];
  float site_xpos[3][D::NS], site_xmat[9][D::NS];
  float M[D::NV][D::NVP];                     // mass matrix
  union {
    float W[D::NV][D::NVP];                   // work matrix (Cholesky factors)
    // dual Newton path: AS[r][s], s >= r: A = J M^-1 J^T (upper triangle); s < r: factor of the reduced Hessian
    float AS[D::NDUAL][D::NDS];
  };
#if !B2S_DEVICE
  float Linv_diag[D::NV];                     // 1 / diag of the factor in W (host build's loop Cholesky only)
#endif
  // ---- contacts ----
  float con_dist[D::NCON], con_pos[3][D::NCON], con_frame[9][D::NCON];
  int con_pair[D::NCON], con_efc[D::NCON];
  // ---- velocity / force stage ----
  float cvel[6][D::NB];
  union {
    struct { float crb[10][D::NB], cacc[6][D::NB], cfrc[6][D::NB], cfrc_ext[6][D::NB]; };
    float Z[D::NDUAL][D::NVP];                // dual Newton path: row s = M^-1 J_s^T (alive from acceleration() to the end of solve())
  };
  float xfrc[6];                              // Cartesian wrench of the ForcePush event (one body)
  int xfrc_body;
  float qfrc_passive[D::NV], qfrc_bias[D::NV], qfrc_actuator[D::NV], qfrc_smooth[D::NV], qacc_smooth[D::NV];
  float qfrc_constraint[D::NV], actuator_force[D::NU];
  // ---- constraints ----
  int nefc, nsparse;                          // rows [0,nsparse) are single-dof rows, [nsparse,nefc) dense
  int efc_type[D::NEFC], efc_dof[D::NEFC];
  float efc_sign[D::NEFC], efc_aref[D::NEFC], efc_D[D::NEFC], efc_R[D::NEFC], efc_floss[D::NEFC];
  float efc_force[D::NEFC], efc_Jaref[D::NEFC], efc_jv[D::NEFC];
  int efc_active[D::NEFC];
  float Jd[D::NDENSE][D::NVP];
  // ---- solver vectors ----
  union {
    struct { float s_qacc[D::NV], s_Ma[D::NV], s_grad[D::NV], s_Mgrad[D::NV], s_search[D::NV], s_mv[D::NV], s_tmp[D::NV]; };
    // dual Newton path: delta0 = warmstart - qacc_smooth and M delta0 (dof space); per-row c0 = J qacc_smooth - aref,
    // j0 = J delta0, c (coordinates of qacc - qacc_smooth - beta delta0 in the columns of Z), d (search direction), 1/diag
    struct { float du_d0[D::NV], du_Md0[D::NV], du_c0[D::NDUAL], du_j0[D::NDUAL], du_c[D::NDUAL], du_d[D::NDUAL], du_dinv[D::NDUAL]; };
  };
  float sensordata[D::NSD];
  float torque[D::NU], action[D::NA], ctrl_blend[D::NA];
  unsigned drop_mask[(D::NA + 31) / 32];      // per-element action drop of this control step, one bit each
  int scratch_i[4];                           // [0] friction rows of this sub-step, [1..3] spare
#if defined(B2S_PROFILE_STAGES) && B2S_DEVICE
  unsigned prof[B2S_NSTAGE];
#endif
};

}  // namespace b2s
