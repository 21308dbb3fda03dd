/*
 * CPU oracle for the batched-rollout hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C, one-environment-at-a-time restatement of the path the reference runs through
 * ksim/engine.py:205-361 (MjxEngine.reset/step), ksim/task/rl.py:153-331,1023-1211 (observations,
 * terminations, commands, reset-on-done), ksim/task/rl.py:189-245 (rewards) and ksim/task/ppo.py:54-121
 * (GAE).  The arithmetic of `mjx.step` / `mjx.forward` lives in the un-vendored PyPI packages
 * `mujoco-mjx` / `mujoco` (ksim/requirements.txt:19-21, no version pin; absent from /root/reference and
 * from this image), so the physics below restates MuJoCo's published computation pipeline
 * (kinematics -> comPos -> crb -> factorM -> collision -> makeConstraint -> comVel -> passive -> rne ->
 * actuation -> acceleration -> Newton solve -> sensors -> implicitfast) from its documentation.
 *
 * PARITY UNPINNED for physics: no reference test pins a physics result and MJX cannot
 * be run here; the oracle is pinned only by analytic cases (tests/test_oracle_physics.py) and by the
 * reference's own known-answer vectors for rewards / terminations / commands (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product path (ksim_b200/csrc) never links or calls it.
 *
 * Build twice: -DREAL=double (liboracle_f64.so) and -DREAL=float (liboracle_f32.so).
 */
#ifndef B2S_ORACLE_H
#define B2S_ORACLE_H

#include <stdint.h>

#ifndef REAL
#define REAL double
#endif

#define O_MAXNV 32
#define O_MAXNQ 33
#define O_MAXNB 32
#define O_MAXNJ 32
#define O_MAXNG 64
#define O_MAXNS 8
#define O_MAXNU 32
#define O_MAXSENS 32
#define O_MAXSD 96
#define O_MAXPAIR 8
#define O_MAXCON 16
#define O_MAXEFC 128

#define O_MINVAL ((REAL)1e-15)
#define O_MINIMP ((REAL)0.0001)
#define O_MAXIMP ((REAL)0.9999)

enum { O_JNT_FREE = 0, O_JNT_HINGE = 3 };
enum { O_GEOM_PLANE = 0, O_GEOM_SPHERE = 2, O_GEOM_CAPSULE = 3 };
enum { O_EFC_FRICTION = 1, O_EFC_LIMIT = 2, O_EFC_CONTACT = 3 };
enum {
  O_SENS_TOUCH = 0, O_SENS_ACCEL = 1, O_SENS_VELOCIMETER = 2, O_SENS_GYRO = 3, O_SENS_FORCE = 4,
  O_SENS_TORQUE = 5, O_SENS_MAGNETOMETER = 6, O_SENS_FRAMEPOS = 10, O_SENS_FRAMEQUAT = 11, O_SENS_FRAMEXAXIS = 12,
  O_SENS_FRAMEYAXIS = 13, O_SENS_FRAMEZAXIS = 14, O_SENS_FRAMELINVEL = 15, O_SENS_FRAMEANGVEL = 16,
  O_SENS_FRAMELINACC = 17, O_SENS_FRAMEANGACC = 18
};
enum { O_OBJ_BODY = 1, O_OBJ_XBODY = 2, O_OBJ_SITE = 6 };

typedef struct {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nsensordata, npair;
  /* options */
  REAL timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  int iterations, ls_iterations, disable_refsafe, disable_warmstart;
  /* bodies */
  int body_parentid[O_MAXNB], body_rootid[O_MAXNB], body_jntnum[O_MAXNB], body_jntadr[O_MAXNB];
  int body_dofnum[O_MAXNB], body_dofadr[O_MAXNB];
  REAL body_pos[O_MAXNB][3], body_quat[O_MAXNB][4], body_ipos[O_MAXNB][3], body_iquat[O_MAXNB][4];
  REAL body_mass[O_MAXNB], body_inertia[O_MAXNB][3], body_subtreemass[O_MAXNB], body_invweight0[O_MAXNB][2];
  /* joints / dofs */
  int jnt_type[O_MAXNJ], jnt_qposadr[O_MAXNJ], jnt_dofadr[O_MAXNJ], jnt_bodyid[O_MAXNJ], jnt_limited[O_MAXNJ];
  REAL jnt_pos[O_MAXNJ][3], jnt_axis[O_MAXNJ][3], jnt_stiffness[O_MAXNJ], jnt_range[O_MAXNJ][2];
  REAL jnt_margin[O_MAXNJ], jnt_solref[O_MAXNJ][2], jnt_solimp[O_MAXNJ][5];
  int dof_bodyid[O_MAXNV], dof_jntid[O_MAXNV], dof_parentid[O_MAXNV];
  REAL dof_armature[O_MAXNV], dof_damping[O_MAXNV], dof_frictionloss[O_MAXNV], dof_invweight0[O_MAXNV];
  REAL dof_solref[O_MAXNV][2], dof_solimp[O_MAXNV][5];
  REAL qpos0[O_MAXNQ], qpos_spring[O_MAXNQ];
  /* geoms, sites */
  int geom_type[O_MAXNG], geom_bodyid[O_MAXNG];
  REAL geom_pos[O_MAXNG][3], geom_quat[O_MAXNG][4], geom_size[O_MAXNG][3];
  int site_bodyid[O_MAXNS];
  REAL site_pos[O_MAXNS][3], site_quat[O_MAXNS][4], site_size[O_MAXNS][3];
  int site_type[O_MAXNS];
  /* collision pairs */
  int pair_geom1[O_MAXPAIR], pair_geom2[O_MAXPAIR], pair_dim[O_MAXPAIR];
  REAL pair_friction[O_MAXPAIR][5], pair_solref[O_MAXPAIR][2], pair_solimp[O_MAXPAIR][5];
  REAL pair_margin[O_MAXPAIR], pair_gap[O_MAXPAIR];
  /* actuators (motors on hinge joints) */
  int actuator_trnid[O_MAXNU], actuator_ctrllimited[O_MAXNU], actuator_forcelimited[O_MAXNU];
  REAL actuator_gear[O_MAXNU], actuator_ctrlrange[O_MAXNU][2], actuator_forcerange[O_MAXNU][2];
  /* sensors */
  int sensor_type[O_MAXSENS], sensor_objtype[O_MAXSENS], sensor_objid[O_MAXSENS], sensor_dim[O_MAXSENS];
  int sensor_adr[O_MAXSENS];
} OModel;

typedef struct {
  int geom1, geom2, dim, efc_address, pair;
  REAL dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5];
} OContact;

typedef struct {
  /* state */
  REAL qpos[O_MAXNQ], qvel[O_MAXNV], qacc[O_MAXNV], qacc_warmstart[O_MAXNV], ctrl[O_MAXNU], time;
  REAL xfrc_applied[O_MAXNB][6];
  /* position-dependent */
  REAL xpos[O_MAXNB][3], xquat[O_MAXNB][4], xmat[O_MAXNB][9], xipos[O_MAXNB][3], ximat[O_MAXNB][9];
  REAL xanchor[O_MAXNJ][3], xaxis[O_MAXNJ][3];
  REAL geom_xpos[O_MAXNG][3], geom_xmat[O_MAXNG][9], site_xpos[O_MAXNS][3], site_xmat[O_MAXNS][9];
  REAL subtree_com[O_MAXNB][3], cinert[O_MAXNB][10], crb[O_MAXNB][10], cdof[O_MAXNV][6];
  REAL qM[O_MAXNV][O_MAXNV], qL[O_MAXNV][O_MAXNV];
  int ncon, nefc;
  OContact contact[O_MAXCON];
  /* velocity-dependent */
  REAL cvel[O_MAXNB][6], cdof_dot[O_MAXNV][6];
  REAL qfrc_passive[O_MAXNV], qfrc_bias[O_MAXNV], qfrc_actuator[O_MAXNV], actuator_force[O_MAXNU];
  REAL qfrc_smooth[O_MAXNV], qacc_smooth[O_MAXNV], qfrc_constraint[O_MAXNV];
  /* constraints */
  int efc_type[O_MAXEFC], efc_id[O_MAXEFC];
  REAL efc_J[O_MAXEFC][O_MAXNV], efc_pos[O_MAXEFC], efc_aref[O_MAXEFC], efc_D[O_MAXEFC], efc_R[O_MAXEFC];
  REAL efc_frictionloss[O_MAXEFC], efc_force[O_MAXEFC];
  int solver_niter;
  /* acceleration-dependent */
  REAL cacc[O_MAXNB][6], cfrc_int[O_MAXNB][6], cfrc_ext[O_MAXNB][6];
  REAL sensordata[O_MAXSD];
} OData;

/* physics.c */
void o_make_data(const OModel* m, OData* d);
void o_forward(const OModel* m, OData* d);
void o_step(const OModel* m, OData* d);
void o_kinematics_only(const OModel* m, OData* d); /* kinematics + comPos + crb, for tests */

/* math helpers shared inside the oracle */
void o_quat_mul(REAL* r, const REAL* a, const REAL* b);
void o_quat_to_mat(REAL* m, const REAL* q);
void o_rot_vec_quat(REAL* r, const REAL* v, const REAL* q, int inverse);
void o_euler_to_quat(REAL* q, const REAL* rpy);

#endif
