/*
 * Oracle engine: ksim's control-step semantics around the oracle physics.  TEST INFRASTRUCTURE (oracle.h).
 *
 * Follows, per environment and in scalar code:
 *   - RNG: jax.random threefry2x32 (split / uniform / normal / bernoulli), partitionable counter layout
 *     [restated from memory of the un-vendored `jax` package; pinned by the Random123 known-answer
 *     vectors in tests/test_rng.py];
 *   - MjxEngine.reset / MjxEngine.step                  ksim/engine.py:205-361
 *   - PositionActuators / TorqueActuators               ksim/actuators.py:70-238
 *   - events, resets, noise, scales                     ksim/events.py, resets.py, noise.py, scales.py
 *   - get_observation + Observation.add_noise           ksim/task/rl.py:153-186, ksim/observation.py:110-142
 *   - get_terminations, done/success                    ksim/task/rl.py:288-299,1042-1050
 *   - reset-on-done, commands, obs carry / bias keys    ksim/task/rl.py:1072-1123
 *   - get_rewards                                       ksim/task/rl.py:189-245, ksim/rewards.py
 *   - compute_ppo_inputs (GAE)                          ksim/task/ppo.py:54-121
 *
 * Buffers are env-major (buf[env][size]) with the layouts the host writes into ti[] (include/b200sim_codes.h).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200sim_codes.h"
#include "oracle.h"

#if defined(ORACLE_F32)
#define SQRT sqrtf
#define SIN sinf
#define FLOOR floorf
#define CEIL ceilf
#define COS cosf
#define FABS fabsf
#define EXP expf
#define LOG logf
#define ATAN2 atan2f
#define ACOS acosf
#define ASIN asinf
#define RINT rintf
#else
#define SQRT sqrt
#define SIN sin
#define FLOOR floor
#define CEIL ceil
#define COS cos
#define FABS fabs
#define EXP exp
#define LOG log
#define ATAN2 atan2
#define ACOS acos
#define ASIN asin
#define RINT rint
#endif

#define PI_R ((REAL)3.14159265358979323846)

/* ------------------------------------------------------------------------------------------------ RNG */

typedef struct { uint32_t k0, k1; } Key;

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void o_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0, uint32_t* o1) {
  static const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0]; x1 += ks[1];
  for (int s = 0; s < 5; s++) {
    const int* r = (s % 2 == 0) ? R : R + 4;
    for (int i = 0; i < 4; i++) { x0 += x1; x1 = rotl32(x1, r[i]); x1 ^= x0; }
    x0 += ks[(s + 1) % 3];
    x1 += ks[(s + 2) % 3] + (uint32_t)(s + 1);
  }
  *o0 = x0; *o1 = x1;
}

/* jax.random.split(key, n)[i] with jax_threefry_partitionable: threefry(key, (hi=0, lo=i)) */
static Key key_split(Key k, int i) {
  Key o;
  o_threefry2x32(k.k0, k.k1, 0u, (uint32_t)i, &o.k0, &o.k1);
  return o;
}
/* jax.random.bits(key, shape)[i] (32-bit): xor of the two output words */
static uint32_t key_bits(Key k, int i) {
  uint32_t a, b;
  o_threefry2x32(k.k0, k.k1, 0u, (uint32_t)i, &a, &b);
  return a ^ b;
}
static float bits_to_unit(uint32_t bits) {
  union { uint32_t u; float f; } c;
  c.u = (bits >> 9) | 0x3F800000u;
  return c.f - 1.0f;
}
/* jax.random.uniform(key, shape, minval, maxval)[i] */
static REAL key_uniform(Key k, int i, REAL lo, REAL hi) {
  REAL u = (REAL)bits_to_unit(key_bits(k, i));
  REAL v = u * (hi - lo) + lo;
  return v > lo ? v : lo;
}
static REAL erfinv_f32poly(REAL x) {
  /* Giles' single-precision polynomial, the form XLA uses for f32 erf_inv */
  REAL w = -LOG(((REAL)1 - x) * ((REAL)1 + x)), p;
  if (w < (REAL)5) {
    w = w - (REAL)2.5;
    p = (REAL)2.81022636e-08;
    p = (REAL)3.43273939e-07 + p * w;
    p = (REAL)-3.5233877e-06 + p * w;
    p = (REAL)-4.39150654e-06 + p * w;
    p = (REAL)0.00021858087 + p * w;
    p = (REAL)-0.00125372503 + p * w;
    p = (REAL)-0.00417768164 + p * w;
    p = (REAL)0.246640727 + p * w;
    p = (REAL)1.50140941 + p * w;
  } else {
    w = SQRT(w) - (REAL)3;
    p = (REAL)-0.000200214257;
    p = (REAL)0.000100950558 + p * w;
    p = (REAL)0.00134934322 + p * w;
    p = (REAL)-0.00367342844 + p * w;
    p = (REAL)0.00573950773 + p * w;
    p = (REAL)-0.0076224613 + p * w;
    p = (REAL)0.00943887047 + p * w;
    p = (REAL)1.00167406 + p * w;
    p = (REAL)2.83297682 + p * w;
  }
  return p * x;
}
/* jax.random.normal(key, shape)[i] = sqrt(2) * erf_inv(uniform(key, minval=nextafter(-1,0), maxval=1)) */
static REAL key_normal(Key k, int i) {
  const REAL lo = (REAL)-0.99999994f;
  REAL u = key_uniform(k, i, lo, (REAL)1);
  return (REAL)1.41421356237309504880 * erfinv_f32poly(u);
}
static int key_bernoulli(Key k, int i, REAL p) { return (REAL)bits_to_unit(key_bits(k, i)) < p; }

/* exported for tests */
void o_rng_split(const uint32_t* key, int n, uint32_t* out) {
  Key k = {key[0], key[1]};
  for (int i = 0; i < n; i++) { Key o = key_split(k, i); out[2 * i] = o.k0; out[2 * i + 1] = o.k1; }
}
void o_rng_uniform(const uint32_t* key, int n, double lo, double hi, double* out) {
  Key k = {key[0], key[1]};
  for (int i = 0; i < n; i++) out[i] = (double)key_uniform(k, i, (REAL)lo, (REAL)hi);
}
void o_rng_normal(const uint32_t* key, int n, double* out) {
  Key k = {key[0], key[1]};
  for (int i = 0; i < n; i++) out[i] = (double)key_normal(k, i);
}

/* ----------------------------------------------------------------------------- noise / scale decoding */

static REAL scale_get(int kind, const REAL* f, REAL level) {
  switch (kind) {
    case SCALE_CONST: return f[0];
    case SCALE_LINEAR: return f[0] * level + f[1];
    case SCALE_QUADRATIC: return f[0] * level * level + f[1];
    case SCALE_SQRT: return f[0] * SQRT(level) + f[1];
    case SCALE_NONZERO: return level > 0 ? f[0] : f[1];
  }
  return f[0];
}

/* RandomVariable.get_random_variable(shape, key, level)[i]  (ksim/noise.py:51-121) */
static REAL rv_sample(int kind, const REAL* f, Key key, int i, REAL level) {
  REAL mean = f[0], std = f[1], mag = f[2];
  switch (kind) {
    case RV_UNIFORM: return key_uniform(key, i, mean - mag * level, mean + mag * level);
    case RV_GAUSSIAN: return key_normal(key, i) * (std * level) + mean;
    case RV_UNIFORM_GAUSSIAN: {
      REAL n = key_normal(key, i) * (std * level);
      REAL u = key_uniform(key, i, -mag * level, mag * level); /* same key for both draws (noise.py:119-120) */
      return n + u + mean;
    }
  }
  return 0;
}

/* Noise.add_noise incl. ChainedNoise (same key for every link, noise.py:223-229) */
static REAL noise_apply(const int* ni, const REAL* nf, REAL x, Key key, int i, REAL level) {
  int n = ni[0];
  for (int l = 0; l < n; l++) {
    int mul = ni[1 + 2 * l], kind = ni[2 + 2 * l];
    REAL r = rv_sample(kind, nf + 3 * l, key, i, level);
    x = mul ? x * r : x + r;
  }
  return x;
}

/* ------------------------------------------------------------------------------------------ task view */

typedef struct {
  const int* ti;
  const REAL* tf;
  int nsub, act_period, act_type;
  int S, R, RPRE, RAND, K;
} Task;

static void task_init(Task* t, const int* ti, const REAL* tf) {
  t->ti = ti; t->tf = tf;
  t->nsub = ti[TI_NSUB]; t->act_period = ti[TI_ACT_PERIOD]; t->act_type = ti[TI_ACT_TYPE];
  t->S = ti[TI_STATE_SIZE]; t->R = ti[TI_REC_SIZE]; t->RPRE = ti[TI_REC_PRE_SIZE];
  t->RAND = ti[TI_RAND_SIZE]; t->K = ti[TI_KEY_SIZE];
}
#define TAB(t, base, i, f) ((t)->ti[(t)->ti[base] + (i)*TAB_STRIDE + (f)])

/* per-env scratch */
typedef struct {
  OModel model;
  OData d;
  REAL level;
  REAL last_ctrl[2 * O_MAXNU], latency, zero_off[O_MAXNU];
  REAL event[32], act_state[2 * O_MAXNU + 3], cmd[32], obs_carry[512];  /* delayed observations: delay_steps x joints each */
  Key rng, bias_key[32];
  REAL nonfinite;
} Env;

static void apply_overrides(const Task* t, const OModel* base, OModel* m, const REAL* rand) {
  memcpy(m, base, sizeof(OModel));
  int n = t->ti[TI_N_RAND], off = 0;
  for (int i = 0; i < n; i++) {
    int code = TAB(t, TI_RAND_TAB, i, TAB_AUX0), cnt = TAB(t, TI_RAND_TAB, i, TAB_AUX1);
    REAL* dst = 0;
    switch (code) {
      case RAND_DOF_FRICTIONLOSS: dst = m->dof_frictionloss; break;
      case RAND_DOF_ARMATURE: dst = m->dof_armature; break;
      case RAND_DOF_DAMPING: dst = m->dof_damping; break;
      case RAND_BODY_MASS: dst = m->body_mass; break;
      case RAND_BODY_INERTIA: dst = &m->body_inertia[0][0]; break;
      case RAND_BODY_IPOS: dst = &m->body_ipos[0][0]; break;
      case RAND_GEOM_SIZE: dst = &m->geom_size[0][0]; break;
      case RAND_GEOM_POS: dst = &m->geom_pos[0][0]; break;
      case RAND_SITE_QUAT: dst = &m->site_quat[0][0]; break;  /* IMUAlignmentRandomizer, randomization.py:245-283 */
      case RAND_SITE_POS: dst = &m->site_pos[0][0]; break;
    }
    if (dst) for (int k = 0; k < cnt; k++) dst[k] = rand[off + k];
    off += cnt;
  }
}

static void env_unpack(const Task* t, Env* e, const REAL* s, const uint32_t* keys) {
  const int* ti = t->ti;
  const OModel* m = &e->model;
  OData* d = &e->d;
  memcpy(d->qpos, s + ti[TI_S_QPOS], m->nq * sizeof(REAL));
  memcpy(d->qvel, s + ti[TI_S_QVEL], m->nv * sizeof(REAL));
  memcpy(d->qacc, s + ti[TI_S_QACC], m->nv * sizeof(REAL));
  memcpy(d->qacc_warmstart, s + ti[TI_S_WARM], m->nv * sizeof(REAL));
  d->time = s[ti[TI_S_TIME]];
  memcpy(d->ctrl, s + ti[TI_S_CTRL], m->nu * sizeof(REAL));
  memcpy(e->last_ctrl, s + ti[TI_S_LAST_CTRL], ti[TI_ACTION_SIZE] * sizeof(REAL));
  e->latency = s[ti[TI_S_LATENCY]];
  memcpy(e->zero_off, s + ti[TI_S_ZERO_OFF], m->nu * sizeof(REAL));
  memcpy(e->event, s + ti[TI_S_EVENT], ti[TI_EVENT_SIZE] * sizeof(REAL));
  memcpy(e->act_state, s + ti[TI_S_ACT_STATE], ti[TI_ACT_STATE_SIZE] * sizeof(REAL));
  e->level = s[ti[TI_S_LEVEL]];
  memcpy(e->cmd, s + ti[TI_S_CMD], ti[TI_CMD_SIZE] * sizeof(REAL));
  memcpy(e->obs_carry, s + ti[TI_S_OBS_CARRY], ti[TI_OBS_CARRY_SIZE] * sizeof(REAL));
  e->nonfinite = s[ti[TI_S_NONFINITE]];
  e->rng.k0 = keys[0]; e->rng.k1 = keys[1];
  for (int i = 0; i < ti[TI_N_OBS]; i++) { e->bias_key[i].k0 = keys[2 + 2 * i]; e->bias_key[i].k1 = keys[3 + 2 * i]; }
}

static void env_pack(const Task* t, const Env* e, REAL* s, uint32_t* keys) {
  const int* ti = t->ti;
  const OModel* m = &e->model;
  const OData* d = &e->d;
  memcpy(s + ti[TI_S_QPOS], d->qpos, m->nq * sizeof(REAL));
  memcpy(s + ti[TI_S_QVEL], d->qvel, m->nv * sizeof(REAL));
  memcpy(s + ti[TI_S_QACC], d->qacc, m->nv * sizeof(REAL));
  memcpy(s + ti[TI_S_WARM], d->qacc_warmstart, m->nv * sizeof(REAL));
  s[ti[TI_S_TIME]] = d->time;
  memcpy(s + ti[TI_S_CTRL], d->ctrl, m->nu * sizeof(REAL));
  memcpy(s + ti[TI_S_LAST_CTRL], e->last_ctrl, ti[TI_ACTION_SIZE] * sizeof(REAL));
  s[ti[TI_S_LATENCY]] = e->latency;
  memcpy(s + ti[TI_S_ZERO_OFF], e->zero_off, m->nu * sizeof(REAL));
  memcpy(s + ti[TI_S_EVENT], e->event, ti[TI_EVENT_SIZE] * sizeof(REAL));
  memcpy(s + ti[TI_S_ACT_STATE], e->act_state, ti[TI_ACT_STATE_SIZE] * sizeof(REAL));
  s[ti[TI_S_LEVEL]] = e->level;
  memcpy(s + ti[TI_S_CMD], e->cmd, ti[TI_CMD_SIZE] * sizeof(REAL));
  memcpy(s + ti[TI_S_OBS_CARRY], e->obs_carry, ti[TI_OBS_CARRY_SIZE] * sizeof(REAL));
  s[ti[TI_S_NONFINITE]] = e->nonfinite;
  keys[0] = e->rng.k0; keys[1] = e->rng.k1;
  for (int i = 0; i < ti[TI_N_OBS]; i++) { keys[2 + 2 * i] = e->bias_key[i].k0; keys[3 + 2 * i] = e->bias_key[i].k1; }
}

/* ------------------------------------------------------------------------------------------- actuators */

static void actuators_ctrl(const Task* t, Env* e, const REAL* ctrl_in, Key rng, REAL* torque) {
  const OModel* m = &e->model;
  const OData* d = &e->d;
  int nu = m->nu;
  const int* ai = t->ti + t->ti[TI_ACT_IOFF];
  const REAL* af = t->tf + t->ti[TI_ACT_FOFF];
  REAL level = e->level;
  if (t->act_type == ACT_TORQUE) {
    /* TorqueActuators.get_ctrl: noise.add_noise(action, level, rng)   (actuators.py:78-87) */
    for (int i = 0; i < nu; i++) torque[i] = noise_apply(ai, af, ctrl_in[i], rng, i, level);
    return;
  }
  /* PositionActuators.get_stateful_ctrl (actuators.py:182-217) */
  const int* an_i = ai; const int* tn_i = ai + NOISE_NI;
  REAL action_scale = af[0];
  const REAL* an_f = af + 1; const REAL* tn_f = an_f + NOISE_NF;
  const REAL* kp = tn_f + NOISE_NF + 15; const REAL* kd = kp + nu; const REAL* clip = kd + nu;
  if (t->act_type == ACT_POSITION_VELOCITY) {
    /* PositionVelocityActuator.get_ctrl (actuators.py:261-293): split(rng, 3); action = [position | velocity] targets */
    const int* vn_i = ai + 2 * NOISE_NI + 5;
    const REAL* vn_f = clip + nu;
    Key pos_rng = key_split(rng, 0), vel_rng = key_split(rng, 1), tor_rng = key_split(rng, 2);
    for (int i = 0; i < nu; i++) {
      REAL tp = noise_apply(an_i, an_f, ctrl_in[i], pos_rng, i, level);
      REAL tv = noise_apply(vn_i, vn_f, ctrl_in[nu + i], vel_rng, i, level);
      REAL c = kp[i] * (tp - (d->qpos[7 + i] - e->zero_off[i])) + kd[i] * (tv - d->qvel[6 + i]);
      c = noise_apply(tn_i, tn_f, c, tor_rng, i, level);
      if (c < -clip[i]) c = -clip[i];
      if (c > clip[i]) c = clip[i];
      torque[i] = c;
    }
    return;
  }
  const REAL* action_bias = e->act_state; const REAL* torque_bias = e->act_state + nu;
  REAL kp_scale = e->act_state[2 * nu], kd_scale = e->act_state[2 * nu + 1], tl_scale = e->act_state[2 * nu + 2];
  Key pos_rng = key_split(rng, 0), tor_rng = key_split(rng, 1);
  for (int i = 0; i < nu; i++) {
    REAL scaled = ctrl_in[i] * action_scale;
    REAL target = noise_apply(an_i, an_f, scaled, pos_rng, i, level) + action_bias[i];
    REAL cur_pos = d->qpos[7 + i] - e->zero_off[i]; /* joint order, engine.py:297 */
    REAL cur_vel = d->qvel[6 + i];
    REAL c = (kp[i] * kp_scale) * (target - cur_pos) + (kd[i] * kd_scale) * ((REAL)0 - cur_vel);
    c = noise_apply(tn_i, tn_f, c, tor_rng, i, level) + torque_bias[i];
    REAL lim = clip[i] * tl_scale;
    if (c < -lim) c = -lim;
    if (c > lim) c = lim;
    torque[i] = c;
  }
}

static void actuators_initial_state(const Task* t, Env* e, Key rng) {
  if (t->act_type != ACT_POSITION) return;
  int nu = e->model.nu;
  const int* ai = t->ti + t->ti[TI_ACT_IOFF];
  const REAL* af = t->tf + t->ti[TI_ACT_FOFF];
  const int* kinds = ai + 2 * NOISE_NI;
  const REAL* rvf = af + 1 + 2 * NOISE_NF;
  /* get_initial_state (actuators.py:219-238): split(rng, 5); level defaults to 1.0 */
  for (int i = 0; i < nu; i++) {
    e->act_state[i] = kinds[0] >= 0 ? rv_sample(kinds[0], rvf + 0, key_split(rng, 0), i, 1) : 0;
    e->act_state[nu + i] = kinds[1] >= 0 ? rv_sample(kinds[1], rvf + 3, key_split(rng, 1), i, 1) : 0;
  }
  e->act_state[2 * nu + 0] = kinds[2] >= 0 ? rv_sample(kinds[2], rvf + 6, key_split(rng, 2), 0, 1) : 1;
  e->act_state[2 * nu + 1] = kinds[3] >= 0 ? rv_sample(kinds[3], rvf + 9, key_split(rng, 3), 0, 1) : 1;
  e->act_state[2 * nu + 2] = kinds[4] >= 0 ? rv_sample(kinds[4], rvf + 12, key_split(rng, 4), 0, 1) : 1;
}

/* ---------------------------------------------------------------------------------------------- events */

static void event_apply(const Task* t, Env* e, int idx, Key rng) {
  int type = TAB(t, TI_EVENT_TAB, idx, TAB_TYPE);
  const int* ei = t->ti + TAB(t, TI_EVENT_TAB, idx, TAB_IOFF);
  const REAL* ef = t->tf + TAB(t, TI_EVENT_TAB, idx, TAB_FOFF);
  REAL* st = e->event + TAB(t, TI_EVENT_TAB, idx, TAB_OUT_OFF);
  OData* d = &e->d;
  REAL dt = (REAL)(float)e->model.timestep; /* jnp.float32(model.opt.timestep) */
  if (type == EVENT_LINEAR_PUSH || type == EVENT_ANGULAR_PUSH || type == EVENT_JUMP) {
    REAL tr = st[0] - dt;
    if (tr <= 0) {
      if (type == EVENT_LINEAR_PUSH) {
        /* events.py:140-166 */
        Key urng = key_split(rng, 0), brng = key_split(rng, 1), trng = key_split(rng, 2);
        REAL lvl = scale_get(ei[0], ef + 5, e->level);
        REAL theta = key_uniform(urng, 0, 0, (REAL)2 * PI_R);
        REAL mag = key_uniform(brng, 0, ef[1], ef[2]) * ef[0];
        d->qvel[0] += COS(theta) * mag * lvl;
        d->qvel[1] += SIN(theta) * mag * lvl;
        d->qvel[2] += (REAL)0 * mag * lvl;
        tr = key_uniform(trng, 0, ef[3], ef[4]);
      } else if (type == EVENT_ANGULAR_PUSH) {
        /* events.py:82-108 */
        Key frng = key_split(rng, 0), brng = key_split(rng, 1), trng = key_split(rng, 2);
        REAL lvl = scale_get(ei[0], ef + 5, e->level);
        int flip = key_bernoulli(frng, 0, (REAL)0.5);
        REAL mag = key_uniform(brng, 0, ef[1], ef[2]) * ef[0];
        if (flip) mag = -mag;
        d->qvel[5] += mag * lvl;
        tr = key_uniform(trng, 0, ef[3], ef[4]);
      } else {
        /* JumpEvent._apply_jump events.py:198-228: sqrt(2 * -g * h) per gravity component */
        Key urng = key_split(rng, 0), trng = key_split(rng, 1);
        REAL lvl = scale_get(ei[0], ef + 4, e->level);
        REAL h = key_uniform(urng, 0, ef[0], ef[1]) * lvl;
        for (int k = 0; k < 3; k++) d->qvel[k] += SQRT((REAL)2 * -e->model.gravity[k] * h);
        tr = key_uniform(trng, 0, ef[2], ef[3]);
      }
    }
    st[0] = tr;
  } else if (type == EVENT_FORCE_PUSH) {
    /* events.py:280-352 */
    REAL lvl = scale_get(ei[0], ef + 8, e->level);
    int body = ei[1];
    REAL tr = st[0] - dt, dur = st[1] - dt;
    REAL force[6];
    for (int k = 0; k < 6; k++) force[k] = st[2 + k] * lvl;
    memset(d->xfrc_applied, 0, sizeof(d->xfrc_applied));
    for (int k = 0; k < 6; k++) d->xfrc_applied[body][k] = dur > 0 ? force[k] * lvl : 0;
    if (tr <= 0) {
      Key k0 = key_split(rng, 0), k1 = key_split(rng, 1), k2 = key_split(rng, 2), k3 = key_split(rng, 3);
      Key k4 = key_split(rng, 4), k5 = key_split(rng, 5);
      REAL dir[3], tq[3];
      for (int k = 0; k < 3; k++) dir[k] = key_uniform(k0, k, -1, 1);
      REAL fs = key_uniform(k1, 0, ef[2], ef[3]);
      for (int k = 0; k < 3; k++) tq[k] = key_uniform(k2, k, -1, 1);
      REAL nd = SQRT(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
      REAL nt = SQRT(tq[0] * tq[0] + tq[1] * tq[1] + tq[2] * tq[2]);
      /* jax.random.randint(key, (), 0, 3): span 3; uses two 32-bit draws (split) and modular arithmetic */
      Key ra = key_split(k3, 0), rb = key_split(k3, 1);
      uint32_t hi = key_bits(ra, 0), lo = key_bits(rb, 0);
      uint32_t span = 3u, mult = (uint32_t)((65536u % span) * (65536u % span)) % span;
      uint32_t off = ((hi % span) * mult + (lo % span)) % span;
      for (int k = 0; k < 3; k++) {
        REAL f = dir[k] / nd * fs * ef[0], q = tq[k] / nt * fs * ef[1];
        force[k] = (off == 1) ? 0 : f;
        force[3 + k] = (off == 0) ? 0 : q;
      }
      tr = key_uniform(k4, 0, ef[6], ef[7]);
      dur = key_uniform(k5, 0, ef[4], ef[5]);
    }
    st[0] = tr; st[1] = dur;
    for (int k = 0; k < 6; k++) st[2 + k] = force[k];
  }
}

static void event_initial_state(const Task* t, Env* e, int idx, Key rng) {
  int type = TAB(t, TI_EVENT_TAB, idx, TAB_TYPE);
  const REAL* ef = t->tf + TAB(t, TI_EVENT_TAB, idx, TAB_FOFF);
  REAL* st = e->event + TAB(t, TI_EVENT_TAB, idx, TAB_OUT_OFF);
  switch (type) {
    case EVENT_LINEAR_PUSH: case EVENT_ANGULAR_PUSH: st[0] = key_uniform(rng, 0, ef[3], ef[4]); break;
    case EVENT_JUMP: st[0] = key_uniform(rng, 0, ef[2], ef[3]); break;
    case EVENT_FORCE_PUSH:
      st[0] = key_uniform(rng, 0, ef[6], ef[7]);
      for (int k = 1; k < 8; k++) st[k] = 0;
      break;
  }
}

/* ----------------------------------------------------------------------------------- engine.step / reset */

/* MjxEngine.step (engine.py:259-361) */
static void engine_step(const Task* t, Env* e, const REAL* action, Key rng) {
  const OModel* m = &e->model;
  OData* d = &e->d;
  int na = t->ti[TI_ACTION_SIZE], nu = m->nu;
  const REAL* ef = t->tf + t->ti[TI_ENGINE_FOFF];
  Key drop_rng = key_split(rng, 0), scan_rng = key_split(rng, 1);
  int drop[2 * O_MAXNU];
  for (int i = 0; i < na; i++) drop[i] = key_bernoulli(drop_rng, i, ef[2] * e->level);
  REAL torques[O_MAXNU];
  for (int i = 0; i < nu; i++) torques[i] = 0;
  Key r = scan_rng;
  for (int step = 0; step < t->nsub; step++) {
    REAL prct = (REAL)step - e->latency;
    if (prct < 0) prct = 0;
    if (prct > 1) prct = 1;
    REAL ctrl[2 * O_MAXNU];
    for (int i = 0; i < na; i++)
      ctrl[i] = drop[i] ? e->last_ctrl[i] : e->last_ctrl[i] * ((REAL)1 - prct) + action[i] * prct;
    Key rng_update = key_split(r, 1);
    r = key_split(r, 0);
    if (step % t->act_period == 0) actuators_ctrl(t, e, ctrl, rng_update, torques);
    for (int i = 0; i < nu; i++) d->ctrl[i] = torques[i];
    for (int ev = 0; ev < t->ti[TI_N_EVENT]; ev++) {
      Key rng_event = key_split(r, 1);
      r = key_split(r, 0);
      event_apply(t, e, ev, rng_event);
    }
    o_step(m, d);
  }
  for (int i = 0; i < na; i++) e->last_ctrl[i] = action[i];
}

/* element i of jax.random.randint(key, (n,), lo, hi): offset in [0, span) */
static uint32_t key_randint_at(Key k, int i, uint32_t span) {
  Key ra = key_split(k, 0), rb = key_split(k, 1);
  uint32_t hi = key_bits(ra, i), lo = key_bits(rb, i);
  uint32_t mult = (uint32_t)((65536u % span) * (65536u % span)) % span;
  return ((hi % span) * mult + (lo % span)) % span;
}

static void reset_apply(const Task* t, Env* e, int idx, Key rng) {
  int type = TAB(t, TI_RESET_TAB, idx, TAB_TYPE);
  const int* ri = t->ti + TAB(t, TI_RESET_TAB, idx, TAB_IOFF);
  const REAL* rf = t->tf + TAB(t, TI_RESET_TAB, idx, TAB_FOFF);
  const OModel* m = &e->model;
  OData* d = &e->d;
  int nj = m->nq - 7;
  switch (type) {
    case RESET_RANDOM_JOINT_POSITION: {
      /* resets.py:132-145 */
      const REAL *zeros = rf + 1, *mins = rf + 1 + nj, *maxs = rf + 1 + 2 * nj;
      for (int i = 0; i < nj; i++) {
        REAL p = key_uniform(rng, i, -rf[0], rf[0]);
        if (ri[0]) p = p * e->level;
        if (ri[1]) p = p + zeros[i];
        if (ri[2] && p < mins[i]) p = mins[i];
        if (ri[3] && p > maxs[i]) p = maxs[i];
        d->qpos[7 + i] = p;
      }
    } break;
    case RESET_RANDOM_JOINT_VELOCITY:
      for (int i = 0; i < m->nv - 6; i++) {
        REAL n = key_uniform(rng, i, -rf[0], rf[0]);
        if (ri[0]) n = n * e->level;
        d->qvel[6 + i] = n;
      }
      break;
    case RESET_RANDOM_HEADING: {
      REAL ang = key_uniform(rng, 0, -PI_R, PI_R);
      REAL rpy[3] = {0, 0, ang}, q[4], r[4];
      o_euler_to_quat(q, rpy);
      o_quat_mul(r, d->qpos + 3, q);
      for (int k = 0; k < 4; k++) d->qpos[3 + k] = r[k];
    } break;
    case RESET_RANDOM_BASE_VELOCITY_XY:
      /* no-op on the MJX path: `qvel.at[0:2].set(noise)` result is discarded (resets.py:204-206) */
      break;
    case RESET_PLANE_XY_POSITION: {
      /* resets.py:103-119; floats [z, x_range, y_range, lower_x, upper_x, lower_y, upper_y] */
      Key kx = key_split(rng, 0), ky = key_split(rng, 1);
      REAL x = key_uniform(kx, 0, -rf[1], rf[1]), y = key_uniform(ky, 0, -rf[2], rf[2]);
      if (x < rf[3]) x = rf[3]; if (x > rf[4]) x = rf[4];
      if (y < rf[5]) y = rf[5]; if (y > rf[6]) y = rf[6];
      d->qpos[0] = x; d->qpos[1] = y; d->qpos[2] = rf[0];
    } break;
    case RESET_HFIELD_XY_POSITION: {
      /* resets.py:56-89; ints [nx, ny]; floats [x_bound, y_bound, z_top, base_height, x_range, y_range, lower_x, upper_x,
         lower_y, upper_y, heights[nx][ny]] */
      Key kx = key_split(rng, 0), ky = key_split(rng, 1);
      REAL x = key_uniform(kx, 0, -rf[4], rf[4]), y = key_uniform(ky, 0, -rf[5], rf[5]);
      if (x < rf[6]) x = rf[6]; if (x > rf[7]) x = rf[7];
      if (y < rf[8]) y = rf[8]; if (y > rf[9]) y = rf[9];
      int nx = ri[0], ny = ri[1];
      int xi = (int)(((x + rf[0]) / ((REAL)2 * rf[0])) * (REAL)(nx - 1));
      int yi = (int)(((y + rf[1]) / ((REAL)2 * rf[1])) * (REAL)(ny - 1));
      xi = xi < 0 ? 0 : (xi > nx - 1 ? nx - 1 : xi);
      yi = yi < 0 ? 0 : (yi > ny - 1 ? ny - 1 : yi);
      d->qpos[0] = x; d->qpos[1] = y;
      d->qpos[2] = rf[10 + xi * ny + yi] + rf[2] + rf[3];
    } break;
    case RESET_INITIAL_MOTION_STATE: {
      /* resets.py:290-304, freejoint = False; ints [num_frames, nq]; floats [ctrl_dt, qpos[frames][nq], qvel[frames][nq]] */
      int frames = ri[0], wq = ri[1];
      int frame = (int)key_randint_at(rng, 0, (uint32_t)frames);
      const REAL* qp = rf + 1 + frame * wq;
      const REAL* qv = rf + 1 + (frames + frame) * wq;
      for (int i = 0; i < nj; i++) { d->qpos[7 + i] = qp[7 + i]; d->qvel[6 + i] = qv[7 + i]; }
      d->time = (REAL)frame * rf[0];
    } break;
    case RESET_RANDOM_HEIGHT: d->qpos[2] += key_uniform(rng, 0, rf[0], rf[1]); break;
    case RESET_RANDOM_PITCH_ROLL: {
      /* same key for pitch and roll draws (resets.py:346-347) */
      REAL pitch = key_uniform(rng, 0, rf[0], rf[1]), roll = key_uniform(rng, 0, rf[2], rf[3]);
      REAL rpy[3] = {roll, pitch, 0}, q[4], r[4];
      o_euler_to_quat(q, rpy);
      o_quat_mul(r, d->qpos + 3, q);
      for (int k = 0; k < 4; k++) d->qpos[3 + k] = r[k];
    } break;
  }
}

/* MjxEngine.reset (engine.py:205-246) */
static void engine_reset(const Task* t, Env* e, Key rng) {
  const OModel* m = &e->model;
  OData* d = &e->d;
  const REAL* ef = t->tf + t->ti[TI_ENGINE_FOFF];
  o_make_data(m, d);
  for (int i = 0; i < t->ti[TI_N_RESET]; i++) {
    Key reset_rng = key_split(rng, 1);
    rng = key_split(rng, 0);
    reset_apply(t, e, i, reset_rng);
  }
  o_forward(m, d);
  for (int i = 0; i < m->nv; i++) { d->qvel[i] = 0; d->qacc[i] = 0; }
  actuators_initial_state(t, e, rng);
  Key events_rng = key_split(rng, 0), latency_rng = key_split(rng, 1), zero_rng = key_split(rng, 2);
  for (int i = 0; i < t->ti[TI_ACTION_SIZE]; i++) e->last_ctrl[i] = 0;
  for (int ev = 0; ev < t->ti[TI_N_EVENT]; ev++) {
    Key event_rng = key_split(events_rng, 1);
    events_rng = key_split(events_rng, 0);
    event_initial_state(t, e, ev, event_rng);
  }
  e->latency = key_uniform(latency_rng, 0, ef[0] * e->level, ef[1] * e->level);
  for (int i = 0; i < m->nu; i++) e->zero_off[i] = rv_sample(t->ti[TI_ZERO_OFFSET_KIND], ef + 3, zero_rng, i, e->level);
}

/* ------------------------------------------------------------------------------------------ commands */

/* xax.quat_to_euler(wxyz) -> [roll, pitch, yaw], normalising with eps 1e-6 [restated from memory, like o_euler_to_quat] */
static void quat_to_euler(REAL* rpy, const REAL* q) {
  REAL n = SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + (REAL)1e-6;
  REAL w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  rpy[0] = ATAN2((REAL)2 * (w * x + y * z), (REAL)1 - (REAL)2 * (x * x + y * y));
  REAL sinp = (REAL)2 * (w * y - z * x);
  rpy[1] = FABS(sinp) >= (REAL)1 ? (sinp > 0 ? PI_R / 2 : -PI_R / 2) : ASIN(sinp);
  rpy[2] = ATAN2((REAL)2 * (w * z + x * y), (REAL)1 - (REAL)2 * (y * y + z * z));
}

/* jax.random.randint(key, (), 0, span): two 32-bit draws from split(key) and modular arithmetic */
static uint32_t key_randint(Key k, uint32_t span) {
  Key ra = key_split(k, 0), rb = key_split(k, 1);
  uint32_t hi = key_bits(ra, 0), lo = key_bits(rb, 0);
  uint32_t mult = (uint32_t)((65536u % span) * (65536u % span)) % span;
  return ((hi % span) * mult + (lo % span)) % span;
}

/* CartesianCoordinateCommand._sample_box (commands.py:557-581); floats [box_min 3, box_max 3, dt, min_speed, max_speed,
   switch_prob, jump_prob, curriculum_scale] */
static void cmd_cartesian_sample(const REAL* cf, REAL level, Key rng, REAL* out) {
  REAL sf = cf[11] * level + (REAL)1;
  for (int k = 0; k < 3; k++) {
    REAL center = (cf[k] + cf[3 + k]) / (REAL)2, half = (cf[3 + k] - cf[k]) / (REAL)2 * sf;
    out[k] = key_uniform(rng, k, center - half, center + half);
  }
}
/* SinusoidalGaitCommand._get_height (commands.py:712-733); floats [gait_period, ctrl_dt, max_height, height_offset,
   stance_ratio, moving_threshold] */
static void cmd_gait_heights(const REAL* cf, int nf, REAL phase, REAL* out) {
  REAL max_h = cf[2], offset = cf[3], sr = cf[4];
  for (int f = 0; f < nf; f++) {
    REAL pf = phase + (REAL)f / (REAL)nf;
    pf = pf - FLOOR(pf);
    REAL prog = (pf - sr) / ((REAL)1 - sr);
    prog = prog < 0 ? 0 : (prog > 1 ? 1 : prog);
    out[f] = (pf >= sr ? max_h * SIN((REAL)3.14159265358979323846 * prog) : 0) + offset;
  }
}

/* UnifiedCommand.initial_command (examples/kbot/train.py:714-756).  floats: 6 x [lo, hi] (vx vy wz bh rx ry),
 * arms lo[10], arms hi[10], switch_prob.  The arm draw and its 50 % mask share one key (train.py:725-728). */
static void cmd_unified_initial(const REAL* cf, Key rng, REAL* out) {
  REAL v[6], arms[10];
  for (int k = 0; k < 6; k++) v[k] = key_uniform(key_split(rng, 1 + k), 0, cf[2 * k], cf[2 * k + 1]);
  Key rh = key_split(rng, 7);
  for (int k = 0; k < 10; k++) {
    REAL a = key_uniform(rh, k, cf[12 + k], cf[22 + k]);
    arms[k] = key_bernoulli(rh, k, (REAL)0.5) ? a : 0;
  }
  uint32_t mode = key_randint(key_split(rng, 0), 6u);
  for (int k = 0; k < 16; k++) out[k] = 0;
  switch (mode) {
    case 0: out[0] = v[0]; break;
    case 1: out[1] = v[1]; break;
    case 2: out[2] = v[2]; break;
    case 3: out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; for (int k = 0; k < 10; k++) out[6 + k] = arms[k]; break;
    case 4: out[3] = v[3]; out[4] = v[4]; out[5] = v[5]; for (int k = 0; k < 10; k++) out[6 + k] = arms[k]; break;
    default: break;
  }
}

/* COMDistanceObservation (examples/kbot/train.py:500-649): Andrew's monotone chain over the n contact points
 * (pts[i] = contact i's xy, or (0,0) when its geom1 is not the floor), area centroid of the hull polygon (mean of its
 * vertices when the area vanishes), distance to `com`. */
static REAL hull_cross(const REAL* a, const REAL* b, const REAL* c) {
  return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}
static REAL com_distance(const REAL (*pts)[2], int n, const REAL* com) {
  int order[O_MAXCON];
  for (int i = 0; i < n; i++) order[i] = i;
  for (int i = 1; i < n; i++) {  /* stable insertion sort == lexsort by x then y */
    int o = order[i], j = i - 1;
    while (j >= 0 && (pts[order[j]][0] > pts[o][0] || (pts[order[j]][0] == pts[o][0] && pts[order[j]][1] > pts[o][1]))) {
      order[j + 1] = order[j]; j--;
    }
    order[j + 1] = o;
  }
  REAL poly[2 * O_MAXCON][2];
  int np_ = 0;
  for (int pass = 0; pass < 2; pass++) {
    int stack[O_MAXCON], ptr = 0;
    for (int t = 0; t < n; t++) {
      int idx = pass == 0 ? t : n - 1 - t;
      while (ptr >= 2 && hull_cross(pts[order[stack[ptr - 2]]], pts[order[stack[ptr - 1]]], pts[order[idx]]) <= 0) ptr--;
      stack[ptr++] = idx;
    }
    for (int t = 0; t < ptr - 1; t++) { poly[np_][0] = pts[order[stack[t]]][0]; poly[np_][1] = pts[order[stack[t]]][1]; np_++; }
  }
  REAL area = 0, sx = 0, sy = 0, mx = 0, my = 0;
  for (int i = 0; i < np_; i++) {
    int j = i + 1 < np_ ? i + 1 : 0;
    REAL cr = poly[i][0] * poly[j][1] - poly[j][0] * poly[i][1];
    area += cr;
    sx += (poly[i][0] + poly[j][0]) * cr;
    sy += (poly[i][1] + poly[j][1]) * cr;
    mx += poly[i][0]; my += poly[i][1];
  }
  area *= (REAL)0.5;
  REAL cnt = np_ > 0 ? (REAL)np_ : (REAL)1;
  REAL cx = FABS(area) < (REAL)1e-12 ? mx / cnt : sx / ((REAL)6 * area);
  REAL cy = FABS(area) < (REAL)1e-12 ? my / cnt : sy / ((REAL)6 * area);
  return SQRT((cx - com[0]) * (cx - com[0]) + (cy - com[1]) * (cy - com[1]));
}

static void cmd_linvel_initial(const int* ci, const REAL* cf, const Env* e, Key rng, const REAL* prev, REAL* out) {
  /* commands.py:295-342; value order: vel, yaw, xvel, yvel, target_vel, target_yaw */
  REAL lvl = e->level;
  REAL min_vel = scale_get(ci[0], cf + 0, lvl), max_vel = scale_get(ci[1], cf + 2, lvl);
  REAL max_yaw = scale_get(ci[2], cf + 4, lvl), zero_prob = scale_get(ci[3], cf + 6, lvl);
  REAL backward_prob = scale_get(ci[4], cf + 8, lvl);
  Key rv = key_split(rng, 0), ry = key_split(rng, 1), rz = key_split(rng, 2), rb = key_split(rng, 3);
  REAL cur_vel, cur_yaw;
  if (!prev) {
    REAL lv[3];
    o_rot_vec_quat(lv, e->d.qvel, e->d.qpos + 3, 1);
    cur_vel = SQRT(lv[0] * lv[0] + lv[1] * lv[1]);
    cur_yaw = ATAN2(lv[1], lv[0]);
  } else { cur_vel = prev[0]; cur_yaw = prev[1]; }
  REAL trg_vel = key_uniform(rv, 0, min_vel, max_vel);
  REAL trg_yaw = key_uniform(ry, 0, -max_yaw, max_yaw);
  int zero = key_bernoulli(rz, 0, zero_prob), back = key_bernoulli(rb, 0, backward_prob);
  trg_vel = zero ? 0 : (back ? -trg_vel : trg_vel);
  trg_yaw = zero ? 0 : trg_yaw;
  out[0] = cur_vel; out[1] = cur_yaw; out[2] = cur_vel * COS(cur_yaw); out[3] = cur_vel * SIN(cur_yaw);
  out[4] = trg_vel; out[5] = trg_yaw;
}

static REAL clipr(REAL x, REAL lo, REAL hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void cmd_initial(const Task* t, Env* e, int idx, Key rng) {
  int type = TAB(t, TI_CMD_TAB, idx, TAB_TYPE);
  const int* ci = t->ti + TAB(t, TI_CMD_TAB, idx, TAB_IOFF);
  const REAL* cf = t->tf + TAB(t, TI_CMD_TAB, idx, TAB_FOFF);
  REAL* st = e->cmd + TAB(t, TI_CMD_TAB, idx, TAB_OUT_OFF);
  int dim = TAB(t, TI_CMD_TAB, idx, TAB_OUT_DIM);
  if (type == CMD_LINEAR_VELOCITY) cmd_linvel_initial(ci, cf, e, rng, 0, st);
  else if (type == CMD_ANGULAR_VELOCITY) {
    /* commands.py:461-479; value order: vel, target_vel */
    REAL lvl = e->level;
    REAL min_vel = scale_get(ci[0], cf + 0, lvl), max_vel = scale_get(ci[1], cf + 2, lvl);
    REAL zero_prob = scale_get(ci[2], cf + 4, lvl);
    Key rv = key_split(rng, 0), rfl = key_split(rng, 1), rz = key_split(rng, 2);
    REAL trg = key_uniform(rv, 0, min_vel, max_vel);
    if (key_bernoulli(rfl, 0, (REAL)0.5)) trg = -trg;
    if (key_bernoulli(rz, 0, zero_prob)) trg = 0;
    st[0] = e->d.qvel[5]; st[1] = trg;
  } else if (type == CMD_FLOAT_VECTOR) {
    /* commands.py:112-121: bernoulli and uniform draw share the key; floats [lo,hi]*n, switch_prob, zero_prob */
    int zero = key_bernoulli(rng, 0, cf[2 * dim + 1]);
    for (int i = 0; i < dim; i++) st[i] = zero ? 0 : key_uniform(rng, i, cf[2 * i], cf[2 * i + 1]);
  } else if (type == CMD_UNIFIED) cmd_unified_initial(cf, rng, st);
  else if (type == CMD_INT_VECTOR) {
    /* commands.py:147-157: jnp.where(bernoulli(rng, zero_prob), 0.0, randint(rng, (n,), lo, hi + 1)); ints [lo, hi]*n,
       floats [switch_prob, zero_prob] */
    int zero = key_bernoulli(rng, 0, cf[1]);
    for (int i = 0; i < dim; i++)
      st[i] = zero ? 0 : (REAL)(ci[2 * i] + (int)key_randint_at(rng, i, (uint32_t)(ci[2 * i + 1] + 1 - ci[2 * i])));
  } else if (type == CMD_START_POSITION) { for (int k = 0; k < 3; k++) st[k] = e->d.qpos[k]; }      /* commands.py:176-182 */
  else if (type == CMD_START_QUATERNION) { for (int k = 0; k < 4; k++) st[k] = e->d.qpos[3 + k]; }  /* commands.py:198-204 */
  else if (type == CMD_BASE_HEIGHT) st[0] = key_uniform(rng, 0, cf[0], cf[1]);                       /* commands.py:778-784 */
  else if (type == CMD_CARTESIAN_COORDINATE) {
    /* commands.py:583-597: [target, target, speed] */
    REAL tg[3];
    cmd_cartesian_sample(cf, e->level, key_split(rng, 0), tg);
    for (int k = 0; k < 3; k++) { st[k] = tg[k]; st[3 + k] = tg[k]; }
    st[6] = key_uniform(key_split(rng, 1), 0, cf[7], cf[8]);
  } else if (type == CMD_SINUSOIDAL_GAIT) {
    /* commands.py:735-746 */
    st[0] = 1; st[1] = 0;
    cmd_gait_heights(cf, ci[0], 0, st + 2);
  } else if (type == CMD_JOINT_POSITION) {
    /* commands.py:834-851; ints [n, indices]; floats [lo, hi]*n, ctrl_dt, min_time, max_time */
    int n = ci[0];
    Key ra = key_split(rng, 0), rb = key_split(rng, 1);
    for (int i = 0; i < n; i++) { st[i] = e->d.qpos[7 + ci[1 + i]]; st[n + i] = key_uniform(ra, i, cf[2 * i], cf[2 * i + 1]); }
    REAL tt = key_uniform(rb, 0, cf[2 * n + 1], cf[2 * n + 2]);
    st[2 * n] = tt; st[2 * n + 1] = CEIL(tt / cf[2 * n]); st[2 * n + 2] = 0;
  }
}

static void cmd_update(const Task* t, Env* e, int idx, Key rng) {
  int type = TAB(t, TI_CMD_TAB, idx, TAB_TYPE);
  const int* ci = t->ti + TAB(t, TI_CMD_TAB, idx, TAB_IOFF);
  const REAL* cf = t->tf + TAB(t, TI_CMD_TAB, idx, TAB_FOFF);
  REAL* st = e->cmd + TAB(t, TI_CMD_TAB, idx, TAB_OUT_OFF);
  int dim = TAB(t, TI_CMD_TAB, idx, TAB_OUT_DIM);
  Key ra = key_split(rng, 0), rb = key_split(rng, 1);
  if (type == CMD_LINEAR_VELOCITY) {
    /* commands.py:344-386 */
    REAL sp = scale_get(ci[5], cf + 10, e->level);
    REAL ctrl_dt = cf[12], lin_acc = cf[13], ang_acc = cf[14];
    int sw = key_bernoulli(ra, 0, sp);
    REAL fresh[6];
    cmd_linvel_initial(ci, cf, e, rb, st, fresh);
    REAL nv = st[0] + clipr(st[4] - st[0], -lin_acc, lin_acc) * ctrl_dt;
    REAL ny = st[1] + clipr(st[5] - st[1], -ang_acc, ang_acc) * ctrl_dt;
    REAL upd[6] = {nv, ny, nv * COS(ny), nv * SIN(ny), st[4], st[5]};
    for (int k = 0; k < 6; k++) st[k] = sw ? fresh[k] : upd[k];
  } else if (type == CMD_ANGULAR_VELOCITY) {
    REAL sp = scale_get(ci[3], cf + 6, e->level);
    REAL ctrl_dt = cf[8], ang_acc = cf[9];
    int sw = key_bernoulli(ra, 0, sp);
    REAL prev0 = st[0], prev1 = st[1];
    cmd_initial(t, e, idx, rb);
    REAL fresh0 = st[0], fresh1 = st[1];
    REAL nv = prev0 + clipr(prev1 - prev0, -ang_acc, ang_acc) * ctrl_dt;
    st[0] = sw ? fresh0 : nv;
    st[1] = sw ? fresh1 : prev1;
  } else if (type == CMD_FLOAT_VECTOR) {
    REAL prev[32];
    for (int i = 0; i < dim; i++) prev[i] = st[i];
    int sw = key_bernoulli(ra, 0, cf[2 * dim]);
    cmd_initial(t, e, idx, rb);
    if (!sw) for (int i = 0; i < dim; i++) st[i] = prev[i];
  } else if (type == CMD_INT_VECTOR) {
    /* commands.py:159-169 */
    REAL prev[32];
    for (int i = 0; i < dim; i++) prev[i] = st[i];
    int sw = key_bernoulli(ra, 0, cf[0]);
    cmd_initial(t, e, idx, rb);
    if (!sw) for (int i = 0; i < dim; i++) st[i] = prev[i];
  } else if (type == CMD_BASE_HEIGHT) {
    /* commands.py:786-796: new height and switch from the same, unsplit key */
    REAL fresh = key_uniform(rng, 0, cf[0], cf[1]);
    if (key_bernoulli(rng, 0, cf[2])) st[0] = fresh;
  } else if (type == CMD_CARTESIAN_COORDINATE) {
    /* commands.py:599-643 */
    REAL dx = st[3] - st[0], dy = st[4] - st[1], dz = st[5] - st[2];
    REAL distance = SQRT(dx * dx + dy * dy + dz * dz);
    Key rc = key_split(rng, 2);
    int reached = distance < cf[6] * st[6] * (REAL)0.5;
    REAL ntg[3], nsp = key_uniform(rb, 0, cf[7], cf[8]);
    cmd_cartesian_sample(cf, e->level, ra, ntg);
    int change = (reached && key_bernoulli(rc, 0, cf[9])) || key_bernoulli(rc, 0, cf[10]);
    if (change) { for (int k = 0; k < 3; k++) st[3 + k] = ntg[k]; st[6] = nsp; }
    REAL step = st[6] * cf[6];
    REAL dd[3] = {st[3] - st[0], st[4] - st[1], st[5] - st[2]};
    REAL dn = SQRT(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
    REAL adv = step < distance ? step : distance;
    for (int k = 0; k < 3; k++) st[k] = st[k] + (dn > 0 ? dd[k] / dn : dd[k]) * adv;
  } else if (type == CMD_SINUSOIDAL_GAIT) {
    /* commands.py:748-770; physics_data is the post-step (post-reset) state */
    int moving = st[0] != 0;
    REAL ph = st[1] + cf[1] / cf[0];
    ph = moving ? ph - FLOOR(ph) : 0;
    int is_moving = SQRT(e->d.qvel[0] * e->d.qvel[0] + e->d.qvel[1] * e->d.qvel[1]) > cf[5];
    if (!is_moving) { st[0] = 0; ph = 0; }
    st[1] = ph;
    cmd_gait_heights(cf, ci[0], ph, st + 2);
  } else if (type == CMD_JOINT_POSITION) {
    /* commands.py:853-880: jax.tree.map(where(at_target, new, next)) with new = initial_command(rng) */
    int n = ci[0];
    REAL step_id = st[2 * n + 2] + 1;
    if (step_id > st[2 * n + 1]) cmd_initial(t, e, idx, rng);
    else st[2 * n + 2] = step_id;
  } else if (type == CMD_UNIFIED) {
    /* train.py:758-775 */
    REAL fresh[16];
    int sw = key_bernoulli(ra, 0, cf[32]);
    cmd_unified_initial(cf, rb, fresh);
    if (sw) for (int i = 0; i < 16; i++) st[i] = fresh[i];
  }
}

/* ---------------------------------------------------------------------------------------- observations */

static int geoms_colliding(const OData* d, const int* g1, int n1, const int* g2, int n2) {
  /* ksim/utils/mujoco.py:175-217, followed by `.any()` over all pairs */
  for (int c = 0; c < d->ncon; c++) {
    if (!(d->contact[c].dist < 0)) continue;
    int a = d->contact[c].geom1, b = d->contact[c].geom2;
    for (int i = 0; i < n1; i++)
      for (int j = 0; j < n2; j++)
        if ((g1[i] == a && g2[j] == b) || (g1[i] == b && g2[j] == a)) return 1;
  }
  return 0;
}

static void obs_carry_initial(const Task* t, Env* e, int idx, Key rng) {
  int type = TAB(t, TI_OBS_TAB, idx, TAB_TYPE);
  int coff = TAB(t, TI_OBS_TAB, idx, TAB_AUX1);
  if (coff < 0) return;
  const REAL* of = t->tf + TAB(t, TI_OBS_TAB, idx, TAB_FOFF) + NOISE_NF + BIAS_NF;
  REAL* c = e->obs_carry + coff;
  if (type == OBS_PROJECTED_GRAVITY) {
    /* observation.py:412-416; carry = x[3], lag, bias[3] */
    Key lrng = key_split(rng, 0), brng = key_split(rng, 1);
    c[0] = c[1] = c[2] = 0;
    c[3] = key_uniform(lrng, 0, of[3], of[4]);
    for (int k = 0; k < 3; k++) c[4 + k] = key_uniform(brng, k, -of[5], of[5]);
  } else if (type == OBS_DELAYED_JOINT_POSITION || type == OBS_DELAYED_JOINT_VELOCITY) {
    /* observation.py:232-236 / 264-268: jnp.tile(current[None, :], (delay_steps, 1)) */
    int dim = TAB(t, TI_OBS_TAB, idx, TAB_OUT_DIM);
    int delay = (t->ti + TAB(t, TI_OBS_TAB, idx, TAB_IOFF) + NOISE_NI + BIAS_NI)[0];
    for (int r = 0; r < delay; r++)
      for (int k = 0; k < dim; k++) c[r * dim + k] = type == OBS_DELAYED_JOINT_POSITION ? e->d.qpos[7 + k] : e->d.qvel[6 + k];
  }
}

/* get_observation (rl.py:153-186): writes obs block [OBS_SIZE] */
static void observe(const Task* t, Env* e, Key rng, REAL* out) {
  const OModel* m = &e->model;
  const OData* d = &e->d;
  int nobs = t->ti[TI_N_OBS];
  for (int o = 0; o < nobs; o++) {
    Key noise_rng = key_split(rng, 2);
    /* obs_rng = split(rng,3)[1] is unused by every supported observe() */
    rng = key_split(rng, 0);
    int type = TAB(t, TI_OBS_TAB, o, TAB_TYPE);
    const int* oi_all = t->ti + TAB(t, TI_OBS_TAB, o, TAB_IOFF);
    const REAL* of_all = t->tf + TAB(t, TI_OBS_TAB, o, TAB_FOFF);
    const int* oi = oi_all + NOISE_NI + BIAS_NI;
    const REAL* of = of_all + NOISE_NF + BIAS_NF;
    int off = TAB(t, TI_OBS_TAB, o, TAB_OUT_OFF), dim = TAB(t, TI_OBS_TAB, o, TAB_OUT_DIM);
    int noff = TAB(t, TI_OBS_TAB, o, TAB_AUX0), coff = TAB(t, TI_OBS_TAB, o, TAB_AUX1);
    REAL* v = out + off;
    switch (type) {
      case OBS_BASE_POSITION: for (int k = 0; k < 3; k++) v[k] = d->qpos[k]; break;
      case OBS_BASE_ORIENTATION: for (int k = 0; k < 4; k++) v[k] = d->qpos[3 + k]; break;
      case OBS_BASE_LINEAR_VELOCITY:
        if (oi[0]) o_rot_vec_quat(v, d->qvel, d->qpos + 3, 1); else for (int k = 0; k < 3; k++) v[k] = d->qvel[k];
        break;
      case OBS_BASE_ANGULAR_VELOCITY:
        if (oi[0]) o_rot_vec_quat(v, d->qvel + 3, d->qpos + 3, 1); else for (int k = 0; k < 3; k++) v[k] = d->qvel[3 + k];
        break;
      case OBS_JOINT_POSITION: for (int k = 0; k < dim; k++) v[k] = d->qpos[7 + k] - e->zero_off[k]; break;
      case OBS_JOINT_VELOCITY: for (int k = 0; k < dim; k++) v[k] = d->qvel[6 + k]; break;
      case OBS_CENTER_OF_MASS_INERTIA: for (int k = 0; k < dim; k++) v[k] = d->cinert[1 + k / 10][k % 10]; break;
      case OBS_CENTER_OF_MASS_VELOCITY: for (int k = 0; k < dim; k++) v[k] = d->cvel[1 + k / 6][k % 6]; break;
      case OBS_ACTUATOR_FORCE: for (int k = 0; k < dim; k++) v[k] = d->actuator_force[k]; break;
      case OBS_SENSOR: for (int k = 0; k < dim; k++) v[k] = d->sensordata[oi[0] + k]; break;
      case OBS_BASE_LINEAR_ACCELERATION: for (int k = 0; k < 3; k++) v[k] = d->qacc[k]; break;
      case OBS_BASE_ANGULAR_ACCELERATION: for (int k = 0; k < 3; k++) v[k] = d->qacc[3 + k]; break;
      case OBS_ACTUATOR_ACCELERATION: for (int k = 0; k < dim; k++) v[k] = d->qacc[6 + k]; break;
      case OBS_PROJECTED_GRAVITY: {
        /* observation.py:418-442 */
        REAL* c = e->obs_carry + coff;
        REAL bq[4], pg[3], pg2[3];
        o_euler_to_quat(bq, c + 4);
        o_rot_vec_quat(pg, of, d->sensordata + oi[0], 1);
        o_rot_vec_quat(pg2, pg, bq, 0);
        for (int k = 0; k < 3; k++) { v[k] = c[k] * c[3] + pg2[k] * ((REAL)1 - c[3]); c[k] = v[k]; }
      } break;
      case OBS_FEET_CONTACT: {
        /* observation.py:515-521: stack([any_floor(left geoms), any_floor(right geoms)], -1) -> (rows, 2) */
        const int* ids = oi + 3;
        for (int r = 0; r < oi[0]; r++) {
          v[2 * r + 0] = (REAL)geoms_colliding(d, ids + r, 1, ids + oi[0] + oi[1], oi[2]);
          v[2 * r + 1] = (REAL)geoms_colliding(d, ids + oi[0] + r, 1, ids + oi[0] + oi[1], oi[2]);
        }
      } break;
      case OBS_BODY_POSITION: for (int b = 0; b < oi[0]; b++) for (int k = 0; k < 3; k++) v[3 * b + k] = d->xpos[oi[1 + b]][k]; break;
      case OBS_BODY_VELOCITY: for (int b = 0; b < oi[0]; b++) for (int k = 0; k < 6; k++) v[6 * b + k] = d->cvel[oi[1 + b]][k]; break;
      case OBS_FEET_FORCE: for (int s = 0; s < 2; s++) for (int k = 0; k < 3; k++) v[3 * s + k] = d->cfrc_ext[oi[s]][k]; break;
      case OBS_FEET_TORQUE: for (int s = 0; s < 2; s++) for (int k = 0; k < 3; k++) v[3 * s + k] = d->cfrc_ext[oi[s]][3 + k]; break;
      case OBS_TIMESTEP: v[0] = d->time; break;
      case OBS_FEET_POSITION: {
        /* train.py:672-689; ints [base, left, right] */
        REAL rpy[3], yawq[4], e0[3] = {0, 0, 0};
        quat_to_euler(rpy, d->xquat[oi[0]]);
        e0[2] = rpy[2];
        o_euler_to_quat(yawq, e0);
        for (int f = 0; f < 2; f++) {
          REAL rel[3];
          for (int k = 0; k < 3; k++) rel[k] = d->xpos[oi[1 + f]][k] - d->xpos[oi[0]][k];
          o_rot_vec_quat(v + 3 * f, rel, yawq, 1);
        }
      } break;
      case OBS_BASE_HEIGHT: v[0] = d->xpos[1][2]; break;
      case OBS_COM_DISTANCE: {
        /* train.py:627-649; ints [enough distinct geom2 values (static: MJX keeps every candidate contact)] */
        if (!oi[0]) { v[0] = -1; break; }
        REAL pts[O_MAXCON][2];
        for (int c = 0; c < d->ncon; c++) {
          int floor = d->contact[c].geom1 == 0;
          pts[c][0] = floor ? d->contact[c].pos[0] : 0;
          pts[c][1] = floor ? d->contact[c].pos[1] : 0;
        }
        v[0] = com_distance((const REAL (*)[2])pts, d->ncon, d->subtree_com[2]);
      } break;
      case OBS_DELAYED_JOINT_POSITION: case OBS_DELAYED_JOINT_VELOCITY: {
        /* observation.py:238-252 / 270-283: reading = carry[0] (- zero_offset); carry = roll(carry, -1).at[-1].set(current) */
        REAL* c = e->obs_carry + coff;
        int delay = oi[0], pos = type == OBS_DELAYED_JOINT_POSITION;
        for (int k = 0; k < dim; k++) v[k] = pos ? c[k] - e->zero_off[k] : c[k];
        for (int r = 0; r + 1 < delay; r++) for (int k = 0; k < dim; k++) c[r * dim + k] = c[(r + 1) * dim + k];
        for (int k = 0; k < dim; k++) c[(delay - 1) * dim + k] = pos ? d->qpos[7 + k] : d->qvel[6 + k];
      } break;
      case OBS_CONTACT:
        /* observation.py:474-478: geoms_colliding(data, idxs, idxs).any(axis=-1) */
        for (int k = 0; k < oi[0]; k++) v[k] = (REAL)geoms_colliding(d, oi + 1 + k, 1, oi + 1, oi[0]);
        break;
      case OBS_BODY_ORIENTATION: for (int b = 0; b < oi[0]; b++) for (int k = 0; k < 4; k++) v[4 * b + k] = d->xquat[oi[1 + b]][k]; break;
      case OBS_SITE_ORIENTATION:
        /* observation.py:666-669; xax.rotation_matrix_to_rotation6d (un-vendored) restated as rows 0 and 1 of the matrix */
        for (int b = 0; b < oi[0]; b++) for (int k = 0; k < 6; k++) v[6 * b + k] = d->site_xmat[oi[1 + b]][k];
        break;
      case OBS_ACT_POS: v[0] = e->last_ctrl[oi[0]]; v[1] = d->qpos[oi[1]]; break;
      default: for (int k = 0; k < dim; k++) v[k] = 0; break;
    }
    if (noff >= 0) {
      /* Observation.add_noise (observation.py:110-142): noise with this step's key, bias with the episode key */
      const int* bi = oi_all + NOISE_NI;
      const REAL* bf = of_all + NOISE_NF;
      for (int k = 0; k < dim; k++) {
        REAL x = noise_apply(oi_all, of_all, v[k], noise_rng, k, e->level);
        if (bi[0] != RV_ZERO) {
          REAL b = rv_sample(bi[0], bf, e->bias_key[o], k, e->level);
          x = bi[1] ? x * b : x + b;
        }
        out[noff + k] = x;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------- terminations */

static int termination(const Task* t, const Env* e, int idx) {
  int type = TAB(t, TI_TERM_TAB, idx, TAB_TYPE);
  const int* xi = t->ti + TAB(t, TI_TERM_TAB, idx, TAB_IOFF);
  const REAL* xf = t->tf + TAB(t, TI_TERM_TAB, idx, TAB_FOFF);
  const OData* d = &e->d;
  switch (type) {
    case TERM_NOT_UPRIGHT: {
      REAL g[3] = {0, 0, 1}, r[3];
      o_rot_vec_quat(r, g, d->qpos + 3, 1);
      return ACOS(r[2]) > xf[0] ? -1 : 0;
    }
    case TERM_MINIMUM_HEIGHT: return d->qpos[2] < xf[0] ? -1 : 0;
    case TERM_ILLEGAL_CONTACT: {
      for (int c = 0; c < d->ncon; c++) {
        int hit = 0;
        for (int k = 0; k < xi[0]; k++) if (d->contact[c].geom1 == xi[1 + k] || d->contact[c].geom2 == xi[1 + k]) hit = 1;
        if (hit && d->contact[c].dist < xf[0]) return -1;
      }
      return 0;
    }
    case TERM_BAD_Z: {
      REAL lo = (xf[1] - xf[0]) * e->level + xf[0], hi = (xf[3] - xf[2]) * e->level + xf[2];
      return (d->qpos[2] < lo || d->qpos[2] > hi) ? -1 : 0;
    }
    case TERM_BAD_VELOCITY: {
      REAL n = SQRT(d->qvel[0] * d->qvel[0] + d->qvel[1] * d->qvel[1] + d->qvel[2] * d->qvel[2]);
      return n > xf[0] ? -1 : 0;
    }
    case TERM_FAR_FROM_ORIGIN: {
      REAL n = SQRT(d->qpos[0] * d->qpos[0] + d->qpos[1] * d->qpos[1] + d->qpos[2] * d->qpos[2]);
      return n > xf[0] ? xi[0] : 0;
    }
    case TERM_TERRAIN_BAD_Z: {
      /* train.py:807-813; ints [base, left, right] */
      REAL lf = d->xpos[xi[1]][2], rf_ = d->xpos[xi[2]][2];
      REAL height = d->xpos[xi[0]][2] - (lf < rf_ ? lf : rf_);
      return height < xf[0] ? -1 : 0;
    }
    case TERM_EPISODE_LENGTH: {
      int v = d->time > xf[0] ? xi[0] : 0;
      if (xi[1] >= 0 && e->level < (REAL)xi[1]) return 0;
      return v;
    }
  }
  return 0;
}

/* -------------------------------------------------------------------------------- fused control step */

static void write_pre_block(const Task* t, Env* e, REAL* rec_next, uint32_t* act_key) {
  /* step_engine (rl.py:1180) + _sample_action (rl.py:967): observation for the NEXT policy call */
  Key action_rng = key_split(e->rng, 0);
  Key obs_rng = key_split(action_rng, 0), act_rng = key_split(action_rng, 1);
  if (act_key) { act_key[0] = act_rng.k0; act_key[1] = act_rng.k1; }
  const int* ti = t->ti;
  observe(t, e, obs_rng, rec_next + ti[TI_R_OBS]);
  memcpy(rec_next + ti[TI_R_CMD], e->cmd, ti[TI_CMD_SIZE] * sizeof(REAL));
  rec_next[ti[TI_R_LEVEL]] = e->level;
}

static void do_reset(const Task* t, Env* e, Key reset_rng, Key cmd_rng, Key obs_carry_rng, Key obs_bias_rng) {
  engine_reset(t, e, reset_rng);
  for (int c = 0; c < t->ti[TI_N_CMD]; c++) {
    Key k = key_split(cmd_rng, 1);
    cmd_rng = key_split(cmd_rng, 0);
    cmd_initial(t, e, c, k);
  }
  int nobs = t->ti[TI_N_OBS];
  for (int o = 0; o < nobs; o++) {
    obs_carry_initial(t, e, o, key_split(obs_carry_rng, o));
    e->bias_key[o] = key_split(obs_bias_rng, o);
  }
}

/* one fused control step of one env: engine.step -> terminations -> record -> reset-on-done -> commands
 * -> next observation.  rl.py:1164-1211 (step_engine) with the observation of step t+1 hoisted to the end
 * of step t (same keys, same values; it only changes WHEN it is evaluated). */
static void env_step(const Task* t, Env* e, const REAL* action_in, REAL* rec_cur, REAL* rec_next, uint32_t* act_key) {
  const int* ti = t->ti;
  const OModel* m = &e->model;
  OData* d = &e->d;
  const REAL* ef = t->tf + ti[TI_ENGINE_FOFF];
  int na = ti[TI_ACTION_SIZE];
  REAL action[2 * O_MAXNU];
  for (int i = 0; i < na; i++) {
    REAL a = action_in[i];
    if (a != a) a = 0; /* NaN -> 0, then clip (rl.py:998-1001) */
    if (a < -ef[6]) a = -ef[6];
    if (a > ef[6]) a = ef[6];
    action[i] = a;
  }
  Key physics_rng = key_split(e->rng, 1), state_rng = key_split(e->rng, 2);
  engine_step(t, e, action, physics_rng);

  Key cmd_rng = key_split(state_rng, 0), obs_carry_rng = key_split(state_rng, 2);
  Key obs_bias_rng = key_split(state_rng, 3), reset_rng = key_split(state_rng, 4), next_rng = key_split(state_rng, 5);

  int done = 0, all_not_fail = 1, nterm = ti[TI_N_TERM];
  for (int k = 0; k < nterm; k++) {
    int v = termination(t, e, k);
    rec_cur[ti[TI_R_TERM] + k] = (REAL)v;
    done |= (v != 0);
    all_not_fail &= (v != -1);
  }
  int success = all_not_fail && done;
  int nonfinite = 0;
  for (int i = 0; i < m->nq; i++) if (!isfinite((double)d->qpos[i])) nonfinite = 1;
  for (int i = 0; i < m->nv; i++) if (!isfinite((double)d->qvel[i])) nonfinite = 1;
  if (nonfinite) e->nonfinite = 1;

  memcpy(rec_cur + ti[TI_R_QPOS], d->qpos, m->nq * sizeof(REAL));
  memcpy(rec_cur + ti[TI_R_QVEL], d->qvel, m->nv * sizeof(REAL));
  for (int b = 0; b < m->nbody; b++) {
    for (int k = 0; k < 3; k++) rec_cur[ti[TI_R_XPOS] + 3 * b + k] = d->xpos[b][k];
    for (int k = 0; k < 4; k++) rec_cur[ti[TI_R_XQUAT] + 4 * b + k] = d->xquat[b][k];
  }
  memcpy(rec_cur + ti[TI_R_CTRL], d->ctrl, m->nu * sizeof(REAL));
  memcpy(rec_cur + ti[TI_R_EVENT], e->event, ti[TI_EVENT_SIZE] * sizeof(REAL));
  memcpy(rec_cur + ti[TI_R_ACTION], action, na * sizeof(REAL));
  rec_cur[ti[TI_R_DONE]] = (REAL)done;
  rec_cur[ti[TI_R_SUCCESS]] = (REAL)success;
  rec_cur[ti[TI_R_TIME]] = d->time;

  if (done) {
    do_reset(t, e, reset_rng, cmd_rng, obs_carry_rng, obs_bias_rng);
  } else {
    for (int c = 0; c < ti[TI_N_CMD]; c++) {
      Key k = key_split(cmd_rng, 1);
      cmd_rng = key_split(cmd_rng, 0);
      cmd_update(t, e, c, k);
    }
  }
  e->rng = next_rng;
  write_pre_block(t, e, rec_next, act_key);
}

/* ------------------------------------------------------------------------------------- batch entry points */

/* init keys per env: [reset, cmd, obs_carry, obs_bias, env_rng] x 2 words  (rl.py:2175-2272 hands each env
 * one key per purpose; apply_randomizations' reset key = split(rand_key)[1], rl.py:368) */
int o_init_envs(const OModel* base, const int* ti, const REAL* tf, int N, const uint32_t* init_keys, const REAL* levels,
                const REAL* rand, REAL* state, uint32_t* keys, REAL* rec0, uint32_t* act_keys) {
  Task t;
  task_init(&t, ti, tf);
  if (ti[TI_MAGIC] != B2S_TASK_MAGIC) return B2S_ERR_BAD_BLOB;
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    Env* e = (Env*)calloc(1, sizeof(Env));
    apply_overrides(&t, base, &e->model, rand + (size_t)n * t.RAND);
    e->level = levels[n];
    const uint32_t* ik = init_keys + (size_t)n * 10;
    Key kr = {ik[0], ik[1]}, kc = {ik[2], ik[3]}, ko = {ik[4], ik[5]}, kb = {ik[6], ik[7]};
    do_reset(&t, e, kr, kc, ko, kb);
    e->rng.k0 = ik[8]; e->rng.k1 = ik[9];
    write_pre_block(&t, e, rec0 + (size_t)n * t.R, act_keys ? act_keys + 2 * n : 0);
    env_pack(&t, e, state + (size_t)n * t.S, keys + (size_t)n * t.K);
    free(e);
  }
  return B2S_OK;
}

int o_step_ctrl(const OModel* base, const int* ti, const REAL* tf, int N, const REAL* action, const REAL* rand,
                REAL* state, uint32_t* keys, REAL* rec_cur, REAL* rec_next, uint32_t* act_keys) {
  Task t;
  task_init(&t, ti, tf);
  if (ti[TI_MAGIC] != B2S_TASK_MAGIC) return B2S_ERR_BAD_BLOB;
  int na = ti[TI_ACTION_SIZE];
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    Env* e = (Env*)calloc(1, sizeof(Env));
    apply_overrides(&t, base, &e->model, rand + (size_t)n * t.RAND);
    env_unpack(&t, e, state + (size_t)n * t.S, keys + (size_t)n * t.K);
    env_step(&t, e, action + (size_t)n * na, rec_cur + (size_t)n * t.R, rec_next + (size_t)n * t.R,
             act_keys ? act_keys + 2 * n : 0);
    env_pack(&t, e, state + (size_t)n * t.S, keys + (size_t)n * t.K);
    free(e);
  }
  return B2S_OK;
}

/* ---- the engine alone (PhysicsEngine.reset / .step, ksim/engine.py:205-361), the checker of b200sim_engine_reset / _step and
 * the entry point MJX fixtures are compared through (scripts/gen_mjx_fixtures.py): `rng` [N][2] is the key the caller hands to
 * engine.reset / engine.step; `data` [N][DATA] uses the B2S_DATA_* block layout of include/b200sim_codes.h. */
static int data_offsets(const OModel* m, int ncon, int* off) {
  const int size[B2S_DATA_N_FIELDS] = {m->nq, m->nv, m->nv, m->nu, 1, 3 * m->nbody, 4 * m->nbody, 10 * m->nbody, 6 * m->nbody, m->nu,
                                       m->nsensordata, 6 * m->nbody, 3 * m->nsite, 9 * m->nsite, 3 * m->nbody, 6 * m->nbody, 1, ncon, ncon,
                                       ncon, 3 * ncon};
  int o = 0;
  for (int i = 0; i < B2S_DATA_N_FIELDS; i++) { off[i] = o; o += size[i]; }
  off[B2S_DATA_N_FIELDS] = o;
  return o;
}
static void write_data(const OModel* m, const OData* d, int ncon, REAL* out) {
  int off[B2S_DATA_N_FIELDS + 1];
  data_offsets(m, ncon, off);
  memcpy(out + off[B2S_DATA_QPOS], d->qpos, m->nq * sizeof(REAL));
  memcpy(out + off[B2S_DATA_QVEL], d->qvel, m->nv * sizeof(REAL));
  memcpy(out + off[B2S_DATA_QACC], d->qacc, m->nv * sizeof(REAL));
  memcpy(out + off[B2S_DATA_CTRL], d->ctrl, m->nu * sizeof(REAL));
  out[off[B2S_DATA_TIME]] = d->time;
  memcpy(out + off[B2S_DATA_ACTUATOR_FORCE], d->actuator_force, m->nu * sizeof(REAL));
  memcpy(out + off[B2S_DATA_SENSORDATA], d->sensordata, m->nsensordata * sizeof(REAL));
  for (int b = 0; b < m->nbody; b++) {
    memcpy(out + off[B2S_DATA_XPOS] + 3 * b, d->xpos[b], 3 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_XQUAT] + 4 * b, d->xquat[b], 4 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_CINERT] + 10 * b, d->cinert[b], 10 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_CVEL] + 6 * b, d->cvel[b], 6 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_CFRC_EXT] + 6 * b, d->cfrc_ext[b], 6 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_SUBTREE_COM] + 3 * b, d->subtree_com[b], 3 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_XFRC_APPLIED] + 6 * b, d->xfrc_applied[b], 6 * sizeof(REAL));
  }
  for (int s = 0; s < m->nsite; s++) {
    memcpy(out + off[B2S_DATA_SITE_XPOS] + 3 * s, d->site_xpos[s], 3 * sizeof(REAL));
    memcpy(out + off[B2S_DATA_SITE_XMAT] + 9 * s, d->site_xmat[s], 9 * sizeof(REAL));
  }
  out[off[B2S_DATA_NCON]] = (REAL)ncon;
  for (int c = 0; c < ncon; c++) {
    const int live = c < d->ncon;
    out[off[B2S_DATA_CONTACT_GEOM1] + c] = live ? (REAL)d->contact[c].geom1 : 0;
    out[off[B2S_DATA_CONTACT_GEOM2] + c] = live ? (REAL)d->contact[c].geom2 : 0;
    out[off[B2S_DATA_CONTACT_DIST] + c] = live ? d->contact[c].dist : 1;
    for (int k = 0; k < 3; k++) out[off[B2S_DATA_CONTACT_POS] + 3 * c + k] = live ? d->contact[c].pos[k] : 0;
  }
}
int o_data_size(const OModel* base, int ncon) {
  int off[B2S_DATA_N_FIELDS + 1];
  return data_offsets(base, ncon, off);
}

int o_engine_reset(const OModel* base, const int* ti, const REAL* tf, int N, const uint32_t* rng, const REAL* levels, const REAL* rand,
                   REAL* state, uint32_t* keys, int ncon, REAL* data) {
  Task t;
  task_init(&t, ti, tf);
  if (ti[TI_MAGIC] != B2S_TASK_MAGIC) return B2S_ERR_BAD_BLOB;
  const int DATA = o_data_size(base, ncon);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    Env* e = (Env*)calloc(1, sizeof(Env));
    apply_overrides(&t, base, &e->model, rand + (size_t)n * t.RAND);
    e->level = levels[n];
    Key k = {rng[2 * n], rng[2 * n + 1]};
    engine_reset(&t, e, k);
    write_data(&e->model, &e->d, ncon, data + (size_t)n * DATA);
    env_pack(&t, e, state + (size_t)n * t.S, keys + (size_t)n * t.K);
    free(e);
  }
  return B2S_OK;
}

int o_engine_step(const OModel* base, const int* ti, const REAL* tf, int N, const REAL* action, const uint32_t* rng, const REAL* rand,
                  REAL* state, uint32_t* keys, int ncon, REAL* data) {
  Task t;
  task_init(&t, ti, tf);
  if (ti[TI_MAGIC] != B2S_TASK_MAGIC) return B2S_ERR_BAD_BLOB;
  const int na = ti[TI_ACTION_SIZE], DATA = o_data_size(base, ncon);
  const REAL aclip = (t.tf + ti[TI_ENGINE_FOFF])[6];
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    Env* e = (Env*)calloc(1, sizeof(Env));
    apply_overrides(&t, base, &e->model, rand + (size_t)n * t.RAND);
    env_unpack(&t, e, state + (size_t)n * t.S, keys + (size_t)n * t.K);
    REAL act[O_MAXNU * 2];
    for (int i = 0; i < na; i++) {
      REAL a = action[(size_t)n * na + i];
      if (a != a) a = 0;
      act[i] = a < -aclip ? -aclip : (a > aclip ? aclip : a);
    }
    Key k = {rng[2 * n], rng[2 * n + 1]};
    engine_step(&t, e, act, k);
    write_data(&e->model, &e->d, ncon, data + (size_t)n * DATA);
    env_pack(&t, e, state + (size_t)n * t.S, keys + (size_t)n * t.K);
    free(e);
  }
  return B2S_OK;
}

/* ------------------------------------------------------------------------------------------- rewards */

static REAL exp_kernel_pen(REAL x, REAL scale, REAL sq, REAL ab) {
  return FABS(x) * -ab + x * x * -sq + EXP(-(x * x) / (2 * scale * scale));
}
static REAL exp_kernel(REAL x, REAL scale) { return EXP(-(x * x) / (2 * scale * scale)); }
static REAL norm_fn(REAL x, int l2) { return l2 ? x * x : FABS(x); }

/* get_rewards (rl.py:189-245) for one env over T records.  rec points at record 0 of this env; consecutive
 * time steps are `tstride` floats apart.  comps: [ncomp][T] (stride cstride between components). */
static void env_rewards(const Task* t, const REAL* rec, size_t tstride, int T, REAL level, REAL* carry,
                        REAL* total, REAL* comps, size_t cstride, int nq, int nv) {
  const int* ti = t->ti;
  const REAL* ef = t->tf + ti[TI_ENGINE_FOFF];
  int nrew = ti[TI_N_REW];
  int done_last = rec[(size_t)(T - 1) * tstride + ti[TI_R_DONE]] != 0;
#define REC(tt, off) rec[(size_t)(tt) * tstride + (off)]
  for (int r = 0; r < nrew; r++) {
    int type = TAB(t, TI_REW_TAB, r, TAB_TYPE);
    const int* xi = t->ti + TAB(t, TI_REW_TAB, r, TAB_IOFF) + SCALE_NI;
    const REAL* xf = t->tf + TAB(t, TI_REW_TAB, r, TAB_FOFF) + SCALE_NF;
    int c0 = TAB(t, TI_REW_TAB, r, TAB_OUT_OFF), ncomp = TAB(t, TI_REW_TAB, r, TAB_OUT_DIM);
    int coff = TAB(t, TI_REW_TAB, r, TAB_AUX0), cdim = TAB(t, TI_REW_TAB, r, TAB_AUX1);
    REAL scale = scale_get(t->ti[TAB(t, TI_REW_TAB, r, TAB_IOFF)], t->tf + TAB(t, TI_REW_TAB, r, TAB_FOFF), level);
    if (TAB(t, TI_REW_TAB, r, TAB_AUX2)) scale = -scale;
    REAL* out[8];
    for (int c = 0; c < ncomp; c++) out[c] = comps + (size_t)(c0 + c) * cstride;
    REAL* cr = coff >= 0 ? carry + coff : 0;
    if (cr && done_last) for (int k = 0; k < cdim; k++) cr[k] = 0; /* initial carry of every supported reward is 0 */
    int QP = ti[TI_R_QPOS], QV = ti[TI_R_QVEL], OB = ti[TI_R_OBS], CM = ti[TI_R_CMD], AC = ti[TI_R_ACTION];
    int DN = ti[TI_R_DONE], SC = ti[TI_R_SUCCESS];
    int na = ti[TI_ACTION_SIZE];
    switch (type) {
      case REW_AMP: /* ksim/task/amp.py:100-112 */
        for (int s = 0; s < T; s++) {
          REAL dd = REC(s, ti[TI_R_AUX] + xi[0]) - 1;
          REAL v = 1 - (REAL)0.25 * dd * dd;
          out[0][s] = v > 0 ? v : 0;
        }
        break;
      case REW_POSITION_TRACKING: {
        /* rewards.py:715-721; ints [tracked body, base body, command offset, l2]; floats [kernel_scale] */
        int XP = ti[TI_R_XPOS];
        for (int s = 0; s < T; s++) {
          REAL err = 0;
          for (int k = 0; k < 3; k++) {
            REAL v = (REC(s, XP + 3 * xi[0] + k) - REC(s, XP + 3 * xi[1] + k)) - REC(s, CM + xi[2] + k);
            err += xi[3] ? v * v : FABS(v);
          }
          out[0][s] = exp_kernel(err, xf[0]);
        }
      } break;
      case REW_SINUSOIDAL_GAIT:
        /* rewards.py:1336-1340; ints [position obs offset, num_feet, command height offset]; floats [max_height] */
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int f = 0; f < xi[1]; f++) acc += FABS(REC(s, OB + xi[0] + 3 * f + 2) - REC(s, CM + xi[2] + f));
          out[0][s] = (REAL)1 - acc / xf[0];
        }
        break;
      case REW_JOINT_POSITION: {
        /* rewards.py:1370-1385, commands.py:806-820; ints [n, command offset, indices]; floats [pos_scale, vel_scale, sq, abs] */
        int n = xi[0], c = CM + xi[1];
        for (int s = 0; s < T; s++) {
          REAL total_time = REC(s, c + 2 * n), num_steps = REC(s, c + 2 * n + 1), step_id = REC(s, c + 2 * n + 2);
          REAL ap = 0, av = 0;
          for (int i = 0; i < n; i++) {
            REAL st0 = REC(s, c + i), dd = REC(s, c + n + i) - st0;
            REAL cpos = st0 + dd * step_id / num_steps, cvel = dd / total_time;
            ap += exp_kernel_pen(cpos - REC(s, QP + 7 + xi[2 + i]), xf[0], xf[2], xf[3]);
            av += exp_kernel_pen(cvel - REC(s, QV + 6 + xi[2 + i]), xf[1], xf[2], xf[3]);
          }
          out[0][s] = ap / (REAL)n;
          out[1][s] = av / (REAL)n;
        }
      } break;
      case REW_INTERSECTION:
        /* rewards.py:1424-1430: the upper triangle (incl. the diagonal) is pushed beyond the threshold */
        for (int s = 0; s < T; s++) {
          int close = 0;
          for (int i = 1; i < xi[1]; i++)
            for (int j = 0; j < i; j++) {
              REAL dx = REC(s, OB + xi[0] + 3 * i) - REC(s, OB + xi[0] + 3 * j);
              REAL dy = REC(s, OB + xi[0] + 3 * i + 1) - REC(s, OB + xi[0] + 3 * j + 1);
              REAL dz = REC(s, OB + xi[0] + 3 * i + 2) - REC(s, OB + xi[0] + 3 * j + 2);
              close |= SQRT(dx * dx + dy * dy + dz * dz) < xf[0];
            }
          out[0][s] = close ? 1 : 0;
        }
        break;
      case REW_STAY_ALIVE:
        for (int s = 0; s < T; s++) {
          REAL alive = (REAL)1 / xf[0];
          REAL v = REC(s, DN) != 0 ? (REC(s, SC) != 0 ? (xi[0] ? xf[1] : alive) : (REAL)-1) : alive;
          out[0][s] = v;
        }
        break;
      case REW_UPRIGHT:
        for (int s = 0; s < T; s++) {
          REAL z[3] = {0, 0, 1}, zb[3];
          o_rot_vec_quat(zb, z, &REC(s, QP + 3), 1);
          out[0][s] = exp_kernel_pen(REC(s, QV + 4), xf[0], xf[3], xf[4]);
          out[1][s] = exp_kernel_pen(REC(s, QV + 2), xf[1], xf[5], xf[6]);
          out[2][s] = exp_kernel_pen(ATAN2(zb[0], zb[2]), xf[2], xf[7], xf[8]);
        }
        break;
      case REW_LINEAR_VELOCITY:
        for (int s = 0; s < T; s++) {
          const REAL* cmd = &REC(s, CM + xi[0]);
          REAL lv[3];
          o_rot_vec_quat(lv, &REC(s, QV), &REC(s, QP + 3), 1);
          REAL vel = SQRT(lv[0] * lv[0] + lv[1] * lv[1]), yaw = ATAN2(lv[1], lv[0]);
          int is_zero = FABS(cmd[0]) < xf[2];
          out[0][s] = exp_kernel_pen(vel - cmd[0], xf[0], xf[3], xf[4]);
          out[1][s] = is_zero ? 0 : exp_kernel_pen(yaw - cmd[1], xf[1], xf[3], xf[4]);
          out[2][s] = is_zero ? 0 : exp_kernel_pen(lv[0] - cmd[2], xf[0], xf[3], xf[4]);
          out[3][s] = is_zero ? 0 : exp_kernel_pen(lv[1] - cmd[3], xf[0], xf[3], xf[4]);
        }
        break;
      case REW_ANGULAR_VELOCITY:
        for (int s = 0; s < T; s++) {
          REAL cv = REC(s, CM + xi[0]), av = REC(s, QV + 5);
          REAL rs = FABS(av) < xf[3] ? 0 : (av > 0 ? 1 : (av < 0 ? -1 : 0));
          REAL cs = FABS(cv) < xf[3] ? 0 : (cv > 0 ? 1 : (cv < 0 ? -1 : 0));
          out[0][s] = exp_kernel_pen(av - cv, xf[0], xf[1], xf[2]);
          out[1][s] = rs == cs ? 1 : -1;
        }
        break;
      case REW_FEET_AIR_TIME: {
        /* rewards.py:993-1085; ints [contact_off, nfeet, rows, linvel_cmd_off, angvel_cmd_off];
         * floats [gait_period scale(2 after kind in ints[5])..]: see host encoder */
        int coffs = xi[0], nf = xi[1], rows = xi[2], lcmd = xi[3], acmd = xi[4], gp_kind = xi[5];
        REAL air_pct = xf[0], ctrl_dt = xf[1], bias = xf[2], gpen = xf[4], lthr = xf[5], athr = xf[6];
        REAL gp = scale_get(gp_kind, xf + 7, level);
        for (int s = 0; s < T; s++) {
          /* gait_period uses trajectory.curriculum_level of each step */
          REAL lvl_s = REC(s, ti[TI_R_LEVEL]);
          gp = scale_get(gp_kind, xf + 7, lvl_s);
          REAL air_steps = RINT(gp * air_pct / ctrl_dt), gnd_steps = RINT(gp * ((REAL)1 - air_pct) / ctrl_dt);
          int mlin = lcmd < 0 ? (SQRT(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) > lthr)
                              : (FABS(REC(s, CM + lcmd)) > lthr);
          int mang = acmd < 0 ? (REC(s, QV + 5) > athr) : (FABS(REC(s, CM + acmd)) > athr);
          int moving = (mlin || mang) && !(REC(s, DN) != 0);
          REAL air_r = -INFINITY, gnd_r = -INFINITY, st_r = -INFINITY;
          for (int f = 0; f < nf; f++) {
            int contact = 0;
            for (int rr = 0; rr < rows; rr++) contact |= REC(s, OB + coffs + rr * nf + f) > (REAL)0.5;
            REAL air = cr[f], gnd = cr[nf + f], stg = cr[2 * nf + f];
            air = (!contact && moving) ? air + 1 : 0;
            gnd = (contact && moving) ? gnd + 1 : 0;
            stg = (contact && !moving) ? stg + 1 : 0;
            cr[f] = air; cr[nf + f] = gnd; cr[2 * nf + f] = stg;
            REAL a = air < air_steps ? air / air_steps + bias : 0;
            REAL g = gnd < gnd_steps ? gnd / gnd_steps + bias : -gpen;
            REAL q = stg / gnd_steps;
            q = clipr(q, 0, 1) + bias;
            if (a > air_r) air_r = a;
            if (g > gnd_r) gnd_r = g;
            if (q > st_r) st_r = q;
          }
          out[0][s] = air_r; out[1][s] = gnd_r; out[2][s] = st_r;
        }
      } break;
      case REW_SPARSE_TARGET_HEIGHT: {
        /* rewards.py:1149-1192; ints [contact_off, pos_off, nbodies, rows]; floats [height, lin_thr, ang_thr] */
        int coffs = xi[0], poff = xi[1], nb = xi[2], rows = xi[3];
        for (int s = 0; s < T; s++) {
          int nm_lin = SQRT(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) < xf[1];
          int nm_ang = REC(s, QV + 5) < xf[2];
          int not_moving = (nm_lin && nm_ang) || (REC(s, DN) != 0);
          REAL best = -INFINITY;
          for (int b = 0; b < nb; b++) {
            int contact = 0;
            for (int rr = 0; rr < rows; rr++) contact |= REC(s, OB + coffs + rr * nb + b) > (REAL)0.5;
            REAL h = REC(s, OB + poff + 3 * b + 2);
            REAL rw = contact ? cr[b] : 0;
            if (rw > xf[0]) rw = xf[0];
            REAL mh = cr[b] > h ? cr[b] : h;
            if (not_moving || contact) mh = 0;
            cr[b] = mh;
            if (rw > best) best = rw;
          }
          out[0][s] = best;
        }
      } break;
      case REW_FORCE_PENALTY: {
        /* rewards.py:1224-1239; ints [obs_off, rows, dim]; floats [ctrl_dt, expected_weight, bias] */
        for (int s = 0; s < T; s++) {
          REAL best = -INFINITY;
          for (int rr = 0; rr < xi[1]; rr++) {
            REAL n2 = 0;
            for (int k = 0; k < xi[2]; k++) { REAL v = REC(s, OB + xi[0] + rr * xi[2] + k); n2 += v * v; }
            REAL o = SQRT(n2) - xf[2];
            if (o < 0) o = 0;
            REAL p = (o / xf[1]) * (o / xf[1]);
            if (p > best) best = p;
          }
          out[0][s] = best;
        }
      } break;
      case REW_MOTIONLESS_AT_REST:
        for (int s = 0; s < T; s++) {
          int nm_lin = SQRT(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) < xf[0];
          int nm_ang = REC(s, QV + 5) < xf[1];
          REAL n2 = 0;
          for (int k = 6; k < nv; k++) n2 += REC(s, QV + k) * REC(s, QV + k);
          out[0][s] = (nm_lin && nm_ang) ? SQRT(n2) : 0;
        }
        break;
      case REW_ACTION_VELOCITY: case REW_ACTION_ACCELERATION: case REW_ACTION_JERK: {
        /* rewards.py:400-463 (edge-padded finite differences with done masking); ints [l2] */
        int order = type - REW_ACTION_VELOCITY + 1;
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < na; k++) {
            /* padded index p maps to time max(p - order, 0); done_pad[p] = done[max(p-order,0)], p < T+order-1 */
#define APAD(p) REC(((p) - order) > 0 ? ((p) - order) : 0, AC + k)
#define DPAD(p) (REC(((p) - order) > 0 ? ((p) - order) : 0, DN) != 0)
            /* vel[i] = done_pad[i] ? 0 : a_pad[i+1]-a_pad[i], i in [0, T+order-1) */
#define VEL(i) (DPAD(i) ? (REAL)0 : APAD((i) + 1) - APAD(i))
#define ACC(i) (DPAD((i) + 1) ? (REAL)0 : VEL((i) + 1) - VEL(i))
#define JRK(i) (DPAD((i) + 2) ? (REAL)0 : ACC((i) + 1) - ACC(i))
            REAL v = order == 1 ? VEL(s) : (order == 2 ? ACC(s) : JRK(s));
            acc += norm_fn(v, xi[0]);
          }
          REAL pen = acc / (REAL)na;
          out[0][s] = s >= order ? pen : 0;
        }
#undef APAD
#undef DPAD
#undef VEL
#undef ACC
#undef JRK
      } break;
      case REW_BASE_HEIGHT: for (int s = 0; s < T; s++) out[0][s] = exp_kernel(REC(s, QP + 2) - xf[0], xf[1]); break;
      case REW_BASE_HEIGHT_RANGE:
        for (int s = 0; s < T; s++) {
          REAL h = REC(s, QP + 2), lo = xf[0] - h, hi = h - xf[1];
          REAL mx = lo > hi ? lo : hi;
          if (mx < 0) mx = 0;
          REAL v = (REAL)1 - mx * xf[2];
          out[0][s] = v < 0 ? 0 : v;
        }
        break;
      case REW_OFF_AXIS_VELOCITY:
        for (int s = 0; s < T; s++) {
          out[0][s] = exp_kernel(REC(s, QV + 2), xf[0]);
          out[1][s] = exp_kernel(REC(s, QV + 4), xf[1]);
          out[2][s] = exp_kernel(REC(s, QV + 5), xf[1]);
        }
        break;
      case REW_SMALL_JOINT_VELOCITY:
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          int prev = s > 0 ? s - 1 : 0;
          int dn = REC(prev, DN) != 0;
          for (int k = 7; k < nq; k++) {
            REAL v = (dn || s == 0) ? 0 : REC(s, QP + k) - REC(prev, QP + k);
            acc += exp_kernel(v, xf[0]);
          }
          out[0][s] = acc / (REAL)(nq - 7);
        }
        break;
      case REW_SMALL_JOINT_ACCELERATION: case REW_SMALL_JOINT_JERK: {
        /* rewards.py:479-508: edge-padded second / third differences of qpos[7:] with done masking, exp kernel, mean over
         * joints; unlike the action penalties the first `order` steps are NOT masked out.  floats [kernel_scale] */
        int order = type == REW_SMALL_JOINT_ACCELERATION ? 2 : 3;
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 7; k < nq; k++) {
#define QPAD(p) REC(((p) - order) > 0 ? ((p) - order) : 0, QP + k)
#define DPAD(p) (REC(((p) - order) > 0 ? ((p) - order) : 0, DN) != 0)
#define VEL(i) (DPAD(i) ? (REAL)0 : QPAD((i) + 1) - QPAD(i))
#define ACC(i) (DPAD((i) + 1) ? (REAL)0 : VEL((i) + 1) - VEL(i))
#define JRK(i) (DPAD((i) + 2) ? (REAL)0 : ACC((i) + 1) - ACC(i))
            REAL v = order == 2 ? ACC(s) : JRK(s);
            acc += exp_kernel(v, xf[0]);
#undef QPAD
#undef VEL
#undef ACC
#undef JRK
          }
          out[0][s] = acc / (REAL)(nq - 7);
        }
      } break;
      case REW_LINK_ACCELERATION: case REW_LINK_JERK: {
        /* rewards.py:809-841: speed of every body but the world (norm of the edge-padded position difference), second /
         * third difference with done masking, get_norm, mean over bodies.  ints [l2, nbody] */
        int order = type == REW_LINK_ACCELERATION ? 2 : 3;
        int XP = ti[TI_R_XPOS], nb = xi[1];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int b = 1; b < nb; b++) {
#define TPAD(p) (((p) - order) > 0 ? ((p) - order) : 0)
#define PDIF(i, c) (REC(TPAD((i) + 1), XP + 3 * b + (c)) - REC(TPAD(i), XP + 3 * b + (c)))
#define VEL(i) (DPAD(i) ? (REAL)0 : SQRT(PDIF(i, 0) * PDIF(i, 0) + PDIF(i, 1) * PDIF(i, 1) + PDIF(i, 2) * PDIF(i, 2)))
#define ACC(i) (DPAD((i) + 1) ? (REAL)0 : VEL((i) + 1) - VEL(i))
#define JRK(i) (DPAD((i) + 2) ? (REAL)0 : ACC((i) + 1) - ACC(i))
            REAL v = order == 2 ? ACC(s) : JRK(s);
            acc += norm_fn(v, xi[0]);
#undef TPAD
#undef PDIF
#undef VEL
#undef ACC
#undef JRK
#undef DPAD
          }
          out[0][s] = acc / (REAL)(nb - 1);
        }
      } break;
      case REW_AVOID_LIMITS: {
        /* rewards.py:522-554; ints [n]; floats lo[n], hi[n] (+-inf for unlimited joints) */
        int n = xi[0];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < n; k++) {
            REAL q = REC(s, QP + 7 + k), a = q - xf[k], b = q - xf[n + k];
            acc += -(a < 0 ? a : (REAL)0) + (b > 0 ? b : (REAL)0);
          }
          out[0][s] = acc / (REAL)n;
        }
      } break;
      case REW_TORQUE: case REW_ENERGY: {
        /* rewards.py:557-593; ints [nu]; floats ctrl_scales[nu] */
        int n = xi[0], CT = ti[TI_R_CTRL];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < n; k++) {
            REAL c = REC(s, CT + k) / xf[k];
            if (type == REW_ENERGY) c = c * REC(s, QV + 6 + k);
            acc += FABS(c);
          }
          out[0][s] = acc / (REAL)n;
        }
      } break;
      case REW_JOINT_DEVIATION: {
        /* rewards.py:596-662; ints [l2, n, has_weights, idx[n]]; floats targets[n], weights[n] */
        int n = xi[1];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < n; k++) {
            REAL d = REC(s, QP + 7 + xi[3 + k]) - xf[k];
            if (xi[2]) d *= xf[n + k];
            acc += norm_fn(d, xi[0]);
          }
          out[0][s] = acc;
        }
      } break;
      case REW_FLAT_BODY: {
        /* rewards.py:666-700; ints [n, body[n]]; floats plane[3] */
        int n = xi[0], XQ = ti[TI_R_XQUAT];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < n; k++) {
            REAL q[4], u[3], pl[3] = {xf[0], xf[1], xf[2]};
            for (int c = 0; c < 4; c++) q[c] = REC(s, XQ + 4 * xi[1 + k] + c);
            o_rot_vec_quat(u, pl, q, 1);
            acc += u[0] * pl[0] + u[1] * pl[1] + u[2] * pl[2];
          }
          out[0][s] = acc / (REAL)n;
        }
      } break;
      case REW_NO_ROLL:
        /* rewards.py:781-805; floats [angvel_scale, roll_scale, angvel_sq, angvel_abs, roll_sq, roll_abs] */
        for (int s = 0; s < T; s++) {
          REAL q[4] = {REC(s, QP + 3), REC(s, QP + 4), REC(s, QP + 5), REC(s, QP + 6)}, z[3] = {0, 0, 1}, zb[3];
          o_rot_vec_quat(zb, z, q, 1);
          out[0][s] = exp_kernel_pen(REC(s, QV + 3), xf[0], xf[2], xf[3]);
          out[1][s] = exp_kernel_pen(-ATAN2(zb[1], zb[2]), xf[1], xf[4], xf[5]);
        }
        break;
      case REW_FEET_SLIP: {
        /* rewards.py:1089-1106; ints [contact_off, contact_rows, n_feet, velocity_off]; contact obs [rows][n], velocity obs [n][6] */
        int co = xi[0], rows = xi[1], nf = xi[2], vo = xi[3];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int f = 0; f < nf; f++) {
            int contact = 0;
            for (int c = 0; c < rows; c++) contact |= REC(s, OB + co + c * nf + f) > (REAL)0.5;
            REAL n2 = 0;
            for (int k = 0; k < 6; k++) { REAL v = REC(s, OB + vo + 6 * f + k); n2 += v * v; }
            if (contact) acc += SQRT(n2);
          }
          out[0][s] = acc;
        }
      } break;
      case REW_REACHABILITY: {
        /* rewards.py:970-989; ints [squared, n]; floats delta_max[n] */
        int n = xi[1];
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < n; k++) {
            REAL ex = FABS(REC(s, AC + k) - REC(s, QP + 7 + k)) - xf[k];
            if (ex < 0) ex = 0;
            acc += xi[0] ? ex * ex : ex;
          }
          out[0][s] = acc;
        }
      } break;
      case REW_SYMMETRY: {
        /* rewards.py:844-897 (stateful); ints [n, idx[n]]; floats [alpha, targets[n]]; carry[n] holds reward_carry - targets
         * (the reference's initial carry is the targets; every carry here starts and resets at 0) */
        int n = xi[0];
        for (int k = 0; k < n; k++) {
          REAL mean = 0;
          for (int s = 0; s < T; s++) mean += REC(s, QP + 7 + xi[1 + k]) - xf[1 + k];
          mean /= (REAL)T;
          cr[k] = (cr[k] + xf[1 + k]) * ((REAL)1 - xf[0]) + mean * xf[0] - xf[1 + k];
        }
        for (int s = 0; s < T; s++) {
          REAL acc = 0;
          for (int k = 0; k < n; k++) acc += (REC(s, QP + 7 + xi[1 + k]) - xf[1 + k]) * -(cr[k] + xf[1 + k]);
          out[0][s] = acc;
        }
      } break;
      case REW_PAIRWISE_SYMMETRY: {
        /* rewards.py:901-966 (stateful); ints [left, right, flipped]; floats [left_zero, right_zero, alpha, kernel, sq, abs];
         * carry[2] holds reward_carry - (left_zero, right_zero) */
        REAL zero[2] = {xf[0], xf[1]};
        REAL mean[2] = {0, 0};
        for (int s = 0; s < T; s++) {
          REAL l = REC(s, QP + 7 + xi[0]) - xf[0], r_ = REC(s, QP + 7 + xi[1]) - xf[1];
          if (xi[2]) r_ = -r_;
          mean[0] += l; mean[1] += r_;
        }
        for (int k = 0; k < 2; k++) cr[k] = (cr[k] + zero[k]) * ((REAL)1 - xf[2]) + (mean[k] / (REAL)T) * xf[2] - zero[k];
        for (int s = 0; s < T; s++) {
          REAL l = REC(s, QP + 7 + xi[0]) - xf[0], r_ = REC(s, QP + 7 + xi[1]) - xf[1];
          if (xi[2]) r_ = -r_;
          out[0][s] = exp_kernel_pen(l - r_, xf[3], xf[4], xf[5]);
          out[1][s] = l * -(cr[0] + zero[0]) + r_ * -(cr[1] + zero[1]);
        }
      } break;
      /* ---- the K-Bot task's own rewards (examples/kbot/train.py:128-496).  ZERO_CMD(s): |cmd[:3]| < 1e-3.
       * Foot contact: touch-sensor observation > 0.1.  Observations / commands are pre-step, qpos / qvel / xpos /
       * xquat post-step (types.py:51-54). */
#define UCMD(s, k) REC(s, CM + ucmd + (k))
#define ZERO_CMD(s) (SQRT(UCMD(s, 0) * UCMD(s, 0) + UCMD(s, 1) * UCMD(s, 1) + UCMD(s, 2) * UCMD(s, 2)) < (REAL)1e-3)
      case REW_KBOT_SINGLE_FOOT_CONTACT: {
        /* train.py:128-156; ints [left obs, right obs, cmd]; floats [ctrl_dt, grace]; carry [time since single contact] */
        int ucmd = xi[2];
        for (int s = 0; s < T; s++) {
          int l = REC(s, OB + xi[0]) > (REAL)0.1, r_ = REC(s, OB + xi[1]) > (REAL)0.1, zero = ZERO_CMD(s);
          REAL tm = (l != r_) ? 0 : cr[0] + xf[0];
          if (zero) tm = xf[1];
          cr[0] = tm;
          out[0][s] = zero ? 1 : (tm < xf[1] ? 1 : 0);
        }
      } break;
      case REW_KBOT_NO_CONTACT: {
        /* train.py:159-167 */
        int ucmd = xi[2];
        for (int s = 0; s < T; s++) {
          int l = REC(s, OB + xi[0]) > (REAL)0.1, r_ = REC(s, OB + xi[1]) > (REAL)0.1;
          out[0][s] = ZERO_CMD(s) ? 0 : ((l || r_) ? 0 : 1);
        }
      } break;
      case REW_KBOT_FEET_AIRTIME: {
        /* train.py:170-218; carry [airtime l, airtime r, NOT prev contact l, r] (initial carry: airtime 0, contact True) */
        int ucmd = xi[2];
        for (int s = 0; s < T; s++) {
          int dn = REC(s, DN) != 0;
          REAL rew = 0;
          for (int f = 0; f < 2; f++) {
            int contact = REC(s, OB + xi[f]) > (REAL)0.1;
            int prev_contact = cr[2 + f] == 0;
            int first = contact && !prev_contact && !dn;
            rew += (cr[f] - xf[1]) * (first ? 1 : 0);       /* airtime of the previous step */
            cr[f] = (contact || dn) ? 0 : cr[f] + xf[0];
            cr[2 + f] = contact ? 0 : 1;
          }
          out[0][s] = ZERO_CMD(s) ? 0 : rew;
        }
      } break;
      case REW_KBOT_ARM_POSITION: {
        /* train.py:221-270; ints [cmd, 10 qpos indices]; floats [error_scale, 10 biases] */
        int ucmd = xi[0];
        for (int s = 0; s < T; s++) {
          REAL err = 0;
          for (int k = 0; k < 10; k++) {
            REAL dlt = REC(s, QP + xi[1 + k]) - (UCMD(s, 6 + k) + xf[1 + k]);
            err += dlt * dlt;
          }
          out[0][s] = EXP(-err / xf[0]);
        }
      } break;
      case REW_KBOT_LINVEL_TRACKING: {
        /* train.py:273-299 */
        int ucmd = xi[0], XQ = ti[TI_R_XQUAT];
        for (int s = 0; s < T; s++) {
          REAL rpy[3], zq[4], cmdv[3] = {UCMD(s, 0), UCMD(s, 1), 0}, g[3];
          quat_to_euler(rpy, &REC(s, XQ + 4));
          rpy[0] = rpy[1] = 0;
          o_euler_to_quat(zq, rpy);
          o_rot_vec_quat(g, cmdv, zq, 0);
          REAL ex = REC(s, QV) - g[0], ey = REC(s, QV + 1) - g[1];
          REAL ve = SQRT(ex * ex + ey * ey);
          out[0][s] = EXP(-(ZERO_CMD(s) ? ve : ve * ve) / xf[0]);
        }
      } break;
      case REW_KBOT_ANGVEL_TRACKING: {
        /* train.py:302-313 */
        int ucmd = xi[0];
        for (int s = 0; s < T; s++) out[0][s] = EXP(-FABS(REC(s, QV + 5) - UCMD(s, 2)) / xf[0]);
      } break;
      case REW_KBOT_XY_ORIENTATION: {
        /* train.py:316-343; floats [error_scale, error_scale_zero_cmd] */
        int ucmd = xi[0], XQ = ti[TI_R_XQUAT];
        for (int s = 0; s < T; s++) {
          REAL rpy[3], q[4], ce[3] = {UCMD(s, 4), UCMD(s, 5), 0}, qc[4];
          quat_to_euler(rpy, &REC(s, XQ + 4));
          rpy[2] = 0;
          o_euler_to_quat(q, rpy);
          o_euler_to_quat(qc, ce);
          REAL dt_ = qc[0] * q[0] + qc[1] * q[1] + qc[2] * q[2] + qc[3] * q[3];
          out[0][s] = EXP(-((REAL)1 - dt_ * dt_) / (ZERO_CMD(s) ? xf[1] : xf[0]));
        }
      } break;
      case REW_KBOT_TERRAIN_BASE_HEIGHT: {
        /* train.py:346-401; ints [cmd, base, left, right]; floats [error_scale, standard_height, foot_origin_height] */
        int ucmd = xi[0], XP = ti[TI_R_XPOS];
        for (int s = 0; s < T; s++) {
          REAL lf = REC(s, XP + 3 * xi[2] + 2) - xf[2], rf_ = REC(s, XP + 3 * xi[3] + 2) - xf[2];
          REAL cur = REC(s, XP + 3 * xi[1] + 2) - (lf < rf_ ? lf : rf_);
          out[0][s] = EXP(-FABS(cur - (UCMD(s, 3) + xf[1])) / xf[0]);
        }
      } break;
      case REW_KBOT_FEET_ORIENTATION: {
        /* train.py:404-467; ints [cmd, left, right]; `is_rotating` is ONE flag per trajectory: the reference takes the
         * norm of the (T,) yaw-rate command column over axis -1, i.e. over time (train.py:464) */
        int ucmd = xi[0], XQ = ti[TI_R_XQUAT];
        REAL n2 = 0;
        for (int s = 0; s < T; s++) n2 += UCMD(s, 2) * UCMD(s, 2);
        int rotating = SQRT(n2) > (REAL)1e-3;
        for (int s = 0; s < T; s++) {
          REAL rpy[3];
          quat_to_euler(rpy, &REC(s, XQ + 4));
          REAL yaw = rpy[2], err = 0;
          for (int f = 0; f < 2; f++) {
            REAL se[3] = {f == 0 ? -PI_R / 2 : PI_R / 2, 0, yaw - PI_R}, sq[4];
            const REAL* fq = &REC(s, XQ + 4 * xi[1 + f]);
            if (rotating) {
              REAL fe[3], fq2[4];
              quat_to_euler(fe, fq);
              fe[2] = 0; se[2] = 0;
              o_euler_to_quat(fq2, fe);
              o_euler_to_quat(sq, se);
              REAL dt_ = sq[0] * fq2[0] + sq[1] * fq2[1] + sq[2] * fq2[2] + sq[3] * fq2[3];
              err += (REAL)1 - dt_ * dt_;
            } else {
              o_euler_to_quat(sq, se);
              REAL dt_ = sq[0] * fq[0] + sq[1] * fq[1] + sq[2] * fq[2] + sq[3] * fq[3];
              err += (REAL)1 - dt_ * dt_;
            }
          }
          out[0][s] = EXP(-err / xf[0]);
        }
      } break;
      case REW_KBOT_COM_DISTANCE: {
        /* train.py:470-484; ints [cmd, com_distance obs] */
        int ucmd = xi[0];
        for (int s = 0; s < T; s++) {
          REAL dist = REC(s, OB + xi[1]);
          out[0][s] = (dist >= 0 && ZERO_CMD(s)) ? EXP(-dist / xf[0]) : 0;
        }
      } break;
      case REW_KBOT_BASE_ACCELERATION: {
        /* train.py:487-497: edge-padded difference of qvel[:6], zeroed after a done */
        for (int s = 0; s < T; s++) {
          int prev = s > 0 ? s - 1 : 0;
          int dn = REC(prev, DN) != 0;
          REAL err = 0;
          for (int k = 0; k < 6; k++) err += (dn || s == 0) ? 0 : FABS(REC(s, QV + k) - REC(prev, QV + k));
          out[0][s] = EXP(-err / xf[0]);
        }
      } break;
#undef UCMD
#undef ZERO_CMD
      default: for (int c = 0; c < ncomp; c++) for (int s = 0; s < T; s++) out[c][s] = 0; break;
    }
    for (int c = 0; c < ncomp; c++) for (int s = 0; s < T; s++) out[c][s] *= scale;
  }
  int ncomp_total = ti[TI_N_REW_COMP];
  for (int s = 0; s < T; s++) {
    REAL sum = 0;
    for (int c = 0; c < ncomp_total; c++) sum += comps[(size_t)c * cstride + s];
    if (ef[7] == ef[7] && sum < ef[7]) sum = ef[7];
    if (ef[8] == ef[8] && sum > ef[8]) sum = ef[8];
    total[s] = sum;
  }
#undef REC
}

/* rec: [T][N][R]; levels: [N]; carry: [N][C]; total: [N][T]; comps: [K][N][T] */
int o_rewards(const OModel* base, const int* ti, const REAL* tf, int N, int T, const REAL* rec, const REAL* levels,
              REAL* carry, REAL* total, REAL* comps) {
  Task t;
  task_init(&t, ti, tf);
  int C = ti[TI_REW_CARRY_SIZE];
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++)
    env_rewards(&t, rec + (size_t)n * t.R, (size_t)N * t.R, T, levels[n], carry + (size_t)n * C, total + (size_t)n * T,
                comps + (size_t)n * T, (size_t)N * T, base->nq, base->nv);
  return B2S_OK;
}

/* ----------------------------------------------------------------------------------------------- GAE */

/* compute_ppo_inputs (ppo.py:54-121).  All arrays [B][T]. */
int o_gae(int B, int T, const REAL* values, const REAL* rewards, const uint8_t* dones, const uint8_t* successes,
          REAL gamma, REAL lam, int flags, REAL* adv, REAL* targets, REAL* returns) {
  for (int b = 0; b < B; b++) {
    const REAL* v = values + (size_t)b * T;
    const REAL* r = rewards + (size_t)b * T;
    REAL ret_next = 0, adv_next = 0;
    for (int s = T - 1; s >= 0; s--) {
      REAL v_next = s + 1 < T ? v[s + 1] : v[T - 1];
      REAL trunc = successes[(size_t)b * T + s] ? (REAL)1 : (REAL)0;
      REAL mask = dones[(size_t)b * T + s] ? (REAL)0 : (REAL)1;
      REAL br = ((REAL)1 - gamma) * r[s] + gamma * v[s] * trunc;
      REAL delta = br + gamma * v_next * mask - v[s];
      REAL ret = br + gamma * mask * ret_next;
      REAL a = delta + gamma * lam * mask * adv_next;
      returns[(size_t)b * T + s] = ret;
      adv[(size_t)b * T + s] = a;
      ret_next = ret; adv_next = a;
    }
  }
  size_t n = (size_t)B * T;
  for (size_t i = 0; i < n; i++) targets[i] = (flags & GAE_MONTE_CARLO_RETURNS) ? returns[i] : adv[i] + values[i];
  if (flags & GAE_NORMALIZE_ADVANTAGES) {
    REAL mean = 0;
    for (size_t i = 0; i < n; i++) mean += adv[i];
    mean /= (REAL)n;
    REAL var = 0;
    for (size_t i = 0; i < n; i++) var += (adv[i] - mean) * (adv[i] - mean);
    REAL std = SQRT(var / (REAL)n);
    if (std < (REAL)1e-6) std = (REAL)1e-6;
    for (size_t i = 0; i < n; i++) adv[i] = (adv[i] - mean) / std;
  }
  return B2S_OK;
}

/* ------------------------------------------------------------------------------------------ PPO loss */

/* compute_ppo_loss (ppo.py:149-246; clipped_value_loss :124-135), the mean over (B, T) of the summed terms
 * (_get_ppo_loss_and_metrics ppo.py:471-486, incl. the kl_scale factor) and the gradient of that mean with respect to the
 * new policy's log-probs / values -- what jax.grad derives at ppo.py:521-534; ppo_inputs are constants (stop_gradient,
 * ppo.py:121).  Gradient conventions at the measure-zero kinks: clip passes 1 strictly inside its range (lax.clamp),
 * min / max ties split evenly (lax.min / lax.max).  logp_* / entropy [B][T][A]; the rest [B][T]; entropy may be NULL.
 * losses[4][B][T] (policy, value, kl, entropy); out[5] = mean total, mean of each term; d_logp[B][T] is the gradient for
 * every logp_off[b][t][a]; d_values[B][T]. */
int o_ppo_loss(int B, int T, int A, const REAL* logp_on, const REAL* logp_off, const REAL* values_on, const REAL* values_off,
               const REAL* adv, const REAL* targets, const REAL* entropy, REAL clip_param, REAL value_loss_coef,
               REAL entropy_coef, REAL kl_coef, REAL log_clip_value, REAL kl_scale, int flags, REAL* losses, REAL* out,
               REAL* d_logp, REAL* d_values) {
  const size_t n = (size_t)B * T;
  double acc[5] = {0, 0, 0, 0, 0};
  const REAL inv_n = (REAL)1 / (REAL)n;
  for (size_t e = 0; e < n; e++) {
    REAL lr = 0, es = 0;
    for (int a = 0; a < A; a++) lr += logp_off[e * A + a] - logp_on[e * A + a];
    if (entropy) for (int a = 0; a < A; a++) es += entropy[e * A + a];
    const REAL ad = adv[e], vt = targets[e], von = values_on[e], voff = values_off[e];
    REAL lrc = lr < -log_clip_value ? -log_clip_value : (lr > log_clip_value ? log_clip_value : lr);
    const REAL ratio = EXP(lrc);
    const REAL lo = (REAL)1 - clip_param, hi = (REAL)1 + clip_param;
    const REAL rcl = ratio < lo ? lo : (ratio > hi ? hi : ratio);
    const REAL s1 = ratio * ad, s2 = rcl * ad;
    const REAL pol = -(s1 < s2 ? s1 : s2);
    const REAL dr_dlr = (lr > -log_clip_value && lr < log_clip_value) ? ratio : (REAL)0;
    const REAL g1 = -ad, g2 = (ratio > lo && ratio < hi) ? -ad : (REAL)0;
    const REAL dpol = (s1 < s2 ? g1 : (s2 < s1 ? g2 : (REAL)0.5 * (g1 + g2))) * dr_dlr;
    REAL val, dval;
    const REAL err = voff - vt;
    if (flags & PPO_CLIPPED_VALUE_LOSS) {
      const REAL d = voff - von;
      const REAL dc = d < -clip_param ? -clip_param : (d > clip_param ? clip_param : d);
      const REAL cerr = (von + dc) - vt;
      const REAL e2 = err * err, c2 = cerr * cerr;
      const REAL gc = (d > -clip_param && d < clip_param) ? cerr : (REAL)0;
      val = (REAL)0.5 * (e2 > c2 ? e2 : c2) * value_loss_coef;
      dval = value_loss_coef * (e2 > c2 ? err : (c2 > e2 ? gc : (REAL)0.5 * (err + gc)));
    } else {
      val = (REAL)0.5 * (err * err) * value_loss_coef;
      dval = value_loss_coef * err;
    }
    const REAL kl = (-lr * kl_coef) * kl_scale;
    const REAL ent = entropy ? -es * entropy_coef : (REAL)0;
    if (losses) { losses[e] = pol; losses[n + e] = val; losses[2 * n + e] = kl; losses[3 * n + e] = ent; }
    if (d_logp) d_logp[e] = (dpol - kl_coef * kl_scale) * inv_n;
    if (d_values) d_values[e] = dval * inv_n;
    acc[0] += (double)(((pol + val) + kl) + ent);
    acc[1] += (double)pol; acc[2] += (double)val; acc[3] += (double)kl; acc[4] += (double)ent;
  }
  for (int k = 0; k < 5; k++) out[k] = n ? (REAL)(acc[k] / (double)n) : (REAL)0;
  return B2S_OK;
}

/* ------------------------------------------------------------------------------------------ test hooks */

/* evaluates termination `idx` of the task program on a hand-built state (tests/test_golden_terminations.py);
 * contacts are given as (geom1, geom2, dist) triples through a one-pair-per-contact scratch model. */
int o_test_termination(const int* ti, const REAL* tf, int idx, const REAL* qpos, int nq, const REAL* qvel, int nv,
                       REAL time, REAL level, int ncon, const int* geom1, const int* geom2, const REAL* dist) {
  Task t;
  task_init(&t, ti, tf);
  Env* e = (Env*)calloc(1, sizeof(Env));
  for (int i = 0; i < nq && i < O_MAXNQ; i++) e->d.qpos[i] = qpos[i];
  for (int i = 0; i < nv && i < O_MAXNV; i++) e->d.qvel[i] = qvel[i];
  e->d.time = time;
  e->level = level;
  e->d.ncon = ncon < O_MAXCON ? ncon : O_MAXCON;
  for (int c = 0; c < e->d.ncon; c++) {
    e->d.contact[c].geom1 = geom1[c];
    e->d.contact[c].geom2 = geom2[c];
    e->d.contact[c].dist = dist[c];
  }
  int v = termination(&t, e, idx);
  free(e);
  return v;
}

/* exported for tests: UnifiedCommand.initial_command for one key; COMDistanceObservation's geometry for one point set */
void o_test_cmd_unified(const REAL* cf, const uint32_t* key, REAL* out) {
  Key k = {key[0], key[1]};
  cmd_unified_initial(cf, k, out);
}
REAL o_test_com_distance(const REAL* pts, int n, const REAL* com) {
  return com_distance((const REAL (*)[2])pts, n < O_MAXCON ? n : O_MAXCON, com);
}
