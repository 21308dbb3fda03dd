// b200sim: ksim's control-step semantics around the warp physics -- one warp, one environment.
//
// What each block replaces in the reference:
//   engine_step / engine_reset      MjxEngine.step / MjxEngine.reset            ksim/engine.py:205-361
//   actuators_*                     PositionActuators / TorqueActuators         ksim/actuators.py:70-238
//   event_*                         LinearPush / AngularPush / Jump / ForcePush ksim/events.py:59-365
//   reset_apply                     Reset plugins                               ksim/resets.py:93-354
//   cmd_*                           LinearVelocity / AngularVelocity / FloatVector commands  ksim/commands.py
//   observe                         get_observation + Observation.add_noise    ksim/task/rl.py:153-186, observation.py
//   termination                     Termination plugins                         ksim/terminations.py:61-192
//   env_step / env_init             step_engine + _get_step_env_state          ksim/task/rl.py:1023-1211
// The plugin lists arrive as the task program ti[]/tf[] (include/b200sim_codes.h).  Scalar, warp-uniform
// values (keys, flags) are computed redundantly by every lane; shared-memory writes of them go through lane 0.
#pragma once

#include "b2s_physics.cuh"

namespace b2s {

#define TAB(ti, base, i, f) ((ti)[(ti)[base] + (i) * TAB_STRIDE + (f)])

// ------------------------------------------------------------------------------------- HBM <-> workspace

template <class D>
DEV void apply_overrides(const Mdl& m, Work<D>& w, const float* rand, LANE_ARG) {
  const int nv = DIM(NV, MH_NV), nb = DIM(NB, MH_NBODY);
  const float* arm = MF(m, MH_DOF_ARMATURE);
  const float* damp = MF(m, MH_DOF_DAMPING);
  const float* fl = MF(m, MH_DOF_FRICTIONLOSS);
  const float* mass = MF(m, MH_BODY_MASS);
  const float* inertia = MF(m, MH_BODY_INERTIA);
  const float* ipos = MF(m, MH_BODY_IPOS);
  LANE_FOR(i, nv) { w.armature[i] = arm[i]; w.damping[i] = damp[i]; w.frictionloss[i] = fl[i]; }
  // the mass matrix is only ever written on its ancestor pattern (crb_mass_matrix), which is a property of the model:
  // zero it once per workspace instead of once per sub-step.  ALL D::NV rows, not only the model's nv: the register-resident
  // factor_solve loads one row per lane, and a stale NaN / Inf in a row beyond nv reaches the live lanes through the
  // unguarded 0 * x products of its Z substitutions (found with tools/debug_poison.py in the larger instantiations)
  LANE_FOR(i, D::NV) {
    for (int j = 0; j < D::NVP; j++) w.M[i][j] = 0.0f;
  }
  LANE_FOR(b, nb) {
    w.body_mass[b] = mass[b];
#pragma unroll
    for (int k = 0; k < 3; k++) { w.body_inertia[k][b] = inertia[3 * b + k]; w.body_ipos[k][b] = ipos[3 * b + k]; }
  }
  const int* pgeom[2] = {MI(m, MH_PAIR_GEOM1), MI(m, MH_PAIR_GEOM2)};
  if constexpr (!D::SMALL) {
    const float* gpos = MF(m, MH_GEOM_POS);
    const float* gsize = MF(m, MH_GEOM_SIZE);
    LANE_FOR(p, DIM(NPAIR, MH_NPAIR)) {
#pragma unroll
      for (int sd = 0; sd < 2; sd++)
#pragma unroll
        for (int k = 0; k < 3; k++) { w.pair_gpos[sd][k][p] = gpos[3 * pgeom[sd][p] + k]; w.pair_gsize[sd][k][p] = gsize[3 * pgeom[sd][p] + k]; }
    }
    const float* spos = MF(m, MH_SITE_POS);
    const float* squat = MF(m, MH_SITE_QUAT);
    LANE_FOR(s, DIM(NS, MH_NSITE)) {
#pragma unroll
      for (int k = 0; k < 3; k++) w.site_pos[k][s] = spos[3 * s + k];
#pragma unroll
      for (int k = 0; k < 4; k++) w.site_quat[k][s] = squat[4 * s + k];
    }
  }
  WSYNC();
  const int* ti = m.ti;
  const int n = ti[TI_N_RAND];
  int off = 0;
  for (int r = 0; r < n; r++) {
    const int code = TAB(ti, TI_RAND_TAB, r, TAB_AUX0), cnt = TAB(ti, TI_RAND_TAB, r, TAB_AUX1);
    if (code == RAND_GEOM_SIZE || code == RAND_GEOM_POS) {
      if constexpr (!D::SMALL) {
        LANE_FOR(p, DIM(NPAIR, MH_NPAIR)) {
#pragma unroll
          for (int sd = 0; sd < 2; sd++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const float v = rand[off + 3 * pgeom[sd][p] + k];
              if (code == RAND_GEOM_SIZE) w.pair_gsize[sd][k][p] = v; else w.pair_gpos[sd][k][p] = v;
            }
        }
      }
      off += cnt;
      continue;
    }
    if (code == RAND_SITE_QUAT || code == RAND_SITE_POS) {
      if constexpr (!D::SMALL) {
        LANE_FOR(k, cnt) {
          if (code == RAND_SITE_QUAT) w.site_quat[k & 3][k >> 2] = rand[off + k]; else w.site_pos[k % 3][k / 3] = rand[off + k];
        }
      }
      off += cnt;
      continue;
    }
    LANE_FOR(k, cnt) {
      const float v = rand[off + k];
      switch (code) {
        case RAND_DOF_FRICTIONLOSS: w.frictionloss[k] = v; break;
        case RAND_DOF_ARMATURE: w.armature[k] = v; break;
        case RAND_DOF_DAMPING: w.damping[k] = v; break;
        case RAND_BODY_MASS: w.body_mass[k] = v; break;
        case RAND_BODY_INERTIA: w.body_inertia[k % 3][k / 3] = v; break;
        case RAND_BODY_IPOS: w.body_ipos[k % 3][k / 3] = v; break;
      }
    }
    off += cnt;
  }
  WSYNC();
}

#define B2S_STATE_FIELDS(X)                                        \
  X(w.qpos, TI_S_QPOS, DIM(NQ, MH_NQ))                                 \
  X(w.qvel, TI_S_QVEL, DIM(NV, MH_NV))                                 \
  X(w.qacc, TI_S_QACC, DIM(NV, MH_NV))                                 \
  X(w.warm, TI_S_WARM, DIM(NV, MH_NV))                                 \
  X(w.ctrl, TI_S_CTRL, DIM(NU, MH_NU))                                 \
  X(w.last_ctrl, TI_S_LAST_CTRL, ti[TI_ACTION_SIZE])               \
  X(w.zero_off, TI_S_ZERO_OFF, DIM(NU, MH_NU))                         \
  X(w.event, TI_S_EVENT, ti[TI_EVENT_SIZE])                        \
  X(w.act_state, TI_S_ACT_STATE, ti[TI_ACT_STATE_SIZE])            \
  X(w.cmd, TI_S_CMD, ti[TI_CMD_SIZE])                              \
  X(w.obs_carry, TI_S_OBS_CARRY, ti[TI_OBS_CARRY_SIZE])

template <class D>
DEV void env_load(const Mdl& m, Work<D>& w, const float* s, const uint32_t* keys, LANE_ARG) {
  const int* ti = m.ti;
#define X(dst, off, n) LANE_FOR(i, n) dst[i] = s[ti[off] + i];
  B2S_STATE_FIELDS(X)
#undef X
  LANE_FOR(i, ti[TI_KEY_SIZE]) w.keys[i] = keys[i];
  IF_LANE0 {
    w.time = s[ti[TI_S_TIME]]; w.latency = s[ti[TI_S_LATENCY]]; w.level = s[ti[TI_S_LEVEL]];
    w.nonfinite = s[ti[TI_S_NONFINITE]];
    w.xfrc_body = -1;
#pragma unroll
    for (int k = 0; k < 6; k++) w.xfrc[k] = 0.0f;
  }
  WSYNC();
}

template <class D>
DEV void env_store(const Mdl& m, const Work<D>& w, float* s, uint32_t* keys, LANE_ARG) {
  const int* ti = m.ti;
#define X(src, off, n) LANE_FOR(i, n) s[ti[off] + i] = src[i];
  B2S_STATE_FIELDS(X)
#undef X
  LANE_FOR(i, ti[TI_KEY_SIZE]) keys[i] = w.keys[i];
  IF_LANE0 {
    s[ti[TI_S_TIME]] = w.time; s[ti[TI_S_LATENCY]] = w.latency; s[ti[TI_S_LEVEL]] = w.level;
    s[ti[TI_S_NONFINITE]] = w.nonfinite;
  }
}

// ---------------------------------------------------------------------------------------------- actuators

template <class D>
DEV void actuators_ctrl(const Mdl& m, Work<D>& w, const float* ctrl_in, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int nu = DIM(NU, MH_NU);
  const int* ai = ti + ti[TI_ACT_IOFF];
  const float* af = m.tf + ti[TI_ACT_FOFF];
  const float level = w.level;
  if (ti[TI_ACT_TYPE] == ACT_TORQUE) {
    LANE_FOR(i, nu) w.torque[i] = noise_apply(ai, af, ctrl_in[i], rng, i, level);
    WSYNC();
    return;
  }
  const int* an_i = ai; const int* tn_i = ai + NOISE_NI;
  const float action_scale = af[0];
  const float* an_f = af + 1; const float* tn_f = an_f + NOISE_NF;
  const float* kp = tn_f + NOISE_NF + 15; const float* kd = kp + nu; const float* clip = kd + nu;
  if (ti[TI_ACT_TYPE] == ACT_POSITION_VELOCITY) {
    // PositionVelocityActuator.get_ctrl (actuators.py:261-293): action = [position targets | velocity targets]
    const int* vn_i = ai + 2 * NOISE_NI + 5;
    const float* vn_f = clip + nu;
    const KeyFan fan3 = key_fan(rng, LANE_PASS);
    const Key pos_rng = key_child(fan3, 0), vel_rng = key_child(fan3, 1), tor_rng = key_child(fan3, 2);
    LANE_FOR(i, nu) {
      const float tp = noise_apply(an_i, an_f, ctrl_in[i], pos_rng, i, level);
      const float tv = noise_apply(vn_i, vn_f, ctrl_in[nu + i], vel_rng, i, level);
      float c = kp[i] * (tp - (w.qpos[7 + i] - w.zero_off[i])) + kd[i] * (tv - w.qvel[6 + i]);
      c = noise_apply(tn_i, tn_f, c, tor_rng, i, level);
      w.torque[i] = clipf(c, -clip[i], clip[i]);
    }
    WSYNC();
    return;
  }
  const float kp_scale = w.act_state[2 * nu], kd_scale = w.act_state[2 * nu + 1], tl_scale = w.act_state[2 * nu + 2];
  const KeyFan fan = key_fan(rng, LANE_PASS);
  const Key pos_rng = key_child(fan, 0), tor_rng = key_child(fan, 1);
  LANE_FOR(i, nu) {
    const float scaled = ctrl_in[i] * action_scale;
    const float target = noise_apply(an_i, an_f, scaled, pos_rng, i, level) + w.act_state[i];
    const float cur_pos = w.qpos[7 + i] - w.zero_off[i];
    const float cur_vel = w.qvel[6 + i];
    float c = (kp[i] * kp_scale) * (target - cur_pos) + (kd[i] * kd_scale) * (0.0f - cur_vel);
    c = noise_apply(tn_i, tn_f, c, tor_rng, i, level) + w.act_state[nu + i];
    const float lim = clip[i] * tl_scale;
    w.torque[i] = clipf(c, -lim, lim);
  }
  WSYNC();
}

template <class D>
DEV void actuators_initial_state(const Mdl& m, Work<D>& w, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  if (ti[TI_ACT_TYPE] != ACT_POSITION) return;
  const int nu = DIM(NU, MH_NU);
  const int* kinds = ti + ti[TI_ACT_IOFF] + 2 * NOISE_NI;
  const float* rvf = m.tf + ti[TI_ACT_FOFF] + 1 + 2 * NOISE_NF;
  const Key k0 = key_split(rng, 0), k1 = key_split(rng, 1);
  LANE_FOR(i, nu) {
    w.act_state[i] = kinds[0] >= 0 ? rv_sample(kinds[0], rvf + 0, k0, i, 1.0f) : 0.0f;
    w.act_state[nu + i] = kinds[1] >= 0 ? rv_sample(kinds[1], rvf + 3, k1, i, 1.0f) : 0.0f;
  }
  IF_LANE0 {
    w.act_state[2 * nu + 0] = kinds[2] >= 0 ? rv_sample(kinds[2], rvf + 6, key_split(rng, 2), 0, 1.0f) : 1.0f;
    w.act_state[2 * nu + 1] = kinds[3] >= 0 ? rv_sample(kinds[3], rvf + 9, key_split(rng, 3), 0, 1.0f) : 1.0f;
    w.act_state[2 * nu + 2] = kinds[4] >= 0 ? rv_sample(kinds[4], rvf + 12, key_split(rng, 4), 0, 1.0f) : 1.0f;
  }
  WSYNC();
}

// ------------------------------------------------------------------------------------------------- events

// jax.random.randint(key, (), 0, span) (see oracle/engine.c)
DEV uint32_t key_randint(Key k, uint32_t span) {
  const Key ra = key_split(k, 0), rb = key_split(k, 1);
  const uint32_t hi = key_bits(ra, 0), lo = key_bits(rb, 0);
  const uint32_t mult = ((65536u % span) * (65536u % span)) % span;
  return ((hi % span) * mult + (lo % span)) % span;
}
DEV uint32_t randint3(Key k) { return key_randint(k, 3u); }
template <class D>
DEVNI void event_apply(const Mdl& m, Work<D>& w, int idx, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_EVENT_TAB, idx, TAB_TYPE);
  const int* ei = ti + TAB(ti, TI_EVENT_TAB, idx, TAB_IOFF);
  const float* ef = m.tf + TAB(ti, TI_EVENT_TAB, idx, TAB_FOFF);
  float* st = w.event + TAB(ti, TI_EVENT_TAB, idx, TAB_OUT_OFF);
  const float dt = m.mf[MF_TIMESTEP];
  IF_LANE0 {
    if (type == EVENT_LINEAR_PUSH || type == EVENT_ANGULAR_PUSH || type == EVENT_JUMP) {
      float tr = st[0] - dt;
      if (tr <= 0.0f) {
        if (type == EVENT_LINEAR_PUSH) {
          const Key urng = key_split(rng, 0), brng = key_split(rng, 1), trng = key_split(rng, 2);
          const float lvl = scale_get(ei[0], ef + 5, w.level);
          const float theta = key_uniform(urng, 0, 0.0f, 2.0f * PI_F);
          const float mag = key_uniform(brng, 0, ef[1], ef[2]) * ef[0];
          w.qvel[0] += cosf(theta) * mag * lvl;
          w.qvel[1] += sinf(theta) * mag * lvl;
          w.qvel[2] += 0.0f * mag * lvl;
          tr = key_uniform(trng, 0, ef[3], ef[4]);
        } else if (type == EVENT_ANGULAR_PUSH) {
          const Key frng = key_split(rng, 0), brng = key_split(rng, 1), trng = key_split(rng, 2);
          const float lvl = scale_get(ei[0], ef + 5, w.level);
          const int flip = key_bernoulli(frng, 0, 0.5f);
          float mag = key_uniform(brng, 0, ef[1], ef[2]) * ef[0];
          if (flip) mag = -mag;
          w.qvel[5] += mag * lvl;
          tr = key_uniform(trng, 0, ef[3], ef[4]);
        } else {
          const Key urng = key_split(rng, 0), trng = key_split(rng, 1);
          const float lvl = scale_get(ei[0], ef + 4, w.level);
          const float h = key_uniform(urng, 0, ef[0], ef[1]) * lvl;
#pragma unroll
          for (int k = 0; k < 3; k++) w.qvel[k] += sqrtf(2.0f * -m.mf[MF_GRAVITY + k] * h);
          tr = key_uniform(trng, 0, ef[2], ef[3]);
        }
      }
      st[0] = tr;
    } else if (type == EVENT_FORCE_PUSH) {
      const float lvl = scale_get(ei[0], ef + 8, w.level);
      float tr = st[0] - dt, dur = st[1] - dt;
      float force[6];
#pragma unroll
      for (int k = 0; k < 6; k++) force[k] = st[2 + k] * lvl;
      w.xfrc_body = ei[1];
#pragma unroll
      for (int k = 0; k < 6; k++) w.xfrc[k] = dur > 0.0f ? force[k] * lvl : 0.0f;
      if (tr <= 0.0f) {
        const Key k0 = key_split(rng, 0), k1 = key_split(rng, 1), k2 = key_split(rng, 2), k3 = key_split(rng, 3);
        const Key k4 = key_split(rng, 4), k5 = key_split(rng, 5);
        float dir[3], tq[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { dir[k] = key_uniform(k0, k, -1.0f, 1.0f); tq[k] = key_uniform(k2, k, -1.0f, 1.0f); }
        const float fs = key_uniform(k1, 0, ef[2], ef[3]);
        const float nd = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        const float nt = sqrtf(tq[0] * tq[0] + tq[1] * tq[1] + tq[2] * tq[2]);
        const uint32_t off = randint3(k3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float f = dir[k] / nd * fs * ef[0], q = tq[k] / nt * fs * ef[1];
          force[k] = (off == 1) ? 0.0f : f;
          force[3 + k] = (off == 0) ? 0.0f : q;
        }
        tr = key_uniform(k4, 0, ef[6], ef[7]);
        dur = key_uniform(k5, 0, ef[4], ef[5]);
      }
      st[0] = tr; st[1] = dur;
#pragma unroll
      for (int k = 0; k < 6; k++) st[2 + k] = force[k];
    }
  }
  WSYNC();
}

template <class D>
DEV void event_initial_state(const Mdl& m, Work<D>& w, int idx, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_EVENT_TAB, idx, TAB_TYPE);
  const float* ef = m.tf + TAB(ti, TI_EVENT_TAB, idx, TAB_FOFF);
  float* st = w.event + TAB(ti, TI_EVENT_TAB, idx, TAB_OUT_OFF);
  IF_LANE0 {
    switch (type) {
      case EVENT_LINEAR_PUSH: case EVENT_ANGULAR_PUSH: st[0] = key_uniform(rng, 0, ef[3], ef[4]); break;
      case EVENT_JUMP: st[0] = key_uniform(rng, 0, ef[2], ef[3]); break;
      case EVENT_FORCE_PUSH:
        st[0] = key_uniform(rng, 0, ef[6], ef[7]);
        for (int k = 1; k < 8; k++) st[k] = 0.0f;
        break;
    }
  }
}

// ----------------------------------------------------------------------------------- engine.step / reset

template <class D>
DEV void engine_step(const Mdl& m, Work<D>& w, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int na = ti[TI_ACTION_SIZE], nu = DIM(NU, MH_NU), nsub = ti[TI_NSUB], act_period = ti[TI_ACT_PERIOD];
  const int nevent = ti[TI_N_EVENT];
  const float* ef = m.tf + ti[TI_ENGINE_FOFF];
  const KeyFan fan0 = key_fan(rng, LANE_PASS);
  const Key drop_rng = key_child(fan0, 0);
  Key r = key_child(fan0, 1);
  const float drop_p = ef[2] * w.level;
  for (int base = 0; base < na; base += 32) {
#if B2S_DEVICE
    const int i = base + lane;
    const unsigned bits = __ballot_sync(0xffffffffu, i < na && drop_p > 0.0f && key_bernoulli(drop_rng, i, drop_p));
#else
    unsigned bits = 0u;
    for (int i = base; i < na && i < base + 32; i++) bits |= (drop_p > 0.0f && key_bernoulli(drop_rng, i, drop_p)) ? (1u << (i - base)) : 0u;
#endif
    IF_LANE0 w.drop_mask[base >> 5] = bits;
  }
  LANE_FOR(i, nu) w.torque[i] = 0.0f;
  WSYNC();
  for (int step = 0; step < nsub; step++) {
    STAGE_BEGIN();
    if ((m.sync_mask >> 6) & 1) CTA_SYNC();  // keep the CTA's warps on the same code so they share instruction-cache lines
    const float prct = clipf((float)step - w.latency, 0.0f, 1.0f);
    LANE_FOR(i, na) {
      const float lc = w.last_ctrl[i];
      w.ctrl_blend[i] = ((w.drop_mask[i >> 5] >> (i & 31)) & 1u) ? lc : lc * (1.0f - prct) + w.action[i] * prct;
    }
    WSYNC();
    const KeyFan fan = key_fan(r, LANE_PASS);
    const Key rng_update = key_child(fan, 1);
    r = key_child(fan, 0);
    if (step % act_period == 0) actuators_ctrl(m, w, w.ctrl_blend, rng_update, LANE_PASS);
    LANE_FOR(i, nu) w.ctrl[i] = w.torque[i];
    WSYNC();
    for (int ev = 0; ev < nevent; ev++) {
      const KeyFan fe = key_fan(r, LANE_PASS);
      const Key rng_event = key_child(fe, 1);
      r = key_child(fe, 0);
      // a push / jump event whose countdown has not expired only counts down: skip the call
      const int etype = TAB(ti, TI_EVENT_TAB, ev, TAB_TYPE);
      float* st = w.event + TAB(ti, TI_EVENT_TAB, ev, TAB_OUT_OFF);
      const float tr = st[0] - m.mf[MF_TIMESTEP];
      if (etype != EVENT_FORCE_PUSH && tr > 0.0f) {
        WSYNC();
        IF_LANE0 st[0] = tr;
        WSYNC();
      } else {
        event_apply(m, w, ev, rng_event, LANE_PASS);
      }
    }
    STAGE_MARK(w, 9);
    physics_step(m, w, step == nsub - 1, LANE_PASS);
  }
  LANE_FOR(i, na) w.last_ctrl[i] = w.action[i];
  WSYNC();
}

// element i of jax.random.randint(key, (n,), lo, hi): offset in [0, span)
DEV uint32_t key_randint_at(Key k, int i, uint32_t span) {
  const Key ra = key_split(k, 0), rb = key_split(k, 1);
  const uint32_t hi = key_bits(ra, i), lo = key_bits(rb, i);
  const uint32_t mult = ((65536u % span) * (65536u % span)) % span;
  return ((hi % span) * mult + (lo % span)) % span;
}

template <class D>
DEV void reset_apply(const Mdl& m, Work<D>& w, int idx, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_RESET_TAB, idx, TAB_TYPE);
  const int* ri = ti + TAB(ti, TI_RESET_TAB, idx, TAB_IOFF);
  const float* rf = m.tf + TAB(ti, TI_RESET_TAB, idx, TAB_FOFF);
  const int nj = DIM(NQ, MH_NQ) - 7;
  switch (type) {
    case RESET_RANDOM_JOINT_POSITION: {
      const float *zeros = rf + 1, *mins = rf + 1 + nj, *maxs = rf + 1 + 2 * nj;
      LANE_FOR(i, nj) {
        float p = key_uniform(rng, i, -rf[0], rf[0]);
        if (ri[0]) p = p * w.level;
        if (ri[1]) p = p + zeros[i];
        if (ri[2] && p < mins[i]) p = mins[i];
        if (ri[3] && p > maxs[i]) p = maxs[i];
        w.qpos[7 + i] = p;
      }
    } break;
    case RESET_RANDOM_JOINT_VELOCITY:
      LANE_FOR(i, DIM(NV, MH_NV) - 6) {
        float n = key_uniform(rng, i, -rf[0], rf[0]);
        if (ri[0]) n = n * w.level;
        w.qvel[6 + i] = n;
      }
      break;
    case RESET_RANDOM_HEADING:
      IF_LANE0 {
        const float ang = key_uniform(rng, 0, -PI_F, PI_F);
        Q4 q; q.w = w.qpos[3]; q.x = w.qpos[4]; q.y = w.qpos[5]; q.z = w.qpos[6];
        q = qmul(q, euler_to_quat(0.0f, 0.0f, ang));
        w.qpos[3] = q.w; w.qpos[4] = q.x; w.qpos[5] = q.y; w.qpos[6] = q.z;
      }
      break;
    case RESET_RANDOM_BASE_VELOCITY_XY: break;  // no-op on the reference's MJX path (ksim/resets.py:204-206)
    case RESET_PLANE_XY_POSITION:
      IF_LANE0 {
        const Key kx = key_split(rng, 0), ky = key_split(rng, 1);
        w.qpos[0] = clipf(key_uniform(kx, 0, -rf[1], rf[1]), rf[3], rf[4]);
        w.qpos[1] = clipf(key_uniform(ky, 0, -rf[2], rf[2]), rf[5], rf[6]);
        w.qpos[2] = rf[0];
      }
      break;
    case RESET_HFIELD_XY_POSITION:
      // resets.py:56-89: xy like the plane reset, z read from the height field's cell under (x, y)
      IF_LANE0 {
        const Key kx = key_split(rng, 0), ky = key_split(rng, 1);
        const float x = clipf(key_uniform(kx, 0, -rf[4], rf[4]), rf[6], rf[7]);
        const float y = clipf(key_uniform(ky, 0, -rf[5], rf[5]), rf[8], rf[9]);
        const int nx = ri[0], ny = ri[1];
        int xi = (int)(((x + rf[0]) / (2.0f * rf[0])) * (float)(nx - 1));
        int yi = (int)(((y + rf[1]) / (2.0f * rf[1])) * (float)(ny - 1));
        xi = xi < 0 ? 0 : (xi > nx - 1 ? nx - 1 : xi);
        yi = yi < 0 ? 0 : (yi > ny - 1 ? ny - 1 : yi);
        w.qpos[0] = x; w.qpos[1] = y;
        w.qpos[2] = rf[10 + xi * ny + yi] + rf[2] + rf[3];
      }
      break;
    case RESET_INITIAL_MOTION_STATE: {
      // resets.py:290-304 (freejoint = False): joints from a random frame of the motion bank, the base is left alone; the
      // bank stores qvel nq wide, the joints' velocities are its columns 7.. ; time = frame * ctrl_dt
      const int frames = ri[0], wq = ri[1];
      const int frame = (int)key_randint_at(rng, 0, (uint32_t)frames);
      const float* qp = rf + 1 + frame * wq;
      const float* qv = rf + 1 + (frames + frame) * wq;
      LANE_FOR(i, nj) { w.qpos[7 + i] = qp[7 + i]; w.qvel[6 + i] = qv[7 + i]; }
      IF_LANE0 w.time = (float)frame * rf[0];
    } break;
    case RESET_RANDOM_HEIGHT:
      IF_LANE0 w.qpos[2] += key_uniform(rng, 0, rf[0], rf[1]);
      break;
    case RESET_RANDOM_PITCH_ROLL:
      IF_LANE0 {
        const float pitch = key_uniform(rng, 0, rf[0], rf[1]), roll = key_uniform(rng, 0, rf[2], rf[3]);
        Q4 q; q.w = w.qpos[3]; q.x = w.qpos[4]; q.y = w.qpos[5]; q.z = w.qpos[6];
        q = qmul(q, euler_to_quat(roll, pitch, 0.0f));
        w.qpos[3] = q.w; w.qpos[4] = q.x; w.qpos[5] = q.y; w.qpos[6] = q.z;
      }
      break;
  }
  WSYNC();
}

template <class D>
DEVNI void engine_reset(const Mdl& m, Work<D>& w, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int nq = DIM(NQ, MH_NQ), nv = DIM(NV, MH_NV), nu = DIM(NU, MH_NU);
  const float* ef = m.tf + ti[TI_ENGINE_FOFF];
  const float* qpos0 = MF(m, MH_QPOS0);
  // mjx.make_data
  LANE_FOR(i, nq) w.qpos[i] = qpos0[i];
  LANE_FOR(i, nv) { w.qvel[i] = 0.0f; w.qacc[i] = 0.0f; w.warm[i] = 0.0f; }
  LANE_FOR(i, nu) w.ctrl[i] = 0.0f;
  IF_LANE0 {
    w.time = 0.0f; w.xfrc_body = -1;
#pragma unroll
    for (int k = 0; k < 6; k++) w.xfrc[k] = 0.0f;
  }
  WSYNC();
  const int nreset = ti[TI_N_RESET];
  for (int i = 0; i < nreset; i++) {
    const KeyFan fr = key_fan(rng, LANE_PASS);
    const Key reset_rng = key_child(fr, 1);
    rng = key_child(fr, 0);
    reset_apply(m, w, i, reset_rng, LANE_PASS);
  }
  DBG_STOP(m, 10);
  forward(m, w, true, false, LANE_PASS);
  DBG_STOP(m, 11);
  LANE_FOR(i, nv) { w.qvel[i] = 0.0f; w.qacc[i] = 0.0f; }
  WSYNC();
  actuators_initial_state(m, w, rng, LANE_PASS);
  const KeyFan fz = key_fan(rng, LANE_PASS);
  Key events_rng = key_child(fz, 0);
  const Key latency_rng = key_child(fz, 1), zero_rng = key_child(fz, 2);
  LANE_FOR(i, ti[TI_ACTION_SIZE]) w.last_ctrl[i] = 0.0f;
  const int nevent = ti[TI_N_EVENT];
  for (int ev = 0; ev < nevent; ev++) {
    const KeyFan fe = key_fan(events_rng, LANE_PASS);
    const Key event_rng = key_child(fe, 1);
    events_rng = key_child(fe, 0);
    event_initial_state(m, w, ev, event_rng, LANE_PASS);
  }
  IF_LANE0 w.latency = key_uniform(latency_rng, 0, ef[0] * w.level, ef[1] * w.level);
  const int zkind = ti[TI_ZERO_OFFSET_KIND];
  LANE_FOR(i, nu) w.zero_off[i] = rv_sample(zkind, ef + 3, zero_rng, i, w.level);
  WSYNC();
}

// ----------------------------------------------------------------------------------------------- commands

template <class D>
DEVNI void cmd_linvel_initial(const int* ci, const float* cf, const Work<D>& w, Key rng, const float* prev, float* out) {
  const float lvl = w.level;
  const float min_vel = scale_get(ci[0], cf + 0, lvl), max_vel = scale_get(ci[1], cf + 2, lvl);
  const float max_yaw = scale_get(ci[2], cf + 4, lvl), zero_prob = scale_get(ci[3], cf + 6, lvl);
  const float backward_prob = scale_get(ci[4], cf + 8, lvl);
  const Key rv = key_split(rng, 0), ry = key_split(rng, 1), rz = key_split(rng, 2), rb = key_split(rng, 3);
  float cur_vel, cur_yaw;
  if (!prev) {
    Q4 q; q.w = w.qpos[3]; q.x = w.qpos[4]; q.y = w.qpos[5]; q.z = w.qpos[6];
    V3 lv = rot_vec_quat(v3(w.qvel[0], w.qvel[1], w.qvel[2]), q, true);
    cur_vel = sqrtf(lv.x * lv.x + lv.y * lv.y);
    cur_yaw = atan2f(lv.y, lv.x);
  } else { cur_vel = prev[0]; cur_yaw = prev[1]; }
  float trg_vel = key_uniform(rv, 0, min_vel, max_vel);
  float trg_yaw = key_uniform(ry, 0, -max_yaw, max_yaw);
  const int zero = key_bernoulli(rz, 0, zero_prob), back = key_bernoulli(rb, 0, backward_prob);
  trg_vel = zero ? 0.0f : (back ? -trg_vel : trg_vel);
  trg_yaw = zero ? 0.0f : trg_yaw;
  out[0] = cur_vel; out[1] = cur_yaw; out[2] = cur_vel * cosf(cur_yaw); out[3] = cur_vel * sinf(cur_yaw);
  out[4] = trg_vel; out[5] = trg_yaw;
}

template <class D>
DEV void cmd_angvel_initial(const int* ci, const float* cf, const Work<D>& w, Key rng, float* out) {
  const float lvl = w.level;
  const float min_vel = scale_get(ci[0], cf + 0, lvl), max_vel = scale_get(ci[1], cf + 2, lvl);
  const float zero_prob = scale_get(ci[2], cf + 4, lvl);
  const Key rv = key_split(rng, 0), rfl = key_split(rng, 1), rz = key_split(rng, 2);
  float trg = key_uniform(rv, 0, min_vel, max_vel);
  if (key_bernoulli(rfl, 0, 0.5f)) trg = -trg;
  if (key_bernoulli(rz, 0, zero_prob)) trg = 0.0f;
  out[0] = w.qvel[5]; out[1] = trg;
}

DEV void cmd_floatvec_initial(const float* cf, int dim, Key rng, float* out) {
  const int zero = key_bernoulli(rng, 0, cf[2 * dim + 1]);
  for (int i = 0; i < dim; i++) out[i] = zero ? 0.0f : key_uniform(rng, i, cf[2 * i], cf[2 * i + 1]);
}

// UnifiedCommand.initial_command (examples/kbot/train.py:714-756); floats: 6 x [lo, hi], arms lo[10], arms hi[10], switch_prob
DEVNI void cmd_unified_initial(const float* cf, Key rng, float* out) {
  float v[6], arms[10];
#pragma unroll
  for (int k = 0; k < 6; k++) v[k] = key_uniform(key_split(rng, 1 + k), 0, cf[2 * k], cf[2 * k + 1]);
  const Key rh = key_split(rng, 7);
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const float a = key_uniform(rh, k, cf[12 + k], cf[22 + k]);
    arms[k] = key_bernoulli(rh, k, 0.5f) ? a : 0.0f;
  }
  const uint32_t mode = key_randint(key_split(rng, 0), 6u);
#pragma unroll
  for (int k = 0; k < 16; k++) out[k] = 0.0f;
  if (mode == 0u || mode == 3u) out[0] = v[0];
  if (mode == 1u || mode == 3u) out[1] = v[1];
  if (mode == 2u || mode == 3u) out[2] = v[2];
  if (mode == 4u) { out[3] = v[3]; out[4] = v[4]; out[5] = v[5]; }
  if (mode == 3u || mode == 4u) {
#pragma unroll
    for (int k = 0; k < 10; k++) out[6 + k] = arms[k];
  }
}

// CartesianCoordinateCommand._sample_box (commands.py:557-581)
DEV void cmd_cartesian_sample(const float* cf, float level, Key rng, float* out) {
  const float sf = cf[11] * level + 1.0f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float center = (cf[k] + cf[3 + k]) / 2.0f, half = (cf[3 + k] - cf[k]) / 2.0f * sf;
    out[k] = key_uniform(rng, k, center - half, center + half);
  }
}
// SinusoidalGaitCommand._get_height (commands.py:712-733)
DEV void cmd_gait_heights(const float* cf, int nf, float phase, float* out) {
  const float max_h = cf[2], offset = cf[3], sr = cf[4];
  for (int f = 0; f < nf; f++) {
    float pf = phase + (float)f / (float)nf;
    pf = pf - floorf(pf);
    const float prog = clipf((pf - sr) / (1.0f - sr), 0.0f, 1.0f);
    out[f] = (pf >= sr ? max_h * sinf(PI_F * prog) : 0.0f) + offset;
  }
}

// lane 0 only
template <class D>
DEV void cmd_initial(const Mdl& m, Work<D>& w, int idx, Key rng) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_CMD_TAB, idx, TAB_TYPE);
  const int* ci = ti + TAB(ti, TI_CMD_TAB, idx, TAB_IOFF);
  const float* cf = m.tf + TAB(ti, TI_CMD_TAB, idx, TAB_FOFF);
  float* st = w.cmd + TAB(ti, TI_CMD_TAB, idx, TAB_OUT_OFF);
  const int dim = TAB(ti, TI_CMD_TAB, idx, TAB_OUT_DIM);
  if (type == CMD_LINEAR_VELOCITY) cmd_linvel_initial(ci, cf, w, rng, nullptr, st);
  else if (type == CMD_ANGULAR_VELOCITY) cmd_angvel_initial(ci, cf, w, rng, st);
  else if (type == CMD_FLOAT_VECTOR) cmd_floatvec_initial(cf, dim, rng, st);
  else if (type == CMD_UNIFIED) cmd_unified_initial(cf, rng, st);
  else if (type == CMD_INT_VECTOR) {
    const int zero = key_bernoulli(rng, 0, cf[1]);
    for (int i = 0; i < dim; i++)
      st[i] = zero ? 0.0f : (float)(ci[2 * i] + (int)key_randint_at(rng, i, (uint32_t)(ci[2 * i + 1] + 1 - ci[2 * i])));
  } else if (type == CMD_START_POSITION) { st[0] = w.qpos[0]; st[1] = w.qpos[1]; st[2] = w.qpos[2]; }
  else if (type == CMD_START_QUATERNION) { st[0] = w.qpos[3]; st[1] = w.qpos[4]; st[2] = w.qpos[5]; st[3] = w.qpos[6]; }
  else if (type == CMD_BASE_HEIGHT) st[0] = key_uniform(rng, 0, cf[0], cf[1]);
  else if (type == CMD_CARTESIAN_COORDINATE) {
    float tg[3];
    cmd_cartesian_sample(cf, w.level, key_split(rng, 0), tg);
#pragma unroll
    for (int k = 0; k < 3; k++) { st[k] = tg[k]; st[3 + k] = tg[k]; }
    st[6] = key_uniform(key_split(rng, 1), 0, cf[7], cf[8]);
  } else if (type == CMD_SINUSOIDAL_GAIT) {
    st[0] = 1.0f; st[1] = 0.0f;
    cmd_gait_heights(cf, ci[0], 0.0f, st + 2);
  } else if (type == CMD_JOINT_POSITION) {
    const int n = ci[0];
    const Key ra = key_split(rng, 0), rb = key_split(rng, 1);
    for (int i = 0; i < n; i++) { st[i] = w.qpos[7 + ci[1 + i]]; st[n + i] = key_uniform(ra, i, cf[2 * i], cf[2 * i + 1]); }
    const float tt = key_uniform(rb, 0, cf[2 * n + 1], cf[2 * n + 2]);
    st[2 * n] = tt; st[2 * n + 1] = ceilf(tt / cf[2 * n]); st[2 * n + 2] = 0.0f;
  }
}

// lane 0 only
template <class D>
DEV void cmd_update(const Mdl& m, Work<D>& w, int idx, Key rng) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_CMD_TAB, idx, TAB_TYPE);
  const int* ci = ti + TAB(ti, TI_CMD_TAB, idx, TAB_IOFF);
  const float* cf = m.tf + TAB(ti, TI_CMD_TAB, idx, TAB_FOFF);
  float* st = w.cmd + TAB(ti, TI_CMD_TAB, idx, TAB_OUT_OFF);
  const int dim = TAB(ti, TI_CMD_TAB, idx, TAB_OUT_DIM);
  // the resampled command (key rb) only matters when the switch (key ra) fires -- once in hundreds of steps -- so it is
  // drawn only then; the values are the same as drawing it always
  const Key ra = key_split(rng, 0);
  if (type == CMD_LINEAR_VELOCITY) {
    const float sp = scale_get(ci[5], cf + 10, w.level);
    const float ctrl_dt = cf[12], lin_acc = cf[13], ang_acc = cf[14];
    if (key_bernoulli(ra, 0, sp)) {
      float fresh[6];
      cmd_linvel_initial(ci, cf, w, key_split(rng, 1), st, fresh);
#pragma unroll
      for (int k = 0; k < 6; k++) st[k] = fresh[k];
    } else {
      const float nv = st[0] + clipf(st[4] - st[0], -lin_acc, lin_acc) * ctrl_dt;
      const float ny = st[1] + clipf(st[5] - st[1], -ang_acc, ang_acc) * ctrl_dt;
      st[0] = nv; st[1] = ny; st[2] = nv * cosf(ny); st[3] = nv * sinf(ny);
    }
  } else if (type == CMD_ANGULAR_VELOCITY) {
    const float sp = scale_get(ci[3], cf + 6, w.level);
    const float ctrl_dt = cf[8], ang_acc = cf[9];
    if (key_bernoulli(ra, 0, sp)) {
      float fresh[2];
      cmd_angvel_initial(ci, cf, w, key_split(rng, 1), fresh);
      st[0] = fresh[0]; st[1] = fresh[1];
    } else {
      st[0] = st[0] + clipf(st[1] - st[0], -ang_acc, ang_acc) * ctrl_dt;
    }
  } else if (type == CMD_FLOAT_VECTOR) {
    if (key_bernoulli(ra, 0, cf[2 * dim])) {
      float fresh[16];
      cmd_floatvec_initial(cf, dim, key_split(rng, 1), fresh);
      for (int i = 0; i < dim; i++) st[i] = fresh[i];
    }
  } else if (type == CMD_INT_VECTOR) {
    if (key_bernoulli(ra, 0, cf[0])) cmd_initial(m, w, idx, key_split(rng, 1));
  } else if (type == CMD_BASE_HEIGHT) {
    // commands.py:790-796: the fresh height and the switch are drawn from the SAME (unsplit) key
    if (key_bernoulli(rng, 0, cf[2])) st[0] = key_uniform(rng, 0, cf[0], cf[1]);
  } else if (type == CMD_CARTESIAN_COORDINATE) {
    // commands.py:599-643
    const float dx = st[3] - st[0], dy = st[4] - st[1], dz = st[5] - st[2];
    const float distance = sqrtf(dx * dx + dy * dy + dz * dz);
    const Key rc = key_split(rng, 2);
    const bool reached = distance < cf[6] * st[6] * 0.5f;
    const bool change = (reached && key_bernoulli(rc, 0, cf[9])) || key_bernoulli(rc, 0, cf[10]);
    if (change) {
      cmd_cartesian_sample(cf, w.level, ra, st + 3);
      st[6] = key_uniform(key_split(rng, 1), 0, cf[7], cf[8]);
    }
    const float step = st[6] * cf[6];
    float d[3] = {st[3] - st[0], st[4] - st[1], st[5] - st[2]};
    const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float adv = fminf(step, distance);
#pragma unroll
    for (int k = 0; k < 3; k++) st[k] = st[k] + (dn > 0.0f ? d[k] / dn : d[k]) * adv;
  } else if (type == CMD_SINUSOIDAL_GAIT) {
    // commands.py:748-770: phase advances while the robot moves, restarts when it stops (the flag never comes back on)
    const bool moving = st[0] != 0.0f;
    float ph = st[1] + cf[1] / cf[0];
    ph = moving ? ph - floorf(ph) : 0.0f;
    const bool is_moving = sqrtf(w.qvel[0] * w.qvel[0] + w.qvel[1] * w.qvel[1]) > cf[5];
    if (!is_moving) { st[0] = 0.0f; ph = 0.0f; }
    st[1] = ph;
    cmd_gait_heights(cf, ci[0], ph, st + 2);
  } else if (type == CMD_JOINT_POSITION) {
    // commands.py:853-880: a fresh command (same, unsplit key) once the step counter has passed num_steps
    const int n = ci[0];
    const float step_id = st[2 * n + 2] + 1.0f;
    if (step_id > st[2 * n + 1]) cmd_initial(m, w, idx, rng);
    else st[2 * n + 2] = step_id;
  } else if (type == CMD_UNIFIED) {
    if (key_bernoulli(ra, 0, cf[32])) cmd_unified_initial(cf, key_split(rng, 1), st);
  }
}

// ------------------------------------------------------------------------------------------- observations

template <class D>
DEV int geoms_colliding(const Mdl& m, const Work<D>& w, const int* g1, int n1, const int* g2, int n2) {
  const int ncon = DIM(NCON, MH_NCON);
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  const int* pg2 = MI(m, MH_PAIR_GEOM2);
  for (int c = 0; c < ncon; c++) {
    if (!(w.con_dist[c] < 0.0f)) continue;
    const int a = pg1[w.con_pair[c]], b = pg2[w.con_pair[c]];
    for (int i = 0; i < n1; i++)
      for (int j = 0; j < n2; j++)
        if ((g1[i] == a && g2[j] == b) || (g1[i] == b && g2[j] == a)) return 1;
  }
  return 0;
}

template <class D>
DEV void obs_carry_initial(const Mdl& m, Work<D>& w, int idx, Key rng, LANE_ARG) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_OBS_TAB, idx, TAB_TYPE);
  const int coff = TAB(ti, TI_OBS_TAB, idx, TAB_AUX1);
  if (coff < 0) return;
  const float* of = m.tf + TAB(ti, TI_OBS_TAB, idx, TAB_FOFF) + NOISE_NF + BIAS_NF;
  float* c = w.obs_carry + coff;
  if (type == OBS_PROJECTED_GRAVITY) {
    IF_LANE0 {
      const Key lrng = key_split(rng, 0), brng = key_split(rng, 1);
      c[0] = c[1] = c[2] = 0.0f;
      c[3] = key_uniform(lrng, 0, of[3], of[4]);
#pragma unroll
      for (int k = 0; k < 3; k++) c[4 + k] = key_uniform(brng, k, -of[5], of[5]);
    }
  } else if (type == OBS_DELAYED_JOINT_POSITION || type == OBS_DELAYED_JOINT_VELOCITY) {
    // observation.py:232-236 / 264-268: every row of the ring starts as the joint state right after the reset
    const int dim = TAB(ti, TI_OBS_TAB, idx, TAB_OUT_DIM);
    const int delay = (ti + TAB(ti, TI_OBS_TAB, idx, TAB_IOFF) + NOISE_NI + BIAS_NI)[0];
    const bool pos = type == OBS_DELAYED_JOINT_POSITION;
    LANE_FOR(k, dim) {
      const float cur = pos ? w.qpos[7 + k] : w.qvel[6 + k];
      for (int r = 0; r < delay; r++) c[r * dim + k] = cur;
    }
  }
}

// COMDistanceObservation (examples/kbot/train.py:500-649): monotone-chain hull of the contact points' xy (points of
// contacts whose geom1 is not the floor are (0,0)), area centroid of the hull polygon (vertex mean when the area
// vanishes), distance to subtree_com[2].  A few dozen operations on <= NCON points, once per control step; serial.
template <class D>
DEVNI float com_distance(const Mdl& m, const Work<D>& w) {
  constexpr int NC = D::NCON;
  const int n = DIM(NCON, MH_NCON);
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  float px[NC], py[NC];
  for (int c = 0; c < n; c++) {
    const bool floor = pg1[w.con_pair[c]] == 0;
    px[c] = floor ? w.con_pos[0][c] : 0.0f;
    py[c] = floor ? w.con_pos[1][c] : 0.0f;
  }
  for (int i = 1; i < n; i++) {  // stable insertion sort == lexsort by x then y
    const float x = px[i], y = py[i];
    int j = i - 1;
    while (j >= 0 && (px[j] > x || (px[j] == x && py[j] > y))) { px[j + 1] = px[j]; py[j + 1] = py[j]; j--; }
    px[j + 1] = x; py[j + 1] = y;
  }
  float hx[2 * NC], hy[2 * NC];
  int np = 0;
  for (int pass = 0; pass < 2; pass++) {
    int stack[NC], ptr = 0;
    for (int t = 0; t < n; t++) {
      const int idx = pass == 0 ? t : n - 1 - t;
      while (ptr >= 2) {
        const int a = stack[ptr - 2], b = stack[ptr - 1];
        const float cr = (px[b] - px[a]) * (py[idx] - py[a]) - (py[b] - py[a]) * (px[idx] - px[a]);
        if (cr <= 0.0f) ptr--; else break;
      }
      stack[ptr++] = idx;
    }
    for (int t = 0; t < ptr - 1; t++) { hx[np] = px[stack[t]]; hy[np] = py[stack[t]]; np++; }
  }
  float area = 0.0f, sx = 0.0f, sy = 0.0f, mx = 0.0f, my = 0.0f;
  for (int i = 0; i < np; i++) {
    const int j = i + 1 < np ? i + 1 : 0;
    const float cr = hx[i] * hy[j] - hx[j] * hy[i];
    area += cr;
    sx += (hx[i] + hx[j]) * cr;
    sy += (hy[i] + hy[j]) * cr;
    mx += hx[i]; my += hy[i];
  }
  area *= 0.5f;
  const float cnt = np > 0 ? (float)np : 1.0f;
  const bool flat = fabsf(area) < 1e-12f;
  const float cx = flat ? mx / cnt : sx / (6.0f * area), cy = flat ? my / cnt : sy / (6.0f * area);
  const float dx = cx - w.subtree_com[0][2], dy = cy - w.subtree_com[1][2];
  return sqrtf(dx * dx + dy * dy);
}

// get_observation: writes the obs block of a record (global memory)
template <class D>
DEVNI void observe(const Mdl& m, Work<D>& w, Key rng, float* out, LANE_ARG) {
  const int* ti = m.ti;
  const int nobs = ti[TI_N_OBS];
  const int nb = DIM(NB, MH_NBODY);
  (void)nb;
  for (int o = 0; o < nobs; o++) {
    const KeyFan fo = key_fan(rng, LANE_PASS);
    const Key noise_rng = key_child(fo, 2);
    rng = key_child(fo, 0);
    const int type = TAB(ti, TI_OBS_TAB, o, TAB_TYPE);
    const int* oi_all = ti + TAB(ti, TI_OBS_TAB, o, TAB_IOFF);
    const float* of_all = m.tf + TAB(ti, TI_OBS_TAB, o, TAB_FOFF);
    const int* oi = oi_all + NOISE_NI + BIAS_NI;
    const float* of = of_all + NOISE_NF + BIAS_NF;
    const int off = TAB(ti, TI_OBS_TAB, o, TAB_OUT_OFF), dim = TAB(ti, TI_OBS_TAB, o, TAB_OUT_DIM);
    const int noff = TAB(ti, TI_OBS_TAB, o, TAB_AUX0), coff = TAB(ti, TI_OBS_TAB, o, TAB_AUX1);
    float* v = out + off;
    const int* bi = oi_all + NOISE_NI;
    const float* bf = of_all + NOISE_NF;
    Key bias_key; bias_key.k0 = w.keys[2 + 2 * o]; bias_key.k1 = w.keys[3 + 2 * o];
    Q4 bq; bq.w = w.qpos[3]; bq.x = w.qpos[4]; bq.y = w.qpos[5]; bq.z = w.qpos[6];
    // stateful observation first (updates its carry through lane 0)
    if (type == OBS_PROJECTED_GRAVITY) {
      float* c = w.obs_carry + coff;
      Q4 biasq = euler_to_quat(c[4], c[5], c[6]);
      Q4 fq; fq.w = w.sensordata[oi[0]]; fq.x = w.sensordata[oi[0] + 1]; fq.y = w.sensordata[oi[0] + 2]; fq.z = w.sensordata[oi[0] + 3];
      V3 pg = rot_vec_quat(v3(of[0], of[1], of[2]), fq, true);
      pg = rot_vec_quat(pg, biasq, false);
      const float lag = c[3];
      const float x0 = c[0] * lag + pg.x * (1.0f - lag), x1 = c[1] * lag + pg.y * (1.0f - lag), x2 = c[2] * lag + pg.z * (1.0f - lag);
      WSYNC();
      IF_LANE0 { c[0] = x0; c[1] = x1; c[2] = x2; }
      WSYNC();
    }
    LANE_FOR(k, dim) {
      float x = 0.0f;
      switch (type) {
        case OBS_BASE_POSITION: x = w.qpos[k]; break;
        case OBS_BASE_ORIENTATION: x = w.qpos[3 + k]; break;
        case OBS_BASE_LINEAR_VELOCITY:
          if (oi[0]) { V3 r = rot_vec_quat(v3(w.qvel[0], w.qvel[1], w.qvel[2]), bq, true); x = k == 0 ? r.x : (k == 1 ? r.y : r.z); }
          else x = w.qvel[k];
          break;
        case OBS_BASE_ANGULAR_VELOCITY:
          if (oi[0]) { V3 r = rot_vec_quat(v3(w.qvel[3], w.qvel[4], w.qvel[5]), bq, true); x = k == 0 ? r.x : (k == 1 ? r.y : r.z); }
          else x = w.qvel[3 + k];
          break;
        case OBS_JOINT_POSITION: x = w.qpos[7 + k] - w.zero_off[k]; break;
        case OBS_JOINT_VELOCITY: x = w.qvel[6 + k]; break;
        case OBS_CENTER_OF_MASS_INERTIA: x = w.cinert[k % 10][1 + k / 10]; break;
        case OBS_CENTER_OF_MASS_VELOCITY: x = w.cvel[k % 6][1 + k / 6]; break;
        case OBS_ACTUATOR_FORCE: x = w.actuator_force[k]; break;
        case OBS_SENSOR: x = w.sensordata[oi[0] + k]; break;
        case OBS_BASE_LINEAR_ACCELERATION: x = w.qacc[k]; break;
        case OBS_BASE_ANGULAR_ACCELERATION: x = w.qacc[3 + k]; break;
        case OBS_ACTUATOR_ACCELERATION: x = w.qacc[6 + k]; break;
        case OBS_PROJECTED_GRAVITY: x = w.obs_carry[coff + k]; break;
        case OBS_FEET_CONTACT: {
          const int* ids = oi + 3;
          const int r = k >> 1, side = k & 1;
          x = (float)geoms_colliding(m, w, ids + (side ? oi[0] : 0) + r, 1, ids + oi[0] + oi[1], oi[2]);
        } break;
        case OBS_BODY_POSITION: x = w.xpos[k % 3][oi[1 + k / 3]]; break;
        case OBS_BODY_VELOCITY: x = w.cvel[k % 6][oi[1 + k / 6]]; break;
        case OBS_FEET_FORCE: x = w.cfrc_ext[k % 3][oi[k / 3]]; break;
        case OBS_FEET_TORQUE: x = w.cfrc_ext[3 + k % 3][oi[k / 3]]; break;
        case OBS_TIMESTEP: x = w.time; break;
        case OBS_FEET_POSITION: {
          // examples/kbot/train.py:672-689; ints [base, left, right]
          Q4 bqx; bqx.w = w.xquat[0][oi[0]]; bqx.x = w.xquat[1][oi[0]]; bqx.y = w.xquat[2][oi[0]]; bqx.z = w.xquat[3][oi[0]];
          const Q4 yawq = euler_to_quat(0.0f, 0.0f, quat_to_euler(bqx).z);
          const int fb = oi[1 + k / 3];
          const V3 rel = v3(w.xpos[0][fb] - w.xpos[0][oi[0]], w.xpos[1][fb] - w.xpos[1][oi[0]], w.xpos[2][fb] - w.xpos[2][oi[0]]);
          const V3 r = rot_vec_quat(rel, yawq, true);
          x = k % 3 == 0 ? r.x : (k % 3 == 1 ? r.y : r.z);
        } break;
        case OBS_BASE_HEIGHT: x = w.xpos[2][1]; break;
        case OBS_COM_DISTANCE: x = oi[0] ? com_distance(m, w) : -1.0f; break;
        case OBS_DELAYED_JOINT_POSITION: case OBS_DELAYED_JOINT_VELOCITY: {
          // observation.py:238-252 / 270-283: the oldest row is the reading, the ring shifts by one and takes the current
          // state; lane k owns column k of the ring, so the shift needs no cross-lane ordering
          float* c = w.obs_carry + coff;
          const int delay = oi[0];
          const bool pos = type == OBS_DELAYED_JOINT_POSITION;
          x = pos ? c[k] - w.zero_off[k] : c[k];
          for (int r = 0; r + 1 < delay; r++) c[r * dim + k] = c[(r + 1) * dim + k];
          c[(delay - 1) * dim + k] = pos ? w.qpos[7 + k] : w.qvel[6 + k];
        } break;
        case OBS_CONTACT: x = (float)geoms_colliding(m, w, oi + 1 + k, 1, oi + 1, oi[0]); break;
        case OBS_BODY_ORIENTATION: x = w.xquat[k & 3][oi[1 + (k >> 2)]]; break;
        case OBS_SITE_ORIENTATION: x = w.site_xmat[k % 6][oi[1 + k / 6]]; break;  // rows 0 and 1 of the row-major 3x3
        case OBS_ACT_POS: x = k == 0 ? w.last_ctrl[oi[0]] : w.qpos[oi[1]]; break;
        default: x = 0.0f; break;
      }
      v[k] = x;
      if (noff >= 0) {
        float y = noise_apply(oi_all, of_all, x, noise_rng, k, w.level);
        if (bi[0] != RV_ZERO) {
          const float b = rv_sample(bi[0], bf, bias_key, k, w.level);
          y = bi[1] ? y * b : y + b;
        }
        out[noff + k] = y;
      }
    }
  }
  WSYNC();
}

// ------------------------------------------------------------------------------------------ terminations

template <class D>
DEV int termination(const Mdl& m, const Work<D>& w, int idx) {
  const int* ti = m.ti;
  const int type = TAB(ti, TI_TERM_TAB, idx, TAB_TYPE);
  const int* xi = ti + TAB(ti, TI_TERM_TAB, idx, TAB_IOFF);
  const float* xf = m.tf + TAB(ti, TI_TERM_TAB, idx, TAB_FOFF);
  switch (type) {
    case TERM_NOT_UPRIGHT: {
      Q4 q; q.w = w.qpos[3]; q.x = w.qpos[4]; q.y = w.qpos[5]; q.z = w.qpos[6];
      V3 r = rot_vec_quat(v3(0, 0, 1), q, true);
      return acosf(r.z) > xf[0] ? -1 : 0;
    }
    case TERM_MINIMUM_HEIGHT: return w.qpos[2] < xf[0] ? -1 : 0;
    case TERM_ILLEGAL_CONTACT: {
      const int ncon = DIM(NCON, MH_NCON);
      const int* pg1 = MI(m, MH_PAIR_GEOM1);
      const int* pg2 = MI(m, MH_PAIR_GEOM2);
      for (int c = 0; c < ncon; c++) {
        const int a = pg1[w.con_pair[c]], b = pg2[w.con_pair[c]];
        int hit = 0;
        for (int k = 0; k < xi[0]; k++) if (a == xi[1 + k] || b == xi[1 + k]) hit = 1;
        if (hit && w.con_dist[c] < xf[0]) return -1;
      }
      return 0;
    }
    case TERM_BAD_Z: {
      const float lo = (xf[1] - xf[0]) * w.level + xf[0], hi = (xf[3] - xf[2]) * w.level + xf[2];
      return (w.qpos[2] < lo || w.qpos[2] > hi) ? -1 : 0;
    }
    case TERM_BAD_VELOCITY: {
      const float n = sqrtf(w.qvel[0] * w.qvel[0] + w.qvel[1] * w.qvel[1] + w.qvel[2] * w.qvel[2]);
      return n > xf[0] ? -1 : 0;
    }
    case TERM_FAR_FROM_ORIGIN: {
      const float n = sqrtf(w.qpos[0] * w.qpos[0] + w.qpos[1] * w.qpos[1] + w.qpos[2] * w.qpos[2]);
      return n > xf[0] ? xi[0] : 0;
    }
    case TERM_TERRAIN_BAD_Z: {
      const float height = w.xpos[2][xi[0]] - fminf(w.xpos[2][xi[1]], w.xpos[2][xi[2]]);
      return height < xf[0] ? -1 : 0;
    }
    case TERM_EPISODE_LENGTH: {
      const int v = w.time > xf[0] ? xi[0] : 0;
      if (xi[1] >= 0 && w.level < (float)xi[1]) return 0;
      return v;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------- fused control step

template <class D>
DEV void write_pre_block(const Mdl& m, Work<D>& w, float* rec_next, uint32_t* act_key, LANE_ARG) {
  const int* ti = m.ti;
  Key rng; rng.k0 = w.keys[0]; rng.k1 = w.keys[1];
  const Key action_rng = key_child(key_fan(rng, LANE_PASS), 0);
  const KeyFan fa = key_fan(action_rng, LANE_PASS);
  const Key obs_rng = key_child(fa, 0), act_rng = key_child(fa, 1);
  if (act_key) { IF_LANE0 { act_key[0] = act_rng.k0; act_key[1] = act_rng.k1; } }
  observe(m, w, obs_rng, rec_next + ti[TI_R_OBS], LANE_PASS);
  LANE_FOR(i, ti[TI_CMD_SIZE]) rec_next[ti[TI_R_CMD] + i] = w.cmd[i];
  IF_LANE0 rec_next[ti[TI_R_LEVEL]] = w.level;
}

template <class D>
DEV void do_reset(const Mdl& m, Work<D>& w, Key reset_rng, Key cmd_rng, Key obs_carry_rng, Key obs_bias_rng, LANE_ARG) {
  const int* ti = m.ti;
  engine_reset(m, w, reset_rng, LANE_PASS);
  DBG_STOP(m, 6);
  const int ncmd = ti[TI_N_CMD];
  for (int c = 0; c < ncmd; c++) {
    const KeyFan fc = key_fan(cmd_rng, LANE_PASS);
    const Key k = key_child(fc, 1);
    cmd_rng = key_child(fc, 0);
    IF_LANE0 cmd_initial(m, w, c, k);
  }
  const int nobs = ti[TI_N_OBS];
  const KeyFan foc = key_fan(obs_carry_rng, LANE_PASS);
  for (int o = 0; o < nobs; o++) obs_carry_initial(m, w, o, key_child(foc, o), LANE_PASS);
  LANE_FOR(o, nobs) {
    const Key k = key_split(obs_bias_rng, o);
    w.keys[2 + 2 * o] = k.k0; w.keys[3 + 2 * o] = k.k1;
  }
  WSYNC();
}

// One control step of one env: engine.step -> terminations -> record -> reset-on-done -> commands -> next obs.
// ksim/task/rl.py:1164-1211 with the observation of step t+1 evaluated at the end of step t (same keys/values).
template <class D>
DEV int env_step(const Mdl& m, Work<D>& w, const float* action_in, float* rec_cur, float* rec_next, uint32_t* act_key,
                 LANE_ARG) {
  const int* ti = m.ti;
  const int nq = DIM(NQ, MH_NQ), nv = DIM(NV, MH_NV), nb = DIM(NB, MH_NBODY), nu = DIM(NU, MH_NU);
  const float* ef = m.tf + ti[TI_ENGINE_FOFF];
  const int na = ti[TI_ACTION_SIZE];
  const float aclip = ef[6];
  STAGE_BEGIN();
  LANE_FOR(i, na) {
    float a = action_in[i];
    if (a != a) a = 0.0f;  // NaN -> 0, then clip (rl.py:998-1001)
    w.action[i] = clipf(a, -aclip, aclip);
  }
  WSYNC();
  Key rng; rng.k0 = w.keys[0]; rng.k1 = w.keys[1];
  const Key physics_rng = key_split(rng, 1), state_rng = key_split(rng, 2);
  STAGE_MARK(w, 10);
  engine_step(m, w, physics_rng, LANE_PASS);
  STAGE_MARK(w, 15);  // double-counts stages 0-9; used as a cross-check only

  const KeyFan fs1 = key_fan(state_rng, LANE_PASS);
  Key cmd_rng = key_child(fs1, 0);
  const Key obs_carry_rng = key_child(fs1, 2), obs_bias_rng = key_child(fs1, 3);
  const Key reset_rng = key_child(fs1, 4), next_rng = key_child(fs1, 5);

  int done = 0, all_not_fail = 1;
  const int nterm = ti[TI_N_TERM];
  for (int k = 0; k < nterm; k++) {
    const int v = termination(m, w, k);
    IF_LANE0 rec_cur[ti[TI_R_TERM] + k] = (float)v;
    done |= (v != 0);
    all_not_fail &= (v != -1);
  }
  const int success = all_not_fail && done;
  int bad = 0;
  LANE_FOR(i, nq) bad |= !isfinite(w.qpos[i]);
  LANE_FOR(i, nv) bad |= !isfinite(w.qvel[i]);
  bad = wany(bad);
  LANE_FOR(i, nq) rec_cur[ti[TI_R_QPOS] + i] = w.qpos[i];
  LANE_FOR(i, nv) rec_cur[ti[TI_R_QVEL] + i] = w.qvel[i];
  LANE_FOR(i, 3 * nb) rec_cur[ti[TI_R_XPOS] + i] = w.xpos[i % 3][i / 3];
  LANE_FOR(i, 4 * nb) rec_cur[ti[TI_R_XQUAT] + i] = w.xquat[i & 3][i >> 2];
  LANE_FOR(i, nu) rec_cur[ti[TI_R_CTRL] + i] = w.ctrl[i];
  LANE_FOR(i, ti[TI_EVENT_SIZE]) rec_cur[ti[TI_R_EVENT] + i] = w.event[i];
  LANE_FOR(i, na) rec_cur[ti[TI_R_ACTION] + i] = w.action[i];
  IF_LANE0 {
    if (bad) w.nonfinite = 1.0f;
    rec_cur[ti[TI_R_DONE]] = (float)done;
    rec_cur[ti[TI_R_SUCCESS]] = (float)success;
    rec_cur[ti[TI_R_TIME]] = w.time;
  }
  WSYNC();
  if (done) {
    do_reset(m, w, reset_rng, cmd_rng, obs_carry_rng, obs_bias_rng, LANE_PASS);
  } else {
    const int ncmd = ti[TI_N_CMD];
    for (int c = 0; c < ncmd; c++) {
      const KeyFan fc = key_fan(cmd_rng, LANE_PASS);
      const Key k = key_child(fc, 1);
      cmd_rng = key_child(fc, 0);
      IF_LANE0 cmd_update(m, w, c, k);
    }
    WSYNC();
  }
  IF_LANE0 { w.keys[0] = next_rng.k0; w.keys[1] = next_rng.k1; }
  WSYNC();
  STAGE_MARK(w, 11);
  write_pre_block(m, w, rec_next, act_key, LANE_PASS);
  STAGE_MARK(w, 12);
  return done;
}

// init keys per env: [reset, cmd, obs_carry, obs_bias, env_rng] x 2 words (oracle/engine.c o_init_envs)
template <class D>
DEV void env_init(const Mdl& m, Work<D>& w, const uint32_t* ik, float level, float* rec0, uint32_t* act_key, LANE_ARG) {
  const int* ti = m.ti;
  IF_LANE0 { w.level = level; w.nonfinite = 0.0f; }
  LANE_FOR(i, ti[TI_EVENT_SIZE]) w.event[i] = 0.0f;
  LANE_FOR(i, ti[TI_ACT_STATE_SIZE]) w.act_state[i] = 0.0f;
  LANE_FOR(i, ti[TI_CMD_SIZE]) w.cmd[i] = 0.0f;
  LANE_FOR(i, ti[TI_OBS_CARRY_SIZE]) w.obs_carry[i] = 0.0f;
  LANE_FOR(i, ti[TI_KEY_SIZE]) w.keys[i] = 0u;
  WSYNC();
  Key kr, kc, ko, kb;
  kr.k0 = ik[0]; kr.k1 = ik[1]; kc.k0 = ik[2]; kc.k1 = ik[3]; ko.k0 = ik[4]; ko.k1 = ik[5]; kb.k0 = ik[6]; kb.k1 = ik[7];
  DBG_STOP(m, 4);
  do_reset(m, w, kr, kc, ko, kb, LANE_PASS);
  DBG_STOP(m, 5);
  IF_LANE0 { w.keys[0] = ik[8]; w.keys[1] = ik[9]; }
  WSYNC();
  write_pre_block(m, w, rec0, act_key, LANE_PASS);
}

}  // namespace b2s
