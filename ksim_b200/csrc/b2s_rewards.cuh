// b200sim: reward evaluation over a finished rollout -- one warp per environment, lanes over time steps.
//
// Replaces get_rewards (ksim/task/rl.py:189-245) and the reward plugins of ksim/rewards.py listed in
// include/b200sim_codes.h (REW_*).  Stateless rewards are evaluated lane-parallel over the T steps; the two
// stateful ones (FeetAirTimeReward rewards.py:993-1085, SparseTargetHeightReward rewards.py:1149-1192) are
// short sequential scans done by lane 0.  rec: [T][N][R]; total: [N][T]; comps: [K][N][T]; carry: [N][C].
#pragma once

#include "b2s_common.cuh"

namespace b2s {

#ifndef TAB
#define TAB(ti, base, i, f) ((ti)[(ti)[base] + (i) * TAB_STRIDE + (f)])
#endif

DEV float exp_kernel_pen(float x, float scale, float sq, float ab) {
  return fabsf(x) * -ab + x * x * -sq + expf(-(x * x) / (2 * scale * scale));
}
DEV float exp_kernel(float x, float scale) { return expf(-(x * x) / (2 * scale * scale)); }
DEV float norm_fn(float x, int l2) { return l2 ? x * x : fabsf(x); }

// Staged copy of an environment's records (b200sim_score): row t of the read-set columns at srec + t * width; col[c] = position of
// record column c in a staged row, or 0xFFFF when the column is not staged (it is then read from global memory).
struct RecStage {
  const float* srec;
  const unsigned short* col;
  int width;
};

DEV const float& rec_at(const float* rec, size_t tstride, const RecStage* stg, int tt, int off) {
  if (stg) {
    const unsigned c = stg->col[off];
    if (c != 0xFFFFu) return stg->srec[tt * stg->width + (int)c];
  }
  return rec[(size_t)tt * tstride + off];
}

// Rewards r0, r0 + rstride, ... of one environment (one warp).  with_total: this warp runs every reward and also forms the clipped
// total (the single-warp pass of the host build and the oracle's order); otherwise only the scaled components are written and
// the caller sums them in component order with env_total (the fused scoring kernel: the rewards of an env are spread over the
// warps of a CTA; carries of different rewards are disjoint).
DEV void env_rewards(const Mdl& m, const float* rec, size_t tstride, int T, float level, float* carry, float* total,
                     float* comps, size_t cstride, LANE_ARG, int r0 = 0, int rstride = 1, bool with_total = true,
                     const RecStage* stg = nullptr) {
  const int* ti = m.ti;
  const float* tf = m.tf;
  const float* ef = tf + ti[TI_ENGINE_FOFF];
  const int nq = m.h[MH_NQ], nv = m.h[MH_NV];
  const int nrew = ti[TI_N_REW];
  const int QP = ti[TI_R_QPOS], QV = ti[TI_R_QVEL], OB = ti[TI_R_OBS], CM = ti[TI_R_CMD], AC = ti[TI_R_ACTION];
  const int DN = ti[TI_R_DONE], SC = ti[TI_R_SUCCESS], LV = ti[TI_R_LEVEL];
  const int na = ti[TI_ACTION_SIZE];
#define REC(tt, off) rec_at(rec, tstride, stg, (tt), (off))
  const int done_last = REC(T - 1, DN) != 0.0f;
  if (with_total) { LANE_FOR(s, T) total[s] = 0.0f; }
  for (int r = r0; r < nrew; r += rstride) {
    const int type = TAB(ti, TI_REW_TAB, r, TAB_TYPE);
    const int* xi = ti + TAB(ti, TI_REW_TAB, r, TAB_IOFF) + SCALE_NI;
    const float* xf = tf + TAB(ti, TI_REW_TAB, r, TAB_FOFF) + SCALE_NF;
    const int c0 = TAB(ti, TI_REW_TAB, r, TAB_OUT_OFF), ncomp = TAB(ti, TI_REW_TAB, r, TAB_OUT_DIM);
    const int coff = TAB(ti, TI_REW_TAB, r, TAB_AUX0), cdim = TAB(ti, TI_REW_TAB, r, TAB_AUX1);
    float scale = scale_get(ti[TAB(ti, TI_REW_TAB, r, TAB_IOFF)], tf + TAB(ti, TI_REW_TAB, r, TAB_FOFF), level);
    if (TAB(ti, TI_REW_TAB, r, TAB_AUX2)) scale = -scale;
    float* o0 = comps + (size_t)c0 * cstride;
#define OUT(c, s) o0[(size_t)(c) * cstride + (s)]
    float* cr = coff >= 0 ? carry + coff : nullptr;
    if (cr && done_last) { LANE_FOR(k, cdim) cr[k] = 0.0f; }
    WSYNC();
    switch (type) {
      case REW_AMP:
        LANE_FOR(s, T) {
          const float d = REC(s, ti[TI_R_AUX] + xi[0]) - 1.0f;
          OUT(0, s) = fmaxf(0.0f, 1.0f - 0.25f * d * d);
        }
        break;
      case REW_POSITION_TRACKING: {
        // rewards.py:715-721: exp_kernel(sum_k norm((xpos[tracked] - xpos[base])_k - command[:3]_k)); command[:3] is the
        // CURRENT point of the CartesianCoordinateCommand
        const int XP = ti[TI_R_XPOS];
        LANE_FOR(s, T) {
          float err = 0.0f;
#pragma unroll
          for (int k = 0; k < 3; k++)
            err += norm_fn((REC(s, XP + 3 * xi[0] + k) - REC(s, XP + 3 * xi[1] + k)) - REC(s, CM + xi[2] + k), xi[3]);
          OUT(0, s) = exp_kernel(err, xf[0]);
        }
      } break;
      case REW_SINUSOIDAL_GAIT:
        // rewards.py:1336-1340: 1 - sum_f |obs[f].z - cmd.height[f]| / max_height
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int f = 0; f < xi[1]; f++) acc += fabsf(REC(s, OB + xi[0] + 3 * f + 2) - REC(s, CM + xi[2] + f));
          OUT(0, s) = 1.0f - acc / xf[0];
        }
        break;
      case REW_JOINT_POSITION: {
        // rewards.py:1370-1385 with JointPositionCommandValue.current_position / current_velocity (commands.py:806-820)
        const int n = xi[0], c = CM + xi[1];
        LANE_FOR(s, T) {
          const float total_time = REC(s, c + 2 * n), num_steps = REC(s, c + 2 * n + 1), step_id = REC(s, c + 2 * n + 2);
          float ap = 0.0f, av = 0.0f;
          for (int i = 0; i < n; i++) {
            const float st0 = REC(s, c + i), d = REC(s, c + n + i) - st0;
            const float cpos = st0 + d * step_id / num_steps, cvel = d / total_time;
            ap += exp_kernel_pen(cpos - REC(s, QP + 7 + xi[2 + i]), xf[0], xf[2], xf[3]);
            av += exp_kernel_pen(cvel - REC(s, QV + 6 + xi[2 + i]), xf[1], xf[2], xf[3]);
          }
          OUT(0, s) = ap / (float)n;
          OUT(1, s) = av / (float)n;
        }
      } break;
      case REW_INTERSECTION:
        // rewards.py:1424-1430: any pair (i > j) of the observed points closer than min_distance
        LANE_FOR(s, T) {
          bool close = false;
          for (int i = 1; i < xi[1]; i++)
            for (int j = 0; j < i; j++) {
              const float dx = REC(s, OB + xi[0] + 3 * i) - REC(s, OB + xi[0] + 3 * j);
              const float dy = REC(s, OB + xi[0] + 3 * i + 1) - REC(s, OB + xi[0] + 3 * j + 1);
              const float dz = REC(s, OB + xi[0] + 3 * i + 2) - REC(s, OB + xi[0] + 3 * j + 2);
              close |= sqrtf(dx * dx + dy * dy + dz * dz) < xf[0];
            }
          OUT(0, s) = close ? 1.0f : 0.0f;
        }
        break;
      case REW_STAY_ALIVE:
        LANE_FOR(s, T) {
          const float alive = 1.0f / xf[0];
          OUT(0, s) = REC(s, DN) != 0.0f ? (REC(s, SC) != 0.0f ? (xi[0] ? xf[1] : alive) : -1.0f) : alive;
        }
        break;
      case REW_UPRIGHT:
        LANE_FOR(s, T) {
          Q4 q; q.w = REC(s, QP + 3); q.x = REC(s, QP + 4); q.y = REC(s, QP + 5); q.z = REC(s, QP + 6);
          V3 zb = rot_vec_quat(v3(0, 0, 1), q, true);
          OUT(0, s) = exp_kernel_pen(REC(s, QV + 4), xf[0], xf[3], xf[4]);
          OUT(1, s) = exp_kernel_pen(REC(s, QV + 2), xf[1], xf[5], xf[6]);
          OUT(2, s) = exp_kernel_pen(atan2f(zb.x, zb.z), xf[2], xf[7], xf[8]);
        }
        break;
      case REW_LINEAR_VELOCITY:
        LANE_FOR(s, T) {
          const float* cmd = &REC(s, CM + xi[0]);
          Q4 q; q.w = REC(s, QP + 3); q.x = REC(s, QP + 4); q.y = REC(s, QP + 5); q.z = REC(s, QP + 6);
          V3 lv = rot_vec_quat(v3(REC(s, QV), REC(s, QV + 1), REC(s, QV + 2)), q, true);
          const float vel = sqrtf(lv.x * lv.x + lv.y * lv.y), yaw = atan2f(lv.y, lv.x);
          const bool is_zero = fabsf(cmd[0]) < xf[2];
          OUT(0, s) = exp_kernel_pen(vel - cmd[0], xf[0], xf[3], xf[4]);
          OUT(1, s) = is_zero ? 0.0f : exp_kernel_pen(yaw - cmd[1], xf[1], xf[3], xf[4]);
          OUT(2, s) = is_zero ? 0.0f : exp_kernel_pen(lv.x - cmd[2], xf[0], xf[3], xf[4]);
          OUT(3, s) = is_zero ? 0.0f : exp_kernel_pen(lv.y - cmd[3], xf[0], xf[3], xf[4]);
        }
        break;
      case REW_ANGULAR_VELOCITY:
        LANE_FOR(s, T) {
          const float cv = REC(s, CM + xi[0]), av = REC(s, QV + 5);
          const float rs = fabsf(av) < xf[3] ? 0.0f : (av > 0 ? 1.0f : (av < 0 ? -1.0f : 0.0f));
          const float cs = fabsf(cv) < xf[3] ? 0.0f : (cv > 0 ? 1.0f : (cv < 0 ? -1.0f : 0.0f));
          OUT(0, s) = exp_kernel_pen(av - cv, xf[0], xf[1], xf[2]);
          OUT(1, s) = rs == cs ? 1.0f : -1.0f;
        }
        break;
#if B2S_DEVICE
      // The two stateful scans of the walking task.  What is serial in them is only the per-foot counters / running maxima; the
      // per-step inputs (gait period, moving / not-moving flags: square roots, scale look-ups, a dozen record reads) are computed by
      // the 32 lanes for 32 steps at once and handed to the scan by shuffle.  Every lane runs the (cheap) scan redundantly on the
      // broadcast inputs, lane j keeps step j's result, and the outputs leave as coalesced rows.  Same operations on the same
      // values in the same order as the one-lane loop of the host build below.
      case REW_FEET_AIR_TIME: {
        const int coffs = xi[0], nf = xi[1], rows = xi[2], lcmd = xi[3], acmd = xi[4], gp_kind = xi[5];
        const float air_pct = xf[0], ctrl_dt = xf[1], bias = xf[2], gpen = xf[4], lthr = xf[5], athr = xf[6];
        for (int s0 = 0; s0 < T; s0 += 32) {
          const int s = s0 + lane;
          float air_steps = 1.0f, gnd_steps = 1.0f;
          int moving = 0;
          unsigned cmask = 0;  // bit f: foot f in contact (nf <= 32: one bit per foot; more feet take the per-step reads below)
          if (s < T) {
            const float gp = scale_get(gp_kind, xf + 7, REC(s, LV));
            air_steps = rintf(gp * air_pct / ctrl_dt); gnd_steps = rintf(gp * (1.0f - air_pct) / ctrl_dt);
            const bool mlin = lcmd < 0 ? (sqrtf(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) > lthr)
                                       : (fabsf(REC(s, CM + lcmd)) > lthr);
            const bool mang = acmd < 0 ? (REC(s, QV + 5) > athr) : (fabsf(REC(s, CM + acmd)) > athr);
            moving = (mlin || mang) && !(REC(s, DN) != 0.0f);
            for (int f = 0; f < nf && f < 32; f++) {
              bool contact = false;
              for (int rr = 0; rr < rows; rr++) contact |= REC(s, OB + coffs + rr * nf + f) > 0.5f;
              cmask |= (unsigned)contact << f;
            }
          }
          float o_air = 0.0f, o_gnd = 0.0f, o_st = 0.0f;
          const int cnt = T - s0 < 32 ? T - s0 : 32;
          for (int j = 0; j < cnt; j++) {
            const float as = __shfl_sync(0xffffffffu, air_steps, j), gs = __shfl_sync(0xffffffffu, gnd_steps, j);
            const unsigned cm = __shfl_sync(0xffffffffu, cmask, j);
            const bool mv = __shfl_sync(0xffffffffu, moving, j) != 0;
            float air_r = -INFINITY, gnd_r = -INFINITY, st_r = -INFINITY;
            for (int f = 0; f < nf; f++) {
              bool contact;
              if (f < 32) contact = (cm >> f) & 1u;
              else { contact = false; for (int rr = 0; rr < rows; rr++) contact |= REC(s0 + j, OB + coffs + rr * nf + f) > 0.5f; }
              float air = cr[f], gnd = cr[nf + f], stg = cr[2 * nf + f];
              air = (!contact && mv) ? air + 1 : 0.0f;
              gnd = (contact && mv) ? gnd + 1 : 0.0f;
              stg = (contact && !mv) ? stg + 1 : 0.0f;
              WSYNC();
              IF_LANE0 { cr[f] = air; cr[nf + f] = gnd; cr[2 * nf + f] = stg; }
              WSYNC();
              const float a = air < as ? air / as + bias : 0.0f;
              const float g = gnd < gs ? gnd / gs + bias : -gpen;
              const float q = clipf(stg / gs, 0.0f, 1.0f) + bias;
              air_r = fmaxf(air_r, a); gnd_r = fmaxf(gnd_r, g); st_r = fmaxf(st_r, q);
            }
            if (lane == j) { o_air = air_r; o_gnd = gnd_r; o_st = st_r; }
          }
          if (s < T) { OUT(0, s) = o_air; OUT(1, s) = o_gnd; OUT(2, s) = o_st; }
        }
      } break;
      case REW_SPARSE_TARGET_HEIGHT: {
        const int coffs = xi[0], poff = xi[1], nbd = xi[2], rows = xi[3];
        for (int s0 = 0; s0 < T; s0 += 32) {
          const int s = s0 + lane;
          int not_moving = 0;
          unsigned cmask = 0;
          if (s < T) {
            const bool nm_lin = sqrtf(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) < xf[1];
            const bool nm_ang = REC(s, QV + 5) < xf[2];
            not_moving = (nm_lin && nm_ang) || (REC(s, DN) != 0.0f);
            for (int b = 0; b < nbd && b < 32; b++) {
              bool contact = false;
              for (int rr = 0; rr < rows; rr++) contact |= REC(s, OB + coffs + rr * nbd + b) > 0.5f;
              cmask |= (unsigned)contact << b;
            }
          }
          float o_best = 0.0f;
          const int cnt = T - s0 < 32 ? T - s0 : 32;
          for (int j = 0; j < cnt; j++) {
            const unsigned cm = __shfl_sync(0xffffffffu, cmask, j);
            const bool nm = __shfl_sync(0xffffffffu, not_moving, j) != 0;
            float best = -INFINITY;
            for (int b = 0; b < nbd; b++) {
              bool contact;
              if (b < 32) contact = (cm >> b) & 1u;
              else { contact = false; for (int rr = 0; rr < rows; rr++) contact |= REC(s0 + j, OB + coffs + rr * nbd + b) > 0.5f; }
              const float h = REC(s0 + j, OB + poff + 3 * b + 2);
              const float cur = cr[b];
              float rw = contact ? cur : 0.0f;
              if (rw > xf[0]) rw = xf[0];
              float mh = cur > h ? cur : h;
              if (nm || contact) mh = 0.0f;
              WSYNC();
              IF_LANE0 cr[b] = mh;
              WSYNC();
              best = fmaxf(best, rw);
            }
            if (lane == j) o_best = best;
          }
          if (s < T) OUT(0, s) = o_best;
        }
      } break;
#else
      case REW_FEET_AIR_TIME:
        IF_LANE0 {
          const int coffs = xi[0], nf = xi[1], rows = xi[2], lcmd = xi[3], acmd = xi[4], gp_kind = xi[5];
          const float air_pct = xf[0], ctrl_dt = xf[1], bias = xf[2], gpen = xf[4], lthr = xf[5], athr = xf[6];
          for (int s = 0; s < T; s++) {
            const float gp = scale_get(gp_kind, xf + 7, REC(s, LV));
            const float air_steps = rintf(gp * air_pct / ctrl_dt), gnd_steps = rintf(gp * (1.0f - air_pct) / ctrl_dt);
            const bool mlin = lcmd < 0 ? (sqrtf(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) > lthr)
                                       : (fabsf(REC(s, CM + lcmd)) > lthr);
            const bool mang = acmd < 0 ? (REC(s, QV + 5) > athr) : (fabsf(REC(s, CM + acmd)) > athr);
            const bool moving = (mlin || mang) && !(REC(s, DN) != 0.0f);
            float air_r = -INFINITY, gnd_r = -INFINITY, st_r = -INFINITY;
            for (int f = 0; f < nf; f++) {
              bool contact = false;
              for (int rr = 0; rr < rows; rr++) contact |= REC(s, OB + coffs + rr * nf + f) > 0.5f;
              float air = cr[f], gnd = cr[nf + f], stg = cr[2 * nf + f];
              air = (!contact && moving) ? air + 1 : 0.0f;
              gnd = (contact && moving) ? gnd + 1 : 0.0f;
              stg = (contact && !moving) ? stg + 1 : 0.0f;
              cr[f] = air; cr[nf + f] = gnd; cr[2 * nf + f] = stg;
              const float a = air < air_steps ? air / air_steps + bias : 0.0f;
              const float g = gnd < gnd_steps ? gnd / gnd_steps + bias : -gpen;
              const float q = clipf(stg / gnd_steps, 0.0f, 1.0f) + bias;
              air_r = fmaxf(air_r, a); gnd_r = fmaxf(gnd_r, g); st_r = fmaxf(st_r, q);
            }
            OUT(0, s) = air_r; OUT(1, s) = gnd_r; OUT(2, s) = st_r;
          }
        }
        break;
      case REW_SPARSE_TARGET_HEIGHT:
        IF_LANE0 {
          const int coffs = xi[0], poff = xi[1], nbd = xi[2], rows = xi[3];
          for (int s = 0; s < T; s++) {
            const bool nm_lin = sqrtf(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) < xf[1];
            const bool nm_ang = REC(s, QV + 5) < xf[2];
            const bool not_moving = (nm_lin && nm_ang) || (REC(s, DN) != 0.0f);
            float best = -INFINITY;
            for (int b = 0; b < nbd; b++) {
              bool contact = false;
              for (int rr = 0; rr < rows; rr++) contact |= REC(s, OB + coffs + rr * nbd + b) > 0.5f;
              const float h = REC(s, OB + poff + 3 * b + 2);
              float rw = contact ? cr[b] : 0.0f;
              if (rw > xf[0]) rw = xf[0];
              float mh = cr[b] > h ? cr[b] : h;
              if (not_moving || contact) mh = 0.0f;
              cr[b] = mh;
              best = fmaxf(best, rw);
            }
            OUT(0, s) = best;
          }
        }
        break;
#endif
      case REW_FORCE_PENALTY:
        LANE_FOR(s, T) {
          float best = -INFINITY;
          for (int rr = 0; rr < xi[1]; rr++) {
            float n2 = 0.0f;
            for (int k = 0; k < xi[2]; k++) { const float v = REC(s, OB + xi[0] + rr * xi[2] + k); n2 += v * v; }
            float o = sqrtf(n2) - xf[2];
            if (o < 0) o = 0.0f;
            const float p = (o / xf[1]) * (o / xf[1]);
            best = fmaxf(best, p);
          }
          OUT(0, s) = best;
        }
        break;
      case REW_MOTIONLESS_AT_REST:
        LANE_FOR(s, T) {
          const bool nm_lin = sqrtf(REC(s, QV) * REC(s, QV) + REC(s, QV + 1) * REC(s, QV + 1)) < xf[0];
          const bool nm_ang = REC(s, QV + 5) < xf[1];
          float n2 = 0.0f;
          for (int k = 6; k < nv; k++) n2 += REC(s, QV + k) * REC(s, QV + k);
          OUT(0, s) = (nm_lin && nm_ang) ? sqrtf(n2) : 0.0f;
        }
        break;
      case REW_ACTION_VELOCITY: case REW_ACTION_ACCELERATION: case REW_ACTION_JERK: {
        // rewards.py:400-463: edge-padded finite differences with done masking (tests/test_rewards.py:60-271)
        const int order = type - REW_ACTION_VELOCITY + 1;
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 0; k < na; k++) {
#define APAD(p) REC(((p) - order) > 0 ? ((p) - order) : 0, AC + k)
#define DPAD(p) (REC(((p) - order) > 0 ? ((p) - order) : 0, DN) != 0.0f)
#define VEL(i) (DPAD(i) ? 0.0f : APAD((i) + 1) - APAD(i))
#define ACC(i) (DPAD((i) + 1) ? 0.0f : VEL((i) + 1) - VEL(i))
#define JRK(i) (DPAD((i) + 2) ? 0.0f : ACC((i) + 1) - ACC(i))
            const float v = order == 1 ? VEL(s) : (order == 2 ? ACC(s) : JRK(s));
            acc += norm_fn(v, xi[0]);
#undef APAD
#undef DPAD
#undef VEL
#undef ACC
#undef JRK
          }
          OUT(0, s) = s >= order ? acc / (float)na : 0.0f;
        }
      } break;
      case REW_BASE_HEIGHT:
        LANE_FOR(s, T) OUT(0, s) = exp_kernel(REC(s, QP + 2) - xf[0], xf[1]);
        break;
      case REW_BASE_HEIGHT_RANGE:
        LANE_FOR(s, T) {
          const float h = REC(s, QP + 2), lo = xf[0] - h, hi = h - xf[1];
          float mx = fmaxf(fmaxf(lo, hi), 0.0f);
          const float v = 1.0f - mx * xf[2];
          OUT(0, s) = v < 0 ? 0.0f : v;
        }
        break;
      case REW_OFF_AXIS_VELOCITY:
        LANE_FOR(s, T) {
          OUT(0, s) = exp_kernel(REC(s, QV + 2), xf[0]);
          OUT(1, s) = exp_kernel(REC(s, QV + 4), xf[1]);
          OUT(2, s) = exp_kernel(REC(s, QV + 5), xf[1]);
        }
        break;
      case REW_SMALL_JOINT_VELOCITY:
        LANE_FOR(s, T) {
          float acc = 0.0f;
          const int prev = s > 0 ? s - 1 : 0;
          const bool dn = REC(prev, DN) != 0.0f;
          for (int k = 7; k < nq; k++) {
            const float v = (dn || s == 0) ? 0.0f : REC(s, QP + k) - REC(prev, QP + k);
            acc += exp_kernel(v, xf[0]);
          }
          OUT(0, s) = acc / (float)(nq - 7);
        }
        break;
      // ---- more of ksim/rewards.py (same structure as oracle/engine.c) ----
      case REW_SMALL_JOINT_ACCELERATION: case REW_SMALL_JOINT_JERK: {
        // rewards.py:479-508: edge-padded differences of qpos[7:] with done masking; no mask on the first steps
        const int order = type == REW_SMALL_JOINT_ACCELERATION ? 2 : 3;
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 7; k < nq; k++) {
#define QPAD(p) REC(((p) - order) > 0 ? ((p) - order) : 0, QP + k)
#define DPAD(p) (REC(((p) - order) > 0 ? ((p) - order) : 0, DN) != 0.0f)
#define VEL(i) (DPAD(i) ? 0.0f : QPAD((i) + 1) - QPAD(i))
#define ACC(i) (DPAD((i) + 1) ? 0.0f : VEL((i) + 1) - VEL(i))
#define JRK(i) (DPAD((i) + 2) ? 0.0f : ACC((i) + 1) - ACC(i))
            const float v = order == 2 ? ACC(s) : JRK(s);
            acc += exp_kernel(v, xf[0]);
#undef QPAD
#undef VEL
#undef ACC
#undef JRK
          }
          OUT(0, s) = acc / (float)(nq - 7);
        }
      } break;
      case REW_LINK_ACCELERATION: case REW_LINK_JERK: {
        // rewards.py:809-841: body speeds from edge-padded position differences, second / third difference
        const int order = type == REW_LINK_ACCELERATION ? 2 : 3;
        const int XP = ti[TI_R_XPOS], nb = xi[1];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int b = 1; b < nb; b++) {
#define TPAD(p) (((p) - order) > 0 ? ((p) - order) : 0)
#define PDIF(i, c) (REC(TPAD((i) + 1), XP + 3 * b + (c)) - REC(TPAD(i), XP + 3 * b + (c)))
#define VEL(i) (DPAD(i) ? 0.0f : sqrtf(PDIF(i, 0) * PDIF(i, 0) + PDIF(i, 1) * PDIF(i, 1) + PDIF(i, 2) * PDIF(i, 2)))
#define ACC(i) (DPAD((i) + 1) ? 0.0f : VEL((i) + 1) - VEL(i))
#define JRK(i) (DPAD((i) + 2) ? 0.0f : ACC((i) + 1) - ACC(i))
            const float v = order == 2 ? ACC(s) : JRK(s);
            acc += norm_fn(v, xi[0]);
#undef TPAD
#undef PDIF
#undef VEL
#undef ACC
#undef JRK
#undef DPAD
          }
          OUT(0, s) = acc / (float)(nb - 1);
        }
      } break;
      case REW_AVOID_LIMITS: {
        const int n = xi[0];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 0; k < n; k++) {
            const float q = REC(s, QP + 7 + k), a = q - xf[k], b = q - xf[n + k];
            acc += -(a < 0.0f ? a : 0.0f) + (b > 0.0f ? b : 0.0f);
          }
          OUT(0, s) = acc / (float)n;
        }
      } break;
      case REW_TORQUE: case REW_ENERGY: {
        const int n = xi[0], CT = ti[TI_R_CTRL];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 0; k < n; k++) {
            float c = REC(s, CT + k) / xf[k];
            if (type == REW_ENERGY) c = c * REC(s, QV + 6 + k);
            acc += fabsf(c);
          }
          OUT(0, s) = acc / (float)n;
        }
      } break;
      case REW_JOINT_DEVIATION: {
        const int n = xi[1];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 0; k < n; k++) {
            float d = REC(s, QP + 7 + xi[3 + k]) - xf[k];
            if (xi[2]) d *= xf[n + k];
            acc += norm_fn(d, xi[0]);
          }
          OUT(0, s) = acc;
        }
      } break;
      case REW_FLAT_BODY: {
        const int n = xi[0], XQ = ti[TI_R_XQUAT];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          const V3 pl = v3(xf[0], xf[1], xf[2]);
          for (int k = 0; k < n; k++) {
            const int o = XQ + 4 * xi[1 + k];
            Q4 q; q.w = REC(s, o); q.x = REC(s, o + 1); q.y = REC(s, o + 2); q.z = REC(s, o + 3);
            acc += dot(rot_vec_quat(pl, q, true), pl);
          }
          OUT(0, s) = acc / (float)n;
        }
      } break;
      case REW_NO_ROLL:
        LANE_FOR(s, T) {
          Q4 q; q.w = REC(s, QP + 3); q.x = REC(s, QP + 4); q.y = REC(s, QP + 5); q.z = REC(s, QP + 6);
          const V3 zb = rot_vec_quat(v3(0, 0, 1), q, true);
          OUT(0, s) = exp_kernel_pen(REC(s, QV + 3), xf[0], xf[2], xf[3]);
          OUT(1, s) = exp_kernel_pen(-atan2f(zb.y, zb.z), xf[1], xf[4], xf[5]);
        }
        break;
      case REW_FEET_SLIP: {
        const int co = xi[0], rows = xi[1], nf = xi[2], vo = xi[3];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int f = 0; f < nf; f++) {
            bool contact = false;
            for (int c = 0; c < rows; c++) contact |= REC(s, OB + co + c * nf + f) > 0.5f;
            float n2 = 0.0f;
            for (int k = 0; k < 6; k++) { const float v = REC(s, OB + vo + 6 * f + k); n2 += v * v; }
            if (contact) acc += sqrtf(n2);
          }
          OUT(0, s) = acc;
        }
      } break;
      case REW_REACHABILITY: {
        const int n = xi[1];
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 0; k < n; k++) {
            float ex = fabsf(REC(s, AC + k) - REC(s, QP + 7 + k)) - xf[k];
            if (ex < 0.0f) ex = 0.0f;
            acc += xi[0] ? ex * ex : ex;
          }
          OUT(0, s) = acc;
        }
      } break;
      case REW_SYMMETRY: {
        // stateful (rewards.py:844-897): one carry update per rollout from the time mean, then the per-step product;
        // the carry is kept relative to the targets so that it starts and resets at 0 like every other carry
        const int n = xi[0];
        for (int k = 0; k < n; k++) {
          float mean = 0.0f;
          LANE_FOR(s, T) mean += REC(s, QP + 7 + xi[1 + k]) - xf[1 + k];
          mean = wsum(mean) / (float)T;
          WSYNC();
          IF_LANE0 cr[k] = (cr[k] + xf[1 + k]) * (1.0f - xf[0]) + mean * xf[0] - xf[1 + k];
        }
        WSYNC();
        LANE_FOR(s, T) {
          float acc = 0.0f;
          for (int k = 0; k < n; k++) acc += (REC(s, QP + 7 + xi[1 + k]) - xf[1 + k]) * -(cr[k] + xf[1 + k]);
          OUT(0, s) = acc;
        }
      } break;
      case REW_PAIRWISE_SYMMETRY: {
        float m0 = 0.0f, m1 = 0.0f;
        LANE_FOR(s, T) {
          const float l = REC(s, QP + 7 + xi[0]) - xf[0];
          float r_ = REC(s, QP + 7 + xi[1]) - xf[1];
          if (xi[2]) r_ = -r_;
          m0 += l; m1 += r_;
        }
        m0 = wsum(m0) / (float)T; m1 = wsum(m1) / (float)T;
        WSYNC();
        IF_LANE0 {
          cr[0] = (cr[0] + xf[0]) * (1.0f - xf[2]) + m0 * xf[2] - xf[0];
          cr[1] = (cr[1] + xf[1]) * (1.0f - xf[2]) + m1 * xf[2] - xf[1];
        }
        WSYNC();
        LANE_FOR(s, T) {
          const float l = REC(s, QP + 7 + xi[0]) - xf[0];
          float r_ = REC(s, QP + 7 + xi[1]) - xf[1];
          if (xi[2]) r_ = -r_;
          OUT(0, s) = exp_kernel_pen(l - r_, xf[3], xf[4], xf[5]);
          OUT(1, s) = l * -(cr[0] + xf[0]) + r_ * -(cr[1] + xf[1]);
        }
      } break;
      // ---- the K-Bot task's own rewards (examples/kbot/train.py:128-496; same structure as oracle/engine.c)
#define UCMD(s, k) REC(s, CM + ucmd + (k))
#define ZERO_CMD(s) (sqrtf(UCMD(s, 0) * UCMD(s, 0) + UCMD(s, 1) * UCMD(s, 1) + UCMD(s, 2) * UCMD(s, 2)) < 1e-3f)
      case REW_KBOT_SINGLE_FOOT_CONTACT: {
        const int ucmd = xi[2];
        IF_LANE0 {
          for (int s = 0; s < T; s++) {
            const bool l = REC(s, OB + xi[0]) > 0.1f, r_ = REC(s, OB + xi[1]) > 0.1f, zero = ZERO_CMD(s);
            float tm = (l != r_) ? 0.0f : cr[0] + xf[0];
            if (zero) tm = xf[1];
            cr[0] = tm;
            OUT(0, s) = zero ? 1.0f : (tm < xf[1] ? 1.0f : 0.0f);
          }
        }
      } break;
      case REW_KBOT_NO_CONTACT: {
        const int ucmd = xi[2];
        LANE_FOR(s, T) {
          const bool l = REC(s, OB + xi[0]) > 0.1f, r_ = REC(s, OB + xi[1]) > 0.1f;
          OUT(0, s) = ZERO_CMD(s) ? 0.0f : ((l || r_) ? 0.0f : 1.0f);
        }
      } break;
      case REW_KBOT_FEET_AIRTIME: {
        const int ucmd = xi[2];
        IF_LANE0 {
          for (int s = 0; s < T; s++) {
            const bool dn = REC(s, DN) != 0.0f;
            float rew = 0.0f;
            for (int f = 0; f < 2; f++) {
              const bool contact = REC(s, OB + xi[f]) > 0.1f;
              const bool prev_contact = cr[2 + f] == 0.0f;
              const bool first = contact && !prev_contact && !dn;
              rew += (cr[f] - xf[1]) * (first ? 1.0f : 0.0f);
              cr[f] = (contact || dn) ? 0.0f : cr[f] + xf[0];
              cr[2 + f] = contact ? 0.0f : 1.0f;
            }
            OUT(0, s) = ZERO_CMD(s) ? 0.0f : rew;
          }
        }
      } break;
      case REW_KBOT_ARM_POSITION: {
        const int ucmd = xi[0];
        LANE_FOR(s, T) {
          float err = 0.0f;
          for (int k = 0; k < 10; k++) {
            const float dlt = REC(s, QP + xi[1 + k]) - (UCMD(s, 6 + k) + xf[1 + k]);
            err += dlt * dlt;
          }
          OUT(0, s) = expf(-err / xf[0]);
        }
      } break;
      case REW_KBOT_LINVEL_TRACKING: {
        const int ucmd = xi[0], XQ = ti[TI_R_XQUAT];
        LANE_FOR(s, T) {
          Q4 bq; bq.w = REC(s, XQ + 4); bq.x = REC(s, XQ + 5); bq.y = REC(s, XQ + 6); bq.z = REC(s, XQ + 7);
          const Q4 zq = euler_to_quat(0.0f, 0.0f, quat_to_euler(bq).z);
          const V3 g = rot_vec_quat(v3(UCMD(s, 0), UCMD(s, 1), 0.0f), zq, false);
          const float ex = REC(s, QV) - g.x, ey = REC(s, QV + 1) - g.y;
          const float ve = sqrtf(ex * ex + ey * ey);
          OUT(0, s) = expf(-(ZERO_CMD(s) ? ve : ve * ve) / xf[0]);
        }
      } break;
      case REW_KBOT_ANGVEL_TRACKING: {
        const int ucmd = xi[0];
        LANE_FOR(s, T) OUT(0, s) = expf(-fabsf(REC(s, QV + 5) - UCMD(s, 2)) / xf[0]);
      } break;
      case REW_KBOT_XY_ORIENTATION: {
        const int ucmd = xi[0], XQ = ti[TI_R_XQUAT];
        LANE_FOR(s, T) {
          Q4 bq; bq.w = REC(s, XQ + 4); bq.x = REC(s, XQ + 5); bq.y = REC(s, XQ + 6); bq.z = REC(s, XQ + 7);
          const V3 e = quat_to_euler(bq);
          const float dt_ = qdot(euler_to_quat(UCMD(s, 4), UCMD(s, 5), 0.0f), euler_to_quat(e.x, e.y, 0.0f));
          OUT(0, s) = expf(-(1.0f - dt_ * dt_) / (ZERO_CMD(s) ? xf[1] : xf[0]));
        }
      } break;
      case REW_KBOT_TERRAIN_BASE_HEIGHT: {
        const int ucmd = xi[0], XP = ti[TI_R_XPOS];
        LANE_FOR(s, T) {
          const float lf = REC(s, XP + 3 * xi[2] + 2) - xf[2], rf_ = REC(s, XP + 3 * xi[3] + 2) - xf[2];
          const float cur = REC(s, XP + 3 * xi[1] + 2) - fminf(lf, rf_);
          OUT(0, s) = expf(-fabsf(cur - (UCMD(s, 3) + xf[1])) / xf[0]);
        }
      } break;
      case REW_KBOT_FEET_ORIENTATION: {
        // `is_rotating` is one flag per trajectory: the norm of the yaw-rate command column over time (train.py:464)
        const int ucmd = xi[0], XQ = ti[TI_R_XQUAT];
        float n2 = 0.0f;
        LANE_FOR(s, T) n2 += UCMD(s, 2) * UCMD(s, 2);
        const bool rotating = sqrtf(wsum(n2)) > 1e-3f;
        const float hp = 1.57079632679489661923f, pi = 3.14159265358979323846f;
        LANE_FOR(s, T) {
          Q4 bq; bq.w = REC(s, XQ + 4); bq.x = REC(s, XQ + 5); bq.y = REC(s, XQ + 6); bq.z = REC(s, XQ + 7);
          const float yaw = quat_to_euler(bq).z;
          float err = 0.0f;
          for (int f = 0; f < 2; f++) {
            const int fo = XQ + 4 * xi[1 + f];
            Q4 fq; fq.w = REC(s, fo); fq.x = REC(s, fo + 1); fq.y = REC(s, fo + 2); fq.z = REC(s, fo + 3);
            const float roll = f == 0 ? -hp : hp;
            float dt_;
            if (rotating) {
              const V3 fe = quat_to_euler(fq);
              dt_ = qdot(euler_to_quat(roll, 0.0f, 0.0f), euler_to_quat(fe.x, fe.y, 0.0f));
            } else {
              dt_ = qdot(euler_to_quat(roll, 0.0f, yaw - pi), fq);
            }
            err += 1.0f - dt_ * dt_;
          }
          OUT(0, s) = expf(-err / xf[0]);
        }
      } break;
      case REW_KBOT_COM_DISTANCE: {
        const int ucmd = xi[0];
        LANE_FOR(s, T) {
          const float dist = REC(s, OB + xi[1]);
          OUT(0, s) = (dist >= 0.0f && ZERO_CMD(s)) ? expf(-dist / xf[0]) : 0.0f;
        }
      } break;
      case REW_KBOT_BASE_ACCELERATION:
        LANE_FOR(s, T) {
          const int prev = s > 0 ? s - 1 : 0;
          const bool dn = REC(prev, DN) != 0.0f;
          float err = 0.0f;
          for (int k = 0; k < 6; k++) err += (dn || s == 0) ? 0.0f : fabsf(REC(s, QV + k) - REC(prev, QV + k));
          OUT(0, s) = expf(-err / xf[0]);
        }
        break;
#undef UCMD
#undef ZERO_CMD
      default:
        for (int c = 0; c < ncomp; c++) { LANE_FOR(s, T) OUT(c, s) = 0.0f; }
        break;
    }
    WSYNC();
    for (int c = 0; c < ncomp; c++) {
      LANE_FOR(s, T) {
        const float v = OUT(c, s) * scale;
        OUT(c, s) = v;
        if (with_total) total[s] += v;
      }
    }
    WSYNC();
#undef OUT
  }
  if (with_total) {
    LANE_FOR(s, T) {
      float sum = total[s];
      if (ef[7] == ef[7] && sum < ef[7]) sum = ef[7];
      if (ef[8] == ef[8] && sum > ef[8]) sum = ef[8];
      total[s] = sum;
    }
  }
#undef REC
}

// Total reward of step s from the scaled components, summed in component (= reward, then component) order from 0.0f and clipped:
// the same order of additions as env_rewards(with_total).
DEV float env_total(const Mdl& m, const float* comps, size_t cstride, int s) {
  const int* ti = m.ti;
  const float* ef = m.tf + ti[TI_ENGINE_FOFF];
  const int ncomp = ti[TI_N_REW_COMP];
  float sum = 0.0f;
  for (int k = 0; k < ncomp; k++) sum += comps[(size_t)k * cstride + s];
  if (ef[7] == ef[7] && sum < ef[7]) sum = ef[7];
  if (ef[8] == ef[8] && sum > ef[8]) sum = ef[8];
  return sum;
}

}  // namespace b2s
