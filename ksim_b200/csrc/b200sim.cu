// b200sim: kernels + C-ABI (include/b200sim.h).  sm_100a only; no CPU path -- every entry point launches CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b200sim.h"
#include "b2s_engine.cuh"
#include "b2s_rewards.cuh"
#include "b2s_host.cuh"
#include "b2s_tc.cuh"

using namespace b2s;


// ------------------------------------------------------------------------------------------------ kernels

#ifdef B2S_PROFILE_STAGES
__device__ unsigned long long g_stage_cycles[B2S_NSTAGE];
extern "C" int b200sim_debug_stage_cycles(unsigned long long* out, int reset) {
  if (out && cudaMemcpyFromSymbol(out, g_stage_cycles, sizeof(g_stage_cycles)) != cudaSuccess) return 1;
  if (reset) { unsigned long long z[B2S_NSTAGE] = {0}; if (cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z)) != cudaSuccess) return 1; }
  return 0;
}
#endif

template <class D>
__global__ void __launch_bounds__(launch_threads<D>()) k_init_envs(const Mdl mp, int n_envs, const uint32_t* __restrict__ init_keys,
                                                   const float* __restrict__ levels, const float* __restrict__ rand,
                                                   float* __restrict__ state, uint32_t* __restrict__ keys,
                                                   float* __restrict__ rec0, uint32_t* __restrict__ act_keys) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int env = blockIdx.x * wpb + warp;
  const Mdl& m = stage_model<D>(mp, smem, wpb);
  if (env >= n_envs) return;
  Work<D>& w = reinterpret_cast<Work<D>*>(smem)[warp];
  POISON_WORK(mp, w);
  const int* ti = m.ti;
  const size_t S = ti[TI_STATE_SIZE], R = ti[TI_REC_SIZE], K = ti[TI_KEY_SIZE], RAND = ti[TI_RAND_SIZE];
  DBG_STOP(m, 1);
  apply_overrides(m, w, rand + env * RAND, lane);
  DBG_STOP(m, 2);
  env_init(m, w, init_keys + (size_t)env * 10, levels[env], rec0 + env * R, act_keys ? act_keys + 2 * env : nullptr, lane);
  DBG_STOP(m, 3);
  env_store(m, w, state + env * S, keys + env * K, lane);
}

// n_steps control steps per launch; the environment stays in shared memory between them.
template <class D>
__global__ void __launch_bounds__(launch_threads<D>()) k_rollout(const Mdl mp, int n_envs, int n_steps, const float* __restrict__ actions,
                                                 const float* __restrict__ rand, float* __restrict__ state,
                                                 uint32_t* __restrict__ keys, float* __restrict__ rec,
                                                 float* __restrict__ rec_next_single, uint32_t* __restrict__ act_keys) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int env = blockIdx.x * wpb + warp;
  const Mdl& m = stage_model<D>(mp, smem, wpb);
  const int* ti = m.ti;
  if (env >= n_envs) {
    // idle warps of the last CTA still take part in the per-sub-step CTA barriers and votes
    const int nsub = ti[TI_NSUB];
    for (int i = 0; i < n_steps * nsub; i++) {
      if ((m.sync_mask >> 6) & 1) CTA_SYNC();
      forward_follower(m);
    }
    return;
  }
  Work<D>& w = reinterpret_cast<Work<D>*>(smem)[warp];
  POISON_WORK(mp, w);
  const size_t S = ti[TI_STATE_SIZE], R = ti[TI_REC_SIZE], K = ti[TI_KEY_SIZE], RAND = ti[TI_RAND_SIZE];
  const size_t NA = ti[TI_ACTION_SIZE];
  apply_overrides(m, w, rand + env * RAND, lane);
  env_load(m, w, state + env * S, keys + env * K, lane);
#ifdef B2S_PROFILE_STAGES
  if (lane < B2S_NSTAGE) w.prof[lane] = 0;
  __syncwarp();
#endif
  for (int t = 0; t < n_steps; t++) {
    float* cur = rec + ((size_t)t * n_envs + env) * R;
    float* nxt = rec_next_single ? rec_next_single + env * R : rec + ((size_t)(t + 1) * n_envs + env) * R;
    env_step(m, w, actions + ((size_t)t * n_envs + env) * NA, cur, nxt, act_keys ? act_keys + 2 * env : nullptr, lane);
  }
  env_store(m, w, state + env * S, keys + env * K, lane);
#ifdef B2S_PROFILE_STAGES
  __syncwarp();
  if (lane < B2S_NSTAGE) atomicAdd(&g_stage_cycles[lane], (unsigned long long)w.prof[lane]);
#endif
}

// Scoring pass of a rollout in ONE launch (north_star: "fuse reward/termination evaluation and the reverse-time GAE scan"):
// get_rewards (rl.py:189-245), the done / success flags, and -- when values are given -- compute_ppo_inputs (ppo.py:54-121).
// One CTA per environment:
//   * the env's T records are rows of rec[T][N][R], N R floats apart.  The columns the rewards read (the program's scoring
//     read-set: TI_SCORE_SEG) are STAGED in shared memory by bulk async copies (cp.async.bulk, one per row and segment, all in
//     flight at once, completing on one mbarrier): HBM sees full-segment sequential reads instead of one 4-byte field per
//     32-byte sector per lane, and the stateful rewards' serial scans over time (one lane walking T steps) run out of shared
//     memory instead of paying a DRAM round trip per dependent load -- that serial chain was what the round-1 kernel's time was;
//   * the env's rewards are spread over the CTA's warps (reward r on warp r mod SCORE_WARPS; carries of different rewards are
//     disjoint); after a CTA barrier the threads turn to time steps: total = ordered sum of the components, flags and the
//     values row, coalesced along [N][T]; one thread runs the reverse-time recurrence out of shared memory (T dependent steps in
//     the oracle's order of operations) and the CTA writes advantages / targets / returns coalesced.
// Unstaged (staged == 0: the default, see b200sim_score; also when T staged rows do not fit the SM) the same code reads the
// records from global memory through L1.
constexpr int SCORE_WARPS_MAX = 8;  // CTA size is a launch parameter (SCORE_WARPS_MAX warps at most): fewer warps per CTA = more envs per SM
__global__ void __launch_bounds__(SCORE_WARPS_MAX * 32) k_score(const Mdl m, int n_envs, int n_steps, int staged, const float* rec,
                                                            const float* __restrict__ levels, float* carry, const float* __restrict__ values,
                                                            float gamma, float lam, int flags, float* total, float* comps,
                                                            uint8_t* __restrict__ dones, uint8_t* __restrict__ successes,
                                                            float* __restrict__ adv, float* __restrict__ targets, float* __restrict__ returns) {
  extern __shared__ __align__(16) unsigned char sc_raw[];  // [mbarrier][staged rows T x width][6 T floats for the scan][carries][column table R]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = blockDim.x, SCORE_WARPS = NT >> 5;
  const int env = blockIdx.x;
  const int* ti = m.ti;
  const int R = ti[TI_REC_SIZE];
  const size_t C = ti[TI_REW_CARRY_SIZE];
  const int T = n_steps;
  const size_t cstride = (size_t)n_envs * T, row = (size_t)env * T, tstride = (size_t)n_envs * R;
  const float* erec = rec + (size_t)env * R;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sc_raw);
  float* srec = reinterpret_cast<float*>(sc_raw + 16);
  const int width = staged ? ti[TI_SCORE_WIDTH] : 0;
  float* s_scan = srec + (size_t)T * width;
  float* s_carry = s_scan + (values ? 6 * T : 0);  // the rewards' carries: their scans are read-modify-write chains over T steps
  unsigned short* s_col = reinterpret_cast<unsigned short*>(s_carry + ((C + 3) & ~(size_t)3));
  for (int k = tid; k < (int)C; k += NT) s_carry[k] = carry[env * C + k];
  RecStage stage = {srec, s_col, width};
  if (staged) {
    const int nseg = ti[TI_SCORE_NSEG];
    const int* seg = ti + ti[TI_SCORE_SEG];
    if (warp == 0) {
      if (lane == 0) {
        b2s_tc::mbar_init(bar, 1);
        b2s_tc::mbar_fence_init();
        int cols = 0;
        for (int g = 0; g < nseg; g++) cols += seg[2 * g + 1];
        b2s_tc::mbar_arrive_expect_tx(bar, (uint32_t)((size_t)T * cols * sizeof(float)));
      }
      __syncwarp();
      for (int idx = lane; idx < T * nseg; idx += 32) {
        const int t = idx / nseg, g = idx - t * nseg;
        int dst = 0;
        for (int k = 0; k < g; k++) dst += seg[2 * k + 1];
        b2s_tc::bulk_g2s(srec + (size_t)t * width + dst, erec + (size_t)t * tstride + seg[2 * g], (uint32_t)(seg[2 * g + 1] * sizeof(float)), bar);
      }
    }
    for (int c = tid; c < R; c += NT) {  // column table, built while the copies fly
      unsigned short pos = 0xFFFFu;
      int dst = 0;
      for (int g = 0; g < nseg; g++) {
        if (c >= seg[2 * g] && c < seg[2 * g] + seg[2 * g + 1]) pos = (unsigned short)(dst + c - seg[2 * g]);
        dst += seg[2 * g + 1];
      }
      s_col[c] = pos;
    }
    __syncthreads();  // the barrier is initialised and the table complete for every thread
    if (!b2s_tc::mbar_wait(bar, 0, 1u << 26)) __trap();  // a lost bulk copy ends the launch with an error, not a hang
  }
  const RecStage* stg = staged ? &stage : nullptr;
  __syncthreads();
  env_rewards(m, erec, tstride, T, levels[env], s_carry, nullptr, comps + row, cstride, lane, warp, SCORE_WARPS, false, stg);
  __syncthreads();
  for (int k = tid; k < (int)C; k += NT) carry[env * C + k] = s_carry[k];
  float* s_rew = s_scan, *s_val = s_scan + T, *s_mask = s_scan + 2 * T, *s_trunc = s_scan + 3 * T;
  float* s_adv = s_scan + 4 * T, *s_ret = s_scan + 5 * T;
  const int DN = ti[TI_R_DONE], SC = ti[TI_R_SUCCESS];
  for (int s = tid; s < T; s += NT) {
    const float tot = env_total(m, comps + row, cstride, s);
    total[row + s] = tot;
    const bool d = rec_at(erec, tstride, stg, s, DN) != 0.0f, su = rec_at(erec, tstride, stg, s, SC) != 0.0f;
    if (dones) dones[row + s] = d;
    if (successes) successes[row + s] = su;
    if (values) { s_rew[s] = tot; s_val[s] = values[row + s]; s_mask[s] = d ? 0.0f : 1.0f; s_trunc[s] = su ? 1.0f : 0.0f; }
  }
  if (!values) return;
  __syncthreads();
  if (tid == 0) {
    float ret_next = 0.0f, adv_next = 0.0f, v_next = s_val[T - 1];
    for (int s = T - 1; s >= 0; s--) {
      const float v = s_val[s], mask = s_mask[s];
      const float br = __fadd_rn(__fmul_rn(1.0f - gamma, s_rew[s]), __fmul_rn(__fmul_rn(gamma, v), s_trunc[s]));
      const float delta = __fsub_rn(__fadd_rn(br, __fmul_rn(__fmul_rn(gamma, v_next), mask)), v);
      const float ret = __fadd_rn(br, __fmul_rn(__fmul_rn(gamma, mask), ret_next));
      const float a = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(gamma, lam), mask), adv_next));
      s_ret[s] = ret; s_adv[s] = a;
      ret_next = ret; adv_next = a; v_next = v;
    }
  }
  __syncthreads();
  for (int s = tid; s < T; s += NT) {
    const float a = s_adv[s], ret = s_ret[s];
    returns[row + s] = ret;
    adv[row + s] = a;
    targets[row + s] = (flags & GAE_MONTE_CARLO_RETURNS) ? ret : __fadd_rn(a, s_val[s]);
  }
}

__global__ void k_extract_flags(int n_envs, int n_steps, size_t R, int off_done, int off_success,
                                const float* __restrict__ rec, uint8_t* __restrict__ dones, uint8_t* __restrict__ successes) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_envs * n_steps) return;
  const int n = idx / n_steps, t = idx - (size_t)n * n_steps;
  const float* r = rec + ((size_t)t * n_envs + n) * R;
  dones[idx] = r[off_done] != 0.0f;
  successes[idx] = r[off_success] != 0.0f;
}

// compute_ppo_inputs (ksim/task/ppo.py:54-121) on its own (the PPO update recomputes it from the new values, ppo.py:597-601):
// one warp per trajectory row.  The row's inputs are read coalesced ([B][T]: lanes along time) into shared memory, lane 0 runs
// the reverse-time recurrence there (T dependent steps in the oracle's order of operations), the warp writes the three outputs
// coalesced.  (The round-1 kernel ran one THREAD per row: every access had stride T.)
constexpr int GAE_WARPS = 4;
__global__ void __launch_bounds__(GAE_WARPS * 32) k_gae(int batch, int n_steps, const float* __restrict__ values, const float* __restrict__ rewards,
                      const uint8_t* __restrict__ dones, const uint8_t* __restrict__ successes, float gamma, float lam,
                      int flags, float* __restrict__ adv, float* __restrict__ targets, float* __restrict__ returns) {
  extern __shared__ float gae_smem[];  // per warp [5][T]: br, value, gamma * mask | advantage, return
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * GAE_WARPS + warp;
  if (b >= batch) return;
  const int T = n_steps;
  float* s_br = gae_smem + (size_t)warp * 5 * T, *s_val = s_br + T, *s_gm = s_br + 2 * T, *s_adv = s_br + 3 * T, *s_ret = s_br + 4 * T;
  const size_t base = (size_t)b * T;
  for (int s = lane; s < T; s += 32) {
    const float v = values[base + s];
    const float trunc = successes[base + s] ? 1.0f : 0.0f;
    s_br[s] = __fadd_rn(__fmul_rn(1.0f - gamma, rewards[base + s]), __fmul_rn(__fmul_rn(gamma, v), trunc));
    s_val[s] = v;
    s_gm[s] = dones[base + s] ? 0.0f : 1.0f;
  }
  __syncwarp();
  if (lane == 0) {
    float ret_next = 0.0f, adv_next = 0.0f, v_next = s_val[T - 1];
    for (int s = T - 1; s >= 0; s--) {
      const float v = s_val[s], mask = s_gm[s], br = s_br[s];
      const float delta = __fsub_rn(__fadd_rn(br, __fmul_rn(__fmul_rn(gamma, v_next), mask)), v);
      const float ret = __fadd_rn(br, __fmul_rn(__fmul_rn(gamma, mask), ret_next));
      const float a = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(gamma, lam), mask), adv_next));
      s_ret[s] = ret; s_adv[s] = a;
      ret_next = ret; adv_next = a; v_next = v;
    }
  }
  __syncwarp();
  for (int s = lane; s < T; s += 32) {
    const float a = s_adv[s], ret = s_ret[s];
    returns[base + s] = ret;
    adv[base + s] = a;
    targets[base + s] = (flags & GAE_MONTE_CARLO_RETURNS) ? ret : __fadd_rn(a, s_val[s]);
  }
}

// advantage normalisation over the whole batch (ppo.py:116-119): deterministic single-CTA two-pass reduction
__global__ void __launch_bounds__(1024) k_gae_normalize(size_t n, float* __restrict__ adv) {
  __shared__ double red[1024];
  const int tid = threadIdx.x;
  double s = 0.0;
  for (size_t i = tid; i < n; i += 1024) s += (double)adv[i];
  red[tid] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  const float mean = (float)(red[0] / (double)n);
  __syncthreads();
  s = 0.0;
  for (size_t i = tid; i < n; i += 1024) { const float d = adv[i] - mean; s += (double)(d * d); }
  red[tid] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  float sd = sqrtf((float)(red[0] / (double)n));
  if (sd < 1e-6f) sd = 1e-6f;
  for (size_t i = tid; i < n; i += 1024) adv[i] = (adv[i] - mean) / sd;
}

// compute_ppo_loss (ksim/task/ppo.py:149-246), the mean of the summed terms (ppo.py:484-486) and that mean's gradient, fused.
// HBM-bound: per (b, t) it reads 2A (+A with entropy) + 4 floats and writes up to 6.  A CTA walks tiles of PPO_TILE
// elements: the A-wide rows of the [B][T][A] arrays are streamed with coalesced loads (consecutive threads, consecutive
// floats) into shared memory as log-ratio terms (odd row stride: conflict-free), then thread r reduces row r in index
// order (the oracle's order).  Partial sums are fp64 per CTA; k_ppo_reduce adds them in CTA order (deterministic).
constexpr int PPO_TILE = 128;
__global__ void __launch_bounds__(PPO_TILE) k_ppo_loss(size_t n, int A, const float* __restrict__ logp_on,
                                                       const float* __restrict__ logp_off, const float* __restrict__ values_on,
                                                       const float* __restrict__ values_off, const float* __restrict__ adv,
                                                       const float* __restrict__ targets, const float* __restrict__ entropy,
                                                       float clip_param, float value_loss_coef, float entropy_coef, float kl_coef,
                                                       float log_clip_value, float kl_scale, int flags, float* __restrict__ losses,
                                                       float* __restrict__ d_logp, float* __restrict__ d_values,
                                                       double* __restrict__ partial, int vec_ok) {
  extern __shared__ __align__(16) float ppo_sh[];
  __shared__ double red[PPO_TILE];
  const int tid = threadIdx.x, AP = A | 1;
  float* sd = ppo_sh;
  float* se = ppo_sh + PPO_TILE * AP;
  const float inv_n = 1.0f / (float)n;
  double acc[1 + PPO_N_TERMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (size_t e0 = (size_t)blockIdx.x * PPO_TILE; e0 < n; e0 += (size_t)gridDim.x * PPO_TILE) {
    const int cnt = (int)(n - e0 < (size_t)PPO_TILE ? n - e0 : (size_t)PPO_TILE);
    const size_t f0 = e0 * A;
    const int nf = cnt * A;
    // a full tile starts on a 16-byte boundary (128 A floats per tile) and holds a multiple of 4 floats: 128-bit loads
    const int nf4 = vec_ok ? (nf >> 2) : 0;
#pragma unroll 4
    for (int v = tid; v < nf4; v += PPO_TILE) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(logp_off + f0) + v);
      const float4 y = __ldcs(reinterpret_cast<const float4*>(logp_on + f0) + v);
      float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (entropy) z = __ldcs(reinterpret_cast<const float4*>(entropy + f0) + v);
      int r = (4 * v) / A, c = 4 * v - r * A;
      const float dx[4] = {__fsub_rn(x.x, y.x), __fsub_rn(x.y, y.y), __fsub_rn(x.z, y.z), __fsub_rn(x.w, y.w)};
      const float ze[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        sd[r * AP + c] = dx[k];
        if (entropy) se[r * AP + c] = ze[k];
        if (++c == A) { c = 0; r++; }
      }
    }
    for (int f = 4 * nf4 + tid; f < nf; f += PPO_TILE) {
      const int r = f / A, c = f - r * A;
      sd[r * AP + c] = __fsub_rn(logp_off[f0 + f], logp_on[f0 + f]);
      if (entropy) se[r * AP + c] = entropy[f0 + f];
    }
    __syncthreads();
    if (tid < cnt) {
      const size_t e = e0 + tid;
      float lr = 0.0f, es = 0.0f;
      for (int a = 0; a < A; a++) lr = __fadd_rn(lr, sd[tid * AP + a]);
      if (entropy) for (int a = 0; a < A; a++) es = __fadd_rn(es, se[tid * AP + a]);
      const float ad = adv[e], vt = targets[e], von = values_on[e], voff = values_off[e];
      // policy term (ppo.py:205-210); clip gradients follow lax.clamp (1 strictly inside), min / max ties split evenly
      const float lrc = fminf(fmaxf(lr, -log_clip_value), log_clip_value);
      const float ratio = expf(lrc);
      const float lo = 1.0f - clip_param, hi = 1.0f + clip_param;
      const float rcl = fminf(fmaxf(ratio, lo), hi);
      const float s1 = __fmul_rn(ratio, ad), s2 = __fmul_rn(rcl, ad);
      const float pol = -fminf(s1, s2);
      const float dr_dlr = (lr > -log_clip_value && lr < log_clip_value) ? ratio : 0.0f;
      const float g1 = -ad, g2 = (ratio > lo && ratio < hi) ? -ad : 0.0f;
      const float dpol = (s1 < s2 ? g1 : (s2 < s1 ? g2 : 0.5f * (g1 + g2))) * dr_dlr;
      // value term (ppo.py:212-222, clipped_value_loss :124-135)
      float val, dval;
      const float err = __fsub_rn(voff, vt);
      if (flags & PPO_CLIPPED_VALUE_LOSS) {
        const float d = __fsub_rn(voff, von);
        const float cerr = __fsub_rn(__fadd_rn(von, fminf(fmaxf(d, -clip_param), clip_param)), vt);
        const float e2 = __fmul_rn(err, err), c2 = __fmul_rn(cerr, cerr);
        const float gc = (d > -clip_param && d < clip_param) ? cerr : 0.0f;
        val = __fmul_rn(__fmul_rn(0.5f, fmaxf(e2, c2)), value_loss_coef);
        dval = value_loss_coef * (e2 > c2 ? err : (c2 > e2 ? gc : 0.5f * (err + gc)));
      } else {
        val = __fmul_rn(__fmul_rn(0.5f, __fmul_rn(err, err)), value_loss_coef);
        dval = value_loss_coef * err;
      }
      // kl and entropy terms (ppo.py:224-238, 471-473)
      const float kl = __fmul_rn(__fmul_rn(-lr, kl_coef), kl_scale);
      const float ent = entropy ? __fmul_rn(-es, entropy_coef) : 0.0f;
      if (losses) { losses[e] = pol; losses[n + e] = val; losses[2 * n + e] = kl; losses[3 * n + e] = ent; }
      if (d_logp) d_logp[e] = (dpol - kl_coef * kl_scale) * inv_n;
      if (d_values) d_values[e] = dval * inv_n;
      acc[0] += (double)__fadd_rn(__fadd_rn(__fadd_rn(pol, val), kl), ent);
      acc[1] += (double)pol; acc[2] += (double)val; acc[3] += (double)kl; acc[4] += (double)ent;
    }
    __syncthreads();
  }
  for (int k = 0; k < 1 + PPO_N_TERMS; k++) {
    red[tid] = acc[k];
    __syncthreads();
    for (int o = PPO_TILE / 2; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    if (tid == 0) partial[(size_t)blockIdx.x * (1 + PPO_N_TERMS) + k] = red[0];
    __syncthreads();
  }
}

// one warp per term: lane l adds partials l, l + 32, ... in order, then a fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(32 * (1 + PPO_N_TERMS)) k_ppo_reduce(int n_blocks, size_t n, const double* __restrict__ partial,
                                                                       float* __restrict__ out) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < n_blocks; b += 32) s += partial[(size_t)b * (1 + PPO_N_TERMS) + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[k] = (float)(s / (double)n);
}

// Trajectory metrics, stage 1: one warp per environment, lanes over time steps (warp-shuffle reductions).
__global__ void __launch_bounds__(256) k_metrics_env(int n_envs, int n_steps, int R, int off_done, int off_time, int off_qpos,
                                                     int off_term, int n_term, int n_comp, const float* __restrict__ rec,
                                                     const float* __restrict__ total, const float* __restrict__ comps,
                                                     float* __restrict__ per_env) {
  const int lane = threadIdx.x & 31;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (env >= n_envs) return;
  const int W = B2S_METRICS_FIXED + n_comp;
  float tsum = 0.0f, dm = 0.0f, deaths = 0.0f, maxd = 0.0f, anyt = 0.0f, rsum = 0.0f;
  float ts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = lane; t < n_steps; t += 32) {
    const float* r = rec + ((size_t)t * n_envs + env) * R;
    const bool done = r[off_done] != 0.0f;
    const bool mask = done || t == n_steps - 1;   // done.at[..., -1].set(True)
    tsum += mask ? r[off_time] : 0.0f;
    dm += mask ? 1.0f : 0.0f;
    deaths += done ? 1.0f : 0.0f;
    maxd = fmaxf(maxd, sqrtf(r[off_qpos] * r[off_qpos] + r[off_qpos + 1] * r[off_qpos + 1] + r[off_qpos + 2] * r[off_qpos + 2]));
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < n_term) { const float c = r[off_term + k]; ts[k] += c; any |= c != 0.0f; }
    anyt += any ? 1.0f : 0.0f;
    rsum += total[(size_t)env * n_steps + t];
  }
  tsum = wsum(tsum); dm = wsum(dm); deaths = wsum(deaths); maxd = wmax(maxd); anyt = wsum(anyt); rsum = wsum(rsum);
#pragma unroll
  for (int k = 0; k < 8; k++) ts[k] = wsum(ts[k]);
  float* o = per_env + (size_t)env * W;
  if (lane == 0) {
    o[0] = tsum / dm; o[1] = deaths; o[2] = maxd; o[3] = anyt; o[12] = rsum;
#pragma unroll
    for (int k = 0; k < 8; k++) o[4 + k] = ts[k];
  }
  for (int c = 0; c < n_comp; c++) {
    float cs = 0.0f;
    const float* row = comps + ((size_t)c * n_envs + env) * n_steps;
    for (int t = lane; t < n_steps; t += 32) cs += row[t];
    cs = wsum(cs);
    if (lane == 0) o[B2S_METRICS_FIXED + c] = cs;
  }
}

// stage 2: one CTA, fixed order, fp64 accumulators -> out[]
__global__ void __launch_bounds__(256) k_metrics_reduce(int n_envs, int n_steps, int n_comp, const float* __restrict__ per_env,
                                                        float* __restrict__ out) {
  __shared__ double red[256];
  const int W = B2S_METRICS_FIXED + n_comp, tid = threadIdx.x;
  for (int j = 0; j < W; j++) {
    double a = 0.0;
    for (int e = tid; e < n_envs; e += 256) { const double v = per_env[(size_t)e * W + j]; a = j == 2 ? (v > a ? v : a) : a + v; }
    red[tid] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) red[tid] = j == 2 ? (red[tid + o] > red[tid] ? red[tid + o] : red[tid]) : red[tid] + red[tid + o];
      __syncthreads();
    }
    if (tid == 0) out[j] = (float)red[0];
    __syncthreads();
  }
  if (tid == 0) {
    const float n = (float)n_envs, nt = (float)n_envs * (float)n_steps;
    out[0] /= n; out[1] /= n;
    const float num = fmaxf(out[3], 1.0f);
    out[3] = num;
    for (int k = 0; k < 8; k++) out[4 + k] /= num;
    for (int j = 12; j < W; j++) out[j] /= nt;
  }
}

// TrajectoryDatasetWriter.write (ksim/dataset.py:59-103): one CTA per environment gathers its T records (and its reward
// rows) into one flattened sample; writes are contiguous, record reads contiguous along each field.
__global__ void __launch_bounds__(256) k_dataset_pack(int n_envs, int n_steps, int R, const float* __restrict__ rec,
                                                      const float* __restrict__ total, const float* __restrict__ comps,
                                                      const int* __restrict__ fields, int n_fields, int sample_size,
                                                      float* __restrict__ out) {
  const int env = blockIdx.x;
  float* o = out + (size_t)env * sample_size;
  for (int f = 0; f < n_fields; f++) {
    const int src = fields[4 * f], off = fields[4 * f + 1], dim = fields[4 * f + 2], oo = fields[4 * f + 3];
    if (src == 0) {
      for (int e = threadIdx.x; e < n_steps * dim; e += blockDim.x) {
        const int t = e / dim, k = e - t * dim;
        o[oo + e] = rec[((size_t)t * n_envs + env) * R + off + k];
      }
    } else {
      const float* row = src == 1 ? total + (size_t)env * n_steps : comps + ((size_t)off * n_envs + env) * n_steps;
      for (int t = threadIdx.x; t < n_steps; t += blockDim.x) o[oo + t] = row[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------- host side

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
extern "C" int b200sim_model_set_option(b200sim_model* mo, int option, int value) {
  if (!mo) return fail(B2S_ERR_BAD_ARG, "null model");
  if (option == B2S_OPT_SOLVER && (value == B2S_SOLVER_FAST || value == B2S_SOLVER_REFERENCE)) {
    // bits 8 / 10 / 11 of sync_mask: constraint-space Newton off, exact line search off, reference-shaped iteration in the
    // factor-reusing dof-space Newton (b2s_physics.cuh: dual_enabled / exact_ls_enabled / newton_dof_refshape).  Models whose
    // rows always fit the constraint-space path (humanoid) have no newton_dof and run solve_primal; the others run the same
    // iteration through newton_dof's bookkeeping.  B2S_SYNC_MASK bit 9 (tuning) forces solve_primal everywhere.
    const int ref_bits = (1 << 8) | (1 << 10) | (1 << 11);
    mo->mdl.sync_mask = value == B2S_SOLVER_REFERENCE ? (mo->mdl.sync_mask | ref_bits) : (mo->mdl.sync_mask & ~ref_bits);
    return B2S_OK;
  }
  return fail(B2S_ERR_BAD_ARG, "unknown option or value");
}

#define CUDA_OK(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) return fail(B2S_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char* b200sim_last_error(void) { return g_last_error.c_str(); }

template <class D>
static bool fits(const int* h) {
  if (D::EXACT)
    return h[MH_NQ] == D::NQ && h[MH_NV] == D::NV && h[MH_NBODY] == D::NB && h[MH_NJNT] == D::NJ && h[MH_NU] == D::NU &&
           h[MH_NSITE] == D::NS && h[MH_NPAIR] == D::NPAIR && h[MH_NCON] == D::NCON && h[MH_NLIMIT] == D::NLIMIT &&
           h[MH_NSENSORDATA] <= D::NSD && h[MH_NEFC_DENSE] <= D::NDENSE && h[MH_NFRICTION] <= D::NFRIC;
  return h[MH_NQ] <= D::NQ && h[MH_NV] <= D::NV && h[MH_NBODY] <= D::NB && h[MH_NJNT] <= D::NJ && h[MH_NU] <= D::NU &&
         h[MH_NSITE] <= D::NS && h[MH_NSENSORDATA] <= D::NSD && h[MH_NPAIR] <= D::NPAIR && h[MH_NCON] <= D::NCON &&
         h[MH_NEFC_DENSE] <= D::NDENSE && h[MH_NLIMIT] <= D::NLIMIT && h[MH_NFRICTION] <= D::NFRIC;
}
// per-env geom / site overrides live only in the workspaces of the non-SMALL instantiations (b2s_work.cuh: PairGeom): a task
// that randomises them must not be served by the humanoid instantiation, which would silently ignore the override
static bool needs_env_geometry(const int* ti) {
  for (int r = 0; r < ti[TI_N_RAND]; r++) {
    const int code = TAB(ti, TI_RAND_TAB, r, TAB_AUX0);
    if (code == RAND_GEOM_SIZE || code == RAND_GEOM_POS || code == RAND_SITE_QUAT || code == RAND_SITE_POS) return true;
  }
  return false;
}

template <class D>
static bool task_fits(const int* ti) {
  return ti[TI_EVENT_SIZE] <= D::EVS && ti[TI_ACT_STATE_SIZE] <= D::ACS && ti[TI_CMD_SIZE] <= D::CMDS &&
         ti[TI_OBS_CARRY_SIZE] <= D::OCS && ti[TI_N_OBS] <= D::NOBS && ti[TI_ACTION_SIZE] <= D::NA &&
         !(D::SMALL && needs_env_geometry(ti));
}

template <class D>
static int configure(b200sim_model* mo, const size_t* words) {
  mo->smem_per_env = sizeof(Work<D>);
  const size_t fixed = sizeof(Mdl) + 16;
  constexpr int mw = max_warps<D>();
  static_assert(mw >= 1, "workspace does not fit in shared memory");
  int wpb = mw;  // one CTA per SM: all its warps run in lockstep (shared i-cache fills)
  if (const char* e = getenv("B2S_WPB")) { const int v = atoi(e); if (v >= 1 && v <= mw) wpb = v; }  // tuning experiments
  mo->warps_per_cta = wpb;
  // model arrays staged in shared memory, in priority order mi, mf, ti, tf, while they fit
  size_t left = (kSmemBudget - fixed - wpb * sizeof(Work<D>)) / 4, staged = 0;
  const bool no_stage = getenv("B2S_NO_STAGE") != nullptr;  // tuning experiments
  for (int a = 0; a < 4; a++) {
    const bool fit = !no_stage && words[a] <= left;
    mo->mdl.stage[a] = fit ? (int)words[a] : 0;
    if (fit) { left -= words[a]; staged += words[a]; }
  }
  mo->smem_bytes = wpb * sizeof(Work<D>) + fixed + staged * 4;
  CUDA_OK(cudaFuncSetAttribute(k_init_envs<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mo->smem_bytes));
  CUDA_OK(cudaFuncSetAttribute(k_rollout<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mo->smem_bytes));
  return B2S_OK;
}

extern "C" int b200sim_model_create(const void* blob, size_t nbytes, b200sim_model** out) {
  if (!blob || !out) return fail(B2S_ERR_BAD_ARG, "null argument");
  if (nbytes < MH_SIZE * sizeof(int)) return fail(B2S_ERR_BAD_BLOB, "blob smaller than its header");
  const int* h = (const int*)blob;
  if (h[MH_MAGIC] != B2S_MODEL_MAGIC || h[MH_VERSION] != B2S_MODEL_VERSION) return fail(B2S_ERR_BAD_BLOB, "bad magic/version");
  const size_t ni = h[MH_NI], nf = h[MH_NF], nti = h[MH_NTI], ntf = h[MH_NTF];
  if (nbytes != (MH_SIZE + ni + nf + nti + ntf) * 4) return fail(B2S_ERR_BAD_BLOB, "blob size does not match its header");
  const int* ti = h + MH_SIZE + ni + nf;
  if (nti < TI_HEADER_SIZE || ti[TI_MAGIC] != B2S_TASK_MAGIC) return fail(B2S_ERR_BAD_BLOB, "bad task program");
  if (h[MH_NV] > B2S_MAXNV || h[MH_NBODY] > B2S_MAXNB) return fail(B2S_ERR_TOO_LARGE, "nv / nbody exceed one-per-lane limit");
  b200sim_model* mo = new b200sim_model();
  memcpy(mo->mdl.h, h, MH_SIZE * sizeof(int));
  mo->ti_host.assign(ti, ti + nti);
  cudaError_t e = cudaMalloc(&mo->dev_blob, nbytes);
  if (e != cudaSuccess) { delete mo; return fail(B2S_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
  e = cudaMemcpy(mo->dev_blob, blob, nbytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(mo->dev_blob); delete mo; return fail(B2S_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
  const int* d = (const int*)mo->dev_blob;
  mo->mdl.mi = d + MH_SIZE;
  mo->mdl.mf = (const float*)(d + MH_SIZE + ni);
  mo->mdl.ti = d + MH_SIZE + ni + nf;
  mo->mdl.tf = (const float*)(d + MH_SIZE + ni + nf + nti);
  // Measured on B200 (profiles/r02k_barrier_masks.txt): the CTA barriers that pay are the ones after the two stages whose length
  // varies per environment -- acceleration (bit 4: the Z solves depend on the row count) and the Newton solve (bit 5) -- plus the
  // control-step one (bit 6, neutral); the barriers after the fixed-length stages (bits 0-3) cost 2.5 % on the humanoid kernel
  // (33.9 -> 33.0 ms) now that the stages between them are short, and nothing either way on the K-Bot kernel.  Without bit 4 or 5
  // the warps stay out of step through the following stages and every warp pays its own instruction-cache misses (+3 % / +33 %).
  mo->mdl.sync_mask = 0x70;
#ifdef B2S_POISON_SMEM
  mo->mdl.poison_lo = getenv("B2S_POISON_LO") ? atoi(getenv("B2S_POISON_LO")) : 0;
  mo->mdl.poison_hi = getenv("B2S_POISON_HI") ? atoi(getenv("B2S_POISON_HI")) : 0;
#endif
#ifdef B2S_DEBUG_STOP
  if (const char* e = getenv("B2S_SYNC_MASK")) mo->mdl.sync_mask = (int)strtol(e, nullptr, 0) & 0xff0fff;  // + DBG_STOP point
#else
  if (const char* e = getenv("B2S_SYNC_MASK")) mo->mdl.sync_mask = (int)strtol(e, nullptr, 0) & 0xfff;  // tuning experiments
#endif
  int rc;
  const size_t words[4] = {ni, nf, nti, ntf};
  int vmin = 0;  // B2S_MIN_VARIANT: tests run a small model through a larger instantiation than the first that fits
  if (const char* e = getenv("B2S_MIN_VARIANT")) vmin = atoi(e);
  if (vmin <= 0 && fits<DimsHumanoid>(h) && task_fits<DimsHumanoid>(ti)) { mo->variant = 0; rc = configure<DimsHumanoid>(mo, words); }
  else if (vmin <= 1 && fits<DimsKbot>(h) && task_fits<DimsKbot>(ti)) { mo->variant = 1; rc = configure<DimsKbot>(mo, words); }
  else if (fits<DimsGeneric>(h) && task_fits<DimsGeneric>(ti)) { mo->variant = 2; rc = configure<DimsGeneric>(mo, words); }
  else rc = fail(B2S_ERR_TOO_LARGE, "model / task sizes exceed the largest kernel instantiation");
  if (rc != B2S_OK) { cudaFree(mo->dev_blob); delete mo; return rc; }
  *out = mo;
  return B2S_OK;
}

extern "C" void b200sim_model_destroy(b200sim_model* model) {
  if (!model) return;
  cudaFree(model->dev_blob);
  delete model;
}

extern "C" int b200sim_model_info(const b200sim_model* mo, int what) {
  if (!mo) return -1;
  const int* ti = mo->ti_host.data();
  switch (what) {
    case B2S_INFO_STATE_SIZE: return ti[TI_STATE_SIZE];
    case B2S_INFO_REC_SIZE: return ti[TI_REC_SIZE];
    case B2S_INFO_KEY_SIZE: return ti[TI_KEY_SIZE];
    case B2S_INFO_RAND_SIZE: return ti[TI_RAND_SIZE];
    case B2S_INFO_ACTION_SIZE: return ti[TI_ACTION_SIZE];
    case B2S_INFO_N_REW_COMP: return ti[TI_N_REW_COMP];
    case B2S_INFO_REW_CARRY_SIZE: return ti[TI_REW_CARRY_SIZE];
    case B2S_INFO_SMEM_PER_ENV: return (int)mo->smem_per_env;
    case B2S_INFO_WARPS_PER_CTA: return mo->warps_per_cta;
    case B2S_INFO_VARIANT: return mo->variant;
    case B2S_INFO_STAGED_ARRAYS: return (mo->mdl.stage[0] > 0) + (mo->mdl.stage[1] > 0) + (mo->mdl.stage[2] > 0) + (mo->mdl.stage[3] > 0);
  }
  return -1;
}

extern "C" int b200sim_init_envs(const b200sim_model* mo, void* stream, int n_envs, const uint32_t* init_keys,
                                 const float* levels, const float* rand, float* state, uint32_t* keys, float* rec0,
                                 uint32_t* act_keys) {
  if (!mo || n_envs < 0 || !init_keys || !levels || !state || !keys || !rec0) return fail(B2S_ERR_BAD_ARG, "null argument");
  if (n_envs == 0) return B2S_OK;
  if (!rand && mo->ti_host[TI_RAND_SIZE] > 0) return fail(B2S_ERR_BAD_ARG, "rand is required for this task");
  const int wpb = mo->warps_per_cta;
  const dim3 grid((n_envs + wpb - 1) / wpb), block(wpb * 32);
  const size_t smem = mo->smem_bytes;
  DISPATCH(mo, (k_init_envs<D><<<grid, block, smem, (cudaStream_t)stream>>>(mo->mdl, n_envs, init_keys, levels, rand, state, keys, rec0, act_keys)));
  CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

static int launch_rollout(const b200sim_model* mo, void* stream, int n_envs, int n_steps, const float* actions,
                          const float* rand, float* state, uint32_t* keys, float* rec, float* rec_next_single,
                          uint32_t* act_keys) {
  const int wpb = mo->warps_per_cta;
  const dim3 grid((n_envs + wpb - 1) / wpb), block(wpb * 32);
  const size_t smem = mo->smem_bytes;
  DISPATCH(mo, (k_rollout<D><<<grid, block, smem, (cudaStream_t)stream>>>(mo->mdl, n_envs, n_steps, actions, rand, state, keys, rec, rec_next_single, act_keys)));
  CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

extern "C" int b200sim_step_ctrl(const b200sim_model* mo, void* stream, int n_envs, const float* action, const float* rand,
                                 float* state, uint32_t* keys, float* rec_cur, float* rec_next, uint32_t* act_keys) {
  if (!mo || n_envs < 0 || !action || !state || !keys || !rec_cur || !rec_next) return fail(B2S_ERR_BAD_ARG, "null argument");
  if (n_envs == 0) return B2S_OK;
  if (!rand && mo->ti_host[TI_RAND_SIZE] > 0) return fail(B2S_ERR_BAD_ARG, "rand is required for this task");
  return launch_rollout(mo, stream, n_envs, 1, action, rand, state, keys, rec_cur, rec_next, act_keys);
}

extern "C" int b200sim_rollout(const b200sim_model* mo, void* stream, int n_envs, int n_steps, const float* actions,
                               const float* rand, float* state, uint32_t* keys, float* rec) {
  if (!mo || n_envs < 0 || n_steps < 0) return fail(B2S_ERR_BAD_ARG, "bad argument");
  if (n_envs == 0 || n_steps == 0) return B2S_OK;
  if (!actions || !state || !keys || !rec) return fail(B2S_ERR_BAD_ARG, "null argument");
  if (!rand && mo->ti_host[TI_RAND_SIZE] > 0) return fail(B2S_ERR_BAD_ARG, "rand is required for this task");
  return launch_rollout(mo, stream, n_envs, n_steps, actions, rand, state, keys, rec, nullptr, nullptr);
}

// dynamic shared memory of the scoring kernels grows with T: opt in above the 48 KB default, refuse beyond the SM
template <class K>
static int scoring_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > kSmemBudget) return fail(B2S_ERR_BAD_ARG, std::string(what) + ": n_steps too large for the shared-memory scan");
  if (bytes > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B2S_OK;
}

extern "C" int b200sim_score(const b200sim_model* mo, void* stream, int n_envs, int n_steps, const float* rec, const float* levels,
                             float* carry, const float* values, float gamma, float lam, int flags, float* total, float* comps,
                             uint8_t* dones, uint8_t* successes, float* adv, float* targets, float* returns) {
  if (!mo || n_envs < 0 || n_steps < 0 || !rec || !levels || !total || !comps) return fail(B2S_ERR_BAD_ARG, "null argument");
  if (values && (!adv || !targets || !returns)) return fail(B2S_ERR_BAD_ARG, "values given without adv / targets / returns");
  if (flags & GAE_NORMALIZE_ADVANTAGES && !values) return fail(B2S_ERR_BAD_ARG, "normalisation asked for without values");
  if (n_envs == 0 || n_steps == 0) return B2S_OK;
  if (!carry && mo->ti_host[TI_REW_CARRY_SIZE] > 0) return fail(B2S_ERR_BAD_ARG, "carry is required for this task");
  // shared memory: [16 B barrier][staged rows][scan arrays][column table]; staging needs 16-byte aligned rows and has to fit the SM
  const int* ti = mo->ti_host.data();
  const size_t scan = (values ? 6 * (size_t)n_steps : 0) * sizeof(float) + (((size_t)ti[TI_REW_CARRY_SIZE] + 3) & ~(size_t)3) * sizeof(float);
  const size_t rows = (size_t)n_steps * ti[TI_SCORE_WIDTH] * sizeof(float), table = ((size_t)ti[TI_REC_SIZE] * 2 + 15) & ~(size_t)15;
  int staged = ti[TI_SCORE_NSEG] > 0 && ti[TI_REC_SIZE] % 4 == 0 && (reinterpret_cast<uintptr_t>(rec) & 15) == 0 &&
               16 + rows + scan + table <= kSmemBudget;
  // Measured (gpurun_out/r02h_score_sweep.txt, humanoid 4096 x 64): staging costs more than it saves at this shape -- the rows'
  // sectors are re-read through L1 (98 % hit rate) and 63 KB of shared memory per CTA caps the SM at 3 environments, so the
  // copy latency is not hidden: 0.48 ms staged vs 0.22 ms unstaged at 2 warps per CTA.  Opt-in (B2S_SCORE_STAGE=1).
  if (!getenv("B2S_SCORE_STAGE")) staged = 0;
  const size_t smem = 16 + (staged ? rows + table : 0) + scan;
  if (int rc = scoring_smem(k_score, smem, "b200sim_score")) return rc;
  int warps = 2;  // measured best of 1 / 2 / 4 / 8: the stateful scans are the critical path, more CTAs per SM hide them better
  if (const char* e = getenv("B2S_SCORE_WARPS")) { const int v = atoi(e); if (v >= 1 && v <= SCORE_WARPS_MAX) warps = v; }  // tuning experiments
  k_score<<<n_envs, warps * 32, smem, (cudaStream_t)stream>>>(mo->mdl, n_envs, n_steps, staged, rec, levels, carry, values, gamma, lam,
                                                                     flags, total, comps, dones, successes, adv, targets, returns);
  CUDA_OK(cudaGetLastError());
  if (flags & GAE_NORMALIZE_ADVANTAGES) {
    k_gae_normalize<<<1, 1024, 0, (cudaStream_t)stream>>>((size_t)n_envs * n_steps, adv);
    CUDA_OK(cudaGetLastError());
  }
  return B2S_OK;
}

extern "C" int b200sim_rewards(const b200sim_model* mo, void* stream, int n_envs, int n_steps, const float* rec,
                               const float* levels, float* carry, float* total, float* comps) {
  return b200sim_score(mo, stream, n_envs, n_steps, rec, levels, carry, nullptr, 0.0f, 0.0f, 0, total, comps, nullptr, nullptr, nullptr,
                       nullptr, nullptr);
}

extern "C" int b200sim_extract_flags(const b200sim_model* mo, void* stream, int n_envs, int n_steps, const float* rec,
                                     uint8_t* dones, uint8_t* successes) {
  if (!mo || !rec || !dones || !successes || n_envs < 0 || n_steps < 0) return fail(B2S_ERR_BAD_ARG, "null argument");
  const size_t n = (size_t)n_envs * n_steps;
  if (n == 0) return B2S_OK;
  const int* ti = mo->ti_host.data();
  k_extract_flags<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_envs, n_steps, ti[TI_REC_SIZE], ti[TI_R_DONE],
                                                                                 ti[TI_R_SUCCESS], rec, dones, successes);
  CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

extern "C" int b200sim_gae(void* stream, int batch, int n_steps, const float* values, const float* rewards,
                           const uint8_t* dones, const uint8_t* successes, float gamma, float lam, int flags, float* adv,
                           float* targets, float* returns) {
  if (batch < 0 || n_steps < 0 || !values || !rewards || !dones || !successes || !adv || !targets || !returns)
    return fail(B2S_ERR_BAD_ARG, "null argument");
  if (batch == 0 || n_steps == 0) return B2S_OK;
  const size_t smem = (size_t)GAE_WARPS * 5 * n_steps * sizeof(float);
  if (int rc = scoring_smem(k_gae, smem, "b200sim_gae")) return rc;
  k_gae<<<(batch + GAE_WARPS - 1) / GAE_WARPS, GAE_WARPS * 32, smem, (cudaStream_t)stream>>>(batch, n_steps, values, rewards, dones, successes,
                                                                                            gamma, lam, flags, adv, targets, returns);
  CUDA_OK(cudaGetLastError());
  if (flags & GAE_NORMALIZE_ADVANTAGES) {
    k_gae_normalize<<<1, 1024, 0, (cudaStream_t)stream>>>((size_t)batch * n_steps, adv);
    CUDA_OK(cudaGetLastError());
  }
  return B2S_OK;
}

extern "C" int b200sim_ppo_loss(void* stream, int batch, int n_steps, int n_act, const float* logp_on, const float* logp_off,
                                const float* values_on, const float* values_off, const float* adv, const float* targets,
                                const float* entropy, float clip_param, float value_loss_coef, float entropy_coef, float kl_coef,
                                float log_clip_value, float kl_scale, int flags, float* losses, float* out, float* d_logp,
                                float* d_values, void* scratch) {
  if (batch < 0 || n_steps < 0) return fail(B2S_ERR_BAD_ARG, "bad argument");
  if (n_act < 1 || n_act > PPO_MAX_ACTIONS) return fail(B2S_ERR_UNSUPPORTED, "n_act outside [1, PPO_MAX_ACTIONS]");
  if (batch == 0 || n_steps == 0) return B2S_OK;
  if (!logp_on || !logp_off || !values_on || !values_off || !adv || !targets || !out || !scratch)
    return fail(B2S_ERR_BAD_ARG, "null argument");
  const size_t n = (size_t)batch * n_steps;
  const size_t tiles = (n + PPO_TILE - 1) / PPO_TILE;
  static_assert(PPO_MAX_CTAS * (1 + PPO_N_TERMS) * sizeof(double) <= PPO_SCRATCH_BYTES, "scratch too small");
  const int vec_ok = ((((uintptr_t)logp_on | (uintptr_t)logp_off | (uintptr_t)entropy) & 15) == 0) ? 1 : 0;
  const size_t smem = (size_t)(entropy ? 2 : 1) * PPO_TILE * (n_act | 1) * sizeof(float);
  if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(k_ppo_loss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one wave of resident CTAs, each striding over the tiles (no tail wave): the grid, and with it the order of the
  // reduction, is a function of (n, n_act, entropy, device model) only -- reproducible run to run
  int dev = 0, sms = 0, per_sm = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ppo_loss, PPO_TILE, smem));
  size_t resident = (size_t)sms * (per_sm > 0 ? per_sm : 1);
  if (resident > PPO_MAX_CTAS) resident = PPO_MAX_CTAS;
  const int n_blocks = (int)(tiles < resident ? tiles : resident);
  k_ppo_loss<<<n_blocks, PPO_TILE, smem, (cudaStream_t)stream>>>(n, n_act, logp_on, logp_off, values_on, values_off, adv, targets,
                                                                 entropy, clip_param, value_loss_coef, entropy_coef, kl_coef,
                                                                 log_clip_value, kl_scale, flags, losses, d_logp, d_values,
                                                                 (double*)scratch, vec_ok);
  CUDA_OK(cudaGetLastError());
  k_ppo_reduce<<<1, 32 * (1 + PPO_N_TERMS), 0, (cudaStream_t)stream>>>(n_blocks, n, (const double*)scratch, out);
  CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

extern "C" int b200sim_dataset_pack(void* stream, int n_envs, int n_steps, int rec_size, const float* rec, const float* total,
                                    const float* comps, const int32_t* fields, int n_fields, int sample_size, float* out) {
  if (n_envs < 0 || n_steps < 0 || n_fields < 0 || rec_size <= 0 || sample_size < 0) return fail(B2S_ERR_BAD_ARG, "bad argument");
  if (n_envs == 0 || n_steps == 0 || n_fields == 0) return B2S_OK;
  if (!rec || !fields || !out) return fail(B2S_ERR_BAD_ARG, "null argument");
  k_dataset_pack<<<n_envs, 256, 0, (cudaStream_t)stream>>>(n_envs, n_steps, rec_size, rec, total, comps, fields, n_fields,
                                                          sample_size, out);
  CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

extern "C" int b200sim_metrics(const b200sim_model* mo, void* stream, int n_envs, int n_steps, const float* rec, const float* total,
                               const float* comps, float* per_env, float* out) {
  if (!mo || n_envs < 0 || n_steps < 0) return fail(B2S_ERR_BAD_ARG, "bad argument");
  if (n_envs == 0 || n_steps == 0) return B2S_OK;
  if (!rec || !total || !per_env || !out) return fail(B2S_ERR_BAD_ARG, "null argument");
  const int* ti = mo->ti_host.data();
  const int n_term = ti[TI_N_TERM], n_comp = ti[TI_N_REW_COMP];
  if (n_term > 8) return fail(B2S_ERR_UNSUPPORTED, "more than 8 termination components");
  if (n_comp > 0 && !comps) return fail(B2S_ERR_BAD_ARG, "comps is required for this task");
  k_metrics_env<<<(unsigned)(((size_t)n_envs * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n_envs, n_steps, ti[TI_REC_SIZE], ti[TI_R_DONE], ti[TI_R_TIME], ti[TI_R_QPOS], ti[TI_R_TERM], n_term, n_comp, rec, total, comps, per_env);
  CUDA_OK(cudaGetLastError());
  k_metrics_reduce<<<1, 256, 0, (cudaStream_t)stream>>>(n_envs, n_steps, n_comp, per_env, out);
  CUDA_OK(cudaGetLastError());
  return B2S_OK;
}
