// Recurrent policy replay on the 5th-generation tensor cores (ksim/task/ppo.py:417-488 calls the task's
// get_ppo_variables, examples/walking.py:674-728, which scans two stacks of GRU(512) cells over the T steps of a minibatch).
//
// What is sequential in that scan is only   gh_t = h_{t-1} W_hh^T   and the gate arithmetic behind it; everything else
// (input projections, x_t W_ih^T for all t at once, the output MLP) is a plain batched GEMM.  These two kernels are the
// sequential part, forward and backward, as persistent cooperative kernels:
//
//   * a GROUP of 16 CTAs owns 128 batch rows of one GRU layer; CTA c of the group owns hidden units [32 c, 32 c + 32) of all
//     three gates and keeps its slice of W_hh (forward: 96 x 512, backward: 32 x 1536 of W_hh^T; 96 KB of bf16) in shared
//     memory for the whole sequence;
//   * every step the CTA gathers the group's 16 state chunks (bf16, already laid out as tcgen05 K-major core matrices) with
//     1-D bulk copies (cp.async.bulk + mbarrier complete_tx), issues the step's tcgen05.mma chain (M = 128, fp32 accumulator
//     in TMEM) from one thread, and its 8 epilogue warps read the accumulator back with tcgen05.ld, apply the gates with the
//     fp32 state held in registers, write the step's outputs and publish the CTA's own chunk of the next state to the peers
//     through L2 (release / acquire flags);
//   * groups never talk to each other, so 2 networks x 4 row tiles = 8 groups = 128 CTAs run side by side.
//
// GRU cell (equinox.nn.GRUCell as used by examples/walking.py:78-91, 124-127):
//   r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r (gh_n + b_hn)), h' = n + z (h - n),
//   gi = x W_ih^T + b_i.  The carry handed to the next step restarts from zero where the episode ended
//   (examples/walking.py:715-719): h_next = keep_t h'_t.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "../../include/b200sim.h"
#include "../../include/b200sim_codes.h"
#include "b2s_tc.cuh"

using namespace b2s_tc;
typedef __nv_bfloat16 bf16;

namespace {

constexpr int H = B2S_GRU_HIDDEN;    // hidden size (512)
constexpr int U = 32;                // hidden units per CTA
constexpr int NCH = H / U;           // CTAs (= state chunks) per group
constexpr int MT = 128;              // batch rows per group = UMMA M
constexpr int NEPI = 256;            // epilogue threads: warps 0-7; warp w reads TMEM lanes 32 (w % 4).., units 16 (w / 4)..
constexpr int NTHREADS = NEPI + 64;  // backward: + warp 8 (bulk-copy producer) + warp 9 (MMA issuer, TMEM owner)
constexpr int FWD_THREADS = NEPI + 32;  // forward: + warp 8 (MMA issuer, TMEM owner); chunks arrive by multicast, no producer
constexpr uint32_t SPIN = 1u << 22;  // bounded waits: a lost signal ends the kernel with an error code, not a hung GPU

constexpr int W_BYTES = 3 * U * H * 2;         // 98304: the CTA's weight slice
constexpr int FWD_CH = MT * U * 2;             // 8192: one forward state chunk (128 rows x 32 units, bf16)
constexpr int FWD_A = NCH * FWD_CH;            // 131072: the whole 128 x 512 state tile
constexpr int BWD_CH = MT * 3 * U * 2;         // 24576: one backward chunk (128 rows x 3 gates x 32 units)
constexpr int BWD_STAGES = 4;
constexpr int FWD_SMEM = W_BYTES + FWD_A + 1024;
constexpr int BWD_SMEM = W_BYTES + BWD_STAGES * BWD_CH + 1024;
constexpr int SAVED_Q = 5;                     // r, z, n, hn = gh_n + b_hn, hmn = h - n
constexpr int SAVED_BLOCK = SAVED_Q * MT * U;  // floats per (step, group, chunk)

struct FwdProblem {
  const bf16* w;       // W_hh [3H][H]
  const float* gi;     // feature-major [3H][T B] = W_ih x^T (bias not included)
  const float* b_i;    // [3H]
  const float* b_hn;   // [H]
  const float* h0;     // [B][H] or null (zeros)
  bf16* y;             // feature-major [H][T B]: h'_t (what the next layer / the output MLP reads)
  bf16* hst;           // feature-major [H][T B]: the state step t started from (keep_{t-1} h'_{t-1}; h0 for t = 0)
  float* h_last;       // [B][H] or null: keep_{T-1} h'_{T-1}
  float* saved;        // [T][n_mt][NCH][SAVED_BLOCK] or null (no backward wanted)
  bf16* xchg;          // [(T+1)][n_mt][NCH][FWD_CH / 2] state chunks in core-matrix order
  int* flags;          // [n_mt][NCH] slots published
};
struct FwdParams {
  FwdProblem pr[B2S_GRU_MAX_PROBLEMS];
  int n_problems, T, B, n_mt;
  const float* keep;   // [T][B] or null (all ones)
  int* status;
  uint32_t lbo_sbo_swap;  // debug: exchange the two descriptor strides
};

struct BwdProblem {
  const bf16* wt;      // W_hh^T [H][3H]
  const float* dy;     // feature-major [H][T B]: gradient with respect to y
  const float* saved;
  bf16* dgi;           // feature-major [3H][T B]: gradient w.r.t. gi (its r and z blocks are also the gradient w.r.t. gh_r, gh_z)
  bf16* dghn;          // feature-major [H][T B]: gradient with respect to gh_n
  float* dh0;          // [B][H] or null
  bf16* xchg;          // [T][n_mt][NCH][BWD_CH / 2]
  int* flags;
};
struct BwdParams {
  BwdProblem pr[B2S_GRU_MAX_PROBLEMS];
  int n_problems, T, B, n_mt;
  const float* keep;
  int* status;
  uint32_t lbo_sbo_swap;
};

#ifdef B2S_GRU_TRACE
// debug build only (B2S_EXTRA_NVCC=-DB2S_GRU_TRACE): SM-clock stamps of CTA 0's roles, 8 points per step (tools/gru_trace.py)
__device__ long long g_trace[8 * 256];
__device__ long long g_trace_ch[16 * 256];
__device__ long long g_trace_ch0[16 * 256];
#define TRACE_CH0(t, ch) do { if (blockIdx.x == 0 && (t) < 256) g_trace_ch0[(t) * 16 + (ch)] = clock64(); } while (0)
#define TRACE_CH(t, ch) do { if (blockIdx.x == 0 && (t) < 256) g_trace_ch[(t) * 16 + (ch)] = clock64(); } while (0)
#define TRACE(t, point) do { if (blockIdx.x == 0 && (t) < 256) g_trace[(t) * 8 + (point)] = clock64(); } while (0)
#else
#define TRACE(t, point) do { } while (0)
#define TRACE_CH(t, ch) do { } while (0)
#define TRACE_CH0(t, ch) do { } while (0)
#endif

__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* status, int code) {
  if (*s_abort) return false;
  if (mbar_wait(bar, parity, SPIN)) return true;
  *s_abort = 1;
  atomicCAS(status, 0, code);
  return false;
}
__device__ __forceinline__ bool wait_flag(const int* flag, int target, volatile int* s_abort, int* status, int code) {
  if (*s_abort) return false;
  if (flag_wait_ge(flag, target, SPIN)) return true;
  *s_abort = 1;
  atomicCAS(status, 0, code);
  return false;
}

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swap) {
  return (swap & 1) ? smem_desc(addr, sbo, lbo) : smem_desc(addr, lbo, sbo);
}

// ------------------------------------------------------------------------------------------------------ forward
// One thread-block CLUSTER of 16 CTAs runs a PAIR of groups (two 128-row tiles of the same GRU layer, which share the weight
// slices) interleaved step by step: while the tensor cores of the cluster multiply one tile's state, the epilogue warps apply
// the other tile's gates and put its next state on its way, so the exchange latency of one tile hides behind the other's MMAs
// (only 7 sixteen-CTA clusters are co-resident on a B200: a pair per cluster also lets 4 clusters cover the 8 tiles of a
// 512-row minibatch of two networks in one round).  State exchange, with no flags and no polling of global memory:
//   * the epilogue threads of CTA c store the CTA's bf16 chunk of a tile's next state to the exchange buffer (L2), and one
//     thread then issues ONE multicast bulk copy of that chunk (cp.async.bulk ... .multicast::cluster) which lands at
//     sA + c * FWD_CH in all 16 CTAs and completes on `full` of each of them: one L2 read feeds 16 SMs, the arrival is the signal;
//   * the state tile sA is shared by the two tiles in turn (micro-step m = t * NG + g): before overwriting it the sender waits
//     for `peers_done`, on which the MMA issuer of every CTA of the cluster arrives (tcgen05.commit ... .multicast::cluster)
//     when its MMAs of the previous micro-step have completed;
//   * two accumulators in TMEM (columns 0.. and 128..), one per tile, each with its own ready / drained barrier.
struct FwdEpi {   // per-tile epilogue state of one thread: (row, 16 hidden units)
  float h[16];
  int grow, mt;
  bool live;
};

__global__ void __launch_bounds__(FWD_THREADS, 1) k_gru_fwd(const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + W_BYTES + FWD_A);  // all 16 chunks of a micro-step
  uint64_t* acc_full = full + 1;      // [2]
  uint64_t* tmem_empty = full + 3;    // [2]
  uint64_t* peers_done = full + 5;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(full + 6);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  float* s_bias = reinterpret_cast<float*>(smem + W_BYTES + FWD_A + 256);  // b_i slice [3][U], b_hn slice [U]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = (int)cluster_ctarank(), slot = (int)cluster_id_x(), n_slots = (int)cluster_count_x();
  const int T = p.T, B = p.B;
  const int pairs_per_problem = (p.n_mt + 1) / 2, n_pairs = p.n_problems * pairs_per_problem;
  const size_t TB = (size_t)T * B;
  constexpr uint16_t ALL = (uint16_t)((1u << NCH) - 1);
  constexpr uint32_t ACC_STRIDE = 128;  // TMEM columns between the two tiles' accumulators

  if (tid == 0) {
    mbar_init(full, 1);
    for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&tmem_empty[i], NEPI); }
    mbar_init(peers_done, NCH);
    *s_abort = 0;
    mbar_fence_init();
    mbar_arrive_expect_tx(full, NCH * FWD_CH);  // phase 0 armed: every micro-step takes the 16 chunks of one state tile
  }
  if (warp == 8) tmem_alloc(s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peers' barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  constexpr uint32_t IDESC = idesc_bf16(MT, 3 * U);

  int loaded = -1;
  // phase bookkeeping, identical in every role and every CTA of the cluster: micro-steps run so far (full, peers_done) and steps
  // run so far on each accumulator (acc_full[g], tmem_empty[g])
  uint32_t mu_base = 0, st_base[2] = {0, 0};
  for (int pr = slot; pr < n_pairs; pr += n_slots) {
    const int pi = pr / pairs_per_problem, mt0 = 2 * (pr % pairs_per_problem);
    const int NG = (mt0 + 1 < p.n_mt) ? 2 : 1;
    const FwdProblem& q = p.pr[pi];
    if (pi != loaded) {
      __syncthreads();
      // weight slice -> K-major core matrices: B row n = gate * U + j  <->  W_hh row gate * H + c U + j
      for (int idx = tid; idx < 3 * U * (H / 8); idx += FWD_THREADS) {
        const int n = idx / (H / 8), kc = idx % (H / 8);
        const uint4 v = *reinterpret_cast<const uint4*>(q.w + ((size_t)((n / U) * H + c * U + (n % U)) * H + kc * 8));
        *reinterpret_cast<uint4*>(sB + kc * (3 * U * 16) + n * 16) = v;
      }
      if (tid < 3 * U) s_bias[tid] = q.b_i[(tid / U) * H + c * U + (tid % U)];
      else if (tid < 4 * U) s_bias[tid] = q.b_hn[c * U + tid - 3 * U];
      fence_proxy_async_smem();
      __syncthreads();
      loaded = pi;
    }
    if (warp < 8) {
      // ------------------------------------------------------------------ epilogue: gates, state, outputs
      const int row = (warp & 3) * 32 + lane, half = warp >> 2, j0 = half * 16;
      const size_t ucol = (size_t)c * U + j0;
      FwdEpi e[2];
#pragma unroll
      for (int g = 0; g < 2; g++) {
        e[g].mt = mt0 + g;
        e[g].grow = e[g].mt * MT + row;
        e[g].live = g < NG && e[g].grow < B;
#pragma unroll
        for (int i = 0; i < 16; i++) e[g].h[i] = 0.0f;
        if (e[g].live && q.h0) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 v = *reinterpret_cast<const float4*>(q.h0 + (size_t)e[g].grow * H + ucol + 4 * i);
            e[g].h[4 * i] = v.x; e[g].h[4 * i + 1] = v.y; e[g].h[4 * i + 2] = v.z; e[g].h[4 * i + 3] = v.w;
          }
        }
      }
      // state s of tile g (what step s starts from) goes out first; the step's bulky outputs are stored afterwards
      uint4 v0, v1;
      auto publish = [&](FwdEpi& st, int g, int s) {
        v0 = make_uint4(pack_bf16(st.h[0], st.h[1]), pack_bf16(st.h[2], st.h[3]), pack_bf16(st.h[4], st.h[5]), pack_bf16(st.h[6], st.h[7]));
        v1 = make_uint4(pack_bf16(st.h[8], st.h[9]), pack_bf16(st.h[10], st.h[11]), pack_bf16(st.h[12], st.h[13]), pack_bf16(st.h[14], st.h[15]));
        bf16* dst = q.xchg + (((size_t)s * p.n_mt + st.mt) * NCH + c) * (FWD_CH / 2);
        *reinterpret_cast<uint4*>(dst + ((2 * half) * MT + row) * 8) = v0;
        *reinterpret_cast<uint4*>(dst + ((2 * half + 1) * MT + row) * 8) = v1;
        __syncwarp();
        named_bar_sync(1, NEPI);
        if (tid == 0 && !(p.lbo_sbo_swap & 2)) {
          fence_proxy_async();  // the chunk was written through the generic proxy; the bulk copy reads it through the async proxy
          // the micro-step that consumes this state is mc: every CTA of the cluster must have finished the MMAs of micro-step
          // mc - 1 (they read the state tile this copy overwrites)
          const uint32_t mc = mu_base + (uint32_t)(s * NG + g);
          if (g == 0 && s > 0) TRACE(s - 1, 3);
          if (mc > 0) wait_bar(peers_done, (mc - 1) & 1, s_abort, p.status, B2S_GRU_ERR_PEERS);
          if (g == 0 && s > 0) TRACE(s - 1, 7);
          bulk_g2s_multicast(sA + c * FWD_CH, dst, FWD_CH, full, ALL);
        }
      };
      auto store_cols = [&](const FwdEpi& st, bf16* base_ptr, int s, const uint4& a, const uint4& b) {  // feature-major [H][T B]
        if (st.live) {
          uint16_t* o = reinterpret_cast<uint16_t*>(base_ptr) + (size_t)ucol * TB + (size_t)s * B + st.grow;
          const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 8; i++) {
            o[(size_t)(2 * i) * TB] = (uint16_t)(w[i] & 0xffffu);
            o[(size_t)(2 * i + 1) * TB] = (uint16_t)(w[i] >> 16);
          }
        }
      };
#pragma unroll
      for (int g = 0; g < 2; g++) {
        if (g < NG) {
          publish(e[g], g, 0);
          if (!(p.lbo_sbo_swap & (4 | 16))) store_cols(e[g], q.hst, 0, v0, v1);
        }
      }
      for (int t = 0; t < T; t++) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
          if (g >= NG) continue;
          FwdEpi& st = e[g];
          float gr[16], gz[16], gn[16];
          float kp = 1.0f;
          if (st.live && !(p.lbo_sbo_swap & 8)) {
            // feature-major gi [3H][T B]: consecutive lanes (batch rows) read consecutive addresses
            const float* gp = q.gi + (size_t)ucol * TB + (size_t)t * B + st.grow;
#pragma unroll
            for (int i = 0; i < 16; i++) {
              gr[i] = ld_stream_f32(gp + (size_t)i * TB);
              gz[i] = ld_stream_f32(gp + (size_t)(H + i) * TB);
              gn[i] = ld_stream_f32(gp + (size_t)(2 * H + i) * TB);
            }
            if (p.keep) kp = p.keep[(size_t)t * B + st.grow];
          } else {
#pragma unroll
            for (int i = 0; i < 16; i++) gr[i] = gz[i] = gn[i] = 0.0f;
          }
          if (lane == 0) wait_bar(&acc_full[g], (st_base[g] + t) & 1, s_abort, p.status, B2S_GRU_ERR_ACC);  // one poller per warp
          if (tid == 0 && g == 0) TRACE(t, 0);
          __syncwarp();
          tc_fence_after();
          const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * ACC_STRIDE + j0);
          float a[16];
          tmem_ld16(ta, a);
#pragma unroll
          for (int i = 0; i < 16; i++) gr[i] = sigmoid_fast(gr[i] + a[i] + s_bias[j0 + i]);
          tmem_ld16(ta + U, a);
#pragma unroll
          for (int i = 0; i < 16; i++) gz[i] = sigmoid_fast(gz[i] + a[i] + s_bias[U + j0 + i]);
          tmem_ld16(ta + 2 * U, a);
          tc_fence_before();
          mbar_arrive(&tmem_empty[g]);
#pragma unroll
          for (int i = 0; i < 16; i++) {
            a[i] += s_bias[3 * U + j0 + i];                                   // hn = gh_n + b_hn
            gn[i] = tanh_fast(gn[i] + s_bias[2 * U + j0 + i] + gr[i] * a[i]);  // n
          }
          float hm[16];   // h - n (saved for the backward pass)
#pragma unroll
          for (int i = 0; i < 16; i++) { hm[i] = st.h[i] - gn[i]; st.h[i] = fmaf(gz[i], hm[i], gn[i]); }  // h' = n + z (h - n)
          const uint4 y0 = make_uint4(pack_bf16(st.h[0], st.h[1]), pack_bf16(st.h[2], st.h[3]), pack_bf16(st.h[4], st.h[5]), pack_bf16(st.h[6], st.h[7]));
          const uint4 y1 = make_uint4(pack_bf16(st.h[8], st.h[9]), pack_bf16(st.h[10], st.h[11]), pack_bf16(st.h[12], st.h[13]), pack_bf16(st.h[14], st.h[15]));
#pragma unroll
          for (int i = 0; i < 16; i++) st.h[i] *= kp;  // the carry restarts from zero after a terminal step
          if (tid == 0 && g == 0) TRACE(t, 1);
          if (t + 1 < T) {
            publish(st, g, t + 1);
            if (tid == 0 && g == 0) TRACE(t, 2);
            if (!(p.lbo_sbo_swap & (4 | 16))) store_cols(st, q.hst, t + 1, v0, v1);
          }
          if (!(p.lbo_sbo_swap & (4 | 64))) store_cols(st, q.y, t, y0, y1);
          if (q.saved && !(p.lbo_sbo_swap & (4 | 32))) {
            float* sv = q.saved + (((size_t)t * p.n_mt + st.mt) * NCH + c) * SAVED_BLOCK;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int o = ((half * 4 + i) * MT + row) * 4;
              __stcs(reinterpret_cast<float4*>(sv + o), make_float4(gr[4 * i], gr[4 * i + 1], gr[4 * i + 2], gr[4 * i + 3]));
              __stcs(reinterpret_cast<float4*>(sv + MT * U + o), make_float4(gz[4 * i], gz[4 * i + 1], gz[4 * i + 2], gz[4 * i + 3]));
              __stcs(reinterpret_cast<float4*>(sv + 2 * MT * U + o), make_float4(gn[4 * i], gn[4 * i + 1], gn[4 * i + 2], gn[4 * i + 3]));
              __stcs(reinterpret_cast<float4*>(sv + 3 * MT * U + o), make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]));
              __stcs(reinterpret_cast<float4*>(sv + 4 * MT * U + o), make_float4(hm[4 * i], hm[4 * i + 1], hm[4 * i + 2], hm[4 * i + 3]));
            }
          }
        }
      }
#pragma unroll
      for (int g = 0; g < 2; g++) {
        if (e[g].live && q.h_last) {
#pragma unroll
          for (int i = 0; i < 4; i++)
            *reinterpret_cast<float4*>(q.h_last + (size_t)e[g].grow * H + ucol + 4 * i) =
                make_float4(e[g].h[4 * i], e[g].h[4 * i + 1], e[g].h[4 * i + 2], e[g].h[4 * i + 3]);
        }
      }
    } else if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer: gh_t[128 x 96] = h_{t-1}[128 x 512] W^T
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (int t = 0; t < T; t++) {
        for (int g = 0; g < NG; g++) {
          const uint32_t gm = mu_base + (uint32_t)(t * NG + g), gs = st_base[g] + t;
          if (gs > 0) wait_bar(&tmem_empty[g], (gs - 1) & 1, s_abort, p.status, B2S_GRU_ERR_TMEM);
          if (!(p.lbo_sbo_swap & 2)) {
            if (g == 0) TRACE_CH0(t, 0);
            wait_bar(full, gm & 1, s_abort, p.status, B2S_GRU_ERR_FULL);
            mbar_arrive_expect_tx(full, NCH * FWD_CH);  // arms the next phase (the next micro-step's chunks)
            if (g == 0) { TRACE(t, 4); TRACE(t, 5); TRACE_CH(t, 0); }
          }
          tc_fence_after();
          for (int ch = 0; ch < NCH; ch++) {
#pragma unroll
            for (int k2 = 0; k2 < U / 16; k2++) {
              const uint64_t da = desc(a0 + ch * FWD_CH + k2 * 2 * (MT * 16), MT * 16, 128, p.lbo_sbo_swap);
              const uint64_t db = desc(b0 + (ch * (U / 8) + k2 * 2) * (3 * U * 16), 3 * U * 16, 128, p.lbo_sbo_swap);
              mma_bf16_ss(tmem + g * ACC_STRIDE, da, db, IDESC, (ch | k2) != 0);
            }
          }
          mma_commit(&acc_full[g]);
          if (!(p.lbo_sbo_swap & 2)) mma_commit_multicast(peers_done, ALL);
          if (g == 0) TRACE(t, 6);
        }
      }
    }
    __syncwarp();  // roles reconverge before the next CTA-wide barrier
    mu_base += (uint32_t)(NG * T);
    st_base[0] += T;
    if (NG == 2) st_base[1] += T;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still signal its barriers
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// ----------------------------------------------------------------------------------------------------- backward
// Step t (descending), per element: dh = dy_t + keep_t dhs  (dhs: gradient with respect to the state step t + 1 started from)
//   dn = dh (1 - z), dz = dh (h - n), dnp = dn (1 - n^2), dr = dnp hn,
//   dgi_n = dnp, dgh_n = dnp r, dgi_z = dgh_z = dz z (1 - z), dgi_r = dgh_r = dr r (1 - r),
//   dhs <- dh z + dgh_t W_hh      (the sequential GEMM: [128 x 1536] x [1536 x 32] per CTA)
__global__ void __launch_bounds__(NTHREADS, 1) k_gru_bwd(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + W_BYTES + BWD_STAGES * BWD_CH);  // [BWD_STAGES]
  uint64_t* empty = full + BWD_STAGES;
  uint64_t* acc_full = full + 2 * BWD_STAGES;
  uint64_t* tmem_empty = acc_full + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x % NCH, slot = blockIdx.x / NCH, n_slots = gridDim.x / NCH;
  const int n_groups = p.n_problems * p.n_mt, T = p.T, B = p.B;

  const size_t TB = (size_t)T * B;
  if (tid == 0) {
    for (int i = 0; i < BWD_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(tmem_empty, NEPI);
    *s_abort = 0;
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(s_tmem, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  constexpr uint32_t IDESC = idesc_bf16(MT, U);
  constexpr int KC = 3 * H / 8;  // 192 core-matrix columns of K

  int loaded = -1;
  uint32_t base = 0;
  for (int g = slot; g < n_groups; g += n_slots, base += T) {
    const int pi = g / p.n_mt, mt = g % p.n_mt;
    const BwdProblem& q = p.pr[pi];
    if (pi != loaded) {
      __syncthreads();
      // B row n = own unit j; K index = chunk c' * 96 + gate * 32 + j'  <->  W_hh^T[c U + j][gate * H + c' U + j']
      for (int idx = tid; idx < U * KC; idx += NTHREADS) {
        const int n = idx / KC, kk = idx % KC;
        const int cp = kk / 12, gate = (kk % 12) / 4, j8 = kk % 4;
        const uint4 v = *reinterpret_cast<const uint4*>(q.wt + ((size_t)(c * U + n) * (3 * H) + gate * H + cp * U + j8 * 8));
        *reinterpret_cast<uint4*>(sB + kk * (U * 16) + n * 16) = v;
      }
      fence_proxy_async_smem();
      __syncthreads();
      loaded = pi;
    }
    int* flags = q.flags + mt * NCH;
    if (warp < 8) {
      const int row = (warp & 3) * 32 + lane, half = warp >> 2, j0 = half * 16;
      const int grow = mt * MT + row;
      const bool live = grow < B;
      const size_t ucol = (size_t)c * U + j0;
      float dhs[16];
#pragma unroll
      for (int i = 0; i < 16; i++) dhs[i] = 0.0f;
      for (int s = 0; s < T; s++) {
        const int t = T - 1 - s;
        float r[16], z[16], n[16], hn[16], hmn[16], dy[16];
        float kp = 1.0f;
        const float* sv = q.saved + (((size_t)t * p.n_mt + mt) * NCH + c) * SAVED_BLOCK;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int o = ((half * 4 + i) * MT + row) * 4;
          const float4 a = __ldcs(reinterpret_cast<const float4*>(sv + o));
          const float4 b = __ldcs(reinterpret_cast<const float4*>(sv + MT * U + o));
          const float4 d = __ldcs(reinterpret_cast<const float4*>(sv + 2 * MT * U + o));
          const float4 e = __ldcs(reinterpret_cast<const float4*>(sv + 3 * MT * U + o));
          const float4 f = __ldcs(reinterpret_cast<const float4*>(sv + 4 * MT * U + o));
          r[4 * i] = a.x; r[4 * i + 1] = a.y; r[4 * i + 2] = a.z; r[4 * i + 3] = a.w;
          z[4 * i] = b.x; z[4 * i + 1] = b.y; z[4 * i + 2] = b.z; z[4 * i + 3] = b.w;
          n[4 * i] = d.x; n[4 * i + 1] = d.y; n[4 * i + 2] = d.z; n[4 * i + 3] = d.w;
          hn[4 * i] = e.x; hn[4 * i + 1] = e.y; hn[4 * i + 2] = e.z; hn[4 * i + 3] = e.w;
          hmn[4 * i] = f.x; hmn[4 * i + 1] = f.y; hmn[4 * i + 2] = f.z; hmn[4 * i + 3] = f.w;
        }
        if (live) {
          const float* dp = q.dy + (size_t)ucol * TB + (size_t)t * B + grow;  // feature-major dy [H][T B]
#pragma unroll
          for (int i = 0; i < 16; i++) dy[i] = ld_stream_f32(dp + (size_t)i * TB);
          if (p.keep) kp = p.keep[(size_t)t * B + grow];
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) dy[i] = 0.0f;
        }
        float dr_[16], dz_[16], dn_[16], dg_[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float dh = fmaf(kp, dhs[i], dy[i]);
          const float dnp = dh * (1.0f - z[i]) * (1.0f - n[i] * n[i]);
          dn_[i] = dnp;
          dg_[i] = dnp * r[i];
          dz_[i] = dh * hmn[i] * z[i] * (1.0f - z[i]);
          dr_[i] = dnp * hn[i] * r[i] * (1.0f - r[i]);
          dhs[i] = dh * z[i];  // the direct path; the GEMM term is added below
        }
        auto pk = [](const float* v, int o) { return make_uint4(pack_bf16(v[o], v[o + 1]), pack_bf16(v[o + 2], v[o + 3]), pack_bf16(v[o + 4], v[o + 5]), pack_bf16(v[o + 6], v[o + 7])); };
        const uint4 r0 = pk(dr_, 0), r1 = pk(dr_, 8), z0 = pk(dz_, 0), z1 = pk(dz_, 8), n0 = pk(dn_, 0), n1 = pk(dn_, 8), g0 = pk(dg_, 0), g1 = pk(dg_, 8);
        // the chunk the peers multiply with W_hh goes out first (K order gate * 32 + j -> core-matrix column gate * 4 + j / 8);
        // the step's own outputs (dgi, dghn) are stored after the flag, off the exchange's critical path
        bf16* dst = q.xchg + (((size_t)s * p.n_mt + mt) * NCH + c) * (BWD_CH / 2);
        *reinterpret_cast<uint4*>(dst + ((0 + 2 * half) * MT + row) * 8) = r0;
        *reinterpret_cast<uint4*>(dst + ((1 + 2 * half) * MT + row) * 8) = r1;
        *reinterpret_cast<uint4*>(dst + ((4 + 2 * half) * MT + row) * 8) = z0;
        *reinterpret_cast<uint4*>(dst + ((5 + 2 * half) * MT + row) * 8) = z1;
        *reinterpret_cast<uint4*>(dst + ((8 + 2 * half) * MT + row) * 8) = g0;
        *reinterpret_cast<uint4*>(dst + ((9 + 2 * half) * MT + row) * 8) = g1;
        __syncwarp();
        named_bar_sync(1, NEPI);
        if (tid == 0) {
          __threadfence();
          st_release(&flags[c], s + 1);
        }
        if (live) {  // feature-major dgi [3H][T B], dghn [H][T B]: lanes = batch rows
          uint16_t* o = reinterpret_cast<uint16_t*>(q.dgi) + (size_t)ucol * TB + (size_t)t * B + grow;
          uint16_t* og = reinterpret_cast<uint16_t*>(q.dghn) + (size_t)ucol * TB + (size_t)t * B + grow;
          const uint32_t wr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w}, wz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
          const uint32_t wn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w}, wg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int i = 0; i < 8; i++) {
            o[(size_t)(2 * i) * TB] = (uint16_t)(wr[i] & 0xffffu);              o[(size_t)(2 * i + 1) * TB] = (uint16_t)(wr[i] >> 16);
            o[(size_t)(H + 2 * i) * TB] = (uint16_t)(wz[i] & 0xffffu);          o[(size_t)(H + 2 * i + 1) * TB] = (uint16_t)(wz[i] >> 16);
            o[(size_t)(2 * H + 2 * i) * TB] = (uint16_t)(wn[i] & 0xffffu);      o[(size_t)(2 * H + 2 * i + 1) * TB] = (uint16_t)(wn[i] >> 16);
            og[(size_t)(2 * i) * TB] = (uint16_t)(wg[i] & 0xffffu);             og[(size_t)(2 * i + 1) * TB] = (uint16_t)(wg[i] >> 16);
          }
        }
        if (lane == 0) wait_bar(acc_full, (base + s) & 1, s_abort, p.status, B2S_GRU_ERR_ACC);  // one poller per warp
        __syncwarp();
        tc_fence_after();
        float a[16];
        tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)j0, a);
        tc_fence_before();
        mbar_arrive(tmem_empty);
#pragma unroll
        for (int i = 0; i < 16; i++) dhs[i] += a[i];
      }
      if (live && q.dh0) {
#pragma unroll
        for (int i = 0; i < 4; i++)
          *reinterpret_cast<float4*>(q.dh0 + (size_t)grow * H + ucol + 4 * i) = make_float4(dhs[4 * i], dhs[4 * i + 1], dhs[4 * i + 2], dhs[4 * i + 3]);
      }
    } else if (warp == 8) {
      // Producer.  The 16 peers' flags are polled AT ONCE, one lane each, and the generic -> async proxy fence is paid once per
      // step: one thread polling them in turn (an L2 round trip each) with a fence per chunk was the longest serial chain of a
      // step (profiles/r02m_gru_bwd.txt: the producer warp never waited for a free stage).  The polling lanes' acquires are
      // ordered before lane 0's copies by the warp barrier; lane 0 then only waits for free stages.
      uint32_t use = base * NCH;
      for (int s = 0; s < T; s++) {
        if (lane < NCH) {
          wait_flag(&flags[lane], s + 1, s_abort, p.status, B2S_GRU_ERR_FLAG);
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) {
          fence_proxy_async();
          for (int ch = 0; ch < NCH; ch++) {
            const uint32_t u = use + ch, st = u % BWD_STAGES, nuse = u / BWD_STAGES;
            if (nuse > 0) wait_bar(&empty[st], (nuse - 1) & 1, s_abort, p.status, B2S_GRU_ERR_EMPTY);
            mbar_arrive_expect_tx(&full[st], BWD_CH);
            bulk_g2s(sA + st * BWD_CH, q.xchg + (((size_t)s * p.n_mt + mt) * NCH + ch) * (BWD_CH / 2), BWD_CH, &full[st]);
          }
        }
        use += NCH;
        __syncwarp();
      }
    } else if (lane == 0) {
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      uint32_t use = base * NCH;
      for (int s = 0; s < T; s++) {
        const uint32_t gs = base + s;
        if (gs > 0) wait_bar(tmem_empty, (gs - 1) & 1, s_abort, p.status, B2S_GRU_ERR_TMEM);
        tc_fence_after();
        for (int ch = 0; ch < NCH; ch++, use++) {
          const uint32_t st = use % BWD_STAGES, nuse = use / BWD_STAGES;
          wait_bar(&full[st], nuse & 1, s_abort, p.status, B2S_GRU_ERR_FULL);
          tc_fence_after();
#pragma unroll
          for (int k6 = 0; k6 < 3 * U / 16; k6++) {
            const uint64_t da = desc(a0 + st * BWD_CH + k6 * 2 * (MT * 16), MT * 16, 128, p.lbo_sbo_swap);
            const uint64_t db = desc(b0 + (ch * 12 + k6 * 2) * (U * 16), U * 16, 128, p.lbo_sbo_swap);
            mma_bf16_ss(tmem, da, db, IDESC, (ch | k6) != 0);
          }
          mma_commit(&empty[st]);
        }
        mma_commit(acc_full);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 32);
}

// -------------------------------------------------------------------------------------------------------- host
thread_local std::string g_gru_error;
int gfail(int code, const std::string& msg) { g_gru_error = msg; return code; }

int slots_for(int n_groups) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  int slots = sms / NCH;
  if (const char* e = getenv("B2S_GRU_SLOTS")) { const int v = atoi(e); if (v >= 1 && v < slots) slots = v; }
  return n_groups < slots ? n_groups : slots;
}

template <class P>
int launch(void (*kernel)(const P), const P& prm, int smem_bytes, int n_groups, void* stream) {
  static thread_local void (*configured)(const P) = nullptr;
  if (configured != kernel) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
      return gfail(B2S_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(cudaGetLastError()));
    configured = kernel;
  }
  const int slots = slots_for(n_groups);
  if (slots < 1) return gfail(B2S_ERR_CUDA, "the device has fewer than 16 SMs");
  void* args[] = {const_cast<P*>(&prm)};
  // every CTA of a group waits for its 15 peers each step: the launch must guarantee co-residency
  cudaError_t e = getenv("B2S_GRU_NO_COOP")
                      ? cudaLaunchKernel((const void*)kernel, dim3(slots * NCH), dim3(NTHREADS), args, smem_bytes, (cudaStream_t)stream)
                      : cudaLaunchCooperativeKernel((const void*)kernel, dim3(slots * NCH), dim3(NTHREADS), args, smem_bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) return gfail(B2S_ERR_CUDA, std::string("gru launch: ") + cudaGetErrorString(e));
  return B2S_OK;
}

// The forward kernel runs one 16-CTA thread-block cluster per group (non-portable cluster size: one cluster per GPC).
int launch_fwd(const FwdParams& prm, int n_groups, void* stream) {
  static thread_local int max_clusters = -1;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCH; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(FWD_THREADS); cfg.dynamicSmemBytes = FWD_SMEM; cfg.stream = (cudaStream_t)stream; cfg.attrs = attr; cfg.numAttrs = 1;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(k_gru_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_gru_fwd, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return gfail(B2S_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(cudaGetLastError()));
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cfg.gridDim = dim3((sms / NCH) * NCH);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_gru_fwd, &cfg) != cudaSuccess)
      return gfail(B2S_ERR_CUDA, std::string("cudaOccupancyMaxActiveClusters: ") + cudaGetErrorString(cudaGetLastError()));
    max_clusters = n;
    if (getenv("B2S_GRU_DEBUG")) fprintf(stderr, "[b200sim] k_gru_fwd: %d co-resident 16-CTA clusters on %d SMs\n", n, sms);
  }
  int slots = max_clusters;
  if (const char* e = getenv("B2S_GRU_SLOTS")) { const int v = atoi(e); if (v >= 1 && v < slots) slots = v; }
  if (n_groups < slots) slots = n_groups;
  if (slots < 1) return gfail(B2S_ERR_CUDA, "no 16-CTA cluster of the forward kernel fits on this device");
  cfg.gridDim = dim3(slots * NCH);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_gru_fwd, prm);
  if (e != cudaSuccess) return gfail(B2S_ERR_CUDA, std::string("gru forward launch: ") + cudaGetErrorString(e));
  return B2S_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" const char* b200sim_gru_last_error(void) { return g_gru_error.c_str(); }

#ifdef B2S_GRU_TRACE
extern "C" int b200sim_debug_gru_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * (n < 8 * 256 ? n : 8 * 256));
}
extern "C" int b200sim_debug_gru_trace_ch0(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_ch0, sizeof(long long) * (n < 16 * 256 ? n : 16 * 256));
}
extern "C" int b200sim_debug_gru_trace_ch(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_ch, sizeof(long long) * (n < 16 * 256 ? n : 16 * 256));
}
#endif

extern "C" size_t b200sim_gru_workspace_bytes(int n_problems, int n_steps, int batch, int backward) {
  if (n_problems < 1 || n_steps < 0 || batch < 0) return 0;
  const size_t n_mt = (batch + MT - 1) / MT;
  const size_t per = backward ? align256((size_t)n_steps * n_mt * NCH * BWD_CH) : align256((size_t)(n_steps + 1) * n_mt * NCH * FWD_CH);
  return 256 + (size_t)n_problems * (per + align256(n_mt * NCH * sizeof(int)));
}

extern "C" int b200sim_gru_forward(void* stream, int n_problems, const b200sim_gru_fwd* pr, int n_steps, int batch, const float* keep,
                                   void* workspace, size_t workspace_bytes) {
  if (n_problems < 1 || n_problems > B2S_GRU_MAX_PROBLEMS || !pr || n_steps < 0 || batch < 0) return gfail(B2S_ERR_BAD_ARG, "bad argument");
  if (n_steps == 0 || batch == 0) return B2S_OK;
  if (!workspace || workspace_bytes < b200sim_gru_workspace_bytes(n_problems, n_steps, batch, 0)) return gfail(B2S_ERR_BAD_ARG, "workspace too small");
  FwdParams prm{};
  prm.n_problems = n_problems; prm.T = n_steps; prm.B = batch; prm.n_mt = (batch + MT - 1) / MT; prm.keep = keep;
  prm.lbo_sbo_swap = (getenv("B2S_GRU_SWAP") ? 1 : 0) | (getenv("B2S_GRU_NOXCHG") ? 2 : 0) | (getenv("B2S_GRU_NOSTORE") ? 4 : 0) |
                     (getenv("B2S_GRU_NOLOAD") ? 8 : 0) | (getenv("B2S_GRU_NOHST") ? 16 : 0) | (getenv("B2S_GRU_NOSAVED") ? 32 : 0) |
                     (getenv("B2S_GRU_NOY") ? 64 : 0);  // debug / timing experiments only
  uint8_t* ws = (uint8_t*)workspace;
  prm.status = (int*)ws;
  const size_t flag_bytes = align256((size_t)prm.n_mt * NCH * sizeof(int));
  const size_t per = align256((size_t)(n_steps + 1) * prm.n_mt * NCH * FWD_CH);
  uint8_t* flags0 = ws + 256;
  uint8_t* x0 = flags0 + (size_t)n_problems * flag_bytes;
  for (int i = 0; i < n_problems; i++) {
    const b200sim_gru_fwd& s = pr[i];
    if (!s.w_hh || !s.gi || !s.b_i || !s.b_hn || !s.y || !s.hst) return gfail(B2S_ERR_BAD_ARG, "null argument");
    FwdProblem& d = prm.pr[i];
    d.w = (const bf16*)s.w_hh; d.gi = s.gi; d.b_i = s.b_i; d.b_hn = s.b_hn; d.h0 = s.h0; d.y = (bf16*)s.y; d.hst = (bf16*)s.hst;
    d.h_last = s.h_last; d.saved = s.saved;
    d.flags = (int*)(flags0 + i * flag_bytes);
    d.xchg = (bf16*)(x0 + i * per);
  }
  cudaError_t e = cudaMemsetAsync(ws, 0, 256 + (size_t)n_problems * flag_bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) return gfail(B2S_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
  return launch_fwd(prm, n_problems * ((prm.n_mt + 1) / 2), stream);  // work items: pairs of row tiles
}

extern "C" int b200sim_gru_backward(void* stream, int n_problems, const b200sim_gru_bwd* pr, int n_steps, int batch, const float* keep,
                                    void* workspace, size_t workspace_bytes) {
  if (n_problems < 1 || n_problems > B2S_GRU_MAX_PROBLEMS || !pr || n_steps < 0 || batch < 0) return gfail(B2S_ERR_BAD_ARG, "bad argument");
  if (n_steps == 0 || batch == 0) return B2S_OK;
  if (!workspace || workspace_bytes < b200sim_gru_workspace_bytes(n_problems, n_steps, batch, 1)) return gfail(B2S_ERR_BAD_ARG, "workspace too small");
  BwdParams prm{};
  prm.n_problems = n_problems; prm.T = n_steps; prm.B = batch; prm.n_mt = (batch + MT - 1) / MT; prm.keep = keep;
  prm.lbo_sbo_swap = getenv("B2S_GRU_SWAP") ? 1 : 0;
  uint8_t* ws = (uint8_t*)workspace;
  prm.status = (int*)ws;
  const size_t flag_bytes = align256((size_t)prm.n_mt * NCH * sizeof(int));
  const size_t per = align256((size_t)n_steps * prm.n_mt * NCH * BWD_CH);
  uint8_t* flags0 = ws + 256;
  uint8_t* x0 = flags0 + (size_t)n_problems * flag_bytes;
  for (int i = 0; i < n_problems; i++) {
    const b200sim_gru_bwd& s = pr[i];
    if (!s.w_hh_t || !s.dy || !s.saved || !s.dgi || !s.dghn) return gfail(B2S_ERR_BAD_ARG, "null argument");
    BwdProblem& d = prm.pr[i];
    d.wt = (const bf16*)s.w_hh_t; d.dy = s.dy; d.saved = s.saved; d.dgi = (bf16*)s.dgi; d.dghn = (bf16*)s.dghn; d.dh0 = s.dh0;
    d.flags = (int*)(flags0 + i * flag_bytes);
    d.xchg = (bf16*)(x0 + i * per);
  }
  cudaError_t e = cudaMemsetAsync(ws, 0, 256 + (size_t)n_problems * flag_bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) return gfail(B2S_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
  return launch<BwdParams>(k_gru_bwd, prm, BWD_SMEM, n_problems * prm.n_mt, stream);
}

extern "C" size_t b200sim_gru_saved_floats(int n_steps, int batch) {
  return (size_t)n_steps * ((batch + MT - 1) / MT) * NCH * SAVED_BLOCK;
}
