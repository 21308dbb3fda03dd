// Optimiser step of the PPO update on one FLAT parameter buffer: the optax chain the walking and K-Bot
// tasks build (examples/walking.py:373-387, examples/kbot/train.py get_optimizer)
//
//   zero_nans -> clip_by_global_norm(max_norm) -> add_decayed_weights(wd) -> scale_by_adam(b1, b2, eps) ->
//   scale_by_schedule(warmup_constant_schedule(lr_init, lr_peak, warmup_steps)) -> scale(-1)
//
// as two HBM-bound passes over the flat buffers: (1) sum of squares of the NaN-cleaned gradient, one partial per CTA;
// (2) every CTA re-reduces the partials in a fixed order (deterministic), derives the clip factor and the step's learning
// rate from the device-side step counter, and updates param / mu / nu in place, also writing the bf16 shadow copy the
// tensor-core GEMMs of the next replay read.  Algorithmic bytes per parameter: pass 1 reads 4; pass 2 reads 16, writes 12 + 2.
// With data-parallel ranks the gradient buffer is all-reduced (one NCCL call on the same flat buffer) between backward and
// this step; `grad_scale` folds the 1 / world_size of the mean into pass 1 and 2.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/b200sim.h"
#include "../../include/b200sim_codes.h"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float clean(float g, float s) {
  g *= s;
  return (g != g) ? 0.0f : g;  // optax.zero_nans
}

__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.0f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < NT / 32 ? sm[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// partial[blockIdx.x] = sum over this CTA's grid-stride slice of clean(g)^2; CTA 0 also advances the step counter.
__global__ void __launch_bounds__(NT) k_optim_sqnorm(size_t n, const float* __restrict__ g, float grad_scale, float* __restrict__ partial,
                                                     int* __restrict__ count) {
  __shared__ float sm[NT / 32];
  float acc = 0.0f;
  const size_t n4 = n / 4, stride = (size_t)gridDim.x * NT;
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float a = clean(v.x, grad_scale), b = clean(v.y, grad_scale), c = clean(v.z, grad_scale), d = clean(v.w, grad_scale);
    acc += (a * a + b * b) + (c * c + d * d);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
    const float a = clean(g[4 * n4 + threadIdx.x], grad_scale);
    acc += a * a;
  }
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
    if (blockIdx.x == 0) count[0] += 1;
  }
}

struct Hyper {
  float lr_init, lr_peak, b1, b2, eps, wd, max_norm, grad_scale;
  int warmup_steps;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const Hyper& h, float clip, float lr, float c1, float c2) {
  g = clean(g, h.grad_scale) * clip + h.wd * p;  // clip_by_global_norm, then add_decayed_weights
  m = h.b1 * m + (1.0f - h.b1) * g;
  v = h.b2 * v + (1.0f - h.b2) * g * g;
  p -= lr * (m * c1) / (sqrtf(v * c2) + h.eps);
}

__global__ void __launch_bounds__(NT) k_optim_apply(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
                                                    float* __restrict__ nu, __nv_bfloat16* __restrict__ p16, const float* __restrict__ partial,
                                                    int n_partial, const int* __restrict__ count, Hyper h, float* __restrict__ stats) {
  __shared__ float sm[NT / 32];
  __shared__ float s_clip, s_lr, s_c1, s_c2;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < n_partial; i += NT) acc += partial[i];  // same order in every CTA
  const float tot = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(tot);
    const int t = count[0];  // 1 for the first step (scale_by_adam's bias correction); the schedule sees t - 1
    const float frac = h.warmup_steps > 0 ? fminf((float)(t - 1) / (float)h.warmup_steps, 1.0f) : 1.0f;
    s_clip = norm < h.max_norm ? 1.0f : h.max_norm / norm;  // optax.clip_by_global_norm
    s_lr = h.lr_init + (h.lr_peak - h.lr_init) * frac;
    s_c1 = 1.0f / (1.0f - powf(h.b1, (float)t));
    s_c2 = 1.0f / (1.0f - powf(h.b2, (float)t));
    if (blockIdx.x == 0 && stats) { stats[0] = norm; stats[1] = s_lr; stats[2] = s_clip; }
  }
  __syncthreads();
  const float clip = s_clip, lr = s_lr, c1 = s_c1, c2 = s_c2;
  const size_t n4 = n / 4, stride = (size_t)gridDim.x * NT;
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(mu)[i], vv = reinterpret_cast<float4*>(nu)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    adam_one(pv.x, gv.x, mv.x, vv.x, h, clip, lr, c1, c2);
    adam_one(pv.y, gv.y, mv.y, vv.y, h, clip, lr, c1, c2);
    adam_one(pv.z, gv.z, mv.z, vv.z, h, clip, lr, c1, c2);
    adam_one(pv.w, gv.w, mv.w, vv.w, h, clip, lr, c1, c2);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(mu)[i] = mv;
    reinterpret_cast<float4*>(nu)[i] = vv;
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&lo);
      w.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p16)[i] = w;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
    const size_t i = 4 * n4 + threadIdx.x;
    float pv = p[i], mv = mu[i], vv = nu[i];
    adam_one(pv, g[i], mv, vv, h, clip, lr, c1, c2);
    p[i] = pv; mu[i] = mv; nu[i] = vv;
    if (p16) p16[i] = __float2bfloat16_rn(pv);
  }
}

thread_local std::string g_opt_error;
int ofail(int code, const char* msg) { g_opt_error = msg; return code; }

int grid_for(size_t n) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const size_t want = (n / 4 + NT - 1) / NT;
  const size_t cap = (size_t)sms * 4;  // a multiple of the SM count, at most B2S_OPTIM_MAX_PARTIALS
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

extern "C" const char* b200sim_optimizer_last_error(void) { return g_opt_error.c_str(); }

extern "C" size_t b200sim_optimizer_scratch_bytes(void) { return sizeof(float) * B2S_OPTIM_MAX_PARTIALS; }

extern "C" int b200sim_optimizer_step(void* stream, size_t n, float* params, const float* grads, float* mu, float* nu, void* params_bf16,
                                      int32_t* step_count, float grad_scale, float lr_init, float lr_peak, int warmup_steps, float b1,
                                      float b2, float eps, float weight_decay, float max_norm, void* scratch, float* stats) {
  if (n == 0) return B2S_OK;
  if (!params || !grads || !mu || !nu || !step_count || !scratch) return ofail(B2S_ERR_BAD_ARG, "null argument");
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)mu | (uintptr_t)nu) & 15 || ((uintptr_t)params_bf16 & 7))
    return ofail(B2S_ERR_BAD_ARG, "flat buffers must be 16-byte aligned (bf16 shadow: 8)");
  const int grid = grid_for(n);
  if (grid > B2S_OPTIM_MAX_PARTIALS) return ofail(B2S_ERR_UNSUPPORTED, "more CTAs than partial-sum slots");
  Hyper h{lr_init, lr_peak, b1, b2, eps, weight_decay, max_norm, grad_scale, warmup_steps};
  cudaStream_t s = (cudaStream_t)stream;
  k_optim_sqnorm<<<grid, NT, 0, s>>>(n, grads, grad_scale, (float*)scratch, step_count);
  k_optim_apply<<<grid, NT, 0, s>>>(n, params, grads, mu, nu, (__nv_bfloat16*)params_bf16, (const float*)scratch, grid, step_count, h, stats);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ofail(B2S_ERR_CUDA, cudaGetErrorString(e));
  return B2S_OK;
}
