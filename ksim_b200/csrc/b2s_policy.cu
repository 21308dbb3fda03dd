// Policy-head kernels of the replay path: the actor's mixture-of-Gaussians head of examples/walking.py:129-146
// (prediction -> means / standard deviations / logits -> xax.MixtureOfGaussians) as one pass each for log-prob, its backward
// and sampling.  Elementwise HBM-bound work: per (row, joint) 3 M inputs, one output.
//
//   mean_k = clip(pred[j M + k] + mean_bias[j], lo[j], hi[j])                       (walking.py:140-143: ZEROS bias, ClipPositions)
//   std_k  = min((softplus(pred[J M + j M + k]) + min_std) var_scale, max_std)      (walking.py:137)
//   logw_k = log_softmax(pred[2 J M + j M + ..])_k
//   log p(x) = logsumexp_k(logw_k - 0.5 ((x - mean_k) / std_k)^2 - log std_k - 0.5 log 2 pi)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <string>

#include "../../include/b200sim.h"
#include "b2s_common.cuh"

namespace {

constexpr int MAXM = 16;
constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

struct MixArgs {
  int rows, J, M;
  const float* pred;       // [rows][3 J M]
  const float* mean_bias;  // [J] or null
  const float* lo;         // [J] or null (no clip)
  const float* hi;
  float min_std, var_scale, max_std;
};

struct Comp {
  float mean[MAXM], std[MAXM], logw[MAXM];
  bool mean_free[MAXM], std_free[MAXM];
  float sraw[MAXM];
};

__device__ __forceinline__ void load_components(const MixArgs& a, int row, int j, Comp& c) {
  const int JM = a.J * a.M;
  const float* p = a.pred + (size_t)row * 3 * JM + j * a.M;
  const float bias = a.mean_bias ? a.mean_bias[j] : 0.0f;
  const float lo = a.lo ? a.lo[j] : -INFINITY, hi = a.hi ? a.hi[j] : INFINITY;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAXM; k++) {
    if (k < a.M) {
      const float raw = p[k] + bias;
      c.mean_free[k] = raw > lo && raw < hi;
      c.mean[k] = fminf(fmaxf(raw, lo), hi);
      const float s = p[JM + k];
      c.sraw[k] = s;
      const float sp = s > 20.0f ? s : log1pf(expf(s));
      const float sd = (sp + a.min_std) * a.var_scale;
      c.std_free[k] = sd < a.max_std;
      c.std[k] = fminf(sd, a.max_std);
      c.logw[k] = p[2 * JM + k];
      mx = fmaxf(mx, c.logw[k]);
    }
  }
  float se = 0.0f;
#pragma unroll
  for (int k = 0; k < MAXM; k++) if (k < a.M) se += expf(c.logw[k] - mx);
  const float lse = mx + logf(se);
#pragma unroll
  for (int k = 0; k < MAXM; k++) if (k < a.M) c.logw[k] -= lse;
}

// log p(x) and the per-component terms (logw_k + log N(x; mean_k, std_k)) left in t[]
__device__ __forceinline__ float mixture_logp(const MixArgs& a, const Comp& c, float x, float* t) {
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAXM; k++) {
    if (k < a.M) {
      const float u = (x - c.mean[k]) / c.std[k];
      t[k] = c.logw[k] - 0.5f * u * u - logf(c.std[k]) - HALF_LOG_2PI;
      mx = fmaxf(mx, t[k]);
    }
  }
  float se = 0.0f;
#pragma unroll
  for (int k = 0; k < MAXM; k++) if (k < a.M) se += expf(t[k] - mx);
  return mx + logf(se);
}

__global__ void __launch_bounds__(256) k_mixture_logprob(const MixArgs a, const float* __restrict__ action, float* __restrict__ logp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)a.rows * a.J) return;
  const int row = idx / a.J, j = idx - (size_t)row * a.J;
  Comp c;
  float t[MAXM];
  load_components(a, row, j, c);
  logp[idx] = mixture_logp(a, c, action[idx], t);
}

// d_pred[row][..] from g = d loss / d logp[row][j] (d_logp_row: one value per row, the PPO loss kernel's layout)
template <class OutT>
__global__ void __launch_bounds__(256) k_mixture_logprob_bwd(const MixArgs a, const float* __restrict__ action, const float* __restrict__ d_logp,
                                                             int per_row, OutT* __restrict__ d_pred) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)a.rows * a.J) return;
  const int row = idx / a.J, j = idx - (size_t)row * a.J;
  Comp c;
  float t[MAXM];
  load_components(a, row, j, c);
  const float x = action[idx];
  const float lp = mixture_logp(a, c, x, t);
  const float g = per_row ? d_logp[row] : d_logp[idx];
  const int JM = a.J * a.M;
  OutT* o = d_pred + (size_t)row * 3 * JM + j * a.M;
#pragma unroll
  for (int k = 0; k < MAXM; k++) {
    if (k < a.M) {
      const float w = expf(t[k] - lp);  // responsibility of component k
      const float d = x - c.mean[k], inv = 1.0f / c.std[k];
      const float dmean = c.mean_free[k] ? g * w * d * inv * inv : 0.0f;
      const float dstd = g * w * (d * d * inv * inv * inv - inv);
      const float sig = 1.0f / (1.0f + expf(-c.sraw[k]));
      const float dsraw = c.std_free[k] ? dstd * a.var_scale * sig : 0.0f;
      const float dlogit = g * (w - expf(c.logw[k]));
      o[k] = (OutT)dmean;
      o[JM + k] = (OutT)dsraw;
      o[2 * JM + k] = (OutT)dlogit;
    }
  }
}

// Sampling (xax.MixtureOfGaussians.sample: component ~ Categorical(softmax(logits)) by Gumbel-max, then a normal draw) or the
// mode of the heaviest component (argmax); threefry counter = (call counter, row J + j).  Also returns log p(action).
__global__ void __launch_bounds__(256) k_mixture_sample(const MixArgs a, const uint32_t* __restrict__ key, const int* __restrict__ counter,
                                                        int argmax, float* __restrict__ action, float* __restrict__ logp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)a.rows * a.J) return;
  const int row = idx / a.J, j = idx - (size_t)row * a.J;
  Comp c;
  float t[MAXM];
  load_components(a, row, j, c);
  const b2s::U2 sub = b2s::threefry2x32(key[0], key[1], (uint32_t)(counter ? counter[0] : 0), (uint32_t)idx);
  b2s::Key k2; k2.k0 = sub.a; k2.k1 = sub.b;
  int best = 0;
  float bv = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAXM; k++) {
    if (k < a.M) {
      float v = c.logw[k];
      if (!argmax) {
        const float u = fmaxf(b2s::bits_to_unit(b2s::key_bits(k2, k)), 1e-20f);
        v -= logf(-logf(u));
      }
      if (v > bv) { bv = v; best = k; }
    }
  }
  float mean = 0.0f, sd = 0.0f;
#pragma unroll
  for (int k = 0; k < MAXM; k++) if (k == best) { mean = c.mean[k]; sd = c.std[k]; }
  const float x = argmax ? mean : mean + sd * b2s::key_normal(k2, MAXM);
  action[idx] = x;
  if (logp) logp[idx] = mixture_logp(a, c, x, t);
}

thread_local std::string g_pol_error;
int pfail(int code, const std::string& msg) { g_pol_error = msg; return code; }

int fill(MixArgs& a, int rows, int J, int M, const float* pred, const float* mean_bias, const float* lo, const float* hi, float min_std,
         float var_scale, float max_std) {
  if (rows < 0 || J < 1 || M < 1 || M > MAXM) return pfail(B2S_ERR_UNSUPPORTED, "mixture: 1 <= n_mix <= 16, n_joints >= 1");
  if (!pred || ((lo == nullptr) != (hi == nullptr))) return pfail(B2S_ERR_BAD_ARG, "null argument");
  a.rows = rows; a.J = J; a.M = M; a.pred = pred; a.mean_bias = mean_bias; a.lo = lo; a.hi = hi;
  a.min_std = min_std; a.var_scale = var_scale; a.max_std = max_std;
  return B2S_OK;
}
#define LAUNCH_OK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return pfail(B2S_ERR_CUDA, cudaGetErrorString(e_)); } while (0)

}  // namespace

extern "C" const char* b200sim_policy_last_error(void) { return g_pol_error.c_str(); }

extern "C" int b200sim_mixture_logprob(void* stream, int rows, int n_joints, int n_mix, const float* pred, const float* action,
                                       const float* mean_bias, const float* lo, const float* hi, float min_std, float var_scale,
                                       float max_std, float* logp) {
  MixArgs a;
  if (int rc = fill(a, rows, n_joints, n_mix, pred, mean_bias, lo, hi, min_std, var_scale, max_std)) return rc;
  if (rows == 0) return B2S_OK;
  if (!action || !logp) return pfail(B2S_ERR_BAD_ARG, "null argument");
  const size_t n = (size_t)rows * n_joints;
  k_mixture_logprob<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, action, logp);
  LAUNCH_OK();
  return B2S_OK;
}

extern "C" int b200sim_mixture_logprob_backward(void* stream, int rows, int n_joints, int n_mix, const float* pred, const float* action,
                                                const float* mean_bias, const float* lo, const float* hi, float min_std, float var_scale,
                                                float max_std, const float* d_logp, int d_logp_per_row, void* d_pred, int d_pred_bf16) {
  MixArgs a;
  if (int rc = fill(a, rows, n_joints, n_mix, pred, mean_bias, lo, hi, min_std, var_scale, max_std)) return rc;
  if (rows == 0) return B2S_OK;
  if (!action || !d_logp || !d_pred) return pfail(B2S_ERR_BAD_ARG, "null argument");
  const size_t n = (size_t)rows * n_joints;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (d_pred_bf16)
    k_mixture_logprob_bwd<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a, action, d_logp, d_logp_per_row, (__nv_bfloat16*)d_pred);
  else
    k_mixture_logprob_bwd<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a, action, d_logp, d_logp_per_row, (float*)d_pred);
  LAUNCH_OK();
  return B2S_OK;
}

extern "C" int b200sim_mixture_sample(void* stream, int rows, int n_joints, int n_mix, const float* pred, const float* mean_bias,
                                      const float* lo, const float* hi, float min_std, float var_scale, float max_std,
                                      const uint32_t* key, const int32_t* counter, int argmax, float* action, float* logp) {
  MixArgs a;
  if (int rc = fill(a, rows, n_joints, n_mix, pred, mean_bias, lo, hi, min_std, var_scale, max_std)) return rc;
  if (rows == 0) return B2S_OK;
  if (!action || !key) return pfail(B2S_ERR_BAD_ARG, "null argument");
  const size_t n = (size_t)rows * n_joints;
  k_mixture_sample<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, key, counter, argmax, action, logp);
  LAUNCH_OK();
  return B2S_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Metric histograms (RLTask.get_histogram, ksim/task/rl.py:1521-1541: every logged metric array with more than one element is
// reduced to jnp.histogram(arr, bins) + min / max / sum / sum of squares / mean).  Two HBM-bound passes over the array:
// (1) per-CTA min / max / sum / sum of squares -> partials; (2) every CTA re-reduces the partials in a fixed order, derives the
// bin edges (numpy's rule: [min, max], widened by +-0.5 when they coincide; the last bin is closed on the right), counts
// into a shared-memory histogram and merges it with integer atomics (exact and order-independent).
namespace {

constexpr int HNT = 256;
constexpr int HMAXBINS = 1024;

__global__ void __launch_bounds__(HNT) k_hist_stats(size_t n, const float* __restrict__ x, float* __restrict__ partial) {
  __shared__ float s_mn[HNT / 32], s_mx[HNT / 32];
  __shared__ double s_sum[HNT / 32], s_sq[HNT / 32];
  float mn = INFINITY, mx = -INFINITY;
  double sum = 0.0, sq = 0.0;
  for (size_t i = (size_t)blockIdx.x * HNT + threadIdx.x; i < n; i += (size_t)gridDim.x * HNT) {
    const float v = x[i];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
    sum += v; sq += (double)v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_mn[w] = mn; s_mx[w] = mx; s_sum[w] = sum; s_sq[w] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < HNT / 32; k++) { mn = fminf(mn, s_mn[k]); mx = fmaxf(mx, s_mx[k]); sum += s_sum[k]; sq += s_sq[k]; }
    float* p = partial + 6 * blockIdx.x;   // min, max, sum (double), sum of squares (double)
    p[0] = mn; p[1] = mx;
    *reinterpret_cast<double*>(p + 2) = sum;
    *reinterpret_cast<double*>(p + 4) = sq;
  }
}

__global__ void __launch_bounds__(HNT) k_hist_count(size_t n, const float* __restrict__ x, const float* __restrict__ partial, int n_partial,
                                                    int bins, int* __restrict__ counts, float* __restrict__ limits, float* __restrict__ stats) {
  __shared__ int s_cnt[HMAXBINS];
  __shared__ float s_lo, s_hi;
  for (int b = threadIdx.x; b < bins; b += HNT) s_cnt[b] = 0;
  if (threadIdx.x == 0) {
    float mn = INFINITY, mx = -INFINITY;
    double sum = 0.0, sq = 0.0;
    for (int k = 0; k < n_partial; k++) {
      const float* p = partial + 6 * k;
      mn = fminf(mn, p[0]); mx = fmaxf(mx, p[1]);
      sum += *reinterpret_cast<const double*>(p + 2);
      sq += *reinterpret_cast<const double*>(p + 4);
    }
    float lo = mn, hi = mx;
    if (lo == hi) { lo -= 0.5f; hi += 0.5f; }
    s_lo = lo; s_hi = hi;
    if (blockIdx.x == 0) {
      stats[0] = mn; stats[1] = mx; stats[2] = (float)sum; stats[3] = (float)sq; stats[4] = (float)(sum / (double)n);
      for (int b = 0; b < bins; b++) limits[b] = b + 1 == bins ? hi : lo + (hi - lo) * ((float)(b + 1) / (float)bins);  // edges[1:]
    }
  }
  __syncthreads();
  const float lo = s_lo, hi = s_hi, scale = (float)bins / (hi - lo);
  for (size_t i = (size_t)blockIdx.x * HNT + threadIdx.x; i < n; i += (size_t)gridDim.x * HNT) {
    const float v = x[i];
    int b = (int)((v - lo) * scale);
    b = b < 0 ? 0 : (b >= bins ? bins - 1 : b);
    // settle the candidate against the edges themselves (searchsorted semantics; the product above can be one bin off)
    const float e0 = lo + (hi - lo) * ((float)b / (float)bins), e1 = lo + (hi - lo) * ((float)(b + 1) / (float)bins);
    if (v < e0 && b > 0) b--;
    else if (v >= e1 && b + 1 < bins) b++;
    atomicAdd(&s_cnt[b], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += HNT)
    if (s_cnt[b]) atomicAdd(&counts[b], s_cnt[b]);
}

// Row sums of a row-major bf16 matrix, fp32 accumulation in a fixed order: the bias gradients of the GRU replay (sum over the T B
// columns of the feature-major gate gradients).  HBM-bound: one CTA per row, 128-bit loads, 2 bytes read per element.
constexpr int RS_NT = 256;
__global__ void __launch_bounds__(RS_NT) k_row_sums_bf16(int cols, const __nv_bfloat16* __restrict__ x, float* __restrict__ out) {
  const __nv_bfloat16* row = x + (size_t)blockIdx.x * cols;
  float acc = 0.0f;
  if ((cols & 7) == 0 && ((size_t)row & 15) == 0) {
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    for (int i = threadIdx.x; i < cols / 8; i += RS_NT) {
      const uint4 v = __ldcs(r4 + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; k++) acc += __uint_as_float(w[k] << 16) + __uint_as_float(w[k] & 0xffff0000u);
    }
  } else {
    for (int i = threadIdx.x; i < cols; i += RS_NT) acc += __bfloat162float(row[i]);
  }
  __shared__ float s_part[RS_NT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < RS_NT / 32; k++) t += s_part[k];
    out[blockIdx.x] = t;
  }
}

}  // namespace

extern "C" int b200sim_row_sums(void* stream, int rows, int cols, const void* x, float* out) {
  if (rows < 0 || cols < 0 || (rows > 0 && (!x || !out))) return pfail(B2S_ERR_BAD_ARG, "row_sums: null argument");
  if (rows == 0) return B2S_OK;
  k_row_sums_bf16<<<rows, RS_NT, 0, (cudaStream_t)stream>>>(cols, (const __nv_bfloat16*)x, out);
  LAUNCH_OK();
  return B2S_OK;
}

extern "C" size_t b200sim_histogram_scratch_bytes(void) { return (size_t)B2S_HIST_MAX_CTAS * 6 * sizeof(float); }

extern "C" int b200sim_histogram(void* stream, size_t n, const float* x, int bins, int32_t* counts, float* limits, float* stats, void* scratch) {
  if (bins < 1 || bins > HMAXBINS) return pfail(B2S_ERR_UNSUPPORTED, "histogram: 1 <= bins <= 1024");
  if (n == 0 || !x || !counts || !limits || !stats || !scratch) return pfail(B2S_ERR_BAD_ARG, "null argument or empty array");
  size_t want = (n + HNT - 1) / HNT;
  const int grid = (int)(want < (size_t)B2S_HIST_MAX_CTAS ? want : (size_t)B2S_HIST_MAX_CTAS);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, sizeof(int32_t) * bins, s) != cudaSuccess) return pfail(B2S_ERR_CUDA, "cudaMemsetAsync");
  k_hist_stats<<<grid, HNT, 0, s>>>(n, x, (float*)scratch);
  k_hist_count<<<grid, HNT, 0, s>>>(n, x, (const float*)scratch, grid, bins, counts, limits, stats);
  LAUNCH_OK();
  return B2S_OK;
}
