// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, SS) for the shapes / shared-memory layouts the GRU
// kernels use.  One CTA per SM (grid = 1 or 148), timing only (operands are whatever is in shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ksim_b200/csrc/b2s_tc.cuh"
using namespace b2s_tc;

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

// the same chain with the other warps of the GRU kernel emulated: mode 1 = one lane per warp polls an mbarrier (try_wait) that
// completes when the chain is done, mode 2 = all lanes poll, mode 3 = the warps stream global stores, mode 4 = global loads
__global__ void __launch_bounds__(288, 1) k_gru_busy(int reps, int mode, float* gbuf, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, done;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); stop = 0; mbar_fence_init(); }
  if (warp == 8) tmem_alloc(&s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int N = 96;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t b0 = smem_u32(smem), a0 = smem_u32(smem + 98304);
      long long t0 = clock64();
      for (int r = 0; r < reps; r++) {
        for (int ch = 0; ch < 16; ch++)
          for (int k2 = 0; k2 < 2; k2++)
            mma_bf16_ss(tmem, mk_desc(a0 + ch * 8192 + k2 * 4096, 2048, 128, 0), mk_desc(b0 + (ch * 4 + k2 * 2) * (N * 16), N * 16, 128, 0), idesc,
                        (ch | k2) != 0);
        mma_commit(&bar);
        mbar_wait(&bar, r & 1, 1u << 24);
      }
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = 0; out[1] = t2 - t0; }
      stop = 1;
      mbar_arrive(&done);
    }
  } else {
    if (mode == 1) { if (lane == 0) mbar_wait(&done, 0, 1u << 26); }
    else if (mode == 2) mbar_wait(&done, 0, 1u << 26);
    else if (mode == 3) { float* q = gbuf + ((size_t)blockIdx.x * 256 + threadIdx.x) * 4; size_t i = 0; while (!stop) { __stcs(reinterpret_cast<float4*>(q + (i & 63) * 148 * 1024), make_float4(1, 2, 3, 4)); i++; } }
    else if (mode == 4) { const float* q = gbuf + ((size_t)blockIdx.x * 256 + threadIdx.x) * 4; size_t i = 0; float acc = 0; while (!stop) { float4 v = __ldcs(reinterpret_cast<const float4*>(q + (i & 63) * 148 * 1024)); acc += v.x; i++; } if (acc == 1234.5f) out[3] = 1; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// exactly the forward GRU kernel's per-step MMA chain: B (weights) at the base of shared memory, A (state tile) behind it
__global__ void __launch_bounds__(128, 1) k_gru_like(int reps, int N, int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint32_t b0 = smem_u32(smem), a0 = smem_u32(smem + 98304);
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      for (int ch = 0; ch < 16; ch++)
        for (int k2 = 0; k2 < 2; k2++) {
          uint32_t a = a0 + ch * 8192 + k2 * 2 * 2048, b = b0 + (ch * 4 + k2 * 2) * (N * 16);
          if (variant == 1) b = b0;                       // B fixed
          if (variant == 2) a = a0;                       // A fixed
          mma_bf16_ss(tmem, mk_desc(a, 2048, 128, 0), mk_desc(b, N * 16, 128, 0), idesc, (ch | k2) != 0);
        }
      mma_commit(&bar);
      mbar_wait(&bar, r & 1, 1u << 24);
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = 0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

__global__ void __launch_bounds__(128, 1) k(int n_mma, int N, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, uint32_t layout,
                                            uint32_t kstep_bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 131072);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i++) {
      const uint32_t off = (i % 16) * kstep_bytes;
      mma_bf16_ss(tmem, mk_desc(a0 + off, lbo_a, sbo_a, layout), mk_desc(b0 + off, lbo_b, sbo_b, layout), idesc, i != 0);
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0, 1u << 24);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 229376);
  struct Cfg { const char* name; int N; uint32_t lbo_a, sbo_a, lbo_b, sbo_b, layout, kstep; };
  const Cfg cfgs[] = {
      {"no-swizzle  M128 N96  (GRU fwd today: LBO 2048/1536, SBO 128)", 96, 2048, 128, 1536, 128, 0, 4096},
      {"no-swizzle  M128 N32  (GRU bwd today)", 32, 2048, 128, 512, 128, 0, 4096},
      {"no-swizzle  M128 N128", 128, 2048, 128, 2048, 128, 0, 4096},
      {"no-swizzle  M128 N256", 256, 2048, 128, 4096, 128, 0, 4096},
      {"swizzle128B M128 N96  (SBO 1024)", 96, 16, 1024, 16, 1024, 2, 32},
      {"swizzle128B M128 N32", 32, 16, 1024, 16, 1024, 2, 32},
      {"swizzle128B M128 N128", 128, 16, 1024, 16, 1024, 2, 32},
      {"swizzle128B M128 N256", 256, 16, 1024, 16, 1024, 2, 32},
      {"swizzle64B  M128 N96  (SBO 512)", 96, 16, 512, 16, 512, 4, 32},
      {"swizzle64B  M128 N32", 32, 16, 512, 16, 512, 4, 32},
      {"swizzle32B  M128 N96  (SBO 256)", 96, 16, 256, 16, 256, 6, 32},
  };
  cudaFuncSetAttribute(k_gru_like, cudaFuncAttributeMaxDynamicSharedMemorySize, 229376);
  for (int N : {96, 32})
    for (int variant = 0; variant < 3; variant++) {
      for (int rep = 0; rep < 2; rep++) {
        k_gru_like<<<128, 128, 229376>>>(8, N, variant, out);
        cudaDeviceSynchronize();
      }
      printf("gru-like chain N=%d variant %d (0 = as the kernel, 1 = B fixed, 2 = A fixed): %6.1f cyc/MMA incl. commit+wait per 32\n", N, variant,
             out[1] / 256.0);
    }
  cudaFuncSetAttribute(k_gru_busy, cudaFuncAttributeMaxDynamicSharedMemorySize, 229376);
  float* gbuf;
  cudaMalloc(&gbuf, (size_t)64 * 148 * 1024 * 4 + 148 * 256 * 16);
  for (int mode = 0; mode < 5; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      k_gru_busy<<<128, 288, 229376>>>(8, mode, gbuf, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("busy mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    printf("gru chain with busy neighbours, mode %d (0 idle, 1 lane-0 pollers, 2 all-lane pollers, 3 streaming stores, 4 streaming loads): %6.1f cyc/MMA\n",
           mode, out[1] / 256.0);
  }
  for (int grid : {1}) {
    for (const Cfg& c : cfgs) {
      for (int rep = 0; rep < 2; rep++) {
        k<<<grid, 128, 229376>>>(256, c.N, c.lbo_a, c.sbo_a, c.lbo_b, c.sbo_b, c.layout, c.kstep, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      }
      printf("grid %3d  %-66s issue %6.1f cyc/MMA   complete %6.1f cyc/MMA   (floor %d)\n", grid, c.name, out[0] / 256.0, out[1] / 256.0,
             128 * c.N / 256);
    }
  }
  return 0;
}
