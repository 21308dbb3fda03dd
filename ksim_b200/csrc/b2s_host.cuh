// Host-side pieces shared by the translation units that launch the per-environment kernels (b200sim.cu: fused rollout /
// scoring; b2s_engine_api.cu: the PhysicsEngine-shaped entry points): the model handle, the shared-memory model view, launch
// geometry and the per-model-size dispatch.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "b2s_engine.cuh"

using namespace b2s;

// The model view the device code reads lives in shared memory: the kernel parameter is copied there once per CTA,
// together with as much of the model / task-program arrays as fits after the per-env workspaces (Mdl::stage), and
// the view's pointers are redirected to the copies.  Model constants then cost a shared-memory access instead of an
// L1-thrashed global load on the critical path of every stage.
template <class D>
__device__ __forceinline__ const Mdl& stage_model(const Mdl& mp, unsigned char* smem, int wpb) {
  Mdl* ms = reinterpret_cast<Mdl*>(smem + (((size_t)wpb * sizeof(Work<D>) + 15) & ~(size_t)15));
  int* dst = reinterpret_cast<int*>(ms + 1);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < (int)(sizeof(Mdl) / 4); i += nt) reinterpret_cast<int*>(ms)[i] = reinterpret_cast<const int*>(&mp)[i];
  const int* src[4] = {mp.mi, reinterpret_cast<const int*>(mp.mf), mp.ti, reinterpret_cast<const int*>(mp.tf)};
  int off = 0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int n = mp.stage[a];
    for (int i = tid; i < n; i += nt) dst[off + i] = src[a][i];
    off += n;
  }
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    if (mp.stage[0]) ms->mi = dst + o; o += mp.stage[0];
    if (mp.stage[1]) ms->mf = reinterpret_cast<const float*>(dst + o); o += mp.stage[1];
    if (mp.stage[2]) ms->ti = dst + o; o += mp.stage[2];
    if (mp.stage[3]) ms->tf = reinterpret_cast<const float*>(dst + o);
  }
  __syncthreads();
  return *ms;
}

static const size_t kSmemBudget = 227 * 1024;
// One CTA per SM; its size is what shared memory allows (Work<D> per warp), at most 16 warps.  Telling the compiler
// lets it use the registers a smaller CTA leaves free (K-Bot: 8 warps -> up to 255 registers per thread).
template <class D>
constexpr int max_warps() {

  return (int)((kSmemBudget - sizeof(Mdl) - 16) / sizeof(Work<D>)) > 16 ? 16 : (int)((kSmemBudget - sizeof(Mdl) - 16) / sizeof(Work<D>));
}

// Launch bound per instantiation (measured, gpurun_out/variants1.txt): the K-Bot kernel gains 20 % from the registers its
// 8-warp CTA leaves free (256 threads -> up to 255 registers); the humanoid kernel (14 warps) is 3 % faster when held
// to 128 registers (bound 512) than with the 144 its 448 threads would allow.
template <class D>
constexpr int launch_threads() { return D::SMALL ? 512 : 32 * max_warps<D>(); }

// -DB2S_POISON_SMEM (debug): a workspace starts as whatever the previous kernel left in shared memory; filling a byte range
// of it with 0xFF (NaN floats, -1 ints) makes a read of a field nobody wrote visible in the outputs (tools/debug_poison.py)
#ifdef B2S_POISON_SMEM
#define POISON_WORK(mp, w) do { unsigned* pw_ = reinterpret_cast<unsigned*>(&(w)); \
    for (int i_ = (mp).poison_lo / 4 + lane; i_ < (mp).poison_hi / 4 && i_ < (int)(sizeof(w) / 4); i_ += 32) pw_[i_] = 0xffffffffu; \
    __syncwarp(); } while (0)
#else
#define POISON_WORK(mp, w) do { } while (0)
#endif

struct b200sim_model {
  Mdl mdl;
  int variant;
  int warps_per_cta;
  size_t smem_per_env;
  size_t smem_bytes;   // dynamic shared memory per CTA: workspaces + model view + staged model arrays
  void* dev_blob;
  std::vector<int> ti_host;
};


#define DISPATCH(mo, CALL)                       \
  switch ((mo)->variant) {                       \
    case 0: { using D = DimsHumanoid; CALL; } break; \
    case 1: { using D = DimsKbot; CALL; } break;     \
    default: { using D = DimsGeneric; CALL; } break; \
  }

