// Host-side SERIAL EMULATION of the warp kernels -- TEST INFRASTRUCTURE ONLY.
//
// The device code in ksim_b200/csrc/*.cuh is written as lane-parallel phases (LANE_FOR / WSYNC).  Compiled
// with B2S_EMU those phases become plain loops, which lets GPU-less machines run the *same source* against
// the CPU oracle to catch logic errors before spending GPU time.  It is never imported by ksim_b200, is not a
// fallback of any product path, and says nothing about races (on the GPU those are covered by the bit-identical
// repeat / shard / step-vs-rollout checks of tests/test_gpu_parity.py::test_*full_size_properties; compute-sanitizer is
// not available on the GPU pool); only tests/ and tools/emu_accuracy.py build and load it.
#define B2S_EMU 1
#include <stdlib.h>
#include <string.h>

#include "../../ksim_b200/csrc/b2s_engine.cuh"
#include "../../ksim_b200/csrc/b2s_rewards.cuh"

using namespace b2s;
typedef DimsGeneric D;

// the workspace starts as whatever the previous kernel left in shared memory: B2S_EMU_POISON=1 fills it with 0xFF bytes
// (NaN floats, -1 ints) so that a read of a field nobody wrote shows up in the parity tests
static int poison_byte() { const char* e = getenv("B2S_EMU_POISON"); return (e && atoi(e)) ? 0xFF : 0; }

static Mdl view(const void* blob) {
  Mdl m;
  const int* h = (const int*)blob;
  memcpy(m.h, h, MH_SIZE * sizeof(int));
  m.mi = h + MH_SIZE;
  m.mf = (const float*)(h + MH_SIZE + h[MH_NI]);
  m.ti = h + MH_SIZE + h[MH_NI] + h[MH_NF];
  m.tf = (const float*)(h + MH_SIZE + h[MH_NI] + h[MH_NF] + h[MH_NTI]);
  m.sync_mask = 0xff;  // barriers / votes are no-ops on the host; the lockstep code paths still run
  if (const char* e = getenv("B2S_EMU_SYNC_MASK")) m.sync_mask = (int)strtol(e, nullptr, 0);  // e.g. 0x7f: the product default
  return m;
}

extern "C" int emu_init_envs(const void* blob, int n, const uint32_t* init_keys, const float* levels, const float* rand,
                             float* state, uint32_t* keys, float* rec0, uint32_t* act_keys) {
  Mdl m = view(blob);
  const int* ti = m.ti;
  const size_t S = ti[TI_STATE_SIZE], R = ti[TI_REC_SIZE], K = ti[TI_KEY_SIZE], RAND = ti[TI_RAND_SIZE];
#pragma omp parallel for schedule(static)
  for (int e = 0; e < n; e++) {
    Work<D>* w = new Work<D>();
    memset(w, poison_byte(), sizeof(*w));
    apply_overrides(m, *w, rand + e * RAND, 0);
    env_init(m, *w, init_keys + (size_t)e * 10, levels[e], rec0 + e * R, act_keys ? act_keys + 2 * e : nullptr, 0);
    env_store(m, *w, state + e * S, keys + e * K, 0);
    delete w;
  }
  return 0;
}

extern "C" int emu_step_ctrl(const void* blob, int n, const float* action, const float* rand, float* state, uint32_t* keys,
                             float* rec_cur, float* rec_next, uint32_t* act_keys) {
  Mdl m = view(blob);
  const int* ti = m.ti;
  const size_t S = ti[TI_STATE_SIZE], R = ti[TI_REC_SIZE], K = ti[TI_KEY_SIZE], RAND = ti[TI_RAND_SIZE], NA = ti[TI_ACTION_SIZE];
#pragma omp parallel for schedule(static)
  for (int e = 0; e < n; e++) {
    Work<D>* w = new Work<D>();
    memset(w, poison_byte(), sizeof(*w));
    apply_overrides(m, *w, rand + e * RAND, 0);
    env_load(m, *w, state + e * S, keys + e * K, 0);
    env_step(m, *w, action + e * NA, rec_cur + e * R, rec_next + e * R, act_keys ? act_keys + 2 * e : nullptr, 0);
    env_store(m, *w, state + e * S, keys + e * K, 0);
    delete w;
  }
  return 0;
}

extern "C" int emu_rewards(const void* blob, int n, int T, const float* rec, const float* levels, float* carry, float* total,
                           float* comps) {
  Mdl m = view(blob);
  const int* ti = m.ti;
  const size_t R = ti[TI_REC_SIZE], C = ti[TI_REW_CARRY_SIZE];
  for (int e = 0; e < n; e++)
    env_rewards(m, rec + e * R, (size_t)n * R, T, levels[e], carry + e * C, total + (size_t)e * T, comps + (size_t)e * T,
                (size_t)n * T, 0);
  return 0;
}

extern "C" int emu_sizeof_work(void) { return (int)sizeof(Work<DimsHumanoid>); }
extern "C" int emu_sizeof_work_kbot(void) { return (int)sizeof(Work<DimsKbot>); }
extern "C" int emu_sizeof_mdl(void) { return (int)sizeof(Mdl); }

// Debug aid for tests: forward kinematics of one configuration with the base model; xpos[nb][3], xquat[nb][4],
// xanchor[nj][3], xaxis[nj][3]; rand = one env's override block.
extern "C" int emu_kinematics(const void* blob, const float* rand, const float* qpos, float* xpos, float* xquat, float* xanchor, float* xaxis) {
  Mdl m = view(blob);
  Work<D>* w = new Work<D>();
  memset(w, poison_byte(), sizeof(*w));
  apply_overrides(m, *w, rand, 0);
  for (int i = 0; i < m.h[MH_NQ]; i++) w->qpos[i] = qpos[i];
  kinematics(m, *w, 0);
  for (int b = 0; b < m.h[MH_NBODY]; b++) {
    for (int k = 0; k < 3; k++) xpos[3 * b + k] = w->xpos[k][b];
    for (int k = 0; k < 4; k++) xquat[4 * b + k] = w->xquat[k][b];
  }
  for (int j = 0; j < m.h[MH_NJNT]; j++)
    for (int k = 0; k < 3; k++) { xanchor[3 * j + k] = w->xanchor[k][j]; xaxis[3 * j + k] = w->xaxis[k][j]; }
  delete w;
  return 0;
}
