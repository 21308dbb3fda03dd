// PhysicsEngine-shaped entry points: what ksim's `PhysicsEngine.reset` / `.step` are (ksim/engine.py:171-199,
// 205-361) WITHOUT the task layer the fused control step adds -- actuators + events + the physics sub-steps only -- returning
// the `physics_state.data` fields the reference's plugins read (qpos, qvel, qacc, ctrl, time, xpos, xquat, cinert,
// cvel, actuator_force, sensordata, cfrc_ext, site_xpos / site_xmat, subtree_com, xfrc_applied, ncon, contact.{geom1, geom2,
// dist, pos}) as one flat fp32 block per environment, so that a user's own Observation / Reward / Termination written against
// `physics_state.data.*` runs on views of that block.  Also the task program's terminations evaluated on such a block
// (`get_terminations`, ksim/task/rl.py:288-299).  Same device code as the fused rollout (engine_step / engine_reset /
// termination of b2s_engine.cuh): one environment per warp, workspace in shared memory.
#include <cuda_runtime.h>

#include <string>

#include "../../include/b200sim.h"
#include "b2s_host.cuh"

namespace {

struct DataLayout {
  int off[B2S_DATA_N_FIELDS + 1];  // float offset of every field inside an environment's block; [N_FIELDS] = block size
};

DataLayout layout_of(const int* h) {
  const int nq = h[MH_NQ], nv = h[MH_NV], nu = h[MH_NU], nb = h[MH_NBODY], ns = h[MH_NSITE], nsd = h[MH_NSENSORDATA], nc = h[MH_NCON];
  const int size[B2S_DATA_N_FIELDS] = {nq, nv, nv, nu, 1, 3 * nb, 4 * nb, 10 * nb, 6 * nb, nu, nsd, 6 * nb, 3 * ns, 9 * ns, 3 * nb, 6 * nb,
                                       1, nc, nc, nc, 3 * nc};
  DataLayout l;
  int o = 0;
  for (int i = 0; i < B2S_DATA_N_FIELDS; i++) { l.off[i] = o; o += size[i]; }
  l.off[B2S_DATA_N_FIELDS] = o;
  return l;
}

template <class D>
__device__ __forceinline__ void write_data(const Mdl& m, const Work<D>& w, const DataLayout& l, float* __restrict__ out, const int lane) {
  const int nq = DIM(NQ, MH_NQ), nv = DIM(NV, MH_NV), nu = DIM(NU, MH_NU), nb = DIM(NB, MH_NBODY), ns = DIM(NS, MH_NSITE);
  const int nsd = m.h[MH_NSENSORDATA], nc = DIM(NCON, MH_NCON);
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  const int* pg2 = MI(m, MH_PAIR_GEOM2);
  LANE_FOR(i, nq) out[l.off[B2S_DATA_QPOS] + i] = w.qpos[i];
  LANE_FOR(i, nv) { out[l.off[B2S_DATA_QVEL] + i] = w.qvel[i]; out[l.off[B2S_DATA_QACC] + i] = w.qacc[i]; }
  LANE_FOR(i, nu) { out[l.off[B2S_DATA_CTRL] + i] = w.ctrl[i]; out[l.off[B2S_DATA_ACTUATOR_FORCE] + i] = w.actuator_force[i]; }
  LANE_FOR(i, 3 * nb) {
    out[l.off[B2S_DATA_XPOS] + i] = w.xpos[i % 3][i / 3];
    if (i >= 3) out[l.off[B2S_DATA_SUBTREE_COM] + i] = w.subtree_com[i % 3][i / 3];
  }
  LANE_FOR(i, 4 * nb) out[l.off[B2S_DATA_XQUAT] + i] = w.xquat[i & 3][i >> 2];
  LANE_FOR(i, 10 * nb) out[l.off[B2S_DATA_CINERT] + i] = w.cinert[i % 10][i / 10];
  LANE_FOR(i, 6 * nb) {
    out[l.off[B2S_DATA_CVEL] + i] = w.cvel[i % 6][i / 6];
    out[l.off[B2S_DATA_CFRC_EXT] + i] = w.cfrc_ext[i % 6][i / 6];
    out[l.off[B2S_DATA_XFRC_APPLIED] + i] = (i / 6 == w.xfrc_body) ? w.xfrc[i % 6] : 0.0f;
  }
  LANE_FOR(i, nsd) out[l.off[B2S_DATA_SENSORDATA] + i] = w.sensordata[i];
  LANE_FOR(i, 3 * ns) out[l.off[B2S_DATA_SITE_XPOS] + i] = w.site_xpos[i % 3][i / 3];
  LANE_FOR(i, 9 * ns) out[l.off[B2S_DATA_SITE_XMAT] + i] = w.site_xmat[i % 9][i / 9];
  LANE_FOR(c, nc) {
    out[l.off[B2S_DATA_CONTACT_GEOM1] + c] = (float)pg1[w.con_pair[c]];
    out[l.off[B2S_DATA_CONTACT_GEOM2] + c] = (float)pg2[w.con_pair[c]];
    out[l.off[B2S_DATA_CONTACT_DIST] + c] = w.con_dist[c];
    out[l.off[B2S_DATA_CONTACT_POS] + 3 * c] = w.con_pos[0][c];
    out[l.off[B2S_DATA_CONTACT_POS] + 3 * c + 1] = w.con_pos[1][c];
    out[l.off[B2S_DATA_CONTACT_POS] + 3 * c + 2] = w.con_pos[2][c];
  }
  if (lane == 0) {
    out[l.off[B2S_DATA_TIME]] = w.time;
    out[l.off[B2S_DATA_NCON]] = (float)nc;
    // subtree_com[0] (the world's subtree = the whole model): the kernels never need it and leave it unaccumulated
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    for (int b = 1; b < nb; b++) { cx += w.body_mass[b] * w.xipos[0][b]; cy += w.body_mass[b] * w.xipos[1][b]; cz += w.body_mass[b] * w.xipos[2][b]; }
    const float sm = fmaxf(MF(m, MH_BODY_SUBTREEMASS)[0], MINVAL);  // nominal subtree mass, as for the other bodies
    out[l.off[B2S_DATA_SUBTREE_COM]] = cx / sm; out[l.off[B2S_DATA_SUBTREE_COM] + 1] = cy / sm; out[l.off[B2S_DATA_SUBTREE_COM] + 2] = cz / sm;
  }
}

// engine.step (ksim/engine.py:259-361): action -> ctrl (latency blend, drop mask), actuators, events, nsub x mjx.step.
template <class D>
__global__ void __launch_bounds__(launch_threads<D>()) k_engine_step(const Mdl mp, int n_envs, const float* __restrict__ action,
                                                                     const uint32_t* __restrict__ rng, const float* __restrict__ rand,
                                                                     float* __restrict__ state, uint32_t* __restrict__ keys,
                                                                     float* __restrict__ data, const DataLayout l) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int env = blockIdx.x * wpb + warp;
  const Mdl& m = stage_model<D>(mp, smem, wpb);
  const int* ti = m.ti;
  if (env >= n_envs) {
    for (int i = 0; i < ti[TI_NSUB]; i++) {  // idle warps of the last CTA still take part in the per-sub-step CTA barriers
      if ((m.sync_mask >> 6) & 1) CTA_SYNC();
      forward_follower(m);
    }
    return;
  }
  Work<D>& w = reinterpret_cast<Work<D>*>(smem)[warp];
  const size_t S = ti[TI_STATE_SIZE], K = ti[TI_KEY_SIZE], RAND = ti[TI_RAND_SIZE], NA = ti[TI_ACTION_SIZE];
  apply_overrides(m, w, rand + env * RAND, lane);
  env_load(m, w, state + env * S, keys + env * K, lane);
  const float aclip = (m.tf + ti[TI_ENGINE_FOFF])[6];
  LANE_FOR(i, (int)NA) {
    float a = action[env * NA + i];
    if (a != a) a = 0.0f;  // the caller's NaN cleaning + clip (rl.py:998-1001), as in the fused step
    w.action[i] = clipf(a, -aclip, aclip);
  }
  WSYNC();
  Key k; k.k0 = rng[2 * env]; k.k1 = rng[2 * env + 1];
  engine_step(m, w, k, lane);
  write_data(m, w, l, data + (size_t)env * l.off[B2S_DATA_N_FIELDS], lane);
  env_store(m, w, state + env * S, keys + env * K, lane);
}

// engine.reset (ksim/engine.py:205-246): mjx.make_data, the Reset plugins, mjx.forward, zeroed velocities, fresh actuator /
// event states, latency and zero offsets.  Task-side state (commands, observation carries, keys) is zeroed.
template <class D>
__global__ void __launch_bounds__(launch_threads<D>()) k_engine_reset(const Mdl mp, int n_envs, const uint32_t* __restrict__ rng,
                                                                      const float* __restrict__ levels, const float* __restrict__ rand,
                                                                      float* __restrict__ state, uint32_t* __restrict__ keys,
                                                                      float* __restrict__ data, const DataLayout l) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int env = blockIdx.x * wpb + warp;
  const Mdl& m = stage_model<D>(mp, smem, wpb);
  if (env >= n_envs) return;
  Work<D>& w = reinterpret_cast<Work<D>*>(smem)[warp];
  const int* ti = m.ti;
  const size_t S = ti[TI_STATE_SIZE], K = ti[TI_KEY_SIZE], RAND = ti[TI_RAND_SIZE];
  apply_overrides(m, w, rand + env * RAND, lane);
  IF_LANE0 { w.level = levels[env]; w.nonfinite = 0.0f; }
  LANE_FOR(i, ti[TI_EVENT_SIZE]) w.event[i] = 0.0f;
  LANE_FOR(i, ti[TI_ACT_STATE_SIZE]) w.act_state[i] = 0.0f;
  LANE_FOR(i, ti[TI_CMD_SIZE]) w.cmd[i] = 0.0f;
  LANE_FOR(i, ti[TI_OBS_CARRY_SIZE]) w.obs_carry[i] = 0.0f;
  LANE_FOR(i, ti[TI_KEY_SIZE]) w.keys[i] = 0u;
  WSYNC();
  Key k; k.k0 = rng[2 * env]; k.k1 = rng[2 * env + 1];
  engine_reset(m, w, k, lane);
  write_data(m, w, l, data + (size_t)env * l.off[B2S_DATA_N_FIELDS], lane);
  env_store(m, w, state + env * S, keys + env * K, lane);
}

// get_terminations (rl.py:288-299) + the done / success reduction (rl.py:1049-1050) on data blocks.
template <class D>
__global__ void __launch_bounds__(launch_threads<D>()) k_terminations(const Mdl mp, int n_envs, const float* __restrict__ data,
                                                                      const DataLayout l, const float* __restrict__ levels,
                                                                      int32_t* __restrict__ codes, uint8_t* __restrict__ done,
                                                                      uint8_t* __restrict__ success) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int env = blockIdx.x * wpb + warp;
  const Mdl& m = stage_model<D>(mp, smem, wpb);
  if (env >= n_envs) return;
  Work<D>& w = reinterpret_cast<Work<D>*>(smem)[warp];
  const int* ti = m.ti;
  const float* in = data + (size_t)env * l.off[B2S_DATA_N_FIELDS];
  const int nq = DIM(NQ, MH_NQ), nv = DIM(NV, MH_NV), nb = DIM(NB, MH_NBODY), nc = DIM(NCON, MH_NCON), npair = m.h[MH_NPAIR];
  const int* pg1 = MI(m, MH_PAIR_GEOM1);
  const int* pg2 = MI(m, MH_PAIR_GEOM2);
  LANE_FOR(i, nq) w.qpos[i] = in[l.off[B2S_DATA_QPOS] + i];
  LANE_FOR(i, nv) w.qvel[i] = in[l.off[B2S_DATA_QVEL] + i];
  LANE_FOR(i, 3 * nb) w.xpos[i % 3][i / 3] = in[l.off[B2S_DATA_XPOS] + i];
  LANE_FOR(c, nc) {
    // a contact is identified by its geom pair; the device code works with indices into the model's pair list
    const int g1 = (int)in[l.off[B2S_DATA_CONTACT_GEOM1] + c], g2 = (int)in[l.off[B2S_DATA_CONTACT_GEOM2] + c];
    int pair = -1;
    for (int p = 0; p < npair; p++)
      if ((pg1[p] == g1 && pg2[p] == g2) || (pg1[p] == g2 && pg2[p] == g1)) { pair = p; break; }
    w.con_pair[c] = pair < 0 ? 0 : pair;
    w.con_dist[c] = pair < 0 ? 1.0f : in[l.off[B2S_DATA_CONTACT_DIST] + c];  // unknown pair: not in contact
  }
  IF_LANE0 { w.time = in[l.off[B2S_DATA_TIME]]; w.level = levels[env]; }
  WSYNC();
  const int nterm = ti[TI_N_TERM];
  int any = 0, all_not_fail = 1;
  for (int t = 0; t < nterm; t++) {
    const int v = termination(m, w, t);
    IF_LANE0 codes[(size_t)env * nterm + t] = v;
    any |= (v != 0);
    all_not_fail &= (v != -1);
  }
  IF_LANE0 {
    if (done) done[env] = (uint8_t)any;
    if (success) success[env] = (uint8_t)(all_not_fail && any);
  }
}

thread_local std::string g_api_error;
int afail(int code, const std::string& msg) { g_api_error = msg; return code; }
#define API_CUDA_OK(expr)                                                                                  \
  do {                                                                                                     \
    cudaError_t e_ = (expr);                                                                               \
    if (e_ != cudaSuccess) return afail(B2S_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <class D>
int prepare(const b200sim_model* mo) {
  static thread_local size_t configured = 0;  // per instantiation: the largest dynamic shared-memory size set so far
  if (configured < mo->smem_bytes) {
    API_CUDA_OK(cudaFuncSetAttribute(k_engine_step<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mo->smem_bytes));
    API_CUDA_OK(cudaFuncSetAttribute(k_engine_reset<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mo->smem_bytes));
    API_CUDA_OK(cudaFuncSetAttribute(k_terminations<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mo->smem_bytes));
    configured = mo->smem_bytes;
  }
  return B2S_OK;
}

const int* header_of(const b200sim_model* mo) { return mo->mdl.h; }

}  // namespace

extern "C" const char* b200sim_engine_last_error(void) { return g_api_error.c_str(); }

extern "C" int b200sim_data_layout(const b200sim_model* mo, int32_t* offsets) {
  if (!mo || !offsets) return afail(B2S_ERR_BAD_ARG, "null argument");
  const DataLayout l = layout_of(header_of(mo));
  for (int i = 0; i <= B2S_DATA_N_FIELDS; i++) offsets[i] = l.off[i];
  return B2S_OK;
}

extern "C" int b200sim_engine_step(const b200sim_model* mo, void* stream, int n_envs, const float* action, const uint32_t* rng,
                                   const float* rand, float* state, uint32_t* keys, float* data) {
  if (!mo || n_envs < 0 || !action || !rng || !state || !keys || !data) return afail(B2S_ERR_BAD_ARG, "null argument");
  if (n_envs == 0) return B2S_OK;
  if (!rand && mo->ti_host[TI_RAND_SIZE] > 0) return afail(B2S_ERR_BAD_ARG, "rand is required for this task");
  const DataLayout l = layout_of(header_of(mo));
  const int wpb = mo->warps_per_cta;
  const dim3 grid((n_envs + wpb - 1) / wpb), block(wpb * 32);
  int rc = B2S_OK;
  DISPATCH(mo, (rc = prepare<D>(mo)));
  if (rc) return rc;
  DISPATCH(mo, (k_engine_step<D><<<grid, block, mo->smem_bytes, (cudaStream_t)stream>>>(mo->mdl, n_envs, action, rng, rand, state, keys, data, l)));
  API_CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

extern "C" int b200sim_engine_reset(const b200sim_model* mo, void* stream, int n_envs, const uint32_t* rng, const float* levels,
                                    const float* rand, float* state, uint32_t* keys, float* data) {
  if (!mo || n_envs < 0 || !rng || !levels || !state || !keys || !data) return afail(B2S_ERR_BAD_ARG, "null argument");
  if (n_envs == 0) return B2S_OK;
  if (!rand && mo->ti_host[TI_RAND_SIZE] > 0) return afail(B2S_ERR_BAD_ARG, "rand is required for this task");
  const DataLayout l = layout_of(header_of(mo));
  const int wpb = mo->warps_per_cta;
  const dim3 grid((n_envs + wpb - 1) / wpb), block(wpb * 32);
  int rc = B2S_OK;
  DISPATCH(mo, (rc = prepare<D>(mo)));
  if (rc) return rc;
  DISPATCH(mo, (k_engine_reset<D><<<grid, block, mo->smem_bytes, (cudaStream_t)stream>>>(mo->mdl, n_envs, rng, levels, rand, state, keys, data, l)));
  API_CUDA_OK(cudaGetLastError());
  return B2S_OK;
}

extern "C" int b200sim_terminations(const b200sim_model* mo, void* stream, int n_envs, const float* data, const float* levels,
                                    int32_t* codes, uint8_t* done, uint8_t* success) {
  if (!mo || n_envs < 0 || !data || !levels || !codes) return afail(B2S_ERR_BAD_ARG, "null argument");
  if (n_envs == 0 || mo->ti_host[TI_N_TERM] == 0) return B2S_OK;
  const DataLayout l = layout_of(header_of(mo));
  const int wpb = mo->warps_per_cta;
  const dim3 grid((n_envs + wpb - 1) / wpb), block(wpb * 32);
  int rc = B2S_OK;
  DISPATCH(mo, (rc = prepare<D>(mo)));
  if (rc) return rc;
  DISPATCH(mo, (k_terminations<D><<<grid, block, mo->smem_bytes, (cudaStream_t)stream>>>(mo->mdl, n_envs, data, l, levels, codes, done, success)));
  API_CUDA_OK(cudaGetLastError());
  return B2S_OK;
}
