/*
 * b200sim C-ABI: the drop-in boundary of the B200 batched-rollout engine.
 *
 * Every entry point takes an opaque model handle, a CUDA stream (as void*; NULL = default stream) and flat
 * DEVICE buffers owned by the caller.  No host synchronisation, no exceptions: the return value is a B2S_*
 * error code (include/b200sim_codes.h; 0 = OK).  This is the shape an XLA-FFI / jax.ffi handler binds
 * (INTEGRATION.md shows the stub); the repo's own host code binds it with ctypes (ksim_b200/binding.py).
 *
 * Per-environment blocks are contiguous and 16-byte aligned (sizes from b200sim_model_info):
 *   state[N][S] f32, keys[N][K] u32, rand[N][RAND] f32 (per-env model overrides), action[N][NA] f32,
 *   rec[T(+1)][N][R] f32 (trajectory records: [pre-step block: obs, command, level | post-step block]).
 *
 * What each entry point replaces in the reference (kscalelabs/ksim):
 *   b200sim_init_envs   RLTask._get_env_state -> apply_randomizations -> MjxEngine.reset + initial commands /
 *                       obs carry / obs bias + first get_observation
 *                       (ksim/task/rl.py:2175-2313, 361-372; ksim/engine.py:205-246)
 *   b200sim_step_ctrl   RLTask.step_engine for all envs: MjxEngine.step (5 x mjx.step with actuators + events),
 *                       get_terminations, Trajectory record, reset-on-done, commands, next get_observation
 *                       (ksim/task/rl.py:1164-1211, 1023-1136; ksim/engine.py:259-361)
 *   b200sim_rollout     RLTask._single_unroll's scan over T of the above with given actions (rl.py:1604-1665)
 *   b200sim_rewards     get_rewards vmapped over envs (ksim/task/rl.py:189-245, 1663; ksim/rewards.py)
 *   b200sim_gae         compute_ppo_inputs (ksim/task/ppo.py:54-121)
 *   b200sim_score       the three above fused: rewards + done / success flags + compute_ppo_inputs in one launch
 *   b200sim_ppo_loss    compute_ppo_loss + the mean over (B, T) of the summed terms, and that mean's gradient with respect to
 *                       the new policy's log-probs / values / entropy (ksim/task/ppo.py:149-246, 471-488; the backward
 *                       pass jax.grad derives from it at ppo.py:521-534)
 */
#ifndef B200SIM_H
#define B200SIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200sim_model b200sim_model;

/* b200sim_model_info selectors */
enum {
  B2S_INFO_STATE_SIZE = 0, /* S  */
  B2S_INFO_REC_SIZE = 1,   /* R  */
  B2S_INFO_KEY_SIZE = 2,   /* K  */
  B2S_INFO_RAND_SIZE = 3,  /* RAND */
  B2S_INFO_ACTION_SIZE = 4,
  B2S_INFO_N_REW_COMP = 5,
  B2S_INFO_REW_CARRY_SIZE = 6,
  B2S_INFO_SMEM_PER_ENV = 7,    /* bytes of shared memory one environment (warp) occupies */
  B2S_INFO_WARPS_PER_CTA = 8,
  B2S_INFO_VARIANT = 9,         /* 0 humanoid-sized, 1 kbot-sized, 2 generic */
  B2S_INFO_STAGED_ARRAYS = 10   /* how many of the 4 model / task-program arrays the kernels keep in shared memory */
};

/* blob: include/b200sim_model.h layout (host memory).  Uploads the model + task program to the current device. */
int b200sim_model_create(const void* blob, size_t nbytes, b200sim_model** out);
void b200sim_model_destroy(b200sim_model* model);
int b200sim_model_info(const b200sim_model* model, int what);
/* Options of a model handle (B2S_OPT_* / values in include/b200sim_codes.h); affects launches issued after the call. */
int b200sim_model_set_option(b200sim_model* model, int option, int value);
const char* b200sim_last_error(void);

/* init_keys[N][10] u32: (reset, command, obs-carry, obs-bias, env) keys; levels[N]; rec0[N][R] receives the first
 * pre-step block; act_keys[N][2] (may be NULL) receives the key the policy samples its first action with. */
int b200sim_init_envs(const b200sim_model* model, void* stream, int n_envs, const uint32_t* init_keys,
                      const float* levels, const float* rand, float* state, uint32_t* keys, float* rec0,
                      uint32_t* act_keys);

/* one control step for every env; rec_cur gets the post-step block, rec_next the next pre-step block. */
int b200sim_step_ctrl(const b200sim_model* model, void* stream, int n_envs, const float* action, const float* rand,
                      float* state, uint32_t* keys, float* rec_cur, float* rec_next, uint32_t* act_keys);

/* T control steps with given actions[T][N][NA]; rec[T+1][N][R] (slot 0's pre-step block must be valid on entry,
 * slot T's pre-step block is valid on exit).  Environment state stays on chip for the whole rollout. */
int b200sim_rollout(const b200sim_model* model, void* stream, int n_envs, int n_steps, const float* actions,
                    const float* rand, float* state, uint32_t* keys, float* rec);

/* rec[T][N][R]; levels[N]; carry[N][C] in/out; total[N][T]; comps[K][N][T]. */
int b200sim_rewards(const b200sim_model* model, void* stream, int n_envs, int n_steps, const float* rec,
                    const float* levels, float* carry, float* total, float* comps);

/* The scoring pass of a rollout in one launch: get_rewards (ksim/task/rl.py:189-245, 1663), the done / success flags of the
 * records, and -- when `values` is given -- compute_ppo_inputs on them (ksim/task/ppo.py:54-121; what the three calls
 * b200sim_rewards, b200sim_extract_flags, b200sim_gae produce).  rec[T][N][R]; levels[N]; carry[N][C] in/out; values[N][T] or
 * NULL; total[N][T]; comps[K][N][T]; dones / successes [N][T] bytes (either may be NULL); adv / targets / returns [N][T]
 * (required with values).  flags as for b200sim_gae. */
int b200sim_score(const b200sim_model* model, void* stream, int n_envs, int n_steps, const float* rec, const float* levels,
                  float* carry, const float* values, float gamma, float lam, int flags, float* total, float* comps,
                  uint8_t* dones, uint8_t* successes, float* adv, float* targets, float* returns);

/* all arrays [B][T]; flags: GAE_NORMALIZE_ADVANTAGES | GAE_MONTE_CARLO_RETURNS (include/b200sim_codes.h). */
int b200sim_gae(void* stream, int batch, int n_steps, const float* values, const float* rewards,
                const uint8_t* dones, const uint8_t* successes, float gamma, float lam, int flags, float* adv,
                float* targets, float* returns);

/* PPO loss, forward and backward in one pass (the loss half).  Device arrays, fp32:
 *   logp_on / logp_off / entropy [B][T][A] (entropy may be NULL: no entropy term, ppo.py:236-238);
 *   values_on / values_off / adv / targets [B][T]  (adv, targets: b200sim_gae's outputs, constants of the loss, ppo.py:121).
 * Terms per (b, t) (ppo.py:205-244): policy = -min(r A, clip(r, 1 +- clip_param) A), r = exp(clip(sum_a(logp_off - logp_on),
 * +-log_clip_value)); value = value_loss_coef * 0.5 * (clipped or plain squared error; flags & PPO_CLIPPED_VALUE_LOSS);
 * kl = kl_coef * kl_scale * sum_a(logp_on - logp_off) (kl_scale: the adaptive-KL factor of ppo.py:471-473);
 * entropy = -entropy_coef * sum_a entropy.
 * Outputs: losses[PPO_N_TERMS][B][T]; out[1 + PPO_N_TERMS] = mean over (B, T) of the summed terms, then of each term
 * (deterministic: fixed grid, fp64 partial sums in `scratch`, PPO_SCRATCH_BYTES of device memory);
 * d_logp[B][T] = d mean / d logp_off[b][t][a] (the same for every a); d_values[B][T] = d mean / d values_off;
 * d mean / d entropy[b][t][a] is the constant -entropy_coef / (B T) and is not materialised.
 * Any of losses, d_logp, d_values may be NULL (not wanted). */
int b200sim_ppo_loss(void* stream, int batch, int n_steps, int n_act, const float* logp_on, const float* logp_off,
                     const float* values_on, const float* values_off, const float* adv, const float* targets,
                     const float* entropy, float clip_param, float value_loss_coef, float entropy_coef, float kl_coef,
                     float log_clip_value, float kl_scale, int flags, float* losses, float* out, float* d_logp,
                     float* d_values, void* scratch);

/* done / success flags of a record buffer as bytes: out[N][T] from rec[T][N][R] (feeds b200sim_gae). */
int b200sim_extract_flags(const b200sim_model* model, void* stream, int n_envs, int n_steps, const float* rec,
                          uint8_t* dones, uint8_t* successes);

/* Per-iteration reductions over the (N, T) trajectory that the curricula and the logged metrics consume
 * (Trajectory.episode_length types.py:72-76; get_termination_metrics / get_reward_metrics rl.py:1544-1573;
 * curriculum.py:125,187,262).  per_env[N][B2S_METRICS_FIXED + 8 + 1 + K] scratch/out:
 *   [0] episode_length, [1] deaths = sum(done), [2] max |qpos[:3]|, [3] steps with any termination,
 *   [4..12) sum of each termination component's codes (up to 8), [12] sum of the total reward, [13..13+K) component sums.
 * out[4 + 8 + 1 + K]: [0] mean episode length, [1] mean deaths (mean_terminations), [2] max distance from the origin,
 *   [3] num_terminations (clipped to >= 1), [4..12) prct/<termination> = code sum / num_terminations,
 *   [12] mean total reward per (env, step), [13..) mean of each reward component.  Deterministic (fixed reduction order). */
#define B2S_METRICS_FIXED 13
int b200sim_metrics(const b200sim_model* model, void* stream, int n_envs, int n_steps, const float* rec, const float* total,
                    const float* comps, float* per_env, float* out);

/* TrajectoryDatasetWriter.write (ksim/dataset.py:59-103) for every env at once: out[N][sample_size] fp32, one flattened
 * trajectory sample per env, fields in the order of the table.  fields[n_fields][4] int32 in DEVICE memory:
 * {source, offset, dim, out_offset}; source 0 = record (float offset inside a record, emits [T][dim]), 1 = reward total
 * (emits [T]), 2 = reward component number `offset` (emits [T]).  rec[T][N][R]; total[N][T]; comps[K][N][T]. */
int b200sim_dataset_pack(void* stream, int n_envs, int n_steps, int rec_size, const float* rec, const float* total,
                         const float* comps, const int32_t* fields, int n_fields, int sample_size, float* out);

/* ---- The engine alone: ksim's PhysicsEngine.reset / .step (ksim/engine.py:171-199; MjxEngine :205-361) without the task layer
 * the fused control step adds -- what RLTask._step_physics (rl.py:1143-1157) and the reset path (rl.py:1072-1080) call.  `rng`
 * [N][2] u32 is the key the caller hands to engine.step / engine.reset for each environment; `state` / `keys` are the same
 * opaque per-environment blocks as above (task-side fields are left untouched by step and zeroed by reset); `data` [N][DATA]
 * receives the physics_state.data fields the reference's plugins read, laid out as B2S_DATA_* (include/b200sim_codes.h);
 * b200sim_data_layout writes the B2S_DATA_N_FIELDS + 1 float offsets (the last one is DATA, the block size). */
int b200sim_data_layout(const b200sim_model* model, int32_t* offsets);
int b200sim_engine_reset(const b200sim_model* model, void* stream, int n_envs, const uint32_t* rng, const float* levels,
                         const float* rand, float* state, uint32_t* keys, float* data);
int b200sim_engine_step(const b200sim_model* model, void* stream, int n_envs, const float* action, const uint32_t* rng,
                        const float* rand, float* state, uint32_t* keys, float* data);
/* get_terminations (rl.py:288-299) of the task program on data blocks: codes[N][n_terminations] int32 in {-1, 0, +1};
 * done[N] = any(code != 0), success[N] = all(code != -1) & done (rl.py:1049-1050); done / success may be NULL.  A contact is
 * matched to the model's contact pairs by its (geom1, geom2); a pair the model does not hold counts as not in contact. */
int b200sim_terminations(const b200sim_model* model, void* stream, int n_envs, const float* data, const float* levels,
                         int32_t* codes, uint8_t* done, uint8_t* success);
const char* b200sim_engine_last_error(void);

/* ---- Recurrent policy replay: the sequential part of get_ppo_variables (ksim/task/ppo.py:417-488 ->
 * examples/walking.py:674-728: xax.scan of two stacks of GRU(512) cells over the T steps of a minibatch) as persistent
 * tcgen05 kernels.  One call advances up to B2S_GRU_MAX_PROBLEMS independent GRU layers (e.g. the actor's and the critic's
 * layer l) over all T steps for B rows.  Sequence arrays are FEATURE-major, [features][T B] with column t B + b, so that the
 * kernels' accesses (one thread per batch row) are coalesced; they are plain transposed operands for the GEMMs around.  Cell = equinox.nn.GRUCell:
 *   r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r (gh_n + b_hn)), h' = n + z (h - n),
 *   gi = x W_ih^T (+ b_i, added here), gh = h W_hh^T; the carry handed on is keep_t h'_t (it restarts from zero after a
 *   terminal step, walking.py:715-719).  bf16 tensor-core operands, fp32 accumulation, fp32 state and gate arithmetic. */
typedef struct {
  const void* w_hh;   /* bf16 [3H][H] (gate order r, z, n) */
  const float* gi;    /* [3H][T B] = W_ih x^T, without bias (a plain batched GEMM over all T, done by the caller) */
  const float* b_i;   /* [3H] */
  const float* b_hn;  /* [H] */
  const float* h0;    /* [B][H] or NULL (zeros) */
  void* y;            /* bf16 [H][T B]: h'_t */
  void* hst;          /* bf16 [H][T B]: the state step t started from (what dW_hh = sum_t dgh_t^T hst_t needs) */
  float* h_last;      /* [B][H] or NULL: the carry after the last step */
  float* saved;       /* b200sim_gru_saved_floats(T, B) floats for the backward pass, or NULL (inference) */
} b200sim_gru_fwd;
typedef struct {
  const void* w_hh_t; /* bf16 [H][3H] = W_hh^T */
  const float* dy;    /* [H][T B]: gradient with respect to y */
  const float* saved; /* written by the forward call */
  void* dgi;          /* bf16 [3H][T B]: gradient w.r.t. gi; its r and z blocks are also the gradient w.r.t. gh_r, gh_z */
  void* dghn;         /* bf16 [H][T B]: gradient w.r.t. gh_n (and, summed over (t, b), b_hn) */
  float* dh0;         /* [B][H] or NULL */
} b200sim_gru_bwd;
size_t b200sim_gru_workspace_bytes(int n_problems, int n_steps, int batch, int backward);
size_t b200sim_gru_saved_floats(int n_steps, int batch);
/* keep[T][B] (1 = the episode continues after step t, 0 = it ended) or NULL.  The first int of `workspace` is a device-side
 * status word (B2S_GRU_ERR_*, 0 = ok) the caller may read after synchronising. */
int b200sim_gru_forward(void* stream, int n_problems, const b200sim_gru_fwd* problems, int n_steps, int batch, const float* keep,
                        void* workspace, size_t workspace_bytes);
int b200sim_gru_backward(void* stream, int n_problems, const b200sim_gru_bwd* problems, int n_steps, int batch, const float* keep,
                         void* workspace, size_t workspace_bytes);
const char* b200sim_gru_last_error(void);

/* ---- Policy head of the replay: the actor's mixture-of-Gaussians distribution of examples/walking.py:129-146 (prediction ->
 * means + ZEROS bias, clipped to the joint ranges; std = min((softplus + min_std) var_scale, max_std); logits) and
 * xax.MixtureOfGaussians.log_prob / sample on it.  pred[rows][3 J M] fp32 = [means | raw stds | logits], each [J][M];
 * action / logp [rows][J]; mean_bias / lo / hi [J] (NULL = none).  M <= 16. */
int b200sim_mixture_logprob(void* stream, int rows, int n_joints, int n_mix, const float* pred, const float* action,
                            const float* mean_bias, const float* lo, const float* hi, float min_std, float var_scale,
                            float max_std, float* logp);
/* d_pred[rows][3 J M] (fp32, or bf16 when d_pred_bf16) from d_logp[rows][J], or from one value per row (d_logp_per_row: the
 * layout b200sim_ppo_loss writes, the same gradient for every joint of a row). */
int b200sim_mixture_logprob_backward(void* stream, int rows, int n_joints, int n_mix, const float* pred, const float* action,
                                     const float* mean_bias, const float* lo, const float* hi, float min_std, float var_scale,
                                     float max_std, const float* d_logp, int d_logp_per_row, void* d_pred, int d_pred_bf16);
/* action ~ the mixture (argmax: the mean of the heaviest component) with a threefry stream keyed by key[2] and the device
 * word counter[0] (NULL = 0); logp (may be NULL) = log p(action). */
int b200sim_mixture_sample(void* stream, int rows, int n_joints, int n_mix, const float* pred, const float* mean_bias,
                           const float* lo, const float* hi, float min_std, float var_scale, float max_std,
                           const uint32_t* key, const int32_t* counter, int argmax, float* action, float* logp);
const char* b200sim_policy_last_error(void);

/* ---- Metric histograms: RLTask.get_histogram (ksim/task/rl.py:1521-1541) = jnp.histogram(arr.flatten(), bins) over [min, max]
 * (widened by +-0.5 when min == max; last bin closed on the right) plus min / max / sum / sum of squares / mean.  counts[bins]
 * int32; limits[bins] = the upper edge of every bin (edges[1:], as the reference stores them); stats[5] = {min, max, sum,
 * sum_squares, mean}; scratch: b200sim_histogram_scratch_bytes() bytes.  bins <= 1024.  Counts are exact and deterministic. */
size_t b200sim_histogram_scratch_bytes(void);
/* out[r] = sum over the columns of row r of a row-major bf16 matrix x[rows][cols], fp32 accumulation in a fixed order: the bias
 * gradients of the GRU replay (the reference gets them from jax.grad through `+ bias`, equinox.nn.GRUCell). */
int b200sim_row_sums(void* stream, int rows, int cols, const void* x, float* out);
int b200sim_histogram(void* stream, size_t n, const float* x, int bins, int32_t* counts, float* limits, float* stats, void* scratch);

/* ---- Optimiser step of the PPO update (ksim/task/ppo.py:536-565 apply_gradients_with_clipping -> optax chain of
 * examples/walking.py:373-387: zero_nans, clip_by_global_norm, add_decayed_weights, scale_by_adam, warmup_constant_schedule,
 * scale(-1)) on flat fp32 buffers of n elements, in place.  step_count[0] (device int32, starts at 0) is advanced by the call;
 * grads are read as grad_scale * grads (1 / world_size after a sum all-reduce).  params_bf16 (may be NULL): bf16 shadow copy of
 * the updated parameters.  scratch: b200sim_optimizer_scratch_bytes() bytes.  stats (may be NULL): [0] global gradient norm
 * before clipping, [1] learning rate used, [2] clip factor.  Deterministic (fixed reduction order). */
size_t b200sim_optimizer_scratch_bytes(void);
int b200sim_optimizer_step(void* stream, size_t n, float* params, const float* grads, float* mu, float* nu, void* params_bf16,
                           int32_t* step_count, float grad_scale, float lr_init, float lr_peak, int warmup_steps, float b1, float b2,
                           float eps, float weight_decay, float max_norm, void* scratch, float* stats);
const char* b200sim_optimizer_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
