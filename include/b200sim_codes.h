/*
 * b200sim task-program encoding: the integer codes and table layouts shared by the host (Python plugin
 * classes), the CUDA engine and the CPU oracle.  This header is the single source of truth; the Python side
 * parses the #defines (ksim_b200/codes.py).
 *
 * A "task program" is two flat arrays, ti[] (int32) and tf[] (fp32 on the device, fp64 in the oracle), that
 * describe the reference's plugin lists -- Actuators / Events / Resets / Observations / Terminations /
 * Commands / Rewards (ksim/actuators.py, events.py, resets.py, observation.py, terminations.py,
 * commands.py, rewards.py) and EngineConfig (ksim/engine.py:50-139) -- plus the per-env state / record
 * layouts derived from them.
 */
#ifndef B200SIM_CODES_H
#define B200SIM_CODES_H

#define B2S_TASK_MAGIC 0x42325354 /* "B2ST" */

/* ---- ti[] header slots ---- */
#define TI_MAGIC 0
#define TI_NSUB 1            /* phys_steps_per_ctrl_steps (engine.py:108-109) */
#define TI_ACT_PERIOD 2      /* phys_steps_per_actuator_step (engine.py:111-115) */
#define TI_ACT_TYPE 3        /* ACT_* */
#define TI_ACT_IOFF 4
#define TI_ACT_FOFF 5
#define TI_ENGINE_FOFF 6     /* tf: min_latency_step, max_latency_step, drop_prob, zero_offset rv(mean,std,mag),
                              *     action_clip_max, reward_clip_min, reward_clip_max (NaN = no clip) */
#define TI_ZERO_OFFSET_KIND 7 /* RV_* */
#define TI_N_EVENT 8
#define TI_EVENT_TAB 9
#define TI_N_RESET 10
#define TI_RESET_TAB 11
#define TI_N_OBS 12
#define TI_OBS_TAB 13
#define TI_N_TERM 14
#define TI_TERM_TAB 15
#define TI_N_CMD 16
#define TI_CMD_TAB 17
#define TI_N_REW 18
#define TI_REW_TAB 19
#define TI_N_RAND 20         /* per-env model override fields */
#define TI_RAND_TAB 21
/* sizes (floats); state/record/key buffers are env-major: buf[env][size] */
#define TI_STATE_SIZE 22     /* padded to a multiple of 4 */
#define TI_REC_SIZE 23       /* padded to a multiple of 4 */
#define TI_RAND_SIZE 24
#define TI_OBS_SIZE 25
#define TI_CMD_SIZE 26
#define TI_EVENT_SIZE 27
#define TI_OBS_CARRY_SIZE 28
#define TI_ACT_STATE_SIZE 29
#define TI_ACTION_SIZE 30
#define TI_N_REW_COMP 31
#define TI_REW_CARRY_SIZE 32
/* state offsets (floats; flags and counters are stored as exactly-representable float values) */
#define TI_S_QPOS 33
#define TI_S_QVEL 34
#define TI_S_QACC 35
#define TI_S_WARM 36
#define TI_S_TIME 37
#define TI_S_CTRL 38
#define TI_S_LAST_CTRL 39
#define TI_S_LATENCY 40
#define TI_S_ZERO_OFF 41
#define TI_S_EVENT 42
#define TI_S_ACT_STATE 43
#define TI_KEY_SIZE 44       /* u32 key buffer per env: [0:2] env rng, [2+2i:4+2i] bias key of observation i */
#define TI_S_LEVEL 45
#define TI_S_CMD 46
#define TI_S_OBS_CARRY 47
#define TI_REC_PRE_SIZE 48   /* record = [pre-step block | post-step block]; pre = obs, cmd, level */
#define TI_S_NONFINITE 49    /* per-env flag: 1 if a non-finite qpos/qvel was ever produced */
/* record offsets */
#define TI_R_QPOS 50
#define TI_R_QVEL 51
#define TI_R_XPOS 52
#define TI_R_XQUAT 53
#define TI_R_CTRL 54
#define TI_R_OBS 55
#define TI_R_CMD 56
#define TI_R_EVENT 57
#define TI_R_ACTION 58
#define TI_R_DONE 59
#define TI_R_SUCCESS 60
#define TI_R_TIME 61
#define TI_R_LEVEL 62
#define TI_R_TERM 63
#define TI_R_AUX 65          /* aux_outputs region of a record (Trajectory.aux_outputs, types.py:48-70): written by the host's
                                postprocess_trajectory hook between rollout and rewards, never by the kernels */
#define TI_AUX_SIZE 66
#define TI_S_XFRC 64         /* 6 floats wrench + body id handled via event params; -1 if unused */
#define TI_SCORE_NSEG 67     /* scoring read-set: the record columns the rewards read, as TI_SCORE_NSEG 16-byte aligned segments */
#define TI_SCORE_SEG 68      /* offset in ti of the segments: (first column, columns) pairs, ascending, disjoint */
#define TI_SCORE_WIDTH 69    /* floats per staged row (sum of the segments' columns, padded so that width / 4 is odd) */
#define TI_HEADER_SIZE 72

/* ---- plugin table entry: 8 ints each ---- */
#define TAB_STRIDE 8
#define TAB_TYPE 0
#define TAB_IOFF 1
#define TAB_FOFF 2
#define TAB_OUT_OFF 3   /* obs: offset inside obs block; cmd/event: offset inside their state block; rew: first comp */
#define TAB_OUT_DIM 4   /* obs dim / cmd floats / event floats / reward components */
#define TAB_AUX0 5      /* obs: noisy offset or -1; rew: carry offset or -1; rand: model field code */
#define TAB_AUX1 6      /* obs: carry offset or -1; rew: carry dim;          rand: count */
#define TAB_AUX2 7      /* obs: has_bias;           rew: is_penalty */

/* ---- random variables / noise / scales ---- */
#define RV_ZERO 0
#define RV_UNIFORM 1
#define RV_GAUSSIAN 2
#define RV_UNIFORM_GAUSSIAN 3
/* noise block: ints [n_links, (mul, kind) x 3]; floats [(mean, std, mag) x 3] */
#define NOISE_MAX_LINKS 3
#define NOISE_NI 7
#define NOISE_NF 9
/* bias block: ints [kind, mul]; floats [mean, std, mag] (kind RV_ZERO == no bias) */
#define BIAS_NI 2
#define BIAS_NF 3
/* scale block: ints [kind]; floats [scale, bias] */
#define SCALE_CONST 0
#define SCALE_LINEAR 1
#define SCALE_QUADRATIC 2
#define SCALE_SQRT 3
#define SCALE_NONZERO 4
#define SCALE_NI 1
#define SCALE_NF 2

/* ---- actuators (ksim/actuators.py) ---- */
#define ACT_TORQUE 0
#define ACT_POSITION 1
#define ACT_POSITION_VELOCITY 2
/* ACT_POSITION ints: [action_noise(NOISE_NI), torque_noise(NOISE_NI), rv kinds: action_bias, torque_bias,
 *   kp_scale, kd_scale, torque_limit_scale (present flag folded: -1 = absent)]
 * floats: [action_scale, action_noise(NOISE_NF), torque_noise(NOISE_NF), 5 x rv(3), kp[nu], kd[nu], clip[nu]] */

/* ---- events (ksim/events.py) ---- */
#define EVENT_LINEAR_PUSH 1
#define EVENT_ANGULAR_PUSH 2
#define EVENT_JUMP 3
#define EVENT_FORCE_PUSH 4

/* ---- resets (ksim/resets.py) ---- */
#define RESET_RANDOM_JOINT_POSITION 1
#define RESET_RANDOM_JOINT_VELOCITY 2
#define RESET_RANDOM_HEADING 3
#define RESET_RANDOM_BASE_VELOCITY_XY 4
#define RESET_PLANE_XY_POSITION 5
#define RESET_RANDOM_HEIGHT 6
#define RESET_RANDOM_PITCH_ROLL 7
#define RESET_INITIAL_MOTION_STATE 9 /* resets.py:284-304: ints [num_frames, nq]; floats [ctrl_dt, qpos frames*nq, qvel frames*nq (nq wide, as MotionReferenceData stores it)] */
#define RESET_HFIELD_XY_POSITION 8  /* resets.py:45-89: ints [nx, ny]; floats [x_bound, y_bound, z_top, base_height, x_range, y_range, lower_x, upper_x, lower_y, upper_y, heights nx*ny] */

/* ---- observations (ksim/observation.py) ---- */
#define OBS_BASE_POSITION 1
#define OBS_BASE_ORIENTATION 2
#define OBS_BASE_LINEAR_VELOCITY 3
#define OBS_BASE_ANGULAR_VELOCITY 4
#define OBS_JOINT_POSITION 5
#define OBS_JOINT_VELOCITY 6
#define OBS_CENTER_OF_MASS_INERTIA 7
#define OBS_CENTER_OF_MASS_VELOCITY 8
#define OBS_ACTUATOR_FORCE 9
#define OBS_SENSOR 10
#define OBS_BASE_LINEAR_ACCELERATION 11
#define OBS_BASE_ANGULAR_ACCELERATION 12
#define OBS_ACTUATOR_ACCELERATION 13
#define OBS_PROJECTED_GRAVITY 14
#define OBS_FEET_CONTACT 15
#define OBS_BODY_POSITION 16
#define OBS_BODY_VELOCITY 17
#define OBS_FEET_FORCE 18
#define OBS_FEET_TORQUE 19
#define OBS_TIMESTEP 20
#define OBS_CONTACT 21           /* observation.py:452-478: ints [n, geom ids]; out[i] = any_j colliding(g_i, g_j) */
#define OBS_FEET_POSITION 22     /* examples/kbot/train.py:653-689: feet relative to the base, in the base's yaw frame */
#define OBS_BASE_HEIGHT 23       /* examples/kbot/train.py:693-697: xpos[1][2] */
#define OBS_COM_DISTANCE 24      /* examples/kbot/train.py:500-649: |hull centroid of the contact points - subtree_com[2]| */
#define OBS_DELAYED_JOINT_POSITION 25  /* observation.py:229-252: ints [delay_steps]; carry = delay_steps x (nq - 7) ring of qpos[7:] */
#define OBS_DELAYED_JOINT_VELOCITY 26  /* observation.py:261-283: ints [delay_steps]; carry = delay_steps x (nv - 6) ring of qvel[6:] */
#define OBS_BODY_ORIENTATION 27  /* observation.py:619-642 (+ FeetOrientationObservation 672-703): ints [n, body ids]; xquat rows */
#define OBS_SITE_ORIENTATION 28  /* observation.py:645-669: ints [n, site ids]; first two rows of site_xmat (6-D rotation) */
#define OBS_ACT_POS 29           /* observation.py:717-747: ints [ctrl_idx, qpos_idx]; (last_ctrl[ctrl_idx], qpos[qpos_idx]) */

/* ---- terminations (ksim/terminations.py) ---- */
#define TERM_NOT_UPRIGHT 1
#define TERM_MINIMUM_HEIGHT 2
#define TERM_ILLEGAL_CONTACT 3
#define TERM_BAD_Z 4
#define TERM_BAD_VELOCITY 5
#define TERM_FAR_FROM_ORIGIN 6
#define TERM_EPISODE_LENGTH 7
#define TERM_TERRAIN_BAD_Z 8     /* examples/kbot/train.py:779-813: base z - lowest foot z < unhealthy_z */

/* ---- commands (ksim/commands.py) ---- */
#define CMD_LINEAR_VELOCITY 1
#define CMD_ANGULAR_VELOCITY 2
#define CMD_FLOAT_VECTOR 3
#define CMD_INT_VECTOR 5         /* commands.py:137-169: ints [lo, hi]*n; floats [switch_prob, zero_prob]; randint and the zero mask share the key */
#define CMD_START_POSITION 6     /* commands.py:173-192: qpos[:3] at the reset, constant afterwards */
#define CMD_START_QUATERNION 7   /* commands.py:195-214: qpos[3:7] at the reset, constant afterwards */
#define CMD_BASE_HEIGHT 8        /* commands.py:773-796: floats [min, max, switch_prob]; the redraw and its switch share the key */
#define CMD_CARTESIAN_COORDINATE 9  /* commands.py:537-692: state [current xyz, target xyz, speed]; floats [box_min 3, box_max 3, dt, min_speed, max_speed, switch_prob, jump_prob, curriculum_scale] */
#define CMD_SINUSOIDAL_GAIT 10   /* commands.py:703-770: state [moving_flag, phase, height[num_feet]]; ints [num_feet]; floats [gait_period, ctrl_dt, max_height, height_offset, stance_ratio, moving_threshold] */
#define CMD_JOINT_POSITION 11    /* commands.py:799-880: state [start n, target n, total_time, num_steps, step_id]; ints [n, indices]; floats [lo, hi]*n, ctrl_dt, min_time, max_time */
#define CMD_UNIFIED 4            /* examples/kbot/train.py:701-775: [vx vy wz bh rx ry arms[10]], 6-way mode switch */

/* ---- rewards (ksim/rewards.py) ---- */
#define REW_STAY_ALIVE 1
#define REW_UPRIGHT 2
#define REW_LINEAR_VELOCITY 3
#define REW_ANGULAR_VELOCITY 4
#define REW_FEET_AIR_TIME 5
#define REW_SPARSE_TARGET_HEIGHT 6
#define REW_FORCE_PENALTY 7
#define REW_MOTIONLESS_AT_REST 8
#define REW_ACTION_VELOCITY 9
#define REW_ACTION_ACCELERATION 10
#define REW_ACTION_JERK 11
#define REW_BASE_HEIGHT 12
#define REW_BASE_HEIGHT_RANGE 13
#define REW_OFF_AXIS_VELOCITY 14
#define REW_SMALL_JOINT_VELOCITY 15
/* the K-Bot task's own rewards (examples/kbot/train.py:128-496) */
/* more of ksim/rewards.py (not used by the two example tasks): ints / floats listed with each class in ksim_b200/rewards.py */
#define REW_SMALL_JOINT_ACCELERATION 27
#define REW_SMALL_JOINT_JERK 28
#define REW_AVOID_LIMITS 29
#define REW_TORQUE 30
#define REW_ENERGY 31
#define REW_JOINT_DEVIATION 32
#define REW_FLAT_BODY 33
#define REW_NO_ROLL 34
#define REW_LINK_ACCELERATION 35
#define REW_LINK_JERK 36
#define REW_FEET_SLIP 37
#define REW_REACHABILITY 38
#define REW_SYMMETRY 39
#define REW_PAIRWISE_SYMMETRY 40
#define REW_AMP 41               /* ksim/task/amp.py:100-112: max(0, 1 - 0.25 (logit - 1)^2) of an aux output; ints [aux offset] */
#define REW_POSITION_TRACKING 42  /* rewards.py:704-744: ints [tracked body, base body, command offset, l2]; floats [kernel_scale] */
#define REW_SINUSOIDAL_GAIT 43   /* rewards.py:1320-1348: ints [position obs offset, num_feet, command height offset]; floats [max_height] */
#define REW_JOINT_POSITION 44    /* rewards.py:1352-1414: ints [n, command offset, indices]; floats [pos_scale, vel_scale, sq_scale, abs_scale]; comps pos, vel */
#define REW_INTERSECTION 45      /* rewards.py:1418-1430: ints [position obs offset, n]; floats [min_distance] */
#define REW_KBOT_SINGLE_FOOT_CONTACT 16
#define REW_KBOT_NO_CONTACT 17
#define REW_KBOT_FEET_AIRTIME 18
#define REW_KBOT_ARM_POSITION 19
#define REW_KBOT_LINVEL_TRACKING 20
#define REW_KBOT_ANGVEL_TRACKING 21
#define REW_KBOT_XY_ORIENTATION 22
#define REW_KBOT_TERRAIN_BASE_HEIGHT 23
#define REW_KBOT_FEET_ORIENTATION 24
#define REW_KBOT_COM_DISTANCE 25
#define REW_KBOT_BASE_ACCELERATION 26

/* ---- per-env model override fields (ksim/randomization.py) ---- */
#define RAND_DOF_FRICTIONLOSS 1
#define RAND_DOF_ARMATURE 2
#define RAND_DOF_DAMPING 3
#define RAND_BODY_MASS 4
#define RAND_BODY_INERTIA 5
#define RAND_BODY_IPOS 6
#define RAND_GEOM_SIZE 7         /* [ngeom][3]; the engine keeps the rows of the geoms that appear in contact pairs */
#define RAND_GEOM_POS 8          /* [ngeom][3] */
#define RAND_SITE_QUAT 9         /* [nsite][4] (IMUAlignmentRandomizer, randomization.py:245-283) */
#define RAND_SITE_POS 10         /* [nsite][3] */

/* ---- GAE flags (ksim/task/ppo.py:54-121) ---- */
#define GAE_NORMALIZE_ADVANTAGES 1
#define GAE_MONTE_CARLO_RETURNS 2

/* ---- PPO loss flags / layout (ksim/task/ppo.py:149-246) ---- */
#define PPO_CLIPPED_VALUE_LOSS 1
#define PPO_N_TERMS 4            /* losses[PPO_N_TERMS][B][T]: policy, value, kl (already x kl_scale), entropy */
#define PPO_MAX_ACTIONS 128      /* widest log-prob row the kernel stages */
#define PPO_MAX_CTAS 2368        /* grid cap: 16 resident CTAs (the 2048-thread limit) on each of 148 SMs */
#define PPO_SCRATCH_BYTES 94720  /* caller-owned device scratch: per-CTA partial sums (<= PPO_MAX_CTAS x 5 doubles) */

/* ---- recurrent replay kernels (b200sim_gru_forward / _backward) ---- */
#define B2S_GRU_HIDDEN 512       /* hidden size the tensor-core kernels are built for (examples/walking.py:262, kbot train.py) */
#define B2S_GRU_MAX_PROBLEMS 2   /* GRU layers of different networks run side by side in one launch (actor + critic) */
/* device-side status word (first int of the workspace): which bounded wait gave up (0 = none) */
#define B2S_GRU_ERR_ACC 1        /* accumulator-ready barrier */
#define B2S_GRU_ERR_FLAG 2       /* a peer CTA's state chunk never arrived */
#define B2S_GRU_ERR_TMEM 3       /* accumulator never drained */
#define B2S_GRU_ERR_FULL 4       /* bulk copy never landed */
#define B2S_GRU_ERR_EMPTY 5      /* pipeline stage never freed */
#define B2S_GRU_ERR_PEERS 6      /* a peer CTA of the cluster never finished reading the previous state */

/* ---- b200sim_model_set_option ---- */
#define B2S_OPT_SOLVER 0
#define B2S_SOLVER_FAST 0        /* constraint-space / factor-reusing Newton, exact one-pass line search (default) */
#define B2S_SOLVER_REFERENCE 1   /* the reference-shaped iteration: primal Newton, bracketing line search, both at the model's caps */

/* ---- fields of the per-environment data block of b200sim_engine_step / _reset (the physics_state.data.* the reference's
 * plugins read), in block order; b200sim_data_layout returns their float offsets ---- */
#define B2S_DATA_QPOS 0            /* [nq] */
#define B2S_DATA_QVEL 1            /* [nv] */
#define B2S_DATA_QACC 2            /* [nv] */
#define B2S_DATA_CTRL 3            /* [nu] */
#define B2S_DATA_TIME 4            /* [1] */
#define B2S_DATA_XPOS 5            /* [nbody][3] */
#define B2S_DATA_XQUAT 6           /* [nbody][4] wxyz */
#define B2S_DATA_CINERT 7          /* [nbody][10] */
#define B2S_DATA_CVEL 8            /* [nbody][6] */
#define B2S_DATA_ACTUATOR_FORCE 9  /* [nu] */
#define B2S_DATA_SENSORDATA 10     /* [nsensordata] */
#define B2S_DATA_CFRC_EXT 11       /* [nbody][6] */
#define B2S_DATA_SITE_XPOS 12      /* [nsite][3] */
#define B2S_DATA_SITE_XMAT 13      /* [nsite][9] */
#define B2S_DATA_SUBTREE_COM 14    /* [nbody][3] */
#define B2S_DATA_XFRC_APPLIED 15   /* [nbody][6] */
#define B2S_DATA_NCON 16           /* [1]: candidate contacts (fixed per model, as in MJX); a contact is active when dist < 0 */
#define B2S_DATA_CONTACT_GEOM1 17  /* [ncon] geom ids, stored as floats */
#define B2S_DATA_CONTACT_GEOM2 18  /* [ncon] */
#define B2S_DATA_CONTACT_DIST 19   /* [ncon] */
#define B2S_DATA_CONTACT_POS 20    /* [ncon][3] */
#define B2S_DATA_N_FIELDS 21

/* ---- metric histograms (b200sim_histogram) ---- */
#define B2S_HIST_MAX_CTAS 592    /* 4 x 148: per-CTA partial statistics in the scratch buffer */

/* ---- flat-buffer optimiser step (b200sim_optimizer_step) ---- */
#define B2S_OPTIM_MAX_PARTIALS 1024  /* per-CTA partial sums of the gradient norm (grid = a multiple of the SM count, <= this) */

/* ---- error codes ---- */
#define B2S_OK 0
#define B2S_ERR_BAD_ARG 1
#define B2S_ERR_BAD_BLOB 2
#define B2S_ERR_UNSUPPORTED 3
#define B2S_ERR_CUDA 4
#define B2S_ERR_TOO_LARGE 5

#endif
