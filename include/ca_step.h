/*
 * ca_step.h — C-ABI of libcastep.so, the B200-native (sm_100a) vectorised replacement of the
 * reference's env.step() hot path.  Plain pointers and sizes only; no torch types.
 *
 * What each entry point replaces in mit-acl/rl_collision_avoidance (paths under /root/reference,
 * GCA = gym-collision-avoidance/gym_collision_avoidance, GA3C = ga3c/GA3C):
 *
 *   ca_create / ca_destroy     CollisionAvoidanceEnv.__init__           GCA/envs/collision_avoidance_env.py:41-129
 *                              (+ the Config values it reads            GCA/envs/config.py:30-47,64-76,171)
 *   ca_set_world_state         CollisionAvoidanceEnv.set_agents         GCA/envs/collision_avoidance_env.py:260-268
 *                              + Agent.__init__/Agent.reset             GCA/envs/agent.py:29-136
 *   ca_reset                   CollisionAvoidanceEnv.reset              GCA/envs/collision_avoidance_env.py:196-215
 *   ca_step / ca_step_host     CollisionAvoidanceEnv.step               GCA/envs/collision_avoidance_env.py:131-194
 *                                _take_action                           :217-252
 *                                Agent.take_action                      GCA/envs/agent.py:190-238
 *                                UnicycleDynamics.step                  GCA/envs/dynamics/UnicycleDynamics.py:14-47
 *                                Dynamics.update_ego_frame              GCA/envs/dynamics/Dynamics.py:24-41
 *                                _compute_rewards/_check_for_collisions GCA/envs/collision_avoidance_env.py:319-409
 *                                OtherAgentsStatesSensor.sense          GCA/envs/sensors/OtherAgentsStatesSensor.py:20-144
 *                                MultiagentDictToMultiagentArrayWrapper GCA/envs/wrappers.py:111-139
 *                                _check_which_agents_done               GCA/envs/collision_avoidance_env.py:411-439
 *                              and, with auto_reset, DummyVecEnv.step_wait (openai/baselines@ea25b9e, not vendored)
 *   ca_get_state               reading Agent attributes                 GCA/envs/agent.py:66-136
 *   ca_nstep_returns           ProcessAgent._accumulate_rewards         GA3C/ProcessAgent.py:54-79
 *   ca_ga3c_record             ProcessAgent.run_episode bookkeeping     GA3C/ProcessAgent.py:105-211
 *   ca_ga3c_episode_stats      ProcessAgent.run / ProcessStats.run      GA3C/ProcessAgent.py:220-243, ProcessStats.py:62-117
 *
 * Conventions
 *   - every function returns 0 (CA_OK) or a negative ca_status; nothing throws across the ABI;
 *     ca_strerror(code) names the code, ca_last_error() gives the last detailed message (thread local).
 *   - the library owns the internal state arrays; the caller owns every I/O buffer and the stream.
 *   - a handle is bound to one CUDA device; calls on one handle must be serialised by the caller;
 *     device-pointer entry points are asynchronous (stream ordered); distinct handles are independent.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - there is NO CPU fallback: without a usable CUDA device ca_create fails with CA_ERR_CUDA.
 */
#ifndef CA_STEP_H_
#define CA_STEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_ABI_VERSION 1

/* hard limit: one world's agents live in the lanes of one warp */
#define CA_MAX_AGENTS 32

/* status codes */
typedef enum ca_status {
  CA_OK = 0,
  CA_ERR_INVALID_ARG = -1,
  CA_ERR_CUDA = -2,
  CA_ERR_NOT_INITIALISED = -3, /* step/reset before ca_set_world_state */
  CA_ERR_UNSUPPORTED = -4,
  CA_ERR_ALLOC = -5
} ca_status;

/* AGENT_SORTING_METHOD, GCA/envs/config.py:169-172 */
typedef enum ca_sort_method {
  CA_SORT_CLOSEST_FIRST = 0,
  CA_SORT_CLOSEST_LAST = 1,
  CA_SORT_TIME_TO_IMPACT = 2
} ca_sort_method;

/* game_over rule, GCA/envs/collision_avoidance_env.py:427-437 */
typedef enum ca_game_over_mode {
  CA_OVER_ALL_LEARNING_DONE = 0, /* training (default) */
  CA_OVER_ALL_DONE = 1,          /* Config.EVALUATE_MODE */
  CA_OVER_FIRST_AGENT_DONE = 2   /* Config.TRAIN_SINGLE_AGENT */
} ca_game_over_mode;

/* per-agent policy type (what acts inside step), GCA/envs/test_cases.py:48-58 */
typedef enum ca_policy {
  CA_POLICY_LEARNING_GA3C = 0, /* external, discrete action 0..10    policies/LearningPolicyGA3C.py:13-27 */
  CA_POLICY_NONCOOP = 1,       /* internal                           policies/NonCooperativePolicy.py:9-22 */
  CA_POLICY_STATIC = 2,        /* internal, goal := pos              policies/StaticPolicy.py:9-23 */
  CA_POLICY_LEARNING = 3       /* external, continuous [speed_frac, heading_frac]  policies/LearningPolicy.py:13-33 */
} ca_policy;

/* agent flag bits (state column CA_S_FLAGS, and the bits of the `done` output's source) */
#define CA_F_AT_GOAL 1u              /* Agent.is_at_goal */
#define CA_F_WAS_AT_GOAL 2u          /* Agent.was_at_goal_already */
#define CA_F_IN_COLLISION 4u         /* Agent.in_collision */
#define CA_F_WAS_IN_COLLISION 8u     /* Agent.was_in_collision_already */
#define CA_F_RAN_OUT_OF_TIME 16u     /* Agent.ran_out_of_time */
#define CA_F_DONE_MASK (CA_F_AT_GOAL | CA_F_IN_COLLISION | CA_F_RAN_OUT_OF_TIME)

/* columns of the `init` tensor of ca_set_world_state: double[W][A][CA_INIT_STRIDE] */
enum {
  CA_I_PX = 0, CA_I_PY, CA_I_GX, CA_I_GY, CA_I_PREF_SPEED, CA_I_RADIUS, CA_I_HEADING, CA_I_POLICY,
  CA_I_TIME_REMAINING, /* Agent.time_remaining_to_reach_goal at reset; NaN => library computes
                          max(max_time_ratio*(|p-g|-near_goal_threshold)/pref_speed, dt)  (agent.py:98-103) */
  CA_I_RESERVED,
  CA_INIT_STRIDE
};

/* columns of ca_get_state: double[W][A][CA_STATE_STRIDE] */
enum {
  CA_S_PX = 0, CA_S_PY, CA_S_HEADING, CA_S_VX, CA_S_VY, CA_S_TIME_REMAINING, CA_S_GX, CA_S_GY,
  CA_S_RADIUS, CA_S_PREF_SPEED, CA_S_FLAGS, CA_S_POLICY,
  CA_STATE_STRIDE
};

/* Observation row (float32), order = Config.STATES_IN_OBS of GA3C/Config.py:40, flattened like
 * GCA/envs/wrappers.py:115-139:  [is_learning, num_other_agents, dist_to_goal, heading_ego_frame,
 * pref_speed, radius, M x (p_prll, p_orth, v_prll, v_orth, other_radius, combined_radius, dist_2_other)]
 * L = 6 + 7*M.  Rows of absent agents and unused other-slots are zero. */
#define CA_OBS_HOST_LEN 6
#define CA_OBS_OTHER_LEN 7
#define CA_OBS_LEN(M) (CA_OBS_HOST_LEN + CA_OBS_OTHER_LEN * (M))

typedef struct ca_config {
  int32_t abi_version;          /* must be CA_ABI_VERSION */
  int32_t num_worlds;           /* W: independent environments advanced per launch */
  int32_t max_agents;           /* A: Config.MAX_NUM_AGENTS_IN_ENVIRONMENT, 1..CA_MAX_AGENTS */
  int32_t max_others_observed;  /* M: Config.MAX_NUM_OTHER_AGENTS_OBSERVED, 1..A-1 (>=1) */
  int32_t sort_method;          /* ca_sort_method */
  int32_t game_over_mode;       /* ca_game_over_mode */
  int32_t auto_reset;           /* 1: on game_over a world reloads its injected initial state inside the same
                                   launch and `obs` carries the NEW episode's first observation while
                                   reward/done/game_over describe the finished step (DummyVecEnv semantics) */
  int32_t device;               /* CUDA device ordinal */
  double dt;                    /* Config.DT */
  double near_goal_threshold;   /* Config.NEAR_GOAL_THRESHOLD */
  double getting_close_range;   /* Config.GETTING_CLOSE_RANGE */
  double reward_at_goal;        /* Config.REWARD_AT_GOAL */
  double reward_collision_with_agent; /* Config.REWARD_COLLISION_WITH_AGENT */
  double reward_time_step;      /* Config.REWARD_TIME_STEP */
  double min_possible_reward;   /* clip bounds, collision_avoidance_env.py:463-483 */
  double max_possible_reward;
  double max_time_ratio;        /* Config.MAX_TIME_RATIO (only used when CA_I_TIME_REMAINING is NaN) */
  double max_heading_change;    /* env.max_heading_change = pi/3 (continuous LearningPolicy only) */
  double sensing_horizon;       /* Config.SENSING_HORIZON (inf) */
} ca_config;

typedef struct ca_env ca_env;

/* Fill *cfg with the reference's default training configuration (GCA/envs/config.py) for W worlds of A agents. */
int ca_default_config(ca_config* cfg, int32_t num_worlds, int32_t max_agents);

int ca_create(const ca_config* cfg, ca_env** out);
int ca_destroy(ca_env* env);

/* Inject the initial configuration of every world (≙ env.set_agents + Agent.__init__) and reset.
 * init: double[W][A][CA_INIT_STRIDE]; num_agents: int32[W] with 1 <= n_w <= A (rows >= n_w ignored).
 * on_device != 0: both pointers are device pointers (copied stream-ordered on `stream`).
 * Either way the call returns after `stream` has drained (it reads back whether every world has all A agents, which
 * selects the step kernel's loop rendering); not capturable into a CUDA graph. */
int ca_set_world_state(ca_env* env, const double* init, const int32_t* num_agents, int on_device, void* stream);

/* Replace only the reset snapshot (same tensors as ca_set_world_state): live worlds keep running and pick the new
 * scenario — which may have a different agent count — up at their next reset / auto-reset.  This is how fresh
 * scenarios are streamed in (≙ test_case_fn being called again on every env.reset(), collision_avoidance_env.py:283). */
int ca_set_reset_state(ca_env* env, const double* init, const int32_t* num_agents, int on_device, void* stream);

/* On-device scenario generator (≙ get_testcase_random + cadrl_test_case_to_agents, GCA/envs/test_cases.py:95-118,263-326,
 * and generate_rand_test_case_multi, GCA/envs/policies/CADRL/scripts/multi/gen_rand_testcases.py:104-437): writes fresh
 * random test cases into the reset snapshot.  Distributional parity with the reference (Philox instead of MT19937). */
typedef struct ca_scenario_config {
  int32_t min_agents, max_agents;       /* num_agents ~ U{min..max} (reference: 2..MAX_NUM_AGENTS_IN_ENVIRONMENT) */
  int32_t side_split_agents;            /* worlds with fewer agents use the small side range (config.py:57-60: 5) */
  int32_t ensure_learner;               /* policy_to_ensure = 'learning_ga3c' (config.py:51) */
  double side_small_lo, side_small_hi;  /* [4, 5] */
  double side_large_lo, side_large_hi;  /* [6, 8] */
  double p_swap, p_circle;              /* 0.15, 0.15 (gen_rand_testcases.py:123-133); the rest are random cases */
  double speed_lo, speed_hi;            /* speed_bnds [0.5, 2.0]; pref_speed = max of two draws */
  double radius_lo, radius_hi;          /* radius_bnds [0.2, 0.8] */
  double p_noncoop, p_learning;         /* policy_distr [0.05, 0.9, 0.05] over noncoop / learning_ga3c / static */
} ca_scenario_config;

int ca_default_scenario_config(ca_scenario_config* cfg, int32_t max_agents);

/* Fill the reset snapshot of every world (only_consumed == 0) or of the worlds that have reset since the last call
 * (only_consumed != 0) with new scenarios.  Stream ordered; call ca_reset afterwards to start all worlds from them. */
int ca_generate_scenarios(ca_env* env, const ca_scenario_config* cfg, uint64_t seed, int only_consumed, void* stream);

/* Reset worlds to their injected initial state (world_mask: device uint8[W], NULL = all) and write the
 * first observation of every world (unmasked worlds: their current observation) to obs (device float[W][A][L]);
 * sorted_idx (device int32[W][A][M], may be NULL) receives the neighbour order (-1 = empty slot). */
int ca_reset(ca_env* env, const uint8_t* world_mask, float* obs, int32_t* sorted_idx, void* stream);

/* Advance every world one time step.  All pointers are DEVICE pointers.
 *   actions      int32[W][A]      discrete action 0..10 per CA_POLICY_LEARNING_GA3C agent (others ignored)
 *   cont_actions double[W][A][2]  [speed_frac, heading_frac] per CA_POLICY_LEARNING agent; NULL = every such
 *                                 agent gets the no-op command [0, 0.5] (zero speed, zero heading change)
 *   obs          float[W][A][L]   next observation
 *   reward       float[W][A]      0 for absent agents
 *   done         uint8[W][A]      which_agents_done; 1 for absent agents
 *   game_over    uint8[W]
 *   sorted_idx   int32[W][A][M]   optional (NULL to skip): indices chosen by the sensor, -1 = empty slot */
int ca_step(ca_env* env, const int32_t* actions, const double* cont_actions, float* obs, float* reward,
            uint8_t* done, uint8_t* game_over, int32_t* sorted_idx, void* stream);

/* Same step through HOST buffers (what a host-side env.step() caller sees): copies actions host->device,
 * launches, copies obs/reward/done/game_over device->host and synchronises.  Buffers registered with
 * cudaHostRegister / allocated pinned make the copies asynchronous DMA; pageable memory also works. */
int ca_step_host(ca_env* env, const int32_t* actions, const double* cont_actions, float* obs, float* reward,
                 uint8_t* done, uint8_t* game_over, int32_t* sorted_idx);

/* The same call split in two, mirroring VecEnv.step_async / step_wait (openai/baselines vec_env.py — the interface
 * MultiagentDummyVecEnv implements, GCA/envs/wrappers.py:104-109): _async enqueues H2D actions -> kernel -> D2H results on
 * the handle's private stream and returns; _wait blocks until the results are in the caller's buffers.  The buffers must
 * stay valid until _wait (page-locked buffers make the copies overlap with other handles' work).  One step may be in
 * flight per handle; distinct handles are independent, so a caller driving several environments keeps the PCIe link
 * busy by waiting on env k only after it has enqueued env k+1. */
int ca_step_host_async(ca_env* env, const int32_t* actions, const double* cont_actions, float* obs, float* reward,
                       uint8_t* done, uint8_t* game_over, int32_t* sorted_idx);
int ca_step_host_wait(ca_env* env);

/* ca_reset through HOST buffers (world_mask host uint8[W] or NULL; obs/sorted_idx host, sorted_idx may be NULL). */
int ca_reset_host(ca_env* env, const uint8_t* world_mask, float* obs, int32_t* sorted_idx);

/* Copy the full agent state out: double[W][A][CA_STATE_STRIDE] (device pointer if on_device, else host; host
 * copies synchronise).  CA_S_TIME_REMAINING is rebuilt by repeating the `-= dt` subtractions the agent's steps stand for
 * from the budget of its last reset (same float64 value as the reference's attribute; budgets of more than 65 534
 * steps are reported as they were at the reset). */
int ca_get_state(ca_env* env, double* out, int on_device, void* stream);

/* Time step used by the steps that follow (the value is read when a step is launched).  The library keeps every agent's
 * time budget as the number of `-= dt` steps it has left, so a call that CHANGES dt recounts them on the device: it
 * drains the device before and after (no stream argument; not capturable).  A call with the current dt is free.
 * Replaces the per-call `dt` argument of CollisionAvoidanceEnv.step(actions, dt=None)
 * (GCA/envs/collision_avoidance_env.py:131-138; Agent.take_action(action, dt), GCA/envs/agent.py:190).  dt must be > 0. */
int ca_set_dt(ca_env* env, double dt);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int ca_launch_count(const ca_env* env, int64_t* out);

/* Pinned (page-locked) host memory for the *_host entry points, so that a ctypes/cgo caller gets
 * asynchronous DMA copies without depending on another CUDA binding. */
int ca_host_alloc(void** out, uint64_t bytes);
int ca_host_free(void* ptr);

/* n-step discounted returns over a rollout segment (≙ ProcessAgent._accumulate_rewards, GA3C/ProcessAgent.py:54-79),
 * batched over N independent agent streams, device pointers:
 *   reward  float[T][N], bootstrap float[N] (terminal_reward: V(s_T) or 0 when done), out float[T][N]
 *   out[t] = reward[t] + gamma * out[t+1], out[T] := bootstrap. */
int ca_nstep_returns(const float* reward, const float* bootstrap, float* out, int32_t T, int32_t N,
                     float gamma, int device, void* stream);

/* ---- GA3C actor bookkeeping (vectorised ProcessAgent.run_episode, GA3C/ProcessAgent.py:105-211) ----
 * All pointers are device pointers owned by the caller.  N = W*A agent slots, R ring slots (R >= time_max + 2),
 * L = observation length.  Ring slot of env step t is (t mod R); the env must have written the observation that
 * step t's prediction was made from into obs_ring[t mod R] (pass obs_ring + ((t+1) mod R)*N*L as ca_step's obs). */
typedef struct ca_ga3c_buffers {
  const float* obs_ring;  /* [R][N][L] */
  int32_t* act_ring;      /* [R][N] */
  float* rew_ring;        /* [R][N]  rewards, overwritten in place by discounted returns like the reference */
  int32_t* length;        /* [N] len(experiences[i]) */
  int32_t* tcount;        /* [N] time_counts[i] */
  uint8_t* done_trained;  /* [N] which_agents_done_and_trained[i] */
  float* out_x;           /* [capacity][L-1] emitted x_ rows (observation without the is_learning column) */
  float* out_r;           /* [capacity] emitted r_ (discounted returns) */
  int32_t* out_a;         /* [capacity] emitted action indices (the reference one-hot encodes them) */
  int32_t* out_count;     /* [1] rows emitted so far (caller zeroes it); > capacity means rows were dropped */
  int32_t capacity;
  int32_t reserved;
  int32_t* out_src;       /* [capacity] scratch: (ring slot * N + agent slot) of every emitted row */
  int32_t* gathered;      /* [2] scratch: [0] rows whose x_/r_/a_ are already copied, [1] internal; zero both with out_count */
} ca_ga3c_buffers;

/* Append step t's experience of every learning agent and emit the training rows the reference's
 * run_episode()/_accumulate_rewards() would yield at this step. */
int ca_ga3c_record(const ca_ga3c_buffers* bufs, int64_t t, int32_t ring_slots, int32_t num_slots, int32_t agents_per_world,
                   int32_t obs_len, int32_t time_max, float gamma, const int32_t* actions, const float* values,
                   const float* reward, const uint8_t* done, const uint8_t* game_over, int device, void* stream);

/* Episode statistics (ProcessAgent.run :220-243): obs_now = obs_ring slot of step t, ep_reward float[W] and
 * ep_steps int32[W] are running accumulators, stats double[3] += {episodes, sum of scores, learning-agent steps}. */
int ca_ga3c_episode_stats(const float* obs_now, const float* reward, const uint8_t* game_over, float* ep_reward,
                          int32_t* ep_steps, double* stats, int32_t num_worlds, int32_t agents_per_world,
                          int32_t obs_len, int device, void* stream);

/* One fused LSTM time step of the predictor (ThreadPredictor / NetworkVP_rnn forward, GA3C/NetworkVP_rnn.py:58-66), device
 * pointers, float32: obs rows [B][obs_stride] (raw observation: column 1 = num_other_agents, other agent t at columns
 * 6+7t..6+7t+6), zh = h @ kernel[7:71] of the previous state ([B][256], NULL when h = 0), Kx = kernel[0:7] ([7][256]),
 * bias [256], avg7/std7 = normalisation of the 7 other-agent columns, c/h [B][64] updated in place. */
int ca_lstm_step(const float* obs, int32_t obs_stride, const float* zh, const float* Kx, const float* bias,
                 const float* avg7, const float* std7, float* c, float* h, int32_t batch, int32_t t, int device,
                 void* stream);

/* ---- trainer: the LSTM cell of NetworkVP_rnn (tf.nn.rnn_cell.LSTMCell(64) under dynamic_rnn, GA3C/NetworkVP_rnn.py:63-66;
 * gradients as tf.train.AdamOptimizer.minimize differentiates it, GA3C/NetworkVPCore.py:100-123) as one forward and one
 * backward launch per time step.  Device pointers, float32.  z [B][256] = x_t Kx + h Kh + b, gate order i, j, f, o; seq_len =
 * sequence_length of row r at seq_len[r * seq_stride] (the raw num_other_agents column); rows with t >= sequence_length
 * keep their state.  forward writes the gate activations (sigmoid(i), tanh(j), sigmoid(f + 1), sigmoid(o)) to gates
 * [B][256] and the new state to c, h [B][64].  backward takes the gradients w.r.t. the new state (dc, dh: nullable = 0) and
 * writes dz [B][256], dc_prev [B][64] and dh_pass [B][64] (the gradient that reaches h_prev without going through z:
 * dh for masked rows, 0 otherwise). */
int ca_lstm_cell_forward(const float* z, const float* c_prev, const float* h_prev, const float* seq_len, int32_t seq_stride,
                         int32_t t, float* gates, float* c, float* h, int32_t batch, int device, void* stream);
int ca_lstm_cell_backward(const float* gates, const float* c_prev, const float* c_new, const float* seq_len,
                          int32_t seq_stride, int32_t t, const float* dc, const float* dh, float* dz, float* dc_prev,
                          float* dh_pass, int32_t batch, int device, void* stream);

/* ---- fused predictor: ThreadPredictor.run (GA3C/ThreadPredictor.py:40-75) -> NetworkVP_rnn forward
 * (GA3C/NetworkVP_rnn.py:39-108, GA3C/NetworkVPCore.py:64-77) -> ProcessAgent.select_action (GA3C/ProcessAgent.py:98-103)
 * as ONE kernel launch over all observation rows (tcgen05 tensor-core products, fp16 operands / fp32 accumulation).
 *
 * ca_predictor_params: device pointers to the float32 parameters in the TF-1.x variable layout (kernels are [in][out],
 * row-major) — rnn/lstm_cell/{kernel [71][256], bias [256]} (rows 0..6 other-agent state, 7..70 h; gate order i, j, f, o;
 * forget bias 1.0 is added by the library), layer1/{kernel [68][256] (rows 0..3 host state, 4..67 h), bias},
 * layer2, fullyconnected1 ([256][256]), logits_p ([256][11]), logits_v ([256][1]) and the NN-input normalisation
 * vectors NN_INPUT_AVG_VECTOR / NN_INPUT_STD_VECTOR (GA3C/Config.py:64-71; at least 12 entries are read). */
typedef struct ca_predictor_params {
  const float* lstm_kernel;
  const float* lstm_bias;
  const float* layer1_kernel;
  const float* layer1_bias;
  const float* layer2_kernel;
  const float* layer2_bias;
  const float* fc1_kernel;
  const float* fc1_bias;
  const float* logits_p_kernel;
  const float* logits_p_bias;
  const float* logits_v_kernel;
  const float* logits_v_bias;
  const float* input_avg;
  const float* input_std;
} ca_predictor_params;

/* size of the packed parameter image ca_predictor_pack writes and ca_predict reads (device memory, 128-byte aligned) */
#define CA_PREDICTOR_BLOB_BYTES 357536

/* Re-pack the parameters into the kernel's shared-memory images (call once per weight update; stream-ordered). */
int ca_predictor_pack(const ca_predictor_params* params, void* blob, int device, void* stream);

/* One forward pass over `batch` raw observation rows (float32, row stride obs_stride floats; column 0 is_learning,
 * 1 num_other_agents = LSTM sequence length, 2..5 host state, 6+7t..12+7t other agent t; num_others <= 22 LSTM steps).
 * Outputs (each nullable): p [batch][11] softmax policy with MIN_POLICY mixing, v [batch], actions int32 [batch] =
 * argmax(p) when greedy != 0 (PLAY_MODE / EVALUATE_MODE), else one draw from p per row (counter-based generator keyed by
 * (seed, offset, row): pass a new offset every call).  error_flag (nullable, device int32) is set to 1 if an internal
 * barrier wait times out (the kernel then traps instead of hanging). */
int ca_predict(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, const void* blob, float* p, float* v,
               int32_t* actions, int32_t greedy, float min_policy, uint64_t seed, uint64_t offset, int32_t* error_flag,
               int device, void* stream);

/* Row plan for ca_predict_rows.  In the reference only LEARNING agents ever ask the predictor (ProcessAgent.run_episode,
 * GA3C/ProcessAgent.py:128-133: rows whose is_learning observation is set), and dynamic_rnn runs num_other_agents steps
 * per row (GA3C/NetworkVP_rnn.py:58-66).  ca_predict_plan lists the rows with is_learning != 0 sorted by descending LSTM
 * sequence length into row_index int32[batch] (device) and writes counters int32[CA_PREDICT_PLAN_COUNTERS] (device):
 * counters[0] = number of planned rows, counters[1 + b] = rows with sequence length b.  Rows that need no prediction get
 * v = 0 and actions = 0 (both nullable).  Stream-ordered, no host synchronisation. */
#define CA_PREDICT_PLAN_COUNTERS 64
int ca_predict_plan(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, int32_t* row_index,
                    int32_t* counters, float* v, int32_t* actions, int device, void* stream);

/* ca_predict over the planned rows only: row_index / n_rows = the plan's row_index and &counters[0] (device pointers; the
 * row count is read on the device).  Outputs are written at the ORIGINAL row positions ([batch]-sized arrays, same
 * meaning as ca_predict); rows outside the plan are not touched.  Tiles of 128 plan entries share one sequence length,
 * so each tile runs exactly the LSTM steps its rows need. */
int ca_predict_rows(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, const void* blob,
                    const int32_t* row_index, const int32_t* n_rows, float* p, float* v, int32_t* actions, int32_t greedy,
                    float min_policy, uint64_t seed, uint64_t offset, int32_t* error_flag, int device, void* stream);

const char* ca_strerror(int code);
const char* ca_last_error(void);
int ca_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CA_STEP_H_ */
