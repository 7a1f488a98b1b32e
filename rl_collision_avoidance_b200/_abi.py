"""ctypes mirror of include/ca_step.h (struct ca_config, enums, column indices).

Kept in lock-step with the header by tests/test_abi.py (sizes, field order, exported symbols).
"""
import ctypes as C
import math

CA_ABI_VERSION = 1
CA_MAX_AGENTS = 32

CA_OK = 0
CA_ERR_INVALID_ARG = -1
CA_ERR_CUDA = -2
CA_ERR_NOT_INITIALISED = -3
CA_ERR_UNSUPPORTED = -4
CA_ERR_ALLOC = -5

SORT_CLOSEST_FIRST, SORT_CLOSEST_LAST, SORT_TIME_TO_IMPACT = 0, 1, 2
SORT_METHODS = {"closest_first": SORT_CLOSEST_FIRST, "closest_last": SORT_CLOSEST_LAST,
                "time_to_impact": SORT_TIME_TO_IMPACT}

OVER_ALL_LEARNING_DONE, OVER_ALL_DONE, OVER_FIRST_AGENT_DONE = 0, 1, 2

POLICY_LEARNING_GA3C, POLICY_NONCOOP, POLICY_STATIC, POLICY_LEARNING = 0, 1, 2, 3
# policy strings of GCA/envs/test_cases.py:48-58 that are supported inside step()
POLICY_IDS = {"learning_ga3c": POLICY_LEARNING_GA3C, "noncoop": POLICY_NONCOOP, "static": POLICY_STATIC,
              "learning": POLICY_LEARNING}

F_AT_GOAL, F_WAS_AT_GOAL, F_IN_COLLISION, F_WAS_IN_COLLISION, F_RAN_OUT_OF_TIME = 1, 2, 4, 8, 16
F_DONE_MASK = F_AT_GOAL | F_IN_COLLISION | F_RAN_OUT_OF_TIME

(I_PX, I_PY, I_GX, I_GY, I_PREF_SPEED, I_RADIUS, I_HEADING, I_POLICY, I_TIME_REMAINING, I_RESERVED) = range(10)
INIT_STRIDE = 10
(S_PX, S_PY, S_HEADING, S_VX, S_VY, S_TIME_REMAINING, S_GX, S_GY, S_RADIUS, S_PREF_SPEED, S_FLAGS,
 S_POLICY) = range(12)
STATE_STRIDE = 12

OBS_HOST_LEN = 6
OBS_OTHER_LEN = 7


def obs_len(M):
    return OBS_HOST_LEN + OBS_OTHER_LEN * M


class CaConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("num_worlds", C.c_int32),
        ("max_agents", C.c_int32),
        ("max_others_observed", C.c_int32),
        ("sort_method", C.c_int32),
        ("game_over_mode", C.c_int32),
        ("auto_reset", C.c_int32),
        ("device", C.c_int32),
        ("dt", C.c_double),
        ("near_goal_threshold", C.c_double),
        ("getting_close_range", C.c_double),
        ("reward_at_goal", C.c_double),
        ("reward_collision_with_agent", C.c_double),
        ("reward_time_step", C.c_double),
        ("min_possible_reward", C.c_double),
        ("max_possible_reward", C.c_double),
        ("max_time_ratio", C.c_double),
        ("max_heading_change", C.c_double),
        ("sensing_horizon", C.c_double),
    ]


class CaGa3cBuffers(C.Structure):
    _fields_ = [
        ("obs_ring", C.c_void_p), ("act_ring", C.c_void_p), ("rew_ring", C.c_void_p), ("length", C.c_void_p),
        ("tcount", C.c_void_p), ("done_trained", C.c_void_p), ("out_x", C.c_void_p), ("out_r", C.c_void_p),
        ("out_a", C.c_void_p), ("out_count", C.c_void_p), ("capacity", C.c_int32), ("reserved", C.c_int32),
        ("out_src", C.c_void_p), ("gathered", C.c_void_p),
    ]


class CaPredictorParams(C.Structure):
    """ca_predictor_params: device pointers to the float32 parameters in the TF variable layout."""
    FIELDS = ("lstm_kernel", "lstm_bias", "layer1_kernel", "layer1_bias", "layer2_kernel", "layer2_bias", "fc1_kernel",
              "fc1_bias", "logits_p_kernel", "logits_p_bias", "logits_v_kernel", "logits_v_bias", "input_avg", "input_std")
    _fields_ = [(n, C.c_void_p) for n in FIELDS]


CA_PREDICTOR_BLOB_BYTES = 357536
CA_PREDICT_PLAN_COUNTERS = 64
CA_PREDICTOR_MAX_OTHERS = 22


class CaScenarioConfig(C.Structure):
    _fields_ = [("min_agents", C.c_int32), ("max_agents", C.c_int32), ("side_split_agents", C.c_int32),
                ("ensure_learner", C.c_int32)] + \
               [(n, C.c_double) for n in ("side_small_lo", "side_small_hi", "side_large_lo", "side_large_hi", "p_swap",
                                          "p_circle", "speed_lo", "speed_hi", "radius_lo", "radius_hi", "p_noncoop",
                                          "p_learning")]


def default_config(num_worlds, max_agents, max_others_observed=None, **overrides):
    """Reference defaults, GCA/envs/config.py:30-47,64-76,171 and collision_avoidance_env.py:76,463-483."""
    cfg = CaConfig()
    cfg.abi_version = CA_ABI_VERSION
    cfg.num_worlds = int(num_worlds)
    cfg.max_agents = int(max_agents)
    cfg.max_others_observed = int(max_others_observed if max_others_observed is not None else max(max_agents - 1, 1))
    cfg.sort_method = SORT_CLOSEST_FIRST
    cfg.game_over_mode = OVER_ALL_LEARNING_DONE
    cfg.auto_reset = 0
    cfg.device = 0
    cfg.dt = 0.2
    cfg.near_goal_threshold = 0.2
    cfg.getting_close_range = 0.2
    cfg.reward_at_goal = 1.0
    cfg.reward_collision_with_agent = -0.25
    cfg.reward_time_step = 0.0
    cfg.min_possible_reward = -0.25
    cfg.max_possible_reward = 1.0
    cfg.max_time_ratio = 2.0
    cfg.max_heading_change = math.pi / 3
    cfg.sensing_horizon = math.inf
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("ca_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg
