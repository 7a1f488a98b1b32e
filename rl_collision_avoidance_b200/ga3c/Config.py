"""GA3C configuration classes with the reference's names and values (GA3C/Config.py:31-184): `Train(EnvConfig)`,
`TrainPhase1`, `TrainPhase2`.  Selected by GYM_CONFIG_CLASS / GYM_CONFIG_PATH exactly like the reference
(GA3C/__init__.py:1-12); `get_config()` instantiates the chosen class once per process.

Extra, GPU-only knobs (not in the reference) are prefixed GPU_.
"""
import os

import numpy as np

from rl_collision_avoidance_b200.config import Config as EnvConfig


class Train(EnvConfig):
    def __init__(self):
        s = self
        if not hasattr(s, "MAX_NUM_AGENTS_IN_ENVIRONMENT"):
            s.MAX_NUM_AGENTS_IN_ENVIRONMENT = 4
        if not hasattr(s, "MAX_NUM_AGENTS_TO_SIM"):
            s.MAX_NUM_AGENTS_TO_SIM = 4
        s.STATES_IN_OBS = ['is_learning', 'num_other_agents', 'dist_to_goal', 'heading_ego_frame', 'pref_speed', 'radius',
                           'other_agents_states']
        s.STATES_NOT_USED_IN_POLICY = ['is_learning']
        s.MULTI_AGENT_ARCH_RNN, s.MULTI_AGENT_ARCH_WEIGHT_SHARING, s.MULTI_AGENT_ARCH_LASERSCAN = range(3)
        s.MULTI_AGENT_ARCH = s.MULTI_AGENT_ARCH_RNN
        s.MAX_NUM_OTHER_AGENTS_OBSERVED = s.MAX_NUM_AGENTS_IN_ENVIRONMENT - 1   # RNN architecture (:48-49)
        EnvConfig.__init__(s)

        s.game_grid, s.game_ale, s.game_collision_avoidance = range(3)
        s.GAME_CHOICE = s.game_collision_avoidance
        s.USE_WANDB = False
        s.WANDB_PROJECT_NAME = "ga3c_cadrl"
        s.DEBUG = False
        s.RANDOM_SEED_1000 = 0

        # observation -> network input (:60-75)
        s.USE_IMAGE = False
        avg, std, size = [], [], 0
        for state in s.STATES_IN_OBS:
            if state in s.STATES_NOT_USED_IN_POLICY:
                continue
            info = s.STATE_INFO_DICT[state]
            size += int(np.prod(info['size']))
            avg.append(np.asarray(info['mean']).flatten())
            std.append(np.asarray(info['std']).flatten())
        s.NN_INPUT_SIZE = size
        s.NN_INPUT_AVG_VECTOR = np.hstack(avg)
        s.NN_INPUT_STD_VECTOR = np.hstack(std)
        s.FIRST_STATE_INDEX = 1
        s.HOST_AGENT_OBSERVATION_LENGTH = 4
        s.OTHER_AGENT_OBSERVATION_LENGTH = 7
        s.OTHER_AGENT_FULL_OBSERVATION_LENGTH = s.OTHER_AGENT_OBSERVATION_LENGTH
        s.HOST_AGENT_STATE_SIZE = s.HOST_AGENT_OBSERVATION_LENGTH
        s.NUM_ACTIONS = 11
        s.LOAD_RL_THEN_TRAIN_RL, s.TRAIN_ONLY_REGRESSION, s.LOAD_REGRESSION_THEN_TRAIN_RL = range(3)

        s.NET_ARCH = 'NetworkVP_rnn'
        s.ALL_ARCHS = ['NetworkVP_rnn']
        s.NORMALIZE_INPUT = True
        s.USE_DROPOUT = False
        s.USE_REGULARIZATION = True

        # The reference runs AGENTS OS processes with one env each; here AGENTS worlds would be a waste of a GPU,
        # see GPU_NUM_WORLDS below.  PREDICTORS / TRAINERS are kept for the stats line only.
        s.AGENTS, s.PREDICTORS, s.TRAINERS = 32, 2, 2
        s.DEVICE = 'cuda:0'
        s.DYNAMIC_SETTINGS = False
        s.DYNAMIC_SETTINGS_STEP_WAIT, s.DYNAMIC_SETTINGS_INITIAL_WAIT = 20, 10

        s.DISCOUNT = 0.97
        s.TIME_MAX = int(4 / s.DT)
        s.MAX_QUEUE_SIZE = 100
        s.PREDICTION_BATCH_SIZE = 128
        s.MIN_POLICY = 0.0
        s.OPT_RMSPROP, s.OPT_ADAM = range(2)
        s.OPTIMIZER = s.OPT_ADAM
        s.LEARNING_RATE_RL_START = s.LEARNING_RATE_RL_END = 2e-5
        s.RMSPROP_DECAY, s.RMSPROP_MOMENTUM, s.RMSPROP_EPSILON = 0.99, 0.0, 0.1
        s.BETA_START = s.BETA_END = 1e-4
        s.USE_GRAD_CLIP, s.GRAD_CLIP_NORM = False, 40.0
        s.LOG_EPSILON = 1e-6
        s.TRAINING_MIN_BATCH_SIZE = 100

        s.TENSORBOARD, s.TENSORBOARD_UPDATE_FREQUENCY = False, 100
        s.SAVE_MODELS, s.SAVE_FREQUENCY = True, 50000
        s.SPECIAL_EPISODES_TO_SAVE = []
        s.PRINT_STATS_FREQUENCY = 1
        s.STAT_ROLLING_MEAN_WINDOW = 1000
        s.RESULTS_FILENAME = 'results.txt'
        s.NETWORK_NAME = 'network'

        # ---- GPU-only knobs
        s.GPU_NUM_WORLDS = int(os.environ.get('GA3C_GPU_NUM_WORLDS', 4096))   # worlds stepped per launch on each GPU
        # rows per optimiser step (>= TRAINING_MIN_BATCH_SIZE); 0 = auto = 8192 rows, the batch the committed learning curve
        # (profiles/r02_learning_curve.json) was validated with
        s.GPU_TRAIN_BATCH = int(os.environ.get('GA3C_GPU_TRAIN_BATCH', 0))
        # The reference's trainer takes one Adam step per ~100-200 rows (ThreadTrainer.py:44) with SUM losses, i.e. about
        # a million updates over a TrainPhase1 run.  A vectorised trainer that steps once per GPU_TRAIN_BATCH rows takes
        # GPU_TRAIN_BATCH / GPU_REF_BATCH times fewer steps; Adam normalises the gradient scale, so each step moves the
        # weights by ~lr whatever the batch — the learning rate therefore has to grow with the batch to cover the same
        # distance in weight space.  GPU_LR_SCALE: 'linear' (lr * B / GPU_REF_BATCH, the update budget of the reference
        # cadence), 'sqrt', 'none', or a number (explicit multiplier); capped by GPU_LR_MAX.  DESIGN.md §6 has the
        # learning curves this default was chosen from.
        s.GPU_REF_BATCH = 128
        s.GPU_LR_SCALE = os.environ.get('GA3C_GPU_LR_SCALE', 'sqrt')
        s.GPU_LR_MAX = float(os.environ.get('GA3C_GPU_LR_MAX', 3e-3))
        s.GPU_PRINT_EVERY_S = 2.0
        # optimiser steps on exactly GPU_TRAIN_BATCH rows replayed from CUDA graphs (NetworkVP_rnn.GraphedTrainStep)
        s.GPU_TRAIN_GRAPH = int(os.environ.get('GA3C_GPU_TRAIN_GRAPH', 1))
        # trainer matmuls on the tensor cores in TF32 (10-bit significand, fp32 accumulation); off = fp32 like the reference
        s.GPU_TRAIN_TF32 = int(os.environ.get('GA3C_GPU_TRAIN_TF32', 0))


class TrainPhase1(Train):
    def __init__(self):
        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 4
        self.MAX_NUM_AGENTS_TO_SIM = 4
        Train.__init__(self)
        self.TRAIN_VERSION = self.LOAD_REGRESSION_THEN_TRAIN_RL
        self.LOAD_FROM_WANDB_RUN_ID = 'run-rnn'
        self.EPISODE_NUMBER_TO_LOAD = 0
        self.EPISODES = 1500000
        self.ANNEALING_EPISODE_COUNT = 1500000
        self.SPECIAL_EPISODES_TO_SAVE = [1490000, 1500000]


class TrainPhase2(Train):
    def __init__(self):
        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 10
        self.MAX_NUM_AGENTS_TO_SIM = 10
        Train.__init__(self)
        self.EPISODES = 2000000
        self.ANNEALING_EPISODE_COUNT = 2000000
        self.TRAIN_VERSION = self.LOAD_RL_THEN_TRAIN_RL
        self.LOAD_FROM_WANDB_RUN_ID = 'run-20200324_221727-2tz70xqi'
        self.EPISODE_NUMBER_TO_LOAD = 1490000
        self.SPECIAL_EPISODES_TO_SAVE = [1990000, 2000000]


_instance = None


def get_config():
    """GA3C/__init__.py:1-12: class named by GYM_CONFIG_CLASS (default TrainPhase1) from GYM_CONFIG_PATH (default: here)."""
    global _instance
    if _instance is None:
        name = os.environ.get('GYM_CONFIG_CLASS', 'TrainPhase1')
        path = os.environ.get('GYM_CONFIG_PATH')
        if path and os.path.abspath(path) != os.path.abspath(__file__):
            import importlib.util
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            cls = getattr(mod, name, None)
        else:
            cls = globals().get(name)
        assert callable(cls), "config class %r not found" % name
        _instance = cls()
    return _instance


def set_config(cfg):
    global _instance
    _instance = cfg
