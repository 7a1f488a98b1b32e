"""Environment configuration with the reference's attribute names and defaults.

Mirrors `Config` of GCA/envs/config.py:3-185 (same names, same values, same "set before __init__ to
override a derived value" idiom via hasattr) so that a reference Config subclass — e.g. GA3C's
`Train(EnvConfig)`, GA3C/Config.py:31 — can subclass this class unchanged.  Only the fields that the
step path reads are interpreted by this package (see `to_ca_config`); the rest are carried for API parity.
"""
import os

import numpy as np

from . import _abi

_OTHER_STD = np.array([5.0, 5.0, 1.0, 1.0, 1.0, 5.0, 1.0], dtype=np.float32)
_OTHER_MEAN = np.array([0.0, 0.0, 0.0, 0.0, 0.5, 0.0, 1.0], dtype=np.float32)

SUPPORTED_STATES_IN_OBS = ['is_learning', 'num_other_agents', 'dist_to_goal', 'heading_ego_frame', 'pref_speed',
                           'radius', 'other_agents_states']


def _scalar_state(attr, mean, std, bounds):
    return {'dtype': np.float32, 'size': 1, 'bounds': bounds, 'attr': attr,
            'std': np.array([std], dtype=np.float32), 'mean': np.array([mean], dtype=np.float32)}


class Config(object):
    def __init__(self):
        s = self
        s.COLLISION_AVOIDANCE = True
        s.continuous, s.discrete = range(2)
        s.ACTION_SPACE_TYPE = s.continuous
        # display (never on in training; plotting is out of scope here)
        s.ANIMATE_EPISODES = s.SHOW_EPISODE_PLOTS = s.SAVE_EPISODE_PLOTS = False
        if not hasattr(s, 'PLOT_CIRCLES_ALONG_TRAJ'):
            s.PLOT_CIRCLES_ALONG_TRAJ = True
        s.ANIMATION_PERIOD_STEPS = 5
        s.PLT_LIMITS, s.PLT_FIG_SIZE = None, (10, 8)
        if not hasattr(s, 'USE_STATIC_MAP'):
            s.USE_STATIC_MAP = False
        s.TRAIN_MODE, s.PLAY_MODE, s.EVALUATE_MODE = True, False, False
        # rewards
        s.REWARD_AT_GOAL = 1.0
        s.REWARD_COLLISION_WITH_AGENT = -0.25
        s.REWARD_COLLISION_WITH_WALL = -0.25
        s.REWARD_GETTING_CLOSE = -0.1
        s.REWARD_ENTERED_NORM_ZONE = -0.05
        s.REWARD_TIME_STEP = 0.0
        s.REWARD_WIGGLY_BEHAVIOR = 0.0
        s.WIGGLY_BEHAVIOR_THRESHOLD = np.inf
        s.COLLISION_DIST = 0.0
        s.GETTING_CLOSE_RANGE = 0.2
        s.SOCIAL_NORMS = "none"
        # simulation
        s.DT = 0.2
        s.NEAR_GOAL_THRESHOLD = 0.2
        s.MAX_TIME_RATIO = 2.
        # scenarios
        s.TEST_CASE_FN = "get_testcase_random"
        s.TEST_CASE_ARGS = {
            'policy_to_ensure': 'learning_ga3c',
            'policies': ['noncoop', 'learning_ga3c', 'static'],
            'policy_distr': [0.05, 0.9, 0.05],
            'speed_bnds': [0.5, 2.0],
            'radius_bnds': [0.2, 0.8],
            'side_length': [{'num_agents': [0, 5], 'side_length': [4, 5]},
                            {'num_agents': [5, np.inf], 'side_length': [6, 8]}],
        }
        if not hasattr(s, 'MAX_NUM_AGENTS_IN_ENVIRONMENT'):
            s.MAX_NUM_AGENTS_IN_ENVIRONMENT = 4
        if not hasattr(s, 'MAX_NUM_AGENTS_TO_SIM'):
            s.MAX_NUM_AGENTS_TO_SIM = 4
        s.MAX_NUM_OTHER_AGENTS_IN_ENVIRONMENT = s.MAX_NUM_AGENTS_IN_ENVIRONMENT - 1
        if not hasattr(s, 'MAX_NUM_OTHER_AGENTS_OBSERVED'):
            s.MAX_NUM_OTHER_AGENTS_OBSERVED = s.MAX_NUM_AGENTS_IN_ENVIRONMENT - 1
        s.PLOT_EVERY_N_EPISODES = 100
        # sensors
        s.SENSING_HORIZON = np.inf
        s.LASERSCAN_LENGTH, s.LASERSCAN_NUM_PAST = 512, 3
        s.NUM_STEPS_IN_OBS_HISTORY = 1
        s.NUM_PAST_ACTIONS_IN_STATE = 0
        s.RVO_TIME_HORIZON, s.RVO_COLLAB_COEFF, s.RVO_ANTI_COLLAB_T = 5.0, 0.5, 1.0
        # observation vector
        s.TRAIN_SINGLE_AGENT = False
        M = s.MAX_NUM_OTHER_AGENTS_OBSERVED
        s.STATE_INFO_DICT = {
            'dist_to_goal': _scalar_state('get_agent_data("dist_to_goal")', 0., 5., [-np.inf, np.inf]),
            'radius': _scalar_state('get_agent_data("radius")', 0.5, 1.0, [0, np.inf]),
            'heading_ego_frame': _scalar_state('get_agent_data("heading_ego_frame")', 0., 3.14, [-np.pi, np.pi]),
            'pref_speed': _scalar_state('get_agent_data("pref_speed")', 1.0, 1.0, [0, np.inf]),
            'num_other_agents': _scalar_state('get_agent_data("num_other_agents_observed")', 1.0, 1.0, [0, np.inf]),
            'other_agent_states': {'dtype': np.float32, 'size': 7, 'bounds': [-np.inf, np.inf],
                                   'attr': 'get_agent_data("other_agent_states")',
                                   'std': _OTHER_STD.copy(), 'mean': _OTHER_MEAN.copy()},
            'other_agents_states': {'dtype': np.float32, 'size': (M, 7), 'bounds': [-np.inf, np.inf],
                                    'attr': 'get_sensor_data("other_agents_states")',
                                    'std': np.tile(_OTHER_STD, (M, 1)), 'mean': np.tile(_OTHER_MEAN, (M, 1))},
            'laserscan': {'dtype': np.float32, 'size': (s.LASERSCAN_NUM_PAST, s.LASERSCAN_LENGTH), 'bounds': [0., 6.],
                          'attr': 'get_sensor_data("laserscan")',
                          'std': 5. * np.ones((s.LASERSCAN_NUM_PAST, s.LASERSCAN_LENGTH), dtype=np.float32),
                          'mean': 5. * np.ones((s.LASERSCAN_NUM_PAST, s.LASERSCAN_LENGTH), dtype=np.float32)},
            'is_learning': {'dtype': np.float32, 'size': 1, 'bounds': [0., 1.],
                            'attr': 'get_agent_data_equiv("policy.str", "learning")'},
            'other_agents_states_encoded': {'dtype': np.float32, 'size': 100., 'bounds': [0., 1.],
                                            'attr': 'get_sensor_data("other_agents_states_encoded")'},
        }
        s.setup_obs()
        s.AGENT_SORTING_METHOD = "closest_first"

    def setup_obs(self):
        if not hasattr(self, "STATES_IN_OBS"):
            self.STATES_IN_OBS = list(SUPPORTED_STATES_IN_OBS)
        if not hasattr(self, "STATES_NOT_USED_IN_POLICY"):
            self.STATES_NOT_USED_IN_POLICY = ['is_learning']
        self.MEAN_OBS, self.STD_OBS = {}, {}
        for state in self.STATES_IN_OBS:
            info = self.STATE_INFO_DICT[state]
            if 'mean' in info:
                self.MEAN_OBS[state] = info['mean']
            if 'std' in info:
                self.STD_OBS[state] = info['std']


class EvaluateConfig(Config):
    def __init__(self):
        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 19
        Config.__init__(self)
        self.EVALUATE_MODE, self.TRAIN_MODE = True, False
        self.DT = 0.1
        self.MAX_TIME_RATIO = 8.


class Example(EvaluateConfig):
    def __init__(self):
        EvaluateConfig.__init__(self)
        self.SAVE_EPISODE_PLOTS = self.PLOT_CIRCLES_ALONG_TRAJ = self.ANIMATE_EPISODES = True


def to_ca_config(cfg, num_worlds, device=0, auto_reset=0):
    """Translate a (reference-shaped) Config object into the C-ABI struct.  Raises on anything the CUDA step
    path does not implement, instead of silently computing something else."""
    if list(cfg.STATES_IN_OBS) != SUPPORTED_STATES_IN_OBS:
        raise NotImplementedError("STATES_IN_OBS %r: only %r is implemented on the GPU path (laserscan / encoded "
                                  "sensors are out of scope)" % (cfg.STATES_IN_OBS, SUPPORTED_STATES_IN_OBS))
    if getattr(cfg, 'USE_STATIC_MAP', False):
        raise NotImplementedError("USE_STATIC_MAP=True (static maps / wall collisions) is out of scope")
    if np.isfinite(getattr(cfg, 'WIGGLY_BEHAVIOR_THRESHOLD', np.inf)) and cfg.REWARD_WIGGLY_BEHAVIOR != 0.0:
        raise NotImplementedError("the wiggly-behaviour reward term is not implemented (inert in the reference defaults)")
    if cfg.AGENT_SORTING_METHOD not in _abi.SORT_METHODS:
        raise ValueError("Did not supply proper AGENT_SORTING_METHOD")
    mode = _abi.OVER_ALL_LEARNING_DONE
    if cfg.EVALUATE_MODE:
        mode = _abi.OVER_ALL_DONE
    elif cfg.TRAIN_SINGLE_AGENT:
        mode = _abi.OVER_FIRST_AGENT_DONE
    # clip bounds as CollisionAvoidanceEnv._initialize_rewards computes them (collision_avoidance_env.py:463-483)
    possible = np.array([cfg.REWARD_AT_GOAL, cfg.REWARD_COLLISION_WITH_AGENT, cfg.REWARD_TIME_STEP,
                         cfg.REWARD_COLLISION_WITH_WALL, cfg.REWARD_WIGGLY_BEHAVIOR])
    return _abi.default_config(
        num_worlds, cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT, cfg.MAX_NUM_OTHER_AGENTS_OBSERVED,
        sort_method=_abi.SORT_METHODS[cfg.AGENT_SORTING_METHOD], game_over_mode=mode, auto_reset=int(auto_reset),
        device=int(device), dt=float(cfg.DT), near_goal_threshold=float(cfg.NEAR_GOAL_THRESHOLD),
        getting_close_range=float(cfg.GETTING_CLOSE_RANGE), reward_at_goal=float(cfg.REWARD_AT_GOAL),
        reward_collision_with_agent=float(cfg.REWARD_COLLISION_WITH_AGENT), reward_time_step=float(cfg.REWARD_TIME_STEP),
        min_possible_reward=float(possible.min()), max_possible_reward=float(possible.max()),
        max_time_ratio=float(cfg.MAX_TIME_RATIO), sensing_horizon=float(cfg.SENSING_HORIZON))


def load_config_from_env(default_class=Config):
    """Same selection mechanism as GCA/envs/__init__.py:1-13: GYM_CONFIG_PATH / GYM_CONFIG_CLASS."""
    path = os.environ.get('GYM_CONFIG_PATH')
    name = os.environ.get('GYM_CONFIG_CLASS')
    if not path:
        cls = globals().get(name, default_class) if name else default_class
        return cls()
    import importlib.util
    spec = importlib.util.spec_from_file_location(name or 'Config', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cls = getattr(mod, name or 'Config', None)
    assert callable(cls), "config class %r not found in %s" % (name, path)
    return cls()
