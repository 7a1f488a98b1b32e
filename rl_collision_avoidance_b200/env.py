"""`gym.Env`-shaped single-world façade with the reference's interface.

Mirrors `CollisionAvoidanceEnv` (GCA/envs/collision_avoidance_env.py:22-521): no-argument constructor that
reads the global Config, `reset() -> obs dict`, `step(actions, dt=None) -> (obs dict, rewards, game_over,
info)`, `set_agents`, `set_testcase`, `set_static_map`, `set_plot_save_dir`, `set_perturbed_info`, and the
light-weight `Agent` / policy / dynamics / sensor classes a caller uses to describe a scenario
(GCA/envs/agent.py:29-55, GCA/envs/test_cases.py:48-67).  The arithmetic happens in the CUDA library through
`HostVecEnv` (one world); this file only translates between the reference's object/dict API and the C-ABI
tensors.  Many worlds at once: use `vec_env.VecCollisionAvoidanceEnv` directly.
"""
import numpy as np

from . import _abi
from . import config as _config
from .vec_env import HostVecEnv

Config = None  # instantiated lazily like GCA/envs/__init__.py does at import time


def get_config():
    global Config
    if Config is None:
        Config = _config.load_config_from_env()
    return Config


def set_config(cfg):
    global Config
    Config = cfg


# ---------------------------------------------------------------- policies / dynamics / sensors (descriptors only)

class Policy(object):
    """GCA/envs/policies/Policy.py:4-16 — the three attributes the env reads."""
    str = "NoPolicy"
    is_external = False
    is_still_learning = False
    ca_policy = None

    def __init__(self):
        pass


class LearningPolicy(Policy):          # GCA/envs/policies/LearningPolicy.py
    str = "learning"
    is_external = True
    is_still_learning = True
    ca_policy = _abi.POLICY_LEARNING


class LearningPolicyGA3C(LearningPolicy):   # GCA/envs/policies/LearningPolicyGA3C.py
    ca_policy = _abi.POLICY_LEARNING_GA3C


class NonCooperativePolicy(Policy):    # GCA/envs/policies/NonCooperativePolicy.py
    str = "NonCooperativePolicy"
    ca_policy = _abi.POLICY_NONCOOP


class StaticPolicy(Policy):            # GCA/envs/policies/StaticPolicy.py
    str = "Static"
    ca_policy = _abi.POLICY_STATIC


policy_dict = {'noncoop': NonCooperativePolicy, 'learning': LearningPolicy, 'learning_ga3c': LearningPolicyGA3C,
               'static': StaticPolicy}


class UnicycleDynamics(object):        # the only entry of the reference's dynamics_dict (test_cases.py:65-67)
    def __init__(self, agent=None):
        self.agent = agent


class OtherAgentsStatesSensor(object):
    name = 'other_agents_states'


class Agent(object):
    """Scenario descriptor + read-back of the simulated state, constructor as GCA/envs/agent.py:29-55."""

    def __init__(self, start_x, start_y, goal_x, goal_y, radius, pref_speed, initial_heading, policy,
                 dynamics_model=UnicycleDynamics, sensors=(OtherAgentsStatesSensor,), id=0):
        cfg = get_config()
        self.policy = policy() if isinstance(policy, type) else policy
        if getattr(self.policy, 'ca_policy', None) is None:
            raise NotImplementedError("policy %r has no GPU implementation (supported: %s)" % (policy, sorted(policy_dict)))
        if dynamics_model is not UnicycleDynamics and not isinstance(dynamics_model, UnicycleDynamics):
            raise NotImplementedError("only UnicycleDynamics is implemented")
        self.dynamics_model = UnicycleDynamics(self)
        self.sensors = [s() if isinstance(s, type) else s for s in sensors]
        self.id = id
        self.near_goal_threshold = cfg.NEAR_GOAL_THRESHOLD
        self.dt_nominal = cfg.DT
        self.reset(px=start_x, py=start_y, gx=goal_x, gy=goal_y, pref_speed=pref_speed, radius=radius,
                   heading=initial_heading)

    def reset(self, px=None, py=None, gx=None, gy=None, pref_speed=None, radius=None, heading=None):
        """Agent.reset, GCA/envs/agent.py:57-136 (only the fields with a meaning on the GPU path)."""
        cfg = get_config()
        if px is not None and py is not None:
            self.pos_global_frame = np.array([px, py], dtype='float64')
        if gx is not None and gy is not None:
            self.goal_global_frame = np.array([gx, gy], dtype='float64')
        self.vel_global_frame = np.array([0.0, 0.0], dtype='float64')
        if heading is None:
            v = self.goal_global_frame - self.pos_global_frame
            self.heading_global_frame = np.arctan2(v[1], v[0])
        else:
            self.heading_global_frame = heading
        if radius is not None:
            self.radius = radius
        if pref_speed is not None:
            self.pref_speed = pref_speed
        self.straight_line_time_to_reach_goal = \
            (np.linalg.norm(self.pos_global_frame - self.goal_global_frame) - self.near_goal_threshold) / self.pref_speed
        self.time_remaining_to_reach_goal = max(cfg.MAX_TIME_RATIO * self.straight_line_time_to_reach_goal, self.dt_nominal)
        self.t = 0.0
        self.step_num = 0
        self.is_at_goal = self.was_at_goal_already = False
        self.in_collision = self.was_in_collision_already = False
        self.ran_out_of_time = False
        self.is_done = False
        self.dist_to_goal = 0.0
        self.heading_ego_frame = 0.0
        self.num_other_agents_observed = 0

    def _init_row(self):
        row = np.zeros(_abi.INIT_STRIDE)
        row[_abi.I_PX], row[_abi.I_PY] = self.pos_global_frame
        row[_abi.I_GX], row[_abi.I_GY] = self.goal_global_frame
        row[_abi.I_PREF_SPEED], row[_abi.I_RADIUS] = self.pref_speed, self.radius
        row[_abi.I_HEADING] = self.heading_global_frame
        row[_abi.I_POLICY] = self.policy.ca_policy
        row[_abi.I_TIME_REMAINING] = self.time_remaining_to_reach_goal
        return row

    def _read_back(self, s, obs_row, moved, dt):
        self.pos_global_frame = s[[_abi.S_PX, _abi.S_PY]].copy()
        self.goal_global_frame = s[[_abi.S_GX, _abi.S_GY]].copy()
        self.vel_global_frame = s[[_abi.S_VX, _abi.S_VY]].copy()
        self.heading_global_frame = s[_abi.S_HEADING]
        self.speed_global_frame = float(np.hypot(*self.vel_global_frame))
        self.time_remaining_to_reach_goal = s[_abi.S_TIME_REMAINING]
        f = int(s[_abi.S_FLAGS])
        self.is_at_goal = bool(f & _abi.F_AT_GOAL)
        self.was_at_goal_already = bool(f & _abi.F_WAS_AT_GOAL)
        self.in_collision = bool(f & _abi.F_IN_COLLISION)
        self.was_in_collision_already = bool(f & _abi.F_WAS_IN_COLLISION)
        self.ran_out_of_time = bool(f & _abi.F_RAN_OUT_OF_TIME)
        self.is_done = bool(f & _abi.F_DONE_MASK)
        self.num_other_agents_observed = int(obs_row[1])
        self.dist_to_goal = float(obs_row[2])
        self.heading_ego_frame = float(obs_row[3])
        if moved:
            self.t += dt
            self.step_num += 1


# ---------------------------------------------------------------- minimal spaces (gym is optional)

class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)


class DictSpace(object):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})


class CollisionAvoidanceEnv(object):
    """Single-world environment, interface of GCA/envs/collision_avoidance_env.py:22-521."""

    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 30}

    def __init__(self, device=0):
        cfg = get_config()
        self.id = 0
        self._device = device
        self.num_agents = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
        self.dt_nominal = cfg.DT
        self.collision_dist = cfg.COLLISION_DIST
        self.getting_close_range = cfg.GETTING_CLOSE_RANGE
        self.evaluate = cfg.EVALUATE_MODE
        self.max_heading_change = np.pi / 3
        self.min_heading_change = -self.max_heading_change
        self.min_speed, self.max_speed = 0.0, 1.0
        self.action_space_type = cfg.ACTION_SPACE_TYPE
        self.action_space = Box(np.array([self.min_speed, self.min_heading_change]),
                                np.array([self.max_speed, self.max_heading_change]), dtype=np.float32)
        self.observation_space = DictSpace({})
        for state in cfg.STATES_IN_OBS:
            info = cfg.STATE_INFO_DICT[state]
            self.observation_space.spaces[state] = Box(info['bounds'][0] * np.ones(info['size']),
                                                       info['bounds'][1] * np.ones(info['size']), dtype=info['dtype'])
        self._ca_cfg = _config.to_ca_config(cfg, 1, device=device)
        self._env = HostVecEnv(self._ca_cfg, want_sorted_idx=True)
        self._dt_in_use = self.dt_nominal
        self.agents = None
        self.default_agents = None
        self.prev_episode_agents = None
        self.static_map_filename = None
        self.map = None
        self.episode_step_number = None
        self.episode_number = 0
        self.plot_save_dir = None
        self.plot_policy_name = None
        self.perturbed_obs = None
        self.test_case_index = 0
        self._rng = np.random.default_rng()
        self.set_testcase(cfg.TEST_CASE_FN, dict(cfg.TEST_CASE_ARGS))
        self.observation = {}
        self._zero_observation()

    @property
    def unwrapped(self):
        return self

    # ---- reference API
    def set_agents(self, agents):
        self.default_agents = agents

    def set_testcase(self, test_case_fn_str, test_case_args):
        """GCA/envs/collision_avoidance_env.py:294-298: the generator reset() calls when no agents were injected.
        Available: get_testcase_random (test_cases.py:95-118) and get_testcase_two_agents (:77-84); the preset suites
        (:176-512) read pickled test sets and use CADRL / RVO policies, which are out of scope — build Agents and call
        set_agents() for those."""
        if test_case_fn_str not in ("get_testcase_random", "get_testcase_two_agents"):
            raise NotImplementedError("test case generator %r is not available on the GPU path; build Agents and call "
                                      "set_agents() instead" % (test_case_fn_str,))
        self.test_case_fn_str = test_case_fn_str
        self.test_case_args = dict(test_case_args or {})

    def set_static_map(self, map_filename):
        self.static_map_filename = map_filename

    def set_plot_save_dir(self, plot_save_dir):
        self.plot_save_dir = plot_save_dir

    def set_perturbed_info(self, perturbed_obs):
        self.perturbed_obs = perturbed_obs

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def reset(self):
        cfg = get_config()
        if self.episode_step_number is not None and self.episode_step_number > 0:
            self.episode_number += 1
        self.episode_step_number = 0
        if self.default_agents is None:
            self.agents = self._two_agents() if self.test_case_fn_str == "get_testcase_two_agents" else self._random_agents()
        else:
            self.agents = self.default_agents
        if len(self.agents) > cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT:
            raise ValueError("%d agents > MAX_NUM_AGENTS_IN_ENVIRONMENT=%d" % (len(self.agents), cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT))
        for agent in self.agents:
            agent.max_heading_change = self.max_heading_change
            agent.max_speed = self.max_speed
        A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
        init = np.zeros((1, A, _abi.INIT_STRIDE))
        for i, a in enumerate(self.agents):
            init[0, i] = a._init_row()
        self._env.set_world_state(init, np.array([len(self.agents)], dtype=np.int32))
        obs = self._env.reset()
        self._sync_agents(moved=None)
        return self._obs_dict(obs[0])

    def step(self, actions, dt=None):
        """CollisionAvoidanceEnv.step, :131-194.  `actions`: dict agent index -> discrete action (learning_ga3c)
        or [speed_frac, heading_frac] (learning); agents with internal policies need no entry."""
        dt = self.dt_nominal if dt is None else float(dt)   # :137-138
        if dt != self._dt_in_use:
            self._env.handle.set_dt(dt)
            self._dt_in_use = dt
        self.episode_step_number += 1
        A = self._env.A
        act = np.zeros((1, A), dtype=np.int32)
        cont = None
        for i, agent in enumerate(self.agents):
            if agent.is_done or not agent.policy.is_external:
                continue
            a = actions[i]  # KeyError for a missing external action, like the reference (:245)
            if agent.policy.ca_policy == _abi.POLICY_LEARNING_GA3C:
                act[0, i] = int(a)
            else:
                if cont is None:
                    cont = np.zeros((1, A, 2))
                    cont[..., 1] = 0.5
                cont[0, i] = np.asarray(a, dtype=np.float64)[:2]
        moved = [not a.is_done for a in self.agents]
        obs, rew, done, over = self._env.step(act, cont)
        self._sync_agents(moved=moved, dt=dt)
        n = len(self.agents)
        rewards = rew[0, :n].astype(np.float64)
        if get_config().TRAIN_SINGLE_AGENT:
            rewards = rewards[0]
        info = {'which_agents_done': {a.id: np.bool_(done[0, i]) for i, a in enumerate(self.agents)},
                'which_agents_learning': {a.id: a.policy.is_still_learning for a in self.agents}}
        return self._obs_dict(obs[0]), rewards, bool(over[0]), info

    def close(self):
        self._env.close()

    # ---- helpers
    def _zero_observation(self):
        cfg = get_config()
        for i in range(cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT):
            self.observation[i] = {s: np.zeros(cfg.STATE_INFO_DICT[s]['size'], dtype=cfg.STATE_INFO_DICT[s]['dtype'])
                                   for s in cfg.STATES_IN_OBS}

    def _obs_dict(self, rows):
        M = self._env.M
        for i in range(self._env.A):
            r = rows[i]
            if i < len(self.agents):
                self.observation[i] = {
                    'is_learning': np.array(bool(r[0])), 'num_other_agents': np.array(int(r[1])),
                    'dist_to_goal': np.array(r[2]), 'heading_ego_frame': np.array(r[3]), 'pref_speed': np.array(r[4]),
                    'radius': np.array(r[5]), 'other_agents_states': r[6:].reshape(M, 7).copy()}
        return self.observation

    def _sync_agents(self, moved, dt=0.0):
        st = self._env.get_state()[0]
        for i, a in enumerate(self.agents):
            a._read_back(st[i], self._env.obs[0, i], False if moved is None else moved[i], dt)

    def _two_agents(self):
        """get_testcase_two_agents (GCA/envs/test_cases.py:77-84): two agents swapping corners of a 6 m square."""
        policies = self.test_case_args.get('policies', ['learning', 'learning_ga3c'])
        g = 3
        return [Agent(-g, -g, g, g, 0.5, 1.0, 0.0, policy_dict[policies[0]], UnicycleDynamics, [OtherAgentsStatesSensor], 0),
                Agent(g, g, -g, -g, 0.5, 1.0, np.pi, policy_dict[policies[1]], UnicycleDynamics, [OtherAgentsStatesSensor], 1)]

    def _random_agents(self):
        """get_testcase_random (GCA/envs/test_cases.py:95-118) for one world, via scenarios.random_worlds."""
        from .scenarios import random_worlds
        cfg = get_config()
        args = self.test_case_args
        A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
        n = args.get('num_agents') or int(self._rng.integers(2, A + 1))
        policies = args.get('policies', 'learning')
        policies = [policies] if isinstance(policies, str) else list(policies)
        init, _ = random_worlds(1, A, self._rng, num_agents=n, speed_bnds=args.get('speed_bnds', (0.5, 2.0)),
                                radius_bnds=args.get('radius_bnds', (0.2, 0.8)), policies=policies,
                                policy_distr=args.get('policy_distr'), policy_to_ensure=args.get('policy_to_ensure'))
        inv = {v: k for k, v in _abi.POLICY_IDS.items()}
        agents = []
        for i in range(n):
            r = init[0, i]
            agents.append(Agent(r[_abi.I_PX], r[_abi.I_PY], r[_abi.I_GX], r[_abi.I_GY], r[_abi.I_RADIUS],
                                r[_abi.I_PREF_SPEED], r[_abi.I_HEADING], policy_dict[inv[int(r[_abi.I_POLICY])]],
                                UnicycleDynamics, [OtherAgentsStatesSensor], i))
        return agents
