"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference (mit-acl/rl_collision_avoidance) from
/root/reference so that golden input/output vectors can be generated from the
reference itself (oracle/gen_golden.py).  /root/reference only exists in the
build container, never on the GPU box, so nothing in tests/-m gpu, smoke() or
bench.py imports this module.

Third-party modules the reference imports but never uses on the step() path
(gym, matplotlib, imageio, moviepy, rvo2, tensorflow, baselines, scipy.misc)
are fabricated as permissive dummies; only gym.Env / gym.spaces.{Box,Dict,
Discrete} / gym.ObservationWrapper / DummyVecEnv get minimal real classes.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("CA_REFERENCE_ROOT", "/root/reference")
GCA_ROOT = os.path.join(REFERENCE_ROOT, "gym-collision-avoidance")
GA3C_ROOT = os.path.join(REFERENCE_ROOT, "ga3c")

_FAKE_ROOTS = ("gym", "matplotlib", "mpl_toolkits", "imageio", "moviepy", "rvo2",
               "tensorflow", "baselines", "wandb")


class _Whatever(object):
    """Object that tolerates any attribute access / call / subscripting."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Whatever()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Whatever()

    def __getitem__(self, k):
        return _Whatever()

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _FakeModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Whatever()


class _FakeFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _FAKE_ROOTS or fullname == "scipy.misc":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _FakeModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _Env(object):
    @property
    def unwrapped(self):
        return self


class _Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)


class _DictSpace(object):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})


class _Discrete(object):
    def __init__(self, n, **kw):
        self.n = n


class _ObservationWrapper(_Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kw):
        return self.observation(self.env.reset(**kw))

    def step(self, action):
        obs, rew, done, info = self.env.step(action)
        return self.observation(obs), rew, done, info


class _DummyVecEnv(object):
    """Restatement of openai/baselines DummyVecEnv (@ea25b9e, pinned by
    GCA/install.sh:26; its source is NOT in /root/reference): step() steps each
    env, and when an env reports done it is reset immediately and the NEW
    episode's first observation is returned alongside the OLD reward/done."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        shape = self.envs[0].observation_space.shape
        self.buf_obs = np.zeros((self.num_envs,) + tuple(shape), dtype=np.float32)
        self.buf_dones = np.zeros((self.num_envs,), dtype=bool)
        self.buf_rews = np.zeros((self.num_envs,), dtype=np.float32)
        self.buf_infos = [{} for _ in range(self.num_envs)]

    def reset(self):
        for e in range(self.num_envs):
            self.buf_obs[e] = self.envs[e].reset()
        return np.copy(self.buf_obs)

    def step(self, actions):
        for e in range(self.num_envs):
            obs, self.buf_rews[e], self.buf_dones[e], self.buf_infos[e] = self.envs[e].step(actions[e])
            if self.buf_dones[e]:
                obs = self.envs[e].reset()
            self.buf_obs[e] = obs
        return (np.copy(self.buf_obs), np.copy(self.buf_rews), np.copy(self.buf_dones),
                list(self.buf_infos))


_installed = False


def install(config_class=None, config_path=None):
    """Make `import gym_collision_avoidance` / `import GA3C` resolve to the reference.

    config_class / config_path select the reference Config subclass exactly as
    train.sh does (env vars GYM_CONFIG_CLASS / GYM_CONFIG_PATH); they must be
    final BEFORE the first import (OtherAgentsStatesSensor.py:14 binds its
    defaults at class-definition time).
    """
    global _installed
    if _installed:
        return
    if not os.path.isdir(GCA_ROOT):
        raise RuntimeError("reference not present at %s (only exists in the build container)" % REFERENCE_ROOT)
    if config_class is not None:
        os.environ["GYM_CONFIG_CLASS"] = config_class
    if config_path is not None:
        os.environ["GYM_CONFIG_PATH"] = config_path
    warnings.filterwarnings("ignore")
    if not hasattr(np, "product"):
        np.product = np.prod          # GA3C/Config.py:69 uses the removed alias
    sys.meta_path.insert(0, _FakeFinder())
    gym = importlib.import_module("gym")
    gym.Env = _Env
    gym.ObservationWrapper = _ObservationWrapper
    spaces = importlib.import_module("gym.spaces")
    spaces.Box, spaces.Dict, spaces.Discrete = _Box, _DictSpace, _Discrete
    gym.spaces = spaces
    importlib.import_module("gym.envs.registration").register = lambda **kw: None
    dve = importlib.import_module("baselines.common.vec_env.dummy_vec_env")
    dve.DummyVecEnv = _DummyVecEnv
    for p in (GCA_ROOT, GA3C_ROOT, os.path.join(GA3C_ROOT, "GA3C")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def reference_modules():
    """Returns a namespace with the reference classes used by the golden generator."""
    ns = types.SimpleNamespace()
    from gym_collision_avoidance.envs import Config
    from gym_collision_avoidance.envs.collision_avoidance_env import CollisionAvoidanceEnv
    from gym_collision_avoidance.envs.agent import Agent
    from gym_collision_avoidance.envs.dynamics.UnicycleDynamics import UnicycleDynamics
    from gym_collision_avoidance.envs.sensors.OtherAgentsStatesSensor import OtherAgentsStatesSensor
    from gym_collision_avoidance.envs.policies.LearningPolicyGA3C import LearningPolicyGA3C
    from gym_collision_avoidance.envs.policies.LearningPolicy import LearningPolicy
    from gym_collision_avoidance.envs.policies.NonCooperativePolicy import NonCooperativePolicy
    from gym_collision_avoidance.envs.policies.StaticPolicy import StaticPolicy
    from gym_collision_avoidance.envs import test_cases
    ns.Config = Config
    ns.CollisionAvoidanceEnv = CollisionAvoidanceEnv
    ns.Agent = Agent
    ns.UnicycleDynamics = UnicycleDynamics
    ns.OtherAgentsStatesSensor = OtherAgentsStatesSensor
    ns.policies = {0: LearningPolicyGA3C, 1: NonCooperativePolicy, 2: StaticPolicy, 3: LearningPolicy}
    ns.test_cases = test_cases
    return ns
