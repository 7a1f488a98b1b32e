"""create_env with the return convention of GCA/experiments/src/env_utils.py:8-43: (vec_env, one_env)."""
import numpy as np

from rl_collision_avoidance_b200 import env as _env
from rl_collision_avoidance_b200.config import to_ca_config
from rl_collision_avoidance_b200.vec_env import HostVecEnv


class _SingleWorldVecEnv(object):
    """num_envs = 1 stand-in for MultiagentDummyVecEnv: obs (1, A, L) float32, auto-reset on game_over."""

    def __init__(self, one_env):
        self.envs = [one_env]
        self.num_envs = 1
        self._env = one_env

    def reset(self):
        self._env.reset()
        return self._env._env.obs.copy()

    def step(self, actions):
        _, rewards, game_over, info = self._env.step(actions[0])
        if game_over:
            self._env.reset()
        rews = np.empty((1,), dtype=object)
        rews[0] = rewards
        return self._env._env.obs.copy(), rews, np.array([game_over]), [info]


def create_env():
    one_env = _env.CollisionAvoidanceEnv()
    one_env.id = 0
    return _SingleWorldVecEnv(one_env), one_env
