"""create_env with the return convention of GCA/experiments/src/env_utils.py:8-43: (vec_env, one_env)."""
from rl_collision_avoidance_b200 import env as _env
from rl_collision_avoidance_b200.ga3c.Environment import _SingleWorldVecEnv


def create_env():
    one_env = _env.CollisionAvoidanceEnv()
    one_env.id = 0
    return _SingleWorldVecEnv(one_env), one_env
