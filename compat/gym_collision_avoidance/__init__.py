# gym id 'CollisionAvoidance-v0' (GCA/__init__.py:6-9) is registered when gym is installed
try:
    from gym.envs.registration import register
    register(id='CollisionAvoidance-v0', entry_point='gym_collision_avoidance.envs.collision_avoidance_env:CollisionAvoidanceEnv')
except Exception:
    pass
