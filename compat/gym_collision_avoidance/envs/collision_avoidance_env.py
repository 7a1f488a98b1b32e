from rl_collision_avoidance_b200.env import CollisionAvoidanceEnv  # noqa: F401
