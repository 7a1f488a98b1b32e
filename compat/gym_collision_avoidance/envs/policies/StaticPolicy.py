from rl_collision_avoidance_b200.env import StaticPolicy  # noqa: F401
