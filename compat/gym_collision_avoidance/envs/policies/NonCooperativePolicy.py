from rl_collision_avoidance_b200.env import NonCooperativePolicy  # noqa: F401
