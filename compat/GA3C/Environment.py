from rl_collision_avoidance_b200.ga3c.Environment import Environment  # noqa: F401
