from rl_collision_avoidance_b200.ga3c.Server import Server  # noqa: F401
