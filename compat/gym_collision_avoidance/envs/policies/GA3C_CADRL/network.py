from rl_collision_avoidance_b200.ga3c.Server import Actions  # noqa: F401
