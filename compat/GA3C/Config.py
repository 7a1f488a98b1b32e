from rl_collision_avoidance_b200.ga3c.Config import Train, TrainPhase1, TrainPhase2  # noqa: F401
