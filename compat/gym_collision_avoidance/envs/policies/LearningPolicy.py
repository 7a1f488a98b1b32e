from rl_collision_avoidance_b200.env import LearningPolicy  # noqa: F401
