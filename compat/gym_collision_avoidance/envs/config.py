from rl_collision_avoidance_b200.config import Config, EvaluateConfig, Example  # noqa: F401
