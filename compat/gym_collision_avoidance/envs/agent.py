from rl_collision_avoidance_b200.env import Agent  # noqa: F401
