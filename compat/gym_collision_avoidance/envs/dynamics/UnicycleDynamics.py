from rl_collision_avoidance_b200.env import UnicycleDynamics  # noqa: F401
