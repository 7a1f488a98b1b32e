from rl_collision_avoidance_b200.env import OtherAgentsStatesSensor  # noqa: F401
