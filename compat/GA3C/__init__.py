from rl_collision_avoidance_b200.ga3c.Config import get_config as _get
Config = _get()
