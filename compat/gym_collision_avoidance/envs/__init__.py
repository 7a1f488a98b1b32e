from rl_collision_avoidance_b200 import env as _env
Config = _env.get_config()
