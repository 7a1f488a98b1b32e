from rl_collision_avoidance_b200.env import LearningPolicyGA3C  # noqa: F401
