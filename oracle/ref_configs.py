"""TEST INFRASTRUCTURE — extra Config subclasses for golden-vector generation.

Loaded by the reference's own config mechanism (GYM_CONFIG_PATH / GYM_CONFIG_CLASS,
GCA/envs/__init__.py:1-13) so that settings are final before the sensor class binds its
defaults (GCA/envs/sensors/OtherAgentsStatesSensor.py:14).  Only subclasses the
reference's Config; contains no reference code.
"""
from gym_collision_avoidance.envs.config import Config as EnvConfig
from gym_collision_avoidance.envs.config import EvaluateConfig as EnvEvaluateConfig


class ClosestLast4(EnvConfig):
    def __init__(self):
        EnvConfig.__init__(self)
        self.AGENT_SORTING_METHOD = "closest_last"


class TimeToImpact4(EnvConfig):
    def __init__(self):
        EnvConfig.__init__(self)
        self.AGENT_SORTING_METHOD = "time_to_impact"


class Clip6Obs3(EnvConfig):
    """6 agents in the world but only the 3 closest are observed (exercises the clip)."""

    def __init__(self):
        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 6
        self.MAX_NUM_AGENTS_TO_SIM = 6
        self.MAX_NUM_OTHER_AGENTS_OBSERVED = 3
        EnvConfig.__init__(self)


class Clip6Obs3ClosestLast(Clip6Obs3):
    def __init__(self):
        Clip6Obs3.__init__(self)
        self.AGENT_SORTING_METHOD = "closest_last"


class Evaluate19(EnvEvaluateConfig):
    """EVALUATE_MODE: DT=0.1, game_over = all agents done, A_max = 19."""

    def __init__(self):
        EnvEvaluateConfig.__init__(self)


class SingleAgent4(EnvConfig):
    def __init__(self):
        EnvConfig.__init__(self)
        self.TRAIN_SINGLE_AGENT = True
