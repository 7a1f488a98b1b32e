"""The scenario helpers of GCA/envs/test_cases.py that only need the supported policies."""
import numpy as np

from rl_collision_avoidance_b200.env import (Agent, OtherAgentsStatesSensor, UnicycleDynamics, policy_dict)  # noqa: F401


def get_testcase_two_agents(policies=('learning_ga3c', 'learning_ga3c')):
    """Geometry of GCA/envs/test_cases.py:77-84."""
    gx = gy = 3
    return [Agent(-gx, -gy, gx, gy, 0.5, 1.0, 0.0, policy_dict[policies[0]], UnicycleDynamics, [OtherAgentsStatesSensor], 0),
            Agent(gx, gy, -gx, -gy, 0.5, 1.0, np.pi, policy_dict[policies[1]], UnicycleDynamics, [OtherAgentsStatesSensor], 1)]
