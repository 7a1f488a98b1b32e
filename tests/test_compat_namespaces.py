"""CPU: the reference's import namespaces resolve through compat/ without a GPU (imports only)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_namespaces_importable():
    code = """
import sys
from gym_collision_avoidance.envs import Config
from gym_collision_avoidance.envs.config import Config as EnvConfig
from gym_collision_avoidance.envs.collision_avoidance_env import CollisionAvoidanceEnv
from gym_collision_avoidance.envs.agent import Agent
from gym_collision_avoidance.envs.policies.LearningPolicyGA3C import LearningPolicyGA3C
from gym_collision_avoidance.envs.policies.NonCooperativePolicy import NonCooperativePolicy
from gym_collision_avoidance.envs.policies.StaticPolicy import StaticPolicy
from gym_collision_avoidance.envs.policies.GA3C_CADRL.network import Actions
from gym_collision_avoidance.envs.dynamics.UnicycleDynamics import UnicycleDynamics
from gym_collision_avoidance.envs.sensors.OtherAgentsStatesSensor import OtherAgentsStatesSensor
from gym_collision_avoidance.envs import test_cases as tc
from gym_collision_avoidance.experiments.src.env_utils import create_env
from GA3C import Config as GConfig
assert GConfig.NN_INPUT_SIZE == 26 and GConfig.TIME_MAX == 20 and type(GConfig).__name__ == 'TrainPhase1'
assert Config.MAX_NUM_AGENTS_IN_ENVIRONMENT == 4 and Config.DT == 0.2
assert Actions().num_actions == 11
agents = tc.get_testcase_two_agents()
assert agents[0].policy.str == 'learning' and agents[0].policy.is_external
assert abs(agents[0].time_remaining_to_reach_goal - 2 * (72 ** 0.5 - 0.2)) < 1e-12
sys.path.insert(0, sys.argv[1])
from Server import Server
from NetworkVP_rnn import NetworkVP_rnn
print('ok')
"""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "compat"), REPO, env.get("PYTHONPATH", "")])
    env["GYM_CONFIG_CLASS"] = "TrainPhase1"
    env["GYM_CONFIG_PATH"] = os.path.join(REPO, "rl_collision_avoidance_b200", "ga3c", "Config.py")
    out = subprocess.run([sys.executable, "-c", code, os.path.join(REPO, "compat", "GA3C")], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout
