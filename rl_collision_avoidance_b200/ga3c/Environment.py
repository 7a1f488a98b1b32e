"""`Environment(id)` — the adapter GA3C's actors hold (GA3C/Environment.py:37-116), over this package's env.

The reference's ProcessAgent talks to the simulator only through this class (ProcessAgent.py:107,128-149):
    env.reset();  env.latest_observations  -> (A, L) rows, column 0 = is_learning
    rewards, game_over, infos = env.step([actions_dict], pid, count)
    env.previous_state / env.current_state -> (1, A, L-1) network inputs before / after the step
With it a reference ProcessAgent runs unchanged against the CUDA environment (one world per actor, host buffers
through the C-ABI).  The vectorised on-device loop (ga3c/rollout.py) does not go through this class: it keeps the same
quantities for W worlds at once in its observation ring.
"""
import numpy as np

from .Config import get_config


class Environment(object):
    def __init__(self, id):
        self.id = id
        self._set_env(id)
        self.nb_frames = 1                # the reference keeps a queue of the last nb_frames = 1 frames (:41-42)
        self.total_reward = 0
        self.latest_observations = None
        self.previous_state = self.current_state = None

    def _set_env(self, id):
        cfg = get_config()
        if cfg.GAME_CHOICE != cfg.game_collision_avoidance:
            raise ValueError("[ ERROR ] Invalid choice of game. Check Config.py for choices")
        from .. import env as _env
        from ..config import to_ca_config  # noqa: F401  (fails early if the package is not importable)
        _env.set_config(cfg)
        one_env = _env.CollisionAvoidanceEnv()
        one_env.id = id
        self.game = _SingleWorldVecEnv(one_env)

    def _process_obs(self, observations):
        obs = observations[0]                      # undo the VecEnv axis (:85)
        if obs.ndim == 3:
            obs = obs[0]                           # undo the multi-agent VecEnv wrapper (:86-87)
        self.latest_observations = obs
        self.previous_state = self.current_state
        self.current_state = np.array([obs[:, 1:]])   # one frame in the queue (:67-68,89-91)

    def reset(self):
        self.total_reward = 0
        self._process_obs(self.game.reset())

    def step(self, action, pid, count):
        observations, rewards, game_over, info = self.game.step(action)
        self.total_reward += np.sum(rewards[0])
        self._process_obs(observations)
        return rewards, game_over, info

    def print_frame_q(self):
        return 0 if self.current_state is None else self.nb_frames


class _SingleWorldVecEnv(object):
    """num_envs = 1 stand-in for MultiagentDummyVecEnv (GCA/envs/wrappers.py:104-109 over baselines DummyVecEnv): obs
    (1, A, L) float32, rewards / done / info wrapped in a length-1 axis, auto-reset on game_over."""

    def __init__(self, one_env):
        self.envs = [one_env]
        self.num_envs = 1
        self._env = one_env

    def reset(self):
        self._env.reset()
        return self._env._env.obs.copy()

    def step(self, actions):
        _, rewards, game_over, info = self._env.step(actions[0])
        if game_over:
            self._env.reset()
        rews = np.empty((1,), dtype=object)
        rews[0] = rewards
        return self._env._env.obs.copy(), rews, np.array([game_over]), [info]
