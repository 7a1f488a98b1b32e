"""The GA3C actor -> predictor loop, vectorised on one GPU.

Replaces, for W worlds at once:
  ProcessAgent.run_episode / predict / select_action / _accumulate_rewards   GA3C/ProcessAgent.py:54-211
  ThreadPredictor.run (dynamic batching of <=128 requests)                    GA3C/ThreadPredictor.py:40-75
  Environment (VecEnv adapter, previous_state/current_state)                  GA3C/Environment.py:37-116
Every env step is: the row plan of the learning agents (ca_predict_plan) -> ONE fused predictor launch (NetworkVP forward +
action sampling, argmax in PLAY/EVALUATE mode; ca_predict_rows) -> one fused env.step launch that writes the next
observation straight into the rollout's observation ring -> the bookkeeping launches (ca_ga3c_record: experience lists,
n-step returns, then the row gather) that emit the training rows (x_, r_, a_) the reference's generator would yield at
this step; the scenario refresh of the worlds that just reset runs on a side stream behind the env step.
Nothing leaves the GPU; torch is the allocator / stream library (and the network under GA3C_PREDICTOR=composed).
"""
import ctypes as C

import numpy as np

from .. import _abi
from .._lib import check, lib
from ..vec_env import VecCollisionAvoidanceEnv


class ExperienceRecorder(object):
    """Device-side experience lists + n-step returns for N = W*A agent slots (ca_ga3c_record)."""

    def __init__(self, num_worlds, agents_per_world, obs_len, time_max, gamma, device, capacity=None):
        import torch
        self.torch = torch
        self.W, self.A, self.L = num_worlds, agents_per_world, obs_len
        self.N = num_worlds * agents_per_world
        self.time_max, self.gamma = int(time_max), float(gamma)
        self.R = self.time_max + 2
        self.device = torch.device(device)
        dev = self.device
        self.obs_ring = torch.zeros((self.R, self.W, self.A, self.L), dtype=torch.float32, device=dev)
        self.act_ring = torch.zeros((self.R, self.N), dtype=torch.int32, device=dev)
        self.rew_ring = torch.zeros((self.R, self.N), dtype=torch.float32, device=dev)
        self.length = torch.zeros(self.N, dtype=torch.int32, device=dev)
        self.tcount = torch.zeros(self.N, dtype=torch.int32, device=dev)
        self.done_trained = torch.zeros(self.N, dtype=torch.uint8, device=dev)
        # worst case: every agent flushes TIME_MAX+1 rows in the same step (synchronised start)
        self.capacity = int(capacity) if capacity else self.N * (self.time_max + 1)
        self.out_x = torch.empty((self.capacity, self.L - 1), dtype=torch.float32, device=dev)
        self.out_r = torch.empty(self.capacity, dtype=torch.float32, device=dev)
        self.out_a = torch.empty(self.capacity, dtype=torch.int32, device=dev)
        self.out_src = torch.empty(self.capacity, dtype=torch.int32, device=dev)
        self._counters = torch.zeros(4, dtype=torch.int32, device=dev)   # [0] rows emitted, [1..2] gather watermark / ticket
        self.out_count = self._counters[0:1]
        self.ep_reward = torch.zeros(self.W, dtype=torch.float32, device=dev)
        self.ep_steps = torch.zeros(self.W, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(3, dtype=torch.float64, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        self._bufs = _abi.CaGa3cBuffers(p(self.obs_ring), p(self.act_ring), p(self.rew_ring), p(self.length), p(self.tcount),
                                        p(self.done_trained), p(self.out_x), p(self.out_r), p(self.out_a), p(self.out_count),
                                        self.capacity, 0, p(self.out_src), C.c_void_p(self._counters.data_ptr() + 4))

    def obs_slot(self, t):
        return self.obs_ring[t % self.R]

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def record(self, t, actions, values, reward, done, game_over):
        """actions int32 [N], values/reward float32 [N], done uint8 [N], game_over uint8 [W] (device tensors)."""
        p = lambda x: C.c_void_p(x.data_ptr())
        dev_index = self.device.index or 0
        check(lib().ca_ga3c_episode_stats(p(self.obs_slot(t)), p(reward), p(game_over), p(self.ep_reward), p(self.ep_steps),
                                          p(self.stats), self.W, self.A, self.L, dev_index, self._stream()),
              "ca_ga3c_episode_stats")
        check(lib().ca_ga3c_record(C.byref(self._bufs), int(t), self.R, self.N, self.A, self.L, self.time_max, self.gamma,
                                   p(actions), p(values), p(reward), p(done), p(game_over), dev_index, self._stream()),
              "ca_ga3c_record")

    def take(self):
        """Returns views (x_ [k, L-1], r_ [k], a_ [k]) of the rows emitted since the last take() and resets the counter.
        Synchronises on the row count (one 4-byte D2H read).  Call it after every step (or at least before the rows of
        several steps can exceed `capacity` = one step's worst case, N * (TIME_MAX + 1)); overflow raises instead of
        silently dropping rows."""
        k = int(self.out_count.item())
        if k > self.capacity:
            raise RuntimeError("experience output overflow: %d rows emitted, capacity %d" % (k, self.capacity))
        self._counters.zero_()
        return self.out_x[:k], self.out_r[:k], self.out_a[:k]

    def discard(self):
        """Drop the rows emitted since the last take() without reading them."""
        self._counters.zero_()

    def pop_stats(self):
        s = self.stats.cpu().numpy().copy()
        self.stats.zero_()
        return {"episodes": int(s[0]), "score_sum": float(s[1]), "frames": int(s[2])}


class GpuRollout(object):
    """W worlds x A agents stepped in lock-step by one policy network on one GPU."""

    def __init__(self, cfg, model, num_worlds, init, num_agents, device=0, seed=0):
        import torch
        from ..config import to_ca_config
        self.torch = torch
        self.cfg, self.model = cfg, model
        self.device = torch.device("cuda", device)
        ca_cfg = to_ca_config(cfg, num_worlds, device=device, auto_reset=1)
        self.env = VecCollisionAvoidanceEnv(ca_cfg)
        self.W, self.A, self.L = self.env.W, self.env.A, self.env.L
        self.N = self.W * self.A
        self.rec = ExperienceRecorder(self.W, self.A, self.L, cfg.TIME_MAX, cfg.DISCOUNT, self.device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        self.greedy = bool(cfg.PLAY_MODE or cfg.EVALUATE_MODE)
        self.seed = int(seed)
        # predictor: the fused tcgen05 kernel (forward pass + action selection in one launch) unless GA3C_PREDICTOR=composed
        # asks for the fp32 path (ca_lstm_step kernel + cuBLAS dense layers + torch.multinomial)
        import os
        self.fused = os.environ.get("GA3C_PREDICTOR", "fused") != "composed" and \
            hasattr(model, "fused_supported") and model.fused_supported()
        # only learning agents ask the predictor (ProcessAgent.py:128-133); GA3C_PREDICT_ALL=1 predicts every slot
        self.learning_only = os.environ.get("GA3C_PREDICT_ALL", "0") != "1"
        self.t = 0
        self.env.set_world_state(init, num_agents)
        self.env.reset(out_obs=self.rec.obs_slot(0))
        self.last_actions = None
        self.last_values = None
        self._gen_cfg = None
        self._gen_done = None

    def attach_scenario_generator(self, scenario_cfg, seed):
        """After every env step, the worlds that consumed their reset snapshot are handed a fresh test case
        (ca_generate_scenarios ≙ test_case_fn(**TEST_CASE_ARGS) on env.reset(), collision_avoidance_env.py:283).  The
        generator runs on a side stream right behind the env step and overlaps the experience bookkeeping of the same
        step; the next env step waits for it."""
        torch = self.torch
        self._gen_cfg, self._gen_seed = scenario_cfg, int(seed)
        self._gen_stream = torch.cuda.Stream(device=self.device)
        self._gen_after_step = torch.cuda.Event()
        self._gen_done = None

    def step(self):
        """One env step for every world; returns (reward, done, game_over) device tensors of this step."""
        torch = self.torch
        obs = self.rec.obs_slot(self.t)                           # [W, A, L]; column 0 = is_learning
        if self.fused:   # ThreadPredictor + select_action: one launch over all slots
            _, v, actions = self.model.predict_fused(obs.reshape(self.N, self.L), want_p=False, want_actions=True,
                                                     greedy=self.greedy, seed=self.seed, learning_only=self.learning_only)
        else:
            p, v = self.model.predict_from_obs(obs.reshape(self.N, self.L))
            if self.greedy:
                actions = torch.argmax(p, dim=1).to(torch.int32)      # ProcessAgent.select_action (:98-103)
            else:
                actions = torch.multinomial(p, 1, generator=self.gen).squeeze(1).to(torch.int32)
        main = torch.cuda.current_stream(self.device)
        if self._gen_done is not None:
            main.wait_event(self._gen_done)       # the snapshots the step may adopt are complete
        _, reward, done, over = self.env.step(actions.view(self.W, self.A), out_obs=self.rec.obs_slot(self.t + 1))
        if self._gen_cfg is not None:
            self._gen_after_step.record(main)
            self._gen_stream.wait_event(self._gen_after_step)
            with torch.cuda.stream(self._gen_stream):
                self.env.generate_scenarios(self._gen_cfg, self._gen_seed, only_consumed=True)
                if self._gen_done is None:
                    self._gen_done = torch.cuda.Event()
                self._gen_done.record(self._gen_stream)
        self.rec.record(self.t, actions, v.contiguous(), reward.view(-1), done.view(-1), over)
        self.last_actions, self.last_values = actions, v
        self.t += 1
        return reward, done, over

    def close(self):
        if self._gen_done is not None:
            self._gen_done.synchronize()
        self.env.close()
