#!/usr/bin/env python
"""Measured error of the fused predictor against the fp32 network (same rows as tests/test_gpu_predictor.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests.test_gpu_predictor import _cfg, _net, _synthetic_obs
from rl_collision_avoidance_b200.config import to_ca_config
from rl_collision_avoidance_b200.scenarios import random_worlds
from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv


def err(net, t_obs):
    p_ref, v_ref = net.predict_p_and_v_device(t_obs[:, 1:])
    p, v, a = net.predict_fused(t_obs, want_p=True, want_actions=True, greedy=True)
    torch.cuda.synchronize()
    dp = (p - p_ref).abs().max(dim=1).values
    dv = (v - v_ref).abs() / (1 + v_ref.abs())
    return "dp max %.2e p99 %.2e | dv max %.2e p99 %.2e" % (dp.max().item(), dp.quantile(0.99).item(), dv.max().item(),
                                                            dv.quantile(0.99).item())


for phase, trained, B in ((1, False, 20000), (1, True, 20000), (2, False, 20000)):
    cfg = _cfg(phase)
    net = _net(trained)
    obs = _synthetic_obs(cfg, B, cfg.MAX_NUM_OTHER_AGENTS_OBSERVED, np.random.default_rng(B))
    print("synthetic phase %d trained %d: %s" % (phase, trained, err(net, torch.from_numpy(obs).cuda())), flush=True)
for phase in (1, 2):
    cfg = _cfg(phase)
    W = 4096
    env = VecCollisionAvoidanceEnv(to_ca_config(cfg, W, device=0, auto_reset=1))
    rng = np.random.default_rng(phase)
    init, nag = random_worlds(W, env.A, rng, num_agents=rng.integers(2, env.A + 1, W))
    env.set_world_state(init, nag)
    obs = env.reset()
    net = _net(trained=(phase == 1))
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(13):
        if t % 4 == 0:
            print("env rows phase %d step %d (trained %d): %s" % (phase, t, phase == 1, err(net, obs.reshape(W * env.A, env.L))), flush=True)
        act = torch.randint(0, 11, (W, env.A), generator=gen, device="cuda", dtype=torch.int32)
        obs, _, _, _ = env.step(act)
    env.close()
