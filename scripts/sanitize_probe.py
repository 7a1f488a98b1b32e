"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel variant, ragged worlds,
auto-reset with streamed scenarios, the generator, and (argument `rollout`, or no argument = everything) the GA3C loop:
row plan + fused tcgen05 predictor, env step, experience bookkeeping + row gather, side-stream scenario refresh, and one
optimiser step through the fused LSTM cell kernels.
    python scripts/sanitize_probe.py [step|rollout]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rl_collision_avoidance_b200 import _abi
from rl_collision_avoidance_b200.vec_env import HostVecEnv, VecCollisionAvoidanceEnv
from rl_collision_avoidance_b200.scenarios import random_worlds

rng = np.random.default_rng(0)
PART = sys.argv[1] if len(sys.argv) > 1 else "all"
for kern in (("oneshot", "stream", "generic") if PART in ("all", "step") else ()):
    os.environ["CA_STEP_KERNEL"] = kern
    for A, W in ((4, 531), (10, 77), (3, 100)):
        init, nag = random_worlds(W, A, rng, num_agents=rng.integers(2, A + 1, W), policies=['noncoop', 'learning_ga3c', 'static'],
                                  policy_distr=[0.2, 0.6, 0.2], policy_to_ensure='learning_ga3c')
        env = HostVecEnv(_abi.default_config(W, A, auto_reset=1), want_sorted_idx=(kern != "oneshot"))
        env.set_world_state(init, nag)
        env.reset()
        for t in range(12):
            env.step(rng.integers(0, 11, (W, A)).astype(np.int32))
        env.get_state()
        env.close()
    print(kern, "ok", flush=True)
os.environ["CA_STEP_KERNEL"] = "oneshot"
if PART == "step":
    sys.exit(0)
env = VecCollisionAvoidanceEnv(_abi.default_config(300, 4, auto_reset=1))
sc = env.scenario_config({'policies': ['noncoop', 'learning_ga3c', 'static'], 'policy_distr': [0.05, 0.9, 0.05],
                          'policy_to_ensure': 'learning_ga3c'})
env.generate_scenarios(sc, 3)
env.reset()
for t in range(30):
    env.step(torch.randint(0, 11, (300, 4), dtype=torch.int32, device="cuda"))
    env.generate_scenarios(sc, 3, only_consumed=True)
torch.cuda.synchronize()
print("generator ok", flush=True)
from rl_collision_avoidance_b200.ga3c import Config as cfgmod
from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
from rl_collision_avoidance_b200.ga3c.rollout import GpuRollout
for cls, W in (("TrainPhase1", 200), ("TrainPhase2", 90)):
    cfg = getattr(cfgmod, cls)(); cfgmod.set_config(cfg)
    A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
    init, nag = random_worlds(W, A, rng)
    net = NetworkVP_rnn("cuda:0", "network", 11)
    ro = GpuRollout(cfg, net, W, init, nag)
    sc = ro.env.scenario_config(cfg.TEST_CASE_ARGS)
    ro.env.generate_scenarios(sc, 3, only_consumed=False)
    ro.env.reset(out_obs=ro.rec.obs_slot(ro.t))
    ro.attach_scenario_generator(sc, 3)
    rows = []
    for t in range(30):
        ro.step()
        x, r, a = ro.rec.take()      # the recorder is drained every step (capacity = one step's worst case)
        if x.shape[0]:
            rows.append((x.clone(), r.clone(), a.clone()))
    x = torch.cat([q[0] for q in rows]); r = torch.cat([q[1] for q in rows]); a = torch.cat([q[2] for q in rows])
    net.train(x, r, a)               # fused LSTM cell forward / backward kernels
    torch.cuda.synchronize()
    ro.close()
    cfgmod.set_config(None)
    print(cls, "rollout + optimiser step ok,", int(x.shape[0]), "rows", flush=True)
