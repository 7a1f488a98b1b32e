"""GPU: throughput of the full GA3C actor->predictor loop (BASELINE configs[2]: 10 agents x 16384 worlds with the
NetworkVP LSTM forward; configs[1] size as a second point).  Reports env-only, predictor-only and full rollout rates.
Not the bench.py line (that is the env.step hot path); numbers are quoted in DESIGN.md §6."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from rl_collision_avoidance_b200.ga3c import Config as cfgmod
from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
from rl_collision_avoidance_b200.ga3c.rollout import GpuRollout
from rl_collision_avoidance_b200.scenarios import random_worlds


def run(cls, W, steps, tf32):
    cfg = getattr(cfgmod, cls)()
    cfgmod.set_config(cfg)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
    rng = np.random.default_rng(0)
    init, nag = random_worlds(W, A, rng)
    model = NetworkVP_rnn("cuda:0", "network", 11, seed=0)
    ro = GpuRollout(cfg, model, W, init, nag, device=0, seed=1)
    sc = ro.env.scenario_config(cfg.TEST_CASE_ARGS)
    ro.env.generate_scenarios(sc, 5, only_consumed=False)
    for _ in range(5):
        ro.step(); ro.env.generate_scenarios(sc, 5, only_consumed=True); ro.rec.out_count.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = 0
    for k in range(steps):
        ro.step()
        ro.env.generate_scenarios(sc, 5, only_consumed=True)
        if k % 8 == 7:
            rows += ro.rec.take()[0].shape[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # predictor alone
    x = ro.rec.obs_slot(0).reshape(W * A, -1)[:, 1:]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(20):
        model.predict_p_and_v_device(x)
    torch.cuda.synchronize(); dp = (time.perf_counter() - t1) / 20
    print("%s W=%d A=%d tf32=%s: rollout %.2f ms/step = %.1f M agent-steps/s (%.1f M learner rows/s emitted); predictor alone "
          "%.2f ms/batch of %d rows" % (cls, W, A, tf32, 1e3 * dt / steps, W * A * steps / dt / 1e6, rows / dt / 1e6, 1e3 * dp, W * A),
          flush=True)
    ro.close()
    cfgmod.set_config(None)


if __name__ == "__main__":
    run("TrainPhase2", 16384, 60, False)
    run("TrainPhase2", 16384, 60, True)
    run("TrainPhase1", 65536, 60, False)
    run("TrainPhase1", 65536, 60, True)
