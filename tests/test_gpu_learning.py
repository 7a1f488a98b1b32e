"""The GPU GA3C loop LEARNS with the default trainer recipe (Config.GPU_TRAIN_BATCH = 8192 rows per optimiser step,
learning rate scaled by sqrt(batch / 128), see Config.GPU_LR_SCALE): from a random initialisation the rolling episode
score of TrainPhase1 rises by more than 0.3 within a bounded number of optimiser steps (reference: rolling reward
-0.05 -> 0.90-0.95 over 1.5 M episodes from a regression-initialised network, README.md:24 and
ga3c/GA3C/checkpoints/RL/wandb/run-2018-backup/checkpoints/index.txt).  The committed curve of a longer run is
profiles/r02_learning_curve.json."""
import os

import pytest

pytestmark = pytest.mark.gpu


def test_trainphase1_score_rises_from_random_init(monkeypatch, tmp_path):
    monkeypatch.setenv("GYM_CONFIG_CLASS", "TrainPhase1")
    monkeypatch.setenv("GA3C_GPU_NUM_WORLDS", "4096")
    monkeypatch.setenv("GA3C_CHECKPOINT_DIR", str(tmp_path))
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    from rl_collision_avoidance_b200.ga3c.Server import Server
    cfgmod.set_config(None)
    cfg = cfgmod.get_config()
    cfg.SAVE_MODELS = False
    cfg.EPISODES = 10 ** 12
    try:
        srv = Server(cfg)
        assert srv.train_batch_rows() == 8192 and abs(srv.lr_multiplier() - 8.0) < 1e-9
        srv.main(max_steps=200, quiet=True)
        start = srv.history[0][2] if srv.history else srv.stats.roll_reward_log
        res = srv.main(max_seconds=150, until_score=start + 0.45, quiet=True)
        best = max(h[2] for h in srv.history)
        print("rolling score %.3f -> %.3f (best %.3f) after %d optimiser steps, %d episodes, %.0f s"
              % (start, srv.stats.roll_reward_log, best, srv.training_step, res["episodes"], res["seconds"]))
        assert start < 0.1, "a random policy does not reach goals"
        assert best >= start + 0.3, "rolling score did not rise: %.3f -> best %.3f" % (start, best)
        assert srv.training_step <= 120000
        srv.rollout.close()
    finally:
        cfgmod.set_config(None)
