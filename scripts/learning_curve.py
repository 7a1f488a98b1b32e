#!/usr/bin/env python
"""Learning curve of the GPU GA3C loop (Server.main) from a random initialisation: rolling episode score vs episodes,
optimiser steps and wall time.  One GPU, TrainPhase1 by default.

    python scripts/learning_curve.py --seconds 120 --worlds 16384 --batch 16384 --lr-scale sqrt --out gpurun_out/curve.json
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--worlds", type=int, default=16384)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--lr-scale", default="sqrt")
    ap.add_argument("--lr-max", type=float, default=3e-3)
    ap.add_argument("--beta", type=float, default=None)
    ap.add_argument("--config", default="TrainPhase1")
    ap.add_argument("--tf32", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    os.environ["GYM_CONFIG_CLASS"] = a.config
    os.environ["GA3C_GPU_NUM_WORLDS"] = str(a.worlds)
    os.environ["GA3C_GPU_TRAIN_BATCH"] = str(a.batch)
    os.environ["GA3C_GPU_LR_SCALE"] = a.lr_scale
    os.environ["GA3C_GPU_LR_MAX"] = str(a.lr_max)
    os.environ["GA3C_GPU_TRAIN_TF32"] = str(a.tf32)
    os.environ.setdefault("GA3C_CHECKPOINT_DIR", "/tmp/ga3c_curve_ckpt")
    from rl_collision_avoidance_b200.ga3c.Config import get_config
    from rl_collision_avoidance_b200.ga3c.Server import Server
    cfg = get_config()
    cfg.SAVE_MODELS = False
    cfg.EPISODES = 10 ** 12
    if a.beta is not None:
        cfg.BETA_START = cfg.BETA_END = a.beta
    srv = Server(cfg)
    res = srv.main(max_seconds=a.seconds, quiet=True)
    hist = srv.history
    rec = {"config": a.config, "worlds": a.worlds, "batch_rows": srv.train_batch_rows(), "lr_scale": a.lr_scale,
           "lr_multiplier": srv.lr_multiplier(), "learning_rate": srv.model.learning_rate, "beta": cfg.BETA_START,
           "result": res,
           "curve": [{"s": round(h[0], 2), "episodes": h[1], "rscore": round(h[2], 4), "opt_steps": h[3], "frames": h[4]}
                     for h in hist[:: max(1, len(hist) // 60)]]}
    if hist:
        rec["rscore_first"], rec["rscore_last"], rec["rscore_max"] = hist[0][2], hist[-1][2], max(h[2] for h in hist)
    line = json.dumps(rec)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")
    print(json.dumps({k: rec[k] for k in rec if k != "curve"}))
    for c in rec["curve"][:: max(1, len(rec["curve"]) // 20)]:
        print(c)


if __name__ == "__main__":
    main()
