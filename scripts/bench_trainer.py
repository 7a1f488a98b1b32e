"""GPU: throughput of one A3C optimiser step (NetworkVP_rnn.train: forward + sum-losses + backward + TF-Adam) on rows shaped
like the rollout's training rows, per configuration and batch size.  CUDA events, median of repeats.
    python scripts/bench_trainer.py [--json gpurun_out/trainer.json]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rl_collision_avoidance_b200.ga3c import Config as cfgmod
from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn


def rows(cfg, B, rng):
    M = cfg.MAX_NUM_OTHER_AGENTS_OBSERVED
    L1 = 5 + 7 * M
    avg = np.asarray(cfg.NN_INPUT_AVG_VECTOR, dtype=np.float32)
    std = np.asarray(cfg.NN_INPUT_STD_VECTOR, dtype=np.float32)
    x = (avg + std * rng.normal(size=(B, L1))).astype(np.float32)
    x[:, 0] = rng.integers(1, M + 1, B)
    return (torch.from_numpy(x).cuda(), torch.from_numpy(rng.normal(size=B).astype(np.float32)).cuda(),
            torch.from_numpy(rng.integers(0, 11, B).astype(np.int64)).cuda())


def run(cls, B, fused, tf32=False):
    os.environ["GA3C_FUSED_TRAIN_CELL"] = "1" if fused else "0"
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    cfg = getattr(cfgmod, cls)()
    cfgmod.set_config(cfg)
    net = NetworkVP_rnn("cuda:0", "network", 11, seed=0)
    x, r, a = rows(cfg, B, np.random.default_rng(0))
    for _ in range(3):
        net.train(x, r, a)
    torch.cuda.synchronize()
    times = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            net.train(x, r, a)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 4)
    ms = float(np.median(times))
    mem = torch.cuda.max_memory_allocated() / 2 ** 20
    print("%s B=%d fused_cell=%d tf32_matmul=%d: %.3f ms per optimiser step = %.2f M rows/s (peak memory %.0f MiB)" %
          (cls, B, fused, tf32, ms, B / ms / 1e3, mem), flush=True)
    cfgmod.set_config(None)
    torch.cuda.reset_peak_memory_stats()
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"config": cls, "batch": B, "fused_cell": bool(fused), "tf32_matmul": bool(tf32), "ms_per_step": ms, "rows_per_s": B / ms * 1e3}


if __name__ == "__main__":
    out = []
    for cls in ("TrainPhase1", "TrainPhase2"):
        for B in (8192, 65536, 262144):
            for fused in (0, 1):
                out.append(run(cls, B, fused))
            out.append(run(cls, B, 1, tf32=True))   # GPU_TRAIN_TF32: matmuls on the tensor cores
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)
