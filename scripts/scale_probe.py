"""GPU probe: step time vs number of worlds (fixed overhead + slope) for the step kernel. Not a bench line."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from rl_collision_avoidance_b200 import _abi
from rl_collision_avoidance_b200.scenarios import random_worlds
from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv

A = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.default_rng(0)
for W in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "8192,16384,32768,65536,131072,262144,524288".split(","))]:
    R = max(2, min(8, int(600e6 // (W * A * 280)) + 1))
    envs = []
    init, nag = random_worlds(W, A, rng)
    for r in range(R):
        e = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1))
        e.set_world_state(init, nag); e.reset(); envs.append(e)
    acts = [torch.randint(0, 11, (W, A), dtype=torch.int32, device="cuda") for _ in range(4)]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for k in range(2 * R): envs[k % R].step(acts[k % 4])
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for k in range(2 * R): envs[k % R].step(acts[k % 4])
        for _ in range(5): g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(20, int(4000 * 65536 / W / (2 * R)))
        e0.record(st)
        for _ in range(n): g.replay()
        e1.record(st); st.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (n * 2 * R)
    print("A=%d W=%7d  %.2f us/step  %.2f G agent-steps/s  alg %.0f GB/s" % (A, W, us, W * A / us / 1e3, (100 + 28 * (A - 1)) * W * A / us / 1e3), flush=True)
    for e in envs: e.close()
    del envs; torch.cuda.empty_cache()
