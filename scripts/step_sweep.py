#!/usr/bin/env python
"""Device-resident timing of the step kernel variants (one GPU): the same measurement as bench.py's `value`
(CUDA graph of 12 steps rotating over 6 world sets, CUDA events on the launching stream), repeated for a list of
(workload, environment overrides).  Usage:
    python scripts/step_sweep.py phase1:CA_STEP_KERNEL=stream,CA_PIPE_MINBLOCKS=7 phase1:CA_STEP_KERNEL=oneshot ...
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from rl_collision_avoidance_b200 import _abi  # noqa: E402
from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv  # noqa: E402


def time_steps(workload, overrides, steps=1200, sets_cache={}):
    bench.select_workload(workload)
    W, A = bench.WORLDS_PER_GPU, bench.AGENTS
    R, G = 6, 12
    if workload not in sets_cache:
        sets_cache[workload] = bench.make_inputs(0, R, W)[0]
    sets = sets_cache[workload]
    envvars = {k: v for k, v in overrides.items() if not k.startswith('_')}
    saved = {k: os.environ.get(k) for k in envvars}
    os.environ.update(envvars)
    try:
        envs = []
        for init, nag in sets:
            e = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1, device=0))
            e.set_world_state(init, nag)
            e.reset()
            envs.append(e)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    actions = [torch.randint(0, 11, (W, A), dtype=torch.int32, device="cuda", generator=gen) for _ in range(G)]
    n_streams = int(overrides.get("_STREAMS", "1"))
    if n_streams == 1:
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for k in range(G):
                envs[k % R].step(actions[k])
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for k in range(G):
                    envs[k % R].step(actions[k])
            for _ in range(8):
                graph.replay()
            stream.synchronize()
            best = None
            for rep in range(3):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
                for _ in range(steps // G):
                    graph.replay()
                ev1.record(stream)
                stream.synchronize()
                ms = ev0.elapsed_time(ev1) / (steps // G * G)
                best = ms if best is None else min(best, ms)
    else:
        # the same 12 steps per replay, but the world sets are split over n_streams streams that run concurrently
        # (independent vectorised envs, e.g. a double-buffered rollout): the tail of one launch meets the head of another
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        graphs = []
        for si, st in enumerate(streams):
            mine = [k for k in range(G) if (k % R) % n_streams == si]
            with torch.cuda.stream(st):
                for k in mine:
                    envs[k % R].step(actions[k])
                st.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    for k in mine:
                        envs[k % R].step(actions[k])
                graphs.append(g)
        main = streams[0]
        best = None
        for rep in range(4):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(main):
                ev0.record(main)
            for st in streams[1:]:
                st.wait_event(ev0)
            for _ in range(steps // G):
                for st, g in zip(streams, graphs):
                    with torch.cuda.stream(st):
                        g.replay()
            for st in streams[1:]:
                main.wait_stream(st)
            with torch.cuda.stream(main):
                ev1.record(main)
            main.synchronize()
            ms = ev0.elapsed_time(ev1) / (steps // G * G)
            if rep > 0:
                best = ms if best is None else min(best, ms)
        graph = graphs
    chk = float(envs[0].obs.double().sum().item()) if hasattr(envs[0], "obs") else 0.0
    for e in envs:
        e.close()
    del graph
    torch.cuda.empty_cache()
    live = bench.live_agents(W)
    alg = bench.ALG_BYTES_PER_AGENT_STEP * live
    peak, _ = bench.measured_peak()
    return {"workload": workload, "env": overrides, "us_per_step": 1e3 * best, "G_agent_steps_s": live / best / 1e6,
            "frac": alg / (best * 1e-3) / 1e9 / peak, "checksum": chk}


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        wl, _, rest = spec.partition(":")
        ov = dict(kv.split("=", 1) for kv in rest.split(",") if kv)
        print(json.dumps(time_steps(wl, ov)), flush=True)
