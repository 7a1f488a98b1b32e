#!/usr/bin/env python
"""Timeline of one step launch, per warp-chunk (experiment tool, one GPU).

Builds build/libcastep_trace.so = the library with -DCA_TRACE (lane 0 of every warp writes %globaltimer stamps: kernel
entry / after griddepcontrol.wait / state consumed / reward done / rows assembled / tile stored, plus its SM id), runs the
same graph of 12 steps over 6 world sets as bench.py, and reads the stamps of the LAST launch of each set.  Prints, per
variant, when the phases of the launch happen relative to its first stamp (percentiles over the chunks) and how long the
phases of a chunk take.  Usage:
    python scripts/step_timeline.py build                      (here, needs nvcc only)
    python scripts/step_timeline.py phase1:CA_STEP_KERNEL=oneshot phase1:CA_STEP_KERNEL=stream,CA_STREAM_STATIC=1 ...
"""
import ctypes as C
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
TRACE_LIB = os.path.join(REPO, "build", "libcastep_trace.so")


def build():
    from rl_collision_avoidance_b200 import _lib
    os.makedirs(os.path.join(REPO, "build"), exist_ok=True)
    objs = []
    procs = []
    for src in _lib.SOURCES:
        obj = os.path.join(REPO, "build", "trace_" + src.replace(".cu", ".o"))
        cmd = [_lib.nvcc_path()] + _lib.NVCC_FLAGS + ["-DCA_TRACE=1", "-c", "-o", obj, os.path.join(_lib.CSRC_DIR, src)]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([_lib.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", TRACE_LIB] + objs)
    print("built", TRACE_LIB)


def run(specs):
    import numpy as np
    import torch
    from rl_collision_avoidance_b200 import _lib
    _lib.LIB_PATH = TRACE_LIB          # the experiment build, never the shipped one
    import bench
    from rl_collision_avoidance_b200 import _abi
    from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv
    L = _lib.lib()
    L.ca_trace_read.argtypes = [C.c_void_p, C.c_void_p]
    L.ca_trace_read.restype = C.c_int
    out = []
    for spec in specs:
        wl, _, rest = spec.partition(":")
        ov = dict(kv.split("=", 1) for kv in rest.split(",") if kv)
        bench.select_workload(wl)
        W, A = bench.WORLDS_PER_GPU, bench.AGENTS
        R, G = 6, 12
        sets = bench.make_inputs(0, R, W)[0]
        saved = {k: os.environ.get(k) for k in ov}
        os.environ.update(ov)
        envs = []
        for init, nag in sets:
            e = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1, device=0))
            e.set_world_state(init, nag)
            e.reset()
            envs.append(e)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1234)
        actions = [torch.randint(0, 11, (W, A), dtype=torch.int32, device="cuda", generator=gen) for _ in range(G)]
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for k in range(G):
                envs[k % R].step(actions[k])
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for k in range(G):
                    envs[k % R].step(actions[k])
            for _ in range(20):
                graph.replay()
            stream.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(50):
                graph.replay()
            ev1.record(stream)
            stream.synchronize()
            us = 1e3 * ev0.elapsed_time(ev1) / (50 * G)
        wpw = min(32 // A, 16)
        n_chunks = (W + wpw - 1) // wpw
        tr = []
        for e in envs:
            buf = np.zeros((n_chunks, 8), dtype=np.uint64)
            rc = L.ca_trace_read(e.handle._h, buf.ctypes.data_as(C.c_void_p))
            assert rc == 0
            tr.append(buf.astype(np.int64))
        # launches of the last graph replay: set k ran at positions k and k + 6; the stamps are of position k + 6
        starts = [int(t[:, 0].min()) for t in tr]
        order = np.argsort(starts)
        rep = {"spec": spec, "us_per_step_traced": us, "launches": []}
        pct = [0, 10, 50, 90, 99, 100]
        for pos, k in enumerate(order):
            t = tr[k]
            t0 = starts[k]
            rel = (t[:, :6] - t0) / 1e3
            d = {"first_stamp_after_previous_launch_first_us": None if pos == 0 else (t0 - starts[order[pos - 1]]) / 1e3,
                 "span_us": float(rel[:, 5].max())}
            names = ["entry", "after_pdl_wait", "state_consumed", "reward_done", "rows_done", "stored"]
            for j, nm in enumerate(names):
                d[nm] = [round(float(np.percentile(rel[:, j], q)), 2) for q in pct]
            dur = np.diff(t[:, :6], axis=1) / 1e3
            for j, nm in enumerate(["pdl_wait", "load+action", "pairs+reward", "rows", "store"]):
                d["dur_" + nm] = [round(float(np.percentile(dur[:, j], q)), 2) for q in pct]
            d["frac_rows_over_1p5us"] = float((dur[:, 3] > 1.5).mean())
            d["frac_load_over_2us"] = float((dur[:, 1] > 2.0).mean())
            # chunks per SM and busy span per SM
            sm = t[:, 6]
            d["chunks_per_sm_minmax"] = [int(np.bincount(sm).min()), int(np.bincount(sm).max())]
            # how many chunks are between state_consumed and stored at each 0.5 us tick (compute concurrency)
            ticks = np.arange(0, rel[:, 5].max(), 0.5)
            d["in_load_wait"] = [int(((rel[:, 1] <= x) & (rel[:, 2] > x)).sum()) for x in ticks]
            d["in_compute"] = [int(((rel[:, 2] <= x) & (rel[:, 4] > x)).sum()) for x in ticks]
            d["in_store"] = [int(((rel[:, 4] <= x) & (rel[:, 5] > x)).sum()) for x in ticks]
            rep["launches"].append(d)
        out.append(rep)
        print(json.dumps(rep), flush=True)
        if os.environ.get("CA_TIMELINE_RAW"):
            np.savez_compressed(os.path.join(REPO, "gpurun_out", "timeline_raw_%d.npz" % len(out)), t=tr[order[3]], spec=spec)
        for e in envs:
            e.close()
        del graph
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    if sys.argv[1:] == ["build"]:
        build()
    else:
        run(sys.argv[1:])
