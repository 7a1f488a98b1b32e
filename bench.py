#!/usr/bin/env python
"""bench.py — agent-steps/s of the fused env.step() kernel on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (one fused kernel launch) over one batch of W worlds x A agents.
Workload (config.workload): 4 agents x 65 536 worlds per GPU, random-policy discrete actions, auto-reset on
game_over (DummyVecEnv semantics) so worlds keep running.  N>1 (torchrun): worlds shard across ranks with no
data-path collective (weak scaling); NCCL is used for the barrier and the max-over-ranks reduction only.

Timing: device-side CUDA events on the launching stream, W >= 3 warm-up steps, barrier + synchronize on
both sides, max over ranks.  L2: steps rotate over a ring of R independent world sets whose combined
working set (state + observations) exceeds the 126 MB L2, so no step finds its data in L2.
  value      K steps replayed from CUDA graphs, inputs resident in HBM
  e2e        the same K steps through ca_step_host: actions from pinned host memory (H2D), observations,
             rewards, done and game_over back to pinned host memory (D2H), synchronous per step
  roofline   algorithmic bytes per launch (100 + 28*M bytes per live agent-step) / mean launch
             duration from the same CUDA events; peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference   the C oracle port (oracle/ca_oracle.c) on all host cores — the only places
             where bench.py executes oracle/; it is the reported baseline, never the thing measured as "ours"
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

# name -> (agent slots per world, worlds per GPU, ragged agent counts, config.workload text, one-shot kernel occupancy)
WORKLOADS = {
    "phase1": (4, 65536, False,
               "4-agent x 65536 worlds vectorised env.step, random-policy actions, auto-reset (BASELINE configs[1])", 7),
    "phase2": (10, 16384, False,
               "10-agent x 16384 worlds vectorised env.step only (the env half of BASELINE configs[2]), random-policy "
               "actions, auto-reset", 5),
    "ragged": (10, 32768, True,
               "variable 2-10 agents per world (ragged, n_w = 2 + w mod 9, mean 6 live agents) x 32768 worlds, random-policy "
               "actions, auto-reset (BASELINE configs[3]); agent-steps count live agents only", 5),
}
AGENTS = 4
WORLDS_PER_GPU = 65536
OTHERS = AGENTS - 1
RAGGED = False
MINBLOCKS = 7
ALG_BYTES_PER_AGENT_STEP = 100 + 28 * OTHERS      # 100 + 28*M bytes; 184 B at M = 3, 352 B at M = 9
WORKLOAD = WORKLOADS["phase1"][3]


def select_workload(name):
    """The default (phase1 = BASELINE configs[1]) is the bench line the driver records; the others are extra lines."""
    global AGENTS, WORLDS_PER_GPU, OTHERS, RAGGED, MINBLOCKS, ALG_BYTES_PER_AGENT_STEP, WORKLOAD
    AGENTS, WORLDS_PER_GPU, RAGGED, WORKLOAD, MINBLOCKS = WORKLOADS[name]
    OTHERS = AGENTS - 1
    ALG_BYTES_PER_AGENT_STEP = 100 + 28 * OTHERS


def live_agents(worlds):
    return int(agent_counts(worlds).sum()) if RAGGED else worlds * AGENTS


def agent_counts(worlds):
    """n_w = 2 + (w mod 9) for the ragged workload, else every slot is live."""
    if RAGGED:
        return (2 + (np.arange(worlds) % 9)).astype(np.int32)
    return None
FALLBACK_HBM_GBS = 6650.0                          # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture, or None."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["ca_world_kernel_step"]["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Polls SM clock / throttle reasons through NVML every ~2 ms while the timed regions run."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, util))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        mhz = [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(mhz), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(mhz)}


def make_inputs(rank, n_sets, worlds):
    from rl_collision_avoidance_b200.scenarios import random_worlds
    rng = np.random.default_rng(20261017 + 1000 * rank)
    return [random_worlds(worlds, AGENTS, rng, num_agents=agent_counts(worlds)) for _ in range(n_sets)], rng


def cpu_baseline_run(seconds, worlds):
    """C oracle port on all host cores, bounded sample of the same workload."""
    from oracle.ca_oracle import OracleEnv
    from rl_collision_avoidance_b200 import _abi
    cores = os.cpu_count() or 1
    (sets, rng) = make_inputs(0, 1, worlds)
    init, nag = sets[0]
    env = OracleEnv(_abi.default_config(worlds, AGENTS, auto_reset=1))
    env.set_world_state(init, nag)
    env.reset()
    acts = rng.integers(0, 11, (8, worlds, AGENTS)).astype(np.int32)
    env.step(acts[0], nthreads=cores)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        env.step(acts[n % 8], nthreads=cores)
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds and n >= 3:
            break
    env.close()
    return {"value": n * live_agents(worlds) / el, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": "oracle/ca_oracle.c (C restatement of the reference env.step), %d pthreads, %d worlds x %d agents, "
                      "%d steps in %.1f s; the reference's own Python/NumPy env runs ~3.5k agent-steps/s per core "
                      "(BASELINE.md §2) and cannot travel to the GPU box" % (cores, worlds, AGENTS, n, el)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.ca_oracle import OracleEnv
    from rl_collision_avoidance_b200 import _abi
    cores = os.cpu_count() or 1
    worlds = WORLDS_PER_GPU
    (sets, rng) = make_inputs(0, 1, worlds)
    init, nag = sets[0]
    env = OracleEnv(_abi.default_config(worlds, AGENTS, auto_reset=1))
    env.set_world_state(init, nag)
    env.reset()
    acts = rng.integers(0, 11, (8, worlds, AGENTS)).astype(np.int32)
    for k in range(args.warmup):
        env.step(acts[k % 8], nthreads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        env.step(acts[k % 8], nthreads=cores)
    el = time.perf_counter() - t0
    env.close()
    value = args.steps * live_agents(worlds) / el
    sample = ("oracle/ca_oracle.c port of the reference env.step on %d host threads; each step = all %d worlds x %d agents"
              % (cores, worlds, AGENTS))
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "worlds": worlds, "agents_per_world": AGENTS},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rl_collision_avoidance_b200 import _abi, _lib
    from rl_collision_avoidance_b200.vec_env import HostVecEnv, VecCollisionAvoidanceEnv

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    torch.cuda.set_device(local_rank)
    distributed = world_size > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, A, K, WU = WORLDS_PER_GPU, AGENTS, args.steps, max(args.warmup, 3)
    R = 6          # ring of world sets: 6 x (21 MB state + 28 MB obs + ...) ~ 300 MB >> 126 MB L2
    G = 2 * R      # steps per captured CUDA graph
    T = G          # ring of action tensors
    sets, rng = make_inputs(rank, R, W)
    envs = []
    for init, nag in sets:
        e = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1, device=local_rank))
        e.set_world_state(init, nag)
        e.reset()
        envs.append(e)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    actions = [torch.randint(0, 11, (W, A), dtype=torch.int32, device="cuda", generator=gen) for _ in range(T)]
    bytes_per_set = 2 * (W // max(1, min(32 // A, 16))) * 2304 + W * A * 4 + W * A * _abi.obs_len(OTHERS) * 4 + W * A * 9   # state blocks x2, actions, obs, outputs

    def eager_step(k):
        envs[k % R].step(actions[k % T])

    stream = torch.cuda.Stream()
    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for k in range(G):                       # lazy init before capture
            eager_step(k)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(G):
                eager_step(k)
        n_graph, n_tail = K // G, K % G
        for _ in range(max(1, (WU + G - 1) // G)):   # warm-up (>= W steps)
            graph.replay()
        barrier()
        sampler.start()
        launches0 = sum(e.handle.launch_count for e in envs)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(n_graph):
            graph.replay()
        for k in range(n_tail):
            eager_step(k)
        ev1.record(stream)
        stream.synchronize()
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
    # graph replays launch the captured kernels without going through ca_step: count them explicitly
    gpu_launches = n_graph * G + (sum(e.handle.launch_count for e in envs) - launches0)
    dev_ms = max_over_ranks(dev_ms)
    agent_steps = K * live_agents(W)
    value = world_size * agent_steps / (dev_ms * 1e-3)

    # ---- e2e: host buffers through ca_step_host (H2D actions, D2H obs/reward/done/game_over every step)
    for e in envs:
        e.close()
    del envs
    torch.cuda.empty_cache()
    Rh = 3
    henvs = []
    for init, nag in sets[:Rh]:
        h = HostVecEnv(_abi.default_config(W, A, auto_reset=1, device=local_rank))
        h.set_world_state(init, nag)
        h.reset()
        henvs.append(h)
    host_actions = rng.integers(0, 11, (T, W, A)).astype(np.int32)
    Ke = min(K, 400)
    for k in range(3):
        henvs[k % Rh].step(host_actions[k % T])
    barrier()
    # (a) strictly synchronous: one ca_step_host (= VecEnv.step) at a time, reported as e2e.sync_value
    Ks = min(Ke, 100)
    t0 = time.perf_counter()
    for k in range(Ks):
        h = henvs[k % Rh]
        np.copyto(h.actions_buf, host_actions[k % T])
        obs, rew, done, over = h.step(h.actions_buf)
    sync_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # (b) The caller drives Rh independent vectorised envs round-robin through VecEnv.step_async / step_wait: env k's results
    # are awaited (and read) only after env k+1's step has been enqueued, so the D2H of one env overlaps the H2D +
    # kernel of the next.  Every step still moves its actions host->device and its results device->host.
    t0 = time.perf_counter()
    checksum = 0.0
    np.copyto(henvs[0].actions_buf, host_actions[0])
    henvs[0].step_async(henvs[0].actions_buf)
    for k in range(Ke):
        if k + 1 < Ke:
            hn = henvs[(k + 1) % Rh]
            np.copyto(hn.actions_buf, host_actions[(k + 1) % T])   # the step's inputs land in pinned memory ...
            hn.step_async(hn.actions_buf)                          # ... H2D, kernel, D2H enqueued
        obs, rew, done, over = henvs[k % Rh].step_wait()           # results of step k are in host memory
        checksum += float(rew[0, 0]) + float(over[0]) + float(obs[-1, -1, 2])   # the host reads the result
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    clocks = sampler.stop()
    e2e_value = world_size * Ke * live_agents(W) / e2e_s
    h2d, d2h = henvs[0].h2d_bytes_per_step, henvs[0].d2h_bytes_per_step
    gpu_launches += Ke + Ks + 3
    for h in henvs:
        h.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        launch_ms = dev_ms / K
        achieved = ALG_BYTES_PER_AGENT_STEP * live_agents(W) / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world_size, "steps": K,
            "warmup": WU, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "worlds_per_gpu": W, "agents_per_world": A, "others_observed": OTHERS,
                       "obs_len": _abi.obs_len(OTHERS), "launch": "CUDA graph of %d steps replayed" % G,
                       "l2": "inputs larger than L2: steps rotate over %d independent world sets (%.0f MB total per GPU)"
                             % (R, R * bytes_per_set / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke,
                    "sync_value": world_size * Ks * live_agents(W) / sync_s,
                    "api": "ca_step_host_async/_wait = VecEnv.step_async/step_wait over %d independent host envs, pinned host buffers" % Rh},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch() if WORKLOAD == WORKLOADS["phase1"][3] else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_AGENT_STEP * live_agents(W),
                         "kernel": "ca::ca_step_kernel<%d, %d, false>" % (A, MINBLOCKS), "launch_ms": launch_ms},
        }
        if world_size == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_run(args.cpu_seconds, 16384)
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2400)
    ap.add_argument("--warmup", type=int, default=48)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="phase1", choices=sorted(WORKLOADS),
                    help="phase1 = BASELINE configs[1] (the recorded bench line); phase2 / ragged = extra lines")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
