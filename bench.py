#!/usr/bin/env python
"""bench.py — agent-steps/s of the fused env.step() kernel on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload phase1|phase2|ragged]

One "step" = one pass of the hot path (one fused kernel launch) over one batch of W worlds x A agents.
Workload (config.workload): 4 agents x 65 536 worlds per GPU, random-policy discrete actions, auto-reset on
game_over (DummyVecEnv semantics) so worlds keep running.  N>1 (torchrun): worlds shard across ranks with no
data-path collective in the step (weak scaling); NCCL carries the barrier, the max-over-ranks reduction and — in the
`train_loop` extra workload — the gradient all-reduce of the GA3C learner.

Timing: device-side CUDA events on the launching stream, W >= 3 warm-up steps, barrier + synchronize on both sides,
max over ranks.  Every one of the K timed steps is replayed from a CUDA graph (whole graphs of G steps plus one graph
of K mod G steps): no eager launch and no in-process polling thread runs inside the timed region; clocks / throttle
reasons are sampled by a CHILD process through NVML while the timed regions run.  L2: steps rotate over a ring of R
independent world sets whose combined working set exceeds the 126 MB L2, so no step finds its data in L2.
  value      K steps replayed from CUDA graphs, inputs resident in HBM
  e2e        the same metric through ca_step_host_async / ca_step_host_wait: actions from pinned host memory (H2D),
             observations, rewards, done and game_over back to pinned host memory (D2H) every step
  roofline   algorithmic bytes per launch (100 + 28*M bytes per live agent-step) / mean launch duration from
             the same CUDA events; peak = MEASURED_PEAKS.json hbm_gbs; traffic = steady-state DRAM bytes per launch (ncu)
  cpu_baseline            the C oracle port (oracle/ca_oracle.c) on all host cores, same workload (kind "port")
  cpu_baseline_reference  the UNMODIFIED reference Python/NumPy env staged under oracle/_ref/ (kind "reference")
  config.extra_workloads  BASELINE configs[2] (env half and full GA3C rollout), configs[3] (ragged) and the GA3C training
             loop of configs[4] on this rank count, each with its own numbers (sub-records, not the headline)
  --impl reference        the reference arm: the oracle port on all host threads — the only places where bench.py
             executes oracle/; it is the reported baseline, never the thing measured as "ours"
"""
import argparse
import json
import os
import signal
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

# name -> (agent slots per world, worlds per GPU, ragged agent counts, config.workload text)
WORKLOADS = {
    "phase1": (4, 65536, False,
               "4-agent x 65536 worlds vectorised env.step, random-policy actions, auto-reset (BASELINE configs[1])"),
    "phase2": (10, 16384, False,
               "10-agent x 16384 worlds vectorised env.step only (the env half of BASELINE configs[2]), random-policy "
               "actions, auto-reset"),
    "ragged": (10, 32768, True,
               "variable 2-10 agents per world (ragged, n_w = 2 + w mod 9, mean 6 live agents) x 32768 worlds, random-policy "
               "actions, auto-reset (BASELINE configs[3]); agent-steps count live agents only"),
}
AGENTS = 4
WORLDS_PER_GPU = 65536
OTHERS = AGENTS - 1
RAGGED = False
ALG_BYTES_PER_AGENT_STEP = 100 + 28 * OTHERS      # 100 + 28*M bytes; 184 B at M = 3, 352 B at M = 9
WORKLOAD = WORKLOADS["phase1"][3]
WORKLOAD_NAME = "phase1"
STATE_BLOCK_BYTES = 2560                           # csrc/ca_kernels.cuh kBlkBytes
FALLBACK_HBM_GBS = 6650.0                          # /opt/skills/guides/B200_PROFILING.md fallback


def select_workload(name):
    """The default (phase1 = BASELINE configs[1]) is the bench line the driver records; the others are extra records."""
    global AGENTS, WORLDS_PER_GPU, OTHERS, RAGGED, ALG_BYTES_PER_AGENT_STEP, WORKLOAD, WORKLOAD_NAME
    AGENTS, WORLDS_PER_GPU, RAGGED, WORKLOAD = WORKLOADS[name]
    OTHERS = AGENTS - 1
    ALG_BYTES_PER_AGENT_STEP = 100 + 28 * OTHERS
    WORKLOAD_NAME = name


def live_agents(worlds):
    return int(agent_counts(worlds).sum()) if RAGGED else worlds * AGENTS


def agent_counts(worlds):
    """n_w = 2 + (w mod 9) for the ragged workload, else every slot is live."""
    if RAGGED:
        return (2 + (np.arange(worlds) % 9)).astype(np.int32)
    return None


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def step_kernel_name():
    kind = os.environ.get("CA_STEP_KERNEL", "oneshot")
    name = {"stream": "ca_step_stream_kernel", "pipe": "ca_step_stream_kernel", "generic": "ca_world_kernel"}.get(kind, "ca_step_kernel")
    return "ca::%s<%d, ...>" % (name, AGENTS) if name != "ca_world_kernel" else "ca::ca_world_kernel<true>"


def ncu_traffic_per_launch(workload):
    """Steady-state DRAM bytes per launch of the step kernel (ncu --cache-control none over consecutive rotated launches,
    profiles/traffic.json), or None when no capture of this workload is committed."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["ca_step_kernel"][workload]["dram_bytes_per_launch"])
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- clocks (child process)
_SAMPLER_SRC = r'''
import json, signal, sys, time
idx = int(sys.argv[1]); period = float(sys.argv[2])
samples, reasons, stop = [], set(), [False]
def _term(*a): stop[0] = True
signal.signal(signal.SIGTERM, _term); signal.signal(signal.SIGINT, _term)
out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(idx)
    out["sm_max_mhz"] = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
    names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}
    print("ready", flush=True)
    while not stop[0]:
        try:
            mhz = int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            util = int(nv.nvmlDeviceGetUtilizationRates(h).gpu)
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            samples.append((mhz, util))
            for bit, name in names.items():
                if mask & bit: reasons.add(name)
        except Exception:
            pass
        time.sleep(period)
except Exception as e:
    out["error"] = repr(e)
    print("ready", flush=True)
busy = [m for m, u in samples if u > 0] or [m for m, u in samples]
if busy:
    busy.sort(); out["sm_mhz"] = busy[len(busy) // 2]
out["reasons"] = sorted(reasons); out["samples"] = len(samples); out["samples_under_load"] = len([1 for m, u in samples if u > 0])
print(json.dumps(out), flush=True)
'''


class ClockSampler(object):
    """SM clock / throttle reasons polled through NVML by a child process (no thread, no GIL contention in this one)."""

    def __init__(self, index, period=0.01):
        self.proc = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index), str(period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.proc.stdout.readline()   # "ready": NVML is initialised, sampling has started
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        try:
            self.proc.send_signal(signal.SIGTERM)
            out, _ = self.proc.communicate(timeout=10)
            return json.loads(out.strip().split("\n")[-1])
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ---------------------------------------------------------------------------------------------- inputs / CPU arms
def make_inputs(rank, n_sets, worlds):
    from rl_collision_avoidance_b200.scenarios import random_worlds
    rng = np.random.default_rng(20261017 + 1000 * rank)
    return [random_worlds(worlds, AGENTS, rng, num_agents=agent_counts(worlds)) for _ in range(n_sets)], rng


def port_run(worlds, steps=None, seconds=None, warmup=1):
    """The C oracle port on all host cores over `worlds` worlds: `steps` steps, or as many as fit in `seconds`."""
    from oracle.ca_oracle import OracleEnv
    from rl_collision_avoidance_b200 import _abi
    cores = os.cpu_count() or 1
    (sets, rng) = make_inputs(0, 1, worlds)
    init, nag = sets[0]
    env = OracleEnv(_abi.default_config(worlds, AGENTS, auto_reset=1))
    env.set_world_state(init, nag)
    env.reset()
    acts = rng.integers(0, 11, (8, worlds, AGENTS)).astype(np.int32)
    for k in range(max(1, warmup)):
        env.step(acts[k % 8], nthreads=cores)
    n, t0 = 0, time.perf_counter()
    while True:
        env.step(acts[n % 8], nthreads=cores)
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and el >= seconds and n >= 3):
            break
    env.close()
    return n, el, cores


def cpu_baseline_run(seconds):
    """cpu_baseline: the port over the SAME workload as the GPU line and the reference arm (all worlds, every step)."""
    worlds = WORLDS_PER_GPU
    n, el, cores = port_run(worlds, seconds=seconds)
    return {"value": n * live_agents(worlds) / el, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": "oracle/ca_oracle.c (C restatement of the reference env.step), %d pthreads, %d worlds x %d agents, "
                      "%d steps in %.1f s" % (cores, worlds, AGENTS, n, el)}


def cpu_reference_run(seconds):
    """The unmodified reference env on all host cores (oracle/ref_bench.py), or a one-line reason why not."""
    try:
        from oracle import ref_bench
        if not ref_bench.available():
            return {"unavailable": "oracle/_ref/ not staged (oracle/stage_ref.py runs where /root/reference exists)"}
        return ref_bench.run(agents=AGENTS, seconds=seconds)
    except Exception as e:   # a reported baseline must never take the bench line down
        return {"unavailable": "reference env failed to run: %r" % (e,)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    worlds = WORLDS_PER_GPU
    n, el, cores = port_run(worlds, steps=args.steps, warmup=args.warmup)
    value = n * live_agents(worlds) / el
    sample = ("oracle/ca_oracle.c port of the reference env.step on %d host threads; each step = all %d worlds x %d agents"
              % (cores, worlds, AGENTS))
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "worlds": worlds, "agents_per_world": AGENTS},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- device-resident steps
class GraphedSteps(object):
    """R independent world sets (vectorised envs) stepped round-robin over S CUDA streams: step k advances set k mod R on
    stream (k mod R) mod S, so the steps of one set stay ordered on one stream while steps of different sets — which do
    not depend on each other — may overlap (the tail of one launch meets the ramp-up of the next: a double-buffered
    rollout).  S = 1 is the strictly serialised sequence.  All `steps` steps are captured into ONE CUDA graph (fork at
    the start, S branches, join at the end), so the timed region holds exactly one graph launch for ANY step count:
    no eager launch, no per-step host work."""

    def __init__(self, sets, device, seed, streams=1):
        import torch
        from rl_collision_avoidance_b200 import _abi
        from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv
        self.torch = torch
        W, A = WORLDS_PER_GPU, AGENTS
        self.R = len(sets)
        self.S = max(1, min(int(streams), self.R))
        self.G = 2 * self.R
        self.envs = []
        for init, nag in sets:
            e = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1, device=device))
            e.set_world_state(init, nag)
            e.reset()
            self.envs.append(e)
        gen = torch.Generator(device="cuda")
        gen.manual_seed(seed)
        self.actions = [torch.randint(0, 11, (W, A), dtype=torch.int32, device="cuda", generator=gen) for _ in range(self.G)]
        self.streams = [torch.cuda.Stream() for _ in range(self.S)]
        self.stream = self.streams[0]
        self.graphs = {}
        for k in range(self.G):                           # lazy init before capture, on the stream the set will use
            with torch.cuda.stream(self.streams[(k % self.R) % self.S]):
                self._eager(k)
        torch.cuda.synchronize()

    def _eager(self, k):
        self.envs[k % self.R].step(self.actions[k % self.G])

    def _graph(self, n, S):
        torch = self.torch
        key = (n, S)
        if key not in self.graphs:
            main = self.stream
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(main):
                # thread-local capture mode: CUDA calls of OTHER threads (the NCCL watchdog polling its events) must not
                # invalidate a capture of a few thousand launches
                with torch.cuda.graph(g, stream=main, capture_error_mode="thread_local"):
                    fork = torch.cuda.Event()
                    fork.record(main)
                    for st in self.streams[1:S]:
                        st.wait_event(fork)
                    for k in range(n):
                        with torch.cuda.stream(self.streams[(k % self.R) % S]):
                            self._eager(k)
                    for st in self.streams[1:S]:
                        join = torch.cuda.Event()
                        join.record(st)
                        main.wait_event(join)
                g.replay()                                # upload + first run outside any timed region
                main.synchronize()
            self.graphs[key] = g
        return self.graphs[key]

    def run(self, steps, barrier=None, streams=None):
        """Replays exactly `steps` steps (one graph launch); returns the device time in ms (CUDA events around it)."""
        torch = self.torch
        g = self._graph(steps, self.S if streams is None else max(1, min(int(streams), self.S)))
        with torch.cuda.stream(self.stream):
            if barrier:
                barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # device-side spinning first (~0.2 ms + ~1 us per captured launch: submitting a graph of a few thousand nodes
            # takes the host about that long), so that the events and the graph launch are all enqueued before the device
            # reaches them: no host latency lands inside the timed region, however short it is (--steps 20)
            torch.cuda._sleep(400000 + 2000 * int(steps))
            ev0.record(self.stream)
            g.replay()
            ev1.record(self.stream)
            self.stream.synchronize()
            if barrier:
                barrier()
        return ev0.elapsed_time(ev1)

    def close(self):
        self.graphs.clear()
        for e in self.envs:
            e.close()
        self.envs = []
        self.torch.cuda.empty_cache()


def env_workload_record(name, rank, local_rank, steps, streams):
    """One extra env.step workload (device-resident, same method as `value`).  Rank-local: returns the two timings to be
    max-reduced over ranks and a function that builds the sub-record (with its own roofline) from the reduced values."""
    saved = WORKLOAD_NAME
    select_workload(name)
    try:
        sets, _ = make_inputs(rank, 6, WORLDS_PER_GPU)
        gs = GraphedSteps(sets, local_rank, 4321 + rank, streams=streams)
        gs.run(2 * gs.G)
        gs.run(steps)
        ms = gs.run(steps)
        gs.run(steps, streams=1)
        ms1 = gs.run(steps, streams=1)
        S = gs.S
        gs.close()
        live = live_agents(WORLDS_PER_GPU)
        alg, worlds, agents, text = ALG_BYTES_PER_AGENT_STEP, WORLDS_PER_GPU, AGENTS, WORKLOAD
        kernel, traffic = step_kernel_name(), ncu_traffic_per_launch(name)
    finally:
        select_workload(saved)

    def finish(vals, world_size):
        ms, ms1 = vals
        peak, _ = measured_peak()
        launch_ms = ms / steps
        achieved = alg * live / (launch_ms * 1e-3) / 1e9
        achieved1 = alg * live / (ms1 / steps * 1e-3) / 1e9
        return {"workload": text, "value": world_size * live * steps / (ms * 1e-3), "unit": "agent-steps/s",
                "ms_per_step": launch_ms, "steps": steps, "streams": S, "worlds_per_gpu": worlds, "agent_slots": agents,
                "live_agents_per_step": live,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "algorithmic_bytes_per_launch": alg * live, "traffic": traffic, "kernel": kernel},
                "single_stream": {"value": world_size * live * steps / (ms1 * 1e-3), "ms_per_step": ms1 / steps,
                                  "roofline_frac": achieved1 / peak}}
    return [ms, ms1], finish


def rollout_record(cls, worlds, steps, fixed_agents, local_rank):
    """BASELINE configs[2]: the GA3C actor -> predictor loop on the device (row plan, fused NetworkVP forward + action
    sampling, env step, statistics, experience bookkeeping, row gather, scenario refresh), random-init weights."""
    import torch
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    from rl_collision_avoidance_b200.ga3c.rollout import GpuRollout
    from rl_collision_avoidance_b200.scenarios import random_worlds
    cfg = getattr(cfgmod, cls)()
    cfgmod.set_config(cfg)
    try:
        A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
        rng = np.random.default_rng(7 + local_rank)
        init, nag = random_worlds(worlds, A, rng)
        model = NetworkVP_rnn("cuda:%d" % local_rank, "network", 11, seed=0)
        ro = GpuRollout(cfg, model, worlds, init, nag, device=local_rank, seed=1 + local_rank)
        sc = ro.env.scenario_config(cfg.TEST_CASE_ARGS)
        if fixed_agents:
            sc.min_agents = sc.max_agents = A
        ro.env.generate_scenarios(sc, 5, only_consumed=False)
        ro.env.reset(out_obs=ro.rec.obs_slot(ro.t))
        ro.attach_scenario_generator(sc, 5)
        for _ in range(8):
            ro.step()
            ro.rec.discard()
        o = ro.rec.obs_slot(ro.t)
        live = float((o[..., 5] > 0).sum())
        learning = float((o[..., 0] != 0).sum())
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rows = 0
        ev0.record()
        for k in range(steps):
            ro.step()
            if k % 8 == 7:
                rows += int(ro.rec.take()[0].shape[0])
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        o = ro.rec.obs_slot(ro.t)
        live = 0.5 * (live + float((o[..., 5] > 0).sum()))
        learning = 0.5 * (learning + float((o[..., 0] != 0).sum()))
        ro.close()
    finally:
        cfgmod.set_config(None)

    def finish(vals, world_size):
        ms = vals[0]
        return {"workload": "%s GA3C rollout, %d worlds x %d agent slots per GPU, %s, fused tcgen05 predictor "
                            "(BASELINE configs[2])" % (cls, worlds, A, "all agents present" if fixed_agents else
                                                       "training mix (2..A agents, 5/90/5 policies)"),
                "value": world_size * live * steps / (ms * 1e-3), "unit": "live agent-steps/s", "ms_per_step": ms / steps,
                "steps": steps, "live_agents_per_step": live, "learning_agents_per_step": learning,
                "training_rows_per_s": world_size * rows / (ms * 1e-3)}
    return [ms], finish


def reduce_extra(key, n_vals, fn, rank, world_size, device, distributed):
    """One extra workload: a rank-local measurement fn() -> (values, finish) followed by EXACTLY ONE collective whatever
    happens on a rank — [failed, values...] max-reduced over ranks — so an exception on one rank can neither deadlock the
    others nor shift the sequence of collectives (the 8-GPU hang of this round: two ranks skipped the six collectives of
    an extra and paired their later ones with the wrong peers).  Returns finish(reduced values, world_size), or an
    {"error": ...} record on every rank when any rank failed."""
    import torch
    import torch.distributed as dist
    vals, finish, err = [0.0] * n_vals, None, None
    try:
        vals, finish = fn()
    except Exception as e:   # an extra record must never take the headline down
        err = repr(e)
        print("[bench] extra workload %s failed on rank %d: %s" % (key, rank, err), file=sys.stderr, flush=True)
    t = torch.tensor([1.0 if err else 0.0] + [float(v) for v in vals][:n_vals], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.tolist()
    if t[0] != 0.0 or finish is None:
        return {"error": err or "failed on another rank"}
    return finish(t[1:], world_size)


def train_loop_record(seconds, worlds, local_rank, world_size):
    """BASELINE configs[4] on this rank count: Server.main (TrainPhase1) with `worlds` worlds per GPU — rollout, A3C
    updates, and (N > 1) the gradient all-reduce over NCCL."""
    import torch
    os.environ["GYM_CONFIG_CLASS"] = "TrainPhase1"
    os.environ["GA3C_GPU_NUM_WORLDS"] = str(worlds)
    os.environ.setdefault("GA3C_CHECKPOINT_DIR", "/tmp/ga3c_bench_ckpt")
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    from rl_collision_avoidance_b200.ga3c.Server import Server
    cfgmod.set_config(None)
    cfg = cfgmod.get_config()
    cfg.SAVE_MODELS = False
    cfg.EPISODES = 10 ** 12
    try:
        srv = Server(cfg, device=local_rank)
        srv.main(max_steps=8, quiet=True)          # lazy init, allocator warm-up
        f0, s0 = srv.stats.total_frame_count, srv.training_step
        res = srv.main(max_seconds=seconds, quiet=True)
        n_param = sum(p.numel() for p in srv.model.net.parameters())
        frames = srv.stats.total_frame_count - f0
        rec = {"workload": "Server.main TrainPhase1 loop (BASELINE configs[4] at %d GPU%s): %d worlds x 4 agents per GPU, "
                           "fused predictor rollout + A3C updates" % (world_size, "s" if world_size > 1 else "", worlds),
               "frames_per_s": frames / res["seconds"], "env_steps_per_s": res["steps"] / res["seconds"],
               "optimiser_steps_per_s": (srv.training_step - s0) / res["seconds"], "seconds": res["seconds"],
               "rows_per_optimiser_step": srv.train_batch_rows() * world_size,
               "learning_rate": srv.model.learning_rate, "lr_scale": str(cfg.GPU_LR_SCALE),
               "allreduce_bytes_per_optimiser_step": 4 * n_param if world_size > 1 else 0,
               "collective": "NCCL all_reduce of the flat fp32 gradient" if world_size > 1 else "none (1 GPU)",
               "episodes": res["episodes"], "rolling_score": srv.stats.roll_reward_log}
        srv.rollout.close()
        return rec
    finally:
        cfgmod.set_config(None)
        torch.cuda.empty_cache()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rl_collision_avoidance_b200 import _abi, _lib
    from rl_collision_avoidance_b200.vec_env import HostVecEnv

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
        import datetime
        # a rank that dies must not hold the others (and the box) for the default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))

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
    R = 6          # ring of world sets: 6 x (19 MB state + 28 MB obs + ...) ~ 300 MB >> 126 MB L2
    sets, rng = make_inputs(rank, R, W)
    S = max(1, min(args.streams, R))
    gs = GraphedSteps(sets, local_rank, 1234 + rank, streams=S)
    G = gs.G
    bytes_per_set = 2 * (W // max(1, min(32 // A, 16))) * STATE_BLOCK_BYTES + W * A * 4 + W * A * _abi.obs_len(OTHERS) * 4 + W * A * 9
    sampler = ClockSampler(local_rank)
    launches0 = sum(e.handle.launch_count for e in gs.envs)
    gs.run(max(WU, G))                           # warm-up (>= W steps), same replay path as the timed region
    gs.run(K)                                    # one untimed pass of exactly the timed graph (uploaded, clocks up)
    dev_ms = max_over_ranks(gs.run(K, barrier))  # THE timed region: K steps = one graph launch
    # the same K steps strictly serialised on one stream (every launch waits for the previous one): reported beside it
    gs.run(K, streams=1)
    dev_ms_1 = max_over_ranks(gs.run(K, barrier, streams=1)) if S > 1 else dev_ms
    eager_launches = sum(e.handle.launch_count for e in gs.envs) - launches0
    gpu_launches = K                             # kernels launched inside the timed region (all from graph replays)
    agent_steps = K * live_agents(W)
    value = world_size * agent_steps / (dev_ms * 1e-3)
    gs.close()

    # ---- e2e: host buffers through ca_step_host (H2D actions, D2H obs/reward/done/game_over every step)
    Rh = 3
    T = 12
    henvs = []
    for init, nag in sets[:Rh]:
        h = HostVecEnv(_abi.default_config(W, A, auto_reset=1, device=local_rank))
        h.set_world_state(init, nag)
        h.reset()
        henvs.append(h)
    host_actions = rng.integers(0, 11, (T, W, A)).astype(np.int32)
    Ke = max(min(K, 400), 120)                   # e2e is timed over at least 120 steps whatever --steps is
    for k in range(6):
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
    e2e_value = world_size * Ke * live_agents(W) / e2e_s
    h2d, d2h = henvs[0].h2d_bytes_per_step, henvs[0].d2h_bytes_per_step
    for h in henvs:
        h.close()
    torch.cuda.empty_cache()

    # ---- extra workloads (sub-records; every rank takes part so that N > 1 numbers are whole-job numbers)
    extras = {}
    if not args.no_extras and WORKLOAD_NAME == "phase1":
        def reduced(key, n_vals, fn):
            extras[key] = reduce_extra(key, n_vals, fn, rank, world_size, "cuda", distributed)
        reduced("phase2_env_step", 2, lambda: env_workload_record("phase2", rank, local_rank, 240, S))
        reduced("ragged_env_step", 2, lambda: env_workload_record("ragged", rank, local_rank, 240, S))
        reduced("rollout_phase2_all_present", 1, lambda: rollout_record("TrainPhase2", 16384, 96, True, local_rank))
        reduced("rollout_phase2_training_mix", 1, lambda: rollout_record("TrainPhase2", 16384, 96, False, local_rank))
        # the training loop has collectives of its own (gradient all-reduce): every rank must enter it, and a rank-local
        # failure inside it cannot be isolated — the process group's timeout bounds the damage
        barrier()
        try:
            extras["train_loop"] = train_loop_record(args.train_seconds, 65536, local_rank, world_size)
        except Exception as e:
            extras["train_loop"] = {"error": repr(e)}
            print("[bench] train_loop failed on rank %d: %r" % (rank, e), file=sys.stderr, flush=True)
        barrier()
    clocks = sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peak()
        launch_ms = dev_ms / K
        achieved = ALG_BYTES_PER_AGENT_STEP * live_agents(W) / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world_size, "steps": K,
            "warmup": WU, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "worlds_per_gpu": W, "agents_per_world": A, "others_observed": OTHERS,
                       "obs_len": _abi.obs_len(OTHERS),
                       "launch": "all %d timed steps replayed as ONE CUDA graph launch: %d independent world sets (vectorised "
                                 "envs) stepped round-robin, set r on stream r mod %d (steps of one set stay ordered, "
                                 "launches of different sets may overlap: the tail of one meets the ramp-up of the next); "
                                 "single_stream = the same steps strictly serialised on one stream; no eager launch in "
                                 "the timed region (%d eager + capture launches before it)" % (K, R, S, eager_launches),
                       "streams": S,
                       "l2": "inputs larger than L2: steps rotate over %d independent world sets (%.0f MB total per GPU)"
                             % (R, R * bytes_per_set / 1e6),
                       "extra_workloads": extras},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke,
                    "sync_value": world_size * Ks * live_agents(W) / sync_s,
                    "api": "ca_step_host_async/_wait = VecEnv.step_async/step_wait over %d independent host envs, pinned host buffers" % Rh},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(WORKLOAD_NAME), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_AGENT_STEP * live_agents(W),
                         "kernel": step_kernel_name(), "launch_ms": launch_ms,
                         "frac_single_stream": ALG_BYTES_PER_AGENT_STEP * live_agents(W) / (dev_ms_1 / K * 1e-3) / 1e9 / peak},
            "single_stream": {"value": world_size * agent_steps / (dev_ms_1 * 1e-3), "unit": "agent-steps/s",
                              "ms_per_step": dev_ms_1 / K},
        }
        if world_size == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_run(args.cpu_seconds)
            line["cpu_baseline_reference"] = cpu_reference_run(args.cpu_seconds)
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
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall time of each cpu_baseline sample")
    ap.add_argument("--train-seconds", type=float, default=6.0, help="wall time of the train_loop extra workload")
    ap.add_argument("--streams", type=int, default=3,
                    help="CUDA streams the 6 independent world sets are stepped on (1 = strictly serialised launches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip config.extra_workloads")
    ap.add_argument("--workload", default="phase1", choices=sorted(WORKLOADS),
                    help="phase1 = BASELINE configs[1] (the recorded bench line); phase2 / ragged = standalone extra lines")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
