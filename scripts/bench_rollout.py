"""GPU: throughput of the full GA3C actor->predictor loop (BASELINE configs[2]: 10 agents x 16384 worlds with the
NetworkVP LSTM forward fused; configs[1] size as a second point).  Reports the full rollout rate (predictor + action
selection + env step + experience bookkeeping + scenario refresh) with the fused tcgen05 predictor and with the
composed fp32 one (ca_lstm_step + cuBLAS + torch.multinomial), and each predictor alone (CUDA events).
Not the bench.py line (that is the env.step hot path); numbers are quoted in DESIGN.md §6.
    python scripts/bench_rollout.py [--json gpurun_out/rollout.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from rl_collision_avoidance_b200.ga3c import Config as cfgmod
from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
from rl_collision_avoidance_b200.ga3c.rollout import GpuRollout
from rl_collision_avoidance_b200.scenarios import random_worlds


def _events(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(cls, W, steps, predictor, tf32=False, fixed_agents=False):
    os.environ["GA3C_PREDICTOR"] = predictor
    cfg = getattr(cfgmod, cls)()
    cfgmod.set_config(cfg)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
    rng = np.random.default_rng(0)
    init, nag = random_worlds(W, A, rng)
    model = NetworkVP_rnn("cuda:0", "network", 11, seed=0)
    ro = GpuRollout(cfg, model, W, init, nag, device=0, seed=1)
    sc = ro.env.scenario_config(cfg.TEST_CASE_ARGS)
    if fixed_agents:   # every world has all A agents (BASELINE configs[2] "10-agent worlds"); policy mix as in training
        sc.min_agents = sc.max_agents = A
    # every world starts from a scenario of the on-device generator (the training distribution, or all A agents present)
    ro.env.generate_scenarios(sc, 5, only_consumed=False)
    ro.env.reset(out_obs=ro.rec.obs_slot(ro.t))
    ro.attach_scenario_generator(sc, 5)    # consumed worlds get a new scenario after every step (side stream)
    for _ in range(8):
        ro.step(); ro.rec.discard()
    o = ro.rec.obs_slot(ro.t)   # first use of these torch reductions loads their kernels: keep that out of the timing
    float((o[..., 5] > 0).sum()); float((o[..., 0] != 0).sum())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = 0
    live = learning = 0.0
    samples = 0
    for k in range(steps):
        ro.step()
        if k % 8 == 7:
            rows += ro.rec.take()[0].shape[0]
            o = ro.rec.obs_slot(ro.t)
            live += float((o[..., 5] > 0).sum()); learning += float((o[..., 0] != 0).sum()); samples += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    live_per_step, learning_per_step = live / max(samples, 1), learning / max(samples, 1)
    obs = ro.rec.obs_slot(0).reshape(W * A, -1)
    if predictor == "fused":
        ms_pred = _events(lambda: model.predict_fused(obs, want_p=False, want_actions=True))
    else:
        ms_pred = _events(lambda: torch.multinomial(model.predict_from_obs(obs)[0], 1))
    # per-phase wall time with a device synchronise after every phase (host launch overhead + kernel time of each)
    import ctypes as C
    phases = {"predict": 0.0, "env_step": 0.0, "record": 0.0, "generate": 0.0, "take": 0.0}
    def tick(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        phases[name] += t1 - t0
        return t1
    torch.cuda.synchronize()
    ro._gen_cfg = None          # the breakdown launches the generator itself, on the main stream
    ro._gen_done = None
    nb = 24
    for k in range(nb):
        t0 = time.perf_counter()
        o = ro.rec.obs_slot(ro.t)
        if ro.fused:
            _, v, actions = model.predict_fused(o.reshape(ro.N, ro.L), want_p=False, want_actions=True, greedy=ro.greedy, seed=ro.seed)
        else:
            p_, v = model.predict_from_obs(o.reshape(ro.N, ro.L))
            actions = torch.multinomial(p_, 1, generator=ro.gen).squeeze(1).to(torch.int32)
        t0 = tick("predict", t0)
        _, reward, done, over = ro.env.step(actions.view(ro.W, ro.A), out_obs=ro.rec.obs_slot(ro.t + 1))
        t0 = tick("env_step", t0)
        ro.rec.record(ro.t, actions, v.contiguous(), reward.view(-1), done.view(-1), over)
        ro.t += 1
        t0 = tick("record", t0)
        ro.env.generate_scenarios(sc, 5, only_consumed=True)
        t0 = tick("generate", t0)
        if k % 8 == 7:
            ro.rec.take()
            t0 = tick("take", t0)
    phases = {k_: 1e3 * v_ / nb for k_, v_ in phases.items()}
    print("   per-phase ms/step (synchronised after each phase): " + ", ".join("%s %.3f" % kv for kv in phases.items()), flush=True)
    res = {"phases_ms": phases, "config": cls, "worlds": W, "agents": A, "predictor": predictor, "tf32_matmul": tf32,
           "fixed_agents": fixed_agents,
           "rollout_ms_per_step": 1e3 * dt / steps, "agent_steps_per_s": live_per_step * steps / dt,
           "slot_steps_per_s": W * A * steps / dt, "live_agents_per_step": live_per_step,
           "learning_agents_per_step": learning_per_step,
           "learner_rows_per_s": rows / dt, "predictor_ms": ms_pred, "predictor_rows": W * A}
    print("%s W=%d A=%d %s predictor=%s%s: rollout %.3f ms/step = %.1f M live agent-steps/s (%.0f live / %.0f learning of %d slots "
          "per step; %.1f M learner rows/s emitted); predictor + action selection alone over all slots %.3f ms" %
          (cls, W, A, "all-agents-present" if fixed_agents else "training mix (2..A agents)", predictor,
           " (tf32 matmul)" if tf32 else "", res["rollout_ms_per_step"], res["agent_steps_per_s"] / 1e6, live_per_step,
           learning_per_step, W * A, res["learner_rows_per_s"] / 1e6, ms_pred), flush=True)
    ro.close()
    cfgmod.set_config(None)
    return res


if __name__ == "__main__":
    out = []
    if "--only" in sys.argv:   # e.g. --only TrainPhase2:16384:fused:40  (one configuration, for profiling)
        cls, W, pred, steps = sys.argv[sys.argv.index("--only") + 1].split(":")[:4]
        out.append(run(cls, int(W), int(steps), pred, fixed_agents="--fixed" in sys.argv))
    else:
        for cls, W in (("TrainPhase2", 16384), ("TrainPhase1", 65536)):
            out.append(run(cls, W, 60, "fused", fixed_agents=True))
            out.append(run(cls, W, 60, "fused"))
            out.append(run(cls, W, 60, "composed"))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)
