"""TEST INFRASTRUCTURE — times the UNMODIFIED reference env (gym_collision_avoidance CollisionAvoidanceEnv.step,
GCA/envs/collision_avoidance_env.py:131-194) on the host cores: BASELINE.md §3 / SURVEY §8(d) "CPU baseline".

    python oracle/ref_bench.py --agents 4 --seconds 10 [--procs P]

P processes (default os.cpu_count()), each imports the reference staged under oracle/_ref/ (oracle/stage_ref.py)
through the stub finder of oracle/ref_harness.py, builds worlds with set_testcase-style get_testcase_random
(num_agents = A, all learning_ga3c), seeds NumPy with its rank, and steps with randint(11) actions for `seconds` of
wall time, resetting on game_over.  Prints one JSON line: total agent-steps/s, per-core, P."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "gym-collision-avoidance", "gym_collision_avoidance", "envs"))


def _worker(rank, agents, seconds, out):
    os.environ["CA_REFERENCE_ROOT"] = REF
    os.environ.pop("GYM_CONFIG_CLASS", None)   # the reference's default Config (a caller may have selected a GA3C class)
    os.environ.pop("GYM_CONFIG_PATH", None)
    sys.path.insert(0, os.path.dirname(HERE))
    import numpy as np
    from oracle import ref_harness
    ref_harness.install()
    ns = ref_harness.reference_modules()
    tc = ns.test_cases
    np.random.seed(rank)
    env = ns.CollisionAvoidanceEnv()

    def new_agents():
        return tc.get_testcase_random(num_agents=agents, policies="learning_ga3c", policy_distr=None,
                                      agents_sensors=["other_agents_states"])

    env.set_agents(new_agents())
    env.reset()
    n_steps, n_agent_steps = 0, 0
    t0 = time.perf_counter()
    while True:
        actions = {i: np.random.randint(11) for i in range(agents)}
        obs, rew, over, info = env.step(actions)
        n_steps += 1
        n_agent_steps += agents
        if over:
            env.set_agents(new_agents())
            env.reset()
        if (n_steps & 15) == 0 and time.perf_counter() - t0 >= seconds:
            break
    out.put((n_agent_steps, time.perf_counter() - t0))


def run(agents=4, seconds=10.0, procs=None):
    procs = procs or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, agents, seconds, q)) for r in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = []
    deadline = time.perf_counter() + seconds * 3 + 120
    while len(res) < len(ps):
        try:
            res.append(q.get(timeout=1.0))
        except Exception:
            if any(p.exitcode not in (None, 0) for p in ps):
                for p in ps:
                    p.terminate()
                raise RuntimeError("a reference worker died (exit codes %r)" % [p.exitcode for p in ps])
            if time.perf_counter() > deadline:
                for p in ps:
                    p.terminate()
                raise RuntimeError("reference workers timed out")
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    total = sum(n / el for n, el in res)
    return {"value": total, "unit": "agent-steps/s", "cores": procs, "kind": "reference", "per_core": total / procs,
            "sample": "UNMODIFIED reference CollisionAvoidanceEnv.step (Python/NumPy, staged from /root/reference into "
                      "oracle/_ref/), %d processes x one %d-agent world each (get_testcase_random, all learning_ga3c, "
                      "randint(11) actions, reset on game_over), %.0f s each, %.0f s wall incl. imports"
                      % (procs, agents, seconds, wall)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--procs", type=int, default=None)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "oracle/_ref/ not staged (run oracle/stage_ref.py where /root/reference exists)"}))
        sys.exit(0)
    print(json.dumps(run(a.agents, a.seconds, a.procs)))
