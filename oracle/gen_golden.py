"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # regenerates every golden file

Each golden file holds several episodes ("cases").  Per case we record the injected
initial state, the action sequence, and after every reference `env.step()`
(GCA/envs/collision_avoidance_env.py:131-194) the full agent state, flags, rewards,
dense observation (through the reference's own MultiagentDictToMultiagentArrayWrapper,
GCA/envs/wrappers.py:111-139), the neighbour order chosen by
OtherAgentsStatesSensor.get_clipped_sorted_inds (sensors/OtherAgentsStatesSensor.py:20-55)
and game_over.

Oracle modes: mode "B" injects the heading as np.float64 (pure
float64 dynamics = the authors' NumPy-1.x behaviour); mode "A" injects a Python float
(float32-contaminated under NumPy>=2).
"""
import argparse
import os
import pickle
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")

POL_LEARNING_GA3C, POL_NONCOOP, POL_STATIC, POL_LEARNING = 0, 1, 2, 3

FLAG_NAMES = ("is_at_goal", "was_at_goal_already", "in_collision", "was_in_collision_already",
              "ran_out_of_time", "is_done")


def _load_cases(ref_root, n_agents):
    fn = os.path.join(ref_root, "gym-collision-avoidance", "gym_collision_avoidance", "envs",
                      "test_cases", "%d_agents_500_cases.p" % n_agents)
    with open(fn, "rb") as f:
        return pickle.load(f, encoding="latin1")


def run_case(ns, init, actions, mode="B", cont_actions=None):
    """init: (n, 8) [px,py,gx,gy,pref_speed,radius,heading,policy]; actions (T, n) int.
    Returns dict of recorded arrays; stops at game_over."""
    Config = ns.Config
    A = Config.MAX_NUM_AGENTS_IN_ENVIRONMENT
    M = Config.MAX_NUM_OTHER_AGENTS_OBSERVED
    n = init.shape[0]
    T = actions.shape[0]
    agents = []
    for i in range(n):
        px, py, gx, gy, ps, rad, h, pol = init[i]
        heading = np.float64(h) if mode == "B" else float(h)
        agents.append(ns.Agent(px, py, gx, gy, rad, ps, heading, ns.policies[int(pol)],
                               ns.UnicycleDynamics, [ns.OtherAgentsStatesSensor], i))
    sort_log = {}
    for i, ag in enumerate(agents):
        sensor = ag.sensors[0]

        def wrapped(crit, _orig=sensor.get_clipped_sorted_inds, _i=i):
            out = _orig(crit)
            sort_log[_i] = list(out)
            return out
        sensor.get_clipped_sorted_inds = wrapped

    env = ns.CollisionAvoidanceEnv()
    from gym_collision_avoidance.envs.wrappers import MultiagentDictToMultiagentArrayWrapper
    wrap = MultiagentDictToMultiagentArrayWrapper(env, dict_keys=Config.STATES_IN_OBS, max_num_agents=A)
    env.set_agents(agents)
    obs0 = wrap.observation(env.reset())
    L = obs0.shape[1]

    def sorted_now():
        s = -np.ones((n, M), dtype=np.int32)
        for i in range(n):
            idx = sort_log.get(i, [])
            s[i, :len(idx)] = idx
        return s

    rec = dict(
        init=np.asarray(init, dtype=np.float64),
        t_rem0=np.array([a.time_remaining_to_reach_goal for a in agents], dtype=np.float64),
        obs0=obs0.astype(np.float64), sorted0=sorted_now(),
        pos=np.zeros((T, n, 2)), heading=np.zeros((T, n)), vel=np.zeros((T, n, 2)),
        goal=np.zeros((T, n, 2)), t_rem=np.zeros((T, n)),
        flags=np.zeros((T, n, len(FLAG_NAMES)), dtype=np.uint8),
        reward=np.zeros((T, n)), obs=np.zeros((T, A, L)),
        sorted=-np.ones((T, n, M), dtype=np.int32), game_over=np.zeros((T,), dtype=np.uint8),
        dist_to_goal=np.zeros((T, n)), heading_ego=np.zeros((T, n)),
    )
    steps = 0
    for t in range(T):
        act = {}
        for i in range(n):
            pol = int(init[i, 7])
            if pol == POL_LEARNING_GA3C:
                act[i] = int(actions[t, i])
            elif pol == POL_LEARNING:
                act[i] = np.array(cont_actions[t, i], dtype=np.float64)
        try:
            obs, rew, over, info = env.step(act)
        except IndexError:
            # reference crashes in _update_state_history for very short goals
            return None
        d = wrap.observation(obs)
        for i, a in enumerate(agents):
            rec["pos"][t, i] = a.pos_global_frame
            rec["heading"][t, i] = a.heading_global_frame
            rec["vel"][t, i] = a.vel_global_frame
            rec["goal"][t, i] = a.goal_global_frame
            rec["t_rem"][t, i] = a.time_remaining_to_reach_goal
            rec["flags"][t, i] = [int(bool(getattr(a, f))) for f in FLAG_NAMES]
            rec["dist_to_goal"][t, i] = a.dist_to_goal
            rec["heading_ego"][t, i] = a.heading_ego_frame
            assert bool(info["which_agents_done"][a.id]) == bool(a.is_done)
        if np.ndim(rew) == 0:
            rec["reward"][t, 0] = rew
        else:
            rec["reward"][t] = rew
        rec["obs"][t] = d
        rec["sorted"][t] = sorted_now()
        rec["game_over"][t] = int(bool(over))
        steps = t + 1
        if over:
            break
    rec["steps"] = np.int32(steps)
    rec["actions"] = np.asarray(actions, dtype=np.int32)
    if cont_actions is not None:
        rec["cont_actions"] = np.asarray(cont_actions, dtype=np.float64)
    rec["mode"] = np.array(mode)
    return rec


def meta_of(ns):
    C = ns.Config
    env = ns.CollisionAvoidanceEnv()
    return dict(
        A=np.int32(C.MAX_NUM_AGENTS_IN_ENVIRONMENT), M=np.int32(C.MAX_NUM_OTHER_AGENTS_OBSERVED),
        dt=np.float64(C.DT), near_goal_threshold=np.float64(C.NEAR_GOAL_THRESHOLD),
        getting_close_range=np.float64(C.GETTING_CLOSE_RANGE),
        reward_at_goal=np.float64(C.REWARD_AT_GOAL),
        reward_collision_with_agent=np.float64(C.REWARD_COLLISION_WITH_AGENT),
        reward_time_step=np.float64(C.REWARD_TIME_STEP),
        min_possible_reward=np.float64(env.min_possible_reward),
        max_possible_reward=np.float64(env.max_possible_reward),
        max_time_ratio=np.float64(C.MAX_TIME_RATIO),
        sort_method=np.array(C.AGENT_SORTING_METHOD),
        evaluate_mode=np.int32(bool(C.EVALUATE_MODE)),
        train_single_agent=np.int32(bool(C.TRAIN_SINGLE_AGENT)),
        states_in_obs=np.array(list(C.STATES_IN_OBS)),
    )


def _mk_init(case, headings, policies):
    """case rows are [px,py,gx,gy,pref_speed,radius] (GCA/envs/test_cases.py:328-345)."""
    n = case.shape[0]
    init = np.zeros((n, 8))
    init[:, :6] = case
    init[:, 6] = headings
    init[:, 7] = policies
    return init


def build_scenarios(kind, ref_root, A):
    """Returns list of (name, init, actions, mode, cont_actions)."""
    out = []
    T = 120
    if kind == "phase1":
        # BASELINE config #1: get_testcase_two_agents geometry (GCA/envs/test_cases.py:77-84),
        # 100 random discrete actions from RandomState(0) (recipe: experiments/src/example.py:28-53)
        init = np.array([[-3., -3., 3., 3., 1.0, 0.5, 0.0, 0], [3., 3., -3., -3., 1.0, 0.5, np.pi, 0]])
        rng = np.random.RandomState(0)
        acts = np.array([[rng.randint(11), rng.randint(11)] for _ in range(100)])
        out.append(("config1_modeB", init, acts, "B", None))
        out.append(("config1_modeA", init, acts, "A", None))
        cases4 = _load_cases(ref_root, 4)
        hrng = np.random.default_rng(1234)
        arng = np.random.default_rng(5678)
        for k in range(24):
            init = _mk_init(cases4[k], hrng.uniform(-np.pi, np.pi, 4), np.zeros(4))
            out.append(("rand4_%02d" % k, init, arng.integers(0, 11, (T, 4)), "B", None))
        # goal-seeking actions: mostly straight/full speed so agents reach goals and collide
        for k in range(24, 40):
            c = cases4[k]
            h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0])
            init = _mk_init(c, h, np.zeros(4))
            acts = arng.choice([2, 2, 2, 2, 1, 3, 6], size=(T, 4))
            out.append(("seek4_%02d" % k, init, acts, "B", None))
        # mixed policies (noncoop / learning_ga3c / static), headings random
        prng = np.random.default_rng(42)
        for k in range(40, 56):
            pol = prng.choice([POL_NONCOOP, POL_LEARNING_GA3C, POL_STATIC], size=4, p=[0.3, 0.4, 0.3])
            if POL_LEARNING_GA3C not in pol:
                pol[prng.integers(4)] = POL_LEARNING_GA3C
            c = cases4[k]
            h = np.where(prng.random(4) < 0.5, np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0]),
                         prng.uniform(-np.pi, np.pi, 4))
            init = _mk_init(c, h, pol)
            acts = arng.choice([2, 2, 2, 1, 3, 6, 0, 4, 9], size=(T, 4))
            out.append(("mixed4_%02d" % k, init, acts, "B", None))
        # ragged worlds: 2 and 3 agents in a 4-agent env
        for n_ag in (2, 3):
            cs = _load_cases(ref_root, n_ag)
            for k in range(6):
                c = cs[k]
                h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0]) + prng.normal(0, 0.3, n_ag)
                init = _mk_init(c, h, np.zeros(n_ag))
                acts = arng.choice([2, 2, 2, 1, 3, 6, 5, 7], size=(T, n_ag))
                out.append(("ragged%d_%02d" % (n_ag, k), init, acts, "B", None))
        # column of 3, boundary gaps 0.15 m, driving straight (close-range penalty asymmetry)
        init = np.array([[0., 0., 10., 0., 1.0, 0.5, 0.0, 0],
                         [0., 1.15, 10., 1.15, 1.0, 0.5, 0.0, 0],
                         [0., 2.30, 10., 2.30, 1.0, 0.5, 0.0, 0]])
        out.append(("column3", init, np.full((40, 3), 2), "B", None))
        # head-on collision
        init = np.array([[-2., 0., 2., 0., 1.0, 0.5, 0.0, 0], [2., 0.05, -2., 0.05, 1.2, 0.4, np.pi, 0]])
        out.append(("headon2", init, np.full((40, 2), 2), "B", None))
        # parked agent times out, later gets hit
        init = np.array([[0., 0., 1.2, 0., 1.0, 0.5, 0.0, 0], [-4., 0.1, 4., 0.1, 1.0, 0.5, 0.0, 0]])
        acts = np.stack([np.full(60, 9), np.full(60, 2)], axis=1)
        out.append(("parked_hit", init, acts, "B", None))
        # continuous-action LearningPolicy (example.py:44 drives [1, 0.5]) + noncoop partner
        init = np.array([[-3., -3., 3., 3., 1.0, 0.5, 0.0, POL_LEARNING], [3., 3., -3., -3., 1.0, 0.5, np.pi, POL_NONCOOP]])
        cont = np.zeros((80, 2, 2))
        crng = np.random.default_rng(7)
        cont[:, 0, 0] = crng.uniform(0.3, 1.0, 80)
        cont[:, 0, 1] = crng.uniform(0.2, 0.8, 80)
        out.append(("continuous2", init, np.zeros((80, 2), dtype=int), "B", cont))
        # 4-agent Mode A (as-shipped float32 contamination), a few cases
        for k in range(56, 60):
            init = _mk_init(cases4[k], hrng.uniform(-np.pi, np.pi, 4), np.zeros(4))
            out.append(("rand4_modeA_%02d" % k, init, arng.integers(0, 11, (T, 4)), "A", None))
    elif kind == "phase2":
        cases10 = _load_cases(ref_root, 10)
        hrng = np.random.default_rng(4321)
        arng = np.random.default_rng(8765)
        prng = np.random.default_rng(99)
        for k in range(6):
            init = _mk_init(cases10[k], hrng.uniform(-np.pi, np.pi, 10), np.zeros(10))
            out.append(("rand10_%02d" % k, init, arng.integers(0, 11, (T, 10)), "B", None))
        for k in range(6, 12):
            c = cases10[k]
            h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0])
            pol = prng.choice([POL_NONCOOP, POL_LEARNING_GA3C, POL_STATIC], size=10, p=[0.2, 0.6, 0.2])
            pol[0] = POL_LEARNING_GA3C
            init = _mk_init(c, h, pol)
            out.append(("seekmixed10_%02d" % k, init, arng.choice([2, 2, 2, 1, 3, 6], size=(T, 10)), "B", None))
        for n_ag in (5, 6, 8):
            cs = _load_cases(ref_root, n_ag)
            for k in range(3):
                c = cs[k]
                h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0]) + prng.normal(0, 0.2, n_ag)
                init = _mk_init(c, h, np.zeros(n_ag))
                out.append(("ragged%d_%02d" % (n_ag, k), init, arng.choice([2, 2, 1, 3, 6, 5, 7], size=(T, n_ag)), "B", None))
    elif kind in ("closest_last", "tti"):
        cases4 = _load_cases(ref_root, 4)
        hrng = np.random.default_rng(11)
        arng = np.random.default_rng(12)
        for k in range(100, 112):
            c = cases4[k]
            h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0]) + hrng.normal(0, 0.3, 4)
            init = _mk_init(c, h, np.zeros(4))
            out.append(("case4_%03d" % k, init, arng.choice([2, 2, 2, 1, 3, 6, 0, 4], size=(T, 4)), "B", None))
    elif kind in ("clip", "clip_last"):
        cases6 = _load_cases(ref_root, 6)
        hrng = np.random.default_rng(21)
        arng = np.random.default_rng(22)
        for k in range(12):
            c = cases6[k]
            h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0]) + hrng.normal(0, 0.3, 6)
            init = _mk_init(c, h, np.zeros(6))
            out.append(("case6_%03d" % k, init, arng.choice([2, 2, 2, 1, 3, 6, 0, 4], size=(T, 6)), "B", None))
    elif kind == "evaluate":
        # EVALUATE_MODE: dt 0.1, game_over when ALL agents are done
        cases8 = _load_cases(ref_root, 8)
        arng = np.random.default_rng(32)
        prng = np.random.default_rng(33)
        for k in range(4):
            c = cases8[k]
            h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0])
            pol = prng.choice([POL_NONCOOP, POL_LEARNING_GA3C], size=8, p=[0.5, 0.5])
            init = _mk_init(c, h, pol)
            out.append(("eval8_%02d" % k, init, arng.choice([2, 2, 2, 1, 3], size=(400, 8)), "B", None))
    elif kind == "single":
        cases4 = _load_cases(ref_root, 4)
        arng = np.random.default_rng(52)
        for k in range(200, 206):
            c = cases4[k]
            h = np.arctan2(c[:, 3] - c[:, 1], c[:, 2] - c[:, 0])
            init = _mk_init(c, h, np.zeros(4))
            out.append(("single4_%03d" % k, init, arng.choice([2, 2, 2, 1, 3], size=(T, 4)), "B", None))
    else:
        raise ValueError(kind)
    return out


KINDS = {
    # kind: (GYM_CONFIG_CLASS, GYM_CONFIG_PATH relative to the reference or to oracle/)
    "phase1": ("TrainPhase1", "ref:ga3c/GA3C/Config.py"),
    "phase2": ("TrainPhase2", "ref:ga3c/GA3C/Config.py"),
    "closest_last": ("ClosestLast4", "oracle:ref_configs.py"),
    "tti": ("TimeToImpact4", "oracle:ref_configs.py"),
    "clip": ("Clip6Obs3", "oracle:ref_configs.py"),
    "clip_last": ("Clip6Obs3ClosestLast", "oracle:ref_configs.py"),
    "evaluate": ("Evaluate19", "oracle:ref_configs.py"),
    "single": ("SingleAgent4", "oracle:ref_configs.py"),
}


def generate(kind):
    sys.path.insert(0, HERE)
    import ref_harness as rh
    cls, path = KINDS[kind]
    where, rel = path.split(":")
    cfg_path = os.path.join(rh.REFERENCE_ROOT if where == "ref" else HERE, rel)
    rh.install(config_class=cls, config_path=cfg_path)
    ns = rh.reference_modules()
    arrays = {"meta_" + k: v for k, v in meta_of(ns).items()}
    names = []
    for name, init, acts, mode, cont in build_scenarios(kind, rh.REFERENCE_ROOT, int(ns.Config.MAX_NUM_AGENTS_IN_ENVIRONMENT)):
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            rec = run_case(ns, init, acts, mode=mode, cont_actions=cont)
        if rec is None:
            print("  skipped %s (reference IndexError, SURVEY N8)" % name)
            continue
        s = int(rec["steps"])
        for k, v in rec.items():
            if isinstance(v, np.ndarray) and v.ndim >= 1 and k in ("pos", "heading", "vel", "goal", "t_rem", "flags",
                                                                     "reward", "obs", "sorted", "game_over",
                                                                     "dist_to_goal", "heading_ego", "actions", "cont_actions"):
                v = v[:s]
            arrays["%s/%s" % (name, k)] = v
        names.append(name)
        print("  %-20s n=%d steps=%d game_over=%d" % (name, init.shape[0], s, int(rec["game_over"][s - 1])))
    arrays["names"] = np.array(names)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = os.path.join(GOLDEN_DIR, "%s.npz" % kind)
    np.savez_compressed(out, **arrays)
    print("wrote %s (%.1f KB, %d cases)" % (out, os.path.getsize(out) / 1024.0, len(names)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default=None, help="one of %s (default: all, each in its own process)" % list(KINDS))
    args = ap.parse_args()
    if args.kind:
        generate(args.kind)
        return
    for kind in KINDS:
        print("== %s" % kind)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--kind", kind])


if __name__ == "__main__":
    main()
