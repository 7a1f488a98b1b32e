"""Helpers shared by the CPU (oracle) and GPU (CUDA) parity tests: load golden vectors recorded
from the unmodified reference (oracle/gen_golden.py) and replay them through an env that follows
the tensor contract of include/ca_step.h."""
import os

import numpy as np

from rl_collision_avoidance_b200 import _abi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_KINDS = ["phase1", "phase2", "closest_last", "tti", "clip", "clip_last", "evaluate", "single"]


class Golden(object):
    def __init__(self, kind):
        self.kind = kind
        self.z = np.load(os.path.join(GOLDEN_DIR, kind + ".npz"), allow_pickle=False)
        self.names = [str(s) for s in self.z["names"]]
        self.meta = {k[5:]: self.z[k] for k in self.z.files if k.startswith("meta_")}
        self.A = int(self.meta["A"])
        self.M = int(self.meta["M"])
        self.L = _abi.obs_len(self.M)
        assert [str(s) for s in self.meta["states_in_obs"]] == [
            "is_learning", "num_other_agents", "dist_to_goal", "heading_ego_frame", "pref_speed", "radius",
            "other_agents_states"]

    def get(self, name, key):
        return self.z["%s/%s" % (name, key)]

    def has(self, name, key):
        return "%s/%s" % (name, key) in self.z.files

    def config(self, num_worlds, **over):
        m = self.meta
        mode = _abi.OVER_ALL_LEARNING_DONE
        if int(m["evaluate_mode"]):
            mode = _abi.OVER_ALL_DONE
        elif int(m["train_single_agent"]):
            mode = _abi.OVER_FIRST_AGENT_DONE
        cfg = _abi.default_config(
            num_worlds, self.A, self.M,
            sort_method=_abi.SORT_METHODS[str(m["sort_method"])], game_over_mode=mode,
            dt=float(m["dt"]), near_goal_threshold=float(m["near_goal_threshold"]),
            getting_close_range=float(m["getting_close_range"]), reward_at_goal=float(m["reward_at_goal"]),
            reward_collision_with_agent=float(m["reward_collision_with_agent"]),
            reward_time_step=float(m["reward_time_step"]), min_possible_reward=float(m["min_possible_reward"]),
            max_possible_reward=float(m["max_possible_reward"]), max_time_ratio=float(m["max_time_ratio"]))
        for k, v in over.items():
            setattr(cfg, k, v)
        return cfg

    def cases(self, mode="B"):
        return [n for n in self.names if str(self.get(n, "mode")) == mode]

    def batch(self, names):
        """Stack cases into the C-ABI init tensor: returns init[Wc,A,INIT_STRIDE], num_agents[Wc], T_max."""
        Wc = len(names)
        init = np.zeros((Wc, self.A, _abi.INIT_STRIDE))
        nag = np.zeros((Wc,), dtype=np.int32)
        T = 0
        for w, name in enumerate(names):
            g = self.get(name, "init")
            n = g.shape[0]
            nag[w] = n
            init[w, :n, _abi.I_PX] = g[:, 0]
            init[w, :n, _abi.I_PY] = g[:, 1]
            init[w, :n, _abi.I_GX] = g[:, 2]
            init[w, :n, _abi.I_GY] = g[:, 3]
            init[w, :n, _abi.I_PREF_SPEED] = g[:, 4]
            init[w, :n, _abi.I_RADIUS] = g[:, 5]
            init[w, :n, _abi.I_HEADING] = g[:, 6]
            init[w, :n, _abi.I_POLICY] = g[:, 7]
            init[w, :n, _abi.I_TIME_REMAINING] = self.get(name, "t_rem0")
            T = max(T, int(self.get(name, "steps")))
        return init, nag, T


def obs_abs_diff(a, b):
    """|a - b| per observation element, with column 3 (heading_ego_frame, an angle in [-pi, pi)) compared on the
    circle: -pi and +pi are the same heading.  The reference itself lands on either side of that seam depending
    on the last ulp of atan2 (e.g. a non-cooperative agent that overshoots its goal faces exactly away from it),
    so a 2*pi jump there is a representation artefact, not a discrepancy."""
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    d[..., 3] = np.minimum(d[..., 3], np.abs(2 * np.pi - d[..., 3]))
    return d


def assert_obs_close(a, b, atol, err_msg=""):
    d = obs_abs_diff(a, b)
    if not np.all(d <= atol):
        idx = np.unravel_index(np.argmax(d), d.shape)
        raise AssertionError("%s: max obs abs diff %.3e at %s (tol %.1e): %r vs %r" % (
            err_msg, d[idx], idx, atol, np.asarray(a)[idx], np.asarray(b)[idx]))


GOLDEN_FLAG_BITS = (_abi.F_AT_GOAL, _abi.F_WAS_AT_GOAL, _abi.F_IN_COLLISION, _abi.F_WAS_IN_COLLISION,
                    _abi.F_RAN_OUT_OF_TIME)


def replay_and_compare(gold, names, env, state_tol, obs_tol, reward_tol, check_state=True):
    """env: object with set_world_state/reset/step/get_state and attributes obs, reward, done, game_over,
    sorted_idx (numpy, shapes of include/ca_step.h).  Asserts flags / done / game_over / neighbour
    indices bit-exact and floats within the given absolute tolerances, for every step of every case."""
    A, M = gold.A, gold.M
    init, nag, T = gold.batch(names)
    env.set_world_state(init, nag)
    env.reset()
    obs = np.asarray(env.obs, dtype=np.float64)
    for w, name in enumerate(names):
        n = nag[w]
        assert_obs_close(obs[w], gold.get(name, "obs0"), obs_tol, "%s obs0" % name)
        np.testing.assert_array_equal(np.asarray(env.sorted_idx)[w, :n], gold.get(name, "sorted0"), err_msg="%s sorted0" % name)
    checked = 0
    for t in range(T):
        actions = np.zeros((len(names), A), dtype=np.int32)
        cont = None
        for w, name in enumerate(names):
            if t < int(gold.get(name, "steps")):
                a = gold.get(name, "actions")[t]
                actions[w, :a.shape[0]] = a
                if gold.has(name, "cont_actions"):
                    if cont is None:
                        cont = np.zeros((len(names), A, 2))
                    ca = gold.get(name, "cont_actions")[t]
                    cont[w, :ca.shape[0]] = ca
        env.step(actions, cont)
        obs = np.asarray(env.obs, dtype=np.float64)
        rew = np.asarray(env.reward, dtype=np.float64)
        done = np.asarray(env.done)
        over = np.asarray(env.game_over)
        sidx = np.asarray(env.sorted_idx)
        st = env.get_state() if check_state else None
        for w, name in enumerate(names):
            if t >= int(gold.get(name, "steps")):
                continue
            n = nag[w]
            tag = "%s/%s step %d" % (gold.kind, name, t)
            gflags = gold.get(name, "flags")[t]
            # bit-exact: done, game_over, neighbour order, every agent flag
            np.testing.assert_array_equal(done[w, :n], gflags[:, 5], err_msg=tag + " done")
            assert np.all(done[w, n:] == 1), tag + " absent agents must read done"
            assert int(over[w]) == int(gold.get(name, "game_over")[t]), tag + " game_over"
            np.testing.assert_array_equal(sidx[w, :n], gold.get(name, "sorted")[t], err_msg=tag + " sorted idx")
            assert np.all(sidx[w, n:] == -1), tag
            np.testing.assert_allclose(rew[w, :n], gold.get(name, "reward")[t], rtol=0, atol=reward_tol, err_msg=tag + " reward")
            assert np.all(rew[w, n:] == 0), tag
            assert_obs_close(obs[w], gold.get(name, "obs")[t], obs_tol, tag + " obs")
            if st is not None:
                s = st[w, :n]
                fl = s[:, _abi.S_FLAGS].astype(np.int64)
                for b, bit in enumerate(GOLDEN_FLAG_BITS):
                    np.testing.assert_array_equal((fl & bit) != 0, gflags[:, b] != 0, err_msg=tag + " flag bit %d" % bit)
                np.testing.assert_allclose(s[:, [_abi.S_PX, _abi.S_PY]], gold.get(name, "pos")[t], rtol=0, atol=state_tol, err_msg=tag + " pos")
                np.testing.assert_allclose(s[:, _abi.S_HEADING], gold.get(name, "heading")[t], rtol=0, atol=state_tol, err_msg=tag + " heading")
                np.testing.assert_allclose(s[:, [_abi.S_VX, _abi.S_VY]], gold.get(name, "vel")[t], rtol=0, atol=state_tol, err_msg=tag + " vel")
                np.testing.assert_allclose(s[:, [_abi.S_GX, _abi.S_GY]], gold.get(name, "goal")[t], rtol=0, atol=state_tol, err_msg=tag + " goal")
                np.testing.assert_allclose(s[:, _abi.S_TIME_REMAINING], gold.get(name, "t_rem")[t], rtol=0, atol=state_tol, err_msg=tag + " t_rem")
            checked += 1
    return checked
