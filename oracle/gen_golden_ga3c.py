"""TEST INFRASTRUCTURE — golden vectors for the GA3C actor bookkeeping, recorded from the UNMODIFIED
reference `ProcessAgent.run_episode` (GA3C/ProcessAgent.py:105-211) driven by a scripted environment and
a scripted predictor (so that only the experience / n-step-return logic is exercised).

    python oracle/gen_golden_ga3c.py        # writes tests/golden/ga3c_actor.npz (build container only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ga3c_actor.npz")


class ScriptedEnv(object):
    """Stands in for GA3C/Environment.py: replays obs/reward/done arrays."""

    def __init__(self, obs, rewards, done, learning):
        self.obs, self.rewards, self.done, self.learning = obs, rewards, done, learning
        self.t = 0
        self.previous_state = self.current_state = None

    def _set(self, o):
        self.latest_observations = o
        self.previous_state = self.current_state
        self.current_state = o[None, :, 1:].copy()

    def reset(self):
        self.t = 0
        self.current_state = None
        self._set(self.obs[0])

    def step(self, action, pid, count):
        t = self.t
        n = self.rewards.shape[1]
        self.taken = dict(action[0])
        info = {'which_agents_done': {i: bool(self.done[t, i]) for i in range(n)},
                'which_agents_learning': {i: bool(self.learning[i]) for i in range(n)}}
        over = all(self.done[t, i] for i in range(n) if self.learning[i])
        self.t += 1
        self._set(self.obs[t + 1])
        return [self.rewards[t].astype(np.float64)], over, [info]


def make_script(rng, A, n, T, L, done_at):
    learning = np.zeros(A, dtype=bool)
    learning[:n] = rng.random(n) < 0.8
    if not learning.any():
        learning[0] = True
    obs = rng.normal(size=(T + 1, A, L)).astype(np.float32)
    obs[:, :, 0] = learning.astype(np.float32)
    obs[:, n:, :] = 0
    rewards = np.where(rng.random((T, n)) < 0.2, rng.uniform(-0.25, 1.0, (T, n)), 0.0)
    done = np.zeros((T, n), dtype=bool)
    for i in range(n):
        done[done_at[i]:, i] = True
    # the episode ends when every learning agent is done
    last = max(done_at[i] for i in range(n) if learning[i])
    T_eff = last + 1
    values = rng.normal(size=(T, A)).astype(np.float32)
    probs = rng.dirichlet(np.ones(11), size=(T, A)).astype(np.float64)
    return dict(learning=learning, obs=obs[:T_eff + 1], rewards=rewards[:T_eff], done=done[:T_eff],
                values=values[:T_eff], probs=probs[:T_eff], n=n)


def run_reference(PA, script, A, seed):
    agent = PA(0, None, None, None, 11)
    agent.env = ScriptedEnv(script["obs"], script["rewards"], script["done"], script["learning"][:script["n"]])
    state = {"k": 0}
    T = script["rewards"].shape[0]

    def predict(obs_row):
        # called once per learning agent per step, in agent order
        t = agent.env.t
        i = state["order"][state["k"] % len(state["order"])]
        state["k"] += 1
        return script["probs"][t, i], script["values"][t, i]
    state["order"] = [i for i in range(A) if script["obs"][0, i, 0] != 0]
    agent.predict = predict
    np.random.seed(seed)
    rows = []   # (emit_step, x, r, a)
    actions = -np.ones((T, A), dtype=np.int32)
    gen = agent.run_episode()
    orig_step = agent.env.step

    def step(action, pid, count):
        t = agent.env.t
        for i, a in action[0].items():
            actions[t, i] = a
        return orig_step(action, pid, count)
    agent.env.step = step
    yields = []
    for x_, r_, a_, reward_sum in gen:
        yields.append((agent.env.t - 1, np.array(x_, dtype=np.float32), np.array(r_, dtype=np.float64),
                       np.argmax(a_, axis=1).astype(np.int32), float(reward_sum)))
    return actions, yields


def main():
    rh.install(config_class="TrainPhase1", config_path=os.path.join(rh.GA3C_ROOT, "GA3C", "Config.py"))
    from GA3C import Config
    from ProcessAgent import ProcessAgent
    A = Config.MAX_NUM_AGENTS_IN_ENVIRONMENT
    L = 6 + 7 * Config.MAX_NUM_OTHER_AGENTS_OBSERVED
    TMAX = Config.TIME_MAX
    rng = np.random.default_rng(2718)
    arrays = {"meta_A": np.int32(A), "meta_L": np.int32(L), "meta_time_max": np.int32(TMAX),
              "meta_gamma": np.float64(Config.DISCOUNT)}
    specs = []
    # handcrafted: done on first step, done exactly when the list holds TIME_MAX+1, long survivors, early finisher
    specs.append((4, [0, 5, 30, 63]))
    specs.append((4, [TMAX, 2 * TMAX + 1, 3, 45]))
    specs.append((2, [TMAX - 1, TMAX + 1]))
    specs.append((3, [41, 20, 21]))
    for _ in range(20):
        n = int(rng.integers(2, A + 1))
        specs.append((n, [int(v) for v in rng.integers(0, 70, n)]))
    names = []
    for k, (n, done_at) in enumerate(specs):
        script = make_script(rng, A, n, 75, L, done_at)
        actions, yields = run_reference(ProcessAgent, script, A, seed=k)
        name = "ep%02d" % k
        names.append(name)
        for key in ("learning", "obs", "rewards", "done", "values", "probs"):
            arrays["%s/%s" % (name, key)] = script[key]
        arrays["%s/n" % name] = np.int32(n)
        arrays["%s/actions" % name] = actions
        arrays["%s/emit_step" % name] = np.concatenate([np.full(len(r), t, dtype=np.int32) for t, _, r, _, _ in yields])
        arrays["%s/emit_x" % name] = np.concatenate([x for _, x, _, _, _ in yields], axis=0)
        arrays["%s/emit_r" % name] = np.concatenate([r for _, _, r, _, _ in yields])
        arrays["%s/emit_a" % name] = np.concatenate([a for _, _, _, a, _ in yields])
        arrays["%s/yield_len" % name] = np.array([len(r) for _, _, r, _, _ in yields], dtype=np.int32)
        arrays["%s/yield_reward_sum" % name] = np.array([s for _, _, _, _, s in yields])
        print("%s n=%d steps=%d yields=%d rows=%d" % (name, n, script["rewards"].shape[0], len(yields),
                                                       len(arrays["%s/emit_r" % name])))
    arrays["names"] = np.array(names)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
