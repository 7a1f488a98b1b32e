"""TEST INFRASTRUCTURE — CPU restatement of the GA3C actor's experience bookkeeping
(ProcessAgent.run_episode, GA3C/ProcessAgent.py:105-211, and _accumulate_rewards, :54-79), pinned against
tests/golden/ga3c_actor.npz which was recorded from the unmodified reference (oracle/gen_golden_ga3c.py).
Pure-Python loops: only for small cases."""
import numpy as np


class _Exp(object):
    __slots__ = ("x", "a", "r")

    def __init__(self, x, a, r):
        self.x, self.a, self.r = x, a, r


def _accumulate(exps, gamma, terminal_reward, done, time_max):
    """ProcessAgent._accumulate_rewards (:54-79): returns (emitted, leftover)."""
    if len(exps) == 1:
        return exps, None
    leftover = None
    returned = exps[:-1]
    n_exps = len(exps) - 1
    if done and len(exps) == time_max + 1:
        leftover = [exps[-1]]
    if done and len(exps) != time_max + 1:
        n_exps = len(exps)
        returned = exps
    R = terminal_reward
    for t in reversed(range(n_exps)):
        R = gamma * R + exps[t].r
        exps[t].r = R  # the reference overwrites the stored reward with the return
    return returned, leftover


def actor_rows(obs, rewards, done, values, actions, time_max, gamma):
    """One episode.  obs [T+1, A, L] (column 0 = is_learning), rewards/done [T, n], values/actions [T, A].
    Returns a list over steps of lists of (agent, x[L-1], r, a) rows emitted at that step."""
    T = rewards.shape[0]
    n = rewards.shape[1]
    learning = [i for i in range(n) if obs[0, i, 0] != 0]
    lists = {i: [] for i in learning}
    tcount = {i: 0 for i in learning}
    trained = {i: False for i in learning}
    out = []
    for t in range(T):
        rows = []
        for i in learning:
            d = bool(done[t, i])
            lists[i].append(_Exp(obs[t, i, 1:].copy(), int(actions[t, i]), float(rewards[t, i])))
            if d or (tcount[i] == time_max and not trained[i]):  # precedence as in ProcessAgent.py:186
                if d:
                    term = 0.0
                    trained[i] = True
                else:
                    term = float(values[t, i])
                emitted, leftover = _accumulate(lists[i], gamma, term, d, time_max)
                for e in emitted:
                    rows.append((i, e.x, e.r, e.a))
                if leftover is not None:
                    for e in leftover:
                        rows.append((i, e.x, e.r, e.a))
                tcount[i] = 0
                lists[i] = [lists[i][-1]]
            tcount[i] += 1
        out.append(rows)
    return out
