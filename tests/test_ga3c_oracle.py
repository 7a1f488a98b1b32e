"""CPU: the Python restatement of the GA3C actor bookkeeping (oracle/ga3c_oracle.py) reproduces the rows that the
UNMODIFIED reference ProcessAgent.run_episode yielded for scripted episodes (tests/golden/ga3c_actor.npz)."""
import os

import numpy as np

from oracle import ga3c_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ga3c_actor.npz")


def load_ga3c_golden():
    z = np.load(GOLD)
    eps = []
    for name in [str(s) for s in z["names"]]:
        eps.append({k: z["%s/%s" % (name, k)] for k in ("learning", "obs", "rewards", "done", "values", "probs", "n",
                                                         "actions", "emit_step", "emit_x", "emit_r", "emit_a",
                                                         "yield_len", "yield_reward_sum")})
        eps[-1]["name"] = name
    meta = dict(A=int(z["meta_A"]), L=int(z["meta_L"]), time_max=int(z["meta_time_max"]), gamma=float(z["meta_gamma"]))
    return meta, eps


def canonical_rows(rows):
    """Sort emitted rows so that two implementations can be compared as multisets per step."""
    return sorted(rows, key=lambda r: (r[0], round(float(r[2]), 4), int(r[3]), float(r[1][0])))


def test_restatement_matches_reference_process_agent():
    meta, eps = load_ga3c_golden()
    assert meta["time_max"] == 20 and abs(meta["gamma"] - 0.97) < 1e-12
    total = 0
    for ep in eps:
        out = ga3c_oracle.actor_rows(ep["obs"], ep["rewards"], ep["done"], ep["values"], ep["actions"],
                                     meta["time_max"], meta["gamma"])
        T = ep["rewards"].shape[0]
        # golden rows are in yield order: agent order within a step, list order within a yield -> same as ours
        k = 0
        for t in range(T):
            sel = np.nonzero(ep["emit_step"] == t)[0]
            assert len(sel) == len(out[t]), "%s step %d: %d rows vs reference %d" % (ep["name"], t, len(out[t]), len(sel))
            for j, row in zip(sel, out[t]):
                np.testing.assert_array_equal(ep["emit_x"][j], row[1])
                assert abs(ep["emit_r"][j] - row[2]) <= 1e-12
                assert int(ep["emit_a"][j]) == row[3]
                k += 1
        assert k == len(ep["emit_r"])
        total += k
    assert total > 2000
