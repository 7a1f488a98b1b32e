"""TEST INFRASTRUCTURE — distribution statistics of the UNMODIFIED reference's scenario generator
(get_testcase_random with Config.TEST_CASE_ARGS, GCA/envs/test_cases.py:95-118) for the on-device generator test.

    python oracle/gen_golden_scenarios.py TrainPhase1|TrainPhase2     (build container only)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def stats_of(cases):
    """cases: list of dict(px,py,gx,gy,ps,rad,policy,heading,t0) arrays per world."""
    q = [5, 25, 50, 75, 95]
    n = np.array([len(c["px"]) for c in cases])
    cat = lambda k: np.concatenate([c[k] for c in cases])
    d_sg = np.concatenate([np.hypot(c["px"] - c["gx"], c["py"] - c["gy"]) for c in cases])
    min_pair = []
    swap = circle = 0
    for c in cases:
        m = len(c["px"])
        dd = [np.hypot(c["px"][i] - c["px"][j], c["py"][i] - c["py"][j]) for i in range(m) for j in range(i + 1, m)]
        min_pair.append(min(dd))
        is_swap = m >= 2 and c["py"][0] == 0 and c["gy"][0] == 0 and c["px"][0] == -c["gx"][0] and c["px"][1] == -c["px"][0]
        is_circle = (not is_swap) and np.allclose(c["px"], -c["gx"], atol=1e-9) and np.allclose(c["py"], -c["gy"], atol=1e-9)
        swap += is_swap
        circle += is_circle
    pol = cat("policy")
    out = {
        "worlds": len(cases),
        "num_agents_hist": {int(k): float(np.mean(n == k)) for k in np.unique(n)},
        "policy_frac": {int(k): float(np.mean(pol == k)) for k in (0, 1, 2)},
        "worlds_with_learner": float(np.mean([np.any(c["policy"] == 0) for c in cases])),
        "frac_swap": swap / len(cases), "frac_circle": circle / len(cases),
        "q_pref_speed": np.percentile(cat("ps"), q).tolist(), "q_radius": np.percentile(cat("rad"), q).tolist(),
        "q_start_goal_dist": np.percentile(d_sg, q).tolist(), "q_abs_start_x": np.percentile(np.abs(cat("px")), q).tolist(),
        "q_min_pair_start_dist": np.percentile(min_pair, q).tolist(), "q_time_remaining": np.percentile(cat("t0"), q).tolist(),
        "q_heading": np.percentile(cat("heading"), q).tolist(),
    }
    return out


def main():
    cls = sys.argv[1] if len(sys.argv) > 1 else "TrainPhase1"
    rh.install(config_class=cls, config_path=os.path.join(rh.GA3C_ROOT, "GA3C", "Config.py"))
    ns = rh.reference_modules()
    C = ns.Config
    np.random.seed(12345)
    pol_id = {"learning": 0, "NonCooperativePolicy": 1, "Static": 2}
    cases = []
    args = dict(C.TEST_CASE_ARGS)
    for _ in range(6000):
        agents = ns.test_cases.get_testcase_random(**args)
        cases.append(dict(px=np.array([a.pos_global_frame[0] for a in agents]), py=np.array([a.pos_global_frame[1] for a in agents]),
                          gx=np.array([a.goal_global_frame[0] for a in agents]), gy=np.array([a.goal_global_frame[1] for a in agents]),
                          ps=np.array([a.pref_speed for a in agents]), rad=np.array([a.radius for a in agents]),
                          policy=np.array([pol_id[a.policy.str] for a in agents]),
                          heading=np.array([float(a.heading_global_frame) for a in agents]),
                          t0=np.array([a.time_remaining_to_reach_goal for a in agents])))
    out = stats_of(cases)
    out["config"] = cls
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "scenario_stats_%s.json" % cls)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
