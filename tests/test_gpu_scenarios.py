"""GPU: the on-device scenario generator (ca_generate_scenarios) has the distribution of the UNMODIFIED reference's
get_testcase_random (statistics recorded by oracle/gen_golden_scenarios.py): case-type mix, agent-count and policy
mix, quantiles of speeds, radii, start-goal distances, spacing and time budgets.  Tolerances: fractions +-0.02 abs,
quantiles 6 % relative (+0.03 abs) — Monte-Carlo noise of 6000 reference worlds vs 30000 generated ones."""
import json
import os

import numpy as np
import pytest

from rl_collision_avoidance_b200 import _abi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _stats(st, nag):
    q = [5, 25, 50, 75, 95]
    W, A = nag.shape[0], st.shape[1]
    live = np.arange(A)[None, :] < nag[:, None]
    px, py, gx, gy = st[..., _abi.S_PX], st[..., _abi.S_PY], st[..., _abi.S_GX], st[..., _abi.S_GY]
    d_sg = np.hypot(px - gx, py - gy)[live]
    dmin = np.full(W, np.inf)
    for i in range(A):
        for j in range(i + 1, A):
            d = np.hypot(px[:, i] - px[:, j], py[:, i] - py[:, j])
            ok = (nag > j)
            dmin = np.where(ok, np.minimum(dmin, d), dmin)
    swap = (py[:, 0] == 0) & (gy[:, 0] == 0) & (px[:, 0] == -gx[:, 0]) & (px[:, 1] == -px[:, 0])
    sym = np.all(np.where(live, np.abs(px + gx) + np.abs(py + gy), 0.0) < 1e-9, axis=1)
    pol = st[..., _abi.S_POLICY][live]
    return {
        "num_agents_hist": {int(k): float(np.mean(nag == k)) for k in np.unique(nag)},
        "policy_frac": {int(k): float(np.mean(pol == k)) for k in (0, 1, 2)},
        "worlds_with_learner": float(np.mean(np.any((st[..., _abi.S_POLICY] == 0) & live, axis=1))),
        "frac_swap": float(np.mean(swap)), "frac_circle": float(np.mean(sym & ~swap)),
        "q_pref_speed": np.percentile(st[..., _abi.S_PREF_SPEED][live], q), "q_radius": np.percentile(st[..., _abi.S_RADIUS][live], q),
        "q_start_goal_dist": np.percentile(d_sg, q), "q_abs_start_x": np.percentile(np.abs(px[live]), q),
        "q_min_pair_start_dist": np.percentile(dmin, q), "q_time_remaining": np.percentile(st[..., _abi.S_TIME_REMAINING][live], q),
        "q_heading": np.percentile(st[..., _abi.S_HEADING][live], q),
    }


@pytest.mark.parametrize("cls,A", [("TrainPhase1", 4), ("TrainPhase2", 10)])
def test_generator_matches_reference_distribution(cls, A):
    import torch
    from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv
    ref = json.load(open(os.path.join(GOLD, "scenario_stats_%s.json" % cls)))
    W = 30000
    env = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1))
    sc = env.scenario_config({'policy_to_ensure': 'learning_ga3c', 'policies': ['noncoop', 'learning_ga3c', 'static'],
                              'policy_distr': [0.05, 0.9, 0.05], 'speed_bnds': [0.5, 2.0], 'radius_bnds': [0.2, 0.8]})
    env.generate_scenarios(sc, seed=2024)
    env.reset()
    st = env.get_state()
    nag = np.rint(env.obs[:, 0, 1].cpu().numpy()).astype(int) + 1          # num_other_agents + 1
    got = _stats(st, nag)
    for k, v in ref["num_agents_hist"].items():
        assert abs(got["num_agents_hist"].get(int(k), 0.0) - v) < 0.02, ("num_agents", k)
    # policy ids: reference golden uses 0=learning, 1=noncoop, 2=static like the C-ABI
    for k, v in ref["policy_frac"].items():
        assert abs(got["policy_frac"][int(k)] - v) < 0.02, ("policy", k, got["policy_frac"], v)
    assert got["worlds_with_learner"] == 1.0 == ref["worlds_with_learner"]
    assert abs(got["frac_swap"] - ref["frac_swap"]) < 0.02 and abs(got["frac_circle"] - ref["frac_circle"]) < 0.02
    for key in ("q_pref_speed", "q_radius", "q_start_goal_dist", "q_abs_start_x", "q_min_pair_start_dist",
                "q_time_remaining"):
        np.testing.assert_allclose(got[key], ref[key], rtol=0.06, atol=0.03, err_msg=key)
    np.testing.assert_allclose(got["q_heading"], ref["q_heading"], rtol=0, atol=0.08)
    # hard constraints of the generator
    live = np.arange(A)[None, :] < nag[:, None]
    assert np.all(st[..., _abi.S_PREF_SPEED][live] >= 0.5) and np.all(st[..., _abi.S_PREF_SPEED][live] <= 2.0)
    assert np.all(st[..., _abi.S_RADIUS][live] >= 0.2) and np.all(st[..., _abi.S_RADIUS][live] <= 0.8)
    for i in range(A):
        for j in range(i + 1, A):
            ok = nag > j
            d = np.hypot(st[:, i, _abi.S_PX] - st[:, j, _abi.S_PX], st[:, i, _abi.S_PY] - st[:, j, _abi.S_PY])
            lim = st[:, i, _abi.S_RADIUS] + st[:, j, _abi.S_RADIUS] + 0.2
            assert np.all(d[ok] >= lim[ok] - 1e-12), "starts closer than r_i + r_j + 0.2"
    env.close()


def test_consumed_worlds_get_new_scenarios():
    """ca_reset / auto-reset mark a world's snapshot as consumed; generate(only_consumed=True) refills exactly those,
    so consecutive episodes of a world start from different scenarios, while without a refill they repeat."""
    import torch
    from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv
    W, A = 2048, 4
    gen = torch.Generator(device="cuda")
    for refill in (True, False):
        env = VecCollisionAvoidanceEnv(_abi.default_config(W, A, auto_reset=1))
        sc = env.scenario_config({'policies': 'learning_ga3c'})
        env.generate_scenarios(sc, seed=7)
        env.reset()
        last_start = env.get_state()[:, :, :2].copy()
        gen.manual_seed(1)
        same = changed = 0
        for t in range(150):
            if refill:
                env.generate_scenarios(sc, seed=7, only_consumed=True)
            a = torch.randint(0, 11, (W, A), dtype=torch.int32, device="cuda", generator=gen)
            _, _, _, over = env.step(a)
            o = over.cpu().numpy().astype(bool)
            if o.any():
                st = env.get_state()[:, :, :2]
                for w in np.nonzero(o)[0]:
                    if np.array_equal(st[w], last_start[w]):
                        same += 1
                    else:
                        changed += 1
                    last_start[w] = st[w]
        assert same + changed > W
        if refill:
            assert same == 0, "every new episode must start from a fresh scenario"
        else:
            assert changed == 0, "without a refill a world repeats its snapshot"
        env.close()
