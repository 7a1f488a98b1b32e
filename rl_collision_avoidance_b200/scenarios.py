"""Synthetic scenario generation (host side, setup time only — not on the step hot path).

`random_worlds` draws W worlds at once with the same distributions as the reference's random test
cases (GCA/envs/test_cases.py:95-118 get_testcase_random -> policies/CADRL/scripts/multi/
gen_rand_testcases.py:137-226 generate_rand_case): radius ~ U[radius_bnds], pref_speed = max of two
U[speed_bnds] draws, start/goal ~ U[-side, side]^2 with the side growing 1% per rejected draw,
rejection of starts/goals closer than r_i + r_j + GETTING_CLOSE_RANGE to an earlier agent's, and
|start - goal| > side/2; heading ~ U(-pi, pi) (test_cases.py:315); side length by agent count
(GCA/envs/config.py:57-60).  Distributional parity only, and a subset: the swap/circle cases (15 % each) and the
"straight line must not already be a solution" rejection of the reference are reproduced by the ON-DEVICE generator
(csrc/ca_scenarios.cuh, ca_generate_scenarios — what training uses for every episode after the first), not here.
"""
import numpy as np

from . import _abi
from .vec_env import make_init

GETTING_CLOSE_RANGE = 0.2


def random_worlds(num_worlds, max_agents, rng, num_agents=None, speed_bnds=(0.5, 2.0), radius_bnds=(0.2, 0.8),
                  policies=("learning_ga3c",), policy_distr=None, policy_to_ensure=None):
    """Returns (init[W, A, INIT_STRIDE] float64, num_agents[W] int32) for ca_set_world_state."""
    W, A = int(num_worlds), int(max_agents)
    if num_agents is None:
        nag = np.full(W, A, dtype=np.int32)
    elif np.isscalar(num_agents):
        nag = np.full(W, int(num_agents), dtype=np.int32)
    else:
        nag = np.asarray(num_agents, dtype=np.int32)
    # side length by agent count, config.py:57-60
    side = np.where(nag < 5, rng.uniform(4.0, 5.0, W), rng.uniform(6.0, 8.0, W))
    radius = rng.uniform(radius_bnds[0], radius_bnds[1], (W, A))
    speed = np.maximum(rng.uniform(speed_bnds[0], speed_bnds[1], (W, A)), rng.uniform(speed_bnds[0], speed_bnds[1], (W, A)))
    start = np.zeros((W, A, 2))
    goal = np.zeros((W, A, 2))
    for i in range(A):
        todo = np.arange(W)
        while todo.size:
            side[todo] *= 1.01
            s = side[todo, None] * (2 * rng.random((todo.size, 2)) - 1)
            g = side[todo, None] * (2 * rng.random((todo.size, 2)) - 1)
            ok = np.linalg.norm(s - g, axis=1) > side[todo] * 0.5
            for j in range(i):
                lim = radius[todo, j] + radius[todo, i] + GETTING_CLOSE_RANGE
                ok &= np.linalg.norm(s - start[todo, j], axis=1) >= lim
                ok &= np.linalg.norm(g - goal[todo, j], axis=1) >= lim
            start[todo[ok], i] = s[ok]
            goal[todo[ok], i] = g[ok]
            todo = todo[~ok]
    heading = rng.uniform(-np.pi, np.pi, (W, A))
    ids = np.array([_abi.POLICY_IDS[p] for p in policies])
    if len(ids) == 1:
        policy = np.full((W, A), ids[0])
    else:
        policy = rng.choice(ids, size=(W, A), p=policy_distr)
        if policy_to_ensure is not None:  # test_cases.py:286-293
            want = _abi.POLICY_IDS[policy_to_ensure]
            live = np.arange(A)[None, :] < nag[:, None]
            missing = ~np.any((policy == want) & live, axis=1)
            pick = (rng.random(W) * nag).astype(int)
            policy[missing, pick[missing]] = want
    init = make_init(start[..., 0], start[..., 1], goal[..., 0], goal[..., 1], speed, radius, heading, policy)
    return init, nag
