"""The time budget is kept on the device as a step countdown (csrc/ca_kernels.cuh: countdown_steps / budget_after) while
the reference subtracts dt from a float64 every step an agent is still running (GCA/envs/agent.py:232-236).  These tests
check, independently of the C oracle, that the two are the same thing: the float64 value ca_get_state rebuilds equals the
sequence of rounded subtractions EXACTLY, the ran_out_of_time flag rises at exactly the step where that value becomes
<= 0, and a change of dt in the middle of an episode (ca_set_dt recounts the countdowns) keeps both properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(dts, seed):
    from rl_collision_avoidance_b200 import _abi
    from rl_collision_avoidance_b200.vec_env import HostVecEnv, make_init
    W, A = 64, 4
    rng = np.random.default_rng(seed)
    px, py = rng.uniform(-6, 6, (W, A)), rng.uniform(-6, 6, (W, A))
    init = make_init(px, py, px + rng.uniform(3, 8, (W, A)), py - rng.uniform(3, 8, (W, A)), rng.uniform(0.5, 1.5, (W, A)),
                     rng.uniform(0.2, 0.5, (W, A)), rng.uniform(-np.pi, np.pi, (W, A)), np.zeros((W, A)))
    init[..., _abi.I_TIME_REMAINING] = rng.uniform(0.3, 5.0, (W, A))       # short explicit budgets: many time-outs
    nag = np.full(W, A, dtype=np.int32)
    env = HostVecEnv(_abi.default_config(W, A, dt=dts[0]))
    env.set_world_state(init, nag)
    env.reset()
    tr = init[..., _abi.I_TIME_REMAINING].copy()                            # float64, the reference's attribute
    np.testing.assert_array_equal(env.get_state()[..., _abi.S_TIME_REMAINING], tr)
    was_done = np.zeros((W, A), dtype=bool)
    ran_out = np.zeros((W, A), dtype=bool)
    for t, dt in enumerate(dts):
        env.handle.set_dt(dt)                                               # free when unchanged, a recount when not
        act = np.full((W, A), 8 + (t % 3), dtype=np.int32)                  # speed 0: nobody reaches a goal or collides
        _, _, done, _ = env.step(act)
        running = ~was_done
        tr = np.where(running, tr - np.float64(dt), tr)                     # one rounded subtraction per running agent
        ran_out |= running & (tr <= 0)
        st = env.get_state()
        np.testing.assert_array_equal(st[..., _abi.S_TIME_REMAINING], tr, err_msg="step %d (dt %g)" % (t, dt))
        flags = st[..., _abi.S_FLAGS].astype(np.int64)
        np.testing.assert_array_equal((flags & 16) != 0, ran_out, err_msg="ran_out_of_time at step %d" % t)
        was_done = done.astype(bool).copy()
    env.close()
    return int(ran_out.sum())


def test_countdown_is_the_repeated_subtraction():
    assert _run([0.2] * 30, seed=1) > 100


def test_dt_change_in_the_middle_of_an_episode_recounts_the_budgets():
    assert _run([0.2] * 6 + [0.1] * 10 + [0.25] * 8 + [0.1] * 12, seed=2) > 100
