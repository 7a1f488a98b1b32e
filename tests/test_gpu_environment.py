"""The GA3C `Environment(id)` adapter (GA3C/Environment.py:37-116) driven the way the reference's ProcessAgent drives it
(GA3C/ProcessAgent.py:107,128-149), per-call dt (CollisionAvoidanceEnv.step(actions, dt), collision_avoidance_env.py:131-138)
and the get_testcase_two_agents generator (test_cases.py:77-84) on the CUDA env."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def phase1_cfg():
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    cfg = cfgmod.TrainPhase1()
    cfgmod.set_config(cfg)
    yield cfg
    cfgmod.set_config(None)


def test_environment_adapter_runs_an_episode_like_process_agent(phase1_cfg):
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat"))
    try:
        from GA3C.Environment import Environment          # the drop-in namespace a reference ProcessAgent imports
    finally:
        sys.path.pop(0)
    cfg = phase1_cfg
    A, L = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT, cfg.FULL_STATE_LENGTH if hasattr(cfg, "FULL_STATE_LENGTH") else None
    env = Environment(3)
    assert env.previous_state is None and env.current_state is None
    env.reset()
    obs = env.latest_observations
    assert obs.shape[0] == A and obs.dtype == np.float32
    L = obs.shape[1]
    assert env.current_state.shape == (1, A, L - 1) and env.previous_state is None
    np.testing.assert_array_equal(env.current_state[0], obs[:, 1:])
    rng = np.random.default_rng(0)
    game_over, steps, total = False, 0, 0.0
    while not game_over and steps < 400:
        actions = {}
        for i, row in enumerate(env.latest_observations):     # ProcessAgent.run_episode :128-142
            if row[0]:
                actions[i] = int(rng.integers(0, 11))
        assert actions, "at least one learning agent (policy_to_ensure)"
        before = env.current_state
        rewards, game_over, infos = env.step([actions], 0, steps)
        rewards = rewards[0]
        done = infos[0]['which_agents_done']
        learning = infos[0]['which_agents_learning']
        assert set(done) == set(learning) and len(done) == len(rewards) <= A
        assert env.previous_state is before and env.current_state.shape == (1, A, L - 1)
        total += float(np.sum(rewards))
        game_over = bool(game_over[0]) if hasattr(game_over, "__len__") else bool(game_over)
        steps += 1
    assert game_over and steps > 1
    assert abs(env.total_reward - total) < 1e-6
    env.game.envs[0].close()


def test_step_with_a_per_call_dt_matches_an_env_configured_with_that_dt():
    """ca_set_dt: stepping a DT = 0.2 env with dt = 0.1 per call is the oracle run at DT = 0.1 (time budgets are computed
    at reset from Config.DT only through its floor, which the 2 m+ goals of these worlds never reach)."""
    from oracle.ca_oracle import OracleEnv
    from rl_collision_avoidance_b200 import _abi
    from rl_collision_avoidance_b200.vec_env import HostVecEnv, make_init
    W, A = 96, 4
    rng = np.random.default_rng(5)
    px, py = rng.uniform(-4, 4, (W, A)), rng.uniform(-4, 4, (W, A))
    init = make_init(px, py, px + rng.uniform(2, 5, (W, A)), py - rng.uniform(2, 5, (W, A)), rng.uniform(0.5, 2.0, (W, A)),
                     rng.uniform(0.2, 0.8, (W, A)), rng.uniform(-np.pi, np.pi, (W, A)), rng.choice([0, 0, 1, 2], size=(W, A)))
    nag = rng.integers(2, A + 1, W).astype(np.int32)
    gpu = HostVecEnv(_abi.default_config(W, A, dt=0.2), want_sorted_idx=True)
    cpu = OracleEnv(_abi.default_config(W, A, dt=0.1))
    gpu.set_world_state(init, nag); cpu.set_world_state(init, nag)
    gpu.reset(); cpu.reset()
    gpu.handle.set_dt(0.1)
    for t in range(30):
        act = rng.integers(0, 11, (W, A)).astype(np.int32)
        gpu.step(act); cpu.step(act)
        np.testing.assert_array_equal(gpu.done, cpu.done)
        np.testing.assert_array_equal(gpu.sorted_idx, cpu.sorted_idx)
        np.testing.assert_allclose(gpu.reward, cpu.reward, rtol=0, atol=1e-5)
    np.testing.assert_allclose(gpu.get_state()[..., :6], cpu.get_state()[..., :6], rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        gpu.handle.set_dt(0.0)
    gpu.close(); cpu.close()


def test_gym_facade_two_agent_testcase_and_per_call_dt(phase1_cfg):
    from rl_collision_avoidance_b200 import env as envmod
    envmod.set_config(phase1_cfg)
    e = envmod.CollisionAvoidanceEnv()
    e.set_testcase("get_testcase_two_agents", {"policies": ["learning_ga3c", "learning_ga3c"]})
    obs = e.reset()
    assert len(e.agents) == 2
    np.testing.assert_allclose(e.agents[0].pos_global_frame, [-3, -3])
    np.testing.assert_allclose(e.agents[1].goal_global_frame, [-3, -3])
    assert float(obs[0]['dist_to_goal']) == pytest.approx(np.hypot(6, 6), abs=1e-5)
    x0 = e.agents[0].pos_global_frame.copy()
    e.step({0: 2, 1: 2}, dt=0.05)                     # straight ahead at pref_speed 1.0 for 0.05 s
    np.testing.assert_allclose(e.agents[0].pos_global_frame - x0, [0.05, 0.0], atol=1e-12)
    e.step({0: 2, 1: 2})                              # back to Config.DT
    np.testing.assert_allclose(e.agents[0].pos_global_frame - x0, [0.25, 0.0], atol=1e-12)
    with pytest.raises(NotImplementedError):
        e.set_testcase("full_test_suite", {})
    e.close()
    envmod.set_config(None)
