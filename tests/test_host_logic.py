"""CPU: host-side logic that needs no GPU — Config translation and its loud failures, the host scenario generator's
constraints, the stats line of the Server, the Actions table."""
import numpy as np
import pytest

from rl_collision_avoidance_b200 import _abi
from rl_collision_avoidance_b200 import config as C
from rl_collision_avoidance_b200.scenarios import random_worlds


def test_to_ca_config_maps_reference_fields():
    cfg = C.Config()
    ca = C.to_ca_config(cfg, 77, device=0, auto_reset=1)
    assert (ca.num_worlds, ca.max_agents, ca.max_others_observed, ca.auto_reset) == (77, 4, 3, 1)
    assert ca.dt == 0.2 and ca.near_goal_threshold == 0.2 and ca.getting_close_range == 0.2
    assert ca.min_possible_reward == -0.25 and ca.max_possible_reward == 1.0      # _initialize_rewards clip bounds
    assert ca.sort_method == _abi.SORT_CLOSEST_FIRST and ca.game_over_mode == _abi.OVER_ALL_LEARNING_DONE
    ev = C.EvaluateConfig()
    ca = C.to_ca_config(ev, 1)
    assert ca.max_agents == 19 and ca.dt == 0.1 and ca.max_time_ratio == 8.0 and ca.game_over_mode == _abi.OVER_ALL_DONE
    cfg.TRAIN_SINGLE_AGENT = True
    assert C.to_ca_config(cfg, 1).game_over_mode == _abi.OVER_FIRST_AGENT_DONE


def test_to_ca_config_refuses_what_the_gpu_path_does_not_implement():
    cfg = C.Config()
    cfg.STATES_IN_OBS = ['is_learning', 'num_other_agents', 'dist_to_goal', 'heading_ego_frame', 'pref_speed', 'radius', 'laserscan']
    with pytest.raises(NotImplementedError):
        C.to_ca_config(cfg, 1)
    cfg = C.Config(); cfg.USE_STATIC_MAP = True
    with pytest.raises(NotImplementedError):
        C.to_ca_config(cfg, 1)
    cfg = C.Config(); cfg.AGENT_SORTING_METHOD = "nearest_please"
    with pytest.raises(ValueError):
        C.to_ca_config(cfg, 1)
    cfg = C.Config(); cfg.WIGGLY_BEHAVIOR_THRESHOLD = 0.5; cfg.REWARD_WIGGLY_BEHAVIOR = -0.01
    with pytest.raises(NotImplementedError):
        C.to_ca_config(cfg, 1)


def test_config_class_selection_by_env(monkeypatch, tmp_path):
    monkeypatch.delenv("GYM_CONFIG_PATH", raising=False)
    monkeypatch.setenv("GYM_CONFIG_CLASS", "EvaluateConfig")
    assert type(C.load_config_from_env()).__name__ == "EvaluateConfig"
    f = tmp_path / "my_config.py"
    f.write_text("from rl_collision_avoidance_b200.config import Config\nclass Mine(Config):\n    def __init__(self):\n"
                 "        self.MAX_NUM_AGENTS_IN_ENVIRONMENT = 7\n        Config.__init__(self)\n        self.DT = 0.05\n")
    monkeypatch.setenv("GYM_CONFIG_PATH", str(f))
    monkeypatch.setenv("GYM_CONFIG_CLASS", "Mine")
    c = C.load_config_from_env()
    assert c.MAX_NUM_AGENTS_IN_ENVIRONMENT == 7 and c.MAX_NUM_OTHER_AGENTS_OBSERVED == 6 and c.DT == 0.05
    assert c.STATE_INFO_DICT['other_agents_states']['size'] == (6, 7)


def test_host_scenarios_respect_generator_constraints():
    rng = np.random.default_rng(0)
    W, A = 2000, 4
    nag = rng.integers(2, A + 1, W)
    init, n = random_worlds(W, A, rng, num_agents=nag, policies=['noncoop', 'learning_ga3c', 'static'],
                            policy_distr=[0.05, 0.9, 0.05], policy_to_ensure='learning_ga3c')
    assert init.shape == (W, A, _abi.INIT_STRIDE) and np.array_equal(n, nag)
    live = np.arange(A)[None, :] < nag[:, None]
    ps, rad = init[..., _abi.I_PREF_SPEED], init[..., _abi.I_RADIUS]
    assert np.all((ps[live] >= 0.5) & (ps[live] <= 2.0)) and np.all((rad[live] >= 0.2) & (rad[live] <= 0.8))
    assert np.all(np.any((init[..., _abi.I_POLICY] == _abi.POLICY_LEARNING_GA3C) & live, axis=1))
    assert np.all(np.isnan(init[..., _abi.I_TIME_REMAINING]))
    for i in range(A):
        for j in range(i + 1, A):
            ok = nag > j
            d = np.hypot(init[:, i, 0] - init[:, j, 0], init[:, i, 1] - init[:, j, 1])
            assert np.all(d[ok] >= (rad[:, i] + rad[:, j] + 0.2)[ok])
    frac_learning = np.mean(init[..., _abi.I_POLICY][live] == _abi.POLICY_LEARNING_GA3C)
    assert 0.86 < frac_learning < 0.95


def test_server_stats_line_and_save_trigger():
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    from rl_collision_avoidance_b200.ga3c.Server import Actions, Stats
    cfg = cfgmod.TrainPhase1()
    st = Stats(cfg)
    st.add(100, 12.5, 6000)
    st.add(49950, 30000.0, 3_000_000)
    assert st.episode_count == 50050 and st.total_frame_count == 3_006_000
    assert st.should_save_model == 1                      # crossed SAVE_FREQUENCY = 50000
    line = st.line(4096)
    for token in ("[Time:", "[Episode:    50050", "RScore:", "RPPS:", "PPS:", "TPS:", "NA: 4096"):
        assert token in line, line
    acts = Actions()
    assert acts.num_actions == 11 and acts.actions.shape == (11, 2)
    np.testing.assert_allclose(acts.actions[:, 0], [1, 1, 1, 1, 1, .5, .5, .5, 0, 0, 0])
    np.testing.assert_allclose(acts.actions[[0, 1, 2, 3, 4], 1], [-np.pi / 6, -np.pi / 12, 0, np.pi / 12, np.pi / 6])


def test_bench_workloads_and_reference_arm_line(capsys):
    """bench.py's workload table (BASELINE configs[1] is the default line; configs[2] env half and configs[3] are extra
    lines), the live-agent accounting of the ragged workload, and the JSON contract of the --impl reference arm."""
    import json
    import types
    import bench
    try:
        bench.select_workload("phase1")
        assert (bench.AGENTS, bench.WORLDS_PER_GPU, bench.ALG_BYTES_PER_AGENT_STEP) == (4, 65536, 184)
        assert bench.live_agents(1000) == 4000 and bench.agent_counts(10) is None
        bench.select_workload("phase2")
        assert (bench.AGENTS, bench.WORLDS_PER_GPU, bench.ALG_BYTES_PER_AGENT_STEP) == (10, 16384, 352)
        bench.select_workload("ragged")
        n = bench.agent_counts(18)
        assert n.tolist() == [2, 3, 4, 5, 6, 7, 8, 9, 10] * 2 and bench.live_agents(18) == 108
        assert bench.live_agents(32768) == int((2 + np.arange(32768) % 9).sum())
        # the reference arm on a reduced world count (the oracle port on the host cores); contract keys of the line
        bench.WORLDS_PER_GPU = 64
        bench.run_reference_arm(types.SimpleNamespace(steps=2, warmup=1, gpus=1))
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["unit"] == "agent-steps/s" and line["higher_is_better"] is True
        assert line["e2e"] == {"value": line["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
        assert line["config"]["workload"] == bench.WORKLOAD and line["gpu_launches"] == 0
    finally:
        bench.select_workload("phase1")
