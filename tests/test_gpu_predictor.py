"""GPU: the fused tcgen05 predictor kernel (csrc/ca_predict.cu: ca_predictor_pack + ca_predict) against the plain
PyTorch fp32 network (`PolicyValueNet`, itself pinned on the NumPy/TF1 oracle in tests/test_network.py and
test_gpu_ga3c.py::test_fused_lstm_predictor_matches_torch_and_numpy).

The kernel multiplies fp16 operands (11-bit significand, like TF32) with fp32 accumulation and evaluates the LSTM
gates with `tanh.approx.f32` (2^-11 relative error); the VALUE head is float32 (fullyconnected1 outputs before their fp16
rounding times the float32 logits_v kernel, on the CUDA cores).  It is compared with a tolerance, stated per test:
  * observations produced by the env itself (what the rollout feeds it): |dp| <= 5e-3, |dv| <= 3e-3 * (1 + |v|)
    (measured with the trained IROS18 weights: dp max 3.8e-3 / p99 1.3e-3, dv max 0.6-2.2e-3 / p99 2.2e-4);
  * synthetic rows drawn from the normalisation statistics, random-init and trained IROS18 weights: |dp| <= 2e-2,
    |dv| <= 5e-3 * (1 + |v|) (measured: dp 1.4e-2, dv 2.7e-3 / p99 1e-3 trained; 4e-5 and 3e-4 random-init; with the
    round-1 fp16 value head dv was 1.4-2e-2; trained weights amplify: |v| reaches ~20 on such out-of-distribution rows).
Index outputs: the greedy action equals argmax of the kernel's own p bit-exactly, and argmax of the fp32 p wherever the
fp32 top-2 gap exceeds the p tolerance; sampled actions follow p (chi-square style bound) and are reproducible per
(seed, offset)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(phase):
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    c = cfgmod.TrainPhase1() if phase == 1 else cfgmod.TrainPhase2()
    cfgmod.set_config(c)
    return c


@pytest.fixture
def phase_cfg():
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    yield _cfg
    cfgmod.set_config(None)


def _net(trained, seed=2):
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    net = NetworkVP_rnn("cuda:0", "network", 11, seed=seed)
    if trained:
        from tests.test_pretrained_policy import load_iros18
        net.net.load_tf_variables(load_iros18())
        net.mark_weights_changed()
    return net


def _synthetic_obs(cfg, B, M, rng):
    L = 6 + 7 * M
    avg = np.asarray(cfg.NN_INPUT_AVG_VECTOR, dtype=np.float32)
    std = np.asarray(cfg.NN_INPUT_STD_VECTOR, dtype=np.float32)
    obs = np.zeros((B, L), dtype=np.float32)
    obs[:, 1:] = avg + std * rng.normal(size=(B, L - 1)).astype(np.float32)
    obs[:, 0] = 1
    obs[:, 1] = rng.integers(0, M + 1, B)      # ragged sequence lengths, including rows with no other agent
    return obs


def _check(net, t_obs, p_tol, v_tol):
    import torch
    p_ref, v_ref = net.predict_p_and_v_device(t_obs[:, 1:])
    p, v, a = net.predict_fused(t_obs, want_p=True, want_actions=True, greedy=True)
    torch.cuda.synchronize()
    assert int(net._pred_error.item()) == 0
    dp = (p - p_ref).abs().max().item()
    dv = ((v - v_ref).abs() / (1 + v_ref.abs())).max().item()
    assert dp <= p_tol, "policy differs by %.3e" % dp
    assert dv <= v_tol, "value differs by %.3e (relative to 1 + |v|)" % dv
    assert torch.allclose(p.sum(1), torch.ones_like(p[:, 0]), atol=1e-5)
    assert torch.equal(a.long(), p.argmax(1))            # index output: exact w.r.t. the kernel's own p
    top2 = p_ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2 * p_tol
    assert torch.equal(a.long()[clear], p_ref.argmax(1)[clear])
    return dp, dv


@pytest.mark.parametrize("phase,trained,B", [(1, False, 1000), (1, True, 5003), (2, False, 4099), (1, True, 128), (1, False, 1)])
def test_fused_predictor_matches_fp32_network_on_synthetic_rows(phase_cfg, phase, trained, B):
    import torch
    cfg = phase_cfg(phase)
    M = cfg.MAX_NUM_OTHER_AGENTS_OBSERVED
    net = _net(trained)
    obs = _synthetic_obs(cfg, B, M, np.random.default_rng(B))
    _check(net, torch.from_numpy(obs).cuda(), 2e-2, 5e-3)


@pytest.mark.parametrize("phase", [1, 2])
def test_fused_predictor_on_env_observations(phase_cfg, phase):
    """Rows the env produces (ragged agent counts, padded absent agents, done agents) after a few steps."""
    import torch
    from rl_collision_avoidance_b200.config import to_ca_config
    from rl_collision_avoidance_b200.scenarios import random_worlds
    from rl_collision_avoidance_b200.vec_env import VecCollisionAvoidanceEnv
    cfg = phase_cfg(phase)
    W = 2048
    env = VecCollisionAvoidanceEnv(to_ca_config(cfg, W, device=0, auto_reset=1))
    rng = np.random.default_rng(phase)
    init, nag = random_worlds(W, env.A, rng, num_agents=rng.integers(2, env.A + 1, W))
    env.set_world_state(init, nag)
    obs = env.reset()
    net = _net(trained=(phase == 1))
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(12):
        flat = obs.reshape(W * env.A, env.L)
        if t % 4 == 0:
            _check(net, flat, 5e-3, 3e-3)
        act = torch.randint(0, 11, (W, env.A), generator=gen, device="cuda", dtype=torch.int32)
        obs, _, _, _ = env.step(act)
    env.close()


@pytest.mark.parametrize("phase", [1, 2])
def test_learning_only_plan_predicts_exactly_the_learning_rows(phase_cfg, phase):
    """ca_predict_plan + ca_predict_rows: the plan lists the is_learning rows sorted by descending sequence length; the
    planned rows get bit-identical outputs to the all-rows launch (a row's result does not depend on its tile-mates), the
    others get v = 0 / action 0 / p = 0."""
    import torch
    cfg = phase_cfg(phase)
    M = cfg.MAX_NUM_OTHER_AGENTS_OBSERVED
    net = _net(False)
    rng = np.random.default_rng(5 + phase)
    B = 20011
    obs = _synthetic_obs(cfg, B, M, rng)
    learning = rng.random(B) < 0.6
    obs[:, 0] = learning
    obs[~learning & (rng.random(B) < 0.5)] = 0.0            # absent slots: all-zero rows
    t_obs = torch.from_numpy(obs).cuda()
    p0, v0, a0 = net.predict_fused(t_obs, want_p=True, want_actions=True, seed=3)
    net._pred_calls -= 1                                     # same (seed, offset) -> same uniform draws per row
    p1, v1, a1 = net.predict_fused(t_obs, want_p=True, want_actions=True, seed=3, learning_only=True)
    torch.cuda.synchronize()
    assert int(net._pred_error.item()) == 0
    lm = torch.from_numpy(learning).cuda()
    assert torch.equal(p1[lm], p0[lm]) and torch.equal(v1[lm], v0[lm]) and torch.equal(a1[lm], a0[lm])
    assert float(p1[~lm].abs().max()) == 0.0 and float(v1[~lm].abs().max()) == 0.0 and int(a1[~lm].abs().max()) == 0
    counters = net._plan_counters.cpu().numpy()
    rows = net._plan_rows.cpu().numpy()[:counters[0]]
    assert counters[0] == learning.sum()
    assert np.array_equal(np.sort(rows), np.nonzero(learning)[0])
    seq = np.ceil(np.clip(obs[rows, 1], 0, M)).astype(int)
    assert np.all(np.diff(seq) <= 0), "plan must be sorted by descending sequence length"
    assert np.array_equal(np.bincount(seq, minlength=M + 1), counters[1:M + 2])


def test_fused_predictor_repacks_after_training_step(phase_cfg):
    import torch
    cfg = phase_cfg(1)
    net = _net(False)
    rng = np.random.default_rng(7)
    obs = torch.from_numpy(_synthetic_obs(cfg, 512, 3, rng)).cuda()
    _check(net, obs, 2e-2, 2e-2)
    for _ in range(3):   # weights move; predict_fused must see the new ones without being told
        net.learning_rate = 1e-2
        net.train(obs[:, 1:], torch.ones(512, device="cuda"), torch.zeros(512, dtype=torch.int64, device="cuda"))
    p_old = net.predict_fused(obs)[0].clone()
    _check(net, obs, 2e-2, 2e-2)
    with torch.no_grad():
        net.net.w("logits_p/bias").add_(torch.arange(11, device="cuda", dtype=torch.float32))
    net.mark_weights_changed()
    assert not torch.allclose(net.predict_fused(obs)[0], p_old, atol=1e-3)
    _check(net, obs, 2e-2, 2e-2)


def test_sampled_actions_follow_the_policy_and_are_reproducible(phase_cfg):
    import torch
    cfg = phase_cfg(1)
    net = _net(True)
    rng = np.random.default_rng(11)
    row = _synthetic_obs(cfg, 4, 3, rng)
    B = 200000
    obs = torch.from_numpy(np.repeat(row, B // 4, axis=0)).cuda()
    p, _, a1 = net.predict_fused(obs, want_actions=True, seed=5)
    calls = net._pred_calls
    net._pred_calls = calls - 1
    _, _, a2 = net.predict_fused(obs, want_actions=True, seed=5)     # same (seed, offset) -> same draws
    assert torch.equal(a1, a2)
    _, _, a3 = net.predict_fused(obs, want_actions=True, seed=5)     # next offset -> different draws
    assert not torch.equal(a1, a3)
    assert int(a1.min()) >= 0 and int(a1.max()) <= 10
    for g in range(4):
        sl = slice(g * (B // 4), (g + 1) * (B // 4))
        freq = torch.bincount(a1[sl].long(), minlength=11).double() / (B // 4)
        pg = p[sl][0].double()
        # binomial standard deviation per action is <= 0.5 / sqrt(50000) = 2.2e-3; 5 sigma
        assert (freq - pg).abs().max().item() < 1.2e-2, (freq, pg)


def test_predict_rejects_bad_arguments(phase_cfg):
    import ctypes as C
    import torch
    from rl_collision_avoidance_b200 import _abi
    from rl_collision_avoidance_b200._lib import lib
    phase_cfg(1)
    net = _net(False)
    obs = torch.zeros((8, 27), device="cuda")
    net.predict_fused(obs)
    L = lib()
    blob = C.c_void_p(net._blob.data_ptr())
    v = torch.empty(8, device="cuda")
    args = lambda stride, batch, M: (C.c_void_p(obs.data_ptr()), stride, batch, M, blob, None, C.c_void_p(v.data_ptr()), None, 0,
                                     0.0, 0, 0, None, 0, None)
    assert L.ca_predict(*args(27, 8, 3)) == _abi.CA_OK
    assert L.ca_predict(*args(20, 8, 3)) == _abi.CA_ERR_INVALID_ARG      # row shorter than 6 + 7 M
    assert L.ca_predict(*args(27, 0, 3)) == _abi.CA_ERR_INVALID_ARG
    assert L.ca_predict(*args(400, 8, 23)) == _abi.CA_ERR_UNSUPPORTED
    assert L.ca_predictor_pack(None, blob, 0, None) == _abi.CA_ERR_INVALID_ARG
    torch.cuda.synchronize()
