"""GPU: the GA3C actor -> predictor loop.
  * ca_ga3c_record (vectorised ProcessAgent.run_episode bookkeeping) against rows recorded from the UNMODIFIED
    reference ProcessAgent (tests/golden/ga3c_actor.npz) and against the Python oracle on longer random scripts;
  * GpuRollout end to end: every emitted training row is re-derived on the CPU from the observations, actions,
    values, rewards and done flags the rollout produced;
  * Server().main() runs, trains and reports; the gym.Env-shaped façade replays BASELINE config #1.
Tolerance for returns: 1e-5 abs (float32 recursion on the GPU vs float64 in the reference)."""
import numpy as np
import pytest

from oracle import ga3c_oracle
from tests.test_ga3c_oracle import load_ga3c_golden

pytestmark = pytest.mark.gpu


def _rows_close(got, want, tag):
    """Compare two lists of (agent_slot, x, r, a) as multisets."""
    assert len(got) == len(want), "%s: %d rows, expected %d" % (tag, len(got), len(want))
    key = lambda r: (int(r[0]), int(r[3]), np.asarray(r[1], dtype=np.float32).tobytes(), float(r[2]))
    for g, w in zip(sorted(got, key=key), sorted(want, key=key)):
        assert int(g[3]) == int(w[3]), tag
        np.testing.assert_array_equal(np.asarray(g[1], dtype=np.float32), np.asarray(w[1], dtype=np.float32), err_msg=tag)
        assert abs(float(g[2]) - float(w[2])) <= 1e-5, "%s: return %r vs %r" % (tag, g[2], w[2])


def _feed_scripts(rec, scripts, A, L, T):
    """scripts: list (one per world) of lists of episodes {obs [Te+1,A,L], rewards [Te,n], done [Te,n], values, actions};
    plays them back to back (auto-reset semantics) and returns per step the rows the recorder emitted, tagged with
    the agent slot they came from (recovered from x rows)."""
    import torch
    W = len(scripts)
    N = W * A
    cur = [(0, 0)] * W   # (episode index, step inside episode)
    out = []
    obs0 = np.zeros((W, A, L), dtype=np.float32)
    for w in range(W):
        obs0[w] = scripts[w][0]["obs"][0]
    rec.obs_slot(0).copy_(torch.from_numpy(obs0).cuda())
    for t in range(T):
        actions = np.zeros((W, A), dtype=np.int32); values = np.zeros((W, A), dtype=np.float32)
        reward = np.zeros((W, A), dtype=np.float32); done = np.ones((W, A), dtype=np.uint8)
        over = np.zeros(W, dtype=np.uint8); nxt = np.zeros((W, A, L), dtype=np.float32)
        for w in range(W):
            e, k = cur[w]
            if e >= len(scripts[w]):
                continue
            ep = scripts[w][e]
            n = ep["rewards"].shape[1]
            actions[w] = np.maximum(ep["actions"][k], 0); values[w] = ep["values"][k]
            reward[w, :n] = ep["rewards"][k]; done[w, :n] = ep["done"][k]
            last = k + 1 == ep["rewards"].shape[0]
            over[w] = 1 if last else 0
            if last:
                cur[w] = (e + 1, 0)
                if e + 1 < len(scripts[w]):
                    nxt[w] = scripts[w][e + 1]["obs"][0]     # DummyVecEnv: the new episode's first observation
            else:
                cur[w] = (e, k + 1)
                nxt[w] = ep["obs"][k + 1]
        rec.obs_slot(t + 1).copy_(torch.from_numpy(nxt).cuda())
        rec.record(t, torch.from_numpy(actions.reshape(N)).cuda(), torch.from_numpy(values.reshape(N)).cuda(),
                   torch.from_numpy(reward.reshape(N)).cuda(), torch.from_numpy(done.reshape(N)).cuda(),
                   torch.from_numpy(over).cuda())
        x, r, a = rec.take()
        out.append((x.cpu().numpy().copy(), r.cpu().numpy().copy(), a.cpu().numpy().copy()))
    return out


def _expected_rows(scripts, A, time_max, gamma, T):
    """Per global step: list of (slot, x, r, a) from the Python oracle, episodes played back to back."""
    W = len(scripts)
    exp = [[] for _ in range(T)]
    for w in range(W):
        t0 = 0
        for ep in scripts[w]:
            rows = ga3c_oracle.actor_rows(ep["obs"], ep["rewards"], ep["done"], ep["values"], ep["actions"], time_max, gamma)
            for k, rr in enumerate(rows):
                if t0 + k < T:
                    exp[t0 + k].extend((w * A + i, x, r, a) for (i, x, r, a) in rr)
            t0 += len(rows)
    return exp


def _tag_slots(x_rows, r, a, exp_rows):
    """The kernel does not output the agent slot; rows are matched to expected rows by their x vector."""
    lookup = {}
    for slot, x, _, _ in exp_rows:
        lookup.setdefault(np.asarray(x, dtype=np.float32).tobytes(), slot)
    return [(lookup.get(x_rows[k].tobytes(), -1), x_rows[k], r[k], a[k]) for k in range(len(r))]


def test_recorder_matches_reference_process_agent_rows():
    from rl_collision_avoidance_b200.ga3c.rollout import ExperienceRecorder
    meta, eps = load_ga3c_golden()
    A, L = meta["A"], meta["L"]
    scripts = [[ep] for ep in eps]
    T = max(ep["rewards"].shape[0] for ep in eps) + 2
    rec = ExperienceRecorder(len(scripts), A, L, meta["time_max"], meta["gamma"], "cuda:0")
    got = _feed_scripts(rec, scripts, A, L, T)
    total = 0
    for t in range(T):
        want = []
        for w, ep in enumerate(eps):
            sel = np.nonzero(ep["emit_step"] == t)[0]
            # golden rows carry no agent index; recover it through the oracle (already pinned to the golden on CPU)
            want.extend((None, ep["emit_x"][j], ep["emit_r"][j], ep["emit_a"][j]) for j in sel)
        x, r, a = got[t]
        assert len(r) == len(want), "step %d: %d rows, reference yielded %d" % (t, len(r), len(want))
        key = lambda row: (row[1].astype(np.float32).tobytes(), int(row[3]), float(row[2]))
        for g, w_ in zip(sorted(((None, x[k], r[k], a[k]) for k in range(len(r))), key=key), sorted(want, key=key)):
            np.testing.assert_array_equal(g[1], w_[1].astype(np.float32))
            assert int(g[3]) == int(w_[3])
            assert abs(float(g[2]) - float(w_[2])) <= 1e-5
        total += len(want)
    assert total == sum(len(ep["emit_r"]) for ep in eps) and total > 2000


def test_recorder_matches_oracle_on_back_to_back_episodes():
    from rl_collision_avoidance_b200.ga3c.rollout import ExperienceRecorder
    rng = np.random.default_rng(11)
    A, L, TMAX, gamma, W, T = 4, 27, 20, 0.97, 96, 150
    scripts = []
    for w in range(W):
        eps, tot = [], 0
        while tot < T + 5:
            n = int(rng.integers(1, A + 1))
            learning = rng.random(n) < 0.8
            if not learning.any():
                learning[rng.integers(n)] = True
            done_at = rng.integers(0, 60, n)
            Te = int(max(done_at[i] for i in range(n) if learning[i])) + 1
            obs = rng.normal(size=(Te + 1, A, L)).astype(np.float32)
            obs[:, :, 0] = 0; obs[:, :n, 0] = learning; obs[:, n:, :] = 0
            done = np.zeros((Te, n), dtype=bool)
            for i in range(n):
                done[done_at[i]:, i] = True
            eps.append(dict(obs=obs, rewards=np.where(rng.random((Te, n)) < 0.3, rng.uniform(-0.25, 1, (Te, n)), 0.0),
                            done=done, values=rng.normal(size=(Te, A)).astype(np.float32),
                            actions=rng.integers(0, 11, (Te, A)).astype(np.int32)))
            tot += Te
        scripts.append(eps)
    rec = ExperienceRecorder(W, A, L, TMAX, gamma, "cuda:0")
    got = _feed_scripts(rec, scripts, A, L, T)
    exp = _expected_rows(scripts, A, TMAX, gamma, T)
    n_rows = 0
    for t in range(T):
        x, r, a = got[t]
        _rows_close(_tag_slots(x, r, a, exp[t]), exp[t], "step %d" % t)
        n_rows += len(r)
    assert n_rows > 10000
    stats = rec.pop_stats()
    assert stats["episodes"] == sum(1 for w in range(W) for k in range(len(scripts[w]))
                                    if sum(e["rewards"].shape[0] for e in scripts[w][:k + 1]) <= T)


@pytest.fixture
def phase1_cfg():
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    c = cfgmod.TrainPhase1()
    cfgmod.set_config(c)
    yield c
    cfgmod.set_config(None)


def test_rollout_rows_rederived_on_cpu(phase1_cfg):
    """Drive GpuRollout (network + env + recorder on the GPU) and re-derive every emitted row on the CPU from the
    observations / actions / values / rewards / done flags it produced."""
    import torch
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    from rl_collision_avoidance_b200.ga3c.rollout import GpuRollout
    from rl_collision_avoidance_b200.scenarios import random_worlds
    cfg = phase1_cfg
    W, A, T = 64, 4, 90
    rng = np.random.default_rng(5)
    init, nag = random_worlds(W, A, rng, num_agents=rng.integers(2, A + 1, W),
                              policies=['noncoop', 'learning_ga3c', 'static'], policy_distr=[0.05, 0.9, 0.05],
                              policy_to_ensure='learning_ga3c')
    model = NetworkVP_rnn("cuda:0", "network", 11, seed=1)
    ro = GpuRollout(cfg, model, W, init, nag, device=0, seed=123)
    L = ro.L
    obs_hist = [ro.rec.obs_slot(0).cpu().numpy().copy()]
    hist, got = [], []
    for t in range(T):
        reward, done, over = ro.step()
        hist.append((ro.last_actions.cpu().numpy().reshape(W, A).copy(), ro.last_values.cpu().numpy().reshape(W, A).copy(),
                     reward.cpu().numpy().copy(), done.cpu().numpy().copy(), over.cpu().numpy().copy()))
        obs_hist.append(ro.rec.obs_slot(t + 1).cpu().numpy().copy())
        x, r, a = ro.rec.take()
        got.append((x.cpu().numpy().copy(), r.cpu().numpy().copy(), a.cpu().numpy().copy()))
    # observations fed to the network are what the env wrote; policy outputs are a distribution
    p, v = model.predict_p_and_v_device(torch.from_numpy(obs_hist[3].reshape(W * A, L)[:, 1:]).cuda())
    assert torch.allclose(p.sum(dim=1), torch.ones(W * A, device="cuda"), atol=1e-5)
    # cut the history into episodes per world and run the oracle
    scripts = []
    for w in range(W):
        eps, start = [], 0
        for t in range(T):
            if hist[t][4][w]:
                eps.append((start, t + 1)); start = t + 1
        if start < T:
            eps.append((start, T))
        world_eps = []
        for (s, e) in eps:
            obs = np.stack([obs_hist[k][w] for k in range(s, e)] + [obs_hist[e][w]])
            n = A
            world_eps.append(dict(obs=obs, rewards=np.stack([hist[k][2][w] for k in range(s, e)]).astype(np.float64),
                                  done=np.stack([hist[k][3][w] for k in range(s, e)]).astype(bool),
                                  values=np.stack([hist[k][1][w] for k in range(s, e)]),
                                  actions=np.stack([hist[k][0][w] for k in range(s, e)])))
        scripts.append(world_eps)
    exp = _expected_rows(scripts, A, cfg.TIME_MAX, cfg.DISCOUNT, T)
    total = 0
    for t in range(T):
        x, r, a = got[t]
        # the final partial episode of each world has not flushed its tail in `got`; the oracle agrees step by step
        _rows_close(_tag_slots(x, r, a, exp[t]), exp[t], "rollout step %d" % t)
        total += len(r)
    assert total > W * 10
    assert np.isfinite(np.concatenate([g[1] for g in got])).all()
    ro.close()


def test_fused_lstm_predictor_matches_torch_and_numpy(phase1_cfg):
    """predict_from_obs (ca_lstm_step kernel + cuBLAS dense layers) == the plain PyTorch forward == the NumPy oracle."""
    import torch
    from oracle import network_oracle
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    from tests.test_pretrained_policy import load_iros18
    cfg = phase1_cfg
    rng = np.random.default_rng(0)
    net = NetworkVP_rnn("cuda:0", "network", 11, seed=2)
    net.net.load_tf_variables(load_iros18())
    B, L = 5003, 27
    obs = rng.normal(size=(B, L)).astype(np.float32) * 2
    obs[:, 0] = 1
    obs[:, 1] = rng.integers(0, 4, B)
    t_obs = torch.from_numpy(obs).cuda()
    p1, v1 = net.predict_from_obs(t_obs)
    p2, v2 = net.predict_p_and_v_device(t_obs[:, 1:])
    # two float32 evaluation orders of the same function (the framework path sums x Kx + b first, then adds h Kh)
    assert torch.allclose(p1, p2, atol=2e-5) and torch.allclose(v1, v2, atol=2e-4)
    p3, v3 = network_oracle.forward(net.net.tf_variables(), obs[:, 1:], cfg.NN_INPUT_AVG_VECTOR, cfg.NN_INPUT_STD_VECTOR, 3)
    # float32 network with the trained (large) weights vs the float64 oracle: a few 1e-6 of accumulated rounding
    for p_, v_ in ((p1, v1), (p2, v2)):
        np.testing.assert_allclose(p_.cpu().numpy(), p3, rtol=0, atol=3e-5)
        np.testing.assert_allclose(v_.cpu().numpy(), v3, rtol=0, atol=3e-4)
    # a strided view of a wider buffer (the rollout's observation ring) works too
    wide = torch.zeros((B, L + 5), device="cuda"); wide[:, :L] = t_obs
    p4, _ = net.predict_from_obs(wide[:, :L])
    assert torch.equal(p4, p1)


@pytest.mark.parametrize("trained", [False, True])
def test_fused_training_cell_matches_framework_autograd(phase1_cfg, monkeypatch, trained):
    """The trainer's fused LSTM cell (ca_lstm_cell_forward / _backward behind an autograd Function) against the same network
    differentiated op by op by the framework: outputs, losses and every parameter gradient agree to float32 rounding, with
    ragged sequence lengths (including rows without any other agent) and a non-trivial loss."""
    import torch
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    from tests.test_pretrained_policy import load_iros18
    cfg = phase1_cfg
    rng = np.random.default_rng(3)
    B, L1 = 4099, 26
    avg = np.asarray(cfg.NN_INPUT_AVG_VECTOR, dtype=np.float32)
    std = np.asarray(cfg.NN_INPUT_STD_VECTOR, dtype=np.float32)
    x = (avg + std * rng.normal(size=(B, L1))).astype(np.float32)
    x[:, 0] = rng.integers(0, 4, B)
    x = torch.from_numpy(x).cuda()
    y_r = torch.from_numpy(rng.normal(size=B).astype(np.float32)).cuda()
    a = torch.from_numpy(rng.integers(0, 11, B)).cuda()
    results = []
    for fused in ("1", "0"):
        monkeypatch.setenv("GA3C_FUSED_TRAIN_CELL", fused)
        net = NetworkVP_rnn("cuda:0", "network", 11, seed=2)
        if trained:
            net.net.load_tf_variables(load_iros18())
        p, v, _ = net.net(x)
        costs = net.losses(x, y_r, a)
        costs["cost_all"].backward()
        grads = {k: prm.grad.detach().clone() for k, prm in net.net.params.items()}
        results.append((p.detach(), v.detach(), float(costs["cost_all"]), grads))
    (p1, v1, c1, g1), (p0, v0, c0, g0) = results
    assert torch.allclose(p1, p0, atol=2e-6) and torch.allclose(v1, v0, atol=1e-4, rtol=1e-5)
    assert abs(c1 - c0) <= 1e-5 * abs(c0) + 1e-3
    for k in g0:
        scale = float(g0[k].abs().max()) + 1e-12
        err = float((g1[k] - g0[k]).abs().max())
        assert err <= 2e-4 * scale, "gradient of %s differs by %.3e (scale %.3e)" % (k, err, scale)
        assert float(g0[k].abs().max()) > 0


def test_rollout_refreshes_scenarios_on_its_side_stream(phase1_cfg):
    """GpuRollout.attach_scenario_generator: every new episode of a world starts from a scenario it has not started from
    before (the refill runs on a side stream between two env steps and must be complete when the next step adopts it)."""
    import torch
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    from rl_collision_avoidance_b200.ga3c.rollout import GpuRollout
    from rl_collision_avoidance_b200.scenarios import random_worlds
    cfg = phase1_cfg
    W, A = 1024, cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
    init, nag = random_worlds(W, A, np.random.default_rng(0))
    ro = GpuRollout(cfg, NetworkVP_rnn("cuda:0", "network", 11, seed=0), W, init, nag, device=0, seed=1)
    sc = ro.env.scenario_config(cfg.TEST_CASE_ARGS)
    ro.env.generate_scenarios(sc, 11, only_consumed=False)
    ro.env.reset(out_obs=ro.rec.obs_slot(ro.t))
    ro.attach_scenario_generator(sc, 11)
    last_start = ro.env.get_state()[:, :, :2].copy()
    same = changed = 0
    for t in range(120):
        _, _, over = ro.step()
        ro.rec.discard()
        o = over.cpu().numpy().astype(bool)
        if o.any():
            st = ro.env.get_state()[:, :, :2]
            for w in np.nonzero(o)[0]:
                if np.array_equal(st[w], last_start[w]):
                    same += 1
                else:
                    changed += 1
                last_start[w] = st[w]
    assert changed > W // 2 and same == 0, (same, changed)
    ro.close()


def test_server_main_trains(phase1_cfg, tmp_path, monkeypatch):
    monkeypatch.setenv("GA3C_CHECKPOINT_DIR", str(tmp_path))
    from rl_collision_avoidance_b200.ga3c.Server import Server
    phase1_cfg.GPU_TRAIN_BATCH = 512
    srv = Server(cfg=phase1_cfg, num_worlds=256, seed=3)
    before = {k: v.copy() for k, v in srv.model.net.tf_variables().items()}
    out = srv.main(max_steps=80, quiet=True)
    assert out["steps"] == 80 and out["training_steps"] >= 3 and out["episodes"] > 0
    after = srv.model.net.tf_variables()
    assert any(np.abs(after[k] - before[k]).max() > 0 for k in before)
    assert all(np.isfinite(v).all() for v in after.values())
    assert np.isfinite(float(srv.model.last_costs["cost_all"].detach()))
    line = srv.stats.line(256)
    assert "[Episode:" in line and "PPS:" in line and "RScore:" in line
    srv.save_model()
    assert any(f.startswith("network_") for f in __import__("os").listdir(tmp_path))


def test_env_facade_replays_config1_golden():
    """The gym.Env-shaped façade (reference API: Agent objects, dict actions, dict observations) on BASELINE config #1."""
    from rl_collision_avoidance_b200 import env as E
    from rl_collision_avoidance_b200 import config as C
    from tests.golden_util import Golden
    E.set_config(C.Config())
    gold = Golden("phase1")
    name = "config1_modeB"
    g = gold.get(name, "init")
    agents = [E.Agent(g[i, 0], g[i, 1], g[i, 2], g[i, 3], g[i, 5], g[i, 4], np.float64(g[i, 6]), E.LearningPolicyGA3C,
                      E.UnicycleDynamics, [E.OtherAgentsStatesSensor], i) for i in range(2)]
    np.testing.assert_array_equal([a.time_remaining_to_reach_goal for a in agents], gold.get(name, "t_rem0"))
    env = E.CollisionAvoidanceEnv()
    env.set_agents(agents)
    obs = env.reset()
    assert set(obs[0].keys()) == set(E.get_config().STATES_IN_OBS)
    np.testing.assert_allclose(obs[0]['other_agents_states'], gold.get(name, "obs0")[0, 6:].reshape(3, 7), atol=1e-5)
    acts = gold.get(name, "actions")
    for t in range(int(gold.get(name, "steps"))):
        obs, rewards, game_over, info = env.step({0: int(acts[t, 0]), 1: int(acts[t, 1])})
        np.testing.assert_allclose(rewards, gold.get(name, "reward")[t], atol=1e-5)
        assert game_over == bool(gold.get(name, "game_over")[t])
        fl = gold.get(name, "flags")[t]
        for i, a in enumerate(env.agents):
            assert bool(info['which_agents_done'][i]) == bool(fl[i, 5]) and info['which_agents_learning'][i] is True
            np.testing.assert_allclose(a.pos_global_frame, gold.get(name, "pos")[t, i], atol=1e-9)
            assert (a.is_at_goal, a.in_collision, a.ran_out_of_time) == (bool(fl[i, 0]), bool(fl[i, 2]), bool(fl[i, 4]))
            np.testing.assert_allclose(obs[i]['dist_to_goal'], gold.get(name, "obs")[t, i, 2], atol=1e-5)
    with pytest.raises(KeyError):
        env2 = E.CollisionAvoidanceEnv()
        env2.set_agents([E.Agent(0, 0, 3, 0, 0.5, 1.0, 0.0, E.LearningPolicyGA3C, id=0),
                         E.Agent(0, 3, 3, 3, 0.5, 1.0, 0.0, E.LearningPolicyGA3C, id=1)])
        env2.reset()
        env2.step({0: 2})        # missing action for external agent 1 -> KeyError like the reference
    env.close()
    E.set_config(None)


def test_graphed_train_step_matches_the_eager_one(phase1_cfg):
    """NetworkVP_rnn.train on exactly `graph_rows` rows is replayed from CUDA graphs (zero / forward / losses / backward /
    TF-Adam with learning rate, entropy weight and step count on the device): the weights after several steps with
    changing data, learning rate and beta are the eager trainer's."""
    import torch
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    cfg = phase1_cfg
    rng = np.random.default_rng(11)
    B, L1 = 1024, cfg.NN_INPUT_SIZE
    avg = np.asarray(cfg.NN_INPUT_AVG_VECTOR, dtype=np.float32)
    std = np.asarray(cfg.NN_INPUT_STD_VECTOR, dtype=np.float32)
    eager = NetworkVP_rnn("cuda:0", "network", 11, seed=6)
    graphed = NetworkVP_rnn("cuda:0", "network", 11, seed=6)
    graphed.enable_graphed_training(B)
    start = {k: v.copy() for k, v in eager.net.tf_variables().items()}
    for step in range(5):
        x = (avg + std * rng.normal(size=(B, L1))).astype(np.float32)
        x[:, 0] = rng.integers(0, 4, B)
        x = torch.from_numpy(x).cuda()
        y_r = torch.from_numpy(rng.normal(size=B).astype(np.float32)).cuda()
        a = torch.from_numpy(rng.integers(0, 11, B).astype(np.int32)).cuda()
        for net in (eager, graphed):
            net.learning_rate = 1e-3 * (1 + step)          # annealing needs no re-capture
            net.beta = 1e-4 * (1 + step)
        c_e = eager.train(x, y_r, a)
        c_g = graphed.train(x, y_r, a)
        assert abs(float(c_e["cost_all"]) - float(c_g["cost_all"])) <= 1e-4 * abs(float(c_e["cost_all"])) + 1e-3
    assert graphed._graphed and graphed.global_step == eager.global_step == 5 and graphed.opt.t == 5
    ve, vg = eager.net.tf_variables(), graphed.net.tf_variables()
    for k in ve:
        moved = np.abs(ve[k] - start[k]).max()
        assert moved > 1e-4, k
        np.testing.assert_allclose(vg[k], ve[k], rtol=0, atol=2e-2 * moved + 1e-7, err_msg=k)
    # a batch of another size falls back to the eager path and keeps the optimiser state consistent
    x2 = x[:100]
    graphed.train(x2, y_r[:100], a[:100]); eager.train(x2, y_r[:100], a[:100])
    assert graphed.opt.t == eager.opt.t == 6
    np.testing.assert_allclose(graphed.net.tf_variables()["layer2/kernel"], eager.net.tf_variables()["layer2/kernel"], rtol=0, atol=1e-4)
