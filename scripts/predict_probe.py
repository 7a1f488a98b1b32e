"""GPU probe of the fused predictor kernel (csrc/ca_predict.cu): error against the fp32 PyTorch network and timing.
Each configuration runs in its own process (a trapped kernel poisons the CUDA context).
    python scripts/predict_probe.py            # all configurations
    python scripts/predict_probe.py one M B    # one configuration in this process
"""
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def one(M, B, trained):
    import numpy as np
    import torch
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    cfg = cfgmod.TrainPhase1() if M == 3 else cfgmod.TrainPhase2()
    cfgmod.set_config(cfg)
    assert cfg.MAX_NUM_OTHER_AGENTS_OBSERVED == M, cfg.MAX_NUM_OTHER_AGENTS_OBSERVED
    net = NetworkVP_rnn("cuda:0", "network", 11, seed=2)
    if trained and M == 3:
        from tests.test_pretrained_policy import load_iros18
        net.net.load_tf_variables(load_iros18())
    L = 6 + 7 * M
    rng = np.random.default_rng(0)
    avg = np.asarray(cfg.NN_INPUT_AVG_VECTOR, dtype=np.float32)
    std = np.asarray(cfg.NN_INPUT_STD_VECTOR, dtype=np.float32)
    obs = np.zeros((B, L), dtype=np.float32)
    obs[:, 1:] = avg + std * rng.normal(size=(B, L - 1)).astype(np.float32)
    obs[:, 0] = 1
    obs[:, 1] = rng.integers(0, M + 1, B)
    t_obs = torch.from_numpy(obs).cuda()
    p_ref, v_ref = net.predict_p_and_v_device(t_obs[:, 1:])
    p, v, a = net.predict_fused(t_obs, want_p=True, want_actions=True, greedy=True)
    torch.cuda.synchronize()
    err = int(net._pred_error.item())
    dp = (p - p_ref).abs().max().item()
    dv = (v - v_ref).abs().max().item()
    agree = (a.long() == p_ref.argmax(1)).float().mean().item()
    print("M=%d B=%d trained=%d: max|dp|=%.3e max|dv|=%.3e (|v|max %.2f) argmax agreement %.4f err=%d" %
          (M, B, trained, dp, dv, v_ref.abs().max().item(), agree, err), flush=True)
    # timing
    for name, fn in (("fused", lambda: net.predict_fused(t_obs, want_p=False, want_actions=True)),
                     ("torch+lstm_step", lambda: net.predict_from_obs(t_obs)),
                     ("torch fp32", lambda: net.predict_p_and_v_device(t_obs[:, 1:]))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("   %-16s %.3f ms per forward (%d rows)" % (name, e0.elapsed_time(e1) / n, B), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        return
    for (M, B, trained) in ((3, 1000, 0), (3, 5003, 1), (3, 262144, 1), (9, 163840, 0)):
        r = subprocess.run([sys.executable, __file__, "one", str(M), str(B), str(trained)], timeout=600,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(r.stdout[-3000:], "rc", r.returncode, flush=True)


if __name__ == "__main__":
    main()
