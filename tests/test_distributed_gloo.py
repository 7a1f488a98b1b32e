"""CPU, world_size 2, gloo: the N>1 host logic — world sharding, rollout-row gather + weight broadcast, and the
equivalence of summed per-rank gradients with one trainer seeing all rows (the A3C loss is a sum over rows)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, tmpdir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from rl_collision_avoidance_b200.ga3c import Config as cfgmod
    from rl_collision_avoidance_b200.ga3c import parallel
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    cfg = cfgmod.TrainPhase1()
    cfgmod.set_config(cfg)
    rng = np.random.default_rng(7)            # same stream on both ranks: the full batch
    B = 300
    x = rng.normal(size=(B, cfg.NN_INPUT_SIZE)).astype(np.float32)
    x[:, 0] = rng.integers(0, 4, B)
    y = rng.normal(size=B).astype(np.float32)
    a = rng.integers(0, 11, B).astype(np.int32)
    lo, hi = (0, 117) if rank == 0 else (117, B)   # ragged split

    # (1) gather_rows reproduces the concatenation in rank order
    gx, gr, ga = parallel.gather_rows(torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]), torch.from_numpy(a[lo:hi]))
    assert gx.shape[0] == B and torch.equal(gx, torch.from_numpy(x)) and torch.equal(ga, torch.from_numpy(a))
    assert torch.equal(gr, torch.from_numpy(y))

    # (2) summed per-rank gradients + one Adam step == single trainer on all rows
    ref = NetworkVP_rnn("cpu", "network", 11, seed=4)
    ref.train(x, y, a, 0)
    net = NetworkVP_rnn("cpu", "network", 11, seed=4 + 10 * rank)     # different init per rank ...
    parallel.broadcast_weights(list(net.net.parameters()), src=0)     # ... until rank 0's weights are broadcast
    if rank == 1:
        pass
    net.backward(torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]), torch.from_numpy(a[lo:hi]))
    parallel.allreduce_flat(net.flat_grad)
    net.apply_gradients()
    ref0 = NetworkVP_rnn("cpu", "network", 11, seed=4)
    ref0.train(x, y, a, 0)
    for k, v in net.net.tf_variables().items():
        np.testing.assert_allclose(v, ref0.net.tf_variables()[k], rtol=0, atol=2e-7, err_msg=k)

    # (2b) a rank with NO rows still takes part: it contributes zeros of the full fixed-length gradient buffer (the
    # distributed trainer agrees on the number of optimiser steps from the largest row count, so this happens)
    lo2, hi2 = (0, B) if rank == 0 else (B, B)
    net.backward(torch.from_numpy(x[lo2:hi2]), torch.from_numpy(y[lo2:hi2]), torch.from_numpy(a[lo2:hi2]))
    parallel.allreduce_flat(net.flat_grad)
    net.apply_gradients()
    ref0.train(x, y, a, 0)
    for k, v in net.net.tf_variables().items():
        np.testing.assert_allclose(v, ref0.net.tf_variables()[k], rtol=0, atol=4e-7, err_msg=k)

    # (3) sharding and max-over-ranks
    assert parallel.shard_range(524288, rank, world_size) == ((0, 262144) if rank == 0 else (262144, 524288))
    assert parallel.max_over_ranks(1.0 + rank, "cpu") == 2.0
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmpdir, "ok%d" % rank), "w").write("ok")


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _extras_worker(rank, world_size, port, tmpdir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import bench

    def measure(ms):
        return lambda: ([ms, 2 * ms], lambda v, ws: {"ms": v[0], "ms1": v[1], "ranks": ws})

    def broken():
        raise RuntimeError("capture invalidated")

    # (1) healthy: the slowest rank's timings, on every rank
    r1 = bench.reduce_extra("a", 2, measure(1.0 + rank), rank, world_size, "cpu", True)
    assert r1 == {"ms": 2.0, "ms1": 4.0, "ranks": 2}
    # (2) rank 1 fails before producing anything: BOTH ranks report the failure and nobody hangs ...
    r2 = bench.reduce_extra("b", 2, broken if rank == 1 else measure(1.0), rank, world_size, "cpu", True)
    assert "error" in r2
    # (3) ... and the sequence of collectives is still aligned: the next extra pairs with its peer's
    r3 = bench.reduce_extra("c", 1, lambda: ([5.0 + rank], lambda v, ws: {"ms": v[0]}), rank, world_size, "cpu", True)
    assert r3 == {"ms": 6.0}
    t = torch.tensor([float(rank)])
    dist.all_reduce(t)
    assert float(t) == 1.0
    dist.destroy_process_group()
    open(os.path.join(tmpdir, "extras_ok%d" % rank), "w").write("ok")


def test_bench_extras_survive_a_failing_rank(tmp_path):
    """bench.py's extra workloads at N > 1: one collective per extra on every rank, whether its measurement worked or not."""
    port = _free_port()
    mp.spawn(_extras_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "extras_ok0") and os.path.exists(tmp_path / "extras_ok1")


def test_shard_range_is_a_partition():
    from rl_collision_avoidance_b200.ga3c.parallel import shard_range
    for W, G in ((65536, 8), (10, 3), (7, 8), (524288, 8), (1, 1)):
        ranges = [shard_range(W, r, G) for r in range(G)]
        assert ranges[0][0] == 0 and ranges[-1][1] == W
        assert all(ranges[k][1] == ranges[k + 1][0] for k in range(G - 1))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
