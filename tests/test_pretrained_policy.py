"""The reference's trained GA3C-CADRL weights (IROS18 checkpoint, imported with the product's own TF-checkpoint reader,
fixture tests/golden/iros18_weights.npz) must actually solve the task in this environment: that ties together the
observation layout and normalisation, the closest_last neighbour order the reference uses for trained policies
(GCA/experiments/src/env_utils.py:106-109), the TF-1.15 LSTMCell semantics of the PyTorch network, the action table and
the dynamics.  With a wrong gate order, obs column or sort order the success rate collapses (random init: 0 %)."""
import os
import struct

import numpy as np
import pytest

from rl_collision_avoidance_b200 import _abi
from rl_collision_avoidance_b200.ga3c import Config as cfgmod
from rl_collision_avoidance_b200.ga3c import tf_checkpoint
from rl_collision_avoidance_b200.scenarios import random_worlds

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iros18_weights.npz")
REF_PREFIX = "/root/reference/gym-collision-avoidance/gym_collision_avoidance/envs/policies/GA3C_CADRL/checkpoints/IROS18/network_01900000"


def load_iros18():
    z = np.load(GOLD)
    return {k.replace("__", "/"): z[k] for k in z.files}


@pytest.fixture
def phase1():
    c = cfgmod.TrainPhase1()
    cfgmod.set_config(c)
    yield c
    cfgmod.set_config(None)


def _write_bundle(prefix, tensors):
    """Minimal writer of the TF bundle format (one uncompressed data block), to test the reader without TensorFlow."""
    def varint(n):
        out = b""
        while True:
            b = n & 0x7F
            n >>= 7
            out += bytes([b | (0x80 if n else 0)])
            if not n:
                return out

    def block(entries):
        body = b""
        for k, v in entries:                       # no prefix sharing, one restart point
            body += varint(0) + varint(len(k)) + varint(len(v)) + k + v
        return body + struct.pack("<I", 0) + struct.pack("<I", 1)

    data, entries = b"", [(b"", b"\x08\x01")]
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name])
        shape = b"".join(b"\x12" + varint(len(d)) + d for d in (b"\x08" + varint(s) for s in a.shape))
        dtype = {np.dtype(np.float32): 1, np.dtype(np.int32): 3}[a.dtype]
        msg = b"\x08" + varint(dtype) + b"\x12" + varint(len(shape)) + shape + b"\x20" + varint(len(data)) + \
            b"\x28" + varint(a.nbytes) + b"\x35" + struct.pack("<I", 0)
        entries.append((name.encode(), msg))
        data += a.tobytes()
    blk = block(entries)
    idx = block([(b"\xff", varint(0) + varint(len(blk)))])
    meta = block([])
    f = blk + b"\x00" + b"\x00" * 4
    meta_off = len(f)
    f += meta + b"\x00" + b"\x00" * 4
    idx_off = len(f)
    f += idx + b"\x00" + b"\x00" * 4
    footer = varint(meta_off) + varint(len(meta)) + varint(idx_off) + varint(len(idx))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    open(prefix + ".index", "wb").write(f + footer)
    open(prefix + ".data-00000-of-00001", "wb").write(data)


def test_tf_checkpoint_reader_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"rnn/lstm_cell/kernel:0": rng.normal(size=(71, 256)).astype(np.float32),
               "layer1/bias:0": rng.normal(size=(256,)).astype(np.float32), "step:0": np.array(1234567, dtype=np.int32),
               "logits_v/kernel:0": rng.normal(size=(256, 1)).astype(np.float32)}
    prefix = str(tmp_path / "network_00000042")
    _write_bundle(prefix, tensors)
    got = tf_checkpoint.load_checkpoint(prefix)
    assert set(got) == set(tensors)
    for k in tensors:
        np.testing.assert_array_equal(got[k], tensors[k])
    with pytest.raises(ValueError):
        open(prefix + ".index", "ab").write(b"garbage")
        tf_checkpoint.read_index(prefix + ".index")


@pytest.mark.skipif(not os.path.exists(REF_PREFIX + ".index"), reason="reference checkpoint only exists in the build container")
def test_reader_on_the_reference_checkpoint_matches_fixture():
    v = tf_checkpoint.network_variables(REF_PREFIX)
    gold = load_iros18()
    assert set(v) == set(gold) == set(tf_checkpoint.NETWORK_VARIABLES)
    for k in v:
        np.testing.assert_array_equal(v[k], gold[k])
    assert v["rnn/lstm_cell/kernel"].shape == (71, 256) and v["layer1/kernel"].shape == (68, 256)
    assert int(tf_checkpoint.load_checkpoint(REF_PREFIX)["step:0"]) == 4688585


def _evaluate(env_factory, predict, W, A, sort, seed):
    rng = np.random.default_rng(seed)
    init, nag = random_worlds(W, A, rng, num_agents=rng.integers(2, A + 1, W))
    env = env_factory(_abi.default_config(W, A, sort_method=_abi.SORT_METHODS[sort]))
    env.set_world_state(init, nag)
    env.reset()
    for t in range(220):
        x = np.asarray(env.obs, dtype=np.float32).reshape(W * A, -1)[:, 1:]
        act = predict(x).argmax(axis=1).reshape(W, A).astype(np.int32)      # GA3CCADRLPolicy.find_next_action: argmax
        env.step(act)
        if np.all(env.game_over):
            break
    fl = env.get_state()[..., _abi.S_FLAGS].astype(int)
    live = np.arange(A)[None, :] < nag[:, None]
    env.close()
    return (((fl & _abi.F_AT_GOAL) != 0)[live].mean(), ((fl & _abi.F_IN_COLLISION) != 0)[live].mean(),
            ((fl & _abi.F_RAN_OUT_OF_TIME) != 0)[live].mean())


def test_trained_policy_solves_the_task_in_the_oracle_env(phase1):
    from oracle.ca_oracle import OracleEnv
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    net = NetworkVP_rnn("cpu", "network", 11, seed=0)
    net.net.load_tf_variables(load_iros18())
    goal, coll, timeout = _evaluate(OracleEnv, net.predict_p, 150, 4, "closest_last", seed=3)
    assert goal > 0.95 and coll < 0.02, (goal, coll, timeout)
    # the same weights with the neighbours in the wrong order are clearly worse: the order convention is load-bearing
    goal_wrong, coll_wrong, _ = _evaluate(OracleEnv, net.predict_p, 150, 4, "closest_first", seed=3)
    assert goal_wrong < goal - 0.1 and coll_wrong > coll + 0.05
    # and an untrained network never gets there
    rnd = NetworkVP_rnn("cpu", "network", 11, seed=0)
    assert _evaluate(OracleEnv, rnd.predict_p, 60, 4, "closest_last", seed=3)[0] < 0.1


@pytest.mark.gpu
def test_trained_policy_solves_the_task_on_the_gpu(phase1):
    from rl_collision_avoidance_b200.ga3c.NetworkVP_rnn import NetworkVP_rnn
    from rl_collision_avoidance_b200.vec_env import HostVecEnv
    net = NetworkVP_rnn("cuda:0", "network", 11, seed=0)
    net.net.load_tf_variables(load_iros18())
    goal, coll, timeout = _evaluate(HostVecEnv, net.predict_p, 4096, 4, "closest_last", seed=5)
    assert goal > 0.96 and coll < 0.02, (goal, coll, timeout)
