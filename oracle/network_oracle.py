"""TEST INFRASTRUCTURE — NumPy forward pass of the GA3C-CADRL policy/value network written from the TF-1.15
definitions the reference uses (GA3C/NetworkVP_rnn.py:39-108, GA3C/NetworkVPCore.py:64-77):
tf.contrib.rnn.LSTMCell (gate order i, j, f, o; forget_bias 1.0; one kernel on concat(x, h)),
tf.nn.dynamic_rnn with sequence_length (state frozen past the end of a row's sequence), tf.layers.dense.
Parity is UNPINNED by the reference (TensorFlow is not installable here and the reference has no network
tests); the definition above is the published TF-1.15 algorithm."""
import numpy as np


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def forward(variables, x, avg, std, M, host_len=4, other_len=7, first=1, min_policy=0.0):
    """variables: {tf variable name: array}; x [B, 1+host_len+M*other_len] float -> (p [B, 11], v [B])."""
    x = np.asarray(x, dtype=np.float64)
    xn = (x - avg) / std
    seq_len = x[:, 0]
    host = xn[:, first:first + host_len]
    others = xn[:, first + host_len:].reshape(-1, M, other_len)
    B, H = x.shape[0], 64
    K, b = variables["rnn/lstm_cell/kernel"].astype(np.float64), variables["rnn/lstm_cell/bias"].astype(np.float64)
    h = np.zeros((B, H)); c = np.zeros((B, H))
    for t in range(M):
        z = np.concatenate([others[:, t], h], axis=1) @ K + b
        i, j, f, o = np.split(z, 4, axis=1)
        c_new = _sigmoid(f + 1.0) * c + _sigmoid(i) * np.tanh(j)
        h_new = _sigmoid(o) * np.tanh(c_new)
        live = (seq_len > t)[:, None]
        c = np.where(live, c_new, c)
        h = np.where(live, h_new, h)
    a = np.concatenate([host, h], axis=1)
    for name in ("layer1", "layer2", "fullyconnected1"):
        a = np.maximum(a @ variables[name + "/kernel"].astype(np.float64) + variables[name + "/bias"], 0.0)
    logits = a @ variables["logits_p/kernel"].astype(np.float64) + variables["logits_p/bias"]
    v = (a @ variables["logits_v/kernel"].astype(np.float64) + variables["logits_v/bias"])[:, 0]
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    p = (p + min_policy) / (1.0 + min_policy * p.shape[1])
    return p, v


def _h(a):
    """round to fp16 and back (what an operand of the fused predictor's tensor-core products carries)"""
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float64)


def forward_fp16_operands(variables, x, avg, std, M, host_len=4, other_len=7, first=1, min_policy=0.0, tanh_rel_err=0.0,
                          rng=None, value_head_fp32=True):
    """The same network with the ARITHMETIC of the fused predictor kernel (csrc/ca_predict.cu): every operand of a matrix
    product — weights, the normalised inputs, h, the ReLU activations, and the LSTM bias, which rides in the product as a
    weight row — is rounded to fp16; products accumulate in (at least) fp32; dense biases, gates and the softmax are fp32.
    value_head_fp32 (the kernel since round 2): the value head reads the fullyconnected1 outputs before their fp16
    rounding and the unrounded logits_v kernel; False reproduces the round-1 kernel (value from the fp16 head product).
    tanh_rel_err > 0 additionally perturbs every tanh by a random relative error of that size (tanh.approx: 2^-11).
    Not bit-exact with the kernel (accumulation order, the hardware's tanh), but it carries the same rounding sources, so
    its distance from `forward` is the error the kernel is expected to have against the fp32 network."""
    x = np.asarray(x, dtype=np.float64)
    xn = (x - avg) / std
    seq_len = x[:, 0]
    host = _h(xn[:, first:first + host_len])
    others = _h(xn[:, first + host_len:].reshape(-1, M, other_len))
    B, H = x.shape[0], 64
    K = _h(variables["rnn/lstm_cell/kernel"])
    bias = variables["rnn/lstm_cell/bias"].astype(np.float64).copy()
    bias[2 * H:3 * H] += 1.0                      # forget_bias folded into the packed bias row
    b = _h(bias)

    def tanh(z):
        t = np.tanh(z)
        if tanh_rel_err > 0.0:
            t = t * (1.0 + tanh_rel_err * (rng or np.random.default_rng(0)).uniform(-1, 1, size=t.shape))
        return t

    def sig(z):                                    # 0.5 tanh(z / 2) + 0.5 as the kernel evaluates it
        return 0.5 * tanh(0.5 * z) + 0.5

    h = np.zeros((B, H)); c = np.zeros((B, H))
    for t in range(M):
        z = np.concatenate([others[:, t], _h(h)], axis=1) @ K + b
        i, j, f, o = np.split(z, 4, axis=1)
        c_new = sig(f) * c + sig(i) * tanh(j)
        h_new = sig(o) * tanh(c_new)
        live = (seq_len > t)[:, None]
        c = np.where(live, c_new, c)
        h = np.where(live, h_new, h)
    a = np.concatenate([host, _h(h)], axis=1)
    for name in ("layer1", "layer2", "fullyconnected1"):
        a32 = np.maximum(a @ _h(variables[name + "/kernel"]) + variables[name + "/bias"], 0.0)
        a = _h(a32)
    logits = a @ _h(variables["logits_p/kernel"]) + variables["logits_p/bias"]
    if value_head_fp32:
        v = (a32 @ variables["logits_v/kernel"].astype(np.float64) + variables["logits_v/bias"])[:, 0]
    else:
        v = (a @ _h(variables["logits_v/kernel"]) + variables["logits_v/bias"])[:, 0]
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    p = (p + min_policy) / (1.0 + min_policy * p.shape[1])
    return p, v


def a3c_costs(p, v, y_r, a_onehot, beta, log_epsilon=1e-6):
    """NetworkVPCore.py:71-98 (sums, not means)."""
    sel = (p * a_onehot).sum(axis=1)
    cost_v = 0.5 * ((y_r - v) ** 2).sum()
    adv = np.log(np.maximum(sel, log_epsilon)) * (y_r - v)
    ent = -beta * (np.log(np.maximum(p, log_epsilon)) * p).sum(axis=1)
    cost_p = -(adv.sum() + ent.sum())
    return cost_p + cost_v, cost_p, cost_v
