"""Policy/value network of GA3C-CADRL in PyTorch, with the reference's interface.

Reference: `NetworkVP_rnn._create_graph` (GA3C/NetworkVP_rnn.py:39-108) + `NetworkVPCore` outputs, loss and
optimiser (GA3C/NetworkVPCore.py:60-123,160-248), TensorFlow 1.15:
  x_n = (x - AVG) / STD; seq_len = x[:, 0] (raw num_other_agents); host = x_n[:, 1:5]; others = x_n[:, 5:] as
  (B, M, 7) -> tf.nn.dynamic_rnn(LSTMCell(64), sequence_length=seq_len) -> final h -> concat(host, h) ->
  Dense256 ReLU (layer1) -> Dense256 ReLU (layer2) -> Dense256 ReLU (fullyconnected1) -> {logits_p(11) softmax,
  logits_v(1)}.
TF1 `LSTMCell` semantics are kept so that reference checkpoints map one to one: a single kernel
[(7+64), 4*64] applied to concat(x_t, h), gate order i, j, f, o, forget bias 1.0:
  c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j);  h' = sigmoid(o) * tanh(c');
rows whose sequence ended (t >= seq_len) keep their state.  Parameter names follow the TF variable names
(`rnn/lstm_cell/kernel`, `layer1/kernel`, ...).  The A3C loss uses sums, not means (NetworkVPCore.py:71-98);
Adam follows TF's formulation (epsilon outside the bias-corrected sqrt).
"""
import math
import os
import re

import torch

from .Config import get_config
from .. import _abi


def _glorot_uniform(fan_in, fan_out, generator):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(fan_in, fan_out, generator=generator) * 2 - 1) * limit


class _FusedLSTMCell(torch.autograd.Function):
    """One LSTM time step (TF-1.x LSTMCell semantics under dynamic_rnn's sequence mask) through the fused CUDA cell kernels
    (ca_lstm_cell_forward / ca_lstm_cell_backward): (z [B, 256], c [B, 64], h [B, 64], x_raw, t) -> (c', h').  x_raw's
    column 0 is the sequence length.  The matmuls that produce z stay in the framework (cuBLAS), so autograd differentiates
    them as usual."""

    @staticmethod
    def forward(ctx, z, c_prev, h_prev, x_raw, t):
        import ctypes as C
        from .._lib import check, lib
        z, c_prev, h_prev = z.contiguous(), c_prev.contiguous(), h_prev.contiguous()
        B = z.shape[0]
        gates = torch.empty_like(z)
        c, h = torch.empty_like(c_prev), torch.empty_like(h_prev)
        ptr = lambda a: C.c_void_p(a.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
        check(lib().ca_lstm_cell_forward(ptr(z), ptr(c_prev), ptr(h_prev), ptr(x_raw), int(x_raw.stride(0)), int(t), ptr(gates),
                                         ptr(c), ptr(h), B, z.device.index or 0, stream), "ca_lstm_cell_forward")
        ctx.save_for_backward(gates, c_prev, c, x_raw)
        ctx.t = int(t)
        return c, h

    @staticmethod
    def backward(ctx, dc, dh):
        import ctypes as C
        from .._lib import check, lib
        gates, c_prev, c, x_raw = ctx.saved_tensors
        B = gates.shape[0]
        dc = None if dc is None else dc.contiguous()
        dh = None if dh is None else dh.contiguous()
        dz = torch.empty_like(gates)
        dc_prev, dh_pass = torch.empty_like(c_prev), torch.empty_like(c_prev)
        ptr = lambda a: None if a is None else C.c_void_p(a.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(gates.device).cuda_stream)
        check(lib().ca_lstm_cell_backward(ptr(gates), ptr(c_prev), ptr(c), ptr(x_raw), int(x_raw.stride(0)), ctx.t, ptr(dc),
                                          ptr(dh), ptr(dz), ptr(dc_prev), ptr(dh_pass), B, gates.device.index or 0, stream),
              "ca_lstm_cell_backward")
        return dz, dc_prev, dh_pass, None, None


class PolicyValueNet(torch.nn.Module):
    """The function only (no optimiser): x [B, NN_INPUT_SIZE] float32 -> (softmax_p [B, num_actions], v [B])."""

    HIDDEN = 64

    def __init__(self, cfg, num_actions, seed=0):
        super().__init__()
        self.M = cfg.MAX_NUM_OTHER_AGENTS_OBSERVED
        self.other_len = cfg.OTHER_AGENT_FULL_OBSERVATION_LENGTH
        self.host_len = cfg.HOST_AGENT_STATE_SIZE
        self.first = cfg.FIRST_STATE_INDEX
        self.normalize = bool(cfg.NORMALIZE_INPUT)
        self.min_policy = float(cfg.MIN_POLICY)
        self.num_actions = num_actions
        g = torch.Generator().manual_seed(seed)
        H = self.HIDDEN
        P = torch.nn.Parameter
        # TF variable name -> parameter (kernels are [in, out] like TF); '/' becomes '__' in the torch names
        shapes = [("rnn/lstm_cell", self.other_len + H, 4 * H), ("layer1", self.host_len + H, 256), ("layer2", 256, 256),
                  ("fullyconnected1", 256, 256), ("logits_p", 256, num_actions), ("logits_v", 256, 1)]
        self.params = torch.nn.ParameterDict()
        for name, fan_in, fan_out in shapes:
            self.params[(name + "/kernel").replace("/", "__")] = P(_glorot_uniform(fan_in, fan_out, g))
            self.params[(name + "/bias").replace("/", "__")] = P(torch.zeros(fan_out))
        self.register_buffer("avg", torch.tensor(cfg.NN_INPUT_AVG_VECTOR, dtype=torch.float32))
        self.register_buffer("std", torch.tensor(cfg.NN_INPUT_STD_VECTOR, dtype=torch.float32))

    def w(self, tf_name):
        return self.params[tf_name.replace("/", "__")]

    def tf_variables(self):
        """{TF variable name: numpy array} (for checkpoint exchange and the NumPy oracle)."""
        return {k.replace("__", "/"): v.detach().cpu().numpy() for k, v in self.params.items()}

    def load_tf_variables(self, variables):
        with torch.no_grad():
            for name, value in variables.items():
                self.w(name).copy_(torch.as_tensor(value, dtype=torch.float32))
        self.weights_version = getattr(self, "weights_version", 0) + 1   # invalidates the packed predictor image

    def features(self, x):
        H = self.HIDDEN
        xn = (x - self.avg) / self.std if self.normalize else x
        seq_len = x[:, 0]
        host = xn[:, self.first:self.first + self.host_len]
        others = xn[:, self.first + self.host_len:].reshape(-1, self.M, self.other_len)
        B = x.shape[0]
        h = x.new_zeros((B, H))
        c = x.new_zeros((B, H))
        K, b = self.w("rnn/lstm_cell/kernel"), self.w("rnn/lstm_cell/bias")
        Kx, Kh = K[:self.other_len], K[self.other_len:]
        if x.is_cuda and H == 64 and x.dtype == torch.float32 and os.environ.get("GA3C_FUSED_TRAIN_CELL", "1") != "0":
            # CUDA: the input projections of all M steps are one batched matmul, each step is one cuBLAS addmm plus one
            # fused cell kernel (forward) / one fused cell kernel plus the addmm gradients (backward)
            xc = x.contiguous()
            zx = (torch.matmul(others, Kx) + b).unbind(1)
            for t in range(self.M):
                z = zx[t] if t == 0 else torch.addmm(zx[t], h, Kh)
                c, h = _FusedLSTMCell.apply(z, c, h, xc, t)
            l1_in = torch.cat([host, h], dim=1)
            l1 = torch.relu(l1_in @ self.w("layer1/kernel") + self.w("layer1/bias"))
            l2 = torch.relu(l1 @ self.w("layer2/kernel") + self.w("layer2/bias"))
            return torch.relu(l2 @ self.w("fullyconnected1/kernel") + self.w("fullyconnected1/bias"))
        for t in range(self.M):
            z = others[:, t] @ Kx + h @ Kh + b
            i, j, f, o = z.split(H, dim=1)
            c_new = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
            h_new = torch.sigmoid(o) * torch.tanh(c_new)
            live = (seq_len > t).unsqueeze(1)
            c = torch.where(live, c_new, c)
            h = torch.where(live, h_new, h)
        l1_in = torch.cat([host, h], dim=1)
        l1 = torch.relu(l1_in @ self.w("layer1/kernel") + self.w("layer1/bias"))
        l2 = torch.relu(l1 @ self.w("layer2/kernel") + self.w("layer2/bias"))
        return torch.relu(l2 @ self.w("fullyconnected1/kernel") + self.w("fullyconnected1/bias"))

    def forward(self, x):
        fc1 = self.features(x)
        logits_p = fc1 @ self.w("logits_p/kernel") + self.w("logits_p/bias")
        v = (fc1 @ self.w("logits_v/kernel") + self.w("logits_v/bias")).squeeze(1)
        p = (torch.softmax(logits_p, dim=1) + self.min_policy) / (1.0 + self.min_policy * self.num_actions)
        return p, v, logits_p


def flatten_parameters(params):
    """Rebinds the parameters (and their .grad) as views into one flat buffer each; returns (flat_param, flat_grad).
    The optimiser and the gradient all-reduce then work on single tensors, and gradients accumulate in place."""
    params = list(params)
    n = sum(p.numel() for p in params)
    dev = params[0].device
    flat_param = torch.empty(n, dtype=torch.float32, device=dev)
    flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
    off = 0
    with torch.no_grad():
        for p in params:
            k = p.numel()
            flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat_param[off:off + k].view_as(p)
            p.grad = flat_grad[off:off + k].view_as(p)
            off += k
    return flat_param, flat_grad


class TFAdam(object):
    """tf.train.AdamOptimizer update rule (beta1 .9, beta2 .999, eps 1e-8):
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  var -= lr_t * m / (sqrt(v) + eps).
    Works on ONE flat parameter / gradient buffer (the parameters are views into it, see NetworkVP_rnn._flatten), so a
    step is a handful of element-wise kernels; step_device() takes the learning rate as a device scalar and keeps the
    step count on the device, which makes it capturable in a CUDA graph."""

    def __init__(self, params, flat_param=None, flat_grad=None, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params = [p for p in params]
        if flat_param is None:
            flat_param, flat_grad = flatten_parameters(self.params)
        self.flat_param, self.flat_grad = flat_param, flat_grad
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.flat_m = torch.zeros_like(flat_param)
        self.flat_v = torch.zeros_like(flat_param)
        self.m, self.v, off = [], [], 0
        for p in self.params:   # per-parameter views (checkpoint layout)
            n = p.numel()
            self.m.append(self.flat_m[off:off + n].view_as(p))
            self.v.append(self.flat_v[off:off + n].view_as(p))
            off += n
        self.t = 0
        self.t_dev = torch.zeros((), dtype=torch.float64, device=flat_param.device)

    @torch.no_grad()
    def step(self, lr):
        self.t += 1
        self.t_dev.fill_(self.t)
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        g = self.flat_grad
        self.flat_m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
        self.flat_v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
        self.flat_param.addcdiv_(self.flat_m, self.flat_v.sqrt().add_(self.eps), value=-lr_t)

    @torch.no_grad()
    def step_device(self, lr_dev):
        """The same update with lr a 0-dim float64 device tensor and the step count read from / advanced on the device."""
        self.t_dev.add_(1.0)
        lr_t = (lr_dev * torch.sqrt(1.0 - torch.pow(self.b2, self.t_dev)) / (1.0 - torch.pow(self.b1, self.t_dev))).to(torch.float32)
        g = self.flat_grad
        self.flat_m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
        self.flat_v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
        self.flat_param.sub_(lr_t * self.flat_m / self.flat_v.sqrt().add_(self.eps))

    def state_dict(self):
        return {"m": self.m, "v": self.v, "t": self.t}

    def load_state_dict(self, sd):
        for dst, src in zip(self.m, sd["m"]):
            dst.copy_(src)
        for dst, src in zip(self.v, sd["v"]):
            dst.copy_(src)
        self.t = int(sd["t"])
        self.t_dev.fill_(self.t)


class GraphedTrainStep(object):
    """One optimiser step on a FIXED number of rows — gradient zeroing, forward, sum losses, backward, (clip,) TF-Adam —
    captured once and replayed as CUDA graphs: a step is then bound by its ~100 small kernels' GPU time instead of their
    launch overhead (8192 rows: 2.1 ms eager -> a few hundred microseconds).  With `allreduce` given (distributed
    trainer) the capture is split around the gradient all-reduce: graph A = zero / forward / backward, the collective runs
    eagerly on the flat gradient buffer, graph B = clip / Adam.  Learning rate and entropy weight are device scalars, so
    annealing them needs no re-capture."""

    def __init__(self, model, rows, allreduce=None):
        self.model, self.rows, self.allreduce = model, int(rows), allreduce
        dev = model.device
        L1 = model.cfg.NN_INPUT_SIZE
        self.x = torch.zeros((self.rows, L1), dtype=torch.float32, device=dev)
        self.r = torch.zeros(self.rows, dtype=torch.float32, device=dev)
        self.a = torch.zeros(self.rows, dtype=torch.int64, device=dev)
        self.lr = torch.zeros((), dtype=torch.float64, device=dev)
        self.beta = torch.zeros((), dtype=torch.float32, device=dev)
        self._lr_val = self._beta_val = None
        self.costs = None
        m = model
        saved = (m.flat_param.clone(), m.opt.flat_m.clone(), m.opt.flat_v.clone(), m.opt.t)
        self._set_scalars(m.learning_rate, m.beta)
        try:   # the .grad views were created on another stream than the capture stream: intended, not a hazard
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):   # warm-up on the side stream (allocator, cuBLAS workspaces, lazy kernel loading)
                self._backward_body()
                self._update_body()
            side.synchronize()
            self.g_bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_bwd, stream=side, capture_error_mode="thread_local"):
                self._backward_body()
                if allreduce is None:
                    self._update_body()
            self.g_upd = None
            if allreduce is not None:
                self.g_upd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_upd, stream=side, capture_error_mode="thread_local"):
                    self._update_body()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():   # the warm-up steps must not count
            m.flat_param.copy_(saved[0]); m.opt.flat_m.copy_(saved[1]); m.opt.flat_v.copy_(saved[2])
        m.opt.t = saved[3]
        m.opt.t_dev.fill_(m.opt.t)

    def _set_scalars(self, lr, beta):
        if lr != self._lr_val:
            self.lr.fill_(lr)
            self._lr_val = lr
        if beta != self._beta_val:
            self.beta.fill_(beta)
            self._beta_val = beta

    def _backward_body(self):
        m = self.model
        m.flat_grad.zero_()
        beta_saved, m.beta = m.beta, self.beta
        try:
            self.costs = m.losses(self.x, self.r, self.a)
            self.costs["cost_all"].backward()
        finally:
            m.beta = beta_saved

    def _update_body(self):
        m = self.model
        m._clip_gradients()
        m.opt.step_device(self.lr)

    def run(self, x, y_r, a):
        m = self.model
        self.x.copy_(x); self.r.copy_(y_r); self.a.copy_(a)
        self._set_scalars(m.learning_rate, m.beta)
        self.g_bwd.replay()
        if self.g_upd is not None:
            self.allreduce(m.flat_grad)
            self.g_upd.replay()
        m.opt.t += 1
        m.global_step += 1
        m.last_costs = self.costs
        return self.costs


class NetworkVP_rnn(object):
    """Interface of GA3C/NetworkVPCore.py:151-268: predict_p_and_v / predict_p / predict_v / predict_single /
    train / log / save / load / get_global_step / get_variables_names / get_variable_value."""

    def __init__(self, device, model_name, num_actions, seed=0):
        cfg = get_config()
        self.cfg = cfg
        if isinstance(device, str) and device.startswith('/'):  # TF style '/cpu:0', '/gpu:0'
            m = re.match(r"/(cpu|gpu):(\d+)", device)
            device = "cpu" if m.group(1) == "cpu" else "cuda:%s" % m.group(2)
        self.device = torch.device(device)
        self.model_name = model_name
        self.num_actions = num_actions
        self.learning_rate_rl = cfg.LEARNING_RATE_RL_START
        self.learning_rate = cfg.LEARNING_RATE_RL_START
        self.beta = cfg.BETA_START
        self.log_epsilon = cfg.LOG_EPSILON
        self.net = PolicyValueNet(cfg, num_actions, seed=seed).to(self.device)
        self._flatten()
        self.opt = TFAdam(self.net.parameters(), self.flat_param, self.flat_grad)
        self._graphed = {}     # rows -> GraphedTrainStep
        self.graph_rows = 0    # batches of exactly this many rows are trained from CUDA graphs (enable_graphed_training)
        self.global_step = 0
        self.checkpoints_save_dir = os.environ.get(
            "GA3C_CHECKPOINT_DIR", os.path.join(os.getcwd(), "checkpoints", "RL_tmp"))
        self.last_costs = {}

    def _flatten(self):
        """All parameters become views into one flat buffer, all gradients views into another: the optimiser and the
        gradient all-reduce work on single tensors, and gradients accumulate in place (CUDA-graph friendly)."""
        self.flat_param, self.flat_grad = flatten_parameters(self.net.parameters())

    def enable_graphed_training(self, rows):
        """Batches of exactly `rows` rows are trained from CUDA graphs from now on (CUDA only; other sizes run eagerly)."""
        self.graph_rows = int(rows) if self.device.type == "cuda" else 0

    def _clip_gradients(self):
        if self.cfg.USE_GRAD_CLIP:
            for p in self.net.parameters():   # tf.clip_by_average_norm
                avg_norm = p.grad.norm() / p.grad.numel()
                p.grad.mul_(torch.clamp(self.cfg.GRAD_CLIP_NORM / (avg_norm + 1e-12), max=1.0))

    # ---- prediction (GA3C/NetworkVPCore.py:160-176)
    def _as_input(self, x):
        return torch.as_tensor(x, dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def predict_p_and_v_device(self, x):
        """Device tensors in and out (used by the on-GPU rollout): x [B, NN_INPUT_SIZE] -> p [B, 11], v [B]."""
        p, v, _ = self.net(x)
        return p, v

    @torch.no_grad()
    def predict_from_obs(self, obs):
        """Predictor on raw observation rows obs [B, L] (column 0 = is_learning), CUDA only: the LSTM runs through the
        fused ca_lstm_step kernel (input projection + gates + state update + sequence mask in one pass per step, no
        normalised or per-gate intermediates in HBM); the dense layers are cuBLAS GEMMs.  Same function as
        `predict_p_and_v_device(obs[:, 1:])` (tests/test_gpu_ga3c.py checks them against each other)."""
        import ctypes as C
        from .._lib import check, lib
        net = self.net
        B, L = obs.shape
        if not obs.is_cuda or obs.dtype != torch.float32 or obs.stride(1) != 1:
            raise ValueError("obs must be a float32 CUDA tensor with contiguous rows")
        H, M = net.HIDDEN, net.M
        K = net.w("rnn/lstm_cell/kernel")
        Kx, Kh = K[:net.other_len].contiguous(), K[net.other_len:].contiguous()
        bias = net.w("rnn/lstm_cell/bias")
        off = net.first + net.host_len                      # first other-agent column of the NN input
        avg7, std7 = net.avg[off:off + 7].contiguous(), net.std[off:off + 7].contiguous()
        c = torch.zeros((B, H), dtype=torch.float32, device=obs.device)
        h = torch.zeros((B, H), dtype=torch.float32, device=obs.device)
        p_ = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        dev = obs.device.index or 0
        zh = None
        for t in range(M):
            if t > 0:
                zh = h @ Kh
            check(lib().ca_lstm_step(p_(obs), int(obs.stride(0)), p_(zh) if zh is not None else None, p_(Kx), p_(bias),
                                     p_(avg7), p_(std7), p_(c), p_(h), B, t, dev, stream), "ca_lstm_step")
        host = (obs[:, 1 + net.first:1 + net.first + net.host_len] - net.avg[net.first:net.first + net.host_len]) \
            / net.std[net.first:net.first + net.host_len]
        a = torch.relu(torch.addmm(net.w("layer1/bias"), torch.cat([host, h], dim=1), net.w("layer1/kernel")))
        a = torch.relu(torch.addmm(net.w("layer2/bias"), a, net.w("layer2/kernel")))
        a = torch.relu(torch.addmm(net.w("fullyconnected1/bias"), a, net.w("fullyconnected1/kernel")))
        logits = torch.addmm(net.w("logits_p/bias"), a, net.w("logits_p/kernel"))
        v = torch.addmv(net.w("logits_v/bias"), a, net.w("logits_v/kernel").squeeze(1))
        p = (torch.softmax(logits, dim=1) + net.min_policy) / (1.0 + net.min_policy * net.num_actions)
        return p, v

    # ---- fused predictor (csrc/ca_predict.cu): the whole forward pass + action selection in one tcgen05 kernel launch
    def fused_supported(self):
        return self.device.type == "cuda" and self.net.M <= _abi.CA_PREDICTOR_MAX_OTHERS and self.net.HIDDEN == 64 \
            and self.net.host_len == 4 and self.net.other_len == 7 and self.net.first == 1 and self.num_actions == 11 \
            and self.net.normalize

    def mark_weights_changed(self):
        """Call after modifying the parameters outside train() / load(): the packed image is rebuilt on next use."""
        self._packed_version = -1

    def _packed_blob(self):
        import ctypes as C
        from .._lib import check, lib
        version = (self.global_step, getattr(self, "_load_count", 0), getattr(self.net, "weights_version", 0))
        if getattr(self, "_packed_version", None) == version:
            return self._blob
        net = self.net
        if not hasattr(self, "_blob"):
            self._blob = torch.empty(_abi.CA_PREDICTOR_BLOB_BYTES, dtype=torch.uint8, device=self.device)
            self._pred_error = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._pred_calls = 0
        keep = [net.w("rnn/lstm_cell/kernel"), net.w("rnn/lstm_cell/bias"), net.w("layer1/kernel"), net.w("layer1/bias"),
                net.w("layer2/kernel"), net.w("layer2/bias"), net.w("fullyconnected1/kernel"), net.w("fullyconnected1/bias"),
                net.w("logits_p/kernel"), net.w("logits_p/bias"), net.w("logits_v/kernel"), net.w("logits_v/bias"),
                net.avg, net.std]
        keep = [t.detach().contiguous() for t in keep]
        params = _abi.CaPredictorParams(*[C.c_void_p(t.data_ptr()) for t in keep])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib().ca_predictor_pack(C.byref(params), C.c_void_p(self._blob.data_ptr()), self.device.index or 0, stream),
              "ca_predictor_pack")
        self._packed_version = version
        return self._blob

    @torch.no_grad()
    def predict_fused(self, obs, want_p=True, want_actions=False, greedy=False, seed=0, learning_only=False):
        """ThreadPredictor + select_action for raw observation rows obs [B, L] (float32 CUDA, contiguous rows) in ONE
        launch of the fused tcgen05 kernel: returns (p [B, 11] or None, v [B], actions int32 [B] or None).  fp16
        operands / fp32 accumulation (tests/test_gpu_predictor.py states the tolerance against `net(obs[:, 1:])`).
        learning_only=True predicts only the rows whose is_learning column is set — the rows the reference's actors send
        to ThreadPredictor (ProcessAgent.py:128-133) — grouped by LSTM sequence length (ca_predict_plan + ca_predict_rows);
        the other rows get v = 0, action 0 and p = 0."""
        import ctypes as C
        from .._lib import check, lib
        if not self.fused_supported():
            raise RuntimeError("fused predictor needs CUDA and the GA3C-CADRL architecture (64 LSTM units, <= 22 others)")
        B, L = obs.shape
        if not obs.is_cuda or obs.dtype != torch.float32 or obs.stride(1) != 1:
            raise ValueError("obs must be a float32 CUDA tensor with contiguous rows")
        blob = self._packed_blob()
        alloc = torch.zeros if learning_only else torch.empty
        p = alloc((B, self.num_actions), dtype=torch.float32, device=obs.device) if want_p else None
        v = torch.empty(B, dtype=torch.float32, device=obs.device)
        actions = torch.empty(B, dtype=torch.int32, device=obs.device) if want_actions else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        self._pred_calls += 1
        if learning_only:
            if getattr(self, "_plan_rows", None) is None or self._plan_rows.numel() < B or self._plan_rows.device != obs.device:
                self._plan_rows = torch.empty(B, dtype=torch.int32, device=obs.device)
                self._plan_counters = torch.zeros(_abi.CA_PREDICT_PLAN_COUNTERS, dtype=torch.int32, device=obs.device)
            dev = obs.device.index or 0
            check(lib().ca_predict_plan(ptr(obs), int(obs.stride(0)), B, self.net.M, ptr(self._plan_rows),
                                        ptr(self._plan_counters), ptr(v), ptr(actions), dev, stream), "ca_predict_plan")
            check(lib().ca_predict_rows(ptr(obs), int(obs.stride(0)), B, self.net.M, ptr(blob), ptr(self._plan_rows),
                                        ptr(self._plan_counters), ptr(p), ptr(v), ptr(actions), 1 if greedy else 0,
                                        float(self.net.min_policy), int(seed) & (2 ** 64 - 1), self._pred_calls,
                                        ptr(self._pred_error), dev, stream), "ca_predict_rows")
            return p, v, actions
        check(lib().ca_predict(ptr(obs), int(obs.stride(0)), B, self.net.M, ptr(blob), ptr(p), ptr(v), ptr(actions),
                               1 if greedy else 0, float(self.net.min_policy), int(seed) & (2 ** 64 - 1), self._pred_calls,
                               ptr(self._pred_error), obs.device.index or 0, stream), "ca_predict")
        return p, v, actions

    def check_predictor_error(self):
        """Raises if the fused predictor flagged a tile (an mbarrier wait that timed out leaves garbage actions / values);
        clears the flag.  One 4-byte D2H read: call it where the loop synchronises anyway."""
        err = getattr(self, "_pred_error", None)
        if err is not None and int(err.item()) != 0:
            err.zero_()
            raise RuntimeError("fused predictor reported a timed-out tile (device error flag set)")

    def predict_p_and_v(self, x):
        p, v = self.predict_p_and_v_device(self._as_input(x))
        return p.cpu().numpy(), v.cpu().numpy()

    def predict_p(self, x):
        return self.predict_p_and_v(x)[0]

    def predict_v(self, x):
        return self.predict_p_and_v(x)[1]

    def predict_single(self, x):
        return self.predict_p(x[None, :])[0]

    # ---- A3C loss (GA3C/NetworkVPCore.py:64-98)
    def losses(self, x, y_r, action_index):
        """action_index: one-hot float [B, num_actions] (reference layout) or int64 [B] action ids."""
        p, v, _ = self.net(x)
        if action_index.dim() == 2:
            sel = (p * action_index).sum(dim=1)
        else:
            sel = p.gather(1, action_index.long().unsqueeze(1)).squeeze(1)
        cost_v = 0.5 * ((y_r - v) ** 2).sum()
        cost_p_advant = torch.log(torch.clamp(sel, min=self.log_epsilon)) * (y_r - v.detach())
        cost_p_entrop = -1.0 * self.beta * (torch.log(torch.clamp(p, min=self.log_epsilon)) * p).sum(dim=1)
        cost_p = -(cost_p_advant.sum() + cost_p_entrop.sum())
        return {"cost_all": cost_p + cost_v, "cost_p": cost_p, "cost_v": cost_v,
                "cost_p_advant_agg": cost_p_advant.sum(), "cost_p_entrop_agg": cost_p_entrop.sum()}

    def train(self, x, y_r, a, trainer_id=0, learning_method='RL'):
        if learning_method != 'RL':
            raise NotImplementedError("regression pre-training is out of scope (datasets are git-LFS pointers)")
        x = self._as_input(x)
        y_r = self._as_input(y_r)
        a = torch.as_tensor(a, device=self.device)
        if self.graph_rows and x.shape[0] == self.graph_rows and a.dim() == 1:
            return self.train_graphed(x, y_r, a)
        costs = self.backward(x, y_r, a)
        self.apply_gradients()
        return costs

    def train_graphed(self, x, y_r, a, allreduce=None):
        key = (int(x.shape[0]), allreduce is not None)
        step = self._graphed.get(key)
        if step is None:
            step = self._graphed[key] = GraphedTrainStep(self, x.shape[0], allreduce)
        return step.run(x, y_r, a)

    def backward(self, x, y_r, a):
        """Sum-loss gradients of one batch of rows into p.grad for EVERY parameter (an empty batch gives zeros, so that
        ranks with nothing to contribute still take part in the gradient all-reduce)."""
        self.flat_grad.zero_()    # p.grad are views into it: backward accumulates in place
        if x.shape[0] == 0:
            zero = torch.zeros((), device=self.device)
            self.last_costs = {k: zero for k in ("cost_all", "cost_p", "cost_v", "cost_p_advant_agg", "cost_p_entrop_agg")}
            return self.last_costs
        costs = self.losses(x, y_r, a)
        costs["cost_all"].backward()
        self.last_costs = costs
        return costs

    def apply_gradients(self):
        """tf.clip_by_average_norm (if configured) + the TF-Adam step on whatever is in p.grad (NetworkVPCore.py:100-123);
        the distributed trainer calls it after the gradient all-reduce, so both paths train the same way."""
        self._clip_gradients()
        self.opt.step(self.learning_rate)
        self.global_step += 1

    def get_global_step(self):
        return self.global_step

    def log(self, x, y_r, a, reward, roll_reward, episode):
        c = {k: float(v) for k, v in self.last_costs.items()}
        print("[NetworkVP] step %d episode %d reward %.4f roll_reward %.4f %s" % (self.global_step, episode, reward, roll_reward, c))

    # ---- checkpoints: same file naming as GA3C/NetworkVPCore.py:219-248 ('<model_name>_%08d')
    def _checkpoint_filename(self, episode, mode='save', learning_method='RL', wandb_runid_for_loading=None):
        d = self.checkpoints_save_dir
        if mode == 'load' and wandb_runid_for_loading is not None:
            d = os.path.join(os.path.dirname(self.checkpoints_save_dir), learning_method, 'wandb', wandb_runid_for_loading,
                             'checkpoints')
        return os.path.join(d, '%s_%08d' % (self.model_name, episode))

    @staticmethod
    def _get_episode_from_filename(filename):
        return int(re.split(r'/|_|\.', filename)[-2 if filename.endswith('.pt') else -1])

    def save(self, episode, learning_method='RL'):
        path = self._checkpoint_filename(episode, mode='save', learning_method=learning_method) + ".pt"
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"variables": {k: torch.as_tensor(v) for k, v in self.net.tf_variables().items()},
                    "adam": self.opt.state_dict(), "step": self.global_step, "episode": episode}, path)
        return path

    def load(self, learning_method='RL', path=None):
        if path is None:
            d = self.checkpoints_save_dir
            cands = sorted(f for f in (os.listdir(d) if os.path.isdir(d) else []) if f.startswith(self.model_name + "_"))
            if not cands:
                raise FileNotFoundError("no checkpoint '%s_*' in %s" % (self.model_name, d))
            ep = self.cfg.EPISODE_NUMBER_TO_LOAD if getattr(self.cfg, 'EPISODE_NUMBER_TO_LOAD', 0) > 0 else None
            path = os.path.join(d, ('%s_%08d.pt' % (self.model_name, ep)) if ep else cands[-1])
        print("[NetworkVPCore] Loading checkpoint file:", path)
        ck = torch.load(path, map_location=self.device)
        self.net.load_tf_variables(ck["variables"])
        self._load_count = getattr(self, "_load_count", 0) + 1
        self.opt.load_state_dict(ck["adam"])
        self.global_step = int(ck["step"])
        return int(ck.get("episode", self._get_episode_from_filename(path)))

    def get_variables_names(self):
        return [k + ":0" for k in self.net.tf_variables()]

    def get_variable_value(self, name):
        return self.net.tf_variables()[name.split(":")[0]]
