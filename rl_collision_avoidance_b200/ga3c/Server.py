"""`Server().main()` — the training loop with the reference's surface (GA3C/Server.py:39-170), single process.

The reference starts 32 ProcessAgent processes, 2 ThreadPredictor and 2 ThreadTrainer threads around two
multiprocessing queues.  Here one process per GPU runs: GpuRollout.step() (batched prediction + fused env step +
experience bookkeeping on the device) and, whenever enough rows were emitted, A3C updates on the same device.
With torch.distributed initialised (one rank per GPU, torchrun) worlds are sharded across ranks, gradients are
all-reduced over NCCL (equivalent to gathering the rollouts on one trainer and broadcasting the weights, see
DESIGN.md §7), and rank 0 prints the reference's stats line and writes checkpoints.
"""
import sys
import time

import numpy as np

from .Config import get_config
from .NetworkVP_rnn import NetworkVP_rnn
from .parallel import allreduce_flat
from .rollout import GpuRollout
from ..scenarios import random_worlds


class Actions(object):
    """GCA/envs/policies/GA3C_CADRL/network.py:7-16: 11 discrete (speed fraction, heading change) actions."""

    def __init__(self):
        pi = np.pi
        self.actions = np.array([[1.0, -pi / 6], [1.0, -pi / 12], [1.0, 0.0], [1.0, pi / 12], [1.0, pi / 6],
                                 [0.5, -pi / 6], [0.5, 0.0], [0.5, pi / 6], [0.0, -pi / 6], [0.0, 0.0], [0.0, pi / 6]])
        self.num_actions = len(self.actions)


class Stats(object):
    """Counters and the stdout line of GA3C/ProcessStats.py:62-117 (rolling window over episodes)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.episode_count = 0
        self.training_count = 0
        self.total_frame_count = 0
        self.should_save_model = 0
        self.start_time = time.time()
        self.window = []            # (time, episodes, score_sum, frames) aggregates, newest last
        self.window_episodes = 0
        self.reward_log = 0.0
        self.roll_reward_log = 0.0
        self._last_save_bucket = 0

    def add(self, episodes, score_sum, frames):
        if episodes <= 0:
            return
        cfg = self.cfg
        self.episode_count += episodes
        self.total_frame_count += frames
        self.window.append((time.time(), episodes, score_sum, frames))
        self.window_episodes += episodes
        while self.window_episodes - self.window[0][1] >= cfg.STAT_ROLLING_MEAN_WINDOW and len(self.window) > 1:
            self.window_episodes -= self.window.pop(0)[1]
        self.reward_log = score_sum / episodes
        self.roll_reward_log = sum(w[2] for w in self.window) / max(self.window_episodes, 1)
        bucket = self.episode_count // cfg.SAVE_FREQUENCY
        special = any(self.episode_count - episodes < e <= self.episode_count for e in cfg.SPECIAL_EPISODES_TO_SAVE)
        if bucket > self._last_save_bucket or special:
            self._last_save_bucket = bucket
            self.should_save_model = 1

    def FPS(self):
        return np.ceil(self.total_frame_count / max(time.time() - self.start_time, 1e-9))

    def TPS(self):
        return np.ceil(self.training_count / max(time.time() - self.start_time, 1e-9))

    def return_reward_log(self):
        return self.reward_log, self.roll_reward_log

    def line(self, n_worlds):
        span = max(time.time() - self.window[0][0], 1e-9) if self.window else 1e-9
        rolling_frames = sum(w[3] for w in self.window)
        return ('[Time: %8d] [Episode: %8d Score: %10.4f] [RScore: %10.4f RPPS: %5d] [PPS: %5d TPS: %5d] '
                '[NT: %2d NP: %2d NA: %2d]' % (int(time.time() - self.start_time), self.episode_count, self.reward_log,
                                               self.roll_reward_log, rolling_frames / span, self.FPS(), self.TPS(),
                                               1, 1, n_worlds))


class Server(object):
    def __init__(self, cfg=None, device=None, num_worlds=None, seed=None):
        import torch
        self.torch = torch
        self.cfg = cfg = cfg or get_config()
        if not torch.cuda.is_available():
            raise RuntimeError("the GPU GA3C loop needs a CUDA device (there is no CPU fallback)")
        torch.backends.cuda.matmul.allow_tf32 = bool(getattr(cfg, "GPU_TRAIN_TF32", 0))
        self.dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world_size = self.dist.get_world_size() if self.dist else 1
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.stats = Stats(cfg)
        self.actions = Actions()
        self.num_actions = self.actions.num_actions
        print("[Server] Making model...")
        self.model = self.make_model()
        if cfg.TRAIN_VERSION in (cfg.LOAD_REGRESSION_THEN_TRAIN_RL, cfg.LOAD_RL_THEN_TRAIN_RL):
            try:
                self.stats.episode_count = self.model.load(
                    learning_method='regression' if cfg.TRAIN_VERSION == cfg.LOAD_REGRESSION_THEN_TRAIN_RL else 'RL')
                if cfg.TRAIN_VERSION == cfg.LOAD_REGRESSION_THEN_TRAIN_RL:
                    self.stats.episode_count = 0
            except FileNotFoundError as e:
                # the reference's regression / RL checkpoints are git-LFS pointers in this checkout
                print("[Server] no checkpoint to load (%s); starting from the random initialisation" % e)
        elif cfg.TRAIN_VERSION == cfg.TRAIN_ONLY_REGRESSION:
            raise NotImplementedError("TRAIN_ONLY_REGRESSION is out of scope (datasets are git-LFS pointers)")
        if self.dist:  # every rank starts from rank 0's weights
            for p in self.model.net.parameters():
                self.dist.broadcast(p.data, src=0)
        self.training_step = 0
        self.frame_counter = 0
        self.num_worlds = int(num_worlds or cfg.GPU_NUM_WORLDS)
        seed = cfg.RANDOM_SEED_1000 * 1000 + self.rank if seed is None else seed   # ProcessAgent.py:218
        self.rng = np.random.default_rng(seed)
        init, nag = self._new_scenarios()
        self.rollout = GpuRollout(cfg, self.model, self.num_worlds, init, nag, device=self.device_index, seed=seed)
        # from here on scenarios come from the on-device generator: every world that finishes an episode is handed a
        # fresh test case for its next one (≙ test_case_fn(**TEST_CASE_ARGS) on each env.reset())
        self._scenario_cfg = self.rollout.env.scenario_config(cfg.TEST_CASE_ARGS)
        self._scenario_seed = int(seed) * 7919 + 17
        self.rollout.env.generate_scenarios(self._scenario_cfg, self._scenario_seed, only_consumed=False)
        self.rollout.attach_scenario_generator(self._scenario_cfg, self._scenario_seed)   # refill after every step
        self._pending = []
        self._pending_rows = 0
        if int(getattr(cfg, "GPU_TRAIN_GRAPH", 1)):
            self.model.enable_graphed_training(self.train_batch_rows())
        self.history = []   # (seconds, episodes, rolling score, optimiser steps, frames) at every stats refresh

    def train_batch_rows(self):
        """Rows per optimiser step."""
        cfg = self.cfg
        return max(cfg.GPU_TRAIN_BATCH if cfg.GPU_TRAIN_BATCH > 0 else 8192, cfg.TRAINING_MIN_BATCH_SIZE + 1)

    def lr_multiplier(self):
        """Learning-rate factor that compensates for taking one optimiser step per train_batch_rows() * world_size rows
        instead of one per ~GPU_REF_BATCH rows (Config.GPU_LR_SCALE)."""
        cfg = self.cfg
        ratio = max(1.0, self.train_batch_rows() * self.world_size / float(cfg.GPU_REF_BATCH))
        mode = str(cfg.GPU_LR_SCALE)
        if mode == 'none':
            return 1.0
        if mode == 'linear':
            return ratio
        if mode == 'sqrt':
            return ratio ** 0.5
        return float(mode)

    def make_model(self):
        cfg = self.cfg
        if cfg.NET_ARCH not in cfg.ALL_ARCHS:
            raise Exception('The model name %s does not exist' % cfg.NET_ARCH)
        return NetworkVP_rnn("cuda:%d" % self.device_index, cfg.NETWORK_NAME, self.num_actions)

    def _new_scenarios(self):
        """get_testcase_random with Config.TEST_CASE_ARGS (GCA/envs/test_cases.py:95-118, config.py:50-62)."""
        cfg = self.cfg
        args = cfg.TEST_CASE_ARGS
        A = cfg.MAX_NUM_AGENTS_IN_ENVIRONMENT
        nag = self.rng.integers(2, A + 1, self.num_worlds) if A >= 2 else np.ones(self.num_worlds, dtype=int)
        policies = args.get('policies', 'learning_ga3c')
        policies = [policies] if isinstance(policies, str) else list(policies)
        return random_worlds(self.num_worlds, A, self.rng, num_agents=nag, speed_bnds=args['speed_bnds'],
                             radius_bnds=args['radius_bnds'], policies=policies, policy_distr=args.get('policy_distr'),
                             policy_to_ensure=args.get('policy_to_ensure'))

    # ---- reference surface
    def train_model(self, x_, r_, a_, trainer_id=0):
        """Server.train_model (:114-124): one optimiser step on a batch of rows."""
        if self.dist:
            costs = self._train_distributed(x_, r_, a_)
        else:
            costs = self.model.train(x_, r_, a_, trainer_id)
        self.training_step += 1
        self.frame_counter += int(x_.shape[0])
        self.stats.training_count += 1
        return costs

    def _train_distributed(self, x_, r_, a_):
        """Sum-loss gradients are summed over ranks (== one trainer seeing the concatenation of all ranks' rows)."""
        m = self.model
        x, r, a = m._as_input(x_), m._as_input(r_), self.torch.as_tensor(a_, device=m.device)
        if m.graph_rows and x.shape[0] == m.graph_rows and a.dim() == 1:
            return m.train_graphed(x, r, a, allreduce=allreduce_flat)   # graphs around the eager collective
        costs = m.backward(x, r, a)
        allreduce_flat(m.flat_grad)     # one flat buffer over EVERY parameter: zeros from ranks without rows
        m.apply_gradients()             # clip (if configured) + Adam, as on the single-GPU path
        return costs

    def save_model(self):
        if self.rank == 0:
            self.model.save(self.stats.episode_count)

    def _train_pending(self, force=False):
        cfg = self.cfg
        batch = self.train_batch_rows()
        torch = self.torch
        if self.dist:
            return self._train_pending_distributed(batch, force)
        while self._pending_rows >= batch or (force and self._pending_rows > cfg.TRAINING_MIN_BATCH_SIZE):
            x = torch.cat([p[0] for p in self._pending])
            r = torch.cat([p[1] for p in self._pending])
            a = torch.cat([p[2] for p in self._pending])
            n = min(batch, x.shape[0])
            if cfg.TRAIN_MODE:
                self.train_model(x[:n], r[:n], a[:n], 0)
            self._pending = [(x[n:], r[n:], a[n:])] if x.shape[0] > n else []
            self._pending_rows = x.shape[0] - n

    def _train_pending_distributed(self, batch, force):
        """Every rank must take part in the same number of collectives: the number of optimiser steps is agreed on
        from the all-gathered pending row counts; a rank with fewer rows contributes smaller (possibly empty) chunks."""
        torch, cfg = self.torch, self.cfg
        mine = torch.tensor([self._pending_rows], dtype=torch.int64, device="cuda")
        counts = [torch.zeros_like(mine) for _ in range(self.world_size)]
        self.dist.all_gather(counts, mine)
        counts = [int(c.item()) for c in counts]
        total = sum(counts)
        if not (total >= batch * self.world_size or (force and total > cfg.TRAINING_MIN_BATCH_SIZE)):
            return
        n_chunks = max(1, -(-max(counts) // batch))
        L1 = self.rollout.L - 1
        if self._pending:
            x = torch.cat([p[0] for p in self._pending]); r = torch.cat([p[1] for p in self._pending])
            a = torch.cat([p[2] for p in self._pending])
        else:
            x = torch.zeros((0, L1), device="cuda"); r = torch.zeros(0, device="cuda")
            a = torch.zeros(0, dtype=torch.int32, device="cuda")
        for xs, rs, as_ in zip(torch.tensor_split(x, n_chunks), torch.tensor_split(r, n_chunks), torch.tensor_split(a, n_chunks)):
            if cfg.TRAIN_MODE:
                self.train_model(xs, rs, as_, 0)
        self._pending, self._pending_rows = [], 0

    def main(self, max_steps=None, max_seconds=None, quiet=False, until_score=None):
        """Runs until Config.EPISODES (or max_steps / max_seconds / a rolling score of until_score, for tests and
        benchmarks)."""
        cfg = self.cfg
        lr_mult = (cfg.LEARNING_RATE_RL_END - cfg.LEARNING_RATE_RL_START) / cfg.ANNEALING_EPISODE_COUNT
        beta_mult = (cfg.BETA_END - cfg.BETA_START) / cfg.ANNEALING_EPISODE_COUNT
        t0 = last_print = time.time()
        steps = 0
        refresh_every = max(8, cfg.TIME_MAX)
        lr_scale = self.lr_multiplier()
        while self.stats.episode_count < cfg.EPISODES:
            step = min(self.stats.episode_count, cfg.ANNEALING_EPISODE_COUNT - 1)   # Server.main :143-148
            self.model.learning_rate = min((cfg.LEARNING_RATE_RL_START + lr_mult * step) * lr_scale, cfg.GPU_LR_MAX)
            self.model.beta = cfg.BETA_START + beta_mult * step
            self.rollout.step()
            steps += 1
            x, r, a = self.rollout.rec.take()
            if x.shape[0]:
                self._pending.append((x.clone(), r.clone(), a.clone()))
                self._pending_rows += int(x.shape[0])
            self._train_pending()
            if steps % refresh_every == 0:
                s = self.rollout.rec.pop_stats()
                if self.dist:
                    t = self.torch.tensor([s["episodes"], s["score_sum"], s["frames"]], dtype=self.torch.float64, device="cuda")
                    self.dist.all_reduce(t)
                    s = {"episodes": int(t[0].item()), "score_sum": float(t[1].item()), "frames": int(t[2].item())}
                self.stats.add(s["episodes"], s["score_sum"], s["frames"])
                self.model.check_predictor_error()
                self.history.append((time.time() - t0, self.stats.episode_count, self.stats.roll_reward_log,
                                     self.training_step, self.stats.total_frame_count))
                if cfg.SAVE_MODELS and self.stats.should_save_model > 0:
                    self.save_model()
                    self.stats.should_save_model = 0
            now = time.time()
            if not quiet and self.rank == 0 and now - last_print >= cfg.GPU_PRINT_EVERY_S and self.stats.window:
                print(self.stats.line(self.num_worlds * self.world_size))
                sys.stdout.flush()
                last_print = now
            if max_steps is not None and steps >= max_steps:
                break
            if (max_seconds is not None or until_score is not None) and steps % refresh_every == 0:
                stop = max_seconds is not None and now - t0 >= max_seconds
                if until_score is not None and self.stats.window and self.stats.roll_reward_log >= until_score:
                    stop = True
                if self.dist:  # all ranks must leave the loop in the same iteration
                    flag = self.torch.tensor([1.0 if stop else 0.0], device="cuda")
                    self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX)
                    stop = bool(flag.item() > 0)
                if stop:
                    break
        self._train_pending(force=True)
        self.torch.cuda.synchronize()
        return {"steps": steps, "seconds": time.time() - t0, "episodes": self.stats.episode_count,
                "training_steps": self.training_step, "frames": self.stats.total_frame_count}
