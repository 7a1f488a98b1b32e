"""Vectorised collision-avoidance environment: W independent worlds advanced by one CUDA launch.

Three layers, thinnest first:
  CaHandle                    the ctypes handle around ca_create/ca_destroy + raw-pointer calls
  HostVecEnv                  numpy in / numpy out through the *_host C-ABI entry points (pinned buffers);
                              this is what a host-side env.step() caller of the reference sees
  VecCollisionAvoidanceEnv    device tensors in / device tensors out (torch is only the allocator/stream
                              plumbing); this is what the on-GPU GA3C rollout uses

Replaces (vectorised): CollisionAvoidanceEnv.reset/step (GCA/envs/collision_avoidance_env.py:131-215) wrapped by
MultiagentDictToMultiagentArrayWrapper + MultiagentDummyVecEnv (GCA/envs/wrappers.py:104-139).
"""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, lib


def make_init(px, py, gx, gy, pref_speed, radius, heading, policy, time_remaining=None):
    """Pack per-agent columns (arrays of shape [W, A]) into the init tensor of ca_set_world_state."""
    px = np.asarray(px, dtype=np.float64)
    init = np.zeros(px.shape + (_abi.INIT_STRIDE,), dtype=np.float64)
    for col, v in ((_abi.I_PX, px), (_abi.I_PY, py), (_abi.I_GX, gx), (_abi.I_GY, gy), (_abi.I_PREF_SPEED, pref_speed),
                   (_abi.I_RADIUS, radius), (_abi.I_HEADING, heading), (_abi.I_POLICY, policy)):
        init[..., col] = np.asarray(v, dtype=np.float64)
    init[..., _abi.I_TIME_REMAINING] = np.nan if time_remaining is None else np.asarray(time_remaining, dtype=np.float64)
    return init


class CaHandle(object):
    """Owns one ca_env* (one GPU)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.W, self.A, self.M = cfg.num_worlds, cfg.max_agents, cfg.max_others_observed
        self.L = _abi.obs_len(self.M)
        self._h = C.c_void_p()
        check(lib().ca_create(C.byref(cfg), C.byref(self._h)), "ca_create")

    def close(self):
        if getattr(self, "_h", None):
            lib().ca_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        out = C.c_int64()
        check(lib().ca_launch_count(self._h, C.byref(out)), "ca_launch_count")
        return out.value

    def set_dt(self, dt):
        """Time step of the steps that follow (the per-call dt of CollisionAvoidanceEnv.step)."""
        check(lib().ca_set_dt(self._h, float(dt)), "ca_set_dt")

    def set_world_state(self, init_ptr, nag_ptr, on_device, stream=None, snapshot_only=False):
        if snapshot_only:
            check(lib().ca_set_reset_state(self._h, init_ptr, nag_ptr, int(on_device), stream), "ca_set_reset_state")
        else:
            check(lib().ca_set_world_state(self._h, init_ptr, nag_ptr, int(on_device), stream), "ca_set_world_state")

    def get_state_host(self):
        out = np.zeros((self.W, self.A, _abi.STATE_STRIDE), dtype=np.float64)
        check(lib().ca_get_state(self._h, out.ctypes.data_as(C.c_void_p), 0, None), "ca_get_state")
        return out


class _PinnedArray(object):
    """numpy view over ca_host_alloc'ed page-locked memory."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = C.c_void_p()
        check(lib().ca_host_alloc(C.byref(self._ptr), max(self.nbytes, 1)), "ca_host_alloc")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.array[...] = 0

    def free(self):
        if self._ptr:
            self.array = None
            lib().ca_host_free(self._ptr)
            self._ptr = C.c_void_p()


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostVecEnv(object):
    """numpy/host-buffer environment over ca_step_host (host<->device copies inside every call)."""

    def __init__(self, cfg, want_sorted_idx=False):
        self.handle = CaHandle(cfg)
        h = self.handle
        self.W, self.A, self.M, self.L = h.W, h.A, h.M, h.L
        self._pins = {
            "actions": _PinnedArray((h.W, h.A), np.int32),
            "cont": _PinnedArray((h.W, h.A, 2), np.float64),
            "obs": _PinnedArray((h.W, h.A, h.L), np.float32),
            "reward": _PinnedArray((h.W, h.A), np.float32),
            "done": _PinnedArray((h.W, h.A), np.uint8),
            "game_over": _PinnedArray((h.W,), np.uint8),
            "mask": _PinnedArray((h.W,), np.uint8),
        }
        if want_sorted_idx:
            self._pins["sorted_idx"] = _PinnedArray((h.W, h.A, h.M), np.int32)
        self.actions_buf = self._pins["actions"].array
        self.obs = self._pins["obs"].array
        self.reward = self._pins["reward"].array
        self.done = self._pins["done"].array
        self.game_over = self._pins["game_over"].array
        self.sorted_idx = self._pins["sorted_idx"].array if want_sorted_idx else None

    def close(self):
        if self.handle is not None:
            self.handle.close()
            for p in self._pins.values():
                p.free()
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def h2d_bytes_per_step(self):
        return self.actions_buf.nbytes

    @property
    def d2h_bytes_per_step(self):
        return self.obs.nbytes + self.reward.nbytes + self.done.nbytes + self.game_over.nbytes

    def set_world_state(self, init, num_agents, snapshot_only=False):
        init = np.ascontiguousarray(init, dtype=np.float64)
        num_agents = np.ascontiguousarray(num_agents, dtype=np.int32)
        if init.shape != (self.W, self.A, _abi.INIT_STRIDE):
            raise ValueError("init must have shape %s, got %s" % ((self.W, self.A, _abi.INIT_STRIDE), init.shape))
        if num_agents.shape != (self.W,):
            raise ValueError("num_agents must have shape (%d,)" % self.W)
        self.handle.set_world_state(_vp(init), _vp(num_agents), False, snapshot_only=snapshot_only)

    def set_reset_state(self, init, num_agents):
        """New scenarios that worlds pick up at their next reset / auto-reset (ca_set_reset_state)."""
        self.set_world_state(init, num_agents, snapshot_only=True)

    def reset(self, world_mask=None):
        m = None
        if world_mask is not None:
            self._pins["mask"].array[...] = np.asarray(world_mask, dtype=np.uint8)
            m = self._pins["mask"].array
        check(lib().ca_reset_host(self.handle._h, _vp(m), _vp(self.obs), _vp(self.sorted_idx)), "ca_reset_host")
        return self.obs

    def _stage_actions(self, actions, cont_actions):
        if actions is not self.actions_buf:
            self.actions_buf[...] = actions
        c = None
        if cont_actions is not None:
            self._pins["cont"].array[...] = cont_actions
            c = self._pins["cont"].array
        return c

    def step(self, actions, cont_actions=None):
        c = self._stage_actions(actions, cont_actions)
        check(lib().ca_step_host(self.handle._h, _vp(self.actions_buf), _vp(c), _vp(self.obs), _vp(self.reward),
                                 _vp(self.done), _vp(self.game_over), _vp(self.sorted_idx)), "ca_step_host")
        return self.obs, self.reward, self.done, self.game_over

    def step_async(self, actions, cont_actions=None):
        """VecEnv.step_async (baselines vec_env.py; MultiagentDummyVecEnv, GCA/envs/wrappers.py:104-109): enqueue the
        step (H2D actions, kernel, D2H results) and return; the result buffers are valid after step_wait().  With
        several HostVecEnv objects, calling step_async on the next one before step_wait on this one keeps the PCIe
        link and the GPU busy at the same time."""
        c = self._stage_actions(actions, cont_actions)
        check(lib().ca_step_host_async(self.handle._h, _vp(self.actions_buf), _vp(c), _vp(self.obs), _vp(self.reward),
                                       _vp(self.done), _vp(self.game_over), _vp(self.sorted_idx)), "ca_step_host_async")

    def step_wait(self):
        """VecEnv.step_wait: block until the step enqueued by step_async has delivered its results."""
        check(lib().ca_step_host_wait(self.handle._h), "ca_step_host_wait")
        return self.obs, self.reward, self.done, self.game_over

    def get_state(self):
        return self.handle.get_state_host()


class VecCollisionAvoidanceEnv(object):
    """Device-resident environment: actions and observations are torch CUDA tensors; one kernel launch per step
    on torch's current stream.  obs/reward/done/game_over are persistent buffers that each step overwrites."""

    def __init__(self, cfg, want_sorted_idx=False):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("VecCollisionAvoidanceEnv needs a CUDA device (there is no CPU fallback)")
        self.torch = torch
        self.handle = CaHandle(cfg)
        h = self.handle
        self.W, self.A, self.M, self.L = h.W, h.A, h.M, h.L
        self.device = torch.device("cuda", cfg.device)
        self.obs = torch.zeros((h.W, h.A, h.L), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros((h.W, h.A), dtype=torch.float32, device=self.device)
        self.done = torch.zeros((h.W, h.A), dtype=torch.uint8, device=self.device)
        self.game_over = torch.zeros((h.W,), dtype=torch.uint8, device=self.device)
        self.sorted_idx = (torch.full((h.W, h.A, h.M), -1, dtype=torch.int32, device=self.device)
                           if want_sorted_idx else None)

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def set_reset_state(self, init, num_agents):
        """New scenarios that worlds pick up at their next reset / auto-reset (ca_set_reset_state)."""
        self.set_world_state(init, num_agents, snapshot_only=True)

    def set_world_state(self, init, num_agents, snapshot_only=False):
        torch = self.torch
        init = torch.as_tensor(init, dtype=torch.float64).to(self.device).contiguous()
        nag = torch.as_tensor(num_agents, dtype=torch.int32).to(self.device).contiguous()
        if tuple(init.shape) != (self.W, self.A, _abi.INIT_STRIDE):
            raise ValueError("init must have shape %s, got %s" % ((self.W, self.A, _abi.INIT_STRIDE), tuple(init.shape)))
        if tuple(nag.shape) != (self.W,):
            raise ValueError("num_agents must have shape (%d,)" % self.W)
        if int(nag.min()) < 1 or int(nag.max()) > self.A:
            raise ValueError("num_agents must be within 1..%d" % self.A)
        self.handle.set_world_state(self._ptr(init), self._ptr(nag), True, self._stream(), snapshot_only=snapshot_only)
        if not snapshot_only:
            self.num_agents = nag
        torch.cuda.current_stream(self.device).synchronize()  # init/nag temporaries may be freed after return

    def _check_out_obs(self, out_obs):
        if out_obs is None:
            return self.obs
        if (tuple(out_obs.shape) != (self.W, self.A, self.L) or out_obs.dtype != self.torch.float32
                or not out_obs.is_cuda or not out_obs.is_contiguous()):
            raise ValueError("out_obs must be a contiguous float32 CUDA tensor of shape (%d, %d, %d)" % (self.W, self.A, self.L))
        return out_obs

    def scenario_config(self, test_case_args=None):
        """ca_scenario_config from the reference's Config.TEST_CASE_ARGS dict (GCA/envs/config.py:50-62)."""
        sc = _abi.CaScenarioConfig()
        check(lib().ca_default_scenario_config(C.byref(sc), self.A), "ca_default_scenario_config")
        a = test_case_args or {}
        if 'speed_bnds' in a:
            sc.speed_lo, sc.speed_hi = a['speed_bnds']
        if 'radius_bnds' in a:
            sc.radius_lo, sc.radius_hi = a['radius_bnds']
        if a.get('num_agents'):
            sc.min_agents = sc.max_agents = int(a['num_agents'])
        pol = a.get('policies')
        if pol is not None:
            pol = [pol] if isinstance(pol, str) else list(pol)
            distr = a.get('policy_distr') or [1.0 / len(pol)] * len(pol)
            probs = {"noncoop": 0.0, "learning_ga3c": 0.0, "static": 0.0}
            for name, pr in zip(pol, distr):
                if name not in probs:
                    raise NotImplementedError("policy %r is not available in the on-device generator" % name)
                probs[name] += pr
            sc.p_noncoop, sc.p_learning = probs["noncoop"], probs["learning_ga3c"]
            sc.ensure_learner = 1 if a.get('policy_to_ensure') == 'learning_ga3c' else 0
        return sc

    def generate_scenarios(self, scenario_cfg, seed, only_consumed=False):
        """On-device scenario generator (ca_generate_scenarios): refills the reset snapshot."""
        check(lib().ca_generate_scenarios(self.handle._h, C.byref(scenario_cfg), int(seed) & (2 ** 64 - 1),
                                          1 if only_consumed else 0, self._stream()), "ca_generate_scenarios")

    def reset(self, world_mask=None, out_obs=None):
        """out_obs: optional destination for the observation (e.g. a slot of a rollout's observation ring)."""
        obs = self._check_out_obs(out_obs)
        m = None
        if world_mask is not None:
            m = self.torch.as_tensor(world_mask, dtype=self.torch.uint8).to(self.device).contiguous()
        check(lib().ca_reset(self.handle._h, self._ptr(m), self._ptr(obs), self._ptr(self.sorted_idx),
                             self._stream()), "ca_reset")
        return obs

    def step(self, actions, cont_actions=None, out_obs=None):
        torch = self.torch
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if tuple(actions.shape) != (self.W, self.A):
            raise ValueError("actions must have shape (%d, %d)" % (self.W, self.A))
        if cont_actions is not None:
            cont_actions = cont_actions.to(device=self.device, dtype=torch.float64).contiguous()
        obs = self._check_out_obs(out_obs)
        check(lib().ca_step(self.handle._h, self._ptr(actions), self._ptr(cont_actions), self._ptr(obs),
                            self._ptr(self.reward), self._ptr(self.done), self._ptr(self.game_over),
                            self._ptr(self.sorted_idx), self._stream()), "ca_step")
        return obs, self.reward, self.done, self.game_over

    def get_state(self):
        self.torch.cuda.current_stream(self.device).synchronize()
        return self.handle.get_state_host()
