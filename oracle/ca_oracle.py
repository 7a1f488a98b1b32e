"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/_build/libca_oracle.so (C restatement of the
reference's env.step path, see ca_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rl_collision_avoidance_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libca_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "ca_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "ca_step.h")
    if (not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(hdr)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-B"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ca_oracle_create.argtypes = [C.POINTER(_abi.CaConfig), C.POINTER(vp)]
        L.ca_oracle_destroy.argtypes = [vp]
        L.ca_oracle_destroy.restype = None
        L.ca_oracle_set_world_state.argtypes = [vp, vp, vp]
        L.ca_oracle_reset.argtypes = [vp, vp, vp, vp]
        L.ca_oracle_set_reset_state.argtypes = [vp, vp, vp]
        L.ca_oracle_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.ca_oracle_get_state.argtypes = [vp, vp]
        L.ca_oracle_nstep_returns.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_double]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv(object):
    """Batched CPU oracle with the same tensor contract as the C-ABI of include/ca_step.h,
    except that obs and reward are float64 (the reference's own precision)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.W, self.A, self.M = cfg.num_worlds, cfg.max_agents, cfg.max_others_observed
        self.L = _abi.obs_len(self.M)
        self._h = C.c_void_p()
        rc = lib().ca_oracle_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError("ca_oracle_create failed: %d" % rc)
        self.obs = np.zeros((self.W, self.A, self.L), dtype=np.float64)
        self.reward = np.zeros((self.W, self.A), dtype=np.float64)
        self.done = np.zeros((self.W, self.A), dtype=np.uint8)
        self.game_over = np.zeros((self.W,), dtype=np.uint8)
        self.sorted_idx = np.zeros((self.W, self.A, self.M), dtype=np.int32)

    def close(self):
        if self._h:
            lib().ca_oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_world_state(self, init, num_agents):
        init = np.ascontiguousarray(init, dtype=np.float64)
        num_agents = np.ascontiguousarray(num_agents, dtype=np.int32)
        assert init.shape == (self.W, self.A, _abi.INIT_STRIDE), init.shape
        assert num_agents.shape == (self.W,)
        rc = lib().ca_oracle_set_world_state(self._h, _p(init), _p(num_agents))
        if rc != 0:
            raise RuntimeError("ca_oracle_set_world_state failed: %d" % rc)

    def set_reset_state(self, init, num_agents):
        init = np.ascontiguousarray(init, dtype=np.float64)
        num_agents = np.ascontiguousarray(num_agents, dtype=np.int32)
        assert init.shape == (self.W, self.A, _abi.INIT_STRIDE), init.shape
        rc = lib().ca_oracle_set_reset_state(self._h, _p(init), _p(num_agents))
        if rc != 0:
            raise RuntimeError("ca_oracle_set_reset_state failed: %d" % rc)

    def reset(self, world_mask=None):
        m = None if world_mask is None else np.ascontiguousarray(world_mask, dtype=np.uint8)
        rc = lib().ca_oracle_reset(self._h, _p(m), _p(self.obs), _p(self.sorted_idx))
        if rc != 0:
            raise RuntimeError("ca_oracle_reset failed: %d" % rc)
        return self.obs

    def step(self, actions, cont_actions=None, nthreads=1):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        assert actions.shape == (self.W, self.A)
        if cont_actions is not None:
            cont_actions = np.ascontiguousarray(cont_actions, dtype=np.float64)
            assert cont_actions.shape == (self.W, self.A, 2)
        rc = lib().ca_oracle_step(self._h, _p(actions), _p(cont_actions), _p(self.obs), _p(self.reward),
                                  _p(self.done), _p(self.game_over), _p(self.sorted_idx), int(nthreads))
        if rc != 0:
            raise RuntimeError("ca_oracle_step failed: %d" % rc)
        return self.obs, self.reward, self.done, self.game_over

    def get_state(self):
        out = np.zeros((self.W, self.A, _abi.STATE_STRIDE), dtype=np.float64)
        lib().ca_oracle_get_state(self._h, _p(out))
        return out


def nstep_returns(reward, bootstrap, gamma):
    reward = np.ascontiguousarray(reward, dtype=np.float64)
    bootstrap = np.ascontiguousarray(bootstrap, dtype=np.float64)
    T, N = reward.shape
    out = np.zeros_like(reward)
    lib().ca_oracle_nstep_returns(_p(reward), _p(bootstrap), _p(out), T, N, float(gamma))
    return out
