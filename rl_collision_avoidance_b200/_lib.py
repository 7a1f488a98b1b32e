"""Loader for libcastep.so (the C-ABI of include/ca_step.h) and its nvcc build recipe.

There is deliberately no fallback: if the shared library is missing or was not built, importing
code gets a RuntimeError telling it to run the build; if no CUDA device is usable, ca_create fails.
"""
import ctypes as C
import os
import subprocess

from . import _abi

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcastep.so")
SOURCES = ["ca_step.cu", "ca_predict.cu"]
OBJ_DIR = os.path.join(CSRC_DIR, "_obj")
HEADERS = [os.path.join(CSRC_DIR, "ca_kernels.cuh"), os.path.join(CSRC_DIR, "ca_step_fast.cuh"), os.path.join(CSRC_DIR, "ca_step_stream.cuh"), os.path.join(CSRC_DIR, "ca_ga3c.cuh"), os.path.join(CSRC_DIR, "ca_scenarios.cuh"), os.path.join(PKG_DIR, "..", "include", "ca_step.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # B200 only
    "-O3", "-lineinfo",
    "-fmad=false",  # float64 parity: never contract a*b+c (explicit __fma_rn where NumPy fuses)
    "-Xcompiler", "-fPIC", "-std=c++17",
]

# every symbol include/ca_step.h declares
EXPORTS = [
    "ca_default_config", "ca_create", "ca_destroy", "ca_set_world_state", "ca_set_reset_state", "ca_reset", "ca_step", "ca_step_host", "ca_step_host_async", "ca_step_host_wait",
    "ca_reset_host", "ca_get_state", "ca_set_dt", "ca_launch_count", "ca_host_alloc", "ca_host_free", "ca_nstep_returns",
    "ca_ga3c_record", "ca_ga3c_episode_stats", "ca_default_scenario_config", "ca_generate_scenarios", "ca_lstm_step", "ca_lstm_cell_forward", "ca_lstm_cell_backward",
    "ca_predictor_pack", "ca_predict", "ca_predict_plan", "ca_predict_rows",
    "ca_strerror", "ca_last_error", "ca_abi_version",
]


def nvcc_path():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a (one object per source, in parallel) and link libcastep.so next to this file
    (in-tree, travels with the repo)."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        path = os.path.join(CSRC_DIR, src)
        deps = [path] + HEADERS
        if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
            cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
            jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), out))
        if verbose:
            print(out)
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), res.stdout))
    return LIB_PATH


_lib = None


def lib():
    """dlopen libcastep.so and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s not found: the CUDA extension is not built. Run `python -c \"import __graft_entry__ as g; g.build()\"` "
            "(needs nvcc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.ca_abi_version.restype = C.c_int
    if L.ca_abi_version() != _abi.CA_ABI_VERSION:
        raise RuntimeError("libcastep.so ABI version %d != python mirror %d; rebuild" % (L.ca_abi_version(), _abi.CA_ABI_VERSION))
    L.ca_default_config.argtypes = [C.POINTER(_abi.CaConfig), i32, i32]
    L.ca_create.argtypes = [C.POINTER(_abi.CaConfig), C.POINTER(vp)]
    L.ca_destroy.argtypes = [vp]
    L.ca_set_world_state.argtypes = [vp, vp, vp, C.c_int, vp]
    L.ca_set_reset_state.argtypes = [vp, vp, vp, C.c_int, vp]
    L.ca_reset.argtypes = [vp, vp, vp, vp, vp]
    L.ca_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ca_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.ca_step_host_async.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.ca_step_host_wait.argtypes = [vp]
    L.ca_reset_host.argtypes = [vp, vp, vp, vp]
    L.ca_get_state.argtypes = [vp, vp, C.c_int, vp]
    L.ca_set_dt.argtypes = [vp, C.c_double]
    L.ca_launch_count.argtypes = [vp, C.POINTER(i64)]
    L.ca_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.ca_host_free.argtypes = [vp]
    L.ca_nstep_returns.argtypes = [vp, vp, vp, i32, i32, C.c_float, C.c_int, vp]
    L.ca_ga3c_record.argtypes = [C.POINTER(_abi.CaGa3cBuffers), i64, i32, i32, i32, i32, i32, C.c_float, vp, vp, vp, vp, vp,
                                 C.c_int, vp]
    L.ca_ga3c_episode_stats.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, C.c_int, vp]
    L.ca_default_scenario_config.argtypes = [C.POINTER(_abi.CaScenarioConfig), i32]
    L.ca_generate_scenarios.argtypes = [vp, C.POINTER(_abi.CaScenarioConfig), C.c_uint64, C.c_int, vp]
    L.ca_lstm_step.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, C.c_int, vp]
    L.ca_lstm_cell_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, C.c_int, vp]
    L.ca_lstm_cell_backward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, i32, C.c_int, vp]
    L.ca_predictor_pack.argtypes = [C.POINTER(_abi.CaPredictorParams), vp, C.c_int, vp]
    L.ca_predict.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32, C.c_float, C.c_uint64, C.c_uint64, vp, C.c_int, vp]
    L.ca_predict_plan.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, C.c_int, vp]
    L.ca_predict_rows.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, C.c_float, C.c_uint64, C.c_uint64, vp, C.c_int, vp]
    L.ca_strerror.argtypes = [C.c_int]
    L.ca_strerror.restype = C.c_char_p
    L.ca_last_error.restype = C.c_char_p
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("ca_strerror", "ca_last_error"):
            fn.restype = C.c_int
    _lib = L
    return L


class CaError(RuntimeError):
    def __init__(self, code, where):
        L = lib()
        self.code = code
        msg = "%s: %s (%d): %s" % (where, L.ca_strerror(code).decode(), code, L.ca_last_error().decode())
        RuntimeError.__init__(self, msg)


def check(code, where):
    if code != _abi.CA_OK:
        if code == _abi.CA_ERR_INVALID_ARG:
            raise ValueError(str(CaError(code, where)))
        raise CaError(code, where)
