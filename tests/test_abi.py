"""CPU: the C-ABI library builds/loads and exports every symbol include/ca_step.h declares; the ctypes
mirror of struct ca_config matches the C layout; no compute calls are made (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from rl_collision_avoidance_b200 import _abi, _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "ca_step.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(ca_[a-z_0-9]+)\s*\(", src, flags=re.M)))


@pytest.fixture(scope="module")
def built_lib():
    _lib.build()
    return _lib.lib()


def test_header_symbols_all_exported(built_lib):
    declared = _declared_symbols()
    assert len(declared) >= 14
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(built_lib, name), "libcastep.so does not export %s" % name


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_config_struct_layout_matches_c(tmp_path):
    fields = [f[0] for f in _abi.CaConfig._fields_]
    prog = ["#include <stdio.h>", "#include <stddef.h>", '#include "ca_step.h"', "int main(void){",
            'printf("%zu\\n", sizeof(ca_config));']
    prog += ['printf("%%zu\\n", offsetof(ca_config, %s));' % f for f in fields]
    prog += ['printf("%d %d %d\\n", CA_INIT_STRIDE, CA_STATE_STRIDE, CA_OBS_LEN(3));', "return 0;}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    assert int(lines[0]) == C.sizeof(_abi.CaConfig)
    for k, f in enumerate(fields):
        assert int(lines[1 + k]) == getattr(_abi.CaConfig, f).offset, f
    a, b, c = (int(x) for x in lines[1 + len(fields)].split())
    assert (a, b, c) == (_abi.INIT_STRIDE, _abi.STATE_STRIDE, _abi.obs_len(3))


def test_other_struct_layouts_and_constants_match_c(tmp_path):
    """ca_ga3c_buffers, ca_predictor_params, ca_scenario_config and the predictor constants: the ctypes mirrors in _abi.py
    have the sizes, field offsets and values the C header gives them (compiled with gcc, no GPU needed)."""
    structs = [("ca_ga3c_buffers", _abi.CaGa3cBuffers), ("ca_predictor_params", _abi.CaPredictorParams),
               ("ca_scenario_config", _abi.CaScenarioConfig)]
    prog = ["#include <stdio.h>", "#include <stddef.h>", '#include "ca_step.h"', "int main(void){"]
    for cname, mirror in structs:
        prog.append('printf("%%zu\\n", sizeof(%s));' % cname)
        prog += ['printf("%%zu\\n", offsetof(%s, %s));' % (cname, f[0]) for f in mirror._fields_]
    prog += ['printf("%d %d %d\\n", (int)CA_PREDICTOR_BLOB_BYTES, (int)CA_PREDICT_PLAN_COUNTERS, (int)CA_MAX_AGENTS);', "return 0;}"]
    src = tmp_path / "probe2.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "probe2"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(src)])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    k = 0
    for cname, mirror in structs:
        assert int(lines[k]) == C.sizeof(mirror), cname
        k += 1
        for f in mirror._fields_:
            assert int(lines[k]) == getattr(mirror, f[0]).offset, (cname, f[0])
            k += 1
    blob, counters, max_agents = (int(x) for x in lines[k].split())
    assert blob == _abi.CA_PREDICTOR_BLOB_BYTES and counters == _abi.CA_PREDICT_PLAN_COUNTERS and max_agents == 32


def test_default_config_matches_python_mirror(built_lib):
    c1 = _abi.CaConfig()
    assert built_lib.ca_default_config(C.byref(c1), 7, 4) == 0
    c2 = _abi.default_config(7, 4)
    for f, _ in _abi.CaConfig._fields_:
        assert getattr(c1, f) == getattr(c2, f), f


def test_strerror_and_argument_validation(built_lib):
    assert built_lib.ca_strerror(0) == b"ok"
    assert b"invalid" in built_lib.ca_strerror(_abi.CA_ERR_INVALID_ARG)
    h = C.c_void_p()
    bad = _abi.default_config(8, 4)
    bad.abi_version = 99
    assert built_lib.ca_create(C.byref(bad), C.byref(h)) == _abi.CA_ERR_INVALID_ARG
    assert b"abi_version" in built_lib.ca_last_error()
    bad = _abi.default_config(8, 40)
    assert built_lib.ca_create(C.byref(bad), C.byref(h)) == _abi.CA_ERR_INVALID_ARG
    bad = _abi.default_config(0, 4)
    assert built_lib.ca_create(C.byref(bad), C.byref(h)) == _abi.CA_ERR_INVALID_ARG
    assert not h


def test_no_cpu_fallback_without_gpu(built_lib):
    """On a box without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    cfg = _abi.default_config(8, 4)
    assert built_lib.ca_create(C.byref(cfg), C.byref(h)) == _abi.CA_ERR_CUDA
    from rl_collision_avoidance_b200.vec_env import HostVecEnv, VecCollisionAvoidanceEnv
    with pytest.raises(RuntimeError):
        HostVecEnv(cfg)
    with pytest.raises(RuntimeError):
        VecCollisionAvoidanceEnv(cfg)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(REPO, "rl_collision_avoidance_b200")
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), fn
                assert "ca_oracle_" not in text and "libca_oracle" not in text, fn
