"""The C-ABI library builds, loads and exports every symbol include/khmm.h declares.
No compute calls (there is no GPU here)."""
import ctypes
import os
import subprocess

import pytest

from kerehmm_b200 import _lib, build


@pytest.fixture(scope="module")
def libpath():
    return build.build_library()


def test_header_parses_all_prototypes():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for must in ("khmm_viterbi_discrete_f64", "khmm_forward_gauss_f32", "khmm_backward_discrete_f64",
                 "khmm_fwdbwd_loglik_gauss_f32", "khmm_estep_discrete_f32", "khmm_xi_accumulate_f32",
                 "khmm_mstep_discrete_f64", "khmm_last_error", "khmm_counts_len"):
        assert must in protos
    ret, argtypes, names = protos["khmm_viterbi_discrete_f64"]
    assert ret is ctypes.c_int and len(argtypes) == 16 and names[0] == "B" and names[-1] == "stream"
    assert argtypes[0] is ctypes.c_int64 and argtypes[2] is ctypes.c_int and argtypes[4] is ctypes.c_void_p


def test_library_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    for name in _lib.parse_header():
        assert hasattr(L, name), "libkhmm.so does not export %s" % name
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "khmm_" in l}
    assert exported == set(_lib.parse_header()), exported ^ set(_lib.parse_header())


def test_library_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if "sm_" in l}
    assert archs == {"sm_100a"}, out


def test_host_side_entry_points_without_a_gpu(libpath):
    L = _lib.lib()
    assert L.khmm_version() >= 100
    N, M = 128, 1024
    assert L.khmm_counts_len(N, M) == N + N * N + N + N * M + N + 3      # SURVEY 8(e) buffer + 3 tail slots
    assert L.khmm_viterbi_workspace_bytes(4, 10, 64) >= 4 * 10 * 64
    assert L.khmm_viterbi_workspace_bytes(4, 10, 300) >= 2 * 4 * 10 * 300
    # argument validation happens before any CUDA call
    with pytest.raises(_lib.KhmmError) as e:
        _lib.call("khmm_forward_discrete_f32", 1, 0, 3, 4, None, None, None, None, 0, None, None, None, None, None)
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.KhmmError) as e:
        _lib.call("khmm_viterbi_discrete_f64", 1, 5, 2000, 4, 1, 1, 1, 1, 0, None, 1, 1 << 30, 1, 2, 1, None)
    assert e.value.code == _lib.ERR_UNSUPPORTED


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "kerehmm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "hmm_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
