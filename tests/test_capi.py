"""CPU tests of the C-ABI library: it loads without a GPU, exports every symbol include/qtesla_b200.h
declares, serves parameters and the constants.h tables, and fails loudly (no CPU fallback) when asked
to compute without a device."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


def declared_symbols():
    syms = set()
    for h in ("qtesla_b200.h", "qtesla_b200_reference_api.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b((?:qt_|test_)[A-Za-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_library_exports_every_declared_symbol(qt):
    L = qt.lib()
    syms = declared_symbols()
    assert len(syms) >= 35 and "test_NTT_GS_CT_nega_gpu" in syms and "qt_polymul" in syms
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/qtesla_b200.h but not exported"
    assert b"sm_100a" in L.qt_version()


def test_params_and_tables_without_gpu(qt, golden):
    p = qt.get_params(qt.SET_III)
    assert (p.n, p.q, p.qinv_neg, p.barrett_mu48, p.omega, p.n_inv) == (1024, 8404993, 4034936831, 33489019, 2893, 8396785)
    g = golden["constants_h_sha256"]
    assert sha(qt.get_table(qt.SET_III, qt.TABLE_BITREV)) == g["bitrev_tbl_gpu"]
    assert sha(qt.get_table(qt.SET_III, qt.TABLE_PHI)) == g["Phi_gpu"]
    assert sha(qt.get_table(qt.SET_III, qt.TABLE_INVPHI)) == g["invPhi_gpu"]
    assert sha(qt.get_table(qt.SET_III, qt.TABLE_TF0)) == g["tf0_gpu"]
    assert sha(qt.get_table(qt.SET_III, qt.TABLE_TI0)) == g["ti0_gpu"]


def test_tables_match_oracle_for_all_sets(qt, oracle):
    for s in range(4):
        t = oracle.tables(s)
        for which, k in enumerate(("bitrev", "Phi", "invPhi", "tf0", "ti0")):
            assert np.array_equal(qt.get_table(s, which), t[k])
        p, po = qt.get_params(s), oracle.params(s)
        for f in ("n", "logn", "q", "psi", "psi_inv", "omega", "omega_inv", "n_inv", "qinv_neg", "barrett_mu48"):
            assert getattr(p, f) == getattr(po, f)


def test_error_behaviour(qt):
    L = qt.lib()
    assert L.qt_get_params(9, C.byref(qt.engine.Params())) == -1
    assert b"parameter set" in L.qt_error_string(-1)
    h = C.c_void_p()
    assert L.qt_create(9, 0, C.byref(h)) == -1


def test_no_cpu_fallback_without_device(qt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qt.QtError):
        qt.Engine(qt.SET_III, 0)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "ntt-gpu-qtesla_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f"{f} mentions the oracle"


def test_multi_handle_and_placement_fail_loudly_without_device(qt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qt.QtError):
        qt.MultiEngine(qt.SET_III, 0)
    x = np.zeros(1024, np.uint32)
    with pytest.raises(qt.QtError):
        qt.polymul_host_multi(qt.SET_III, x, x, 0)
    with pytest.raises(qt.QtError):
        qt.numa.gpu_numa_node(0)


def test_dropin_binary_is_the_reference_main_linked_against_the_library():
    """`make -C oracle dropin` (needs /root/reference at build time): the reference's own main.cu with the
    INTEGRATION.md patch, linked against libqtesla_b200.so — here it must resolve the library and, with no GPU,
    fail with the library's error message instead of computing anything on the CPU."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_main")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_main not built")
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libqtesla_b200.so" in ldd and "not found" not in ldd
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-speedgpu", "3"], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and "no usable CUDA device" in r.stderr
