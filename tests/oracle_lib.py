"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and, when it has been built, of the
unmodified reference CPU functions (oracle/_ref/libqtref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libqtref.so")

SET_I, SET_III, SET_P_I, SET_P_III = 0, 1, 2, 3
SET_NAMES = {SET_I: "qTESLA-I", SET_III: "qTESLA-III", SET_P_I: "qTESLA-p-I", SET_P_III: "qTESLA-p-III"}
VARIANT_GS_CT, VARIANT_GS_GS, VARIANT_CT_CT, VARIANT_STOCKHAM = 0, 1, 2, 3

_u32p = C.POINTER(C.c_uint32)


def _p(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


class Params(C.Structure):
    _fields_ = [("set", C.c_int)] + [
        (k, C.c_uint32)
        for k in ("n", "logn", "q", "psi", "psi_inv", "omega", "omega_inv", "n_inv", "qinv_neg", "barrett_mu48")
    ]


def build_oracle():
    """(Re)build liboracle.so (and _ref when /root/reference exists). Building is not using."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.qto_get_params.argtypes = [C.c_int, C.POINTER(Params)]
        L.qto_tables.argtypes = [C.c_int] + [_u32p] * 5
        for name in ("qto_gs_dif", "qto_ct_dit"):
            getattr(L, name).argtypes = [C.c_int, _u32p, C.c_size_t, _u32p]
            getattr(L, name).restype = None
        L.qto_stockham.argtypes = [C.c_int, _u32p, C.c_size_t, _u32p, _u32p]
        L.qto_bitrev_copy.argtypes = [C.c_int, _u32p, _u32p, C.c_size_t]
        for name in ("qto_ntt_forward", "qto_ntt_inverse", "qto_ntt_forward_natural", "qto_ntt_inverse_natural"):
            getattr(L, name).argtypes = [C.c_int, _u32p, C.c_size_t]
            getattr(L, name).restype = None
        L.qto_pointwise.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t]
        L.qto_polymul.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t, C.c_int]
        L.qto_polymul_omp.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t, C.c_int]
        L.qto_schoolbook.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t]
        L.qto_nussbaumer.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, C.c_size_t]
        L.qto_ring_schoolbook.argtypes = [C.c_uint32, _u32p, _u32p, _u32p, C.c_size_t]
        L.qto_nussbaumer_modq.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t]
        L.qto_fill_xorshift_pair.argtypes = [C.c_uint64, C.c_uint32, _u32p, _u32p, C.c_size_t]
        L.qto_fill_xorshift_pair.restype = C.c_uint64
        L.qto_fill_splitmix.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, _u32p, C.c_size_t]
        L.qto_fill_splitmix.restype = None
        L.qto_bitrev.argtypes = [C.c_uint32, C.c_uint32]
        L.qto_bitrev.restype = C.c_uint32
        L.qto_fast_polymul.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t, C.c_int]
        L.qto_fast_ntt_forward.argtypes = [C.c_int, _u32p, C.c_size_t]
        L.qto_fast_ntt_forward.restype = None

    def params(self, s):
        p = Params()
        assert self.lib.qto_get_params(s, C.byref(p)) == 0
        return p

    def tables(self, s):
        n = self.params(s).n
        t = {k: np.zeros(n, np.uint32) for k in ("bitrev", "Phi", "invPhi", "tf0", "ti0")}
        assert self.lib.qto_tables(s, *[_p(t[k]) for k in ("bitrev", "Phi", "invPhi", "tf0", "ti0")]) == 0
        return t

    def _B(self, s, a):
        n = self.params(s).n
        assert a.size % n == 0
        return a.size // n

    def forward(self, s, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qto_ntt_forward(s, _p(a), self._B(s, a))
        return a

    def inverse(self, s, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qto_ntt_inverse(s, _p(a), self._B(s, a))
        return a

    def forward_natural(self, s, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qto_ntt_forward_natural(s, _p(a), self._B(s, a))
        return a

    def inverse_natural(self, s, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qto_ntt_inverse_natural(s, _p(a), self._B(s, a))
        return a

    def bitrev_copy(self, s, a):
        a = np.ascontiguousarray(a, np.uint32)
        o = np.empty_like(a)
        self.lib.qto_bitrev_copy(s, _p(a), _p(o), self._B(s, a))
        return o

    def pointwise(self, s, a, b):
        a = np.ascontiguousarray(a, np.uint32)
        b = np.ascontiguousarray(b, np.uint32)
        c = np.empty_like(a)
        self.lib.qto_pointwise(s, _p(a), _p(b), _p(c), self._B(s, a))
        return c

    def polymul(self, s, x, y, variant=VARIANT_GS_CT, threads=None):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        if x.size == 0:
            return z
        if threads is None:
            assert self.lib.qto_polymul(s, _p(x), _p(y), _p(z), self._B(s, x), variant) == 0
        else:
            assert self.lib.qto_polymul_omp(s, _p(x), _p(y), _p(z), self._B(s, x), threads) > 0
        return z

    def fast_polymul(self, s, x, y, threads=1):
        """qTESLA-style CPU path (oracle/qt_cpu_fast.c: Montgomery, merged twiddles, lazy ranges)"""
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        if x.size:
            assert self.lib.qto_fast_polymul(s, _p(x), _p(y), _p(z), self._B(s, x), threads) > 0
        return z

    def fast_forward(self, s, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qto_fast_ntt_forward(s, _p(a), self._B(s, a))
        return a

    def schoolbook(self, s, x, y):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        self.lib.qto_schoolbook(s, _p(x), _p(y), _p(z), self._B(s, x))
        return z

    def nussbaumer(self, n, x, y):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        assert self.lib.qto_nussbaumer(n, _p(x), _p(y), _p(z), x.size // n) == 0
        return z

    def ring_schoolbook(self, n, x, y):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        self.lib.qto_ring_schoolbook(n, _p(x), _p(y), _p(z), x.size // n)
        return z

    def nussbaumer_modq(self, s, x, y):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        assert self.lib.qto_nussbaumer_modq(s, _p(x), _p(y), _p(z), self._B(s, x)) == 0
        return z

    def xorshift_pair(self, q, count, state=88172645463325252):
        x = np.empty(count, np.uint32)
        y = np.empty(count, np.uint32)
        st = self.lib.qto_fill_xorshift_pair(state, q, _p(x), _p(y), count)
        return x, y, st

    def splitmix(self, seed, first, q, count):
        a = np.empty(count, np.uint32)
        self.lib.qto_fill_splitmix(seed, first, q, _p(a), count)
        return a

    def max_threads(self):
        return self.lib.qto_max_threads()


class Reference:
    """The unmodified reference CPU functions (n=1024, q=8404993 only)."""

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        L.qtref_q.restype = C.c_uint
        L.qtref_qinv.restype = C.c_uint
        L.qtref_miu.restype = C.c_uint
        L.qtref_table.argtypes = [C.c_int, _u32p]
        L.qtref_forward.argtypes = [_u32p, C.c_size_t]
        L.qtref_inverse.argtypes = [_u32p, C.c_size_t]
        L.qtref_polymul.argtypes = [_u32p, _u32p, _u32p, C.c_size_t, C.c_int, C.c_int]
        L.qtref_nussbaumer.argtypes = [_u32p, _u32p, _u32p, C.c_size_t]
        L.qtref_naive.argtypes = [_u32p, _u32p, _u32p, C.c_uint]
        L.qtref_barrett_cpu.argtypes = [C.c_ulonglong]
        L.qtref_barrett_cpu.restype = C.c_uint
        if hasattr(L, "qtref_gpu_ct_gs"):
            L.qtref_gpu_ct_gs.argtypes = [_u32p, _u32p, _u32p, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self.n = L.qtref_n()
        self.q = L.qtref_q()

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def table(self, which):
        t = np.zeros(self.n, np.uint32)
        assert self.lib.qtref_table(which, _p(t)) == 0
        return t

    def forward(self, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qtref_forward(_p(a), a.size // self.n)
        return a

    def inverse(self, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self.lib.qtref_inverse(_p(a), a.size // self.n)
        return a

    def polymul(self, x, y, variant=0, threads=1):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        self.lib.qtref_polymul(_p(x), _p(y), _p(z), x.size // self.n, variant, threads)
        return z

    def nussbaumer(self, x, y):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        self.lib.qtref_nussbaumer(_p(x), _p(y), _p(z), x.size // self.n)
        return z

    def gpu_ct_gs(self, x, y, reps=3):
        """The reference's own GPU kernels (unmodified), launch order of test_NTT_CT_GS_nega_gpu (NTT.cu:2388-2425).
        Returns (z, kernel_ms, total_ms): kernel-only and copy-inclusive time per batch."""
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        z = np.empty_like(x)
        k, t = C.c_float(0), C.c_float(0)
        rc = self.lib.qtref_gpu_ct_gs(_p(x), _p(y), _p(z), x.size // self.n, reps, C.byref(k), C.byref(t))
        assert rc == 0, f"reference GPU kernels failed: cudaError {rc}"
        return z, k.value, t.value

    def max_threads(self):
        return self.lib.qtref_max_threads()
