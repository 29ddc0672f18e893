"""ctypes binding of libqtesla_b200.so (include/qtesla_b200.h).

PyTorch is used by callers only for device memory and streams; every signature here takes raw
device/host addresses (ints), numpy arrays (host) or torch tensors (device, via .data_ptr()).
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QT_LIB_PATH") or os.path.join(PKG_DIR, "libqtesla_b200.so")  # override: A/B builds

SET_I, SET_III, SET_P_I, SET_P_III = 0, 1, 2, 3
SET_NAMES = {SET_I: "qTESLA-I", SET_III: "qTESLA-III", SET_P_I: "qTESLA-p-I", SET_P_III: "qTESLA-p-III"}
TABLE_BITREV, TABLE_PHI, TABLE_INVPHI, TABLE_TF0, TABLE_TI0 = range(5)
RING_2P32M1, RING_MODQ, RING_2P32M1_LIFT_Q = 0, 1, 2

_vp = C.c_void_p
_sz = C.c_size_t


class QtError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("set", C.c_int)] + [(k, C.c_uint32) for k in (
        "n", "logn", "q", "psi", "psi_inv", "omega", "omega_inv", "n_inv", "qinv_neg", "barrett_mu48")]


_SIGNATURES = {
    "qt_version": (C.c_char_p, []),
    "qt_error_string": (C.c_char_p, [C.c_int]),
    "qt_get_params": (C.c_int, [C.c_int, C.POINTER(Params)]),
    "qt_get_table": (C.c_int, [C.c_int, C.c_int, _vp]),
    "qt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qt_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "qt_destroy": (C.c_int, [_vp]),
    "qt_set_stream": (C.c_int, [_vp, _vp]),
    "qt_synchronize": (C.c_int, [_vp]),
    "qt_set_fused_variant": (C.c_int, [_vp, C.c_int]),
    "qt_set_nussbaumer_variant": (C.c_int, [_vp, C.c_int]),
    "qt_set_launch_overlap": (C.c_int, [_vp, C.c_int]),
    "qt_device_malloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "qt_device_free": (C.c_int, [_vp, _vp]),
    "qt_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "qt_host_free": (C.c_int, [_vp]),
    "qt_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, _sz]),
    "qt_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, _sz]),
    "qt_ntt_forward": (C.c_int, [_vp, _vp, _sz]),
    "qt_ntt_inverse": (C.c_int, [_vp, _vp, _sz]),
    "qt_ntt_forward_natural": (C.c_int, [_vp, _vp, _sz]),
    "qt_ntt_inverse_natural": (C.c_int, [_vp, _vp, _sz]),
    "qt_pointwise": (C.c_int, [_vp, _vp, _vp, _vp, _sz]),
    "qt_polymul": (C.c_int, [_vp, _vp, _vp, _vp, _sz]),
    "qt_bitrev_copy": (C.c_int, [_vp, _vp, _vp, _sz]),
    "qt_polymul_ntt": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _sz]),
    "qt_nussbaumer": (C.c_int, [_vp, _vp, _vp, _vp, _sz, C.c_int]),
    "qt_fill_uniform": (C.c_int, [_vp, _vp, _sz, C.c_uint64, C.c_uint64]),
    "qt_polymul_host": (C.c_int, [_vp, _vp, _vp, _vp, _sz]),
    "qt_polymul_host_multi": (C.c_int, [C.c_int, _vp, _vp, _vp, _sz, C.c_int]),
    "qt_nussbaumer_host": (C.c_int, [_vp, _vp, _vp, _vp, _sz, C.c_int]),
    "qt_shutdown": (C.c_int, []),
    "qt_graph_begin": (C.c_int, [_vp]),
    "qt_graph_end": (C.c_int, [_vp, C.POINTER(_vp)]),
    "qt_graph_launch": (C.c_int, [_vp, _vp]),
    "qt_graph_destroy": (C.c_int, [_vp]),
    "qt_graph_kernel_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "qt_multi_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "qt_multi_destroy": (C.c_int, [_vp]),
    "qt_multi_gpus": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "qt_multi_polymul_host": (C.c_int, [_vp, _vp, _vp, _vp, _sz]),
    "qt_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "qt_kernel_info": (C.c_int, [_vp] + [C.POINTER(C.c_int)] * 5),
    "qt_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, _sz]),
    "qt_device_numa_node": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "qt_bind_thread_to_device": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
}

_lib = None


def lib():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QtError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise QtError(f"qtesla_b200 error {rc}: {lib().qt_error_string(rc).decode()}")


def _addr(a):
    """device/host address of an int, torch tensor or numpy array"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.dtype in (np.uint32, np.int32) and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous() and a.element_size() == 4
        return a.data_ptr()
    raise TypeError(type(a))


def get_params(param_set):
    p = Params()
    _check(lib().qt_get_params(param_set, C.byref(p)))
    return p


def get_table(param_set, which):
    n = get_params(param_set).n
    out = np.empty(n, np.uint32)
    _check(lib().qt_get_table(param_set, which, out.ctypes.data))
    return out


def device_count():
    n = C.c_int(0)
    rc = lib().qt_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def polymul_host_multi(param_set, x, y, ngpus=0):
    """z = x*y for host arrays, batch sharded contiguously over ngpus devices (no collective)."""
    p = get_params(param_set)
    x = np.ascontiguousarray(x, np.uint32)
    y = np.ascontiguousarray(y, np.uint32)
    assert x.size == y.size and x.size % p.n == 0
    z = np.empty_like(x)
    _check(lib().qt_polymul_host_multi(param_set, _addr(x), _addr(y), _addr(z), x.size // p.n, ngpus))
    return z


class MultiEngine:
    """qt_multi: one context + one NUMA-bound host thread per GPU, the batch sharded contiguously (no collective)."""

    def __init__(self, param_set=SET_III, ngpus=0):
        self._h = _vp()
        self.params = get_params(param_set)
        self.n, self.q = self.params.n, self.params.q
        _check(lib().qt_multi_create(param_set, ngpus, C.byref(self._h)))
        k = C.c_int(0)
        _check(lib().qt_multi_gpus(self._h, C.byref(k)))
        self.ngpus = k.value

    def polymul_host(self, x, y, z=None, batch=None):
        if z is None:
            z = np.empty_like(x)
        if batch is None:
            assert x.size % self.n == 0
            batch = x.size // self.n
        _check(lib().qt_multi_polymul_host(self._h, _addr(x), _addr(y), _addr(z), batch))
        return z

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().qt_multi_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One context = one GPU + one parameter set (qt_create).  Not thread-safe."""

    def __init__(self, param_set=SET_III, device=0):
        self._h = _vp()
        self.param_set = param_set
        self.device = device
        self.params = get_params(param_set)
        self.n, self.q = self.params.n, self.params.q
        _check(lib().qt_create(param_set, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().qt_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: the library may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing
    def set_stream(self, cuda_stream):
        _check(lib().qt_set_stream(self._h, cuda_stream))

    def set_fused_variant(self, variant):
        """0 automatic, 1 direct coalesced loads, 2 TMA bulk copies staged through shared memory, 3 / 4 n=2048 as two halves
        (one warp / a pair of warps), 5 FP64-quotient butterflies (qTESLA-I, -III)"""
        _check(lib().qt_set_fused_variant(self._h, variant))

    def set_launch_overlap(self, mode):
        """programmatic dependent launch of the TMA-staged kernels: 0 automatic (non-blocking streams), 1 never, 2 always"""
        _check(lib().qt_set_launch_overlap(self._h, mode))

    def set_nussbaumer_variant(self, variant):
        """row products of the Z_q Nussbaumer kernels: 0 automatic, 1 schoolbook, 2 recursive (split once more),
        3 schoolbook on the FP64 pipe (q < 2^25)"""
        _check(lib().qt_set_nussbaumer_variant(self._h, variant))

    def synchronize(self):
        _check(lib().qt_synchronize(self._h))

    def launch_count(self):
        v = C.c_uint64(0)
        _check(lib().qt_launch_count(self._h, C.byref(v)))
        return v.value

    def kernel_info(self):
        v = [C.c_int(0) for _ in range(5)]
        _check(lib().qt_kernel_info(self._h, *[C.byref(i) for i in v]))
        return dict(zip(("grid", "block", "smem_bytes", "blocks_per_sm", "num_sms"), [i.value for i in v]))

    def _batch(self, a, batch):
        if batch is not None:
            return batch
        sz = a.size if isinstance(a, np.ndarray) else a.numel()
        assert sz % self.n == 0, "array must hold whole polynomials"
        return sz // self.n

    # -- CUDA graphs: record a fixed launch sequence once, replay it with one host call (launch-bound batches)
    def graph_begin(self):
        _check(lib().qt_graph_begin(self._h))

    def graph_end(self):
        g = _vp()
        _check(lib().qt_graph_end(self._h, C.byref(g)))
        return g

    def graph_launch(self, g):
        _check(lib().qt_graph_launch(self._h, g))

    def graph_destroy(self, g):
        _check(lib().qt_graph_destroy(g))

    def graph_kernel_count(self, g):
        v = C.c_uint64(0)
        _check(lib().qt_graph_kernel_count(g, C.byref(v)))
        return v.value

    # -- device-pointer operators (asynchronous on the context stream)
    def ntt_forward(self, d_a, batch=None):
        _check(lib().qt_ntt_forward(self._h, _addr(d_a), self._batch(d_a, batch)))

    def ntt_inverse(self, d_a, batch=None):
        _check(lib().qt_ntt_inverse(self._h, _addr(d_a), self._batch(d_a, batch)))

    def ntt_forward_natural(self, d_a, batch=None):
        _check(lib().qt_ntt_forward_natural(self._h, _addr(d_a), self._batch(d_a, batch)))

    def ntt_inverse_natural(self, d_a, batch=None):
        _check(lib().qt_ntt_inverse_natural(self._h, _addr(d_a), self._batch(d_a, batch)))

    def pointwise(self, d_a, d_b, d_c, batch=None):
        _check(lib().qt_pointwise(self._h, _addr(d_a), _addr(d_b), _addr(d_c), self._batch(d_a, batch)))

    def polymul(self, d_x, d_y, d_z, batch=None):
        _check(lib().qt_polymul(self._h, _addr(d_x), _addr(d_y), _addr(d_z), self._batch(d_x, batch)))

    def polymul_ntt(self, d_a_hat, d_y, d_z, broadcast=True, batch=None):
        """z = a*y with NTT(a) given (broadcast: one a_hat for the whole batch)"""
        _check(lib().qt_polymul_ntt(self._h, _addr(d_a_hat), 1 if broadcast else 0, _addr(d_y), _addr(d_z),
                                    self._batch(d_y, batch)))

    def bitrev_copy(self, d_in, d_out, batch=None):
        _check(lib().qt_bitrev_copy(self._h, _addr(d_in), _addr(d_out), self._batch(d_in, batch)))

    def nussbaumer(self, d_x, d_y, d_z, ring=RING_2P32M1, batch=None):
        _check(lib().qt_nussbaumer(self._h, _addr(d_x), _addr(d_y), _addr(d_z), self._batch(d_x, batch), ring))

    def fill_uniform(self, d_a, seed, first_index=0, count=None):
        if count is None:
            count = d_a.numel()
        _check(lib().qt_fill_uniform(self._h, _addr(d_a), count, seed, first_index))

    # -- host-pointer operators (synchronous; what the reference's test_*_nega_gpu drivers do)
    def polymul_host(self, x, y, z=None, batch=None):
        if z is None:
            z = np.empty_like(x)
        _check(lib().qt_polymul_host(self._h, _addr(x), _addr(y), _addr(z), self._batch(x, batch)))
        return z

    def nussbaumer_host(self, x, y, z=None, ring=RING_2P32M1, batch=None):
        if z is None:
            z = np.empty_like(x)
        _check(lib().qt_nussbaumer_host(self._h, _addr(x), _addr(y), _addr(z), self._batch(x, batch), ring))
        return z

    # -- small conveniences for tests (host numpy in, host numpy out, through device memory)
    def _roundtrip(self, fn, *host_arrays):
        import torch
        dev = torch.device("cuda", self.device)
        ts = [torch.from_numpy(np.ascontiguousarray(a, np.uint32).view(np.int32)).to(dev) for a in host_arrays]
        torch.cuda.current_stream(dev).synchronize()  # the engine's stream is not ordered against torch's
        out = fn(*ts)
        self.synchronize()
        torch.cuda.synchronize(dev)
        return out.cpu().numpy().view(np.uint32)

    def forward_np(self, a):
        return self._roundtrip(lambda t: (self.ntt_forward(t), t)[1], a)

    def inverse_np(self, a):
        return self._roundtrip(lambda t: (self.ntt_inverse(t), t)[1], a)

    def forward_natural_np(self, a):
        return self._roundtrip(lambda t: (self.ntt_forward_natural(t), t)[1], a)

    def inverse_natural_np(self, a):
        return self._roundtrip(lambda t: (self.ntt_inverse_natural(t), t)[1], a)

    def polymul_np(self, x, y):
        import torch
        return self._roundtrip(lambda tx, ty: (lambda tz: (self.polymul(tx, ty, tz), tz)[1])(torch.empty_like(tx)), x, y)

    def pointwise_np(self, a, b):
        import torch
        return self._roundtrip(lambda ta, tb: (lambda tc: (self.pointwise(ta, tb, tc), tc)[1])(torch.empty_like(ta)), a, b)
