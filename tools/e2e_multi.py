#!/usr/bin/env python
"""End-to-end hot path through the IN-PROCESS multi-GPU entry point of the C ABI (qt_multi_polymul_host:
one process, one NUMA-bound host thread + context per GPU, the batch sharded contiguously, no collective)
for 1, 2, 4, ... all visible GPUs.  Weak scaling: 65 536 polynomials of qTESLA-III per GPU, pinned host
arrays, parity of the first / last polynomials of every shard against the CPU oracle.
One JSON line per GPU count."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from qtesla_b200_loader import load  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

qt = load()
o = Oracle()
SET = 1
p = qt.get_params(SET)
PER_GPU = int(os.environ.get("QT_E2E_PER_GPU", "65536"))
ndev = qt.device_count()
reps = 6


def pinned(words):
    ptr = C.c_void_p()
    assert qt.lib().qt_host_alloc(words * 4, C.byref(ptr)) == 0
    return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(words,))


g = 1
while g <= ndev:
    B = PER_GPU * g
    words = B * p.n
    px, x = pinned(words)
    py, y = pinned(words)
    pz, z = pinned(words)
    # cheap host-side fill: one random block tiled (the check below compares against the oracle on what is there)
    rng = np.random.default_rng(g)
    blk = rng.integers(0, p.q, 256 * p.n, dtype=np.uint32)
    for i in range(0, words, blk.size):
        x[i:i + blk.size] = blk
        y[i:i + blk.size] = blk[::-1]
    m = qt.MultiEngine(SET, g)
    for _ in range(2):
        m.polymul_host(x, y, z, B)
    t0 = time.perf_counter()
    for _ in range(reps):
        m.polymul_host(x, y, z, B)
    dt = (time.perf_counter() - t0) / reps
    ok = True
    for s in range(g):  # first and last polynomials of every shard
        lo, hi = B * s // g, B * (s + 1) // g
        for a in (lo, hi - 4):
            sl = slice(a * p.n, (a + 4) * p.n)
            ok &= bool(np.array_equal(z[sl], o.polymul(SET, x[sl].copy(), y[sl].copy())))
    print(json.dumps({"test": "qt_multi_polymul_host", "n_gpus": g, "batch_total": B, "ms_per_call": dt * 1e3,
                      "polymuls_per_s": B / dt, "GBs_host_traffic": 12.0 * words / dt / 1e9, "parity_ok": ok}), flush=True)
    m.close()
    for ptr in (px, py, pz):
        qt.lib().qt_host_free(ptr)
    g *= 2
