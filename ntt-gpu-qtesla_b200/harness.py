"""Python mirror of the reference's harness-level operators (main.cuh:61-70, NTT.cu:2008-2443).

Same names and argument meaning as the reference's `test_NTT_*_nega_gpu` / `test_nussbaumer`
drivers: caller-owned host arrays x, y, z (and the unused X, Y, Z, tf0, ti0, nfg0, nig0, Ni the
reference threads through), result in z (Z for Stockham, NTT.cu:2078).  Like the reference the
drivers overwrite x and y with ones (NTT.cu:2010, 2099, 2183, 2273, 2360) unless `keep_operands`
is set, and report the reference's own metric: wall-clock ms and "Multiplications per second"
including the H2D/D2H copies (NTT.cu:2163-2167).

All five orderings compute the same negacyclic product (the reference's GS-forward GPU variants
omit the psi pre-scale, NTT.cu:969-971 — a defect that is not reproduced); they all map onto the
single fused kernel.  Geometry is run-time (param_set, batch from the array size) instead of the
reference's BATCH/NTTSIZE macros.
"""
import time

import numpy as np

from . import engine as _e

_engines = {}


def _engine(param_set, device):
    key = (param_set, device)
    if key not in _engines:
        _engines[key] = _e.Engine(param_set, device)
    return _engines[key]


def _run(label, x, y, out, param_set, device, keep_operands, verbose, nussbaumer_ring=None):
    eng = _engine(param_set, device)
    if not keep_operands:
        x[...] = 1
        y[...] = 1
    batch = x.size // eng.n
    t0 = time.perf_counter()
    if nussbaumer_ring is None:
        eng.polymul_host(x, y, out, batch)
    else:
        eng.nussbaumer_host(x, y, out, nussbaumer_ring, batch)
    ms = (time.perf_counter() - t0) * 1e3
    if verbose:
        print(f"Performance GPU {label} \n Time\t\t: {ms: .4f} ms. \nThroughput\t: {batch / ms * 1000:.2f} "
              f"Multiplications per second")
    return ms


def test_NTT_Stockham_nega_gpu(x, y, z, X=None, Y=None, Z=None, tf0=None, ti0=None, fg0=0, ig0=0, Ni=0, *,
                               param_set=_e.SET_III, device=0, keep_operands=False, verbose=True):
    out = Z if Z is not None else z  # the reference leaves the Stockham result in Z (NTT.cu:2078)
    return _run("Stockham GPU", x, y, out, param_set, device, keep_operands, verbose)


def test_NTT_GS_CT_nega_gpu(x, y, z, X=None, Y=None, Z=None, tf0=None, ti0=None, nfg0=0, nig0=0, Ni=0, *,
                            param_set=_e.SET_III, device=0, keep_operands=False, verbose=True):
    return _run("GS-CT GPU", x, y, z, param_set, device, keep_operands, verbose)


def test_NTT_CT_CT_nega_gpu(x, y, z, X=None, Y=None, Z=None, tf0=None, ti0=None, nfg0=0, nig0=0, Ni=0, *,
                            param_set=_e.SET_III, device=0, keep_operands=False, verbose=True):
    return _run("CT-CT GPU", x, y, z, param_set, device, keep_operands, verbose)


def test_NTT_GS_GS_nega_gpu(x, y, z, X=None, Y=None, Z=None, tf0=None, ti0=None, nfg0=0, nig0=0, Ni=0, *,
                            param_set=_e.SET_III, device=0, keep_operands=False, verbose=True):
    return _run("GS-GS GPU", x, y, z, param_set, device, keep_operands, verbose)


def test_NTT_CT_GS_nega_gpu(x, y, z, X=None, Y=None, Z=None, tf0=None, ti0=None, nfg0=0, nig0=0, Ni=0, *,
                            param_set=_e.SET_III, device=0, keep_operands=False, verbose=True):
    return _run("CT-GS GPU", x, y, z, param_set, device, keep_operands, verbose)


def test_nussbaumer(x, y, z, X=None, Y=None, Z=None, *, param_set=_e.SET_III, device=0, keep_operands=False,
                    ring=_e.RING_2P32M1, verbose=True):
    """Batched GPU counterpart of test_nussbaumer (NTT.cu:1987-2005), which is CPU-only and handles
    polynomial 0 only; here every polynomial of the batch is multiplied."""
    return _run("Nussbaumer GPU", x, y, z, param_set, device, keep_operands, verbose, nussbaumer_ring=ring)


def init_operand(n):
    """The reference's fixed operand (init_operand, NTT.cu:4-16 with RANDOM=0): x[i]=n/2-i, i<n/2; else 0."""
    x = np.zeros(n, np.uint32)
    x[: n // 2] = n // 2 - np.arange(n // 2, dtype=np.uint32)
    return x
