"""CPU tests of the kernel logic: the warp-tile engine header (qt_tile.cuh) compiled for the host and
executed lane by lane (tests/emu) must be bit-exact with the oracle — index maps, twiddle tables,
lazy-reduction bounds with real 32-bit wrap-around — and its shared-memory patterns conflict-free."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import SET_I, SET_III, SET_P_I, SET_P_III, _p

HERE = os.path.dirname(os.path.abspath(__file__))
ALL_SETS = [SET_I, SET_III, SET_P_I, SET_P_III]


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libqt_emu.so")
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu")], check=True)
    L = C.CDLL(so)
    u = C.POINTER(C.c_uint32)
    L.qtemu_polymul.argtypes = [C.c_int, u, u, u, C.c_size_t]
    L.qtemu_forward.argtypes = [C.c_int, u, C.c_size_t]
    L.qtemu_inverse.argtypes = [C.c_int, u, C.c_size_t]
    L.qtemu_forward_natural.argtypes = [C.c_int, u, C.c_size_t]
    L.qtemu_polymul_split.argtypes = [u, u, u, C.c_size_t]
    L.qtemu_inverse_natural.argtypes = [C.c_int, u, C.c_size_t]
    L.qtemu_nussbaumer.argtypes = [C.c_int, u, u, u, C.c_size_t, C.c_int]
    L.qtemu_nussbaumer_recursive.argtypes = [C.c_int, u, u, u, C.c_size_t]
    L.qtemu_inner_lazy.argtypes = [C.c_int, u, u, u, C.c_size_t]
    L.qtemu_row_f64.argtypes = [C.c_int, u, u, u, C.c_size_t]
    L.qtemu_polymul_dq.argtypes = [C.c_int, u, u, u, C.c_size_t, C.POINTER(C.c_double)]
    L.qtemu_dq_remainder.argtypes = [C.c_int, u, u, C.POINTER(C.c_int32), C.c_size_t]
    return L


@pytest.mark.parametrize("s", ALL_SETS)
def test_emulated_kernel_equals_oracle(emu, oracle, s):
    p = oracle.params(s)
    B = 5  # odd: exercises the half-empty last tile of the 2-polynomials-per-warp layout (n=512)
    rng = np.random.default_rng(s)
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    y = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1   # worst case for the lazy bounds
    y[: p.n] = p.q - 1
    x[p.n: 2 * p.n] = 0
    z = np.zeros_like(x)
    assert emu.qtemu_polymul(s, _p(x), _p(y), _p(z), B) == 0
    assert np.array_equal(z, oracle.polymul(s, x, y))
    f = x.copy()
    emu.qtemu_forward(s, _p(f), B)
    assert np.array_equal(f, oracle.forward(s, x))
    g = f.copy()
    emu.qtemu_inverse(s, _p(g), B)
    assert np.array_equal(g, x)


def test_emulated_split_kernel_equals_oracle(emu, oracle):
    """k_polymul_split: n=2048 as two 1024-point halves joined by one level (SET_P_III_H tile)."""
    p = oracle.params(SET_P_III)
    B = 4
    rng = np.random.default_rng(77)
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    y = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1
    y[: p.n] = p.q - 1
    x[p.n: 2 * p.n] = 0
    z = np.zeros_like(x)
    assert emu.qtemu_polymul_split(_p(x), _p(y), _p(z), B) == 0
    assert np.array_equal(z, oracle.polymul(SET_P_III, x, y))


@pytest.mark.parametrize("s", ALL_SETS)
def test_emulated_natural_order_transforms(emu, oracle, s):
    """Natural-order NTT domain (the reference's Stockham ordering, NTT.cu:1162-1191, 1339-1370)."""
    p = oracle.params(s)
    B = 3
    rng = np.random.default_rng(40 + s)
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1
    f = x.copy()
    assert emu.qtemu_forward_natural(s, _p(f), B) == 0
    assert np.array_equal(f, oracle.forward_natural(s, x))
    assert np.array_equal(f, oracle.bitrev_copy(s, oracle.forward(s, x)))
    g = f.copy()
    assert emu.qtemu_inverse_natural(s, _p(g), B) == 0
    assert np.array_equal(g, x)
    assert np.array_equal(oracle.inverse_natural(s, f), x)


@pytest.mark.parametrize("s", ALL_SETS)
def test_shared_memory_patterns_conflict_free(emu, s):
    assert emu.qtemu_bank_conflicts(s) == 1


@pytest.mark.parametrize("s", ALL_SETS)
def test_emulated_nussbaumer_equals_oracle(emu, oracle, s):
    p = oracle.params(s)
    B = 5
    rng = np.random.default_rng(10 + s)
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    y = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1
    y[: p.n] = p.q - 1
    z = np.zeros_like(x)
    assert emu.qtemu_nussbaumer(s, _p(x), _p(y), _p(z), B, 1) == 0          # Z_q mode == NTT product
    assert np.array_equal(z, oracle.polymul(s, x, y))
    xr = rng.integers(0, 2 ** 32, B * p.n, dtype=np.uint32)
    yr = rng.integers(0, 2 ** 32, B * p.n, dtype=np.uint32)
    xr[: p.n] = 1; yr[: p.n] = 1; xr[p.n: 2 * p.n] = 0xFFFFFFFF
    yr[2 * p.n: 3 * p.n] = 0
    yr[2 * p.n + rng.choice(p.n, 40, replace=False)] = 0xFFFFFFFE
    rc = emu.qtemu_nussbaumer(s, _p(xr), _p(yr), _p(z), B, 0)               # ring 2^32-1, bit-exact incl. zeros
    assert rc == 0 and np.array_equal(z, oracle.nussbaumer(p.n, xr, yr))


@pytest.mark.parametrize("s", ALL_SETS)
def test_emulated_recursive_nussbaumer_equals_oracle(emu, oracle, s):
    """k_nussbaumer<SET, Z_q, recursive>: the 2m row products split once more (NussInner, canonical flavour)."""
    p = oracle.params(s)
    B = 6
    rng = np.random.default_rng(20 + s)
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    y = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1; y[: p.n] = p.q - 1
    x[p.n: 2 * p.n] = p.q - 1; y[p.n: 2 * p.n] = np.where(np.arange(p.n) % 2 == 0, p.q - 1, 0)
    x[2 * p.n: 3 * p.n] = 0
    z = np.zeros_like(x)
    assert emu.qtemu_nussbaumer_recursive(s, _p(x), _p(y), _p(z), B) == 0
    assert np.array_equal(z, oracle.polymul(s, x, y))


def _negacyclic_rows(x, y, q):
    """exact negacyclic products of length-32 rows of Python integers, mod q"""
    out = np.zeros_like(x)
    for r in range(x.shape[0]):
        for k in range(32):
            acc = 0
            for j in range(32):
                t = int(x[r, j]) * int(y[r, (k - j) % 32])
                acc += t if j <= k else -t
            out[r, k] = acc % q
    return out


@pytest.mark.parametrize("s", [SET_I, SET_III])
def test_inner_lazy_product_ranges(emu, oracle, s):
    """Signed-lazy NussInner as the warp-resident kernel feeds it: operands at the bounds the kernel
    static_asserts (|x| <= 2^LOGM * q/2 centred for qTESLA-III, 2^LOGM * q otherwise; |y| <= 2^LOGM * q), all
    sign patterns; the result must be the exact product mod q and lie in [-q/2, 3q/2)."""
    p = oracle.params(s)
    q = p.q
    logm = 4 if p.n == 512 else 5
    bx = ((q // 2 + 1) if s == SET_III else q) << logm
    by = q << logm
    rng = np.random.default_rng(30 + s)
    rows = 24
    x = rng.integers(-bx, bx + 1, (rows, 32), dtype=np.int64)
    y = rng.integers(-by, by + 1, (rows, 32), dtype=np.int64)
    x[0, :] = bx; y[0, :] = by                       # all terms of one sign at maximum magnitude
    x[1, :] = -bx; y[1, :] = by
    x[2, :] = bx; y[2, :] = np.where(np.arange(32) % 2 == 0, by, -by)
    x[3, :] = np.where(np.arange(32) % 8 < 4, bx, -bx); y[3, :] = -by
    x[4, :] = 0
    x[5, :] = 0; x[5, 31] = 1; y[5, :] = 0; y[5, 1] = 1   # X^31 * X = -1
    xu = x.astype(np.int32).view(np.uint32).ravel().copy()
    yu = y.astype(np.int32).view(np.uint32).ravel().copy()
    z = np.zeros_like(xu)
    assert emu.qtemu_inner_lazy(s, _p(xu), _p(yu), _p(z), rows) == 0
    zs = z.view(np.int32).astype(np.int64).reshape(rows, 32)
    assert zs.min() >= -(q // 2) - 1 and zs.max() < 3 * q // 2 + 1
    assert np.array_equal(zs % q, _negacyclic_rows(x, y, q))
    assert zs[5, 0] % q == q - 1


@pytest.mark.parametrize("s", [SET_I, SET_III])
def test_fp64_row_product_exact(emu, oracle, s):
    """NussRowF64: operands anywhere in [-q/2, 3q/2) incl. the extremes with every term of one sign (the largest
    partial sums, 72 q^2 < 2^53); the double-precision accumulation must be exact and |z| <= q/2 + 1."""
    q = oracle.params(s).q
    lo, hi = -(q // 2), (3 * q) // 2 - 1
    rng = np.random.default_rng(40 + s)
    rows = 24
    x = rng.integers(lo, hi + 1, (rows, 32), dtype=np.int64)
    y = rng.integers(lo, hi + 1, (rows, 32), dtype=np.int64)
    k = np.arange(32)
    x[0, :] = hi; y[0, :] = hi                       # output 31: 32 terms of +hi^2
    x[1, :] = hi; y[1, :] = np.where(k % 2 == 0, hi, lo)
    x[2, :] = lo; y[2, :] = hi
    x[3, :] = hi; y[3, :] = hi; x[3, 16:] = 0        # partial wrap
    x[4, :] = 0
    x[5, :] = 0; x[5, 31] = 1; y[5, :] = 0; y[5, 1] = 1
    xu = x.astype(np.int32).view(np.uint32).ravel().copy()
    yu = y.astype(np.int32).view(np.uint32).ravel().copy()
    z = np.zeros_like(xu)
    assert emu.qtemu_row_f64(s, _p(xu), _p(yu), _p(z), rows) == 0
    zs = z.view(np.int32).astype(np.int64).reshape(rows, 32)
    assert np.abs(zs).max() <= q // 2 + 1
    assert np.array_equal(zs % q, _negacyclic_rows(x, y, q))


def test_harvey_range_plan_of_p_I(emu):
    """the compile-time range plan of the 29-bit modulus (HarveyPlan, qt_tile.cuh): no value can leave 32 bits
    (every bound <= floor((2^32-1)/q) = 12), and it needs 276 conditional subtractions per product and thread where the
    classic [0,4q) butterflies need 576"""
    out = (C.c_uint32 * 8)()
    assert emu.qtemu_harvey_plan(out) == 0
    ok, fwd, pw, inv, fout_max, rows_out, icols_out, cap = list(out)
    assert ok == 1 and cap == 12
    assert max(fout_max, rows_out, icols_out) <= cap
    assert (fwd, pw, fout_max) == (64, 32, 4)      # forward: 5 free levels, 32 + 32 corrections; one correction per pointwise pair
    assert inv == 84 and rows_out == 11
    assert 2 * fwd + pw + inv + 32 == 276          # + the 32 final subtractions that make the output canonical


@pytest.mark.parametrize("s", [SET_I, SET_III])
def test_fp64_quotient_remainder(emu, oracle, s):
    """the FP64-quotient product of the DQ butterflies (Tile::dq_quot: the DENORMAL double whose bit pattern is {y, 0}
    times w / q is the denormal whose bit pattern is {rint(y w / q), 0}): for every 32-bit unsigned y and every twiddle
    w in [0, q), y w - qe q is congruent to y w and lies within q/2 + 1 of zero — what Tile::ct_dq and the range
    statement in qt_tile.cuh rest on"""
    q = oracle.params(s).q
    rng = np.random.default_rng(50 + s)
    n = 1 << 18
    y = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    w = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    y[:8] = [0, 1, q - 1, q, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF]
    w[:8] = [q - 1, q - 1, q // 2, q // 2 + 1, q - 1, q // 2, 1, 0]
    out = np.zeros(n, dtype=np.int32)
    assert emu.qtemu_dq_remainder(s, _p(y), _p(w), out.ctypes.data_as(C.POINTER(C.c_int32)), n) == 0
    r = out.astype(np.int64)
    assert np.abs(r).max() <= q // 2 + 1
    assert np.array_equal(r % q, (y.astype(object) * w.astype(object) % q).astype(np.int64))


@pytest.mark.parametrize("s", [SET_I, SET_III])
def test_emulated_fp64_quotient_kernel_equals_oracle(emu, oracle, s):
    """k_polymul_dq (fused variant 5) lane by lane: bit-exact with the oracle, the high half of every register pair stays
    zero, and no value leaves the +-40 q window the offset form is built for (worst-case operands included)"""
    p = oracle.params(s)
    B = 5
    rng = np.random.default_rng(60 + s)
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    y = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1
    y[: p.n] = p.q - 1
    x[p.n: 2 * p.n] = 0
    x[2 * p.n: 3 * p.n] = np.where(np.arange(p.n) % 2 == 0, p.q - 1, 0)
    z = np.zeros_like(x)
    stats = (C.c_double * 1)()
    assert emu.qtemu_polymul_dq(s, _p(x), _p(y), _p(z), B, stats) == 0
    assert np.array_equal(z, oracle.polymul(s, x, y))
    assert stats[0] < 40.0
    assert emu.qtemu_polymul_dq(SET_P_I, _p(x), _p(y), _p(z), B, stats) == -4   # the Harvey sets have no such kernel
