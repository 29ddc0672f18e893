"""CPU tests (-m "not gpu"): pin the oracle against the reference's golden vectors, against the
unmodified reference CPU code (oracle/_ref, when built) and against mathematics (schoolbook)."""
import hashlib
import os

import numpy as np
import pytest

from oracle_lib import (SET_I, SET_III, SET_P_I, SET_P_III, SET_NAMES, VARIANT_GS_CT, VARIANT_GS_GS,
                        VARIANT_CT_CT, VARIANT_STOCKHAM)

ALL_SETS = [SET_I, SET_III, SET_P_I, SET_P_III]
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


def test_params_match_reference_macros(oracle):
    p = oracle.params(SET_III)  # main.cuh:14-21, main.cu:26
    assert (p.n, p.q, p.qinv_neg, p.barrett_mu48) == (1024, 8404993, 4034936831, 33489019)
    assert (p.omega, p.omega_inv, p.n_inv, p.psi, p.psi_inv) == (2893, 7562460, 8396785, 2083362, 5907167)
    # derived constants of the sets the reference has no code for
    exp = {SET_I: (512, 4205569, 3353664, 4197355, 3098553343), SET_P_I: (1024, 343576577, 249751876, 343241053, 2205847551),
           SET_P_III: (2048, 856145921, 89095543, 855727881, 587710463)}
    for s, (n, q, psi, ninv, qinv) in exp.items():
        p = oracle.params(s)
        assert (p.n, p.q, p.psi, p.n_inv, p.qinv_neg) == (n, q, psi, ninv, qinv)
        assert pow(p.psi, p.n, p.q) == p.q - 1 and (p.q * (2 ** 32 - p.qinv_neg)) % 2 ** 32 == 1


def test_tables_equal_constants_h(oracle, golden):
    t = oracle.tables(SET_III)
    g = golden["constants_h_sha256"]
    assert sha(t["bitrev"]) == g["bitrev_tbl"] == g["bitrev_tbl_gpu"]
    assert sha(t["Phi"]) == g["Phi"] == g["Phi_gpu"]
    assert sha(t["invPhi"]) == g["invPhi"] == g["invPhi_gpu"]
    assert sha(t["tf0"]) == g["tf0_gpu"]
    assert sha(t["ti0"]) == g["ti0_gpu"]


def test_tables_equal_linked_reference(oracle, reference):
    t = oracle.tables(SET_III)
    for i, k in enumerate(("bitrev", "Phi", "invPhi", "tf0", "ti0")):
        assert np.array_equal(t[k], reference.table(i)), k


def test_random_golden_III(oracle, golden):
    g = golden["III_random_b2"]
    d = np.load(os.path.join(HERE, "golden", "golden_III_b2.npz"))
    x, y, _ = oracle.xorshift_pair(8404993, 2048)
    assert np.array_equal(x, d["x"]) and np.array_equal(y, d["y"])
    assert list(x[:4]) == g["x_first4"]
    for v in (VARIANT_GS_CT, VARIANT_GS_GS, VARIANT_CT_CT, VARIANT_STOCKHAM):
        z = oracle.polymul(SET_III, x, y, v)
        assert np.array_equal(z, d["z"]) and sha(z) == g["z_sha256"]
    f = oracle.forward(SET_III, x)
    assert np.array_equal(f, d["fwd_x"]) and sha(f) == g["fwd_sha256"] and list(f[:4]) == g["fwd_first4"]
    assert np.array_equal(oracle.inverse(SET_III, f), x)
    assert np.array_equal(oracle.schoolbook(SET_III, x, y), d["z"])


def test_all_ones_kat(oracle, golden):
    # the reference's own fixture: x = y = 1 (NTT.cu:1822, 2099) -> z[k] = (2k+2-n) mod q
    for s in ALL_SETS:
        p = oracle.params(s)
        ones = np.ones(2 * p.n, np.uint32)
        z = oracle.polymul(s, ones, ones)
        exp = ((2 * np.arange(p.n, dtype=np.int64) + 2 - p.n) % p.q).astype(np.uint32)
        assert np.array_equal(z[: p.n], exp) and np.array_equal(z[p.n:], exp)
    assert sha(oracle.polymul(SET_III, np.ones(2048, np.uint32), np.ones(2048, np.uint32))) == golden["III_all_ones"]["z_sha256"]


def test_ramp_roundtrip_and_golden(oracle, golden):
    # init_operand (NTT.cu:11,15): INTT(NTT(x)) == x ("Identical.", NTT.cu:1522-1530)
    for s in ALL_SETS:
        p = oracle.params(s)
        x = np.zeros(2 * p.n, np.uint32)
        for b in range(2):
            x[b * p.n: b * p.n + p.n // 2] = p.n // 2 - np.arange(p.n // 2)
        f = oracle.forward(s, x)
        assert np.array_equal(oracle.inverse(s, f), x)
        assert np.array_equal(oracle.inverse_natural(s, oracle.forward_natural(s, x)), x)
        if s == SET_III:
            assert sha(f) == golden["III_ramp_forward_sha256"]
            assert sha(oracle.polymul(s, x, x)) == golden["III_ramp_square_sha256"]


@pytest.mark.parametrize("s", ALL_SETS)
def test_variants_equal_schoolbook(oracle, golden, s):
    p = oracle.params(s)
    x, y, _ = oracle.xorshift_pair(p.q, p.n)
    z = oracle.schoolbook(s, x, y)
    if s != SET_III:
        g = golden[SET_NAMES[s] + "_random_b1"]
        assert sha(z) == g["z_sha256"] and list(z[:4]) == g["z_first4"]
    for v in (VARIANT_GS_CT, VARIANT_GS_GS, VARIANT_CT_CT, VARIANT_STOCKHAM):
        assert np.array_equal(oracle.polymul(s, x, y, v), z)
    assert np.array_equal(oracle.polymul(s, x, y, threads=2), z)
    assert np.array_equal(oracle.nussbaumer_modq(s, x, y), z)
    # NTT-domain orderings: natural (Stockham) == bit-reversal of the GS output
    assert np.array_equal(oracle.forward_natural(s, x), oracle.bitrev_copy(s, oracle.forward(s, x)))


@pytest.mark.parametrize("s", ALL_SETS)
def test_structural_edges(oracle, s):
    p = oracle.params(s)
    n, q = p.n, p.q
    rng = np.random.default_rng(7)
    y = rng.integers(0, q, n, dtype=np.uint32)
    one = np.zeros(n, np.uint32); one[0] = 1
    assert np.array_equal(oracle.polymul(s, one, y), y)                      # x = 1 -> z = y
    xm = np.zeros(n, np.uint32); xm[n - 1] = 1
    xx = np.zeros(n, np.uint32); xx[1] = 1
    z = oracle.polymul(s, xm, xx)                                             # X^(n-1) * X = -1
    assert z[0] == q - 1 and not z[1:].any()
    assert not oracle.polymul(s, np.zeros(n, np.uint32), y).any()             # zero
    top = np.full(n, q - 1, np.uint32)                                        # all q-1 == all-ones product
    assert np.array_equal(oracle.polymul(s, top, top), oracle.polymul(s, np.ones(n, np.uint32), np.ones(n, np.uint32)))
    # linearity: (a+b)*y == a*y + b*y
    a = rng.integers(0, q, n, dtype=np.uint32); b = rng.integers(0, q, n, dtype=np.uint32)
    ab = ((a.astype(np.uint64) + b) % q).astype(np.uint32)
    lhs = oracle.polymul(s, ab, y)
    rhs = ((oracle.polymul(s, a, y).astype(np.uint64) + oracle.polymul(s, b, y)) % q).astype(np.uint32)
    assert np.array_equal(lhs, rhs)


def test_oracle_equals_reference_cpu(oracle, reference):
    # the restatement against the UNMODIFIED reference functions on fresh random inputs, incl. odd batch
    rng = np.random.default_rng(11)
    for B in (1, 2, 5):
        x = rng.integers(0, 8404993, B * 1024, dtype=np.uint32)
        y = rng.integers(0, 8404993, B * 1024, dtype=np.uint32)
        z = oracle.polymul(SET_III, x, y)
        for v in (0, 1, 2):  # Stockham CPU path prints from barrett_red_cpu; covered by the golden file
            assert np.array_equal(reference.polymul(x, y, v, 1), z)
        assert np.array_equal(reference.forward(x), oracle.forward(SET_III, x))
        assert np.array_equal(reference.inverse(x), oracle.inverse(SET_III, x))


def test_nussbaumer_ring(oracle, golden):
    x, y, _ = oracle.xorshift_pair(8404993, 1024)
    z = oracle.nussbaumer(1024, x, y)
    g = golden["III_nussbaumer_b1"]
    assert sha(z) == g["z_sha256"] and list(z[:4]) == g["z_first4"]
    ones = np.ones(1024, np.uint32)
    zo = oracle.nussbaumer(1024, ones, ones)
    go = golden["III_nussbaumer_all_ones"]
    assert sha(zo) == go["z_sha256"] and list(zo[:3]) == go["z_first3"] == [4294966273, 4294966275, 4294966277]
    norm = lambda a: np.where(a == 0xFFFFFFFF, 0, a).astype(np.uint32)
    assert np.array_equal(norm(z), norm(oracle.ring_schoolbook(1024, x, y)))
    for s in (SET_I, SET_P_I, SET_P_III):
        p = oracle.params(s)
        xx, yy, _ = oracle.xorshift_pair(p.q, p.n)
        assert sha(norm(oracle.nussbaumer(p.n, xx, yy))) == golden[SET_NAMES[s] + "_ring_schoolbook_b1_sha256"]


def test_nussbaumer_equals_reference(oracle, reference):
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2 ** 32, 3 * 1024, dtype=np.uint32)   # the ring accepts any 32-bit word
    y = rng.integers(0, 2 ** 32, 3 * 1024, dtype=np.uint32)
    x[:1024] = 0xFFFFFFFF                                       # non-normalised zeros
    assert np.array_equal(oracle.nussbaumer(1024, x, y), reference.nussbaumer(x, y))
    # sparse ternary second operand (qTESLA-shaped): zeros are frequent in the intermediates
    t = np.zeros(1024, np.uint32); t[rng.choice(1024, 48, replace=False)] = np.where(rng.random(48) < 0.5, 1, 0xFFFFFFFE)
    assert np.array_equal(oracle.nussbaumer(1024, x[1024:2048], t), reference.nussbaumer(x[1024:2048], t))


def test_splitmix_stream(oracle):
    a = oracle.splitmix(1, 0, 8404993, 16)
    b = oracle.splitmix(1, 8, 8404993, 8)
    assert np.array_equal(a[8:], b) and a.max() < 8404993
    # value pinned so that the device generator can be checked against it
    assert int(oracle.splitmix(0, 0, 2 ** 32 - 1, 1)[0]) == (0xE220A8397B1DCDAF % (2 ** 32 - 1))


# ---- qTESLA-style Montgomery / merged-twiddle CPU baseline (oracle/qt_cpu_fast.c) ----------
@pytest.mark.parametrize("s", ALL_SETS)
def test_fast_cpu_path_equals_port_and_schoolbook(oracle, s):
    """the restatement of the qTESLA C poly_ntt/poly_mul is pinned to the port (itself pinned to the reference)
    and to the O(n^2) schoolbook; edge operands: all q-1, all 0, x = 1, x = X^(n-1) * y = X"""
    p = oracle.params(s)
    rng = np.random.default_rng(900 + s)
    B = 12
    x = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    y = rng.integers(0, p.q, B * p.n, dtype=np.uint32)
    x[: p.n] = p.q - 1
    y[: p.n] = p.q - 1
    x[p.n: 2 * p.n] = 0
    x[2 * p.n: 3 * p.n] = 0
    x[2 * p.n] = 1
    x[3 * p.n: 4 * p.n] = 0
    x[4 * p.n - 1] = 1
    y[3 * p.n: 4 * p.n] = 0
    y[3 * p.n + 1] = 1
    z = oracle.fast_polymul(s, x, y, threads=2)
    assert np.array_equal(z, oracle.polymul(s, x, y))
    assert np.array_equal(z[: 4 * p.n], oracle.schoolbook(s, x[: 4 * p.n], y[: 4 * p.n]))
    assert np.array_equal(z[2 * p.n: 3 * p.n], y[2 * p.n: 3 * p.n]) and z[3 * p.n] == p.q - 1
    # NTT-domain layout: bit for bit the reference's Phi-scale + radix2NTTGS output
    assert np.array_equal(oracle.fast_forward(s, x), oracle.forward(s, x))


def test_fast_cpu_path_golden_and_reference(oracle, golden, reference):
    d = np.load(os.path.join(HERE, "golden", "golden_III_b2.npz"))
    assert np.array_equal(oracle.fast_polymul(SET_III, d["x"], d["y"]), d["z"])
    assert sha(oracle.fast_forward(SET_III, d["x"])) == golden["III_random_b2"]["fwd_sha256"]
    rng = np.random.default_rng(31)
    x = rng.integers(0, 8404993, 64 * 1024, dtype=np.uint32)
    y = rng.integers(0, 8404993, 64 * 1024, dtype=np.uint32)
    assert np.array_equal(oracle.fast_polymul(SET_III, x, y, threads=2), reference.polymul(x, y, 0, 2))
