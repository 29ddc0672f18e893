"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs, against the committed golden vectors, and — at BASELINE.json's full batch sizes —
through size-independent properties.  Bit-exact everywhere (integer arithmetic)."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ALL_SETS = [0, 1, 2, 3]
FULL_BATCH = {0: 65536, 1: 65536, 2: 65536, 3: 32768}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def engines(qt):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    # ONE stream for torch and the engines: the tests interleave torch device ops (copy_, clone) with engine
    # calls, and an engine's own stream is non-blocking, i.e. not ordered against torch's default stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    es = {s: qt.Engine(s, 0) for s in ALL_SETS}
    for e in es.values():
        e.set_stream(stream.cuda_stream)
    yield es
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    for e in es.values():
        e.close()


def rand_pair(q, words, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, q, words, dtype=np.uint32), rng.integers(0, q, words, dtype=np.uint32)


# 0 = automatic, 1 = direct coalesced loads, 2 = TMA bulk copies + mbarrier, 3 = n=2048 as two halves (one warp),
# 4 = n=2048 as two halves by a pair of warps, 5 = FP64-quotient butterflies (the 23-bit moduli: sets 0 and 1)
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("s", ALL_SETS)
@pytest.mark.parametrize("B", [1, 2, 3, 67, 1000, 5001])
def test_fused_polymul_equals_oracle(engines, oracle, s, B, variant):
    eng = engines[s]
    if variant in (3, 4) and s != 3:
        with pytest.raises(Exception):
            eng.set_fused_variant(variant)   # the split tile exists for n=2048 only
        return
    if variant == 5 and s not in (0, 1):
        with pytest.raises(Exception):
            eng.set_fused_variant(variant)   # FP64-quotient butterflies exist for the signed-lazy sets only
        return
    eng.set_fused_variant(variant)
    try:
        x, y = rand_pair(eng.q, B * eng.n, 100 * s + B)
        assert np.array_equal(eng.polymul_np(x, y), oracle.polymul(s, x, y, threads=0))
    finally:
        eng.set_fused_variant(0)


def test_fuzz_batches_sets_variants(engines, oracle):
    """Random (set, batch, data path) combinations: ragged batches around the grid sizes (148 SMs x 12/16
    warps, two polynomials per warp for n=512) through fused, cached-transform and unfused entry points."""
    import torch
    rng = np.random.default_rng(20261018)
    special = [147, 148, 149, 295, 296, 297, 2367, 2368, 2369, 1775, 1776, 1777, 4735, 4737]
    for it in range(36):
        s = int(rng.integers(0, 4))
        eng = engines[s]
        B = int(special[it % len(special)] if it % 3 == 0 else rng.integers(1, 3000))
        variant = int(rng.choice([0, 1, 2] + ([3, 4] if s == 3 else []) + ([5] if s in (0, 1) else [])))
        eng.set_fused_variant(variant)
        try:
            x, y = rand_pair(eng.q, B * eng.n, 5000 + it)
            ref = oracle.polymul(s, x, y, threads=0)
            assert np.array_equal(eng.polymul_np(x, y), ref), (s, B, variant)
            tx = torch.from_numpy(x.view(np.int32)).cuda(); ty = torch.from_numpy(y.view(np.int32)).cuda(); tz = torch.empty_like(tx)
            eng.ntt_forward(tx)
            eng.polymul_ntt(tx, ty, tz, broadcast=False)
            eng.synchronize()
            assert np.array_equal(tz.cpu().numpy().view(np.uint32), ref), ("cached", s, B, variant)
            assert np.array_equal(eng.inverse_natural_np(eng.pointwise_np(eng.forward_natural_np(x), eng.forward_natural_np(y))), ref)
        finally:
            eng.set_fused_variant(0)


@pytest.mark.parametrize("s", ALL_SETS)
def test_fused_variants_agree_at_full_size(engines, s):
    import torch
    eng = engines[s]
    B = FULL_BATCH[s]
    x = torch.empty(B * eng.n, dtype=torch.int32, device="cuda")
    y = torch.empty_like(x); z1 = torch.empty_like(x); z2 = torch.empty_like(x)
    eng.fill_uniform(x, 1, 7); eng.fill_uniform(y, 2, 7)
    eng.set_fused_variant(1); eng.polymul(x, y, z1)
    eng.set_fused_variant(2); eng.polymul(x, y, z2)
    eng.set_fused_variant(0); eng.synchronize()
    assert torch.equal(z1, z2)
    eng.polymul(x, y, z2); eng.synchronize()      # automatic choice (n=2048: two warps per polynomial)
    assert torch.equal(z1, z2)
    if s in (0, 1):                               # FP64-quotient butterflies: full batch, worst-case operands, a ragged batch
        eng.set_fused_variant(5)
        z2.zero_(); eng.polymul(x, y, z2); eng.synchronize()
        assert torch.equal(z1, z2)
        x.fill_(eng.q - 1); y.fill_(eng.q - 1)
        eng.set_fused_variant(2); eng.polymul(x, y, z1)
        eng.set_fused_variant(5); z2.zero_(); eng.polymul(x, y, z2, B - 149); eng.synchronize()
        assert torch.equal(z1[: (B - 149) * eng.n], z2[: (B - 149) * eng.n]) and not z2[(B - 149) * eng.n:].any()
        eng.set_fused_variant(0)
    if s == 3:
        for v in (3, 4):                          # both split-tile kernels, full batch and a ragged one
            eng.set_fused_variant(v)
            z2.zero_(); eng.polymul(x, y, z2); eng.synchronize()
            assert torch.equal(z1, z2), v
            z2.zero_(); eng.polymul(x, y, z2, B - 149); eng.synchronize()
            assert torch.equal(z1[: (B - 149) * eng.n], z2[: (B - 149) * eng.n]) and not z2[(B - 149) * eng.n:].any(), v
        eng.set_fused_variant(0)


def test_golden_vectors_III(engines, golden):
    eng = engines[1]
    d = np.load(os.path.join(HERE, "golden", "golden_III_b2.npz"))
    z = eng.polymul_np(d["x"], d["y"])
    assert np.array_equal(z, d["z"]) and sha(z) == golden["III_random_b2"]["z_sha256"]
    f = eng.forward_np(d["x"])
    assert np.array_equal(f, d["fwd_x"]) and sha(f) == golden["III_random_b2"]["fwd_sha256"]
    assert np.array_equal(eng.inverse_np(d["fwd_x"]), d["x"])
    ones = np.ones(2048, np.uint32)
    assert sha(eng.polymul_np(ones, ones)) == golden["III_all_ones"]["z_sha256"]
    ramp = np.zeros(2048, np.uint32)
    for b in range(2):
        ramp[b * 1024: b * 1024 + 512] = 512 - np.arange(512)
    assert sha(eng.forward_np(ramp)) == golden["III_ramp_forward_sha256"]
    assert sha(eng.polymul_np(ramp, ramp)) == golden["III_ramp_square_sha256"]


@pytest.mark.parametrize("s", [0, 2, 3])
def test_golden_vectors_other_sets(engines, oracle, golden, qt, s):
    eng = engines[s]
    x, y, _ = oracle.xorshift_pair(eng.q, eng.n)
    g = golden[qt.SET_NAMES[s] + "_random_b1"]
    z = eng.polymul_np(x, y)
    assert sha(z) == g["z_sha256"] and list(z[:4]) == g["z_first4"]


@pytest.mark.parametrize("s", ALL_SETS)
def test_edge_inputs(engines, oracle, s):
    eng = engines[s]
    n, q = eng.n, eng.q
    rng = np.random.default_rng(5)
    rows_x, rows_y = [], []
    y = rng.integers(0, q, n, dtype=np.uint32)
    one = np.zeros(n, np.uint32); one[0] = 1
    xm = np.zeros(n, np.uint32); xm[n - 1] = 1
    x1 = np.zeros(n, np.uint32); x1[1] = 1
    top = np.full(n, q - 1, np.uint32)
    for a, b in ((one, y), (xm, x1), (np.zeros(n, np.uint32), y), (top, top), (np.ones(n, np.uint32), np.ones(n, np.uint32)), (top, y)):
        rows_x.append(a); rows_y.append(b)
    x = np.concatenate(rows_x); yy = np.concatenate(rows_y)
    z = eng.polymul_np(x, yy)
    assert np.array_equal(z, oracle.polymul(s, x, yy))
    assert np.array_equal(z[:n], y)                       # x = 1 -> z = y
    assert z[n] == q - 1 and not z[n + 1: 2 * n].any()    # X^(n-1) * X = -1
    assert not z[2 * n: 3 * n].any()
    exp = ((2 * np.arange(n, dtype=np.int64) + 2 - n) % q).astype(np.uint32)
    assert np.array_equal(z[3 * n: 4 * n], exp) and np.array_equal(z[4 * n: 5 * n], exp)   # all-ones KAT
    assert z.max() < q


@pytest.mark.parametrize("s", ALL_SETS)
def test_forward_inverse_pointwise_equal_oracle(engines, oracle, s):
    eng = engines[s]
    B = 33
    x, y = rand_pair(eng.q, B * eng.n, 900 + s)
    f = eng.forward_np(x)
    assert np.array_equal(f, oracle.forward(s, x))
    assert np.array_equal(eng.inverse_np(f), x)
    assert np.array_equal(eng.inverse_np(y), oracle.inverse(s, y))
    assert np.array_equal(eng.pointwise_np(x, y), oracle.pointwise(s, x, y))
    # unfused composition == fused kernel == oracle
    z = eng.inverse_np(eng.pointwise_np(eng.forward_np(x), eng.forward_np(y)))
    assert np.array_equal(z, eng.polymul_np(x, y))


@pytest.mark.parametrize("s", ALL_SETS)
def test_bitrev_copy_and_natural_order(engines, oracle, s):
    import torch
    eng = engines[s]
    x, _ = rand_pair(eng.q, 9 * eng.n, 40 + s)
    f = eng.forward_np(x)
    t = torch.from_numpy(f.view(np.int32)).cuda()
    o = torch.empty_like(t)
    eng.bitrev_copy(t, o)
    eng.synchronize()
    nat = o.cpu().numpy().view(np.uint32)
    assert np.array_equal(nat, oracle.forward_natural(s, x))     # what the Stockham pipeline leaves
    assert np.array_equal(nat, oracle.bitrev_copy(s, f))


@pytest.mark.parametrize("B", [1, 3, 257, 4099])
@pytest.mark.parametrize("s", ALL_SETS)
def test_natural_order_transforms_equal_stockham_oracle(engines, oracle, s, B):
    """qt_ntt_forward_natural / qt_ntt_inverse_natural == the reference's Stockham pipeline ordering
    (Phi scale + radix2NTTStock, radix2INTTStock + invPhi scale; NTT.cu:1162-1191, 1339-1370)."""
    eng = engines[s]
    x, y = rand_pair(eng.q, B * eng.n, 7000 + 10 * s + B)
    x[: eng.n] = eng.q - 1
    f = eng.forward_natural_np(x)
    assert np.array_equal(f, oracle.forward_natural(s, x))
    assert np.array_equal(eng.inverse_natural_np(f), x)
    assert np.array_equal(eng.inverse_natural_np(y), oracle.inverse_natural(s, y))
    # natural-order pipeline (Stockham variant of the reference) == fused product
    fy = eng.forward_natural_np(y)
    z = eng.inverse_natural_np(eng.pointwise_np(f, fy))
    assert np.array_equal(z, oracle.polymul(s, x, y, 3))  # variant 3 = the Stockham composition


def test_in_place_and_fill_uniform(engines, oracle):
    import torch
    eng = engines[1]
    B = 50
    x = torch.empty(B * eng.n, dtype=torch.int32, device="cuda")
    y = torch.empty_like(x)
    eng.fill_uniform(x, 1, 12345)
    eng.fill_uniform(y, 2, 12345)
    eng.synchronize()
    xs = x.cpu().numpy().view(np.uint32); ys = y.cpu().numpy().view(np.uint32)
    assert np.array_equal(xs, oracle.splitmix(1, 12345, eng.q, B * eng.n))
    assert np.array_equal(ys, oracle.splitmix(2, 12345, eng.q, B * eng.n))
    eng.polymul(x, y, x)  # z aliases x
    eng.synchronize()
    assert np.array_equal(x.cpu().numpy().view(np.uint32), oracle.polymul(1, xs, ys))


@pytest.mark.parametrize("s", [0, 1, 3])
def test_host_pointer_entry_point(engines, oracle, s):
    eng = engines[s]
    B = (4 << 20) // eng.n * 2 + 37                                 # two full chunks + a ragged tail
    x, y = rand_pair(eng.q, B * eng.n, 77 + s)
    z = eng.polymul_host(x, y)                                      # pageable numpy memory
    idx = np.r_[0:3, B // 2 - 1: B // 2 + 2, B - 3: B]
    pick = lambda a: np.concatenate([a[i * eng.n:(i + 1) * eng.n] for i in idx])
    assert np.array_equal(pick(z), oracle.polymul(s, pick(x), pick(y)))
    zd = eng.polymul_np(x, y)
    assert np.array_equal(z, zd)                                    # whole batch: host path == device path
    assert eng.polymul_host(x[:0], y[:0]).size == 0                  # empty batch
    # pinned buffers take the in-place copy pipeline; mixed pinned / pageable operands are allowed
    import torch
    pin = lambda a: torch.from_numpy(a.view(np.int32)).pin_memory().numpy().view(np.uint32)
    xp, yp = pin(x), pin(y)
    zp = pin(np.zeros_like(x))
    eng.polymul_host(xp, yp, zp)
    assert np.array_equal(zp, zd)
    assert np.array_equal(eng.polymul_host(xp, y), zd)              # x pinned, y and z pageable
    zp[:] = 0
    eng.polymul_host(x, y, zp)                                      # only z pinned
    assert np.array_equal(zp, zd)


def test_context_lifecycle_and_small_host_batches(qt, oracle):
    """Contexts can be created and destroyed repeatedly (tables, streams, pipelines are released), several
    contexts of different sets share a device, and the host entry points handle batches around the chunk
    size of both pipelines (pinned: 4 Mi words per chunk, pageable: 2 Mi words)."""
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    for rep in range(6):
        es = [qt.Engine(s, 0) for s in ALL_SETS]
        for s, e in zip(ALL_SETS, es):
            x, y = rand_pair(e.q, 3 * e.n, 1000 + rep * 10 + s)
            assert np.array_equal(e.polymul_host(x, y), oracle.polymul(s, x, y))   # builds the staged pipeline
        for e in es:
            e.close()
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20), "device memory leaked by create/destroy"
    eng = qt.Engine(1, 0)
    n = eng.n
    chunk = (2 << 20) // n
    pin = lambda a: torch.from_numpy(a.view(np.int32)).pin_memory().numpy().view(np.uint32)
    for B in (1, 2, chunk - 1, chunk, chunk + 1, 2 * chunk + 1):
        x, y = rand_pair(eng.q, B * n, 2000 + B)
        ref = eng.polymul_np(x, y)
        assert np.array_equal(eng.polymul_host(x, y), ref), B                       # pageable
        zp = pin(np.zeros_like(x))
        eng.polymul_host(pin(x), pin(y), zp)                                        # pinned
        assert np.array_equal(zp, ref), B
    eng.close()


def test_multi_gpu_host_sharding(qt, oracle):
    # contiguous batch slices over however many GPUs the box has (1 is fine): no collective involved
    B = 301
    x, y = rand_pair(8404993, B * 1024, 4242)
    z = qt.polymul_host_multi(qt.SET_III, x, y, 0)
    assert np.array_equal(z, oracle.polymul(1, x, y))


def test_harness_mirror_all_ones(qt):
    # reads like the reference's own check: drivers fill x = y = 1 and z must be (2k+2-n) mod q
    n, q, B = 1024, 8404993, 2
    exp = ((2 * np.arange(n, dtype=np.int64) + 2 - n) % q).astype(np.uint32)
    for drv in (qt.harness.test_NTT_Stockham_nega_gpu, qt.harness.test_NTT_GS_CT_nega_gpu, qt.harness.test_NTT_CT_CT_nega_gpu,
                qt.harness.test_NTT_GS_GS_nega_gpu, qt.harness.test_NTT_CT_GS_nega_gpu):
        x = np.zeros(B * n, np.uint32); y = np.zeros(B * n, np.uint32); z = np.zeros(B * n, np.uint32)
        drv(x, y, z, verbose=False)
        assert x.min() == 1 and np.array_equal(z[:n], exp) and np.array_equal(z[n:], exp)
        assert list(z[:3]) == [8403971, 8403973, 8403975]          # what the reference prints


@pytest.mark.parametrize("s", ALL_SETS)
def test_full_size_properties(engines, oracle, s):
    """BASELINE.json batch sizes: properties that need no O(B) CPU work + an oracle sample."""
    import torch
    eng = engines[s]
    B, n, q = FULL_BATCH[s], eng.n, eng.q
    x = torch.empty(B * n, dtype=torch.int32, device="cuda")
    y = torch.empty_like(x); z = torch.empty_like(x); w = torch.empty_like(x)
    eng.fill_uniform(x, 1, 0); eng.fill_uniform(y, 2, 0)
    eng.polymul(x, y, z)
    eng.polymul(y, x, w)                                   # commutativity over the whole batch
    eng.synchronize()
    assert torch.equal(z, w)
    assert int(z.max()) < q and int(z.min()) >= 0          # canonical
    # forward -> inverse round trip over the whole batch ("Identical.")
    w.copy_(x); eng.ntt_forward(w); eng.ntt_inverse(w); eng.synchronize()
    assert torch.equal(w, x)
    # unfused pipeline == fused kernel over the whole batch
    w.copy_(x); eng.ntt_forward(w)
    v = y.clone(); eng.ntt_forward(v); eng.pointwise(w, v, w); eng.ntt_inverse(w); eng.synchronize()
    assert torch.equal(w, z)
    # oracle sample: first / middle / last polynomials
    idx = [0, 1, B // 2, B - 2, B - 1]
    pick = lambda t: np.concatenate([t[i * n:(i + 1) * n].cpu().numpy().view(np.uint32) for i in idx])
    assert np.array_equal(pick(z), oracle.polymul(s, pick(x), pick(y)))


@pytest.mark.parametrize("mode", [1, 2])  # programmatic dependent launch: never / always
@pytest.mark.parametrize("s", ALL_SETS)
def test_dependent_launch_chains(engines, oracle, s, mode):
    """Back-to-back launches on one stream in which every launch consumes the previous launch's output — the case
    programmatic dependent launch must keep ordered (operands are touched only after griddepcontrol.wait)."""
    import torch
    eng = engines[s]
    n, q = eng.n, eng.q
    eng.set_launch_overlap(mode)
    try:
        for B in (5, 300, 9000):
            x, y = rand_pair(q, B * n, 900 + s + B)
            tx = torch.from_numpy(x.view(np.int32)).cuda(); ty = torch.from_numpy(y.view(np.int32)).cuda()
            t1 = torch.empty_like(tx); t2 = torch.empty_like(tx); t3 = torch.empty_like(tx)
            for _ in range(3):                       # repeated: the launches of one round overlap the previous round's tail
                eng.polymul(tx, ty, t1)              # t1 = x y
                eng.polymul(t1, ty, t2)              # t2 = x y^2
                eng.polymul(t2, t1, t3)              # t3 = x^2 y^3
                t4 = t3.clone()
                eng.ntt_forward(t4); eng.ntt_inverse(t4)          # round trip of a fresh output
                ah = t1[:n].clone(); eng.ntt_forward(ah, 1)
                t5 = torch.empty_like(tx)
                eng.polymul_ntt(ah, t3, t5, True)    # cached transform of the first polynomial of t1, broadcast
            eng.synchronize()
            r1 = oracle.polymul(s, x, y); r2 = oracle.polymul(s, r1, y); r3 = oracle.polymul(s, r2, r1)
            assert np.array_equal(t3.cpu().numpy().view(np.uint32), r3)
            assert np.array_equal(t4.cpu().numpy().view(np.uint32), r3)
            assert np.array_equal(t5.cpu().numpy().view(np.uint32), oracle.polymul(s, np.tile(r1[:n], B), r3))
    finally:
        eng.set_launch_overlap(0)


def test_nussbaumer_ring_equals_oracle(engines, oracle, qt):
    import torch
    eng = engines[1]
    rng = np.random.default_rng(9)
    B = 5
    x = rng.integers(0, 2 ** 32, B * 1024, dtype=np.uint32)
    y = rng.integers(0, 2 ** 32, B * 1024, dtype=np.uint32)
    x[:1024] = 1; y[:1024] = 1                                       # the reference's all-ones fixture
    x[1024:2048] = 0xFFFFFFFF                                        # non-normalised zeros
    y[2048:3072] = 0; y[2048 + rng.choice(1024, 40, replace=False)] = 0xFFFFFFFE   # sparse ternary (-1)
    # products that are multiples of 2^32-1 without being zero (the chain must answer 0xFFFFFFFF, not 0),
    # and products whose 64-bit sum overflows
    x[3072:4096] = 0; y[3072:4096] = 0
    x[3072] = 3; y[3072] = 0x55555555; x[3073] = 0x10000; y[3073] = 0x10000; x[3074] = 0xFFFFFFFF; y[3074] = 0xFFFFFFFF
    x[4096:5120] = 0xFFFFFFFE; y[4096:5120] = 0xFFFFFFFE
    tx = torch.from_numpy(x.view(np.int32)).cuda(); ty = torch.from_numpy(y.view(np.int32)).cuda(); tz = torch.empty_like(tx)
    try:
        eng.nussbaumer(tx, ty, tz, qt.RING_2P32M1)
    except qt.QtError as e:
        if "unsupported" in str(e):
            pytest.skip("Nussbaumer kernel not built yet")
        raise
    eng.synchronize()
    z = tz.cpu().numpy().view(np.uint32)
    assert np.array_equal(z, oracle.nussbaumer(1024, x, y))
    assert list(z[:3]) == [4294966273, 4294966275, 4294966277]


@pytest.mark.parametrize("s", [0, 2, 3])
def test_nussbaumer_ring_other_sizes(engines, oracle, golden, qt, s):
    """n-generic Nussbaumer over Z/(2^32-1) (n = 512: 16x32 split, n = 2048: 32x64 split), SURVEY.md 8a/8c."""
    import torch
    eng = engines[s]
    n = eng.n
    rng = np.random.default_rng(90 + s)
    B = 7
    x = rng.integers(0, 2 ** 32, B * n, dtype=np.uint32)
    y = rng.integers(0, 2 ** 32, B * n, dtype=np.uint32)
    xx, yy, _ = oracle.xorshift_pair(eng.q, n)                       # the golden stream of SURVEY.md 8c
    x[:n] = xx; y[:n] = yy
    x[n:2 * n] = 0xFFFFFFFF                                          # non-normalised zeros
    y[2 * n:3 * n] = 0; y[2 * n + rng.choice(n, 40, replace=False)] = 0xFFFFFFFE   # sparse ternary (-1)
    tx = torch.from_numpy(x.view(np.int32)).cuda(); ty = torch.from_numpy(y.view(np.int32)).cuda(); tz = torch.empty_like(tx)
    eng.nussbaumer(tx, ty, tz, qt.RING_2P32M1)
    eng.synchronize()
    z = tz.cpu().numpy().view(np.uint32)
    assert np.array_equal(z, oracle.nussbaumer(n, x, y))
    z0 = z[:n].copy(); z0[z0 == 0xFFFFFFFF] = 0
    name = {0: "qTESLA-I", 2: "qTESLA-p-I", 3: "qTESLA-p-III"}[s]
    assert sha(z0) == golden[name + "_ring_schoolbook_b1_sha256"]


@pytest.mark.parametrize("nv", [0, 1, 2, 3])  # row products: automatic, schoolbook, recursive, FP64 pipe
@pytest.mark.parametrize("s", ALL_SETS)
def test_nussbaumer_modq_equals_ntt(engines, oracle, qt, s, nv):
    import torch
    eng = engines[s]
    if nv == 3 and eng.q >= 1 << 25:
        with pytest.raises(qt.QtError):      # FP64 row products need q < 2^25
            eng.set_nussbaumer_variant(3)
        return
    B = 9
    n, q = eng.n, eng.q
    x, y = rand_pair(q, B * n, 555 + s)
    # worst cases for the unreduced (lazy) stage arithmetic: everything q-1; q-1 against alternating 0 / q-1
    x[:n] = q - 1; y[:n] = q - 1
    x[n:2 * n] = q - 1; y[n:2 * n] = np.where(np.arange(n) % 2 == 0, q - 1, 0)
    x[2 * n:3 * n] = np.where(np.arange(n) % 64 < 32, q - 1, 0); y[2 * n:3 * n] = q - 1
    x[3 * n:4 * n] = 0
    x[4 * n:5 * n] = q // 2 + 1; y[4 * n:5 * n] = q - 1          # just above the centring threshold of the recursive rows
    x[5 * n:6 * n] = q // 2; y[5 * n:6 * n] = np.where(np.arange(n) % 8 < 4, q - 1, 1)
    tx = torch.from_numpy(x.view(np.int32)).cuda(); ty = torch.from_numpy(y.view(np.int32)).cuda(); tz = torch.empty_like(tx)
    eng.set_nussbaumer_variant(nv)
    try:
        eng.nussbaumer(tx, ty, tz, qt.RING_MODQ)
        eng.synchronize()
    finally:
        eng.set_nussbaumer_variant(0)
    assert np.array_equal(tz.cpu().numpy().view(np.uint32), oracle.polymul(s, x, y))


@pytest.mark.parametrize("s", ALL_SETS)
def test_nussbaumer_row_variants_agree_at_size(engines, oracle, qt, s):
    """schoolbook and recursive row products: identical results on a batch that fills every SM several times, a
    sample of it checked against the oracle; the ring 2^32-1 ignores the variant."""
    import torch
    eng = engines[s]
    n, q = eng.n, eng.q
    B = 4099 if n < 2048 else 1031
    x, y = rand_pair(q, B * n, 777 + s)
    tx = torch.from_numpy(x.view(np.int32)).cuda(); ty = torch.from_numpy(y.view(np.int32)).cuda()
    outs = {}
    for nv in (1, 2) + ((3,) if q < 1 << 25 else ()):
        tz = torch.empty_like(tx)
        eng.set_nussbaumer_variant(nv)
        try:
            eng.nussbaumer(tx, ty, tz, qt.RING_MODQ)
            eng.synchronize()
            outs[nv] = tz.cpu().numpy().view(np.uint32)
            tr = torch.empty_like(tx)
            eng.nussbaumer(tx, ty, tr, qt.RING_2P32M1)
            eng.synchronize()
            outs[("ring", nv)] = tr.cpu().numpy().view(np.uint32)
        finally:
            eng.set_nussbaumer_variant(0)
    assert np.array_equal(outs[1], outs[2])
    assert 3 not in outs or np.array_equal(outs[1], outs[3])
    assert np.array_equal(outs[("ring", 1)], outs[("ring", 2)])
    k = 40 * n
    assert np.array_equal(outs[2][:k], oracle.polymul(s, x[:k], y[:k]))
    assert np.array_equal(outs[2][-k:], oracle.polymul(s, x[-k:], y[-k:]))
    with pytest.raises(qt.QtError):
        eng.set_nussbaumer_variant(4)


def test_cxx_harness_reference_command_line():
    """tools/harness/qtesla_harness = the reference's main.cu re-created on the C ABI: the reference's
    own command line `-speedgpu 3` must print the all-ones known answer."""
    import subprocess
    root = os.path.dirname(HERE)
    subprocess.run(["make", "-s", "-C", os.path.join(root, "tools", "harness")], check=True)
    exe = os.path.join(root, "tools", "harness", "qtesla_harness")
    for opt in ("2", "3", "4", "5", "6"):
        out = subprocess.run([exe, "-speedgpu", opt], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        assert "Multiplications per second" in out.stdout
        assert "8403971 8403973 8403975" in out.stdout            # z[k] = 2k+2-n mod q, both batch rows
        assert out.stdout.count("8403971 8403973 8403975") == 2
    out = subprocess.run([exe, "-speedgpu", "9"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "4294966273 4294966275 4294966277" in out.stdout   # Nussbaumer ring KAT
    out = subprocess.run([exe, "-speedgpu", "3", "-set", "p-III", "-batch", "3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and str((2 - 2048) % 856145921) in out.stdout


@pytest.mark.parametrize("B", [41, 2000])
@pytest.mark.parametrize("s,variant", [(0, 0), (1, 0), (2, 0), (3, 0), (3, 2)])  # n=2048: split tile (auto) and 64-wide tile
def test_cached_transform_product(engines, oracle, s, variant, B):
    """qTESLA-shaped caller: one public polynomial a (transformed once) times many y, incl. sparse ternary y."""
    import torch
    eng = engines[s]
    eng.set_fused_variant(variant)
    n, q = eng.n, eng.q
    rng = np.random.default_rng(70 + s)
    a = rng.integers(0, q, n, dtype=np.uint32)
    y = rng.integers(0, q, B * n, dtype=np.uint32)
    for b in range(0, B, 3):                                   # every third y: weight-48 ternary challenge
        row = np.zeros(n, np.uint32)
        row[rng.choice(n, 48, replace=False)] = np.where(rng.random(48) < 0.5, 1, q - 1)
        y[b * n:(b + 1) * n] = row
    ta = torch.from_numpy(a.view(np.int32)).cuda()
    eng.ntt_forward(ta)                                        # NTT(a), once
    ty = torch.from_numpy(y.view(np.int32)).cuda()
    tz = torch.empty_like(ty)
    eng.polymul_ntt(ta, ty, tz, broadcast=True)
    eng.synchronize()
    ref = oracle.polymul(s, np.tile(a, B), y)
    assert np.array_equal(tz.cpu().numpy().view(np.uint32), ref)
    # per-product a_hat
    aa = rng.integers(0, q, B * n, dtype=np.uint32)
    taa = torch.from_numpy(aa.view(np.int32)).cuda()
    eng.ntt_forward(taa)
    eng.polymul_ntt(taa, ty, tz, broadcast=False)
    eng.synchronize()
    eng.set_fused_variant(0)
    assert np.array_equal(tz.cpu().numpy().view(np.uint32), oracle.polymul(s, aa, y))


def test_unaligned_operands_take_the_direct_kernel(engines, oracle):
    """TMA bulk copies need 16-byte alignment; 4-byte aligned views must still work (direct-load kernel)."""
    import torch
    eng = engines[1]
    B = 19
    x, y = rand_pair(eng.q, B * eng.n, 31337)
    tx = torch.zeros(B * eng.n + 1, dtype=torch.int32, device="cuda")
    ty = torch.zeros(B * eng.n + 3, dtype=torch.int32, device="cuda")
    tz = torch.zeros(B * eng.n + 1, dtype=torch.int32, device="cuda")
    tx[1:].copy_(torch.from_numpy(x.view(np.int32)))
    ty[3:].copy_(torch.from_numpy(y.view(np.int32)))
    eng.polymul(tx.data_ptr() + 4, ty.data_ptr() + 12, tz.data_ptr() + 4, B)
    eng.synchronize()
    assert np.array_equal(tz[1:].cpu().numpy().view(np.uint32), oracle.polymul(1, x, y))
    assert int(tz[0]) == 0


def test_batch_beyond_32bit_word_index(engines, oracle):
    """4 Mi + 5 polynomials of n=1024: more than 2^32 coefficients per array (16 GiB each) — 64-bit indexing in
    the kernels, the generator and the TMA addresses.  Checked on the first and last polynomials."""
    import torch
    eng = engines[1]
    n = eng.n
    B = (1 << 22) + 5
    free, _ = torch.cuda.mem_get_info()
    if free < 3 * B * n * 4 + (2 << 30):
        pytest.skip("not enough device memory")
    x = torch.empty(B * n, dtype=torch.int32, device="cuda")
    y = torch.empty_like(x); z = torch.empty_like(x)
    eng.fill_uniform(x, 1, 0); eng.fill_uniform(y, 2, 0)
    eng.polymul(x, y, z)
    eng.synchronize()
    for lo in (0, (1 << 22) - 2, B - 3):
        xs = x[lo * n:(lo + 3) * n].cpu().numpy().view(np.uint32)
        ys = y[lo * n:(lo + 3) * n].cpu().numpy().view(np.uint32)
        assert np.array_equal(xs, oracle.splitmix(1, lo * n, eng.q, 3 * n))
        assert np.array_equal(z[lo * n:(lo + 3) * n].cpu().numpy().view(np.uint32), oracle.polymul(1, xs, ys))
    del x, y, z
    torch.cuda.empty_cache()


@pytest.mark.parametrize("s", ALL_SETS)
def test_full_batch_worst_case_operands(engines, s):
    """Every coefficient q-1 (the largest magnitude the lazy ranges have to absorb), whole bench batch:
    (q-1)*(q-1) == 1*1, so every product is the all-ones known answer z[k] = 2k+2-n (mod q)."""
    import torch
    eng = engines[s]
    B, n, q = FULL_BATCH[s], eng.n, eng.q
    x = torch.full((B * n,), q - 1, dtype=torch.int32, device="cuda")
    z = torch.empty_like(x)
    eng.polymul(x, x, z)
    eng.synchronize()
    exp = torch.from_numpy(((2 * np.arange(n, dtype=np.int64) + 2 - n) % q).astype(np.int32)).cuda()
    assert torch.equal(z.view(B, n), exp.expand(B, n))
    # same through the unfused entry points
    w = x.clone()
    eng.ntt_forward(w); eng.pointwise(w, w, w); eng.ntt_inverse(w); eng.synchronize()
    assert torch.equal(w, z)
