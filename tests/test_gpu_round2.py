"""GPU parity tests (-m gpu), second set: the CUDA path against the UNMODIFIED reference directly (oracle/_ref ships
to the GPU box: its CPU functions and its own GPU kernels), the reference's own main.cu linked against this library,
the in-process multi-GPU handle, and the smaller entry points added after round 1."""
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q3, N3 = 8404993, 1024


def rand_pair(q, words, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, q, words, dtype=np.uint32), rng.integers(0, q, words, dtype=np.uint32)


@pytest.fixture(scope="module")
def eng3(qt):
    e = qt.Engine(qt.SET_III, 0)
    yield e
    e.close()


# ---- CUDA vs the unmodified reference, no port in between ------------------------------------------------------
def test_cuda_equals_reference_cpu_functions(eng3, reference, qt):
    """fresh random qTESLA-III operands: fused product == every CPU composition of the reference
    (NTT.cu:1820-1984), forward / inverse == Phi-scale + radix2NTTGS / radix2INTT + invPhi (NTT.cu:1866-1876,
    1845-1849), Nussbaumer == nussbaumer_fft (NTT.cu:167-277)."""
    import torch
    B = 37
    x, y = rand_pair(Q3, B * N3, 20261019)
    z = eng3.polymul_np(x, y)
    for variant in range(4):  # GS->CT, GS/GS, CT/CT, Stockham
        assert np.array_equal(z, reference.polymul(x, y, variant, 2)), variant
    f = eng3.forward_np(x)
    assert np.array_equal(f, reference.forward(x))
    assert np.array_equal(eng3.inverse_np(f), reference.inverse(f))
    assert np.array_equal(eng3.inverse_np(f), x)
    xr, yr = x[: 6 * N3], y[: 6 * N3]
    tx = torch.from_numpy(xr.view(np.int32)).cuda()
    ty = torch.from_numpy(yr.view(np.int32)).cuda()
    tz = torch.empty_like(tx)
    torch.cuda.synchronize()
    eng3.nussbaumer(tx, ty, tz, qt.RING_2P32M1)
    eng3.synchronize()
    assert np.array_equal(tz.cpu().numpy().view(np.uint32), reference.nussbaumer(xr, yr))


def test_cuda_equals_reference_gpu_kernels(eng3, reference):
    """the reference's OWN GPU kernels, unmodified, in the launch order of test_NTT_CT_GS_nega_gpu
    (NTT.cu:2388-2425) on the same random operands and the same GPU: identical products."""
    if not hasattr(reference.lib, "qtref_gpu_ct_gs"):
        pytest.skip("oracle/_ref/libqtref.so predates the GPU leg")
    B = 300
    x, y = rand_pair(Q3, B * N3, 777)
    z_ref, kernel_ms, total_ms = reference.gpu_ct_gs(x, y, reps=2)
    assert kernel_ms > 0 and total_ms >= kernel_ms * 0.5
    assert np.array_equal(eng3.polymul_np(x, y), z_ref)
    assert np.array_equal(z_ref, reference.polymul(x, y, 0, 2))


# ---- the reference's own main.cu on top of the library ------------------------------------------------------------
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin_main")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/dropin_main not built (make -C oracle dropin needs /root/reference)")
@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
def test_reference_main_cu_runs_on_the_library(k):
    """`make -C oracle dropin`: the UNMODIFIED main.cu (argument parsing, twiddle set-up, malloc'd operands, dispatch
    main.cu:190-226) with the INTEGRATION.md patch; -speedgpu 2..6 must print the all-ones fixture
    (NTT.cu:1822, 2099): z[k] = 2k + 2 - n mod q = 8403971 8403973 8403975 ... for both polynomials of BATCH 2."""
    r = subprocess.run([DROPIN, "-speedgpu", str(k)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    out = r.stdout
    assert "Multiplications per second" in out and "Batch Size is 2" in out
    dump = out.split("z:")[1].split()
    vals = [int(v) for v in dump[: 2 * N3]]
    exp = [(2 * i + 2 - N3) % Q3 for i in range(N3)]
    assert vals[:3] == [8403971, 8403973, 8403975]
    assert vals[:N3] == exp and vals[N3:] == exp


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/dropin_main not built")
def test_reference_main_cu_nussbaumer_option():
    # main.cu routes -speedcpu 6 to test_nussbaumer (main.cu:187-188); with the drop-in it runs batched on the GPU
    r = subprocess.run([DROPIN, "-speedcpu", "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    dump = r.stdout.split("z:")[1].split()
    assert [int(v) for v in dump[:2]] == [4294966273, 4294966275]  # SURVEY.md 8c-1, ring 2^32-1


# ---- in-process multi-GPU handle ---------------------------------------------------------------------------------
def test_multi_engine_handle_all_visible_gpus(qt, oracle):
    B = 1203
    x, y = rand_pair(Q3, B * N3, 4243)
    ref = oracle.polymul(1, x, y, threads=0)
    for g in sorted({1, qt.device_count()}):
        m = qt.MultiEngine(qt.SET_III, g)
        assert m.ngpus == g
        for _ in range(3):  # persistent worker threads: several jobs through the same handle
            assert np.array_equal(m.polymul_host(x, y), ref)
        assert np.array_equal(m.polymul_host(x[: N3], y[: N3]), ref[: N3])  # fewer polynomials than GPUs
        m.close()


def test_two_multi_handles_from_two_threads(qt, oracle):
    """no process-global state: two callers, two handles, concurrently"""
    out = {}

    def work(tag, s, seed):
        p = qt.get_params(s)
        x, y = rand_pair(p.q, 257 * p.n, seed)
        m = qt.MultiEngine(s, 0)
        out[tag] = (np.array_equal(m.polymul_host(x, y), oracle.polymul(s, x, y)), )
        m.close()

    ts = [threading.Thread(target=work, args=("a", qt.SET_III, 1)), threading.Thread(target=work, args=("b", qt.SET_P_I, 2))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert out["a"][0] and out["b"][0]


def test_one_shot_multi_keeps_no_state(qt, oracle):
    import torch
    x, y = rand_pair(Q3, 64 * N3, 5)
    free0 = None
    for i in range(3):
        assert np.array_equal(qt.polymul_host_multi(qt.SET_III, x, y, 0), oracle.polymul(1, x, y))
        torch.cuda.synchronize()
        if i == 0:
            free0 = torch.cuda.mem_get_info()[0]
    assert torch.cuda.mem_get_info()[0] >= free0 - (32 << 20), "the one-shot form must release its contexts"


# ---- smaller entry points -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_pointwise_accepts_unaligned_views(qt, oracle, s):
    """a 4-byte aligned view into a larger array takes the word-wise kernel (no misaligned 128-bit access, which
    would poison the CUDA context)"""
    import torch
    e = qt.Engine(s, 0)
    try:
        B = 5
        a, b = rand_pair(e.q, B * e.n, 60 + s)
        ref = oracle.pointwise(s, a, b)
        big = [torch.zeros(B * e.n + 8, dtype=torch.int32, device="cuda") for _ in range(3)]
        for off in ((1, 0, 0), (0, 3, 0), (0, 0, 1), (1, 2, 3)):
            va, vb, vc = (t[o: o + B * e.n] for t, o in zip(big, off))
            va.copy_(torch.from_numpy(a.view(np.int32)))
            vb.copy_(torch.from_numpy(b.view(np.int32)))
            torch.cuda.synchronize()
            e.pointwise(va, vb, vc, B)
            e.synchronize()
            assert np.array_equal(vc.cpu().numpy().view(np.uint32), ref), off
    finally:
        e.close()


def test_host_placement_queries(qt):
    pci = qt.numa.gpu_pci_bus_id(0)
    assert len(pci) >= 12 and pci.count(":") == 2
    node = qt.numa.gpu_numa_node(0)
    assert node >= -1
    before = len(os.sched_getaffinity(0))
    t = threading.Thread(target=lambda: qt.numa.bind_to_gpu_node(0))  # binds the CALLING thread only
    t.start()
    t.join()
    assert len(os.sched_getaffinity(0)) == before


# ---- sparse / small-operand path: Z_q operands through the ring 2^32-1 with a signed lift ----
CHALLENGE_WEIGHT = {0: 30, 1: 48, 2: 25, 3: 40}  # qTESLA's h per parameter set


def ternary(rng, q, n, B, h):
    y = np.zeros(B * n, np.uint32)
    for b in range(B):
        pos = rng.choice(n, h, replace=False)
        y[b * n + pos] = np.where(rng.integers(0, 2, h) == 1, 1, q - 1).astype(np.uint32)
    return y


def run_lift(qt, e, x, y):
    import torch
    tx = torch.from_numpy(x.view(np.int32)).cuda()
    ty = torch.from_numpy(y.view(np.int32)).cuda()
    tz = torch.empty_like(tx)
    torch.cuda.synchronize()
    e.nussbaumer(tx, ty, tz, qt.RING_2P32M1_LIFT_Q)
    e.synchronize()
    return tz.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_ring_lift_equals_ntt_product_under_its_precondition(qt, oracle, s):
    e = qt.Engine(s, 0)
    try:
        n, q = e.n, e.q
        rng = np.random.default_rng(300 + s)
        B = 41
        h = CHALLENGE_WEIGHT[s]
        # (a) qTESLA's own shape: small secret / error polynomial (|s_i| <= 2^12, stored mod q) x weight-h ternary challenge
        small = rng.integers(-(1 << 12), (1 << 12) + 1, B * n)
        xs = np.where(small < 0, small + q, small).astype(np.uint32)
        c = ternary(rng, q, n, B, h)
        ref = oracle.polymul(s, xs, c)
        assert np.array_equal(run_lift(qt, e, xs, c), ref)
        assert np.array_equal(run_lift(qt, e, c, xs), ref)            # either operand order
        assert np.array_equal(e.polymul_np(xs, c), ref)               # and the NTT path agrees
        # (b) UNIFORM x ternary, the weight the precondition allows: h * (q-1)/2 < 2^31
        hu = min(h, ((1 << 31) - 1) // ((q - 1) // 2))
        assert hu == {0: 30, 1: 48, 2: 12, 3: 5}[s]
        xu = rng.integers(0, q, B * n, dtype=np.uint32)
        xu[:n] = (q - 1) // 2          # extreme centred magnitudes: +-(q-1)/2 everywhere
        xu[n: 2 * n] = (q + 1) // 2
        cu = ternary(rng, q, n, B, hu)
        assert np.array_equal(run_lift(qt, e, xu, cu), oracle.polymul(s, xu, cu))
        # zero results and the two representations of zero in the ring
        z0 = run_lift(qt, e, np.zeros(2 * n, np.uint32), cu[: 2 * n])
        assert not z0.any()
    finally:
        e.close()


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_ring_lift_guard_full_range_operands_are_outside_the_precondition(qt, oracle, s):
    """two uniform operands: integer coefficients reach n q^2 / 4 >> 2^31, the ring wraps, and the lifted result is a
    canonical residue that is NOT the Z_q product — the documented failure mode, not an accident to be relied on"""
    e = qt.Engine(s, 0)
    try:
        x, y = rand_pair(e.q, 3 * e.n, 77 + s)
        z = run_lift(qt, e, x, y)
        assert z.max() < e.q
        ref = oracle.polymul(s, x, y)
        assert (z != ref).sum() > e.n  # essentially every coefficient differs
    finally:
        e.close()


# ---- CUDA graphs for launch-bound batches ------------------------------------------------------------------------
@pytest.mark.parametrize("overlap", [0, 1, 2])
@pytest.mark.parametrize("s", [0, 1, 3])
def test_graph_replay_of_a_dependent_chain(qt, oracle, s, overlap):
    """z0 = x*y, z_{k+1} = z_k * y, ... captured once, replayed twice; every node consumes its predecessor's output
    (with programmatic dependent launch on, off and forced)"""
    import torch
    e = qt.Engine(s, 0)
    try:
        e.set_launch_overlap(overlap)
        n, q, B, K = e.n, e.q, 97, 6
        x, y = rand_pair(q, B * n, 500 + s)
        tx = torch.from_numpy(x.view(np.int32)).cuda()
        ty = torch.from_numpy(y.view(np.int32)).cuda()
        bufs = [torch.empty_like(tx) for _ in range(2)]
        torch.cuda.synchronize()
        l0 = e.launch_count()
        e.graph_begin()
        src = tx
        for k in range(K):
            e.polymul(src, ty, bufs[k & 1], B)
            src = bufs[k & 1]
        w = bufs[(K - 1) & 1]
        e.ntt_forward(w, B)
        e.ntt_inverse(w, B)
        g = e.graph_end()
        assert e.launch_count() == l0, "recording must not count (or run) launches"
        assert e.graph_kernel_count(g) == K + 2
        ref = x
        for k in range(K):
            ref = oracle.polymul(s, ref, y, threads=0)
        for rep in range(2):
            bufs[0].zero_(); bufs[1].zero_()
            torch.cuda.synchronize()
            e.graph_launch(g)
            e.synchronize()
            assert np.array_equal(w.cpu().numpy().view(np.uint32), ref), rep
        assert e.launch_count() == l0 + 2 * (K + 2)
        e.graph_destroy(g)
        # the context works normally after a capture
        assert np.array_equal(e.polymul_np(x, y), oracle.polymul(s, x, y, threads=0))
    finally:
        e.close()


def test_graph_api_misuse_is_an_error(qt):
    e = qt.Engine(1, 0)
    try:
        with pytest.raises(qt.QtError):
            e.graph_end()            # not recording
        e.graph_begin()
        with pytest.raises(qt.QtError):
            e.graph_begin()          # already recording
        g = e.graph_end()            # an empty graph is legal
        assert e.graph_kernel_count(g) == 0
        e.graph_launch(g)
        e.synchronize()
        e.graph_destroy(g)
    finally:
        e.close()


# ---- guard-band check (compute-sanitizer is closed on this pool): no entry point writes outside its operands -------
@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_no_write_outside_the_output_arrays(qt, s):
    """every device-pointer entry point on ragged batches, outputs embedded in sentinel-filled guard bands (also right
    after the LAST polynomial of a partial tile — where a mis-sized bulk copy or an unmasked store would land)"""
    import torch
    e = qt.Engine(s, 0)
    try:
        n, GUARD, SENT = e.n, 4096, 0x5A5A5A5A
        for B in (1, 3, 17, 149, 297, 1777):
            words = B * n

            def guarded():
                t = torch.full((words + 2 * GUARD,), SENT, dtype=torch.int32, device="cuda")
                return t, t[GUARD: GUARD + words]

            def intact(t):
                return bool((t[:GUARD] == SENT).all()) and bool((t[GUARD + words:] == SENT).all())
            x = torch.empty(words, dtype=torch.int32, device="cuda")
            y = torch.empty_like(x)
            e.fill_uniform(x, 1, 0)
            e.fill_uniform(y, 2, 0)
            e.synchronize()
            variants = (0, 1, 2) + ((3, 4) if s == 3 else ())
            for v in variants:
                e.set_fused_variant(v)
                zt, z = guarded()
                e.polymul(x, y, z, B)
                e.synchronize()
                assert intact(zt), ("polymul", v, B)
            e.set_fused_variant(0)
            for name, fn in (("ntt_forward", e.ntt_forward), ("ntt_inverse", e.ntt_inverse),
                             ("ntt_forward_natural", e.ntt_forward_natural), ("ntt_inverse_natural", e.ntt_inverse_natural)):
                wt, w = guarded()
                w.copy_(x)
                torch.cuda.synchronize()
                fn(w, B)
                e.synchronize()
                assert intact(wt), (name, B)
            wt, w = guarded()
            e.pointwise(x, y, w, B)
            e.synchronize()
            assert intact(wt), ("pointwise", B)
            wt, w = guarded()
            e.bitrev_copy(x, w, B)
            e.synchronize()
            assert intact(wt), ("bitrev_copy", B)
            ah = x[:n].clone()
            e.ntt_forward(ah, 1)
            for bc, a in ((True, ah), (False, x)):
                wt, w = guarded()
                e.polymul_ntt(a, y, w, bc, B)
                e.synchronize()
                assert intact(wt), ("polymul_ntt", bc, B)
            for ring in (qt.RING_2P32M1, qt.RING_MODQ, qt.RING_2P32M1_LIFT_Q):
                for nv in (0, 1, 2, 16, 18):
                    e.set_nussbaumer_variant(nv)
                    wt, w = guarded()
                    e.nussbaumer(x, y, w, ring, B)
                    e.synchronize()
                    assert intact(wt), ("nussbaumer", ring, nv, B)
            e.set_nussbaumer_variant(0)
            wt, w = guarded()
            e.fill_uniform(w, 3, 0)
            e.synchronize()
            assert intact(wt), ("fill_uniform", B)
    finally:
        e.close()


def test_graph_with_every_recordable_entry_point(qt, oracle):
    """one graph holding a transform round trip, a pointwise product, a bit-reverse copy, a Nussbaumer product over Z_q,
    the ring lift and a cached-transform product; replayed and compared with the oracle"""
    import torch
    s = 1
    e = qt.Engine(s, 0)
    try:
        n, q, B = e.n, e.q, 53
        rng = np.random.default_rng(9)
        x, y = rand_pair(q, B * n, 91)
        c = ternary(rng, q, n, B, 48)
        dev = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
        tx, ty, tc = dev(x), dev(y), dev(c)
        w, pw, br, nz, lf, ca = (torch.empty_like(tx) for _ in range(6))
        ah = tx[:n].clone()
        torch.cuda.synchronize()
        e.graph_begin()
        e.ntt_forward(ah, 1)
        e.polymul_ntt(ah, ty, ca, True, B)      # ca = x[0] * y[b]
        e.pointwise(tx, ty, pw, B)
        e.bitrev_copy(tx, br, B)
        e.nussbaumer(tx, ty, nz, qt.RING_MODQ, B)
        e.nussbaumer(tx, tc, lf, qt.RING_2P32M1_LIFT_Q, B)
        e.polymul(tx, ty, w, B)
        e.ntt_forward(w, B)
        e.ntt_inverse(w, B)
        g = e.graph_end()
        assert e.graph_kernel_count(g) == 9
        for _ in range(2):
            ah.copy_(tx[:n])
            torch.cuda.synchronize()
            e.graph_launch(g)
            e.synchronize()
        host = lambda t: t.cpu().numpy().view(np.uint32)
        ref = oracle.polymul(s, x, y, threads=0)
        assert np.array_equal(host(w), ref) and np.array_equal(host(nz), ref)
        assert np.array_equal(host(pw), oracle.pointwise(s, x, y))
        assert np.array_equal(host(br), oracle.bitrev_copy(s, x))
        assert np.array_equal(host(lf), oracle.polymul(s, x, c, threads=0))
        assert np.array_equal(host(ca), oracle.polymul(s, np.tile(x[:n], B), y, threads=0))
        e.graph_destroy(g)
    finally:
        e.close()


def test_ring_lift_through_the_host_pointer_form(qt, oracle):
    e = qt.Engine(0, 0)
    try:
        rng = np.random.default_rng(12)
        B = 700
        small = rng.integers(-2000, 2001, B * e.n)
        xs = np.where(small < 0, small + e.q, small).astype(np.uint32)
        c = ternary(rng, e.q, e.n, B, 30)
        z = e.nussbaumer_host(xs, c, ring=qt.RING_2P32M1_LIFT_Q)
        assert np.array_equal(z, oracle.polymul(0, xs, c, threads=0))
    finally:
        e.close()
