#!/usr/bin/env python
"""End-to-end (host pointers) throughput vs pipeline chunk size (development aid)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from qtesla_b200_loader import load
    qt = load()
    eng = qt.Engine(1, 0)
    B = 65536
    words = B * eng.n
    hx = torch.empty(words, dtype=torch.int32).pin_memory(); hy = torch.empty_like(hx).pin_memory(); hz = torch.empty_like(hx).pin_memory()
    hx.random_(0, eng.q); hy.random_(0, eng.q)
    xh, yh, zh = (t.numpy().view(np.uint32) for t in (hx, hy, hz))
    for _ in range(2): eng.polymul_host(xh, yh, zh, B)
    t0 = time.perf_counter()
    for _ in range(8): eng.polymul_host(xh, yh, zh, B)
    dt = (time.perf_counter() - t0) / 8
    print(f"chunk_words={os.environ.get('QT_PIPE_CHUNK_WORDS','default')}: {dt*1e3:.2f} ms/step, {B/dt/1e6:.2f} M polymul/s, {words*12/dt/1e9:.1f} GB/s PCIe total")
    # raw copy ceilings
    d = torch.empty(words, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): d.copy_(hx, non_blocking=True)
    torch.cuda.synchronize(); h2d = 4 * words * 4 / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter()
    for _ in range(4): hz.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); d2h = 4 * words * 4 / (time.perf_counter() - t0) / 1e9
    print(f"   raw pinned copies: H2D {h2d:.1f} GB/s, D2H {d2h:.1f} GB/s")
else:
    for slots in ("2", "3", "4", "6"):
        for cw in ("2097152", "4194304", "8388608"):
            env = dict(os.environ, QT_PIPE_CHUNK_WORDS=cw, QT_PIPE_SLOTS=slots)
            print("slots", slots, end=" ", flush=True)
            subprocess.run([sys.executable, __file__, "child"], env=env)
