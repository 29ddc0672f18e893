#!/usr/bin/env python
"""e2e (pinned host arrays through qt_polymul_host) with / without the ramped chunk schedule; child processes because
the pipeline reads its tuning variables when it is created.   python tools/e2e_ab.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import time
    import numpy as np, torch
    from qtesla_b200_loader import load
    qt = load()
    for s, B in ((1, 65536), (0, 65536), (3, 32768)):
        eng = qt.Engine(s, 0)
        words = B * eng.n
        hx = torch.empty(words, dtype=torch.int32).pin_memory(); hy = torch.empty(words, dtype=torch.int32).pin_memory(); hz = torch.empty(words, dtype=torch.int32).pin_memory()
        hx.random_(0, eng.q); hy.random_(0, eng.q)
        xh, yh, zh = (t.numpy().view(np.uint32) for t in (hx, hy, hz))
        for _ in range(2): eng.polymul_host(xh, yh, zh, B)
        t0 = time.perf_counter()
        for _ in range(10): eng.polymul_host(xh, yh, zh, B)
        dt = (time.perf_counter() - t0) / 10
        d = eng.polymul_np(xh[: 8 * eng.n], yh[: 8 * eng.n])
        ok = bool(np.array_equal(d, zh[: 8 * eng.n])) and bool(np.array_equal(eng.polymul_np(xh[-8 * eng.n:], yh[-8 * eng.n:]), zh[-8 * eng.n:]))
        print(f"roles={os.environ.get('QT_PIPE_ROLES','1')} ramp={os.environ.get('QT_PIPE_RAMP','1')} slots={os.environ.get('QT_PIPE_SLOTS','3')} set {s}: {dt*1e3:.2f} ms/step, {B/dt/1e6:.2f} M polymul/s, {words*12/dt/1e9:.1f} GB/s, ok={ok}", flush=True)
        eng.close()
else:
    for roles, ramp, slots in (("0", "0", "3"), ("0", "1", "3"), ("1", "0", "4"), ("1", "1", "3"), ("1", "1", "4"), ("1", "1", "6")):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, QT_PIPE_ROLES=roles, QT_PIPE_RAMP=ramp, QT_PIPE_SLOTS=slots))
