#!/usr/bin/env python
"""qt_polymul_host (pinned and pageable arrays) at several batch sizes (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qtesla_b200_loader import load
qt = load()
eng = qt.Engine(1, 0)
n = eng.n
for B in (256, 1024, 4096, 16384, 65536):
    words = B * n
    hx = torch.empty(words, dtype=torch.int32).pin_memory(); hy = torch.empty(words, dtype=torch.int32).pin_memory(); hz = torch.empty(words, dtype=torch.int32).pin_memory()
    hx.random_(0, eng.q); hy.random_(0, eng.q)
    xh, yh, zh = (t.numpy().view(np.uint32) for t in (hx, hy, hz))
    xp, yp = xh.copy(), yh.copy(); zp = np.empty_like(xp)
    res = []
    for a, b, c in ((xh, yh, zh), (xp, yp, zp)):
        for _ in range(3): eng.polymul_host(a, b, c, B)
        reps = max(3, min(50, (1 << 18) // B))
        t0 = time.perf_counter()
        for _ in range(reps): eng.polymul_host(a, b, c, B)
        res.append((time.perf_counter() - t0) / reps)
    assert np.array_equal(zh, zp)
    print(f"B={B:6d}: pinned {res[0]*1e6:9.1f} us ({B/res[0]/1e6:5.2f} M/s)   pageable {res[1]*1e6:9.1f} us ({B/res[1]/1e6:5.2f} M/s)")
