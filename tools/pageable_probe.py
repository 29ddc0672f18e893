#!/usr/bin/env python
"""qt_polymul_host with pageable (malloc'd, what the reference harness passes) vs pinned host buffers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qtesla_b200_loader import load
qt = load()
eng = qt.Engine(1, 0)
B, n = 65536, 1024
rng = np.random.default_rng(1)
x = rng.integers(0, eng.q, B * n, dtype=np.uint32); y = rng.integers(0, eng.q, B * n, dtype=np.uint32)
z = np.empty_like(x)
def timeit(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps
t = timeit(lambda: eng.polymul_host(x, y, z))
print(f"pageable: {t*1e3:.2f} ms  {B/t/1e6:.2f} M polymul/s  {3*B*n*4/t/1e9:.1f} GB/s")
zp = z.copy()
px = torch.from_numpy(x).pin_memory() if False else torch.empty(B * n, dtype=torch.int32).pin_memory()
py = torch.empty(B * n, dtype=torch.int32).pin_memory(); pz = torch.empty(B * n, dtype=torch.int32).pin_memory()
px.numpy().view(np.uint32)[:] = x; py.numpy().view(np.uint32)[:] = y
ax, ay, az = px.numpy().view(np.uint32), py.numpy().view(np.uint32), pz.numpy().view(np.uint32)
t = timeit(lambda: eng.polymul_host(ax, ay, az))
print(f"pinned:   {t*1e3:.2f} ms  {B/t/1e6:.2f} M polymul/s  {3*B*n*4/t/1e9:.1f} GB/s")
print("equal:", np.array_equal(zp, az))
