#!/usr/bin/env python
"""Small all-kernel workload for compute-sanitizer (memcheck / racecheck / synccheck), one tool per run:
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from qtesla_b200_loader import load
qt = load()
for s in (0, 1, 2, 3):
    eng = qt.Engine(s, 0)
    B = 37
    x = torch.empty(B * eng.n, dtype=torch.int32, device="cuda"); y = torch.empty_like(x); z = torch.empty_like(x)
    eng.fill_uniform(x, 1, 0); eng.fill_uniform(y, 2, 0)
    for v in (1, 2):
        eng.set_fused_variant(v); eng.polymul(x, y, z)
    eng.set_fused_variant(0)
    w = x.clone(); eng.ntt_forward(w); eng.polymul_ntt(w[: eng.n], y, z, True); eng.polymul_ntt(w, y, z, False)
    eng.ntt_inverse(w); eng.pointwise(x, y, z); eng.bitrev_copy(x, z)
    eng.nussbaumer(x, y, z, qt.RING_MODQ)
    if eng.n != 2048:
        eng.nussbaumer(x, y, z, qt.RING_2P32M1)
    eng.synchronize()
    assert torch.equal(w, x)
    eng.close()
print("sanitize_small: all kernels ran")
