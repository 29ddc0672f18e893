#!/usr/bin/env python
"""unfused entry points + cached-transform product, device-resident, for one library build (QT_LIB_PATH)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qtesla_b200_loader import load
qt = load()
st = torch.cuda.Stream()
for s, B in ((1, 65536), (2, 65536), (0, 65536)):
    e = qt.Engine(s, 0); e.set_stream(st.cuda_stream)
    n = e.n
    x = torch.empty(B * n, dtype=torch.int32, device="cuda"); y = torch.empty_like(x); w = torch.empty_like(x)
    ah = torch.empty(n, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(st):
        e.fill_uniform(x, 1, 0); e.fill_uniform(y, 2, 0); e.fill_uniform(ah, 3, 0); e.ntt_forward(ah, 1); w.copy_(x)
    def timed(fn, reps=20):
        with torch.cuda.stream(st):
            for _ in range(3): fn()
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(st)
            for _ in range(reps): fn()
            a1.record(st)
        a1.synchronize()
        return a0.elapsed_time(a1) / reps
    out = []
    for nm, fn in (("fwd", lambda: e.ntt_forward(w, B)), ("inv", lambda: e.ntt_inverse(w, B)),
                   ("ntt_bcast", lambda: e.polymul_ntt(ah, y, w, True, B)), ("ntt_each", lambda: e.polymul_ntt(x, y, w, False, B))):
        t = timed(fn)
        out.append(f"{nm} {B / (t * 1e-3) / 1e6:.1f} M/s")
    print(os.environ.get("QT_LIB_PATH", "main").split("/")[-2] if "QT_LIB_PATH" in os.environ else "main", "set", s, " | ".join(out), flush=True)
    e.close()
