#!/usr/bin/env python
"""Quick A/B timing of the fused kernel variants on one GPU (development aid, not the bench).
   python tools/ab.py [--sets III,I,p-I,p-III] [--steps 50]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qtesla_b200_loader import load

ap = argparse.ArgumentParser()
ap.add_argument("--sets", default="III,I,p-I,p-III")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--variants", default="1,2")
ap.add_argument("--nuss", action="store_true")
args = ap.parse_args()
qt = load()
SETS = {"I": 0, "III": 1, "p-I": 2, "p-III": 3}
BATCH = {0: 65536, 1: 65536, 2: 65536, 3: 32768}
stream = torch.cuda.Stream()
for name in args.sets.split(","):
    s = SETS[name]
    eng = qt.Engine(s, 0)
    eng.set_stream(stream.cuda_stream)
    B = BATCH[s]
    x = torch.empty(B * eng.n, dtype=torch.int32, device="cuda"); y = torch.empty_like(x); z = torch.empty_like(x)
    with torch.cuda.stream(stream):
        eng.fill_uniform(x, 1, 0); eng.fill_uniform(y, 2, 0)
    def timeit(fn, steps):
        with torch.cuda.stream(stream):
            for _ in range(5): fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps): fn()
            e1.record(stream)
        e1.synchronize()
        return B * steps / (e0.elapsed_time(e1) * 1e-3)
    for v in [int(t) for t in args.variants.split(",")]:
        try:
            eng.set_fused_variant(v)
            r = timeit(lambda: eng.polymul(x, y, z, B), args.steps)
            print(f"{name:6s} variant {v}: {r/1e6:8.2f} M polymul/s  info={eng.kernel_info()}", flush=True)
        except Exception as ex:
            print(f"{name:6s} variant {v}: {ex}")
    try:
        ah = torch.empty(eng.n, dtype=torch.int32, device="cuda")
        with torch.cuda.stream(stream):
            eng.fill_uniform(ah, 3, 0); eng.ntt_forward(ah)
        r = timeit(lambda: eng.polymul_ntt(ah, y, z, True, B), args.steps)
        print(f"{name:6s} cached a_hat (broadcast): {r/1e6:8.2f} M polymul/s", flush=True)
        xh = x.clone()
        with torch.cuda.stream(stream):
            eng.ntt_forward(xh, B)
        r = timeit(lambda: eng.polymul_ntt(xh, y, z, False, B), args.steps)
        print(f"{name:6s} cached a_hat (per product): {r/1e6:8.2f} M polymul/s", flush=True)
        del xh
    except Exception as ex:
        print(f"{name:6s} cached: {ex}")
    w = x.clone()
    for nm, fn in (("ntt_forward", lambda: eng.ntt_forward(w, B)), ("ntt_inverse", lambda: eng.ntt_inverse(w, B)),
                   ("fwd_natural", lambda: eng.ntt_forward_natural(w, B)), ("inv_natural", lambda: eng.ntt_inverse_natural(w, B)),
                   ("pointwise", lambda: eng.pointwise(x, y, z, B)), ("bitrev_copy", lambda: eng.bitrev_copy(x, z, B))):
        r = timeit(fn, args.steps)
        print(f"{name:6s} {nm:12s}: {r/1e6:8.2f} M poly/s  ({r*eng.n*8/1e9:7.1f} GB/s r+w)", flush=True)
    del w
    if args.nuss:
        for ring in (0, 1):
            try:
                r = timeit(lambda: eng.nussbaumer(x, y, z, ring, B), 5)
                print(f"{name:6s} nussbaumer ring {ring}: {r/1e6:8.2f} M polymul/s", flush=True)
            except Exception as ex:
                print(f"{name:6s} nussbaumer ring {ring}: {ex}")
    eng.close()
