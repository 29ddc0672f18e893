#!/usr/bin/env python
"""A/B timing of the Nussbaumer kernels on one GPU (development aid, not the bench): every parameter set, the ring
2^32-1 and Z_q with schoolbook / recursive row products.   python tools/nuss_ab.py [--sets III,I,p-I,p-III] [--steps 5]
QT_LIB_PATH selects an alternative build of the library."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from qtesla_b200_loader import load

ap = argparse.ArgumentParser()
ap.add_argument("--sets", default="III,I,p-I,p-III")
ap.add_argument("--steps", type=int, default=5)
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
    ref = None
    for label, ring, nv in (("ring 2^32-1", 0, 0), ("ring 2^32-1 [whole]", 0, 16), ("ring lift", 2, 0), ("Z_q schoolbook rows", 1, 1), ("Z_q schoolbook [whole]", 1, 17),
                            ("Z_q recursive rows", 1, 2), ("Z_q recursive [whole]", 1, 18),
                            ("Z_q FP64-pipe rows", 1, 3), ("Z_q FP64 [whole]", 1, 19), ("Z_q automatic", 1, 0)):
        try:
            eng.set_nussbaumer_variant(nv)
        except qt.QtError as ex:
            print(f"{name:6s} {label:22s}: {ex}", flush=True)
            continue
        with torch.cuda.stream(stream):
            for _ in range(2): eng.nussbaumer(x, y, z, ring, B)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps): eng.nussbaumer(x, y, z, ring, B)
            e1.record(stream)
        e1.synchronize()
        r = B * args.steps / (e0.elapsed_time(e1) * 1e-3)
        same = ""
        if ring == 1:
            h = z.cpu().numpy()
            if ref is None: ref = h
            else: same = "  identical to schoolbook: %s" % bool(np.array_equal(ref, h))
        elif ring == 0:
            h = z.cpu().numpy()
            if nv == 0: ref0 = h
            else: same = "  identical to block-pass: %s" % bool(np.array_equal(ref0, h))
        print(f"{name:6s} {label:22s}: {r/1e6:8.2f} M polymul/s{same}", flush=True)
    eng.close()
