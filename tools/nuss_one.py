#!/usr/bin/env python
"""One Nussbaumer configuration, a few launches (the program ncu captures):  python tools/nuss_one.py <set> <ring> <variant> [batch]
   set: I | III | p-I | p-III;  ring: 0 = 2^32-1, 1 = Z_q, 2 = lift;  variant: 0 auto, 1 schoolbook, 2 recursive, 3 FP64 rows"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qtesla_b200_loader import load
qt = load()
s = {"I": 0, "III": 1, "p-I": 2, "p-III": 3}[sys.argv[1]]
ring, variant = int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else {0: 65536, 1: 65536, 2: 65536, 3: 32768}[s]
eng = qt.Engine(s, 0)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
eng.set_nussbaumer_variant(variant)
x = torch.empty(B * eng.n, dtype=torch.int32, device="cuda"); y = torch.empty_like(x); z = torch.empty_like(x)
with torch.cuda.stream(stream):
    eng.fill_uniform(x, 1, 0); eng.fill_uniform(y, 2, 0)
    for _ in range(2):
        eng.nussbaumer(x, y, z, ring, B)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        eng.nussbaumer(x, y, z, ring, B)
    e1.record(stream)
e1.synchronize()
print(f"{sys.argv[1]} ring {ring} variant {variant}: {B * 3 / (e0.elapsed_time(e1) * 1e-3) / 1e6:.2f} M polymul/s")
eng.close()
