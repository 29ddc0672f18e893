#!/usr/bin/env python
"""Raw PCIe ceilings of the box: H2D alone, D2H alone, both at once (pinned memory, large copies)."""
import time, torch
MB = 1 << 20
h_in = torch.empty(512 * MB, dtype=torch.uint8).pin_memory(); d_in = torch.empty(512 * MB, dtype=torch.uint8, device="cuda")
h_out = torch.empty(256 * MB, dtype=torch.uint8).pin_memory(); d_out = torch.empty(256 * MB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1, reps=5):
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                for c in range(chunks):
                    n = 512 * MB // chunks
                    d_in[c*n:(c+1)*n].copy_(h_in[c*n:(c+1)*n], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for c in range(chunks):
                    n = 256 * MB // chunks
                    h_out[c*n:(c+1)*n].copy_(d_out[c*n:(c+1)*n], non_blocking=True)
        torch.cuda.synchronize()
    once()
    t0 = time.perf_counter()
    for _ in range(reps): once()
    return (time.perf_counter() - t0) / reps
for chunks in (1, 16):
    a = run(True, False, chunks); b = run(False, True, chunks); c = run(True, True, chunks)
    print(f"chunks={chunks}: H2D 512MiB {a*1e3:.2f} ms ({512*MB/a/1e9:.1f} GB/s)  D2H 256MiB {b*1e3:.2f} ms ({256*MB/b/1e9:.1f} GB/s)  both {c*1e3:.2f} ms ({768*MB/c/1e9:.1f} GB/s total)")
