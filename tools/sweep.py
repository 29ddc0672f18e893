#!/usr/bin/env python
"""Batch sweep (BASELINE.json configs[4]): n=1024, total batch B in {2^10 .. 2^22} sharded contiguously
over the ranks (strong scaling: the TOTAL batch is fixed, each rank owns B/world polynomials), device
resident, CUDA events, max over ranks.  One JSON line per batch size on rank 0.

  python tools/sweep.py                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from qtesla_b200_loader import load

_OUT = os.dup(1)   # JSON lines go to the real stdout; library chatter (NCCL banner) to stderr
os.dup2(2, 1)
ap = argparse.ArgumentParser()
ap.add_argument("--set", default="III")
ap.add_argument("--min-log2", type=int, default=10)
ap.add_argument("--max-log2", type=int, default=22)
args = ap.parse_args()
qt = load()
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
set_id = {"I": 0, "III": 1, "p-I": 2, "p-III": 3}[args.set]
eng = qt.Engine(set_id, local)
stream = torch.cuda.Stream(device=dev)
eng.set_stream(stream.cuda_stream)
n = eng.n
for lg in range(args.min_log2, args.max_log2 + 1):
    B = 1 << lg
    lo, hi = qt.sharding.shard_bounds(B, rank, world)
    b = hi - lo
    words = max(b, 1) * n
    x = torch.empty(words, dtype=torch.int32, device=dev); y = torch.empty_like(x); z = torch.empty_like(x)
    steps = max(5, min(200, (1 << 24) // max(b, 1)))
    with torch.cuda.stream(stream):
        eng.fill_uniform(x, 1, lo * n); eng.fill_uniform(y, 2, lo * n)
        for _ in range(3):
            eng.polymul(x, y, z, b)
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            eng.polymul(x, y, z, b)
        e1.record(stream)
    e1.synchronize()
    ms = qt.sharding.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None, dev)
    if rank == 0:
        line = {"param_set": args.set, "n": n, "total_batch": B, "n_gpus": world, "per_gpu_batch": b, "steps": steps,
                "us_per_step": ms / steps * 1e3, "polymuls_per_s": B * steps / (ms * 1e-3),
                "working_set_MiB": 3 * b * n * 4 / 2 ** 20}
        os.write(_OUT, (json.dumps(line) + "\n").encode())
    del x, y, z
eng.close()
if world > 1:
    dist.destroy_process_group()
