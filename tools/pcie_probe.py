#!/usr/bin/env python
"""Host<->device copy ceiling of the box and the end-to-end hot path beside it, one process per GPU.

  python tools/pcie_probe.py                                                    # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank moves, per step, what one end-to-end step of the bench moves (qTESLA-III, batch 65 536: 512 MiB
host->device for x and y, 256 MiB device->host for z, both directions in flight), first as BARE copies (no
kernel: the ceiling), then through qt_polymul_host (the `e2e` figure of bench.py).  Ranks meet at a gloo
barrier, the time of a case is the max over ranks.  One JSON line per case on rank 0; the last line is the
summary {"n_gpus", "ceiling_steps_per_s", "e2e_steps_per_s", "e2e_over_ceiling"}.

--bind binds each rank to the CPUs of its GPU's NUMA node (sysfs) before any host buffer is allocated.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_OUT = os.dup(1)
os.dup2(2, 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--bind", dest="numa", action="store_true", help="bind each rank to its GPU NUMA node first")
ap.add_argument("--skip-e2e", action="store_true")
ap.add_argument("--tag", default="")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")


def emit(obj):
    if rank == 0:
        os.write(_OUT, (json.dumps(obj) + "\n").encode())


def barrier():
    if world > 1:
        dist.barrier()


def max_ranks(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather(v):
    if world == 1:
        return [v]
    out = [None] * world
    dist.all_gather_object(out, v)
    return out


torch.cuda.set_device(local)
from qtesla_b200_loader import load  # noqa: E402
qt = load()
numa = qt.numa.bind_to_gpu_node(local) if args.numa else None
info = gather({"rank": rank, "gpu": local, "pci": qt.numa.gpu_pci_bus_id(local), "numa_node": qt.numa.gpu_numa_node(local),
               "bound_cpus": numa, "affinity": len(os.sched_getaffinity(0))})
emit({"test": "ranks", "n_gpus": world, "numa_bind": bool(args.numa), "ranks": info, "tag": args.tag})

MB = 1 << 20
IN, OUT = 512 * MB, 256 * MB
h_in = torch.empty(IN, dtype=torch.uint8).pin_memory()
h_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(1)
d_in = torch.empty(IN, dtype=torch.uint8, device="cuda")
d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def bare(h2d, d2h, chunk):
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                for o in range(0, IN, chunk):
                    d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for o in range(0, OUT, chunk // 2):
                    h_out[o:o + chunk // 2].copy_(d_out[o:o + chunk // 2], non_blocking=True)
        torch.cuda.synchronize()
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        once()
    mine = time.perf_counter() - t0
    dt = max_ranks(mine)
    nbytes = (IN if h2d else 0) + (OUT if d2h else 0)
    per = gather(round(nbytes * args.reps / mine / 1e9, 1))
    emit({"test": "bare_copies", "n_gpus": world, "dir": "both" if h2d and d2h else "h2d" if h2d else "d2h", "chunk_MiB": chunk // MB,
          "ms_per_step": dt / args.reps * 1e3, "GBs_total": nbytes * args.reps * world / dt / 1e9,
          "steps_per_s_total": world * args.reps / dt, "per_gpu_GBs": per, "numa_bind": bool(args.numa), "tag": args.tag})
    return world * args.reps / dt


ceil16 = bare(True, True, 16 * MB)
ceil512 = bare(True, True, 512 * MB)
bare(True, False, 512 * MB)
bare(False, True, 512 * MB)
ceiling = max(ceil16, ceil512)

e2e = None
if not args.skip_e2e:
    eng = qt.Engine(1, local)
    B = 65536
    words = B * eng.n
    xh = h_in.numpy().view(np.uint32)[:words]
    yh = h_in.numpy().view(np.uint32)[words:2 * words]
    zh = h_out.numpy().view(np.uint32)[:words]
    rng = np.random.default_rng(7 + rank)
    xh[: 64 * eng.n] = rng.integers(0, eng.q, 64 * eng.n, dtype=np.uint32)
    yh[: 64 * eng.n] = rng.integers(0, eng.q, 64 * eng.n, dtype=np.uint32)
    for _ in range(2):
        eng.polymul_host(xh, yh, zh, B)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        eng.polymul_host(xh, yh, zh, B)
    dt = max_ranks(time.perf_counter() - t0)
    e2e = world * args.reps / dt
    emit({"test": "qt_polymul_host", "n_gpus": world, "ms_per_step": dt / args.reps * 1e3, "polymuls_per_s_total": e2e * B,
          "GBs_total": 768 * MB * e2e / 1e9, "steps_per_s_total": e2e, "chunk_words": os.environ.get("QT_PIPE_CHUNK_WORDS", "default"),
          "slots": os.environ.get("QT_PIPE_SLOTS", "default"), "numa_bind": bool(args.numa), "tag": args.tag})
    eng.close()
emit({"test": "summary", "n_gpus": world, "numa_bind": bool(args.numa), "ceiling_steps_per_s": ceiling,
      "ceiling_polymuls_per_s": ceiling * 65536, "e2e_steps_per_s": e2e,
      "e2e_over_ceiling": (e2e / ceiling) if e2e else None, "tag": args.tag})
if world > 1:
    dist.destroy_process_group()
