"""CPU test of the N>1 host logic (gloo, world size 2): contiguous batch sharding with no data-path
collective, per-rank slices of the synthetic stream, and the max-over-ranks timing reduction that
bench.py uses.  The per-shard products are checked with the oracle — the GPU is not involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from qtesla_b200_loader import load
    from oracle_lib import Oracle
    qt = load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    s, B = 1, 37                                   # odd batch: uneven shards
    p = o.params(s)
    lo, hi = qt.sharding.shard_bounds(B, rank, world)
    # every rank regenerates ONLY its slice of the global stream and multiplies it
    x = o.splitmix(1, lo * p.n, p.q, (hi - lo) * p.n)
    y = o.splitmix(2, lo * p.n, p.q, (hi - lo) * p.n)
    z = o.polymul(s, x, y)
    np.save(os.path.join(out_dir, f"z{rank}.npy"), z)
    np.save(os.path.join(out_dir, f"b{rank}.npy"), np.array([lo, hi]))
    # the only communication of the path: barrier + max of the timings
    dist.barrier()
    t = qt.sharding.max_over_ranks(1.0 + rank, dist)
    assert t == float(world)
    # weak-scaling stream offsets are disjoint and contiguous
    assert qt.sharding.weak_scaling_slice(8, p.n, rank) == rank * 8 * p.n
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    s, B = 1, 37
    p = oracle.params(s)
    bounds = [np.load(tmp_path / f"b{r}.npy") for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == B and bounds[0][1] == bounds[1][0]   # contiguous cover
    z = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)])
    x = oracle.splitmix(1, 0, p.q, B * p.n)
    y = oracle.splitmix(2, 0, p.q, B * p.n)
    assert np.array_equal(z, oracle.polymul(s, x, y))     # sharded result == unsharded result


def test_shard_bounds_properties(qt):
    for B in (0, 1, 7, 65536, 4 * 1024 * 1024 + 3):
        for G in (1, 2, 3, 4, 8):
            cuts = [qt.sharding.shard_bounds(B, g, G) for g in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
