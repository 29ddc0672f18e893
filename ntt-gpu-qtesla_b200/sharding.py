"""Batch partitioning across GPUs: polynomials are independent, so GPU g of G owns the
contiguous slice [g*B/G, (g+1)*B/G) of x, y and z — the same rule qt_polymul_host_multi applies in
C++ (csrc/qt_capi.cu).  No data-path collective exists; ranks only agree on timing."""


def shard_bounds(batch, rank, world):
    """[lo, hi) polynomial range of `rank` when `batch` polynomials are split over `world` ranks."""
    if world < 1 or not 0 <= rank < world or batch < 0:
        raise ValueError("bad shard request")
    return batch * rank // world, batch * (rank + 1) // world


def weak_scaling_slice(per_gpu_batch, n, rank):
    """Weak scaling (bench.py): every rank owns `per_gpu_batch` polynomials; returns the rank's offset
    (in coefficients) into the global synthetic stream a[i] = splitmix64(seed + i) % q."""
    return rank * per_gpu_batch * n


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over all ranks (identity for a single process)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
