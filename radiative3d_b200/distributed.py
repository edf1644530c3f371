"""One process per GPU: phonon index ranges per rank and the single end-of-run reduction.

Phonons are independent and phonon i always uses the draw stream keyed by (seed, i), so a run is sharded by
contiguous index ranges with no data-path collective; the per-rank seismometer bins and loss
counters are then summed once -- what the reference does with files and vis/seisplot/combine.m:26-33 --
by one all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import numpy as np

DIAG = 7   # counters[7] is a bit mask (OR), the others are sums


def shard_range(first, n, rank, world):
    """Contiguous share of phonon indices [first, first+n) for `rank` of `world` -> (first_r, n_r)."""
    lo = n // world * rank + (n % world) * rank // world
    hi = n // world * (rank + 1) + (n % world) * (rank + 1) // world
    return first + lo, hi - lo


def all_reduce_results(energies, counts, counters, group=None):
    """In-place sum of bins and counters over ranks.  Arguments are torch tensors (device or host):
    energies f64, counts i64, counters i64[8] (counters[7] is OR-ed)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    diag = counters[DIAG].clone()
    dist.all_reduce(energies, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    # OR of a bit mask: reduce each of the 7 reason bits with MAX
    bits = torch.stack([(diag >> b) & 1 for b in range(8)])
    dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
    counters[DIAG] = sum(int(bits[b]) << b for b in range(8))


def all_reduce_numpy(energies, counts, counters, group=None):
    """Same for host numpy arrays (used by the gloo tests and by hosts that fetched already)."""
    import torch
    e = torch.from_numpy(energies)
    c = torch.from_numpy(counts.view(np.int64))
    k = torch.from_numpy(counters.view(np.int64))
    all_reduce_results(e, c, k, group)
