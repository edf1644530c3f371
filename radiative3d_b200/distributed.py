"""One process per GPU: phonon index ranges per rank, the model broadcast and the single end-of-run reduction.

Phonons are independent and phonon i always uses the draw stream keyed by (seed, i), so a run is sharded by
contiguous index ranges with no data-path collective; the per-rank seismometer bins and loss
counters are then summed once -- what the reference does with files and vis/seisplot/combine.m:26-33 --
by one reduction (NCCL over NVLink on GPUs; gloo in the CPU tests): TWO collectives, one per element type
(f64 energies; i64 counts + counters + diagnostic-bit lanes), and no host synchronisation.
"""
import numpy as np

from . import abi

DIAG = abi.R3D_CNT_DIAG   # counters[7] is a bit mask (OR), the others are sums
LANES = abi.R3D_NDIAG_LANES


def shard_range(first, n, rank, world):
    """Contiguous share of phonon indices [first, first+n) for `rank` of `world` -> (first_r, n_r)."""
    lo = n // world * rank + (n % world) * rank // world
    hi = n // world * (rank + 1) + (n % world) * (rank + 1) // world
    return first + lo, hi - lo


def _rebuild_diag(i64_block, counters_at):
    """counters[DIAG] = OR over b of (lane[b] != 0) << b, as tensor arithmetic (no host round trip)."""
    import torch
    lanes = i64_block[counters_at + abi.R3D_NCOUNTERS: counters_at + abi.R3D_NCOUNTERS + LANES]
    weights = torch.tensor([1 << b for b in range(LANES)], dtype=torch.int64, device=i64_block.device)
    i64_block[counters_at + DIAG] = ((lanes != 0).to(torch.int64) * weights).sum()


def all_reduce_blocks(f64_block, i64_block, counters_at, group=None, dst=None):
    """In-place sum over ranks of the two accumulator blocks of a handle (Engine.device_accumulator_blocks):
    f64_block = energies; i64_block = counts | counters[8] | diag lanes[8].  Two collectives; the diagnostic word is
    rebuilt from its lanes on the device.  dst=None: all-reduce (every rank ends with the sums); dst=r: reduce to rank
    r only (what a launcher that writes the output files on one rank needs)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if dst is None:
        dist.all_reduce(f64_block, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(i64_block, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(f64_block, dst=dst, op=dist.ReduceOp.SUM, group=group)
        dist.reduce(i64_block, dst=dst, op=dist.ReduceOp.SUM, group=group)
    _rebuild_diag(i64_block, counters_at)


def all_reduce_results(energies, counts, counters, group=None):
    """In-place sum of separately held bins and counters over ranks (torch tensors, device or host): energies f64,
    counts i64, counters i64[8] (counters[7] is OR-ed).  Packs the integers into one buffer with the diagnostic word's
    bit lanes, so it is two collectives as well; use all_reduce_blocks on a handle's own blocks to avoid the packing."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    weights = torch.tensor([1 << b for b in range(LANES)], dtype=torch.int64, device=counters.device)
    lanes = ((counters[DIAG] & weights) != 0).to(torch.int64)
    packed = torch.cat([counts.reshape(-1), counters.reshape(-1), lanes])
    dist.all_reduce(energies, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    n = counts.numel()
    counts.copy_(packed[:n].reshape(counts.shape))
    counters.copy_(packed[n:n + abi.R3D_NCOUNTERS])
    _rebuild_diag(packed, n)
    counters[DIAG] = packed[n + DIAG]


def all_reduce_numpy(energies, counts, counters, group=None):
    """Same for host numpy arrays (used by the gloo tests and by hosts that fetched already)."""
    import torch
    e = torch.from_numpy(energies)
    c = torch.from_numpy(counts.view(np.int64))
    k = torch.from_numpy(counters.view(np.int64))
    all_reduce_results(e, c, k, group)


# the five large tables of a model: the ones r3d_create accepts in device memory (include/r3d_gpu.h)
BIG_TABLES = ("toa_theta", "toa_phi", "src_cdf", "scat_cdf", "scat_spol")


def broadcast_model_tables(model, device, src=0, group=None):
    """The large tables of `model` cross PCIe once per node: rank `src` uploads them (one pinned H2D copy each), every
    rank receives them by broadcast (NCCL over NVLink) into ONE device buffer.  Returns {name: (device pointer, tensor)}
    for Engine(model, device_tables=...).  Every rank must hold a model of the same shapes (the small arrays and scalars
    are host data and are not sent)."""
    import torch
    import torch.distributed as dist
    sizes = [int(getattr(model, n).size) for n in BIG_TABLES]
    total = sum(sizes)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    buf = torch.empty(total, dtype=torch.float64, device=device)
    if rank == src:
        at = 0
        for n, sz in zip(BIG_TABLES, sizes):
            a = getattr(model, n)
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(a)
            buf[at:at + sz].copy_(t.reshape(-1), non_blocking=True)
            at += sz
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(buf, src=src, group=group)
    out, at = {}, 0
    for n, sz in zip(BIG_TABLES, sizes):
        t = buf[at:at + sz]
        out[n] = (t.data_ptr(), t)
        at += sz
    return out
