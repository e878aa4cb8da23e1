"""Slab sharding of uniform grids across ranks (one process per GPU) and the cross-rank OutputMinMax reduction.

Every point of a grid is independent (Cellular's 3^N neighbourhood is computed, not read), so a grid shards along
its outermost axis with no halo and no data-path collective. A rank evaluates its slab with GLOBAL integer indices
(fnb200GenUniformGrid*: slabBegin/slabEnd), because fma(float(i), step, off) != fma(float(i-k), step, off + k*step)
in general (reference position rule: Generator.inl:160-162). The only exchange is the min/max of 2 floats; an
all-gather of the volume happens only if the caller asks for one contiguous buffer.
"""


def slab_range(total, rank, world):
    """Contiguous, balanced split of `total` outermost-axis indices: the first (total % world) ranks get one extra."""
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_reduce_minmax(lo, hi, device=None):
    """min/max over ranks with ONE all-reduce: MIN over (lo, -hi). No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(lo), float(hi)
    t = torch.tensor([lo, -hi], dtype=torch.float32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    lo, neg_hi = t.tolist()
    return lo, -neg_hi


def all_gather_slabs(local, counts, device=None):
    """Optional assembly of the whole volume (only when a caller wants one contiguous buffer): ragged all-gather."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    # slabs differ by at most one outermost index: pad to the largest, gather, trim (all_gather needs equal sizes)
    largest = int(max(counts))
    padded = local if local.numel() == largest else torch.cat([local, local.new_zeros(largest - local.numel())])
    parts = [torch.empty(largest, dtype=local.dtype, device=local.device) for _ in counts]
    dist.all_gather(parts, padded)
    return torch.cat([part[:int(c)] for part, c in zip(parts, counts)])
