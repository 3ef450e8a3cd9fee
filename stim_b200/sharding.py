"""Multi-GPU sharding of the sampling path: shots are independent, so each rank samples its own contiguous range
of the global shot index space (disjoint Philox counters, no inter-GPU traffic). The only collective of the path
is the optional per-detector flip-count sum: allreduce(SUM) over uint64[D+L].

The reference has no multi-process layer for this path (sinter's worker pool sums counts on the host,
/root/reference/glue/sample/src/sinter/_collection/_collection_manager.py:129-185)."""
from typing import Tuple

import numpy as np

COLUMN_SHOTS = 128  # shots per Philox column (program.h GSTIM_COL_SHOTS)


def shard_range(total_shots: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, count) of the global shot range owned by `rank`: contiguous, multiples of 128 (so ptb64 groups and
    Philox columns never straddle ranks), remainder to the last rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    cols = total_shots // COLUMN_SHOTS
    per = cols // world
    first = rank * per * COLUMN_SHOTS
    if rank == world - 1:
        return first, total_shots - first
    return first, per * COLUMN_SHOTS


def rank_shot_offset(rank: int, shots_per_rank_per_call: int, calls: int) -> int:
    """Shot offset that keeps `calls` successive calls of every rank disjoint from all other ranks' shots."""
    span = (shots_per_rank_per_call + 32 * COLUMN_SHOTS) // COLUMN_SHOTS * COLUMN_SHOTS  # blocks round shots up
    return rank * span * max(calls, 1)


def allreduce_counts(counts: np.ndarray, device=None):
    """Sum per-detector flip counts over all ranks of the default torch.distributed group (NCCL on GPUs, gloo on CPU).

    counts: uint64[D+L] host array (or a device int64 tensor for NCCL). Returns the reduced numpy array."""
    import torch
    import torch.distributed as dist

    if isinstance(counts, np.ndarray):
        t = torch.from_numpy(counts.astype(np.int64))
        if device is not None:
            t = t.to(device)
    else:
        t = counts
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)
