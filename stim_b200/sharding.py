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


def check_shard_invariance(circuit, seed: int, shots_per_rank: int, device: int) -> bool:
    """Every rank samples its shard of one global shot range ([rank * shots_per_rank, +shots_per_rank), block size pinned),
    the device-resident per-bit and adjacent-pair flip counts are summed with the group's all_reduce (NCCL on GPUs), and
    rank 0 compares the sums with one sampler running the whole range: they must be identical (disjoint Philox counter
    ranges, no other inter-GPU traffic). Returns the verdict on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    sampler = circuit.compile_detector_sampler(seed=seed, device=device)
    K = 2 * int(sampler.stats.lanes_per_item)
    shots_per_rank = max(shots_per_rank // (K * COLUMN_SHOTS), 1) * K * COLUMN_SHOTS
    sampler.set_block_columns(K)
    sampler.shot_offset = rank * shots_per_rank
    n = int(sampler.stats.num_detectors + sampler.stats.num_observables)
    single = torch.zeros(n, dtype=torch.int64, device=f"cuda:{device}")
    pair = torch.zeros(max(n - 1, 1), dtype=torch.int64, device=f"cuda:{device}")
    sampler.bit_counts(shots_per_rank, single_dev_ptr=single.data_ptr(), pair_dev_ptr=pair.data_ptr())
    torch.cuda.synchronize()
    total_single = allreduce_counts(single)
    total_pair = allreduce_counts(pair)
    ok = torch.ones(1, dtype=torch.int64, device=f"cuda:{device}")
    if rank == 0:
        whole = circuit.compile_detector_sampler(seed=seed, device=device)
        whole.set_block_columns(K)
        want_single, want_pair = whole.bit_counts(world * shots_per_rank)
        good = np.array_equal(total_single, want_single) and np.array_equal(total_pair[: n - 1], want_pair) and want_single.sum() > 0
        ok[0] = 1 if good else 0
    dist.broadcast(ok, src=0)
    return bool(ok.item())
