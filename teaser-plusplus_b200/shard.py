"""Host-side multi-GPU plumbing: independent registration problems are sharded by index over the ranks of one
node (one process per GPU).  There is NO collective on the data path: torch.distributed is used only
for the start/stop barrier, the max-over-ranks timing and the gather of the ~200-byte result records."""
from __future__ import annotations

import numpy as np


def shard_indices(n_problems: int, rank: int, world: int) -> np.ndarray:
    """Problem b belongs to rank b mod world ('sharded by b mod n_gpus')."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_problems, world, dtype=np.int64)


def max_over_ranks(value_ms: float, dist=None, device="cpu") -> float:
    """Timing rule of the benchmark contract: the job time is the MAX over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value_ms)
    import torch
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_solutions(local_records: np.ndarray, local_idx: np.ndarray, n_problems: int, dist=None):
    """Gather per-rank structured solution records to rank 0, restoring the global problem order.
    Returns the full array on rank 0, None elsewhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.zeros(n_problems, dtype=local_records.dtype)
        out[local_idx] = local_records
        return out
    payload = (local_idx.tolist(), local_records.tobytes())
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if dist.get_rank() != 0:
        return None
    out = np.zeros(n_problems, dtype=local_records.dtype)
    for idx, raw in gathered:
        out[np.asarray(idx, dtype=np.int64)] = np.frombuffer(raw, dtype=local_records.dtype)
    return out
