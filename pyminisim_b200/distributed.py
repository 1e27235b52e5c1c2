"""Multi-GPU plumbing: worlds are independent, so they shard across ranks with NO collective on the step
path; the only collective is the gather of per-episode statistics (a few numbers) at the end.

One process per GPU (torchrun), ``torch.distributed`` with NCCL on GPUs (gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def shard_worlds(total_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [first, first + count) of the global world range owned by ``rank``."""
    base, extra = divmod(int(total_worlds), int(world_size))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def gather_episode_stats(local: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum-reduce additive episode statistics (detections, collisions, non-finite worlds, speed sums, world
    counts) over all ranks.  Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    keys = sorted(local.keys())
    t = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def local_episode_stats(det: np.ndarray, col: np.ndarray, nonfinite: np.ndarray, vels: np.ndarray) -> Dict[str, float]:
    return {"detections": float(det.sum()), "robot_ped_collisions": float(col.sum()),
            "nonfinite_worlds": float((nonfinite != 0).sum()), "speed_sum": float(np.linalg.norm(vels[..., :2], axis=-1).sum()),
            "pedestrians": float(vels.shape[0] * vels.shape[1]), "worlds": float(vels.shape[0])}
