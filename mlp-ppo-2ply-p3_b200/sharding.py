"""Multi-GPU host logic: games shard by contiguous global game id, one process per GPU, NO data-path
collective.  Dice/action streams are keyed by the GLOBAL game id (stream_base + local index),
so the union of trajectories is the same for 1, 2, 4 or 8 GPUs.  The only reductions are for reporting."""
from __future__ import annotations


def shard_range(total_games: int, rank: int, world: int):
    """Contiguous shard [base, base+count) of `total_games` owned by `rank` (remainder to the low ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    q, r = divmod(int(total_games), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def reduce_report(local_ms: float, local_units: float, dist=None, device=None):
    """(max over ranks of the elapsed time, sum over ranks of the processed units) -- what bench.py reports."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms), float(local_units)
    import torch
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
