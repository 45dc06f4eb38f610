"""Multi-GPU host logic: games shard by contiguous global game id, one process per GPU, NO data-path
collective.  Dice/action streams are keyed by the GLOBAL game id (stream_base + local index),
so the union of trajectories is the same for 1, 2, 4 or 8 GPUs.  Used by bench.py and scripts/train_ppo.py to
place every rank's env (`stream_base` = the shard's first global id)."""
from __future__ import annotations


def shard_range(total_games: int, rank: int, world: int):
    """Contiguous shard [base, base+count) of `total_games` owned by `rank` (remainder to the low ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    q, r = divmod(int(total_games), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count
