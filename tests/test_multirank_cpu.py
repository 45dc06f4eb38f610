"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding by global game id, max/sum
report reduction, and GPU-count independence of the trajectories (Philox streams keyed by global game id;
the per-shard games are replayed with the oracle, which uses the engine's dice/action conventions)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _traj_digest(stream_ids, steps, seed, aseed):
    from oracle import bg_oracle as O
    out = []
    for g in stream_ids:
        e = O.Env()
        e.set_philox(seed, g)
        e.reset()
        acc = 0
        for t in range(steps):
            s = e.state()
            acc = (acc * 1000003 + int(np.frombuffer(s["board"].tobytes(), np.uint8).sum()) * 31 + int(s["roll"][0]) * 7 + int(s["roll"][1])) % (1 << 61)
            n = s["n_legal"]
            _, done, _ = e.step(O.philox_action(aseed, g, t, n) if n else 0)
            if done:
                e.reset()
        out.append(acc)
    return out


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bg_b200
    base, count = bg_b200.shard_range(total, rank, world)
    digest = _traj_digest(range(base, base + count), 40, 0x5EED, 0xAC7)
    ms, units = bg_b200.reduce_report(10.0 + rank, count * 40, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (base, count, digest))
    if rank == 0:
        q.put((ms, units, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    import bg_b200
    total, world = 13, 2
    assert [bg_b200.shard_range(total, r, world) for r in range(world)] == [(0, 7), (7, 6)]
    assert bg_b200.shard_range(65536 * 8, 3, 8) == (3 * 65536, 65536)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    [p.start() for p in procs]
    ms, units, gathered = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert ms == 11.0 and units == total * 40                     # max over ranks, sum over ranks
    union = [d for _, _, dig in sorted(gathered) for d in dig]
    assert union == _traj_digest(range(total), 40, 0x5EED, 0xAC7)  # same trajectories as one rank owning all games
