"""world_size-2 gloo test (CPU) of the multi-GPU host logic that bench.py and scripts/train_ppo.py actually run:
  * every rank's env is placed by bench.make_env / train_ppo.make_env -> shard_range: contiguous global game ids,
    stream_base = the shard's first id (the Philox dice / action streams are keyed by stream_base + local index, so
    the union over ranks is the single-process id range whatever the world size);
  * PPOTrainer.count_episodes all-reduces the finished-game count, so every rank anneals the entropy coefficient on the
    GLOBAL episode count (ppo_agent.py:193, train.py:73) and holds the same value.
The env constructor is replaced by a recorder (there is no CPU engine: the arguments are what is under test); the
trajectories' independence of the shard count itself is a GPU test (test_gpu_fullsize.py)."""
import importlib.util
import os
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _worker(rank, world, port, games_per_gpu, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bg_b200
    from bg_b200.ppo import PPOTrainer
    calls = []

    class Recorder:
        def __init__(self, **kw):
            calls.append(kw)
            self.num_envs = kw["num_envs"]
    bg_b200.B200BackgammonVecEnv = Recorder
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_under_test")
    train = _load(os.path.join(ROOT, "scripts", "train_ppo.py"), "train_ppo_under_test")
    bench.make_env(bg_b200, "cpu", games_per_gpu, rank, world, 64)
    train.make_env(bg_b200, "cpu", games_per_gpu, rank, world, 3)
    # global episode count: rank r finished 10 + r games with reward sum 1.5 * (r + 1)
    dones = torch.zeros((4, 16), dtype=torch.uint8); dones.view(-1)[: 10 + rank] = 1
    rewards = torch.zeros((4, 16)); rewards[0, 0] = 1.5 * (rank + 1)
    fake = types.SimpleNamespace(buf=types.SimpleNamespace(dones=dones, rewards=rewards), dist=dist, episodes=100,
                                 learner=types.SimpleNamespace(total_episodes=0))
    n_done, reward_sum, w = PPOTrainer.count_episodes(fake)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, calls, n_done, reward_sum, w, fake.episodes, fake.learner.total_episodes))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_global_episode_count():
    import bg_b200
    assert [bg_b200.shard_range(13, r, 2) for r in range(2)] == [(0, 7), (7, 6)]
    assert bg_b200.shard_range(65536 * 8, 3, 8) == (3 * 65536, 65536)
    world, per_gpu = 2, 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_gpu, q)) for r in range(world)]
    [p.start() for p in procs]
    gathered = q.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    covered = []
    for rank, calls, n_done, reward_sum, w, episodes, total in sorted(gathered):
        b_call, t_call = calls
        for c, seed in ((b_call, 0x5EED), (t_call, 0x5EED + 3)):
            assert c["num_envs"] == per_gpu and c["stream_base"] == rank * per_gpu and c["seed"] == seed
        covered += list(range(b_call["stream_base"], b_call["stream_base"] + b_call["num_envs"]))
        # every rank sees the GLOBAL totals
        assert (n_done, w) == (10 + 11, 2) and abs(reward_sum - 4.5) < 1e-12
        assert episodes == 121 and total == 121
    assert covered == list(range(world * per_gpu))               # the union of the shards = one rank owning all games
