#!/usr/bin/env python
"""PPO self-play training (BASELINE config 5): every game resident on the GPUs, one process per GPU.

    python scripts/train_ppo.py --games 16384 --horizon 64 --updates 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_ppo.py ...

Rollout (env step, legal moves, policy forward + sampling, returns) = hand-written CUDA; update = ManualUpdate (library GEMMs + bg_ppo_loss_grad);
the only collective is the flat-bucket gradient all-reduce (NCCL).  Prints one JSON line per update from rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_env(bg_b200, dev, games_per_gpu, rank, world, seed):
    """This rank's shard of the world * games_per_gpu games (contiguous global game ids; Philox streams keyed by them)."""
    base, count = bg_b200.shard_range(world * games_per_gpu, rank, world)
    return bg_b200.B200BackgammonVecEnv(num_envs=count, device=dev, seed=0x5EED + seed, stream_base=base, check_every=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=16384, help="games per GPU")
    ap.add_argument("--horizon", type=int, default=64)
    ap.add_argument("--updates", type=int, default=10)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatches", type=int, default=1)
    ap.add_argument("--lam", type=float, default=1.0)
    ap.add_argument("--bootstrap", action="store_true")
    ap.add_argument("--eval-games", type=int, default=4096)
    ap.add_argument("--eval-every", type=int, default=5)
    ap.add_argument("--save", default="")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--update-impl", default="tcgen05", choices=["tcgen05", "cublas"], help="the update's linear algebra: hand-written kernels / library GEMMs")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bg_b200
    from bg_b200.ppo import PPOConfig, PPOTrainer, evaluate_vs_random
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = make_env(bg_b200, dev, args.games, rank, world, args.seed)
    env.reset()
    net = bg_b200.PolicyValueNet.random_init(dev, seed=args.seed)          # same weights on every rank
    cfg = PPOConfig(t_horizon=args.horizon, num_epochs=args.epochs, num_minibatches=args.minibatches, lam=args.lam, bootstrap=args.bootstrap,
                    update_impl=args.update_impl)
    tr = PPOTrainer(env, net, cfg, dist if world > 1 else None, seed=args.seed)
    for u in range(args.updates):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ret = tr.collect()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        tr.count_episodes()                                                # global (all-reduced) episode count -> entropy anneal
        stats = tr.update(ret)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        env.check_status()
        line = dict(update=u, **stats, episodes=tr.episodes, rollout_s=t1 - t0, update_s=t2 - t1,
                    env_steps_per_s=world * args.games * args.horizon / (t2 - t0), rollout_env_steps_per_s=world * args.games * args.horizon / (t1 - t0))
        if rank == 0 and args.eval_games and (u % args.eval_every == args.eval_every - 1 or u == args.updates - 1):
            line.update(evaluate_vs_random(net, args.eval_games, dev))
        if rank == 0:
            print(json.dumps(line), flush=True)
    if rank == 0 and args.save:
        torch.save(net.state_dict(), args.save)                            # the reference's checkpoint format
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
