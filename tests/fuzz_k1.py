#!/usr/bin/env python
"""Differential fuzz of K1 (GPU, all tiers, through the C ABI) against the CPU oracle at scale: positions from the
engine's own random self-play x all 36 ordered rolls (plus the positions' own dice), afterstate lists compared IN ORDER.
The oracle (oracle/bg_oracle.c) is pinned to the reference by tests/test_oracle_golden.py; this script is a checker,
not a product path (it lives under tests/ because it uses the oracle).   python tests/fuzz_k1.py [--positions 20000] [--threads 16]
tests/test_gpu_fuzz_large.py runs it at 30,000 positions (1.08 M pairs); profiles/r1_l_fuzz_*.json are 30 k / 200 k runs.
Prints one JSON line; exit code 1 on any difference."""
import argparse, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from oracle import bg_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--positions", type=int, default=20000)
ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
ap.add_argument("--seed", type=int, default=2026)
args = ap.parse_args()
dev = torch.device("cuda:0")
N = 4096
env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=args.seed, check_every=0)
env.reset()
boards, players = [], []
t = 0
while sum(b.shape[0] for b in boards) < args.positions:
    for _ in range(7):
        env.step_device(env.random_actions(11, t)); t += 1
    boards.append(env.boards52.cpu().numpy().copy()); players.append(env.players.cpu().numpy().copy())
env.check_status()
b52 = np.concatenate(boards)[: args.positions]; pl = np.concatenate(players)[: args.positions]
rolls = np.array([(a, b) for a in range(1, 7) for b in range(1, 7)], np.int8)        # 36 ordered rolls
P = b52.shape[0]
B52 = np.repeat(b52, 36, axis=0); PL = np.repeat(pl, 36); DC = np.tile(rolls, (P, 1))
t0 = time.perf_counter()
counts, offsets, after = bg_b200.legal_moves(torch.as_tensor(B52).to(dev), torch.as_tensor(PL).to(dev), torch.as_tensor(DC).to(dev))
torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
counts, offsets, after = counts.cpu().numpy(), offsets.cpu().numpy(), after.cpu().numpy()
B4 = O.unpack52(B52)
O.lib()
chunks = np.array_split(np.arange(B52.shape[0]), args.threads * 8)
def work(idx):
    c, o, a = O.legal_moves_batch(B4[idx], PL[idx], DC[idx])
    bad = 0
    if not np.array_equal(c, counts[idx]):
        return int((c != counts[idx]).sum()), int(c.sum())
    lo, hi = offsets[idx[0]], offsets[idx[-1] + 1]
    if not np.array_equal(O.pack52(a), after[lo:hi]):
        bad = 1
    return bad, int(c.sum())
t0 = time.perf_counter()
with ThreadPoolExecutor(args.threads) as ex:
    res = list(ex.map(work, chunks))
t_cpu = time.perf_counter() - t0
bad = sum(r[0] for r in res); rows = sum(r[1] for r in res)
print(json.dumps({"positions": int(P), "position_roll_pairs": int(B52.shape[0]), "afterstates": rows, "max_count": int(counts.max()),
                  "pairs_over_128": int((counts > 128).sum()), "pairs_over_512": int((counts > 512).sum()),
                  "differences": bad, "gpu_seconds_incl_alloc": round(t_gpu, 3), "oracle_seconds": round(t_cpu, 1), "oracle_threads": args.threads}))
sys.exit(1 if bad else 0)
