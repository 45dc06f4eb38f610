"""Does the ORDER in which K1 meets the positions matter (heavy doubles first / last / natural)?  Same positions, permuted in place."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=65536, device=dev, seed=0x5EED, check_every=0, rows_per_game=64)
env.reset()
for t in range(160):
    env.step_random_device(7, t); env.update_legal_plays(obs=True, features=True)
ev = lambda: torch.cuda.Event(enable_timing=True)
b0, p0, d0 = env.boards52.clone(), env.players.clone(), env.dice.clone()
dbl = (d0[:, 0] == d0[:, 1])
def run(tag, perm):
    env.boards52.copy_(b0[perm]); env.players.copy_(p0[perm]); env.dice.copy_(d0[perm])
    ts = []
    for i in range(25):
        a, b = ev(), ev(); a.record(); env._refresh_legal_moves(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    us = sorted(a.elapsed_time(b) * 1e3 for a, b in ts[5:])
    print(f"{tag:28s}: K1 median {us[len(us) // 2]:.1f} us (min {us[0]:.1f}), rows {int(env.alloc_rows[0])}")
N = 65536
ar = torch.arange(N, device=dev)
run("natural", ar)
run("doubles first", torch.argsort((~dbl).to(torch.int8), stable=True))
run("doubles last", torch.argsort(dbl.to(torch.int8), stable=True))
run("natural again", ar)
c = env.legal_counts_true.clone().long()
if c is not None:
    run("by plays, descending", torch.argsort(-c, stable=True))
    run("doubles first, by plays desc", torch.argsort(-(c + dbl.long() * 100000), stable=True))
