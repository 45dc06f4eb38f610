"""K1 alone and the whole step against the team sizes of the overflow tiers (bg_set_team_threads)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from bg_b200._lib import lib
dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=65536, device=dev, seed=0x5EED, check_every=0, rows_per_game=64)
env.reset()
t = 0
for _ in range(160):
    env.step_random_device(7, t); env.update_legal_plays(obs=True, features=True); t += 1
ev = lambda: torch.cuda.Event(enable_timing=True)
acts = torch.empty(65536, dtype=torch.int32, device=dev)
def k1_us():
    global t
    ts = []
    for i in range(30):
        env.random_actions(7, t, out=acts); env._apply_actions(acts)
        a, b = ev(), ev(); a.record(); env._refresh_legal_moves(); b.record(); ts.append((a, b)); t += 1
    torch.cuda.synchronize()
    us = sorted(a.elapsed_time(b) * 1e3 for a, b in ts[5:]); return us[len(us) // 2]
def step_us(n=300):
    global t
    a, b = ev(), ev(); a.record()
    for _ in range(n):
        env.step_random_device(7, t); env.update_legal_plays(obs=True, features=True); t += 1
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for mid, big in [(0, 0), (128, 512), (256, 512), (128, 1024), (128, 256), (64, 512), (256, 1024), (0, 0)]:
    lib().bg_set_team_threads(mid, big)
    step_us(30)
    print(f"team threads mid {mid} big {big}: K1 alone {k1_us():.1f} us, step {step_us():.1f} us", flush=True)
env.check_status()
