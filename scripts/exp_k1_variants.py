"""K1 alone (tier 0 + overflow tiers) and the whole step on a stationary random-play mix: used to compare library variants."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=65536, device=dev, seed=0x5EED, check_every=0, rows_per_game=64)
env.reset()
acts = torch.empty(65536, dtype=torch.int32, device=dev)
for t in range(128):
    env.step_random_device(7, t); env.update_legal_plays(obs=True, features=True)
ev = lambda: torch.cuda.Event(enable_timing=True)
t = 128
k1 = []
for i in range(40):
    env.random_actions(7, t, out=acts); env._apply_actions(acts)
    a, b = ev(), ev(); a.record(); env._refresh_legal_moves(); b.record(); k1.append((a, b)); t += 1
torch.cuda.synchronize()
k1us = sorted(a.elapsed_time(b) * 1e3 for a, b in k1[5:])
a, b = ev(), ev(); a.record()
for i in range(400):
    env.step_random_device(7, t); env.update_legal_plays(obs=True, features=True); t += 1
b.record(); torch.cuda.synchronize()
print(f"{os.environ.get('TAG', '')}: K1 alone median {k1us[len(k1us) // 2]:.1f} us (min {k1us[0]:.1f}), step {a.elapsed_time(b) / 400 * 1e3:.1f} us")
# the same step with K3's afterstate features written by K1's own output stage (FEATS instantiation), observations after it
a, b = ev(), ev(); a.record()
for i in range(400):
    env.step_random_device(7, t); env._refresh_legal_moves(with_features=True); env.encode_resident(obs=True, afterstates=False); t += 1
b.record(); torch.cuda.synchronize()
print(f"{os.environ.get('TAG', '')}: step with features fused into K1 {a.elapsed_time(b) / 400 * 1e3:.1f} us")
k1 = []
for i in range(30):
    env.random_actions(7, t, out=acts); env._apply_actions(acts)
    a, b = ev(), ev(); a.record(); env._refresh_legal_moves(with_features=True); b.record(); k1.append((a, b)); t += 1
torch.cuda.synchronize()
k1us = sorted(a.elapsed_time(b) * 1e3 for a, b in k1[5:])
print(f"{os.environ.get('TAG', '')}: K1 with fused features alone median {k1us[len(k1us) // 2]:.1f} us")
