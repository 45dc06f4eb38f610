"""Policy kernel alone: bg_policy_sample on 65,536 random-play positions (partition + fused kernel), CUDA events."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
dev = torch.device("cuda:0")
N = int(os.environ.get("GAMES", 65536))
env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=0x5EED, check_every=0)
env.reset()
for t in range(128):
    env.step_random_device(7, t); env.update_legal_plays(obs=False, features=False)
net = bg_b200.PolicyValueNet.random_init(dev, seed=0)
ev = lambda: torch.cuda.Event(enable_timing=True)
def med(fn, n=30):
    ts = []
    for i in range(n):
        a, b = ev(), ev(); a.record(); fn(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    us = sorted(a.elapsed_time(b) * 1e3 for a, b in ts[5:]); return us[len(us) // 2], us[0]
outb = (torch.empty(N, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.float32, device=dev), torch.empty(N, dtype=torch.float32, device=dev))
counts = env.legal_counts.to(torch.int32).contiguous()
if os.environ.get('ALL_A1'): counts = counts.clamp(min=1, max=32).contiguous()
if os.environ.get('NO_B'): counts = counts.clamp(min=1, max=128).contiguous()
f = lambda: net.act(env.boards52, env.players, counts, seed=3, step=5, out=outb)
print(os.environ.get("TAG", ""), N, "policy sample: median %.1f us, min %.1f us" % med(f))

# GPU time without the host's launch path: the same call captured in a CUDA graph and replayed
f(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    f(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        f()
torch.cuda.synchronize()
a, b = ev(), ev()
g.replay(); torch.cuda.synchronize()
a.record()
for _ in range(50): g.replay()
b.record(); torch.cuda.synchronize()
print(os.environ.get("TAG", ""), N, "graph replay: %.1f us per call" % (a.elapsed_time(b) / 50 * 1e3))
