"""Does a CUDA graph of the device-resident step (random actions -> K2 -> K1 -> K3) beat the stream-launched one?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
N = 65536
dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=0x5EED, check_every=0); env.reset()
acts = torch.empty(N, dtype=torch.int32, device=dev)
tctr = [0]
def one_step(t):
    env.random_actions(0xAC7, t, out=acts); env._apply_actions(acts); env.update_legal_plays(obs=True, features=True)
for t in range(160): one_step(t)
torch.cuda.synchronize()
def timed(fn, reps=300):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3
print("stream launched: %.1f us/step" % timed(lambda: one_step(200)))
# graph of 4 consecutive steps (fixed action seeds t; the games still advance, only the action stream repeats)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    one_step(300); 
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for t in range(4): one_step(400 + t)
torch.cuda.synchronize()
print("graph of 4 steps: %.1f us/step" % (timed(lambda: g.replay(), 100) / 4))
env.check_status()
# the same loop with a CUDA event pair around K1 (what bench.py does in its timed region), and with the row accumulator
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(301)]
for a, b in evs: a.record(); b.record()
i = [0]
def step_ev():
    env.random_actions(0xAC7, 200, out=acts); env._apply_actions(acts)
    env.update_legal_plays(obs=True, features=True, k1_events=evs[i[0] % 301]); i[0] += 1
print("with K1 event pairs: %.1f us/step" % timed(step_ev))
rows_acc = torch.zeros(1, dtype=torch.int64, device=dev)
def step_acc():
    one_step(200); rows_acc.add_(env.alloc_rows)
print("with rows_acc.add_: %.1f us/step" % timed(step_acc))
tt = [1000]
def step_t():
    one_step(tt[0]); tt[0] += 1
print("varying t: %.1f us/step" % timed(step_t))
print("serial encoders: %.1f us/step" % timed(lambda: (env.random_actions(0xAC7, 200, out=acts), env._apply_actions(acts), env.update_legal_plays(obs=True, features=True, overlap=False))))
