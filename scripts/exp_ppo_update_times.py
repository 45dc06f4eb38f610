"""Per-update device times of PPOTrainer.update at the bench's size (65,536 games x 64 steps = 4.2 M samples x 4 epochs)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from bg_b200.ppo import PPOConfig, PPOTrainer
dev = torch.device("cuda:0")
games = int(os.environ.get("GAMES", 65536))
env = bg_b200.B200BackgammonVecEnv(num_envs=games, device=dev, seed=1, check_every=0); env.reset()
net = bg_b200.PolicyValueNet.random_init(dev, seed=0)
tr = PPOTrainer(env, net, PPOConfig(t_horizon=64), seed=0)
if os.environ.get("BG_PPO_DBG"):                     # experiment switches of the GEMM kernels (results are garbage, timings are not)
    from bg_b200._lib import lib
    lib().bg_ppo_gemm_debug(int(os.environ["BG_PPO_DBG"]))
ev = lambda: torch.cuda.Event(enable_timing=True)
for u in range(6):
    a, b, c = ev(), ev(), ev()
    a.record(); ret = tr.collect(); b.record(); tr.count_episodes(); tr.update(ret); c.record(); torch.cuda.synchronize()
    p = tr.learner._manual._p
    print(f"update {u}: rollout {a.elapsed_time(b):.2f} ms, update {b.elapsed_time(c):.2f} ms  (n_a {p['n_a']}, n_b {p['n_b']}, cap {tr.learner._manual._cap}, "
          f"alloc {torch.cuda.memory_allocated() / 2**30:.1f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB)", flush=True)
from torch.profiler import profile, ProfilerActivity
ret = tr.collect(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.update(ret); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
