"""Where the PPO update's time goes: torch profiler over one update (1 M samples x 4 epochs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from bg_b200.ppo import PPOConfig, PPOTrainer
dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=16384, device=dev, seed=1, check_every=0); env.reset()
net = bg_b200.PolicyValueNet.random_init(dev, seed=0)
tr = PPOTrainer(env, net, PPOConfig(t_horizon=64), seed=0)
ret = tr.collect(); tr.update(ret); ret = tr.collect(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.update(ret); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
