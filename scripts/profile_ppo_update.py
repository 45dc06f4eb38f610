"""Where the PPO update's time goes: torch profiler over one update (1 M samples x 4 epochs), for the tcgen05 path (our GEMM
kernels) and the cuBLAS path; then wall-clock per update of both."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from bg_b200.ppo import PPOConfig, PPOTrainer
dev = torch.device("cuda:0")
games = int(os.environ.get("GAMES", 16384))
for impl in ("tcgen05", "cublas"):
    env = bg_b200.B200BackgammonVecEnv(num_envs=games, device=dev, seed=1, check_every=0); env.reset()
    net = bg_b200.PolicyValueNet.random_init(dev, seed=0)
    tr = PPOTrainer(env, net, PPOConfig(t_horizon=64, update_impl=impl), seed=0)
    ret = tr.collect(); tr.update(ret); ret = tr.collect(); torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.update(ret); torch.cuda.synchronize()
    print(f"==== update_impl={impl}, {games * 64} samples x 4 epochs")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        tr.update(ret)
    torch.cuda.synchronize()
    print(f"update_impl={impl}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per update ({games * 64} samples x 4 epochs)")
    del tr, env, net
    torch.cuda.empty_cache()
