"""Where the host time of one e2e step goes (perf_counter around the pieces, queue empty at the start of each step),
then the plain e2e loop as bench.py runs it, with the encoders beside K1's overflow tiers (overlap) or after K1 (serial)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=0x5EED, check_every=0); env.reset()
acts = torch.empty(N, dtype=torch.int32, device=dev)
for t in range(128):
    env.random_actions(1, t, out=acts); env._apply_actions(acts); env.update_legal_plays(obs=True, features=True)
host = bg_b200.HostStepBuffers(env)
h_acts = torch.empty(N, dtype=torch.int32).pin_memory()
host.legal_counts.copy_(env.legal_counts); torch.cuda.synchronize()
rng = np.random.default_rng(1)
acts_np, counts_np, tmp = h_acts.numpy(), host.legal_counts.numpy(), np.empty(N, dtype=np.int32)
K = 200
for overlap in (True, False):
    T = {}
    def tick(name, t0):
        t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1
    u = rng.integers(0, 65536, size=(K, N), dtype=np.int32)
    for k in range(K):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        np.multiply(u[k], counts_np, out=tmp); np.right_shift(tmp, 16, out=acts_np)
        t0 = tick("numpy policy", t0)
        env.step(h_acts, with_features=True, host=host, overlap=overlap)
        t0 = tick("env.step call", t0)
        host.wait()
        t0 = tick("host.wait", t0)
        torch.cuda.synchronize()
        t0 = tick("rest of GPU work (K3 tail)", t0)
    print("overlap", overlap, " ".join(f"[{k} {v / K * 1e6:.1f} us]" for k, v in T.items()), "total %.1f us" % (sum(T.values()) / K * 1e6))
    for seg in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(100):
            host.wait()
            np.multiply(u[k], counts_np, out=tmp); np.right_shift(tmp, 16, out=acts_np)
            env.step(h_acts, with_features=True, host=host, overlap=overlap)
        host.wait(); torch.cuda.synchronize()
        print("   loop ms/step", (time.perf_counter() - t0) * 10)
        u = rng.integers(0, 65536, size=(K, N), dtype=np.int32)
env.check_status()
