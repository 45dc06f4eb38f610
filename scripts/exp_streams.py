import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
dev = torch.device("cuda:0")
N = 65536
for nstreams in (1, 2, 4):
    envs = [bg_b200.B200BackgammonVecEnv(num_envs=N // nstreams, device=dev, seed=0x5EED, stream_base=i * (N // nstreams), check_every=0) for i in range(nstreams)]
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    acts = [torch.empty(N // nstreams, dtype=torch.int32, device=dev) for _ in range(nstreams)]
    for e in envs: e.reset()
    torch.cuda.synchronize()
    def step(t, feats=True):
        for e, s, a in zip(envs, streams, acts):
            with torch.cuda.stream(s):
                e.random_actions(7, t, out=a); e._apply_actions(a); e._refresh_legal_moves(); e.encode_resident(True, feats)
    for t in range(140): step(t)
    torch.cuda.synchronize()
    for feats in (True, False):
        t0 = time.perf_counter()
        K = 300
        for t in range(K): step(1000 + t, feats)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"streams={nstreams} feats={feats}: {dt/K*1e3:.3f} ms/step  {N*K/dt/1e6:.1f} M steps/s", flush=True)
    for e in envs: e.check_status()
