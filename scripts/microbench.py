"""Per-kernel timings on one B200 (CUDA events, warm, mean of reps): K1 / K3 / K4 on de-phased random-play positions."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200

def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

dev = torch.device("cuda:0")
N = int(os.environ.get("GAMES", 65536))
env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=0x5EED, check_every=0)
env.reset()
acts = torch.empty(N, dtype=torch.int32, device=dev)
for t in range(128):
    env.random_actions(7, t, out=acts); env.step_device(acts)
env.check_status()
rows = env.total_rows()
print(f"games {N}, afterstate rows {rows} ({rows/N:.2f}/game)")
print(f"K1 movegen_slab (3 tiers)      : {timed(env._refresh_legal_moves):8.1f} us")
print(f"K1 movegen_slab + fused bf16   : {timed(lambda: env._refresh_legal_moves(with_features=True)):8.1f} us")
env.encode_resident(True, True)
print(f"K1+K3 update_legal_plays serial : {timed(lambda: env.update_legal_plays(True, True, overlap=False)):8.1f} us")
print(f"K1+K3 update_legal_plays overlap: {timed(lambda: env.update_legal_plays(True, True, overlap=True)):8.1f} us")
from bg_b200.engine import encode
print(f"K3 encode_bf16 {rows} rows     : {timed(lambda: env.encode_resident(False, True)):8.1f} us   -> {rows*469/ (timed(lambda: env.encode_resident(False, True))*1e-6)/1e12:.2f} TB/s")
print(f"K3 encode_f32 obs {N} rows      : {timed(lambda: env.encode_resident(True, False)):8.1f} us")
net = bg_b200.ValueNet.random_init(dev)
vb = torch.empty(rows, dtype=torch.float32, device=dev)
t = timed(lambda: net.values(env.after52[:rows], env.row_players[:rows], out=vb))
print(f"K4 mlp_value {rows} rows        : {t:8.1f} us   -> {rows/t*1e6/1e9:.2f} G pos/s, {rows*53504/t*1e6/1e12:.0f} TFLOP/s")
pnet = bg_b200.PolicyValueNet.random_init(dev)
pout = (torch.empty(N, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.float32, device=dev), torch.empty(N, dtype=torch.float32, device=dev))
t = timed(lambda: pnet.act(env.boards52, env.players, env.legal_counts, seed=1, step=2, out=pout))
print(f"N1 policy_sample {N} rows       : {t:8.1f} us   -> {N/t*1e6/1e9:.2f} G pos/s, {N*(53504+2*128*512)/t*1e6/1e12:.0f} TFLOP/s")
t = timed(lambda: pnet.act(env.boards52, env.players, env.legal_counts, seed=1, step=2, greedy=True, out=pout))
print(f"N1 policy greedy {N} rows       : {t:8.1f} us")
# the same kernel on a batch large enough to amortise the 181 KB weight load per CTA (16 x the positions)
R = 16
bb, pp, cc = env.boards52.repeat(R, 1).contiguous(), env.players.repeat(R).contiguous(), env.legal_counts.repeat(R).contiguous()
pout2 = (torch.empty(N * R, dtype=torch.int32, device=dev), torch.empty(N * R, dtype=torch.float32, device=dev), torch.empty(N * R, dtype=torch.float32, device=dev))
t = timed(lambda: pnet.act(bb, pp, cc, seed=1, step=2, out=pout2), 5)
print(f"N1 policy_sample {N*R} rows     : {t:8.1f} us   -> {N*R/t*1e6/1e9:.2f} G pos/s (useful work follows the mask: one 128-slot chunk of the policy GEMM for 94 % of the rows)")
del bb, pp, cc, pout2
env.random_actions(7, 999, out=acts)
print(f"K2 step                        : {timed(lambda: env._apply_actions(acts), 5):8.1f} us (mutates state)")
env._refresh_legal_moves(); env.check_status()
# calibration: write-only and copy bandwidth of torch on the same buffers
f = env.after_feats[:rows]
t = timed(lambda: f.zero_())
print(f"torch zero_ {f.numel()*2/1e6:.0f} MB           : {t:8.1f} us   -> {f.numel()*2/t*1e6/1e12:.2f} TB/s (write only)")
g = torch.empty_like(f)
t = timed(lambda: g.copy_(f))
print(f"torch copy_ {f.numel()*2/1e6:.0f} MB           : {t:8.1f} us   -> {2*f.numel()*2/t*1e6/1e12:.2f} TB/s (read+write)")
big = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
t = timed(lambda: big.zero_(), 5)
print(f"torch zero_ 2 GiB              : {t:8.1f} us   -> {big.numel()*2/t*1e6/1e12:.2f} TB/s (write only)")
