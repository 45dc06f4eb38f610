"""2-ply stage timings on one B200: K1 replicate-21 alone, K4 on its replies alone, the fused call (serial / overlapped)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from bg_b200._lib import lib, check
from bg_b200.engine import _stream

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

dev = torch.device("cuda:0")
env = bg_b200.B200BackgammonVecEnv(num_envs=8192, device=dev, seed=0x5EED, check_every=0)
env.reset()
for t in range(128):
    env.step_device(env.random_actions(7, t))
net = bg_b200.ValueNet.random_init(dev)
counts, offsets, A, rowp = bg_b200.legal_moves(env.boards52, env.players, env.dice, with_row_players=True)
M = min(32768, A.shape[0]); A, rowp = A[:M].contiguous(), rowp[:M].contiguous()
s = bg_b200.TwoPlySearch(net, max_afterstates_per_chunk=32768)
b = s._workspace(dev)
L = lib()
def k1():
    b["alloc"].zero_()
    check(L.bg_movegen_replies_slab(A.data_ptr(), rowp.data_ptr(), M, 0, b["replies"].data_ptr(), b["cap"], b["rowp"].data_ptr(), None, None,
                                    b["counts"].data_ptr(), b["starts"].data_ptr(), b["alloc"].data_ptr(), b["ws"].status.data_ptr(),
                                    b["ws"].buf.data_ptr(), b["ws"].nbytes, _stream()), "k1")
t1 = timed(k1); rows = int(b["alloc"].item())
print(f"root afterstates {M}, work items {M*21}, replies {rows} ({rows/M:.0f}/afterstate)")
print(f"K1 replies (3 tiers)     : {t1:8.1f} us  -> {M*21/t1:.1f} M movegen/s")
t4 = timed(lambda: net.values(b["replies"], b["rowp"], terminal_aware=True, n_rows_dev=b["alloc"], out=b["leaf_v"]))
print(f"K4 leaves                : {t4:8.1f} us  -> {rows/t4/1e3:.2f} G leaves/s")
out = torch.empty(M, dtype=torch.float32, device=dev)
for ov in (False, True):
    s.overlap = ov
    t = timed(lambda: s._score_chunk(A, rowp, out))
    print(f"score_chunk overlap={ov!s:5}: {t:8.1f} us  -> {M/t:.2f} M root afterstates/s")

# ---- the whole search: bg_twoply (fused, one call) vs the unfused chunked pipeline, 4,096 roots
R = 4096
rb, rp, rd = env.boards52[:R].clone(), env.players[:R].clone(), env.dice[:R].clone()
f = s.search_device(rb, rp, rd); torch.cuda.synchronize()
assert int(f["status"].item()) == 0, int(f["status"].item())
na, nl, oi, ol = (int(x) for x in f["stats"].tolist())
tf = timed(lambda: s.search_device(rb, rp, rd), reps=10)
print(f"bg_twoply fused          : {tf:8.1f} us  -> {na/tf:.2f} M root afterstates/s, {nl/tf/1e3:.2f} G leaves/s "
      f"({na} afterstates, {nl} leaves, overflow items {oi} = {oi/(na*21)*100:.2f} %, overflow leaves {ol/nl*100:.1f} %)")
u = bg_b200.TwoPlySearch(net, max_afterstates_per_chunk=98304)
u.search_unfused(rb, rp, rd); torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    u.search_unfused(rb, rp, rd)
torch.cuda.synchronize()
tu = (time.perf_counter() - t0) / 5 * 1e6
print(f"unfused search (wall)    : {tu:8.1f} us  -> {na/tu:.2f} M root afterstates/s")
