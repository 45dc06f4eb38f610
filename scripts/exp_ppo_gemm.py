"""Where the time of the PPO GEMM kernels goes: each op on 1 M rows (8,192 tiles) with parts switched off (bg_ppo_gemm_debug)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
from bg_b200._lib import lib, check
dev = torch.device("cuda:0")
L = lib(); st = torch.cuda.current_stream().cuda_stream
T = 8192; B = T * 128
tc = bg_b200.TensorCoreUpdate(dev)
flat = torch.randn(90101, device=dev) * 0.1
check(L.bg_ppo_pack_weights(flat.data_ptr(), tc.w1p.data_ptr(), tc.wap_a.data_ptr(), tc.wap_b.data_ptr(), tc.bias_a.data_ptr(), tc.bias_b.data_ptr(), st))
r = lambda n: torch.randn(n, device=dev).to(torch.bfloat16)
x, h, la, dpre = r(B * 208), torch.relu(r(B * 128)), r(B * 144), torch.empty(B * 128, dtype=torch.bfloat16, device=dev)
g = torch.zeros(90101, device=dev)
scr = torch.empty(199 * 128, device=dev)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
ops = {
    "HIDDEN   (672 B/row)": (lambda: check(L.bg_ppo_gemm_nt(0, x.data_ptr(), 0, T, tc.w1p.data_ptr(), None, None, h.data_ptr(), st)), 672),
    "LOGITS_A (544 B/row)": (lambda: check(L.bg_ppo_gemm_nt(1, h.data_ptr(), 0, T, tc.wap_a.data_ptr(), tc.bias_a.data_ptr(), None, la.data_ptr(), st)), 544),
    "DPRE_A   (800 B/row)": (lambda: check(L.bg_ppo_gemm_nt(3, la.data_ptr(), 0, T, tc.wap_a.data_ptr(), None, h.data_ptr(), dpre.data_ptr(), st)), 800),
    "GRAD_WA_A(544 B/row)": (lambda: check(L.bg_ppo_gemm_tn(5, h.data_ptr(), la.data_ptr(), 0, T, g.data_ptr(), None, st)), 544),
    "GRAD_W1  (672 B/row)": (lambda: check(L.bg_ppo_gemm_tn(7, dpre.data_ptr(), x.data_ptr(), 0, T, g.data_ptr(), scr.data_ptr(), st)), 672),
}
for flags, what in ((0, "normal"), (8, "A operand from shared memory (SS MMAs) where the default is tensor memory"), (1, "no MMAs"), (2, "no stores"), (3, "no MMAs, no stores"), (4, "no loads"), (7, "nothing but the loop")):
    L.bg_ppo_gemm_debug(flags)
    print(f"--- {what}")
    for name, (fn, bpr) in ops.items():
        t = timed(fn)
        print(f"   {name}: {t:7.1f} us  {B * bpr / t / 1e6:6.2f} TB/s")
L.bg_ppo_gemm_debug(0)
hh = h.view(B, 128); dd = dpre.view(B, 128)
t = timed(lambda: hh.copy_(dd))
print(f"torch copy 2 x 256 B/row: {t:7.1f} us  {B * 512 / t / 1e6:6.2f} TB/s")
w = torch.randn((128, 208), device=dev).to(torch.bfloat16)
xx = x.view(B, 208)
t = timed(lambda: torch.mm(xx, w.t(), out=hh))
print(f"cuBLAS x @ W1p^T        : {t:7.1f} us  {B * 672 / t / 1e6:6.2f} TB/s")
