"""Write-only bandwidth of this GPU by three means: torch.zero_ (fill kernel), cudaMemsetAsync (driver), and a read+write copy for scale."""
import ctypes, torch
rt = ctypes.CDLL("libcudart.so")
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")          # 1 GiB
y = torch.empty_like(x)
def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
n = x.numel()
t = timed(lambda: x.zero_()); print(f"torch zero_        : {n / t / 1e12:.2f} TB/s")
st = torch.cuda.current_stream().cuda_stream
t = timed(lambda: rt.cudaMemsetAsync(ctypes.c_void_p(x.data_ptr()), 0, ctypes.c_size_t(n), ctypes.c_void_p(st))); print(f"cudaMemsetAsync    : {n / t / 1e12:.2f} TB/s")
t = timed(lambda: y.copy_(x)); print(f"torch copy_ (r+w)  : {2 * n / t / 1e12:.2f} TB/s")
x16 = x.view(torch.int32)
t = timed(lambda: x16.fill_(7)); print(f"torch fill_ int32  : {n / t / 1e12:.2f} TB/s")
