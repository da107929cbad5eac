"""HBM bandwidth by access mix (CUDA events): write-only (memset), read-only (reduction), copy (read+write)."""
import torch
n = 1 << 30   # 8 GiB of fp64
x = torch.empty(n, dtype=torch.float64, device="cuda"); y = torch.empty(n // 2, dtype=torch.float64, device="cuda")
def t(fn, bytes_, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return reps * bytes_ / (e0.elapsed_time(e1) * 1e-3) / 1e9
print("write-only (zero_)  GB/s", round(t(lambda: x.zero_(), n * 8)))
print("read-only  (sum)    GB/s", round(t(lambda: x.sum(), n * 8)))
print("copy (r+w)          GB/s", round(t(lambda: y.copy_(x[: n // 2]), n * 8)))
