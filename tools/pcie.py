import torch, time
n = 1 << 27  # 1 GiB of fp64
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda"); d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n * 8 / dt / 1e9
run(True, True, 1)
print("H2D only  GB/s", run(True, False)); print("D2H only  GB/s", run(False, True)); print("both (each dir) GB/s", run(True, True))
