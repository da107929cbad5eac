"""BASELINE.json config 5: BarotropicField equivalent latitude + LWA on 2-D 0.25-degree absolute-vorticity fields,
batched over 1460 time steps (fields synthesised on the device: 2*Omega*sin(phi) + a localised wave + noise)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from hn2016_falwa_b200.barotropic_field import BarotropicBatch

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1460
nlat, nlon, chunk = 721, 1440, 146
ylat = np.linspace(-90, 90, nlat); xlon = np.linspace(0, 360, nlon, endpoint=False)
phi = torch.deg2rad(torch.tensor(ylat, device="cuda"))[None, :, None]
lam = torch.deg2rad(torch.tensor(xlon, device="cuda"))[None, None, :]
g = torch.Generator(device="cuda").manual_seed(1234)
def make(t0, n):
    ph = (torch.arange(t0, t0 + n, device="cuda", dtype=torch.float64) * 0.05)[:, None, None]
    return (2 * 7.29e-5 * torch.sin(phi) + 8e-5 * torch.cos(phi) * torch.exp(-((torch.rad2deg(phi) - 36) / 10) ** 2)
            * torch.cos(3 * lam + ph) + 1e-6 * torch.randn((n, nlat, nlon), device="cuda", dtype=torch.float64, generator=g))
BarotropicBatch(xlon, ylat, make(0, 2)).compute()   # warm-up
torch.cuda.synchronize(); t_all = 0.0
for t0 in range(0, nt, chunk):
    pv = make(t0, min(chunk, nt - t0))
    torch.cuda.synchronize(); a = time.perf_counter()
    b = BarotropicBatch(xlon, ylat, pv).compute()
    torch.cuda.synchronize(); t_all += time.perf_counter() - a
print(f"config 5: {nt} fields of {nlat}x{nlon}: {t_all:.2f} s = {nt / t_all:.0f} fields/s "
      f"(reference NumPy: 23.9 s per field on one core); lwa mean {float(b.lwa.mean()):.4f}")
