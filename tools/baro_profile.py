"""Per-kernel CUDA-event times of the 2-D (barotropic) path for one chunk of 0.25-degree fields."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from hn2016_falwa_b200 import _lib
from hn2016_falwa_b200.barotropic_field import BarotropicBatch
nlat, nlon, n = 721, 1440, int(sys.argv[1]) if len(sys.argv) > 1 else 146
ylat = np.linspace(-90, 90, nlat); xlon = np.linspace(0, 360, nlon, endpoint=False)
phi = torch.deg2rad(torch.tensor(ylat, device="cuda"))[None, :, None]
lam = torch.deg2rad(torch.tensor(xlon, device="cuda"))[None, None, :]
g = torch.Generator(device="cuda").manual_seed(1)
ph = (torch.arange(n, device="cuda", dtype=torch.float64) * 0.05)[:, None, None]
pv = (2 * 7.29e-5 * torch.sin(phi) + 8e-5 * torch.cos(phi) * torch.exp(-((torch.rad2deg(phi) - 36) / 10) ** 2)
      * torch.cos(3 * lam + ph) + 1e-6 * torch.randn((n, nlat, nlon), device="cuda", dtype=torch.float64, generator=g))
BarotropicBatch(xlon, ylat, pv[:2]).compute()
_lib.profile_enable(True)
BarotropicBatch(xlon, ylat, pv).compute()
for k, (cnt, ms) in sorted(_lib.profile_read().items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} launches {cnt:3d}  {ms:9.3f} ms  = {ms / n * 1e3:8.1f} us per field")
