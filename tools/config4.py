"""BASELINE.json config 4: one year of 6-hourly 1.5-degree snapshots (1460 time steps) through QGDataset on one GPU
(time shards over more GPUs: one process per GPU, see tests/test_dataset_sharding.py)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from hn2016_falwa_b200 import xarrayinterface as xi
from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
from hn2016_falwa_b200.synthetic import synthetic_snapshot

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1460
base = [synthetic_snapshot(nlon=240, nlat=121, seed=100 + k) for k in range(8)]
xlon, ylat, plev = base[0][:3]
u, v, t = (np.empty((nt,) + base[0][3].shape) for _ in range(3))
for k in range(nt):
    s = base[k % 8]
    sh = (k // 8) % 240
    u[k], v[k], t[k] = (np.roll(s[i], sh, axis=-1) for i in (3, 4, 5))
coords = dict(time=np.arange(nt), level=plev, latitude=ylat, longitude=xlon)
dims = ("time", "level", "latitude", "longitude")
mk = lambda a, n: xi.SimpleDataArray(a, dims, coords, name=n)
ds = xi.SimpleDataset(dict(u=mk(u, "u"), v=mk(v, "v"), t=mk(t, "t")))
w = xi.SimpleDataset(dict(u=mk(u[:256], "u"), v=mk(v[:256], "v"), t=mk(t[:256], "t")))
w["u"].coords["time"] = w["v"].coords["time"] = w["t"].coords["time"] = np.arange(256)
xi.QGDataset(w, qgfield=QGFieldNHN22, keep_3d=False, chunk=128).compute_lwa_and_barotropic_fluxes(return_dataset=False)  # warm-up
for cls in (QGFieldNHN22, QGFieldNH18):
    for keep in (False, True):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        q = xi.QGDataset(ds, qgfield=cls, keep_3d=keep, chunk=128)
        q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{cls.__name__:14s} keep_3d={keep}: {nt} snapshots in {dt:.2f} s = {nt / dt:.0f} snapshots/s "
              f"(lwa_baro mean {float(np.mean(q.lwa_baro.data)):.4f})", flush=True)
        del q
