import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hn2016_falwa_b200 import engine
from hn2016_falwa_b200.oopinterface import QGFieldNHN22
from hn2016_falwa_b200.synthetic import synthetic_snapshot
xlon, ylat, plev, u, v, t = synthetic_snapshot(nlon=1440, nlat=721, seed=1234)
plan = QGFieldNHN22(xlon, ylat, plev, u, v, t)._plan
B = 4
def host(dt):
    out = []
    for x in (u, v, t):
        h = engine.HostStream.pinned((B,) + x.shape, dt)
        for b in range(B): h[b].copy_(torch.from_numpy(np.roll(x, 37 * b, axis=-1)))
        out.append(h)
    return out
for dt in (torch.float64, torch.float32, torch.int16):
    hin = host(dt if dt != torch.int16 else torch.float64)
    kw = {}
    if dt == torch.int16:
        kw = dict(packing=[(0.01, 0.0)] * 3, flip_lat=True, flip_lev=True)
        hin = [engine.HostStream.pinned(x.shape, dt).copy_((x / 0.01).round().clamp(-32767, 32767)) for x in hin]
    for lwa3d in (True, False):
        for chunk in (1, 2):
            hs = engine.HostStream(plan, chunk=chunk, want_lwa3d=lwa3d, input_dtype=dt, **kw)
            out = hs.alloc_outputs(B)
            hs.run(*hin, out=out)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            hs.run(*hin, out=out, cycles=4)
            torch.cuda.synchronize(); dtm = time.perf_counter() - t0
            print(f"{str(dt):14s} lwa3d={lwa3d!s:5s} chunk={chunk}: {16 / dtm:6.1f} snapshots/s  ({dtm / 16 * 1e3:.1f} ms each)", flush=True)
            del hs, out
