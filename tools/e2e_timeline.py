"""Timeline (CUDA events) of HostStream's copy-in / compute / copy-out phases.  usage: e2e_timeline.py [f64|f32|i16]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hn2016_falwa_b200 import engine
from hn2016_falwa_b200.oopinterface import QGFieldNHN22
from hn2016_falwa_b200.synthetic import synthetic_snapshot
mode = sys.argv[1] if len(sys.argv) > 1 else "f32"
xlon, ylat, plev, u, v, t = synthetic_snapshot(nlon=1440, nlat=721, seed=1234)
plan = QGFieldNHN22(xlon, ylat, plev, u, v, t)._plan
B = 4
dt = {"f64": torch.float64, "f32": torch.float32, "i16": torch.int16}[mode]
kw, hin = {}, []
for x in (u, v, t):
    if mode == "i16":
        lo, hi = x.min(), x.max(); sc, off = (hi - lo) / 65534., 0.5 * (hi + lo)
        kw.setdefault("packing", []).append((sc, off))
        x = np.round((x[::-1, ::-1] - off) / sc).astype(np.int16)
    h = engine.HostStream.pinned((B,) + x.shape, dt)
    for b in range(B): h[b].copy_(torch.from_numpy(np.roll(x, 37 * b, axis=-1).copy()))
    hin.append(h)
if mode == "i16":
    kw.update(flip_lat=True, flip_lev=True)
hs = engine.HostStream(plan, chunk=1, want_lwa3d=True, input_dtype=dt, **kw)
out = hs.alloc_outputs(B)
hs.run(*hin, out=out, cycles=2)
hs.trace = []
torch.cuda.synchronize()
z = torch.cuda.Event(enable_timing=True); z.record()
t0 = time.perf_counter()
hs.run(*hin, out=out, cycles=3)
torch.cuda.synchronize()
print(f"{mode}: {12 / (time.perf_counter() - t0):.1f} snapshots/s")
rows = sorted((z.elapsed_time(a), z.elapsed_time(b), ph, c) for ph, c, a, b in hs.trace)
for i, (a, b, ph, c) in enumerate(rows):
    print(f"{ph:8s} chunk@{c}  {a:8.2f} -> {b:8.2f}  ({b - a:6.2f} ms)")
