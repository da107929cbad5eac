"""Per-stage CUDA-event timing of the product pipeline (diagnostic; bench.py is the contract)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import testdata
from hn2016_falwa_b200 import engine
from hn2016_falwa_b200.oopinterface import QGFieldNHN22, QGFieldNH18


def main(nlon=1440, nlat=721, kind="nhn22", batch=1, method=1, reps=3):
    t0 = time.time()
    xlon, ylat, plev, u, v, t = testdata.synthetic_snapshot(nlon=nlon, nlat=nlat, seed=1234)
    print(f"synthetic {nlon}x{nlat} generated in {time.time() - t0:.1f}s", flush=True)
    cls = QGFieldNHN22 if kind == "nhn22" else QGFieldNH18
    f = cls(xlon, ylat, plev, u, v, t)
    plan = f._plan
    ub = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(u, (batch,) + u.shape))).cuda()
    vb = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(v, (batch,) + v.shape))).cuda()
    tb = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(t, (batch,) + t.shape))).cuda()
    res = {}
    for rep in range(reps):
        b = engine.QGBatch(plan, ub, vb, tb)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize()
        ev[0].record()
        prof = b.static_stability_profiles()
        ev[1].record()
        b.interpolate_fields()
        ev[2].record()
        b.compute_reference_states()
        ev[3].record()
        b.compute_lwa_and_barotropic_fluxes(method=method)
        ev[4].record()
        torch.cuda.synchronize()
        names = ["theta_means+spline", "interpolate_fields(total)", "reference_states", "lwa_fluxes"]
        res = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
        print(json.dumps(dict(rep=rep, kind=kind, grid=[nlon, nlat], batch=batch, method=method, ms=res)), flush=True)
    return res


if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, val = a.split("=")
        kw[k] = val if k == "kind" else int(val)
    main(**kw)
