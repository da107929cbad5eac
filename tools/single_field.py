"""Wall-clock of the drop-in class API for single snapshots (BASELINE.json configs 1-3): NumPy in -> NumPy out."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from oracle import testdata
from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
from hn2016_falwa_b200.synthetic import synthetic_snapshot

def run(tag, cls, args, reps=4):
    for r in range(reps):
        torch.cuda.synchronize(); t = [time.perf_counter()]
        f = cls(*args); torch.cuda.synchronize(); t.append(time.perf_counter())
        f.interpolate_fields(return_named_tuple=False); torch.cuda.synchronize(); t.append(time.perf_counter())
        f.compute_reference_states(return_named_tuple=False); torch.cuda.synchronize(); t.append(time.perf_counter())
        f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False); torch.cuda.synchronize(); t.append(time.perf_counter())
        x = (f.lwa_baro, f.u_baro, f.adv_flux_f1, f.qref, f.uref, f.lwa); t.append(time.perf_counter())
        d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
        print(f"{tag:28s} rep{r}: ctor {d[0]:7.1f}  interp {d[1]:7.1f}  refstates {d[2]:7.1f}  flux {d[3]:7.1f}  "
              f"outputs {d[4]:7.1f}  total {sum(d):8.1f} ms", flush=True)

era = testdata.load_era_snapshot()
run("NH18  ERA-Interim 1.5deg", QGFieldNH18, era)
run("NHN22 ERA-Interim 1.5deg", QGFieldNHN22, era)
big = synthetic_snapshot(nlon=1440, nlat=721, seed=1234)
run("NHN22 synthetic 0.25deg", QGFieldNHN22, big)
