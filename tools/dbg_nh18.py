import sys; sys.path.insert(0, ".")
import numpy as np
from hn2016_falwa_b200.synthetic import synthetic_snapshot
from hn2016_falwa_b200.oopinterface import QGFieldNH18
from oracle import pipeline
snap = synthetic_snapshot(nlon=96, nlat=49, seed=21, t_index=3)
want = pipeline.run_qgfield("nh18", *snap)
for method in (0, 1):
    f = QGFieldNH18(*snap)
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False, method=method)
    print("method", method, "iters", f._batch.num_of_iter, want["num_of_iter"])
    for name in pipeline.USER_FIELDS:
        g, w = np.asarray(getattr(f, name)), np.asarray(want[name])
        err = np.abs(g - w).max() / max(np.abs(w).max(), 1e-300)
        if err > 1e-9:
            bad = np.argwhere(np.abs(g - w) > 1e-8 * np.abs(w).max())
            print(f"  {name:40s} relerr {err:.3e} nbad {len(bad)} first {bad[:3].tolist()} rows {sorted(set(bad[:, -2].tolist()))[:12] if bad.ndim == 2 and bad.shape[1] >= 2 else ''}")
