"""TEST INFRASTRUCTURE.  Golden fixture of the layer-wise flux terms (compute_layerwise_lwa_fluxes,
compute_ncforce_from_heating_rate; /root/reference/src/falwa/oopinterface.py:985-1156): the reference's unmodified
Python driven on the ERA-Interim snapshot with the nine native names bound to the C++ restatement.
Run in the build container only:  ``python -m oracle.gen_golden_layerwise``"""
import contextlib
import io
import os
import warnings

import numpy as np

from . import f2py_shim as shim
from . import testdata
from .gen_golden import LEVELS_3D, LON_STRIDE

NAMES = ("lwa", "ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term")


def main():
    warnings.filterwarnings("ignore")
    shim.set_precision(np.float64)
    shim.install_as_falwa_modules()
    from falwa.oopinterface import QGFieldNH18, QGFieldNHN22
    inputs = testdata.load_era_snapshot()
    xlon, ylat, plev, u, v, t = inputs
    rng = np.random.default_rng(11)
    heating = 1e-5 * rng.standard_normal(u.shape)
    for kind, cls in (("nh18", QGFieldNH18), ("nhn22", QGFieldNHN22)):
        f = cls(xlon, ylat, plev, u, v, t)
        with contextlib.redirect_stdout(io.StringIO()):
            f.interpolate_fields(return_named_tuple=False)
            f.compute_reference_states(return_named_tuple=False)
            ncf = f.compute_ncforce_from_heating_rate(heating)
            f.compute_layerwise_lwa_fluxes(ncforce=ncf)
        store = {n: np.asarray(getattr(f, n))[list(LEVELS_3D)][:, :, ::LON_STRIDE] for n in NAMES}
        store["ncforce_input"] = np.asarray(ncf)[list(LEVELS_3D)][:, :, ::LON_STRIDE]
        store["heating_seed"] = 11
        out = os.path.join(testdata.GOLDEN_DIR, f"layerwise_{kind}_era_interim.npz")
        np.savez_compressed(out, **store)
        print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
