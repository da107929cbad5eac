"""TEST INFRASTRUCTURE.  Fixture for inputs on an even latitude grid without the equator (the constructor regrids to
an odd grid and the outputs are interpolated back; /root/reference/src/falwa/oopinterface.py:147-154, :345-403;
reference tests/test_oopinterface.py:230-287).  Reference's unmodified Python + the C++ restatement.
Run in the build container only:  ``python -m oracle.gen_golden_evengrid``"""
import contextlib
import io
import os
import warnings

import numpy as np

from . import f2py_shim as shim
from . import testdata

NAMES = ("qgpv", "qref", "uref", "ptref", "lwa_baro", "u_baro", "adv_flux_f1", "adv_flux_f2", "adv_flux_f3",
         "convergence_zonal_advective_flux", "divergence_eddy_momentum_flux", "meridional_heat_flux", "lwa")


def even_grid_inputs():
    nlat, nlon = 48, 96
    xlon, _, plev, _, _, _ = testdata.synthetic_snapshot(nlon=nlon, nlat=49, seed=9)
    ylat = np.linspace(-90. + 90. / nlat, 90. - 90. / nlat, nlat)      # cell centres: no equator, no poles
    # sample the analytic generator on the even grid by generating on a fine odd grid and interpolating
    _, yf, _, u, v, t = testdata.synthetic_snapshot(nlon=nlon, nlat=193, seed=9)
    from scipy.interpolate import interp1d
    u, v, t = (interp1d(yf, x, axis=1)(ylat) for x in (u, v, t))
    return xlon, ylat, plev, u, v, t


def main():
    warnings.filterwarnings("ignore")
    shim.set_precision(np.float64)
    shim.install_as_falwa_modules()
    from falwa.oopinterface import QGFieldNH18, QGFieldNHN22
    inputs = even_grid_inputs()
    store = {}
    for kind, cls in (("nh18", QGFieldNH18), ("nhn22", QGFieldNHN22)):
        f = cls(*inputs)
        assert f.need_latitude_interpolation
        with contextlib.redirect_stdout(io.StringIO()):
            f.interpolate_fields(return_named_tuple=False)
            f.compute_reference_states(return_named_tuple=False)
            f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False)
        for n in NAMES:
            a = np.asarray(getattr(f, n))
            store[f"{kind}_{n}"] = a[::6] if a.ndim == 3 else a
    out = os.path.join(testdata.GOLDEN_DIR, "evengrid_synthetic.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
