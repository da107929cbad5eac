"""TEST INFRASTRUCTURE.  Fixtures for the remaining constructor / call modes of the QGField classes, produced by the
reference's unmodified Python (+ the C++ restatement of the native routines) on a seeded 96x49 snapshot:
  * data_on_evenly_spaced_pseudoheight_grid=True (reference tests/test_oopinterface.py:290-335),
  * northern_hemisphere_results_only=True for NHN22 and NH18,
  * compute_lwa_and_barotropic_fluxes(ncforce=...).
Run in the build container only:  ``python -m oracle.gen_golden_modes``"""
import contextlib
import io
import os
import warnings

import numpy as np

from . import f2py_shim as shim
from . import testdata

NAMES = ("qgpv", "qref", "uref", "ptref", "lwa_baro", "u_baro", "adv_flux_f1", "adv_flux_f2", "adv_flux_f3",
         "convergence_zonal_advective_flux", "divergence_eddy_momentum_flux", "meridional_heat_flux", "ncforce_baro")


def inputs():
    return testdata.synthetic_snapshot(nlon=96, nlat=49, seed=17)


def height_grid_inputs(kmax=33):
    """The same snapshot pre-interpolated to an evenly spaced pseudo-height grid (the user does the interpolation)."""
    from scipy.interpolate import interp1d
    xlon, ylat, plev, u, v, t = inputs()
    zlev = -7000. * np.log(plev / 1000.)
    height = np.arange(0, kmax) * 1000.
    new_plev = 1000. * np.exp(-height / 7000.)
    f = lambda a: interp1d(zlev, a, axis=0, kind="linear", fill_value="extrapolate")(height)
    return xlon, ylat, new_plev, f(u), f(v), f(t)


def ncforce_field(shape):
    return 1e-9 * np.random.default_rng(23).standard_normal(shape)


def run(f, ncforce=None):
    with contextlib.redirect_stdout(io.StringIO()):
        f.interpolate_fields(return_named_tuple=False)
        f.compute_reference_states(return_named_tuple=False)
        f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False, ncforce=ncforce)
    return {n: np.asarray(getattr(f, n)) for n in NAMES}


def main():
    warnings.filterwarnings("ignore")
    shim.set_precision(np.float64)
    shim.install_as_falwa_modules()
    from falwa.oopinterface import QGFieldNH18, QGFieldNHN22
    store = {}
    for kind, cls in (("nh18", QGFieldNH18), ("nhn22", QGFieldNHN22)):
        for k, v in run(cls(*height_grid_inputs(), data_on_evenly_spaced_pseudoheight_grid=True)).items():
            store[f"{kind}_zgrid_{k}"] = v[::4] if v.ndim == 3 else v
        for k, v in run(cls(*inputs(), northern_hemisphere_results_only=True)).items():
            store[f"{kind}_nhonly_{k}"] = v[::6] if v.ndim == 3 else v
        for k, v in run(cls(*inputs()), ncforce=ncforce_field((49, 49, 96))).items():
            store[f"{kind}_ncforce_{k}"] = v[::6] if v.ndim == 3 else v
    out = os.path.join(testdata.GOLDEN_DIR, "modes_synthetic.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
