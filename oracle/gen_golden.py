"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz — run in the build container only
(needs /root/reference):  ``python -m oracle.gen_golden``

What it pins
------------
1. 3-D path (configs 1 and 2): the reference's *unmodified* Python (falwa/oopinterface.py,
   data_storage.py, utilities.py, with SciPy's interp1d / UnivariateSpline / dgetrf / dgetri) is
   imported from /root/reference and driven on the repo's ERA-Interim snapshot, with the nine
   native names bound to the C++ restatement (the Fortran cannot be compiled here).  The outputs
   are (a) required to equal ``oracle.pipeline.run_qgfield`` bit for bit and (b) stored,
   sub-sampled, as fixtures.
2. 2-D path (config 5): the reference's unmodified ``BarotropicField`` (pure NumPy) on the
   analytic field of tests/test_barotropic_field.py and on a seeded field; required to equal
   ``oracle.basis2d`` and stored.  The reference test's inline answer key is stored alongside.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

from . import f2py_shim as shim
from . import pipeline, testdata

LON_STRIDE = 8
LEVELS_3D = (1, 10, 24, 40, 47)


def _subsample(name, a):
    a = np.asarray(a)
    if a.ndim == 3:
        return a[list(LEVELS_3D)][:, :, ::LON_STRIDE]
    if a.ndim == 2 and a.shape[-1] > 200:  # (nlat, nlon)
        return a[:, ::LON_STRIDE]
    return a


def _run_reference(kind, inputs):
    from falwa.oopinterface import QGFieldNH18, QGFieldNHN22
    cls = QGFieldNH18 if kind == "nh18" else QGFieldNHN22
    xlon, ylat, plev, u, v, t = inputs
    f = cls(xlon, ylat, plev, u, v, t, northern_hemisphere_results_only=False)
    with contextlib.redirect_stdout(io.StringIO()):
        f.interpolate_fields(return_named_tuple=False)
        f.compute_reference_states(return_named_tuple=False)
        f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False)
    res = {n: np.asarray(getattr(f, n)) for n in pipeline.USER_FIELDS}
    ss = f.static_stability
    res["static_stability"] = np.asarray(ss)
    return res


def main():
    warnings.filterwarnings("ignore")
    shim.set_precision(np.float64)
    shim.install_as_falwa_modules()
    inputs = testdata.load_era_snapshot()
    for kind in ("nh18", "nhn22"):
        ref = _run_reference(kind, inputs)
        mine = pipeline.run_qgfield(kind, *inputs)
        store = {}
        for name in pipeline.USER_FIELDS:
            a, b = ref[name], np.asarray(mine[name])
            assert a.shape == b.shape, (name, a.shape, b.shape)
            same = np.array_equal(a, b, equal_nan=True)
            print(f"{kind:6s} {name:36s} bit-identical to reference python: {same}"
                  + ("" if same else f"   max|diff|={np.nanmax(np.abs(a - b)):.3e}"))
            assert same, name
            store[name] = _subsample(name, a)
        assert np.array_equal(ref["static_stability"], np.asarray(mine["static_stability"]))
        store["static_stability"] = ref["static_stability"]
        if kind == "nh18":
            store["num_of_iter"] = np.asarray(mine["num_of_iter"])
        store["mask_count"] = np.asarray(mine["mask_count"])
        store["lon_stride"] = LON_STRIDE
        store["levels_3d"] = np.asarray(LEVELS_3D)
        out = os.path.join(testdata.GOLDEN_DIR, f"qgfield_{kind}_era_interim.npz")
        np.savez_compressed(out, **store)
        print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB")

    # ---- 2-D path -------------------------------------------------------------------------
    from falwa.barotropic_field import BarotropicField
    from . import basis2d
    cases = {}
    nlat, nlon = 31, 60  # tests/test_barotropic_field.py:7-22
    xlon = np.linspace(0, 360, nlon, endpoint=False)
    ylat = np.linspace(-90., 90., nlat, endpoint=True)
    lat_var = 8.e-5 * np.cos(np.deg2rad(ylat)) * np.exp(-(ylat - 36) ** 2 / 10 ** 2)
    pv = np.multiply(lat_var.reshape(nlat, 1), np.cos(3 * np.deg2rad(xlon)).reshape(1, nlon)) \
        + 2 * pipeline.EARTH_OMEGA * np.sin(np.deg2rad(ylat[:, np.newaxis]))
    cases["analytic"] = (xlon, ylat, pv)
    xl, yl, field = basis2d.synthetic_vorticity(nlat=121, nlon=240, seed=7)
    cases["seeded"] = (xl, yl, field)
    store = {}
    for name, (xl, yl, field) in cases.items():
        bf = BarotropicField(xl, yl, field)
        eq_ref, lwa_ref = np.asarray(bf.equivalent_latitudes), np.asarray(bf.lwa)
        eq, lwa = basis2d.barotropic_eqvlat_lwa(xl, yl, field)
        ok = np.array_equal(eq_ref, eq) and np.array_equal(lwa_ref, lwa)
        print(f"2-D {name}: oracle.basis2d bit-identical to reference BarotropicField: {ok}")
        assert ok
        store[f"{name}_xlon"], store[f"{name}_ylat"], store[f"{name}_pv"] = xl, yl, field
        store[f"{name}_eqvlat"], store[f"{name}_lwa"] = eq_ref, lwa_ref
    out = os.path.join(testdata.GOLDEN_DIR, "barotropic_field.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB")


if __name__ == "__main__":
    sys.exit(main())
