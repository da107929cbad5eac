"""The prefix-sum formulation of the LWA sums (csrc/lwa_prefix.cu) restated in NumPy against the oracle's double loop
(compute_flux_dirinv.f90:70-116) on the reference's real snapshot: exact pair counts, sums within rounding.  The NH18
thresholds of its northern hemisphere are non-monotone in 51 rows — the exception path."""
import numpy as np
import pytest

from oracle import f2py_shim as ora
from oracle import pipeline, testdata
from tests.lwa_prefix_numpy import prefix_lwa


def _f(a):
    return np.swapaxes(a, 0, 2)


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_prefix_formulation_matches_the_double_loop(kind):
    o = pipeline.run_qgfield(kind, *testdata.load_era_snapshot(), stages=("interp", "ref"))
    c = o["config"]
    eq, jd = c["equator_idx"], c["jd"]
    pv, uu, vv, pt = (_f(o[k]) for k in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"))
    ncf = np.random.default_rng(5).standard_normal(pv.shape) * 1e-5
    qref_cu, uref, ptref = o["qref"].T, o["uref"].T, o["ptref"].T
    nexc = 0
    for hemi in ("north", "south"):
        if hemi == "north":
            kw = dict(pv=pv, uu=uu, vv=vv, pt=pt, ncforce=ncf, tn0=o["tn0"], qref=qref_cu[-eq:], uref=uref[-jd:],
                      tref=ptref[-jd:])
        else:
            kw = dict(pv=-pv[:, ::-1, :], uu=uu[:, ::-1, :], vv=-vv[:, ::-1, :], pt=pt[:, ::-1, :], ncforce=ncf[:, ::-1, :],
                      tn0=o["ts0"], qref=-qref_cu[(eq - 1)::-1], uref=uref[(jd - 1)::-1], tref=ptref[(jd - 1)::-1])
        kw.update(jb=c["jb"], is_nhem=True, a=c["a"], om=c["omega"], dz=c["dz"], h=c["h"], rr=c["rr"], cp=c["cp"],
                  prefac=c["prefactor"], return_mask_count=True)
        w = ora.compute_flux_dirinv_nshem(**kw)
        a1, a2, nc3, u2, counts, ne = prefix_lwa(np.asarray(kw["pv"]), np.asarray(kw["uu"]), np.asarray(kw["ncforce"]),
                                                 np.asarray(kw["qref"]), np.asarray(kw["uref"]), c["jb"], c["a"], None)
        nexc += ne
        assert tuple(counts) == tuple(int(x) for x in w[9])                       # set membership is exact
        for got, want in ((a1, w[0]), (a2, w[1]), (nc3, w[2]), (u2, w[4])):
            np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-13 * np.abs(want).max())
    if kind == "nh18":
        assert nexc > 0     # the non-monotone thresholds of this snapshot exercise the exception rows
