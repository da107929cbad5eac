"""CPU tests: the oracle against the committed golden fixtures (generated from the reference's own
unmodified Python, oracle/gen_golden.py), the reference tests' structural assertions
(/root/reference/tests/test_oopinterface.py) and the 2-D inline answer key
(/root/reference/tests/test_barotropic_field.py:26-54)."""
import os

import numpy as np
import pytest

from oracle import basis2d, pipeline, testdata

GOLD = testdata.GOLDEN_DIR


@pytest.fixture(scope="module")
def era():
    return testdata.load_era_snapshot()


@pytest.fixture(scope="module", params=["nh18", "nhn22"])
def run(request, era):
    return request.param, pipeline.run_qgfield(request.param, *era)


def test_pipeline_matches_reference_python_fixture(run):
    kind, out = run
    g = np.load(os.path.join(GOLD, f"qgfield_{kind}_era_interim.npz"))
    stride, levels = int(g["lon_stride"]), list(g["levels_3d"])
    for name in pipeline.USER_FIELDS:
        a = np.asarray(out[name])
        if a.ndim == 3:
            a = a[levels][:, :, ::stride]
        elif a.ndim == 2 and a.shape[-1] > 200:
            a = a[:, ::stride]
        # bit-for-bit: same NumPy/SciPy + the same compiled oracle; tolerate a different libm only
        np.testing.assert_allclose(a, g[name], rtol=1e-12, atol=0, err_msg=name)
    assert np.array_equal(np.asarray(out["mask_count"]), g["mask_count"])
    if kind == "nh18":
        assert tuple(out["num_of_iter"]) == tuple(g["num_of_iter"])


def test_structural_assertions_of_reference_tests(run, era):
    kind, out = run
    xlon, ylat, plev, u, v, t = era
    kmax, nlat, nlon = 49, ylat.size, xlon.size
    # test_oopinterface.py:46-67
    for name, src in (("interpolated_u", u), ("interpolated_v", v)):
        a = out[name][1:-1]
        assert src.min() <= a.min() and a.max() <= src.max()
    assert out["qgpv"].shape == (kmax, nlat, nlon)
    assert np.isnan(out["qgpv"]).sum() == 0 and np.isinf(out["qgpv"]).sum() == 0
    # qref non-decreasing in the northern hemisphere, test_oopinterface.py:78, :179-180
    nh = out["qref"][1:-1, nlat // 2 + 6:-1] if kind == "nhn22" else out["qref"][1:-1, nlat // 2 + 1:-1]
    assert (np.diff(nh, axis=-1) >= 0.).all()
    # flux_vector_lambda_baro == f1 + f2 + f3 exactly, test_oopinterface.py:410-425
    assert np.array_equal(out["flux_vector_lambda_baro"],
                          out["adv_flux_f1"] + out["adv_flux_f2"] + out["adv_flux_f3"])
    # convergence consistency, test_oopinterface.py:429-452
    c = out["config"]
    clat = np.abs(np.cos(np.deg2rad(ylat)))
    conv = pipeline.zonal_convergence(out["flux_vector_lambda_baro"], clat, c["dlambda"], c["a"])
    np.testing.assert_allclose(conv, out["convergence_zonal_advective_flux"], rtol=1e-10)


def test_sor_iteration_count_is_near_the_notebook_value(era):
    """demo_script_for_nh2018.ipynb cell 12 prints 951 (NH) / 727 (SH) for this very snapshot with the
    shipped fp32 build started from uninitialised memory; the zero-started restatement must land close."""
    out = pipeline.run_qgfield("nh18", *era, stages=("interp", "ref"))
    n, s = out["num_of_iter"]
    assert abs(n - 951) < 40 and abs(s - 727) < 40, (n, s)


def test_barotropic_answer_key():
    g = np.load(os.path.join(GOLD, "barotropic_field.npz"))
    eq, lwa = basis2d.barotropic_eqvlat_lwa(g["analytic_xlon"], g["analytic_ylat"], g["analytic_pv"])
    key = basis2d.ANSWER_KEY
    assert np.allclose(eq, key["eqv_lat"], rtol=1e-5, atol=1e-8)
    assert np.allclose(lwa.mean(axis=-1), key["lwa_zonal_mean"], rtol=1e-5, atol=1e-8)
    assert np.allclose(lwa[20, :], key["lwa_row20"], rtol=1e-5, atol=1e-8)
    for case in ("analytic", "seeded"):
        eq, lwa = basis2d.barotropic_eqvlat_lwa(g[f"{case}_xlon"], g[f"{case}_ylat"], g[f"{case}_pv"])
        np.testing.assert_allclose(eq, g[f"{case}_eqvlat"], rtol=1e-13)
        np.testing.assert_allclose(lwa, g[f"{case}_lwa"], rtol=1e-12, atol=1e-18)


def test_lwa_is_zero_for_zonally_symmetric_field():
    """reference tests/test_basis.py:27-38."""
    vort = (np.ones((5, 5)) * np.array([1, 2, 3, 4, 5])).swapaxes(0, 1)
    out = basis2d.lwa(5, 5, vort, np.array([1, 2, 3, 4, 5]), np.ones(5))
    assert np.array_equal(out, np.zeros((5, 5)))
