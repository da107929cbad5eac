"""GPU parity of hn2016_falwa_b200.basis / .wrapper (the reference's falwa.basis / falwa.wrapper)
with the golden vectors of the reference's own modules (tests/golden/wrapper_2d.npz) and the oracle.

Tolerances: the area histograms are accumulated with floating-point atomics (order-dependent in the last bits), so
quantities derived from them carry rtol 1e-10; the LWA line integrals for a GIVEN qref are sequential sums in the
reference's order and must agree to the last bit."""
import os

import numpy as np
import pytest

from oracle import basis2d, testdata

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(testdata.GOLDEN_DIR, "wrapper_2d.npz"))
Y, PV, AREA, DMU = G["ylat"], G["pv"], G["area"], G["dmu"]
NLAT, NLON = PV.shape


def close(a, b, rtol=1e-10):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * np.abs(b).max())


def test_basis_eqvlat_fawa_vgrad():
    from hn2016_falwa_b200 import basis
    q, f = basis.eqvlat_fawa(Y, PV, AREA, NLAT)
    assert q.shape == (NLAT,) and f.shape == (NLAT - 1,)
    close(q, G["fawa_qref"]); close(f, G["fawa_fawa"], 1e-9)
    q, f = basis.eqvlat_fawa(Y, PV, AREA, 31, output_fawa=False)
    assert f is None
    close(q, G["fawa_qref_31"])
    q, b = basis.eqvlat_vgrad(Y, PV, AREA, NLAT, vgrad=G["vgrad"])
    close(q, G["vgrad_q"]); close(b, G["vgrad_brac"], 1e-9)
    q, b = basis.eqvlat_vgrad(Y, PV, AREA, 40)
    assert b is None
    close(q, G["vgrad_q_none"])


def test_basis_lwa_is_bit_exact_for_a_given_qref():
    from hn2016_falwa_b200 import basis
    parts, sig = basis.lwa(NLON, NLAT, PV, G["fawa_qref"], DMU, return_partitioned_lwa=True, ncforce=G["ncforce"])
    np.testing.assert_array_equal(parts, G["lwa_parts"])
    np.testing.assert_array_equal(sig, G["lwa_sigma"])
    total, none = basis.lwa(NLON, NLAT, PV, G["fawa_qref"], DMU)
    assert none is None
    np.testing.assert_array_equal(total, G["lwa_parts"][0] + G["lwa_parts"][1])


def test_basis_batched_and_device_tensors():
    import torch
    from hn2016_falwa_b200 import basis
    fields = np.stack([PV, PV[:, ::-1].copy(), 0.5 * PV])
    q, f = basis.eqvlat_fawa(Y, torch.from_numpy(fields).cuda(), AREA, NLAT)
    assert q.is_cuda and tuple(q.shape) == (3, NLAT) and tuple(f.shape) == (3, NLAT - 1)
    for i in range(3):
        qo, fo = basis2d.eqvlat_fawa(Y, fields[i], AREA, NLAT)
        close(q[i].cpu().numpy(), qo); close(f[i].cpu().numpy(), fo, 1e-9)


def test_wrappers_against_reference_golden():
    from hn2016_falwa_b200 import wrapper
    q, l = wrapper.barotropic_eqlat_lwa(Y, PV, AREA, DMU, NLAT)
    close(q, G["w_baro_qref"]); close(l, G["w_baro_lwa"], 1e-9)
    l, none = wrapper.barotropic_input_qref_to_compute_lwa(Y, G["fawa_qref"], PV, AREA, DMU)
    assert none is None
    np.testing.assert_array_equal(l, G["w_baro_input_qref_lwa"])
    close(wrapper.eqvlat_hemispheric(Y, PV, AREA), G["w_eqvlat_hemi"])
    q, l = wrapper.qgpv_eqlat_lwa(Y, PV, AREA, DMU)
    close(q, G["w_qgpv_qref"]); close(l, G["w_qgpv_lwa"], 1e-9)
    q, l, c = wrapper.qgpv_eqlat_lwa_ncforce(Y, PV, G["ncforce"], AREA, DMU, n_points=30)
    close(q, G["w_ncf_qref"]); close(l, G["w_ncf_lwa"], 1e-9); close(c, G["w_ncf_capsigma"], 1e-9)
    o = wrapper.qgpv_eqlat_lwa_options(Y, PV, AREA, DMU, ncforce=G["ncforce"])
    assert set(o) == {"qref", "lwa_result", "capsigma"}
    close(o["qref"], G["w_opt_qref"]); close(o["lwa_result"], G["w_opt_lwa"], 1e-9)
    close(o["capsigma"], G["w_opt_capsigma"], 1e-9)
    np.testing.assert_array_equal(wrapper.qgpv_input_qref_to_compute_lwa(Y, G["w_qgpv_qref"], PV, AREA, DMU),
                                  G["w_input_qref_lwa"])
    q, l = wrapper.theta_lwa(Y, G["theta"], AREA, DMU)
    close(q, G["w_theta_qref"]); close(l, G["w_theta_lwa"], 1e-9)


def test_wrappers_that_fail_upstream_follow_the_oracle():
    """eqvlat_bracket_hemispheric and qgpv_eqlat_lwa_options(vgrad=...) raise TypeError upstream; here they do what
    their docstrings say (basis.eqvlat_vgrad per hemisphere)."""
    from hn2016_falwa_b200 import wrapper
    q, b = wrapper.eqvlat_bracket_hemispheric(Y, PV, AREA, vgrad=G["vgrad"])
    o = basis2d.hemispheric(Y, PV, AREA, DMU, negate_back=False, vgrad=G["vgrad"], with_lwa=False, vgrad_whole=True)
    close(q, o["qref"]); close(b, o["brac_result"], 1e-9)
    r = wrapper.qgpv_eqlat_lwa_options(Y, PV, AREA, DMU, vgrad=G["vgrad"], ncforce=G["ncforce"])
    o = basis2d.hemispheric(Y, PV, AREA, DMU, vgrad=G["vgrad"], ncforce=G["ncforce"])
    assert set(r) == {"qref", "lwa_result", "brac_result", "capsigma"}
    close(r["qref"], o["qref"]); close(r["brac_result"], o["brac_result"], 1e-9)
    close(r["lwa_result"], o["lwa_result"], 1e-9); close(r["capsigma"], o["capsigma"], 1e-9)


def test_errors():
    from hn2016_falwa_b200 import basis
    with pytest.raises(AssertionError):
        basis.eqvlat_fawa(Y[:-1], PV, AREA, NLAT)
    with pytest.raises(ValueError):
        basis.eqvlat_fawa(Y, PV, AREA, 1)


def test_even_latitude_count_and_explicit_hemisphere_size():
    from hn2016_falwa_b200 import basis, wrapper
    y, pv, area, dmu, th = (G["even_" + k] for k in ("ylat", "pv", "area", "dmu", "theta"))
    q, l = wrapper.qgpv_eqlat_lwa(y, pv, area, dmu)
    close(q, G["even_qgpv_qref"]); close(l, G["even_qgpv_lwa"], 1e-9)
    q, l = wrapper.theta_lwa(y, th, area, dmu, nlat_s=20, n_points=25)
    close(q, G["even_theta_qref"]); close(l, G["even_theta_lwa"], 1e-9)
    assert np.all(l[20:28] == 0.0)
    q, f = basis.eqvlat_fawa(y, pv, area, 48)
    close(q, G["even_fawa_qref"]); close(f, G["even_fawa"], 1e-9)


def test_reference_basis_tests_on_the_gpu():
    """The reference's own tests/test_basis.py, run against hn2016_falwa_b200.basis: LWA of a zonally symmetric field
    is exactly zero; eqvlat_vgrad gives a non-decreasing Q(y) that is IDENTICAL with and without vgrad (which needs
    the area histogram to be bit-reproducible from launch to launch)."""
    from math import pi
    from hn2016_falwa_b200 import basis
    from hn2016_falwa_b200.constant import EARTH_RADIUS
    test_vort = (np.ones((5, 5)) * np.array([1, 2, 3, 4, 5])).swapaxes(0, 1)
    res, _ = basis.lwa(5, 5, test_vort, np.array([1, 2, 3, 4, 5]), np.ones(5))
    assert np.array_equal(res, np.zeros((5, 5)))
    nlat, nlon, planet_radius = 241, 480, 1.
    ylat = np.linspace(-90, 90, nlat, endpoint=True)
    xlon = np.linspace(0, 360, nlon, endpoint=False)
    vort = np.sin(3. * np.deg2rad(xlon[np.newaxis, :])) * np.cos(np.deg2rad(ylat[:, np.newaxis]))
    dummy_vgrad = 3. / (planet_radius * np.cos(np.deg2rad(ylat[:, np.newaxis]))) \
        * np.cos(np.deg2rad(3. * xlon[np.newaxis, :])) * np.cos(np.deg2rad(ylat[:, np.newaxis])) \
        - 1. / planet_radius * np.sin(np.deg2rad(ylat[:, np.newaxis])) * np.sin(np.deg2rad(3. * xlon[np.newaxis, :]))
    dphi = np.deg2rad(np.diff(ylat)[0])
    area = 2. * pi * planet_radius ** 2 * (np.cos(np.deg2rad(ylat[:, np.newaxis])) * dphi) / float(nlon) \
        * np.ones((nlat, nlon))
    q1, vg = basis.eqvlat_vgrad(ylat, vort, area, nlat, planet_radius=EARTH_RADIUS, vgrad=dummy_vgrad)
    q2, none = basis.eqvlat_vgrad(ylat, vort, area, nlat, planet_radius=EARTH_RADIUS, vgrad=None)
    assert np.all(np.diff(q1) >= 0.)
    assert q1.tolist() == q2.tolist()
    assert none is None and vg is not None and vg.shape == q1.shape
    qo, vo = basis2d.eqvlat_vgrad(ylat, vort, area, nlat, planet_radius=EARTH_RADIUS, vgrad=dummy_vgrad)
    close(q1, qo); close(vg, vo, 1e-9)
    # launch-to-launch reproducibility of the whole 2-D path
    a = basis.eqvlat_fawa(ylat, vort, area, nlat, planet_radius=EARTH_RADIUS)
    b = basis.eqvlat_fawa(ylat, vort, area, nlat, planet_radius=EARTH_RADIUS)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
