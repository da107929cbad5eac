"""GPU mirror of the structural assertions of the reference's tests/test_oopinterface.py:338-490 on the product classes
(real ERA-Interim 1.5 degree snapshot, like the reference's tests): property guards, shapes / finiteness of the
layer-wise and flux-vector outputs for both classes and both hemisphere modes, the exact lambda identity, and the
consistency of the flux vectors with the stored convergence / divergence."""
import numpy as np
import pytest

from oracle import pipeline, testdata

pytestmark = pytest.mark.gpu

XLON, YLAT, PLEV, U, V, T = testdata.load_era_snapshot()
NLAT, NLON, KMAX = YLAT.size, XLON.size, 49


def _cls(kind):
    from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
    return QGFieldNH18 if kind == "nh18" else QGFieldNHN22


def test_layerwise_flux_guard():
    f = _cls("nh18")(XLON, YLAT, PLEV, U, V, T, northern_hemisphere_results_only=True)
    f.interpolate_fields(); f.compute_reference_states()
    for name in ("ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term"):
        with pytest.raises(ValueError):
            getattr(f, name)


@pytest.mark.parametrize("nh_only", [False, True])
@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_layerwise_and_flux_vector_properties(kind, nh_only):
    f = _cls(kind)(XLON, YLAT, PLEV, U, V, T, northern_hemisphere_results_only=nh_only)
    f.interpolate_fields(); f.compute_reference_states()
    f.compute_layerwise_lwa_fluxes()
    shape3 = (KMAX, NLAT // 2 + 1, NLON) if nh_only else (KMAX, NLAT, NLON)
    for name in ("ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term"):
        a = getattr(f, name)
        assert a.shape == shape3, name
        assert np.isnan(a).sum() == 0 and np.abs(a).sum() > 0, name
    f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False)
    shape2 = shape3[1:]
    for name in ("flux_vector_lambda_baro", "flux_vector_phi_baro"):
        a = getattr(f, name)
        assert a.shape == shape2 and np.isnan(a).sum() == 0 and np.abs(a).sum() > 0, name


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_flux_vector_identity_and_divergence_consistency(kind):
    f = _cls(kind)(XLON, YLAT, PLEV, U, V, T)
    f.interpolate_fields(); f.compute_reference_states(); f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False)
    np.testing.assert_array_equal(f.flux_vector_lambda_baro, f.adv_flux_f1 + f.adv_flux_f2 + f.adv_flux_f3)
    clat = np.cos(np.deg2rad(YLAT))
    conv = pipeline.zonal_convergence(f.flux_vector_lambda_baro, clat, np.deg2rad(XLON[1] - XLON[0]), f.planet_radius)
    np.testing.assert_allclose(conv, f.convergence_zonal_advective_flux, rtol=1e-10,
                               atol=1e-12 * np.abs(f.convergence_zonal_advective_flux).max())
    # meridional divergence of -flux_vector_phi_baro against the stored eddy-momentum-flux divergence away from the
    # equatorial and polar boundary rows (utilities.py:236-279; reference test: atol 1e-5)
    field = -f.flux_vector_phi_baro
    dphi = np.deg2rad(YLAT[1] - YLAT[0])
    d = np.zeros_like(field)
    d[1:-1] = field[2:] * clat[2:, None] - field[:-2] * clat[:-2, None]
    fin = np.abs(clat) > 1e-5
    d[fin] = d[fin] / (f.planet_radius * clat[fin, None] * 2. * dphi)
    eqb, pole = 6, 3
    emf = f.divergence_eddy_momentum_flux
    assert np.abs(f.flux_vector_phi_baro).sum() > 0 and np.abs(emf).sum() > 0
    np.testing.assert_allclose(d[60 + eqb:-pole], emf[60 + eqb:-pole], atol=1e-5)
    np.testing.assert_allclose(d[pole:61 - eqb], emf[pole:61 - eqb], atol=1e-5)


@pytest.mark.parametrize("nh_only", [False, True])
@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_basic_qgdataset_calls(kind, nh_only):
    """tests/test_xarrayinterface.py:108-143, :216-275 of the reference: a dataset without leading dimensions, every
    compute method, datasets agree with the accessors, layer-wise and flux-vector variables with the right dims."""
    from hn2016_falwa_b200 import xarrayinterface as xi
    dims = ("plev", "ylat", "xlon")
    coords = dict(plev=PLEV, ylat=YLAT, xlon=XLON)
    ds = xi.SimpleDataset({n: xi.SimpleDataArray(a, dims, coords, name=n) for n, a in zip("uvt", (U, V, T))})
    q = xi.QGDataset(ds, qgfield=_cls(kind), qgfield_kwargs={"northern_hemisphere_results_only": nh_only})
    out1 = q.interpolate_fields()
    for name in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"):
        np.testing.assert_allclose(out1[name].data, getattr(q, name).data)
    assert "static_stability" in out1 or ("static_stability_n" in out1 and "static_stability_s" in out1)
    out2 = q.compute_reference_states()
    for name in ("qref", "uref", "ptref"):
        np.testing.assert_allclose(out2[name].data, getattr(q, name).data)
    out3 = q.compute_lwa_and_barotropic_fluxes()
    for name in ("adv_flux_f1", "adv_flux_f2", "adv_flux_f3", "convergence_zonal_advective_flux",
                 "divergence_eddy_momentum_flux", "meridional_heat_flux", "lwa_baro", "u_baro", "lwa",
                 "flux_vector_lambda_baro", "flux_vector_phi_baro"):
        np.testing.assert_allclose(out3[name].data, getattr(q, name).data)
    nrow = NLAT // 2 + 1 if nh_only else NLAT
    assert out3["flux_vector_lambda_baro"].dims == ("ylat", "xlon") and out3["lwa_baro"].shape == (nrow, NLON)
    layer = q.compute_layerwise_lwa_fluxes()
    for name in ("lwa", "ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term"):
        assert name in layer and layer[name].dims == ("height", "ylat", "xlon")
        assert layer[name].shape == (KMAX, nrow, NLON) and np.isfinite(layer[name].data).all()
    out4 = q.compute_lwa_only()
    for name in ("lwa_baro", "u_baro", "lwa"):
        np.testing.assert_allclose(out4[name].data, getattr(q, name).data)
    assert out4["lwa_baro"].shape == (nrow, NLON)
