"""CPU tests of integrate_budget / hemisphere_to_globe (reference xarrayinterface.py:664-791 and
tests/test_xarrayinterface.py:168-213) on the xarray-free stand-ins."""
import numpy as np

from hn2016_falwa_b200 import xarrayinterface as xi


def test_hemisphere_to_globe_mirrors_and_negates_v():
    ylat = np.array([0., 30., 60., 90.])
    dims = ("time", "ylat", "xlon")
    coords = dict(time=np.arange(2), ylat=ylat, xlon=np.arange(3.))
    u = np.arange(24.).reshape(2, 4, 3)
    ds = xi.SimpleDataset(dict(u=xi.SimpleDataArray(u, dims, coords, "u"), v=xi.SimpleDataArray(u + 100, dims, coords, "v")))
    g = xi.hemisphere_to_globe(ds)
    np.testing.assert_array_equal(g["u"].coords["ylat"], [-90., -60., -30., 0., 30., 60., 90.])
    np.testing.assert_array_equal(g["u"].data[:, :3], u[:, :0:-1])
    np.testing.assert_array_equal(g["u"].data[:, 3:], u)
    np.testing.assert_array_equal(g["v"].data[:, :3], -(u + 100)[:, :0:-1])
    # southern input (ends at the equator)
    ds_s = xi.SimpleDataset(dict(u=xi.SimpleDataArray(u, dims, dict(coords, ylat=-ylat[::-1]), "u")))
    g = xi.hemisphere_to_globe(ds_s)
    np.testing.assert_array_equal(g["u"].coords["ylat"], [-90., -60., -30., 0., 30., 60., 90.])
    np.testing.assert_array_equal(g["u"].data[:, 4:], u[:, -2::-1])


def test_integrate_budget_closes():
    nt, ny, nx = 5, 3, 4
    t = np.arange(nt) * 21600.0            # seconds, 6-hourly
    rng = np.random.default_rng(0)
    dims = ("time", "ylat", "xlon")
    coords = dict(time=t, ylat=np.arange(ny), xlon=np.arange(nx))
    mk = lambda n: xi.SimpleDataArray(rng.standard_normal((nt, ny, nx)), dims, coords, n)
    ds = xi.SimpleDataset({n: mk(n) for n in ("lwa_baro", "convergence_zonal_advective_flux",
                                               "divergence_eddy_momentum_flux", "meridional_heat_flux")})
    out = xi.integrate_budget(ds)
    assert out["delta_lwa"].dims == ("ylat", "xlon")
    np.testing.assert_allclose(out["delta_lwa"].data, ds["lwa_baro"].data[-1] - ds["lwa_baro"].data[0])
    want = np.trapezoid(ds["meridional_heat_flux"].data, x=t, axis=0)
    np.testing.assert_allclose(out["integrated_meridional_heat_flux"].data, want)
    total = sum(out[k].data for k in ("integrated_convergence_zonal_advective_flux",
                                      "integrated_divergence_eddy_momentum_flux", "integrated_meridional_heat_flux",
                                      "residual"))
    np.testing.assert_allclose(total, out["delta_lwa"].data, atol=1e-9)
    assert out.attrs["integration_seconds"] == 4 * 21600.0
