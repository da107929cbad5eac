"""GPU test of QGDataset's device runner (BASELINE.json config 4 in miniature): a (time, member) batch through
hn2016_falwa_b200.xarrayinterface.QGDataset equals each snapshot through QGFieldNHN22 / QGFieldNH18 on its own and the
CPU oracle; leading dimensions are restored like the reference (tests/test_xarrayinterface.py:75-106)."""
import numpy as np
import pytest

from hn2016_falwa_b200.synthetic import synthetic_snapshot

pytestmark = pytest.mark.gpu


def _make(nt, nm):
    from hn2016_falwa_b200 import xarrayinterface as xi
    snaps = [synthetic_snapshot(nlon=96, nlat=49, seed=21, t_index=k) for k in range(nt * nm)]
    xlon, ylat, plev = snaps[0][:3]
    u, v, t = (np.stack([s[i] for s in snaps]).reshape((nt, nm) + snaps[0][3].shape) for i in (3, 4, 5))
    # files usually come north->south and top->bottom
    u, v, t = (x[:, :, ::-1, ::-1, :] for x in (u, v, t))
    coords = dict(time=np.arange(nt), number=np.arange(nm), level=plev[::-1], latitude=ylat[::-1], longitude=xlon)
    dims = ("time", "number", "level", "latitude", "longitude")
    mk = lambda a, n: xi.SimpleDataArray(a, dims, coords, name=n)
    return xi.SimpleDataset(dict(u=mk(u, "u"), v=mk(v, "v"), t=mk(t, "t"))), snaps


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_dataset_equals_single_fields_and_oracle(kind):
    from hn2016_falwa_b200 import xarrayinterface as xi
    from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
    from oracle import pipeline
    cls = QGFieldNH18 if kind == "nh18" else QGFieldNHN22
    ds, snaps = _make(2, 2)
    q = xi.QGDataset(ds, qgfield=cls, chunk=3)          # 4 snapshots in chunks of 3 + 1
    interp = q.interpolate_fields()
    assert interp["qgpv"].dims == ("time", "number", "height", "ylat", "xlon")
    assert interp["qgpv"].shape == (2, 2, 49, 49, 96)
    ref = q.compute_reference_states()
    flux = q.compute_lwa_and_barotropic_fluxes()
    assert set(flux) >= {"lwa_baro", "u_baro", "adv_flux_f1", "convergence_zonal_advective_flux", "lwa"}
    assert len(q.fields) == 4 and q.attrs["protocol"] == cls.__name__
    idx = 3
    f = cls(*snaps[idx])
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
    want = pipeline.run_qgfield(kind, *snaps[idx])
    for name in ("qgpv", "uref", "lwa_baro", "adv_flux_f2", "lwa", "divergence_eddy_momentum_flux"):
        got = np.asarray(getattr(q, name).data).reshape((4,) + np.asarray(getattr(f, name)).shape)[idx]
        np.testing.assert_allclose(got, getattr(f, name), rtol=1e-11, atol=1e-12 * np.abs(want[name]).max())
        np.testing.assert_allclose(got, want[name], rtol=1e-8, atol=1e-9 * np.abs(want[name]).max())
    stab = q.static_stability
    if kind == "nh18":
        assert stab.shape == (2, 2, 49)
    else:
        assert stab[0].name == "static_stability_n" and stab[0].shape == (2, 2, 49)


def test_dataset_layerwise_and_lwa_only():
    from hn2016_falwa_b200 import xarrayinterface as xi
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22, QGFieldNH18
    ds, snaps = _make(2, 1)
    q = xi.QGDataset(ds, qgfield=QGFieldNHN22)
    out = q.compute_layerwise_lwa_fluxes()
    assert set(out) == {"lwa", "ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term"}
    assert out["ua2"].shape == (2, 1, 49, 49, 96)
    f = QGFieldNHN22(*snaps[1])
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_layerwise_lwa_fluxes()
    for name in ("lwa", "ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term"):   # (histogram sums use atomics)
        want = np.asarray(getattr(f, name))
        np.testing.assert_allclose(out[name].data[1, 0], want, rtol=1e-9, atol=1e-10 * np.abs(want).max())
    # heating rate -> ncforce (device-batched) -> layer-wise fluxes with the non-conservative term
    hr = 1e-5 * np.random.default_rng(4).standard_normal((2, 1) + snaps[0][3].shape)
    # (the reference hands heating_rate.sel(coord) to the field as it is — xarrayinterface.py:655-660 — so it is
    # given in the field's orientation, unlike u, v, t)
    nc = q.compute_ncforce_from_heating_rate(hr)
    assert nc.shape == (2, 1, 49, 49, 96)
    want = f.compute_ncforce_from_heating_rate(hr[1, 0])
    np.testing.assert_allclose(nc[1, 0], want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    out = q.compute_layerwise_lwa_fluxes(ncforce=nc)
    f.compute_layerwise_lwa_fluxes(ncforce=want)
    for name in ("lwa", "ua2", "ep3"):
        w = np.asarray(getattr(f, name))
        np.testing.assert_allclose(out[name].data[1, 0], w, rtol=1e-9, atol=1e-10 * np.abs(w).max())
    q18 = xi.QGDataset(ds, qgfield=QGFieldNH18, qgfield_kwargs=dict(maxit=5, raise_error_for_nonconvergence=False))
    lo = q18.compute_lwa_only()
    assert lo["lwa"].shape == (2, 1, 49, 49, 96) and np.isfinite(lo["lwa_baro"].data).all()


def test_config4_shape_at_the_real_grid_against_the_oracle():
    """BASELINE.json config 4 at its real grid (240 x 121, 37 levels -> kmax 49): a time series holding the reference
    repo's ERA-Interim snapshot among seeded synthetic ones goes through QGDataset in chunks; every snapshot is compared
    with the CPU oracle (all 2-D budget terms, the reference states and the 3-D LWA)."""
    from hn2016_falwa_b200 import xarrayinterface as xi
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22
    from oracle import pipeline, testdata
    era = testdata.load_era_snapshot()
    snaps = [synthetic_snapshot(nlon=240, nlat=121, seed=31, t_index=k) for k in range(5)]
    snaps.insert(2, era)
    xlon, ylat, plev = era[:3]
    u, v, t = (np.stack([s[i] for s in snaps]) for i in (3, 4, 5))
    coords = dict(time=np.arange(len(snaps)), level=plev, latitude=ylat, longitude=xlon)
    dims = ("time", "level", "latitude", "longitude")
    mk = lambda a, n: xi.SimpleDataArray(a, dims, coords, name=n)
    q = xi.QGDataset(xi.SimpleDataset(dict(u=mk(u, "u"), v=mk(v, "v"), t=mk(t, "t"))), qgfield=QGFieldNHN22, chunk=4)
    q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    names = ("qref", "uref", "ptref", "lwa_baro", "u_baro", "adv_flux_f1", "adv_flux_f2", "adv_flux_f3",
             "convergence_zonal_advective_flux", "divergence_eddy_momentum_flux", "meridional_heat_flux", "lwa")
    for k, s in enumerate(snaps):
        want = pipeline.run_qgfield("nhn22", *s)
        for name in names:
            rtol, atol = (1e-9, 0.0) if name == "ptref" else (1e-8, 1e-9 * float(np.abs(want[name]).max()))
            np.testing.assert_allclose(np.asarray(getattr(q, name).data)[k], want[name], rtol=rtol, atol=atol,
                                       err_msg=f"snapshot {k}: {name}")
