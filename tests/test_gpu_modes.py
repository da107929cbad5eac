"""Constructor / call modes of the QGField classes against fixtures from the reference's own unmodified Python
(oracle/gen_golden_modes.py): data already on an evenly spaced pseudo-height grid (reference tests/test_oopinterface.py:
290-335), northern_hemisphere_results_only, and the non-conservative forcing argument of
compute_lwa_and_barotropic_fluxes."""
import os

import numpy as np
import pytest

from oracle import testdata
from oracle.gen_golden_modes import NAMES, height_grid_inputs, inputs, ncforce_field
from tests.parity import compare

pytestmark = pytest.mark.gpu


def _cls(kind):
    from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
    return QGFieldNH18 if kind == "nh18" else QGFieldNHN22


def _check(tag, f, g, prefix, stride):
    for n in NAMES:
        got = np.asarray(getattr(f, n))
        got = got[::stride] if got.ndim == 3 else got
        want = g[f"{prefix}_{n}"]
        assert got.shape == want.shape, (n, got.shape, want.shape)
        atol = 1e-9 * np.abs(want).max() if n != "qgpv" else 1e-17
        compare(f"modes/{tag}/{n}", got, want, rtol=1e-8, atol=atol)


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_modes(kind):
    g = np.load(os.path.join(testdata.GOLDEN_DIR, "modes_synthetic.npz"))
    cls = _cls(kind)
    f = cls(*height_grid_inputs(), data_on_evenly_spaced_pseudoheight_grid=True)
    assert f.kmax == 33
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
    _check(f"{kind}/zgrid", f, g, f"{kind}_zgrid", 4)
    f = cls(*inputs(), northern_hemisphere_results_only=True)
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
    assert f.qref.shape == (49, 25) and f.lwa_baro.shape == (25, 96)
    _check(f"{kind}/nhonly", f, g, f"{kind}_nhonly", 6)
    f = cls(*inputs())
    f.interpolate_fields(False); f.compute_reference_states(False)
    f.compute_lwa_and_barotropic_fluxes(False, ncforce=ncforce_field((49, 49, 96)))
    _check(f"{kind}/ncforce", f, g, f"{kind}_ncforce", 6)


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_height_grid_and_nh_only_through_the_batched_runners(kind):
    """The same modes through QGDataset's device runner and HostStream with chunks of one snapshot: stage 3 of chunk c
    is queued after the input slot of chunk c has been handed to chunk c+2 (data on the height grid is read by all three
    stages, so it must have left the slot)."""
    import torch
    from hn2016_falwa_b200 import engine, xarrayinterface as xi
    cls = _cls(kind)
    xlon, ylat, height_p, u, v, t = height_grid_inputs()
    shifts = (0, 5, 11, 17)
    fields = [tuple(np.roll(x, s, axis=-1) for x in (u, v, t)) for s in shifts]
    dims = ("time", "level", "latitude", "longitude")
    coords = dict(time=np.arange(len(shifts)), level=height_p, latitude=ylat, longitude=xlon)
    ds = xi.SimpleDataset({n: xi.SimpleDataArray(np.stack([f[i] for f in fields]), dims, coords, name=n)
                           for i, n in enumerate("uvt")})
    kw = dict(data_on_evenly_spaced_pseudoheight_grid=True)
    q = xi.QGDataset(ds, qgfield=cls, qgfield_kwargs=kw, chunk=1)
    q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    singles = []
    for f3 in fields:
        f = cls(xlon, ylat, height_p, *f3, **kw)
        f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
        singles.append(f)
    for name in ("qgpv", "uref", "lwa_baro", "adv_flux_f3", "lwa"):
        got = np.asarray(getattr(q, name).data)
        for i, f in enumerate(singles):
            want = np.asarray(getattr(f, name))
            np.testing.assert_allclose(got[i], want, rtol=1e-11, atol=1e-12 * np.abs(want).max())
    hs = engine.HostStream(singles[0]._plan, chunk=1, want_lwa3d=True, on_height_grid=True)
    out = hs.run(*(np.stack([f[i] for f in fields]) for i in range(3)), cycles=2)
    for i, f in enumerate(singles):
        want = np.asarray(f.lwa)
        np.testing.assert_allclose(out["lwa"][i].numpy(), want, rtol=1e-11, atol=1e-12 * np.abs(want).max())
    # northern-hemisphere-only results through the dataset
    xlon2, ylat2, plev2, u2, v2, t2 = inputs()
    coords2 = dict(time=np.arange(3), level=plev2, latitude=ylat2, longitude=xlon2)
    ds2 = xi.SimpleDataset({n: xi.SimpleDataArray(np.stack([np.roll(a, s, axis=-1) for s in (0, 3, 9)]), dims, coords2,
                                                  name=n) for n, a in zip("uvt", (u2, v2, t2))})
    q2 = xi.QGDataset(ds2, qgfield=cls, qgfield_kwargs=dict(northern_hemisphere_results_only=True), chunk=2)
    q2.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    f = cls(xlon2, ylat2, plev2, np.roll(u2, 9, axis=-1), np.roll(v2, 9, axis=-1), np.roll(t2, 9, axis=-1),
            northern_hemisphere_results_only=True)
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
    for name in ("qref", "lwa_baro", "lwa"):
        want = np.asarray(getattr(f, name))
        got = np.asarray(getattr(q2, name).data)[2]
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12 * np.abs(want).max())
