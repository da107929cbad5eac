"""CPU mirror of the reference's tests/test_xarrayinterface.py:24-106 (argument handling, flips, leading dimensions)
for hn2016_falwa_b200.xarrayinterface.QGDataset, with the chunk runner backed by the CPU oracle (test infrastructure)
and the Simple* stand-ins for xarray objects."""
import numpy as np
import pytest

from hn2016_falwa_b200 import xarrayinterface as xi
from hn2016_falwa_b200.synthetic import synthetic_snapshot
from tests.test_dataset_sharding import OracleRunner, _Template

XLON, YLAT, PLEV, U, V, T = synthetic_snapshot(nlon=24, nlat=13, seed=7)
DIMS = ("plev", "ylat", "xlon")


def _arrays(lead=(), lead_coords=None):
    """u, v, t DataArrays with optional leading dimensions (name, size) filled with shifted copies of the snapshot."""
    names = tuple(n for n, _ in lead)
    shape = tuple(s for _, s in lead)
    coords = dict(plev=PLEV, ylat=YLAT, xlon=XLON, **{n: np.arange(s) for n, s in lead})
    out = []
    for name, a in (("u", U), ("v", V), ("t", T)):
        data = np.empty(shape + a.shape)
        for idx in np.ndindex(*shape):
            data[idx] = np.roll(a, sum(idx), axis=-1)
        out.append(xi.SimpleDataArray(data, names + DIMS, coords, name=name))
    return out


def _ds(*arrays):
    return xi.SimpleDataset({a.name: a for a in arrays})


def _q(*args, **kw):
    return xi.QGDataset(*args, qgfield=_Template, runner=OracleRunner(XLON, YLAT, PLEV), **kw)


def test_with_dataset_dataarrays_and_mixed_arguments():
    u, v, t = _arrays()
    q = _q(_ds(u, v, t))                      # a dataset without leading dimensions is one snapshot
    interp = q.interpolate_fields()
    assert interp["interpolated_u"].dims == ("height", "ylat", "xlon") and interp["qgpv"].shape == (13, 13, 24)
    q.compute_reference_states(); q.compute_lwa_and_barotropic_fluxes()
    assert q.lwa_baro.shape == (13, 24)
    a = _q(u, v, t).interpolate_fields()      # three data arrays
    b = _q(_ds(u, t), da_v=v).interpolate_fields()   # dataset + the missing array
    for x in (a, b):
        np.testing.assert_array_equal(x["qgpv"].data, interp["qgpv"].data)


def test_rejects_incomplete_and_mismatching_inputs():
    u, v, t = _arrays()
    for drop in "uvt":
        with pytest.raises(KeyError):
            _q(_ds(*[a for a in (u, v, t) if a.name != drop]))
    renamed = xi.SimpleDataArray(t.data, ("plev", "latitude", "xlon"),
                                 dict(plev=PLEV, latitude=YLAT, xlon=XLON), name="t")
    with pytest.raises(AssertionError):       # dimension names differ (reference: test_qgdataset_with_coordinate_mismatch)
        _q(u, v, renamed)
    shifted = xi.SimpleDataArray(t.data, DIMS, dict(plev=PLEV, ylat=YLAT + 0.5, xlon=XLON), name="t")
    with pytest.raises((AssertionError, ValueError)):   # same names, different coordinate values (xr.merge join="exact")
        _q(u, v, shifted)
    for order in (("xlon", "plev", "ylat"), ("ylat", "xlon", "plev")):
        perm = [DIMS.index(d) for d in order]
        tr = [xi.SimpleDataArray(np.transpose(a.data, perm), order, a.coords, name=a.name) for a in (u, v, t)]
        with pytest.raises(AssertionError):
            _q(_ds(*tr))


def test_flipped_latitude_and_pressure_give_the_same_fields():
    u, v, t = _arrays()
    ref = _q(_ds(u, v, t)).interpolate_fields()
    flip_y = [xi.SimpleDataArray(a.data[:, ::-1, :], DIMS, dict(plev=PLEV, ylat=YLAT[::-1], xlon=XLON), name=a.name)
              for a in (u, v, t)]
    flip_p = [xi.SimpleDataArray(a.data[::-1], DIMS, dict(plev=PLEV[::-1], ylat=YLAT, xlon=XLON), name=a.name)
              for a in (u, v, t)]
    for arrays in (flip_y, flip_p):
        got = _q(_ds(*arrays)).interpolate_fields()
        assert np.allclose(got["qgpv"].coords["ylat"], ref["qgpv"].coords["ylat"])
        assert np.allclose(got["qgpv"].coords["height"], ref["qgpv"].coords["height"])
        np.testing.assert_array_equal(got["interpolated_theta"].data, ref["interpolated_theta"].data)


def test_leading_dimensions_are_preserved_4d_5d():
    q = _q(_ds(*_arrays(lead=(("time", 3),))))
    interp = q.interpolate_fields()
    assert interp["interpolated_u"].dims == ("time", "height", "ylat", "xlon")
    assert interp["interpolated_u"].shape == (3, q.attrs["kmax"], 13, 24) and "time" in interp["interpolated_u"].coords
    q = _q(_ds(*_arrays(lead=(("number", 4), ("time", 3)))))
    interp = q.interpolate_fields()
    assert interp["interpolated_u"].dims == ("number", "time", "height", "ylat", "xlon")
    assert interp["interpolated_u"].shape == (4, 3, q.attrs["kmax"], 13, 24)
    single = _q(_ds(*_arrays())).interpolate_fields()["interpolated_u"].data
    np.testing.assert_array_equal(interp["interpolated_u"].data[2, 1], np.roll(single, 3, axis=-1))
