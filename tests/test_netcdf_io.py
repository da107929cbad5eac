"""CPU tests of hn2016_falwa_b200.netcdf_io: undecoded NetCDF-3 input (big-endian packed int16 + scale/offset, file
order) reaches QGDataset's runner as the golden decoded snapshot, and result files round-trip."""
import numpy as np
from scipy.io import netcdf_file

from oracle import testdata
from hn2016_falwa_b200 import netcdf_io, xarrayinterface as xi


def write_packed_era_file(path, names="uvt", nt=2):
    """The reference's packed ERA-Interim snapshot (tests/data/2005-01-23-0000-{u,v,t}.nc: int16 + scale_factor /
    add_offset, north -> south, top -> bottom) re-written as a NetCDF-3 file with ``nt`` identical time steps."""
    lon, lat, lev, raw, packing = testdata.load_era_packed()
    with netcdf_file(path, "w", version=2) as f:
        for d, c in (("time", np.arange(nt, dtype=np.int32)), ("level", lev.astype(np.int32)),
                     ("latitude", lat.astype(np.float32)), ("longitude", lon.astype(np.float32))):
            f.createDimension(d, c.size)
            v = f.createVariable(d, c.dtype, (d,))
            v[:] = c
        for n in names:
            v = f.createVariable(n, np.dtype(np.int16), ("time", "level", "latitude", "longitude"))
            v[:] = np.stack([raw[n]] * nt)
            v.scale_factor, v.add_offset = (np.float64(x) for x in packing[n])   # (plain floats become float32)
            v._FillValue = np.int16(-32767)
    return path


def test_open_fields_keeps_storage_type_and_dataset_decodes_lazily(tmp_path):
    from tests.test_dataset_sharding import _Template
    paths = [write_packed_era_file(str(tmp_path / f"{n}.nc"), names=n) for n in "uvt"]   # one file per variable
    ds = netcdf_io.open_fields(*paths)
    assert list(ds) == ["u", "v", "t"] and ds["u"].data.dtype == np.dtype(">i2") and ds["u"].shape == (2, 37, 121, 240)
    _, _, _, raw, packing = testdata.load_era_packed()
    assert ds["t"].attrs["scale_factor"] == packing["t"][0] and ds["t"].attrs["add_offset"] == packing["t"][1]
    seen = []

    class Recorder:
        def run(self, u, v, t, stages, want, ncforce=None):
            assert u.byteswap and u.raw_native.dtype == np.int16 and u.flip_lat and u.flip_lev
            assert np.array_equal(u.raw_native.byteswap(), raw["u"][None].repeat(2, 0))
            seen.extend((u[i], v[i], t[i]) for i in range(u.shape[0]))
            return iter(())

    q = xi.QGDataset(ds, qgfield=_Template, runner=Recorder(), shard=(0, 1))
    q._run("interp")
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    assert len(seen) == 2
    for got, want in zip(seen[1], (u, v, t)):
        assert np.array_equal(got, want)


def test_write_results_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    coords = dict(time=np.array(["2005-01-23T00", "2005-01-23T06"], dtype="datetime64[h]"), height=np.arange(5) * 1e3,
                  ylat=np.linspace(-90, 90, 7), xlon=np.arange(8) * 45.)
    lwa = xi.SimpleDataArray(rng.standard_normal((2, 5, 7, 8)), ("time", "height", "ylat", "xlon"), coords, name="lwa")
    baro = xi.SimpleDataArray(rng.standard_normal((2, 7, 8)), ("time", "ylat", "xlon"),
                              {k: coords[k] for k in ("time", "ylat", "xlon")}, name="lwa_baro")
    ds = xi.SimpleDataset(dict(lwa=lwa, lwa_baro=baro), attrs=dict(kmax=5, protocol="QGFieldNHN22", dz=1000.))
    path = netcdf_io.write_results(str(tmp_path / "out.nc"), ds, dtype=np.float64)
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.variables["lwa"].dimensions == ("time", "height", "ylat", "xlon")
        assert np.array_equal(f.variables["lwa"][:], lwa.data) and np.array_equal(f.variables["lwa_baro"][:], baro.data)
        assert f.variables["lwa"].units == b"m/s" and f.protocol == b"QGFieldNHN22" and f.kmax == 5
        assert np.array_equal(f.variables["ylat"][:], coords["ylat"])
        assert f.variables["time"][1] - f.variables["time"][0] == 6 * 3600
    path32 = netcdf_io.write_results(str(tmp_path / "out32.nc"), ds, names=["lwa_baro"])
    with netcdf_file(path32, "r", mmap=False) as f:
        assert f.variables["lwa_baro"][:].dtype == np.dtype(">f4") and "lwa" not in f.variables
