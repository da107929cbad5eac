"""CPU tests of the ingest restatement (oracle/ingest.py) and of QGDataset's handling of undecoded inputs.

The pin: the reference's tests decode the packed ERA-Interim snapshot with xarray and hand it over flipped
(tests/test_output_results.py:26-45, :94-97); ``oracle.testdata.load_era_snapshot`` is that array, and every golden
fixture of the 3-D path was generated from it.  Decoding the raw int16 values of the same file must give it back
bit for bit."""
import numpy as np

from oracle import ingest, testdata
from hn2016_falwa_b200 import xarrayinterface as xi


def test_decode_of_packed_snapshot_is_the_golden_input():
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    lon, lat, lev, raw, packing = testdata.load_era_packed()
    assert lat[0] > lat[-1] and lev[0] < lev[-1]          # file order: north -> south, top -> bottom
    for name, want in zip("uvt", (u, v, t)):
        got = ingest.decode_field(raw[name], *packing[name], flip_lat=True, flip_lev=True)
        assert got.dtype == np.float64 and np.array_equal(got, want)
    assert np.array_equal(lat[::-1], ylat) and np.array_equal(lev[::-1], plev)


def test_decode_fill_value_and_plain_arrays():
    rng = np.random.default_rng(3)
    raw = rng.integers(-32768, 32767, size=(2, 3, 4, 5), dtype=np.int16)
    raw[0, 1, 2, 3] = -32767
    got = ingest.decode_field(raw, 0.01, 250.0, fill_value=-32767)
    assert np.isnan(got[0, 1, 2, 3]) and np.isnan(got).sum() == (raw == -32767).sum()
    x = rng.standard_normal((2, 3, 4, 5))
    assert np.array_equal(ingest.decode_field(x, flip_lev=True), x[:, ::-1])
    assert np.array_equal(ingest.decode_field(x.astype(np.float32), flip_lat=True), x.astype(np.float32)[:, :, ::-1])


def test_host_decoder_of_the_package_equals_the_oracle():
    lon, lat, lev, raw, packing = testdata.load_era_packed()
    for name in "uvt":
        lazy = xi._LazyField(raw[name][None], packing[name], True, True)
        assert lazy.shape == (1, 37, 121, 240) and len(lazy) == 1
        want = ingest.decode_field(raw[name], *packing[name], flip_lat=True, flip_lev=True)
        assert np.array_equal(lazy[0], want)
        assert np.array_equal(np.asarray(lazy[0:1])[0], want)


def test_packing_attributes_are_picked_up_only_when_undecoded():
    a = xi.SimpleDataArray(np.zeros((1, 2, 3, 4), np.int16), ("time", "level", "latitude", "longitude"),
                           attrs=dict(scale_factor=0.5, add_offset=3.0, _FillValue=-32767))
    assert xi._packing_of(a) == (0.5, 3.0, -32767.0)
    b = xi.SimpleDataArray(np.zeros((1, 2, 3, 4)), ("time", "level", "latitude", "longitude"), attrs=dict(units="K"))
    assert xi._packing_of(b) is None


def test_dataset_on_packed_inputs_equals_dataset_on_decoded_inputs():
    """QGDataset keeps packed inputs as they are and decodes lazily; a (CPU) runner sees the same float64 snapshots as
    with decoded inputs."""
    lon, lat, lev, raw, packing = testdata.load_era_packed()
    dims = ("time", "level", "latitude", "longitude")
    coords = dict(time=np.arange(2), level=lev, latitude=lat, longitude=lon)
    mk = lambda n: xi.SimpleDataArray(np.stack([raw[n], raw[n]]), dims, coords, name=n,
                                      attrs=dict(scale_factor=packing[n][0], add_offset=packing[n][1]))
    seen = []

    class Recorder:
        def run(self, u, v, t, stages, want, ncforce=None):
            for i in range(u.shape[0]):
                seen.append((u[i], v[i], t[i]))
            return iter(())

    from tests.test_dataset_sharding import _Template   # stands in for a QGField (no GPU here)
    q = xi.QGDataset(xi.SimpleDataset(dict(u=mk("u"), v=mk("v"), t=mk("t"))), qgfield=_Template, runner=Recorder(),
                     shard=(0, 1))
    assert q._u.raw.dtype == np.int16 and q._u.flip_lat and q._u.flip_lev
    q._run("interp")
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    assert len(seen) == 2
    for got, want in zip(seen[1], (u, v, t)):
        assert np.array_equal(got, want)
    assert np.array_equal(q._ylat, ylat) and np.array_equal(q._plev, plev)
