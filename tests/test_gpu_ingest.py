"""GPU parity of the ingest kernel (Section D of include/falwa_b200.h) with oracle/ingest.py — bit-exact — and of the
packed-input path of HostStream / QGDataset with the decoded-input path."""
import numpy as np
import pytest
import torch

from oracle import ingest as oracle_ingest, testdata

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


@pytest.mark.parametrize("dtype", [np.int16, np.float32, np.float64])
@pytest.mark.parametrize("flip_lat,flip_lev", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(2, 5, 7, 12), (1, 3, 6, 9), (3, 37, 13, 24)])
def test_ingest_kernel_bit_exact(dtype, flip_lat, flip_lev, shape):
    from hn2016_falwa_b200 import engine
    rng = np.random.default_rng(hash((str(dtype), flip_lat, flip_lev, shape)) % (2 ** 32))
    if dtype == np.int16:
        raw = rng.integers(-32768, 32767, size=shape, dtype=np.int16)
        packing = (0.0021312222851791987, 249.9451242692285)
    else:
        raw = (rng.standard_normal(shape) * 30).astype(dtype)
        packing = None if dtype == np.float64 else (1.0, 0.0)
    want = oracle_ingest.decode_field(raw, *(packing or ()), flip_lat=flip_lat, flip_lev=flip_lev)
    got = engine.ingest_field(torch.from_numpy(raw).cuda(), packing, flip_lat, flip_lev).cpu().numpy()
    assert got.dtype == np.float64 and np.array_equal(_bits(got), _bits(want))


def test_ingest_fill_value_unaligned_and_errors():
    from hn2016_falwa_b200 import engine
    rng = np.random.default_rng(5)
    raw = rng.integers(-3000, 3000, size=(2, 4, 5, 8), dtype=np.int16)
    raw[1, 2, 3, 4] = raw[0, 0, 0, 0] = -32767
    want = oracle_ingest.decode_field(raw, 0.5, -3.0, fill_value=-32767, flip_lat=True)
    got = engine.ingest_field(torch.from_numpy(raw).cuda(), (0.5, -3.0, -32767), True, False).cpu().numpy()
    assert np.isnan(got).sum() == 2 and np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(want))
    # a source that starts at an odd element (2-byte aligned only) takes the scalar path
    base = torch.from_numpy(rng.integers(-100, 100, size=1 + raw.size, dtype=np.int16)).cuda()
    view = base[1:].view(raw.shape)
    got = engine.ingest_field(view, (2.0, 1.0), False, True).cpu().numpy()
    assert np.array_equal(got, oracle_ingest.decode_field(view.cpu().numpy(), 2.0, 1.0, flip_lev=True))
    with pytest.raises(AssertionError):
        engine.ingest_field(torch.zeros(2, 3, 4, dtype=torch.int16, device="cuda"))
    with pytest.raises(AssertionError):
        engine.ingest_field(torch.zeros(1, 2, 3, 4, dtype=torch.int32, device="cuda"))


def test_packed_snapshot_decodes_to_the_golden_input():
    from hn2016_falwa_b200 import engine
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    _, _, _, raw, packing = testdata.load_era_packed()
    for name, want in zip("uvt", (u, v, t)):
        got = engine.ingest_field(torch.from_numpy(raw[name][None]).cuda(), packing[name], True, True)
        assert np.array_equal(_bits(got[0].cpu().numpy()), _bits(want))


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_dataset_on_packed_file_order_inputs(kind):
    """QGDataset fed with the undecoded int16 variables in file order (north -> south, top -> bottom) gives what the
    golden fixture holds for the decoded snapshot, and the same as the decoded float64 route."""
    from hn2016_falwa_b200 import xarrayinterface as xi
    from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
    import os
    cls = QGFieldNH18 if kind == "nh18" else QGFieldNHN22
    lon, lat, lev, raw, packing = testdata.load_era_packed()
    dims = ("time", "level", "latitude", "longitude")
    coords = dict(time=np.arange(3), level=lev, latitude=lat, longitude=lon)
    packed = xi.SimpleDataset({n: xi.SimpleDataArray(
        np.stack([raw[n]] * 3), dims, coords, name=n,
        attrs=dict(scale_factor=packing[n][0], add_offset=packing[n][1], _FillValue=-32767)) for n in "uvt"})
    q = xi.QGDataset(packed, qgfield=cls, chunk=2)
    assert q._u.raw.dtype == np.int16
    q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    coords2 = dict(time=np.arange(1), level=plev, latitude=ylat, longitude=xlon)
    decoded = xi.SimpleDataset({n: xi.SimpleDataArray(a[None], dims, coords2, name=n)
                                for n, a in zip("uvt", (u, v, t))})
    q2 = xi.QGDataset(decoded, qgfield=cls)
    q2.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    gold = np.load(os.path.join(testdata.GOLDEN_DIR, f"qgfield_{kind}_era_interim.npz"))
    stride, levels = int(gold["lon_stride"]), list(gold["levels_3d"])
    for name in ("qgpv", "interpolated_theta", "qref", "uref", "lwa_baro", "adv_flux_f1", "lwa"):
        a, b = np.asarray(getattr(q, name).data), np.asarray(getattr(q2, name).data)
        scale = np.abs(b).max()
        for i in range(3):
            np.testing.assert_allclose(a[i], b[0], rtol=1e-11, atol=1e-12 * scale)
        g = gold[name]      # subsampled like tests/test_gpu_pipeline.py::test_golden_fixture_era
        x = a[2]
        x = x[levels][:, :, ::stride] if x.ndim == 3 else (x[:, ::stride] if x.shape[-1] > 200 else x)
        np.testing.assert_allclose(x, g, rtol=1e-8, atol=1e-9 * np.abs(g).max())


def test_hoststream_packed_int16_equals_float64_stream():
    from hn2016_falwa_b200 import engine
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    _, _, _, raw, packing = testdata.load_era_packed()
    plan = QGFieldNHN22(xlon, ylat, plev, u, v, t)._plan
    hs = engine.HostStream(plan, chunk=2, want_lwa3d=True)
    o64 = hs.run(*(np.stack([x] * 3) for x in (u, v, t)))
    hp = engine.HostStream(plan, chunk=2, want_lwa3d=True, input_dtype=torch.int16,
                           packing=[packing[n] for n in "uvt"], flip_lat=True, flip_lev=True)
    o16 = hp.run(*(np.stack([raw[n]] * 3) for n in "uvt"))
    assert hp.h2d_bytes * 4 == hs.h2d_bytes
    for name in ("qref", "uref", "out2d", "lwa"):
        a, b = o16[name].numpy(), o64[name].numpy()
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-12 * np.abs(b).max())
    with pytest.raises(TypeError):
        hp.run(*(np.stack([x] * 3) for x in (u, v, t)))


@pytest.mark.parametrize("dtype", [">i2", ">f4", ">f8"])
@pytest.mark.parametrize("shape", [(2, 3, 6, 10), (1, 4, 5, 7)])
def test_ingest_big_endian_storage(dtype, shape):
    """NetCDF-3 files store big-endian values: the bytes are shipped as they are and swapped on the device."""
    from hn2016_falwa_b200 import engine
    rng = np.random.default_rng(len(dtype) + shape[-1])
    native = (rng.integers(-32768, 32767, size=shape).astype(np.int16) if dtype == ">i2"
              else (rng.standard_normal(shape) * 50).astype(dtype[1:].replace("f4", "float32").replace("f8", "float64")))
    stored = native.astype(dtype)                                 # big-endian bytes
    packing = (0.003455173064643958, 28.886333936905178) if dtype == ">i2" else None
    want = oracle_ingest.decode_field(stored, *(packing or ()), flip_lat=True, flip_lev=True)
    as_native = torch.from_numpy(stored.view(stored.dtype.newbyteorder("="))).cuda()
    got = engine.ingest_field(as_native, packing, True, True, byteswap=True).cpu().numpy()
    assert np.array_equal(_bits(got), _bits(want))


def test_netcdf3_file_through_dataset_and_back(tmp_path):
    """Packed big-endian NetCDF-3 input -> QGDataset (device decode) -> golden results -> NetCDF-3 output file."""
    import os
    from scipy.io import netcdf_file
    from hn2016_falwa_b200 import netcdf_io, xarrayinterface as xi
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22
    from tests.test_netcdf_io import write_packed_era_file
    ds = netcdf_io.open_fields(write_packed_era_file(str(tmp_path / "era.nc"), nt=3))
    q = xi.QGDataset(ds, qgfield=QGFieldNHN22, chunk=2)
    assert q._u.byteswap and q._u.raw.dtype == np.dtype(">i2")
    out = q.compute_lwa_and_barotropic_fluxes()
    gold = np.load(os.path.join(testdata.GOLDEN_DIR, "qgfield_nhn22_era_interim.npz"))
    stride = int(gold["lon_stride"])
    for name in ("lwa_baro", "u_baro", "adv_flux_f1", "meridional_heat_flux"):
        g = gold[name]
        for i in range(3):
            x = np.asarray(out[name].data)[i]
            x = x[:, ::stride] if x.shape[-1] > 200 else x
            np.testing.assert_allclose(x, g, rtol=1e-8, atol=1e-9 * np.abs(g).max())
    path = netcdf_io.write_results(str(tmp_path / "out.nc"), out, names=["lwa_baro", "u_baro", "lwa"])
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.variables["lwa"].dimensions == ("time", "height", "ylat", "xlon")
        np.testing.assert_array_equal(f.variables["lwa_baro"][:], np.asarray(out["lwa_baro"].data, dtype=np.float32))
        assert f.protocol == b"QGFieldNHN22"


def test_example_workflow_script(tmp_path):
    """examples/lwa_budget_nhn22.py: three packed NetCDF-3 files in, one result file out."""
    import importlib.util
    import os
    from scipy.io import netcdf_file
    from tests.test_netcdf_io import write_packed_era_file
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("lwa_budget_nhn22", os.path.join(root, "examples", "lwa_budget_nhn22.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    paths = [write_packed_era_file(str(tmp_path / f"{n}.nc"), names=n, nt=3) for n in "uvt"]
    out = str(tmp_path / "budget.nc")
    mod.main(*paths, out)
    with netcdf_file(out, "r", mmap=False) as f:
        assert f.variables["lwa_baro"][:].shape == (3, 121, 240) and f.variables["qref"][:].shape == (3, 49, 121)
        assert np.isfinite(f.variables["lwa_baro"][:]).all() and f.variables["lwa_baro"][:].max() > 1.0
