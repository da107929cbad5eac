"""Input generators of the bench and the full-size tests (hn2016_falwa_b200/synthetic.py), on the CPU."""
import os
import sys

import numpy as np

from hn2016_falwa_b200.synthetic import load_packed_snapshot, synthetic_snapshot, upsampled_snapshot
from oracle import testdata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_snapshot_loader_matches_the_oracle_side_loader():
    """The product-side decoder of the committed real snapshot gives, bit for bit, what the test infrastructure's
    loader (pinned against the reference's own files, tests/test_ingest_oracle.py) gives."""
    got = load_packed_snapshot(testdata.ERA_NPZ)
    want = testdata.load_era_snapshot()
    assert len(got) == len(want) == 6
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    assert got[1][0] < got[1][-1] and got[2][0] > got[2][-1]      # latitude ascending, pressure descending


def test_upsampling_is_bilinear_periodic_and_keeps_the_coarse_points():
    xlon, ylat, plev, u, v, t = testdata.load_era_snapshot()
    x2, y2, u2, v2, t2 = upsampled_snapshot(xlon, ylat, u, v, t, nlon=480, nlat=241)
    assert u2.shape == (37, 241, 480) and x2[0] == 0.0 and y2[0] == -90.0 and y2[-1] == 90.0
    for fine, coarse in ((u2, u), (v2, v), (t2, t)):
        np.testing.assert_array_equal(fine[:, ::2, ::2], coarse)                     # coinciding points are copied
        mid = 0.5 * (coarse + np.roll(coarse, -1, axis=2))                            # halfway in longitude, wrapping
        np.testing.assert_allclose(fine[:, ::2, 1::2], mid, rtol=1e-14, atol=1e-12)
        midy = 0.5 * (coarse[:, :-1, :] + coarse[:, 1:, :])
        np.testing.assert_allclose(fine[:, 1::2, ::2], midy, rtol=1e-14, atol=1e-12)
    assert np.isfinite(t2).all() and t2.min() > 150.0


def test_bench_batches_of_both_data_kinds():
    sys.path.insert(0, ROOT)
    import bench
    for data in ("analytic", "upsampled"):
        xlon, ylat, plev, u, v, t, shifts = bench.make_batch(240, 121, 3, seed=1234, data=data)
        assert u.shape == v.shape == t.shape == (37, 121, 240) and len(shifts) == 3 and shifts[0] == 0
        assert bench.workload_config(240, 121, data)["grid"] == {"nlon": 240, "nlat": 121, "nlev": 37, "kmax": 49}
    assert bench.workload_config(240, 121, "upsampled") != bench.workload_config(240, 121)
    a = synthetic_snapshot(nlon=48, nlat=25, seed=7)
    b = synthetic_snapshot(nlon=48, nlat=25, seed=7)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)                                          # seeded: reproducible
