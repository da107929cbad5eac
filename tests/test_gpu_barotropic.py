"""GPU parity of the 2-D path (BASELINE.json config 5) against the NumPy oracle of basis.eqvlat_fawa / basis.lwa and
the reference's own inline answer key (tests/test_barotropic_field.py:26-54, rtol 1e-5)."""
import os

import numpy as np
import pytest

from oracle import basis2d, testdata
from tests.parity import compare

pytestmark = pytest.mark.gpu


def test_answer_key_and_oracle():
    from hn2016_falwa_b200.barotropic_field import BarotropicField
    g = np.load(os.path.join(testdata.GOLDEN_DIR, "barotropic_field.npz"))
    f = BarotropicField(g["analytic_xlon"], g["analytic_ylat"], g["analytic_pv"])
    key = basis2d.ANSWER_KEY
    assert np.allclose(f.equivalent_latitudes, key["eqv_lat"], rtol=1e-5, atol=1e-8)
    assert np.allclose(f.lwa.mean(axis=-1), key["lwa_zonal_mean"], rtol=1e-5, atol=1e-8)
    assert np.allclose(f.lwa[20, :], key["lwa_row20"], rtol=1e-5, atol=1e-8)
    for case in ("analytic", "seeded"):
        f = BarotropicField(g[f"{case}_xlon"], g[f"{case}_ylat"], g[f"{case}_pv"])
        compare(f"baro2d/{case}/eqvlat", f.equivalent_latitudes, g[f"{case}_eqvlat"], rtol=1e-8,
                atol=1e-9 * np.abs(g[f"{case}_eqvlat"]).max())
        compare(f"baro2d/{case}/lwa", f.lwa, g[f"{case}_lwa"], rtol=1e-8, atol=1e-9 * np.abs(g[f"{case}_lwa"]).max())


def test_batch_partitioned_and_zonally_symmetric():
    from hn2016_falwa_b200.barotropic_field import BarotropicBatch
    rng = np.random.default_rng(0)
    nlat, nlon, T = 61, 120, 3
    ylat = np.linspace(-90, 90, nlat)
    xlon = np.linspace(0, 360, nlon, endpoint=False)
    base = 2 * 7.29e-5 * np.sin(np.deg2rad(ylat))[:, None]
    pv = np.stack([base + 8e-5 * np.cos(np.deg2rad(ylat))[:, None] * np.exp(-((ylat[:, None] - 36) / 10) ** 2) *
                   np.cos(3 * np.deg2rad(xlon)[None, :] + k) + 1e-6 * rng.standard_normal((nlat, nlon))
                   for k in range(T)])
    b = BarotropicBatch(xlon, ylat, pv, return_partitioned_lwa=True).compute()
    lwa = b.lwa.cpu().numpy()
    for k in range(T):
        eq, want = basis2d.barotropic_eqvlat_lwa(xlon, ylat, pv[k])
        compare(f"baro2d/batch{k}/eqvlat", b.equivalent_latitudes[k].cpu().numpy(), eq, rtol=1e-8, atol=1e-12)
        compare(f"baro2d/batch{k}/lwa", lwa[k].sum(axis=0), want, rtol=1e-8, atol=1e-9 * np.abs(want).max())
        assert (lwa[k] >= 0).all()
    # zonally symmetric field: every column identical, same values as the oracle (the equivalent-latitude Q(y) is not
    # exactly the input profile, so LWA is small but not identically zero)
    sym = np.broadcast_to(base, (nlat, nlon)).copy()
    z = BarotropicBatch(xlon, ylat, sym[None]).compute().lwa.cpu().numpy()[0]
    _, want = basis2d.barotropic_eqvlat_lwa(xlon, ylat, sym)
    compare("baro2d/symmetric/lwa", z, want, rtol=1e-8, atol=1e-9 * np.abs(want).max())
    assert np.array_equal(z, np.broadcast_to(z[:, :1], z.shape))


def test_reference_barotropic_vorticity_file():
    """GPU mirror of the reference's tests/test_integration.py:107-125 on its own sampled field
    (tests/data/barotropic_vorticity.nc, 256 x 512): total == sum of the cyclonic / anticyclonic parts, equal
    equivalent latitudes — and both against what the reference's unmodified BarotropicField returns for that file
    (tests/golden/barotropic_vorticity_256x512.npz, written by oracle.testdata.convert_reference_barotropic)."""
    from hn2016_falwa_b200.barotropic_field import BarotropicField
    g = testdata.load_reference_barotropic()
    av = g["absolute_vorticity"]
    xlon = np.linspace(0, 360., 512, endpoint=False)
    ylat = np.linspace(-90, 90., 256, endpoint=True)
    cc1 = BarotropicField(xlon, ylat, pv_field=av)
    cc2 = BarotropicField(xlon, ylat, pv_field=av, return_partitioned_lwa=True)
    assert np.isclose(cc1.equivalent_latitudes, cc2.equivalent_latitudes).all()
    assert np.isclose(cc1.lwa, cc2.lwa.sum(axis=0)).all()
    assert cc2.lwa.shape == (2, 256, 512) and (cc2.lwa >= 0).all()
    compare("baro2d/reference_file/eqvlat", cc1.equivalent_latitudes, g["equivalent_latitudes"], rtol=1e-8,
            atol=1e-9 * np.abs(g["equivalent_latitudes"]).max())
    compare("baro2d/reference_file/lwa", cc1.lwa, g["lwa"], rtol=1e-8, atol=1e-9 * np.abs(g["lwa"]).max())
    # the reference's own test hands the float32 values over as they are: same answer at its tolerance
    assert np.allclose(cc1.equivalent_latitudes, g["equivalent_latitudes_f32_input"], rtol=1e-5, atol=1e-9)
