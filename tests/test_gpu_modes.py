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
