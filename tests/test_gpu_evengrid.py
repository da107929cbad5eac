"""Inputs on an even latitude grid without the equator: the constructor regrids them to an odd grid and every output is
interpolated back (reference oopinterface.py:147-154, :345-403; tests/test_oopinterface.py:230-287).  Compared with a
fixture produced by the reference's own unmodified Python (oracle/gen_golden_evengrid.py)."""
import os

import numpy as np
import pytest

from oracle import testdata
from oracle.gen_golden_evengrid import NAMES, even_grid_inputs
from tests.parity import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_even_latitude_grid(kind):
    from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
    g = np.load(os.path.join(testdata.GOLDEN_DIR, "evengrid_synthetic.npz"))
    inputs = even_grid_inputs()
    f = (QGFieldNH18 if kind == "nh18" else QGFieldNHN22)(*inputs)
    assert f.need_latitude_interpolation and f.get_latitude_dim() == 48
    f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
    for n in NAMES:
        got = np.asarray(getattr(f, n))
        got = got[::6] if got.ndim == 3 else got
        want = g[f"{kind}_{n}"]
        assert got.shape == want.shape, (n, got.shape, want.shape)
        compare(f"evengrid/{kind}/{n}", got, want, rtol=1e-8, atol=1e-9 * np.abs(want).max())
