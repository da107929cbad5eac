"""The column walk fused into the window kernel (FALWA_FUSED_WALK=1, off by default: measured no faster than the two
kernels) must give the same results as the default two-kernel path."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from hn2016_falwa_b200.oopinterface import QGFieldNHN22
from hn2016_falwa_b200.synthetic import synthetic_snapshot
f = QGFieldNHN22(*synthetic_snapshot(nlon=96, nlat=49, seed=3))
f.interpolate_fields(False); f.compute_reference_states(False); f.compute_lwa_and_barotropic_fluxes(False)
np.savez(sys.argv[1], **{n: np.asarray(getattr(f, n)) for n in ("lwa", "lwa_baro", "u_baro", "adv_flux_f1", "adv_flux_f2",
         "adv_flux_f3", "convergence_zonal_advective_flux", "divergence_eddy_momentum_flux", "meridional_heat_flux",
         "flux_vector_phi_baro")})
""" % ROOT


def test_fused_column_walk_equals_two_kernels(tmp_path):
    outs = []
    for flag in ("0", "1"):   # the knob is read once per process
        path = str(tmp_path / f"walk{flag}.npz")
        env = dict(os.environ, FALWA_FUSED_WALK=flag)
        subprocess.run([sys.executable, "-c", SCRIPT, path], check=True, env=env)
        outs.append(np.load(path))
    for n in outs[0].files:
        a, b = outs[0][n], outs[1][n]
        np.testing.assert_allclose(b, a, rtol=1e-10, atol=1e-12 * np.abs(a).max(), err_msg=n)
