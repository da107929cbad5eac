"""CPU tests of the drop-in boundary: the shared library builds in-tree, loads, and exports every symbol that
include/falwa_b200.h declares; the Python layer mirrors the reference's interface (names, argument order) and refuses
to run without CUDA instead of falling back to the CPU."""
import inspect
import os

import pytest

from hn2016_falwa_b200 import _lib

F2PY_SIGNATURES = {   # SURVEY.md Appendix B (numpy.f2py -h of the reference's f90_modules)
    "compute_qgpv": ["ut", "vt", "theta", "height", "t0", "stat", "aa", "omega", "dz", "hh", "rr", "cp"],
    "compute_qgpv_direct_inv": ["jd", "uq", "vq", "tq", "height", "ts0", "tn0", "stats", "statn", "aa", "omega", "dz",
                                "hh", "rr", "cp"],
    "compute_reference_states": ["pv", "uu", "pt", "stat", "jd", "npart", "maxits", "a", "om", "dz", "eps", "h", "dphi",
                                 "dlambda", "r", "cp", "rjac"],
    "compute_qref_and_fawa_first": ["pv", "uu", "vort", "pt", "tn0", "nd", "nnd", "jb", "jd", "a", "omega", "dz", "h",
                                    "dphi", "dlambda", "rr", "cp"],
    "matrix_b4_inversion": ["k", "jmax", "jb", "z", "statn", "qref", "ckref", "sjk", "a", "om", "dz", "h", "rr", "cp"],
    "matrix_after_inversion": ["k", "qjj", "djj", "cjj", "rj", "sjk", "tjk"],
    "upward_sweep": ["jmax", "jb", "sjk", "tjk", "ckref", "tb", "qref_over_cor", "a", "om", "dz", "h", "rr", "cp"],
    "compute_flux_dirinv_nshem": ["pv", "uu", "vv", "pt", "ncforce", "tn0", "qref", "uref", "tref", "jb", "is_nhem",
                                  "a", "om", "dz", "h", "rr", "cp", "prefac"],
    "compute_lwa_only_nhn22": ["pv", "uu", "qref", "jb", "is_nhem", "a", "om", "dz", "h", "rr", "cp", "prefac"],
}


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from hn2016_falwa_b200 import build
        build.build()
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert b"sm_100a" in lib.falwa_b200_version()


def test_f2py_compatible_signatures():
    from hn2016_falwa_b200 import f90_modules
    for name, args in F2PY_SIGNATURES.items():
        params = list(inspect.signature(getattr(f90_modules, name)).parameters)
        assert params[:len(args)] == args, (name, params)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import numpy as np
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22
    from hn2016_falwa_b200.synthetic import synthetic_snapshot
    with pytest.raises(_lib.FalwaB200Error):
        QGFieldNHN22(*synthetic_snapshot(nlon=16, nlat=9))
    with pytest.raises(_lib.FalwaB200Error):
        from hn2016_falwa_b200 import f90_modules
        z = np.zeros((4, 5, 6))
        f90_modules.compute_qgpv(z, z, z, np.zeros(6), np.zeros(6), np.ones(6), 1., 1., 1., 1., 1., 1.)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(_lib.LIB_PATH)
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libfalwa_oracle" not in text, fn
