"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Parity status: see oracle/falwa_oracle.cpp header.

Python callables with the *f2py* signatures of the nine native routines of falwa v2.3.3
(generated from /root/reference/src/falwa/f90_modules/*.f90 with
``numpy.f2py -h``), backed by the C++ restatement in ``oracle/falwa_oracle.cpp``.

``install_as_falwa_modules()`` registers them as ``falwa.compute_qgpv`` … so that the
reference's own, unmodified ``falwa/oopinterface.py`` can be imported and driven by the oracle
(only possible where /root/reference exists; used by ``oracle/gen_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes
import os
import subprocess
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfalwa_oracle.so")
_lib = None


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, no fast-math, no FMA contraction)."""
    src = os.path.join(_HERE, "falwa_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


class _Mode:
    """dtype switch: float64 = parity oracle, float32 = arithmetic as shipped by the reference."""
    dtype = np.float64


def set_precision(dtype):
    _Mode.dtype = np.dtype(dtype).type


def _suf():
    return "f64" if _Mode.dtype == np.float64 else "f32"


def _creal(x):
    return ctypes.c_double(x) if _Mode.dtype == np.float64 else ctypes.c_float(x)


def _farr(x, ndim=None):
    a = np.asfortranarray(x, dtype=_Mode.dtype)
    if ndim is not None and a.ndim != ndim:
        raise ValueError(f"expected a {ndim}-d array, got shape {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _out(shape, dtype=None):
    return np.zeros(shape, dtype=dtype or _Mode.dtype, order="F")


def _fn(name):
    f = getattr(lib(), f"oracle_{name}_{_suf()}")
    f.restype = None
    return f


# --------------------------------------------------------------------------------------------
def compute_qgpv(ut, vt, theta, height, t0, stat, aa, omega, dz, hh, rr, cp):
    ut, vt, theta = _farr(ut, 3), _farr(vt, 3), _farr(theta, 3)
    nlon, nlat, kmax = ut.shape
    height, t0, stat = _farr(height, 1), _farr(t0, 1), _farr(stat, 1)
    pv, avort = _out(ut.shape), _out(ut.shape)
    _fn("compute_qgpv")(nlon, nlat, kmax, _p(ut), _p(vt), _p(theta), _p(height), _p(t0), _p(stat),
                        _creal(aa), _creal(omega), _creal(dz), _creal(hh), _creal(rr), _creal(cp), _p(pv), _p(avort))
    return pv, avort


def compute_qgpv_direct_inv(jd, uq, vq, tq, height, ts0, tn0, stats, statn, aa, omega, dz, hh, rr, cp):
    uq, vq, tq = _farr(uq, 3), _farr(vq, 3), _farr(tq, 3)
    nlon, nlat, kmax = uq.shape
    height, ts0, tn0, stats, statn = (_farr(x, 1) for x in (height, ts0, tn0, stats, statn))
    pv, avort = _out(uq.shape), _out(uq.shape)
    _fn("compute_qgpv_direct_inv")(nlon, nlat, kmax, int(jd), _p(uq), _p(vq), _p(tq), _p(height), _p(ts0), _p(tn0),
                                   _p(stats), _p(statn), _creal(aa), _creal(omega), _creal(dz), _creal(hh),
                                   _creal(rr), _creal(cp), _p(pv), _p(avort))
    return pv, avort


def compute_reference_states(pv, uu, pt, stat, jd, npart, maxits, a, om, dz, eps, h, dphi, dlambda, r, cp, rjac,
                             return_bins=False):
    pv, uu, pt, stat = _farr(pv, 3), _farr(uu, 3), _farr(pt, 3), _farr(stat, 1)
    nlon, nlat, kmax = pv.shape
    qref, u, tref = _out((jd, kmax)), _out((jd, kmax)), _out((jd, kmax))
    nit = ctypes.c_int(0)
    bins = _out(pv.shape, np.int32) if return_bins else None
    _fn("compute_reference_states")(_p(pv), _p(uu), _p(pt), _p(stat), nlon, nlat, kmax, int(jd), int(npart),
                                    int(maxits), _creal(a), _creal(om), _creal(dz), _creal(eps), _creal(h),
                                    _creal(dphi), _creal(dlambda), _creal(r), _creal(cp), _creal(rjac), _p(qref),
                                    _p(u), _p(tref), ctypes.byref(nit), _p(bins) if return_bins else None)
    if return_bins:
        return qref, u, tref, nit.value, bins
    return qref, u, tref, nit.value


def compute_qref_and_fawa_first(pv, uu, vort, pt, tn0, nd, nnd, jb, jd, a, omega, dz, h, dphi, dlambda, rr, cp,
                                return_bins=False):
    pv, uu, vort, pt, tn0 = _farr(pv, 3), _farr(uu, 3), _farr(vort, 3), _farr(pt, 3), _farr(tn0, 1)
    imax, jmax, kmax = pv.shape
    n = jd - 2
    qref, ubar, tbar, fawa, ckref = (_out((nd, kmax)) for _ in range(5))
    tjk, sjk = _out((n, kmax - 1)), _out((n, n, kmax - 1))
    bpv = _out(pv.shape, np.int32) if return_bins else None
    bvo = _out(pv.shape, np.int32) if return_bins else None
    _fn("compute_qref_and_fawa_first")(_p(pv), _p(uu), _p(vort), _p(pt), _p(tn0), imax, jmax, kmax, int(nd),
                                       int(nnd), int(jb), int(jd), _creal(a), _creal(omega), _creal(dz), _creal(h),
                                       _creal(dphi), _creal(dlambda), _creal(rr), _creal(cp), _p(qref), _p(ubar),
                                       _p(tbar), _p(fawa), _p(ckref), _p(tjk), _p(sjk),
                                       _p(bpv) if return_bins else None, _p(bvo) if return_bins else None)
    if return_bins:
        return qref, ubar, tbar, fawa, ckref, tjk, sjk, bpv, bvo
    return qref, ubar, tbar, fawa, ckref, tjk, sjk


def matrix_b4_inversion(k, jmax, jb, z, statn, qref, ckref, sjk, a, om, dz, h, rr, cp, jd=None):
    z, statn, qref, ckref, sjk = _farr(z, 1), _farr(statn, 1), _farr(qref, 2), _farr(ckref, 2), _farr(sjk, 3)
    kmax = z.shape[0]
    nd = qref.shape[0]
    if jd is None:
        jd = sjk.shape[0] + 2
    n = jd - 2
    qjj, djj, cjj, rj = _out((n, n)), _out((n, n)), _out((n, n)), _out((n,))
    _fn("matrix_b4_inversion")(int(k), int(jmax), kmax, nd, int(jb), int(jd), _p(z), _p(statn), _p(qref), _p(ckref),
                               _p(sjk), _creal(a), _creal(om), _creal(dz), _creal(h), _creal(rr), _creal(cp),
                               _p(qjj), _p(djj), _p(cjj), _p(rj))
    return qjj, djj, cjj, rj


def matrix_after_inversion(k, qjj, djj, cjj, rj, sjk, tjk):
    dt = _Mode.dtype
    for name, arr in (("sjk", sjk), ("tjk", tjk)):  # intent(inout): must be usable in place, like f2py
        if not (isinstance(arr, np.ndarray) and arr.dtype == dt and arr.flags.f_contiguous):
            raise ValueError(f"{name} must be a Fortran-contiguous {np.dtype(dt).name} array (intent(inout))")
    qjj, djj, cjj, rj = _farr(qjj, 2), _farr(djj, 2), _farr(cjj, 2), _farr(rj, 1)
    jd = sjk.shape[0] + 2
    kmax = sjk.shape[2] + 1
    _fn("matrix_after_inversion")(int(k), kmax, jd, _p(qjj), _p(djj), _p(cjj), _p(rj), _p(sjk), _p(tjk))
    return None


def upward_sweep(jmax, jb, sjk, tjk, ckref, tb, qref_over_cor, a, om, dz, h, rr, cp):
    sjk, tjk, ckref, tb, qoc = _farr(sjk, 3), _farr(tjk, 2), _farr(ckref, 2), _farr(tb, 1), _farr(qref_over_cor, 2)
    jd = sjk.shape[0] + 2
    kmax = tb.shape[0]
    nd = ckref.shape[0]
    tref, qref, u = _out((jd, kmax)), _out((nd, kmax)), _out((jd, kmax))
    _fn("upward_sweep")(int(jmax), kmax, nd, int(jb), jd, _p(sjk), _p(tjk), _p(ckref), _p(tb), _p(qoc), _p(tref),
                        _p(qref), _p(u), _creal(a), _creal(om), _creal(dz), _creal(h), _creal(rr), _creal(cp))
    return tref, qref, u


def compute_flux_dirinv_nshem(pv, uu, vv, pt, ncforce, tn0, qref, uref, tref, jb, is_nhem, a, om, dz, h, rr, cp,
                              prefac, return_mask_count=False):
    pv, uu, vv, pt, ncforce = (_farr(x, 3) for x in (pv, uu, vv, pt, ncforce))
    tn0, qref, uref, tref = _farr(tn0, 1), _farr(qref, 2), _farr(uref, 2), _farr(tref, 2)
    imax, jmax, kmax = pv.shape
    nd, jd = qref.shape[0], uref.shape[0]
    outs = [_out((imax, nd, kmax)) for _ in range(8)] + [_out((imax, nd))]
    cnt = (ctypes.c_int64 * 2)(0, 0)
    _fn("compute_flux_dirinv_nshem")(_p(pv), _p(uu), _p(vv), _p(pt), _p(ncforce), _p(tn0), _p(qref), _p(uref),
                                     _p(tref), imax, jmax, kmax, nd, int(jb), jd, int(bool(is_nhem)), _creal(a),
                                     _creal(om), _creal(dz), _creal(h), _creal(rr), _creal(cp), _creal(prefac),
                                     *[_p(o) for o in outs], cnt)
    if return_mask_count:
        return tuple(outs) + ((int(cnt[0]), int(cnt[1])),)
    return tuple(outs)


def compute_lwa_only_nhn22(pv, uu, qref, jb, is_nhem, a, om, dz, h, rr, cp, prefac):
    pv, uu, qref = _farr(pv, 3), _farr(uu, 3), _farr(qref, 2)
    imax, jmax, kmax = pv.shape
    nd = qref.shape[0]
    astarbaro, ubaro = _out((imax, nd)), _out((imax, nd))
    astar1, astar2 = _out((imax, nd, kmax)), _out((imax, nd, kmax))
    _fn("compute_lwa_only_nhn22")(_p(pv), _p(uu), _p(qref), imax, jmax, kmax, nd, int(jb), int(bool(is_nhem)),
                                  _creal(a), _creal(om), _creal(dz), _creal(h), _creal(rr), _creal(cp),
                                  _creal(prefac), _p(astarbaro), _p(ubaro), _p(astar1), _p(astar2))
    return astarbaro, ubaro, astar1, astar2


NATIVE_NAMES = {
    # falwa submodule name -> function names it must export (src/falwa/__init__.py:8-16)
    "compute_qgpv": ["compute_qgpv"],
    "compute_qgpv_direct_inv": ["compute_qgpv_direct_inv"],
    "compute_qref_and_fawa_first": ["compute_qref_and_fawa_first"],
    "matrix_b4_inversion": ["matrix_b4_inversion"],
    "matrix_after_inversion": ["matrix_after_inversion"],
    "upward_sweep": ["upward_sweep"],
    "compute_reference_states": ["compute_reference_states"],
    "compute_flux_dirinv": ["compute_flux_dirinv_nshem"],
    "compute_lwa_only_nhn22": ["compute_lwa_only_nhn22"],
}


def install_as_falwa_modules(reference_src="/root/reference/src", namespace=None):
    """Pre-seed ``sys.modules['falwa.<native module>']`` so the reference's Python imports cleanly.

    ``namespace`` is the object providing the nine callables (default: this module); passing the
    product's ``hn2016_falwa_b200.f90_modules`` makes the reference's unmodified ``oopinterface``
    drive the CUDA kernels instead.
    """
    ns = namespace or sys.modules[__name__]
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
    for modname, funcs in NATIVE_NAMES.items():
        m = types.ModuleType(f"falwa.{modname}")
        for f in funcs:
            setattr(m, f, getattr(ns, f))
        sys.modules[f"falwa.{modname}"] = m
