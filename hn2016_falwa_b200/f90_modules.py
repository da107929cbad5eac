"""Drop-in replacements of the nine F2PY functions of falwa v2.3.3, running on the GPU.

Same names, argument names/order and return tuples as the reference's native modules
(``from falwa import compute_qgpv, …`` — /root/reference/src/falwa/__init__.py:8-16; signatures:
SURVEY.md Appendix B), so the reference's unmodified ``oopinterface.py`` can call them
(see INTEGRATION.md).  Inputs may be NumPy arrays of any memory order (copied to the device, like
f2py copies to a contiguous Fortran array) or CUDA tensors already laid out in the reference's
memory order; outputs are NumPy arrays in Fortran order.  All arithmetic is float64.

The fast path (``hn2016_falwa_b200.oopinterface``) does not go through these per-call copies; it keeps
every field on the device and calls the fused entry points.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, dptr, hptr, stream_ptr

# LWA methods of compute_flux_dirinv_nshem / compute_lwa_only_nhn22: 0 = default (exact windowed sums, lwa_window.cu),
# 1 = the reference's double loop, 2 = interval scan (flux_scan.cu), 4 = prefix sums over threshold intervals
# (lwa_prefix.cu; what the pipeline switches to where the PV contours meander far)
LWA_METHODS = (0, 1, 2, 4)


def _dev(x, ndim):
    """numpy (Fortran index order, any strides) or CUDA tensor -> contiguous CUDA tensor in memory order."""
    _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
        return x
    a = np.asarray(x, dtype=np.float64)
    if a.ndim != ndim:
        raise ValueError(f"expected a {ndim}-d array, got shape {a.shape}")
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def _host(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _new(*fshape, dtype=torch.float64):
    """Zero-filled device array for a Fortran-shaped output (stored reversed = memory order)."""
    return torch.zeros(tuple(reversed(fshape)), dtype=dtype, device="cuda")


def _np(t):
    """device memory-order tensor -> NumPy array with the Fortran shape (F-contiguous view)."""
    return t.cpu().numpy().T


def _fshape(t):
    return tuple(reversed(t.shape))


def compute_qgpv(ut, vt, theta, height, t0, stat, aa, omega, dz, hh, rr, cp):
    lib = _lib.load()
    ut, vt, theta = _dev(ut, 3), _dev(vt, 3), _dev(theta, 3)
    nlon, nlat, kmax = _fshape(ut)
    height, t0, stat = _host(height), _host(t0), _host(stat)
    pv, avort = _new(nlon, nlat, kmax), _new(nlon, nlat, kmax)
    check(lib.falwa_b200_compute_qgpv(nlon, nlat, kmax, dptr(ut), dptr(vt), dptr(theta), hptr(height), hptr(t0),
                                      hptr(stat), aa, omega, dz, hh, rr, cp, dptr(pv), dptr(avort), stream_ptr()),
          "compute_qgpv")
    return _np(pv), _np(avort)


def compute_qgpv_direct_inv(jd, uq, vq, tq, height, ts0, tn0, stats, statn, aa, omega, dz, hh, rr, cp):
    lib = _lib.load()
    uq, vq, tq = _dev(uq, 3), _dev(vq, 3), _dev(tq, 3)
    nlon, nlat, kmax = _fshape(uq)
    height, ts0, tn0, stats, statn = (_host(x) for x in (height, ts0, tn0, stats, statn))
    pv, avort = _new(nlon, nlat, kmax), _new(nlon, nlat, kmax)
    check(lib.falwa_b200_compute_qgpv_direct_inv(nlon, nlat, kmax, int(jd), dptr(uq), dptr(vq), dptr(tq),
                                                 hptr(height), hptr(ts0), hptr(tn0), hptr(stats), hptr(statn), aa,
                                                 omega, dz, hh, rr, cp, dptr(pv), dptr(avort), stream_ptr()),
          "compute_qgpv_direct_inv")
    return _np(pv), _np(avort)


def compute_reference_states(pv, uu, pt, stat, jd, npart, maxits, a, om, dz, eps, h, dphi, dlambda, r, cp, rjac,
                             return_bins=False):
    lib = _lib.load()
    pv, uu, pt = _dev(pv, 3), _dev(uu, 3), _dev(pt, 3)
    nlon, nlat, kmax = _fshape(pv)
    stat = _host(stat)
    qref, u, tref = _new(jd, kmax), _new(jd, kmax), _new(jd, kmax)
    bins = _new(nlon, nlat, kmax, dtype=torch.int32) if return_bins else None
    nit = ctypes.c_int(0)
    check(lib.falwa_b200_compute_reference_states(
        dptr(pv), dptr(uu), dptr(pt), hptr(stat), nlon, nlat, kmax, int(jd), int(npart), int(maxits), a, om, dz, eps,
        h, dphi, dlambda, r, cp, rjac, dptr(qref), dptr(u), dptr(tref), ctypes.cast(ctypes.byref(nit), ctypes.c_void_p),
        dptr(bins), stream_ptr()), "compute_reference_states")
    out = (_np(qref), _np(u), _np(tref), int(nit.value))
    return out + (_np(bins),) if return_bins else out


def compute_qref_and_fawa_first(pv, uu, vort, pt, tn0, nd, nnd, jb, jd, a, omega, dz, h, dphi, dlambda, rr, cp,
                                return_bins=False):
    lib = _lib.load()
    pv, uu, vort, pt = _dev(pv, 3), _dev(uu, 3), _dev(vort, 3), _dev(pt, 3)
    imax, jmax, kmax = _fshape(pv)
    tn0 = _host(tn0)
    n = jd - 2
    qref, ubar, tbar, fawa, ckref = (_new(nd, kmax) for _ in range(5))
    tjk, sjk = _new(n, kmax - 1), _new(n, n, kmax - 1)
    bpv = _new(imax, jmax, kmax, dtype=torch.int32) if return_bins else None
    bvo = _new(imax, jmax, kmax, dtype=torch.int32) if return_bins else None
    check(lib.falwa_b200_compute_qref_and_fawa_first(
        dptr(pv), dptr(uu), dptr(vort), dptr(pt), hptr(tn0), imax, jmax, kmax, int(nd), int(nnd), int(jb), int(jd),
        a, omega, dz, h, dphi, dlambda, rr, cp, dptr(qref), dptr(ubar), dptr(tbar), dptr(fawa), dptr(ckref),
        dptr(tjk), dptr(sjk), dptr(bpv), dptr(bvo), stream_ptr()), "compute_qref_and_fawa_first")
    out = tuple(_np(x) for x in (qref, ubar, tbar, fawa, ckref, tjk, sjk))
    return out + (_np(bpv), _np(bvo)) if return_bins else out


def matrix_b4_inversion(k, jmax, jb, z, statn, qref, ckref, sjk, a, om, dz, h, rr, cp, jd=None):
    lib = _lib.load()
    z, statn = _host(z), _host(statn)
    qref, ckref, sjk = _dev(qref, 2), _dev(ckref, 2), _dev(sjk, 3)
    kmax = z.shape[0]
    nd = _fshape(qref)[0]
    if jd is None:
        jd = _fshape(sjk)[0] + 2
    n = jd - 2
    qjj, djj, cjj, rj = _new(n, n), _new(n, n), _new(n, n), _new(n)
    check(lib.falwa_b200_matrix_b4_inversion(int(k), int(jmax), kmax, nd, int(jb), int(jd), hptr(z), hptr(statn),
                                             dptr(qref), dptr(ckref), dptr(sjk), a, om, dz, h, rr, cp, dptr(qjj),
                                             dptr(djj), dptr(cjj), dptr(rj), stream_ptr()), "matrix_b4_inversion")
    return _np(qjj), _np(djj), _np(cjj), _np(rj)


def matrix_after_inversion(k, qjj, djj, cjj, rj, sjk, tjk):
    """sjk and tjk are intent(inout): Fortran-contiguous float64 NumPy arrays updated in place (as with f2py),
    or CUDA tensors updated in place on the device."""
    lib = _lib.load()
    on_host = not isinstance(sjk, torch.Tensor)
    if on_host:
        for name, arr in (("sjk", sjk), ("tjk", tjk)):
            if not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags.f_contiguous):
                raise ValueError(f"{name} must be a Fortran-contiguous float64 array (intent(inout))")
    qjj, djj, cjj, rj = _dev(qjj, 2), _dev(djj, 2), _dev(cjj, 2), _dev(rj, 1)
    sjk_d, tjk_d = _dev(sjk, 3), _dev(tjk, 2)
    jd = _fshape(sjk_d)[0] + 2
    kmax = _fshape(sjk_d)[2] + 1
    check(lib.falwa_b200_matrix_after_inversion(int(k), kmax, jd, dptr(qjj), dptr(djj), dptr(cjj), dptr(rj),
                                                dptr(sjk_d), dptr(tjk_d), stream_ptr()), "matrix_after_inversion")
    if on_host:  # copy back only the level that changed
        sjk[:, :, k - 2] = sjk_d[k - 2].cpu().numpy().T
        tjk[:, k - 2] = tjk_d[k - 2].cpu().numpy()
    return None


def upward_sweep(jmax, jb, sjk, tjk, ckref, tb, qref_over_cor, a, om, dz, h, rr, cp):
    lib = _lib.load()
    sjk, tjk, ckref, qoc = _dev(sjk, 3), _dev(tjk, 2), _dev(ckref, 2), _dev(qref_over_cor, 2)
    tb = _host(tb)
    jd = _fshape(sjk)[0] + 2
    kmax = tb.shape[0]
    nd = _fshape(ckref)[0]
    tref, qref, u = _new(jd, kmax), _new(nd, kmax), _new(jd, kmax)
    check(lib.falwa_b200_upward_sweep(int(jmax), kmax, nd, int(jb), jd, dptr(sjk), dptr(tjk), dptr(ckref), hptr(tb),
                                      dptr(qoc), dptr(tref), dptr(qref), dptr(u), a, om, dz, h, rr, cp, stream_ptr()),
          "upward_sweep")
    return _np(tref), _np(qref), _np(u)


def compute_flux_dirinv_nshem(pv, uu, vv, pt, ncforce, tn0, qref, uref, tref, jb, is_nhem, a, om, dz, h, rr, cp,
                              prefac, return_mask_count=False, method=0):
    lib = _lib.load()
    pv, uu, vv, pt, ncforce = (_dev(x, 3) for x in (pv, uu, vv, pt, ncforce))
    qref, uref, tref = _dev(qref, 2), _dev(uref, 2), _dev(tref, 2)
    tn0 = _host(tn0)
    imax, jmax, kmax = _fshape(pv)
    nd, jd = _fshape(qref)[0], _fshape(uref)[0]
    outs = [_new(imax, nd, kmax) for _ in range(8)] + [_new(imax, nd)]
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda") if return_mask_count else None
    check(lib.falwa_b200_compute_flux_dirinv_nshem(
        dptr(pv), dptr(uu), dptr(vv), dptr(pt), dptr(ncforce), hptr(tn0), dptr(qref), dptr(uref), dptr(tref), imax,
        jmax, kmax, nd, int(jb), jd, int(bool(is_nhem)), a, om, dz, h, rr, cp, prefac, *[dptr(o) for o in outs],
        dptr(cnt), int(method), stream_ptr()), "compute_flux_dirinv_nshem")
    res = tuple(_np(o) for o in outs)
    if return_mask_count:
        c = cnt.cpu().numpy()
        res = res + ((int(c[0]), int(c[1])),)
    return res


def compute_lwa_only_nhn22(pv, uu, qref, jb, is_nhem, a, om, dz, h, rr, cp, prefac, method=0):
    lib = _lib.load()
    pv, uu, qref = _dev(pv, 3), _dev(uu, 3), _dev(qref, 2)
    imax, jmax, kmax = _fshape(pv)
    nd = _fshape(qref)[0]
    astarbaro, ubaro = _new(imax, nd), _new(imax, nd)
    astar1, astar2 = _new(imax, nd, kmax), _new(imax, nd, kmax)
    check(lib.falwa_b200_compute_lwa_only_nhn22(dptr(pv), dptr(uu), dptr(qref), imax, jmax, kmax, nd, int(jb),
                                                int(bool(is_nhem)), a, om, dz, h, rr, cp, prefac, dptr(astarbaro),
                                                dptr(ubaro), dptr(astar1), dptr(astar2), int(method), stream_ptr()),
          "compute_lwa_only_nhn22")
    return _np(astarbaro), _np(ubaro), _np(astar1), _np(astar2)
