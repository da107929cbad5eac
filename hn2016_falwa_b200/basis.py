"""``basis`` — equivalent latitude and 2-D local wave activity on the GPU with the signatures of falwa.basis
(/root/reference/src/falwa/basis.py:11-205).  Fields are NumPy arrays or CUDA tensors (nlat, nlon); a leading batch
axis (T, nlat, nlon) is accepted as an extension and returns batched results.  NumPy in -> NumPy out, CUDA in -> CUDA
out.  Every number is produced by the kernels of csrc/baro2d.cu (Section C of include/falwa_b200.h)."""
import numpy as np
import torch

from . import _lib
from ._lib import check, dptr, hptr, stream_ptr


def _dev(x, ndim_single=2):
    """-> (contiguous float64 CUDA tensor with a leading batch axis, was_batched, was_numpy)"""
    was_numpy = not isinstance(x, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda() if was_numpy else x
    assert t.is_cuda and t.dtype == torch.float64, "falwa_b200.basis needs float64 data"
    batched = t.dim() == ndim_single + 1
    assert t.dim() in (ndim_single, ndim_single + 1), f"expected {ndim_single}-d fields (or a batch of them)"
    return (t if batched else t[None]).contiguous(), batched, was_numpy


def _ret(t, batched, was_numpy):
    if t is None:
        return None
    t = t if batched else t[0]
    return t.cpu().numpy() if was_numpy else t


def _area(area, nlat, nlon):
    a, _, _ = _dev(area)
    assert tuple(a.shape[1:]) == (nlat, nlon), "area must be (nlat, nlon)"
    return a[0].contiguous()


def _eqvlat(mode, ylat, vort, area, n_points, planet_radius, vgrad, want_aux):
    _lib.require_cuda()
    lib = _lib.load()
    v, batched, was_numpy = _dev(vort)
    T, nlat, nlon = (int(s) for s in v.shape)
    ylat = np.ascontiguousarray(ylat, dtype=np.float64)
    assert ylat.size == nlat, "ylat must have one entry per row of vort"
    a = _area(area, nlat, nlon)
    g = None
    if vgrad is not None:
        g, _, _ = _dev(vgrad)
        assert tuple(g.shape) == tuple(v.shape), "vgrad must have the shape of vort"
    qref = torch.empty((T, nlat), dtype=torch.float64, device=v.device)
    aux = None
    if want_aux:
        aux = torch.empty((T, nlat - 1 if mode == 0 else nlat), dtype=torch.float64, device=v.device)
    check(lib.falwa_b200_eqvlat_2d(mode, T, nlon, nlat, int(n_points), dptr(v), dptr(a), dptr(g), hptr(ylat),
                                   float(planet_radius), dptr(qref), dptr(aux), stream_ptr()), "eqvlat_2d")
    return _ret(qref, batched, was_numpy), _ret(aux, batched, was_numpy)


def eqvlat_fawa(ylat, vort, area, n_points, planet_radius=6.378e+6, output_fawa=True):
    """basis.py:11-64 -> (qref (nlat), fawa (nlat-1) or None)."""
    return _eqvlat(0, ylat, vort, area, n_points, planet_radius, None, bool(output_fawa))


def eqvlat_vgrad(ylat, vort, area, n_points, planet_radius=6.378e+6, vgrad=None):
    """basis.py:67-134 -> (q_part (nlat), brac (nlat) or None)."""
    return _eqvlat(1, ylat, vort, area, n_points, planet_radius, vgrad, vgrad is not None)


def lwa(nlon, nlat, vort, q_part, dmu, return_partitioned_lwa=False, ncforce=None):
    """basis.py:137-205 -> (lwa (nlat, nlon) or (2, nlat, nlon) = (anticyclonic, cyclonic), bigsigma or None)."""
    _lib.require_cuda()
    lib = _lib.load()
    v, batched, was_numpy = _dev(vort)
    T = int(v.shape[0])
    assert tuple(v.shape[1:]) == (nlat, nlon), "vort must be (nlat, nlon)"
    q, _, _ = _dev(q_part, ndim_single=1)
    assert tuple(q.shape) == (T, nlat), "q_part must be (nlat)"
    dmu = np.ascontiguousarray(dmu, dtype=np.float64)
    assert dmu.size == nlat, "dmu must be (nlat)"
    nc = sig = None
    if ncforce is not None:
        nc, _, _ = _dev(ncforce)
        assert tuple(nc.shape) == tuple(v.shape), "ncforce must have the shape of vort"
        sig = torch.empty_like(v)
    shape = (T, 2, nlat, nlon) if return_partitioned_lwa else (T, nlat, nlon)
    out = torch.empty(shape, dtype=torch.float64, device=v.device)
    check(lib.falwa_b200_lwa_2d(T, nlon, nlat, dptr(v), dptr(q), hptr(dmu), dptr(nc), dptr(out), dptr(sig),
                                int(bool(return_partitioned_lwa)), stream_ptr()), "lwa_2d")
    return _ret(out, batched, was_numpy), _ret(sig, batched, was_numpy)
