"""``wrapper`` — the hemispheric 2-D callers of falwa.wrapper (/root/reference/src/falwa/wrapper.py:11-574) on top of
the GPU ``basis`` functions.  Same names, arguments and return values; fields are NumPy arrays (nlat, nlon).

Each hemisphere is analysed on its own sub-domain exactly like the reference: the southern one as it is, the northern
one flipped (and, for vorticity-like fields, negated), with ``ylat[:nlat_s]`` and ``area[:nlat_s]`` in both cases.

Two reference functions cannot run upstream because they pass ``vgrad=`` to ``basis.eqvlat_fawa``, which has no such
parameter (wrapper.py:178-191, :405-436: TypeError).  Here ``eqvlat_bracket_hemispheric`` and ``qgpv_eqlat_lwa_options(
vgrad=...)`` call ``basis.eqvlat_vgrad`` instead, which is what their docstrings describe.
"""
import numpy as np

from . import basis
from .constant import EARTH_RADIUS


def _np(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def barotropic_eqlat_lwa(ylat, vort, area, dmu, n_points, planet_radius=EARTH_RADIUS, return_partitioned_lwa=False):
    """wrapper.py:11-55 -> (qref, lwa)."""
    nlat, nlon = vort.shape
    if n_points is None:
        n_points = nlat
    qref, _ = basis.eqvlat_fawa(ylat, vort, area, n_points, planet_radius=planet_radius, output_fawa=False)
    lwa_result, _ = basis.lwa(nlon, nlat, vort, qref, dmu, return_partitioned_lwa=return_partitioned_lwa)
    return qref, lwa_result


def barotropic_input_qref_to_compute_lwa(ylat, qref, vort, area, dmu, planet_radius=EARTH_RADIUS):
    """wrapper.py:58-86: returns basis.lwa's (lwa, None) tuple, like the reference."""
    nlat, nlon = vort.shape
    return basis.lwa(nlon, nlat, vort, qref, dmu)


def _hemispheres(vort, nlat_s, sign):
    """southern sub-domain, and the northern one flipped (and multiplied by ``sign``) -> two (nlat_s, nlon) arrays"""
    v = _np(vort)
    return v[:nlat_s], np.ascontiguousarray(sign * v[::-1][:nlat_s])


def eqvlat_hemispheric(ylat, vort, area, nlat_s=None, n_points=None, planet_radius=EARTH_RADIUS):
    """wrapper.py:89-135 (the northern qref is NOT negated back, as upstream)."""
    nlat = vort.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    if n_points is None:
        n_points = nlat_s
    south, north = _hemispheres(vort, nlat_s, -1.0)
    q, _ = basis.eqvlat_fawa(ylat[:nlat_s], np.stack([south, north]), _np(area)[:nlat_s], n_points,
                             planet_radius=planet_radius, output_fawa=False)
    qref = np.zeros(nlat)
    qref[:nlat_s] = q[0]
    qref[-nlat_s:] = q[1][::-1]
    return qref


def eqvlat_bracket_hemispheric(ylat, vort, area, nlat_s=None, n_points=None, planet_radius=EARTH_RADIUS, vgrad=None):
    """wrapper.py:138-198 with basis.eqvlat_vgrad in place of the call that fails upstream; ``vgrad`` is passed whole
    and cut to the first nlat_s rows for both hemispheres, as written there."""
    nlat = vort.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    if n_points is None:
        n_points = nlat_s
    south, north = _hemispheres(vort, nlat_s, -1.0)
    g = None if vgrad is None else np.stack([_np(vgrad)[:nlat_s]] * 2)
    q, br = basis.eqvlat_vgrad(ylat[:nlat_s], np.stack([south, north]), _np(area)[:nlat_s], n_points,
                               planet_radius=planet_radius, vgrad=g)
    qref, brac = np.zeros(nlat), np.zeros(nlat)
    qref[:nlat_s] = q[0]
    qref[-nlat_s:] = q[1][::-1]
    if br is not None:
        brac[:nlat_s] = br[0]
        brac[-nlat_s:] = br[1][::-1]
    return qref, brac


def _hemispheric_lwa(vort, qref, dmu, nlat_s, ncforce=None):
    """basis.lwa on vort[:nlat_s] with qref[:nlat_s] and on vort[-nlat_s:] with qref[-nlat_s:] (one batched call)."""
    v, q, d = _np(vort), _np(qref), _np(dmu)
    nlat, nlon = v.shape
    out = np.zeros((nlat, nlon))
    sig = None
    kw_s, kw_n = {}, {}
    if ncforce is not None:
        nc = _np(ncforce)
        kw_s, kw_n = dict(ncforce=nc[:nlat_s]), dict(ncforce=nc[-nlat_s:])
        sig = np.zeros((nlat, nlon))
    ls, ss = basis.lwa(nlon, nlat_s, v[:nlat_s], q[:nlat_s], d[:nlat_s], **kw_s)
    ln, sn = basis.lwa(nlon, nlat_s, v[-nlat_s:], q[-nlat_s:], d[-nlat_s:], **kw_n)
    out[:nlat_s], out[-nlat_s:] = ls, ln
    if sig is not None:
        sig[:nlat_s], sig[-nlat_s:] = ss, sn
    return out, sig


def _qgpv_qref(ylat, vort, area, nlat_s, n_points, planet_radius, vgrad=None):
    """qref of both hemispheres for PV-like fields: the northern hemisphere is analysed on -vort[::-1] and its qref is
    negated back (wrapper.py:233-240)."""
    nlat = vort.shape[0]
    south, north = _hemispheres(vort, nlat_s, -1.0)
    fields = np.stack([south, north])
    a = _np(area)[:nlat_s]
    if vgrad is None:
        q, br = basis.eqvlat_fawa(ylat[:nlat_s], fields, a, n_points, planet_radius=planet_radius, output_fawa=False)
    else:
        gs, gn = _hemispheres(vgrad, nlat_s, -1.0)
        q, br = basis.eqvlat_vgrad(ylat[:nlat_s], fields, a, n_points, planet_radius=planet_radius,
                                   vgrad=np.stack([gs, gn]))
    qref = np.zeros(nlat)
    qref[:nlat_s] = q[0]
    qref[-nlat_s:] = -q[1][::-1]
    brac = None
    if br is not None:
        brac = np.zeros(nlat)
        brac[:nlat_s] = br[0]
        brac[-nlat_s:] = -br[1][::-1]
    return qref, brac


def qgpv_eqlat_lwa(ylat, vort, area, dmu, nlat_s=None, n_points=None, planet_radius=EARTH_RADIUS):
    """wrapper.py:201-267 -> (qref, lwa)."""
    nlat = vort.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    if n_points is None:
        n_points = nlat_s
    qref, _ = _qgpv_qref(ylat, vort, area, nlat_s, n_points, planet_radius)
    lwa_result, _ = _hemispheric_lwa(vort, qref, dmu, nlat_s)
    return qref, lwa_result


def qgpv_eqlat_lwa_ncforce(ylat, vort, ncforce, area, dmu, nlat_s=None, n_points=None, planet_radius=EARTH_RADIUS):
    """wrapper.py:270-341 -> (qref, lwa, capsigma)."""
    nlat = vort.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    if n_points is None:
        n_points = nlat_s
    qref, _ = _qgpv_qref(ylat, vort, area, nlat_s, n_points, planet_radius)
    lwa_result, capsigma = _hemispheric_lwa(vort, qref, dmu, nlat_s, ncforce=ncforce)
    return qref, lwa_result, capsigma


def qgpv_eqlat_lwa_options(ylat, vort, area, dmu, nlat_s=None, n_points=None, vgrad=None, ncforce=None,
                           planet_radius=EARTH_RADIUS):
    """wrapper.py:344-467 -> dict(qref, lwa_result[, brac_result][, capsigma])."""
    nlat = vort.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    if n_points is None:
        n_points = nlat_s
    qref, brac = _qgpv_qref(ylat, vort, area, nlat_s, n_points, planet_radius, vgrad=vgrad)
    lwa_result, capsigma = _hemispheric_lwa(vort, qref, dmu, nlat_s, ncforce=ncforce)
    out = dict(qref=qref, lwa_result=lwa_result)
    if vgrad is not None:
        out["brac_result"] = brac
    if ncforce is not None:
        out["capsigma"] = capsigma
    return out


def qgpv_input_qref_to_compute_lwa(ylat, qref, vort, area, dmu, nlat_s=None, planet_radius=EARTH_RADIUS):
    """wrapper.py:470-514 -> lwa."""
    nlat = vort.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    return _hemispheric_lwa(vort, qref, dmu, nlat_s)[0]


def theta_lwa(ylat, theta, area, dmu, nlat_s=None, n_points=None, planet_radius=EARTH_RADIUS):
    """wrapper.py:517-574: like qgpv_eqlat_lwa without the sign change of the northern hemisphere -> (qref, lwa)."""
    nlat = theta.shape[0]
    if nlat_s is None:
        nlat_s = nlat // 2
    if n_points is None:
        n_points = nlat_s
    south, north = _hemispheres(theta, nlat_s, 1.0)
    q, _ = basis.eqvlat_fawa(ylat[:nlat_s], np.stack([south, north]), _np(area)[:nlat_s], n_points,
                             planet_radius=planet_radius, output_fawa=False)
    qref = np.zeros(nlat)
    qref[:nlat_s] = q[0]
    qref[-nlat_s:] = q[1][::-1]
    lwa_result, _ = _hemispheric_lwa(theta, qref, dmu, nlat_s)
    return qref, lwa_result
