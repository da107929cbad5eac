"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Parity status: see oracle/falwa_oracle.cpp header.

NumPy/SciPy restatement of the *host* side of the reference's QGField pipeline
(/root/reference/src/falwa/oopinterface.py, data_storage.py, utilities.py), in functional form,
on top of the restated native routines (oracle/f2py_shim.py).  It exists because
/root/reference is not present on the GPU box: tests there need the whole pipeline as one
callable.  ``oracle/gen_golden.py`` checks, in the build container, that this file reproduces the
reference's own unmodified Python bit for bit and stores the fixtures under tests/golden/.

Everything is kept in the reference's "Fortran index order" (nlon, nlat, kmax) / (nlat, kmax) /
(nlon, nlat) internally; ``run_qgfield`` returns the user-facing (python-order) arrays.
"""
import math

import numpy as np
from scipy.interpolate import interp1d, UnivariateSpline
from scipy.linalg.lapack import dgetrf, dgetri

from . import f2py_shim as _oracle_native

nat = _oracle_native


class use_native:
    """Context manager: run the same host restatement on top of another set of the nine native callables
    (e.g. ``hn2016_falwa_b200.f90_modules`` — the drop-in test of the GPU kernels)."""

    def __init__(self, namespace):
        self.ns = namespace

    def __enter__(self):
        global nat
        self.prev, nat = nat, self.ns
        return self

    def __exit__(self, *exc):
        global nat
        nat = self.prev
        return False

P_GROUND, SCALE_HEIGHT, CP, DRY_GAS_CONSTANT = 1000., 7000., 1004., 287.  # constant.py:5-8
EARTH_RADIUS, EARTH_OMEGA = 6.378e+6, 7.29e-5  # constant.py:9-10


def static_stability_profiles(theta_field, clat, jd, plev_to_height, height):
    """oopinterface.py:241-261 and :517-518 — hemispheric-mean θ(z_p) → smoothing splines."""
    csm = clat[:jd].sum()
    t0_s = np.mean(theta_field[:, :jd, :] * clat[np.newaxis, :jd, np.newaxis], axis=-1).sum(axis=-1) / csm
    t0_n = np.mean(theta_field[:, -jd:, :] * clat[np.newaxis, -jd:, np.newaxis], axis=-1).sum(axis=-1) / csm
    return spline_profiles(t0_s, t0_n, plev_to_height, height) + (t0_s, t0_n)


def spline_profiles(t0_s_plev, t0_n_plev, plev_to_height, height):
    us = UnivariateSpline(x=plev_to_height, y=t0_s_plev)
    un = UnivariateSpline(x=plev_to_height, y=t0_n_plev)
    return us(height), un(height), us.derivative()(height), un.derivative()(height)


def vertical_interpolation(variable, plev_to_height, height):
    """oopinterface.py:537-540."""
    return interp1d(plev_to_height, variable, axis=0, bounds_error=False, kind="linear",
                    fill_value="extrapolate")(height)


def zonal_convergence(field, clat, dlambda, planet_radius, tol=1.e-5):
    """utilities.py:189-233."""
    zd = np.zeros_like(field)
    zd[:, 1:-1] = field[:, 2:] - field[:, :-2]
    zd[:, 0] = field[:, 1] - field[:, -1]
    zd[:, -1] = field[:, 0] - field[:, -2]
    fin = np.abs(clat) > tol
    zd[fin, :] = zd[fin, :] * (-1. / (planet_radius * clat[fin, np.newaxis] * 2. * dlambda))
    return zd


def _nhem_set(full, value, axis):
    """data_storage.py:42-45 (NHemProperty.__set__)."""
    jdim = value.shape[axis]
    slc = [slice(None)] * full.ndim
    slc[axis] = slice(-jdim, None)
    full[tuple(slc)] = value


def _shem_set(full, value, axis):
    """data_storage.py:55-62 (SHemProperty.__set__): written flipped, after the NH value."""
    jdim = value.shape[axis]
    slc = [slice(None)] * full.ndim
    slc[axis] = slice(None, jdim)
    vlc = [slice(None)] * full.ndim
    vlc[axis] = slice(None, None, -1)
    full[tuple(slc)] = value[tuple(vlc)]


def _nhn22_hemisphere(qgpv, u, avort, theta, t0, static_stability, tn0_for_sweep, c):
    """oopinterface.py:1808-1888."""
    nlat, kmax = c["nlat"], c["kmax"]
    nd = nlat // 2 + nlat % 2
    qref_over_sin, ubar, tbar, fawa, ckref, tjk, sjk = nat.compute_qref_and_fawa_first(
        pv=qgpv, uu=u, vort=avort, pt=theta, tn0=t0, nd=nd, nnd=nlat, jb=c["jb"], jd=c["jd"], a=c["a"],
        omega=c["omega"], dz=c["dz"], h=c["h"], dphi=c["dphi"], dlambda=c["dlambda"], rr=c["rr"], cp=c["cp"])
    for k in range(kmax - 1, 1, -1):
        qjj, djj, cjj, rj = nat.matrix_b4_inversion(
            k=k, jmax=nlat, jb=c["jb"], jd=c["jd"], z=np.arange(0, kmax * c["dz"], c["dz"]),
            statn=static_stability, qref=qref_over_sin, ckref=ckref, sjk=sjk, a=c["a"], om=c["omega"],
            dz=c["dz"], h=c["h"], rr=c["rr"], cp=c["cp"])
        lu, piv, info = dgetrf(qjj)
        qjj, info = dgetri(lu, piv)
        nat.matrix_after_inversion(k=k, qjj=qjj, djj=djj, cjj=cjj, rj=rj, sjk=sjk, tjk=tjk)
    tref, qref, uref = nat.upward_sweep(
        jmax=nlat, jb=c["jb"], sjk=sjk, tjk=tjk, ckref=ckref, tb=tn0_for_sweep, qref_over_cor=qref_over_sin,
        a=c["a"], om=c["omega"], dz=c["dz"], h=c["h"], rr=c["rr"], cp=c["cp"])
    aux = dict(qref_over_sin=qref_over_sin, ubar=ubar, tbar=tbar, fawa=fawa, ckref=ckref, tjk=tjk, sjk=sjk)
    return qref_over_sin / (2. * c["omega"]), uref, tref, aux


def _vertical_average(var_3d, c):
    """oopinterface.py:405-420 with lowest_layer_index=1, height_axis=-1."""
    dc = c["dz"] / c["prefactor"]
    return np.sum(var_3d[..., 1:] * np.exp(-c["height"][1:].reshape(1, 1, -1) / c["h"]) * dc, axis=-1)


def _flux_wrapper(pv, uu, vv, pt, ncforce, tn0, qref, uref, tref, c, keep3d=False):
    """oopinterface.py:429-465."""
    outs = nat.compute_flux_dirinv_nshem(
        pv=pv, uu=uu, vv=vv, pt=pt, ncforce=ncforce, tn0=tn0, qref=qref, uref=uref, tref=tref, jb=c["jb"],
        is_nhem=True, a=c["a"], om=c["omega"], dz=c["dz"], h=c["h"], rr=c["rr"], cp=c["cp"],
        prefac=c["prefactor"], return_mask_count=True)
    astar1, astar2, ncforce3d, ua1, ua2, ep1, ep2, ep3, ep4, mask_count = outs
    r = dict(
        astar1=astar1, astar2=astar2, ep4=ep4, mask_count=mask_count,
        astar1_baro=_vertical_average(astar1, c), astar2_baro=_vertical_average(astar2, c),
        ua1baro=_vertical_average(ua1, c), u_baro=_vertical_average(uu[:, -c["equator_idx"]:, :], c),
        ua2baro=_vertical_average(ua2, c), ep1baro=_vertical_average(ep1, c), ep2baro=_vertical_average(ep2, c),
        ep3baro=_vertical_average(ep3, c), ncforce_baro=_vertical_average(ncforce3d, c))
    r["lwa_baro"] = r["astar1_baro"] + r["astar2_baro"]
    if keep3d:
        r.update(ncforce3d=ncforce3d, ua1=ua1, ua2=ua2, ep1=ep1, ep2=ep2, ep3=ep3)
    return r


def run_qgfield(kind, xlon, ylat, plev, u_field, v_field, t_field, kmax=49, maxit=100000, dz=1000., npart=None,
                tol=1.e-5, rjac=0.95, scale_height=SCALE_HEIGHT, cp=CP, dry_gas_constant=DRY_GAS_CONSTANT,
                omega=EARTH_OMEGA, planet_radius=EARTH_RADIUS, eq_boundary_index=5, ncforce=None,
                stages=("interp", "ref", "flux"), keep_aux=False):
    """The full pipeline of ``QGFieldNH18`` (kind="nh18") or ``QGFieldNHN22`` (kind="nhn22") for an odd
    latitude grid that contains the equator, both hemispheres.  Returns a dict of user-facing arrays.

    Follows QGFieldBase.__init__ (oopinterface.py:93-187), interpolate_fields (:467-535),
    compute_reference_states (:550-615, :1604-1674, :1774-1888) and
    compute_lwa_and_barotropic_fluxes (:716-982).
    """
    assert kind in ("nh18", "nhn22")
    xlon, ylat, plev = np.asarray(xlon, float), np.asarray(ylat, float), np.asarray(plev, float)
    nlev, nlat, nlon = u_field.shape
    assert nlat % 2 == 1 and (ylat == 0).sum() == 1
    plev_to_height = -scale_height * np.log(plev / P_GROUND)
    height = np.array([i * dz for i in range(kmax)])
    theta_field = t_field * np.exp(dry_gas_constant / cp * plev_to_height[:, np.newaxis, np.newaxis] / scale_height)
    clat = np.abs(np.cos(np.deg2rad(ylat)))
    equator_idx = int(np.argwhere(ylat == 0)[0][0]) + 1
    jb = eq_boundary_index if kind == "nhn22" else 0
    jd = nlat // 2 + nlat % 2 - jb
    c = dict(nlon=nlon, nlat=nlat, kmax=kmax, jb=jb, jd=jd, equator_idx=equator_idx, a=planet_radius, omega=omega,
             dz=dz, h=scale_height, rr=dry_gas_constant, cp=cp, height=height,
             dphi=np.deg2rad(180. / (nlat - 1)), dlambda=np.deg2rad(360. / nlon),
             npart=npart if npart is not None else nlat,
             prefactor=sum([math.exp(-k * dz / scale_height) * dz for k in range(1, kmax - 1)]))
    out = dict(config=c)

    # ---- interpolate_fields ---------------------------------------------------------------
    t0_s, t0_n, stat_s, stat_n, t0_s_plev, t0_n_plev = static_stability_profiles(
        theta_field, clat, jd, plev_to_height, height)
    iu = np.swapaxes(vertical_interpolation(u_field, plev_to_height, height), 0, 2)
    iv = np.swapaxes(vertical_interpolation(v_field, plev_to_height, height), 0, 2)
    it = np.swapaxes(vertical_interpolation(theta_field, plev_to_height, height), 0, 2)
    if kind == "nh18":  # oopinterface.py:1563-1594
        t0 = 0.5 * (t0_s + t0_n)
        stat = 0.5 * (stat_s + stat_n)
        tn0 = ts0 = t0
        stat_n_used = stat_s_used = stat
        qgpv, avort = nat.compute_qgpv(iu, iv, it, height, t0, stat, planet_radius, omega, dz, scale_height,
                                       dry_gas_constant, cp)
        out["static_stability"] = stat
    else:  # oopinterface.py:1738-1772
        tn0, ts0 = t0_n, t0_s
        stat_n_used, stat_s_used = stat_n, stat_s
        qgpv, avort = nat.compute_qgpv_direct_inv(equator_idx, iu, iv, it, height, t0_s, t0_n, stat_s, stat_n,
                                                  planet_radius, omega, dz, scale_height, dry_gas_constant, cp)
        out["static_stability"] = (stat_s, stat_n)
    f2p = lambda a: np.swapaxes(a, 0, 2)
    out.update(qgpv=f2p(qgpv), interpolated_u=f2p(iu), interpolated_v=f2p(iv), interpolated_theta=f2p(it),
               interpolated_avort=f2p(avort), t0_s=t0_s, t0_n=t0_n, stat_s=stat_s, stat_n=stat_n,
               t0_s_plev=t0_s_plev, t0_n_plev=t0_n_plev, tn0=tn0, ts0=ts0)
    if "ref" not in stages:
        return out

    # ---- compute_reference_states ---------------------------------------------------------
    qref_st = np.zeros((nlat, kmax))
    uref_st = np.zeros((nlat, kmax))
    ptref_st = np.zeros((nlat, kmax))
    aux = {}
    if kind == "nh18":  # oopinterface.py:1604-1674
        kw = dict(stat=stat, jd=equator_idx, npart=c["npart"], maxits=maxit, a=planet_radius, om=omega, dz=dz,
                  eps=tol, h=scale_height, dphi=c["dphi"], dlambda=c["dlambda"], r=dry_gas_constant, cp=cp, rjac=rjac)
        qn, un, tn, nit_n = nat.compute_reference_states(pv=qgpv, uu=iu, pt=it, **kw)
        qs, us, ts, nit_s = nat.compute_reference_states(pv=-qgpv[:, ::-1, :], uu=iu[:, ::-1, :], pt=it[:, ::-1, :],
                                                         **kw)
        out["num_of_iter"] = (nit_n, nit_s)
        out["nonconvergent"] = (nit_n >= maxit) or (nit_s >= maxit)
    else:  # oopinterface.py:1774-1806
        qn, un, tn, aux["n"] = _nhn22_hemisphere(qgpv, iu, avort, it, tn0, stat_n, tn0, c)
        qs, us, ts, aux["s"] = _nhn22_hemisphere(-qgpv[:, ::-1, :], iu[:, ::-1, :], avort[:, ::-1, :],
                                                 it[:, ::-1, :], ts0, stat_s, tn0, c)
        out["nonconvergent"] = False
    for full, vn, vs in ((qref_st, qn, qs), (uref_st, un, us), (ptref_st, tn, ts)):
        _nhem_set(full, vn, 0)
        _shem_set(full, vs, 0)
    # data_storage.py:170-179
    qref_correct_unit = qref_st * 2 * omega * np.sin(np.deg2rad(ylat[:, np.newaxis]))
    out.update(qref=qref_correct_unit.T, uref=uref_st.T, ptref=ptref_st.T, qref_stored=qref_st.T)
    if keep_aux:
        out["aux"] = aux
    if "flux" not in stages or out["nonconvergent"]:
        return out

    # ---- compute_lwa_and_barotropic_fluxes ------------------------------------------------
    if ncforce is None:
        ncf = np.zeros((nlon, nlat, kmax))
    else:
        ncf = np.swapaxes(ncforce, 0, 2)
    rn = _flux_wrapper(qgpv, iu, iv, it, ncf, tn0, qref_correct_unit[-equator_idx:, :], uref_st[-jd:, :],
                       ptref_st[-jd:, :], c)
    rs = _flux_wrapper(-qgpv[:, ::-1, :], iu[:, ::-1, :], -iv[:, ::-1, :], it[:, ::-1, :], ncf[:, ::-1, :], ts0,
                       -qref_correct_unit[(equator_idx - 1)::-1, :], uref_st[(jd - 1)::-1, :],
                       ptref_st[(jd - 1)::-1, :], c)
    # oopinterface.py:764-793: SH ep2/ep3 swapped and negated, ncforce negated
    rs["ep2baro"], rs["ep3baro"] = -rs["ep3baro"], -rs["ep2baro"]
    rs["ncforce_baro"] = -rs["ncforce_baro"]
    baro = {}
    for name in ("lwa_baro", "u_baro", "ua1baro", "ua2baro", "ep1baro", "ep2baro", "ep3baro", "ep4",
                 "astar1_baro", "astar2_baro", "ncforce_baro"):
        full = np.zeros((nlon, nlat))
        _nhem_set(full, rn[name], 1)
        _shem_set(full, rs[name], 1)
        baro[name] = full
    lwa = np.zeros((nlon, nlat, kmax))
    _nhem_set(lwa, np.abs(rn["astar1"] + rn["astar2"]), 1)
    _shem_set(lwa, np.abs(rs["astar1"] + rs["astar2"]), 1)
    # oopinterface.py:906-944
    div_emf = np.swapaxes((baro["ep2baro"] - baro["ep3baro"]) / (2 * planet_radius * c["dphi"] * clat), 0, 1)
    zsum = np.swapaxes(baro["ua1baro"] + baro["ua2baro"] + baro["ep1baro"], 0, 1)
    conv = zonal_convergence(zsum, clat, c["dlambda"], planet_radius)
    # oopinterface.py:964-982
    fv3d = -(((iu - uref_st[np.newaxis, :, :]) * iv) * clat[np.newaxis, :, np.newaxis])
    out.update(
        adv_flux_f1=baro["ua1baro"].T, adv_flux_f2=baro["ua2baro"].T, adv_flux_f3=baro["ep1baro"].T,
        convergence_zonal_advective_flux=conv, divergence_eddy_momentum_flux=div_emf,
        meridional_heat_flux=baro["ep4"].T, lwa_baro=baro["lwa_baro"].T, u_baro=baro["u_baro"].T, lwa=f2p(lwa),
        ncforce_baro=baro["ncforce_baro"].T, flux_vector_lambda_baro=(baro["ua1baro"] + baro["ua2baro"] +
                                                                      baro["ep1baro"]).T,
        flux_vector_phi_baro=_vertical_average(fv3d, c).T, ep2baro=baro["ep2baro"].T, ep3baro=baro["ep3baro"].T,
        astar1_baro=baro["astar1_baro"].T, astar2_baro=baro["astar2_baro"].T,
        mask_count=(rn["mask_count"], rs["mask_count"]))
    return out


USER_FIELDS = ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta", "qref", "uref", "ptref",
               "adv_flux_f1", "adv_flux_f2", "adv_flux_f3", "convergence_zonal_advective_flux",
               "divergence_eddy_momentum_flux", "meridional_heat_flux", "lwa_baro", "u_baro", "lwa",
               "ncforce_baro", "flux_vector_lambda_baro", "flux_vector_phi_baro")
