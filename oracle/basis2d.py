"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

NumPy restatement of the reference's 2-D (barotropic) path: ``BarotropicField`` →
``basis.eqvlat_fawa(output_fawa=False)`` + ``basis.lwa`` (/root/reference/src/falwa/
barotropic_field.py:44-125, basis.py:11-64, :137-205).  Pinned by the inline answer key of the
reference's tests/test_barotropic_field.py:26-54 (stored in ANSWER_KEY below) and, bit for bit,
by the reference module itself in oracle/gen_golden.py.
"""
from math import pi

import numpy as np

EARTH_RADIUS, EARTH_OMEGA = 6.378e+6, 7.29e-5


def grid_area(ylat, nlon, planet_radius=EARTH_RADIUS):
    """barotropic_field.py:84-91 (dphi = pi/(nlat-1) for every row)."""
    nlat = ylat.size
    dphi = pi / (nlat - 1) * np.ones((nlat))
    area = 2. * pi * planet_radius ** 2 * (np.cos(ylat[:, np.newaxis] * pi / 180.) * dphi[:, np.newaxis]) \
        / float(nlon) * np.ones((nlat, nlon))
    return area, dphi


def eqvlat(ylat, vort, area, n_points, planet_radius=EARTH_RADIUS):
    """basis.py:37-53: uniform bins between min and max, np.digitize (right-open: the maximum
    element falls outside the last bin and is dropped), cumulative area from the low end,
    arcsin, np.interp."""
    vort_min, vort_max = np.amin(vort), np.amax(vort)
    vort_bins = np.linspace(vort_min, vort_max, n_points, endpoint=True)
    vort_flat, area_flat = vort.flatten(), area.flatten()
    inds = np.digitize(vort_flat, vort_bins)
    aq = np.cumsum([0] + [area_flat[np.where(inds == i)].sum() for i in range(1, vort_bins.size)])
    eqv_gaussian_grid = aq / (2 * pi * planet_radius ** 2) - 1.0
    eqv_latitude = np.rad2deg(np.arcsin(eqv_gaussian_grid))
    return np.interp(ylat, eqv_latitude, vort_bins)


def lwa(nlon, nlat, vort, q_part, dmu):
    """basis.py:176-205 with return_partitioned_lwa=False, ncforce=None: row j belongs to the
    cyclonic branch (rows <= j, vort_e > 0), rows > j with vort_e < 0 are anticyclonic; the last
    row stays 0.  Vectorised over j in blocks; the per-(j, lon) sums run over
    latitude in the same order as the reference's np.sum(axis=0)."""
    out = np.zeros((nlat, nlon))
    rows = np.arange(nlat)[:, np.newaxis]
    for j in range(0, nlat - 1):
        vort_e = vort - q_part[j]
        anti = np.where((vort_e < 0) & (rows > j), -1.0, 0.0)
        cyc = np.where((vort_e > 0) & (rows <= j), 1.0, 0.0)
        out[j, :] = np.sum(vort_e * anti * dmu[:, np.newaxis], axis=0) + \
            np.sum(vort_e * cyc * dmu[:, np.newaxis], axis=0)
    return out


def barotropic_eqvlat_lwa(xlon, ylat, pv_field, planet_radius=EARTH_RADIUS):
    """BarotropicField(xlon, ylat, pv_field).equivalent_latitudes / .lwa with all defaults."""
    nlat, nlon = ylat.size, xlon.size
    area, dphi = grid_area(ylat, nlon, planet_radius)
    clat = np.abs(np.cos(np.deg2rad(ylat)))
    eq = eqvlat(ylat, pv_field, area, nlat, planet_radius)
    return eq, lwa(nlon, nlat, pv_field, eq, planet_radius * clat * dphi)


def synthetic_vorticity(nlat=721, nlon=1440, seed=0, t_index=0):
    """2Ω sinφ + 8e-5 cosφ exp(-((φ-36)/10)^2) cos(3λ + phase(t)) + 1e-6 N(0,1)."""
    rng = np.random.default_rng(seed + t_index)
    xlon = np.linspace(0, 360, nlon, endpoint=False)
    ylat = np.linspace(-90., 90., nlat, endpoint=True)
    phase = 2 * np.pi * t_index / 40.0
    lat_var = 8.e-5 * np.cos(np.deg2rad(ylat)) * np.exp(-(ylat - 36.) ** 2 / 10. ** 2)
    pv = lat_var[:, None] * np.cos(3 * np.deg2rad(xlon) + phase)[None, :] \
        + 2 * EARTH_OMEGA * np.sin(np.deg2rad(ylat[:, None])) + 1e-6 * rng.standard_normal((nlat, nlon))
    return xlon, ylat, pv


# tests/test_barotropic_field.py:26-54 (np.allclose rtol=1e-5, atol=1e-8 in the reference test)
ANSWER_KEY = dict(
    eqv_lat=[-1.45800000e-04, -1.42949614e-04, -1.40099229e-04, -1.37248843e-04, -1.33297467e-04, -1.28392619e-04,
             -1.20907978e-04, -1.11063638e-04, -1.01211730e-04, -8.14807659e-05, -7.16202750e-05, -6.17571838e-05,
             -4.20180950e-05, -3.21514559e-05, -1.24095227e-05, -2.54052114e-06, 1.70060059e-05, 2.58382415e-05,
             4.22608624e-05, 4.86083082e-05, 6.36631843e-05, 8.87920495e-05, 1.11234061e-04, 1.17873494e-04,
             1.25186315e-04, 1.31933407e-04, 1.36129493e-04, 1.40325579e-04, 1.45482242e-04, 1.50420449e-04,
             1.50420449e-04],
    lwa_zonal_mean=[0.00000000e+00, 0.00000000e+00, 0.00000000e+00, 0.00000000e+00, 2.78559787e-02, 7.10018930e-01,
                    1.15941636e+00, 1.21253264e+00, 1.81290612e+00, 0.00000000e+00, 0.00000000e+00, 1.49792983e+00,
                    0.00000000e+00, 1.20073402e+00, 0.00000000e+00, 1.69682067e+00, 0.00000000e+00, 3.18338839e+00,
                    5.71579135e+00, 1.32708939e+01, 2.01584444e+01, 1.95740843e+01, 1.18914496e+01, 6.79453254e+00,
                    3.79448248e+00, 2.08640756e+00, 1.20716646e+00, 6.47041885e-01, 2.62663095e-01, 1.88963620e-16,
                    0.00000000e+00],
    lwa_row20=[41.20590115, 39.32040089, 33.84846603, 25.32572766, 14.58645081, 5.34276464, 0., 8.64894682,
               16.38582659, 24.10885087, 26.82811918, 24.10885087, 16.38582659, 8.64894682, 0., 5.34276464,
               14.58645081, 25.32572766, 33.84846603, 39.32040089, 41.20590115, 39.32040089, 33.84846603,
               25.32572766, 14.58645081, 5.34276464, 0., 8.64894682, 16.38582659, 24.10885087, 26.82811918,
               24.10885087, 16.38582659, 8.64894682, 0., 5.34276464, 14.58645081, 25.32572766, 33.84846603,
               39.32040089, 41.20590115, 39.32040089, 33.84846603, 25.32572766, 14.58645081, 5.34276464, 0.,
               8.64894682, 16.38582659, 24.10885087, 26.82811918, 24.10885087, 16.38582659, 8.64894682, 0.,
               5.34276464, 14.58645081, 25.32572766, 33.84846603, 39.32040089])


# ---- the rest of basis.py / wrapper.py --------------------------------------------------------
def _bin_sums(vort, area, n_points, weights=None):
    """np.digitize membership of every grid point in the n_points uniform bins between min and max (the maximum
    falls outside and is dropped) -> bin edges, area per bin, and optionally sum(area * weights) per bin; index 0 is
    the never-populated bin below the minimum (basis.py:41-50, :100-120)."""
    edges = np.linspace(np.amin(vort), np.amax(vort), n_points, endpoint=True)
    inds = np.digitize(vort.ravel(), edges)
    a = area.ravel()
    keep = inds < n_points
    asum = np.zeros(n_points)
    np.add.at(asum, inds[keep], a[keep])
    wsum = None
    if weights is not None:
        wsum = np.zeros(n_points)
        np.add.at(wsum, inds[keep], (a * weights.ravel())[keep])
    return edges, asum, wsum


def _eqv_latitude(asum, planet_radius):
    return np.rad2deg(np.arcsin(np.cumsum(asum) / (2 * pi * planet_radius ** 2) - 1.0))


def eqvlat_fawa(ylat, vort, area, n_points, planet_radius=EARTH_RADIUS, output_fawa=True):
    """basis.py:11-64.  FAWA = -(cbar + cref)/(2 pi a) on rows 1..nlat-1, with cref the cumulative area-weighted
    vorticity interpolated to ylat and cbar the trapezoidal sum of zonal-mean vorticity x row area from the pole."""
    edges, asum, wsum = _bin_sums(vort, area, n_points, weights=vort if output_fawa else None)
    lat = _eqv_latitude(asum, planet_radius)
    qref = np.interp(ylat, lat, edges)
    if not output_fawa:
        return qref, None
    cref = np.interp(ylat, lat, np.cumsum(wsum))
    rowint = vort.mean(axis=-1) * area.sum(axis=1)
    cbar = np.zeros_like(cref)
    for j in range(ylat.size - 2, 0, -1):
        cbar[j] = cbar[j + 1] + 0.5 * (rowint[j + 1] + rowint[j])
    return qref, -(cbar[1:] + cref[1:]) / (2 * pi * planet_radius)


def eqvlat_vgrad(ylat, vort, area, n_points, planet_radius=EARTH_RADIUS, vgrad=None):
    """basis.py:67-134: bracket = mean over the two neighbouring bins of the area-weighted bin average of vgrad."""
    edges, asum, wsum = _bin_sums(vort, area, n_points, weights=vgrad)
    lat = _eqv_latitude(asum, planet_radius)
    q_part = np.interp(ylat, lat, edges)
    if vgrad is None:
        return q_part, None
    dp = np.divide(wsum, asum, out=np.zeros_like(wsum), where=asum != 0.)
    brac = np.zeros(n_points)
    brac[1:-1] = 0.5 * (dp[:-2] + dp[2:])
    return q_part, np.interp(ylat, lat, brac)


def lwa_full(nlon, nlat, vort, q_part, dmu, return_partitioned_lwa=False, ncforce=None):
    """basis.py:137-205 including the partitioned output (anticyclonic, cyclonic) and the non-conservative term."""
    parts = np.zeros((2, nlat, nlon))
    sigma = np.zeros((nlat, nlon)) if ncforce is not None else None
    rows = np.arange(nlat)[:, np.newaxis]
    w = dmu[:, np.newaxis]
    for j in range(nlat - 1):
        e = vort - q_part[j]
        anti = np.where((e < 0) & (rows > j), -1.0, 0.0)
        cyc = np.where((e > 0) & (rows <= j), 1.0, 0.0)
        parts[0, j] = np.sum(e * anti * w, axis=0)
        parts[1, j] = np.sum(e * cyc * w, axis=0)
        if ncforce is not None:
            sigma[j] = np.sum(ncforce * (anti + cyc) * w, axis=0)
    return (parts if return_partitioned_lwa else parts.sum(axis=0)), sigma


def hemispheric(ylat, field, area, dmu, nlat_s=None, n_points=None, planet_radius=EARTH_RADIUS, sign=-1.0,
                negate_back=True, vgrad=None, ncforce=None, with_lwa=True, vgrad_whole=False):
    """The common structure of wrapper.py:89-574: southern rows as they are, northern rows flipped and multiplied by
    ``sign``; the northern qref (and bracket) flipped back and, for ``negate_back``, negated.  Returns dict(qref,
    lwa_result, brac_result, capsigma) with None for what was not asked."""
    nlat, nlon = field.shape
    nlat_s = nlat // 2 if nlat_s is None else nlat_s
    n_points = nlat_s if n_points is None else n_points
    back = -1.0 if negate_back else 1.0
    south, north = field[:nlat_s], (sign * field[::-1])[:nlat_s]
    gs = gn = None
    if vgrad is not None:
        gs = vgrad[:nlat_s]
        gn = gs if vgrad_whole else (-vgrad[::-1])[:nlat_s]
    fn = eqvlat_vgrad if vgrad is not None else (lambda *a, **k: eqvlat_fawa(*a, output_fawa=False, **k))
    kw = (lambda g: dict(vgrad=g)) if vgrad is not None else (lambda g: {})
    q1, b1 = fn(ylat[:nlat_s], south, area[:nlat_s], n_points, planet_radius=planet_radius, **kw(gs))
    q2, b2 = fn(ylat[:nlat_s], north, area[:nlat_s], n_points, planet_radius=planet_radius, **kw(gn))
    qref = np.zeros(nlat)
    qref[:nlat_s], qref[-nlat_s:] = q1, back * q2[::-1]
    out = dict(qref=qref, lwa_result=None, brac_result=None, capsigma=None)
    if vgrad is not None:
        brac = np.zeros(nlat)
        brac[:nlat_s], brac[-nlat_s:] = b1, back * b2[::-1]
        out["brac_result"] = brac
    if with_lwa:
        out["lwa_result"], out["capsigma"] = hemispheric_lwa(field, qref, dmu, nlat_s, ncforce)
    return out


def hemispheric_lwa(field, qref, dmu, nlat_s, ncforce=None):
    nlat, nlon = field.shape
    res = np.zeros((nlat, nlon))
    sig = np.zeros((nlat, nlon)) if ncforce is not None else None
    for sl in (slice(None, nlat_s), slice(nlat - nlat_s, None)):
        l, s = lwa_full(nlon, nlat_s, field[sl], qref[sl], dmu[sl], ncforce=None if ncforce is None else ncforce[sl])
        res[sl] = l
        if sig is not None:
            sig[sl] = s
    return res, sig
