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
