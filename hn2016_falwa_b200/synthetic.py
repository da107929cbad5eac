"""Seeded synthetic u/v/T snapshots at any grid size: ERA5 0.25° (1440x721) and ERA-Interim
1.5° (240x121) shapes for benchmarks and tests.  Pure NumPy input generation; no reference or oracle code."""
import numpy as np

ERA_PLEV = np.array([1000, 975, 950, 925, 900, 875, 850, 825, 800, 775, 750, 700, 650, 600, 550, 500, 450, 400,
                     350, 300, 250, 225, 200, 175, 150, 125, 100, 70, 50, 30, 20, 10, 7, 5, 3, 2, 1], dtype=np.float64)


def synthetic_snapshot(nlon=240, nlat=121, plev=ERA_PLEV, seed=1234, t_index=0):
    """Seeded analytic snapshot: a Gaussian jet per hemisphere, six Rossby-like waves per
    hemisphere in u, v, T, and a statically stable T(p) profile (dθ/dz > 0 everywhere).

    Returns xlon, ylat, plev, u, v, t with fields of shape (nlev, nlat, nlon), float64."""
    rng = np.random.default_rng(seed + t_index)
    xlon = np.linspace(0.0, 360.0, nlon, endpoint=False)
    ylat = np.linspace(-90.0, 90.0, nlat, endpoint=True)
    lam = np.deg2rad(xlon)[None, None, :]
    phi = ylat[None, :, None]
    p = plev[:, None, None]
    zp = -7000.0 * np.log(p / 1000.0)
    # temperature: troposphere lapse 6.5 K/km to 11 km (min 210 K), slowly warming stratosphere,
    # plus a pole-to-equator gradient decaying with height
    t_prof = np.maximum(288.0 - 6.5e-3 * zp, 212.0) + np.maximum(zp - 20000.0, 0.0) * 1.2e-3
    t = t_prof + 25.0 * (np.cos(np.deg2rad(phi)) ** 2 - 0.5) * np.exp(-zp / 12000.0)
    vert = np.exp(-((zp - 10500.0) / 9000.0) ** 2)
    u = np.zeros((plev.size, nlat, nlon))
    v = np.zeros_like(u)
    for sgn in (1.0, -1.0):
        env = np.exp(-((phi - sgn * 45.0) / 15.0) ** 2)
        u = u + 40.0 * env * vert
        for m in range(1, 7):
            amp = rng.uniform(2.0, 10.0)
            ph0 = rng.uniform(0.0, 2 * np.pi) + 2 * np.pi * t_index / (4.0 * 5.0 * m)
            tilt = rng.uniform(-0.5, 0.5)
            wave = np.cos(m * lam + ph0 + tilt * np.deg2rad(phi - sgn * 45.0) * 4.0)
            wave_s = np.sin(m * lam + ph0 + tilt * np.deg2rad(phi - sgn * 45.0) * 4.0)
            u = u + 0.5 * amp * env * vert * wave
            v = v + amp * env * vert * wave_s
            t = t + 0.3 * amp * env * np.exp(-zp / 15000.0) * wave
    return xlon, ylat, plev.copy(), u, v, t


def upsampled_snapshot(xlon, ylat, u, v, t, nlon=1440, nlat=721):
    """a real coarse snapshot (e.g. the 1.5° ERA-Interim one) brought to an ERA5-shaped grid by
    bilinear interpolation, periodic in longitude; the levels stay as they are.  ``ylat`` ascending, pole to pole.

    Returns xlon2, ylat2, u2, v2, t2 with fields of shape (nlev, nlat, nlon), float64."""
    xlon = np.asarray(xlon, dtype=np.float64)
    ylat = np.asarray(ylat, dtype=np.float64)
    xlon2 = np.linspace(0.0, 360.0, nlon, endpoint=False)
    ylat2 = np.linspace(-90.0, 90.0, nlat, endpoint=True)
    fx = (xlon2 - xlon[0]) / (xlon[1] - xlon[0])
    i0 = np.floor(fx).astype(np.int64)
    wx = fx - i0
    i1 = (i0 + 1) % xlon.size
    i0 = i0 % xlon.size
    fy = np.clip((ylat2 - ylat[0]) / (ylat[1] - ylat[0]), 0.0, ylat.size - 1.0)
    j0 = np.minimum(np.floor(fy).astype(np.int64), ylat.size - 2)
    wy = (fy - j0)[None, :, None]
    out = []
    for a in (u, v, t):
        a = np.asarray(a, dtype=np.float64)
        rows = a[:, j0, :] * (1.0 - wy) + a[:, j0 + 1, :] * wy
        out.append(np.ascontiguousarray(rows[:, :, i0] * (1.0 - wx) + rows[:, :, i1] * wx))
    return (xlon2, ylat2, *out)


def load_packed_snapshot(npz_path):
    """Decode a CF-packed snapshot kept as .npz (int16 u, v, t with `<name>_scale` / `<name>_offset`, `longitude`,
    `latitude` descending, `level` ascending — the layout of the reference's tests/data files) into the orientation the
    QGField classes expect: latitude ascending, pressure descending.  -> xlon, ylat, plev, u, v, t (float64)."""
    z = np.load(npz_path)
    fields = [np.ascontiguousarray((z[n].astype(np.float64) * float(z[n + "_scale"]) + float(z[n + "_offset"]))[::-1, ::-1, :])
              for n in "uvt"]
    return (z["longitude"].astype(np.float64), z["latitude"].astype(np.float64)[::-1].copy(),
            z["level"].astype(np.float64)[::-1].copy(), *fields)
