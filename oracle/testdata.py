"""TEST INFRASTRUCTURE.  Inputs for parity tests and benchmarks.

* ``load_era_snapshot()``  — the reference repo's own ERA-Interim 2005-01-23 00Z u/v/t snapshot
  (tests/data/2005-01-23-0000-{u,v,t}.nc; 1.5°, 37 levels), repacked by
  ``convert_reference_snapshot()`` into ``tests/golden/era_interim_2005-01-23_00Z.npz`` (the
  original int16 packing + scale/offset, so the decoded values are bit-identical).  Returned in
  the orientation the reference's tests use: latitude ascending, pressure descending
  (tests/test_output_results.py:26-45, 94-97).
* ``synthetic_snapshot()`` — seeded analytic fields of any grid size.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")
ERA_NPZ = os.path.join(GOLDEN_DIR, "era_interim_2005-01-23_00Z.npz")

# byte offsets of the contiguous, unfiltered HDF5 datasets
_REF_FILES = {"u": ("2005-01-23-0000-u.nc", 13673), "v": ("2005-01-23-0000-v.nc", 13674),
              "t": ("2005-01-23-0000-t.nc", 13661)}


def _attr_f8(raw, name):
    key = name.encode() + b"\0"
    end = raw.find(key) + len(key)
    return float(np.frombuffer(raw, dtype="<f8", count=1, offset=end + 40)[0])


def convert_reference_snapshot(ref_data_dir="/root/reference/tests/data", out=ERA_NPZ):
    """Run once in the build container (needs /root/reference); the .npz is committed."""
    pack = {}
    for v, (fn, off) in _REF_FILES.items():
        raw = open(os.path.join(ref_data_dir, fn), "rb").read()
        pack[v] = np.frombuffer(raw, dtype="<i2", count=37 * 121 * 240, offset=off).reshape(37, 121, 240).copy()
        pack[v + "_scale"] = _attr_f8(raw, "scale_factor")
        pack[v + "_offset"] = _attr_f8(raw, "add_offset")
    pack["longitude"] = np.frombuffer(raw, dtype="<f4", count=240, offset=2529).copy()
    pack["latitude"] = np.frombuffer(raw, dtype="<f4", count=121, offset=3489).copy()
    pack["level"] = np.frombuffer(raw, dtype="<i4", count=37, offset=3973).copy()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **pack)
    return out


def load_era_snapshot():
    """-> xlon(240), ylat(121, ascending), plev(37, descending), u, v, t  each (37,121,240) float64."""
    z = np.load(ERA_NPZ)
    fields = []
    for v in "uvt":
        a = z[v].astype(np.float64) * float(z[v + "_scale"]) + float(z[v + "_offset"])
        fields.append(np.ascontiguousarray(a[::-1, ::-1, :]))
    xlon = z["longitude"].astype(np.float64)
    ylat = z["latitude"].astype(np.float64)[::-1].copy()
    plev = z["level"].astype(np.float64)[::-1].copy()
    return (xlon, ylat, plev, *fields)


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
