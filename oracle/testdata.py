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


def load_era_packed():
    """The same snapshot as it lies in the files: -> longitude(240), latitude(121, DEscending), level(37, Ascending),
    raw {u,v,t} int16 (37,121,240) and packing {u,v,t} = (scale_factor, add_offset)."""
    z = np.load(ERA_NPZ)
    raw = {v: z[v].copy() for v in "uvt"}
    packing = {v: (float(z[v + "_scale"]), float(z[v + "_offset"])) for v in "uvt"}
    return (z["longitude"].astype(np.float64), z["latitude"].astype(np.float64), z["level"].astype(np.float64),
            raw, packing)


from hn2016_falwa_b200.synthetic import ERA_PLEV, synthetic_snapshot  # noqa: E402,F401  (shared input generator)


BARO_NPZ = os.path.join(GOLDEN_DIR, "barotropic_vorticity_256x512.npz")


def convert_reference_barotropic(ref_data_dir="/root/reference/tests/data", out=BARO_NPZ):
    """Run once in the build container (needs /root/reference): the reference's sampled absolute vorticity
    (tests/data/barotropic_vorticity.nc, float32 (256, 512) at byte offset 10101) together with
    what the reference's own BarotropicField (pure NumPy) makes of it, the inputs and the answer key of the GPU mirror
    of tests/test_integration.py:107-125."""
    import sys
    raw = open(os.path.join(ref_data_dir, "barotropic_vorticity.nc"), "rb").read()
    av = np.frombuffer(raw, dtype="<f4", count=256 * 512, offset=10101).reshape(256, 512).copy()
    sys.path.insert(0, "/root/reference/src")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_barotropic_field", "/root/reference/src/falwa/barotropic_field.py",
                                                  submodule_search_locations=None)
    import types
    pkg = types.ModuleType("falwa")
    pkg.__path__ = ["/root/reference/src/falwa"]
    sys.modules.setdefault("falwa", pkg)
    from falwa.barotropic_field import BarotropicField      # the reference's unmodified module
    xlon = np.linspace(0, 360., 512, endpoint=False)
    ylat = np.linspace(-90, 90., 256, endpoint=True)
    # the file holds float32; handed over as such (like the reference's test does) NumPy carries float32 through the
    # reference's bin edges and sums, so the answer key proper is made from the same values widened to float64 and the
    # single-precision run is kept for a comparison at the reference's own tolerance (rtol 1e-5)
    c1 = BarotropicField(xlon, ylat, pv_field=av.astype(np.float64))
    c2 = BarotropicField(xlon, ylat, pv_field=av.astype(np.float64), return_partitioned_lwa=True)
    assert np.isclose(c1.lwa, c2.lwa.sum(axis=0)).all()     # the reference's own assertion
    c32 = BarotropicField(xlon, ylat, pv_field=av)
    np.savez_compressed(out, absolute_vorticity=av, equivalent_latitudes=c1.equivalent_latitudes, lwa=c1.lwa,
                        equivalent_latitudes_f32_input=np.asarray(c32.equivalent_latitudes, dtype=np.float64))
    return out


def load_reference_barotropic():
    z = np.load(BARO_NPZ)
    return {k: z[k] for k in z.files}
