"""File formats either side of the path: archived fields in, result files out.

``open_fields``   NetCDF-3 (classic / 64-bit offset) files, memory-mapped and NOT decoded: the variables keep their
                  storage type (big-endian packed int16, float32 or float64) together with ``scale_factor`` /
                  ``add_offset`` / ``_FillValue``.  ``QGDataset`` ships those bytes to the GPU as they are and decodes,
                  byte-swaps and reorders them there (csrc/ingest.cu), replacing xarray's CF decoding on the host and the
                  np.flip of src/falwa/xarrayinterface.py:384-405.
``write_results`` the outputs of a ``QGDataset`` (or any SimpleDataset) as a NetCDF-3 file with the variable layout of
                  the reference's production script (scripts/lubis_et_al_2025/cal_lwa_budget_NHN22.py:79-158).

Reader and writer are SciPy's ``scipy.io.netcdf_file`` (the only NetCDF implementation in this image).  NetCDF-4 /
HDF5 files need xarray + netCDF4: open them with ``xr.open_dataset(path, mask_and_scale=False)`` and pass the dataset
to ``QGDataset`` — undecoded xarray variables take the same device path.
"""
from collections import OrderedDict

import numpy as np
from scipy.io import netcdf_file

from .xarrayinterface import SimpleDataArray, SimpleDataset, _NAMES_T, _NAMES_U, _NAMES_V

UNITS = {  # cal_lwa_budget_NHN22.py:96-130
    "qgpv": "1/s", "interpolated_u": "m/s", "interpolated_v": "m/s", "interpolated_theta": "K", "qref": "1/s",
    "uref": "m/s", "ptref": "K", "lwa": "m/s", "adv_flux_f1": "m**2/s**2", "adv_flux_f2": "m**2/s**2",
    "adv_flux_f3": "m**2/s**2", "convergence_zonal_advective_flux": "m/s**2", "divergence_eddy_momentum_flux": "m/s**2",
    "meridional_heat_flux": "m/s**2", "lwa_baro": "m/s", "u_baro": "m/s", "ncforce_baro": "m/s**2",
    "flux_vector_lambda_baro": "m**2/s**2", "flux_vector_phi_baro": "m**2/s**2"}


def _attrs(var):
    out = {}
    for k, v in var._attributes.items():
        if isinstance(v, bytes):
            v = v.decode("ascii", "replace")
        elif isinstance(v, np.ndarray) and v.size == 1:
            v = v.reshape(()).item()
        out[k] = v
    return out


def open_fields(*paths, var_names=None):
    """-> SimpleDataset with the u, v, t variables of the given NetCDF-3 files (one file holding all three, or one file
    per variable), undecoded and memory-mapped.  The files stay open as long as the dataset lives."""
    files = [netcdf_file(p, "r", mmap=True, maskandscale=False) for p in paths]
    wanted = [list(n) for n in (_NAMES_U, _NAMES_V, _NAMES_T)]
    for i, key in enumerate(("u", "v", "t")):
        if var_names and key in var_names:
            wanted[i] = [var_names[key]]
    data = OrderedDict()
    for names in wanted:
        for f in files:
            hit = next((n for n in names if n in f.variables), None)
            if hit is None:
                continue
            var = f.variables[hit]
            dims = tuple(var.dimensions)
            coords = OrderedDict((d, np.array(f.variables[d].data)) for d in dims if d in f.variables)
            data[hit] = SimpleDataArray(var.data, dims, coords, name=hit, attrs=_attrs(var))
            break
        else:
            raise KeyError(f"none of {names} found in {paths}")
    ds = SimpleDataset(data)
    ds.attrs["_files"] = files      # keeps the memory maps valid
    return ds


def write_results(path, dataset, names=None, dtype=np.float32, attrs=None):
    """Write the variables of ``dataset`` (a SimpleDataset, e.g. what QGDataset.compute_* returns) to a NetCDF-3
    64-bit-offset file: one dimension + coordinate variable per distinct dimension name, data in ``dtype`` (float32
    like the reference's script), ``units`` attributes where known."""
    names = list(dataset.keys()) if names is None else list(names)
    with netcdf_file(path, "w", version=2) as f:
        for k, v in {**getattr(dataset, "attrs", {}), **(attrs or {})}.items():
            if isinstance(v, (str, int, float, np.integer, np.floating)) and not k.startswith("_"):
                # (SciPy stores a plain Python float as float32: hand it a float64 scalar)
                setattr(f, k, np.float64(v) if isinstance(v, float) else v)
        done = set()
        for n in names:
            da = dataset[n]
            for i, d in enumerate(da.dims):
                if d in done:
                    continue
                done.add(d)
                f.createDimension(d, int(da.shape[i]))
                c = np.asarray(da.coords[d]) if d in da.coords else np.arange(da.shape[i])
                if c.dtype.kind == "M":                       # datetime64 -> seconds since the epoch
                    c, unit = c.astype("datetime64[s]").astype(np.int64).astype(np.float64), "seconds since 1970-01-01"
                else:
                    unit = None
                if c.dtype.kind in "iu":
                    c = c.astype(np.int32)
                elif c.dtype.kind != "f":
                    c = np.arange(da.shape[i], dtype=np.int32)
                cv = f.createVariable(d, c.dtype, (d,))
                cv[:] = c
                if unit:
                    cv.units = unit
        for n in names:
            da = dataset[n]
            v = f.createVariable(n, np.dtype(dtype), tuple(da.dims))
            v[:] = np.asarray(da.data, dtype=dtype)
            if n in UNITS:
                v.units = UNITS[n]
    return path
