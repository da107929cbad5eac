"""``QGDataset`` — the batch wrapper of falwa.xarrayinterface (/root/reference/src/falwa/xarrayinterface.py:258-661) on
top of the batched device engine.

The reference builds one ``QGField`` object per (time, member, ...) combination and loops over them in Python
(:376-410, :454-455, :512-513, :546-548).  Here the flattened leading index is the batch axis of ``engine.QGBatch``:
snapshots are processed in chunks on one GPU and — when the job runs one process per GPU — the flattened index is
split into contiguous shards, one per rank, with no communication on the data path and a single gather of the
outputs at the end.

xarray is optional (it is absent from the build image).  With xarray installed the constructor takes and the methods
return ``xarray`` objects exactly like the reference; without it the same classes accept / return ``SimpleDataArray``
/ ``SimpleDataset`` (dims + coords + NumPy data), which is also what the tests use.
"""
from collections import OrderedDict

import numpy as np

_trapezoid = getattr(np, "trapezoid", None) or np.trapz   # NumPy >= 2.0 / 1.x


try:  # pragma: no cover - not available in the build image
    import xarray as xr
except Exception:  # noqa: BLE001
    xr = None

from . import __version__

_NAMES_PLEV = ["plev", "lev", "level", "isobaricInhPa", "pressure_level"]
_NAMES_YLAT = ["ylat", "lat", "latitude"]
_NAMES_XLON = ["xlon", "lon", "longitude"]
_NAMES_U, _NAMES_V, _NAMES_T = ["u", "U"], ["v", "V"], ["t", "T"]

# variable registry (xarrayinterface.py:207-238): core dims and the template attribute holding each coordinate
_VARS = {}
for _v in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"):
    _VARS[_v] = (("height", "ylat", "xlon"), ("height", "ylat", "xlon"))
for _v in ("static_stability", "static_stability_n", "static_stability_s"):
    _VARS[_v] = (("height",), ("height",))
for _v in ("qref", "uref", "ptref"):
    _VARS[_v] = (("height", "ylat"), ("height", "ylat_ref_states"))
for _v in ("u_baro", "lwa_baro", "ncforce_baro", "adv_flux_f1", "adv_flux_f2", "adv_flux_f3",
           "convergence_zonal_advective_flux", "divergence_eddy_momentum_flux", "meridional_heat_flux",
           "flux_vector_lambda_baro", "flux_vector_phi_baro"):
    _VARS[_v] = (("ylat", "xlon"), ("ylat_ref_states", "xlon"))
for _v in ("lwa",):
    _VARS[_v] = (("height", "ylat", "xlon"), ("height", "ylat_ref_states", "xlon"))

STAGE_VARS = {
    "interp": ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"),
    "ref": ("qref", "uref", "ptref"),
    "flux": ("adv_flux_f1", "adv_flux_f2", "adv_flux_f3", "convergence_zonal_advective_flux",
             "divergence_eddy_momentum_flux", "meridional_heat_flux", "lwa_baro", "u_baro", "ncforce_baro",
             "flux_vector_lambda_baro", "flux_vector_phi_baro", "lwa"),
}


class SimpleDataArray:
    """Minimal stand-in for xarray.DataArray: ``data`` (NumPy), ``dims`` (names), ``coords`` (name -> 1-D array)."""

    def __init__(self, data, dims, coords=None, name=None, attrs=None):
        self.data = np.asarray(data)
        self.dims = tuple(dims)
        assert self.data.ndim == len(self.dims)
        self.coords = OrderedDict((k, np.asarray(v)) for k, v in (coords or {}).items())
        self.name, self.attrs = name, dict(attrs or {})

    values = property(lambda self: self.data)
    shape = property(lambda self: self.data.shape)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)


class SimpleDataset(OrderedDict):
    """name -> SimpleDataArray, with ``attrs`` (stand-in for xarray.Dataset)."""

    def __init__(self, data_vars=None, attrs=None):
        super().__init__(data_vars or {})
        self.attrs = dict(attrs or {})


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of the flattened snapshot index owned by ``rank`` (blocks differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _find(names, available, user_names, what):
    if user_names and names[0] in user_names:
        if user_names[names[0]] not in available:
            raise KeyError(f"specified variable '{user_names[names[0]]}' not found")
        return user_names[names[0]]
    for n in names:
        if n in available:
            return n
    raise KeyError(f"no matching variable for '{what}' found")


for _v in ("ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term"):
    _VARS[_v] = (("height", "ylat", "xlon"), ("height", "ylat_ref_states", "xlon"))


_DEVICE_DTYPES = (np.dtype(np.float64), np.dtype(np.float32), np.dtype(np.int16))


def _packing_of(da):
    """(scale_factor, add_offset[, fill_value]) of a variable that was opened WITHOUT CF decoding (the attributes are
    still in ``attrs``; a decoded xarray variable keeps them in ``encoding`` instead), else None."""
    attrs = getattr(da, "attrs", None) or {}
    if "scale_factor" not in attrs and "add_offset" not in attrs:
        return None
    pk = (float(attrs.get("scale_factor", 1.0)), float(attrs.get("add_offset", 0.0)))
    fill = attrs.get("_FillValue", attrs.get("missing_value"))
    return pk + (float(fill),) if fill is not None else pk


def _decode_host(raw, packing, flip_lat, flip_lev):
    """xarray's CF decoding and the reference's np.flip (xarrayinterface.py:384-405) for a few snapshots on the host
    (the template field, the per-field fallbacks); chunks of the batched path are decoded on the device instead."""
    x = np.asarray(raw).astype(np.float64)
    if packing is not None:
        x = x * np.float64(packing[0])
        x = x + np.float64(packing[1])
        if len(packing) > 2:
            x[np.asarray(raw).astype(np.float64) == packing[2]] = np.nan
    if flip_lat:
        x = np.flip(x, axis=-2)
    if flip_lev:
        x = np.flip(x, axis=-3)
    return x


class _LazyField:
    """Snapshots [n][level][lat][lon] as they lie in the file (``raw``: float64 / float32 / packed int16, any axis
    order of latitude and level) that read like the decoded, reordered float64 array: ``x[i]`` and ``x[lo:hi]`` decode on
    the host; the device runner ships ``raw`` and decodes there (engine.ingest_field)."""

    def __init__(self, raw, packing=None, flip_lat=False, flip_lev=False):
        self.raw, self.packing, self.flip_lat, self.flip_lev = raw, packing, bool(flip_lat), bool(flip_lev)

    @property
    def byteswap(self):
        """stored in the other byte order (NetCDF-3 files are big-endian): swapped on the device"""
        return not self.raw.dtype.isnative

    @property
    def raw_native(self):
        """the same bytes viewed as the native type (what torch and the pinned staging buffers can hold)"""
        return self.raw.view(self.raw.dtype.newbyteorder("=")) if self.byteswap else self.raw

    shape = property(lambda self: self.raw.shape)
    ndim = property(lambda self: self.raw.ndim)
    dtype = np.dtype(np.float64)

    def __len__(self):
        return self.raw.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return _LazyField(self.raw[idx], self.packing, self.flip_lat, self.flip_lev)
        return _decode_host(self.raw[idx], self.packing, self.flip_lat, self.flip_lev)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(_decode_host(self.raw, self.packing, self.flip_lat, self.flip_lev), dtype=dtype)


class _FieldView:
    """Per-snapshot read-only view (``QGDataset.fields[i].lwa_baro`` ...), for code that iterates over fields."""

    def __init__(self, qgds, idx):
        self._q, self._i = qgds, idx

    def __getattr__(self, name):
        store = self._q._store
        if name in store:
            return store[name][self._i]
        return getattr(self._q._template, name)


class QGDataset:
    """Batch of snapshots with the interface of falwa.xarrayinterface.QGDataset.

    Extra keyword arguments (all optional): ``chunk`` — snapshots per device batch; ``shard=(rank, world)`` — process
    only this rank's contiguous block of the flattened leading index (defaults to torch.distributed's rank / world size
    when a process group is initialised); ``keep_3d`` — copy the 3-D outputs (qgpv, interpolated fields, lwa) to the
    host (default True, like the reference); ``runner`` — the object that runs a chunk (tests inject a CPU runner).
    """

    def __init__(self, da_u, da_v=None, da_t=None, *, var_names=None, qgfield=None, qgfield_args=None,
                 qgfield_kwargs=None, chunk=None, shard=None, keep_3d=True, runner=None):
        from .oopinterface import QGFieldNH18
        var_names = var_names or {}
        self._qgfield = qgfield or QGFieldNH18
        self._qgfield_args = list(qgfield_args or [])
        self._qgfield_kwargs = dict(qgfield_kwargs or {})
        u, v, t, dims, coords, packing = self._extract(da_u, da_v, da_t, var_names)
        name_p = _find(_NAMES_PLEV, coords, var_names, "plev")
        name_y = _find(_NAMES_YLAT, coords, var_names, "ylat")
        name_x = _find(_NAMES_XLON, coords, var_names, "xlon")
        assert dims[-3] == name_p, f"dimension -3 of input fields must be '{name_p}' (plev)"
        assert dims[-2] == name_y, f"dimension -2 of input fields must be '{name_y}' (ylat)"
        assert dims[-1] == name_x, f"dimension -1 of input fields must be '{name_x}' (xlon)"
        assert u.shape == v.shape == t.shape, "dimensions of the u, v, t fields don't match"
        self._other_dims = tuple(dims[:-3])
        self._other_coords = OrderedDict((d, np.asarray(coords[d]) if d in coords else np.arange(u.shape[i]))
                                         for i, d in enumerate(self._other_dims))
        self._other_shape = tuple(u.shape[:-3])
        n = int(np.prod(self._other_shape, dtype=np.int64))
        assert n > 0, "empty input"
        shape = (n,) + tuple(u.shape[-3:])
        u, v, t = u.reshape(shape), v.reshape(shape), t.reshape(shape)
        ylat, plev, xlon = np.asarray(coords[name_y]), np.asarray(coords[name_p]), np.asarray(coords[name_x])
        # xarrayinterface.py:386-395: latitude ascending, pressure descending.  The fields themselves stay as they lie
        # in the file (packed, unflipped): they are decoded and reordered on the device, chunk by chunk
        flip_lat = not np.all(np.diff(ylat) > 0)
        flip_lev = not np.all(np.diff(plev) < 0)
        if flip_lat:
            ylat = np.flip(ylat)
        if flip_lev:
            plev = np.flip(plev)
        self._n = n
        if shard is None:
            shard = (0, 1)
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    shard = (dist.get_rank(), dist.get_world_size())
            except Exception:  # noqa: BLE001
                pass
        self._rank, self._world = shard
        self._lo, self._hi = shard_bounds(n, *shard)
        self._u, self._v, self._t = (_LazyField(x[self._lo:self._hi], pk, flip_lat, flip_lev)
                                     for x, pk in zip((u, v, t), packing))
        self._xlon, self._ylat, self._plev = xlon, ylat, plev
        self._keep_3d = keep_3d
        self._chunk = chunk
        self._runner = runner
        self._template = None
        self._store = {}      # name -> host array [n_local, ...]
        self._done = set()
        self._global = {}     # name -> gathered array [n, ...] (every rank's shard, in order); see gather()
        if self._hi > self._lo or runner is None:
            self._make_template()

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _extract(da_u, da_v, da_t, var_names):
        """-> u, v, t (NumPy, same shape, storage dtype), dims, coords, packing from xarray objects or the Simple*
        stand-ins."""
        def is_ds(x):
            return (xr is not None and isinstance(x, xr.Dataset)) or isinstance(x, SimpleDataset)

        def pick(ds, names, what):
            return ds[_find(names, list(ds.keys()) if not (xr is not None and isinstance(ds, xr.Dataset)) else ds,
                            var_names, what)]
        if is_ds(da_u):
            src = da_u
            if da_v is None:
                da_v = pick(src, _NAMES_V, "v")
            elif is_ds(da_v):
                da_v = pick(da_v, _NAMES_V, "v")
            if da_t is None:
                da_t = pick(src, _NAMES_T, "t")
            elif is_ds(da_t):
                da_t = pick(da_t, _NAMES_T, "t")
            da_u = pick(src, _NAMES_U, "u")
        assert da_u is not None, "missing u field"
        assert da_v is not None, "missing v field"
        assert da_t is not None, "missing t field"
        assert tuple(da_u.dims) == tuple(da_v.dims), "dimensions of fields u and v don't match"
        assert tuple(da_u.dims) == tuple(da_t.dims), "dimensions of fields u and t don't match"
        coords = {k: np.asarray(getattr(c, "values", c)) for k, c in da_u.coords.items()}
        # the reference merges the three arrays with join="exact" (xarrayinterface.py:358): coordinates must agree
        for other in (da_v, da_t):
            for k, c in other.coords.items():
                if k in coords and k in da_u.dims:
                    c = np.asarray(getattr(c, "values", c))
                    if c.shape != coords[k].shape or not np.array_equal(c, coords[k]):
                        raise ValueError(f"coordinate '{k}' of the u, v, t fields differs (an exact match is required)")
        fields, packing = [], []
        for da in (da_u, da_v, da_t):
            a, pk = np.asarray(da.data), _packing_of(da)
            if a.dtype.newbyteorder("=") not in _DEVICE_DTYPES:      # other storage types: decode here, ship float64
                a, pk = _decode_host(a, pk, False, False), None
            fields.append(a)
            packing.append(pk)
        return (*fields, tuple(da_u.dims), coords, packing)

    def _make_template(self):
        if self._template is None:
            i = 0 if self._hi > self._lo else None
            assert i is not None or self._runner is not None
            self._template = self._qgfield(self._xlon, self._ylat, self._plev, self._u[0], self._v[0], self._t[0],
                                           *self._qgfield_args, **self._qgfield_kwargs)
        return self._template

    @property
    def fields(self):
        """Per-snapshot views of this rank's shard (the reference returns its list of QGField objects)."""
        return [_FieldView(self, i) for i in range(self._hi - self._lo)]

    @property
    def shard(self):
        return self._lo, self._hi

    @property
    def attrs(self):
        f = self._template
        return {"kmax": f.kmax, "dz": f.dz, "maxit": f.maxit, "tol": f.tol, "npart": f.npart, "rjac": f.rjac,
                "scale_height": f.scale_height, "cp": f.cp, "dry_gas_constant": f.dry_gas_constant, "omega": f.omega,
                "planet_radius": f.planet_radius, "prefactor": f.prefactor, "protocol": type(f).__name__,
                "package": f"hn2016_falwa_b200 {__version__}"}

    # ---- execution ---------------------------------------------------------------------------------
    def _run(self, upto, ncforce=None):
        """Run the pipeline up to stage ``upto`` over this rank's shard, chunk by chunk, keeping the outputs of the
        stages that have not been stored yet.  Earlier stages are recomputed on the device rather than re-uploaded:
        they cost ~1 ms per 0.25-degree snapshot, a PCIe round trip of their five 3-D fields costs ~40 ms."""
        order = ("interp", "ref", "flux")
        stages = order[:order.index(upto) + 1]
        todo = [s for s in stages if s not in self._done]
        if not todo:
            return
        nloc = self._hi - self._lo
        if nloc == 0:
            self._done.update(todo)
            return
        runner = self._runner
        if runner is None:
            tpl = self._make_template()
            # inputs on an even latitude grid are regridded per field by the QGField constructor
            # (oopinterface.py:147-154): rare path, one object per snapshot like the reference
            runner = _FieldLoopRunner(self) if tpl.need_latitude_interpolation else _DeviceRunner(tpl, self._chunk)
        want = [v for s in todo for v in STAGE_VARS[s]]
        if "interp" in todo:
            want += ["static_stability"]
        if not self._keep_3d:
            want = [v for v in want if len(_VARS.get(v, ((), ()))[0]) < 3]
        for lo, hi, out in runner.run(self._u, self._v, self._t, stages, want,
                                      None if ncforce is None else ncforce[self._lo:self._hi]):
            for name, arr in out.items():
                if name not in self._store:
                    self._store[name] = np.empty((nloc,) + tuple(arr.shape[1:]), dtype=np.float64)
                if hasattr(out, "copy_into"):     # pinned staging slot: drained by the runner's host threads
                    out.copy_into(name, self._store[name][lo:hi])
                else:
                    self._store[name][lo:hi] = arr
        self._done.update(todo)

    def gather(self):
        """Concatenate every rank's shard of the stored outputs on all ranks (one all_gather per variable; the only
        communication of the whole job).  No-op for a single process.  The local shard stays what the compute methods
        work on: a stage run after a gather() computes this rank's rows as before, and the next gather() collects the
        variables that are new since (all ranks must call it, like any collective)."""
        if self._world == 1:
            return
        import torch.distributed as dist
        # agree on the variable list first (a rank with an empty shard has stored nothing)
        meta = {k: (str(a.dtype), tuple(a.shape[1:])) for k, a in self._store.items()
                if self._global.get(k) is None or self._global[k][0] is not a}
        metas = [None] * self._world
        dist.all_gather_object(metas, meta)
        union = {}
        for m in metas:
            union.update(m)
        for name in sorted(union):
            dtype, tail = union[name]
            loc = self._store.get(name)
            src = loc
            if loc is None:
                loc = np.empty((0,) + tuple(tail), dtype=dtype)
            parts = [None] * self._world
            if int(np.prod(tail, dtype=np.int64)) * max(1, loc.shape[0]) < (1 << 14):
                dist.all_gather_object(parts, (self._lo, loc))
            else:
                self._gather_big(parts, loc)
            parts = sorted((p for p in parts if p[1].shape[0] > 0), key=lambda p: p[0])
            # remember which local array the gathered copy was made from: a recomputed variable is gathered again
            self._global[name] = (src, np.concatenate([p[1] for p in parts], axis=0))

    def _gather_big(self, parts, loc):
        """Tensor all_gather (NCCL on GPUs, gloo on CPU) of equally padded shards."""
        import torch
        import torch.distributed as dist
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        nmax = max(shard_bounds(self._n, r, self._world)[1] - shard_bounds(self._n, r, self._world)[0]
                   for r in range(self._world))
        pad = np.zeros((nmax,) + loc.shape[1:], dtype=loc.dtype)
        pad[:loc.shape[0]] = loc
        mine = torch.from_numpy(pad).to(dev)
        bufs = [torch.empty_like(mine) for _ in range(self._world)]
        dist.all_gather(bufs, mine)
        for r, b in enumerate(bufs):
            lo, hi = shard_bounds(self._n, r, self._world)
            parts[r] = (lo, b[:hi - lo].cpu().numpy())

    # ---- output helpers -------------------------------------------------------------------------------
    def _coords_for(self, var):
        core, names = _VARS[var]
        coords = OrderedDict(self._other_coords)
        for dim, attr in zip(core, names):
            coords[dim] = np.asarray(getattr(self._template, attr))
        return self._other_dims + core, coords

    def _as_dataarray(self, var):
        if var not in self._store and var not in self._global:
            raise ValueError(f"{var} is not computed yet.")
        g = self._global.get(var)
        # the gathered copy, unless the variable has been recomputed locally since
        arr = g[1] if g is not None and (var not in self._store or g[0] is self._store[var]) else self._store[var]
        dims, coords = self._coords_for(var)
        if arr.shape[0] == self._n:          # full (single process or gathered): restore the leading dims
            arr = arr.reshape(self._other_shape + arr.shape[1:])
        else:                                # a shard: one flattened leading dim
            dims = ("snapshot",) + dims[len(self._other_dims):]
            coords = OrderedDict((k, v) for k, v in coords.items() if k not in self._other_dims)
            coords["snapshot"] = np.arange(self._lo, self._hi)
        if xr is not None:  # pragma: no cover
            return xr.DataArray(arr, dims=dims, coords=coords, name=var)
        return SimpleDataArray(arr, dims, coords, name=var)

    def _dataset(self, names):
        data_vars = OrderedDict((n, self._as_dataarray(n)) for n in names
                                if n in self._store or n in self._global or self._keep_3d)
        if xr is not None:  # pragma: no cover
            return xr.Dataset(data_vars, attrs=self.attrs)
        return SimpleDataset(data_vars, attrs=self.attrs)

    def __getattr__(self, name):
        if name in _VARS and not name.startswith("static_stability"):
            return self._as_dataarray(name)
        raise AttributeError(name)

    @property
    def static_stability(self):
        """Global definition (NH18): one DataArray; hemispheric (NHN22): the (north, south) pair like the reference
        (xarrayinterface.py:476-497, whose suffixes follow the order of QGFieldNHN22.static_stability)."""
        st = self._store.get("static_stability")
        g = self._global.get("static_stability")
        if g is not None and (st is None or g[0] is st):
            st = g[1]            # every rank's rows (after gather())
        if st is None:
            raise ValueError("static_stability is not computed yet.")
        if st.ndim == 2:
            self._store["static_stability"] = st
            return self._as_dataarray("static_stability")
        self._store["static_stability_n"], self._store["static_stability_s"] = st[:, 0, :], st[:, 1, :]
        return self._as_dataarray("static_stability_n"), self._as_dataarray("static_stability_s")

    # ---- the reference's methods ---------------------------------------------------------------------
    def interpolate_fields(self, return_dataset=True):
        self._run("interp")
        if return_dataset:
            ds = self._dataset(STAGE_VARS["interp"])
            stab = self.static_stability
            for s in (stab if isinstance(stab, tuple) else (stab,)):
                ds[s.name] = s
            return ds

    def compute_reference_states(self, return_dataset=True):
        self._run("ref")
        if return_dataset:
            return self._dataset(STAGE_VARS["ref"])

    def compute_lwa_and_barotropic_fluxes(self, ncforce=None, return_dataset=True):
        nc = None
        if ncforce is not None:
            nc = np.asarray(getattr(ncforce, "data", ncforce))
            nc = nc.reshape((self._n,) + nc.shape[-3:])
        self._run("flux", ncforce=nc)
        if return_dataset:
            return self._dataset(STAGE_VARS["flux"])


def _per_field(qgds):
    """One QGField object per snapshot of this rank's shard (the reference's structure), for the methods that are not
    batched on the device."""
    q = qgds
    for i in range(q._hi - q._lo):
        f = q._qgfield(q._xlon, q._ylat, q._plev, q._u[i], q._v[i], q._t[i], *q._qgfield_args, **q._qgfield_kwargs)
        f.interpolate_fields(return_named_tuple=False)
        yield i, f


def _loop_method(self, names, call):
    """Shared body of the per-field (not device-batched) QGDataset methods."""
    nloc = self._hi - self._lo
    for i, f in _per_field(self):
        call(i, f)
        for n in names:
            a = np.asarray(getattr(f, n))
            if n not in self._store or self._store[n].shape[0] != nloc:
                self._store[n] = np.empty((nloc,) + a.shape)
            self._store[n][i] = a


def _compute_lwa_only(self, return_dataset=True):
    """``compute_lwa_only`` of every field (xarrayinterface.py:566-585): LWA from QGPV and Qref alone, for snapshots whose
    NH18 reference-state solver did not converge."""
    def call(i, f):
        f._raise_error_for_nonconvergence = False
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f.compute_reference_states(return_named_tuple=False)
        f.compute_lwa_only()
    _loop_method(self, ("lwa_baro", "u_baro", "lwa"), call)
    if return_dataset:
        return self._dataset(("lwa_baro", "u_baro", "lwa"))


def _compute_layerwise(self, ncforce=None, return_dataset=True):
    """``compute_layerwise_lwa_fluxes`` of every field (xarrayinterface.py:587-630).  Batched on the device (stages 1-2
    are recomputed there, then falwa_b200_layerwise_lwa_fluxes per chunk) unless the inputs need the constructor's
    latitude regridding."""
    nc = None
    if ncforce is not None:
        nc = np.asarray(getattr(ncforce, "data", ncforce))
        nc = nc.reshape((self._n,) + nc.shape[-3:])[self._lo:self._hi]
    names = ("lwa", "ua1", "ua2", "ep1", "ep2", "ep3", "stretch_term")
    tpl = self._make_template() if self._hi > self._lo else None
    if self._runner is None and tpl is not None and not tpl.need_latitude_interpolation:
        nloc = self._hi - self._lo
        runner = _DeviceRunner(tpl, self._chunk, layerwise=True)
        for lo, hi, out in runner.run(self._u, self._v, self._t, ("interp", "ref", "layerwise"), list(names), nc):
            for name, arr in out.items():
                if name not in self._store or self._store[name].shape[0] != nloc:
                    self._store[name] = np.empty((nloc,) + tuple(arr.shape[1:]), dtype=np.float64)
                out.copy_into(name, self._store[name][lo:hi])
        if return_dataset:
            return self._dataset(names)
        return None

    def call(i, f):
        f.compute_reference_states(return_named_tuple=False)
        f.compute_layerwise_lwa_fluxes(ncforce=None if nc is None else nc[i])
    _loop_method(self, names, call)
    if return_dataset:
        return self._dataset(names)


def _compute_ncforce(self, heating_rate):
    """``compute_ncforce_from_heating_rate`` of every field (xarrayinterface.py:632-661); returns (n, kmax, nlat, nlon)
    with the leading dimensions restored when the whole batch is local."""
    hr = np.asarray(getattr(heating_rate, "data", heating_rate))
    hr = hr.reshape((self._n,) + hr.shape[-3:])[self._lo:self._hi]
    tpl = self._make_template() if self._hi > self._lo else None
    if self._runner is None and tpl is not None and not tpl.need_latitude_interpolation:
        # device-batched: only the temperature is needed (static stability), then one kernel per chunk
        import torch
        from . import engine
        nloc = self._hi - self._lo
        chunk = self._chunk or _DeviceRunner(tpl).chunk
        arr = np.empty((nloc, tpl.kmax) + hr.shape[-2:])
        on_h = tpl._data_on_evenly_spaced_pseudoheight_grid
        for lo in range(0, nloc, chunk):
            hi = min(nloc, lo + chunk)
            x = self._t[lo:hi]
            raw = torch.from_numpy(np.ascontiguousarray(x.raw_native)).cuda()
            t = engine.ingest_field(raw, x.packing, x.flip_lat, x.flip_lev, byteswap=x.byteswap) \
                if engine.needs_ingest(raw.dtype, x.packing, x.flip_lat, x.flip_lev, x.byteswap) else raw
            b = engine.QGBatch(tpl._plan, t, t, t, on_height_grid=on_h)
            b.compute_profiles()
            arr[lo:hi] = b.ncforce_from_heating_rate(np.ascontiguousarray(hr[lo:hi], dtype=np.float64)).cpu().numpy()
    else:
        arr = np.asarray([f.compute_ncforce_from_heating_rate(hr[i]) for i, f in _per_field(self)])
    return arr.reshape(self._other_shape + arr.shape[1:]) if arr.shape[0] == self._n else arr


QGDataset.compute_lwa_only = _compute_lwa_only
QGDataset.compute_layerwise_lwa_fluxes = _compute_layerwise
QGDataset.compute_ncforce_from_heating_rate = _compute_ncforce


class _DeviceRunner:
    """Runs chunks of snapshots through ``engine.QGBatch`` on the current CUDA device and yields host copies of the
    requested variables.  User arrays are pageable NumPy (or lazily loaded) memory, so both directions go through
    two pinned staging slots: a few host threads fill / drain them (multi-threaded memcpy) while the GPU works on the
    neighbouring chunk, and the PCIe copies run on side streams."""

    def __init__(self, template, chunk=None, layerwise=False):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        from . import engine
        self.torch, self.engine = torch, engine
        self.f = template
        self.plan = template._plan
        if chunk is None:
            # two input slots + decoded copies, and two batches in flight (stage 1-2 of one, stage 3 of the other)
            per = 8 * self.plan.nlon * self.plan.nlat * (12 * self.plan.nlev + (40 if layerwise else 20) * self.plan.kmax
                                                         + 40)
            free = torch.cuda.mem_get_info()[0]
            chunk = int(max(1, min(256, 0.4 * free // per)))
            # ... and at most ~8 GB of page-locked input staging (two slots x three fields)
            stage = 2 * 3 * 8 * self.plan.nlon * self.plan.nlat * self.plan.nlev
            chunk = int(max(1, min(chunk, 8e9 // stage)))
        self.chunk = chunk
        self.pool = ThreadPoolExecutor(max_workers=6)

    def _threaded_copy(self, dst, src):
        """dst[...] = src with the leading axis split over the pool (np.copyto releases the GIL)."""
        n = dst.shape[0]
        step = max(1, -(-n // 6))
        return [self.pool.submit(np.copyto, dst[i:i + step], src[i:i + step]) for i in range(0, n, step)]

    def run(self, u, v, t, stages, want, ncforce=None):
        torch, engine = self.torch, self.engine
        f = self.f
        n = u.shape[0]
        nh_only = f.northern_hemisphere_results_only
        on_h = f._data_on_evenly_spaced_pseudoheight_grid
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        main = torch.cuda.current_stream()
        # the arrays as they lie in the file (float64 / float32 / packed int16, file order) cross PCIe unchanged; CF
        # decoding and the latitude / level flips happen on the device (engine.ingest_field)
        lazy = [x if isinstance(x, _LazyField) else _LazyField(np.asarray(x, dtype=np.float64)) for x in (u, v, t)]
        u, v, t = (x.raw_native for x in lazy)
        tdt = [torch.from_numpy(np.empty(0, dtype=x.dtype)).dtype for x in (u, v, t)]
        ingest = [engine.needs_ingest(dt, x.packing, x.flip_lat, x.flip_lev, x.byteswap) for dt, x in zip(tdt, lazy)]
        chunk = max(1, min(self.chunk, n))       # never stage more snapshots than there are
        shape_in = (chunk,) + tuple(u.shape[1:])
        pin_in = [[torch.empty(shape_in, dtype=dt).pin_memory() for dt in tdt] for _ in range(2)]
        dev_in = [[torch.empty(shape_in, dtype=dt, device="cuda") for dt in tdt] for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        for e in ev_free:
            e.record(main)
        bounds = [(lo, min(n, lo + chunk)) for lo in range(0, n, chunk)]

        def stage(c):   # host threads: pageable user arrays -> pinned slot
            lo, hi = bounds[c]
            futs = []
            for dst, src in zip(pin_in[c & 1], (u, v, t)):
                futs += self._threaded_copy(dst.numpy()[:hi - lo], src[lo:hi])
            return futs

        def upload(c, futs):
            lo, hi = bounds[c]
            for fu in futs:
                fu.result()
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[c & 1])
                for d, p in zip(dev_in[c & 1], pin_in[c & 1]):
                    d[:hi - lo].copy_(p[:hi - lo], non_blocking=True)
                ev_in[c & 1].record(s_in)

        eq = f.equator_idx
        pin_out, out_ev, drain = [dict(), dict()], [None, None], [[], []]
        upload(0, stage(0))
        nxt = stage(1) if len(bounds) > 1 else None

        def finish(c, b):
            """Flux stage and copy-out of chunk c -> (lo, hi, _HostChunk).  Runs after the hemispheric means of chunk
            c+1 are queued, so the host fits the splines of c+1 while the device works on this."""
            lo, hi = bounds[c]
            slot = c & 1
            if "flux" in stages:
                b.compute_lwa_and_barotropic_fluxes(ncforce=None if ncforce is None else ncforce[lo:hi])
            layer = None
            if "layerwise" in stages:
                layer = b.compute_layerwise_lwa_fluxes(ncforce=None if ncforce is None else ncforce[lo:hi])
            done = torch.cuda.Event()
            done.record(main)
            for fu in drain[slot]:
                fu.result()          # the pinned slot's previous contents have been copied out
            drain[slot] = []
            host = {}
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for name in want:
                    if name == "static_stability":
                        continue
                    if layer is not None:     # layer-wise store (its lwa is astar1 + astar2, oopinterface.py:1099-1100)
                        src = layer[:, engine.LAYER3D[name]]
                    else:
                        src = {"qgpv": b.qgpv, "interpolated_u": b.u, "interpolated_v": b.v,
                               "interpolated_theta": b.theta, "qref": b.qref_cu, "uref": b.uref, "ptref": b.ptref,
                               "lwa": b.lwa}.get(name)
                        if src is None:
                            src = b.out(name)
                    if nh_only and name not in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"):
                        src = src[..., -eq:, :] if name in engine.OUT2D or name == "lwa" or layer is not None \
                            else src[..., -eq:]
                    src = src.contiguous()
                    src.record_stream(s_out)
                    buf = pin_out[slot].get(name)
                    if buf is None or buf.shape[1:] != src.shape[1:]:
                        buf = pin_out[slot][name] = torch.empty((chunk,) + tuple(src.shape[1:]),
                                                                dtype=src.dtype).pin_memory()
                    buf[:hi - lo].copy_(src, non_blocking=True)
                    host[name] = buf[:hi - lo]
                out_ev[slot] = torch.cuda.Event()
                out_ev[slot].record(s_out)
            if "static_stability" in want:
                p = b.profiles
                ss = p[:, 2].copy() if self.plan.kind == 0 else np.stack([p[:, 2], p[:, 3]], axis=1)
                if self.plan.kind == 1 and nh_only:
                    ss = p[:, 3].copy()
                host["static_stability"] = ss
            return lo, hi, _HostChunk(host, out_ev[slot], self, drain[slot])

        prev = pending = None
        for c, (lo, hi) in enumerate(bounds):
            slot = c & 1
            main.wait_event(ev_in[slot])
            fields = [x[:hi - lo] for x in dev_in[slot]]
            for i, x in enumerate(lazy):
                if ingest[i] or on_h:    # (data on the height grid is read by all stages: it must leave the slot)
                    fields[i] = engine.ingest_field(fields[i], x.packing, x.flip_lat, x.flip_lev, byteswap=x.byteswap)
            b = engine.QGBatch(self.plan, *fields, on_height_grid=on_h, northern_hemisphere_results_only=nh_only)
            b.launch_theta_means()
            if prev is not None:
                item = finish(*prev)
                # the chunk BEFORE that one, whose copy has had a whole chunk of compute to finish, is handed over
                if pending is not None:
                    yield pending
                pending = item
            b.interpolate_fields()      # waits for the means only; the device is busy with the previous flux stage
            ev_free[slot].record(main)  # stage 1 was the last reader of the input slot
            if c + 1 < len(bounds):      # the pinned slot of chunk c+1 is filled by now (or we wait here), send it
                upload(c + 1, nxt)
                nxt = stage(c + 2) if c + 2 < len(bounds) else None
            if "ref" in stages:
                b.compute_reference_states()
                if self.plan.kind == 0 and (b.num_of_iter >= f.maxit).any():
                    msg = "The reference state does not converge for at least one snapshot of the batch."
                    if f._raise_error_for_nonconvergence:
                        raise ValueError(msg)
                    import warnings
                    warnings.warn(msg)
            prev = (c, b)
        if prev is not None:
            item = finish(*prev)
            if pending is not None:
                yield pending
            yield item
        for d in drain:
            for fu in d:
                fu.result()


class _HostChunk(dict):
    """Mapping name -> array of one chunk living in a pinned staging slot.  ``copy_into`` drains it into the caller's
    store with the runner's host threads (and registers the futures that guard the slot's reuse)."""

    def __init__(self, host, event, runner, drain):
        super().__init__()
        self._host, self._event, self._runner, self._drain = host, event, runner, drain
        for k, v in host.items():
            self[k] = v

    def items(self):
        self._event.synchronize()
        for k, v in self._host.items():
            yield k, (v.numpy() if hasattr(v, "numpy") and not isinstance(v, np.ndarray) else v)

    def copy_into(self, name, dst):
        self._event.synchronize()
        v = self._host[name]
        src = v.numpy() if not isinstance(v, np.ndarray) else v
        self._drain.extend(self._runner._threaded_copy(dst, src))


class _FieldLoopRunner:
    """One QGField object per snapshot (the reference's own structure); used when the inputs need the constructor's
    latitude regridding."""

    def __init__(self, qgds):
        self.q = qgds

    def run(self, u, v, t, stages, want, ncforce=None):
        q = self.q
        for i in range(u.shape[0]):
            f = q._qgfield(q._xlon, q._ylat, q._plev, u[i], v[i], t[i], *q._qgfield_args, **q._qgfield_kwargs)
            f.interpolate_fields(return_named_tuple=False)
            if "ref" in stages:
                f.compute_reference_states(return_named_tuple=False)
            if "flux" in stages:
                f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False,
                                                    ncforce=None if ncforce is None else ncforce[i])
            out = {}
            for name in want:
                val = getattr(f, name)
                out[name] = np.asarray(val)[None] if not isinstance(val, tuple) else np.stack(val)[None]
            yield i, i + 1, out


# ---- the steps after the path (falwa.xarrayinterface.integrate_budget / hemisphere_to_globe, :664-791) --------------
_NAMES_TIME = ["time", "date", "datetime", "valid_time"]


def _seconds(t):
    t = np.asarray(t)
    if np.issubdtype(t.dtype, np.datetime64):
        return (t - t[0]) / np.timedelta64(1, "s")
    return t.astype(np.float64)


def integrate_budget(ds, var_names=None):
    """Time-integrated LWA budget terms of equation (2) of NH18 (xarrayinterface.py:664-735): the change of ``lwa_baro``
    over the interval, the trapezoidal time integrals of the three tendency terms and the residual.  Works on the
    datasets returned by ``QGDataset.compute_lwa_and_barotropic_fluxes`` (xarray when installed, ``SimpleDataset``
    otherwise).  A non-datetime time coordinate is taken to be in seconds."""
    if xr is not None and isinstance(ds, xr.Dataset):  # pragma: no cover - same calls as the reference
        names = {k: _find(v, ds, var_names, v[0]) for k, v in dict(
            time=_NAMES_TIME, lwa=["lwa_baro"], czaf=["convergence_zonal_advective_flux"],
            demf=["divergence_eddy_momentum_flux"], mhf=["meridional_heat_flux"]).items()}
        start, stop = ds[names["time"]].values[0], ds[names["time"]].values[-1]
        dlwa = ds[names["lwa"]].sel({names["time"]: stop}) - ds[names["lwa"]].sel({names["time"]: start})
        terms = {k: ds[names[k]].integrate(coord=names["time"], datetime_unit="s") for k in ("czaf", "demf", "mhf")}
        res = dlwa - terms["czaf"] - terms["demf"] - terms["mhf"]
        attrs = dict(ds.attrs, integration_start=str(start), integration_stop=str(stop),
                     integration_seconds=(stop - start) / np.timedelta64(1000000000))
        return xr.Dataset({"delta_lwa": dlwa, "integrated_convergence_zonal_advective_flux": terms["czaf"],
                           "integrated_divergence_eddy_momentum_flux": terms["demf"],
                           "integrated_meridional_heat_flux": terms["mhf"], "residual": res}, ds.coords, attrs)
    lwa = ds[_find(["lwa_baro"], ds, var_names, "lwa_baro")]
    tname = _find(_NAMES_TIME, lwa.coords, var_names, "time")
    axis = lwa.dims.index(tname)
    tsec = _seconds(lwa.coords[tname])
    rest = tuple(d for d in lwa.dims if d != tname)
    coords = OrderedDict((k, v) for k, v in lwa.coords.items() if k != tname)

    def integ(name):
        return _trapezoid(np.asarray(ds[_find([name], ds, var_names, name)].data), x=tsec, axis=axis)
    dl = np.take(lwa.data, -1, axis=axis) - np.take(lwa.data, 0, axis=axis)
    czaf, demf, mhf = (integ(n) for n in ("convergence_zonal_advective_flux", "divergence_eddy_momentum_flux",
                                          "meridional_heat_flux"))
    out = {"delta_lwa": dl, "integrated_convergence_zonal_advective_flux": czaf,
           "integrated_divergence_eddy_momentum_flux": demf, "integrated_meridional_heat_flux": mhf,
           "residual": dl - czaf - demf - mhf}
    attrs = dict(ds.attrs, integration_start=str(lwa.coords[tname][0]), integration_stop=str(lwa.coords[tname][-1]),
                 integration_seconds=float(tsec[-1] - tsec[0]))
    return SimpleDataset(OrderedDict((k, SimpleDataArray(v, rest, coords, name=k)) for k, v in out.items()), attrs)


def hemisphere_to_globe(ds, var_names=None):
    """Mirror a hemispheric dataset (which must start or end at the equator) to the other hemisphere, negating the
    meridional wind on the created half (xarrayinterface.py:738-791)."""
    if xr is not None and isinstance(ds, xr.Dataset):  # pragma: no cover - same calls as the reference
        yname = _find(_NAMES_YLAT, ds, var_names, "ylat")
        eq0 = abs(float(ds[yname][0])) < 1e-4
        assert eq0 or abs(float(ds[yname][-1])) < 1e-4, "equator not found on the hemisphere"
        sd = ds.reindex({yname: ds[yname][slice(None, 0, -1) if eq0 else slice(-2, None, -1)]})
        sd[yname] = -sd[yname]
        try:
            vname = _find(_NAMES_V, ds, var_names, "v")
            sd[vname] = -sd[vname]
        except KeyError:
            pass
        return xr.concat([sd, ds] if eq0 else [ds, sd], dim=yname)
    first = next(iter(ds.values()))
    yname = _find(_NAMES_YLAT, first.coords, var_names, "ylat")
    ylat = np.asarray(first.coords[yname])
    eq0 = abs(ylat[0]) < 1.0e-4
    assert eq0 or abs(ylat[-1]) < 1.0e-4, \
        "equator not found on the hemisphere; make sure latitudes either begin or end with 0° latitude"
    sel = slice(None, 0, -1) if eq0 else slice(-2, None, -1)
    try:
        vname = _find(_NAMES_V, ds, var_names, "v")
    except KeyError:
        vname = None
    out = SimpleDataset(attrs=ds.attrs)
    for name, da in ds.items():
        if yname not in da.dims:     # e.g. static_stability: carried through unchanged, as xarray does
            out[name] = da
            continue
        ax = da.dims.index(yname)
        idx = [slice(None)] * da.data.ndim
        idx[ax] = sel
        mirrored = da.data[tuple(idx)]
        if name == vname:
            mirrored = -mirrored
        data = np.concatenate([mirrored, da.data] if eq0 else [da.data, mirrored], axis=ax)
        coords = OrderedDict(da.coords)
        coords[yname] = np.concatenate([-ylat[sel], ylat] if eq0 else [ylat, -ylat[sel]])
        out[name] = SimpleDataArray(data, da.dims, coords, name=name, attrs=da.attrs)
    return out
