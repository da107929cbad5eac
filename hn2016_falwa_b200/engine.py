"""Batched, device-resident QGField engine: the host side of Section B of include/falwa_b200.h.

One ``QGBatch`` holds B independent snapshots on one GPU (fields [B][level][lat][lon], fp64) and runs the
three stages of the reference pipeline on all of them per call:

    interpolate_fields  ->  compute_reference_states  ->  compute_lwa_and_barotropic_fluxes
    (oopinterface.py:467-535)  (:550-615, :1604-1674, :1774-1888)   (:716-982)

``QGFieldNH18`` / ``QGFieldNHN22`` (oopinterface.py of this package) are a batch of one; ``QGDataset`` shards
its snapshots over batches (and over GPUs, one process per GPU).  PyTorch is used for device memory, streams
and host<->device copies only; every number is produced by the kernels of libfalwa_b200.so, except the
static-stability smoothing spline (SciPy FITPACK on 2 x nlev numbers per snapshot), which is the
same host call the reference makes.
"""
import ctypes
import math

import numpy as np
import torch
from scipy.interpolate import UnivariateSpline

from . import _lib
from ._lib import c_d, c_dp, c_i, check, dptr, hptr, stream_ptr
from .constant import P_GROUND

OUT2D = {  # slot order of falwa_b200.h (enum FALWA_OUT2D_*)
    "astar1_baro": 0, "astar2_baro": 1, "lwa_baro": 2, "adv_flux_f1": 3, "adv_flux_f2": 4, "adv_flux_f3": 5,
    "ep2baro": 6, "ep3baro": 7, "meridional_heat_flux": 8, "ncforce_baro": 9, "u_baro": 10,
    "convergence_zonal_advective_flux": 11, "divergence_eddy_momentum_flux": 12, "flux_vector_lambda_baro": 13,
    "flux_vector_phi_baro": 14}
N_OUT2D = 15
LAYER3D = {"astar1": 0, "astar2": 1, "ncforce": 2, "ua1": 3, "ua2": 4, "ep1": 5, "ep2": 6, "ep3": 7, "stretch_term": 8,
           "lwa": 9}
FUSE_ZONAL_STATS = False

_registered = False


def _register():
    global _registered
    if _registered:
        return
    _lib.register("falwa_b200_plan_create", [c_dp] + [c_i] * 8 + [c_dp] * 4 + [c_d] * 9)
    _lib.register("falwa_b200_plan_destroy", [c_dp])
    _lib.register("falwa_b200_theta_hemispheric_means", [c_dp, c_i, c_dp, c_dp, c_dp])
    _lib.register("falwa_b200_interpolate_fields", [c_dp, c_i, c_dp, c_dp, c_dp, c_i, c_dp] + [c_dp] * 7 + [c_dp])
    _lib.register("falwa_b200_reference_states", [c_dp, c_i] + [c_dp] * 5 + [c_i] + [c_dp] * 4 + [c_dp] * 4)
    _lib.register("falwa_b200_lwa_and_barotropic_fluxes",
                  [c_dp, c_i] + [c_dp] * 9 + [c_i, c_dp, c_dp, c_i, c_dp])
    _lib.register("falwa_b200_layerwise_lwa_fluxes", [c_dp, c_i] + [c_dp] * 9 + [c_i, c_dp, c_i, c_dp])
    _lib.register("falwa_b200_ncforce_from_heating_rate", [c_dp, c_i, c_dp, c_i, c_dp, c_dp, c_dp])
    _registered = True


_SRC_TYPES = {torch.float64: 0, torch.float32: 1, torch.int16: 2}


def needs_ingest(dtype, packing=None, flip_lat=False, flip_lev=False, byteswap=False):
    """False when the raw array already is what the pipeline reads (native float64, ordered, not packed)."""
    return dtype != torch.float64 or packing is not None or flip_lat or flip_lev or byteswap


def ingest_field(raw, packing=None, flip_lat=False, flip_lev=False, out=None, byteswap=False):
    """Archived values on the device -> the pipeline's float64 [B][level][lat][lon], latitude ascending, pressure
    descending (Section D of falwa_b200.h).  ``raw``: CUDA tensor of float64 / float32 / int16 as it lies in the file;
    ``packing`` = (scale_factor, add_offset[, fill_value]) of CF-packed variables, value = raw * scale + offset and
    fill -> NaN, like xarray's decoder; ``flip_lat`` / ``flip_lev`` reverse those axes (the np.flip of
    xarrayinterface.py:384-405); ``byteswap``: the values are stored big-endian (NetCDF-3 files) and ``raw`` is the
    same bytes viewed as the native type."""
    assert raw.is_cuda and raw.dim() == 4 and raw.dtype in _SRC_TYPES, "ingest_field: [B][lev][lat][lon] f64/f32/i16"
    raw = raw.contiguous()
    if out is None:
        out = torch.empty(raw.shape, dtype=torch.float64, device=raw.device)
    assert out.dtype == torch.float64 and out.is_contiguous() and tuple(out.shape) == tuple(raw.shape)
    assert out.data_ptr() != raw.data_ptr(), "ingest_field works out of place"
    scale, offset, fill = 1.0, 0.0, None
    if packing is not None:
        scale, offset = float(packing[0]), float(packing[1])
        fill = packing[2] if len(packing) > 2 else None
    B, nl, ny, nx = (int(x) for x in raw.shape)
    check(_lib.load().falwa_b200_ingest_field(
        _SRC_TYPES[raw.dtype], B, nl, ny, nx, ctypes.c_void_p(raw.data_ptr()), int(bool(byteswap)), scale, offset,
        int(fill is not None), float(fill) if fill is not None else 0.0, int(bool(flip_lat)), int(bool(flip_lev)),
        dptr(out), stream_ptr()), "ingest_field")
    return out


def to_host(t):
    """Device tensor -> NumPy array at PCIe speed: the copy lands in page-locked memory from torch's caching host
    allocator (a fresh pageable array of a 0.25-degree 3-D field costs ~200 ms of page faults, this ~8 ms); the array
    keeps the block alive and hands it back to the cache when it is garbage-collected."""
    t = t.contiguous()
    try:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:          # page-locked memory exhausted: ordinary memory
        host = torch.empty(t.shape, dtype=t.dtype)
    host.copy_(t)
    return host.numpy()


_upload_pool = None


def to_device(a):
    """Contiguous NumPy array -> CUDA tensor through page-locked staging filled by a few host threads (np.copyto
    releases the GIL), slice by slice, each slice sent as soon as it is staged."""
    global _upload_pool
    if a.nbytes < (32 << 20) or a.ndim < 2:
        return torch.from_numpy(a).pin_memory().cuda(non_blocking=True)
    if _upload_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _upload_pool = ThreadPoolExecutor(max_workers=4)
    flat = a.reshape(-1)
    stage = torch.empty(flat.shape, dtype=torch.from_numpy(flat[:0]).dtype, pin_memory=True)
    dev = torch.empty(flat.shape, dtype=stage.dtype, device="cuda")
    sn = stage.numpy()
    n = flat.size
    step = -(-n // 8)
    futs = [(lo, _upload_pool.submit(np.copyto, sn[lo:lo + step], flat[lo:lo + step])) for lo in range(0, n, step)]
    for lo, fu in futs:
        fu.result()
        dev[lo:lo + step].copy_(stage[lo:lo + step], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    ev.synchronize()              # the staging block goes back to the cache: the copies must have read it
    return dev.view(a.shape)


class Plan:
    """Grid-dependent device tables (falwa_b200_plan_create)."""

    def __init__(self, kind, xlon, ylat, plev, kmax, dz, jb, npart, maxit, tol, rjac, scale_height, cp,
                 dry_gas_constant, omega, planet_radius, on_height_grid=False):
        _lib.require_cuda()
        _register()
        self.lib = _lib.load()
        self.kind = kind
        self.nlon, self.nlat, self.nlev, self.kmax = int(xlon.size), int(ylat.size), int(plev.size), int(kmax)
        self.jb = int(jb) if kind == 1 else 0
        self.nd = self.nlat // 2 + 1
        self.jd = self.nd - self.jb
        self.dz, self.h = float(dz), float(scale_height)
        self.height = np.array([i * dz for i in range(kmax)])                      # oopinterface.py:121
        self.plev_to_height = -scale_height * np.log(np.asarray(plev, float) / P_GROUND)   # :120
        self.theta_factor = np.exp(dry_gas_constant / cp * self.plev_to_height / scale_height)  # :127-128
        self.clat = np.abs(np.cos(np.deg2rad(ylat)))                               # :342
        self.sinlat = np.sin(np.deg2rad(ylat))                                     # data_storage.py:175-176
        self.prefactor = sum([math.exp(-k * dz / scale_height) * dz for k in range(1, kmax - 1)])  # :270
        if on_height_grid:  # data_on_evenly_spaced_pseudoheight_grid: no interpolation tables needed
            p2h = np.ascontiguousarray(self.height)
        else:
            p2h = np.ascontiguousarray(self.plev_to_height)
        self._keep = (p2h, np.ascontiguousarray(self.theta_factor), np.ascontiguousarray(self.clat),
                      np.ascontiguousarray(self.sinlat))
        handle = ctypes.c_void_p()
        check(self.lib.falwa_b200_plan_create(
            ctypes.cast(ctypes.byref(handle), ctypes.c_void_p), kind, self.nlon, self.nlat,
            p2h.size, self.kmax, self.jb, int(npart) if npart is not None else 0, int(maxit), hptr(self._keep[0]),
            hptr(self._keep[1]), hptr(self._keep[2]), hptr(self._keep[3]), float(planet_radius), float(omega),
            float(dz), float(scale_height), float(dry_gas_constant), float(cp), float(tol), float(rjac),
            float(self.prefactor)), "plan_create")
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.falwa_b200_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown
            pass


class QGBatch:
    """B snapshots on the current CUDA device.  Inputs: u, v, t as [B][nlev][nlat][nlon] float64
    (NumPy -> copied from pinned memory, or CUDA tensors -> used as is), latitude ascending, pressure descending."""

    def __init__(self, plan, u, v, t, on_height_grid=False, northern_hemisphere_results_only=False):
        self.plan = plan
        self.lib = plan.lib
        self.on_height_grid = on_height_grid
        self.north_only = int(bool(northern_hemisphere_results_only))
        self.u_in, self.v_in, self.t_in = (self._to_device(x) for x in (u, v, t))
        self.B = int(self.u_in.shape[0])
        exp = (self.B, plan.nlev if not on_height_grid else plan.kmax, plan.nlat, plan.nlon)
        for name, x in (("u", self.u_in), ("v", self.v_in), ("t", self.t_in)):
            if tuple(x.shape) != exp:
                raise TypeError(f"Incorrect dimension of {name}_field. Expected dimension: {exp}")
        self._means_host = self._means_dev = self._means_ev = None
        self.profiles = None          # host [B][4][kmax]: t0_s, t0_n, stat_s, stat_n
        self.profiles_hemi = None
        self.u = self.v = self.theta = self.avort = self.qgpv = self.zm = self.ext = None
        self.qref_st = self.qref_cu = self.uref = self.ptref = None
        self.num_of_iter = None
        self.lwa = self.out2d = None

    @staticmethod
    def _to_device(x):
        if isinstance(x, torch.Tensor):
            assert x.is_cuda and x.dtype == torch.float64
            return x.contiguous()
        return to_device(np.ascontiguousarray(x, dtype=np.float64))

    def _new(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.u_in.device)

    # ---- stage 1 -------------------------------------------------------------------------------
    def launch_theta_means(self):
        """Queue the hemispheric-mean reduction and its copy to pinned host memory without waiting for it, so that a
        caller can queue other work (the previous batch's flux kernels) behind it and fit the splines meanwhile."""
        if self._means_host is not None:
            return
        p = self.plan
        # the reduction kernel stores its B x 2 x nlev results straight into pinned (device-mapped) host memory from a
        # small per-plan ring: no cudaHostAlloc per batch (it would synchronise the device) and no copy-engine
        # transfer (it would queue behind another stream's bulk PCIe traffic)
        ring = p.__dict__.setdefault("_means_ring", {})
        key = (self.B, 2, int(self.t_in.shape[1]))
        bufs = ring.setdefault(key, [torch.empty(key, dtype=torch.float64).pin_memory() for _ in range(4)])
        p._means_next = (getattr(p, "_means_next", -1) + 1) % len(bufs)
        self._means_host = bufs[p._means_next]
        check(self.lib.falwa_b200_theta_hemispheric_means(p.handle, self.B, dptr(self.t_in),
                                                          ctypes.c_void_p(self._means_host.data_ptr()),
                                                          stream_ptr()), "theta_hemispheric_means")
        self._means_ev = torch.cuda.Event()
        self._means_ev.record()

    def hemispheric_theta_means(self):
        """[B][2][nlev] cos-weighted hemispheric means of potential temperature (device reduction)."""
        self.launch_theta_means()
        self._means_ev.synchronize()
        return self._means_host.numpy().copy()

    def static_stability_profiles(self):
        """oopinterface.py:241-261 and :530-531: smoothing splines of the hemispheric-mean theta(z) and their
        derivatives, evaluated on the height grid.  Returns [B][4][kmax] (t0_s, t0_n, stat_s, stat_n)."""
        p = self.plan
        means = self.hemispheric_theta_means()
        x = p.height if self.on_height_grid else p.plev_to_height
        prof = np.empty((self.B, 4, p.kmax))

        def fit(b):
            us = UnivariateSpline(x=x, y=means[b, 0])
            un = UnivariateSpline(x=x, y=means[b, 1])
            prof[b, 0], prof[b, 1] = us(p.height), un(p.height)
            prof[b, 2], prof[b, 3] = us.derivative()(p.height), un.derivative()(p.height)
        for b in range(self.B):   # ~0.17 ms per snapshot; the FITPACK wrappers hold the GIL, threads do not help
            fit(b)
        return prof

    def compute_profiles(self):
        """Static-stability profiles of the batch (the only host arithmetic of the path: the reference's spline)."""
        if self.profiles is not None:
            return
        prof = self.static_stability_profiles()
        self.profiles_hemi = np.ascontiguousarray(prof)   # t0_s, t0_n, stat_s, stat_n before any averaging
        if self.plan.kind == 0:  # NH18 uses the average of the two hemispheres everywhere (oopinterface.py:1567-1580)
            t0 = 0.5 * (prof[:, 0] + prof[:, 1])
            st = 0.5 * (prof[:, 2] + prof[:, 3])
            prof = np.stack([t0, t0, st, st], axis=1)
        self.profiles = np.ascontiguousarray(prof)

    def interpolate_fields(self):
        p = self.plan
        self.compute_profiles()
        shape = (self.B, p.kmax, p.nlat, p.nlon)
        if self.on_height_grid:
            # the inputs already sit on the height grid (nlev == kmax): theta = T * factor(level) is formed by the
            # library (theta_from_t_kernel) inside falwa_b200_interpolate_fields
            self.u, self.v, self.theta = self.u_in, self.v_in, self._new(*shape)
        else:
            self.u, self.v, self.theta = self._new(*shape), self._new(*shape), self._new(*shape)
        self.avort, self.qgpv = self._new(*shape), self._new(*shape)
        # zonal means of (qgpv, u, theta) and per-level extrema of (qgpv, avort) can be fused into the QGPV kernel
        # (FUSE_ZONAL_STATS); measured slower than the separate row-reduction kernel on B200, so off by default
        if FUSE_ZONAL_STATS:
            self.zm = self._new(self.B, 3, p.kmax, p.nlat)
            self.ext = torch.empty((self.B, 2, p.kmax, 2), dtype=torch.int64, device=self.u_in.device)
        check(self.lib.falwa_b200_interpolate_fields(
            p.handle, self.B, dptr(self.u_in), dptr(self.v_in), dptr(self.t_in), int(self.on_height_grid),
            hptr(self.profiles), dptr(self.u), dptr(self.v), dptr(self.theta), dptr(self.avort), dptr(self.qgpv),
            dptr(self.zm), dptr(self.ext), stream_ptr()), "interpolate_fields")

    # ---- stage 2 -------------------------------------------------------------------------------
    def compute_reference_states(self):
        p = self.plan
        if self.qgpv is None:
            raise ValueError("QGField.interpolate_fields has to be called before QGField.compute_reference_states.")
        shape = (self.B, p.kmax, p.nlat)
        self.qref_st, self.qref_cu, self.uref, self.ptref = (self._new(*shape) for _ in range(4))
        nit = np.zeros((self.B, 2), dtype=np.int32)
        check(self.lib.falwa_b200_reference_states(
            p.handle, self.B, dptr(self.qgpv), dptr(self.u), dptr(self.theta), dptr(self.avort), hptr(self.profiles),
            self.north_only, dptr(self.qref_st), dptr(self.qref_cu), dptr(self.uref), dptr(self.ptref),
            nit.ctypes.data_as(ctypes.c_void_p), dptr(self.zm), dptr(self.ext), stream_ptr()), "reference_states")
        self.num_of_iter = nit

    # ---- stage 3 -------------------------------------------------------------------------------
    def compute_lwa_and_barotropic_fluxes(self, ncforce=None, method=0):
        p = self.plan
        if p.nd > 416 and method != 1:   # lwa_window.cu holds 384 hemispheric rows, flux_scan.cu 416
            import warnings
            warnings.warn(f"{p.nd} latitudes per hemisphere exceed what the fast LWA kernels hold on chip (384 / 416): "
                          "falling back to the reference's O(nd^2) double loop on the device (about 100x slower)",
                          RuntimeWarning, stacklevel=2)
        if self.qref_cu is None:
            raise ValueError("Reference states have not been computed yet.")
        nc = None
        if ncforce is not None:
            nc = self._to_device(ncforce)
            assert tuple(nc.shape) == tuple(self.theta.shape)
        self.lwa = self._new(self.B, p.kmax, p.nlat, p.nlon)
        self.out2d = self._new(self.B, N_OUT2D, p.nlat, p.nlon)
        check(self.lib.falwa_b200_lwa_and_barotropic_fluxes(
            p.handle, self.B, dptr(self.qgpv), dptr(self.u), dptr(self.v), dptr(self.theta), dptr(nc),
            hptr(self.profiles), dptr(self.qref_cu), dptr(self.uref), dptr(self.ptref), self.north_only,
            dptr(self.lwa), dptr(self.out2d), int(method), stream_ptr()), "lwa_and_barotropic_fluxes")

    def compute_layerwise_lwa_fluxes(self, ncforce=None, method=0):
        """Layer-wise flux terms of the batch (oopinterface.py:1063-1156): -> device tensor [B][10][kmax][nlat][nlon] in
        the slot order of ``LAYER3D``."""
        p = self.plan
        if self.qref_cu is None:
            raise ValueError("Reference states have not been computed yet.")
        nc = None
        if ncforce is not None:
            nc = self._to_device(ncforce)
            assert tuple(nc.shape) == tuple(self.theta.shape)
        self.layer3d = self._new(self.B, len(LAYER3D), p.kmax, p.nlat, p.nlon)
        check(self.lib.falwa_b200_layerwise_lwa_fluxes(
            p.handle, self.B, dptr(self.qgpv), dptr(self.u), dptr(self.v), dptr(self.theta), dptr(nc),
            hptr(self.profiles), dptr(self.qref_cu), dptr(self.uref), dptr(self.ptref), self.north_only,
            dptr(self.layer3d), int(method), stream_ptr()), "layerwise_lwa_fluxes")
        return self.layer3d

    def ncforce_from_heating_rate(self, heating_rate):
        """q-dot on the height grid from the heating rate on the input levels, [B][nlev][nlat][nlon] ->
        [B][kmax][nlat][nlon] (oopinterface.py:985-1019)."""
        p = self.plan
        if self.profiles is None:
            raise ValueError("QGField.interpolate_fields has to be called first.")
        hr = self._to_device(heating_rate)
        assert tuple(hr.shape) == tuple(self.t_in.shape), "heating_rate must have the shape of the input fields"
        out = self._new(self.B, p.kmax, p.nlat, p.nlon)
        check(self.lib.falwa_b200_ncforce_from_heating_rate(p.handle, self.B, dptr(hr), 0, hptr(self.profiles),
                                                            dptr(out), stream_ptr()), "ncforce_from_heating_rate")
        return out

    def out(self, name):
        """2-D output `name` as a [B][nlat][nlon] device view."""
        return self.out2d[:, OUT2D[name]]

    def run(self, ncforce=None, method=0):
        self.interpolate_fields()
        self.compute_reference_states()
        self.compute_lwa_and_barotropic_fluxes(ncforce=ncforce, method=method)
        return self


class HostStream:
    """Streams host-resident snapshots through one GPU in chunks: pinned host -> device copy, the three stages,
    device -> pinned host copy of the results, with the copies of neighbouring chunks overlapping the kernels
    (separate copy-in / compute / copy-out streams, two input slots).  This is what ``QGDataset`` uses for a shard
    of time steps and what bench.py times as the end-to-end number.

    Results (per snapshot, the return values of the reference's ``compute_reference_states`` and
    ``compute_lwa_and_barotropic_fluxes``, oopinterface.py:550-615, :796-962):
      qref, uref, ptref [N][kmax][nlat];  the 2-D fields of ``OUT2D`` [N][15][nlat][nlon];  lwa [N][kmax][nlat][nlon]
      (optional, ``want_lwa3d``).
    """

    def __init__(self, plan, chunk=1, want_lwa3d=True, method=0, on_height_grid=False, input_dtype=torch.float64,
                 packing=None, flip_lat=False, flip_lev=False, byteswap=False):
        self.plan, self.chunk, self.want_lwa3d, self.method = plan, int(chunk), want_lwa3d, method
        self.on_height_grid = on_height_grid
        # float32 or CF-packed int16 host inputs in file order cross PCIe as they are and are decoded / reordered on
        # the device (ingest_field); ``packing`` = one (scale_factor, add_offset[, fill_value]) per field or None
        self.input_dtype = input_dtype
        self.packing = tuple(packing) if packing is not None else (None, None, None)
        self.flip_lat, self.flip_lev, self.byteswap = bool(flip_lat), bool(flip_lev), bool(byteswap)
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        nl = plan.kmax if on_height_grid else plan.nlev
        shape = (self.chunk, nl, plan.nlat, plan.nlon)
        self.slots = [tuple(torch.empty(shape, dtype=input_dtype, device="cuda") for _ in range(3)) for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = self.d2h_bytes = 0
        self.trace = None    # set to a list to collect (phase, chunk, start event, end event) for a timeline

    def _mark(self, stream):
        if self.trace is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    @staticmethod
    def pinned(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype).pin_memory()

    def alloc_outputs(self, n):
        p = self.plan
        out = dict(qref=self.pinned((n, p.kmax, p.nlat)), uref=self.pinned((n, p.kmax, p.nlat)),
                   ptref=self.pinned((n, p.kmax, p.nlat)), out2d=self.pinned((n, N_OUT2D, p.nlat, p.nlon)),
                   num_of_iter=np.zeros((n, 2), dtype=np.int32))
        if self.want_lwa3d:
            out["lwa"] = self.pinned((n, p.kmax, p.nlat, p.nlon))
        return out

    def _h2d(self, slot, host, lo, hi):
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_free[slot])
            e0 = self._mark(self.s_in)
            for dst, src in zip(self.slots[slot], host):
                dst[:hi - lo].copy_(src[lo:hi], non_blocking=True)
                self.h2d_bytes += (hi - lo) * dst[0].numel() * dst.element_size()
            self.ev_in[slot].record(self.s_in)
            if e0 is not None:
                self.trace.append(("h2d", lo, e0, self._mark(self.s_in)))

    def copy_only(self, u, v, t, out, cycles=1):
        """The host <-> device traffic of ``run`` with no kernels in between: the same pinned inputs, chunk by chunk
        into the two device slots on the input stream, and result-sized device buffers back into ``out`` on the output
        stream.  What this sustains is the ceiling of the end-to-end rate on this box (bench.py: ``e2e.ceiling``)."""
        host = [u, v, t]
        n = int(u.shape[0])
        self.h2d_bytes = self.d2h_bytes = 0
        names = [k for k in ("qref", "uref", "ptref", "out2d", "lwa") if k in out and isinstance(out[k], torch.Tensor)]
        dev = {k: torch.empty((self.chunk,) + tuple(out[k].shape[1:]), dtype=out[k].dtype, device="cuda") for k in names}
        main = torch.cuda.current_stream()
        for e in self.ev_free:
            e.record(main)
        bounds = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)] * int(cycles)
        for c, (lo, hi) in enumerate(bounds):
            slot = c & 1
            self._h2d(slot, host, lo, hi)
            self.ev_free[slot].record(self.s_in)       # nothing reads the slot: free as soon as the copy is done
            with torch.cuda.stream(self.s_out):
                for k in names:
                    out[k][lo:hi].copy_(dev[k][:hi - lo], non_blocking=True)
                    self.d2h_bytes += (hi - lo) * dev[k][0].numel() * 8
        self.s_in.synchronize()
        self.s_out.synchronize()

    def run(self, u, v, t, out=None, cycles=1):
        """u, v, t: pinned host tensors [N][nlev][nlat][nlon] of ``input_dtype`` (NumPy arrays are staged through
        pinned memory first).  Returns the dict of pinned host result tensors (``out`` is reused when given).
        ``cycles`` > 1 streams the same N host snapshots that many times through one uninterrupted pipeline (the
        results of the last cycle remain in ``out``): a long time series without holding it all in host memory."""
        host = []
        for x in (u, v, t):
            if not isinstance(x, torch.Tensor):
                x = torch.from_numpy(np.ascontiguousarray(x))
            if x.dtype != self.input_dtype:
                if not self.input_dtype.is_floating_point:
                    raise TypeError(f"HostStream(input_dtype={self.input_dtype}) was given {x.dtype} values")
                x = x.to(self.input_dtype)
            if not x.is_pinned():
                x = x.pin_memory()
            host.append(x)
        n = int(host[0].shape[0])
        if out is None:
            out = self.alloc_outputs(n)
        self.h2d_bytes = self.d2h_bytes = 0
        main = torch.cuda.current_stream()
        for e in self.ev_free:
            e.record(main)
        bounds = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)] * int(cycles)
        marks = {}

        def chunks():
            """Device batches in order; called back by run_stream when stage 1 and 2 of the previous chunk are queued."""
            self._h2d(0, host, *bounds[0])
            for c, (lo, hi) in enumerate(bounds):
                slot = c & 1
                if c >= 1:   # the other slot was last read by stage 1 of chunk c-1, which is queued by now
                    self.ev_free[slot ^ 1].record(main)
                if c + 1 < len(bounds):
                    self._h2d(slot ^ 1, host, *bounds[c + 1])
                main.wait_event(self.ev_in[slot])
                marks[c] = self._mark(main)
                fields = [x[:hi - lo] for x in self.slots[slot]]
                for i, pk in enumerate(self.packing):
                    # (data on the height grid is read by all three stages: it must leave the input slot too)
                    if needs_ingest(self.input_dtype, pk, self.flip_lat, self.flip_lev, self.byteswap) \
                            or self.on_height_grid:
                        fields[i] = ingest_field(fields[i], pk, self.flip_lat, self.flip_lev, byteswap=self.byteswap)
                yield fields

        # run_stream queues the hemispheric means of chunk c+1 before the flux stage of chunk c: the host fits the
        # splines of c+1 while the device is busy, and the copies of the neighbouring chunks overlap both
        for c, b in enumerate(run_stream(self.plan, chunks(), method=self.method, on_height_grid=self.on_height_grid)):
            lo, hi = bounds[c]
            done = torch.cuda.Event()
            done.record(main)
            if marks.get(c) is not None:
                self.trace.append(("compute", lo, marks.pop(c), self._mark(main)))
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                e0 = self._mark(self.s_out)
                pairs = [("qref", b.qref_cu), ("uref", b.uref), ("ptref", b.ptref), ("out2d", b.out2d)]
                if self.want_lwa3d:
                    pairs.append(("lwa", b.lwa))
                for name, dev in pairs:
                    dev.record_stream(self.s_out)
                    out[name][lo:hi].copy_(dev, non_blocking=True)
                    self.d2h_bytes += dev.numel() * 8
                if e0 is not None:
                    self.trace.append(("d2h", lo, e0, self._mark(self.s_out)))
            out["num_of_iter"][lo:hi] = b.num_of_iter
        self.s_out.synchronize()
        return out


def run_stream(plan, batches, method=0, ncforce=None, on_height_grid=False, northern_hemisphere_results_only=False):
    """Run a sequence of device-resident batches through the whole pipeline, yielding each finished ``QGBatch``.

    ``batches`` yields (u, v, t) CUDA tensors [B][nlev][nlat][nlon].  The only host work on the critical path is the
    static-stability spline (SciPy, ~0.17 ms per snapshot); here the hemispheric means of batch n+1 are queued BEFORE
    the flux kernels of batch n, so the splines of n+1 are fitted while the GPU is busy with n's LWA / flux stage."""
    prev = None
    for u, v, t in batches:
        b = QGBatch(plan, u, v, t, on_height_grid=on_height_grid,
                    northern_hemisphere_results_only=northern_hemisphere_results_only)
        b.launch_theta_means()
        if prev is not None:
            prev.compute_lwa_and_barotropic_fluxes(ncforce=ncforce, method=method)
            yield prev
        b.interpolate_fields()          # waits for the means only, fits the splines, queues stage 1
        b.compute_reference_states()
        prev = b
    if prev is not None:
        prev.compute_lwa_and_barotropic_fluxes(ncforce=ncforce, method=method)
        yield prev
