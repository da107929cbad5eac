#!/usr/bin/env python
"""bench.py — snapshots/s of the full QGFieldNHN22 pipeline at ERA5 0.25° (BASELINE.json configs[2]).

One "step" = interpolate_fields + compute_reference_states + compute_lwa_and_barotropic_fluxes (both
hemispheres, default jb=5, kmax=49) over one batch of synthetic snapshots per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--grid 0.25|1.5] [--impl b200|reference]

  value      snapshots/s with the batch already resident in HBM (CUDA events on the launching stream, barrier +
             synchronize on both sides, max over ranks).  Inputs of one step are B x 0.92 GB >> 126 MB L2.
  e2e        the same pipeline through hn2016_falwa_b200.engine.HostStream: pinned HOST inputs, H2D and D2H copies
             (reference states, the 2-D flux fields and 3-D LWA) inside the timed region.
  roofline   the kernel with the largest share of the step, timed live with CUDA events inside the library
             (falwa_b200_profile_*), against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline / --impl reference
             the CPU oracle (C++/NumPy restatement of the reference's Fortran/Python path; the reference's own
             Fortran cannot be compiled in this image) on the host cores, bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRIDS = {"0.25": (1440, 721), "1.5": (240, 121)}
KMAX, NLEV = 49, 37
CPU_THREADS_PER_PROCESS_025 = 2   # see cpu_split(): 8 x 2 is the best split of the 16-core box (profiles/r02_cpu_sweep.jsonl)


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (first sample at once, then
    every 250 ms: denser NVML / nvidia-smi polling measurably slows the kernel launches of the step loop)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.err = index, [], set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def sample_now(self):
        """One sample (also called by the main thread right after the last launch of the timed region has been
        queued, while the device is still working through it)."""
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for b, name in self.REASONS.items():
                if bits & b:
                    self.reasons.add(name)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.25)   # every NVML query stalls the CUDA driver for a few ms: sample sparsely

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": self.err or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


_ORIG_AFFINITY = {}


def bind_to_gpu_numa_node(index):
    """Best effort: restrict this process to the CPUs NVML reports as local to GPU `index`, so that the pinned host
    buffers (first touch) and the copy threads sit on the socket the GPU hangs off.  With 8 ranks on a two-socket box
    the end-to-end rate is otherwise limited by cross-socket traffic."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        _ORIG_AFFINITY.setdefault("cpus", os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:  # noqa: BLE001 - affinity is an optimisation only
        pass
    return None


def workload_config(nlon, nlat, data="analytic"):
    """The `config` object of the JSON line: identical for the B200 arm and the reference arm."""
    what = ("seeded synthetic snapshots (hn2016_falwa_b200.synthetic), zonally rotated copies per batch" if data == "analytic"
            else "the reference's ERA-Interim 2005-01-23 00Z snapshot bilinearly interpolated to the grid, zonally rotated "
                 "copies per batch")
    return {"workload": f"QGFieldNHN22 {nlon}x{nlat}x{NLEV}plev->kmax{KMAX}, jb=5, both hemispheres",
            "grid": {"nlon": nlon, "nlat": nlat, "nlev": NLEV, "kmax": KMAX}, "protocol": "NHN22", "jb": 5,
            "data": what}


def cpu_split(nlon, cores):
    """processes x OpenMP threads of the CPU arm — the best of the sweep kept in profiles/r02_cpu_sweep.jsonl
    (tools/cpu_sweep.py on the GPU box's host cores); used by `cpu_baseline` and by `--impl reference` alike."""
    if nlon > 1000:
        workers = max(1, cores // CPU_THREADS_PER_PROCESS_025)
    else:
        workers = min(8, cores)
    return workers


REAL_SNAPSHOT = os.path.join(ROOT, "tests", "golden", "era_interim_2005-01-23_00Z.npz")


def make_batch(nlon, nlat, batch, seed, data="analytic"):
    """One seeded synthetic snapshot (hn2016_falwa_b200.synthetic) + zonally rotated copies of it: B distinct
    snapshots [B][nlev][nlat][nlon] without B x 8 s of host trigonometry.  data="upsampled": SURVEY.md §8(d)(i), the
    reference's own 1.5-degree ERA-Interim snapshot (tests/golden) bilinearly interpolated to the grid — large
    meanders, i.e. LWA windows an order of magnitude longer than the analytic waves'."""
    from hn2016_falwa_b200.synthetic import load_packed_snapshot, synthetic_snapshot, upsampled_snapshot
    if data == "upsampled":
        x0, y0, plev, u0, v0, t0 = load_packed_snapshot(REAL_SNAPSHOT)
        xlon, ylat, u, v, t = upsampled_snapshot(x0, y0, u0, v0, t0, nlon, nlat)
    else:
        xlon, ylat, plev, u, v, t = synthetic_snapshot(nlon=nlon, nlat=nlat, seed=seed)
    shifts = [(b * 37 * nlon) // 360 for b in range(batch)]
    return xlon, ylat, plev, u, v, t, shifts


def algorithmic_bytes(nlon, nlat, nd, kmax, nlev, has_avort=True):
    """Compulsory bytes per snapshot of each kernel, fp64."""
    f3, f3in, f2 = 8.0 * nlon * nlat * kmax, 8.0 * nlon * nlat * nlev, 8.0 * nlon * nlat
    hemi3, hemi2 = 8.0 * nlon * nd * kmax, 8.0 * nlon * nd
    return {
        "theta_rowmean_kernel": f3in,
        "interp_kernel": 3 * f3in + 3 * f3,
        "interp_qgpv_kernel": 3 * f3in + 5 * f3,                                # read u, v, T; write u, v, theta, avort, qgpv
        "avort_pv_kernel": 3 * f3 + 2 * f3,
        "zonal_stats_kernel": 4 * f3,
        "area_hist3_kernel": 2 * f3,
        "fused_flux_bruteforce_kernel": 2 * (4 * hemi3 + hemi3 + 11 * hemi2),
        "lwa_scan_kernel": 2 * (2 * hemi3 + 3 * hemi3),                        # read pv, u; write a1, a2, ua2
        "flux_column_kernel": 2 * (6 * hemi3 + hemi3 + 12 * hemi2),            # read a1,a2,ua2,u,v,theta; write lwa + 2-D
        "lwa_window_kernel": 2 * (2 * hemi3 + hemi3 + 5 * hemi2),              # read pv, u; write lwa + 5 vertical sums
        "lwa_prefix_kernel": 2 * (2 * hemi3 + hemi3 + 5 * hemi2),              # same traffic, other algorithm
        "flux_uvt_column_kernel": 2 * (3 * hemi3 + 8 * hemi2),                 # read u, v, theta; write 2-D sums
        "baro_post_kernel": 2 * f3 + 8 * f2,
        "pipeline": 3 * f3in + 4 * f3 + f3 + 12 * f2,                         # 3056.6 MB at 0.25°
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle pipeline on the host cores (reported baseline; see module docstring)
# ------------------------------------------------------------------------------------------------
_CPU_INPUT = {}


def _cpu_init(nlon, nlat, threads, precision, data="analytic"):
    """Pool initialiser: OpenMP width of this worker, arithmetic type of the native routines, and the worker's synthetic
    snapshot (generated once, reused every step)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from hn2016_falwa_b200.synthetic import synthetic_snapshot
    from oracle import f2py_shim
    f2py_shim.set_precision(np.float32 if precision == "f32" else np.float64)
    if data == "upsampled":
        xlon, ylat, plev, u, v, t, _ = make_batch(nlon, nlat, 1, 0, data=data)
        _CPU_INPUT["inp"] = (xlon, ylat, plev, u, v, t)
    else:
        _CPU_INPUT["inp"] = synthetic_snapshot(nlon=nlon, nlat=nlat, seed=1234 + os.getpid() % 97)


def _cpu_one(_):
    from oracle import pipeline
    t0 = time.perf_counter()
    pipeline.run_qgfield("nhn22", *_CPU_INPUT["inp"])
    return time.perf_counter() - t0


class CpuReference:
    """The CPU oracle pipeline on the host cores: `workers` processes x (cores // workers) OpenMP threads."""

    def __init__(self, nlon, nlat, workers, precision="f64", threads=None, data="analytic"):
        import multiprocessing as mp
        import subprocess
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        self.cores = os.cpu_count() or 1
        self.workers = max(1, min(workers, self.cores))
        self.threads = threads or max(1, self.cores // self.workers)
        self.precision = precision
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_init,
                                                 initargs=(nlon, nlat, self.threads, precision, data))
        self.pool.map(_cpu_one, [0] * 0)   # start the workers

    def step(self, snapshots):
        """-> snapshots/s of one step of `snapshots` pipeline runs (wall clock over the pool)."""
        t0 = time.perf_counter()
        self.pool.map(_cpu_one, range(snapshots), chunksize=1)
        return snapshots / (time.perf_counter() - t0)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_reference_rate(nlon, nlat, snapshots, workers, precision="f64", threads=None, data="analytic"):
    ref = CpuReference(nlon, nlat, workers, precision, threads, data)
    ref.step(ref.workers)            # untimed: worker start-up + input synthesis
    t0 = time.perf_counter()
    rate = ref.step(snapshots)
    wall = time.perf_counter() - t0
    ref.close()
    return rate, ref.cores, ref.workers, ref.threads, wall


def run_reference(args, nlon, nlat, rank, world):
    if rank != 0:
        return
    workers = cpu_split(nlon, os.cpu_count() or 1)
    per_step = workers if nlon > 1000 else 4 * workers
    ref = CpuReference(nlon, nlat, workers, data=args.data)
    for _ in range(max(1, args.warmup and 1)):
        ref.step(ref.workers)        # warm-up: worker start-up, input synthesis, library load (one pass is enough on a CPU)
    rates = [ref.step(per_step) for _ in range(max(1, args.steps))]
    ref.close()
    value = float(np.mean(rates))
    cores, workers, threads = ref.cores, ref.workers, ref.threads
    sample = f"{per_step} synthetic {nlon}x{nlat} snapshot(s) per step, {workers} processes x {threads} OpenMP threads, fp64"
    line = {
        "impl": "reference", "metric": "snapshots/sec, QGFieldNHN22 full pipeline", "value": value,
        "unit": "snapshots/s", "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": args.warmup,
        "ms_per_step": 1e3 * per_step / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(nlon, nlat, args.data),
        "run": {"snapshots_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "snapshots/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "snapshots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle (C++/NumPy restatement, fp64); the reference's Fortran cannot be built in this image",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, nlon, nlat, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from hn2016_falwa_b200 import _lib, engine
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22

    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)   # pinned staging buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up: send it to stderr so that
        # stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = args.batch
    xlon, ylat, plev, u, v, t, shifts = make_batch(nlon, nlat, B, seed=1234 + rank, data=args.data)
    # plan via the public class (validates the grid exactly like the reference's constructor)
    f0 = QGFieldNHN22(xlon, ylat, plev, u, v, t)
    plan = f0._plan
    del f0
    hu, hv, ht = (engine.HostStream.pinned((B,) + u.shape) for _ in range(3))
    for b, s in enumerate(shifts):
        hu[b].copy_(torch.from_numpy(np.roll(u, s, axis=-1)))
        hv[b].copy_(torch.from_numpy(np.roll(v, s, axis=-1)))
        ht[b].copy_(torch.from_numpy(np.roll(t, s, axis=-1)))
    du, dv, dt_ = hu.cuda(), hv.cuda(), ht.cuda()
    torch.cuda.synchronize()

    def step():
        return engine.QGBatch(plan, du, dv, dt_).run(method=args.method)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in engine.run_stream(plan, ((du, dv, dt_) for _ in range(args.warmup)), method=args.method):
        pass   # warm-up through the same pipelined path (allocator caches both batch slots, kernels loaded)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in engine.run_stream(plan, ((du, dv, dt_) for _ in range(args.steps)), method=args.method):
        pass   # K steps back to back; the host-side spline of step n+1 overlaps the flux kernels of step n
    e1.record()
    sampler.sample_now()   # everything is queued, the device is still busy with the last steps: a sample under load
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    clocks = sampler.result()
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: pinned host inputs -> results in pinned host memory ---------------------
    if args.kernels_only:   # tuning runs: device-resident rate and the per-kernel times only
        _lib.profile_enable(True)
        step()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms / args.steps,
                              "kernels": {k: m for k, (n, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}}))
        return
    hs = engine.HostStream(plan, chunk=args.chunk, want_lwa3d=True, method=args.method)
    out = hs.alloc_outputs(B)
    hs.run(hu, hv, ht, out=out)   # warm-up (allocator, pinned output pages)
    barrier()
    def time_e2e(stream_runner, srcs):
        w0 = torch.cuda.Event(enable_timing=True); w1 = torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        w0.record()
        stream_runner.run(*srcs, out=out, cycles=e2e_steps)   # one uninterrupted stream of e2e_steps x B snapshots
        w1.record()
        barrier()
        sec = max(time.perf_counter() - t0, w0.elapsed_time(w1) / 1e3)
        te = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * B * e2e_steps / float(te.item())

    e2e_steps = max(1, args.steps)
    e2e = {"value": time_e2e(hs, (hu, hv, ht)), "unit": "snapshots/s",
           "h2d_bytes_per_step": int(hs.h2d_bytes // e2e_steps), "d2h_bytes_per_step": int(hs.d2h_bytes // e2e_steps),
           "chunk": args.chunk, "steps": e2e_steps, "input_dtype": "f64"}

    class _CopyOnly:   # the same transfers with no kernels: what the host <-> device path of this box sustains
        def __init__(self, stream):
            self.stream = stream

        def run(self, *srcs, out, cycles):
            self.stream.copy_only(*srcs, out, cycles=cycles)
    e2e["ceiling"] = {"value": time_e2e(_CopyOnly(hs), (hu, hv, ht)), "unit": "snapshots/s",
                      "what": "the same H2D + D2H copies (all ranks at once), no kernels"}
    # the same stream with float32 host inputs (ERA5's native precision), widened to fp64 on the device (ingest kernel)
    hs32 = engine.HostStream(plan, chunk=args.chunk, want_lwa3d=True, method=args.method, input_dtype=torch.float32)
    h32 = tuple(engine.HostStream.pinned(x.shape, torch.float32).copy_(x) for x in (hu, hv, ht))
    hs32.run(*h32, out=out)
    e2e["f32_inputs"] = {"value": time_e2e(hs32, h32), "h2d_bytes_per_step": int(hs32.h2d_bytes // e2e_steps)}
    del hs32, h32
    # ... and with the fields as ERA5 files hold them: CF-packed int16 (scale_factor / add_offset), north -> south and
    # top -> bottom; decoded and reordered by the ingest kernel.  The arithmetic after decoding is fp64 as above.
    packing, h16 = [], []
    for x in (hu, hv, ht):
        lo_, hi_ = float(x.min()), float(x.max())
        sc, off = (hi_ - lo_) / 65534.0, 0.5 * (hi_ + lo_)
        q = torch.round((x.flip(1, 2) - off) / sc).clamp_(-32767, 32767).to(torch.int16)
        h16.append(engine.HostStream.pinned(q.shape, torch.int16).copy_(q))
        packing.append((sc, off))
        del q
    hs16 = engine.HostStream(plan, chunk=args.chunk, want_lwa3d=True, method=args.method, input_dtype=torch.int16,
                             packing=packing, flip_lat=True, flip_lev=True)
    hs16.run(*h16, out=out)
    e2e["packed_int16_inputs"] = {"value": time_e2e(hs16, h16),
                                  "h2d_bytes_per_step": int(hs16.h2d_bytes // e2e_steps),
                                  "ceiling": time_e2e(_CopyOnly(hs16), h16)}
    # ... and when only the reference states and the 2-D budget terms are wanted back (no 3-D LWA: 125 MB instead of
    # 532 MB of results per snapshot)
    hs2d = engine.HostStream(plan, chunk=args.chunk, want_lwa3d=False, method=args.method, input_dtype=torch.int16,
                             packing=packing, flip_lat=True, flip_lev=True)
    out2 = hs2d.alloc_outputs(B)
    hs2d.run(*h16, out=out2)
    saved_out, out = out, out2
    e2e["packed_int16_inputs_2d_outputs"] = {"value": time_e2e(hs2d, h16),
                                             "h2d_bytes_per_step": int(hs2d.h2d_bytes // e2e_steps),
                                             "d2h_bytes_per_step": int(hs2d.d2h_bytes // e2e_steps),
                                             "ceiling": time_e2e(_CopyOnly(hs2d), h16)}
    out = saved_out
    del hs16, h16, hs2d, out2

    # ---- per-kernel shares (CUDA events inside the library) and the roofline of the top kernel --
    roofline, kernels = None, {}
    if rank == 0:
        _lib.profile_enable(True)
        psteps = max(1, min(args.steps, 2))
        for _ in range(psteps):
            step()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        total = sum(ms_ for _, ms_ in prof.values()) or 1.0
        kernels = {k: {"launches_per_step": n / psteps, "ms_per_step": m / psteps, "share": m / total}
                   for k, (n, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        top = next(iter(kernels))
        ab = algorithmic_bytes(nlon, nlat, plan.nd, KMAX, NLEV)
        peak, which = peaks()
        n_launch, top_ms = prof[top]
        per_launch_ms = top_ms / n_launch
        bytes_launch = ab.get(top, 0.0) * B / (n_launch / psteps)   # table entries are per snapshot, all launches
        achieved = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and nlon == 1440:   # ncu --set full capture at batch 1, scaled to this launch
            rec = json.load(open(tpath)).get(top)
            if rec:
                traffic = rec["dram_bytes_per_snapshot"] * B / (n_launch / psteps)
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                    "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch": per_launch_ms,
                    "pipeline_frac": (ab["pipeline"] * B * args.steps / (ms * 1e-3) / 1e9) / peak}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if "cpus" in _ORIG_AFFINITY:   # the CPU baseline uses every host core again
                os.sched_setaffinity(0, _ORIG_AFFINITY["cpus"])
            workers = cpu_split(nlon, os.cpu_count() or 1)   # the same split as `--impl reference`
            snaps = workers if nlon > 1000 else 4 * workers
            rate, cores, workers, threads, wall = cpu_reference_rate(nlon, nlat, snaps, workers=workers, data=args.data)
            cpu = {"value": rate, "unit": "snapshots/s", "cores": cores, "kind": "port",
                   "sample": f"{snaps} synthetic {nlon}x{nlat} snapshot(s), {workers} process(es) x {threads} OpenMP "
                             f"threads, fp64, {wall:.1f} s wall"}
            if args.cpu_extra:   # the arithmetic type the reference ships (fp32) and a single core
                r32, _, _, _, w32 = cpu_reference_rate(nlon, nlat, snaps, workers=workers, precision="f32", data=args.data)
                r1, _, _, _, w1 = cpu_reference_rate(nlon, nlat, 1, workers=1, threads=1, data=args.data)
                cpu["fp32_as_shipped"] = {"value": r32, "sample": f"same split, fp32 native routines, {w32:.1f} s wall"}
                cpu["single_core"] = {"value": r1, "sample": f"1 process x 1 thread, fp64, {w1:.1f} s wall"}
        line = {
            "metric": "snapshots/sec, QGFieldNHN22 full pipeline", "value": value, "unit": "snapshots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(nlon, nlat, args.data),
            "run": {"snapshots_per_step_per_gpu": B, "l2": "inputs of one step exceed L2 (B x 0.92 GB at 0.25 deg)",
                    "lwa_method": args.method},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------
# BASELINE.json configs 4 and 5 (kept as JSON lines under profiles/; the headline is config 3)
# ------------------------------------------------------------------------------------------------
def _year_inputs(nt):
    from hn2016_falwa_b200.synthetic import synthetic_snapshot
    base = [synthetic_snapshot(nlon=240, nlat=121, seed=100 + k) for k in range(8)]
    xlon, ylat, plev = base[0][:3]
    u, v, t = (np.empty((nt,) + base[0][3].shape) for _ in range(3))
    for k in range(nt):
        sn, sh = base[k % 8], (k // 8) % 240
        u[k], v[k], t[k] = (np.roll(sn[i], sh, axis=-1) for i in (3, 4, 5))
    return xlon, ylat, plev, u, v, t


def run_config4(args, rank, world, local_rank):
    """One year of 6-hourly 1.5-degree snapshots (1460 time steps) through QGDataset, time-sharded over the ranks:
    host arrays in, host arrays out (reference states + barotropic terms; 3-D LWA with --keep-3d)."""
    nt = args.nt or 1460
    cfg = {"workload": f"QGDataset(QGFieldNHN22), {nt} snapshots 240x121x{NLEV}plev->kmax{KMAX} (BASELINE config 4)",
           "grid": {"nlon": 240, "nlat": 121, "nlev": NLEV, "kmax": KMAX}, "protocol": "NHN22", "jb": 5}
    metric = "snapshots/sec, one year of 6-hourly 1.5-degree snapshots through QGDataset"
    if args.impl == "reference":
        if rank != 0:
            return
        workers = cpu_split(240, os.cpu_count() or 1)
        ref = CpuReference(240, 121, workers)
        ref.step(workers)
        rates = [ref.step(4 * workers) for _ in range(max(1, args.steps))]
        ref.close()
        v = float(np.mean(rates))
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "snapshots/s", "n_gpus": args.gpus,
                          "steps": max(1, args.steps), "warmup": args.warmup, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "snapshots/s", "cores": ref.cores, "kind": "port",
                                           "sample": f"{4 * workers} snapshots per step, {workers} processes x "
                                                     f"{ref.threads} threads"},
                          "e2e": {"value": v, "unit": "snapshots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return
    import torch
    import torch.distributed as dist
    from hn2016_falwa_b200 import _lib, xarrayinterface as xi
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    xlon, ylat, plev, u, v, t = _year_inputs(nt)
    coords = dict(time=np.arange(nt), level=plev, latitude=ylat, longitude=xlon)
    dims = ("time", "level", "latitude", "longitude")
    mk = lambda a, n: xi.SimpleDataArray(a, dims, coords, name=n)
    ds = xi.SimpleDataset(dict(u=mk(u, "u"), v=mk(v, "v"), t=mk(t, "t")))
    times = []
    n0 = _lib.launch_count()
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        q = xi.QGDataset(ds, qgfield=QGFieldNHN22, keep_3d=args.keep_3d, chunk=128)
        q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if it >= args.warmup:
            times.append(float(tt.item()))
        lo, hi = q.shard
        nbytes_in = 3 * (hi - lo) * u[0].nbytes
        nbytes_out = sum(a.nbytes for a in q._store.values())
        del q
    launches = _lib.launch_count() - n0
    if rank == 0:
        sec = float(np.mean(times))
        print(json.dumps({"metric": metric, "value": nt / sec, "unit": "snapshots/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                          "run": {"keep_3d": bool(args.keep_3d), "chunk": 128},
                          "e2e": {"value": nt / sec, "unit": "snapshots/s", "h2d_bytes_per_step": int(nbytes_in),
                                  "d2h_bytes_per_step": int(nbytes_out),
                                  "what": "QGDataset is host arrays in, host arrays out: value is end to end"},
                          "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config5(args, rank, world, local_rank):
    """BarotropicField / basis equivalent latitude + LWA of 1460 2-D 0.25-degree absolute-vorticity fields, batched."""
    nt = args.nt or 1460
    nlat, nlon, chunk = 721, 1440, 146
    cfg = {"workload": f"BarotropicField eqvlat + LWA, {nt} fields {nlon}x{nlat} (BASELINE config 5)",
           "grid": {"nlon": nlon, "nlat": nlat}}
    metric = "fields/sec, 2-D equivalent latitude + LWA of 0.25-degree absolute vorticity"
    ylat = np.linspace(-90, 90, nlat)
    xlon = np.linspace(0, 360, nlon, endpoint=False)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import basis2d
        rng = np.random.default_rng(1234)
        pv = (2 * 7.29e-5 * np.sin(np.deg2rad(ylat))[:, None] + 8e-5 * np.cos(np.deg2rad(ylat))[:, None] *
              np.exp(-((ylat[:, None] - 36) / 10) ** 2) * np.cos(3 * np.deg2rad(xlon)[None, :]) +
              1e-6 * rng.standard_normal((nlat, nlon)))
        t0 = time.perf_counter()
        basis2d.barotropic_eqvlat_lwa(xlon, ylat, pv)
        v = 1.0 / (time.perf_counter() - t0)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "fields/s", "n_gpus": args.gpus,
                          "steps": 1, "warmup": 0, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "fields/s", "cores": 1, "kind": "port",
                                           "sample": "one field, NumPy restatement of basis.eqvlat_fawa + basis.lwa"},
                          "e2e": {"value": v, "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return
    import torch
    from hn2016_falwa_b200 import _lib
    from hn2016_falwa_b200.barotropic_field import BarotropicBatch
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    lo, hi = (nt * rank) // world, (nt * (rank + 1)) // world      # time shards
    phi = torch.deg2rad(torch.tensor(ylat, device="cuda"))[None, :, None]
    lam = torch.deg2rad(torch.tensor(xlon, device="cuda"))[None, None, :]
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)

    def make(t0, n):
        ph = (torch.arange(t0, t0 + n, device="cuda", dtype=torch.float64) * 0.05)[:, None, None]
        return (2 * 7.29e-5 * torch.sin(phi) + 8e-5 * torch.cos(phi) * torch.exp(-((torch.rad2deg(phi) - 36) / 10) ** 2)
                * torch.cos(3 * lam + ph) + 1e-6 * torch.randn((n, nlat, nlon), device="cuda", dtype=torch.float64,
                                                               generator=g))
    BarotropicBatch(xlon, ylat, make(0, 2)).compute()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for t0 in range(lo, hi, chunk):
        pv = make(t0, min(chunk, hi - t0))
        torch.cuda.synchronize()
        e0.record()
        BarotropicBatch(xlon, ylat, pv).compute()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"metric": metric, "value": nt / (ms / 1e3), "unit": "fields/s", "n_gpus": world, "steps": 1,
                          "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic (generated on the device)",
                          "config": cfg, "gpu_launches": int(_lib.launch_count() - n0), "e2e": None}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="snapshots per step per GPU")
    ap.add_argument("--chunk", type=int, default=1, help="snapshots per H2D/D2H chunk of the end-to-end run")
    ap.add_argument("--grid", default="0.25", choices=sorted(GRIDS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--data", default="analytic", choices=["analytic", "upsampled"],
                    help="analytic: seeded synthetic waves (default, SURVEY 8(d)(ii)); upsampled: the real 1.5-degree "
                         "snapshot interpolated to the grid (SURVEY 8(d)(i))")
    ap.add_argument("--method", type=int, default=0, help="LWA algorithm: 0 automatic (windowed sums), 1 double loop, 2 interval scan")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernels-only", action="store_true", help="tuning: skip the end-to-end and CPU parts")
    ap.add_argument("--cpu-extra", action="store_true", help="cpu_baseline also in fp32 (as shipped) and on one core")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE.json config: 3 = the headline (default), 4 = one year of 1.5-degree snapshots through "
                         "QGDataset, 5 = 2-D BarotropicField batch of 1460 0.25-degree fields")
    ap.add_argument("--nt", type=int, default=0, help="configs 4 / 5: number of time steps (default 1460)")
    ap.add_argument("--keep-3d", action="store_true", help="config 4: also bring the 3-D fields back to the host")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nlon, nlat = GRIDS[args.grid]
    if args.config == 4:
        return run_config4(args, rank, world, local_rank)
    if args.config == 5:
        return run_config5(args, rank, world, local_rank)
    if args.impl == "reference":
        run_reference(args, nlon, nlat, rank, world)
    else:
        run_b200(args, nlon, nlat, rank, world, local_rank)


if __name__ == "__main__":
    main()
