"""BASELINE.json config 3 at its full size (ERA5 0.25 deg: 1440 x 721, 37 levels -> kmax 49) against the CPU oracle.

Every user-facing field of QGFieldNHN22 (the benchmarked configuration: default LWA method, jb = 5, the separable
spectral solve in place of the block recursion + LAPACK of /root/reference/src/falwa/oopinterface.py:1839-1885 at
n = jd - 2 = 354) and of QGFieldNH18 (red-black SOR with a 361 x 49 grid in shared memory; iteration counts must be
equal) is compared with ``oracle.pipeline.run_qgfield`` — the literal restatement of the reference — on the bench's
seeded synthetic snapshot.  Equivalent-latitude bin assignments and LWA set membership (pair counts) must agree
exactly; the mismatch counts and the per-field error statistics are appended to gpurun_out/parity_fullsize.jsonl
(a copy of the GPU box's run is kept under profiles/).  rtol 1e-8 as stated by BASELINE.json's north_star; atol =
1e-9 of the field's maximum unless tighter (values that are differences of large sums)."""
import json
import os
import time

import numpy as np
import pytest

from oracle import f2py_shim as ora
from oracle import pipeline
from tests.parity import compare, ROOT

pytestmark = pytest.mark.gpu

REPORT = os.path.join(ROOT, "gpurun_out", "parity_fullsize.jsonl")
NLON, NLAT = 1440, 721
TOL = dict(qgpv=(1e-10, 1e-17), interpolated_u=(1e-13, 1e-13), interpolated_v=(1e-13, 1e-13),
           interpolated_theta=(1e-13, 0.0), ptref=(1e-9, 0.0))


def _note(**rec):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(rec) + "\n")


@pytest.fixture(scope="module")
def snap():
    from hn2016_falwa_b200.synthetic import synthetic_snapshot
    return synthetic_snapshot(nlon=NLON, nlat=NLAT, seed=1234)


@pytest.fixture(scope="module", params=["nhn22", "nh18"])
def case(request, snap):
    t = time.time()
    want = pipeline.run_qgfield(request.param, *snap)
    _note(what="oracle_seconds", kind=request.param, seconds=round(time.time() - t, 1), cores=os.cpu_count())
    return request.param, want


def _cls(kind):
    from hn2016_falwa_b200.oopinterface import QGFieldNH18, QGFieldNHN22
    return QGFieldNH18 if kind == "nh18" else QGFieldNHN22


def _f(a):  # python (kmax, nlat, nlon) -> Fortran-order view (nlon, nlat, kmax)
    return np.swapaxes(a, 0, 2)


def test_all_user_fields_match_the_oracle(snap, case):
    kind, want = case
    f = _cls(kind)(*snap)
    f.interpolate_fields(return_named_tuple=False)
    f.compute_reference_states(return_named_tuple=False)
    f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False)
    if kind == "nh18":
        got_it = tuple(int(x) for x in f._batch.num_of_iter[0])
        _note(what="sor_iterations", kind=kind, gpu=got_it, oracle=[int(x) for x in want["num_of_iter"]])
        assert got_it == tuple(want["num_of_iter"])
    failures = []
    for name in pipeline.USER_FIELDS:
        rtol, atol = TOL.get(name, (1e-8, 1e-9 * float(np.abs(want[name]).max())))
        try:
            rec = compare(f"fullsize/{kind}/{name}", getattr(f, name), want[name], rtol=rtol, atol=atol)
        except AssertionError as e:   # keep going: the report should list every field
            failures.append(str(e))
            continue
        _note(what="field", kind=kind, grid=[NLON, NLAT], **rec)
    assert not failures, "\n".join(failures)
    np.testing.assert_array_equal(f.flux_vector_lambda_baro, f.adv_flux_f1 + f.adv_flux_f2 + f.adv_flux_f3)


@pytest.mark.parametrize("hemisphere", ["north", "south"])
def test_lwa_set_membership_is_exact(case, hemisphere):
    """Pair counts of the cyclonic / anticyclonic sets (compute_flux_dirinv.f90:90-113) of the default LWA method and
    of the interval scan against the oracle's double loop, and the LWA pieces themselves."""
    from hn2016_falwa_b200 import f90_modules as gpu
    kind, o = case
    c = o["config"]
    eq, jd = c["equator_idx"], c["jd"]
    pv, uu, vv, pt = (_f(o[k]) for k in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"))
    ncf = np.zeros(pv.shape)
    qref_cu, uref, ptref = o["qref"].T, o["uref"].T, o["ptref"].T
    if hemisphere == "north":
        kw = dict(pv=pv, uu=uu, vv=vv, pt=pt, ncforce=ncf, tn0=o["tn0"], qref=qref_cu[-eq:], uref=uref[-jd:],
                  tref=ptref[-jd:])
    else:  # oopinterface.py:776-789
        kw = dict(pv=-pv[:, ::-1, :], uu=uu[:, ::-1, :], vv=-vv[:, ::-1, :], pt=pt[:, ::-1, :], ncforce=ncf[:, ::-1, :],
                  tn0=o["ts0"], qref=-qref_cu[(eq - 1)::-1], uref=uref[(jd - 1)::-1], tref=ptref[(jd - 1)::-1])
    kw.update(jb=c["jb"], is_nhem=True, a=c["a"], om=c["omega"], dz=c["dz"], h=c["h"], rr=c["rr"], cp=c["cp"],
              prefac=c["prefactor"], return_mask_count=True)
    want_cnt = tuple(int(x) for x in o["mask_count"][0 if hemisphere == "north" else 1])
    w = ora.compute_flux_dirinv_nshem(**kw)
    assert tuple(int(x) for x in w[9]) == want_cnt
    for method in gpu.LWA_METHODS:
        g = gpu.compute_flux_dirinv_nshem(method=method, **kw)
        got_cnt = tuple(int(x) for x in g[9])
        _note(what="lwa_mask_count", kind=kind, hemisphere=hemisphere, method=method, gpu=got_cnt, oracle=want_cnt,
              mismatch=[abs(a - b) for a, b in zip(got_cnt, want_cnt)])
        assert got_cnt == want_cnt
        for nm, a_, b_ in zip(("astar1", "astar2", "ncforce3d", "ua1", "ua2"), g[:5], w[:5]):
            rec = compare(f"fullsize/{kind}/{hemisphere}/method{method}/{nm}", a_, b_, rtol=1e-8,
                          atol=1e-11 * float(np.abs(b_).max()))
            _note(what="lwa_piece", kind=kind, hemisphere=hemisphere, method=method, **rec)


@pytest.mark.parametrize("hemisphere", ["north", "south"])
def test_equivalent_latitude_bins_are_exact(case, hemisphere):
    """Histogram bin of every grid point (compute_qref_and_fawa_first.f90:58-142 / compute_reference_states.f90:
    68-100): the number of mismatching assignments is reported and must be 0."""
    from hn2016_falwa_b200 import f90_modules as gpu
    kind, o = case
    c = o["config"]
    pv, uu, av, pt = (_f(o[k]) for k in ("qgpv", "interpolated_u", "interpolated_avort", "interpolated_theta"))
    if hemisphere == "south":
        pv, uu, av, pt = -pv[:, ::-1, :], uu[:, ::-1, :], av[:, ::-1, :], pt[:, ::-1, :]
    if kind == "nhn22":
        kw = dict(pv=pv, uu=uu, vort=av, pt=pt, tn0=o["t0_n"], nd=c["equator_idx"], nnd=c["nlat"], jb=c["jb"],
                  jd=c["jd"], a=c["a"], omega=c["omega"], dz=c["dz"], h=c["h"], dphi=c["dphi"], dlambda=c["dlambda"],
                  rr=c["rr"], cp=c["cp"], return_bins=True)
        g, w = gpu.compute_qref_and_fawa_first(**kw), ora.compute_qref_and_fawa_first(**kw)
        pairs = (("bins_pv", g[7], w[7]), ("bins_vort", g[8], w[8]))
        compare(f"fullsize/{kind}/{hemisphere}/first/qref", g[0], w[0], rtol=1e-8, atol=1e-16)
    else:
        kw = dict(pv=pv, uu=uu, pt=pt, stat=o["static_stability"], jd=c["equator_idx"], npart=c["npart"],
                  maxits=100000, a=c["a"], om=c["omega"], dz=c["dz"], eps=1e-5, h=c["h"], dphi=c["dphi"],
                  dlambda=c["dlambda"], r=c["rr"], cp=c["cp"], rjac=0.95, return_bins=True)
        g, w = gpu.compute_reference_states(**kw), ora.compute_reference_states(**kw)
        pairs = (("bins_pv", g[4], w[4]),)
        assert g[3] == w[3]
    for nm, a_, b_ in pairs:
        mism = int((a_[:, :, 1:-1] != b_[:, :, 1:-1]).sum())
        _note(what="bin_mismatch", kind=kind, hemisphere=hemisphere, which=nm, mismatch=mism,
              points=int(a_[:, :, 1:-1].size))
        assert mism == 0


def test_prefix_sums_on_every_problem_match_the_oracle(snap, case):
    """The benchmarked snapshot again with the prefix-sum LWA kernel (lwa_prefix.cu, method 4) on every (snapshot,
    hemisphere) instead of the automatic choice."""
    kind, want = case
    f = _cls(kind)(*snap)
    f.interpolate_fields(return_named_tuple=False)
    f.compute_reference_states(return_named_tuple=False)
    f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False, method=4)
    for name in ("lwa", "lwa_baro", "u_baro", "adv_flux_f1", "adv_flux_f2", "adv_flux_f3", "convergence_zonal_advective_flux"):
        rec = compare(f"fullsize/{kind}/prefix/{name}", getattr(f, name), want[name], rtol=1e-8,
                      atol=1e-9 * float(np.abs(want[name]).max()))
        _note(what="field_prefix_kernel", kind=kind, grid=[NLON, NLAT], **rec)


def test_upsampled_real_snapshot_matches_the_oracle():
    """the reference's ERA-Interim snapshot interpolated to 1440 x 721.  Its PV contours meander by tens
    of rows (17 x the LWA pairs of the analytic waves), the regime the automatic method hands to the prefix-sum kernel;
    the windowed sums (method 3) must agree as well."""
    from hn2016_falwa_b200.oopinterface import QGFieldNHN22
    from hn2016_falwa_b200.synthetic import load_packed_snapshot, upsampled_snapshot
    x0, y0, plev, u0, v0, t0 = load_packed_snapshot(os.path.join(ROOT, "tests", "golden", "era_interim_2005-01-23_00Z.npz"))
    xlon, ylat, u, v, t = upsampled_snapshot(x0, y0, u0, v0, t0, NLON, NLAT)
    want = pipeline.run_qgfield("nhn22", xlon, ylat, plev, u, v, t)
    _note(what="lwa_pairs_upsampled_real", oracle=[[int(x) for x in row] for row in want["mask_count"]])
    for method in (0, 3):
        f = QGFieldNHN22(xlon, ylat, plev, u, v, t)
        f.interpolate_fields(return_named_tuple=False)
        f.compute_reference_states(return_named_tuple=False)
        f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False, method=method)
        failures = []
        for name in pipeline.USER_FIELDS:
            rtol, atol = TOL.get(name, (1e-8, 1e-9 * float(np.abs(want[name]).max())))
            try:
                rec = compare(f"fullsize/upsampled/method{method}/{name}", getattr(f, name), want[name], rtol=rtol, atol=atol)
            except AssertionError as e:
                failures.append(str(e))
                continue
            _note(what="field_upsampled_real", method=method, grid=[NLON, NLAT], **rec)
        assert not failures, "\n".join(failures)


@pytest.mark.parametrize("method", [0, 4], ids=["automatic", "prefix"])
def test_results_are_bit_reproducible(snap, method):
    """the area histograms and the prefix-sum LWA accumulate 64-bit fixed-point integers, the windowed LWA
    uses no atomics at all — two runs of the pipeline on the same inputs give bit-identical reference states and fluxes."""
    runs = []
    for _ in range(2):
        f = _cls("nhn22")(*snap)
        f.interpolate_fields(return_named_tuple=False)
        f.compute_reference_states(return_named_tuple=False)
        f.compute_lwa_and_barotropic_fluxes(return_named_tuple=False, method=method)
        runs.append({n: np.array(getattr(f, n)) for n in ("qref", "uref", "ptref", "lwa", "lwa_baro", "adv_flux_f2")})
    diff = {n: int((runs[0][n] != runs[1][n]).sum()) for n in runs[0]}
    _note(what="run_to_run_differences", kind="nhn22", method=method, **diff)
    assert not any(diff.values()), diff
