"""GPU parity tests of the nine F2PY-compatible entry points (through the C ABI) against the CPU oracle,
on the reference repo's ERA-Interim snapshot (configs 1 and 2 of BASELINE.json) and on a seeded synthetic
snapshot.  Tolerances: rtol 1e-8 as BASELINE.json's north_star states, plus an atol tied to the field's
magnitude where a quantity is a difference of large sums; equivalent-latitude bins and LWA masks must agree
exactly (the mismatch count is reported and must be 0)."""
import numpy as np
import pytest

from oracle import f2py_shim as ora
from oracle import pipeline, testdata
from tests.parity import compare

pytestmark = pytest.mark.gpu

A, OM, DZ, H, RR, CP = pipeline.EARTH_RADIUS, pipeline.EARTH_OMEGA, 1000., 7000., 287., 1004.


@pytest.fixture(scope="module")
def gpu():
    from hn2016_falwa_b200 import f90_modules
    return f90_modules


@pytest.fixture(scope="module", params=["era", "synthetic"])
def staged(request):
    """Oracle results of every stage for NHN22 and NH18 (inputs of the per-routine tests)."""
    if request.param == "era":
        inputs = testdata.load_era_snapshot()
    else:
        inputs = testdata.synthetic_snapshot(nlon=96, nlat=49, seed=3)
    o22 = pipeline.run_qgfield("nhn22", *inputs, keep_aux=True)
    o18 = pipeline.run_qgfield("nh18", *inputs)
    return request.param, inputs, o18, o22


def _f(a):  # python (kmax,nlat,nlon) -> fortran-order view
    return np.swapaxes(a, 0, 2)


def test_compute_qgpv_nh18(gpu, staged):
    name, inputs, o18, _ = staged
    height = o18["config"]["height"]
    t0 = 0.5 * (o18["t0_s"] + o18["t0_n"])
    stat = o18["static_stability"]
    args = (_f(o18["interpolated_u"]), _f(o18["interpolated_v"]), _f(o18["interpolated_theta"]), height, t0, stat,
            A, OM, DZ, H, RR, CP)
    pv, av = gpu.compute_qgpv(*args)
    pv0, av0 = ora.compute_qgpv(*args)
    compare(f"{name}/nh18/avort", av, av0, rtol=1e-12, atol=1e-19)
    compare(f"{name}/nh18/pv", pv, pv0, rtol=1e-10, atol=1e-17)


def test_compute_qgpv_direct_inv(gpu, staged):
    name, inputs, _, o22 = staged
    c = o22["config"]
    args = (c["equator_idx"], _f(o22["interpolated_u"]), _f(o22["interpolated_v"]), _f(o22["interpolated_theta"]),
            c["height"], o22["t0_s"], o22["t0_n"], o22["stat_s"], o22["stat_n"], A, OM, DZ, H, RR, CP)
    pv, av = gpu.compute_qgpv_direct_inv(*args)
    pv0, av0 = ora.compute_qgpv_direct_inv(*args)
    r = compare(f"{name}/nhn22/avort", av, av0, rtol=1e-12, atol=1e-19)
    compare(f"{name}/nhn22/pv", pv, pv0, rtol=1e-12, atol=1e-19)
    # interior rows use no reductions: they must be bit-identical to the oracle
    assert np.array_equal(av[:, 1:-1, :], av0[:, 1:-1, :])
    assert np.array_equal(pv[:, 1:-1, :], pv0[:, 1:-1, :])


@pytest.mark.parametrize("hemisphere", ["north", "south"])
def test_compute_qref_and_fawa_first(gpu, staged, hemisphere):
    name, inputs, _, o22 = staged
    c = o22["config"]
    pv, uu, av, pt = (_f(o22[k]) for k in ("qgpv", "interpolated_u", "interpolated_avort", "interpolated_theta"))
    if hemisphere == "south":  # oopinterface.py:1800-1806
        pv, uu, av, pt = -pv[:, ::-1, :], uu[:, ::-1, :], av[:, ::-1, :], pt[:, ::-1, :]
    nd = c["nlat"] // 2 + 1
    kw = dict(pv=pv, uu=uu, vort=av, pt=pt, tn0=o22["t0_n"], nd=nd, nnd=c["nlat"], jb=c["jb"], jd=c["jd"], a=A,
              omega=OM, dz=DZ, h=H, dphi=c["dphi"], dlambda=c["dlambda"], rr=RR, cp=CP, return_bins=True)
    got = gpu.compute_qref_and_fawa_first(**kw)
    want = ora.compute_qref_and_fawa_first(**kw)
    names = ("qref", "ubar", "tbar", "fawa", "ckref", "tjk", "sjk")
    # equivalent-latitude bin assignment: exact
    for nm, g, w in (("bins_pv", got[7], want[7]), ("bins_vort", got[8], want[8])):
        mism = int((g[:, :, 1:-1] != w[:, :, 1:-1]).sum())
        print(f"{name}/{hemisphere}/{nm}: {mism} mismatching bin assignments of {g[:, :, 1:-1].size}")
        assert mism == 0
    tag = f"{name}/nhn22/{hemisphere}/first/"
    compare(tag + "qref", got[0], want[0], rtol=1e-8, atol=1e-16)
    compare(tag + "ubar", got[1], want[1], rtol=1e-10, atol=1e-12)
    compare(tag + "tbar", got[2], want[2], rtol=1e-12)
    fscale = np.abs(want[3]).max()
    compare(tag + "fawa", got[3], want[3], rtol=1e-8, atol=1e-9 * fscale)
    compare(tag + "ckref", got[4], want[4], rtol=1e-8, atol=1e-9 * np.abs(want[4]).max())
    compare(tag + "tjk", got[5], want[5], rtol=1e-8, atol=1e-9 * np.abs(want[5]).max())
    assert np.array_equal(got[6], want[6])


def test_nhn22_block_recursion_pieces(gpu, staged):
    """matrix_b4_inversion / matrix_after_inversion / upward_sweep driven exactly like oopinterface.py:1839-1885."""
    from scipy.linalg.lapack import dgetrf, dgetri
    name, inputs, _, o22 = staged
    c = o22["config"]
    aux = o22["aux"]["n"]
    kmax, nlat = c["kmax"], c["nlat"]
    z = np.arange(0, kmax * DZ, DZ)
    first = ora.compute_qref_and_fawa_first(
        pv=_f(o22["qgpv"]), uu=_f(o22["interpolated_u"]), vort=_f(o22["interpolated_avort"]),
        pt=_f(o22["interpolated_theta"]), tn0=o22["t0_n"], nd=nlat // 2 + 1, nnd=nlat, jb=c["jb"], jd=c["jd"], a=A,
        omega=OM, dz=DZ, h=H, dphi=c["dphi"], dlambda=c["dlambda"], rr=RR, cp=CP)
    qref_os, ckref = first[0], first[4]
    sj_g, tj_g = first[6].copy(order="F"), first[5].copy(order="F")
    sj_o, tj_o = first[6].copy(order="F"), first[5].copy(order="F")
    for k in range(kmax - 1, 1, -1):
        kw = dict(k=k, jmax=nlat, jb=c["jb"], jd=c["jd"], z=z, statn=o22["stat_n"], qref=qref_os, ckref=ckref, a=A,
                  om=OM, dz=DZ, h=H, rr=RR, cp=CP)
        g = gpu.matrix_b4_inversion(sjk=sj_o, **kw)   # same inputs for both at every level
        w = ora.matrix_b4_inversion(sjk=sj_o, **kw)
        if k in (kmax - 1, kmax // 2, 2):
            for nm, a_, b_ in zip(("qjj", "djj", "cjj", "rj"), g, w):
                compare(f"{name}/b4/k{k}/{nm}", a_, b_, rtol=1e-13, atol=0)
        qjj, djj, cjj, rj = w
        lu, piv, info = dgetrf(qjj)
        qinv, info = dgetri(lu, piv)
        sj_g[...] = sj_o
        tj_g[...] = tj_o
        gpu.matrix_after_inversion(k=k, qjj=qinv, djj=djj, cjj=cjj, rj=rj, sjk=sj_g, tjk=tj_g)
        ora.matrix_after_inversion(k=k, qjj=qinv, djj=djj, cjj=cjj, rj=rj, sjk=sj_o, tjk=tj_o)
        if k in (kmax - 1, kmax // 2, 2):
            compare(f"{name}/after/k{k}/sjk", sj_g[:, :, k - 2], sj_o[:, :, k - 2], rtol=1e-13)
            compare(f"{name}/after/k{k}/tjk", tj_g[:, k - 2], tj_o[:, k - 2], rtol=1e-12,
                    atol=1e-13 * np.abs(tj_o[:, k - 2]).max())
    kw = dict(jmax=nlat, jb=c["jb"], sjk=sj_o, tjk=tj_o, ckref=ckref, tb=o22["t0_n"], qref_over_cor=qref_os, a=A,
              om=OM, dz=DZ, h=H, rr=RR, cp=CP)
    g = gpu.upward_sweep(**kw)
    w = ora.upward_sweep(**kw)
    compare(f"{name}/sweep/tref", g[0], w[0], rtol=1e-10)
    compare(f"{name}/sweep/qref", g[1], w[1], rtol=1e-13, atol=0)
    compare(f"{name}/sweep/u", g[2], w[2], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("hemisphere", ["north", "south"])
def test_compute_reference_states_nh18(gpu, staged, hemisphere):
    name, inputs, o18, _ = staged
    c = o18["config"]
    pv, uu, pt = (_f(o18[k]) for k in ("qgpv", "interpolated_u", "interpolated_theta"))
    if hemisphere == "south":
        pv, uu, pt = -pv[:, ::-1, :], uu[:, ::-1, :], pt[:, ::-1, :]
    kw = dict(pv=pv, uu=uu, pt=pt, stat=o18["static_stability"], jd=c["equator_idx"], npart=c["npart"],
              maxits=100000, a=A, om=OM, dz=DZ, eps=1e-5, h=H, dphi=c["dphi"], dlambda=c["dlambda"], r=RR, cp=CP,
              rjac=0.95, return_bins=True)
    g = gpu.compute_reference_states(**kw)
    w = ora.compute_reference_states(**kw)
    mism = int((g[4][:, :, 1:-1] != w[4][:, :, 1:-1]).sum())
    print(f"{name}/nh18/{hemisphere}: bins mismatching {mism}; SOR iterations gpu={g[3]} oracle={w[3]}")
    assert mism == 0
    assert g[3] == w[3]
    tag = f"{name}/nh18/{hemisphere}/ref/"
    compare(tag + "qref", g[0], w[0], rtol=1e-8, atol=1e-12)
    compare(tag + "uref", g[1], w[1], rtol=1e-8, atol=1e-8)
    compare(tag + "tref", g[2], w[2], rtol=1e-9)


@pytest.mark.parametrize("method", [0, 1, 2, 4], ids=["window", "bruteforce", "scan", "prefix"])
@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
@pytest.mark.parametrize("hemisphere", ["north", "south"])
def test_compute_flux_dirinv_nshem(gpu, staged, kind, hemisphere, method):
    name, inputs, o18, o22 = staged
    o = o18 if kind == "nh18" else o22
    c = o["config"]
    eq, jd = c["equator_idx"], c["jd"]
    pv, uu, vv, pt = (_f(o[k]) for k in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"))
    rng = np.random.default_rng(5)
    ncf = rng.standard_normal(pv.shape) * 1e-9
    qref_cu, uref, ptref = o["qref"].T, o["uref"].T, o["ptref"].T   # (nlat, kmax)
    if hemisphere == "north":
        kw = dict(pv=pv, uu=uu, vv=vv, pt=pt, ncforce=ncf, tn0=o["tn0"], qref=qref_cu[-eq:], uref=uref[-jd:],
                  tref=ptref[-jd:])
    else:  # oopinterface.py:776-789
        kw = dict(pv=-pv[:, ::-1, :], uu=uu[:, ::-1, :], vv=-vv[:, ::-1, :], pt=pt[:, ::-1, :], ncforce=ncf[:, ::-1, :],
                  tn0=o["ts0"], qref=-qref_cu[(eq - 1)::-1], uref=uref[(jd - 1)::-1], tref=ptref[(jd - 1)::-1])
    kw.update(jb=c["jb"], is_nhem=True, a=A, om=OM, dz=DZ, h=H, rr=RR, cp=CP, prefac=c["prefactor"],
              return_mask_count=True)
    g = gpu.compute_flux_dirinv_nshem(method=method, **kw)
    w = ora.compute_flux_dirinv_nshem(**kw)
    print(f"{name}/{kind}/{hemisphere}/method{method}: LWA mask counts gpu={g[9]} oracle={w[9]}")
    assert g[9] == w[9]   # number of (i, j, jj, k) pairs in the anticyclonic / cyclonic sets: exact
    names = ("astar1", "astar2", "ncforce3d", "ua1", "ua2", "ep1", "ep2", "ep3", "ep4")
    # brute force keeps the reference's summation order; the interval scan sums in a different order
    # (differences of prefix sums), hence the looser absolute term
    rtol, ascale = (1e-12, 1e-14) if method == 1 else (1e-8, 1e-11)
    for nm, a_, b_ in zip(names, g[:9], w[:9]):
        compare(f"{name}/{kind}/{hemisphere}/flux{method}/{nm}", a_, b_, rtol=rtol, atol=ascale * np.abs(b_).max())


@pytest.mark.parametrize("method", [0, 1, 2, 4], ids=["window", "bruteforce", "scan", "prefix"])
def test_compute_lwa_only_nhn22(gpu, staged, method):
    name, inputs, o18, _ = staged
    c = o18["config"]
    eq = c["equator_idx"]
    kw = dict(pv=_f(o18["qgpv"]), uu=_f(o18["interpolated_u"]), qref=o18["qref"].T[-eq:], jb=0, is_nhem=True, a=A,
              om=OM, dz=DZ, h=H, rr=RR, cp=CP, prefac=c["prefactor"])
    g = gpu.compute_lwa_only_nhn22(method=method, **kw)
    w = ora.compute_lwa_only_nhn22(**kw)
    rtol, ascale = (1e-12, 1e-14) if method == 1 else (1e-8, 1e-11)
    for nm, a_, b_ in zip(("astarbaro", "ubaro", "astar1", "astar2"), g, w):
        compare(f"{name}/lwa_only{method}/{nm}", a_, b_, rtol=rtol, atol=ascale * np.abs(b_).max())


@pytest.mark.parametrize("kind", ["nh18", "nhn22"])
def test_drop_in_pipeline_through_f2py_boundary(gpu, staged, kind):
    """The whole reference pipeline (host restatement of oopinterface.py) with the nine native names bound to
    the CUDA entry points, against the same pipeline on the CPU oracle."""
    name, inputs, o18, o22 = staged
    want = o18 if kind == "nh18" else o22
    with pipeline.use_native(gpu):
        got = pipeline.run_qgfield(kind, *inputs)
    if kind == "nh18":
        assert got["num_of_iter"] == want["num_of_iter"]
    assert got["mask_count"] == want["mask_count"]
    tol = dict(qgpv=(1e-10, 1e-17), qref=(1e-8, 1e-14), uref=(1e-8, 1e-7), ptref=(1e-9, 0.0))
    for f in pipeline.USER_FIELDS:
        rtol, atol = tol.get(f, (1e-8, 1e-9 * float(np.abs(want[f]).max())))
        compare(f"{name}/{kind}/dropin/{f}", got[f], want[f], rtol=rtol, atol=atol)


@pytest.mark.parametrize("method", [0, 2, 4], ids=["window", "scan", "prefix"])
def test_lwa_with_nan_points_matches_the_double_loop(gpu, method):
    """A NaN PV value makes both of the reference's comparisons false (compute_flux_dirinv.f90:90, :101): the point
    belongs to no set of any row.  The fast LWA methods must treat it the same way as the double loop (method 1), and
    a NaN in u may only reach the rows whose sets contain that point."""
    inputs = testdata.synthetic_snapshot(nlon=96, nlat=49, seed=3)
    o = pipeline.run_qgfield("nhn22", *inputs, stages=("interp", "ref"))
    c = o["config"]
    eq, jd = c["equator_idx"], c["jd"]
    pv, uu, vv, pt = (np.array(_f(o[k])) for k in ("qgpv", "interpolated_u", "interpolated_v", "interpolated_theta"))
    rng = np.random.default_rng(11)
    for _ in range(40):
        pv[rng.integers(0, 96), rng.integers(24, 49), rng.integers(2, 47)] = np.nan
    if method == 0:   # (the interval scan forms its sums as differences of prefix sums: a NaN in u, once added, cannot
        for _ in range(10):   # leave again — methods 2 and 4 are exact for NaN PV only; the default method has no such limit)
            uu[rng.integers(0, 96), rng.integers(24, 49), rng.integers(2, 47)] = np.nan
    kw = dict(pv=pv, uu=uu, vv=vv, pt=pt, ncforce=np.zeros(pv.shape), tn0=o["tn0"], qref=o["qref"].T[-eq:],
              uref=o["uref"].T[-jd:], tref=o["ptref"].T[-jd:], jb=c["jb"], is_nhem=True, a=A, om=OM, dz=DZ, h=H,
              rr=RR, cp=CP, prefac=c["prefactor"], return_mask_count=True)
    want = gpu.compute_flux_dirinv_nshem(method=1, **kw)
    got = gpu.compute_flux_dirinv_nshem(method=method, **kw)
    assert got[9] == want[9]
    for nm, a_, b_ in zip(("astar1", "astar2", "ncforce3d", "ua1", "ua2"), got[:5], want[:5]):
        assert np.array_equal(np.isnan(a_), np.isnan(b_)), nm
        ok = ~np.isnan(b_)
        compare(f"nan/method{method}/{nm}", a_[ok], b_[ok], rtol=1e-8, atol=1e-11 * np.abs(b_[ok]).max())
