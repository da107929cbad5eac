"""CPU tests of the host logic of QGDataset (hn2016_falwa_b200/xarrayinterface.py): flattening of the leading
dimensions, automatic latitude / pressure flips (reference xarrayinterface.py:376-410), contiguous time shards and the
final gather over a world_size-2 gloo process group.  The chunk runner is replaced by one backed by the CPU oracle
(test infrastructure), so no GPU is needed; the device runner is covered by tests/test_gpu_dataset.py."""
import os
import socket

import numpy as np
import pytest

from hn2016_falwa_b200 import xarrayinterface as xi
from hn2016_falwa_b200.synthetic import synthetic_snapshot


class _Template:
    """The attributes QGDataset reads from its template field."""
    kmax, dz, maxit, tol, npart, rjac = 13, 1000., 1000, 1e-5, None, 0.95
    scale_height, cp, dry_gas_constant, omega, planet_radius, prefactor = 7000., 1004., 287., 7.29e-5, 6.378e6, 1.0
    need_latitude_interpolation = False

    def __init__(self, xlon, ylat, plev, *_, **__):
        self.xlon, self.ylat, self.ylat_ref_states = np.asarray(xlon), np.asarray(ylat), np.asarray(ylat)
        self.height = np.arange(self.kmax) * self.dz
        assert np.all(np.diff(ylat) > 0) and np.all(np.diff(plev) < 0), "QGDataset must hand over flipped coordinates"


class OracleRunner:
    """Runs each snapshot through the CPU oracle pipeline (NH18, small kmax)."""

    def __init__(self, xlon, ylat, plev):
        self.c = (xlon, ylat, plev)

    def run(self, u, v, t, stages, want, ncforce=None):
        from oracle import pipeline
        for i in range(u.shape[0]):
            o = pipeline.run_qgfield("nhn22", *self.c, u[i], v[i], t[i], kmax=_Template.kmax, stages=tuple(stages),
                                     eq_boundary_index=2)
            out = {}
            for n in want:
                val = o[n]
                out[n] = np.stack(val)[None] if isinstance(val, tuple) else np.asarray(val)[None]
            yield i, i + 1, out


def _dataset(nt=5, flip=True):
    snaps = [synthetic_snapshot(nlon=24, nlat=13, seed=7, t_index=k) for k in range(nt)]
    xlon, ylat, plev = snaps[0][:3]
    u, v, t = (np.stack([s[i] for s in snaps]) for i in (3, 4, 5))
    if flip:   # hand the data over north->south and top->bottom, like ERA files
        u, v, t = (x[:, ::-1, ::-1, :] for x in (u, v, t))
        ylat_in, plev_in = ylat[::-1], plev[::-1]
    else:
        ylat_in, plev_in = ylat, plev
    coords = dict(time=np.arange(nt), level=plev_in, latitude=ylat_in, longitude=xlon)
    dims = ("time", "level", "latitude", "longitude")
    mk = lambda a, n: xi.SimpleDataArray(a, dims, coords, name=n)
    return xi.SimpleDataset(dict(u=mk(u, "u"), v=mk(v, "v"), t=mk(t, "t"))), (xlon, ylat, plev), (u, v, t)


def test_shard_bounds_cover_everything_once():
    for n in (1, 2, 7, 1460):
        for world in (1, 2, 3, 8):
            b = [xi.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_flips_and_flattening_match_per_snapshot_oracle():
    ds, (xlon, ylat, plev), _ = _dataset(nt=3, flip=True)
    q = xi.QGDataset(ds, qgfield=_Template, runner=OracleRunner(xlon, ylat, plev))
    out = q.compute_reference_states()
    assert set(out) == {"qref", "uref", "ptref"}
    assert out["qref"].dims == ("time", "height", "ylat") and out["qref"].shape == (3, 13, 13)
    from oracle import pipeline
    snap = synthetic_snapshot(nlon=24, nlat=13, seed=7, t_index=1)
    want = pipeline.run_qgfield("nhn22", *snap, kmax=13, stages=("interp", "ref"), eq_boundary_index=2)
    np.testing.assert_array_equal(out["qref"].data[1], want["qref"])
    flux = q.compute_lwa_and_barotropic_fluxes()
    assert flux["lwa"].shape == (3, 13, 13, 24) and flux["lwa_baro"].dims == ("time", "ylat", "xlon")
    assert flux.attrs["protocol"] == "_Template"
    with pytest.raises(AssertionError):   # transposed input is rejected (reference test_xarrayinterface.py:24-59)
        bad = xi.SimpleDataArray(np.zeros((2, 3, 4, 5)), ("time", "latitude", "level", "longitude"),
                                 dict(time=np.arange(2), latitude=np.arange(3.), level=np.arange(4.),
                                      longitude=np.arange(5.)))
        xi.QGDataset(bad, bad, bad, qgfield=_Template, runner=OracleRunner(xlon, ylat, plev))


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds, (xlon, ylat, plev), _ = _dataset(nt=5, flip=True)
    q = xi.QGDataset(ds, qgfield=_Template, runner=OracleRunner(xlon, ylat, plev))   # shard from the process group
    # the usual multi-rank sequence: a stage, gather, the next stage, gather — the second stage must still work on
    # this rank's own rows and the second gather must pick up the new variables (and leave the old ones intact)
    q.compute_reference_states(return_dataset=False)
    q.gather()
    qref_first = q.qref.data.copy()
    assert q.shard == xi.shard_bounds(5, rank, world)
    q.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    lo, hi = q.shard
    assert q.lwa_baro.data.shape[0] == hi - lo          # not gathered yet: this rank's rows only
    q.gather()
    q.gather()                                           # nothing new: no-op
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), lo=lo, hi=hi, lwa_baro=q.lwa_baro.data, uref=q.uref.data,
             lwa=q.lwa.data, qref=q.qref.data, qref_first=qref_first)
    dist.destroy_process_group()


def test_two_rank_gloo_shards_and_gather(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in (0, 1))
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 3, 3, 5)
    ds, (xlon, ylat, plev), _ = _dataset(nt=5, flip=True)
    single = xi.QGDataset(ds, qgfield=_Template, runner=OracleRunner(xlon, ylat, plev), shard=(0, 1))
    single.compute_lwa_and_barotropic_fluxes(return_dataset=False)
    for r in (r0, r1):   # every rank holds the complete, correctly ordered result
        np.testing.assert_array_equal(r["lwa_baro"], single.lwa_baro.data)
        np.testing.assert_array_equal(r["uref"], single.uref.data)
        np.testing.assert_array_equal(r["lwa"], single.lwa.data)
        np.testing.assert_array_equal(r["qref"], single.qref.data)
        np.testing.assert_array_equal(r["qref_first"], single.qref.data)
