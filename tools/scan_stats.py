"""Event statistics of the interval-scan LWA on the bench's synthetic 0.25-degree snapshot (CPU, oracle fields):
how many rows per column are displaced across a Qref contour, how far, and how many share a bin.  Informs the design
of the running-sum phase of lwa_scan_kernel."""
import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import pipeline
from hn2016_falwa_b200.synthetic import synthetic_snapshot

nlon, nlat = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1440, 721)
snap = synthetic_snapshot(nlon=nlon, nlat=nlat, seed=1234)
o = pipeline.run_qgfield("nhn22", *snap, stages=("interp", "ref"))
qgpv, qref = np.asarray(o["qgpv"]), np.asarray(o["qref"])          # (kmax, nlat, nlon), (kmax, nlat)
nd, jb = nlat // 2 + 1, 5
rows = np.arange(nd - 1, nlat)                                       # northern hemisphere rows, equator -> pole
tot = dict(rows=0, disp=0, cols=0)
dists, mult, warp_multi = [], [], []
for k in (5, 10, 20, 30, 40):
    q = qgpv[k][rows]                                                # (nd, nlon)
    Q = qref[k][rows][jb:nd - 1]                                     # thresholds of rows jb+1 .. nd-1 (1-based)
    Qs = np.maximum.accumulate(Q)                                    # (monotone up to rounding)
    E = np.searchsorted(Qs, q, side="left")                          # thresholds strictly below q
    nxt = np.clip(np.arange(nd)[:, None] - jb + 1 - 1 + 0, 0, Qs.size)  # thresholds at rows <= jj
    nxt = np.clip(np.arange(1, nd + 1)[:, None] - jb, 0, Qs.size)
    d = E - nxt
    disp = d != 0
    tot["rows"] += d.size; tot["disp"] += int(disp.sum()); tot["cols"] += q.shape[1]
    dists.append(np.abs(d[disp]))
    for i in range(0, q.shape[1], 7):
        e = E[disp[:, i], i]
        if e.size:
            c = np.bincount(e, minlength=Qs.size + 1)
            mult.append(c[c > 0])
            # per 32-threshold "warp step" groups (lane-blocked: thresholds lane*R+s): does ANY lane see a bin >= 2?
            R = 13
            cc = np.zeros(32 * R, int); cc[:min(c.size, 32 * R)] = c[:32 * R]
            warp_multi.append((cc.reshape(32, R) >= 2).any(axis=0).mean())
dist = np.concatenate(dists); m = np.concatenate(mult)
print(f"grid {nlon}x{nlat}: displaced rows {100 * tot['disp'] / tot['rows']:.1f}% of all rows")
print("displacement |E - nxt| percentiles 50/90/99/max:", np.percentile(dist, [50, 90, 99]).round(1), dist.max())
print("rows per occupied bin: mean %.2f, share of bins with >= 2 rows %.1f%%, max %d" % (m.mean(), 100 * (m >= 2).mean(), m.max()))
print("share of warp steps in which at least one lane has a bin with >= 2 rows: %.1f%%" % (100 * np.mean(warp_multi)))
