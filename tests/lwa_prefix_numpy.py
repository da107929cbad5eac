"""NumPy restatement of the algorithm of csrc/lwa_prefix.cu (difference arrays over threshold intervals, 64-bit fixed-point
scatter, exception rows) — test infrastructure: tests/test_lwa_prefix_algorithm.py checks it against the oracle's double
loop on the CPU, so the mathematics of the kernel is verified without a GPU."""
import numpy as np


def pow2_scale(maxabs, nd):
    # 2^k with nd * maxabs * 2^k < 2^62
    if maxabs == 0 or not np.isfinite(maxabs): return 1.0
    e = np.frexp(maxabs)[1]          # maxabs < 2^e
    n = int(np.ceil(np.log2(nd + 1)))
    return 2.0 ** (61 - e - n)

def prefix_lwa(pv, uu, nc, qref, uref, jb, a, dphi_nlat):
    """pv, uu, nc: (imax, jmax, kmax) hemispheric-frame Fortran-order views as passed to compute_flux_dirinv_nshem
    (rows: the NH is the last nd rows).  qref (nd,kmax), uref (jd,kmax)."""
    imax, jmax, kmax = pv.shape
    nd = qref.shape[0]
    dp = np.pi / (jmax - 1)
    jstart, jend = jb + 1, nd - 1
    js0, je0 = jstart - 1, jend - 1
    cosphi = np.cos(dp * np.arange(nd))
    aaw = a * dp * cosphi
    ab = a * dp
    a1 = np.zeros((imax, nd, kmax)); a2 = np.zeros_like(a1); u2 = np.zeros_like(a1); ncf = np.zeros_like(a1)
    cnt_c = cnt_a = 0
    nexc = 0
    for k in range(1, kmax - 1):
        q = pv[:, jmax - nd:, k].T.copy()   # [nd, imax]
        u = uu[:, jmax - nd:, k].T.copy()
        n = nc[:, jmax - nd:, k].T.copy()
        Q = qref[:, k]
        ur = np.zeros(nd); ur[js0:je0 + 1] = uref[jstart - jb - 1: jend - jb, k]
        Qt = np.full(nd, np.inf); Qt[:js0] = -np.inf
        pm = np.maximum.accumulate(Q[js0:je0 + 1])
        Qt[js0:je0 + 1] = pm
        exc = np.zeros(nd, bool)
        exc[js0 + 1:je0 + 1] = Q[js0 + 1:je0 + 1] < pm[:-1]
        nexc += int(exc.sum())
        x = np.arange(nd)[:, None]
        E = np.searchsorted(Qt, q.ravel(), side="left").reshape(q.shape)
        E = np.where(np.isnan(q), x + 1, E)
        cyc = E > x + 1
        anti = E <= x
        disp = cyc | anti
        # pair counts over kept thresholds
        excp = np.concatenate([[0], np.cumsum(exc)])   # excp[m] = # exceptions among rows < m
        act_lo = np.maximum(x + 1, js0); 
        lo_c = np.maximum(x + 1, js0); hi_c = np.minimum(E, je0 + 1)        # kept thresholds in [lo, hi)
        cc = np.where(cyc & (hi_c > lo_c), (hi_c - lo_c) - (excp[np.maximum(hi_c, lo_c)] - excp[lo_c]), 0)
        lo_a = np.maximum(E, js0); hi_a = np.minimum(x + 1, je0 + 1)
        ca = np.where(anti & (hi_a > lo_a), (hi_a - lo_a) - (excp[np.maximum(hi_a, lo_a)] - excp[lo_a]), 0)
        cnt_c += int(cc.sum()); cnt_a += int(ca.sum())
        f = [q * aaw[:, None], np.broadcast_to(aaw[:, None], q.shape).copy(), q * u, u.copy(), n * aaw[:, None]]
        sc = [pow2_scale(np.nanmax(np.abs(np.where(disp, v, 0.0))) if disp.any() else 0.0, nd) for v in f]
        fx = [np.where(disp, np.rint(np.where(disp, v, 0.0) * s), 0).astype(np.int64) for v, s in zip(f, sc)]
        D = {name: np.zeros((nd + 1, imax), np.int64) for name in ("c1", "c2", "a1", "a2", "t3", "t4", "t5")}
        cols = np.broadcast_to(np.arange(imax)[None, :], q.shape)
        def add(name, pos, val, mask):
            np.add.at(D[name], (pos[mask], cols[mask]), val[mask])
        xp1 = np.broadcast_to(x + 1, q.shape)
        add("c1", xp1, fx[0], cyc); add("c1", E, -fx[0], cyc)
        add("c2", xp1, fx[1], cyc); add("c2", E, -fx[1], cyc)
        add("a1", E, fx[0], anti); add("a1", xp1, -fx[0], anti)
        add("a2", E, fx[1], anti); add("a2", xp1, -fx[1], anti)
        for nm, v in (("t3", fx[2]), ("t4", fx[3]), ("t5", fx[4])):
            add(nm, xp1, v, disp); add(nm, E, -v, disp)
        S = {nm: np.cumsum(D[nm], axis=0)[:nd] for nm in D}
        S1c, S2c, S1a, S2a, T3, T4, T5 = (S["c1"] / sc[0], S["c2"] / sc[1], S["a1"] / sc[0], S["a2"] / sc[1],
                                            S["t3"] / sc[2], S["t4"] / sc[3], S["t5"] / sc[4])
        Qc = Q[:, None]
        A1 = S1c - Qc * S2c
        A2 = Qc * S2a - S1a
        B = T3 - Qc * T4
        U2 = ab * cosphi[:, None] * B - ur[:, None] * (A1 + A2)
        act = np.zeros(nd, bool); act[js0:je0 + 1] = True
        keep = act & ~exc
        a1[:, keep, k] = A1[keep].T; a2[:, keep, k] = A2[keep].T; u2[:, keep, k] = U2[keep].T; ncf[:, keep, k] = T5[keep].T
        for j0 in np.nonzero(exc)[0]:
            qe = q - Q[j0]
            mc = (qe > 0) & (x < j0); ma = (qe <= 0) & (x >= j0)
            cnt_c += int(mc.sum()); cnt_a += int(ma.sum())
            A1e = (np.where(mc, qe, 0) * aaw[:, None]).sum(0); A2e = (np.where(ma, -qe, 0) * aaw[:, None]).sum(0)
            Be = (np.where(mc, qe, 0) * u).sum(0) - (np.where(ma, qe, 0) * u).sum(0)
            Ne = (np.where(mc, n, 0) * aaw[:, None]).sum(0) - (np.where(ma, n, 0) * aaw[:, None]).sum(0)
            a1[:, j0, k] = A1e; a2[:, j0, k] = A2e; u2[:, j0, k] = ab * cosphi[j0] * Be - ur[j0] * (A1e + A2e); ncf[:, j0, k] = Ne
    return a1, a2, ncf, u2, (cnt_a, cnt_c), nexc
