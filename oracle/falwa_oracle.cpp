// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU oracle for the QGField hot path: a C++ restatement, templated on float /
// double, of the nine F2PY Fortran routines of csyhuang/hn2016_falwa (falwa
// v2.3.3).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product (hn2016_falwa_b200)
// never does.
//
// Parity status: "parity unpinned" for the 3-D routines — the reference's own
// golden vectors for them are absent from /root/reference (.MISSING_LARGE_BLOBS)
// and no Fortran compiler exists in this image, so the restatement can only be
// pinned to (a) the reference's *Python* driver, which runs unmodified on top
// of this library (oracle/gen_golden.py), and (b) the structural assertions of
// the reference's tests.  See DESIGN.md §3.
//
// Conventions
//   * Arrays are Fortran order (column-major), 1-based in the comments; the
//     F3/F2 accessors below do the index arithmetic.
//   * R = double is the parity oracle (== `gfortran -fdefault-real-8`);
//     R = float is the arithmetic as shipped (bare REAL).
//   * Every intent(out) array is zero-filled first (the reference leaves
//     several regions undefined — SURVEY.md Appendix A, U1..U7).
//   * Loop-invariant trigonometric factors are hoisted out of inner loops;
//     the values and the order of the floating-point operations that produce
//     each output are those of the cited Fortran statements.
//
// Each function cites the reference file:line it follows
// (paths relative to /root/reference/src/falwa/f90_modules/).

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

template <typename R> inline R rpi() { return std::acos(R(-1)); }

// 1-based column-major accessors
#define F2(a, i, j, n1) (a)[(size_t)((i)-1) + (size_t)(n1) * (size_t)((j)-1)]
#define F3(a, i, j, k, n1, n2) \
    (a)[(size_t)((i)-1) + (size_t)(n1) * ((size_t)((j)-1) + (size_t)(n2) * (size_t)((k)-1))]

template <typename R> inline void zero(R* p, size_t n) { std::fill(p, p + n, R(0)); }

// ---------------------------------------------------------------------------
// Absolute vorticity shared by compute_qgpv.f90:59-94 and
// compute_qgpv_direct_inv.f90:28-60.  `deg_lat` selects how latitude is built
// (degrees->radians in NH18, radians in NHN22).
// ---------------------------------------------------------------------------
template <typename R>
void avort_common(int nlon, int nlat, int kmax, const R* u, const R* v, R aa, R omega, bool deg_lat,
                  R* avort) {
    const R pi = rpi<R>();
    const R dphi = pi / R(nlat - 1);
    auto lat = [&](int jm1) -> R {  // latitude of (1-based) row jm1+1
        if (deg_lat) {
            R p = R(-90.) + R(jm1) * R(180.) / R(nlat - 1);
            return p * pi / R(180.);
        }
        return R(-0.5) * pi + pi * R(jm1) / R(nlat - 1);
    };
    for (int kk = 1; kk <= kmax; ++kk) {
        for (int j = 2; j <= nlat - 1; ++j) {
            const R phi0 = lat(j - 1), phim = lat(j - 2), phip = lat(j);
            const R s0 = std::sin(phi0), c0 = std::cos(phi0), cp = std::cos(phip), cm = std::cos(phim);
            const R den = R(2.) * aa * c0 * dphi;
            for (int i = 1; i <= nlon; ++i) {
                const int ip = (i == nlon) ? 1 : i + 1;
                const int im = (i == 1) ? nlon : i - 1;
                const R av1 = R(2.) * omega * s0;
                const R av2 = (F3(v, ip, j, kk, nlon, nlat) - F3(v, im, j, kk, nlon, nlat)) / den;
                const R av3 = -(F3(u, i, j + 1, kk, nlon, nlat) * cp - F3(u, i, j - 1, kk, nlon, nlat) * cm) / den;
                F3(avort, i, j, kk, nlon, nlat) = av1 + av2 + av3;
            }
        }
        R avs = 0, avn = 0;
        for (int i = 1; i <= nlon; ++i) {
            avs = avs + F3(avort, i, 2, kk, nlon, nlat) / R(nlon);
            avn = avn + F3(avort, i, nlat - 1, kk, nlon, nlat) / R(nlon);
        }
        for (int i = 1; i <= nlon; ++i) {
            F3(avort, i, 1, kk, nlon, nlat) = avs;
            F3(avort, i, nlat, kk, nlon, nlat) = avn;
        }
    }
}

// compute_qgpv.f90:1-137 (NH18).  pv(:,:,1) and pv(:,:,kmax) are left 0 (U1).
template <typename R>
void compute_qgpv(int nlon, int nlat, int kmax, const R* ut, const R* vt, const R* theta, const R* height,
                  const R* t0, const R* stat, R aa, R omega, R /*dz*/, R hh, R /*rr*/, R /*cp*/, R* pv,
                  R* avort) {
    const size_t n3 = (size_t)nlon * nlat * kmax;
    zero(pv, n3);
    zero(avort, n3);
    avort_common<R>(nlon, nlat, kmax, ut, vt, aa, omega, /*deg_lat=*/true, avort);
    // zonal-mean absolute vorticity, compute_qgpv.f90:98-105
    std::vector<R> zmav((size_t)nlat * kmax, R(0));
    for (int kk = 1; kk <= kmax; ++kk)
        for (int j = 1; j <= nlat; ++j) {
            R s = 0;
            for (int i = 1; i <= nlon; ++i) s = s + F3(avort, i, j, kk, nlon, nlat) / R(nlon);
            F2(zmav, j, kk, nlat) = s;
        }
    // interior pv, compute_qgpv.f90:109-124
    for (int kk = 2; kk <= kmax - 1; ++kk) {
        const R ep = std::exp(-height[kk] / hh), em = std::exp(-height[kk - 2] / hh);
        const R e0 = std::exp(height[kk - 1] / hh);
        const R dh = height[kk] - height[kk - 2];
        for (int j = 1; j <= nlat; ++j) {
            const R zm = F2(zmav, j, kk, nlat);
            for (int i = 1; i <= nlon; ++i) {
                const R altp = ep * (F3(theta, i, j, kk + 1, nlon, nlat) - t0[kk]) / stat[kk];
                const R altm = em * (F3(theta, i, j, kk - 1, nlon, nlat) - t0[kk - 2]) / stat[kk - 2];
                const R strc = (altp - altm) * zm / dh;
                F3(pv, i, j, kk, nlon, nlat) = F3(avort, i, j, kk, nlon, nlat) + e0 * strc;
            }
        }
    }
}

// compute_qgpv_direct_inv.f90:1-112 (NHN22).
template <typename R>
void compute_qgpv_direct_inv(int nlon, int nlat, int kmax, int jd, const R* uq, const R* vq, const R* tq,
                             const R* height, const R* ts0, const R* tn0, const R* stats, const R* statn,
                             R aa, R omega, R /*dz*/, R hh, R /*rr*/, R /*cp*/, R* pv, R* avort) {
    const size_t n3 = (size_t)nlon * nlat * kmax;
    zero(pv, n3);
    zero(avort, n3);
    avort_common<R>(nlon, nlat, kmax, uq, vq, aa, omega, /*deg_lat=*/false, avort);
    const R pi = rpi<R>();
    // interior pv, compute_qgpv_direct_inv.f90:75-100
    for (int kk = 2; kk <= kmax - 1; ++kk) {
        const R ep = std::exp(-height[kk] / hh), em = std::exp(-height[kk - 2] / hh);
        const R e0 = std::exp(height[kk - 1] / hh);
        const R dh = height[kk] - height[kk - 2];
        for (int j = 1; j <= nlat; ++j) {
            const R phi0 = R(-0.5) * pi + pi * R(j - 1) / R(nlat - 1);
            const R f = R(2.) * omega * std::sin(phi0);
            R statp, statm, t00p, t00m;
            if (j <= jd) {
                statp = stats[kk]; statm = stats[kk - 2]; t00p = ts0[kk]; t00m = ts0[kk - 2];
            } else {
                statp = statn[kk]; statm = statn[kk - 2]; t00p = tn0[kk]; t00m = tn0[kk - 2];
            }
            for (int i = 1; i <= nlon; ++i) {
                const R altp = ep * (F3(tq, i, j, kk + 1, nlon, nlat) - t00p) / statp;
                const R altm = em * (F3(tq, i, j, kk - 1, nlon, nlat) - t00m) / statm;
                const R strc = (altp - altm) * f / dh;
                F3(pv, i, j, kk, nlon, nlat) = F3(avort, i, j, kk, nlon, nlat) + e0 * strc;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Area analysis of one level (compute_reference_states.f90:74-106 and
// compute_qref_and_fawa_first.f90:61-96 / :109-141): histogram `fld` into
// nbins bins counted from its maximum downward, cumulative area `aan` and
// circulation `ccn` starting at bin 2, then for each row j (1..nrow-1) keep the
// LAST bin nn with aan(nn) <= alat(j) < aan(nn+1).  Optional outputs receive
// the interpolated bin value (qout) and circulation (cout).  `ind_out`, when
// given, receives the 1-based bin of every grid point (for exact bin-parity
// tests).
// ---------------------------------------------------------------------------
template <typename R>
void area_analysis(int imax, int jmax, const R* fld /*(imax,jmax)*/, int nbins, int nrow, const R* alat,
                   R a, R dphi, R dlambda, R* qout /*stride kstride*/, R* cout, R* qmax_out,
                   std::vector<R>& qn, std::vector<R>& an, std::vector<R>& cn, std::vector<R>& aan,
                   std::vector<R>& ccn, int32_t* ind_out) {
    const R pi = rpi<R>();
    R qmax = fld[0], qmin = fld[0];
    const size_t n2 = (size_t)imax * jmax;
    for (size_t t = 1; t < n2; ++t) {
        qmax = std::max(qmax, fld[t]);
        qmin = std::min(qmin, fld[t]);
    }
    const R dq = (qmax - qmin) / R(nbins - 1);
    for (int nn = 1; nn <= nbins; ++nn) {
        qn[nn - 1] = qmax - dq * R(nn - 1);
        an[nn - 1] = 0;
        cn[nn - 1] = 0;
    }
    for (int j = 1; j <= jmax; ++j) {
        const R phi0 = R(-0.5) * pi + dphi * R(j - 1);
        const R da = a * a * dphi * dlambda * std::cos(phi0);
        for (int i = 1; i <= imax; ++i) {
            const R q = F2(fld, i, j, imax);
            int ind = 1 + (int)((qmax - q) / dq);
            if (ind_out) F2(ind_out, i, j, imax) = ind;
            an[ind - 1] = an[ind - 1] + da;
            cn[ind - 1] = cn[ind - 1] + da * q;
        }
    }
    aan[0] = 0;
    ccn[0] = 0;
    for (int nn = 2; nn <= nbins; ++nn) {
        aan[nn - 1] = aan[nn - 2] + an[nn - 1];
        ccn[nn - 1] = ccn[nn - 2] + cn[nn - 1];
    }
    for (int j = 1; j <= nrow - 1; ++j) {
        for (int nn = 1; nn <= nbins - 1; ++nn) {
            if (aan[nn - 1] <= alat[j - 1] && aan[nn] > alat[j - 1]) {
                const R dd = (alat[j - 1] - aan[nn - 1]) / (aan[nn] - aan[nn - 1]);
                if (qout) qout[j - 1] = qn[nn - 1] * (R(1.) - dd) + qn[nn] * dd;
                if (cout) cout[j - 1] = ccn[nn - 1] * (R(1.) - dd) + ccn[nn] * dd;
            }
        }
    }
    if (qmax_out) *qmax_out = qmax;
}

// compute_reference_states.f90:1-261 (NH18: area analysis + red-black SOR + tref).
// `u` starts from 0 (the reference starts from uninitialised memory, U3).
template <typename R>
void compute_reference_states(const R* pv, const R* uu, const R* pt, const R* stat, int nlon, int nlat, int kmax,
                              int jd, int npart, int maxits, R a, R om, R dz, R eps, R h, R dphi, R dlambda,
                              R r, R cp, R rjac, R* qref, R* u, R* tref, int* num_of_iter,
                              int32_t* ind_out /*(nlon,nlat,kmax) or null*/) {
    const R pi = rpi<R>();
    const R zero_ = 0, half = R(0.5), qtr = R(0.25), one = 1;
    const R rkappa = r / cp;
    const size_t njk = (size_t)jd * kmax;
    zero(qref, njk); zero(u, njk); zero(tref, njk);
    std::vector<R> alat(jd), phi(jd), z(kmax), tb(kmax, R(0)), tg(kmax, R(0));
    std::vector<R> cbar(njk, R(0)), cref(njk, R(0)), qbar(njk, R(0)), ubar(njk, R(0)), tbar(njk, R(0));
    std::vector<R> ajk(njk, R(0)), bjk(njk, R(0)), cjk(njk, R(0)), djk(njk, R(0)), ejk(njk, R(0)), fjk(njk, R(0));
    for (int nn = 1; nn <= jd; ++nn) {
        phi[nn - 1] = dphi * R(nn - 1);
        alat[nn - 1] = R(2.) * pi * a * a * (R(1.) - std::sin(phi[nn - 1]));
    }
    for (int k = 1; k <= kmax; ++k) z[k - 1] = dz * R(k - 1);
    // zonal means, :46-55
    for (int j = jd; j <= nlat; ++j) {
        const int jr = j - (jd - 1);
        for (int i = 1; i <= nlon; ++i)
            for (int k = 1; k <= kmax; ++k) {
                F2(qbar, jr, k, jd) = F2(qbar, jr, k, jd) + F3(pv, i, j, k, nlon, nlat) / R(nlon);
                F2(tbar, jr, k, jd) = F2(tbar, jr, k, jd) + F3(pt, i, j, k, nlon, nlat) / R(nlon);
                F2(ubar, jr, k, jd) = F2(ubar, jr, k, jd) + F3(uu, i, j, k, nlon, nlat) / R(nlon);
            }
    }
    // hemispheric-mean potential temperature, :58-66
    R wt = 0;
    for (int j = jd; j <= nlat; ++j) {
        const R phi0 = dphi * R(j - 1) - R(0.5) * pi;
        const R c = std::cos(phi0);
        for (int k = 1; k <= kmax; ++k) tb[k - 1] = tb[k - 1] + c * F2(tbar, j - (jd - 1), k, jd);
        wt = wt + c;
    }
    for (int k = 1; k <= kmax; ++k) tb[k - 1] = tb[k - 1] / wt;

    std::vector<R> qn(npart), an(npart), cn(npart), aan(npart), ccn(npart), qcol(jd), ccol(jd);
    for (int k = 2; k <= kmax - 1; ++k) {
        const R* pv2 = pv + (size_t)nlon * nlat * (k - 1);
        std::fill(qcol.begin(), qcol.end(), R(0));
        std::fill(ccol.begin(), ccol.end(), R(0));
        R qmax;
        area_analysis<R>(nlon, nlat, pv2, npart, jd, alat.data(), a, dphi, dlambda, qcol.data(), ccol.data(),
                         &qmax, qn, an, cn, aan, ccn,
                         ind_out ? ind_out + (size_t)nlon * nlat * (k - 1) : nullptr);
        for (int j = 1; j <= jd - 1; ++j) {
            F2(qref, j, k, jd) = qcol[j - 1];
            F2(cref, j, k, jd) = ccol[j - 1];
        }
        F2(qref, jd, k, jd) = qmax;  // :108
        F2(cbar, jd, k, jd) = 0;     // :110-115
        for (int j = jd - 1; j >= 1; --j) {
            const R phi0 = dphi * (R(j) - R(0.5));
            F2(cbar, j, k, jd) = F2(cbar, j + 1, k, jd) +
                                 R(0.5) * (F2(qbar, j + 1, k, jd) + F2(qbar, j, k, jd)) * a * dphi * R(2.) * pi * a *
                                     std::cos(phi0);
        }
    }
    // normalise by the Coriolis parameter, :122-130
    for (int j = 2; j <= jd; ++j) {
        const R phi0 = dphi * R(j - 1);
        const R cor = R(2.) * om * std::sin(phi0);
        for (int k = 1; k <= kmax; ++k) F2(qref, j, k, jd) = F2(qref, j, k, jd) / cor;
    }
    for (int k = 2; k <= kmax - 1; ++k) F2(qref, 1, k, jd) = R(2.) * F2(qref, 2, k, jd) - F2(qref, 3, k, jd);

    // SOR coefficients, :135-162
    for (int j = 2; j <= jd - 1; ++j) {
        const R phi0 = R(j - 1) * dphi, phip = (R(j) - R(0.5)) * dphi, phim = (R(j) - R(1.5)) * dphi;
        const R cos0 = std::cos(phi0), cosp = std::cos(phip), cosm = std::cos(phim);
        const R sin0 = std::sin(phi0), sinp = std::sin(phip), sinm = std::sin(phim);
        for (int k = 2; k <= kmax - 1; ++k) {
            const R zp = R(0.5) * (z[k] + z[k - 1]);
            const R zm = R(0.5) * (z[k - 2] + z[k - 1]);
            const R statp = R(0.5) * (stat[k] + stat[k - 1]);
            const R statm = R(0.5) * (stat[k - 2] + stat[k - 1]);
            const R fact = R(4.) * om * om * h * a * a * sin0 * dphi * dphi / (dz * dz * r * cos0);
            R amp = std::exp(-zp / h) * std::exp(rkappa * zp / h) / statp;
            amp = amp * fact * std::exp(z[k - 1] / h);
            R amm = std::exp(-zm / h) * std::exp(rkappa * zm / h) / statm;
            amm = amm * fact * std::exp(z[k - 1] / h);
            F2(ajk, j, k, jd) = R(1.) / (sinp * cosp);
            F2(bjk, j, k, jd) = R(1.) / (sinm * cosm);
            F2(cjk, j, k, jd) = amp;
            F2(djk, j, k, jd) = amm;
            F2(ejk, j, k, jd) = -F2(ajk, j, k, jd) - F2(bjk, j, k, jd) - F2(cjk, j, k, jd) - F2(djk, j, k, jd);
            F2(fjk, j, k, jd) = -om * a * dphi * (F2(qref, j + 1, k, jd) - F2(qref, j - 1, k, jd));
        }
    }
    R anormf = zero_;
    for (int j = 2; j <= jd - 1; ++j)
        for (int k = 2; k <= kmax - 1; ++k) anormf = anormf + std::fabs(F2(fjk, j, k, jd));

    // top-boundary increment of row j, :188-190 (iteration-invariant)
    std::vector<R> uzrow(jd, R(0));
    for (int j = 2; j <= jd - 1; ++j) {
        const R phi0 = dphi * R(j - 1);
        R uz = dz * r * std::cos(phi0) * std::exp(-z[kmax - 2] * rkappa / h);
        uz = uz * (F2(tbar, j + 1, kmax - 1, jd) - F2(tbar, j - 1, kmax - 1, jd)) /
             (R(2.) * om * std::sin(phi0) * dphi * h * a);
        uzrow[j - 1] = uz;
    }

    R omega = one;
    int nnn = 1;
    bool converged = false;
    for (nnn = 1; nnn <= maxits; ++nnn) {  // :174-209
        R anorm = zero_;
        for (int j = 2; j <= jd - 1; ++j) {
            for (int k = 2; k <= kmax - 1; ++k) {
                if (((j + k) % 2) == (nnn % 2)) {
                    const R e = F2(ejk, j, k, jd);
                    const R resid = F2(ajk, j, k, jd) * F2(u, j + 1, k, jd) + F2(bjk, j, k, jd) * F2(u, j - 1, k, jd) +
                                    F2(cjk, j, k, jd) * F2(u, j, k + 1, jd) + F2(djk, j, k, jd) * F2(u, j, k - 1, jd) +
                                    e * F2(u, j, k, jd) - F2(fjk, j, k, jd);
                    anorm = anorm + std::fabs(resid);
                    if (e != R(0)) F2(u, j, k, jd) = F2(u, j, k, jd) - omega * resid / e;
                    if (e == R(0)) F2(u, j, k, jd) = 0;
                }
            }
            F2(u, j, 1, jd) = 0;
            F2(u, j, kmax, jd) = F2(u, j, kmax - 2, jd) - uzrow[j - 1];
        }
        for (int k = 1; k <= kmax; ++k) {
            F2(u, jd, k, jd) = 0;
            F2(u, 1, k, jd) = F2(ubar, 1, k, jd) + (F2(cref, 1, k, jd) - F2(cbar, 1, k, jd)) / (R(2.) * pi * a);
        }
        if (nnn == 1)
            omega = one / (one - half * rjac * rjac);
        else
            omega = one / (one - qtr * rjac * rjac * omega);
        if (nnn > 1 && anorm < eps * anormf) { converged = true; break; }
    }
    if (!converged) {  // :213-215, loop index is maxits+1 after a completed DO loop
        zero(u, njk);
        nnn = maxits + 1;
    }
    *num_of_iter = nnn;

    for (int j = 2; j <= jd - 1; ++j) {  // :224-229
        const R phi0 = dphi * R(j - 1);
        const R c = std::cos(phi0);
        for (int k = 1; k <= kmax; ++k) F2(u, j, k, jd) = F2(u, j, k, jd) / c;
    }
    for (int k = 1; k <= kmax; ++k) {
        F2(u, 1, k, jd) = F2(ubar, 1, k, jd);
        F2(u, jd, k, jd) = R(2.) * F2(u, jd - 1, k, jd) - F2(u, jd - 2, k, jd);
    }
    // tref, :233-258
    for (int k = 2; k <= kmax - 1; ++k) {
        const R t00 = 0;
        const R zz = dz * R(k - 1);
        F2(tref, 1, k, jd) = t00;
        F2(tref, 2, k, jd) = t00;
        for (int j = 2; j <= jd - 1; ++j) {
            const R phi0 = dphi * R(j - 1);
            const R cor = R(2.) * om * std::sin(phi0);
            const R uz = (F2(u, j, k + 1, jd) - F2(u, j, k - 1, jd)) / (R(2.) * dz);
            R ty = -cor * uz * a * h * std::exp(rkappa * zz / h);
            ty = ty / r;
            F2(tref, j + 1, k, jd) = F2(tref, j - 1, k, jd) + R(2.) * ty * dphi;
        }
        tg[k - 1] = 0;
        wt = 0;
        for (int j = 1; j <= jd; ++j) {
            const R phi0 = dphi * R(j - 1);
            tg[k - 1] = tg[k - 1] + std::cos(phi0) * F2(tref, j, k, jd);
            wt = wt + std::cos(phi0);
        }
        tg[k - 1] = tg[k - 1] / wt;
        const R tres = tb[k - 1] - tg[k - 1];
        for (int j = 1; j <= jd; ++j) F2(tref, j, k, jd) = F2(tref, j, k, jd) + tres;
    }
    for (int j = 1; j <= jd; ++j) {
        F2(tref, j, 1, jd) = F2(tref, j, 2, jd) - tb[1] + tb[0];
        F2(tref, j, kmax, jd) = F2(tref, j, kmax - 1, jd) - tb[kmax - 2] + tb[kmax - 1];
    }
}

// compute_qref_and_fawa_first.f90:1-176 (NHN22 first stage).
template <typename R>
void compute_qref_and_fawa_first(const R* pv, const R* uu, const R* vort, const R* pt, const R* /*tn0*/, int imax,
                                 int jmax, int kmax, int nd, int nnd, int jb, int jd, R a, R omega, R dz, R h,
                                 R dphi, R dlambda, R rr, R cp, R* qref, R* ubar, R* tbar, R* fawa, R* ckref,
                                 R* tjk, R* sjk, int32_t* ind_pv, int32_t* ind_vort) {
    const R pi = rpi<R>();
    const R rkappa = rr / cp;
    const size_t njk = (size_t)nd * kmax;
    const int n = jd - 2;
    zero(qref, njk); zero(ubar, njk); zero(tbar, njk); zero(fawa, njk); zero(ckref, njk);
    zero(tjk, (size_t)n * (kmax - 1));
    zero(sjk, (size_t)n * n * (kmax - 1));
    std::vector<R> alat(nd), z(kmax), cref(njk, R(0)), cbar(njk, R(0)), qbar(njk, R(0));
    for (int nn = 1; nn <= nd; ++nn) alat[nn - 1] = R(2.) * pi * a * a * (R(1.) - std::sin(dphi * R(nn - 1)));
    for (int k = 1; k <= kmax; ++k) z[k - 1] = dz * R(k - 1);
    // zonal means over the northern rows, :45-54
    for (int j = nd; j <= jmax; ++j) {
        const int jr = j - (nd - 1);
        for (int i = 1; i <= imax; ++i)
            for (int k = 1; k <= kmax; ++k) {
                F2(qbar, jr, k, nd) = F2(qbar, jr, k, nd) + F3(pv, i, j, k, imax, jmax) / R(imax);
                F2(tbar, jr, k, nd) = F2(tbar, jr, k, nd) + F3(pt, i, j, k, imax, jmax) / R(imax);
                F2(ubar, jr, k, nd) = F2(ubar, jr, k, nd) + F3(uu, i, j, k, imax, jmax) / R(imax);
            }
    }
    std::vector<R> qn(nnd), an(nnd), cn(nnd), aan(nnd), ccn(nnd), qcol(nd), ccol(nd);
    for (int k = 2; k <= kmax - 1; ++k) {
        const size_t off = (size_t)imax * jmax * (k - 1);
        // QGPV area analysis -> qref, cref  (:61-98)
        std::fill(qcol.begin(), qcol.end(), R(0));
        std::fill(ccol.begin(), ccol.end(), R(0));
        R qmax;
        area_analysis<R>(imax, jmax, pv + off, nnd, nd, alat.data(), a, dphi, dlambda, qcol.data(), ccol.data(),
                         &qmax, qn, an, cn, aan, ccn, ind_pv ? ind_pv + off : nullptr);
        for (int j = 1; j <= nd - 1; ++j) {
            F2(qref, j, k, nd) = qcol[j - 1];
            F2(cref, j, k, nd) = ccol[j - 1];
        }
        F2(qref, nd, k, nd) = qmax;
        F2(cbar, nd, k, nd) = 0;  // :99-104
        for (int j = nd - 1; j >= 1; --j) {
            const R phi0 = dphi * (R(j) - R(0.5));
            F2(cbar, j, k, nd) = F2(cbar, j + 1, k, nd) +
                                 R(0.5) * (F2(qbar, j + 1, k, nd) + F2(qbar, j, k, nd)) * a * dphi * R(2.) * pi * a *
                                     std::cos(phi0);
        }
        // absolute-vorticity area analysis -> ckref  (:109-141)
        std::fill(ccol.begin(), ccol.end(), R(0));
        area_analysis<R>(imax, jmax, vort + off, nnd, nd, alat.data(), a, dphi, dlambda, (R*)nullptr, ccol.data(),
                         (R*)nullptr, qn, an, cn, aan, ccn, ind_vort ? ind_vort + off : nullptr);
        for (int j = 1; j <= nd - 1; ++j) F2(ckref, j, k, nd) = ccol[j - 1];
    }
    // normalise by sin(latitude), :146-154
    for (int j = 2; j <= nd; ++j) {
        const R cor = std::sin(dphi * R(j - 1));
        for (int k = 1; k <= kmax; ++k) F2(qref, j, k, nd) = F2(qref, j, k, nd) / cor;
    }
    for (int k = 2; k <= kmax - 1; ++k) F2(qref, 1, k, nd) = R(2.) * F2(qref, 2, k, nd) - F2(qref, 3, k, nd);
    // FAWA, :157
    for (size_t t = 0; t < njk; ++t) fawa[t] = (cref[t] - cbar[t]) / (R(2.) * pi * a);
    // top boundary condition, :164-175
    for (int jj = jb + 2; jj <= nd - 1; ++jj) {
        const int j = jj - jb;
        const R phi0 = R(jj - 1) * dphi;
        const R cos0 = std::cos(phi0), sin0 = std::sin(phi0);
        R t = -dz * rr * cos0 * std::exp(-z[kmax - 2] * rkappa / h);
        t = t * (F2(tbar, j + 1, kmax, nd) - F2(tbar, j - 1, kmax, nd));
        t = t / (R(4.) * omega * sin0 * dphi * h * a);
        F2(tjk, j - 1, kmax - 1, n) = t;
        F3(sjk, j - 1, j - 1, kmax - 1, n, n) = R(1.);
    }
}

// matrix_b4_inversion.f90:1-88.
template <typename R>
void matrix_b4_inversion(int k, int jmax, int kmax, int nd, int jb, int jd, const R* z, const R* statn,
                         const R* qref, const R* ckref, const R* sjk, R a, R om, R dz, R h, R rr, R cp, R* qjj,
                         R* djj, R* cjj, R* rj) {
    const int n = jd - 2;
    const R rkappa = rr / cp;
    const R pi = rpi<R>();
    const R dp = pi / R(jmax - 1);
    const R zp = R(0.5) * (z[k] + z[k - 1]);
    const R zm = R(0.5) * (z[k - 2] + z[k - 1]);
    const R statp = R(0.5) * (statn[k] + statn[k - 1]);
    const R statm = R(0.5) * (statn[k - 2] + statn[k - 1]);
    zero(cjj, (size_t)n * n); zero(djj, (size_t)n * n); zero(qjj, (size_t)n * n); zero(rj, (size_t)n);
    const R* sjj = sjk + (size_t)n * n * (k - 1);
    std::vector<R> cdiag(n, R(0));
    for (int jj = jb + 2; jj <= nd - 1; ++jj) {
        const int j = jj - jb;
        const R phi0 = R(jj - 1) * dp, phip = (R(jj) - R(0.5)) * dp, phim = (R(jj) - R(1.5)) * dp;
        const R cos0 = std::cos(phi0), cosp = std::cos(phip), cosm = std::cos(phim);
        const R sin0 = std::sin(phi0), sinp = std::sin(phip), sinm = std::sin(phim);
        const R fact = R(4.) * om * om * h * a * a * sin0 * dp * dp / (dz * dz * rr * cos0);
        R amp = std::exp(-zp / h) * std::exp(rkappa * zp / h) / statp;
        amp = amp * fact * std::exp(z[k - 1] / h);
        R amm = std::exp(-zm / h) * std::exp(rkappa * zm / h) / statm;
        amm = amm * fact * std::exp(z[k - 1] / h);
        const R ajk = R(1.) / (sinp * cosp);
        const R bjk = R(1.) / (sinm * cosm);
        const R cjk = amp, djk = amm;
        const R ejk = ajk + bjk + cjk + djk;
        const R fjk = R(-0.5) * a * dp * (F2(qref, jj + 1, k, nd) - F2(qref, jj - 1, k, nd));
        // north-south boundary values, :54-57
        const R ujd = 0;
        const R phib = dp * R(jb);
        const R u1 = F2(ckref, jb + 1, k, nd) / (R(2.) * pi * a) - om * a * std::cos(phib);
        rj[j - 2] = fjk;
        if (j == 2) rj[j - 2] = fjk - bjk * u1;
        if (j == jd - 1) rj[j - 2] = fjk - ajk * ujd;
        F2(cjj, j - 1, j - 1, n) = cjk;
        F2(djj, j - 1, j - 1, n) = djk;
        cdiag[j - 2] = cjk;
        F2(qjj, j - 1, j - 1, n) = -ejk;
        if (j - 1 >= 1 && j - 1 < jd - 2) F2(qjj, j - 1, j, n) = ajk;
        if (j - 1 > 1 && j - 1 <= jd - 2) F2(qjj, j - 1, j - 2, n) = bjk;
    }
    // Qk + Ck Sk, :78-86.  Ck is diagonal, so the reference's sum over kk has a
    // single non-zero term cjj(i,i)*sjj(i,j) (all others are exact zeros).
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= n; ++j) {
            const R x = R(0) + cdiag[i - 1] * F2(sjj, i, j, n);
            F2(qjj, i, j, n) = F2(qjj, i, j, n) + x;
        }
}

// matrix_after_inversion.f90:1-39 (sjk, tjk updated in place).
template <typename R>
void matrix_after_inversion(int k, int kmax, int jd, const R* qjj, const R* djj, const R* cjj, const R* rj,
                            R* sjk, R* tjk) {
    const int n = jd - 2;
    (void)kmax;
    std::vector<R> tj(n), yj(n);
    for (int i = 1; i <= n; ++i) tj[i - 1] = F2(tjk, i, k, n);
    R* sout = sjk + (size_t)n * n * (k - 2);
    // S_{k-1} = -Qinv * D ; D diagonal -> one non-zero term per (i,j), :11-19
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= n; ++j) {
            const R x = R(0) + F2(qjj, i, j, n) * F2(djj, j, j, n);
            F2(sout, i, j, n) = -x;
        }
    // r_k - C_k T_k, :22-28
    for (int i = 1; i <= n; ++i) {
        const R y = R(0) + F2(cjj, i, i, n) * tj[i - 1];
        yj[i - 1] = rj[i - 1] - y;
    }
    // T_{k-1} = Qinv * y, :31-37
    for (int i = 1; i <= n; ++i) {
        R t = 0;
        for (int kk = 1; kk <= n; ++kk) t = t + F2(qjj, i, kk, n) * yj[kk - 1];
        F2(tjk, i, k - 1, n) = t;
    }
}

// upward_sweep.f90:1-92.  u(1,2:kmax-1) stays 0 (U4).
template <typename R>
void upward_sweep(int jmax, int kmax, int nd, int jb, int jd, const R* sjk, const R* tjk, const R* ckref,
                  const R* tb, const R* qref_over_cor, R* tref, R* qref, R* u, R a, R om, R dz, R h, R rr, R cp) {
    const int n = jd - 2;
    const R rkappa = rr / cp;
    const R pi = rpi<R>();
    const R dp = pi / R(jmax - 1);
    zero(tref, (size_t)jd * kmax); zero(u, (size_t)jd * kmax); zero(qref, (size_t)nd * kmax);
    std::vector<R> pjk((size_t)n * kmax, R(0)), tg(kmax, R(0));
    for (int k = 1; k <= kmax - 1; ++k) {  // :19-34
        const R* sjj = sjk + (size_t)n * n * (k - 1);
        for (int i = 1; i <= n; ++i) {
            R y = 0;
            for (int kk = 1; kk <= n; ++kk) y = y + F2(sjj, i, kk, n) * F2(pjk, kk, k, n);
            F2(pjk, i, k + 1, n) = y + F2(tjk, i, k, n);
        }
    }
    for (int k = 1; k <= kmax; ++k)
        for (int j = 2; j <= jd - 1; ++j) F2(u, j, k, jd) = F2(pjk, j - 1, k, n);
    // corner boundary conditions, :44-48
    F2(u, 1, 1, jd) = 0;
    F2(u, jd, 1, jd) = 0;
    F2(u, 1, kmax, jd) = F2(ckref, 1 + jb, kmax, nd) / (R(2.) * pi * a) - om * a * std::cos(dp * R(jb));
    F2(u, jd, kmax, jd) = 0;
    for (int jj = jb + 1; jj <= nd - 1; ++jj) {  // :51-56
        const int j = jj - jb;
        const R c = std::cos(dp * R(jj - 1));
        for (int k = 1; k <= kmax; ++k) F2(u, j, k, jd) = F2(u, j, k, jd) / c;
    }
    for (int k = 1; k <= kmax; ++k) F2(u, jd, k, jd) = R(2.) * F2(u, jd - 1, k, jd) - F2(u, jd - 2, k, jd);
    // tref, :59-91
    std::memcpy(qref, qref_over_cor, sizeof(R) * (size_t)nd * kmax);
    for (int k = 2; k <= kmax - 1; ++k) {
        const R t00 = 0;
        const R zz = dz * R(k - 1);
        F2(tref, 1, k, jd) = t00;
        F2(tref, 2, k, jd) = t00;
        for (int j = 2; j <= jd - 1; ++j) {
            const R phi0 = dp * R(j - 1 + jb);
            const R cor = R(2.) * om * std::sin(phi0);
            const R uz = (F2(u, j, k + 1, jd) - F2(u, j, k - 1, jd)) / (R(2.) * dz);
            R ty = -cor * uz * a * h * std::exp(rkappa * zz / h);
            ty = ty / rr;
            F2(tref, j + 1, k, jd) = F2(tref, j - 1, k, jd) + R(2.) * ty * dp;
        }
        for (int j = 1; j <= nd; ++j) F2(qref, j, k, nd) = F2(qref_over_cor, j, k, nd) * std::sin(dp * R(j - 1));
        tg[k - 1] = 0;
        R wt = 0;
        for (int jj = jb + 1; jj <= nd; ++jj) {
            const int j = jj - jb;
            const R c = std::cos(dp * R(jj - 1));
            tg[k - 1] = tg[k - 1] + c * F2(tref, j, k, jd);
            wt = wt + c;
        }
        tg[k - 1] = tg[k - 1] / wt;
        const R tres = tb[k - 1] - tg[k - 1];
        for (int j = 1; j <= jd; ++j) F2(tref, j, k, jd) = F2(tref, j, k, jd) + tres;
    }
    for (int j = 1; j <= jd; ++j) {
        F2(tref, j, 1, jd) = F2(tref, j, 2, jd) - tb[1] + tb[0];
        F2(tref, j, kmax, jd) = F2(tref, j, kmax - 1, jd) - tb[kmax - 2] + tb[kmax - 1];
    }
}

// compute_flux_dirinv.f90:1-182.  Only is_nhem=.true. is ever used by the
// Python layer; the .false. branch is restated too.
// `mask_count` (optional, 2 entries) receives the number of (i,j,jj,k) pairs
// that entered the anticyclonic (qe<=0, jj>=j) and cyclonic (qe>0, jj<j) sums.
template <typename R>
void compute_flux_dirinv_nshem(const R* pv, const R* uu, const R* vv, const R* pt, const R* ncforce,
                               const R* tn0, const R* qref, const R* uref, const R* tref, int imax, int jmax,
                               int kmax, int nd, int jb, int jd, int is_nhem, R a, R om, R dz, R h, R rr, R cp,
                               R prefac, R* astar1, R* astar2, R* ncforce3d, R* ua1, R* ua2, R* ep1, R* ep2,
                               R* ep3, R* ep4, int64_t* mask_count) {
    const R pi = rpi<R>();
    const R dp = pi / R(jmax - 1);
    const R rkappa = rr / cp;
    const size_t n3 = (size_t)imax * nd * kmax;
    zero(astar1, n3); zero(astar2, n3); zero(ncforce3d, n3); zero(ua1, n3); zero(ua2, n3);
    zero(ep1, n3); zero(ep2, n3); zero(ep3, n3); zero(ep4, (size_t)imax * nd);
    const R* tg = tn0;
    int jstart, jend;
    if (is_nhem) { jstart = jb + 1; jend = nd - 1; } else { jstart = 2; jend = nd - jb; }
    const R hemoff = is_nhem ? R(0) : R(0.5) * pi;
    std::vector<R> cosphi1(nd), aaw(nd);
    for (int jj = 1; jj <= nd; ++jj) {
        const R phi1 = dp * R(jj - 1) - hemoff;
        cosphi1[jj - 1] = std::cos(phi1);
        aaw[jj - 1] = a * dp * std::cos(phi1);
    }
    const R ab = a * dp;
    const int joff = is_nhem ? nd - 1 : 0;  // row offset into the global fields
    const int uoff = is_nhem ? jb : 0;      // row offset into uref/tref
    int64_t cnt2 = 0, cnt1 = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : cnt1, cnt2)
    for (int k = 2; k <= kmax - 1; ++k) {
        std::vector<R> qcol(nd), ucol(nd), ncol(nd);
        for (int i = 1; i <= imax; ++i) {
            for (int jj = 1; jj <= nd; ++jj) {
                qcol[jj - 1] = F3(pv, i, jj + joff, k, imax, jmax);
                ucol[jj - 1] = F3(uu, i, jj + joff, k, imax, jmax);
                ncol[jj - 1] = F3(ncforce, i, jj + joff, k, imax, jmax);
            }
            for (int j = jstart; j <= jend; ++j) {
                const R phi0 = dp * R(j - 1) - hemoff;
                const R cphi0 = std::cos(phi0);
                const R qr = F2(qref, j, k, nd);
                const R ur = F2(uref, j - uoff, k, jd);
                R as_pos = 0;  // sum over qe<=0, jj>=j of -qe*aa
                R as_neg = 0;  // sum over qe>0,  jj<j  of +qe*aa
                R ncf = 0, u2 = 0;
                for (int jj = 1; jj <= nd; ++jj) {
                    const R qe = qcol[jj - 1] - qr;
                    const R ue = ucol[jj - 1] * cphi0 - ur * cosphi1[jj - 1];
                    const R aa = aaw[jj - 1];
                    if (qe <= R(0) && jj >= j) {
                        as_pos = as_pos - qe * aa;
                        ncf = ncf - ncol[jj - 1] * aa;
                        u2 = u2 - qe * ue * ab;
                        ++cnt2;
                    }
                    if (qe > R(0) && jj < j) {
                        as_neg = as_neg + qe * aa;
                        ncf = ncf + ncol[jj - 1] * aa;
                        u2 = u2 + qe * ue * ab;
                        ++cnt1;
                    }
                }
                R a1, a2;
                if (is_nhem) { a2 = as_pos; a1 = as_neg; } else { a1 = as_pos; a2 = as_neg; }
                F3(astar1, i, j, k, imax, nd) = a1;
                F3(astar2, i, j, k, imax, nd) = a2;
                F3(ncforce3d, i, j, k, imax, nd) = ncf;
                F3(ua2, i, j, k, imax, nd) = u2;
                // other fluxes, :119-157
                const R uu0 = F3(uu, i, j + joff, k, imax, jmax);
                const R vv0 = F3(vv, i, j + joff, k, imax, jmax);
                const R pt0 = F3(pt, i, j + joff, k, imax, jmax);
                const R tr = F2(tref, j - uoff, k, jd);
                F3(ua1, i, j, k, imax, nd) = ur * (a1 + a2);
                R e1 = R(-0.5) * ((uu0 - ur) * (uu0 - ur));
                e1 = e1 + R(0.5) * (vv0 * vv0);
                R ep11 = R(0.5) * ((pt0 - tr) * (pt0 - tr));
                const R zz = dz * R(k - 1);
                ep11 = ep11 * (rr / h) * std::exp(-rkappa * zz / h);
                ep11 = ep11 * R(2.) * dz / (tg[k] - tg[k - 2]);
                e1 = e1 - ep11;
                const R phip = dp * R(j) - hemoff, phim = dp * R(j - 2) - hemoff;
                const R cosp = std::cos(phip), cos0 = cphi0, cosm = std::cos(phim), sin0 = std::sin(phi0);
                F3(ep1, i, j, k, imax, nd) = e1 * cos0;
                F3(ep2, i, j, k, imax, nd) =
                    (F3(uu, i, j + joff + 1, k, imax, jmax) - F2(uref, j - uoff + 1, k, jd)) * cosp * cosp *
                    F3(vv, i, j + joff + 1, k, imax, jmax);
                F3(ep3, i, j, k, imax, nd) =
                    (F3(uu, i, j + joff - 1, k, imax, jmax) - F2(uref, j - uoff - 1, k, jd)) * cosm * cosm *
                    F3(vv, i, j + joff - 1, k, imax, jmax);
                if (k == 2) {  // low-level meridional eddy heat flux, :160-171
                    const R ep41 = R(2.) * om * sin0 * cos0 * dz / prefac;
                    const R ep42 = std::exp(-dz / h) * F3(vv, i, j + joff, 2, imax, jmax) *
                                   (F3(pt, i, j + joff, 2, imax, jmax) - F2(tref, j - uoff, 2, jd)) / (tg[2] - tg[0]);
                    R ep43 = F3(vv, i, j + joff, 1, imax, jmax) *
                             (F3(pt, i, j + joff, 1, imax, jmax) - F2(tref, j - uoff, 1, jd));
                    ep43 = R(0.5) * ep43 / (tg[1] - tg[0]);
                    F2(ep4, i, j, imax) = ep41 * (ep42 + ep43);
                }
            }
        }
    }
    // boundary row jb, :173-178.  For jb = 0 the reference's (i,0,k) aliases
    // element (i,nd,k-1) of the same array (no bounds checking);
    // the flat index below reproduces exactly that.  Done after the main loop
    // (serially, in k order) because level k's alias lands in level k-1.
    {
        const R phip = dp * R(jb), phi0 = dp * R(jb - 1);
        const R cosp = std::cos(phip), cos0 = std::cos(phi0);
        for (int k = 2; k <= kmax - 1; ++k)
            for (int i = 1; i <= imax; ++i) {
                const ptrdiff_t flat = (ptrdiff_t)(i - 1) + (ptrdiff_t)imax * ((ptrdiff_t)(jb - 1) + (ptrdiff_t)nd * (k - 1));
                ep2[flat] = (F3(uu, i, nd + jb + 1, k, imax, jmax) - F2(uref, 2, k, jd)) * cosp * cosp *
                            F3(vv, i, nd + jb + 1, k, imax, jmax);
                ep3[flat] = (F3(uu, i, nd + jb, k, imax, jmax) - F2(uref, 1, k, jd)) * cos0 * cos0 *
                            F3(vv, i, nd + jb, k, imax, jmax);
            }
    }
    if (mask_count) { mask_count[0] = cnt2; mask_count[1] = cnt1; }
}

// compute_lwa_only_nhn22.f90:1-107.
template <typename R>
void compute_lwa_only_nhn22(const R* pv, const R* uu, const R* qref, int imax, int jmax, int kmax, int nd, int jb,
                            int is_nhem, R a, R om, R dz, R h, R rr, R cp, R prefac, R* astarbaro, R* ubaro,
                            R* astar1, R* astar2) {
    (void)om; (void)rr; (void)cp;
    const R pi = rpi<R>();
    const R dp = pi / R(jmax - 1);
    const size_t n3 = (size_t)imax * nd * kmax;
    zero(astar1, n3); zero(astar2, n3);
    zero(astarbaro, (size_t)imax * nd); zero(ubaro, (size_t)imax * nd);
    const R dc = dz / prefac;
    int jstart, jend;
    if (is_nhem) { jstart = jb + 1; jend = nd - 1; } else { jstart = 2; jend = nd - jb; }
    const R hemoff = is_nhem ? R(0) : R(0.5) * pi;
    const int joff = is_nhem ? nd - 1 : 0;
    std::vector<R> aaw(nd);
    for (int jj = 1; jj <= nd; ++jj) aaw[jj - 1] = a * dp * std::cos(dp * R(jj - 1) - hemoff);
#pragma omp parallel for schedule(dynamic)
    for (int k = 2; k <= kmax - 1; ++k) {
        std::vector<R> qcol(nd);
        for (int i = 1; i <= imax; ++i) {
            for (int jj = 1; jj <= nd; ++jj) qcol[jj - 1] = F3(pv, i, jj + joff, k, imax, jmax);
            for (int j = jstart; j <= jend; ++j) {
                const R qr = F2(qref, j, k, nd);
                R as_pos = 0, as_neg = 0;
                for (int jj = 1; jj <= nd; ++jj) {
                    const R qe = qcol[jj - 1] - qr;
                    if (qe <= R(0) && jj >= j) as_pos = as_pos - qe * aaw[jj - 1];
                    if (qe > R(0) && jj < j) as_neg = as_neg + qe * aaw[jj - 1];
                }
                if (is_nhem) {
                    F3(astar2, i, j, k, imax, nd) = as_pos;
                    F3(astar1, i, j, k, imax, nd) = as_neg;
                } else {
                    F3(astar1, i, j, k, imax, nd) = as_pos;
                    F3(astar2, i, j, k, imax, nd) = as_neg;
                }
            }
        }
    }
    // column averages, :93-104 (serial in k to keep the reference's sum order)
    for (int k = 2; k <= kmax - 1; ++k) {
        const R zk = dz * R(k - 1);
        const R w = std::exp(-zk / h);
        for (int j = 1; j <= nd; ++j)
            for (int i = 1; i <= imax; ++i)
                F2(astarbaro, i, j, imax) = F2(astarbaro, i, j, imax) +
                                            (F3(astar1, i, j, k, imax, nd) + F3(astar2, i, j, k, imax, nd)) * w * dc;
        for (int j = jstart; j <= jend; ++j)
            for (int i = 1; i <= imax; ++i)
                F2(ubaro, i, j, imax) = F2(ubaro, i, j, imax) + F3(uu, i, j + joff, k, imax, jmax) * w * dc;
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// C entry points (loaded with ctypes by oracle/f2py_shim.py)
// ---------------------------------------------------------------------------
#define ORACLE_EXPORTS(SUF, R)                                                                                       \
    extern "C" void oracle_compute_qgpv_##SUF(int nlon, int nlat, int kmax, const R* ut, const R* vt,              \
                                              const R* theta, const R* height, const R* t0, const R* stat, R aa,   \
                                              R omega, R dz, R hh, R rr, R cp, R* pv, R* avort) {                  \
        compute_qgpv<R>(nlon, nlat, kmax, ut, vt, theta, height, t0, stat, aa, omega, dz, hh, rr, cp, pv, avort);  \
    }                                                                                                                \
    extern "C" void oracle_compute_qgpv_direct_inv_##SUF(int nlon, int nlat, int kmax, int jd, const R* uq,         \
                                                         const R* vq, const R* tq, const R* height, const R* ts0,   \
                                                         const R* tn0, const R* stats, const R* statn, R aa,        \
                                                         R omega, R dz, R hh, R rr, R cp, R* pv, R* avort) {        \
        compute_qgpv_direct_inv<R>(nlon, nlat, kmax, jd, uq, vq, tq, height, ts0, tn0, stats, statn, aa, omega, dz, \
                                   hh, rr, cp, pv, avort);                                                          \
    }                                                                                                                \
    extern "C" void oracle_compute_reference_states_##SUF(                                                           \
        const R* pv, const R* uu, const R* pt, const R* stat, int nlon, int nlat, int kmax, int jd, int npart,      \
        int maxits, R a, R om, R dz, R eps, R h, R dphi, R dlambda, R r, R cp, R rjac, R* qref, R* u, R* tref,      \
        int* num_of_iter, int32_t* ind_out) {                                                                        \
        compute_reference_states<R>(pv, uu, pt, stat, nlon, nlat, kmax, jd, npart, maxits, a, om, dz, eps, h, dphi, \
                                    dlambda, r, cp, rjac, qref, u, tref, num_of_iter, ind_out);                      \
    }                                                                                                                \
    extern "C" void oracle_compute_qref_and_fawa_first_##SUF(                                                        \
        const R* pv, const R* uu, const R* vort, const R* pt, const R* tn0, int imax, int jmax, int kmax, int nd,   \
        int nnd, int jb, int jd, R a, R omega, R dz, R h, R dphi, R dlambda, R rr, R cp, R* qref, R* ubar, R* tbar, \
        R* fawa, R* ckref, R* tjk, R* sjk, int32_t* ind_pv, int32_t* ind_vort) {                                     \
        compute_qref_and_fawa_first<R>(pv, uu, vort, pt, tn0, imax, jmax, kmax, nd, nnd, jb, jd, a, omega, dz, h,   \
                                       dphi, dlambda, rr, cp, qref, ubar, tbar, fawa, ckref, tjk, sjk, ind_pv,      \
                                       ind_vort);                                                                    \
    }                                                                                                                \
    extern "C" void oracle_matrix_b4_inversion_##SUF(int k, int jmax, int kmax, int nd, int jb, int jd, const R* z, \
                                                     const R* statn, const R* qref, const R* ckref, const R* sjk,   \
                                                     R a, R om, R dz, R h, R rr, R cp, R* qjj, R* djj, R* cjj,      \
                                                     R* rj) {                                                        \
        matrix_b4_inversion<R>(k, jmax, kmax, nd, jb, jd, z, statn, qref, ckref, sjk, a, om, dz, h, rr, cp, qjj,    \
                               djj, cjj, rj);                                                                        \
    }                                                                                                                \
    extern "C" void oracle_matrix_after_inversion_##SUF(int k, int kmax, int jd, const R* qjj, const R* djj,        \
                                                        const R* cjj, const R* rj, R* sjk, R* tjk) {                \
        matrix_after_inversion<R>(k, kmax, jd, qjj, djj, cjj, rj, sjk, tjk);                                         \
    }                                                                                                                \
    extern "C" void oracle_upward_sweep_##SUF(int jmax, int kmax, int nd, int jb, int jd, const R* sjk,             \
                                              const R* tjk, const R* ckref, const R* tb, const R* qref_over_cor,    \
                                              R* tref, R* qref, R* u, R a, R om, R dz, R h, R rr, R cp) {           \
        upward_sweep<R>(jmax, kmax, nd, jb, jd, sjk, tjk, ckref, tb, qref_over_cor, tref, qref, u, a, om, dz, h,    \
                        rr, cp);                                                                                     \
    }                                                                                                                \
    extern "C" void oracle_compute_flux_dirinv_nshem_##SUF(                                                          \
        const R* pv, const R* uu, const R* vv, const R* pt, const R* ncforce, const R* tn0, const R* qref,          \
        const R* uref, const R* tref, int imax, int jmax, int kmax, int nd, int jb, int jd, int is_nhem, R a, R om, \
        R dz, R h, R rr, R cp, R prefac, R* astar1, R* astar2, R* ncforce3d, R* ua1, R* ua2, R* ep1, R* ep2,        \
        R* ep3, R* ep4, int64_t* mask_count) {                                                                       \
        compute_flux_dirinv_nshem<R>(pv, uu, vv, pt, ncforce, tn0, qref, uref, tref, imax, jmax, kmax, nd, jb, jd,  \
                                     is_nhem, a, om, dz, h, rr, cp, prefac, astar1, astar2, ncforce3d, ua1, ua2,    \
                                     ep1, ep2, ep3, ep4, mask_count);                                                \
    }                                                                                                                \
    extern "C" void oracle_compute_lwa_only_nhn22_##SUF(const R* pv, const R* uu, const R* qref, int imax,          \
                                                        int jmax, int kmax, int nd, int jb, int is_nhem, R a, R om, \
                                                        R dz, R h, R rr, R cp, R prefac, R* astarbaro, R* ubaro,    \
                                                        R* astar1, R* astar2) {                                      \
        compute_lwa_only_nhn22<R>(pv, uu, qref, imax, jmax, kmax, nd, jb, is_nhem, a, om, dz, h, rr, cp, prefac,    \
                                  astarbaro, ubaro, astar1, astar2);                                                 \
    }

ORACLE_EXPORTS(f64, double)
ORACLE_EXPORTS(f32, float)
