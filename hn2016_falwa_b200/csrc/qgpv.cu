// QGPV stage: absolute vorticity + quasi-geostrophic PV stencils, vertical interpolation and
// the hemispheric-mean potential temperature that feeds the host-side static-stability spline.
//
// Reference behaviour restated (not ported): f90_modules/compute_qgpv.f90:1-137,
// f90_modules/compute_qgpv_direct_inv.f90:1-112, oopinterface.py:241-261 (hemispheric means),
// :537-540 (scipy interp1d, kind="linear", extrapolating).
//
// Layout: [k][lat][lon], longitude fastest.  One thread owns one (lon, lat) column of a level
// chunk and walks upward keeping the three theta levels of the vertical stencil in registers, so
// theta is read once; the u(j±1) / v(i±1) neighbours come from L1/L2 (a 32x8 tile shares them).
// All latitude/level dependent factors are host tables (glibc libm, same as the CPU oracle) so the
// device arithmetic is +,-,*,/ only, compiled with -fmad=false: results are bit-comparable.
#include "common.cuh"

namespace falwa {

__global__ void init_ext_kernel(unsigned long long* ext, int n);  // refstate.cu

struct QgpvTables {
    // per latitude row (nlat each): av1 = 2*omega*sin(phi0), den = 2*aa*cos(phi0)*dphi, cosp, cosm, rden = 1/den
    const double* av1;
    const double* den;
    const double* cosp;
    const double* cosm;
    const double* rden;
    // per level (kmax each): ep[k] = exp(-height[k]/hh), e0[k] = exp(height[k]/hh), height[k],
    // rdh[k] = 1/(height[k+1]-height[k-1])
    const double* ep;
    const double* e0;
    const double* height;
    const double* rdh;
};
// Per-snapshot profiles live in a separate device array prof[batch][prof_rows][kmax]:
//   row 0 = t0 south (rows j <= jd), 1 = t0 north, 2 = static stability south, 3 = north
//   (FAST only) 4, 5 = reciprocal static stability south, north.

// Fused zonal statistics (FAST pipeline): per (row, level) sums of pv, u, theta and per-level extrema of pv and
// avort, so that the reference-state stage does not have to read the four 3-D fields again.
struct QgpvStats {
    double* part;              // [batch][3][kmax][nlat][nwx] warp partial sums (f = 0 pv, 1 u, 2 theta); null = off
    unsigned long long* ext;   // [batch][2][kmax][2] ordered {max, min} of pv, avort (initialised by the caller)
    int nwx;
};

// mode 0: NH18  — pv receives (altp - altm), finished by nh18_finish_kernel (needs the zonal-mean avort)
// mode 1: NHN22 — pv = avort + e0 * ((altp - altm) * f / dh)
// FAST = false: the reference's divisions (bit-comparable with the CPU oracle; F2PY-compatible entry points).
// FAST = true : host-rounded reciprocals instead of the five ~40-instruction DDIV sequences per point
//               (<= 2 ulp per term; the kernel becomes HBM-bound instead of issue-bound).
// The static-stability term alt(l) = ep[l]*(theta(l)-t0[l])/stat[l] is evaluated once per level and rolled.
template <int MODE, bool FAST>
__global__ void __launch_bounds__(256)
avort_pv_kernel(int nlon, int nlat, int kmax, int jd, int kchunk, const double* __restrict__ u,
                const double* __restrict__ v, const double* __restrict__ th, QgpvTables tb,
                const double* __restrict__ prof, double* __restrict__ avort, double* __restrict__ pv,
                size_t batch_stride, QgpvStats stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;  // 0-based row
    const int nchunks = (kmax + kchunk - 1) / kchunk;
    const int chunk = blockIdx.z % nchunks;
    const int b = blockIdx.z / nchunks;
    const bool valid = (i < nlon && j < nlat);
    if (!FAST && !valid) return;            // FAST keeps whole warps alive for the shuffles below
    if (FAST && j >= nlat) return;          // warp-uniform (a warp is one row)
    constexpr int PR = FAST ? 6 : 4;
    const size_t plane = (size_t)nlon * nlat;
    const size_t base = (size_t)b * batch_stride;
    u += base; v += base; th += base; avort += base; pv += base;
    const int k0 = chunk * kchunk;
    const int k1 = min(kmax, k0 + kchunk);
    const bool interior_row = (j >= 1 && j <= nlat - 2);
    const int ic = valid ? i : nlon - 1;
    const int ip = (ic == nlon - 1) ? 0 : ic + 1;
    const int im = (ic == 0) ? nlon - 1 : ic - 1;
    const double av1 = tb.av1[j], den = tb.den[j], cp = tb.cosp[j], cm = tb.cosm[j], rden = tb.rden[j];
    const int hemi = (j + 1 <= jd) ? 0 : 1;
    const double* t0 = prof + (size_t)b * PR * kmax + hemi * kmax;
    const double* stat = t0 + 2 * kmax;
    const double* rstat = t0 + 4 * kmax;
    const size_t col = (size_t)j * nlon + ic;
    auto alt = [&](int l, double theta) -> double {
        return FAST ? tb.ep[l] * (theta - t0[l]) * rstat[l] : tb.ep[l] * (theta - t0[l]) / stat[l];
    };
    // rolling static-stability terms: am = alt(k-1), ac = alt(k)
    double am = (k0 >= 1) ? alt(k0 - 1, th[(size_t)(k0 - 1) * plane + col]) : 0.0;
    double tc = th[(size_t)k0 * plane + col];
    double ac = alt(k0, tc);
    const int lane = threadIdx.x;
    // loads of level k+1 are issued before level k is evaluated (register double buffering)
    struct In { double tp, vp, vm, up, um; };
    auto fetch = [&](int k) -> In {
        In r;
        const size_t off = (size_t)k * plane;
        r.tp = (k + 1 < kmax) ? th[off + plane + col] : 0.0;
        r.vp = r.vm = r.up = r.um = 0.0;
        if (interior_row) {
            r.vp = v[off + (size_t)j * nlon + ip]; r.vm = v[off + (size_t)j * nlon + im];
            r.up = u[off + col + nlon]; r.um = u[off + col - nlon];
        }
        return r;
    };
    In cur = fetch(k0);
    for (int k = k0; k < k1; ++k) {
        const size_t off = (size_t)k * plane;
        In nxt = cur;
        if (k + 1 < k1) nxt = fetch(k + 1);
        const double tp = cur.tp;
        const double ap = (k + 1 < kmax) ? alt(k + 1, tp) : 0.0;
        double av = 0.0;
        if (interior_row) {
            const double dv = cur.vp - cur.vm;
            const double du = -(cur.up * cp - cur.um * cm);
            av = FAST ? av1 + dv * rden + du * rden : av1 + dv / den + du / den;
        }
        double q = 0.0;
        if (k >= 1 && k <= kmax - 2) {
            if (MODE == 0) {
                q = ap - am;
            } else {
                const double strc = FAST ? (ap - am) * av1 * tb.rdh[k]
                                         : (ap - am) * av1 / (tb.height[k + 1] - tb.height[k - 1]);
                q = av + tb.e0[k] * strc;
            }
        }
        if (valid) {
            st_stream(&avort[off + col], av);
            st_stream(&pv[off + col], q);
        }
        if (FAST && stats.part) {
            // warp = 32 consecutive longitudes of row j: partial zonal sums and extrema (deterministic two-stage sum)
            const double uc = valid ? u[off + col] : 0.0;
            double s0 = (valid && interior_row && MODE == 1) ? q : 0.0, s1 = uc, s2 = valid ? tc : 0.0;
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
            if (lane == 0) {
                double* pp = stats.part + (size_t)b * 3 * kmax * nlat * stats.nwx;
                const size_t o = ((size_t)k * nlat + j) * stats.nwx + blockIdx.x;
                const size_t fs = (size_t)kmax * nlat * stats.nwx;
                pp[o] = s0; pp[fs + o] = s1; pp[2 * fs + o] = s2;
            }
            if (interior_row && k >= 1 && k <= kmax - 2) {
                const double qx = warp_max(valid ? q : -INFINITY), qn = warp_min(valid ? q : INFINITY);
                const double ax = warp_max(valid ? av : -INFINITY), an = warp_min(valid ? av : INFINITY);
                if (lane == 0) {
                    unsigned long long* e = stats.ext + (size_t)b * 2 * kmax * 2;
                    if (MODE == 1) {
                        atomicMax(&e[k * 2 + 0], f64_to_ordered(qx));
                        atomicMin(&e[k * 2 + 1], f64_to_ordered(qn));
                    }
                    atomicMax(&e[(kmax + k) * 2 + 0], f64_to_ordered(ax));
                    atomicMin(&e[(kmax + k) * 2 + 1], f64_to_ordered(an));
                }
            }
        }
        am = ac; ac = ap; tc = tp;
        cur = nxt;
    }
}

// Polar rows: avort(:,1,k) / avort(:,nlat,k) = zonal mean of the adjacent row (compute_qgpv.f90:86-93).
// For NHN22 the pole rows of pv (which hold only the stretching term so far) receive that mean too.
// With stats: the polar rows' contribution to the pv row sums and to the pv / avort extrema.
__global__ void __launch_bounds__(256)
polar_rows_kernel(int nlon, int nlat, int kmax, int add_to_pv, double* __restrict__ avort,
                  double* __restrict__ pv, size_t batch_stride, QgpvStats stats) {
    __shared__ double red[33];
    const int k = blockIdx.x;
    const int pole = blockIdx.y;  // 0 south, 1 north
    const int b = blockIdx.z;
    const size_t base = (size_t)b * batch_stride + (size_t)k * nlon * nlat;
    const int jsrc = pole ? nlat - 2 : 1;
    const int jdst = pole ? nlat - 1 : 0;
    double s = 0.0;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) s += avort[base + (size_t)jsrc * nlon + i] / (double)nlon;
    s = block_sum(s, red);
    const bool interior = (k >= 1 && k <= kmax - 2);
    double psum = 0.0, pmx = -INFINITY, pmn = INFINITY;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) {
        avort[base + (size_t)jdst * nlon + i] = s;
        if (add_to_pv && interior) {
            const double q = s + pv[base + (size_t)jdst * nlon + i];
            pv[base + (size_t)jdst * nlon + i] = q;
            psum += q; pmx = fmax(pmx, q); pmn = fmin(pmn, q);
        }
    }
    if (stats.part && add_to_pv && interior) {
        psum = block_sum(psum, red);
        pmx = warp_max(pmx); pmn = warp_min(pmn);
        unsigned long long* e = stats.ext + (size_t)b * 2 * kmax * 2;
        if ((threadIdx.x & 31) == 0) {
            atomicMax(&e[k * 2 + 0], f64_to_ordered(pmx));
            atomicMin(&e[k * 2 + 1], f64_to_ordered(pmn));
            atomicMax(&e[(kmax + k) * 2 + 0], f64_to_ordered(s));
            atomicMin(&e[(kmax + k) * 2 + 1], f64_to_ordered(s));
        }
        if (threadIdx.x == 0) {  // slot 0 of the polar row carries the whole row sum; the other slots are zero
            double* pp = stats.part + (size_t)b * 3 * kmax * nlat * stats.nwx;
            pp[((size_t)k * nlat + jdst) * stats.nwx] = psum;
        }
    }
}

// zm[f][k][j] = (sum of the warp partials, fixed order) / nlon
__global__ void zonal_finalize_kernel(int nlon, int nlat, int kmax, int nwx, const double* __restrict__ part,
                                      double* __restrict__ zm) {
    const size_t n = (size_t)3 * kmax * nlat;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double* p = part + ((size_t)blockIdx.y * n + t) * nwx;
    double s = 0.0;
    for (int w = 0; w < nwx; ++w) s += p[w];
    zm[(size_t)blockIdx.y * n + t] = s / (double)nlon;
}

// NH18: pv = avort + e0 * ((altp-altm) * zonal_mean(avort) / dh)   (compute_qgpv.f90:98-124)
__global__ void __launch_bounds__(256)
nh18_finish_kernel(int nlon, int nlat, int kmax, const double* __restrict__ avort, double* __restrict__ pv,
                   QgpvTables tb, size_t batch_stride, QgpvStats stats) {
    __shared__ double red[33];
    const int j = blockIdx.x, k = blockIdx.y + 1;  // interior levels only
    const size_t row = (size_t)blockIdx.z * batch_stride + ((size_t)k * nlat + j) * nlon;
    double s = 0.0;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) s += avort[row + i] / (double)nlon;
    const double zm = block_sum(s, red);
    const double dh = tb.height[k + 1] - tb.height[k - 1];
    const double e0 = tb.e0[k];
    double psum = 0.0, pmx = -INFINITY, pmn = INFINITY;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) {
        const double strc = pv[row + i] * zm / dh;
        const double q = avort[row + i] + e0 * strc;
        pv[row + i] = q;
        psum += q; pmx = fmax(pmx, q); pmn = fmin(pmn, q);
    }
    if (stats.part) {
        psum = block_sum(psum, red);
        pmx = warp_max(pmx); pmn = warp_min(pmn);
        unsigned long long* e = stats.ext + (size_t)blockIdx.z * 2 * kmax * 2;
        if ((threadIdx.x & 31) == 0) {
            atomicMax(&e[k * 2 + 0], f64_to_ordered(pmx));
            atomicMin(&e[k * 2 + 1], f64_to_ordered(pmn));
        }
        if (threadIdx.x == 0) {
            double* pp = stats.part + (size_t)blockIdx.z * 3 * kmax * nlat * stats.nwx;
            pp[((size_t)k * nlat + j) * stats.nwx] = psum;   // MODE 0 leaves the pv slots of avort_pv_kernel at zero
        }
    }
}

static int build_qgpv_tables(int nlat, int kmax, bool deg_lat, const double* height, double aa, double omega,
                             double hh, std::vector<double>& h) {
    const double pi = host_pi();
    const double dphi = pi / (double)(nlat - 1);
    auto lat = [&](int jm1) -> double {
        if (deg_lat) {  // compute_qgpv.f90:61-66
            double p = -90. + (double)jm1 * 180. / (double)(nlat - 1);
            return p * pi / 180.;
        }
        return -0.5 * pi + pi * (double)jm1 / (double)(nlat - 1);  // compute_qgpv_direct_inv.f90:30-32
    };
    h.assign((size_t)5 * nlat + 4 * kmax, 0.0);
    double *av1 = h.data(), *den = av1 + nlat, *cosp = den + nlat, *cosm = cosp + nlat, *rden = cosm + nlat;
    double *ep = rden + nlat, *e0 = ep + kmax, *hg = e0 + kmax, *rdh = hg + kmax;
    for (int j = 1; j <= nlat; ++j) {
        const double phi0 = lat(j - 1), phim = lat(j - 2), phip = lat(j);
        av1[j - 1] = 2. * omega * std::sin(phi0);
        den[j - 1] = 2. * aa * std::cos(phi0) * dphi;
        cosp[j - 1] = std::cos(phip);
        cosm[j - 1] = std::cos(phim);
        rden[j - 1] = 1. / den[j - 1];
    }
    for (int k = 0; k < kmax; ++k) {
        ep[k] = std::exp(-height[k] / hh);
        e0[k] = std::exp(height[k] / hh);
        hg[k] = height[k];
    }
    for (int k = 1; k + 1 < kmax; ++k) rdh[k] = 1. / (height[k + 1] - height[k - 1]);
    return 0;
}

static QgpvTables view_tables(const double* d, int nlat, int kmax) {
    QgpvTables t;
    t.av1 = d; t.den = d + nlat; t.cosp = d + 2 * nlat; t.cosm = d + 3 * nlat; t.rden = d + 4 * nlat;
    t.ep = d + 5 * nlat; t.e0 = t.ep + kmax; t.height = t.e0 + kmax; t.rdh = t.height + kmax;
    return t;
}

static int pick_kchunk(int nlon, int nlat, int kmax, int batch) {
    // enough CTAs for >= 4 waves of 148 SMs x 8 resident CTAs, otherwise split the column in level chunks
    const long long cols = (long long)ceil_div(nlon, 32) * ceil_div(nlat, 8) * batch;
    int kchunk = kmax;
    while (kchunk > 7 && cols * ceil_div(kmax, kchunk) < 148LL * 8 * 4) kchunk = (kchunk + 1) / 2;
    return kchunk;
}

// Shared driver: mode 0 = NH18 (global t0/stat), mode 1 = NHN22 (hemispheric, split at row jd).
// fast = 0: exact divisions, prof has 4 rows per snapshot.  fast = 1: reciprocals, prof has 6 rows; when zm / ext are
// given the zonal statistics of the reference-state stage are produced on the fly (part = scratch
// [batch][3][kmax][nlat][ceil(nlon/32)], zero-filled here).
int qgpv_launch(int mode, int nlon, int nlat, int kmax, int jd, int batch, const double* u, const double* v,
                const double* th, const double* d_tables, const double* d_prof, double* pv, double* avort,
                cudaStream_t st, int fast = 0, double* zm = nullptr, unsigned long long* ext = nullptr,
                double* part = nullptr) {
    const QgpvTables tb = view_tables(d_tables, nlat, kmax);
    const size_t bstride = (size_t)nlon * nlat * kmax;
    const int kchunk = pick_kchunk(nlon, nlat, kmax, batch);
    dim3 block(32, 8);
    dim3 grid(ceil_div(nlon, 32), ceil_div(nlat, 8), ceil_div(kmax, kchunk) * batch);
    QgpvStats stt;
    stt.part = (fast && zm && ext) ? part : nullptr;
    stt.ext = ext;
    stt.nwx = grid.x;
    if (stt.part) {  // every partial-sum slot is written by the kernels below
        { FALWA_KERNEL("init_ext_kernel", st); init_ext_kernel<<<ceil_div((long long)batch * 2 * kmax * 2, 128), 128, 0, st>>>(ext, batch * 2 * kmax * 2); }
    }
    {
        FALWA_KERNEL("avort_pv_kernel", st);
        if (mode == 0 && !fast)
            avort_pv_kernel<0, false><<<grid, block, 0, st>>>(nlon, nlat, kmax, jd, kchunk, u, v, th, tb, d_prof, avort, pv, bstride, stt);
        else if (mode == 0)
            avort_pv_kernel<0, true><<<grid, block, 0, st>>>(nlon, nlat, kmax, jd, kchunk, u, v, th, tb, d_prof, avort, pv, bstride, stt);
        else if (!fast)
            avort_pv_kernel<1, false><<<grid, block, 0, st>>>(nlon, nlat, kmax, jd, kchunk, u, v, th, tb, d_prof, avort, pv, bstride, stt);
        else
            avort_pv_kernel<1, true><<<grid, block, 0, st>>>(nlon, nlat, kmax, jd, kchunk, u, v, th, tb, d_prof, avort, pv, bstride, stt);
    }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("polar_rows_kernel", st); polar_rows_kernel<<<dim3(kmax, 2, batch), 256, 0, st>>>(nlon, nlat, kmax, mode == 1, avort, pv, bstride, stt); }
    FALWA_LAUNCH_CHECK();
    if (mode == 0) {
        { FALWA_KERNEL("nh18_finish_kernel", st); nh18_finish_kernel<<<dim3(nlat, kmax - 2, batch), 256, 0, st>>>(nlon, nlat, kmax, avort, pv, tb, bstride, stt); }
        FALWA_LAUNCH_CHECK();
    }
    if (stt.part) {
        FALWA_KERNEL("zonal_finalize_kernel", st);
        zonal_finalize_kernel<<<dim3(ceil_div((long long)3 * kmax * nlat, 256), batch), 256, 0, st>>>(nlon, nlat, kmax, stt.nwx, part, zm);
        FALWA_LAUNCH_CHECK();
    }
    return 0;
}

int qgpv_tables_host(int mode, int nlat, int kmax, const double* height, double aa, double omega, double hh,
                     std::vector<double>& h) {
    return build_qgpv_tables(nlat, kmax, mode == 0, height, aa, omega, hh, h);
}

size_t qgpv_tables_size(int nlat, int kmax) { return (size_t)5 * nlat + 4 * kmax; }

// ---------------------------------------------------------------------------------------------
// Vertical interpolation plev -> uniform pseudo-height, scipy.interpolate.interp1d(kind="linear",
// fill_value="extrapolate") as implemented by the SciPy of this image (1.18, _call_linear):
//   idx = clip(searchsorted(x, x_new), 1, n-1); lo = idx-1; hi = idx
//   y = ((x_new-x_lo)/(x_hi-x_lo)) * y_hi + ((x_hi-x_new)/(x_hi-x_lo)) * y_lo
// T is turned into potential temperature with the per-level factor exp(kappa z_p / H) first
// (oopinterface.py:127-128).  One thread per (lon, lat) column; each input level is read once.
// Tables (host): lo[kmax], whi[kmax], wlo[kmax], fac[nlev].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
interp_kernel(int nlon, int nlat, int nlev, int kmax, const double* __restrict__ uin,
              const double* __restrict__ vin, const double* __restrict__ tin, const int* __restrict__ lo,
              const double* __restrict__ whi, const double* __restrict__ wlo, const double* __restrict__ fac,
              double* __restrict__ uo, double* __restrict__ vo, double* __restrict__ to, size_t in_stride,
              size_t out_stride) {
    const size_t plane = (size_t)nlon * nlat;
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= plane) return;
    const int b = blockIdx.y;
    uin += (size_t)b * in_stride; vin += (size_t)b * in_stride; tin += (size_t)b * in_stride;
    uo += (size_t)b * out_stride; vo += (size_t)b * out_stride; to += (size_t)b * out_stride;
    int cur = -2;
    double u0 = 0, u1 = 0, v0 = 0, v1 = 0, t0 = 0, t1 = 0;
    for (int k = 0; k < kmax; ++k) {
        const int l = lo[k];
        if (l != cur) {
            if (l == cur + 1) {
                u0 = u1; v0 = v1; t0 = t1;
            } else {
                u0 = ld_stream(&uin[(size_t)l * plane + col]);
                v0 = ld_stream(&vin[(size_t)l * plane + col]);
                t0 = ld_stream(&tin[(size_t)l * plane + col]) * fac[l];
            }
            u1 = ld_stream(&uin[(size_t)(l + 1) * plane + col]);
            v1 = ld_stream(&vin[(size_t)(l + 1) * plane + col]);
            t1 = ld_stream(&tin[(size_t)(l + 1) * plane + col]) * fac[l + 1];
            cur = l;
        }
        const double a1 = whi[k], a0 = wlo[k];
        st_stream(&uo[(size_t)k * plane + col], a1 * u1 + a0 * u0);
        st_stream(&vo[(size_t)k * plane + col], a1 * v1 + a0 * v0);
        to[(size_t)k * plane + col] = a1 * t1 + a0 * t0;
    }
}

// Hemispheric cos-weighted mean potential temperature per pressure level (oopinterface.py:250-256):
//   t0_s[l] = sum_{j<jd} clat[j]*mean_i(theta) / csm ;  t0_n[l] likewise over the last jd rows.
// Pass 1: one block per (row, level) writes clat*rowmean.  Pass 2: one block per (level, hemi).
__global__ void __launch_bounds__(256)
theta_rowmean_kernel(int nlon, int nlat, int nlev, const double* __restrict__ tin, const double* __restrict__ fac,
                     const double* __restrict__ clat, double* __restrict__ rowmean, size_t in_stride) {
    // one warp per latitude row (8 rows per CTA), 16-byte loads, four independent partial sums per lane so that
    // several loads are in flight; the order of the additions is fixed (lane-strided, then the shuffle tree)
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), l = blockIdx.y, b = blockIdx.z, lane = threadIdx.x & 31;
    if (j >= nlat) return;
    const double* row = tin + (size_t)b * in_stride + ((size_t)l * nlat + j) * nlon;
    const double f = fac[l], c = clat[j];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if ((nlon & 1) == 0 && ((uintptr_t)row & 15) == 0) {
        const double2* r2 = reinterpret_cast<const double2*>(row);
        const int n2 = nlon >> 1;
        int i = lane;
        for (; i + 96 < n2; i += 128) {
            const double2 a = __ldcs(r2 + i), bq = __ldcs(r2 + i + 32), cq = __ldcs(r2 + i + 64), d = __ldcs(r2 + i + 96);
            s0 += (a.x * f) * c + (a.y * f) * c;
            s1 += (bq.x * f) * c + (bq.y * f) * c;
            s2 += (cq.x * f) * c + (cq.y * f) * c;
            s3 += (d.x * f) * c + (d.y * f) * c;
        }
        for (; i < n2; i += 32) { const double2 a = __ldcs(r2 + i); s0 += (a.x * f) * c + (a.y * f) * c; }
    } else {
        for (int i = lane; i < nlon; i += 32) s0 += (ld_stream(&row[i]) * f) * c;
    }
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) rowmean[((size_t)b * nlev + l) * nlat + j] = s / (double)nlon;
}

__global__ void __launch_bounds__(128)
theta_hemi_kernel(int nlat, int nlev, int jd, const double* __restrict__ rowmean, const double* __restrict__ clat,
                  double* __restrict__ out /* [batch][2][nlev] */) {
    __shared__ double red[33];
    const int l = blockIdx.x, hemi = blockIdx.y, b = blockIdx.z;
    const double* rm = rowmean + ((size_t)b * nlev + l) * nlat;
    const int j0 = hemi ? nlat - jd : 0;
    double s = 0.0, c = 0.0;
    for (int j = threadIdx.x; j < jd; j += blockDim.x) {
        s += rm[j0 + j];
        c += clat[j];  // csm is the sum over the first jd rows for both hemispheres (oopinterface.py:250)
    }
    s = block_sum(s, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) out[((size_t)b * 2 + hemi) * nlev + l] = s / c;
}

// ---------------------------------------------------------------------------------------------
// Fused stage 1 (pipeline path): vertical interpolation of u, v, theta AND the absolute-vorticity / QGPV stencils in
// one pass, so the interpolated fields are written once and never read back (2957 MB instead of 4178 MB of traffic
// per 0.25-degree snapshot).  Each thread owns one (lon, lat) column and keeps seven rolling interpolation series in
// registers — its own u, v, theta plus u at (j+1), (j-1) and v at (i+1), (i-1), whose input levels the neighbouring
// threads load as their own (L1/L2 hits) — evaluated with exactly the arithmetic of interp_kernel, so the stencil
// sees bit-identical values to the unfused path.  QGPV of level k-1 is emitted at iteration k (it needs theta(k)).
// Input levels are prefetched into two tagged register slots following a host-built schedule (lo[kmax .. 3 kmax):
// the next two levels, in order of need, that the current bracket does not hold), i.e. ~2.5 output levels ahead where
// the pressure levels are sparse and a full bracket ahead near the ground where they are dense.  All tags are uniform
// across the grid, so the slot bookkeeping is branch-uniform register renaming.  The four level records fit the 128
// registers of 2 CTAs x 256 threads per SM without spills.
// ---------------------------------------------------------------------------------------------
constexpr int kStage1Rows = 8;
template <int MODE, int PFD>
__global__ void __launch_bounds__(32 * kStage1Rows, 2)
interp_qgpv_kernel(int nlon, int nlat, int nlev, int kmax, int jd, const double* __restrict__ uin,
                   const double* __restrict__ vin, const double* __restrict__ tin, const int* __restrict__ lo,
                   const double* __restrict__ whi, const double* __restrict__ wlo, const double* __restrict__ fac,
                   QgpvTables tb, const double* __restrict__ prof, double* __restrict__ uo, double* __restrict__ vo,
                   double* __restrict__ to, double* __restrict__ avort, double* __restrict__ pv, size_t in_stride,
                   size_t out_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (i >= nlon || j >= nlat) return;
    const size_t plane = (size_t)nlon * nlat;
    uin += (size_t)b * in_stride; vin += (size_t)b * in_stride; tin += (size_t)b * in_stride;
    uo += (size_t)b * out_stride; vo += (size_t)b * out_stride; to += (size_t)b * out_stride;
    avort += (size_t)b * out_stride; pv += (size_t)b * out_stride;
    const bool interior_row = (j >= 1 && j <= nlat - 2);
    const int ip = (i == nlon - 1) ? 0 : i + 1, im = (i == 0) ? nlon - 1 : i - 1;
    const size_t col = (size_t)j * nlon + i;
    const size_t cvp = (size_t)j * nlon + ip, cvm = (size_t)j * nlon + im;
    const size_t cup = interior_row ? col + nlon : col, cum = interior_row ? col - nlon : col;
    const double av1 = tb.av1[j], cp = tb.cosp[j], cm = tb.cosm[j], rden = tb.rden[j];
    const int hemi = (j + 1 <= jd) ? 0 : 1;
    const double* t0 = prof + (size_t)b * 6 * kmax + hemi * kmax;
    const double* rstat = t0 + 4 * kmax;
    struct Lev { double u, v, t, up, um, vp, vm; };
    auto load = [&](int l) -> Lev {
        Lev r;
        const size_t off = (size_t)l * plane;
        // raw values only: any arithmetic here would make the warp wait for the load it has just issued
        r.u = ld_stream(&uin[off + col]); r.v = ld_stream(&vin[off + col]); r.t = ld_stream(&tin[off + col]);
        r.up = uin[off + cup]; r.um = uin[off + cum]; r.vp = vin[off + cvp]; r.vm = vin[off + cvm];
        return r;
    };
    const int* pf1 = lo + kmax;
    const int* pf2 = lo + 2 * kmax;
    int cur = lo[0], lp = -1, lq = -1;
    Lev L0 = load(cur), L1 = load(cur + 1);
    Lev P = L1, Q = L1;      // prefetch slots holding levels lp, lq
    double am = 0.0, ac = 0.0, av_prev = 0.0;   // alt(k-2), alt(k-1), avort(k-1)
    for (int k = 0; k < kmax; ++k) {
        const int l = lo[k];
        if (l != cur) {      // advance the bracket (the tables only ever move upward)
            Lev n0, n1;
            if (l == cur + 1) n0 = L1; else if (l == lp) n0 = P; else if (l == lq) n0 = Q; else n0 = load(l);
            if (l + 1 == lp) n1 = P; else if (l + 1 == lq) n1 = Q; else n1 = load(l + 1);
            L0 = n0; L1 = n1; cur = l;
        }
        if (PFD > 0) {   // pull this column's own values of a level further ahead into L2 (no registers involved): the
            const int lf = cur + PFD;   // register prefetch below then waits for an L2 hit instead of DRAM
            if (lf < nlev) {
                const size_t off = (size_t)lf * plane + col;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(uin + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vin + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(tin + off));
            }
        }
        {   // keep the two slots filled with what the schedule asks for; the loads land while levels are evaluated
            const int w1 = pf1[k], w2 = pf2[k];
            if (w1 >= 0 && w1 != lp && w1 != lq) { if (lp != w2) { P = load(w1); lp = w1; } else { Q = load(w1); lq = w1; } }
            if (w2 >= 0 && w2 != lp && w2 != lq) { if (lp != w1) { P = load(w2); lp = w2; } else { Q = load(w2); lq = w2; } }
        }
        const double a1 = whi[k], a0 = wlo[k];
        const double u = a1 * L1.u + a0 * L0.u, v = a1 * L1.v + a0 * L0.v;
        const double th = a1 * (L1.t * fac[cur + 1]) + a0 * (L0.t * fac[cur]);   // T -> theta per input level first
        const size_t off = (size_t)k * plane;
        st_stream(&uo[off + col], u);
        st_stream(&vo[off + col], v);
        st_stream(&to[off + col], th);
        double av = 0.0;
        if (interior_row) {
            const double upv = a1 * L1.up + a0 * L0.up, umv = a1 * L1.um + a0 * L0.um;
            const double vpv = a1 * L1.vp + a0 * L0.vp, vmv = a1 * L1.vm + a0 * L0.vm;
            const double dv = vpv - vmv;
            const double du = -(upv * cp - umv * cm);
            av = av1 + dv * rden + du * rden;
        }
        st_stream(&avort[off + col], av);
        const double ap = tb.ep[k] * (th - t0[k]) * rstat[k];      // alt(k)
        if (k == 0) st_stream(&pv[col], 0.0);
        if (k >= 2) {                                               // QGPV of level k-1
            double q;
            if (MODE == 0) q = ap - am;
            else q = av_prev + tb.e0[k - 1] * ((ap - am) * av1 * tb.rdh[k - 1]);
            st_stream(&pv[off - plane + col], q);
        }
        if (k == kmax - 1) st_stream(&pv[off + col], 0.0);
        am = ac; ac = ap; av_prev = av;
    }
}

// fused stage 1: interpolation + avort + QGPV (mode 0 NH18 / 1 NHN22), then the polar rows (and the NH18 finish)
int interp_qgpv_launch(int mode, int nlon, int nlat, int nlev, int kmax, int jd, int batch, const double* uin,
                       const double* vin, const double* tin, const int* lo, const double* whi, const double* wlo,
                       const double* fac, const double* d_tables, const double* d_prof6, double* uo, double* vo,
                       double* to, double* pv, double* avort, cudaStream_t st) {
    const QgpvTables tb = view_tables(d_tables, nlat, kmax);
    const size_t plane = (size_t)nlon * nlat;
    const int bx = tune_int("FALWA_STAGE1_BX", 32), by = 32 * kStage1Rows / bx;
    dim3 block(bx, by), grid(ceil_div(nlon, bx), ceil_div(nlat, by), batch);
    QgpvStats stt;
    stt.part = nullptr; stt.ext = nullptr; stt.nwx = 0;
    {
        FALWA_KERNEL("interp_qgpv_kernel", st);
        static const int pfd = tune_int("FALWA_STAGE1_PFD", 6);
        auto go = [&](auto kern) {
            kern<<<grid, block, 0, st>>>(nlon, nlat, nlev, kmax, jd, uin, vin, tin, lo, whi, wlo, fac, tb, d_prof6, uo, vo, to,
                                         avort, pv, plane * nlev, plane * kmax);
        };
        if (mode == 0) {
            if (pfd == 0) go(interp_qgpv_kernel<0, 0>); else if (pfd == 8) go(interp_qgpv_kernel<0, 8>);
            else if (pfd == 10) go(interp_qgpv_kernel<0, 10>); else if (pfd == 14) go(interp_qgpv_kernel<0, 14>); else go(interp_qgpv_kernel<0, 6>);
        } else {
            if (pfd == 0) go(interp_qgpv_kernel<1, 0>); else if (pfd == 8) go(interp_qgpv_kernel<1, 8>);
            else if (pfd == 10) go(interp_qgpv_kernel<1, 10>); else if (pfd == 14) go(interp_qgpv_kernel<1, 14>); else go(interp_qgpv_kernel<1, 6>);
        }
    }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("polar_rows_kernel", st); polar_rows_kernel<<<dim3(kmax, 2, batch), 256, 0, st>>>(nlon, nlat, kmax, mode == 1, avort, pv, plane * kmax, stt); }
    FALWA_LAUNCH_CHECK();
    if (mode == 0) {
        { FALWA_KERNEL("nh18_finish_kernel", st); nh18_finish_kernel<<<dim3(nlat, kmax - 2, batch), 256, 0, st>>>(nlon, nlat, kmax, avort, pv, tb, plane * kmax, stt); }
        FALWA_LAUNCH_CHECK();
    }
    return 0;
}

int interp_launch(int nlon, int nlat, int nlev, int kmax, int batch, const double* uin, const double* vin,
                  const double* tin, const int* lo, const double* whi, const double* wlo, const double* fac, double* uo,
                  double* vo, double* to, cudaStream_t st) {
    const size_t plane = (size_t)nlon * nlat;
    { FALWA_KERNEL("interp_kernel", st); interp_kernel<<<dim3(ceil_div((long long)plane, 256), batch), 256, 0, st>>>(
        nlon, nlat, nlev, kmax, uin, vin, tin, lo, whi, wlo, fac, uo, vo, to, plane * nlev, plane * kmax); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

int theta_means_launch(int nlon, int nlat, int nlev, int jd, int batch, const double* tin, const double* fac,
                       const double* clat, double* rowmean, double* out, cudaStream_t st) {
    { FALWA_KERNEL("theta_rowmean_kernel", st); theta_rowmean_kernel<<<dim3(ceil_div(nlat, 8), nlev, batch), 256, 0, st>>>(nlon, nlat, nlev, tin, fac, clat, rowmean,
                                                                  (size_t)nlon * nlat * nlev); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("theta_hemi_kernel", st); theta_hemi_kernel<<<dim3(nlev, 2, batch), 128, 0, st>>>(nlat, nlev, jd, rowmean, clat, out); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

}  // namespace falwa

using namespace falwa;

// ================================ C ABI ========================================================
extern "C" int falwa_b200_compute_qgpv(int nlon, int nlat, int kmax, const double* ut, const double* vt,
                                       const double* theta, const double* height_host, const double* t0_host,
                                       const double* stat_host, double aa, double omega, double dz, double hh,
                                       double rr, double cp, double* pv, double* avort, void* stream) {
    (void)dz; (void)rr; (void)cp;
    if (!ut || !vt || !theta || !height_host || !t0_host || !stat_host || !pv || !avort) return FALWA_ERR_NULL;
    if (nlon < 3 || nlat < 3 || kmax < 3) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<double> h;
    qgpv_tables_host(0, nlat, kmax, height_host, aa, omega, hh, h);
    const size_t o_prof = h.size();
    for (const double* src : {t0_host, t0_host, stat_host, stat_host}) h.insert(h.end(), src, src + kmax);
    DeviceTable tb;
    int rc = tb.upload(h, st);
    if (rc) return rc;
    return qgpv_launch(0, nlon, nlat, kmax, nlat, 1, ut, vt, theta, tb.d, tb.d + o_prof, pv, avort, st);
}

extern "C" int falwa_b200_compute_qgpv_direct_inv(int nlon, int nlat, int kmax, int jd, const double* uq,
                                                  const double* vq, const double* tq, const double* height_host,
                                                  const double* ts0_host, const double* tn0_host,
                                                  const double* stats_host, const double* statn_host, double aa,
                                                  double omega, double dz, double hh, double rr, double cp,
                                                  double* pv, double* avort, void* stream) {
    (void)dz; (void)rr; (void)cp;
    if (!uq || !vq || !tq || !height_host || !ts0_host || !tn0_host || !stats_host || !statn_host || !pv || !avort)
        return FALWA_ERR_NULL;
    if (nlon < 3 || nlat < 3 || kmax < 3 || jd < 0 || jd > nlat) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<double> h;
    qgpv_tables_host(1, nlat, kmax, height_host, aa, omega, hh, h);
    const size_t o_prof = h.size();
    for (const double* src : {ts0_host, tn0_host, stats_host, statn_host}) h.insert(h.end(), src, src + kmax);
    DeviceTable tb;
    int rc = tb.upload(h, st);
    if (rc) return rc;
    return qgpv_launch(1, nlon, nlat, kmax, jd, 1, uq, vq, tq, tb.d, tb.d + o_prof, pv, avort, st);
}
