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

struct QgpvTables {
    // per latitude row (nlat each): av1 = 2*omega*sin(phi0), den = 2*aa*cos(phi0)*dphi, cosp, cosm
    const double* av1;
    const double* den;
    const double* cosp;
    const double* cosm;
    // per level (kmax each): ep[k] = exp(-height[k]/hh), e0[k] = exp(height[k]/hh), height[k]
    const double* ep;
    const double* e0;
    const double* height;
};
// Per-snapshot profiles live in a separate device array prof[batch][4][kmax]:
//   row 0 = t0 south (rows j <= jd), 1 = t0 north, 2 = static stability south, 3 = north.

// mode 0: NH18  — pv receives (altp - altm), finished by nh18_finish_kernel (needs the zonal-mean avort)
// mode 1: NHN22 — pv = avort + e0 * ((altp - altm) * f / dh)
template <int MODE>
__global__ void __launch_bounds__(256)
avort_pv_kernel(int nlon, int nlat, int kmax, int jd, int kchunk, const double* __restrict__ u,
                const double* __restrict__ v, const double* __restrict__ th, QgpvTables tb,
                const double* __restrict__ prof, double* __restrict__ avort, double* __restrict__ pv,
                size_t batch_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;  // 0-based row
    const int nchunks = (kmax + kchunk - 1) / kchunk;
    const int chunk = blockIdx.z % nchunks;
    const int b = blockIdx.z / nchunks;
    if (i >= nlon || j >= nlat) return;
    const size_t plane = (size_t)nlon * nlat;
    const size_t base = (size_t)b * batch_stride;
    u += base; v += base; th += base; avort += base; pv += base;
    const int k0 = chunk * kchunk;
    const int k1 = min(kmax, k0 + kchunk);
    const bool interior_row = (j >= 1 && j <= nlat - 2);
    const int ip = (i == nlon - 1) ? 0 : i + 1;
    const int im = (i == 0) ? nlon - 1 : i - 1;
    const double av1 = tb.av1[j], den = tb.den[j], cp = tb.cosp[j], cm = tb.cosm[j];
    const int hemi = (j + 1 <= jd) ? 0 : 1;
    const double* t0 = prof + (size_t)b * 4 * kmax + hemi * kmax;
    const double* stat = t0 + 2 * kmax;
    const size_t col = (size_t)j * nlon + i;
    // rolling theta registers: tm = theta(k-1), tp = theta(k+1)
    double tm = (k0 >= 1) ? th[(size_t)(k0 - 1) * plane + col] : 0.0;
    double tc = th[(size_t)k0 * plane + col];
    for (int k = k0; k < k1; ++k) {
        const size_t off = (size_t)k * plane;
        double tp = (k + 1 < kmax) ? th[off + plane + col] : 0.0;
        double av = 0.0;
        if (interior_row) {
            const double av2 = (v[off + (size_t)j * nlon + ip] - v[off + (size_t)j * nlon + im]) / den;
            const double av3 = -(u[off + col + nlon] * cp - u[off + col - nlon] * cm) / den;
            av = av1 + av2 + av3;
        }
        double q = 0.0;
        if (k >= 1 && k <= kmax - 2) {
            const double altp = tb.ep[k + 1] * (tp - t0[k + 1]) / stat[k + 1];
            const double altm = tb.ep[k - 1] * (tm - t0[k - 1]) / stat[k - 1];
            if (MODE == 0) {
                q = altp - altm;
            } else {
                const double strc = (altp - altm) * av1 / (tb.height[k + 1] - tb.height[k - 1]);
                q = av + tb.e0[k] * strc;
            }
        }
        st_stream(&avort[off + col], av);
        st_stream(&pv[off + col], q);
        tm = tc;
        tc = tp;
    }
}

// Polar rows: avort(:,1,k) / avort(:,nlat,k) = zonal mean of the adjacent row (compute_qgpv.f90:86-93).
// For NHN22 the pole rows of pv (which hold only the stretching term so far) receive that mean too.
__global__ void __launch_bounds__(256)
polar_rows_kernel(int nlon, int nlat, int kmax, int add_to_pv, double* __restrict__ avort,
                  double* __restrict__ pv, size_t batch_stride) {
    __shared__ double red[33];
    const int k = blockIdx.x;
    const int pole = blockIdx.y;  // 0 south, 1 north
    const size_t base = (size_t)blockIdx.z * batch_stride + (size_t)k * nlon * nlat;
    const int jsrc = pole ? nlat - 2 : 1;
    const int jdst = pole ? nlat - 1 : 0;
    double s = 0.0;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) s += avort[base + (size_t)jsrc * nlon + i] / (double)nlon;
    s = block_sum(s, red);
    const bool interior = (k >= 1 && k <= kmax - 2);
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) {
        avort[base + (size_t)jdst * nlon + i] = s;
        if (add_to_pv && interior) pv[base + (size_t)jdst * nlon + i] = s + pv[base + (size_t)jdst * nlon + i];
    }
}

// NH18: pv = avort + e0 * ((altp-altm) * zonal_mean(avort) / dh)   (compute_qgpv.f90:98-124)
__global__ void __launch_bounds__(256)
nh18_finish_kernel(int nlon, int nlat, int kmax, const double* __restrict__ avort, double* __restrict__ pv,
                   QgpvTables tb, size_t batch_stride) {
    __shared__ double red[33];
    const int j = blockIdx.x, k = blockIdx.y + 1;  // interior levels only
    const size_t row = (size_t)blockIdx.z * batch_stride + ((size_t)k * nlat + j) * nlon;
    double s = 0.0;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) s += avort[row + i] / (double)nlon;
    const double zm = block_sum(s, red);
    const double dh = tb.height[k + 1] - tb.height[k - 1];
    const double e0 = tb.e0[k];
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) {
        const double strc = pv[row + i] * zm / dh;
        pv[row + i] = avort[row + i] + e0 * strc;
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
    h.assign((size_t)4 * nlat + 3 * kmax, 0.0);
    double *av1 = h.data(), *den = av1 + nlat, *cosp = den + nlat, *cosm = cosp + nlat;
    double *ep = cosm + nlat, *e0 = ep + kmax, *hg = e0 + kmax;
    for (int j = 1; j <= nlat; ++j) {
        const double phi0 = lat(j - 1), phim = lat(j - 2), phip = lat(j);
        av1[j - 1] = 2. * omega * std::sin(phi0);
        den[j - 1] = 2. * aa * std::cos(phi0) * dphi;
        cosp[j - 1] = std::cos(phip);
        cosm[j - 1] = std::cos(phim);
    }
    for (int k = 0; k < kmax; ++k) {
        ep[k] = std::exp(-height[k] / hh);
        e0[k] = std::exp(height[k] / hh);
        hg[k] = height[k];
    }
    return 0;
}

static QgpvTables view_tables(const double* d, int nlat, int kmax) {
    QgpvTables t;
    t.av1 = d; t.den = d + nlat; t.cosp = d + 2 * nlat; t.cosm = d + 3 * nlat;
    t.ep = d + 4 * nlat; t.e0 = t.ep + kmax; t.height = t.e0 + kmax;
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
int qgpv_launch(int mode, int nlon, int nlat, int kmax, int jd, int batch, const double* u, const double* v,
                const double* th, const double* d_tables, const double* d_prof, double* pv, double* avort,
                cudaStream_t st) {
    const QgpvTables tb = view_tables(d_tables, nlat, kmax);
    const size_t bstride = (size_t)nlon * nlat * kmax;
    const int kchunk = pick_kchunk(nlon, nlat, kmax, batch);
    dim3 block(32, 8);
    dim3 grid(ceil_div(nlon, 32), ceil_div(nlat, 8), ceil_div(kmax, kchunk) * batch);
    {
        FALWA_KERNEL("avort_pv_kernel", st);
        if (mode == 0)
            avort_pv_kernel<0><<<grid, block, 0, st>>>(nlon, nlat, kmax, jd, kchunk, u, v, th, tb, d_prof, avort, pv, bstride);
        else
            avort_pv_kernel<1><<<grid, block, 0, st>>>(nlon, nlat, kmax, jd, kchunk, u, v, th, tb, d_prof, avort, pv, bstride);
    }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("polar_rows_kernel", st); polar_rows_kernel<<<dim3(kmax, 2, batch), 256, 0, st>>>(nlon, nlat, kmax, mode == 1, avort, pv, bstride); }
    FALWA_LAUNCH_CHECK();
    if (mode == 0) {
        { FALWA_KERNEL("nh18_finish_kernel", st); nh18_finish_kernel<<<dim3(nlat, kmax - 2, batch), 256, 0, st>>>(nlon, nlat, kmax, avort, pv, tb, bstride); }
        FALWA_LAUNCH_CHECK();
    }
    return 0;
}

int qgpv_tables_host(int mode, int nlat, int kmax, const double* height, double aa, double omega, double hh,
                     std::vector<double>& h) {
    return build_qgpv_tables(nlat, kmax, mode == 0, height, aa, omega, hh, h);
}

size_t qgpv_tables_size(int nlat, int kmax) { return (size_t)4 * nlat + 3 * kmax; }

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
    __shared__ double red[33];
    const int j = blockIdx.x, l = blockIdx.y, b = blockIdx.z;
    const double* row = tin + (size_t)b * in_stride + ((size_t)l * nlat + j) * nlon;
    const double f = fac[l], c = clat[j];
    double s = 0.0;
    for (int i = threadIdx.x; i < nlon; i += blockDim.x) s += (ld_stream(&row[i]) * f) * c;
    s = block_sum(s, red);
    if (threadIdx.x == 0) rowmean[((size_t)b * nlev + l) * nlat + j] = s / (double)nlon;
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
    { FALWA_KERNEL("theta_rowmean_kernel", st); theta_rowmean_kernel<<<dim3(nlat, nlev, batch), 256, 0, st>>>(nlon, nlat, nlev, tin, fac, clat, rowmean,
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
