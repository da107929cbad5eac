// 2-D (barotropic) path: equivalent latitude and local wave activity of absolute-vorticity fields, batched.
//
// Reference behaviour restated (not ported): src/falwa/barotropic_field.py:44-125 (BarotropicField),
// src/falwa/basis.py:11-64 (eqvlat_fawa, output_fawa=False) and :137-205 (lwa).  Conventions that differ from the
// 3-D routines: bins are np.linspace(min, max, n) with np.digitize membership (right-open; the maximum
// drops out), the cumulative area starts at the LOW-vorticity end, row j itself belongs to the cyclonic branch, both
// inequalities are strict, and the last row stays 0.
#include "common.cuh"

namespace falwa {

// np.linspace(vmin, vmax, n)[i]: arange(n) * step + start with the last element set to stop (numpy/core/function_base.py)
__device__ __forceinline__ double linspace_at(double vmin, double vmax, double step, int n, int i) {
    return (i == n - 1) ? vmax : (double)i * step + vmin;
}

__global__ void __launch_bounds__(256)
baro_minmax_kernel(int npts, const double* __restrict__ vort, unsigned long long* __restrict__ ext) {
    const double* v = vort + (size_t)blockIdx.y * npts;
    double mx = -INFINITY, mn = INFINITY;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < npts; t += gridDim.x * blockDim.x) {
        const double x = v[t];
        mx = fmax(mx, x); mn = fmin(mn, x);
    }
    mx = warp_max(mx); mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&ext[blockIdx.y * 2 + 0], f64_to_ordered(mx));
        atomicMin(&ext[blockIdx.y * 2 + 1], f64_to_ordered(mn));
    }
}

// area per vorticity bin: inds = np.digitize(vort, bins) in 1..n-1 are kept (basis.py:47-48)
// WEIGHTED: a second histogram of area * w (w = the field itself for basis.eqvlat_fawa's cq, basis.py:49-50; the
// Laplacian-type field of basis.eqvlat_vgrad, basis.py:119-120) into the slots [n .. 2n).
//
// Bit-reproducible: no floating-point atomics.  Every warp owns a contiguous range of grid points and a private
// shared-memory histogram; the lanes of a 32-point batch that fall into the same bin are combined by the lowest such
// lane in lane order (__match_any_sync + shuffles), the warp histograms are added in warp order into a per-CTA
// partial, and baro_hist_combine_kernel adds the partials in CTA order.  The reference's own test relies on
// eqvlat_vgrad giving identical Q(y) with and without vgrad (tests/test_basis.py:41-58).
template <bool WEIGHTED>
__global__ void __launch_bounds__(256)
baro_hist_kernel(int nlon, int nlat, int n, const double* __restrict__ vort, const double* __restrict__ area,
                 const double* __restrict__ wfield, const unsigned long long* __restrict__ ext,
                 double* __restrict__ partial /* [B][gridDim.x][NH * n] */) {
    extern __shared__ double sh[];  // [warps][NH * n]
    constexpr int NH = WEIGHTED ? 2 : 1;
    const unsigned FULL = 0xffffffffu;
    const int b = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const size_t npts = (size_t)nlon * nlat;
    const double* v = vort + (size_t)b * npts;
    const double* wf = WEIGHTED ? wfield + (size_t)b * npts : nullptr;
    for (int t = threadIdx.x; t < nw * NH * n; t += blockDim.x) sh[t] = 0.0;
    __syncthreads();
    double* hw = sh + (size_t)w * NH * n;
    const double vmax = ordered_to_f64(ext[b * 2 + 0]), vmin = ordered_to_f64(ext[b * 2 + 1]);
    const double step = (vmax - vmin) / (double)(n - 1);
    const double rstep = 1. / step;
    const size_t nwarps = (size_t)gridDim.x * nw, gw = (size_t)blockIdx.x * nw + w;
    const size_t chunk = ((npts + nwarps - 1) / nwarps + 31) / 32 * 32;
    const size_t t0 = gw * chunk, t1 = (t0 + chunk < npts) ? t0 + chunk : npts;
    for (size_t base = t0; base < t1; base += 32) {
        const size_t t = base + lane;
        int idx = 0;
        double a = 0.0, wa = 0.0;
        if (t < t1) {
            const double x = v[t];
            // searchsorted(bins, x, side='right'): number of bins <= x.  Guess from the uniform spacing, then
            // settle against the exact numpy bin edges.
            idx = min(max(__double2int_rz((x - vmin) * rstep) + 1, 0), n);
            while (idx > 0 && linspace_at(vmin, vmax, step, n, idx - 1) > x) --idx;
            while (idx < n && linspace_at(vmin, vmax, step, n, idx) <= x) ++idx;
        }
        const bool keep = (idx >= 1 && idx <= n - 1);
        if (keep) {
            a = area[t];
            if (WEIGHTED) wa = a * wf[t];
        }
        const unsigned m = __match_any_sync(FULL, keep ? idx : n + lane);   // dropped points group with nobody
        const int cnt = __popc(m), leader = __ffs(m) - 1;
        const int maxc = __reduce_max_sync(FULL, cnt);
        double sa = a, sw = wa;
        if (maxc == 32) {        // (uniform) the whole batch lies in one bin: fixed shuffle tree
            sa = warp_sum(a);
            if (WEIGHTED) sw = warp_sum(wa);
        } else if (maxc > 1) {   // (uniform) add the members of every group in lane order
            sa = 0.0; sw = 0.0;
            for (int k = 0; k < maxc; ++k) {
                const unsigned f = __fns(m, 0, k + 1);
                const int src = (f < 32u) ? (int)f : lane;
                const double av = __shfl_sync(FULL, a, src);
                double wv = 0.0;
                if (WEIGHTED) wv = __shfl_sync(FULL, wa, src);
                if (k < cnt) { sa += av; sw += wv; }
            }
        }
        if (keep && lane == leader) {
            hw[idx] += sa;
            if (WEIGHTED) hw[n + idx] += sw;
        }
        __syncwarp();
    }
    __syncthreads();
    double* out = partial + ((size_t)b * gridDim.x + blockIdx.x) * NH * n;
    for (int t = threadIdx.x; t < NH * n; t += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nw; ++q) s += sh[(size_t)q * NH * n + t];
        out[t] = s;
    }
}

// hist[b][i] = sum of the CTA partials in CTA order
__global__ void __launch_bounds__(256)
baro_hist_combine_kernel(int nslot, int nblk, const double* __restrict__ partial, double* __restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= nslot) return;
    double s = 0.0;
    for (int c = 0; c < nblk; ++c) s += partial[((size_t)b * nblk + c) * nslot + i];
    hist[(size_t)b * nslot + i] = s;
}

// zonal mean of the field and zonal sum of the area per row (basis.py:39, :58-59); one warp per row
__global__ void __launch_bounds__(256)
baro_rowstats_kernel(int nlon, int nlat, const double* __restrict__ vort, const double* __restrict__ area,
                     double* __restrict__ qbar /* [B][nlat] */, double* __restrict__ asum /* [nlat] */) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
    if (row >= nlat) return;
    const double* v = vort + ((size_t)b * nlat + row) * nlon;
    const double* a = area + (size_t)row * nlon;
    double s = 0.0, t = 0.0;
    for (int i = lane; i < nlon; i += 32) { s += v[i]; t += a[i]; }
    s = warp_sum(s); t = warp_sum(t);
    if (lane == 0) {
        qbar[(size_t)b * nlat + row] = s / (double)nlon;
        if (b == 0) asum[row] = t;
    }
}

// np.interp(x, xp, fp) for non-decreasing xp[0..n); fp given by a callable
template <typename F>
__device__ __forceinline__ double np_interp(double x, const double* xp, int n, F fp) {
    if (x > xp[n - 1]) return fp(n - 1);                              // right = fp[-1]
    if (x < xp[0]) return fp(0);                                      // left  = fp[0]
    int lo = 0, hi = n;                                               // largest i with xp[i] <= x
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (xp[mid] <= x) lo = mid + 1; else hi = mid; }
    const int i = lo - 1;
    const double f0 = fp(i);
    if (i == n - 1 || xp[i] == x) return f0;
    const double slope = (fp(i + 1) - f0) / (xp[i + 1] - xp[i]);
    return slope * (x - xp[i]) + f0;
}

// cumulative area -> equivalent latitude of every bin edge -> Q(ylat) by np.interp (basis.py:48-53).
// mode 0 with aux: finite-amplitude wave activity (basis.py:55-62); mode 1 with aux: the bracket average of
// basis.eqvlat_vgrad (basis.py:112-131).  hist is [B][n] (aux == NULL) or [B][2n].
__global__ void __launch_bounds__(256)
baro_qref_kernel(int mode, int nlat, int n, double planet_radius, double pi, const double* __restrict__ hist,
                 const unsigned long long* __restrict__ ext, const double* __restrict__ ylat,
                 const double* __restrict__ qbar, const double* __restrict__ asum,
                 double* __restrict__ qref /* [B][nlat] */, double* __restrict__ aux) {
    extern __shared__ double sh[];  // eqv[n], f2[n], cref[nlat]
    const int b = blockIdx.x;
    double* eqv = sh;
    double* f2 = sh + n;
    double* cref = f2 + n;
    const int nh = aux ? 2 : 1;
    const double* H = hist + (size_t)b * nh * n;
    const double vmax = ordered_to_f64(ext[b * 2 + 0]), vmin = ordered_to_f64(ext[b * 2 + 1]);
    const double step = (vmax - vmin) / (double)(n - 1);
    if (threadIdx.x == 0) {
        double aq = 0.0, cq = 0.0;
        for (int i = 0; i < n; ++i) {
            if (i >= 1) aq = aq + H[i];
            const double g = aq / (2 * pi * planet_radius * planet_radius) - 1.0;
            eqv[i] = asin(g) * (180.0 / pi);   // np.rad2deg(np.arcsin(.))
            if (aux && mode == 0) {            // cq: cumulative area * vorticity
                if (i >= 1) cq = cq + H[n + i];
                f2[i] = cq;
            }
        }
        if (aux && mode == 1) {                // brac[1:-1] = 0.5 (dp[:-2] + dp[2:]),  dp = sum(area*vgrad)/area per bin
            f2[0] = 0.0; f2[n - 1] = 0.0;
            for (int i = 1; i < n - 1; ++i) {
                const double d0 = (H[i - 1] == 0.) ? 0. : H[n + i - 1] / H[i - 1];
                const double d2 = (H[i + 1] == 0.) ? 0. : H[n + i + 1] / H[i + 1];
                f2[i] = 0.5 * (d0 + d2);
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nlat; j += blockDim.x) {
        const double x = ylat[j];
        qref[(size_t)b * nlat + j] = np_interp(x, eqv, n, [&](int i) { return linspace_at(vmin, vmax, step, n, i); });
        if (aux) {
            const double r = np_interp(x, eqv, n, [&](int i) { return f2[i]; });
            if (mode == 1) aux[(size_t)b * nlat + j] = r; else cref[j] = r;
        }
    }
    if (!(aux && mode == 0)) return;
    __syncthreads();
    if (threadIdx.x == 0) {   // cbar from the pole downwards; fawa = -(cbar[1:] + cref[1:]) / (2 pi a)
        double cbar = 0.0;
        double* fawa = aux + (size_t)b * (nlat - 1);
        const double* qb = qbar + (size_t)b * nlat;
        fawa[nlat - 2] = -(0.0 + cref[nlat - 1]) / (2 * pi * planet_radius);
        for (int j = nlat - 2; j >= 1; --j) {
            cbar = cbar + 0.5 * (qb[j + 1] * asum[j + 1] + qb[j] * asum[j]);
            fawa[j - 1] = -(cbar + cref[j]) / (2 * pi * planet_radius);
        }
    }
}

// basis.lwa: out(j, i) = sum_{jj > j, q < Q_j} -(q - Q_j) dmu_jj  +  sum_{jj <= j, q > Q_j} (q - Q_j) dmu_jj
// one thread per (i, j); rows are summed in ascending order like np.sum(axis=0).  out: [B][2][nlat][nlon] (anti,
// cyclonic) when partitioned, else [B][nlat][nlon].
// NC: also bigsigma(j, i) = sum ncforce * (anticyclonic_mask + cyclonic_mask) * dmu  (basis.py:194-196)
template <bool NC>
__global__ void __launch_bounds__(256)
baro_lwa_kernel(int nlon, int nlat, const double* __restrict__ vort, const double* __restrict__ qref,
                const double* __restrict__ dmu, const double* __restrict__ ncforce, double* __restrict__ out,
                double* __restrict__ bigsigma, int partitioned) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (i >= nlon || j >= nlat) return;
    const size_t npts = (size_t)nlon * nlat;
    const double* v = vort + (size_t)b * npts + i;
    const double* nc = NC ? ncforce + (size_t)b * npts + i : nullptr;
    double anti = 0.0, cyc = 0.0, sig = 0.0;
    if (j < nlat - 1) {
        const double Q = qref[(size_t)b * nlat + j];
        for (int jj = 0; jj <= j; ++jj) {
            const double qe = v[(size_t)jj * nlon] - Q;
            if (qe > 0.) {
                cyc = cyc + qe * dmu[jj];
                if (NC) sig = sig + nc[(size_t)jj * nlon] * dmu[jj];
            }
        }
        for (int jj = j + 1; jj < nlat; ++jj) {
            const double qe = v[(size_t)jj * nlon] - Q;
            if (qe < 0.) {
                anti = anti + (-qe) * dmu[jj];
                if (NC) sig = sig + (-nc[(size_t)jj * nlon]) * dmu[jj];
            }
        }
    }
    const size_t o = (size_t)j * nlon + i;
    if (NC) bigsigma[(size_t)b * npts + o] = sig;
    if (partitioned) {
        out[(size_t)b * 2 * npts + o] = anti;
        out[(size_t)b * 2 * npts + npts + o] = cyc;
    } else {
        out[(size_t)b * npts + o] = anti + cyc;
    }
}

}  // namespace falwa

using namespace falwa;

static int eqvlat_2d_launch(int mode, int batch, int nlon, int nlat, int n_points, const double* vort,
                            const double* area, const double* wfield, const double* ylat_host, double planet_radius,
                            double* qref, double* aux, cudaStream_t st) {
    const int npts = nlon * nlat;
    const int nh = aux ? 2 : 1;
    int rc;
    std::vector<double> tab(ylat_host, ylat_host + nlat);
    DeviceTable T;
    if ((rc = T.upload(tab, st))) return rc;
    DeviceScratch<unsigned long long> ext;
    DeviceScratch<double> hist, rows;
    if ((rc = ext.alloc((size_t)batch * 2, st))) return rc;
    if ((rc = hist.alloc((size_t)batch * nh * n_points, st, true))) return rc;
    if ((rc = rows.alloc((size_t)(batch + 1) * nlat, st, true))) return rc;
    { FALWA_KERNEL("init_ext_kernel", st); init_ext_kernel<<<ceil_div(batch * 2, 128), 128, 0, st>>>(ext.d, batch * 2); }
    const int nblk = std::max(1, std::min(ceil_div(npts, 256 * 8), std::max(8, 148 * 8 / std::max(1, batch))));
    { FALWA_KERNEL("baro_minmax_kernel", st); baro_minmax_kernel<<<dim3(nblk, batch), 256, 0, st>>>(npts, vort, ext.d); }
    FALWA_LAUNCH_CHECK();
    // warps per CTA of the histogram kernel: as many private histograms as fit into shared memory
    int hwarps = 8;
    while (hwarps > 1 && sizeof(double) * hwarps * nh * n_points > 200 * 1024) hwarps >>= 1;
    const size_t smem_h = sizeof(double) * hwarps * nh * n_points;
    if (smem_h > 220 * 1024) return FALWA_ERR_UNSUPPORTED;
    const size_t smem_q = sizeof(double) * (2 * (size_t)n_points + nlat);
    if (smem_h > 48 * 1024) {
        FALWA_CUDA_TRY(cudaFuncSetAttribute(baro_hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
        FALWA_CUDA_TRY(cudaFuncSetAttribute(baro_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
    }
    if (smem_q > 48 * 1024)
        FALWA_CUDA_TRY(cudaFuncSetAttribute(baro_qref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
    DeviceScratch<double> partial;
    if ((rc = partial.alloc((size_t)batch * nblk * nh * n_points, st))) return rc;
    {
        FALWA_KERNEL("baro_hist_kernel", st);
        if (aux) baro_hist_kernel<true><<<dim3(nblk, batch), 32 * hwarps, smem_h, st>>>(nlon, nlat, n_points, vort, area, wfield, ext.d, partial.d);
        else baro_hist_kernel<false><<<dim3(nblk, batch), 32 * hwarps, smem_h, st>>>(nlon, nlat, n_points, vort, area, nullptr, ext.d, partial.d);
    }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("baro_hist_combine_kernel", st); baro_hist_combine_kernel<<<dim3(ceil_div(nh * n_points, 256), batch), 256, 0, st>>>(nh * n_points, nblk, partial.d, hist.d); }
    FALWA_LAUNCH_CHECK();
    if (aux && mode == 0) {
        FALWA_KERNEL("baro_rowstats_kernel", st);
        baro_rowstats_kernel<<<dim3(ceil_div(nlat, 8), batch), 256, 0, st>>>(nlon, nlat, vort, area, rows.d, rows.d + (size_t)batch * nlat);
        FALWA_LAUNCH_CHECK();
    }
    { FALWA_KERNEL("baro_qref_kernel", st); baro_qref_kernel<<<batch, 256, smem_q, st>>>(mode, nlat, n_points, planet_radius, host_pi(), hist.d, ext.d, T.d, rows.d, rows.d + (size_t)batch * nlat, qref, aux); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

static int lwa_2d_launch(int batch, int nlon, int nlat, const double* vort, const double* qref, const double* dmu_dev,
                         const double* ncforce, double* lwa, double* bigsigma, int partitioned, cudaStream_t st) {
    dim3 block(32, 8), grid(ceil_div(nlon, 32), ceil_div(nlat, 8), batch);
    FALWA_KERNEL("baro_lwa_kernel", st);
    if (ncforce) baro_lwa_kernel<true><<<grid, block, 0, st>>>(nlon, nlat, vort, qref, dmu_dev, ncforce, lwa, bigsigma, partitioned);
    else baro_lwa_kernel<false><<<grid, block, 0, st>>>(nlon, nlat, vort, qref, dmu_dev, nullptr, lwa, nullptr, partitioned);
    FALWA_LAUNCH_CHECK();
    return 0;
}

// basis.eqvlat_fawa (mode 0; basis.py:11-64) and basis.eqvlat_vgrad (mode 1; basis.py:67-134) for a batch of fields.
//   mode 0: aux = fawa [batch][nlat-1] or NULL (output_fawa=False);  mode 1: aux = brac [batch][nlat] and vgrad
//   [batch][nlat][nlon], or both NULL.
extern "C" int falwa_b200_eqvlat_2d(int mode, int batch, int nlon, int nlat, int n_points, const double* vort,
                                    const double* area, const double* vgrad, const double* ylat_host,
                                    double planet_radius, double* qref, double* aux, void* stream) {
    if (!vort || !area || !ylat_host || !qref) return FALWA_ERR_NULL;
    if (mode < 0 || mode > 1 || batch < 1 || nlon < 1 || nlat < 2 || n_points < 2) return FALWA_ERR_ARG;
    if (mode == 1 && ((aux != nullptr) != (vgrad != nullptr))) return FALWA_ERR_ARG;
    if ((2 * (size_t)n_points + nlat) * sizeof(double) > 200 * 1024) return FALWA_ERR_UNSUPPORTED;
    return eqvlat_2d_launch(mode, batch, nlon, nlat, n_points, vort, area, mode == 0 ? vort : vgrad, ylat_host,
                            planet_radius, qref, aux, (cudaStream_t)stream);
}

// basis.lwa (basis.py:137-205) with a prescribed Q(y) per field: qref [batch][nlat] (device), dmu_host [nlat];
// ncforce [batch][nlat][nlon] and bigsigma [batch][nlat][nlon] are both given or both NULL.
extern "C" int falwa_b200_lwa_2d(int batch, int nlon, int nlat, const double* vort, const double* qref,
                                 const double* dmu_host, const double* ncforce, double* lwa, double* bigsigma,
                                 int partitioned, void* stream) {
    if (!vort || !qref || !dmu_host || !lwa) return FALWA_ERR_NULL;
    if (batch < 1 || nlon < 1 || nlat < 2) return FALWA_ERR_ARG;
    if ((ncforce != nullptr) != (bigsigma != nullptr)) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<double> tab(dmu_host, dmu_host + nlat);
    DeviceTable T;
    int rc;
    if ((rc = T.upload(tab, st))) return rc;
    return lwa_2d_launch(batch, nlon, nlat, vort, qref, T.d, ncforce, lwa, bigsigma, partitioned, st);
}

// replaces BarotropicField.equivalent_latitudes / .lwa (barotropic_field.py:104-125) for a batch of fields.
//   vort [batch][nlat][nlon]; area [nlat][nlon]; ylat_host[nlat] (degrees, ascending); dmu_host[nlat] = a cos(lat) dphi
//   qref [batch][nlat]; lwa [batch][nlat][nlon], or [batch][2][nlat][nlon] (anticyclonic, cyclonic) when partitioned.
extern "C" int falwa_b200_barotropic_eqvlat_lwa(int batch, int nlon, int nlat, int n_points, const double* vort,
                                                const double* area, const double* ylat_host, const double* dmu_host,
                                                double planet_radius, double* qref, double* lwa, int partitioned,
                                                void* stream) {
    if (!vort || !area || !ylat_host || !dmu_host || !qref) return FALWA_ERR_NULL;
    if (batch < 1 || nlon < 1 || nlat < 2 || n_points < 2) return FALWA_ERR_ARG;
    if ((2 * (size_t)n_points + nlat) * sizeof(double) > 200 * 1024) return FALWA_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = eqvlat_2d_launch(0, batch, nlon, nlat, n_points, vort, area, nullptr, ylat_host, planet_radius, qref,
                               nullptr, st))) return rc;
    if (lwa) {
        std::vector<double> tab(dmu_host, dmu_host + nlat);
        DeviceTable T;
        if ((rc = T.upload(tab, st))) return rc;
        if ((rc = lwa_2d_launch(batch, nlon, nlat, vort, qref, T.d, nullptr, lwa, nullptr, partitioned, st))) return rc;
    }
    return 0;
}
