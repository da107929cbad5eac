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
__global__ void __launch_bounds__(256)
baro_hist_kernel(int nlon, int nlat, int n, const double* __restrict__ vort, const double* __restrict__ area,
                 const unsigned long long* __restrict__ ext, double* __restrict__ hist /* [B][n] */) {
    extern __shared__ double sh[];  // [n]
    const int b = blockIdx.y;
    const size_t npts = (size_t)nlon * nlat;
    const double* v = vort + (size_t)b * npts;
    for (int t = threadIdx.x; t < n; t += blockDim.x) sh[t] = 0.0;
    __syncthreads();
    const double vmax = ordered_to_f64(ext[b * 2 + 0]), vmin = ordered_to_f64(ext[b * 2 + 1]);
    const double step = (vmax - vmin) / (double)(n - 1);
    const double rstep = 1. / step;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < npts; t += (size_t)gridDim.x * blockDim.x) {
        const double x = v[t];
        // searchsorted(bins, x, side='right'): number of bins <= x.  Guess from the uniform spacing, then settle
        // against the exact numpy bin edges.
        int idx = min(max(__double2int_rz((x - vmin) * rstep) + 1, 0), n);
        while (idx > 0 && linspace_at(vmin, vmax, step, n, idx - 1) > x) --idx;
        while (idx < n && linspace_at(vmin, vmax, step, n, idx) <= x) ++idx;
        if (idx >= 1 && idx <= n - 1) atomicAdd(&sh[idx], area[t]);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x)
        if (sh[t] != 0.0) atomicAdd(&hist[(size_t)b * n + t], sh[t]);
}

// cumulative area -> equivalent latitude of every bin edge -> Q(ylat) by np.interp (basis.py:48-53)
__global__ void __launch_bounds__(256)
baro_qref_kernel(int nlat, int n, double planet_radius, double pi, const double* __restrict__ hist,
                 const unsigned long long* __restrict__ ext, const double* __restrict__ ylat,
                 double* __restrict__ qref /* [B][nlat] */) {
    extern __shared__ double sh[];  // eqv[n]
    const int b = blockIdx.x;
    double* eqv = sh;
    const double vmax = ordered_to_f64(ext[b * 2 + 0]), vmin = ordered_to_f64(ext[b * 2 + 1]);
    const double step = (vmax - vmin) / (double)(n - 1);
    if (threadIdx.x == 0) {
        double aq = 0.0;
        for (int i = 0; i < n; ++i) {
            if (i >= 1) aq = aq + hist[(size_t)b * n + i];
            const double g = aq / (2 * pi * planet_radius * planet_radius) - 1.0;
            eqv[i] = asin(g) * (180.0 / pi);   // np.rad2deg(np.arcsin(.))
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nlat; j += blockDim.x) {
        const double x = ylat[j];
        double r;
        if (x > eqv[n - 1]) r = vmax;                                     // np.interp: right = fp[-1]
        else if (x < eqv[0]) r = vmin;                                    //            left  = fp[0]
        else {
            int lo = 0, hi = n;                                           // largest i with eqv[i] <= x
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (eqv[mid] <= x) lo = mid + 1; else hi = mid; }
            const int i = lo - 1;
            const double f0 = linspace_at(vmin, vmax, step, n, i);
            if (i == n - 1 || eqv[i] == x) r = f0;
            else {
                const double f1 = linspace_at(vmin, vmax, step, n, i + 1);
                const double slope = (f1 - f0) / (eqv[i + 1] - eqv[i]);
                r = slope * (x - eqv[i]) + f0;
            }
        }
        qref[(size_t)b * nlat + j] = r;
    }
}

// basis.lwa: out(j, i) = sum_{jj > j, q < Q_j} -(q - Q_j) dmu_jj  +  sum_{jj <= j, q > Q_j} (q - Q_j) dmu_jj
// one thread per (i, j); rows are summed in ascending order like np.sum(axis=0).  out: [B][2][nlat][nlon] (anti,
// cyclonic) when partitioned, else [B][nlat][nlon].
__global__ void __launch_bounds__(256)
baro_lwa_kernel(int nlon, int nlat, const double* __restrict__ vort, const double* __restrict__ qref,
                const double* __restrict__ dmu, double* __restrict__ out, int partitioned) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (i >= nlon || j >= nlat) return;
    const size_t npts = (size_t)nlon * nlat;
    const double* v = vort + (size_t)b * npts + i;
    double anti = 0.0, cyc = 0.0;
    if (j < nlat - 1) {
        const double Q = qref[(size_t)b * nlat + j];
        for (int jj = 0; jj <= j; ++jj) {
            const double qe = v[(size_t)jj * nlon] - Q;
            if (qe > 0.) cyc = cyc + qe * dmu[jj];
        }
        for (int jj = j + 1; jj < nlat; ++jj) {
            const double qe = v[(size_t)jj * nlon] - Q;
            if (qe < 0.) anti = anti + (-qe) * dmu[jj];
        }
    }
    const size_t o = (size_t)j * nlon + i;
    if (partitioned) {
        out[(size_t)b * 2 * npts + o] = anti;
        out[(size_t)b * 2 * npts + npts + o] = cyc;
    } else {
        out[(size_t)b * npts + o] = anti + cyc;
    }
}

}  // namespace falwa

using namespace falwa;

// replaces BarotropicField.equivalent_latitudes / .lwa (barotropic_field.py:104-125) for a batch of fields.
//   vort [batch][nlat][nlon]; area [nlat][nlon]; ylat_host[nlat] (degrees, ascending); dmu_host[nlat] = a cos(lat) dphi
//   qref [batch][nlat]; lwa [batch][nlat][nlon], or [batch][2][nlat][nlon] (anticyclonic, cyclonic) when partitioned.
extern "C" int falwa_b200_barotropic_eqvlat_lwa(int batch, int nlon, int nlat, int n_points, const double* vort,
                                                const double* area, const double* ylat_host, const double* dmu_host,
                                                double planet_radius, double* qref, double* lwa, int partitioned,
                                                void* stream) {
    if (!vort || !area || !ylat_host || !dmu_host || !qref) return FALWA_ERR_NULL;
    if (batch < 1 || nlon < 1 || nlat < 2 || n_points < 2) return FALWA_ERR_ARG;
    if ((size_t)n_points * sizeof(double) > 200 * 1024) return FALWA_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int npts = nlon * nlat;
    int rc;
    std::vector<double> tab(ylat_host, ylat_host + nlat);
    tab.insert(tab.end(), dmu_host, dmu_host + nlat);
    DeviceTable T;
    if ((rc = T.upload(tab, st))) return rc;
    DeviceScratch<unsigned long long> ext;
    DeviceScratch<double> hist;
    if ((rc = ext.alloc((size_t)batch * 2, st))) return rc;
    if ((rc = hist.alloc((size_t)batch * n_points, st, true))) return rc;
    { FALWA_KERNEL("init_ext_kernel", st); init_ext_kernel<<<ceil_div(batch * 2, 128), 128, 0, st>>>(ext.d, batch * 2); }
    const int nblk = std::max(1, std::min(ceil_div(npts, 256 * 8), 148 * 8 / std::max(1, std::min(batch, 148 * 8))));
    { FALWA_KERNEL("baro_minmax_kernel", st); baro_minmax_kernel<<<dim3(nblk, batch), 256, 0, st>>>(npts, vort, ext.d); }
    FALWA_LAUNCH_CHECK();
    const size_t smem = sizeof(double) * n_points;
    if (smem > 48 * 1024) {
        FALWA_CUDA_TRY(cudaFuncSetAttribute(baro_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FALWA_CUDA_TRY(cudaFuncSetAttribute(baro_qref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    { FALWA_KERNEL("baro_hist_kernel", st); baro_hist_kernel<<<dim3(nblk, batch), 256, smem, st>>>(nlon, nlat, n_points, vort, area, ext.d, hist.d); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("baro_qref_kernel", st); baro_qref_kernel<<<batch, 256, smem, st>>>(nlat, n_points, planet_radius, host_pi(), hist.d, ext.d, T.d, qref); }
    FALWA_LAUNCH_CHECK();
    if (lwa) {
        dim3 block(32, 8), grid(ceil_div(nlon, 32), ceil_div(nlat, 8), batch);
        { FALWA_KERNEL("baro_lwa_kernel", st); baro_lwa_kernel<<<grid, block, 0, st>>>(nlon, nlat, vort, qref, T.d + nlat, lwa, partitioned); }
        FALWA_LAUNCH_CHECK();
    }
    return 0;
}
