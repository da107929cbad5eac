// LWA / flux stage: local wave activity (cyclonic + anticyclonic), the non-linear zonal flux F2,
// the remaining zonal-flux and eddy-momentum-flux integrands and the low-level heat flux.
//
// Reference behaviour restated (not ported): f90_modules/compute_flux_dirinv.f90:1-182 and
// f90_modules/compute_lwa_only_nhn22.f90:1-107; vertical averages of oopinterface.py:405-420.
//
// Two LWA algorithms produce the same sums:
//   * brute force  — the reference's O(nd^2) (j, jj) double loop per column, same accumulation
//     order as the CPU oracle (bit-comparable; FP64-ALU bound);
//   * interval scan — O(nd log nd) per column, valid where qref is monotone in latitude
//     (flux_scan.cu); used by the fused path.
#include "common.cuh"

namespace falwa {

struct FluxArgs {
    int imax, jmax, kmax, nd, jb, jd, is_nhem;
    int jstart, jend, joff, uoff;
    double a, om, dz, h, rr, prefac, ab;
    const double *pv, *uu, *vv, *pt, *ncforce;
    const double *qref, *uref, *tref;  // [k][nd], [k][jd], [k][jd]
    // host tables: cosphi[nd+2] = cos(phi(jj)), jj = 0..nd+1 (1-based jj at index jj), sinphi likewise,
    // aaw[nd+2]; per level: ep11a[k] = exp(-rkappa*zz/h), tgdiff[k] = tg(k+1)-tg(k-1)
    const double *cosphi, *sinphi, *aaw, *ep11a, *tgdiff;
    double ep42fac, tg31, tg21, rrh;  // exp(-dz/h), tg(3)-tg(1), tg(2)-tg(1), rr/h
    double *astar1, *astar2, *ncforce3d, *ua1, *ua2, *ep1, *ep2, *ep3, *ep4;
    unsigned long long* mask_count;
};

// one thread per (i, j) of one level; jj loop in the reference's order.
// BRUTE = false: astar1 / astar2 (/ ua2 / ncforce3d) were already written by the interval-scan kernel
// (flux_scan.cu) and only the pointwise flux terms are evaluated here.
template <bool BRUTE>
__global__ void __launch_bounds__(256)
flux_bruteforce_kernel(FluxArgs p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = p.jstart + blockIdx.y * blockDim.y + threadIdx.y;  // 1-based row
    const int k = blockIdx.z + 2;                                     // 1-based level 2..kmax-1
    const bool active = (i < p.imax && j <= p.jend);
    unsigned long long c1 = 0, c2 = 0;
    if (active) {
        const size_t lev = (size_t)(k - 1) * p.jmax * p.imax;
        const double* pvk = p.pv + lev + (size_t)p.joff * p.imax + i;  // row jj (1-based) at (jj-1)*imax
        const double* uuk = p.uu + lev + (size_t)p.joff * p.imax + i;
        const double* nck = p.ncforce ? p.ncforce + lev + (size_t)p.joff * p.imax + i : nullptr;
        const double qr = p.qref[(size_t)(k - 1) * p.nd + (j - 1)];
        const double ur = p.uref[(size_t)(k - 1) * p.jd + (j - p.uoff - 1)];
        const double cphi0 = p.cosphi[j];
        double as_pos = 0., as_neg = 0., ncf = 0., u2 = 0.;
        const size_t o = ((size_t)(k - 1) * p.nd + (j - 1)) * p.imax + i;
        if (!BRUTE) {
            as_neg = p.astar1[o];
            as_pos = p.astar2[o];
        }
        for (int jj = 1; BRUTE && jj < j; ++jj) {  // cyclonic candidates: qe > 0, jj < j
            const double q = pvk[(size_t)(jj - 1) * p.imax];
            const double qe = q - qr;
            if (qe > 0.) {
                const double aa = p.aaw[jj];
                const double ue = uuk[(size_t)(jj - 1) * p.imax] * cphi0 - ur * p.cosphi[jj];
                as_neg = as_neg + qe * aa;
                if (nck) ncf = ncf + nck[(size_t)(jj - 1) * p.imax] * aa;
                u2 = u2 + qe * ue * p.ab;
                ++c1;
            }
        }
        for (int jj = j; BRUTE && jj <= p.nd; ++jj) {  // anticyclonic candidates: qe <= 0, jj >= j
            const double q = pvk[(size_t)(jj - 1) * p.imax];
            const double qe = q - qr;
            if (qe <= 0.) {
                const double aa = p.aaw[jj];
                const double ue = uuk[(size_t)(jj - 1) * p.imax] * cphi0 - ur * p.cosphi[jj];
                as_pos = as_pos - qe * aa;
                if (nck) ncf = ncf - nck[(size_t)(jj - 1) * p.imax] * aa;
                u2 = u2 - qe * ue * p.ab;
                ++c2;
            }
        }
        double a1, a2;
        if (p.is_nhem) { a2 = as_pos; a1 = as_neg; } else { a1 = as_pos; a2 = as_neg; }
        if (BRUTE) {
            p.astar1[o] = a1;
            p.astar2[o] = a2;
            if (p.ncforce3d) p.ncforce3d[o] = ncf;
            if (p.ua2) p.ua2[o] = u2;
        }
        if (p.ua1) {
            const size_t g = lev + (size_t)(j + p.joff - 1) * p.imax + i;
            const double uu0 = p.uu[g], vv0 = p.vv[g], pt0 = p.pt[g];
            const double* trefk = p.tref + (size_t)(k - 1) * p.jd;
            const double* urefk = p.uref + (size_t)(k - 1) * p.jd;
            const double tr = trefk[j - p.uoff - 1];
            p.ua1[o] = ur * (a1 + a2);
            double e1 = -0.5 * ((uu0 - ur) * (uu0 - ur));
            e1 = e1 + 0.5 * (vv0 * vv0);
            double ep11 = 0.5 * ((pt0 - tr) * (pt0 - tr));
            ep11 = ep11 * p.rrh * p.ep11a[k - 1];
            ep11 = ep11 * 2. * p.dz / p.tgdiff[k - 1];
            e1 = e1 - ep11;
            const double cosp = p.cosphi[j + 1], cos0 = cphi0, cosm = p.cosphi[j - 1], sin0 = p.sinphi[j];
            p.ep1[o] = e1 * cos0;
            // uref(j-uoff+1) / uref(j-uoff-1): flat indexing reproduces the reference's out-of-bounds
            // read of uref(0,k) == uref(jd,k-1) at the first row (no bounds checking in the F2PY build)
            p.ep2[o] = (p.uu[g + p.imax] - urefk[j - p.uoff]) * cosp * cosp * p.vv[g + p.imax];
            p.ep3[o] = (p.uu[g - p.imax] - urefk[j - p.uoff - 2]) * cosm * cosm * p.vv[g - p.imax];
            if (k == 2) {
                const size_t plane = (size_t)p.jmax * p.imax;
                const size_t g1 = (size_t)(j + p.joff - 1) * p.imax + i;  // level 1
                const double ep41 = 2. * p.om * sin0 * cos0 * p.dz / p.prefac;
                const double ep42 = p.ep42fac * p.vv[g1 + plane] * (p.pt[g1 + plane] - p.tref[p.jd + j - p.uoff - 1]) / p.tg31;
                double ep43 = p.vv[g1] * (p.pt[g1] - p.tref[j - p.uoff - 1]);
                ep43 = 0.5 * ep43 / p.tg21;
                p.ep4[(size_t)(j - 1) * p.imax + i] = ep41 * (ep42 + ep43);
            }
        }
    }
    if (BRUTE && p.mask_count) {
        c1 = __reduce_add_sync(0xffffffffu, (unsigned)c1);
        c2 = __reduce_add_sync(0xffffffffu, (unsigned)c2);
        if ((threadIdx.x & 31) == 0 && (c1 | c2)) {
            atomicAdd(&p.mask_count[0], c2);
            atomicAdd(&p.mask_count[1], c1);
        }
    }
}

// boundary row jb (compute_flux_dirinv.f90:173-178); jb = 0 aliases (i, nd, k-1) — SURVEY Q13
__global__ void flux_boundary_row_kernel(FluxArgs p, double cosp, double cos0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y + 2;
    if (i >= p.imax) return;
    const size_t lev = (size_t)(k - 1) * p.jmax * p.imax;
    const long long flat = (long long)i + (long long)p.imax * ((long long)(p.jb - 1) + (long long)p.nd * (k - 1));
    const size_t r1 = lev + (size_t)(p.nd + p.jb) * p.imax + i;      // row nd+jb+1 (1-based)
    const size_t r0 = lev + (size_t)(p.nd + p.jb - 1) * p.imax + i;  // row nd+jb
    p.ep2[flat] = (p.uu[r1] - p.uref[(size_t)(k - 1) * p.jd + 1]) * cosp * cosp * p.vv[r1];
    p.ep3[flat] = (p.uu[r0] - p.uref[(size_t)(k - 1) * p.jd + 0]) * cos0 * cos0 * p.vv[r0];
}

// density-weighted column sums of compute_lwa_only_nhn22.f90:93-104 (sequential in k)
__global__ void lwa_only_column_kernel(int imax, int jmax, int kmax, int nd, int jstart, int jend, int joff,
                                       const double* __restrict__ astar1, const double* __restrict__ astar2,
                                       const double* __restrict__ uu, const double* __restrict__ wk /* exp(-zk/h) */,
                                       double dc, double* __restrict__ astarbaro, double* __restrict__ ubaro) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y + 1;  // 1-based
    if (i >= imax) return;
    double sa = 0., su = 0.;
    for (int k = 2; k <= kmax - 1; ++k) {
        const size_t o = ((size_t)(k - 1) * nd + (j - 1)) * imax + i;
        sa = sa + (astar1[o] + astar2[o]) * wk[k - 1] * dc;
        if (j >= jstart && j <= jend)
            su = su + uu[((size_t)(k - 1) * jmax + (j + joff - 1)) * imax + i] * wk[k - 1] * dc;
    }
    astarbaro[(size_t)(j - 1) * imax + i] = sa;
    ubaro[(size_t)(j - 1) * imax + i] = su;
}

static int fill_common(FluxArgs& p, int imax, int jmax, int kmax, int nd, int jb, int jd, int is_nhem, double a,
                       double om, double dz, double h, double rr, double cp, double prefac,
                       std::vector<double>& tab, const double* tg) {
    const double pi = host_pi();
    const double dp = pi / (double)(jmax - 1);
    const double rkappa = rr / cp;
    p.imax = imax; p.jmax = jmax; p.kmax = kmax; p.nd = nd; p.jb = jb; p.jd = jd; p.is_nhem = is_nhem;
    if (is_nhem) { p.jstart = jb + 1; p.jend = nd - 1; p.joff = nd - 1; p.uoff = jb; }
    else { p.jstart = 2; p.jend = nd - jb; p.joff = 0; p.uoff = 0; }
    p.a = a; p.om = om; p.dz = dz; p.h = h; p.rr = rr; p.prefac = prefac; p.ab = a * dp;
    const double hemoff = is_nhem ? 0. : 0.5 * pi;
    const int nt = nd + 2;
    tab.assign((size_t)3 * nt + 2 * kmax, 0.0);
    for (int jj = 0; jj <= nd + 1; ++jj) {
        const double phi1 = dp * (double)(jj - 1) - hemoff;
        tab[jj] = std::cos(phi1);
        tab[nt + jj] = std::sin(phi1);
        tab[2 * nt + jj] = a * dp * std::cos(phi1);
    }
    for (int k = 2; k <= kmax - 1; ++k) {
        const double zz = dz * (double)(k - 1);
        tab[3 * nt + k - 1] = std::exp(-rkappa * zz / h);
        if (tg) tab[3 * nt + kmax + k - 1] = tg[k] - tg[k - 2];
    }
    p.ep42fac = std::exp(-dz / h);
    p.rrh = rr / h;
    if (tg) { p.tg31 = tg[2] - tg[0]; p.tg21 = tg[1] - tg[0]; } else { p.tg31 = p.tg21 = 1.; }
    return nt;
}

static void bind_tables(FluxArgs& p, const double* d, int nt, int kmax) {
    p.cosphi = d; p.sinphi = d + nt; p.aaw = d + 2 * nt; p.ep11a = d + 3 * nt; p.tgdiff = d + 3 * nt + kmax;
}

}  // namespace falwa

using namespace falwa;

extern "C" int falwa_b200_compute_flux_dirinv_nshem(
    const double* pv, const double* uu, const double* vv, const double* pt, const double* ncforce,
    const double* tn0_host, const double* qref, const double* uref, const double* tref, int imax, int jmax, int kmax,
    int nd, int jb, int jd, int is_nhem, double a, double om, double dz, double h, double rr, double cp, double prefac,
    double* astar1, double* astar2, double* ncforce3d, double* ua1, double* ua2, double* ep1, double* ep2, double* ep3,
    double* ep4, int64_t* mask_count, int method, void* stream) {
    if (!pv || !uu || !vv || !pt || !ncforce || !tn0_host || !qref || !uref || !tref || !astar1 || !astar2 ||
        !ncforce3d || !ua1 || !ua2 || !ep1 || !ep2 || !ep3 || !ep4)
        return FALWA_ERR_NULL;
    if (imax < 1 || kmax < 4 || nd < 4 || jmax != 2 * nd - 1 || jb < 0 || jd < 3 || jd > nd) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FluxArgs p;
    std::vector<double> tab;
    const int nt = fill_common(p, imax, jmax, kmax, nd, jb, jd, is_nhem, a, om, dz, h, rr, cp, prefac, tab, tn0_host);
    DeviceTable T;
    int rc = T.upload(tab, st);
    if (rc) return rc;
    bind_tables(p, T.d, nt, kmax);
    p.pv = pv; p.uu = uu; p.vv = vv; p.pt = pt; p.ncforce = ncforce; p.qref = qref; p.uref = uref; p.tref = tref;
    p.astar1 = astar1; p.astar2 = astar2; p.ncforce3d = ncforce3d; p.ua1 = ua1; p.ua2 = ua2; p.ep1 = ep1; p.ep2 = ep2;
    p.ep3 = ep3; p.ep4 = ep4; p.mask_count = (unsigned long long*)mask_count;
    const size_t n3 = sizeof(double) * (size_t)imax * nd * kmax;
    double* outs[8] = {astar1, astar2, ncforce3d, ua1, ua2, ep1, ep2, ep3};
    for (double* o : outs) FALWA_CUDA_TRY(cudaMemsetAsync(o, 0, n3, st));
    FALWA_CUDA_TRY(cudaMemsetAsync(ep4, 0, sizeof(double) * (size_t)imax * nd, st));
    if (mask_count) FALWA_CUDA_TRY(cudaMemsetAsync(mask_count, 0, 2 * sizeof(int64_t), st));
    const int nrows = p.jend - p.jstart + 1;
    dim3 block(32, 8);
    dim3 grid(ceil_div(imax, 32), ceil_div(nrows, 8), kmax - 2);
    // method: 0 = automatic (windowed sums, else interval scan, else the double loop), 1 = double loop,
    //         2 = interval scan, 3 = windowed sums, 4 = prefix sums over threshold intervals
    const bool use_window = (method == 0 || method == 3 || method == 4) && is_nhem && win_supported(imax, nd, pv, uu, ncforce);
    const bool use_scan = !use_window && (method != 1) && is_nhem && scan_supported(nd);
    ScanTables tabs;
    if (use_window) {  // astar1, astar2, ua2, ncforce3d by exact windowed sums (lwa_window.cu), the pointwise terms afterwards
        WinArgs w{};
        w.imax = imax; w.nd = nd; w.kmax = kmax; w.jstart = p.jstart; w.jend = p.jend; w.nhemi = 1;
        w.k_first = 2; w.k_count = kmax - 2;
        w.fr[0].in_row0 = p.joff; w.fr[0].out_row0 = 0; w.fr[0].rev = 0; w.fr[0].skip_first = 0; w.fr[0].sg = 1.0;
        w.qref = qref; w.q_batch = 0; w.q_kstride = nd; w.q_row0[0] = 0; w.q_rstep[0] = 1; w.q_sg[0] = 1.0;
        w.uref = uref; w.u_batch = 0; w.u_kstride = jd; w.u_row0[0] = -p.uoff; w.u_rstep[0] = 1;
        w.cosphi = p.cosphi; w.ab = p.ab;
        w.a1 = astar1; w.a2 = astar2; w.u2 = ua2; w.ncf = ncforce3d;
        w.out_batch = 0; w.out_plane = (size_t)nd * imax; w.out_pitch = imax;
        w.mask_count = p.mask_count;
        if ((rc = lwa_window_launch(w, 0, pv, uu, ncforce, jmax, 1, 1, st))) return rc;
        if (method == 4) {   // prefix sums (lwa_prefix.cu) for astar1, astar2, ua2 and the pair counts; it does not carry
                             // ncforce, so ncforce3d stays the window kernel's
            if (mask_count) FALWA_CUDA_TRY(cudaMemsetAsync(mask_count, 0, 2 * sizeof(int64_t), st));
            w.ncf = nullptr;
            if ((rc = lwa_prefix_launch(w, 0, pv, uu, jmax, 1, 1, nullptr, 0, st))) return rc;
        }
        { FALWA_KERNEL("flux_pointwise_kernel", st); flux_bruteforce_kernel<false><<<grid, block, 0, st>>>(p); }
    } else if (use_scan) {  // astar1, astar2, ua2, ncforce3d by the interval scan, the pointwise terms afterwards
        if ((rc = tabs.alloc(1, nd, kmax, st))) return rc;
        ThrArgs t{};
        ScanArgs s{};
        t.nd = nd; t.kmax = kmax; t.jstart = p.jstart; t.jend = p.jend; t.nhemi = 1;
        t.qref = qref; t.q_batch = 0; t.q_kstride = nd; t.q_row0[0] = 0; t.q_rstep[0] = 1; t.q_sg[0] = 1.0;
        t.uref = uref; t.u_batch = 0; t.u_kstride = jd; t.u_row0[0] = -p.uoff; t.u_rstep[0] = 1;
        s.imax = imax; s.nd = nd; s.kmax = kmax; s.jstart = p.jstart; s.jend = p.jend; s.nhemi = 1;
        s.pv = pv; s.uu = uu; s.nc = ncforce; s.in_batch = 0; s.in_plane = (size_t)jmax * imax;
        s.a1 = astar1; s.a2 = astar2; s.u2 = ua2; s.ncf = ncforce3d;
        s.out_batch = 0; s.out_plane = (size_t)nd * imax; s.out_pitch = imax;
        s.fr[0].in_row0 = p.joff; s.fr[0].in_rstep = 1; s.fr[0].out_row0 = 0; s.fr[0].out_rstep = 1;
        s.fr[0].skip_first_row = 0; s.fr[0].sg = 1.0;
        s.cosphi = p.cosphi; s.ab = p.ab; s.mask_count = p.mask_count;
        tabs.bind(t, s, 1, nd, kmax);
        if ((rc = lwa_scan_launch(t, s, 1, st))) return rc;
        { FALWA_KERNEL("flux_pointwise_kernel", st); flux_bruteforce_kernel<false><<<grid, block, 0, st>>>(p); }
    } else {
        FALWA_KERNEL("flux_bruteforce_kernel", st);
        flux_bruteforce_kernel<true><<<grid, block, 0, st>>>(p);
    }
    FALWA_LAUNCH_CHECK();
    {  // boundary row jb: executed for either value of is_nhem, like the reference
        const double pi = host_pi(), dp = pi / (double)(jmax - 1);
        { FALWA_KERNEL("flux_boundary_row_kernel", st); flux_boundary_row_kernel<<<dim3(ceil_div(imax, 128), kmax - 2), 128, 0, st>>>(p, std::cos(dp * (double)jb),
                                                                                     std::cos(dp * (double)(jb - 1))); }
        FALWA_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int falwa_b200_compute_lwa_only_nhn22(const double* pv, const double* uu, const double* qref, int imax,
                                                 int jmax, int kmax, int nd, int jb, int is_nhem, double a, double om,
                                                 double dz, double h, double rr, double cp, double prefac,
                                                 double* astarbaro, double* ubaro, double* astar1, double* astar2,
                                                 int method, void* stream) {
    if (!pv || !uu || !qref || !astarbaro || !ubaro || !astar1 || !astar2) return FALWA_ERR_NULL;
    if (imax < 1 || kmax < 4 || nd < 4 || jmax != 2 * nd - 1 || jb < 0) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FluxArgs p;
    std::vector<double> tab;
    const int nt = fill_common(p, imax, jmax, kmax, nd, jb, nd - jb, is_nhem, a, om, dz, h, rr, cp, prefac, tab, nullptr);
    // column weights exp(-zk/h)
    const size_t o_w = tab.size();
    tab.resize(tab.size() + kmax, 0.0);
    for (int k = 1; k <= kmax; ++k) tab[o_w + k - 1] = std::exp(-(dz * (double)(k - 1)) / h);
    DeviceTable T;
    int rc = T.upload(tab, st);
    if (rc) return rc;
    bind_tables(p, T.d, nt, kmax);
    // the LWA-only routine never touches uref: point it at qref's buffer with uoff chosen in-bounds
    p.pv = pv; p.uu = uu; p.vv = nullptr; p.pt = nullptr; p.ncforce = nullptr; p.qref = qref;
    p.uref = qref; p.jd = nd; p.uoff = 0; p.tref = nullptr;
    p.astar1 = astar1; p.astar2 = astar2; p.ncforce3d = nullptr; p.ua1 = nullptr; p.ua2 = nullptr; p.ep1 = nullptr;
    p.ep2 = nullptr; p.ep3 = nullptr; p.ep4 = nullptr; p.mask_count = nullptr;
    const size_t n3 = sizeof(double) * (size_t)imax * nd * kmax;
    FALWA_CUDA_TRY(cudaMemsetAsync(astar1, 0, n3, st));
    FALWA_CUDA_TRY(cudaMemsetAsync(astar2, 0, n3, st));
    const int nrows = p.jend - p.jstart + 1;
    dim3 block(32, 8);
    dim3 grid(ceil_div(imax, 32), ceil_div(nrows, 8), kmax - 2);
    ScanTables tabs;
    if ((method == 0 || method == 3 || method == 4) && is_nhem && win_supported(imax, nd, pv, uu, nullptr)) {
        WinArgs w{};
        w.imax = imax; w.nd = nd; w.kmax = kmax; w.jstart = p.jstart; w.jend = p.jend; w.nhemi = 1;
        w.k_first = 2; w.k_count = kmax - 2;
        w.fr[0].in_row0 = p.joff; w.fr[0].out_row0 = 0; w.fr[0].rev = 0; w.fr[0].skip_first = 0; w.fr[0].sg = 1.0;
        w.qref = qref; w.q_batch = 0; w.q_kstride = nd; w.q_row0[0] = 0; w.q_rstep[0] = 1; w.q_sg[0] = 1.0;
        w.uref = nullptr;
        w.cosphi = p.cosphi; w.ab = p.ab;
        w.a1 = astar1; w.a2 = astar2; w.u2 = nullptr; w.ncf = nullptr;
        w.out_batch = 0; w.out_plane = (size_t)nd * imax; w.out_pitch = imax;
        w.mask_count = nullptr;
        if (method == 4) rc = lwa_prefix_launch(w, 0, pv, uu, jmax, 1, 1, nullptr, 0, st);
        else rc = lwa_window_launch(w, 0, pv, uu, nullptr, jmax, 1, 1, st);
        if (rc) return rc;
    } else if ((method != 1) && is_nhem && scan_supported(nd)) {
        if ((rc = tabs.alloc(1, nd, kmax, st))) return rc;
        ThrArgs t{};
        ScanArgs s{};
        t.nd = nd; t.kmax = kmax; t.jstart = p.jstart; t.jend = p.jend; t.nhemi = 1;
        t.qref = qref; t.q_batch = 0; t.q_kstride = nd; t.q_row0[0] = 0; t.q_rstep[0] = 1; t.q_sg[0] = 1.0;
        t.uref = nullptr;
        s.imax = imax; s.nd = nd; s.kmax = kmax; s.jstart = p.jstart; s.jend = p.jend; s.nhemi = 1;
        s.pv = pv; s.uu = uu; s.nc = nullptr; s.in_batch = 0; s.in_plane = (size_t)jmax * imax;
        s.a1 = astar1; s.a2 = astar2; s.u2 = nullptr; s.ncf = nullptr;
        s.out_batch = 0; s.out_plane = (size_t)nd * imax; s.out_pitch = imax;
        s.fr[0].in_row0 = p.joff; s.fr[0].in_rstep = 1; s.fr[0].out_row0 = 0; s.fr[0].out_rstep = 1;
        s.fr[0].skip_first_row = 0; s.fr[0].sg = 1.0;
        s.cosphi = p.cosphi; s.ab = p.ab; s.mask_count = nullptr;
        tabs.bind(t, s, 1, nd, kmax);
        if ((rc = lwa_scan_launch(t, s, 1, st))) return rc;
    } else {
        FALWA_KERNEL("flux_bruteforce_kernel", st);
        flux_bruteforce_kernel<true><<<grid, block, 0, st>>>(p);
    }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("lwa_only_column_kernel", st); lwa_only_column_kernel<<<dim3(ceil_div(imax, 128), nd), 128, 0, st>>>(imax, jmax, kmax, nd, p.jstart, p.jend, p.joff,
                                                                        astar1, astar2, uu, T.d + o_w, dz / prefac,
                                                                        astarbaro, ubaro); }
    FALWA_LAUNCH_CHECK();
    return 0;
}
