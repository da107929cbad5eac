// Section B of the C ABI: the fused, batched, device-resident QGField pipeline
// (interpolate_fields -> compute_reference_states -> compute_lwa_and_barotropic_fluxes) behind
// hn2016_falwa_b200.QGFieldNH18 / QGFieldNHN22 / QGDataset.
//
// Reference call stacks restated: SURVEY.md §3.1-3.2 (oopinterface.py:467-535, :1604-1674,
// :1774-1888, :716-982, data_storage.py:32-62,170-179, utilities.py:189-233).
//
// Frames: every 3-D field is kept once, in the global frame [k][lat south->north][lon].  The
// reference processes the southern hemisphere by flipping latitude and negating pv, v and qref
// (SURVEY Q12); here that is index arithmetic: hemispheric row jj (1 = equator .. nd = pole) lives
// in global row  g(jj) = nd-2+jj (north)  or  nd-jj (south)  (0-based), with sign factors.
#include "common.cuh"

namespace falwa {

enum Out2D {
    O_ASTAR1 = 0, O_ASTAR2, O_LWA, O_UA1, O_UA2, O_EP1, O_EP2, O_EP3, O_EP4, O_NCF, O_UBARO,
    O_CONV, O_DIVEMF, O_FLAMBDA, O_FPHI, O_COUNT
};

}  // namespace falwa

struct falwa_plan {
    int kind, nlon, nlat, nlev, kmax, jb, nd, jd, npart, maxit;
    double a, omega, dz, h, rr, cp, tol, rjac, prefactor, dphi, dlambda;
    double hist_area_scale, hist_cbound;  // fixed-point scales of the area histograms (area_hist3_kernel)
    std::vector<double> height, sinlat;   // host copies (tables built per call)
    double* d_tab = nullptr;
    int* d_lo = nullptr;
    // offsets into d_tab
    size_t o_fac, o_clat, o_whi, o_wlo, o_qg, o_alat, o_da, o_cosmid, o_norm, o_vw, o_sinlat;
    size_t o_ajk, o_bjk, o_fact, o_cosj, o_cor, o_cosw, o_eref, o_ez, o_uzc1, o_uzc2;       // NH18 (nd rows)
    size_t o_najk, o_nbjk, o_nfact, o_topcoef, o_topden, o_cosrow, o_sinrow, o_sinnd;       // NHN22
    size_t o_cosphi, o_sinphi, o_aaw, o_ep11a;                                             // flux (nd+2)
    size_t ntab;
};

namespace falwa {

// ---------------------------------------------------------------------------------------------
// Area-analysis histograms in one pass over pv and avort.
// hist layout: [b][4][k][2][nb]   slot 0: +pv (northern frame), slot 2: avort,
//                                  slot 1 / 3: the southern / northern contributions of "edge" points only.
// The southern analysis histograms -pv from its own maximum: ind_S = 1 + int((q - pmin)/dq).  Because
// (pmax - q) + (q - pmin) = (nb-1) dq, ind_S = nb - ind_N unless the quotient lies within rounding distance of an
// integer (area_bin's edge test, |frac| < 1e-9), so the southern histogram is the MIRROR of the northern one and is
// not accumulated at all: refstate_level_kernel forms  S[m] = (N - N_edge)[nb - m] (sign-flipped circulation)
// + S_edge[m],  where the edge points (exactly classified with the reference's divisions) go to slots 1 and 3.
// NEED_QN: the circulation sums of the pv histogram are only needed by the NH18 boundary condition.
//
// Reproducible sums: the bins are accumulated as 64-bit fixed-point integers — every term is rounded once
// to a multiple of 2^-e and integer addition is associative, so the histogram (hence Qref) is bit-identical from run
// to run whatever order the atomics land in — and integer shared-memory / L2 atomics are single instructions, where
// an fp64 shared-memory atomicAdd is a compare-and-swap loop.  Areas use one scale for the whole grid (the sphere's
// area must fit 2^62); circulations use a power-of-two scale per level derived from that level's extrema.  The
// quantisation (relative 2^-40 or better per term against the largest term) is far below the rtol the sums are
// compared with, and is the same for every run.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline double hist_pow2_scale(double bound) {   // 2^e with bound * 2^e in [2^60, 2^61)
    if (!(bound > 0.0) || !(bound < 1e300)) return 1.0;
    int ex;
    frexp(bound, &ex);          // bound = m * 2^ex, m in [0.5, 1)
    return ldexp(1.0, 61 - ex);
}
__device__ __forceinline__ double hist_circ_scale(double vmax, double vmin, double cbound) {
    return hist_pow2_scale(fmax(fabs(vmax), fabs(vmin)) * cbound);
}

// 64-bit fixed-point add into shared memory with native 32-bit atomics (a 64-bit shared-memory atomicAdd, integer or
// floating point, is a compare-and-swap loop): the low word is added first and its carry rides with the high word.
// The pair (lo, hi) is the little-endian image of the 64-bit sum modulo 2^64, whatever the order of the adds.
__device__ __forceinline__ void fx_add(unsigned long long* slot, long long v) {
    unsigned* w = reinterpret_cast<unsigned*>(slot);
    const unsigned lo = (unsigned)v, hi = (unsigned)((unsigned long long)v >> 32);
    const unsigned old = atomicAdd(w, lo);
    const unsigned carry = (old + lo) < old ? 1u : 0u;
    if (hi + carry) atomicAdd(w + 1, hi + carry);
}

// A lane's run of points in one bin: `cnt` points of a row with cell area dai (fixed point) and the run's circulation
// sum (accumulated in the lane's fixed order in fp64, rounded once to fixed point here).
template <bool NEED_QN>
__device__ __forceinline__ void hist_flush(unsigned long long* sh, int nb, int cur, int cnt, long long dai, double sc, double scale) {
    if (cur > 0) {
        fx_add(&sh[cur - 1], (long long)cnt * dai);
        if (NEED_QN) fx_add(&sh[nb + cur - 1], __double2ll_rn(sc * scale));
    }
}

// WITH_P / WITH_V: which of the two histograms (pv, absolute vorticity) this launch accumulates.  One field per launch
// halves the registers of the software-pipelined loads (one more CTA per SM) and the two launches overlap their tails.
template <bool NEED_QN, bool WITH_V, bool WITH_P = true>
__global__ void __launch_bounds__(256, (WITH_V && WITH_P) ? 2 : 3)
area_hist3_kernel(int nlon, int nlat, int kmax, int nb, const double* __restrict__ pv,
                  const double* __restrict__ avort, const unsigned long long* __restrict__ ext,
                  const double* __restrict__ da, int rows_per_block, unsigned long long* __restrict__ hist, int run,
                  double area_scale, double cbound) {
    extern __shared__ unsigned long long shi[];  // [4][2][nb]
    const int k = blockIdx.y + 1, b = blockIdx.z;
    pv += (size_t)b * nlon * nlat * kmax;
    if (WITH_V) avort += (size_t)b * nlon * nlat * kmax;
    ext += (size_t)b * 2 * kmax * 2;
    hist += (size_t)b * 4 * kmax * 2 * nb;
    for (int t = threadIdx.x; t < 4 * 2 * nb; t += blockDim.x) shi[t] = 0ull;
    __syncthreads();
    const double pmax = ordered_to_f64(ext[k * 2 + 0]), pmin = ordered_to_f64(ext[k * 2 + 1]);
    const double dqp = (pmax - pmin) / (double)(nb - 1);
    const double smax = -pmin;                      // southern frame: q' = -pv, max' = -min
    const double dqs = (smax - (-pmax)) / (double)(nb - 1);
    const double rdqp = 1. / dqp;
    const double scp = hist_circ_scale(pmax, pmin, cbound);
    double vmax = 0., dqv = 1., scv = 1.;
    if (WITH_V) {
        vmax = ordered_to_f64(ext[(kmax + k) * 2 + 0]);
        const double vmin = ordered_to_f64(ext[(kmax + k) * 2 + 1]);
        dqv = (vmax - vmin) / (double)(nb - 1);
        scv = hist_circ_scale(vmax, vmin, cbound);
    }
    const double rdqv = 1. / dqv;
    const int j0 = blockIdx.x * rows_per_block, j1 = min(nlat, j0 + rows_per_block);
    const int runs_per_row = (nlon + run - 1) / run;
    const int total = (j1 - j0) * runs_per_row;
    const size_t lev = (size_t)k * nlat * nlon;
    unsigned long long* hN = shi;
    unsigned long long* hSx = shi + 2 * nb;
    unsigned long long* hV = shi + 4 * nb;
    unsigned long long* hNx = shi + 6 * nb;
    const int nrows = j1 - j0;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        // consecutive lanes take the same longitude run of consecutive ROWS: PV differs from row to row, so the
        // lanes of a warp hit different bins and their shared-memory atomics rarely meet
        const int j = j0 + t % nrows;
        const int i0 = (t / nrows) * run, i1 = min(nlon, i0 + run);
        const double daj = da[j];
        const long long dai = __double2ll_rn(daj * area_scale);
        int cn = -1, cv = -1, an = 0, av = 0;
        double qn = 0., qv = 0.;
        auto point = [&](double q, double w) {
            bool edge = false;
            int ind = 0;
            if (WITH_P) {
                ind = area_bin(pmax, q, dqp, rdqp, nb, &edge);
                if (ind != cn) { hist_flush<NEED_QN>(hN, nb, cn, an, dai, qn, scp); cn = ind; an = 0; qn = 0.; }
                ++an;
                if (NEED_QN) qn += daj * q;
            }
            if (WITH_P && __builtin_expect(edge, 0)) {  // rare: classify the southern bin with the reference's own division
                const double qq = -q;
                const int is = max(1, min(nb, 1 + (int)area_bin_exact(smax - qq, dqs)));
                const long long tq = __double2ll_rn((daj * q) * scp);
                fx_add(&hSx[is - 1], dai); fx_add(&hSx[nb + is - 1], -tq);
                fx_add(&hNx[ind - 1], dai); fx_add(&hNx[nb + ind - 1], tq);
            }
            if (WITH_V) {
                ind = area_bin(vmax, w, dqv, rdqv, nb);
                if (ind != cv) { hist_flush<true>(hV, nb, cv, av, dai, qv, scv); cv = ind; av = 0; qv = 0.; }
                ++av; qv += daj * w;
            }
        };
        const size_t row = lev + (size_t)j * nlon;
        int i = i0;
        if (((row + i0) & 1) == 0) {
            // 16-byte vector loads, eight points (four double2 per field) in flight per lane: the lanes of a warp
            // read different rows, so every access is its own sector and memory-level parallelism has to come
            // from the thread itself
            // software-pipelined: the next eight points are requested before the current eight are binned
            double2 a[4], c[4], na[4], nc[4];
            auto fetch8 = [&](int at, double2* A, double2* C) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (WITH_P) A[v] = __ldg(reinterpret_cast<const double2*>(pv + row + at) + v);
                    else A[v] = make_double2(0., 0.);
                    if (WITH_V) C[v] = __ldg(reinterpret_cast<const double2*>(avort + row + at) + v);
                    else C[v] = make_double2(0., 0.);
                }
            };
            if (i + 8 <= i1) fetch8(i, a, c);
            for (; i + 8 <= i1; i += 8) {
                const bool more = (i + 16 <= i1);
                if (more) fetch8(i + 8, na, nc);
#pragma unroll
                for (int v = 0; v < 4; ++v) { point(a[v].x, c[v].x); point(a[v].y, c[v].y); }
                if (more) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) { a[v] = na[v]; c[v] = nc[v]; }
                }
            }
        }
        for (; i < i1; ++i) point(WITH_P ? pv[row + i] : 0.0, WITH_V ? avort[row + i] : 0.0);
        if (WITH_P) hist_flush<NEED_QN>(hN, nb, cn, an, dai, qn, scp);
        if (WITH_V) hist_flush<true>(hV, nb, cv, av, dai, qv, scv);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 4 * 2 * nb; t += blockDim.x) {
        const unsigned long long x = shi[t];
        if (x != 0ull) {
            const int hsel = t / (2 * nb), r = t % (2 * nb);
            atomicAdd(&hist[((size_t)hsel * kmax + k) * 2 * nb + r], x);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// merge hemispheric reference-state profiles into the global-latitude storage
// (data_storage.py:32-62: north written to the last rows, south written flipped afterwards)
//   qref_st[k][nlat] <- stored (normalised) qref ; qref_cu = ((qref_st*2)*omega)*sin(lat)
// ---------------------------------------------------------------------------------------------
__global__ void merge_refstates_kernel(int kmax, int nlat, int nd, int jd, int north_only, double qscale, double omega,
                                       const double* __restrict__ sinlat, const double* __restrict__ work, size_t per,
                                       size_t njk, size_t njd, double* __restrict__ qref_st,
                                       double* __restrict__ qref_cu, double* __restrict__ uref_st,
                                       double* __restrict__ ptref_st) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kmax * nlat) return;
    {   // blockIdx.y = snapshot; hemispheric work areas: qref at 0, uref at 6*njk, tref at 6*njk + njd
        const size_t b = blockIdx.y, o = b * (size_t)kmax * nlat;
        qref_st += o; qref_cu += o; uref_st += o; ptref_st += o;
    }
    const double* wN = work + ((size_t)blockIdx.y * 2 + 0) * per;
    const double* wS = work + ((size_t)blockIdx.y * 2 + 1) * per;
    const double *qN = wN, *qS = wS, *uN = wN + 6 * njk, *uS = wS + 6 * njk, *tN = uN + njd, *tS = uS + njd;
    const int k = t / nlat, j = t % nlat;  // global row
    double q, u = 0.0, th = 0.0;
    // south covers rows [0, nd), north rows [nlat-nd, nlat); the equator row takes the southern value
    if (north_only) {
        q = (j >= nlat - nd) ? qN[(size_t)k * nd + (j - (nlat - nd))] / qscale : 0.0;
        if (j >= nlat - jd) { u = uN[(size_t)k * jd + (j - (nlat - jd))]; th = tN[(size_t)k * jd + (j - (nlat - jd))]; }
    } else {
        if (j < nd) q = qS[(size_t)k * nd + (nd - 1 - j)] / qscale;
        else q = qN[(size_t)k * nd + (j - (nlat - nd))] / qscale;
        if (j < jd) { u = uS[(size_t)k * jd + (jd - 1 - j)]; th = tS[(size_t)k * jd + (jd - 1 - j)]; }
        else if (j >= nlat - jd) { u = uN[(size_t)k * jd + (j - (nlat - jd))]; th = tN[(size_t)k * jd + (j - (nlat - jd))]; }
    }
    qref_st[t] = q;
    qref_cu[t] = q * 2 * omega * sinlat[j];
    uref_st[t] = u;
    ptref_st[t] = th;
}

// ---------------------------------------------------------------------------------------------
// LWA + flux integrands + density-weighted vertical sums, one hemisphere, brute-force LWA.
// One thread per (lon, hemispheric row); levels are walked sequentially so the vertical sums are
// accumulated in the order of numpy's reduction (oopinterface.py:405-420).
// ---------------------------------------------------------------------------------------------
struct FusedFluxArgs {
    int imax, nlat, kmax, nd, jb, jd, nhemi, has_ncforce;
    const double *sa1, *sa2, *su2, *snc;       // BRUTE = false: LWA pieces from the interval-scan kernel, [b][k][lat][lon]
    const double *clat;                        // |cos(lat)| of the global rows
    double a, om, dz, prefac, ab, rrh, ep42fac, dc, cosjb, cosjbm1;
    const double *pv, *uu, *vv, *pt, *ncforce;
    const double *qref_cu, *uref_st, *tref_st;  // [b][k][nlat]
    const double *prof;                         // [b][4][kmax]
    const double *cosphi, *sinphi, *aaw, *ep11a, *vw;
    const double *f11;                          // BRUTE = false: [b][2][kmax] = (R/H) e^{-kappa z/H} 2 dz / (tg(k+1)-tg(k-1))
    double *lwa, *out2d;
    double* l3d;   // optional layer-wise outputs [b][L3_COUNT][kmax][nlat][nlon] (compute_layerwise_lwa_fluxes)
};

// slots of the layer-wise output (oopinterface.py:1063-1156: astar1, astar2, ncforce, ua1, ua2, ep1, ep2, ep3; then
// the stretching term written by zderiv_kernel)
enum Layer3D { L_ASTAR1 = 0, L_ASTAR2, L_NCF, L_UA1, L_UA2, L_EP1, L_EP2, L_EP3, L_STRETCH, L_LWA, L3_COUNT };

// BRUTE = true: the reference's O(nd^2) double loop inline (method 1).  BRUTE = false: a1 / a2 / ua2 / ncforce3d
// come from lwa_scan_kernel (flux_scan.cu) and this kernel is a pure column walk ("flux_column").
// blockIdx.z = snapshot * nhemi + hemisphere (0 = north, 1 = south).  When both hemispheres are computed the
// southern one owns the equator row (the reference writes north first, south second: data_storage.py:42-62).
template <bool BRUTE>
__global__ void __launch_bounds__(256, BRUTE ? 2 : 3)
fused_flux_kernel(FusedFluxArgs p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;  // hemispheric row 1..nd
    const int b = blockIdx.z / p.nhemi;
    const bool south = (blockIdx.z % p.nhemi) == 1;
    if (i >= p.imax || j > p.nd) return;
    if (j == 1 && p.nhemi == 2 && !south) return;
    const int nd = p.nd, nlat = p.nlat, imax = p.imax, kmax = p.kmax, jb = p.jb;
    const size_t plane = (size_t)nlat * imax;
    const size_t f3 = (size_t)b * plane * kmax;
    const double *pv = p.pv + f3, *uu = p.uu + f3, *vv = p.vv + f3, *pt = p.pt + f3;
    const double* nc = p.has_ncforce ? p.ncforce + f3 : nullptr;
    const double* qref = p.qref_cu + (size_t)b * kmax * nlat;
    const double* uref = p.uref_st + (size_t)b * kmax * nlat;
    const double* tref = p.tref_st + (size_t)b * kmax * nlat;
    const double* tg = p.prof + (size_t)b * 4 * kmax + (south ? 0 : kmax);  // tn0 north / ts0 south
    double* lwa = p.lwa + f3;
    double* out = p.out2d + (size_t)b * O_COUNT * plane;
    const double sg = south ? -1.0 : 1.0;  // sign of pv, v and qref in the hemispheric frame
    auto grow = [&](int jj) -> int { return south ? (nd - jj) : (nd - 2 + jj); };  // 0-based global row
    const int g = grow(j);
    const int jstart = jb + 1, jend = nd - 1;
    const bool in_range = (j >= jstart && j <= jend);
    const bool brow = (jb >= 1) ? (j == jb) : (j == nd);  // boundary row jb, or its alias (i,nd,k-1) when jb = 0
    const double cphi0 = p.cosphi[j];
    double s_a1 = 0, s_a2 = 0, s_ua1 = 0, s_ua2 = 0, s_ep1 = 0, s_ep2 = 0, s_ep3 = 0, s_nc = 0, s_u = 0, ep4 = 0;
    double s_fphi = 0;
    const double cl = p.clat[g];
    const size_t o2 = (size_t)g * imax + i;
    lwa[o2] = 0.0;
    lwa[(size_t)(kmax - 1) * plane + o2] = 0.0;
    for (int k = 2; k <= kmax; ++k) {  // 1-based level
        const size_t lev = (size_t)(k - 1) * plane;
        const double w = p.vw[k - 1];
        const double wd = w * p.dc;
        // numpy evaluates (x * w) * dc; the scan path folds the two weights (one rounding less per term)
        auto acc = [&](double sum, double x) -> double { return BRUTE ? sum + x * w * p.dc : sum + x * wd; };
        s_u = acc(s_u, uu[lev + o2]);
        {   // flux_vector_phi_baro = vertical average of -((u - uref) * v) * clat  (oopinterface.py:964-982)
            const double x = -(((uu[lev + o2] - uref[(size_t)(k - 1) * nlat + g]) * vv[lev + o2]) * cl);
            s_fphi = acc(s_fphi, x);
        }
        if (k == kmax) break;
        double a1 = 0, a2 = 0, ncf = 0, u2 = 0, e1 = 0, e2 = 0, e3 = 0, x_ua1 = 0;
        if (in_range) {
            const double qr = sg * qref[(size_t)(k - 1) * nlat + g];
            const double ur = uref[(size_t)(k - 1) * nlat + g];
            if (!BRUTE) {
                a1 = p.sa1[f3 + lev + o2];
                a2 = p.sa2[f3 + lev + o2];
                u2 = p.su2[f3 + lev + o2];
                if (nc) ncf = p.snc[f3 + lev + o2];
            }
            for (int jj = 1; BRUTE && jj < j; ++jj) {
                const size_t gg = lev + (size_t)grow(jj) * imax + i;
                const double qe = sg * pv[gg] - qr;
                if (qe > 0.) {
                    const double aa = p.aaw[jj];
                    const double ue = uu[gg] * cphi0 - ur * p.cosphi[jj];
                    a1 = a1 + qe * aa;
                    if (nc) ncf = ncf + nc[gg] * aa;
                    u2 = u2 + qe * ue * p.ab;
                }
            }
            for (int jj = j; BRUTE && jj <= nd; ++jj) {
                const size_t gg = lev + (size_t)grow(jj) * imax + i;
                const double qe = sg * pv[gg] - qr;
                if (qe <= 0.) {
                    const double aa = p.aaw[jj];
                    const double ue = uu[gg] * cphi0 - ur * p.cosphi[jj];
                    a2 = a2 - qe * aa;
                    if (nc) ncf = ncf - nc[gg] * aa;
                    u2 = u2 - qe * ue * p.ab;
                }
            }
            const size_t gc = lev + o2;
            const double uu0 = uu[gc], vv0 = sg * vv[gc], pt0 = pt[gc];
            const double tr = tref[(size_t)(k - 1) * nlat + g];
            x_ua1 = ur * (a1 + a2);
            e1 = -0.5 * ((uu0 - ur) * (uu0 - ur));
            e1 = e1 + 0.5 * (vv0 * vv0);
            double ep11 = 0.5 * ((pt0 - tr) * (pt0 - tr));
            if (BRUTE) {
                ep11 = ep11 * p.rrh * p.ep11a[k - 1];
                ep11 = ep11 * 2. * p.dz / (tg[k] - tg[k - 2]);
            } else {
                ep11 = ep11 * p.f11[((size_t)b * 2 + (south ? 0 : 1)) * kmax + (k - 1)];
            }
            e1 = e1 - ep11;
            const double cosp = p.cosphi[j + 1], cosm = p.cosphi[j - 1], sin0 = p.sinphi[j];
            e1 = e1 * cphi0;
            const size_t gp = lev + (size_t)grow(j + 1) * imax + i, gm = lev + (size_t)grow(j - 1) * imax + i;
            const double urp = uref[(size_t)(k - 1) * nlat + grow(j + 1)];
            // uref(j-jb-1): for the first row this is the reference's out-of-bounds uref(0,k) == uref(jd,k-1)
            const double urm = (j - jb - 1 >= 1) ? uref[(size_t)(k - 1) * nlat + grow(j - 1)]
                                                 : uref[(size_t)(k - 2) * nlat + grow(nd)];
            e2 = (uu[gp] - urp) * cosp * cosp * (sg * vv[gp]);
            e3 = (uu[gm] - urm) * cosm * cosm * (sg * vv[gm]);
            if (k == 2) {
                const double ep41 = 2. * p.om * sin0 * cphi0 * p.dz / p.prefac;
                const double ep42 = p.ep42fac * (sg * vv[plane + o2]) * (pt[plane + o2] - tref[nlat + g]) / (tg[2] - tg[0]);
                double ep43 = (sg * vv[o2]) * (pt[o2] - tref[g]);
                ep43 = 0.5 * ep43 / (tg[1] - tg[0]);
                ep4 = ep41 * (ep42 + ep43);
            }
        }
        lwa[lev + o2] = fabs(a1 + a2);
        double wk = w;
        if (brow) {  // compute_flux_dirinv.f90:173-178
            const size_t r1 = lev + (size_t)grow(jb + 2) * imax + i, r0 = lev + (size_t)grow(jb + 1) * imax + i;
            e2 = (uu[r1] - uref[(size_t)(k - 1) * nlat + grow(jb + 2)]) * p.cosjb * p.cosjb * (sg * vv[r1]);
            e3 = (uu[r0] - uref[(size_t)(k - 1) * nlat + grow(jb + 1)]) * p.cosjbm1 * p.cosjbm1 * (sg * vv[r0]);
            if (jb == 0) wk = (k >= 3) ? p.vw[k - 2] : 0.0;  // aliased into level k-1; level 1 is not averaged
        }
        if (p.l3d) {   // layer-wise fields in the global frame, southern sign rules of oopinterface.py:1117-1148
            double* L = p.l3d + (size_t)b * L3_COUNT * plane * kmax + lev + o2;
            const size_t fs = plane * kmax;
            L[L_ASTAR1 * fs] = a1;
            L[L_ASTAR2 * fs] = a2;
            L[L_LWA * fs] = a1 + a2;    // the layer-wise store keeps the sum, not its magnitude (oopinterface.py:1099)
            L[L_NCF * fs] = south ? -ncf : ncf;
            L[L_UA1 * fs] = x_ua1;
            L[L_UA2 * fs] = u2;
            L[L_EP1 * fs] = e1;
            // the boundary-row values of a jb = 0 run land on (i, nd, k-1): one level down
            double* Le = (brow && jb == 0) ? L - plane : L;
            Le[L_EP2 * fs] = south ? -e3 : e2;
            Le[L_EP3 * fs] = south ? -e2 : e3;
        }
        s_a1 = acc(s_a1, a1);
        s_a2 = acc(s_a2, a2);
        s_ua1 = acc(s_ua1, x_ua1);
        s_ua2 = acc(s_ua2, u2);
        s_ep1 = acc(s_ep1, e1);
        s_ep2 = BRUTE ? s_ep2 + e2 * wk * p.dc : s_ep2 + e2 * (wk * p.dc);
        s_ep3 = BRUTE ? s_ep3 + e3 * wk * p.dc : s_ep3 + e3 * (wk * p.dc);
        s_nc = acc(s_nc, ncf);
    }
    out[(size_t)O_ASTAR1 * plane + o2] = s_a1;
    out[(size_t)O_ASTAR2 * plane + o2] = s_a2;
    out[(size_t)O_LWA * plane + o2] = s_a1 + s_a2;
    out[(size_t)O_UA1 * plane + o2] = s_ua1;
    out[(size_t)O_UA2 * plane + o2] = s_ua2;
    out[(size_t)O_EP1 * plane + o2] = s_ep1;
    // southern hemisphere: ep2/ep3 swapped and negated, ncforce negated (oopinterface.py:790-792)
    out[(size_t)O_EP2 * plane + o2] = south ? -s_ep3 : s_ep2;
    out[(size_t)O_EP3 * plane + o2] = south ? -s_ep2 : s_ep3;
    out[(size_t)O_EP4 * plane + o2] = ep4;
    out[(size_t)O_NCF * plane + o2] = south ? -s_nc : s_nc;
    out[(size_t)O_UBARO * plane + o2] = s_u;
    out[(size_t)O_FPHI * plane + o2] = s_fphi;
}


// ---------------------------------------------------------------------------------------------
// Column walk of the scan path (method 0): everything compute_lwa_and_barotropic_fluxes needs besides the LWA
// pieces a1 / a2 / ua2 / ncforce3d, which lwa_scan_kernel has already written.  One thread per (lon, hemispheric
// row) walks the levels upward (the vertical sums accumulate in numpy's order, oopinterface.py:405-420) with the
// loads of level k+1 issued before level k is evaluated (register double buffering: the kernel is bound by DRAM
// latency otherwise).  The eddy-momentum terms ep2(j) / ep3(j) are the same pointwise quantity
// G = (u - uref) cos^2 (v) at rows j+1 / j-1 (compute_flux_dirinv.f90:152-157), so each thread sums G for its own
// row only (hs) and baro_post_kernel picks the neighbours; the two rows with special rules (the first row's
// out-of-bounds uref and the boundary row jb, :173-178) are evaluated directly.
// ---------------------------------------------------------------------------------------------
// per-level operands of the column walk (the special-row members exist only in the SPECIAL instantiation)
struct ColumnLevelBase { double u0, v0, t0, a1, a2, u2, nc, ur, tr; };
struct ColumnLevelSpecial : ColumnLevelBase { double um, vm, urm, bu1, bv1, bu0, bv0, bur1, bur0; };
struct ColumnLevelPlain : ColumnLevelBase {
    static constexpr double um = 0, vm = 0, urm = 0, bu1 = 0, bv1 = 0, bu0 = 0, bv0 = 0, bur1 = 0, bur0 = 0;
};

// SPECIAL = false walks every ordinary row (launch: 32 x 8 tiles over all rows; the two special rows return at once);
// SPECIAL = true walks only the first row and the boundary row (launch: grid.y = 2), which need extra neighbour
// loads — keeping those out of the ordinary instantiation saves 36 registers there (3 CTAs per SM instead of 2).
// EXT = true: the LWA pieces never exist as 3-D intermediates — lwa_window_kernel has already written the 3-D LWA and
// the vertical sums of astar1 / astar2 / ua1 / ua2 (/ ncforce) — and this kernel only walks u, v and theta.
template <bool SPECIAL, bool EXT>
__global__ void __launch_bounds__(256, SPECIAL ? 2 : (EXT ? 4 : 3))
flux_column_kernel(FusedFluxArgs p, double* __restrict__ hs /* [b][nhemi][nd][imax] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int jfirst = p.jb + 1, jbrow = (p.jb >= 1) ? p.jb : p.nd;
    int j;
    if (SPECIAL) {
        if (blockIdx.y == 1 && jbrow == jfirst) return;
        j = blockIdx.y == 0 ? jfirst : jbrow;
    } else {
        j = blockIdx.y * blockDim.y + threadIdx.y + 1;  // hemispheric row 1..nd
        if (j == jfirst || j == jbrow) return;
    }
    const int b = blockIdx.z / p.nhemi, hidx = blockIdx.z % p.nhemi;
    const bool south = hidx == 1;
    if (i >= p.imax || j > p.nd) return;
    const bool skip_out = (j == 1 && p.nhemi == 2 && !south);  // the southern launch owns the equator row
    const int nd = p.nd, nlat = p.nlat, imax = p.imax, kmax = p.kmax, jb = p.jb;
    const size_t plane = (size_t)nlat * imax;
    const size_t f3 = (size_t)b * plane * kmax;
    const double *uu = p.uu + f3, *vv = p.vv + f3, *pt = p.pt + f3;
    const double *sa1 = EXT ? nullptr : p.sa1 + f3, *sa2 = EXT ? nullptr : p.sa2 + f3, *su2 = EXT ? nullptr : p.su2 + f3;
    const double* snc = (!EXT && p.has_ncforce) ? p.snc + f3 : nullptr;
    const double* uref = p.uref_st + (size_t)b * kmax * nlat;
    const double* tref = p.tref_st + (size_t)b * kmax * nlat;
    const double* tg = p.prof + (size_t)b * 4 * kmax + (south ? 0 : kmax);  // tn0 north / ts0 south
    const double* f11 = p.f11 + ((size_t)b * 2 + (south ? 0 : 1)) * kmax;
    double* lwa = p.lwa + f3;
    double* out = p.out2d + (size_t)b * O_COUNT * plane;
    const double sg = south ? -1.0 : 1.0;
    auto grow = [&](int jj) -> int { return south ? (nd - jj) : (nd - 2 + jj); };  // 0-based global row
    const int g = grow(j);
    const int jstart = jb + 1, jend = nd - 1;
    const bool in_range = (j >= jstart && j <= jend);
    const bool first = SPECIAL && (j == jstart);           // ep3 needs the reference's uref(0,k) == uref(jd,k-1)
    const bool brow = SPECIAL && ((jb >= 1) ? (j == jb) : (j == nd));   // boundary row jb, or its alias (i,nd,k-1) when jb = 0
    const bool own_g = (j >= jstart);                      // rows whose G sum some neighbour needs
    const double c0 = p.cosphi[j], cm = p.cosphi[j - 1], cl = p.clat[g];
    const size_t o2 = (size_t)g * imax + i;
    const size_t om = (size_t)grow(j - 1) * imax + i;                       // row j-1 (first row only)
    const size_t r1 = (size_t)grow(jb + 2) * imax + i, r0 = (size_t)grow(jb + 1) * imax + i;  // boundary row only
    const int gr1 = grow(jb + 2), gr0 = grow(jb + 1), gnd = grow(nd);

    using Lv = typename std::conditional<SPECIAL, ColumnLevelSpecial, ColumnLevelPlain>::type;
    auto fetch = [&](int k) -> Lv {  // 1-based level k (2..kmax)
        Lv L;
        const size_t lev = (size_t)(k - 1) * plane;
        L.u0 = uu[lev + o2]; L.v0 = vv[lev + o2];
        L.ur = uref[(size_t)(k - 1) * nlat + g];
        L.t0 = L.a1 = L.a2 = L.u2 = L.nc = L.tr = 0.0;
        if constexpr (SPECIAL) {
            L.um = L.vm = L.urm = 0.0;
            L.bu1 = L.bv1 = L.bu0 = L.bv0 = L.bur1 = L.bur0 = 0.0;
        }
        if (k < kmax) {
            if (in_range) {
                L.t0 = pt[lev + o2]; L.tr = tref[(size_t)(k - 1) * nlat + g];
                if (!EXT) {
                    L.a1 = sa1[lev + o2]; L.a2 = sa2[lev + o2]; L.u2 = su2[lev + o2];
                    if (snc) L.nc = snc[lev + o2];
                }
            }
            if constexpr (SPECIAL) {
                if (first) {
                    L.um = uu[lev + om]; L.vm = vv[lev + om];
                    L.urm = (j - jb - 1 >= 1) ? uref[(size_t)(k - 1) * nlat + grow(j - 1)] : uref[(size_t)(k - 2) * nlat + gnd];
                }
                if (brow) {
                    L.bu1 = uu[lev + r1]; L.bv1 = vv[lev + r1]; L.bu0 = uu[lev + r0]; L.bv0 = vv[lev + r0];
                    L.bur1 = uref[(size_t)(k - 1) * nlat + gr1]; L.bur0 = uref[(size_t)(k - 1) * nlat + gr0];
                }
            }
        }
        return L;
    };

    double s_a1 = 0, s_a2 = 0, s_ua1 = 0, s_ua2 = 0, s_ep1 = 0, s_nc = 0, s_u = 0, s_fphi = 0, s_g = 0;
    double s_e2 = 0, s_e3 = 0, ep4 = 0;   // directly evaluated ep2 / ep3 of the two special rows
    if (!EXT && !skip_out) {
        lwa[o2] = 0.0;
        lwa[(size_t)(kmax - 1) * plane + o2] = 0.0;
    }
    if (in_range) {  // low-level meridional heat flux from levels 1 and 2 (compute_flux_dirinv.f90:160-171)
        const double ep41 = 2. * p.om * p.sinphi[j] * c0 * p.dz / p.prefac;
        const double ep42 = p.ep42fac * (sg * vv[plane + o2]) * (pt[plane + o2] - tref[nlat + g]) / (tg[2] - tg[0]);
        double ep43 = (sg * vv[o2]) * (pt[o2] - tref[g]);
        ep43 = 0.5 * ep43 / (tg[1] - tg[0]);
        ep4 = ep41 * (ep42 + ep43);
    }
    // With the LWA pieces gone (EXT) a level is three 8-byte loads per thread: two levels are requested ahead so that
    // enough bytes are in flight per SM to cover the DRAM latency at the full bandwidth.
    constexpr bool DEEP = EXT && !SPECIAL;
    Lv cur = fetch(2);
    Lv nx2 = cur;
    if (DEEP) nx2 = fetch(kmax >= 3 ? 3 : 2);
    for (int k = 2; k <= kmax; ++k) {
        Lv nxt = cur;
        if (DEEP) {
            nxt = nx2;
            if (k + 2 <= kmax) nx2 = fetch(k + 2);
        } else if (k < kmax) nxt = fetch(k + 1);
        const double w = p.vw[k - 1];
        const double wd = w * p.dc;
        // (EXT: the running sums take one fused multiply-add per term, like the LWA sums of lwa_window_kernel)
        auto accw = [&](double sum, double x) -> double { return EXT ? __fma_rn(x, wd, sum) : sum + x * wd; };
        s_u = accw(s_u, cur.u0);
        s_fphi = accw(s_fphi, -(((cur.u0 - cur.ur) * cur.v0) * cl));
        if (k < kmax) {
            const size_t lev = (size_t)(k - 1) * plane;
            if (own_g) s_g = accw(s_g, (cur.u0 - cur.ur) * c0 * c0 * (sg * cur.v0));
            double e1 = 0, x_ua1 = 0;
            if (in_range) {
                const double vv0 = sg * cur.v0;
                x_ua1 = cur.ur * (cur.a1 + cur.a2);
                e1 = -0.5 * ((cur.u0 - cur.ur) * (cur.u0 - cur.ur));
                e1 = e1 + 0.5 * (vv0 * vv0);
                e1 = e1 - 0.5 * ((cur.t0 - cur.tr) * (cur.t0 - cur.tr)) * f11[k - 1];
                e1 = e1 * c0;
            }
            if (!EXT) {
                if (!skip_out) lwa[lev + o2] = fabs(cur.a1 + cur.a2);
                s_a1 = s_a1 + cur.a1 * wd;
                s_a2 = s_a2 + cur.a2 * wd;
                s_ua1 = s_ua1 + x_ua1 * wd;
                s_ua2 = s_ua2 + cur.u2 * wd;
                s_nc = s_nc + cur.nc * wd;
            }
            s_ep1 = accw(s_ep1, e1);
            if (first) s_e3 = s_e3 + ((cur.um - cur.urm) * cm * cm * (sg * cur.vm)) * wd;
            if (brow) {  // compute_flux_dirinv.f90:173-178
                double wk = wd;
                if (jb == 0) wk = (k >= 3) ? p.vw[k - 2] * p.dc : 0.0;  // aliased into level k-1; level 1 is not averaged
                s_e2 = s_e2 + ((cur.bu1 - cur.bur1) * p.cosjb * p.cosjb * (sg * cur.bv1)) * wk;
                s_e3 = s_e3 + ((cur.bu0 - cur.bur0) * p.cosjbm1 * p.cosjbm1 * (sg * cur.bv0)) * wk;
            }
        }
        cur = nxt;
    }
    hs[(((size_t)b * p.nhemi + hidx) * nd + (j - 1)) * imax + i] = s_g;
    if (skip_out) return;
    if (!EXT) {
        out[(size_t)O_ASTAR1 * plane + o2] = s_a1;
        out[(size_t)O_ASTAR2 * plane + o2] = s_a2;
        out[(size_t)O_LWA * plane + o2] = s_a1 + s_a2;
        out[(size_t)O_UA1 * plane + o2] = s_ua1;
        out[(size_t)O_UA2 * plane + o2] = s_ua2;
    }
    out[(size_t)O_EP1 * plane + o2] = s_ep1;
    // hemispheric-frame ep2 / ep3 of the special rows; baro_post_kernel fills the others from hs and applies the
    // southern swap + sign (oopinterface.py:790-792)
    out[(size_t)O_EP2 * plane + o2] = s_e2;
    out[(size_t)O_EP3 * plane + o2] = s_e3;
    out[(size_t)O_EP4 * plane + o2] = ep4;
    if (!EXT || !p.has_ncforce) out[(size_t)O_NCF * plane + o2] = south ? -s_nc : s_nc;
    out[(size_t)O_UBARO * plane + o2] = s_u;
    out[(size_t)O_FPHI * plane + o2] = s_fphi;
}

// barotropic post-processing (oopinterface.py:906-944, :964-982; utilities.py:189-233)
__global__ void __launch_bounds__(256)
baro_post_kernel(int nlon, int nlat, double a, double dphi, double dlambda, const double* __restrict__ clat,
                 double* __restrict__ out2d, const double* __restrict__ hs, int nhemi, int nd, int jb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, b = blockIdx.z;
    if (i >= nlon) return;
    const size_t plane = (size_t)nlat * nlon;
    double* out = out2d + (size_t)b * O_COUNT * plane;
    const size_t o = (size_t)j * nlon + i;
    const double cl = clat[j];
    if (hs) {
        // scan path: ep2(jh) = G(jh+1), ep3(jh) = G(jh-1) from the own-row sums of flux_column_kernel; the first row's
        // ep3 and the boundary row are already in place.  South: swapped and negated (oopinterface.py:790-792).
        const bool south = (nhemi == 2) && (j <= nd - 1);
        const bool covered = south || (j >= nd - 1);
        if (covered) {
            const int jh = south ? (nd - j) : (j - nd + 2);   // hemispheric row 1..nd
            const double* H = hs + (((size_t)b * nhemi + (south ? 1 : 0)) * nd) * nlon + i;
            const int jstart = jb + 1, jend = nd - 1;
            const bool brow = (jb >= 1) ? (jh == jb) : (jh == nd);
            double e2 = out[(size_t)O_EP2 * plane + o], e3 = out[(size_t)O_EP3 * plane + o];
            if (!brow) {
                e2 = (jh >= jstart && jh <= jend) ? H[(size_t)jh * nlon] : 0.0;             // row jh+1
                if (jh > jstart && jh <= jend) e3 = H[(size_t)(jh - 2) * nlon];               // row jh-1
                else if (jh != jstart) e3 = 0.0;
            }
            out[(size_t)O_EP2 * plane + o] = south ? -e3 : e2;
            out[(size_t)O_EP3 * plane + o] = south ? -e2 : e3;
        }
    }
    out[(size_t)O_DIVEMF * plane + o] = (out[(size_t)O_EP2 * plane + o] - out[(size_t)O_EP3 * plane + o]) / (2 * a * dphi * cl);
    auto zsum = [&](int ii) -> double {
        const size_t oo = (size_t)j * nlon + ii;
        return out[(size_t)O_UA1 * plane + oo] + out[(size_t)O_UA2 * plane + oo] + out[(size_t)O_EP1 * plane + oo];
    };
    const int ip = (i == nlon - 1) ? 0 : i + 1, im = (i == 0) ? nlon - 1 : i - 1;
    double zd = zsum(ip) - zsum(im);
    if (fabs(cl) > 1.e-5) zd = zd * (-1. / (a * cl * 2. * dlambda));
    out[(size_t)O_CONV * plane + o] = zd;
    out[(size_t)O_FLAMBDA * plane + o] = zsum(i);
}

// ---------------------------------------------------------------------------------------------
// Reference-state profiles of one (level, hemisphere, snapshot): cumulative area/circulation of the PV histogram
// and its inversion at the equivalent latitudes (qref, cref; ckref from the absolute-vorticity histogram for
// NHN22), the zonal-mean circulation cbar, the normalisation of qref, and the hemispheric ubar / tbar rows
// (compute_reference_states.f90:60-130, compute_qref_and_fawa_first.f90:44-160).  Same arithmetic and summation
// order as qref_invert_kernel + cbar_normalise_kernel + extract_rows_kernel + tb_mean_kernel, one launch.
// work layout per (b, hs): qref, cref, cbar, ubar, tbar, fjk [nd*kmax each]; uref, tref [jd*kmax]; tb[kmax]; uz[nd]
// ---------------------------------------------------------------------------------------------
struct LevelArgs {
    int kind, kmax, nlat, nd, nb;
    double a, dphi, pi, area_scale, cbound;
    const double *zm, *alat, *cosmid, *norm, *cosw;
    const long long* hist;   // fixed-point bins of area_hist3_kernel
    const unsigned long long* ext;
    double *work, *ckref;
    size_t per, njk, njd;
};

__global__ void __launch_bounds__(256)
refstate_level_kernel(LevelArgs p) {
    extern __shared__ double sh[];
    const int nb = p.nb, nd = p.nd, nlat = p.nlat, kmax = p.kmax;
    double* aan = sh;            // [nb]
    double* ccn = aan + nb;      // [nb]
    double* aav = ccn + nb;      // [nb]  (vorticity)
    double* ccv = aav + nb;      // [nb]
    double* qrow = ccv + nb;     // [nd]
    double* crow = qrow + nd;    // [nd]
    double* term = crow + nd;    // [nd]
    const int k = blockIdx.x, hs = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
    const double sgn = hs == 0 ? 1.0 : -1.0;
    const bool interior = (k >= 1 && k <= kmax - 2);
    const bool do_vort = (p.kind != 0) && hs == 0 && interior;
    const double* zmb = p.zm + (size_t)b * 3 * kmax * nlat;
    double* W = p.work + ((size_t)b * 2 + hs) * p.per;
    double *qref = W, *cref = W + p.njk, *cbar = W + 2 * p.njk, *ubar = W + 3 * p.njk, *tbar = W + 4 * p.njk;
    auto grow = [&](int j) -> int { return hs == 0 ? (nd - 1 + j) : (nlat - 1 - (nd - 1 + j)); };  // 0-based hemispheric -> global
    // hemispheric zonal-mean rows (all levels)
    for (int j = tid; j < nd; j += blockDim.x) {
        ubar[(size_t)k * nd + j] = zmb[((size_t)1 * kmax + k) * nlat + grow(j)];
        tbar[(size_t)k * nd + j] = zmb[((size_t)2 * kmax + k) * nlat + grow(j)];
    }
    if (p.kind == 0) {  // hemispheric-mean theta of the SOR routine (compute_reference_states.f90:60-65)
        __syncthreads();
        if (tid == 0) {
            double s = 0., w = 0.;
            for (int j = 0; j < nd; ++j) { s = s + p.cosw[j] * tbar[(size_t)k * nd + j]; w = w + p.cosw[j]; }
            W[6 * p.njk + 2 * p.njd + k] = s / w;
        }
    }
    if (!interior) return;
    const unsigned long long* extb = p.ext + (size_t)b * 2 * kmax * 2;
    const long long* histb = p.hist + (size_t)b * 4 * kmax * 2 * nb;
    double qmax = ordered_to_f64(extb[k * 2 + 0]), qmin = ordered_to_f64(extb[k * 2 + 1]);
    if (sgn < 0) { const double t = qmax; qmax = -qmin; qmin = -t; }
    const double dq = (qmax - qmin) / (double)(nb - 1);
    double vmax = 0., dqv = 1.;
    if (do_vort) {
        vmax = ordered_to_f64(extb[(kmax + k) * 2 + 0]);
        const double vmin = ordered_to_f64(extb[(kmax + k) * 2 + 1]);
        dqv = (vmax - vmin) / (double)(nb - 1);
    }
    {   // cumulative sums in the reference's order (bin 1 is never added)
        const long long* an = histb + ((size_t)0 * kmax + k) * 2 * nb;     // +pv, all points
        const long long* sx = histb + ((size_t)1 * kmax + k) * 2 * nb;     // southern bins of the edge points
        const long long* av = histb + ((size_t)2 * kmax + k) * 2 * nb;     // avort
        const long long* nx = histb + ((size_t)3 * kmax + k) * 2 * nb;     // northern bins of the edge points
        // fixed point -> double (the scales of area_hist3_kernel; the southern frame has the same |extrema|)
        const double ia = 1.0 / p.area_scale;
        const double icp = 1.0 / hist_circ_scale(ordered_to_f64(extb[k * 2 + 0]), ordered_to_f64(extb[k * 2 + 1]), p.cbound);
        const double icv = do_vort ? 1.0 / hist_circ_scale(ordered_to_f64(extb[(kmax + k) * 2 + 0]),
                                                           ordered_to_f64(extb[(kmax + k) * 2 + 1]), p.cbound) : 1.0;
        for (int t = tid; t < nb; t += blockDim.x) {
            if (hs == 0) {
                aan[t] = (double)an[t] * ia; ccn[t] = (double)an[nb + t] * icp;
            } else {   // mirror: southern bin t <- northern bin nb-2-t (circulation changes sign), plus the edge points
                const int tn = nb - 2 - t;
                long long a = sx[t], c = sx[nb + t];
                if (tn >= 0) { a += an[tn] - nx[tn]; c += -(an[nb + tn] - nx[nb + tn]); }
                aan[t] = (double)a * ia; ccn[t] = (double)c * icp;
            }
            if (do_vort) { aav[t] = (double)av[t] * ia; ccv[t] = (double)av[nb + t] * icv; }
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0., c = 0.;
            aan[0] = 0.; ccn[0] = 0.;
            for (int nn = 1; nn < nb; ++nn) { a = a + aan[nn]; c = c + ccn[nn]; aan[nn] = a; ccn[nn] = c; }
        }
        if (tid == 32 && do_vort) {
            double a = 0., c = 0.;
            aav[0] = 0.; ccv[0] = 0.;
            for (int nn = 1; nn < nb; ++nn) { a = a + aav[nn]; c = c + ccv[nn]; aav[nn] = a; ccv[nn] = c; }
        }
        __syncthreads();
    }
    for (int j = tid; j < nd; j += blockDim.x) {
        double qv = 0., cv = 0., kv = 0.;
        if (j < nd - 1) {
            const double x = p.alat[j];
            int lo = 0, hi = nb;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (aan[mid] <= x) lo = mid + 1; else hi = mid; }
            if (lo >= 1 && lo <= nb - 1) {
                const double a0 = aan[lo - 1], a1 = aan[lo];
                const double dd = (x - a0) / (a1 - a0);
                const double qn0 = qmax - dq * (double)(lo - 1), qn1 = qmax - dq * (double)lo;
                qv = qn0 * (1. - dd) + qn1 * dd;
                cv = ccn[lo - 1] * (1. - dd) + ccn[lo] * dd;
            }
            if (do_vort) {
                lo = 0; hi = nb;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (aav[mid] <= x) lo = mid + 1; else hi = mid; }
                if (lo >= 1 && lo <= nb - 1) {
                    const double a0 = aav[lo - 1], a1 = aav[lo];
                    const double dd = (x - a0) / (a1 - a0);
                    kv = ccv[lo - 1] * (1. - dd) + ccv[lo] * dd;
                }
            }
        } else {
            qv = qmax;
        }
        qrow[j] = qv;
        crow[j] = cv;
        if (do_vort) p.ckref[(size_t)b * p.njk + (size_t)k * nd + j] = kv;
        // cbar increment of row j (added when marching from j+1 down to j)
        if (j < nd - 1) {
            const double qb1 = sgn * zmb[(size_t)k * nlat + grow(j + 1)], qb0 = sgn * zmb[(size_t)k * nlat + grow(j)];
            term[j] = 0.5 * (qb1 + qb0) * p.a * p.dphi * 2. * p.pi * p.a * p.cosmid[j];
        }
    }
    __syncthreads();
    if (tid == 0) {
        double c = 0.0;
        cbar[(size_t)k * nd + nd - 1] = 0.0;
        for (int j = nd - 2; j >= 0; --j) { c = c + term[j]; cbar[(size_t)k * nd + j] = c; }
    }
    for (int j = tid; j < nd; j += blockDim.x) {
        cref[(size_t)k * nd + j] = crow[j];
        if (j >= 1) qrow[j] = qrow[j] / p.norm[j];
    }
    __syncthreads();
    for (int j = tid; j < nd; j += blockDim.x) qref[(size_t)k * nd + j] = (j == 0) ? 2. * qrow[1] - qrow[2] : qrow[j];
}

// small helpers -------------------------------------------------------------------------------
static size_t push_tab(std::vector<double>& tab, const std::vector<double>& v) {
    const size_t o = tab.size();
    tab.insert(tab.end(), v.begin(), v.end());
    return o;
}

}  // namespace falwa

using namespace falwa;

// ================================ C ABI, Section B =============================================
extern "C" int falwa_b200_plan_create(falwa_plan** out, int kind, int nlon, int nlat, int nlev, int kmax, int jb,
                                      int npart, int maxit, const double* plev_to_height_host,
                                      const double* theta_factor_host, const double* clat_host,
                                      const double* sinlat_host, double a, double omega, double dz, double h,
                                      double rr, double cp, double tol, double rjac, double prefactor) {
    if (!out || !plev_to_height_host || !theta_factor_host || !clat_host || !sinlat_host) return FALWA_ERR_NULL;
    if ((kind != 0 && kind != 1) || nlon < 4 || nlat < 9 || (nlat % 2) != 1 || nlev < 2 || kmax < 4 || jb < 0)
        return FALWA_ERR_ARG;
    const int nd = nlat / 2 + 1;
    if (kind == 0) jb = 0;
    const int jd = nd - jb;
    if (jd < 4) return FALWA_ERR_ARG;
    falwa_plan* P = new falwa_plan();
    P->kind = kind; P->nlon = nlon; P->nlat = nlat; P->nlev = nlev; P->kmax = kmax; P->jb = jb; P->nd = nd; P->jd = jd;
    P->npart = npart > 0 ? npart : nlat; P->maxit = maxit;
    P->a = a; P->omega = omega; P->dz = dz; P->h = h; P->rr = rr; P->cp = cp; P->tol = tol; P->rjac = rjac;
    P->prefactor = prefactor;
    const double pi = host_pi();
    // oopinterface.py:164-165: np.deg2rad(180/(nlat-1)), np.deg2rad(360/nlon)
    P->dphi = (180. / (double)(nlat - 1)) * (pi / 180.);
    P->dlambda = (360. / (double)nlon) * (pi / 180.);
    const double dphi = P->dphi, rkappa = rr / cp;
    P->height.resize(kmax);
    for (int k = 0; k < kmax; ++k) P->height[k] = (double)k * dz;
    std::vector<double> tab;
    P->o_fac = push_tab(tab, std::vector<double>(theta_factor_host, theta_factor_host + nlev));
    P->o_clat = push_tab(tab, std::vector<double>(clat_host, clat_host + nlat));
    P->o_sinlat = push_tab(tab, std::vector<double>(sinlat_host, sinlat_host + nlat));
    P->sinlat.assign(sinlat_host, sinlat_host + nlat);
    // interpolation tables (scipy interp1d linear, extrapolating)
    std::vector<double> whi(kmax), wlo(kmax);
    std::vector<int> lo(kmax);
    const double* x = plev_to_height_host;
    for (int k = 0; k < kmax; ++k) {
        const double xn = P->height[k];
        int idx = 0;
        while (idx < nlev && x[idx] < xn) ++idx;  // searchsorted(side='left')
        idx = std::max(1, std::min(nlev - 1, idx));
        lo[k] = idx - 1;
        const double xl = x[idx - 1], xh = x[idx];
        whi[k] = (xn - xl) / (xh - xl);
        wlo[k] = (xh - xn) / (xh - xl);
    }
    P->o_whi = push_tab(tab, whi);
    P->o_wlo = push_tab(tab, wlo);
    {
        std::vector<double> qg;
        qgpv_tables_host(kind == 0 ? 0 : 1, nlat, kmax, P->height.data(), a, omega, h, qg);
        P->o_qg = push_tab(tab, qg);
    }
    std::vector<double> alat, da;
    host_alat_da(nd, nlat, a, dphi, P->dlambda, alat, da);
    P->o_alat = push_tab(tab, alat);
    P->o_da = push_tab(tab, da);
    {   // fixed-point scales: the area of all cells, and (per unit of |field|) of all circulation terms, fits 2^61
        double total = 0.0, damax = 0.0;
        for (double x : da) { total += std::fabs(x) * (double)nlon; damax = std::max(damax, std::fabs(x)); }
        P->hist_area_scale = hist_pow2_scale(total);
        P->hist_cbound = damax * (double)nlon * (double)nlat;
    }
    std::vector<double> cosmid(nd, 0.), norm(nd, 1.), vw(kmax);
    for (int j = 1; j <= nd; ++j) {
        cosmid[j - 1] = std::cos(dphi * ((double)j - 0.5));
        const double s = std::sin(dphi * (double)(j - 1));
        norm[j - 1] = (kind == 0) ? 2. * omega * s : s;
    }
    for (int k = 0; k < kmax; ++k) vw[k] = std::exp(-P->height[k] / h);
    P->o_cosmid = push_tab(tab, cosmid);
    P->o_norm = push_tab(tab, norm);
    P->o_vw = push_tab(tab, vw);
    std::vector<double> eref(kmax), ez(kmax);
    for (int k = 1; k <= kmax; ++k) {
        eref[k - 1] = std::exp(rkappa * (dz * (double)(k - 1)) / h);
        ez[k - 1] = std::exp((dz * (double)(k - 1)) / h);
    }
    P->o_eref = push_tab(tab, eref);
    P->o_ez = push_tab(tab, ez);
    if (kind == 0) {  // NH18 SOR tables (compute_reference_states.f90:135-162, :188-190, :60-65)
        std::vector<double> ajk(nd, 0.), bjk(nd, 0.), fact(nd, 0.), cosj(nd), cor(nd), cosw(nd), c1(nd, 0.), c2(nd, 1.);
        for (int j = 1; j <= nd; ++j) {
            const double phi0 = dphi * (double)(j - 1);
            cosj[j - 1] = std::cos(phi0);
            cor[j - 1] = 2. * omega * std::sin(phi0);
            cosw[j - 1] = std::cos(dphi * (double)(j + nd - 2) - 0.5 * pi);
            if (j >= 2 && j <= nd - 1) {
                const double phip = ((double)j - 0.5) * dphi, phim = ((double)j - 1.5) * dphi;
                ajk[j - 1] = 1. / (std::sin(phip) * std::cos(phip));
                bjk[j - 1] = 1. / (std::sin(phim) * std::cos(phim));
                fact[j - 1] = 4. * omega * omega * h * a * a * std::sin(phi0) * dphi * dphi / (dz * dz * rr * std::cos(phi0));
                c1[j - 1] = dz * rr * std::cos(phi0) * std::exp(-(dz * (double)(kmax - 2)) * rkappa / h);
                c2[j - 1] = 2. * omega * std::sin(phi0) * dphi * h * a;
            }
        }
        P->o_ajk = push_tab(tab, ajk); P->o_bjk = push_tab(tab, bjk); P->o_fact = push_tab(tab, fact);
        P->o_cosj = push_tab(tab, cosj); P->o_cor = push_tab(tab, cor); P->o_cosw = push_tab(tab, cosw);
        P->o_uzc1 = push_tab(tab, c1); P->o_uzc2 = push_tab(tab, c2);
    } else {  // NHN22 tables; latitude step of these routines is pi/(jmax-1)
        const int n = jd - 2;
        const double dp = pi / (double)(nlat - 1);
        std::vector<double> ajk(n), bjk(n), fact(n), tc(n), td(n), cosrow(jd), sinrow(jd), sinnd(nd);
        for (int jj = jb + 2; jj <= nd - 1; ++jj) {
            const int i = jj - jb - 2;
            const double phi0 = (double)(jj - 1) * dp, phip = ((double)jj - 0.5) * dp, phim = ((double)jj - 1.5) * dp;
            ajk[i] = 1. / (std::sin(phip) * std::cos(phip));
            bjk[i] = 1. / (std::sin(phim) * std::cos(phim));
            fact[i] = 4. * omega * omega * h * a * a * std::sin(phi0) * dp * dp / (dz * dz * rr * std::cos(phi0));
            const double phi0b = (double)(jj - 1) * dphi;  // compute_qref_and_fawa_first uses the dphi argument
            tc[i] = -dz * rr * std::cos(phi0b) * std::exp(-(dz * (double)(kmax - 2)) * rkappa / h);
            td[i] = 4. * omega * std::sin(phi0b) * dphi * h * a;
        }
        for (int j = 1; j <= jd; ++j) {
            cosrow[j - 1] = std::cos(dp * (double)(j + jb - 1));
            sinrow[j - 1] = std::sin(dp * (double)(j - 1 + jb));
        }
        for (int j = 1; j <= nd; ++j) sinnd[j - 1] = std::sin(dp * (double)(j - 1));
        P->o_najk = push_tab(tab, ajk); P->o_nbjk = push_tab(tab, bjk); P->o_nfact = push_tab(tab, fact);
        P->o_topcoef = push_tab(tab, tc); P->o_topden = push_tab(tab, td);
        P->o_cosrow = push_tab(tab, cosrow); P->o_sinrow = push_tab(tab, sinrow); P->o_sinnd = push_tab(tab, sinnd);
    }
    {  // flux tables (compute_flux_dirinv.f90: dp = pi/(jmax-1), northern-hemisphere branch)
        const double dp = pi / (double)(nlat - 1);
        const int nt = nd + 2;
        std::vector<double> cosphi(nt), sinphi(nt), aaw(nt), ep11a(kmax, 0.);
        for (int jj = 0; jj <= nd + 1; ++jj) {
            const double phi1 = dp * (double)(jj - 1);
            cosphi[jj] = std::cos(phi1);
            sinphi[jj] = std::sin(phi1);
            aaw[jj] = a * dp * std::cos(phi1);
        }
        for (int k = 2; k <= kmax - 1; ++k) ep11a[k - 1] = std::exp(-rkappa * (dz * (double)(k - 1)) / h);
        P->o_cosphi = push_tab(tab, cosphi); P->o_sinphi = push_tab(tab, sinphi); P->o_aaw = push_tab(tab, aaw);
        P->o_ep11a = push_tab(tab, ep11a);
    }
    P->ntab = tab.size();
    cudaError_t e = cudaMalloc((void**)&P->d_tab, tab.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&P->d_lo, 3 * kmax * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(P->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice);
    {   // prefetch schedule of the fused stage-1 kernel: after output level k, the next two input levels (in order of
        // need) that are not part of bracket(k); -1 when there is none.  d_lo = lo[kmax], pf1[kmax], pf2[kmax].
        lo.resize(3 * (size_t)kmax, -1);
        for (int k = 0; k < kmax; ++k) {
            int found = 0, F[2] = {-1, -1};
            for (int k2 = k + 1; k2 < kmax && found < 2; ++k2)
                for (int lev = lo[k2]; lev <= lo[k2] + 1 && found < 2; ++lev)
                    if (lev != lo[k] && lev != lo[k] + 1 && lev != F[0]) F[found++] = lev;
            lo[kmax + k] = F[0];
            lo[2 * (size_t)kmax + k] = F[1];
        }
    }
    if (e == cudaSuccess) e = cudaMemcpy(P->d_lo, lo.data(), 3 * kmax * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {  // keep stream-ordered scratch cached between calls
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    if (e != cudaSuccess) {
        if (P->d_tab) cudaFree(P->d_tab);
        if (P->d_lo) cudaFree(P->d_lo);
        delete P;
        return (int)e;
    }
    *out = P;
    return 0;
}

extern "C" int falwa_b200_plan_destroy(falwa_plan* P) {
    if (!P) return 0;
    if (P->d_tab) cudaFree(P->d_tab);
    if (P->d_lo) cudaFree(P->d_lo);
    delete P;
    return 0;
}

// hemispheric cos-weighted mean theta per pressure level: means [batch][2][nlev] (0 = south, 1 = north)
extern "C" int falwa_b200_theta_hemispheric_means(const falwa_plan* P, int batch, const double* t_in, double* means,
                                                  void* stream) {
    if (!P || !t_in || !means) return FALWA_ERR_NULL;
    if (batch < 1) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceScratch<double> rowmean;
    int rc = rowmean.alloc((size_t)batch * P->nlev * P->nlat, st);
    if (rc) return rc;
    return theta_means_launch(P->nlon, P->nlat, P->nlev, P->jd, batch, t_in, P->d_tab + P->o_fac, P->d_tab + P->o_clat,
                              rowmean.d, means, st);
}

namespace falwa {
// data_on_evenly_spaced_pseudoheight_grid: theta = T * exp(kappa z / H) level by level (oopinterface.py:120-122)
__global__ void __launch_bounds__(256)
theta_from_t_kernel(size_t plane, int nlev, const double* __restrict__ t_in, const double* __restrict__ fac,
                    double* __restrict__ theta) {
    const size_t n = plane * nlev;
    const size_t base = (size_t)blockIdx.y * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        st_stream(&theta[base + t], ld_stream(&t_in[base + t]) * fac[t / plane]);
}
}  // namespace falwa

// vertical interpolation + QGPV.  profiles_host [batch][4][kmax] = t0_s, t0_n, stat_s, stat_n at `height`.
// zm [batch][3][kmax][nlat] / ext (uint64 [batch][2][kmax][2]), both optional: zonal means of qgpv, u, theta and the
// per-level extrema of qgpv and avort, produced on the fly for falwa_b200_reference_states.
extern "C" int falwa_b200_interpolate_fields(const falwa_plan* P, int batch, const double* u_in, const double* v_in,
                                             const double* t_in, int already_on_height_grid,
                                             const double* profiles_host, double* u, double* v, double* theta,
                                             double* avort, double* qgpv, double* zm, void* ext, void* stream) {
    if (!P || !u_in || !v_in || !t_in || !profiles_host || !u || !v || !theta || !avort || !qgpv) return FALWA_ERR_NULL;
    if (batch < 1) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const double* d = P->d_tab;
    const int kmax = P->kmax;
    int rc;
    // profiles + reciprocal static stability: [batch][6][kmax]
    std::vector<double> hp((size_t)batch * 6 * kmax);
    for (int b = 0; b < batch; ++b) {
        const double* src = profiles_host + (size_t)b * 4 * kmax;
        double* dst = hp.data() + (size_t)b * 6 * kmax;
        for (int t = 0; t < 4 * kmax; ++t) dst[t] = src[t];
        for (int t = 0; t < 2 * kmax; ++t) dst[4 * kmax + t] = 1. / src[2 * kmax + t];
    }
    DeviceTable prof;
    if ((rc = prof.upload(hp, st))) return rc;
    const int jdsplit = (P->kind == 0) ? P->nlat : P->nd;  // NHN22: rows j <= equator use the southern profile
    if (!already_on_height_grid && !(zm && ext) && tune_int("FALWA_FUSED_STAGE1", 1)) {
        // one pass: interpolation + absolute vorticity + QGPV
        return interp_qgpv_launch(P->kind == 0 ? 0 : 1, P->nlon, P->nlat, P->nlev, kmax, jdsplit, batch, u_in, v_in, t_in,
                                  P->d_lo, d + P->o_whi, d + P->o_wlo, d + P->o_fac, d + P->o_qg, prof.d, u, v, theta, qgpv,
                                  avort, st);
    }
    if (already_on_height_grid && theta != t_in) {   // u, v are used as they are; T -> theta on the device
        FALWA_KERNEL("theta_from_t_kernel", st);
        theta_from_t_kernel<<<dim3(148 * 8, batch), 256, 0, st>>>((size_t)P->nlat * P->nlon, P->nlev, t_in, d + P->o_fac, theta);
        FALWA_LAUNCH_CHECK();
    }
    if (!already_on_height_grid) {
        if ((rc = interp_launch(P->nlon, P->nlat, P->nlev, kmax, batch, u_in, v_in, t_in, P->d_lo, d + P->o_whi,
                                d + P->o_wlo, d + P->o_fac, u, v, theta, st)))
            return rc;
    }
    DeviceScratch<double> part;
    if (zm && ext) {
        if ((rc = part.alloc((size_t)batch * 3 * kmax * P->nlat * ceil_div(P->nlon, 32), st))) return rc;
    }
    return qgpv_launch(P->kind == 0 ? 0 : 1, P->nlon, P->nlat, kmax, jdsplit, batch, u, v, theta, d + P->o_qg, prof.d,
                       qgpv, avort, st, 1, zm, (unsigned long long*)ext, part.d);
}

// reference states of both hemispheres -> global-latitude storage [batch][kmax][nlat]
//   qref_st: stored (normalised) qref; qref_cu: qref in s^-1; uref_st; ptref_st.
//   num_of_iter_host [batch][2] (NH18 SOR iterations north, south; 0 for NHN22).  Synchronises.
extern "C" int falwa_b200_reference_states(const falwa_plan* P, int batch, const double* qgpv, const double* u,
                                           const double* theta, const double* avort, const double* profiles_host,
                                           int north_only, double* qref_st, double* qref_cu, double* uref_st,
                                           double* ptref_st, int* num_of_iter_host, const double* zm_in,
                                           const void* ext_in, void* stream) {
    if (!P || !qgpv || !u || !theta || !avort || !profiles_host || !qref_st || !qref_cu || !uref_st || !ptref_st ||
        !num_of_iter_host)
        return FALWA_ERR_NULL;
    if (batch < 1) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const double* d = P->d_tab;
    const int nlon = P->nlon, nlat = P->nlat, kmax = P->kmax, nd = P->nd, jd = P->jd, jb = P->jb;
    const int nb = (P->kind == 0) ? P->npart : nlat;  // NHN22 ignores npart
    const double pi = host_pi(), rkappa = P->rr / P->cp;
    const size_t njk = (size_t)nd * kmax, njd = (size_t)jd * kmax;
    int rc;
    DeviceScratch<double> zm, hist, prof, work;
    DeviceScratch<unsigned long long> ext;
    DeviceScratch<int> nit;
    const bool have_stats = zm_in && ext_in;
    const double area_scale = P->hist_area_scale, cbound = P->hist_cbound;
    if (!have_stats) {
        if ((rc = zm.alloc((size_t)batch * 3 * kmax * nlat, st))) return rc;
        if ((rc = ext.alloc((size_t)batch * 2 * kmax * 2, st))) return rc;
    }
    if ((rc = hist.alloc((size_t)batch * 4 * kmax * 2 * nb, st, true))) return rc;
    if ((rc = nit.alloc((size_t)batch * 2, st, true))) return rc;
    // per (b, hemisphere): qref, cref, cbar, ubar, tbar [njk each]; uref, tref [njd]; fjk [njk]; tb[kmax]; uz[nd]
    const size_t per = 6 * njk + 2 * njd + kmax + nd;
    const size_t ckoff = (size_t)batch * 2 * per;  // ckref per snapshot
    if ((rc = work.alloc(ckoff + (size_t)batch * njk, st, true))) return rc;
    const double* zmd = have_stats ? zm_in : zm.d;
    const unsigned long long* extd = have_stats ? (const unsigned long long*)ext_in : ext.d;
    if (!have_stats) {
        { FALWA_KERNEL("init_ext_kernel", st); init_ext_kernel<<<ceil_div((long long)batch * 2 * kmax * 2, 128), 128, 0, st>>>(ext.d, batch * 2 * kmax * 2); }
        { FALWA_KERNEL("zonal_stats_kernel", st); zonal_stats_kernel<<<dim3(ceil_div((long long)nlat * kmax, 8), 1, batch), 256, 0, st>>>(nlon, nlat, kmax, qgpv, u, theta, avort, zm.d, ext.d,
                                                                    P->kind == 0 ? 1 : 3); }
    }
    FALWA_LAUNCH_CHECK();
    {
        int slabs = std::max(1, std::min(nlat, (148 * tune_int("FALWA_HIST_CTAS", 12) + (kmax - 3)) / (kmax - 2) / std::max(1, std::min(batch, 4))));
        const int rows = ceil_div(nlat, slabs);
        slabs = ceil_div(nlat, rows);
        const size_t smem = sizeof(unsigned long long) * 4 * 2 * nb;
        const int hrun = tune_int("FALWA_HIST_RUN", 64);
        unsigned long long* histi = reinterpret_cast<unsigned long long*>(hist.d);
        dim3 hgrid(slabs, kmax - 2, batch);
        FALWA_KERNEL("area_hist3_kernel", st);
        if (P->kind == 0) {  // NH18: pv histogram with circulation sums (SOR boundary condition), no vorticity histogram
            FALWA_CUDA_TRY(cudaFuncSetAttribute(area_hist3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            area_hist3_kernel<true, false><<<hgrid, 256, smem, st>>>(nlon, nlat, kmax, nb, qgpv, avort, extd, d + P->o_da, rows, histi, hrun, area_scale, cbound);
        } else if (tune_int("FALWA_HIST_SPLIT", 1)) {   // NHN22, one field per launch
            FALWA_CUDA_TRY(cudaFuncSetAttribute(area_hist3_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FALWA_CUDA_TRY(cudaFuncSetAttribute(area_hist3_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            area_hist3_kernel<false, false, true><<<hgrid, 256, smem, st>>>(nlon, nlat, kmax, nb, qgpv, avort, extd, d + P->o_da, rows, histi, hrun, area_scale, cbound);
            area_hist3_kernel<false, true, false><<<hgrid, 256, smem, st>>>(nlon, nlat, kmax, nb, qgpv, avort, extd, d + P->o_da, rows, histi, hrun, area_scale, cbound);
        } else {             // NHN22: areas of the pv histogram, areas + circulation of the vorticity histogram
            FALWA_CUDA_TRY(cudaFuncSetAttribute(area_hist3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            area_hist3_kernel<false, true><<<hgrid, 256, smem, st>>>(nlon, nlat, kmax, nb, qgpv, avort, extd, d + P->o_da, rows, histi, hrun, area_scale, cbound);
        }
        FALWA_LAUNCH_CHECK();
    }
    // host-built per-problem tables (the 2*batch eigen-decompositions run on a few host threads)
    std::vector<double> htab;
    std::vector<size_t> o_spec((size_t)batch * 2), o_amp((size_t)batch * 2);
    {
        const int m = kmax - 2;
        const size_t each = (P->kind == 0) ? (size_t)2 * kmax : (size_t)m + 2 * (size_t)m * m + 1;
        htab.assign(each * batch * 2, 0.0);
        std::atomic<int> err{0};
        auto work_fn = [&](int q0, int q1) {
            for (int q = q0; q < q1; ++q) {
                const int b = q / 2, hs = q % 2;
                // hs 0 = north uses stat_n (row 3), hs 1 = south uses stat_s (row 2)
                const double* stat = profiles_host + ((size_t)b * 4 + (hs == 0 ? 3 : 2)) * kmax;
                double* dst = htab.data() + each * q;
                if (P->kind == 0) {
                    double *amp = dst, *amm = dst + kmax;
                    for (int k = 2; k <= kmax - 1; ++k) {
                        const double zk = P->dz * (double)(k - 1), zkp = P->dz * (double)k, zkm = P->dz * (double)(k - 2);
                        const double zp = 0.5 * (zkp + zk), zmm = 0.5 * (zkm + zk);
                        const double statp = 0.5 * (stat[k] + stat[k - 1]), statm = 0.5 * (stat[k - 2] + stat[k - 1]);
                        amp[k - 1] = std::exp(-zp / P->h) * std::exp(rkappa * zp / P->h) / statp;
                        amm[k - 1] = std::exp(-zmm / P->h) * std::exp(rkappa * zmm / P->h) / statm;
                    }
                    o_amp[q] = each * q;
                } else {
                    std::vector<double> spec;
                    if (vertical_operator_spectrum(kmax, stat, P->dz, P->h, rkappa, spec)) { err = 1; continue; }
                    std::copy(spec.begin(), spec.end(), dst);
                    o_spec[q] = each * q;
                }
            }
        };
        const int nq = batch * 2;
        const int nthr = (P->kind == 0) ? 1 : std::max(1, std::min({nq / 4, (int)std::thread::hardware_concurrency(), 16}));
        if (nthr <= 1) work_fn(0, nq);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthr; ++t) pool.emplace_back(work_fn, (int)((long long)nq * t / nthr), (int)((long long)nq * (t + 1) / nthr));
            for (auto& th : pool) th.join();
        }
        if (err) return FALWA_ERR_UNSUPPORTED;
    }
    const size_t o_prof = push_tab(htab, std::vector<double>(profiles_host, profiles_host + (size_t)batch * 4 * kmax));
    DeviceTable HT;
    if ((rc = HT.upload(htab, st))) return rc;
    FALWA_CUDA_TRY(cudaMemsetAsync(uref_st, 0, sizeof(double) * (size_t)batch * kmax * nlat, st));
    // ---- per (snapshot, hemisphere) profile kernels ---------------------------------------------
    auto W = [&](int b, int hs, int which) -> double* {  // which: 0 qref 1 cref 2 cbar 3 ubar 4 tbar 5 fjk
        return work.d + ((size_t)b * 2 + hs) * per + (size_t)which * njk;
    };
    auto WU = [&](int b, int hs) { return work.d + ((size_t)b * 2 + hs) * per + 6 * njk; };
    auto WT = [&](int b, int hs) { return work.d + ((size_t)b * 2 + hs) * per + 6 * njk + njd; };
    auto WTB = [&](int b, int hs) { return work.d + ((size_t)b * 2 + hs) * per + 6 * njk + 2 * njd; };
    auto WUZ = [&](int b, int hs) { return work.d + ((size_t)b * 2 + hs) * per + 6 * njk + 2 * njd + kmax; };
    {
        LevelArgs la;
        la.kind = P->kind; la.kmax = kmax; la.nlat = nlat; la.nd = nd; la.nb = nb;
        la.a = P->a; la.dphi = P->dphi; la.pi = pi; la.area_scale = area_scale; la.cbound = cbound;
        la.zm = zmd; la.hist = reinterpret_cast<const long long*>(hist.d); la.alat = d + P->o_alat; la.cosmid = d + P->o_cosmid; la.norm = d + P->o_norm;
        la.cosw = (P->kind == 0) ? d + P->o_cosw : nullptr;
        la.ext = extd; la.work = work.d; la.ckref = work.d + ckoff; la.per = per; la.njk = njk; la.njd = njd;
        const size_t smem = sizeof(double) * ((size_t)4 * nb + 3 * nd);
        if (smem > 48 * 1024)
            FALWA_CUDA_TRY(cudaFuncSetAttribute(refstate_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { FALWA_KERNEL("refstate_level_kernel", st); refstate_level_kernel<<<dim3(kmax, 2, batch), 256, smem, st>>>(la); }
        FALWA_LAUNCH_CHECK();
        if (P->kind == 0) {
            for (int b = 0; b < batch; ++b)
                for (int hs = 0; hs < 2; ++hs) {
                    { FALWA_KERNEL("sor_uz_kernel", st); sor_uz_kernel<<<ceil_div(nd, 128), 128, 0, st>>>(nd, kmax, d + P->o_uzc1, d + P->o_uzc2, W(b, hs, 4), WUZ(b, hs)); }
                    FALWA_LAUNCH_CHECK();
                }
        }
    }
    // ---- batched solves --------------------------------------------------------------------------
    const int nprob = batch * 2;
    if (P->kind == 0) {
        std::vector<SorArgs> probs(nprob);
        for (int b = 0; b < batch; ++b)
            for (int hs = 0; hs < 2; ++hs) {
                SorArgs& p = probs[(size_t)b * 2 + hs];
                p.jd = nd; p.kmax = kmax; p.maxits = P->maxit;
                p.om = P->omega; p.a = P->a; p.dphi = P->dphi; p.dz = P->dz; p.eps = P->tol; p.rjac = P->rjac; p.pi = pi;
                p.h = P->h; p.r = P->rr;
                p.ajk = d + P->o_ajk; p.bjk = d + P->o_bjk; p.fact = d + P->o_fact; p.uz = WUZ(b, hs);
                p.cosj = d + P->o_cosj; p.cor = d + P->o_cor;
                p.amp = HT.d + o_amp[(size_t)b * 2 + hs]; p.amm = p.amp + kmax; p.ez = d + P->o_ez; p.eref = d + P->o_eref;
                p.qref = W(b, hs, 0); p.ubar = W(b, hs, 3); p.tbar = W(b, hs, 4); p.cref = W(b, hs, 1); p.cbar = W(b, hs, 2);
                p.tb = WTB(b, hs); p.u = WU(b, hs); p.tref = WT(b, hs); p.fjk = W(b, hs, 5);
                p.num_of_iter = nit.d + (size_t)b * 2 + hs;
                p.u_in_smem = (sizeof(double) * njk <= 200 * 1024);
            }
        DeviceScratch<SorArgs> dp;
        if ((rc = dp.alloc(nprob, st))) return rc;
        if ((rc = upload_bytes(dp.d, probs.data(), sizeof(SorArgs) * nprob, st))) return rc;
        FALWA_CUDA_TRY(cudaFuncSetAttribute(sor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        const size_t smem = probs[0].u_in_smem ? sizeof(double) * njk : 0;
        { FALWA_KERNEL("sor_kernel", st); sor_kernel<<<nprob, 1024, smem, st>>>(dp.d); }
        FALWA_LAUNCH_CHECK();
    } else {
        const int n = jd - 2, m = kmax - 2;
        DeviceScratch<double> sw;
        const size_t per_s = 2 * (size_t)m * n + (size_t)n * m + (size_t)kmax * n + njk;  // g, gh, cp, pjk, qref_out
        if ((rc = sw.alloc((size_t)nprob * per_s, st))) return rc;
        std::vector<SpectralProblem> sp(nprob);
        std::vector<SweepArgs> wp(nprob);
        for (int b = 0; b < batch; ++b)
            for (int hs = 0; hs < 2; ++hs) {
                const size_t q = (size_t)b * 2 + hs;
                double* base = sw.d + q * per_s;
                SpectralProblem& s = sp[q];
                s.qref = W(b, hs, 0); s.ckref = work.d + ckoff + (size_t)b * njk; s.tbar = W(b, hs, 4);
                s.spec = HT.d + o_spec[q];
                s.g = base; s.gh = base + (size_t)m * n; s.cp = base + 2 * (size_t)m * n; s.pjk = base + 3 * (size_t)m * n;
                SweepArgs& w = wp[q];
                w.n = n; w.jd = jd; w.nd = nd; w.kmax = kmax; w.jb = jb;
                w.a = P->a; w.om = P->omega; w.dz = P->dz; w.h = P->h; w.rr = P->rr; w.dp = pi / (double)(nlat - 1); w.pi = pi;
                w.sjk = nullptr; w.tjk = nullptr; w.ckref = s.ckref;
                w.tb = HT.d + o_prof + ((size_t)b * 4 + 1) * kmax;  // tn0 for both hemispheres
                w.qoc = W(b, hs, 0);
                w.cosrow = d + P->o_cosrow; w.sinrow = d + P->o_sinrow; w.sinnd = d + P->o_sinnd; w.eref = d + P->o_eref;
                w.tref = WT(b, hs); w.qref = base + 3 * (size_t)m * n + (size_t)kmax * n; w.u = WU(b, hs); w.pjk = s.pjk;
                w.u1top = 0.0; w.u1off = P->omega * P->a * std::cos(w.dp * (double)jb);
            }
        SpectralCommon c;
        c.n = n; c.m = m; c.nd = nd; c.kmax = kmax; c.jb = jb;
        c.ajk = d + P->o_najk; c.bjk = d + P->o_nbjk; c.fact = d + P->o_nfact; c.topcoef = d + P->o_topcoef;
        c.topden = d + P->o_topden;
        c.fcoef = -0.5 * P->a * (pi / (double)(nlat - 1));
        c.u1coef = 2. * pi * P->a;
        c.u1off = P->omega * P->a * std::cos((pi / (double)(nlat - 1)) * (double)jb);
        DeviceScratch<SpectralProblem> dsp;
        DeviceScratch<SweepArgs> dwp;
        if ((rc = dsp.alloc(nprob, st))) return rc;
        if ((rc = dwp.alloc(nprob, st))) return rc;
        if ((rc = upload_bytes(dsp.d, sp.data(), sizeof(SpectralProblem) * nprob, st))) return rc;
        if ((rc = upload_bytes(dwp.d, wp.data(), sizeof(SweepArgs) * nprob, st))) return rc;
        { FALWA_KERNEL("nhn22_spectral_solve_kernel", st); nhn22_spectral_solve_kernel<<<nprob, 512, 0, st>>>(c, dsp.d); }
        FALWA_LAUNCH_CHECK();
        { FALWA_KERNEL("sweep_post_kernel", st); sweep_post_kernel<<<nprob, 512, 0, st>>>(dwp.d); }
        FALWA_LAUNCH_CHECK();
        // dsp/dwp/sw are stream-ordered: freed after the kernels that use them
        { FALWA_KERNEL("merge_refstates_kernel", st); merge_refstates_kernel<<<dim3(ceil_div((long long)kmax * nlat, 256), batch), 256, 0, st>>>(
            kmax, nlat, nd, jd, north_only, 2. * P->omega, P->omega, d + P->o_sinlat, work.d, per, njk, njd, qref_st, qref_cu,
            uref_st, ptref_st); }
        FALWA_LAUNCH_CHECK();
    }
    if (P->kind == 0) {
        { FALWA_KERNEL("merge_refstates_kernel", st); merge_refstates_kernel<<<dim3(ceil_div((long long)kmax * nlat, 256), batch), 256, 0, st>>>(
            kmax, nlat, nd, nd, north_only, 1.0, P->omega, d + P->o_sinlat, work.d, per, njk, njd, qref_st, qref_cu, uref_st,
            ptref_st); }
        FALWA_LAUNCH_CHECK();
    }
    if (P->kind != 0) {
        // the direct solve has no iteration count: nothing to wait for, the host runs ahead of the device
        std::fill(num_of_iter_host, num_of_iter_host + nprob, 0);
        return 0;
    }
    // NH18: the caller's convergence check needs the SOR iteration counts now.  They come back through mapped host
    // memory (not a copy engine, which may be busy with another stream's bulk transfer for milliseconds).
    void* stage = nullptr;
    {
        PinnedRing& R = pinned_ring();
        std::lock_guard<std::mutex> lock(R.mu);
        size_t off = 0;
        stage = R.take(sizeof(int) * nprob, &off);
    }
    if (stage) {
        FALWA_KERNEL("pull_words_kernel", st);
        pull_words_kernel<<<1, 256, 0, st>>>((unsigned long long*)stage, (const unsigned long long*)nit.d, (size_t)batch);
        FALWA_LAUNCH_CHECK();
        FALWA_CUDA_TRY(cudaStreamSynchronize(st));
        std::memcpy(num_of_iter_host, stage, sizeof(int) * nprob);
    } else {
        FALWA_CUDA_TRY(cudaMemcpyAsync(num_of_iter_host, nit.d, sizeof(int) * nprob, cudaMemcpyDeviceToHost, st));
        FALWA_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

// LWA and barotropic fluxes of both hemispheres.
//   lwa [batch][kmax][nlat][nlon]; out2d [batch][15][nlat][nlon] (slot order: enum Out2D / falwa_b200.h)
//   ncforce may be NULL (treated as zeros).  method: 0 = automatic, 1 = brute force.
static int lwa_fluxes_impl(const falwa_plan* P, int batch, const double* qgpv, const double* u, const double* v,
                           const double* theta, const double* ncforce, const double* profiles_host,
                           const double* qref_cu, const double* uref_st, const double* ptref_st, int north_only,
                           double* lwa, double* out2d, int method, double* l3d, cudaStream_t st) {
    const double* d = P->d_tab;
    const double pi = host_pi(), dp = pi / (double)(P->nlat - 1);
    const int nlon = P->nlon, nlat = P->nlat, kmax = P->kmax, nd = P->nd, jb = P->jb;
    const int nhemi = north_only ? 1 : 2;
    const size_t f3 = (size_t)kmax * nlat * nlon;
    int rc;
    DeviceScratch<double> tmp, hsum, lsel;
    DeviceTable prof;
    {   // profiles [batch][4][kmax] followed by the per-level factor of the ep1 heat term, [batch][2][kmax]
        std::vector<double> hp(profiles_host, profiles_host + (size_t)batch * 4 * kmax);
        const double rkappa = P->rr / P->cp;
        hp.resize((size_t)batch * 6 * kmax, 0.0);
        for (int b = 0; b < batch; ++b)
            for (int hs = 0; hs < 2; ++hs) {  // slot 0: south (ts0), slot 1: north (tn0)
                const double* tg = profiles_host + ((size_t)b * 4 + hs) * kmax;
                double* f = hp.data() + (size_t)batch * 4 * kmax + ((size_t)b * 2 + hs) * kmax;
                for (int k = 2; k <= kmax - 1; ++k)
                    f[k - 1] = (P->rr / P->h) * std::exp(-rkappa * (P->dz * (double)(k - 1)) / P->h) * 2. * P->dz / (tg[k] - tg[k - 2]);
            }
        if ((rc = prof.upload(hp, st))) return rc;
    }
    FusedFluxArgs p;
    p.imax = nlon; p.nlat = nlat; p.kmax = kmax; p.nd = nd; p.jb = jb; p.jd = P->jd; p.nhemi = nhemi;
    p.has_ncforce = ncforce != nullptr;
    p.a = P->a; p.om = P->omega; p.dz = P->dz; p.prefac = P->prefactor; p.ab = P->a * dp; p.rrh = P->rr / P->h;
    p.ep42fac = std::exp(-P->dz / P->h); p.dc = P->dz / P->prefactor;
    p.cosjb = std::cos(dp * (double)jb); p.cosjbm1 = std::cos(dp * (double)(jb - 1));
    p.pv = qgpv; p.uu = u; p.vv = v; p.pt = theta; p.ncforce = ncforce;
    p.qref_cu = qref_cu; p.uref_st = uref_st; p.tref_st = ptref_st; p.prof = prof.d;
    p.cosphi = d + P->o_cosphi; p.sinphi = d + P->o_sinphi; p.aaw = d + P->o_aaw; p.ep11a = d + P->o_ep11a; p.vw = d + P->o_vw;
    p.clat = d + P->o_clat;
    p.sa1 = p.sa2 = p.su2 = p.snc = nullptr;
    p.f11 = prof.d + (size_t)batch * 4 * kmax;
    p.lwa = lwa; p.out2d = out2d; p.l3d = l3d;
    if (north_only) {  // rows the northern launch does not cover stay zero
        FALWA_CUDA_TRY(cudaMemsetAsync(lwa, 0, sizeof(double) * (size_t)batch * f3, st));
        FALWA_CUDA_TRY(cudaMemsetAsync(out2d, 0, sizeof(double) * (size_t)batch * O_COUNT * nlat * nlon, st));
    }
    // method: 0 = automatic (windowed sums, else interval scan, else the double loop), 1 = double loop,
    //         2 = interval scan, 3 = windowed sums
    const bool use_window = (method == 0 || method == 3 || method == 4) && !l3d && win_supported(nlon, nd, qgpv, u, ncforce);
    // 0: per (snapshot, hemisphere) the windowed sums or the prefix sums, by the sampled displacement of its PV contours;
    // 3 / 4: one of them everywhere (FALWA_LWA_FORCE=0|1 does the same for method 0)
    static const int lwa_force = tune_int("FALWA_LWA_FORCE", -1);       // -1 automatic, 0 window, 1 prefix
    static const int lwa_rows = tune_int("FALWA_LWA_SWITCH_ROWS", 10);  // mean displacement (rows) above which prefix sums win
    const int pfx_mode = !pfx_supported(nlon, nd, qgpv, u, ncforce) ? 0 : (method == 3 ? 0 : (method == 4 ? 1 : (lwa_force >= 0 ? lwa_force : -1)));
    const bool use_scan = !use_window && (method != 1) && scan_supported(nd);
    ScanTables tabs;
    if (use_window) {
        // 3-D LWA and the vertical sums of astar1, astar2, ua1, ua2 (, ncforce) in one walk over the levels
        WinArgs w{};
        w.imax = nlon; w.nd = nd; w.kmax = kmax; w.jstart = jb + 1; w.jend = nd - 1; w.nhemi = nhemi;
        w.k_first = 2; w.k_count = kmax - 2;
        for (int h = 0; h < 2; ++h) {  // hemispheric row jj -> global row nd-2+jj (north) / nd-jj (south)
            w.fr[h].in_row0 = w.fr[h].out_row0 = h ? 0 : nd - 1;
            w.fr[h].rev = h;
            w.fr[h].skip_first = (nhemi == 2 && h == 0) ? 1 : 0;
            w.fr[h].sg = h ? -1.0 : 1.0;
            w.q_row0[h] = w.u_row0[h] = nd - 1;
            w.q_rstep[h] = w.u_rstep[h] = h ? -1 : 1;
            w.q_sg[h] = h ? -1.0 : 1.0;
        }
        w.qref = qref_cu; w.q_batch = (size_t)kmax * nlat; w.q_kstride = nlat;
        w.uref = uref_st; w.u_batch = (size_t)kmax * nlat; w.u_kstride = nlat;
        w.cosphi = d + P->o_cosphi; w.ab = p.ab;
        w.lwa = lwa; w.out2d = out2d; w.vw = d + P->o_vw; w.dc = p.dc;
        w.lwa_batch = f3; w.lwa_plane = (size_t)nlat * nlon; w.out_pitch = nlon;
        w.o2_batch = (size_t)O_COUNT * nlat * nlon; w.o2_plane = (size_t)nlat * nlon;
        w.slot_a1 = O_ASTAR1; w.slot_a2 = O_ASTAR2; w.slot_lwa = O_LWA; w.slot_ua1 = O_UA1; w.slot_ua2 = O_UA2;
        w.slot_ncf = O_NCF;
        w.mask_count = nullptr;
        const int nprob = batch * nhemi;
        if (pfx_mode == 0) {
            if ((rc = lwa_window_launch(w, 1, qgpv, u, ncforce, nlat, batch, nprob, st))) return rc;
        } else if (pfx_mode == 1) {
            if ((rc = lwa_prefix_launch(w, 1, qgpv, u, nlat, batch, nprob, nullptr, 0, st))) return rc;
        } else {
            if ((rc = lsel.alloc((size_t)3 * nprob, st))) return rc;
            unsigned long long* acc = reinterpret_cast<unsigned long long*>(lsel.d);
            int* sel = reinterpret_cast<int*>(lsel.d + 2 * nprob);
            FALWA_CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(unsigned long long) * 2 * nprob, st));
            PfxProbeArgs pa{};
            pa.imax = nlon; pa.nd = nd; pa.kmax = kmax; pa.jstart = w.jstart; pa.jend = w.jend; pa.nhemi = nhemi; pa.in_rows = nlat;
            pa.pv = qgpv; pa.in_batch = f3; pa.in_plane = (size_t)nlat * nlon;
            pa.qref = w.qref; pa.q_batch = w.q_batch; pa.q_kstride = w.q_kstride;
            for (int h = 0; h < 2; ++h) {
                pa.in_row0[h] = nd - 1; pa.in_rstep[h] = h ? -1 : 1; pa.sg[h] = h ? -1.0 : 1.0;
                pa.q_row0[h] = w.q_row0[h]; pa.q_rstep[h] = w.q_rstep[h]; pa.q_sg[h] = w.q_sg[h];
            }
            pa.acc = acc;
            { FALWA_KERNEL("lwa_probe_kernel", st); lwa_probe_kernel<<<dim3(8 * ceil_div(kmax - 2, 8), nprob), 256, 0, st>>>(pa);
              lwa_select_kernel<<<ceil_div(nprob, 128), 128, 0, st>>>(acc, sel, nprob, (double)lwa_rows, -1); }
            FALWA_LAUNCH_CHECK();
            if ((rc = lwa_window_launch(w, 1, qgpv, u, ncforce, nlat, batch, nprob, st, sel, 0))) return rc;
            if ((rc = lwa_prefix_launch(w, 1, qgpv, u, nlat, batch, nprob, sel, 1, st))) return rc;
        }
    } else if (use_scan) {
        // LWA pieces a1, a2, ua2 (, ncforce3d) of every level by the interval scan, in the global frame
        const int narr = ncforce ? 4 : 3, nprob = batch * nhemi;
        if ((rc = tmp.alloc((size_t)narr * batch * f3, st))) return rc;
        if ((rc = tabs.alloc(nprob, nd, kmax, st))) return rc;
        ThrArgs t{};
        ScanArgs s{};
        t.nd = nd; t.kmax = kmax; t.jstart = jb + 1; t.jend = nd - 1; t.nhemi = nhemi;
        t.qref = qref_cu; t.q_batch = (size_t)kmax * nlat; t.q_kstride = nlat;
        t.uref = uref_st; t.u_batch = (size_t)kmax * nlat; t.u_kstride = nlat;
        for (int h = 0; h < 2; ++h) {  // hemispheric row jj -> global row nd-2+jj (north) / nd-jj (south)
            t.q_row0[h] = t.u_row0[h] = nd - 1;
            t.q_rstep[h] = t.u_rstep[h] = h ? -1 : 1;
            t.q_sg[h] = h ? -1.0 : 1.0;
            s.fr[h].in_row0 = s.fr[h].out_row0 = nd - 1;
            s.fr[h].in_rstep = s.fr[h].out_rstep = h ? -1 : 1;
            s.fr[h].sg = h ? -1.0 : 1.0;
            s.fr[h].skip_first_row = (nhemi == 2 && h == 0) ? 1 : 0;
        }
        s.imax = nlon; s.nd = nd; s.kmax = kmax; s.jstart = t.jstart; s.jend = t.jend; s.nhemi = nhemi;
        s.pv = qgpv; s.uu = u; s.nc = ncforce; s.in_batch = f3; s.in_plane = (size_t)nlat * nlon;
        s.a1 = tmp.d; s.a2 = tmp.d + (size_t)batch * f3; s.u2 = tmp.d + (size_t)2 * batch * f3;
        s.ncf = ncforce ? tmp.d + (size_t)3 * batch * f3 : nullptr;
        s.out_batch = f3; s.out_plane = (size_t)nlat * nlon; s.out_pitch = nlon;
        s.cosphi = d + P->o_cosphi; s.ab = p.ab; s.mask_count = nullptr;
        tabs.bind(t, s, nprob, nd, kmax);
        if ((rc = lwa_scan_launch(t, s, nprob, st))) return rc;
        p.sa1 = s.a1; p.sa2 = s.a2; p.su2 = s.u2; p.snc = s.ncf;
    }
    {
        dim3 block(32, 8);
        dim3 grid(ceil_div(nlon, 32), ceil_div(nd, 8), batch * nhemi);
        if (l3d) {   // layer-wise stores: the plain per-level kernel (scan pieces or the double loop inline)
            FALWA_KERNEL("fused_flux_layerwise_kernel", st);
            if (use_scan) fused_flux_kernel<false><<<grid, block, 0, st>>>(p);
            else fused_flux_kernel<true><<<grid, block, 0, st>>>(p);
        } else if (use_window) {
            if ((rc = hsum.alloc((size_t)batch * nhemi * nd * nlon, st))) return rc;
            FALWA_KERNEL("flux_uvt_column_kernel", st);
            flux_column_kernel<false, true><<<grid, block, 0, st>>>(p, hsum.d);
            flux_column_kernel<true, true><<<dim3(ceil_div(nlon, 256), 2, batch * nhemi), 256, 0, st>>>(p, hsum.d);
        } else if (use_scan) {
            if ((rc = hsum.alloc((size_t)batch * nhemi * nd * nlon, st))) return rc;
            FALWA_KERNEL("flux_column_kernel", st);
            flux_column_kernel<false, false><<<grid, block, 0, st>>>(p, hsum.d);
            flux_column_kernel<true, false><<<dim3(ceil_div(nlon, 256), 2, batch * nhemi), 256, 0, st>>>(p, hsum.d);
        } else {
            FALWA_KERNEL("fused_flux_bruteforce_kernel", st);
            fused_flux_kernel<true><<<grid, block, 0, st>>>(p);
        }
        FALWA_LAUNCH_CHECK();
    }
    { FALWA_KERNEL("baro_post_kernel", st); baro_post_kernel<<<dim3(ceil_div(P->nlon, 256), P->nlat, batch), 256, 0, st>>>(
        P->nlon, P->nlat, P->a, P->dphi, P->dlambda, d + P->o_clat, out2d, ((use_scan || use_window) && !l3d) ? hsum.d : nullptr, nhemi, nd, jb); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

extern "C" int falwa_b200_lwa_and_barotropic_fluxes(const falwa_plan* P, int batch, const double* qgpv,
                                                    const double* u, const double* v, const double* theta,
                                                    const double* ncforce, const double* profiles_host,
                                                    const double* qref_cu, const double* uref_st,
                                                    const double* ptref_st, int north_only, double* lwa,
                                                    double* out2d, int method, void* stream) {
    if (!P || !qgpv || !u || !v || !theta || !profiles_host || !qref_cu || !uref_st || !ptref_st || !lwa || !out2d)
        return FALWA_ERR_NULL;
    if (batch < 1) return FALWA_ERR_ARG;
    return lwa_fluxes_impl(P, batch, qgpv, u, v, theta, ncforce, profiles_host, qref_cu, uref_st, ptref_st, north_only,
                           lwa, out2d, method, nullptr, (cudaStream_t)stream);
}

namespace falwa {

// f e^{z/H} d/dz [ e^{-z/H} g / (d theta~/dz) ]  (utilities.py:305-344: centred differences, one-sided at both ends;
// hemispheric static stability, the equator row takes the average of the two).  One thread per (lon, lat) column.
//   MODE 0: g = a field given on the pressure levels (interpolated like interpolate_fields) or on the height grid:
//           the non-conservative forcing from a heating rate (oopinterface.py:985-1019), out = + the expression.
//   MODE 1: g = v (theta - theta_ref) |cos(lat)|: the layer-wise stretching term (oopinterface.py:1021-1061),
//           out = - the expression.
struct ZDerivArgs {
    int nlon, nlat, nlev, kmax, eq, interp;
    const double *gin, *vv, *pt, *tref;       // MODE 0: gin [b][nlev or kmax][nlat][nlon]; MODE 1: v, theta, tref
    const int* lo; const double *whi, *wlo;   // interpolation tables of the plan
    const double *prof;                       // [b][4][kmax]: t0_s, t0_n, stat_s, stat_n
    const double *dens, *mult, *clat;         // [kmax], [kmax][nlat], [nlat]
    double dz;
    double* out; size_t out_batch;            // out + b * out_batch: [kmax][nlat][nlon]
};

template <int MODE>
__global__ void __launch_bounds__(256)
zderiv_kernel(ZDerivArgs a) {
    const size_t plane = (size_t)a.nlon * a.nlat;
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= plane) return;
    const int b = blockIdx.y, g = (int)(col / a.nlon);
    const double* ss = a.prof + ((size_t)b * 4 + 2) * a.kmax;
    const double* sn = ss + a.kmax;
    auto stab = [&](int k) { return g < a.eq - 1 ? ss[k] : (g == a.eq - 1 ? 0.5 * (ss[k] + sn[k]) : sn[k]); };
    auto x_at = [&](int k) -> double {
        double gv;
        if (MODE == 0) {
            const double* in = a.gin + (size_t)b * plane * (a.interp ? a.nlev : a.kmax);
            if (a.interp) {
                const int l = a.lo[k];
                gv = a.whi[k] * in[(size_t)(l + 1) * plane + col] + a.wlo[k] * in[(size_t)l * plane + col];
            } else gv = in[(size_t)k * plane + col];
        } else {
            const size_t o = (size_t)b * plane * a.kmax + (size_t)k * plane + col;
            gv = a.vv[o] * (a.pt[o] - a.tref[((size_t)b * a.kmax + k) * a.nlat + g]) * a.clat[g];
        }
        return a.dens[k] * gv / stab(k);
    };
    double* out = a.out + (size_t)b * a.out_batch + col;
    const double sgn = MODE == 1 ? -1.0 : 1.0;
    const int kmax = a.kmax;
    double xm = x_at(0), x0 = x_at(1);
    out[0] = sgn * (a.mult[g] * (x0 - xm) / a.dz);
    for (int k = 1; k < kmax - 1; ++k) {
        const double xp = x_at(k + 1);
        out[(size_t)k * plane] = sgn * ((xp - xm) / (2 * a.dz) * a.mult[(size_t)k * a.nlat + g]);
        xm = x0; x0 = xp;
    }
    out[(size_t)(kmax - 1) * plane] = sgn * (a.mult[(size_t)(kmax - 1) * a.nlat + g] * (x0 - xm) / a.dz);
}

// host tables of zderiv_kernel: dens[kmax] = e^{-z/H}, mult[kmax][nlat] = 2 Omega sin(lat) e^{z/H}, then the profiles
static int zderiv_tables(const falwa_plan* P, int batch, const double* profiles_host, DeviceTable& T, ZDerivArgs& a,
                         cudaStream_t st) {
    const int kmax = P->kmax, nlat = P->nlat;
    std::vector<double> h((size_t)kmax + (size_t)kmax * nlat + (size_t)batch * 4 * kmax);
    const std::vector<double>& sinlat = P->sinlat;
    for (int k = 0; k < kmax; ++k) {
        const double z = P->height[k];
        h[k] = std::exp(-z / P->h);
        const double ez = std::exp(z / P->h);
        for (int j = 0; j < nlat; ++j) h[(size_t)kmax + (size_t)k * nlat + j] = 2 * P->omega * sinlat[j] * ez;
    }
    std::copy(profiles_host, profiles_host + (size_t)batch * 4 * kmax, h.begin() + kmax + (size_t)kmax * nlat);
    int rc = T.upload(h, st);
    if (rc) return rc;
    a.nlon = P->nlon; a.nlat = nlat; a.nlev = P->nlev; a.kmax = kmax; a.eq = nlat / 2 + 1;
    a.lo = P->d_lo; a.whi = P->d_tab + P->o_whi; a.wlo = P->d_tab + P->o_wlo; a.clat = P->d_tab + P->o_clat;
    a.dens = T.d; a.mult = T.d + kmax; a.prof = T.d + kmax + (size_t)kmax * nlat; a.dz = P->dz;
    return 0;
}

}  // namespace falwa

// compute_layerwise_lwa_fluxes of a batch (oopinterface.py:1063-1156, :1021-1061).
//   out3d [batch][10][kmax][nlat][nlon]: astar1, astar2, ncforce, ua1, ua2, ep1, ep2, ep3, stretch_term, lwa in the global
//   frame (southern hemisphere: ep2 / ep3 swapped and negated, ncforce negated; the equator row is the southern value).
//   The other arguments are those of falwa_b200_lwa_and_barotropic_fluxes.
extern "C" int falwa_b200_layerwise_lwa_fluxes(const falwa_plan* P, int batch, const double* qgpv, const double* u,
                                               const double* v, const double* theta, const double* ncforce,
                                               const double* profiles_host, const double* qref_cu,
                                               const double* uref_st, const double* ptref_st, int north_only,
                                               double* out3d, int method, void* stream) {
    if (!P || !qgpv || !u || !v || !theta || !profiles_host || !qref_cu || !uref_st || !ptref_st || !out3d)
        return FALWA_ERR_NULL;
    if (batch < 1) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)P->nlat * P->nlon, f3 = plane * P->kmax;
    int rc;
    DeviceScratch<double> lwa, out2d;
    if ((rc = lwa.alloc((size_t)batch * f3, st))) return rc;
    if ((rc = out2d.alloc((size_t)batch * O_COUNT * plane, st))) return rc;
    FALWA_CUDA_TRY(cudaMemsetAsync(out3d, 0, sizeof(double) * (size_t)batch * L3_COUNT * f3, st));
    if ((rc = lwa_fluxes_impl(P, batch, qgpv, u, v, theta, ncforce, profiles_host, qref_cu, uref_st, ptref_st,
                              north_only, lwa.d, out2d.d, method, out3d, st))) return rc;
    DeviceTable T;
    ZDerivArgs a{};
    if ((rc = zderiv_tables(P, batch, profiles_host, T, a, st))) return rc;
    a.vv = v; a.pt = theta; a.tref = ptref_st;
    a.out = out3d + (size_t)L_STRETCH * f3; a.out_batch = (size_t)L3_COUNT * f3;
    { FALWA_KERNEL("zderiv_kernel", st); zderiv_kernel<1><<<dim3(ceil_div((long long)plane, 256), batch), 256, 0, st>>>(a); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

// compute_ncforce_from_heating_rate of a batch (oopinterface.py:985-1019): heating_rate [batch][nlev][nlat][nlon] on the
// pressure levels (on_height_grid = 0; interpolated like the fields) or [batch][kmax][nlat][nlon] on the height grid;
// ncforce [batch][kmax][nlat][nlon].
extern "C" int falwa_b200_ncforce_from_heating_rate(const falwa_plan* P, int batch, const double* heating_rate,
                                                    int on_height_grid, const double* profiles_host, double* ncforce,
                                                    void* stream) {
    if (!P || !heating_rate || !profiles_host || !ncforce) return FALWA_ERR_NULL;
    if (batch < 1) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)P->nlat * P->nlon;
    DeviceTable T;
    ZDerivArgs a{};
    int rc;
    if ((rc = zderiv_tables(P, batch, profiles_host, T, a, st))) return rc;
    a.gin = heating_rate; a.interp = on_height_grid ? 0 : 1;
    a.out = ncforce; a.out_batch = plane * P->kmax;
    { FALWA_KERNEL("zderiv_kernel", st); zderiv_kernel<0><<<dim3(ceil_div((long long)plane, 256), batch), 256, 0, st>>>(a); }
    FALWA_LAUNCH_CHECK();
    return 0;
}
