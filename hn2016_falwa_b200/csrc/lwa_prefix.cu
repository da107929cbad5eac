// Local wave activity by prefix sums over threshold intervals: the companion of lwa_window.cu for strongly meandering
// flows (same sets, same inputs and outputs, cost independent of how far the PV contours are displaced).
//
// For one column (fixed longitude and level) with PV values q_x (hemispheric rows x = 0 .. nd-1, equator -> pole) and
// thresholds Q_j, the reference (compute_flux_dirinv.f90:70-116, compute_lwa_only_nhn22.f90:55-90) sums over
//     cyclonic      { x <  j : q_x >  Q_j }          anticyclonic  { x >= j : q_x <= Q_j }.
// The area analysis makes Q non-decreasing in j (up to a few rows, below), so with E(x) = #{ j : Q_j < q_x } row x
// belongs to the cyclonic set of exactly the thresholds j in [x+1, E(x)) and to the anticyclonic set of those in
// [E(x), x]: a contiguous interval of thresholds.  Every summand is f(x) g(j), so six set sums suffice
// (sum q aa, sum aa over either set; sum q u, sum u with the anticyclonic part negated), and each is the running sum
// over j of a difference array that receives +f(x) at j = x+1 (a slot of its own: a plain store) and -f(x) at j = E(x)
// (a scatter: several rows may meet there).  The scatter adds 64-bit FIXED-POINT integers with native 32-bit
// shared-memory atomics (low word + carry, as the area histograms do): integer addition is associative, so the result
// is bit-reproducible, rows outside every set cancel EXACTLY to zero, and the running sums are plain integer scans.
// The scale of each quantity is a power of two chosen per (tile, level) from the largest exponent present, leaving
// 53 bits below the largest possible term: the rounding of a term is below one ulp of that bound.
// E(x) comes from a 9-step bisection with the reference's own comparison (Q_j < q_x in fp64): set membership is exact.
// Thresholds that break monotonicity (Q_j below an earlier one; NaN) are "exception rows": the search runs on the
// running maximum of Q, and an exception row is evaluated by its warp with the reference's double loop over the column
// (the tile is still in shared memory).  NaN PV belongs to no set, as in the reference; non-finite u or +-Inf PV are not
// supported by this method (the window method is).
//
// Layout, thread -> element map, TMA tiles, tensor-memory accumulators and the output stage are those of
// lwa_window_kernel (one CTA = (snapshot, hemisphere, 8 longitudes) x all rows, walking the levels).  The six
// difference arrays of a tile (6 x 8 columns x (nd+1) x 8 bytes = 149 KB) leave room for ONE tile stage: the load of the
// next level is issued as soon as this level's values are in the arrays and lands during the scatter / scan / output.
#include "common.cuh"

namespace falwa {

constexpr int kPfxWarps = 24;
constexpr int kPfxPitch = 388;            // entries per (array, column): >= kWinMaxRows + 1, == 4 (mod 16) -> conflict-free
constexpr int kPfxArrays = 6;             // c1 c2 a1 a2 t3 t4
constexpr int kPfxQt = 512;               // running-maximum thresholds padded with +inf: bisection without bounds checks
constexpr int kPfxSpan = 13;              // entries per lane in the scans (32 * 13 = 416 >= nd + 1)
constexpr int kPfxExcp = 32 * kPfxSpan;

struct PfxSmem { size_t off_tiles, off_d, off_q, off_u, off_qt, off_cs, off_excp, off_excf, off_misc, total; };
__host__ __device__ inline PfxSmem pfx_smem_layout() {
    PfxSmem s;
    size_t o = 0;
    s.off_tiles = o; o += (size_t)2 * kWinTile * 8;                                  // q | u, one stage
    s.off_d = o; o += (size_t)kPfxArrays * kWinCols * kPfxPitch * 8;
    s.off_q = o; o += (size_t)2 * kWinMaxRows * 8;                                   // thresholds, two levels
    s.off_u = o; o += (size_t)2 * kWinMaxRows * 8;
    s.off_qt = o; o += (size_t)2 * kPfxQt * 8;
    s.off_cs = o; o += (size_t)kWinMaxRows * 8;
    s.off_excp = o; o += (size_t)2 * kPfxExcp * 4;
    s.off_excf = o; o += (size_t)2 * kPfxExcp;
    s.off_misc = o; o += 64;
    s.total = o + 128;
    return s;
}

__device__ __forceinline__ long long lds_s64(uint32_t a) { long long v; asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_s64(uint32_t a, long long v) { asm volatile("st.shared.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
// 64-bit fixed-point add with two native 32-bit shared atomics (the carry rides with the high word)
__device__ __forceinline__ void pfx_add(uint32_t a, long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)((unsigned long long)v >> 32);
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(lo) : "memory");
    const unsigned add = hi + (((old + lo) < old) ? 1u : 0u);
    if (add) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a + 4), "r"(add) : "memory");
}
__device__ __forceinline__ unsigned pfx_expo(double x) { return ((unsigned)__double2hiint(x) >> 20) & 0x7ffu; }
__device__ __forceinline__ double pfx_pow2(int field) { return __hiloint2double(min(max(field, 1), 2045) << 20, 0); }

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kPfxWarps * 32, 1)
lwa_prefix_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_u, const WinArgs p,
                  const int* __restrict__ sel, int sel_want) {
    extern __shared__ __align__(1024) unsigned char pfx_raw[];
    constexpr int NW = kPfxWarps, NT = NW * 32, GPW = kWinMaxRows / (4 * NW);
    constexpr int PD = kPfxPitch, NDM = kWinMaxRows;
    constexpr unsigned FULL = 0xffffffffu;
    const int prob = blockIdx.z;
    if (sel && sel[prob] != sel_want) return;    // the other LWA kernel owns this problem
    const int nd = p.nd;
    const PfxSmem L = pfx_smem_layout();
    double* tiles = reinterpret_cast<double*>(pfx_raw + L.off_tiles);     // [q | u][rows][8]
    double* Qh = reinterpret_cast<double*>(pfx_raw + L.off_q);            // [2][NDM] thresholds of a level
    double* Ur = reinterpret_cast<double*>(pfx_raw + L.off_u);            // [2][NDM] uref of a level
    double* Qt = reinterpret_cast<double*>(pfx_raw + L.off_qt);           // [2][512] running maximum of the thresholds
    double* abcs = reinterpret_cast<double*>(pfx_raw + L.off_cs);         // [NDM] a dphi cos(phi) of hemispheric row x
    int* excp = reinterpret_cast<int*>(pfx_raw + L.off_excp);             // [2][416] exception rows below a position
    unsigned char* excf = pfx_raw + L.off_excf;                           // [2][416] "row is an exception"
    uint64_t* bar = reinterpret_cast<uint64_t*>(pfx_raw + L.off_misc);    // "the tile has landed"
    uint32_t* tbase = reinterpret_cast<uint32_t*>(pfx_raw + L.off_misc + 8);
    unsigned* mexp = reinterpret_cast<unsigned*>(pfx_raw + L.off_misc + 16);   // [2][2] largest exponent of q, u
    int* nexc = reinterpret_cast<int*>(pfx_raw + L.off_misc + 32);             // [2]
    const uint32_t a_d = win_smem(pfx_raw + L.off_d);

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int c = lane & 7, r = lane >> 3;
    const int b = prob / p.nhemi, h = prob % p.nhemi;
    const WinFrame fr = p.fr[h];
    const int i0 = blockIdx.x * kWinCols;
    const bool colok = (i0 + c) < p.imax;
    const int kbeg = p.k_first + blockIdx.y * p.k_per_cta;
    const int kend = min(kbeg + p.k_per_cta, p.k_first + p.k_count);
    const int nlev = kend - kbeg;
    if (nlev <= 0) return;
    const int js0 = p.jstart - 1, je0 = p.jend - 1;                        // active thresholds, 0-based rows
    const unsigned box_bytes = (unsigned)(p.boxrows * kWinCols * 8);

    auto issue = [&](int li) {
        const int k = kbeg + li;
        win_mbar_expect(bar, box_bytes * p.nbox * 2);
        for (int bx = 0; bx < p.nbox; ++bx) {
            const int row = fr.in_row0 + bx * p.boxrows;
            double* dst = tiles + (size_t)bx * p.boxrows * kWinCols;
            win_tma_load(dst, &tm_q, bar, i0, row, k - 1, b);
            win_tma_load(dst + kWinTile, &tm_u, bar, i0, row, k - 1, b);
        }
    };
    // running maximum of the thresholds of buffer `sb`, its exception rows and their prefix count (one warp)
    auto envelope = [&](int sb) {
        const double* Qn = Qh + sb * NDM;
        double inc[kPfxSpan], prv[kPfxSpan];
        double run = -INFINITY;
#pragma unroll
        for (int t = 0; t < kPfxSpan; ++t) {
            const int m = lane * kPfxSpan + t;
            const double v = (m < js0) ? -INFINITY : ((m > je0 || m >= nd) ? INFINITY : Qn[m]);
            prv[t] = run;
            run = fmax(run, v);          // a NaN threshold leaves the running maximum alone
            inc[t] = run;
        }
        double ex = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double n = __shfl_up_sync(FULL, ex, d);
            if (lane >= d) ex = fmax(ex, n);
        }
        double off = __shfl_up_sync(FULL, ex, 1);
        if (lane == 0) off = -INFINITY;
        int cnt = 0;
        unsigned flags = 0;
#pragma unroll
        for (int t = 0; t < kPfxSpan; ++t) {
            const int m = lane * kPfxSpan + t;
            Qt[sb * kPfxQt + m] = fmax(off, inc[t]);
            const bool e = (m >= js0 && m <= je0 && m < nd) && !(Qn[m] >= fmax(off, prv[t]));
            if (e) { flags |= 1u << t; ++cnt; }
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += n;
        }
        int below = incl - cnt;
#pragma unroll
        for (int t = 0; t < kPfxSpan; ++t) {
            const int m = lane * kPfxSpan + t;
            excp[sb * kPfxExcp + m] = below;
            excf[sb * kPfxExcp + m] = (unsigned char)((flags >> t) & 1u);
            below += (flags >> t) & 1u;
        }
        if (lane == 31) nexc[sb] = incl;
    };

    const bool trow = tid < nd && tid + 1 >= p.jstart && tid + 1 <= p.jend;
    const double* qsrc = trow ? p.qref + (size_t)b * p.q_batch + p.q_row0[h] + p.q_rstep[h] * tid : nullptr;
    const double* usrc = (trow && p.uref) ? p.uref + (size_t)b * p.u_batch + p.u_row0[h] + p.u_rstep[h] * tid : nullptr;
    const double qsg = p.q_sg[h];

    if (tid == 0) {
        win_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mexp[0] = mexp[1] = mexp[2] = mexp[3] = 0u;
    }
    __syncthreads();
    if (tid == 0) issue(0);
    if (tid < NDM) {
        abcs[tid] = (tid < nd) ? p.ab * p.cosphi[tid + 1] : 0.0;
        Qh[tid] = qsrc ? __fma_rn(qsg, qsrc[(size_t)(kbeg - 1) * p.q_kstride], 0.0) : 0.0;
        Ur[tid] = usrc ? usrc[(size_t)(kbeg - 1) * p.u_kstride] : 0.0;
        Qh[NDM + tid] = 0.0; Ur[NDM + tid] = 0.0;
    }
    for (int t = kPfxExcp + tid; t < kPfxQt; t += NT) { Qt[t] = INFINITY; Qt[kPfxQt + t] = INFINITY; }
    if (MODE == 1) {   // levels 1 and kmax of the 3-D LWA are zero (compute_flux_dirinv.f90 touches k = 2 .. kmax-1)
        for (int t = tid; t < nd * kWinCols; t += NT) {
            const int m = t >> 3, cc = t & 7;
            const int j0 = fr.rev ? nd - 1 - m : m;
            if (i0 + cc < p.imax && !(fr.skip_first && j0 == 0)) {
                const size_t o = (size_t)b * p.lwa_batch + (size_t)(fr.out_row0 + m) * p.out_pitch + i0 + cc;
                p.lwa[o] = 0.0;
                p.lwa[o + (size_t)(p.kmax - 1) * p.lwa_plane] = 0.0;
            }
        }
    }
    constexpr int TCOL = 8;
    constexpr unsigned TMEM_COLS = 256;     // (NW / 4) * GPW * TCOL = 192
    uint32_t tacc = 0;
    if (MODE == 1) {
        if (w == 0) tmem_alloc(tbase, TMEM_COLS);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (MODE == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tacc = *tbase + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * GPW * TCOL);
        const double zero[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) tmem_st8(tacc + gi * TCOL, zero);
        tmem_wait_st();
    }
    if (w == 0) envelope(0);     // thresholds of the first level are in place (barrier above)
    unsigned cnt_c = 0, cnt_a = 0;

    const int j00 = 4 * w + r;
    unsigned validm = 0, actm = 0, storem = 0;
#pragma unroll
    for (int gi = 0; gi < GPW; ++gi) {
        const int j0 = j00 + 4 * NW * gi;
        const bool valid = j0 < nd;
        if (valid) validm |= 1u << gi;
        if (valid && colok && j0 >= js0 && j0 <= je0) actm |= 1u << gi;
        if (valid && colok && !(fr.skip_first && j0 == 0)) storem |= 1u << gi;
    }
    const int m00 = fr.rev ? nd - 1 - j00 : j00;
    const int mstep = fr.rev ? -4 * NW : 4 * NW;
    const uint32_t a_t0 = win_smem(tiles) + (uint32_t)((m00 * kWinCols + c) * 8);     // q of this lane's element, group 0
    const int gbytes = mstep * kWinCols * 8;
    const uint32_t a_dc = a_d + (uint32_t)(c * PD * 8);                                // this lane's column in array 0
    constexpr uint32_t ARR = kWinCols * PD * 8;                                        // bytes between arrays
    double* out_k = (MODE == 1 ? p.lwa + (size_t)b * p.lwa_batch + (size_t)(kbeg - 1) * p.lwa_plane
                               : p.a1 + (size_t)b * p.out_batch + (size_t)(kbeg - 1) * p.out_plane) +
                    (size_t)(fr.out_row0 + m00) * p.out_pitch + i0 + c;
    const long long out_lev = (long long)((MODE == 1) ? p.lwa_plane : p.out_plane);
    const long long out_gstep = (long long)mstep * p.out_pitch;
    const double sgn = fr.sg;
    const int bab = (int)pfx_expo(p.ab);
    __syncthreads();             // envelope of the first level visible

    double wd = (MODE == 1) ? p.vw[kbeg - 1] * p.dc : 0.0;
    for (int li = 0; li < nlev; ++li) {
        const int s = li & 1, k = kbeg + li;
        const bool more = li + 1 < nlev;
        double Qn = 0.0, Un = 0.0, wdn = 0.0;
        if (more && qsrc) Qn = qsrc[(size_t)k * p.q_kstride];
        if (more && usrc) Un = usrc[(size_t)k * p.u_kstride];
        if (MODE == 1 && more) wdn = p.vw[k];
        win_mbar_wait(bar, (unsigned)(li & 1));

        // ---- A0: this thread's elements in the hemispheric frame; largest exponents of the tile -----------------
        double q[GPW], u[GPW];
        {
            unsigned mq = 0, mu = 0;
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) {
                q[gi] = 0.0; u[gi] = 0.0;
                if ((validm >> gi) & 1) {
                    const uint32_t a = a_t0 + gi * gbytes;
                    q[gi] = __fma_rn(sgn, lds64(a), 0.0);
                    if (sgn != 1.0) sts64(a, q[gi]);                 // the exception rows read the tile in this frame
                    u[gi] = lds64(a + kWinTile * 8);
                    const unsigned eq = pfx_expo(q[gi]), eu = pfx_expo(u[gi]);
                    if (eq != 0x7ffu) mq = max(mq, eq);
                    if (eu != 0x7ffu) mu = max(mu, eu);
                }
            }
            mq = __reduce_max_sync(FULL, mq);
            mu = __reduce_max_sync(FULL, mu);
            if (lane == 0) { atomicMax(&mexp[s * 2], mq); atomicMax(&mexp[s * 2 + 1], mu); }
        }
        __syncthreads();                                                                                   // (1)
        const int bq = (int)mexp[s * 2], bu = (int)mexp[s * 2 + 1];
        const int f1 = 3120 - bq - bab, f2 = 2098 - bab, f3 = 3120 - bq - bu, f4 = 2098 - bu;
        if (tid == 0) { mexp[(s ^ 1) * 2] = 0u; mexp[(s ^ 1) * 2 + 1] = 0u; }
        if (more && tid < nd) {    // next level's thresholds (their last readers finished before barrier 1)
            Qh[(s ^ 1) * NDM + tid] = __fma_rn(qsg, Qn, 0.0);
            Ur[(s ^ 1) * NDM + tid] = Un;
        }

        // ---- A1: E(x) by bisection, fixed-point terms, the slots of this row at position x + 1 ---------------------
        int E[GPW];
        long long v1[GPW], v2[GPW], v3[GPW], v4[GPW];
        {
            const double* qt = Qt + s * kPfxQt;
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) E[gi] = 0;
#pragma unroll
            for (int step = 256; step >= 1; step >>= 1) {
#pragma unroll
                for (int gi = 0; gi < GPW; ++gi)
                    if (qt[E[gi] + step - 1] < q[gi]) E[gi] += step;
            }
            const double s1 = pfx_pow2(f1), s2 = pfx_pow2(f2), s3 = pfx_pow2(f3), s4 = pfx_pow2(f4);
            if (tid < kPfxArrays * kWinCols) sts_s64(a_d + (uint32_t)(tid * PD * 8), 0ll);                 // position 0
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) {
                if ((validm >> gi) & 1) {
                    const int x = j00 + 4 * NW * gi;
                    if (!(q[gi] == q[gi])) E[gi] = x + 1;             // NaN PV: in no set
                    const bool cyc = E[gi] > x + 1, anti = E[gi] <= x, disp = cyc || anti;
                    const double av = abcs[x];
                    v1[gi] = disp ? __double2ll_rn((q[gi] * av) * s1) : 0ll;
                    v2[gi] = disp ? __double2ll_rn(av * s2) : 0ll;
                    v3[gi] = disp ? __double2ll_rn((q[gi] * u[gi]) * s3) : 0ll;
                    v4[gi] = disp ? __double2ll_rn(u[gi] * s4) : 0ll;
                    const uint32_t a = a_dc + (uint32_t)((x + 1) * 8);
                    sts_s64(a, cyc ? v1[gi] : 0ll);
                    sts_s64(a + ARR, cyc ? v2[gi] : 0ll);
                    sts_s64(a + 2 * ARR, anti ? -v1[gi] : 0ll);
                    sts_s64(a + 3 * ARR, anti ? -v2[gi] : 0ll);
                    sts_s64(a + 4 * ARR, v3[gi]);
                    sts_s64(a + 5 * ARR, v4[gi]);
                    if (COUNT && colok) {   // pairs with the kept thresholds of the interval, clipped to the active rows
                        const int* ep = excp + s * kPfxExcp;
                        if (cyc) {
                            const int lo = max(x + 1, js0), hi = min(E[gi], je0 + 1);
                            if (hi > lo) cnt_c += (unsigned)((hi - lo) - (ep[hi] - ep[lo]));
                        } else if (anti) {
                            const int lo = max(E[gi], js0), hi = min(x + 1, je0 + 1);
                            if (hi > lo) cnt_a += (unsigned)((hi - lo) - (ep[hi] - ep[lo]));
                        }
                    }
                } else {
                    E[gi] = -1; v1[gi] = v2[gi] = v3[gi] = v4[gi] = 0ll;
                }
            }
        }
        const bool hasexc = nexc[s] > 0;     // uniform: the tile must outlive this level's output stage
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                                                                   // (2)
        if (tid == 0 && more && !hasexc) issue(li + 1);

        // ---- B: the other end of every interval: -f at position E(x), fixed-point atomics -----------------------
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            const int x = j00 + 4 * NW * gi;
            if (E[gi] >= 0 && E[gi] != x + 1) {
                const uint32_t a = a_dc + (uint32_t)(E[gi] * 8);
                if (E[gi] > x + 1) { pfx_add(a, -v1[gi]); pfx_add(a + ARR, -v2[gi]); }
                else { pfx_add(a + 2 * ARR, v1[gi]); pfx_add(a + 3 * ARR, v2[gi]); }
                pfx_add(a + 4 * ARR, -v3[gi]);
                pfx_add(a + 5 * ARR, -v4[gi]);
            }
        }
        if (w == 0 && more) envelope(s ^ 1);
        __syncthreads();                                                                                   // (3)

        // ---- C: running sums over the threshold positions 0 .. nd-1, one warp per (array, column) ------------------
#pragma unroll 1
        for (int pi = w; pi < kPfxArrays * kWinCols; pi += NW) {
            const uint32_t base = a_d + (uint32_t)(pi * PD * 8) + (uint32_t)(lane * kPfxSpan * 8);
            long long v[kPfxSpan];
            long long run = 0;
#pragma unroll
            for (int t = 0; t < kPfxSpan; ++t) {
                const int m = lane * kPfxSpan + t;
                if (m < nd) run += lds_s64(base + t * 8);
                v[t] = run;
            }
            long long ex = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long n = __shfl_up_sync(FULL, ex, d);
                if (lane >= d) ex += n;
            }
            const long long off = ex - run;
#pragma unroll
            for (int t = 0; t < kPfxSpan; ++t) {
                const int m = lane * kPfxSpan + t;
                if (m < nd) sts_s64(base + t * 8, v[t] + off);
            }
        }
        __syncthreads();                                                                                   // (4)

        // ---- D: the sums of every threshold row, exception rows by the double loop, output ------------------------
        const double i1 = pfx_pow2(2046 - min(max(f1, 1), 2045)), i2 = pfx_pow2(2046 - min(max(f2, 1), 2045)),
                     i3 = pfx_pow2(2046 - min(max(f3, 1), 2045)), i4 = pfx_pow2(2046 - min(max(f4, 1), 2045));
        double* o = out_k;
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            const int j0 = j00 + 4 * NW * gi;
            if (j0 - r < nd) {   // warp-uniform: the group exists
                const bool act = (actm >> gi) & 1;
                const bool valid = (validm >> gi) & 1;
                const int jc = valid ? j0 : 0;
                const double Qj = Qh[s * NDM + jc];
                double A1 = 0.0, A2 = 0.0, B = 0.0;
                if (act) {
                    const uint32_t a = a_dc + (uint32_t)(j0 * 8);
                    const double S1c = (double)lds_s64(a) * i1, S2c = (double)lds_s64(a + ARR) * i2;
                    const double S1a = (double)lds_s64(a + 2 * ARR) * i1, S2a = (double)lds_s64(a + 3 * ARR) * i2;
                    const double T3 = (double)lds_s64(a + 4 * ARR) * i3, T4 = (double)lds_s64(a + 5 * ARR) * i4;
                    A1 = __fma_rn(-Qj, S2c, S1c);
                    A2 = __fma_rn(Qj, S2a, -S1a);
                    B = __fma_rn(-Qj, T4, T3);
                }
                if (hasexc) {
                    const bool ex = valid && excf[s * kPfxExcp + jc];
                    const unsigned em = __ballot_sync(FULL, ex);
                    if (em) {
#pragma unroll 1
                        for (int rr = 0; rr < 4; ++rr) {
                            if (!((em >> (8 * rr)) & 0xffu)) continue;
                            // all 32 lanes: column c, quarter r of the rows, threshold row je
                            const int je = 4 * w + rr + 4 * NW * gi;
                            const double Qe = Qh[s * NDM + je];
                            const int per = (nd + 3) >> 2, x0 = r * per, x1 = min(nd, x0 + per);
                            double e1 = 0.0, e2 = 0.0, eb = 0.0;
                            unsigned ec = 0, ea = 0;
                            for (int x = x0; x < x1; ++x) {
                                const int m = fr.rev ? nd - 1 - x : x;
                                const double qv = tiles[m * kWinCols + c], uv = tiles[kWinTile + m * kWinCols + c];
                                const double d = qv - Qe, av = abcs[x];
                                if (d > 0.0 && x < je) { e1 += d * av; eb += d * uv; ++ec; }
                                if (d <= 0.0 && x >= je) { e2 -= d * av; eb -= d * uv; ++ea; }
                            }
#pragma unroll
                            for (int sh = 8; sh <= 16; sh <<= 1) {
                                e1 += __shfl_xor_sync(FULL, e1, sh);
                                e2 += __shfl_xor_sync(FULL, e2, sh);
                                eb += __shfl_xor_sync(FULL, eb, sh);
                                ec += __shfl_xor_sync(FULL, ec, sh);
                                ea += __shfl_xor_sync(FULL, ea, sh);
                            }
                            if (r == rr && act) {
                                A1 = e1; A2 = e2; B = eb;
                                if (COUNT) { cnt_c += ec; cnt_a += ea; }
                            }
                        }
                    }
                }
                {
                    const double lw = A1 + A2;
                    const double x1 = Ur[s * NDM + jc] * lw;
                    const double u2 = __fma_rn(abcs[jc], B, -x1);
                    const bool store = (storem >> gi) & 1;
                    if (MODE == 1) {
                        double a[4];
                        tmem_ld8(tacc + gi * TCOL, a);
                        a[0] = __fma_rn(A1, wd, a[0]);
                        a[1] = __fma_rn(A2, wd, a[1]);
                        a[2] = __fma_rn(x1, wd, a[2]);
                        a[3] = __fma_rn(u2, wd, a[3]);
                        tmem_st8(tacc + gi * TCOL, a);
                        if (store) st_stream(o, fabs(lw));
                    } else if (store) {
                        o[0] = A1;
                        o[p.a2 - p.a1] = A2;
                        if (p.u2) o[p.u2 - p.a1] = u2;
                    }
                    o += out_gstep;
                }
            }
        }
        out_k += out_lev;
        wd = wdn * p.dc;
        if (MODE == 1) tmem_wait_st();
        if (hasexc && more) {    // the tile was needed until here
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) issue(li + 1);
        }
    }

    if (MODE == 1) {
        double* o2 = p.out2d + (size_t)b * p.o2_batch + (size_t)(fr.out_row0 + m00) * p.out_pitch + i0 + c;
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            double a[4];
            tmem_ld8(tacc + gi * TCOL, a);
            if ((storem >> gi) & 1) {
                double* o = o2 + (long long)gi * out_gstep;
                o[(size_t)p.slot_a1 * p.o2_plane] = a[0];
                o[(size_t)p.slot_a2 * p.o2_plane] = a[1];
                o[(size_t)p.slot_lwa * p.o2_plane] = a[0] + a[1];
                o[(size_t)p.slot_ua1 * p.o2_plane] = a[2];
                o[(size_t)p.slot_ua2 * p.o2_plane] = a[3];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (w == 0) tmem_dealloc(*tbase, TMEM_COLS);
    }
    if (COUNT && p.mask_count) {
        cnt_c = __reduce_add_sync(FULL, cnt_c);
        cnt_a = __reduce_add_sync(FULL, cnt_a);
        if (lane == 0 && (cnt_c | cnt_a)) {
            atomicAdd(&p.mask_count[0], (unsigned long long)cnt_a);
            atomicAdd(&p.mask_count[1], (unsigned long long)cnt_c);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Which of the two kernels a (snapshot, hemisphere) gets: the mean displacement |E(x) - (x+1)| of a sample of its
// points (every 32nd longitude, every 8th level).  The windowed sums cost ~14 warp instructions per row of the longest
// window of 32 neighbouring points, the prefix kernel a constant: 0.39 ms per 0.25-degree snapshot at 2 rows and 3.3 ms at
// 37 rows against 1.2 ms measured, i.e. they break even near 12 rows (the switch sits at 10).
// ---------------------------------------------------------------------------------------------------------------
struct PfxProbeArgs {
    int imax, nd, kmax, jstart, jend, nhemi, in_rows;
    const double* pv; size_t in_batch, in_plane;
    int in_row0[2], in_rstep[2]; double sg[2];
    const double* qref; size_t q_batch; int q_kstride; int q_row0[2], q_rstep[2]; double q_sg[2];
    unsigned long long* acc;   // [nprob][2]: sum of displacements, samples
};

__global__ void __launch_bounds__(256)
lwa_probe_kernel(PfxProbeArgs p) {
    __shared__ double Q[kWinMaxRows];
    __shared__ unsigned long long ssum[2];
    const int prob = blockIdx.y, b = prob / p.nhemi, h = prob % p.nhemi;
    const int k = 2 + 8 * (blockIdx.x >> 3);                // 1-based level; eight CTAs share its rows
    if (k > p.kmax - 1) return;
    const int js0 = p.jstart - 1, je0 = p.jend - 1, nact = je0 - js0 + 1;
    if (threadIdx.x < 2) ssum[threadIdx.x] = 0ull;
    for (int j = threadIdx.x; j < nact; j += blockDim.x)
        Q[j] = p.q_sg[h] * p.qref[(size_t)b * p.q_batch + (size_t)(k - 1) * p.q_kstride + p.q_row0[h] + p.q_rstep[h] * (js0 + j)];
    __syncthreads();
    unsigned long long sum = 0, n = 0;
    const int ncol = (p.imax + 31) / 32;
    const int per = (nact + 7) >> 3, r0 = (blockIdx.x & 7) * per, r1 = min(nact, r0 + per);
    for (int t = r0 * ncol + threadIdx.x; t < r1 * ncol; t += blockDim.x) {
        const int x = js0 + t / ncol, i = (t % ncol) * 32;
        const double q = p.sg[h] * p.pv[(size_t)b * p.in_batch + (size_t)(k - 1) * p.in_plane +
                                        (size_t)(p.in_row0[h] + p.in_rstep[h] * x) * p.imax + i];
        int lo = 0, hi = nact;                               // E = #{Q < q} among the active thresholds (plain bisection)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (Q[mid] < q) lo = mid + 1; else hi = mid; }
        if (q == q) { sum += (unsigned)abs(js0 + lo - (x + 1)); ++n; }
    }
    for (int d = 16; d >= 1; d >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, d); n += __shfl_xor_sync(0xffffffffu, n, d); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&ssum[0], sum); atomicAdd(&ssum[1], n); }
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(&p.acc[prob * 2], ssum[0]); atomicAdd(&p.acc[prob * 2 + 1], ssum[1]); }
}

// sel[prob] = 1 (prefix kernel) when the mean sampled displacement exceeds `rows`; the accumulators are cleared for the next call
__global__ void lwa_select_kernel(unsigned long long* acc, int* sel, int nprob, double rows, int force) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nprob) {
        const double s = (double)acc[2 * t], n = (double)acc[2 * t + 1];
        sel[t] = force >= 0 ? force : ((n > 0.0 && s > rows * n) ? 1 : 0);
        acc[2 * t] = 0ull; acc[2 * t + 1] = 0ull;
    }
}

static bool pfx_supported(int imax, int nd, const void* pv, const void* uu, const void* nc) {
    return nc == nullptr && win_supported(imax, nd, pv, uu, nullptr);
}

// same contract as lwa_window_launch (no ncforce); sel / sel_want: optional per-problem selector (see above)
static int lwa_prefix_launch(WinArgs a, int mode, const double* pv, const double* uu, int in_rows, int batch, int nprob,
                             const int* sel, int sel_want, cudaStream_t st) {
    const int nbox = (a.nd + 255) / 256;
    int boxrows = (a.nd + nbox - 1) / nbox;
    boxrows += boxrows & 1;
    a.nbox = nbox; a.boxrows = boxrows;
    CUtensorMap tq, tu;
    int rc;
    if ((rc = win_tensor_map(&tq, pv, a.imax, in_rows, a.kmax, batch, boxrows))) return rc;
    if ((rc = win_tensor_map(&tu, uu, a.imax, in_rows, a.kmax, batch, boxrows))) return rc;
    const int tiles = ceil_div(a.imax, kWinCols);
    if (mode == 1) a.k_per_cta = a.k_count;
    else a.k_per_cta = std::max(1, std::min(a.k_count, ceil_div((long long)a.k_count * tiles * nprob, 148 * 8)));
    const int chunks = ceil_div(a.k_count, a.k_per_cta);
    const size_t smem = pfx_smem_layout().total;
    dim3 grid(tiles, chunks, nprob);
    auto go = [&](auto kern) -> int {
        FALWA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kPfxWarps * 32, smem, st>>>(tq, tu, a, sel, sel_want);
        return 0;
    };
    FALWA_KERNEL("lwa_prefix_kernel", st);
    const bool count = a.mask_count != nullptr;
    if (mode == 1) rc = count ? go(lwa_prefix_kernel<1, true>) : go(lwa_prefix_kernel<1, false>);
    else rc = count ? go(lwa_prefix_kernel<0, true>) : go(lwa_prefix_kernel<0, false>);
    if (rc) return rc;
    FALWA_LAUNCH_CHECK();
    return 0;
}

}  // namespace falwa
