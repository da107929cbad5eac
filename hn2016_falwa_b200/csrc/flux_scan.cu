// Interval-scan local wave activity: the O(nd) replacement of the reference's O(nd^2) (j, jj) double loop
// (compute_flux_dirinv.f90:70-116, compute_lwa_only_nhn22.f90:55-90).
//
// For one column (fixed longitude i and level k) with PV values q_jj (jj = 1..nd, equator -> pole) and the
// reference-state thresholds Q_j (j = jstart..jend), the reference accumulates, for every row j,
//     cyclonic      sums over { jj <  j : q_jj >  Q_j }      (astar1, and + contributions to ua2 / ncforce3d)
//     anticyclonic  sums over { jj >= j : q_jj <= Q_j }      (astar2, and - contributions)
// of  (q_jj - Q_j) cos(phi_jj) a dphi,  (q_jj - Q_j)(u_jj cos(phi_j) - uref_j cos(phi_jj)) a dphi,  ncforce_jj cos(phi_jj) a dphi.
// Every term is a product of a factor that depends on jj only and one that depends on j only, so all of them
// follow from six (seven with ncforce) set sums  S(j) = sum_{jj in set(j)} f(jj).
//
// Let m_0 < m_1 < ... < m_{M-1} be rows whose thresholds are non-decreasing (all rows when Qref is monotone, which
// the area analysis guarantees up to rounding; rows that break monotonicity are "exception rows" and are done by
// the brute-force fix-up kernel).  With  E(jj) = #{t : Q_{m_t} < q_jj}  (binary search) and
// nxt(jj) = #{t : m_t <= jj},  row jj belongs to the cyclonic set of m_t  iff  nxt(jj) <= t < E(jj)  and to the
// anticyclonic set iff  E(jj) <= t < nxt(jj):  contiguous intervals in t.  Each set sum is therefore a prefix sum
// over t of  (+f at t = nxt(jj), -f at t = E(jj))  — the second part is a scatter by key E(jj).
//
// Kernel layout: one CTA per (snapshot, hemisphere, level, tile of 8 longitudes); the tile is staged through shared
// memory with coalesced 64-byte row segments and transposed so that each warp owns one column with lanes over rows.
// The scatter is a per-warp counting sort on int16 keys: the slot inside a bin comes from a native 32-bit shared-memory
// atomic and the few bins that hold more than one row are insertion-sorted, so every floating-point sum runs in a
// fixed order (no fp64 atomics anywhere: fp64 ATOMS is a CAS loop); rows whose E equals nxt (not displaced across a
// Qref contour) contribute nothing and are skipped.  Each lane then owns R consecutive thresholds, accumulates the
// six running sums sequentially, and a single warp scan of the lane totals supplies the offsets (the outputs are
// linear in the sums).  Results are staged back through the same shared memory and written with coalesced stores.
#include "common.cuh"

namespace falwa {

constexpr int kScanCols = 8;    // longitudes (= warps) per CTA
constexpr int kScanRMax = 13;   // thresholds per lane in the output phase (odd => conflict-free lane strides)
constexpr int kScanMaxRows = 32 * kScanRMax;  // 416 hemispheric rows (0.25 deg: 361)
constexpr int kScanCells = 1024;  // uniform PV grid that brackets the threshold search (cell -> range of thresholds)

// cell of a PV value: monotone in q, so thresholds in lower (higher) cells are < (>) q
__device__ __forceinline__ int scan_cell(double q, double lo, double inv) {
    return min(max(__double2int_rz((q - lo) * inv), 0), kScanCells - 1);
}

struct ScanFrame {
    int in_row0, in_rstep;    // memory row of hemispheric row jj (1-based): in_row0 + in_rstep * (jj - 1)
    int out_row0, out_rstep;  // same for the outputs
    int skip_first_row;       // do not write hemispheric row 1 (the other hemisphere's launch owns the equator)
    double sg;                // sign applied to pv in the hemispheric frame
};

struct ScanArgs {
    int imax, nd, kmax, jstart, jend, nhemi;
    const double *pv, *uu, *nc;          // 3-D inputs; nc may be null
    size_t in_batch, in_plane;           // element strides: snapshot, level
    double *a1, *a2, *u2, *ncf;          // 3-D outputs; u2 / ncf may be null
    size_t out_batch, out_plane;
    int out_pitch;                        // elements per output row
    ScanFrame fr[2];
    const double* cosphi;                 // [nd+2], index jj
    double ab;                            // a * dphi
    // per (problem = b * nhemi + h, level): tables of lwa_thresholds_kernel
    const double *thr, *urt;              // [nprob][kmax][nd]: thresholds (exception rows appended), uref alike
    const short *rowof, *nxt, *exc;       // [nprob][kmax][nd], [..][nd+2], [..][nd]
    const short* cell;                    // [nprob][kmax][kScanCells+1]: #thresholds below each cell of a uniform grid
    const double* srch;                   // [nprob][kmax][2]: grid origin, cells per unit
    const int* meta;                      // [nprob][kmax][4]: M, nexc
    unsigned long long* mask_count;       // optional: [0] anticyclonic pairs, [1] cyclonic pairs
};

struct ThrArgs {
    int nd, kmax, jstart, jend, nhemi;
    const double* qref; size_t q_batch; int q_kstride; int q_row0[2], q_rstep[2]; double q_sg[2];
    const double* uref; size_t u_batch; int u_kstride; int u_row0[2], u_rstep[2];  // uref may be null (-> 0)
    double *thr, *urt; short *rowof, *nxt, *exc; int* meta;
    short* cell; double* srch;
};

// ---------------------------------------------------------------------------------------------
// Thresholds of one (problem, level): the monotone subsequence of Qref rows, its inverse maps and the exceptions.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lwa_thresholds_kernel(ThrArgs p) {
    extern __shared__ double shd[];
    const int nd = p.nd;
    double* Q = shd;                                   // [nd+1], 1-based
    double* U = shd + nd + 1;                          // [nd+1]
    short* s_rowof = (short*)(U + nd + 1);             // [nd]
    short* s_nxt = s_rowof + nd;                       // [nd+2]
    short* s_exc = s_nxt + nd + 2;                     // [nd]
    unsigned char* keepA = (unsigned char*)(s_exc + nd);  // [nd+1]
    unsigned char* keepB = keepA + nd + 1;                // [nd+1]
    __shared__ int s_M, s_nexc;
    __shared__ int s_cc[kScanCells + 1];
    const int k = blockIdx.x + 2;                      // 1-based level
    const int prob = blockIdx.y, b = prob / p.nhemi, h = prob % p.nhemi;
    for (int j = p.jstart + threadIdx.x; j <= p.jend; j += blockDim.x) {
        Q[j] = p.q_sg[h] * p.qref[(size_t)b * p.q_batch + (size_t)(k - 1) * p.q_kstride + p.q_row0[h] + p.q_rstep[h] * (j - 1)];
        U[j] = p.uref ? p.uref[(size_t)b * p.u_batch + (size_t)(k - 1) * p.u_kstride + p.u_row0[h] + p.u_rstep[h] * (j - 1)] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // two greedy non-decreasing subsequences; keep the one with fewer exception rows
        int excA = 0, excB = 0;
        double mx = -INFINITY, mn = INFINITY;
        for (int j = p.jstart; j <= p.jend; ++j) {
            if (Q[j] >= mx) { keepA[j] = 1; mx = Q[j]; } else { keepA[j] = 0; ++excA; }
        }
        for (int j = p.jend; j >= p.jstart; --j) {
            if (Q[j] <= mn) { keepB[j] = 1; mn = Q[j]; } else { keepB[j] = 0; ++excB; }
        }
        const unsigned char* keep = (excB < excA) ? keepB : keepA;
        int M = 0, ne = 0;
        s_nxt[0] = 0;
        for (int jj = 1; jj <= nd; ++jj) {
            if (jj >= p.jstart && jj <= p.jend) {
                if (keep[jj]) s_rowof[M++] = (short)jj; else s_exc[ne++] = (short)jj;
            }
            s_nxt[jj] = (short)M;
        }
        s_nxt[nd + 1] = (short)M;
        s_M = M; s_nexc = ne;
    }
    __syncthreads();
    const size_t pk = (size_t)prob * p.kmax + (k - 1);
    const int M = s_M, ne = s_nexc;
    for (int t = threadIdx.x; t < M; t += blockDim.x) {
        p.thr[pk * nd + t] = Q[s_rowof[t]];
        p.urt[pk * nd + t] = U[s_rowof[t]];
        p.rowof[pk * nd + t] = s_rowof[t];
    }
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        p.thr[pk * nd + M + e] = Q[s_exc[e]];
        p.urt[pk * nd + M + e] = U[s_exc[e]];
        p.exc[pk * nd + e] = s_exc[e];
    }
    for (int jj = threadIdx.x; jj <= nd + 1; jj += blockDim.x) p.nxt[pk * (nd + 2) + jj] = s_nxt[jj];
    if (threadIdx.x == 0) { p.meta[pk * 4 + 0] = M; p.meta[pk * 4 + 1] = ne; }
    // search accelerator: cell[c] = number of thresholds in cells < c of a uniform grid over [thr_0, thr_{M-1}]
    double lo = 0.0, inv = 0.0;
    if (M > 0) {
        lo = Q[s_rowof[0]];
        const double hi = Q[s_rowof[M - 1]];
        if (hi > lo) inv = (double)kScanCells / (hi - lo);
        if (!(inv < 1e300)) inv = 0.0;  // overflow / NaN: every value falls in cell 0 and the search is a plain bisection
    }
    for (int c = threadIdx.x; c <= kScanCells; c += blockDim.x) s_cc[c] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < M; t += blockDim.x) atomicAdd(&s_cc[scan_cell(Q[s_rowof[t]], lo, inv) + 1], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c <= kScanCells; ++c) { run += s_cc[c]; s_cc[c] = run; }
        p.srch[pk * 2 + 0] = lo;
        p.srch[pk * 2 + 1] = inv;
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= kScanCells; c += blockDim.x) p.cell[pk * (kScanCells + 1) + c] = (short)s_cc[c];
}

__host__ __device__ inline int scan_pitch(int nd) { return ((nd - 2 + 15) / 16) * 16 + 2; }  // == 2 (mod 16)

static size_t scan_smem_bytes(int nd, bool has_nc) {
    const int ndp = scan_pitch(nd), narr = has_nc ? 4 : 3;
    size_t doubles = (size_t)2 * nd + (nd + 2) + (size_t)narr * kScanCols * ndp;
    size_t ints = (size_t)kScanCols * (nd + 2);
    size_t shorts = (size_t)2 * kScanCols * ndp + nd + (nd + 2) + (kScanCells + 1);
    return doubles * 8 + ints * 4 + ((shorts * 2 + 7) / 8) * 8;
}

__device__ __forceinline__ double shfl_up_f64(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

// inclusive -> exclusive warp scan of one double (result: sum over lower lanes)
__device__ __forceinline__ double warp_excl_scan(double v, int lane) {
    double incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double n = shfl_up_f64(incl, d);
        if (lane >= d) incl += n;
    }
    return incl - v;
}

template <bool HAS_NC>
__global__ void __launch_bounds__(kScanCols * 32, 2)
lwa_scan_kernel(ScanArgs p) {
    extern __shared__ double shd[];
    const int nd = p.nd, ndp = scan_pitch(nd), cp = nd + 2;
    constexpr int NARR = HAS_NC ? 4 : 3;
    double* s_thr = shd;
    double* s_urt = s_thr + nd;
    double* s_cos = s_urt + nd;                         // [nd+2]
    double* s_q = s_cos + nd + 2;                       // [C][ndp]  pv  -> a2
    double* s_u = s_q + kScanCols * ndp;                // [C][ndp]  u   -> ua2
    double* s_o = s_u + kScanCols * ndp;                // [C][ndp]        a1
    double* s_n = s_o + kScanCols * ndp;                // [C][ndp]  nc  -> ncforce3d (HAS_NC only)
    int* s_cnt = (int*)(shd + (size_t)2 * nd + (nd + 2) + (size_t)NARR * kScanCols * ndp);  // [C][cp]
    short* s_E = (short*)(s_cnt + kScanCols * cp);      // [C][ndp]
    short* s_perm = s_E + kScanCols * ndp;              // [C][ndp]
    short* s_rowof = s_perm + kScanCols * ndp;          // [nd]
    short* s_nxt = s_rowof + nd;                        // [nd+2]
    short* s_cell = s_nxt + nd + 2;                     // [kScanCells+1]

    const int k = blockIdx.y + 2;                       // 1-based level
    const int prob = blockIdx.z, b = prob / p.nhemi, h = prob % p.nhemi;
    const ScanFrame fr = p.fr[h];
    const size_t pk = (size_t)prob * p.kmax + (k - 1);
    const int M = p.meta[pk * 4 + 0];
    const int i0 = blockIdx.x * kScanCols;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;

    // ---- stage the (nd x 8) tile and the tables ------------------------------------------------
    // All global loads of the tile are issued before anything waits on them (one DRAM round trip for the whole
    // tile instead of one per group of rows); the table loads (L2 hits) travel behind them.
    {
        const int col = tid % kScanCols, rr = tid / kScanCols;
        const bool colok = (i0 + col) < p.imax;
        const size_t base = (size_t)b * p.in_batch + (size_t)(k - 1) * p.in_plane + i0 + col;
        double qv[kScanRMax], uv[kScanRMax], cv[HAS_NC ? kScanRMax : 1];
#pragma unroll
        for (int c = 0; c < kScanRMax; ++c) {
            const int r = rr + 32 * c;
            qv[c] = 0.0; uv[c] = 0.0;
            if (HAS_NC) cv[c] = 0.0;
            if (colok && r < nd) {
                const size_t g = base + (size_t)(fr.in_row0 + fr.in_rstep * r) * p.imax;
                qv[c] = ld_stream(&p.pv[g]);
                uv[c] = ld_stream(&p.uu[g]);
                if (HAS_NC) cv[c] = ld_stream(&p.nc[g]);
            }
        }
        for (int t = tid; t < M; t += blockDim.x) {
            s_thr[t] = p.thr[pk * nd + t];
            s_urt[t] = p.urt[pk * nd + t];
            s_rowof[t] = p.rowof[pk * nd + t];
        }
        for (int jj = tid; jj <= nd + 1; jj += blockDim.x) {
            s_cos[jj] = p.cosphi[jj];
            s_nxt[jj] = p.nxt[pk * (nd + 2) + jj];
        }
        for (int c = tid; c <= kScanCells; c += blockDim.x) s_cell[c] = p.cell[pk * (kScanCells + 1) + c];
#pragma unroll
        for (int c = 0; c < kScanRMax; ++c) {
            const int r = rr + 32 * c;
            if (r < nd) {
                s_q[col * ndp + r] = fr.sg * qv[c];
                s_u[col * ndp + r] = uv[c];
                if (HAS_NC) s_n[col * ndp + r] = cv[c];
            }
        }
    }
    const double srch_lo = p.srch[pk * 2 + 0], srch_inv = p.srch[pk * 2 + 1];
    __syncthreads();

    // ---- per-warp column work ----------------------------------------------------------------
    double* __restrict__ q_ = s_q + w * ndp;
    double* __restrict__ u_ = s_u + w * ndp;
    double* __restrict__ o_ = s_o + w * ndp;
    double* __restrict__ n_ = s_n + w * ndp;
    int* __restrict__ cn = s_cnt + w * cp;
    short* __restrict__ E_ = s_E + w * ndp;
    short* __restrict__ pm = s_perm + w * ndp;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int t = lane; t <= M; t += 32) cn[t] = 0;
    for (int r = lane; r < nd; r += 32) o_[r] = 0.0;
    __syncwarp();

    // 1. E(jj) by cell table + bisection (four row chunks interleaved for ILP) and the slot of every displaced row
    //    inside its bin
    unsigned long long ccyc = 0, canti = 0;
    for (int r0 = 0; r0 < nd; r0 += 128) {
        double q4[4];
        int e4[4], b4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int r = r0 + 32 * c + lane;
            q4[c] = (r < nd) ? q_[r] : 0.0;
            const int cl = scan_cell(q4[c], srch_lo, srch_inv);
            e4[c] = s_cell[cl];       // E lies in [cell[cl], cell[cl+1]]
            b4[c] = s_cell[cl + 1];
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (e4[c] < b4[c]) {
                    const int mid = (e4[c] + b4[c]) >> 1;
                    if (s_thr[mid] < q4[c]) e4[c] = mid + 1; else b4[c] = mid;
                }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            while (e4[c] < b4[c]) {
                const int mid = (e4[c] + b4[c]) >> 1;
                if (s_thr[mid] < q4[c]) e4[c] = mid + 1; else b4[c] = mid;
            }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int r = r0 + 32 * c + lane;
            if (r < nd) {
                // a NaN PV value belongs to neither set of any row (the reference's comparisons are false): not displaced
                const int nx = s_nxt[r + 1], e = (q4[c] == q4[c]) ? e4[c] : nx;
                E_[r] = (short)e;
                if (e > nx) ccyc += (unsigned)(e - nx); else canti += (unsigned)(nx - e);
                // slot inside the bin: native 32-bit shared-memory atomic (arrival order is arbitrary; step 4 sorts
                // the few bins that hold more than one row, so every sum is still reproducible bit for bit)
                if ((e != nx) && (e < M)) pm[r] = (short)atomicAdd(&cn[e], 1);
            }
        }
    }
    __syncwarp();
    // 2. exclusive scan of the counts -> bin starts; cn[M] = number of displaced rows
    {
        int carry = 0;
        for (int t0 = 0; t0 <= M; t0 += 32) {
            const int t = t0 + lane;
            const int v = (t < M) ? cn[t] : 0;
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += n;
            }
            if (t <= M) cn[t] = carry + incl - v;
            carry += __shfl_sync(FULL, incl, 31);
        }
    }
    __syncwarp();
    // 3. placement: sorted position = bin start + rank.  The ranks were stashed in pm[row]; all of them are read
    //    into registers before any sorted slot is written (a slot may alias another row's stash).
    {
        int dst[kScanRMax];
#pragma unroll
        for (int c = 0; c < kScanRMax; ++c) {
            const int r = 32 * c + lane;
            dst[c] = -1;
            if (r < nd) {
                const int e = E_[r];
                if (e != s_nxt[r + 1] && e < M) dst[c] = cn[e] + pm[r];
            }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < kScanRMax; ++c)
            if (dst[c] >= 0) pm[dst[c]] = (short)(32 * c + lane);
    }
    __syncwarp();

    // 4. running set sums: lane owns thresholds t = lane*R .. lane*R + R-1
    int R = (M + 31) / 32;
    if ((R & 1) == 0) ++R;
    double S1c = 0, S2c = 0, S1a = 0, S2a = 0, T3 = 0, T4 = 0, T5 = 0;
    double a2l[kScanRMax], u2l[kScanRMax], ncl[kScanRMax];
    const double ab = p.ab;
    int jprev = 1, xcarry = 0;   // threshold row and bin end of the previous step (carried in registers)
#pragma unroll
    for (int s = 0; s < kScanRMax; ++s) {
        a2l[s] = 0.0; u2l[s] = 0.0; ncl[s] = 0.0;
        const int t = lane * R + s;
        if (s < R && t < M) {
            // everything that does not depend on the two loops is loaded first, so that its shared-memory latency
            // overlaps with them (the compiler cannot hoist these loads above the loops' shared-memory stores)
            const int j = s_rowof[t];
            const int jfirst = (s == 0) ? (t ? (int)s_rowof[t - 1] : 1) : jprev;
            const int x0 = (s == 0) ? cn[t] : xcarry, xe = cn[t + 1];
            const double Q = s_thr[t], ur = s_urt[t], cj = s_cos[j];
            jprev = j; xcarry = xe;
            // Common case: ONE row enters at t (nxt(jj) == t  <=>  m_{t-1} <= jj < m_t; a single row when Qref is
            // monotone) and at most ONE displaced row sits in bin t.  Their operands are fetched up front and applied
            // with predicated straight-line code, so a step is two dependent rounds of shared-memory loads instead of
            // six; the rare extra rows follow in rolled loops.  Adding 0.0 for the inactive cases keeps every sum
            // bit-identical to the plain loops (same terms, same order).
            const bool ent = jfirst < j;
            const int ei = ent ? jfirst : j;                       // any valid row when nothing enters
            const int e1 = E_[ei - 1];
            const double qe = q_[ei - 1], ue = u_[ei - 1], ce = s_cos[ei];
            double ne = 0.0;
            if (HAS_NC) ne = n_[ei - 1];
            const bool binr = x0 < xe;
            int r0 = binr ? (int)pm[x0] : 0;
            if (__builtin_expect(xe - x0 >= 2, 0)) {               // ascending row order inside the bin (rare)
#pragma unroll 1
                for (int a = x0 + 1; a < xe; ++a) {
                    const short v = pm[a];
                    int c = a - 1;
#pragma unroll 1
                    for (; c >= x0 && pm[c] > v; --c) pm[c + 1] = pm[c];
                    pm[c + 1] = v;
                }
                r0 = pm[x0];
            }
            const double qb = q_[r0], ub = u_[r0], cb = s_cos[r0 + 1];
            const int nb0 = s_nxt[r0 + 1];
            double nbv = 0.0;
            if (HAS_NC) nbv = n_[r0];
            {   // first entering row
                // class weights 1.0 / 0.0 and one fused multiply-add per sum: fma(1, x, S) == S + x and
                // fma(0, x, S) == S exactly for finite x, at a third of the instructions of select + add on 64-bit
                // values; the operands of a row outside both sets are zeroed so that a NaN / Inf there cannot leak
                const bool cyc = ent && (e1 > t), anti = ent && (e1 < t), on = cyc || anti;
                const double wc = cyc ? 1.0 : 0.0, wa = anti ? -1.0 : 0.0, wd = on ? 1.0 : 0.0;
                const double qc = on ? qe * ce : 0.0, qu = on ? qe * ue : 0.0, uz = on ? ue : 0.0;
                S1c = __fma_rn(wc, qc, S1c); S2c = __fma_rn(wc, ce, S2c);
                S1a = __fma_rn(wa, qc, S1a); S2a = __fma_rn(wa, ce, S2a);
                T3 = __fma_rn(wd, qu, T3); T4 = __fma_rn(wd, uz, T4);
                if (HAS_NC) T5 = __fma_rn(wd, on ? ne * ce : 0.0, T5);
            }
#pragma unroll 1
            for (int jj = jfirst + 1; jj < j; ++jj) {              // further entering rows (exception rows, t = 0)
                const int e = E_[jj - 1];
                if (e != t) {
                    const double q = q_[jj - 1], u = u_[jj - 1], cs = s_cos[jj];
                    if (e > t) { S1c += q * cs; S2c += cs; } else { S1a -= q * cs; S2a -= cs; }
                    T3 += q * u; T4 += u;
                    if (HAS_NC) T5 += n_[jj - 1] * cs;
                }
            }
            {   // first row leaving (cyclonic) / entering (anticyclonic) at t: E(jj) == t
                const bool cyc = binr && (t > nb0), anti = binr && !(t > nb0);
                const double wc = cyc ? -1.0 : 0.0, wa = anti ? 1.0 : 0.0, wd = binr ? -1.0 : 0.0;
                const double qc = binr ? qb * cb : 0.0, qu = binr ? qb * ub : 0.0, uz = binr ? ub : 0.0;
                S1c = __fma_rn(wc, qc, S1c); S2c = __fma_rn(wc, cb, S2c);
                S1a = __fma_rn(wa, qc, S1a); S2a = __fma_rn(wa, cb, S2a);
                T3 = __fma_rn(wd, qu, T3); T4 = __fma_rn(wd, uz, T4);
                if (HAS_NC) T5 = __fma_rn(wd, binr ? nbv * cb : 0.0, T5);
            }
#pragma unroll 1
            for (int x = x0 + 1; x < xe; ++x) {
                const int r = pm[x];
                const double q = q_[r], u = u_[r], cs = s_cos[r + 1];
                if (t > s_nxt[r + 1]) { S1c -= q * cs; S2c -= cs; } else { S1a += q * cs; S2a += cs; }
                T3 -= q * u; T4 -= u;
                if (HAS_NC) T5 -= n_[r] * cs;
            }
            // a1 = ab (S1c - Q S2c),  a2 = ab (Q S2a - S1a),  ua2 = ab cos_j (T3 - Q T4) - uref_j (a1 + a2)
            const double a1 = __fma_rn(-Q, S2c, S1c) * ab, a2 = __fma_rn(Q, S2a, -S1a) * ab;
            o_[j - 1] = a1;
            a2l[s] = a2;
            u2l[s] = __fma_rn(ab * cj, __fma_rn(-Q, T4, T3), -ur * (a1 + a2));
            if (HAS_NC) ncl[s] = ab * T5;
        }
    }
    // offsets from the lower lanes (the outputs are linear in the sums)
    const double O1c = warp_excl_scan(S1c, lane), O2c = warp_excl_scan(S2c, lane), O1a = warp_excl_scan(S1a, lane),
                 O2a = warp_excl_scan(S2a, lane), O3 = warp_excl_scan(T3, lane), O4 = warp_excl_scan(T4, lane);
    double O5 = 0.0;
    if (HAS_NC) O5 = warp_excl_scan(T5, lane);
    __syncwarp();  // every lane is done reading q_, u_, n_
    for (int r = lane; r < nd; r += 32) {
        q_[r] = 0.0; u_[r] = 0.0;
        if (HAS_NC) n_[r] = 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kScanRMax; ++s) {
        const int t = lane * R + s;
        if (s < R && t < M) {
            const int j = s_rowof[t];
            const double Q = s_thr[t], ur = s_urt[t], cj = s_cos[j];
            const double d1 = __fma_rn(-Q, O2c, O1c) * ab, d2 = __fma_rn(Q, O2a, -O1a) * ab;
            o_[j - 1] += d1;
            q_[j - 1] = a2l[s] + d2;
            u_[j - 1] = u2l[s] + __fma_rn(ab * cj, __fma_rn(-Q, O4, O3), -ur * (d1 + d2));
            if (HAS_NC) n_[j - 1] = ncl[s] + ab * O5;
        }
    }
    if (p.mask_count) {
        ccyc = __reduce_add_sync(FULL, (unsigned)ccyc);
        canti = __reduce_add_sync(FULL, (unsigned)canti);
        if (lane == 0 && (i0 + w) < p.imax && (ccyc | canti)) {
            atomicAdd(&p.mask_count[0], canti);
            atomicAdd(&p.mask_count[1], ccyc);
        }
    }
    __syncthreads();

    // ---- coalesced write-out -------------------------------------------------------------------
    {
        const int col = tid % kScanCols, rr = tid / kScanCols;
        if ((i0 + col) < p.imax) {
            const size_t base = (size_t)b * p.out_batch + (size_t)(k - 1) * p.out_plane + i0 + col;
            for (int r = rr; r < nd; r += 32) {
                if (r == 0 && fr.skip_first_row) continue;
                const size_t g = base + (size_t)(fr.out_row0 + fr.out_rstep * r) * p.out_pitch;
                p.a1[g] = s_o[col * ndp + r];
                p.a2[g] = s_q[col * ndp + r];
                if (p.u2) p.u2[g] = s_u[col * ndp + r];
                if (HAS_NC && p.ncf) p.ncf[g] = s_n[col * ndp + r];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Exception rows (thresholds outside the monotone subsequence): the reference's double loop for those rows only.
// grid: (ceil(imax/128), slots, (kmax-2) * nprob)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lwa_exception_rows_kernel(ScanArgs p) {
    const int lev = blockIdx.z % (p.kmax - 2), prob = blockIdx.z / (p.kmax - 2);
    const int k = lev + 2, b = prob / p.nhemi, h = prob % p.nhemi;
    const size_t pk = (size_t)prob * p.kmax + (k - 1);
    const int M = p.meta[pk * 4 + 0], nexc = p.meta[pk * 4 + 1];
    if (nexc == 0) return;
    const ScanFrame fr = p.fr[h];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nd = p.nd;
    unsigned long long c1 = 0, c2 = 0;
    for (int e = blockIdx.y; e < nexc; e += gridDim.y) {
        const int j = p.exc[pk * nd + e];
        if (i < p.imax && !(j == 1 && fr.skip_first_row)) {
            const double qr = p.thr[pk * nd + M + e], ur = p.urt[pk * nd + M + e];
            const double cphi0 = p.cosphi[j];
            const size_t base = (size_t)b * p.in_batch + (size_t)(k - 1) * p.in_plane + i;
            double a1 = 0., a2 = 0., ncf = 0., u2 = 0.;
            for (int jj = 1; jj < j; ++jj) {
                const size_t g = base + (size_t)(fr.in_row0 + fr.in_rstep * (jj - 1)) * p.imax;
                const double qe = fr.sg * p.pv[g] - qr;
                if (qe > 0.) {
                    const double aa = p.ab * p.cosphi[jj];
                    const double ue = p.uu[g] * cphi0 - ur * p.cosphi[jj];
                    a1 = a1 + qe * aa;
                    if (p.nc) ncf = ncf + p.nc[g] * aa;
                    u2 = u2 + qe * ue * p.ab;
                    ++c1;
                }
            }
            for (int jj = j; jj <= nd; ++jj) {
                const size_t g = base + (size_t)(fr.in_row0 + fr.in_rstep * (jj - 1)) * p.imax;
                const double qe = fr.sg * p.pv[g] - qr;
                if (qe <= 0.) {
                    const double aa = p.ab * p.cosphi[jj];
                    const double ue = p.uu[g] * cphi0 - ur * p.cosphi[jj];
                    a2 = a2 - qe * aa;
                    if (p.nc) ncf = ncf - p.nc[g] * aa;
                    u2 = u2 - qe * ue * p.ab;
                    ++c2;
                }
            }
            const size_t g = (size_t)b * p.out_batch + (size_t)(k - 1) * p.out_plane +
                             (size_t)(fr.out_row0 + fr.out_rstep * (j - 1)) * p.out_pitch + i;
            p.a1[g] = a1;
            p.a2[g] = a2;
            if (p.u2) p.u2[g] = u2;
            if (p.nc && p.ncf) p.ncf[g] = ncf;
        }
    }
    if (p.mask_count) {
        c1 = __reduce_add_sync(0xffffffffu, (unsigned)c1);
        c2 = __reduce_add_sync(0xffffffffu, (unsigned)c2);
        if ((threadIdx.x & 31) == 0 && (c1 | c2)) {
            atomicAdd(&p.mask_count[0], c2);
            atomicAdd(&p.mask_count[1], c1);
        }
    }
}

// Scratch of the threshold tables for nprob problems.
struct ScanTables {
    DeviceScratch<double> dbl;   // thr, urt
    DeviceScratch<short> sh;     // rowof, nxt, exc, cell
    DeviceScratch<int> meta;
    int alloc(int nprob, int nd, int kmax, cudaStream_t st) {
        int rc;
        const size_t pk = (size_t)nprob * kmax;
        if ((rc = dbl.alloc(pk * nd * 2 + pk * 2, st))) return rc;
        if ((rc = sh.alloc(pk * ((size_t)2 * nd + nd + 2 + kScanCells + 1), st))) return rc;
        if ((rc = meta.alloc(pk * 4, st, true))) return rc;
        return 0;
    }
    void bind(ThrArgs& t, ScanArgs& s, int nprob, int nd, int kmax) {
        const size_t pk = (size_t)nprob * kmax;
        t.thr = dbl.d; t.urt = dbl.d + pk * nd; t.srch = dbl.d + pk * nd * 2;
        t.rowof = sh.d; t.exc = sh.d + pk * nd; t.nxt = sh.d + pk * nd * 2;
        t.cell = sh.d + pk * ((size_t)2 * nd + nd + 2);
        t.meta = meta.d;
        s.thr = t.thr; s.urt = t.urt; s.rowof = t.rowof; s.exc = t.exc; s.nxt = t.nxt; s.meta = t.meta;
        s.cell = t.cell; s.srch = t.srch;
    }
};

static bool scan_supported(int nd) { return nd >= 4 && nd <= kScanMaxRows; }

// thresholds -> scan -> exception fix-up, for nprob = batch * nhemi problems
static int lwa_scan_launch(const ThrArgs& t, const ScanArgs& s, int nprob, cudaStream_t st) {
    const int nd = s.nd, kmax = s.kmax;
    {
        const size_t smem = sizeof(double) * 2 * (nd + 1) + sizeof(short) * ((size_t)3 * nd + 2) + 2 * (nd + 1) + 16;
        FALWA_KERNEL("lwa_thresholds_kernel", st);
        lwa_thresholds_kernel<<<dim3(kmax - 2, nprob), 128, smem, st>>>(t);
        FALWA_LAUNCH_CHECK();
    }
    const bool has_nc = s.nc != nullptr;
    const size_t smem = scan_smem_bytes(nd, has_nc);
    dim3 grid(ceil_div(s.imax, kScanCols), kmax - 2, nprob);
    {
        FALWA_KERNEL("lwa_scan_kernel", st);
        if (has_nc) {
            FALWA_CUDA_TRY(cudaFuncSetAttribute(lwa_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            lwa_scan_kernel<true><<<grid, kScanCols * 32, smem, st>>>(s);
        } else {
            FALWA_CUDA_TRY(cudaFuncSetAttribute(lwa_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            lwa_scan_kernel<false><<<grid, kScanCols * 32, smem, st>>>(s);
        }
        FALWA_LAUNCH_CHECK();
    }
    {
        FALWA_KERNEL("lwa_exception_rows_kernel", st);
        lwa_exception_rows_kernel<<<dim3(ceil_div(s.imax, 128), 4, (kmax - 2) * nprob), 128, 0, st>>>(s);
        FALWA_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace falwa
