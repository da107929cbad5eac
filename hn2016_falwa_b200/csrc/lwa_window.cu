// Local wave activity by exact windowed sums: the B200-native replacement of the reference's O(nd^2) (j, jj) double
// loop (compute_flux_dirinv.f90:70-116, compute_lwa_only_nhn22.f90:55-90) on the hot path.
//
// For one column (fixed longitude and level) with PV values q_jj (jj = 1..nd, equator -> pole, hemispheric frame) and
// thresholds Q_j, the reference accumulates for every row j
//     cyclonic      { jj <  j : q_jj >  Q_j }      astar1 += (q_jj - Q_j) a dphi cos(phi_jj)
//     anticyclonic  { jj >= j : q_jj <= Q_j }      astar2 -= (q_jj - Q_j) a dphi cos(phi_jj)
// and ua2 / ncforce3d over the same two sets.  In the atmosphere a PV contour is displaced by a few rows only, so the
// two sets of row j live in a short window around j.  The window is found EXACTLY, per threshold, from two monotone
// bound arrays of the column: PE(x) = max q over rows < x and PP(x) = min q over rows >= x.  Going equatorward from
// j-1, no row at or beyond x can be in the cyclonic set once PE(x+1) <= Q_j; going poleward from j, none once
// PP(x) > Q_j.  Inside the window every (j, jj) pair is evaluated with the reference's own predicate and term
// ((q - Q) > 0, (q - Q) * aa), so set membership is exact and no sort, search or atomic is involved; a column whose
// contours are displaced far costs proportionally more steps (worst case the reference's full loop), never a wrong
// answer.  The bounds are carried as the upper 32 bits of the order-preserving integer image of the doubles
// (conservative by construction: q > Q implies key(q) >= key(Q)), so building them is integer min/max work.
//
// Layout: one CTA = (snapshot, hemisphere, 8 longitudes) x all nd rows, walking the levels of its range.  The
// (rows x 8) tiles of q and u (and ncforce) of a level arrive by TMA (cp.async.bulk.tensor, one elected thread, mbarrier
// completion) straight in the layout the math uses — a warp works on 4 rows x 8 longitudes = 256 contiguous bytes of
// shared memory — two levels in flight, so the tile of level k+1 lands while level k is computed.  In the
// accumulating mode (compute_lwa_and_barotropic_fluxes) the density-weighted vertical sums of astar1, astar2, ua1, ua2
// (oopinterface.py:405-420) stay in registers across the level walk and only |astar1 + astar2| — the user-visible 3-D
// LWA — is written per level: the a1 / a2 / ua2 intermediates of the interval-scan path never exist.
#include "common.cuh"
#include <cuda.h>
#include <climits>

namespace falwa {

constexpr int kWinCols = 8;                                       // longitudes per tile: 64-byte row segments
constexpr int kWinMaxRows = 384;                                  // hemispheric rows (0.25 deg: 361): 96 four-row groups
constexpr int kWinSpan = 13;                                      // rows per lane in the bound scans (odd stride)
constexpr int kWinKeyPitch = 420;                                 // >= 32 * 13 + 2 and == 4 (mod 32): conflict-free

struct WinFrame {
    int in_row0, out_row0;   // memory row of tile row 0 in the inputs / outputs
    int rev;                 // tile row m holds hemispheric row jj = rev ? nd - m : m + 1
    int skip_first;          // do not write hemispheric row 1 (the other hemisphere's launch owns the equator)
    double sg;               // sign of pv in the hemispheric frame
};

struct WinArgs {
    int imax, nd, kmax, jstart, jend, nhemi;
    int nbox, boxrows;                      // TMA boxes per tile, rows per box (even; nbox * boxrows >= nd)
    int k_first, k_count, k_per_cta;        // 1-based first level, number of levels, levels per CTA (grid.y chunks)
    WinFrame fr[2];
    const double* qref; size_t q_batch; int q_kstride; int q_row0[2], q_rstep[2]; double q_sg[2];
    const double* uref; size_t u_batch; int u_kstride; int u_row0[2], u_rstep[2];   // uref may be null (-> 0)
    const double* cosphi;                   // [nd+2], index jj
    double ab;                              // a * dphi
    // MODE 0: the 3-D pieces per level (u2 / ncf may be null)
    double *a1, *a2, *u2, *ncf; size_t out_batch, out_plane; int out_pitch;
    // MODE 1: |a1 + a2| per level and the vertical sums, global frame
    double *lwa, *out2d; const double* vw; double dc; size_t lwa_batch, lwa_plane, o2_batch, o2_plane;
    int slot_a1, slot_a2, slot_lwa, slot_ua1, slot_ua2, slot_ncf;
    unsigned long long* mask_count;         // optional: [0] anticyclonic pairs, [1] cyclonic pairs
};

__device__ __forceinline__ uint32_t win_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void win_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(win_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void win_mbar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(win_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void win_mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    const uint32_t a = win_smem(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void win_tma_load(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(win_smem(dst)), "l"(map), "r"(win_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// upper half of the order-preserving integer image of a double (NaN sorts outside the numbers; harmless, see above)
__device__ __forceinline__ int win_key(double x) {
    const int hi = __double2hiint(x);
    return hi ^ ((hi >> 31) & 0x7fffffff);
}

// Shared-memory plan.  The tile stride is a compile-time constant so that, from the address of q(row, col), the
// operands u(row, col) and aa(row) sit at immediate offsets (one address register per window step).
constexpr int kWinTile = kWinMaxRows * kWinCols;     // doubles of one (rows x 8) tile slot
// Bound arrays first, then the tiles — without ncforce [q0 | u0 | aa | q1 | u1 | aa], with it [q0 | u0 | n0 | aa | q1 | u1 | n1]
// (one aa8 between the stages) — then the per-level tables.  Lanes of the last group that lie past the last row compute
// on whatever sits at their (in-bounds, thanks to this order and the tail pad) addresses and never store.
struct WinSmem {
    size_t off_keys, off_pe, off_pp, off_tiles, off_q, off_u, off_cs, off_bar, total;
};
__host__ __device__ inline WinSmem win_smem_layout(int nd, int nfields) {
    WinSmem s;
    size_t o = 0;
    s.off_keys = o; o += (size_t)kWinCols * kWinKeyPitch * 4;
    s.off_pe = o; o += (size_t)kWinCols * kWinKeyPitch * 4;
    s.off_pp = o; o += (size_t)kWinCols * kWinKeyPitch * 4;
    s.off_tiles = o; o += (size_t)(nfields == 3 ? 7 : 6) * kWinTile * 8;
    s.off_q = o; o += (size_t)2 * nd * 8;
    s.off_u = o; o += (size_t)2 * nd * 8;
    s.off_cs = o; o += (size_t)nd * 8;
    s.off_bar = o; o += 32;
    s.total = o + 256;
    return s;
}

// shared-memory accesses with explicit 32-bit addresses (the hot loops carry one address register per operand stream)
__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
template <int OFF> __device__ __forceinline__ double lds64o(uint32_t a) {
    double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF)); return v;
}
__device__ __forceinline__ int lds32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <int OFF> __device__ __forceinline__ int lds32o(uint32_t a) {
    int v; asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF)); return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- tensor memory as the accumulator file -----------------------------------------------------------------------
// The vertical sums of a CTA (4 per grid point of its tile: 92 KB at 0.25 deg) do not fit beside the tiles in shared
// memory, and kept in registers they push the kernel to its register limit (the compiler re-derives addresses instead
// of holding them).  Nothing here uses the tensor cores, so their 256 KB of tensor memory per SM is free: every thread
// owns 8 (10 with ncforce) 32-bit columns of its own TMEM lane per group and updates them once per level with
// tcgen05.ld / tcgen05.st.  A warp reaches lanes 32 (warp % 4) .. + 31 only; warps that share a lane quarter use
// different columns.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(win_smem(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, double& a) {
    uint32_t v0, v1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __hiloint2double((int)v1, (int)v0);
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, double a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                 ::"r"(taddr), "r"(__double2loint(a)), "r"(__double2hiint(a)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double (&a)[4]) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int x = 0; x < 4; ++x) a[x] = __hiloint2double((int)v[2 * x + 1], (int)v[2 * x]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&a)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__double2loint(a[0])), "r"(__double2hiint(a[0])), "r"(__double2loint(a[1])),
                 "r"(__double2hiint(a[1])), "r"(__double2loint(a[2])), "r"(__double2hiint(a[2])),
                 "r"(__double2loint(a[3])), "r"(__double2hiint(a[3])) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One window step of a threshold row: diff = q - Q; in the set (and still inside this lane's window) the pair adds
// diff * aa to the LWA piece and diff * u to the flux sum.  Written with predicated FMAs (no select chains): a pair
// outside the set leaves every sum bit-for-bit untouched, whatever its operands are (NaN-safe like the reference's IF).
template <bool ANTI, bool HAS_NC, bool COUNT>
__device__ __forceinline__ void win_pair(double qv, double Qj, double av, double uv, double nv, bool mine,
                                         double& A, double& B, double& N, unsigned& cnt) {
    unsigned m = mine, c = cnt;
    if (!ANTI) {
        asm("{\n\t.reg .pred p, pm;\n\t.reg .f64 d;\n\t"
            "setp.ne.u32 pm, %6, 0;\n\tsub.rn.f64 d, %4, %5;\n\tsetp.gt.and.f64 p, d, 0d0000000000000000, pm;\n\t"
            "@p fma.rn.f64 %0, d, %7, %0;\n\t@p fma.rn.f64 %1, d, %8, %1;\n\t@p fma.rn.f64 %2, %9, %7, %2;\n\t@p add.u32 %3, %3, 1;\n\t}"
            : "+d"(A), "+d"(B), "+d"(N), "+r"(c) : "d"(qv), "d"(Qj), "r"(m), "d"(av), "d"(uv), "d"(nv));
    } else {
        asm("{\n\t.reg .pred p, pm;\n\t.reg .f64 d;\n\t"
            "setp.ne.u32 pm, %6, 0;\n\tsub.rn.f64 d, %5, %4;\n\tsetp.ge.and.f64 p, d, 0d0000000000000000, pm;\n\t"
            "@p fma.rn.f64 %0, d, %7, %0;\n\t@p fma.rn.f64 %1, d, %8, %1;\n\t@p fma.rn.f64 %2, %9, %7, %2;\n\t@p add.u32 %3, %3, 1;\n\t}"
            : "+d"(A), "+d"(B), "+d"(N), "+r"(c) : "d"(qv), "d"(Qj), "r"(m), "d"(av), "d"(uv), "d"(-nv));
    }
    if (COUNT) cnt = c;
    (void)HAS_NC;
}

template <int MODE, bool HAS_NC, bool COUNT, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
lwa_window_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_u,
                  const __grid_constant__ CUtensorMap tm_n, const WinArgs p, const int* __restrict__ sel, int sel_want) {
    extern __shared__ __align__(1024) unsigned char win_raw[];
    if (sel && sel[blockIdx.z] != sel_want) return;    // the prefix kernel (lwa_prefix.cu) owns this problem
    constexpr int NT = NW * 32;
    constexpr int GPW = kWinMaxRows / (4 * NW);                          // 4-row groups a warp owns
    constexpr int NF = HAS_NC ? 3 : 2;
    constexpr int P = kWinKeyPitch;
    constexpr int TU = kWinTile, TN = 2 * kWinTile;                      // u, nc relative to q (doubles)
    constexpr int SD = (HAS_NC ? 4 : 3) * kWinTile;                      // doubles between the two stages
    constexpr unsigned FULL = 0xffffffffu;
    const int nd = p.nd;
    const WinSmem L = win_smem_layout(nd, NF);
    double* tiles = reinterpret_cast<double*>(win_raw + L.off_tiles);   // [stage][q | u | aa8 | nc][rows][8]
    int* keys = reinterpret_cast<int*>(win_raw + L.off_keys);           // [8][P], hemispheric row index
    int* PE = reinterpret_cast<int*>(win_raw + L.off_pe);               // [8][P]: PE[x] = max key of rows < x
    int* PP = reinterpret_cast<int*>(win_raw + L.off_pp);               // [8][P]: PP[x] = min key of rows >= x
    double* Qh = reinterpret_cast<double*>(win_raw + L.off_q);          // [2][nd] thresholds of a level
    double* Ur = reinterpret_cast<double*>(win_raw + L.off_u);          // [2][nd] uref of a level
    double* abcs = reinterpret_cast<double*>(win_raw + L.off_cs);       // [nd] a dphi cos(phi_j)
    uint64_t* bars = reinterpret_cast<uint64_t*>(win_raw + L.off_bar);  // [2] "tile of this stage has landed"

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int c = lane & 7, r = lane >> 3;
    const int prob = blockIdx.z, b = prob / p.nhemi, h = prob % p.nhemi;
    const WinFrame fr = p.fr[h];
    const int i0 = blockIdx.x * kWinCols;
    const bool colok = (i0 + c) < p.imax;
    const int kbeg = p.k_first + blockIdx.y * p.k_per_cta;
    const int kend = min(kbeg + p.k_per_cta, p.k_first + p.k_count);    // exclusive
    const int nlev = kend - kbeg;
    if (nlev <= 0) return;
    const unsigned box_bytes = (unsigned)(p.boxrows * kWinCols * 8);

    auto issue = [&](int li) {   // one elected thread: all boxes of level kbeg + li into stage li & 1
        const int s = li & 1, k = kbeg + li;
        win_mbar_expect(&bars[s], box_bytes * p.nbox * NF);
        for (int bx = 0; bx < p.nbox; ++bx) {
            const int row = fr.in_row0 + bx * p.boxrows;
            double* dst = tiles + (size_t)s * SD + (size_t)bx * p.boxrows * kWinCols;
            win_tma_load(dst, &tm_q, &bars[s], i0, row, k - 1, b);
            win_tma_load(dst + TU, &tm_u, &bars[s], i0, row, k - 1, b);
            if (HAS_NC) win_tma_load(dst + TN, &tm_n, &bars[s], i0, row, k - 1, b);
        }
    };
    // thresholds of hemispheric row tid + 1: per-thread source pointers, one level stride
    const bool trow = tid < nd && tid + 1 >= p.jstart && tid + 1 <= p.jend;
    const double* qsrc = trow ? p.qref + (size_t)b * p.q_batch + p.q_row0[h] + p.q_rstep[h] * tid : nullptr;
    const double* usrc = (trow && p.uref) ? p.uref + (size_t)b * p.u_batch + p.u_row0[h] + p.u_rstep[h] * tid : nullptr;
    const double qsg = p.q_sg[h];

    if (tid == 0) {
        win_mbar_init(&bars[0], 1);
        win_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        issue(0);
        if (nlev > 1) issue(1);
    }
    for (int t = tid; t < nd * kWinCols; t += NT) {   // a dphi cos(phi) of every tile row
        const int m = t >> 3;
        const double v = p.ab * p.cosphi[fr.rev ? nd - m : m + 1];
        tiles[(HAS_NC ? 3 : 2) * kWinTile + t] = v;
        if (!HAS_NC) tiles[5 * kWinTile + t] = v;
    }
    for (int t = nd + tid; t < kWinSpan * 32; t += NT)   // rows beyond the column: neutral for the minima
#pragma unroll
        for (int cc = 0; cc < kWinCols; ++cc) keys[cc * P + t] = INT_MAX;
    if (tid < nd) {
        abcs[tid] = p.ab * p.cosphi[tid + 1];
        Qh[tid] = qsrc ? __fma_rn(qsg, qsrc[(size_t)(kbeg - 1) * p.q_kstride], 0.0) : 0.0;
        Ur[tid] = usrc ? usrc[(size_t)(kbeg - 1) * p.u_kstride] : 0.0;
    }
    if (tid < kWinCols) { PE[tid * P] = INT_MIN; PP[tid * P + nd] = INT_MAX; }
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

    // accumulator file in tensor memory (MODE 1): this thread's columns of group gi start at tacc + gi * TCOL
    constexpr int TCOL = HAS_NC ? 10 : 8;
    constexpr unsigned TMEM_COLS = (NW / 4) * GPW * TCOL <= 256 ? 256 : 512;
    uint32_t tacc = 0;
    if (MODE == 1) {
        uint32_t* tbase = reinterpret_cast<uint32_t*>(bars + 2);
        if (w == 0) tmem_alloc(tbase, TMEM_COLS);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tacc = *tbase + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * GPW * TCOL);
        const double zero[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            tmem_st8(tacc + gi * TCOL, zero);
            if (HAS_NC) tmem_st2(tacc + gi * TCOL + 8, 0.0);
        }
        tmem_wait_st();
    }
    unsigned cnt_c = 0, cnt_a = 0;
    // this lane's element of group 0 of its warp; groups of a warp are NW groups = 4 NW rows apart.  What a lane does
    // for a group does not depend on the level: one bit per group.
    const int j00 = 4 * w + r;
    const int m00 = fr.rev ? nd - 1 - j00 : j00;
    const int mstep = fr.rev ? -4 * NW : 4 * NW;                      // tile rows between consecutive groups of a warp
    unsigned validm = 0, actm = 0, storem = 0;
#pragma unroll
    for (int gi = 0; gi < GPW; ++gi) {
        const int j0 = j00 + 4 * NW * gi;
        const bool valid = j0 < nd;
        if (valid) validm |= 1u << gi;
        if (valid && colok && j0 + 1 >= p.jstart && j0 + 1 <= p.jend) actm |= 1u << gi;
        if (valid && colok && !(fr.skip_first && j0 == 0)) storem |= 1u << gi;
    }
    asm volatile("" : "+r"(validm), "+r"(actm), "+r"(storem));          // keep the bits; do not re-derive them per level
    const int ngroups = (nd - 4 * w + 4 * NW - 1) / (4 * NW);            // groups of this warp that exist (warp-uniform)
    // Everything the level walk needs is held as 32-bit shared addresses / 64-bit global pointers that the compiler
    // must treat as opaque values (it otherwise re-derives them from the kernel parameters in every group).
    uint32_t stepb = fr.rev ? 64u : 0xffffffc0u;                         // bytes per row step toward the equator
    uint32_t gbytes = (uint32_t)(mstep * kWinCols * 8);                  // bytes between consecutive groups of a warp
    uint32_t a_t0 = win_smem(tiles) + (uint32_t)((m00 * kWinCols + c) * 8);   // this lane's element of group 0, stage 0
    uint32_t a_key = win_smem(keys) + (uint32_t)((c * P + j00) * 4);     // ... its key; PE / PP follow at fixed offsets
    uint32_t a_tab = win_smem(Qh) + (uint32_t)(j00 * 8);                 // ... its threshold; Ur / abcs at fixed offsets
    const uint32_t o_ur = (uint32_t)((Ur - Qh) * 8), o_cs = (uint32_t)((abcs - Qh) * 8);
    double* out_k = (MODE == 1 ? p.lwa + (size_t)b * p.lwa_batch + (size_t)(kbeg - 1) * p.lwa_plane
                               : p.a1 + (size_t)b * p.out_batch + (size_t)(kbeg - 1) * p.out_plane) +
                    (size_t)(fr.out_row0 + m00) * p.out_pitch + i0 + c;
    long long out_lev = (long long)((MODE == 1) ? p.lwa_plane : p.out_plane);
    long long out_gstep = (long long)mstep * p.out_pitch;
    asm volatile("" : "+r"(stepb), "+r"(gbytes), "+r"(a_t0), "+r"(a_key), "+r"(a_tab), "+l"(out_k), "+l"(out_lev), "+l"(out_gstep));
    constexpr int KPE = kWinCols * P * 4, KPP = 2 * kWinCols * P * 4;    // PE, PP relative to keys (bytes)
    constexpr int GK = 4 * NW * 4, GT = 4 * NW * 8;                      // key / table bytes between groups of a warp
    const double sgn = fr.sg;

    double wd = (MODE == 1) ? p.vw[kbeg - 1] * p.dc : 0.0;
    for (int li = 0; li < nlev; ++li) {
        const int s = li & 1, k = kbeg + li;
        uint32_t a_t = a_t0 + (uint32_t)(s * SD * 8);                     // this lane's element of group 0 in this stage
        uint32_t a_q = a_tab + (uint32_t)(s * nd * 8);                    // table buffers alternate like the tile stages
        asm volatile("" : "+r"(a_t), "+r"(a_q));
        const uint32_t TAb = (uint32_t)((HAS_NC ? (s ? -kWinTile : 3 * kWinTile) : 2 * kWinTile) * 8);   // aa8 from q
        // thresholds and weight of the next level travel while this one is computed
        const bool more = li + 1 < nlev;
        double Qn = 0.0, Un = 0.0, wdn = 0.0;
        if (more && qsrc) Qn = qsrc[(size_t)k * p.q_kstride];
        if (more && usrc) Un = usrc[(size_t)k * p.u_kstride];
        if (MODE == 1 && more) wdn = p.vw[k];
        win_mbar_wait(&bars[s], (unsigned)((li >> 1) & 1));

        // ---- phase 1: hemispheric-frame PV (sign, -0 -> +0) and its keys ---------------------------
        {
            double q[GPW];
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) q[gi] = ((validm >> gi) & 1) ? lds64(a_t + gi * gbytes) : 0.0;
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi)
                if ((validm >> gi) & 1) {
                    const double qh = __fma_rn(sgn, q[gi], 0.0);
                    if (sgn != 1.0) sts64(a_t + gi * gbytes, qh);
                    sts32(a_key + gi * GK, win_key(qh));
                }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // see the TMA issue below
        __syncthreads();
        // Every warp has finished the previous level: its stage is free.  The load of level li + 1 has the rest of
        // this level to land.
        if (tid == 0 && li >= 1 && more) issue(li + 1);
        // ---- phase 2: the two monotone bound arrays of every column, one warp per (column, direction) -
        {
            int incl[kWinSpan];
            if (w < 8) {   // PE[x] = max key over rows < x  (PE[0] = -inf); rows beyond nd do not reach entries <= nd
                const int* kp = keys + w * P + kWinSpan * lane;
                int run = INT_MIN;
#pragma unroll
                for (int t = 0; t < kWinSpan; ++t) { run = max(run, kp[t]); incl[t] = run; }
                int ex = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int n = __shfl_up_sync(FULL, ex, d);
                    if (lane >= d) ex = max(ex, n);
                }
                int off = __shfl_up_sync(FULL, ex, 1);
                if (lane == 0) off = INT_MIN;
                int* op = PE + w * P + kWinSpan * lane + 1;
#pragma unroll
                for (int t = 0; t < kWinSpan; ++t) op[t] = max(off, incl[t]);
            } else if (w < 16) {       // PP[x] = min key over rows >= x  (PP[nd] = +inf; rows beyond nd hold +inf keys)
                const int* kp = keys + (w - 8) * P + kWinSpan * lane;
                int run = INT_MAX;
#pragma unroll
                for (int t = kWinSpan - 1; t >= 0; --t) { run = min(run, kp[t]); incl[t] = run; }
                int ex = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int n = __shfl_down_sync(FULL, ex, d);
                    if (lane + d < 32) ex = min(ex, n);
                }
                int off = __shfl_down_sync(FULL, ex, 1);
                if (lane == 31) off = INT_MAX;
                int* op = PP + (w - 8) * P + kWinSpan * lane;
#pragma unroll
                for (int t = 0; t < kWinSpan; ++t) op[t] = min(off, incl[t]);
            }
        }
        __syncthreads();
        // ---- main: the two windowed sums of every threshold row ----------------------------------------
        double* o = out_k;
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            if (gi < ngroups) {   // warp-uniform
                const bool act = (actm >> gi) & 1;
                const uint32_t at = a_t + gi * gbytes, ak = a_key + gi * GK, aq = a_q + gi * GT;
                const double Qj = lds64(aq);
                const int kq = win_key(Qj);
                const int kqc = max(kq, INT_MIN + 1);   // "may exceed Q_j":      key >= kqc  (never true at PE[0])
                const int kqa = min(kq, INT_MAX - 1);   // "may not exceed Q_j":  key <= kqa  (never true at PP[nd])
                double A1 = 0.0, A2 = 0.0, B = 0.0, N = 0.0;
                {   // cyclonic side: rows j0-1, j0-2, ... while some row at or beyond may exceed Q_j
                    uint32_t t = at, pk = ak;
                    bool mine = act && (lds32o<KPE>(pk) >= kqc);
                    while (__any_sync(FULL, mine)) {
                        if (mine) { t += stepb; pk -= 4; }     // finished lanes keep re-reading their last row
                        const double qv = lds64(t), uv = lds64o<TU * 8>(t), av = lds64(t + TAb);
                        const double nv = HAS_NC ? lds64o<TN * 8>(t) : 0.0;
                        const int pe = lds32o<KPE>(pk);
                        win_pair<false, HAS_NC, COUNT>(qv, Qj, av, uv, nv, mine, A1, B, N, cnt_c);
                        mine = mine && (pe >= kqc);
                    }
                }
                {   // anticyclonic side: rows j0, j0+1, ... while some row at or beyond may not exceed Q_j
                    uint32_t t = at, pk = ak;
                    bool mine = act && (lds32o<KPP>(pk) <= kqa);
                    while (__any_sync(FULL, mine)) {
                        const double qv = lds64(t), uv = lds64o<TU * 8>(t), av = lds64(t + TAb);
                        const double nv = HAS_NC ? lds64o<TN * 8>(t) : 0.0;
                        const int pp = lds32o<KPP + 4>(pk);
                        win_pair<true, HAS_NC, COUNT>(qv, Qj, av, uv, nv, mine, A2, B, N, cnt_a);
                        mine = mine && (pp <= kqa);
                        if (mine) { t -= stepb; pk += 4; }
                    }
                }
                {
                    // a1 = A1, a2 = A2;  ua2 = a dphi cos(phi_j) B - uref_j (a1 + a2)   (compute_flux_dirinv.f90:96-113)
                    // (lanes past the last row and inactive rows carry A1 = A2 = B = N = 0; they store nothing)
                    const double lw = A1 + A2;
                    const double x1 = lds64(aq + o_ur) * lw;
                    const double u2 = __fma_rn(lds64(a_tab + o_cs + gi * GT), B, -x1);
                    const bool store = (storem >> gi) & 1;
                    if (MODE == 1) {
                        double a[4];
                        tmem_ld8(tacc + gi * TCOL, a);
                        a[0] = __fma_rn(A1, wd, a[0]);
                        a[1] = __fma_rn(A2, wd, a[1]);
                        a[2] = __fma_rn(x1, wd, a[2]);
                        a[3] = __fma_rn(u2, wd, a[3]);
                        tmem_st8(tacc + gi * TCOL, a);
                        if (HAS_NC) {
                            double an;
                            tmem_ld2(tacc + gi * TCOL + 8, an);
                            tmem_st2(tacc + gi * TCOL + 8, __fma_rn(N, wd, an));
                        }
                        if (store) st_stream(o, fabs(lw));
                    } else if (store) {
                        o[0] = A1;
                        o[p.a2 - p.a1] = A2;
                        if (p.u2) o[p.u2 - p.a1] = u2;
                        if (HAS_NC && p.ncf) o[p.ncf - p.a1] = N;
                    }
                    o += out_gstep;
                }
            }
        }
        out_k += out_lev;
        // next level's thresholds into the other table buffer (its last readers finished a level ago)
        if (more && tid < nd) {
            Qh[(s ^ 1) * nd + tid] = __fma_rn(qsg, Qn, 0.0);
            Ur[(s ^ 1) * nd + tid] = Un;
        }
        wd = wdn * p.dc;
        if (MODE == 1) tmem_wait_st();
    }

    if (MODE == 1) {
        double* o2 = p.out2d + (size_t)b * p.o2_batch + (size_t)(fr.out_row0 + m00) * p.out_pitch + i0 + c;
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            double a[4], an = 0.0;
            tmem_ld8(tacc + gi * TCOL, a);
            if (HAS_NC) tmem_ld2(tacc + gi * TCOL + 8, an);
            if ((storem >> gi) & 1) {
                double* o = o2 + (long long)gi * out_gstep;
                o[(size_t)p.slot_a1 * p.o2_plane] = a[0];
                o[(size_t)p.slot_a2 * p.o2_plane] = a[1];
                o[(size_t)p.slot_lwa * p.o2_plane] = a[0] + a[1];
                o[(size_t)p.slot_ua1 * p.o2_plane] = a[2];
                o[(size_t)p.slot_ua2 * p.o2_plane] = a[3];
                if (HAS_NC) o[(size_t)p.slot_ncf * p.o2_plane] = (fr.sg < 0.0) ? -an : an;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (w == 0) tmem_dealloc(*reinterpret_cast<uint32_t*>(bars + 2), TMEM_COLS);
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
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*WinEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WinEncodeFn win_encoder() {
    static WinEncodeFn fn = []() -> WinEncodeFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        return (WinEncodeFn)f;
    }();
    return fn;
}

// fields [batch][levels][rows][imax] (longitude fastest) as a 4-D tensor with (8 x boxrows) boxes
static int win_tensor_map(CUtensorMap* tm, const double* base, int imax, int rows, int levels, int batch, int boxrows) {
    WinEncodeFn enc = win_encoder();
    if (!enc) return FALWA_ERR_ARG;
    cuuint64_t dims[4] = {(cuuint64_t)imax, (cuuint64_t)rows, (cuuint64_t)levels, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)imax * 8, (cuuint64_t)imax * rows * 8, (cuuint64_t)imax * rows * levels * 8};
    cuuint32_t box[4] = {(cuuint32_t)kWinCols, (cuuint32_t)boxrows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : FALWA_ERR_ARG;
}

// the window kernel needs TMA-compatible fields (16-byte aligned base and row pitch) and at most 384 hemispheric rows
static bool win_supported(int imax, int nd, const void* pv, const void* uu, const void* nc) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    return nd >= 4 && nd <= kWinMaxRows && (imax % 2 == 0) && al(pv) && al(uu) && (!nc || al(nc)) && win_encoder() != nullptr;
}

template <int MODE, int NW>
static int win_launch_mode(const CUtensorMap& tq, const CUtensorMap& tu, const CUtensorMap& tn, const WinArgs& a, bool has_nc,
                           dim3 grid, size_t smem, cudaStream_t st, const int* sel, int sel_want) {
    auto go = [&](auto kern) -> int {
        FALWA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NW * 32, smem, st>>>(tq, tu, tn, a, sel, sel_want);
        return 0;
    };
    const bool count = a.mask_count != nullptr;
    if (has_nc) return count ? go(lwa_window_kernel<MODE, true, true, NW>) : go(lwa_window_kernel<MODE, true, false, NW>);
    return count ? go(lwa_window_kernel<MODE, false, true, NW>) : go(lwa_window_kernel<MODE, false, false, NW>);
}

// pv / uu / nc: [batch][kmax][in_rows][imax]; a.k_first / k_count / frames / thresholds / outputs filled by the caller
static int lwa_window_launch(WinArgs a, int mode, const double* pv, const double* uu, const double* nc, int in_rows,
                             int batch, int nprob, cudaStream_t st, const int* sel = nullptr, int sel_want = 0) {
    const int nbox = (a.nd + 255) / 256;
    int boxrows = (a.nd + nbox - 1) / nbox;
    boxrows += boxrows & 1;
    a.nbox = nbox; a.boxrows = boxrows;
    CUtensorMap tq, tu, tn;
    int rc;
    if ((rc = win_tensor_map(&tq, pv, a.imax, in_rows, a.kmax, batch, boxrows))) return rc;
    if ((rc = win_tensor_map(&tu, uu, a.imax, in_rows, a.kmax, batch, boxrows))) return rc;
    if ((rc = win_tensor_map(&tn, nc ? nc : pv, a.imax, in_rows, a.kmax, batch, boxrows))) return rc;
    const int tiles = ceil_div(a.imax, kWinCols);
    if (mode == 1) a.k_per_cta = a.k_count;
    else a.k_per_cta = std::max(1, std::min(a.k_count, ceil_div((long long)a.k_count * tiles * nprob, 148 * 8)));
    const int chunks = ceil_div(a.k_count, a.k_per_cta);
    const size_t smem = win_smem_layout(a.nd, nc ? 3 : 2).total;
    dim3 grid(tiles, chunks, nprob);
    FALWA_KERNEL("lwa_window_kernel", st);
    static const int nw = tune_int("FALWA_WIN_WARPS", 24);
    if (mode == 1 && nw == 24) rc = win_launch_mode<1, 24>(tq, tu, tn, a, nc != nullptr, grid, smem, st, sel, sel_want);
    else if (mode == 1 && nw == 32) rc = win_launch_mode<1, 32>(tq, tu, tn, a, nc != nullptr, grid, smem, st, sel, sel_want);
    else rc = mode == 1 ? win_launch_mode<1, 16>(tq, tu, tn, a, nc != nullptr, grid, smem, st, sel, sel_want)
                        : win_launch_mode<0, 16>(tq, tu, tn, a, nc != nullptr, grid, smem, st, sel, sel_want);
    if (rc) return rc;
    FALWA_LAUNCH_CHECK();
    return 0;
}

}  // namespace falwa
