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
    const double *pv, *uu, *nc; int in_rows; // the tiled fields [batch][kmax][in_rows][imax] (cp.async loader)
    int dbg;                                // timing experiments only (FALWA_WIN_DEBUG): 1 = no window loops, 2 = no scans
};

__device__ __forceinline__ uint32_t win_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void win_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(win_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void win_mbar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(win_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void win_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(win_smem(bar)) : "memory");
}
__device__ __forceinline__ void win_mbar_wait(uint64_t* bar, unsigned parity) {
    // try_wait with a suspend-time hint parks the warp in hardware; a failed attempt backs off before the next one, so
    // that waiting warps do not take issue slots from the warps they are waiting for
    unsigned ok = 0;
    const uint32_t a = win_smem(bar);
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(2000u) : "memory");
        if (ok) break;
        __nanosleep(200);
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
// Keys and bound arrays first, then the tile stages [q | u (| nc)] — three stages of 368-row slots when the column fits
// and there is no ncforce tile (TMA runs two levels ahead of the compute warps), else two stages of 384-row slots — then
// the per-level tables.  Lanes of the last group that lie past the last row compute on whatever sits at their
// (in-bounds, thanks to this order and the tail pad) addresses and never store.  The bound arrays PE | PP exist twice
// (two levels in flight) unless the ncforce tile takes the room.
__host__ __device__ constexpr int win_tile_rows(int nstage) { return nstage == 3 ? 368 : kWinMaxRows; }
struct WinSmem {
    size_t off_keys, off_pe, off_tiles, off_q, off_u, off_cs, off_bar, total;
};
__host__ __device__ inline WinSmem win_smem_layout(int nd, int nfields, int nstage) {
    WinSmem s;
    const int nbound = nfields == 3 ? 1 : 2;
    size_t o = 0;
    s.off_keys = o; o += (size_t)kWinCols * kWinKeyPitch * 4;
    s.off_pe = o; o += (size_t)nbound * 2 * kWinCols * kWinKeyPitch * 4;
    s.off_tiles = o; o += (size_t)nstage * nfields * win_tile_rows(nstage) * kWinCols * 8;
    s.off_q = o; o += (size_t)nstage * nd * 8;
    s.off_u = o; o += (size_t)nstage * nd * 8;
    s.off_cs = o; o += (size_t)nd * 8;
    s.off_bar = o; o += 128;
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
// issue now, use later: the values may only be read after tmem_ld8_wait (which ties the registers to the wait)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_wait(uint32_t (&v)[8], double (&a)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
#pragma unroll
    for (int x = 0; x < 4; ++x) a[x] = __hiloint2double((int)v[2 * x + 1], (int)v[2 * x]);
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

// Roles.  NW compute warps own the 4-row groups (group g belongs to warp g % NW) and do nothing but the windowed sums;
// kWinScan scan warps prepare each level for them — the hemispheric-frame PV and its keys, the two bound arrays of
// every column, the thresholds of the level — one level ahead, and one of their threads issues the TMA loads.  The
// hand-over is three mbarriers per stage: full (the tile has landed: TMA transaction count), ready (keys, bounds and
// thresholds of the level are in place: one arrival per scan warp) and done (a compute warp has finished the level: one
// arrival per compute warp; the stage and the bound buffer may be refilled).  No warp ever waits at a CTA-wide barrier
// inside the level walk.
constexpr int kWinScan = 8;   // one per column of the tile in the bound scans

template <int MODE, bool HAS_NC, bool COUNT, int NW, int NSTAGE, bool TMA>
__global__ void __launch_bounds__((NW + kWinScan) * 32, 1)
lwa_window_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_u,
                  const __grid_constant__ CUtensorMap tm_n, const WinArgs p) {
    extern __shared__ __align__(1024) unsigned char win_raw[];
    constexpr int NT = (NW + kWinScan) * 32;
    constexpr int GPW = kWinMaxRows / (4 * NW);                          // 4-row groups a compute warp owns
    constexpr int NF = HAS_NC ? 3 : 2;
    constexpr int NB = HAS_NC ? 1 : 2;                                   // bound buffers (win_smem_layout)
    constexpr int P = kWinKeyPitch;
    constexpr int TILE = win_tile_rows(NSTAGE) * kWinCols;               // doubles of one (rows x 8) tile slot
    constexpr int TU = TILE, TN = 2 * TILE;                              // u, nc relative to q (doubles)
    constexpr int SD = NF * TILE;                                        // doubles between consecutive stages
    constexpr int KB = 2 * kWinCols * P;                                 // ints of one (PE | PP) buffer
    constexpr unsigned FULL = 0xffffffffu;
    const int nd = p.nd;
    const WinSmem L = win_smem_layout(nd, NF, NSTAGE);
    double* tiles = reinterpret_cast<double*>(win_raw + L.off_tiles);   // [stage][q | u | nc][rows][8]
    int* keys = reinterpret_cast<int*>(win_raw + L.off_keys);           // [8][P], hemispheric row index
    int* PEPP = reinterpret_cast<int*>(win_raw + L.off_pe);             // [NB][PE | PP][8][P]
    double* Qh = reinterpret_cast<double*>(win_raw + L.off_q);          // [NSTAGE][nd] thresholds of a level
    double* Ur = reinterpret_cast<double*>(win_raw + L.off_u);          // [NSTAGE][nd] uref of a level
    double* abcs = reinterpret_cast<double*>(win_raw + L.off_cs);       // [nd] a dphi cos(phi_j): the LWA weight of a row
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(win_raw + L.off_bar);   // [NSTAGE]
    uint64_t* bar_ready = bar_full + NSTAGE;                                  // [NSTAGE]
    uint64_t* bar_done = bar_full + 2 * NSTAGE;                               // [NSTAGE]
    uint32_t* tbase = reinterpret_cast<uint32_t*>(bar_full + 3 * NSTAGE);
    unsigned* left = reinterpret_cast<unsigned*>(tbase + 2);                  // [NSTAGE] compute warps that left a level

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
    const double sgn = fr.sg;
    // completion of level L by the compute warps: the (L / NSTAGE)-th phase of done[L % NSTAGE]
    auto wait_done = [&](int lev) { win_mbar_wait(&bar_done[lev % NSTAGE], (unsigned)((lev / NSTAGE) & 1)); };
    const unsigned box_bytes = (unsigned)(p.boxrows * kWinCols * 8);
    auto issue = [&](int li) {   // one thread: all boxes of level kbeg + li into stage li % NSTAGE
        const int s = li % NSTAGE, k = kbeg + li;
        win_mbar_expect(&bar_full[s], box_bytes * p.nbox * NF);
        for (int bx = 0; bx < p.nbox; ++bx) {
            const int row = fr.in_row0 + bx * p.boxrows;
            double* dst = tiles + (size_t)s * SD + (size_t)bx * p.boxrows * kWinCols;
            win_tma_load(dst, &tm_q, &bar_full[s], i0, row, k - 1, b);
            win_tma_load(dst + TU, &tm_u, &bar_full[s], i0, row, k - 1, b);
            if (HAS_NC) win_tma_load(dst + TN, &tm_n, &bar_full[s], i0, row, k - 1, b);
        }
    };

    // ---- set-up by everybody -------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            win_mbar_init(&bar_full[s], 1);
            win_mbar_init(&bar_ready[s], kWinScan);
            win_mbar_init(&bar_done[s], NW);
            left[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = nd + tid; t < kWinSpan * 32; t += NT)   // rows beyond the column: neutral for the minima
#pragma unroll
        for (int cc = 0; cc < kWinCols; ++cc) keys[cc * P + t] = INT_MAX;
    if (tid < nd) abcs[tid] = p.ab * p.cosphi[tid + 1];
    if (tid < kWinCols * NB) {   // sentinels of both bound buffers: PE[0] = -inf, PP[nd] = +inf
        int* base = PEPP + (tid / kWinCols) * KB + (tid % kWinCols) * P;
        base[0] = INT_MIN;
        base[kWinCols * P + nd] = INT_MAX;
    }
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
    // accumulator file in tensor memory (MODE 1): a compute thread's columns of group gi start at tacc + gi * TCOL
    constexpr int TCOL = HAS_NC ? 10 : 8;
    constexpr unsigned TMEM_COLS = ((NW + 3) / 4) * GPW * TCOL <= 256 ? 256 : 512;
    if (MODE == 1 && w == 0) tmem_alloc(tbase, TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (TMA && tid == NW * 32)   // the first levels; later ones are requested by the compute warp that frees a stage
        for (int li = 0; li < NSTAGE && li < nlev; ++li) issue(li);

    if (w >= NW) {
        // =====================================================================================================
        // scan warps: tile -> keys -> bounds, thresholds, TMA issue
        // =====================================================================================================
        const int ws = w - NW, stid = tid - NW * 32;
        const double qsg = p.q_sg[h];
        constexpr int TPT = (kWinMaxRows + kWinScan * 32 - 1) / (kWinScan * 32);   // threshold rows per scan thread
        // everything a level needs before the compute warps may enter it
        auto prepare = [&](int li) {
            const int s = li % NSTAGE, k = kbeg + li;
            double* const tq = tiles + (size_t)s * SD;
            int* const PE = PEPP + (li % NB) * KB;
            int* const PP = PE + kWinCols * P;
            // thresholds of this level: requested now, stored once the buffers are known to be free
            double Qv[TPT], Uv[TPT];
#pragma unroll
            for (int x = 0; x < TPT; ++x) {
                const int j0 = stid + x * kWinScan * 32;
                Qv[x] = Uv[x] = 0.0;
                if (j0 < nd && j0 + 1 >= p.jstart && j0 + 1 <= p.jend) {
                    Qv[x] = p.qref[(size_t)b * p.q_batch + (size_t)(k - 1) * p.q_kstride + p.q_row0[h] + p.q_rstep[h] * j0];
                    if (p.uref) Uv[x] = p.uref[(size_t)b * p.u_batch + (size_t)(k - 1) * p.u_kstride + p.u_row0[h] + p.u_rstep[h] * j0];
                }
            }
            if (TMA) win_mbar_wait(&bar_full[s], (unsigned)((li / NSTAGE) & 1));
            else {   // this thread's copies of the level have landed; with the barrier, everybody's
                asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(kWinScan * 32) : "memory");
            }
            // phase 1: hemispheric-frame PV (sign, -0 -> +0) and its keys; four groups of a warp in flight at a time
            for (int g0 = ws; 4 * g0 < nd && p.dbg != 2; g0 += 4 * kWinScan) {
                double qv[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int j0 = 4 * (g0 + x * kWinScan) + r;
                    qv[x] = (j0 < nd) ? tq[(fr.rev ? nd - 1 - j0 : j0) * kWinCols + c] : 0.0;
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int j0 = 4 * (g0 + x * kWinScan) + r;
                    if (j0 < nd) {
                        const double qh = __fma_rn(sgn, qv[x], 0.0);
                        if (sgn != 1.0) tq[(fr.rev ? nd - 1 - j0 : j0) * kWinCols + c] = qh;
                        keys[c * P + j0] = win_key(qh);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile is refilled by the async proxy later
            asm volatile("bar.sync 1, %0;" ::"n"(kWinScan * 32) : "memory");
            // the bound buffer of this level was read by the compute warps NB levels ago (the table buffer NSTAGE ago)
            if (TMA ? li >= NB : (NB == 1 && li >= 1)) wait_done(li - NB);
#pragma unroll
            for (int x = 0; x < TPT; ++x) {
                const int j0 = stid + x * kWinScan * 32;
                if (j0 < nd) { Qh[s * nd + j0] = __fma_rn(qsg, Qv[x], 0.0); Ur[s * nd + j0] = Uv[x]; }
            }
            // phase 2: the two monotone bound arrays of column ws, from one read of its keys: a lane takes 13
            // consecutive rows, PE[x] = max key over rows < x (rows beyond nd do not reach entries <= nd),
            // PP[x] = min key over rows >= x (rows beyond nd hold +inf keys)
            {
                static_assert(kWinScan == kWinCols, "one scan warp per column");
                const int* kp = keys + ws * P + kWinSpan * lane;
                int kv[kWinSpan], mx[kWinSpan], mn[kWinSpan];
#pragma unroll
                for (int t = 0; t < kWinSpan; ++t) kv[t] = kp[t];
                int rmx = INT_MIN, rmn = INT_MAX;
#pragma unroll
                for (int t = 0; t < kWinSpan; ++t) {
                    rmx = max(rmx, kv[t]); mx[t] = rmx;
                    rmn = min(rmn, kv[kWinSpan - 1 - t]); mn[kWinSpan - 1 - t] = rmn;
                }
                int ex = rmx, en = rmn;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int a = __shfl_up_sync(FULL, ex, d), bb = __shfl_down_sync(FULL, en, d);
                    if (lane >= d) ex = max(ex, a);
                    if (lane + d < 32) en = min(en, bb);
                }
                int offx = __shfl_up_sync(FULL, ex, 1), offn = __shfl_down_sync(FULL, en, 1);
                if (lane == 0) offx = INT_MIN;
                if (lane == 31) offn = INT_MAX;
                int* opx = PE + ws * P + kWinSpan * lane + 1;
                int* opn = PP + ws * P + kWinSpan * lane;
#pragma unroll
                for (int t = 0; t < kWinSpan; ++t) { opx[t] = max(offx, mx[t]); opn[t] = min(offn, mn[t]); }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kWinScan * 32) : "memory");   // keys are rewritten next level
            if (lane == 0) win_mbar_arrive(&bar_ready[s]);
        };
        if (TMA) {
            for (int li = 0; li < nlev; ++li) prepare(li);
        } else {
            // The scan warps fetch the tiles themselves: 16-byte cp.async (LDGSTS), four lanes per 64-byte row segment,
            // NSTAGE - 1 levels ahead of the compute warps.  (A TMA box of this shape is one 64-byte request per row and
            // the engine serves ~1 row per 7 clocks: 2.6 us per level and SM, more than the level's arithmetic.)
            const int nchunk = nd * 4;                               // 16-byte pieces of one field of the tile
            auto load = [&](int li) {
                const int s = li % NSTAGE, k = kbeg + li;
                const size_t lev = ((size_t)b * p.kmax + (k - 1)) * (size_t)p.in_rows + fr.in_row0;
                const uint32_t dst0 = win_smem(tiles + (size_t)s * SD);
                for (int q = stid; q < nchunk; q += kWinScan * 32) {
                    const int row = q >> 2, col = i0 + (q & 3) * 2;
                    const size_t src = (lev + row) * (size_t)p.imax + col;
                    const unsigned nbytes = col < p.imax ? 16u : 0u;     // past the last longitude: zeros
                    const uint32_t dst = dst0 + (uint32_t)q * 16u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p.pv + (nbytes ? src : 0)), "r"(nbytes) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + TU * 8), "l"(p.uu + (nbytes ? src : 0)), "r"(nbytes) : "memory");
                    if (HAS_NC)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + TN * 8), "l"(p.nc + (nbytes ? src : 0)), "r"(nbytes) : "memory");
                }
            };
            constexpr int LA = NSTAGE - 1;                           // levels the loads run ahead
            for (int l = 0; l < LA; ++l) {
                if (l < nlev) load(l);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            prepare(0);
            for (int t = 0; t < nlev; ++t) {
                if (t >= 1) wait_done(t - 1);        // the stage of level t + LA and the bound buffer of level t + 1 are free
                if (t + LA < nlev) load(t + LA);
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (t + 1 < nlev) prepare(t + 1);
            }
        }
    } else {
        // =====================================================================================================
        // compute warps
        // =====================================================================================================
        uint32_t tacc = 0;
        if (MODE == 1) {
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
        // this lane's element of group 0 of its warp; groups of a warp are NW groups = 4 NW rows apart.  What a lane
        // does for a group does not depend on the level: one bit per group.
        const int j00 = 4 * w + r;
        const int m00 = fr.rev ? nd - 1 - j00 : j00;
        const int mstep = fr.rev ? -4 * NW : 4 * NW;                  // tile rows between consecutive groups of a warp
        unsigned actm = 0, storem = 0;
#pragma unroll
        for (int gi = 0; gi < GPW; ++gi) {
            const int j0 = j00 + 4 * NW * gi;
            const bool valid = j0 < nd;
            if (valid && colok && j0 + 1 >= p.jstart && j0 + 1 <= p.jend) actm |= 1u << gi;
            if (valid && colok && !(fr.skip_first && j0 == 0)) storem |= 1u << gi;
        }
        if (p.dbg == 1) actm = 0;
        asm volatile("" : "+r"(actm), "+r"(storem));                    // keep the bits; do not re-derive them per level
        const int ngroups = (nd - 4 * w + 4 * NW - 1) / (4 * NW);        // groups of this warp that exist (warp-uniform)
        // Everything the level walk needs is held as 32-bit shared addresses / 64-bit global pointers that the compiler
        // must treat as opaque values (it otherwise re-derives them from the kernel parameters in every group).
        uint32_t stepb = fr.rev ? 64u : 0xffffffc0u;                     // bytes per row step toward the equator
        uint32_t gbytes = (uint32_t)(mstep * kWinCols * 8);              // bytes between consecutive groups of a warp
        uint32_t a_t0 = win_smem(tiles) + (uint32_t)((m00 * kWinCols + c) * 8);   // this lane's element of group 0, stage 0
        uint32_t a_pe0 = win_smem(PEPP) + (uint32_t)((c * P + j00) * 4);  // ... its entry of PE (buffer 0); PP follows
        uint32_t a_tab = win_smem(Qh) + (uint32_t)(j00 * 8);             // ... its threshold; Ur / abcs at fixed offsets
        const uint32_t o_ur = (uint32_t)((Ur - Qh) * 8), o_cs = (uint32_t)((abcs - Qh) * 8);
        double* out_k = (MODE == 1 ? p.lwa + (size_t)b * p.lwa_batch + (size_t)(kbeg - 1) * p.lwa_plane
                                   : p.a1 + (size_t)b * p.out_batch + (size_t)(kbeg - 1) * p.out_plane) +
                        (size_t)(fr.out_row0 + m00) * p.out_pitch + i0 + c;
        long long out_lev = (long long)((MODE == 1) ? p.lwa_plane : p.out_plane);
        long long out_gstep = (long long)mstep * p.out_pitch;
        asm volatile("" : "+r"(stepb), "+r"(gbytes), "+r"(a_t0), "+r"(a_pe0), "+r"(a_tab), "+l"(out_k), "+l"(out_lev), "+l"(out_gstep));
        constexpr int KPP = kWinCols * P * 4;                            // PP relative to PE (bytes)
        constexpr int GK = 4 * NW * 4, GT = 4 * NW * 8;                  // key / table bytes between groups of a warp

        for (int li = 0; li < nlev; ++li) {
            const int s = li % NSTAGE, k = kbeg + li;
            uint32_t a_t = a_t0 + (uint32_t)(s * SD * 8);                 // this lane's element of group 0 in this stage
            uint32_t a_q = a_tab + (uint32_t)(s * nd * 8);                // table buffers rotate like the tile stages
            uint32_t a_k = a_pe0 + (uint32_t)((li % NB) * KB * 4);        // ... and the bound buffers
            asm volatile("" : "+r"(a_t), "+r"(a_q), "+r"(a_k));
            const double wd = (MODE == 1) ? p.vw[k - 1] * p.dc : 0.0;
            win_mbar_wait(&bar_ready[s], (unsigned)((li / NSTAGE) & 1)); // keys, bounds and thresholds are in place
            if (TMA) win_mbar_wait(&bar_full[s], (unsigned)((li / NSTAGE) & 1));  // (complete by now: observed for the TMA writes)
            double* o = out_k;
            // the operands a group starts with (threshold, the two first bound keys, uref, LWA weight of the row) are
            // fetched one group ahead, the group's running sums are requested from tensor memory before its windows
            // are walked: the latencies overlap with the arithmetic instead of heading every group
            double Qn = lds64(a_q), urn = lds64(a_q + o_ur), wn = lds64(a_tab + o_cs);
            int pen = lds32(a_k), ppn = lds32o<KPP>(a_k);
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) {
                if (gi < ngroups) {   // warp-uniform
                    const bool act = (actm >> gi) & 1;
                    const uint32_t at = a_t + gi * gbytes, ak = a_k + gi * GK, aq = a_q + gi * GT;
                    const uint32_t aw = a_tab + o_cs + gi * GT;          // the LWA weight a dphi cos(phi) of this row
                    const double Qj = Qn, urv = urn, wv = wn;
                    const int pe0 = pen, pp0 = ppn;
                    uint32_t tv[8];
                    if (MODE == 1) tmem_ld8_issue(tacc + gi * TCOL, tv);
                    if (gi + 1 < GPW && gi + 1 < ngroups) {
                        Qn = lds64(aq + GT); urn = lds64(aq + GT + o_ur); wn = lds64(aw + GT);
                        pen = lds32(ak + GK); ppn = lds32o<KPP>(ak + GK);
                    }
                    const int kq = win_key(Qj);
                    const int kqc = max(kq, INT_MIN + 1);   // "may exceed Q_j":      key >= kqc  (never true at PE[0])
                    const int kqa = min(kq, INT_MAX - 1);   // "may not exceed Q_j":  key <= kqa  (never true at PP[nd])
                    double A1 = 0.0, A2 = 0.0, B = 0.0, N = 0.0;
                    {   // cyclonic side: rows j0-1, j0-2, ... while some row at or beyond may exceed Q_j
                        uint32_t t = at, pk = ak, pa = aw;
                        bool mine = act && (pe0 >= kqc);
                        while (__any_sync(FULL, mine)) {
                            if (mine) { t += stepb; pk -= 4; pa -= 8; }     // finished lanes keep re-reading their last row
                            const double qv = lds64(t), uv = lds64o<TU * 8>(t), av = lds64(pa);
                            const double nv = HAS_NC ? lds64o<TN * 8>(t) : 0.0;
                            const int pe = lds32(pk);
                            win_pair<false, HAS_NC, COUNT>(qv, Qj, av, uv, nv, mine, A1, B, N, cnt_c);
                            mine = mine && (pe >= kqc);
                        }
                    }
                    {   // anticyclonic side: rows j0, j0+1, ... while some row at or beyond may not exceed Q_j
                        uint32_t t = at, pk = ak, pa = aw;
                        bool mine = act && (pp0 <= kqa);
                        while (__any_sync(FULL, mine)) {
                            const double qv = lds64(t), uv = lds64o<TU * 8>(t), av = lds64(pa);
                            const double nv = HAS_NC ? lds64o<TN * 8>(t) : 0.0;
                            const int pp = lds32o<KPP + 4>(pk);
                            win_pair<true, HAS_NC, COUNT>(qv, Qj, av, uv, nv, mine, A2, B, N, cnt_a);
                            mine = mine && (pp <= kqa);
                            if (mine) { t -= stepb; pk += 4; pa += 8; }
                        }
                    }
                    {
                        // a1 = A1, a2 = A2;  ua2 = a dphi cos(phi_j) B - uref_j (a1 + a2)   (compute_flux_dirinv.f90:96-113)
                        // (lanes past the last row and inactive rows carry A1 = A2 = B = N = 0; they store nothing)
                        const double lw = A1 + A2;
                        const double x1 = urv * lw;
                        const double u2 = __fma_rn(wv, B, -x1);
                        const bool store = (storem >> gi) & 1;
                        if (MODE == 1) {
                            double a[4];
                            tmem_ld8_wait(tv, a);
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
            if (MODE == 1) tmem_wait_st();
            // this warp has left the level (TMA variant: the last one to do so refills the stage with level li + NSTAGE)
            if (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                win_mbar_arrive(&bar_done[s]);
                __threadfence_block();
                if (TMA && atomicAdd(&left[s], 1u) == (unsigned)(NW - 1)) {
                    left[s] = 0;
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (li + NSTAGE < nlev) issue(li + NSTAGE);
                }
            }
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
    if (MODE == 1) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (w == 0) tmem_dealloc(*tbase, TMEM_COLS);
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
                           bool three_stages, dim3 grid, size_t smem, cudaStream_t st) {
    auto go = [&](auto kern) -> int {
        FALWA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, (NW + kWinScan) * 32, smem, st>>>(tq, tu, tn, a);
        return 0;
    };
    const bool count = a.mask_count != nullptr;
    if (has_nc) return count ? go(lwa_window_kernel<MODE, true, true, NW, 2, false>) : go(lwa_window_kernel<MODE, true, false, NW, 2, false>);
    if (count) return go(lwa_window_kernel<MODE, false, true, NW, 2, false>);
    if (MODE == 1 && three_stages) {
        static const int tma = tune_int("FALWA_WIN_TMA", 0);   // the TMA-fed variant of the hot instantiation
        return tma ? go(lwa_window_kernel<1, false, false, NW, 3, true>) : go(lwa_window_kernel<1, false, false, NW, 3, false>);
    }
    return go(lwa_window_kernel<MODE, false, false, NW, 2, false>);
}

// pv / uu / nc: [batch][kmax][in_rows][imax]; a.k_first / k_count / frames / thresholds / outputs filled by the caller
static int lwa_window_launch(WinArgs a, int mode, const double* pv, const double* uu, const double* nc, int in_rows,
                             int batch, int nprob, cudaStream_t st) {
    a.dbg = tune_int("FALWA_WIN_DEBUG", 0);
    a.pv = pv; a.uu = uu; a.nc = nc; a.in_rows = in_rows;
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
    // three tile stages (TMA two levels ahead) where the column fits the smaller slots: the hot path
    const bool three = mode == 1 && !nc && !a.mask_count && nbox * boxrows <= win_tile_rows(3) && tune_int("FALWA_WIN_STAGES", 3) == 3;
    const size_t smem = win_smem_layout(a.nd, nc ? 3 : 2, three ? 3 : 2).total;
    dim3 grid(tiles, chunks, nprob);
    FALWA_KERNEL("lwa_window_kernel", st);
    static const int nw = tune_int("FALWA_WIN_WARPS", 16);
    if (mode == 1 && nw == 24) rc = win_launch_mode<1, 24>(tq, tu, tn, a, nc != nullptr, three, grid, smem, st);
    else rc = mode == 1 ? win_launch_mode<1, 16>(tq, tu, tn, a, nc != nullptr, three, grid, smem, st)
                        : win_launch_mode<0, 16>(tq, tu, tn, a, nc != nullptr, false, grid, smem, st);
    if (rc) return rc;
    FALWA_LAUNCH_CHECK();
    return 0;
}

}  // namespace falwa
