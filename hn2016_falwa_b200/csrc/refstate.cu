// Reference-state stage: zonal means, per-level extrema, the equivalent-latitude ("area analysis")
// histogram, Qref by cumulative-area inversion, the NH18 red-black SOR solve and the NHN22
// block-recursion pieces (matrix_b4_inversion / matrix_after_inversion / upward_sweep).
//
// Reference behaviour restated (not ported): f90_modules/compute_reference_states.f90:1-261,
// compute_qref_and_fawa_first.f90:1-176, matrix_b4_inversion.f90:1-88,
// matrix_after_inversion.f90:1-39, upward_sweep.f90:1-92.
//
// Histogram design: one CTA per (level, latitude slab); every thread walks a contiguous longitude
// run of one row and merges consecutive points that fall in the same bin before touching the
// shared-memory histogram (PV is smooth along a latitude circle, so runs are long and shared
// fp64 atomics are rare); CTAs flush their private histogram to global memory with one fp64
// RED per non-empty bin.  Bin *assignment* is exact IEEE arithmetic (sub, div.rn, trunc) and
// therefore identical to the CPU oracle; the bin *sums* differ from it only by summation order.
#include "common.cuh"

namespace falwa {

// ---------------------------------------------------------------------------------------------
// zonal means of up to 3 fields per (row, level) + per-level min/max of up to 2 fields
// zm[f][k][j] ; ext[f][k] = {max, min} as ordered uint64 (init via init_ext_kernel)
// ---------------------------------------------------------------------------------------------
__global__ void init_ext_kernel(unsigned long long* ext, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) ext[t] = (t & 1) ? ~0ull : 0ull;  // even slots: running max (start lowest); odd: running min
}

__global__ void __launch_bounds__(256)
zonal_stats_kernel(int nlon, int nlat, int kmax, const double* __restrict__ f0, const double* __restrict__ f1,
                   const double* __restrict__ f2, const double* __restrict__ f3, double* __restrict__ zm,
                   unsigned long long* __restrict__ ext, int ext_mask) {
    // One warp per (row, level): 8 independent row reductions per CTA, no block-level synchronisation; every lane
    // streams nlon/32 elements of each field with all loads of an unrolled group in flight.
    // f0..f2: fields whose zonal mean is wanted (may be null); extrema are taken of f0 (ext_mask&1) and f3 (ext_mask&2).
    {   // batch: blockIdx.z selects the snapshot (contiguous fields / zm / ext per snapshot)
        const size_t fs = (size_t)blockIdx.z * nlon * nlat * kmax;
        if (f0) f0 += fs;
        if (f1) f1 += fs;
        if (f2) f2 += fs;
        if (f3) f3 += fs;
        zm += (size_t)blockIdx.z * 3 * kmax * nlat;
        if (ext) ext += (size_t)blockIdx.z * 2 * kmax * 2;
    }
    const int lane = threadIdx.x & 31;
    const long long rowid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (rowid >= (long long)nlat * kmax) return;
    const int k = (int)(rowid / nlat), j = (int)(rowid % nlat);
    const size_t row = ((size_t)k * nlat + j) * nlon;
    const double* fs[3] = {f0, f1, f2};
    double s[3] = {0, 0, 0};
    double mx0 = -INFINITY, mn0 = INFINITY, mx3 = -INFINITY, mn3 = INFINITY;
    const bool want3 = f3 && (ext_mask & 2);
#pragma unroll 4
    for (int i = lane; i < nlon; i += 32) {
#pragma unroll
        for (int f = 0; f < 3; ++f)
            if (fs[f]) {
                const double x = ld_stream(&fs[f][row + i]);
                s[f] += x;
                if (f == 0) { mx0 = fmax(mx0, x); mn0 = fmin(mn0, x); }
            }
        if (want3) {
            const double x = ld_stream(&f3[row + i]);
            mx3 = fmax(mx3, x); mn3 = fmin(mn3, x);
        }
    }
#pragma unroll
    for (int f = 0; f < 3; ++f)
        if (fs[f]) {
            const double t = warp_sum(s[f]);
            if (lane == 0) zm[((size_t)f * kmax + k) * nlat + j] = t / (double)nlon;
        }
    if (ext) {
        mx0 = warp_max(mx0); mn0 = warp_min(mn0); mx3 = warp_max(mx3); mn3 = warp_min(mn3);
        if (lane == 0) {
            if (f0 && (ext_mask & 1)) {
                atomicMax(&ext[(0 * kmax + k) * 2 + 0], f64_to_ordered(mx0));
                atomicMin(&ext[(0 * kmax + k) * 2 + 1], f64_to_ordered(mn0));
            }
            if (want3) {
                atomicMax(&ext[(1 * kmax + k) * 2 + 0], f64_to_ordered(mx3));
                atomicMin(&ext[(1 * kmax + k) * 2 + 1], f64_to_ordered(mn3));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Area-analysis histogram of one field, all interior levels.
//   q = sgn * fld ; qmax/qmin from ext (swapped and negated when sgn < 0)
//   ind = 1 + int((qmax - q)/dq), dq = (qmax-qmin)/(nb-1)      (compute_reference_states.f90:76-91)
//   an[ind] += da(j) ; cn[ind] += da(j)*q
// hist layout: [k][2][nb]  (an then cn), zero-initialised by the caller.
// ---------------------------------------------------------------------------------------------
constexpr int kHistThreads = 256;

__global__ void __launch_bounds__(kHistThreads)
area_hist_kernel(int nlon, int nlat, int kmax, int nb, const double* __restrict__ fld,
                 const unsigned long long* __restrict__ ext /* [k][2] of this field */, double sgn,
                 const double* __restrict__ da /* [nlat] */, int rows_per_block, double* __restrict__ hist,
                 int32_t* __restrict__ bins) {
    extern __shared__ double sh[];  // [2][nb]
    const int k = blockIdx.y + 1;
    const int j0 = blockIdx.x * rows_per_block;
    const int j1 = min(nlat, j0 + rows_per_block);
    for (int t = threadIdx.x; t < 2 * nb; t += blockDim.x) sh[t] = 0.0;
    __syncthreads();
    double qmax = ordered_to_f64(ext[k * 2 + 0]), qmin = ordered_to_f64(ext[k * 2 + 1]);
    if (sgn < 0) { const double t = qmax; qmax = -qmin; qmin = -t; }
    const double dq = (qmax - qmin) / (double)(nb - 1);
    // each thread owns a contiguous run of `run` longitudes of one row
    const int run = 8;
    const int runs_per_row = (nlon + run - 1) / run;
    const int total = (j1 - j0) * runs_per_row;
    const size_t lev = (size_t)k * nlat * nlon;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int j = j0 + t / runs_per_row;
        const int i0 = (t % runs_per_row) * run;
        const int i1 = min(nlon, i0 + run);
        const double daj = da[j];
        int cur = -1;
        double sa = 0.0, sc = 0.0;
        for (int i = i0; i < i1; ++i) {
            const double q = sgn * fld[lev + (size_t)j * nlon + i];
            int ind = 1 + (int)((qmax - q) / dq);
            ind = max(1, min(nb, ind));
            if (bins) bins[lev + (size_t)j * nlon + i] = ind;
            if (ind != cur) {
                if (cur > 0) {
                    atomicAdd(&sh[cur - 1], sa);
                    atomicAdd(&sh[nb + cur - 1], sc);
                }
                cur = ind; sa = 0.0; sc = 0.0;
            }
            sa += daj;
            sc += daj * q;
        }
        if (cur > 0) {
            atomicAdd(&sh[cur - 1], sa);
            atomicAdd(&sh[nb + cur - 1], sc);
        }
    }
    __syncthreads();
    double* out = hist + (size_t)k * 2 * nb;
    for (int t = threadIdx.x; t < 2 * nb; t += blockDim.x) {
        const double x = sh[t];
        if (x != 0.0) atomicAdd(&out[t], x);
    }
}

// ---------------------------------------------------------------------------------------------
// Cumulative area / circulation and inversion at the equivalent latitudes.
//   aan(1)=0, aan(nn)=aan(nn-1)+an(nn)   (bin 1's area is never added — SURVEY Q6)
//   row j (1..nrow-1): the unique nn with aan(nn) <= alat(j) < aan(nn+1) (aan is non-decreasing, so
//   "keep the last match" of the reference == the only match), linear interpolation of qn, ccn.
// One CTA per interior level; the scan is done in the reference's order by one thread.
// qout/cout: [k][ldo] profiles (entries not found keep their previous contents, i.e. 0).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
qref_invert_kernel(int kmax, int nb, int nrow, int ldo, const double* __restrict__ hist,
                   const unsigned long long* __restrict__ ext, double sgn, const double* __restrict__ alat,
                   double* __restrict__ qout, double* __restrict__ cout, int set_qmax_row) {
    extern __shared__ double sh[];  // aan[nb], ccn[nb]
    double* aan = sh;
    double* ccn = sh + nb;
    const int k = blockIdx.x + 1;
    const double* an = hist + (size_t)k * 2 * nb;
    const double* cn = an + nb;
    double qmax = ordered_to_f64(ext[k * 2 + 0]), qmin = ordered_to_f64(ext[k * 2 + 1]);
    if (sgn < 0) { const double t = qmax; qmax = -qmin; qmin = -t; }
    const double dq = (qmax - qmin) / (double)(nb - 1);
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        aan[0] = 0.0; ccn[0] = 0.0;
        for (int nn = 1; nn < nb; ++nn) {
            a = a + an[nn]; c = c + cn[nn];
            aan[nn] = a; ccn[nn] = c;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nrow - 1; j += blockDim.x) {
        const double x = alat[j];
        // upper_bound: first index with aan > x
        int lo = 0, hi = nb;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (aan[mid] <= x) lo = mid + 1; else hi = mid;
        }
        const int nn = lo;  // 1-based bin index of the lower edge (0-based lo-1)
        if (nn >= 1 && nn <= nb - 1) {
            const double a0 = aan[nn - 1], a1 = aan[nn];
            const double dd = (x - a0) / (a1 - a0);
            const double qn0 = qmax - dq * (double)(nn - 1), qn1 = qmax - dq * (double)nn;
            if (qout) qout[(size_t)k * ldo + j] = qn0 * (1. - dd) + qn1 * dd;
            if (cout) cout[(size_t)k * ldo + j] = ccn[nn - 1] * (1. - dd) + ccn[nn] * dd;
        }
    }
    if (set_qmax_row && qout && threadIdx.x == 0) qout[(size_t)k * ldo + (nrow - 1)] = qmax;
}

// ---------------------------------------------------------------------------------------------
// cbar (zonal-mean circulation), normalisation of qref, equatorial extrapolation, FAWA.
//   cbar(nd,k)=0; cbar(j,k)=cbar(j+1,k)+0.5*(qbar(j+1,k)+qbar(j,k))*a*dphi*2*pi*a*cos(dphi*(j-0.5))
//   qref(j,:) /= norm(j) for j>=2 ; qref(1,k)=2*qref(2,k)-qref(3,k) (interior k)
// One thread per level.  tables: cosmid[nd] = cos(dphi*(j-0.5)) (j=1..nd-1), norm[nd].
// ---------------------------------------------------------------------------------------------
__global__ void cbar_normalise_kernel(int kmax, int nd, const double* __restrict__ qbar /* [k][ldq] */, int ldq,
                                      double qsgn, int qflip_n /* 0 or nlat: row j -> qflip_n-1-j */, int qoff,
                                      const double* __restrict__ cosmid, const double* __restrict__ norm,
                                      double a, double dphi, double pi, double* __restrict__ qref,
                                      const double* __restrict__ cref, double* __restrict__ cbar,
                                      double* __restrict__ fawa) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kmax) return;
    const bool interior = (k >= 1 && k <= kmax - 2);
    auto qb = [&](int j) -> double {  // j: 0-based hemispheric row
        const int jg = qflip_n ? (qflip_n - 1 - (qoff + j)) : (qoff + j);
        return qsgn * qbar[(size_t)k * ldq + jg];
    };
    if (interior) {
        double c = 0.0;
        cbar[(size_t)k * nd + nd - 1] = 0.0;
        for (int j = nd - 2; j >= 0; --j) {
            c = c + 0.5 * (qb(j + 1) + qb(j)) * a * dphi * 2. * pi * a * cosmid[j];
            cbar[(size_t)k * nd + j] = c;
        }
    } else {
        for (int j = 0; j < nd; ++j) cbar[(size_t)k * nd + j] = 0.0;
    }
    for (int j = 1; j < nd; ++j) qref[(size_t)k * nd + j] = qref[(size_t)k * nd + j] / norm[j];
    if (interior) qref[(size_t)k * nd + 0] = 2. * qref[(size_t)k * nd + 1] - qref[(size_t)k * nd + 2];
    if (fawa)
        for (int j = 0; j < nd; ++j)
            fawa[(size_t)k * nd + j] = (cref[(size_t)k * nd + j] - cbar[(size_t)k * nd + j]) / (2. * pi * a);
}

// hemispheric profile extraction: out[k][j] = sgn * zm[k][map(j)]  (rows nd..jmax of the caller's frame)
__global__ void extract_rows_kernel(int kmax, int nd, int ldz, int off, int flip_n, double sgn,
                                    const double* __restrict__ zm, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kmax * nd) return;
    const int k = t / nd, j = t % nd;
    const int jg = flip_n ? (flip_n - 1 - (off + j)) : (off + j);
    out[t] = sgn * zm[(size_t)k * ldz + jg];
}

// ---------------------------------------------------------------------------------------------
// NH18 SOR (compute_reference_states.f90:135-258).  One CTA per hemisphere solve.
// u lives in shared memory when it fits, otherwise in the global output array.
// Red-black sweeps are order-independent within a colour, so every point gets exactly the value
// the serial reference produces; only the |resid| sum is reduced in a different order.
// tables per row j (1-based rows 1..jd at index j-1): ajk, bjk, fact, uz, cosj, cor ; per level k:
//   amp[k], amm[k], ez[k], eref[k] (= exp(rkappa*z/h))
// ---------------------------------------------------------------------------------------------
struct SorArgs {
    int jd, kmax, maxits;
    double om, a, dphi, dz, eps, rjac, pi, h, r;
    const double *ajk, *bjk, *fact, *uz, *cosj, *cor;  // [jd]
    const double *amp, *amm, *ez, *eref;               // [kmax]
    const double *qref;                                // [k][jd] normalised
    const double *ubar, *tbar, *cref, *cbar;           // [k][jd]
    const double *tb;                                  // [kmax] hemispheric-mean theta
    double *u, *tref, *fjk;                            // [k][jd]
    int* num_of_iter;
    int u_in_smem;
};

__global__ void __launch_bounds__(1024)
sor_kernel(const SorArgs* __restrict__ probs) {
    const SorArgs p = probs[blockIdx.x];
    extern __shared__ double sh[];
    __shared__ double red[33];
    __shared__ double s_anormf;
    const int jd = p.jd, kmax = p.kmax;
    const int n = jd * kmax;
    double* u = p.u_in_smem ? sh : p.u;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int t = tid; t < n; t += nt) u[t] = 0.0;
    // f(j,k) and its 1-norm
    double part = 0.0;
    for (int t = tid; t < n; t += nt) {
        const int k = t / jd, j = t % jd;
        double f = 0.0;
        if (j >= 1 && j <= jd - 2 && k >= 1 && k <= kmax - 2) {
            f = -p.om * p.a * p.dphi * (p.qref[(size_t)k * jd + j + 1] - p.qref[(size_t)k * jd + j - 1]);
            part += fabs(f);
        }
        p.fjk[t] = f;
    }
    const double anormf = block_sum(part, red);
    if (tid == 0) s_anormf = anormf;
    __syncthreads();
    const int nj = jd - 2, nk = kmax - 2;
    const int npts = nj * nk;
    double omega = 1.0;
    int nnn = 1;
    bool converged = false;
    for (nnn = 1; nnn <= p.maxits; ++nnn) {
        double an = 0.0;
        const int par = nnn & 1;
        for (int t = tid; t < npts; t += nt) {
            const int kk = t / nj, jj = t % nj;
            const int j = jj + 2, k = kk + 2;  // 1-based
            if (((j + k) & 1) != par) continue;
            const int c = (k - 1) * jd + (j - 1);
            const double aj = p.ajk[j - 1], bj = p.bjk[j - 1], fc = p.fact[j - 1];
            const double cj = p.amp[k - 1] * fc * p.ez[k - 1];
            const double dj = p.amm[k - 1] * fc * p.ez[k - 1];
            const double ej = -aj - bj - cj - dj;
            const double resid = aj * u[c + 1] + bj * u[c - 1] + cj * u[c + jd] + dj * u[c - jd] + ej * u[c] - p.fjk[c];
            an += fabs(resid);
            if (ej != 0.) u[c] = u[c] - omega * resid / ej; else u[c] = 0.;
        }
        __syncthreads();
        // boundary refresh (order inside the reference's j loop is immaterial)
        for (int j = tid + 2; j <= jd - 1; j += nt) {
            u[j - 1] = 0.;
            u[(kmax - 1) * jd + j - 1] = u[(kmax - 3) * jd + j - 1] - p.uz[j - 1];
        }
        for (int k = tid; k < kmax; k += nt) {
            u[k * jd + jd - 1] = 0.;
            u[k * jd] = p.ubar[(size_t)k * jd] + (p.cref[(size_t)k * jd] - p.cbar[(size_t)k * jd]) / (2. * p.pi * p.a);
        }
        const double anorm = block_sum(an, red);  // contains the needed __syncthreads
        if (nnn == 1) omega = 1. / (1. - 0.5 * p.rjac * p.rjac);
        else omega = 1. / (1. - 0.25 * p.rjac * p.rjac * omega);
        if (nnn > 1 && anorm < p.eps * s_anormf) { converged = true; break; }
    }
    __syncthreads();
    if (!converged) {
        for (int t = tid; t < n; t += nt) u[t] = 0.0;
        nnn = p.maxits + 1;
    }
    if (tid == 0) *p.num_of_iter = nnn;
    __syncthreads();
    // u / cos, edge rows  (compute_reference_states.f90:224-229)
    for (int t = tid; t < n; t += nt) {
        const int j = t % jd;
        if (j >= 1 && j <= jd - 2) u[t] = u[t] / p.cosj[j];
    }
    __syncthreads();
    for (int k = tid; k < kmax; k += nt) {
        u[k * jd] = p.ubar[(size_t)k * jd];
        u[k * jd + jd - 1] = 2. * u[k * jd + jd - 2] - u[k * jd + jd - 3];
    }
    __syncthreads();
    if (p.u_in_smem)
        for (int t = tid; t < n; t += nt) p.u[t] = u[t];
    __syncthreads();
    // tref by thermal-wind marching, one thread per interior level (:233-256)
    for (int k = tid + 1; k <= kmax - 2; k += nt) {
        double* tr = p.tref + (size_t)k * jd;
        tr[0] = 0.; tr[1] = 0.;
        for (int j = 2; j <= jd - 1; ++j) {  // 1-based j
            const double uzz = (u[(k + 1) * jd + j - 1] - u[(k - 1) * jd + j - 1]) / (2. * p.dz);
            double ty = -p.cor[j - 1] * uzz * p.a * p.h * p.eref[k];
            ty = ty / p.r;
            tr[j] = tr[j - 2] + 2. * ty * p.dphi;
        }
        double tg = 0., wt = 0.;
        for (int j = 1; j <= jd; ++j) {
            tg = tg + p.cosj[j - 1] * tr[j - 1];
            wt = wt + p.cosj[j - 1];
        }
        tg = tg / wt;
        const double tres = p.tb[k] - tg;
        for (int j = 0; j < jd; ++j) tr[j] = tr[j] + tres;
    }
    __syncthreads();
    for (int j = tid; j < jd; j += nt) {
        p.tref[j] = p.tref[jd + j] - p.tb[1] + p.tb[0];
        p.tref[(size_t)(kmax - 1) * jd + j] = p.tref[(size_t)(kmax - 2) * jd + j] - p.tb[kmax - 2] + p.tb[kmax - 1];
    }
}

// top-boundary increment of the SOR solve (compute_reference_states.f90:188-190)
__global__ void sor_uz_kernel(int jd, int kmax, const double* __restrict__ c1, const double* __restrict__ c2,
                              const double* __restrict__ tbar, double* __restrict__ uz) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;  // 0-based
    if (j >= jd) return;
    double v = 0.0;
    if (j >= 1 && j <= jd - 2) {
        v = c1[j];
        v = v * (tbar[(size_t)(kmax - 2) * jd + j + 1] - tbar[(size_t)(kmax - 2) * jd + j - 1]) / c2[j];
    }
    uz[j] = v;
}

// hemispheric-mean theta of the SOR routine: tb(k) = sum_j cos(phi_j) tbar(j,k) / sum_j cos(phi_j)
__global__ void tb_mean_kernel(int kmax, int jd, const double* __restrict__ tbar, const double* __restrict__ cosw,
                               double* __restrict__ tb) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kmax) return;
    double s = 0., w = 0.;
    for (int j = 0; j < jd; ++j) {
        s = s + cosw[j] * tbar[(size_t)k * jd + j];
        w = w + cosw[j];
    }
    tb[k] = s / w;
}

// ---------------------------------------------------------------------------------------------
// NHN22 pieces
// ---------------------------------------------------------------------------------------------
// top boundary condition (compute_qref_and_fawa_first.f90:164-175): tjk(:,kmax-1), sjk(:,:,kmax-1)=I
__global__ void nhn22_top_bc_kernel(int n, int nd, int kmax, int jb, const double* __restrict__ tbar,
                                    const double* __restrict__ coef /* [n]: -dz*rr*cos0*exp(..) */,
                                    const double* __restrict__ den /* [n]: 4*omega*sin0*dphi*h*a */,
                                    double* __restrict__ tjk, double* __restrict__ sjk) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;  // 0-based row of the (jd-2) system
    if (m >= n) return;
    const int j = m + 2;  // 1-based shifted index j = jj - jb
    double t = coef[m];
    t = t * (tbar[(size_t)(kmax - 1) * nd + j] - tbar[(size_t)(kmax - 1) * nd + j - 2]);
    t = t / den[m];
    tjk[(size_t)(kmax - 2) * n + m] = t;
    sjk[(size_t)(kmax - 2) * n * n + (size_t)m * n + m] = 1.0;
}

// matrix_b4_inversion.f90:43-86 — element (i,j) of Q_k + C_k S_k, plus C_k, D_k, r_k
__global__ void __launch_bounds__(256)
b4_kernel(int n, int k /*1-based*/, int nd, int jb, const double* __restrict__ ajk, const double* __restrict__ bjk,
          const double* __restrict__ cjk, const double* __restrict__ djk, double fcoef,
          double u1coef, double u1off, const double* __restrict__ qref, const double* __restrict__ ckref,
          const double* __restrict__ sjj, double* __restrict__ qjj, double* __restrict__ djj,
          double* __restrict__ cjj, double* __restrict__ rj) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * n) return;
    const int i = (int)(t % n), j = (int)(t / n);  // Fortran (i,j), 0-based; i fastest
    const double a = ajk[i], b = bjk[i], c = cjk[i], d = djk[i];
    double q = 0.0;
    if (i == j) q = -(a + b + c + d);
    else if (j == i + 1) q = a;
    else if (j == i - 1) q = b;
    const double x = 0.0 + c * sjj[t];
    qjj[t] = q + x;
    cjj[t] = (i == j) ? c : 0.0;
    djj[t] = (i == j) ? d : 0.0;
    if (j == 0) {
        const int jj = i + 2 + jb;  // 1-based global hemispheric row
        const double fjk = fcoef * (qref[(size_t)(k - 1) * nd + jj] - qref[(size_t)(k - 1) * nd + jj - 2]);
        double r = fjk;
        if (i == 0) {
            const double u1 = ckref[(size_t)(k - 1) * nd + jb] / u1coef - u1off;
            r = fjk - b * u1;
        }
        if (i == n - 1) r = fjk - a * 0.0;
        rj[i] = r;
    }
}

// matrix_after_inversion.f90:11-19: sjk(:,:,k-1) = -(Qinv * D), D diagonal
__global__ void __launch_bounds__(256)
after_s_kernel(int n, const double* __restrict__ qinv, const double* __restrict__ djj, double* __restrict__ sout) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * n) return;
    const int j = (int)(t / n);
    const double x = 0.0 + qinv[t] * djj[(size_t)j * n + j];
    sout[t] = -x;
}
// matrix_after_inversion.f90:22-37: tjk(:,k-1) = Qinv (r - C t_k); one thread per row, sum in kk order
__global__ void after_t_kernel(int n, const double* __restrict__ qinv, const double* __restrict__ cjj,
                               const double* __restrict__ rj, const double* __restrict__ tin,
                               double* __restrict__ tout) {
    extern __shared__ double y[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double c = 0.0 + cjj[(size_t)i * n + i] * tin[i];
        y[i] = rj[i] - c;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double t = 0.0;
        for (int kk = 0; kk < n; ++kk) t = t + qinv[(size_t)kk * n + i] * y[kk];
        tout[i] = t;  // level k-1; the input tin is level k, so there is no aliasing
    }
}

// upward_sweep.f90:19-91.  Single CTA; thread-per-row mat-vec keeps the reference's summation order.
struct SweepArgs {
    int n, jd, nd, kmax, jb;
    double a, om, dz, h, rr, dp, pi;
    const double *sjk, *tjk, *ckref, *tb, *qoc;
    const double *cosrow;  // [jd] cos(dp*(jj-1)), jj=jb+1..nd
    const double *sinrow;  // [jd] sin(dp*(j-1+jb)) (== sin of the same latitude, other expression)
    const double *sinnd;   // [nd] sin(dp*(j-1))
    const double *eref;    // [kmax] exp(rkappa*zz/h)
    double *tref, *qref, *u, *pjk;  // pjk scratch [kmax][n]
    double u1top;                  // -om*a*cos(dp*jb) part handled on host: value = ckref(1+jb,kmax)/(2 pi a) - off
    double u1off;
};

// p_{k+1} = S_k p_k + T_k  (upward_sweep.f90:19-34); thread-per-row keeps the reference's summation order
__global__ void __launch_bounds__(1024)
sweep_matvec_kernel(int n, int kmax, const double* __restrict__ sjk, const double* __restrict__ tjk,
                    double* __restrict__ pjk) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) pjk[i] = 0.0;
    __syncthreads();
    for (int k = 0; k < kmax - 1; ++k) {
        const double* S = sjk + (size_t)k * n * n;
        const double* pk = pjk + (size_t)k * n;
        for (int i = tid; i < n; i += nt) {
            double y = 0.0;
            for (int kk = 0; kk < n; ++kk) y = y + S[(size_t)kk * n + i] * pk[kk];
            pjk[(size_t)(k + 1) * n + i] = y + tjk[(size_t)k * n + i];
        }
        __syncthreads();
    }
}

// upward_sweep.f90:36-91 given p_k: u, corner values, /cos, tref marching, qref * sin(lat).
// One CTA per (snapshot, hemisphere) problem.
__global__ void __launch_bounds__(512)
sweep_post_kernel(const SweepArgs* __restrict__ probs) {
    const SweepArgs p = probs[blockIdx.x];
    const int n = p.n, jd = p.jd, nd = p.nd, kmax = p.kmax;
    const int tid = threadIdx.x, nt = blockDim.x;
    // recover u, corners, /cos, extrapolated last row (:36-57)
    for (int t = tid; t < jd * kmax; t += nt) {
        const int k = t / jd, j = t % jd;  // 0-based
        double v = 0.0;
        if (j >= 1 && j <= jd - 2) v = p.pjk[(size_t)k * n + j - 1];
        if (j == 0 && k == kmax - 1) v = p.ckref[(size_t)(kmax - 1) * nd + p.jb] / (2. * p.pi * p.a) - p.u1off;
        if (j <= jd - 2) v = v / p.cosrow[j];
        p.u[t] = v;
    }
    __syncthreads();
    for (int k = tid; k < kmax; k += nt)
        p.u[(size_t)k * jd + jd - 1] = 2. * p.u[(size_t)k * jd + jd - 2] - p.u[(size_t)k * jd + jd - 3];
    // qref = qref_over_cor * sin(lat) on interior levels, copy elsewhere (:73-76)
    for (int t = tid; t < nd * kmax; t += nt) {
        const int k = t / nd, j = t % nd;
        const bool interior = (k >= 1 && k <= kmax - 2);
        p.qref[t] = interior ? p.qoc[t] * p.sinnd[j] : p.qoc[t];
    }
    __syncthreads();
    for (int k = tid + 1; k <= kmax - 2; k += nt) {
        double* __restrict__ tr = p.tref + (size_t)k * jd;
        const double* __restrict__ up = p.u + (size_t)(k + 1) * jd;
        const double* __restrict__ um = p.u + (size_t)(k - 1) * jd;
        tr[0] = 0.; tr[1] = 0.;
        // thermal-wind marching tr[j] = tr[j-2] + ...: only kmax-2 threads and a sequential chain, so the operands
        // of eight steps are fetched first and the chain runs on registers
        double t0v = 0., t1v = 0.;   // tr[j-2], tr[j-1]
        double tg = 0., wt = 0.;
        tg = tg + p.cosrow[0] * 0.; wt = wt + p.cosrow[0];
        tg = tg + p.cosrow[1] * 0.; wt = wt + p.cosrow[1];
        constexpr int U = 8;
        for (int j0 = 2; j0 <= jd - 1; j0 += U) {
            double a[U], b[U], sn[U], cs[U], out[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int j = min(j0 + q, jd - 1);
                a[q] = up[j - 1]; b[q] = um[j - 1]; sn[q] = p.sinrow[j - 1]; cs[q] = p.cosrow[j];
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int j = j0 + q;
                const double cor = 2. * p.om * sn[q];
                const double uz = (a[q] - b[q]) / (2. * p.dz);
                double ty = -cor * uz * p.a * p.h * p.eref[k];
                ty = ty / p.rr;
                const double v = t0v + 2. * ty * p.dp;
                out[q] = v;
                if (j <= jd - 1) { t0v = t1v; t1v = v; tg = tg + cs[q] * v; wt = wt + cs[q]; }
            }
#pragma unroll
            for (int q = 0; q < U; ++q)
                if (j0 + q <= jd - 1) tr[j0 + q] = out[q];
        }
        tg = tg / wt;
        const double tres = p.tb[k] - tg;
        for (int j = 0; j < jd; ++j) tr[j] = tr[j] + tres;
    }
    __syncthreads();
    for (int j = tid; j < jd; j += nt) {
        p.tref[j] = p.tref[jd + j] - p.tb[1] + p.tb[0];
        p.tref[(size_t)(kmax - 1) * jd + j] = p.tref[(size_t)(kmax - 2) * jd + j] - p.tb[kmax - 2] + p.tb[kmax - 1];
    }
}

// ---------------------------------------------------------------------------------------------
// host-side helpers shared by the C entry points
// ---------------------------------------------------------------------------------------------
static void host_alat_da(int nrow, int nlat, double a, double dphi, double dlambda, std::vector<double>& alat,
                         std::vector<double>& da) {
    const double pi = host_pi();
    alat.resize(nrow);
    for (int nn = 1; nn <= nrow; ++nn) alat[nn - 1] = 2. * pi * a * a * (1. - std::sin(dphi * (double)(nn - 1)));
    da.resize(nlat);
    for (int j = 1; j <= nlat; ++j) {
        const double phi0 = -0.5 * pi + dphi * (double)(j - 1);
        da[j - 1] = a * a * dphi * dlambda * std::cos(phi0);
    }
}

// zonal means (rows nrow..nlat of the given frame), extrema, histograms and inversion for one field set
struct AreaAnalysis {
    DeviceScratch<double> zm;                  // [3][kmax][nlat]
    DeviceScratch<unsigned long long> ext;     // [2 fields][kmax][2]
    DeviceScratch<double> hist;                // [kmax][2][nb]
};

static int launch_hist(int nlon, int nlat, int kmax, int nb, const double* fld, const unsigned long long* ext,
                       double sgn, const double* d_da, double* hist, int32_t* bins, cudaStream_t st) {
    FALWA_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)kmax * 2 * nb, st));
    // ~4 CTAs per SM in total
    int slabs = max(1, min(nlat, (148 * 4 + (kmax - 3)) / (kmax - 2)));
    int rows = ceil_div(nlat, slabs);
    slabs = ceil_div(nlat, rows);
    const size_t smem = sizeof(double) * 2 * nb;
    { FALWA_KERNEL("area_hist_kernel", st); area_hist_kernel<<<dim3(slabs, kmax - 2), kHistThreads, smem, st>>>(nlon, nlat, kmax, nb, fld, ext, sgn, d_da,
                                                                       rows, hist, bins); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

}  // namespace falwa

using namespace falwa;

// ================================ C ABI ========================================================
extern "C" int falwa_b200_compute_reference_states(const double* pv, const double* uu, const double* pt,
                                                   const double* stat_host, int nlon, int nlat, int kmax, int jd,
                                                   int npart, int maxits, double a, double om, double dz, double eps,
                                                   double h, double dphi, double dlambda, double r, double cp,
                                                   double rjac, double* qref, double* u, double* tref,
                                                   int* num_of_iter_host, int32_t* bins, void* stream) {
    if (!pv || !uu || !pt || !stat_host || !qref || !u || !tref || !num_of_iter_host) return FALWA_ERR_NULL;
    if (nlon < 3 || kmax < 4 || jd < 4 || nlat != 2 * jd - 1 || npart < 2) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const double pi = host_pi();
    const double rkappa = r / cp;
    const size_t njk = (size_t)jd * kmax;
    // ---- host tables -------------------------------------------------------------------------
    std::vector<double> alat, da;
    host_alat_da(jd, nlat, a, dphi, dlambda, alat, da);
    std::vector<double> tab;
    auto push = [&](const std::vector<double>& v) { size_t o = tab.size(); tab.insert(tab.end(), v.begin(), v.end()); return o; };
    std::vector<double> cosmid(jd, 0.), norm(jd, 1.), ajk(jd, 0.), bjk(jd, 0.), fact(jd, 0.), cosj(jd, 0.),
        cor(jd, 0.), cosw(jd, 0.), amp(kmax, 0.), amm(kmax, 0.), ez(kmax, 0.), eref(kmax, 0.), z(kmax);
    for (int k = 1; k <= kmax; ++k) z[k - 1] = dz * (double)(k - 1);
    for (int j = 1; j <= jd; ++j) {
        cosmid[j - 1] = std::cos(dphi * ((double)j - 0.5));
        const double phi0 = dphi * (double)(j - 1);
        norm[j - 1] = 2. * om * std::sin(phi0);
        cosj[j - 1] = std::cos(phi0);
        cor[j - 1] = 2. * om * std::sin(phi0);
        cosw[j - 1] = std::cos(dphi * (double)(j + jd - 2) - 0.5 * pi);  // rows j=jd..nlat (:60-65)
        if (j >= 2 && j <= jd - 1) {
            const double phip = ((double)j - 0.5) * dphi, phim = ((double)j - 1.5) * dphi;
            ajk[j - 1] = 1. / (std::sin(phip) * std::cos(phip));
            bjk[j - 1] = 1. / (std::sin(phim) * std::cos(phim));
            fact[j - 1] = 4. * om * om * h * a * a * std::sin(phi0) * dphi * dphi / (dz * dz * r * std::cos(phi0));
        }
    }
    for (int k = 2; k <= kmax - 1; ++k) {
        const double zp = 0.5 * (z[k] + z[k - 1]), zm = 0.5 * (z[k - 2] + z[k - 1]);
        const double statp = 0.5 * (stat_host[k] + stat_host[k - 1]), statm = 0.5 * (stat_host[k - 2] + stat_host[k - 1]);
        amp[k - 1] = std::exp(-zp / h) * std::exp(rkappa * zp / h) / statp;
        amm[k - 1] = std::exp(-zm / h) * std::exp(rkappa * zm / h) / statm;
        ez[k - 1] = std::exp(z[k - 1] / h);
    }
    for (int k = 1; k <= kmax; ++k) eref[k - 1] = std::exp(rkappa * (dz * (double)(k - 1)) / h);
    const size_t o_alat = push(alat), o_da = push(da), o_cosmid = push(cosmid), o_norm = push(norm),
                 o_ajk = push(ajk), o_bjk = push(bjk), o_fact = push(fact), o_cosj = push(cosj), o_cor = push(cor),
                 o_cosw = push(cosw), o_amp = push(amp), o_amm = push(amm), o_ez = push(ez), o_eref = push(eref);
    DeviceTable T;
    int rc = T.upload(tab, st);
    if (rc) return rc;
    const double* d = T.d;
    // ---- scratch -----------------------------------------------------------------------------
    DeviceScratch<double> zm, hist, prof, tb, uzdev;
    DeviceScratch<unsigned long long> ext;
    DeviceScratch<int> nit;
    if ((rc = zm.alloc((size_t)3 * kmax * nlat, st))) return rc;
    if ((rc = hist.alloc((size_t)kmax * 2 * npart, st))) return rc;
    if ((rc = prof.alloc(njk * 6, st, true))) return rc;  // qbar? ubar tbar cref cbar fjk
    if ((rc = tb.alloc(kmax, st))) return rc;
    if ((rc = uzdev.alloc(jd, st))) return rc;
    if ((rc = ext.alloc((size_t)2 * kmax * 2, st))) return rc;
    if ((rc = nit.alloc(1, st))) return rc;
    double *ubar = prof.d, *tbar = prof.d + njk, *cref = prof.d + 2 * njk, *cbar = prof.d + 3 * njk,
           *fjk = prof.d + 4 * njk;
    FALWA_CUDA_TRY(cudaMemsetAsync(qref, 0, sizeof(double) * njk, st));
    FALWA_CUDA_TRY(cudaMemsetAsync(tref, 0, sizeof(double) * njk, st));
    { FALWA_KERNEL("init_ext_kernel", st); init_ext_kernel<<<ceil_div(2 * kmax * 2, 128), 128, 0, st>>>(ext.d, 2 * kmax * 2); }
    { FALWA_KERNEL("zonal_stats_kernel", st); zonal_stats_kernel<<<dim3(ceil_div((long long)nlat * kmax, 8)), 256, 0, st>>>(nlon, nlat, kmax, pv, uu, pt, nullptr, zm.d, ext.d, 1); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("extract_rows_kernel", st); extract_rows_kernel<<<ceil_div(njk, 256), 256, 0, st>>>(kmax, jd, nlat, jd - 1, 0, 1.0, zm.d + (size_t)1 * kmax * nlat, ubar); }
    { FALWA_KERNEL("extract_rows_kernel", st); extract_rows_kernel<<<ceil_div(njk, 256), 256, 0, st>>>(kmax, jd, nlat, jd - 1, 0, 1.0, zm.d + (size_t)2 * kmax * nlat, tbar); }
    { FALWA_KERNEL("tb_mean_kernel", st); tb_mean_kernel<<<ceil_div(kmax, 64), 64, 0, st>>>(kmax, jd, tbar, d + o_cosw, tb.d); }
    if ((rc = launch_hist(nlon, nlat, kmax, npart, pv, ext.d, 1.0, d + o_da, hist.d, bins, st))) return rc;
    { FALWA_KERNEL("qref_invert_kernel", st); qref_invert_kernel<<<kmax - 2, 256, sizeof(double) * 2 * npart, st>>>(kmax, npart, jd, jd, hist.d, ext.d, 1.0,
                                                                           d + o_alat, qref, cref, 1); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("cbar_normalise_kernel", st); cbar_normalise_kernel<<<ceil_div(kmax, 64), 64, 0, st>>>(kmax, jd, zm.d, nlat, 1.0, 0, jd - 1, d + o_cosmid,
                                                             d + o_norm, a, dphi, pi, qref, cref, cbar, nullptr); }
    FALWA_LAUNCH_CHECK();
    // uz(j) = c1(j) * (tbar(j+1,kmax-1) - tbar(j-1,kmax-1)) / c2(j)   (:188-190)
    DeviceTable C;
    {
        std::vector<double> both(2 * (size_t)jd, 0.);
        for (int j = 1; j <= jd; ++j) both[jd + j - 1] = 1.;
        for (int j = 2; j <= jd - 1; ++j) {
            const double phi0 = dphi * (double)(j - 1);
            both[j - 1] = dz * r * std::cos(phi0) * std::exp(-z[kmax - 2] * rkappa / h);
            both[jd + j - 1] = 2. * om * std::sin(phi0) * dphi * h * a;
        }
        if ((rc = C.upload(both, st))) return rc;
        { FALWA_KERNEL("sor_uz_kernel", st); sor_uz_kernel<<<ceil_div(jd, 128), 128, 0, st>>>(jd, kmax, C.d, C.d + jd, tbar, uzdev.d); }
        FALWA_LAUNCH_CHECK();
    }
    SorArgs p;
    p.jd = jd; p.kmax = kmax; p.maxits = maxits;
    p.om = om; p.a = a; p.dphi = dphi; p.dz = dz; p.eps = eps; p.rjac = rjac; p.pi = pi; p.h = h; p.r = r;
    p.ajk = d + o_ajk; p.bjk = d + o_bjk; p.fact = d + o_fact; p.uz = uzdev.d; p.cosj = d + o_cosj; p.cor = d + o_cor;
    p.amp = d + o_amp; p.amm = d + o_amm; p.ez = d + o_ez; p.eref = d + o_eref;
    p.qref = qref; p.ubar = ubar; p.tbar = tbar; p.cref = cref; p.cbar = cbar; p.tb = tb.d;
    p.u = u; p.tref = tref; p.fjk = fjk; p.num_of_iter = nit.d;
    const size_t need = sizeof(double) * njk;
    p.u_in_smem = need <= 200 * 1024;
    const size_t smem = p.u_in_smem ? need : 0;
    FALWA_CUDA_TRY(cudaFuncSetAttribute(sor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    DeviceScratch<SorArgs> dp;
    if ((rc = dp.alloc(1, st))) return rc;
    { int rcu = upload_bytes(dp.d, &p, sizeof(SorArgs), st); if (rcu) return rcu; }
    { FALWA_KERNEL("sor_kernel", st); sor_kernel<<<1, 1024, smem, st>>>(dp.d); }
    FALWA_LAUNCH_CHECK();
    FALWA_CUDA_TRY(cudaMemcpyAsync(num_of_iter_host, nit.d, sizeof(int), cudaMemcpyDeviceToHost, st));
    FALWA_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int falwa_b200_compute_qref_and_fawa_first(const double* pv, const double* uu, const double* vort,
                                                      const double* pt, const double* tn0_host, int imax, int jmax,
                                                      int kmax, int nd, int nnd, int jb, int jd, double a,
                                                      double omega, double dz, double h, double dphi, double dlambda,
                                                      double rr, double cp, double* qref, double* ubar, double* tbar,
                                                      double* fawa, double* ckref, double* tjk, double* sjk,
                                                      int32_t* bins_pv, int32_t* bins_vort, void* stream) {
    (void)tn0_host;
    if (!pv || !uu || !vort || !pt || !qref || !ubar || !tbar || !fawa || !ckref || !tjk || !sjk) return FALWA_ERR_NULL;
    if (imax < 3 || kmax < 4 || nd < 4 || jmax != 2 * nd - 1 || nnd < 2 || jb < 0 || jd != nd - jb || jd < 4)
        return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const double pi = host_pi();
    const double rkappa = rr / cp;
    const int n = jd - 2;
    const size_t njk = (size_t)nd * kmax;
    std::vector<double> alat, da, tab;
    host_alat_da(nd, jmax, a, dphi, dlambda, alat, da);
    auto push = [&](const std::vector<double>& v) { size_t o = tab.size(); tab.insert(tab.end(), v.begin(), v.end()); return o; };
    std::vector<double> cosmid(nd, 0.), norm(nd, 1.), coef(n, 0.), den(n, 1.);
    for (int j = 1; j <= nd; ++j) {
        cosmid[j - 1] = std::cos(dphi * ((double)j - 0.5));
        norm[j - 1] = std::sin(dphi * (double)(j - 1));
    }
    const double ztop = dz * (double)(kmax - 2);
    for (int jj = jb + 2; jj <= nd - 1; ++jj) {
        const double phi0 = (double)(jj - 1) * dphi;
        coef[jj - jb - 2] = -dz * rr * std::cos(phi0) * std::exp(-ztop * rkappa / h);
        den[jj - jb - 2] = 4. * omega * std::sin(phi0) * dphi * h * a;
    }
    const size_t o_alat = push(alat), o_da = push(da), o_cosmid = push(cosmid), o_norm = push(norm),
                 o_coef = push(coef), o_den = push(den);
    DeviceTable T;
    int rc = T.upload(tab, st);
    if (rc) return rc;
    const double* d = T.d;
    DeviceScratch<double> zm, hist, prof;
    DeviceScratch<unsigned long long> ext;
    if ((rc = zm.alloc((size_t)3 * kmax * jmax, st))) return rc;
    if ((rc = hist.alloc((size_t)kmax * 2 * nnd, st))) return rc;
    if ((rc = prof.alloc(njk * 2, st, true))) return rc;  // cref, cbar
    if ((rc = ext.alloc((size_t)2 * kmax * 2, st))) return rc;
    double *cref = prof.d, *cbar = prof.d + njk;
    FALWA_CUDA_TRY(cudaMemsetAsync(qref, 0, sizeof(double) * njk, st));
    FALWA_CUDA_TRY(cudaMemsetAsync(ckref, 0, sizeof(double) * njk, st));
    FALWA_CUDA_TRY(cudaMemsetAsync(tjk, 0, sizeof(double) * (size_t)n * (kmax - 1), st));
    FALWA_CUDA_TRY(cudaMemsetAsync(sjk, 0, sizeof(double) * (size_t)n * n * (kmax - 1), st));
    { FALWA_KERNEL("init_ext_kernel", st); init_ext_kernel<<<ceil_div(2 * kmax * 2, 128), 128, 0, st>>>(ext.d, 2 * kmax * 2); }
    { FALWA_KERNEL("zonal_stats_kernel", st); zonal_stats_kernel<<<dim3(ceil_div((long long)jmax * kmax, 8)), 256, 0, st>>>(imax, jmax, kmax, pv, uu, pt, vort, zm.d, ext.d, 3); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("extract_rows_kernel", st); extract_rows_kernel<<<ceil_div(njk, 256), 256, 0, st>>>(kmax, nd, jmax, nd - 1, 0, 1.0, zm.d + (size_t)1 * kmax * jmax, ubar); }
    { FALWA_KERNEL("extract_rows_kernel", st); extract_rows_kernel<<<ceil_div(njk, 256), 256, 0, st>>>(kmax, nd, jmax, nd - 1, 0, 1.0, zm.d + (size_t)2 * kmax * jmax, tbar); }
    // QGPV area analysis -> qref, cref
    if ((rc = launch_hist(imax, jmax, kmax, nnd, pv, ext.d, 1.0, d + o_da, hist.d, bins_pv, st))) return rc;
    { FALWA_KERNEL("qref_invert_kernel", st); qref_invert_kernel<<<kmax - 2, 256, sizeof(double) * 2 * nnd, st>>>(kmax, nnd, nd, nd, hist.d, ext.d, 1.0,
                                                                         d + o_alat, qref, cref, 1); }
    FALWA_LAUNCH_CHECK();
    // absolute-vorticity area analysis -> ckref
    if ((rc = launch_hist(imax, jmax, kmax, nnd, vort, ext.d + (size_t)kmax * 2, 1.0, d + o_da, hist.d, bins_vort, st)))
        return rc;
    { FALWA_KERNEL("qref_invert_kernel", st); qref_invert_kernel<<<kmax - 2, 256, sizeof(double) * 2 * nnd, st>>>(kmax, nnd, nd, nd, hist.d,
                                                                         ext.d + (size_t)kmax * 2, 1.0, d + o_alat,
                                                                         nullptr, ckref, 0); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("cbar_normalise_kernel", st); cbar_normalise_kernel<<<ceil_div(kmax, 64), 64, 0, st>>>(kmax, nd, zm.d, jmax, 1.0, 0, nd - 1, d + o_cosmid,
                                                             d + o_norm, a, dphi, pi, qref, cref, cbar, fawa); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("nhn22_top_bc_kernel", st); nhn22_top_bc_kernel<<<ceil_div(n, 128), 128, 0, st>>>(n, nd, kmax, jb, tbar, d + o_coef, d + o_den, tjk, sjk); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

extern "C" int falwa_b200_matrix_b4_inversion(int k, int jmax, int kmax, int nd, int jb, int jd,
                                              const double* z_host, const double* statn_host, const double* qref,
                                              const double* ckref, const double* sjk, double a, double om, double dz,
                                              double h, double rr, double cp, double* qjj, double* djj, double* cjj,
                                              double* rj, void* stream) {
    if (!z_host || !statn_host || !qref || !ckref || !sjk || !qjj || !djj || !cjj || !rj) return FALWA_ERR_NULL;
    if (k < 2 || k > kmax - 1 || jd != nd - jb || jd < 4) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = jd - 2;
    const double rkappa = rr / cp, pi = host_pi(), dp = pi / (double)(jmax - 1);
    const double* z = z_host;
    const double zp = 0.5 * (z[k] + z[k - 1]), zm = 0.5 * (z[k - 2] + z[k - 1]);
    const double statp = 0.5 * (statn_host[k] + statn_host[k - 1]), statm = 0.5 * (statn_host[k - 2] + statn_host[k - 1]);
    std::vector<double> tab(4 * (size_t)n);
    for (int jj = jb + 2; jj <= nd - 1; ++jj) {
        const int i = jj - jb - 2;
        const double phi0 = (double)(jj - 1) * dp, phip = ((double)jj - 0.5) * dp, phim = ((double)jj - 1.5) * dp;
        const double fact = 4. * om * om * h * a * a * std::sin(phi0) * dp * dp / (dz * dz * rr * std::cos(phi0));
        double amp = std::exp(-zp / h) * std::exp(rkappa * zp / h) / statp;
        amp = amp * fact * std::exp(z[k - 1] / h);
        double amm = std::exp(-zm / h) * std::exp(rkappa * zm / h) / statm;
        amm = amm * fact * std::exp(z[k - 1] / h);
        tab[i] = 1. / (std::sin(phip) * std::cos(phip));
        tab[n + i] = 1. / (std::sin(phim) * std::cos(phim));
        tab[2 * n + i] = amp;
        tab[3 * n + i] = amm;
    }
    DeviceTable T;
    int rc = T.upload(tab, st);
    if (rc) return rc;
    const double fcoef = -0.5 * a * dp;
    const double u1coef = 2. * pi * a, u1off = om * a * std::cos(dp * (double)jb);
    { FALWA_KERNEL("b4_kernel", st); b4_kernel<<<ceil_div((long long)n * n, 256), 256, 0, st>>>(n, k, nd, jb, T.d, T.d + n, T.d + 2 * n, T.d + 3 * n, fcoef,
                                                              u1coef, u1off, qref, ckref, sjk + (size_t)(k - 1) * n * n,
                                                              qjj, djj, cjj, rj); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

extern "C" int falwa_b200_matrix_after_inversion(int k, int kmax, int jd, const double* qjj, const double* djj,
                                                 const double* cjj, const double* rj, double* sjk, double* tjk,
                                                 void* stream) {
    if (!qjj || !djj || !cjj || !rj || !sjk || !tjk) return FALWA_ERR_NULL;
    if (k < 2 || k > kmax - 1 || jd < 4) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = jd - 2;
    { FALWA_KERNEL("after_s_kernel", st); after_s_kernel<<<ceil_div((long long)n * n, 256), 256, 0, st>>>(n, qjj, djj, sjk + (size_t)(k - 2) * n * n); }
    FALWA_LAUNCH_CHECK();
    { FALWA_KERNEL("after_t_kernel", st); after_t_kernel<<<1, 512, sizeof(double) * n, st>>>(n, qjj, cjj, rj, tjk + (size_t)(k - 1) * n,
                                                        tjk + (size_t)(k - 2) * n); }
    FALWA_LAUNCH_CHECK();
    return 0;
}

extern "C" int falwa_b200_upward_sweep(int jmax, int kmax, int nd, int jb, int jd, const double* sjk,
                                       const double* tjk, const double* ckref, const double* tb_host,
                                       const double* qref_over_cor, double* tref, double* qref, double* u, double a,
                                       double om, double dz, double h, double rr, double cp, void* stream) {
    if (!sjk || !tjk || !ckref || !tb_host || !qref_over_cor || !tref || !qref || !u) return FALWA_ERR_NULL;
    if (jd != nd - jb || jd < 4 || kmax < 4) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = jd - 2;
    const double rkappa = rr / cp, pi = host_pi(), dp = pi / (double)(jmax - 1);
    std::vector<double> tab;
    auto push = [&](const std::vector<double>& v) { size_t o = tab.size(); tab.insert(tab.end(), v.begin(), v.end()); return o; };
    std::vector<double> cosrow(jd), sinrow(jd), sinnd(nd), eref(kmax), tb(tb_host, tb_host + kmax);
    for (int j = 1; j <= jd; ++j) {
        cosrow[j - 1] = std::cos(dp * (double)(j + jb - 1));
        sinrow[j - 1] = std::sin(dp * (double)(j - 1 + jb));
    }
    for (int j = 1; j <= nd; ++j) sinnd[j - 1] = std::sin(dp * (double)(j - 1));
    for (int k = 1; k <= kmax; ++k) eref[k - 1] = std::exp(rkappa * (dz * (double)(k - 1)) / h);
    const size_t o_cos = push(cosrow), o_sin = push(sinrow), o_snd = push(sinnd), o_er = push(eref), o_tb = push(tb);
    DeviceTable T;
    int rc = T.upload(tab, st);
    if (rc) return rc;
    DeviceScratch<double> pjk;
    if ((rc = pjk.alloc((size_t)n * kmax, st, true))) return rc;
    SweepArgs p;
    p.n = n; p.jd = jd; p.nd = nd; p.kmax = kmax; p.jb = jb;
    p.a = a; p.om = om; p.dz = dz; p.h = h; p.rr = rr; p.dp = dp; p.pi = pi;
    p.sjk = sjk; p.tjk = tjk; p.ckref = ckref; p.tb = T.d + o_tb; p.qoc = qref_over_cor;
    p.cosrow = T.d + o_cos; p.sinrow = T.d + o_sin; p.sinnd = T.d + o_snd; p.eref = T.d + o_er;
    p.tref = tref; p.qref = qref; p.u = u; p.pjk = pjk.d;
    p.u1top = 0.0; p.u1off = om * a * std::cos(dp * (double)jb);
    FALWA_CUDA_TRY(cudaMemsetAsync(tref, 0, sizeof(double) * (size_t)jd * kmax, st));
    { FALWA_KERNEL("sweep_matvec_kernel", st); sweep_matvec_kernel<<<1, 1024, 0, st>>>(n, kmax, sjk, tjk, pjk.d); }
    FALWA_LAUNCH_CHECK();
    DeviceScratch<SweepArgs> dprob;
    if ((rc = dprob.alloc(1, st))) return rc;
    { int rcu = upload_bytes(dprob.d, &p, sizeof(SweepArgs), st); if (rcu) return rcu; }
    { FALWA_KERNEL("sweep_post_kernel", st); sweep_post_kernel<<<1, 512, 0, st>>>(dprob.d); }
    FALWA_LAUNCH_CHECK();
    return 0;
}
