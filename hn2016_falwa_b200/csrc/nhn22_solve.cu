// NHN22 reference-state inversion, B200-native formulation.
//
// The reference solves   b_i u(i-1,k) + a_i u(i+1,k) - e_ik u(i,k) + c_ik u(i,k+1) + d_ik u(i,k-1) = r_ik
// (i = latitude row of the (jd-2) system, k = 2..kmax-1) by a block recursion in k with dense
// (jd-2)^2 blocks, a host LAPACK inverse per level and a Python loop (oopinterface.py:1839-1885,
// matrix_b4_inversion.f90, matrix_after_inversion.f90, upward_sweep.f90:19-34): 33 GFLOP per 0.25°
// snapshot.  Because c_ik = fact_i * (amp_k ez_k) and d_ik = fact_i * (amm_k ez_k), the operator is
//     (1/fact_i) (J u)_i  +  K u_i      with  J tridiagonal in latitude, K tridiagonal in height,
// i.e. separable.  K is similar to a symmetric tridiagonal matrix; its m = kmax-2 eigenpairs are
// computed on the host (implicit-shift QL, ~50 us) while the stage's tables are built, and the device
// does   transform (m x m by m x n)  ->  m independent scalar tridiagonal solves in latitude  ->
// back-transform: 2*m*m*n + O(m*n) flops (1.6 MFLOP at 0.25°) instead of 33 GFLOP, same linear
// system, same boundary conditions (u(i,1)=0, u(i,kmax)=u(i,kmax-1)+t_i, u(0,k)=u1_k, u(n+1,k)=0).
// Agreement with the literal recursion (CPU oracle) is at the 1e-12 level (tests/test_gpu_pipeline.py).
#include "common.cuh"

namespace falwa {

// Eigen-decomposition of a symmetric tridiagonal matrix (diagonal d[0..m), off-diagonal e[0..m-1),
// e[i] couples i and i+1) by the implicit-shift QL iteration.  On return d holds the eigenvalues and
// z (row-major m x m) the orthonormal eigenvectors in its columns.  Returns 0, or -1 if an eigenvalue
// needed more than 60 iterations.
static int symtridiag_eig(int m, std::vector<double>& d, std::vector<double> e, std::vector<double>& z) {
    z.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) z[(size_t)i * m + i] = 1.0;
    e.resize(m, 0.0);
    e[m - 1] = 0.0;
    for (int l = 0; l < m; ++l) {
        int iter = 0;
        while (true) {
            int mm = l;
            for (; mm < m - 1; ++mm) {
                const double dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
                if (std::fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm == l) break;
            if (++iter > 60) return -1;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = std::hypot(g, 1.0);
            g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = mm - 1;
            for (; i >= l; --i) {
                double f = s * e[i];
                const double b = c * e[i];
                r = std::hypot(f, g);
                e[i + 1] = r;
                if (r == 0.0) {
                    d[i + 1] -= p;
                    e[mm] = 0.0;
                    break;
                }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
                for (int k = 0; k < m; ++k) {
                    f = z[(size_t)k * m + i + 1];
                    z[(size_t)k * m + i + 1] = s * z[(size_t)k * m + i] + c * f;
                    z[(size_t)k * m + i] = c * z[(size_t)k * m + i] - s * f;
                }
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p;
            e[l] = g;
            e[mm] = 0.0;
        }
    }
    return 0;
}

// Spectral tables of the vertical operator for one (snapshot, hemisphere):
//   K0 (m x m): row r <-> level k = r+2 (1-based);  sub(r,r-1) = dk_r, super(r,r+1) = ck_r,
//   diag = -(ck_r + dk_r), except the last row whose diag is -dk (top BC u(kmax) = u(kmax-1) + t).
//   ck_r = amp_k * ez_k, dk_r = amm_k * ez_k   (matrix_b4_inversion.f90:38-41 without fact).
// Output (packed): lambda[m], V[m][m] (K0 = V diag(lambda) V^-1), Vinv[m][m], ctop (= ck of the last row).
static int vertical_operator_spectrum(int kmax, const double* stat, double dz, double h, double rkappa,
                                      std::vector<double>& out) {
    const int m = kmax - 2;
    std::vector<double> ck(m), dk(m);
    for (int r = 0; r < m; ++r) {
        const int k = r + 2;  // 1-based
        const double zk = dz * (double)(k - 1), zkp = dz * (double)k, zkm = dz * (double)(k - 2);
        const double zp = 0.5 * (zkp + zk), zm = 0.5 * (zkm + zk);
        const double statp = 0.5 * (stat[k] + stat[k - 1]), statm = 0.5 * (stat[k - 2] + stat[k - 1]);
        const double amp = std::exp(-zp / h) * std::exp(rkappa * zp / h) / statp;
        const double amm = std::exp(-zm / h) * std::exp(rkappa * zm / h) / statm;
        ck[r] = amp * std::exp(zk / h);
        dk[r] = amm * std::exp(zk / h);
    }
    // similarity scaling D: S = D^-1 K0 D symmetric with off-diagonal sqrt(ck_r * dk_{r+1})
    std::vector<double> D(m, 1.0), d(m), e(m, 0.0), z;
    for (int r = 0; r + 1 < m; ++r) {
        if (!(ck[r] > 0.0) || !(dk[r + 1] > 0.0)) return -2;  // needs a statically stable profile
        D[r + 1] = D[r] * std::sqrt(dk[r + 1] / ck[r]);
        e[r] = std::sqrt(ck[r] * dk[r + 1]);
    }
    for (int r = 0; r < m; ++r) d[r] = (r == m - 1) ? -dk[r] : -(ck[r] + dk[r]);
    if (symtridiag_eig(m, d, e, z)) return -1;
    out.assign((size_t)m + 2 * (size_t)m * m + 1, 0.0);
    double* lam = out.data();
    double* V = lam + m;
    double* Vi = V + (size_t)m * m;
    for (int r = 0; r < m; ++r) lam[r] = d[r];
    for (int r = 0; r < m; ++r)
        for (int c = 0; c < m; ++c) {
            V[(size_t)r * m + c] = D[r] * z[(size_t)r * m + c];        // V = D Z
            Vi[(size_t)c * m + r] = z[(size_t)r * m + c] / D[r];       // V^-1 = Z^T D^-1
        }
    out[(size_t)m + 2 * (size_t)m * m] = ck[m - 1];
    return 0;
}

struct SpectralProblem {
    const double* qref;    // [kmax][nd]  qref / sin(lat)
    const double* ckref;   // [kmax][nd]
    const double* tbar;    // [kmax][nd]
    const double* spec;    // lambda[m], V[m][m], Vinv[m][m], ctop
    double* g;             // scratch [m][n]   (rhs, then solution in level space)
    double* gh;            // scratch [m][n]   (mode space)
    double* cp;            // scratch [n][m]   Thomas c'
    double* pjk;           // out [kmax][n]
};

struct SpectralCommon {
    int n, m, nd, kmax, jb;
    const double *ajk, *bjk, *fact;      // [n]
    const double *topcoef, *topden;      // [n]  top-BC factors (compute_qref_and_fawa_first.f90:164-175)
    double fcoef, u1coef, u1off;
};

__global__ void __launch_bounds__(512)
nhn22_spectral_solve_kernel(SpectralCommon c, const SpectralProblem* __restrict__ probs) {
    const SpectralProblem p = probs[blockIdx.x];
    const int n = c.n, m = c.m, nd = c.nd, kmax = c.kmax;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* lam = p.spec;
    const double* V = lam + m;
    const double* Vi = V + (size_t)m * m;
    const double ctop = p.spec[(size_t)m + 2 * (size_t)m * m];
    // 1. right-hand side:  g[r][i] = f_ik - [i==0] b_0 u1_k - [r==m-1] fact_i * ctop * t_i
    for (int t = tid; t < m * n; t += nt) {
        const int r = t / n, i = t % n;
        const int k = r + 2;         // 1-based level
        const int jj = i + 2 + c.jb; // 1-based hemispheric row
        double f = c.fcoef * (p.qref[(size_t)(k - 1) * nd + jj] - p.qref[(size_t)(k - 1) * nd + jj - 2]);
        if (i == 0) {
            const double u1 = p.ckref[(size_t)(k - 1) * nd + c.jb] / c.u1coef - c.u1off;
            f = f - c.bjk[0] * u1;
        }
        double val = f;
        if (r == m - 1) {
            double tt = c.topcoef[i];
            tt = tt * (p.tbar[(size_t)(kmax - 1) * nd + i + 2] - p.tbar[(size_t)(kmax - 1) * nd + i]);
            tt = tt / c.topden[i];
            val = val - c.fact[i] * ctop * tt;
        }
        p.g[t] = val;
    }
    __syncthreads();
    // 2. to mode space: gh[mu][i] = sum_r Vinv[mu][r] g[r][i]
    for (int t = tid; t < m * n; t += nt) {
        const int mu = t / n, i = t % n;
        double s = 0.0;
        for (int r = 0; r < m; ++r) s += Vi[(size_t)mu * m + r] * p.g[(size_t)r * n + i];
        p.gh[t] = s;
    }
    __syncthreads();
    // 3. per mode, a scalar tridiagonal solve in latitude (diagonally dominant since lam <= 0):
    //    b_i w(i-1) + a_i w(i+1) + (fact_i*lam - (a_i+b_i)) w(i) = gh(i)
    //    Only m threads work here and the recurrence is sequential, so the operands of eight steps are loaded up
    //    front (independent of the recurrence) and the chain itself runs on registers.
    for (int mu = tid; mu < m; mu += nt) {
        const double l = lam[mu];
        double* __restrict__ gm = p.gh + (size_t)mu * n;
        double cprev = 0.0, dprev = 0.0;
        constexpr int U = 8;
        for (int i0 = 0; i0 < n; i0 += U) {
            double a[U], b[U], fc[U], g[U], cc[U], dd[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int i = min(i0 + q, n - 1);
                a[q] = c.ajk[i]; b[q] = c.bjk[i]; fc[q] = c.fact[i]; g[q] = gm[i];
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int i = i0 + q;
                const double diag = fc[q] * l - (a[q] + b[q]);
                const double den = diag - ((i > 0) ? b[q] * cprev : 0.0);
                cc[q] = a[q] / den;
                dd[q] = (g[q] - ((i > 0) ? b[q] * dprev : 0.0)) / den;
                if (i < n) { cprev = cc[q]; dprev = dd[q]; }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int i = i0 + q;
                if (i < n) { p.cp[(size_t)i * m + mu] = cc[q]; gm[i] = dd[q]; }
            }
        }
        double w = gm[n - 1];
        for (int i1 = n - 2; i1 >= 0; i1 -= U) {
            double g[U], cq[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int i = max(i1 - q, 0);
                g[q] = gm[i]; cq[q] = p.cp[(size_t)i * m + mu];
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int i = i1 - q;
                if (i >= 0) { w = g[q] - cq[q] * w; g[q] = w; }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int i = i1 - q;
                if (i >= 0) gm[i] = g[q];
            }
        }
    }
    __syncthreads();
    // 4. back to level space and assemble p_k (level 1 = 0, level kmax = level kmax-1 + t_i)
    for (int t = tid; t < m * n; t += nt) {
        const int r = t / n, i = t % n;
        double s = 0.0;
        for (int mu = 0; mu < m; ++mu) s += V[(size_t)r * m + mu] * p.gh[(size_t)mu * n + i];
        p.pjk[(size_t)(r + 1) * n + i] = s;
        if (r == 0) p.pjk[i] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        double tt = c.topcoef[i];
        tt = tt * (p.tbar[(size_t)(kmax - 1) * nd + i + 2] - p.tbar[(size_t)(kmax - 1) * nd + i]);
        tt = tt / c.topden[i];
        p.pjk[(size_t)(kmax - 1) * n + i] = p.pjk[(size_t)(kmax - 2) * n + i] + tt;
    }
}

}  // namespace falwa

// Debug/test hook (host only, no GPU needed): eigen-decomposition used by the NHN22 spectral solve.
// d[m], e[m-1] in; eigenvalues out in d, eigenvectors (row-major, columns) in z[m*m].
extern "C" int falwa_b200_debug_symtridiag_eig(int m, double* d, const double* e, double* z) {
    std::vector<double> dv(d, d + m), ev(e, e + m - 1), zv;
    const int rc = falwa::symtridiag_eig(m, dv, ev, zv);
    for (int i = 0; i < m; ++i) d[i] = dv[i];
    for (size_t i = 0; i < (size_t)m * m; ++i) z[i] = zv[i];
    return rc;
}
