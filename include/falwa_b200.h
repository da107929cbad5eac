/*
 * falwa_b200 — C ABI of the B200-native (sm_100a, fp64) QGField hot path.
 *
 * Drop-in boundary
 * ----------------
 * The reference (csyhuang/hn2016_falwa, falwa v2.3.3) crosses from Python into native code through
 * exactly nine F2PY functions (src/falwa/__init__.py:8-16, called from src/falwa/oopinterface.py).
 * Section A below exports one entry point per F2PY function with the same argument meaning
 * (f2py signatures: SURVEY.md Appendix B).  Section B exports the fused, batched, device-resident
 * entry points the host classes (hn2016_falwa_b200.QGFieldNH18 / QGFieldNHN22 / QGDataset) use.
 *
 * Conventions (all entry points)
 * ------------------------------
 *   * Every pointer is a DEVICE pointer to contiguous fp64 (double) data unless the name ends in
 *     `_host`.  The caller (PyTorch) owns all buffers; the library allocates only small,
 *     stream-ordered table/workspace buffers of its own.
 *   * Array layout is the reference's memory order ("Fortran order", first index fastest):
 *       3-D fields   (nlon, nlat, kmax)  == C array [kmax][nlat][nlon]   (longitude fastest)
 *       profiles     (nd|jd, kmax)       == C array [kmax][nd|jd]
 *       2-D fields   (nlon, nd|nlat)     == C array [nd|nlat][nlon]
 *     Index arguments (k, jb, jd, nd …) are the reference's 1-based Fortran values.
 *   * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   * Return value: 0 = ok; < 0 = invalid argument (FALWA_ERR_*); > 0 = cudaError_t.
 *     No exceptions cross this boundary.  There is no CPU fallback: without a CUDA device every
 *     compute entry point returns the CUDA error.
 *   * Regions the reference leaves uninitialised are written as 0.
 */
#ifndef FALWA_B200_H
#define FALWA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FALWA_OK 0
#define FALWA_ERR_ARG (-1)      /* bad dimension / index argument            */
#define FALWA_ERR_NULL (-2)     /* required pointer is NULL                  */
#define FALWA_ERR_UNSUPPORTED (-3) /* size outside what the kernels support  */

/* Library / build identification ("falwa_b200 <version> sm_100a"). */
const char* falwa_b200_version(void);

/* ======================================================================================
 * Section A — the nine F2PY routines (one snapshot, one hemisphere per call, like the reference)
 * ====================================================================================== */

/* replaces falwa.compute_qgpv  (f90_modules/compute_qgpv.f90:1-137; oopinterface.py:1574) */
int falwa_b200_compute_qgpv(int nlon, int nlat, int kmax, const double* ut, const double* vt,
                            const double* theta, const double* height_host, const double* t0_host,
                            const double* stat_host, double aa, double omega, double dz, double hh,
                            double rr, double cp, double* pv, double* avort, void* stream);

/* replaces falwa.compute_qgpv_direct_inv  (compute_qgpv_direct_inv.f90:1-112; oopinterface.py:1748) */
int falwa_b200_compute_qgpv_direct_inv(int nlon, int nlat, int kmax, int jd, const double* uq,
                                       const double* vq, const double* tq, const double* height_host,
                                       const double* ts0_host, const double* tn0_host,
                                       const double* stats_host, const double* statn_host, double aa,
                                       double omega, double dz, double hh, double rr, double cp,
                                       double* pv, double* avort, void* stream);

/* replaces falwa.compute_reference_states  (compute_reference_states.f90:1-261; oopinterface.py:1656)
 * qref,u,tref: (jd,kmax).  num_of_iter_host receives the SOR iteration count (synchronises).
 * bins (optional, int32 (nlon,nlat,kmax), may be NULL) receives the 1-based area-analysis bin of
 * every grid point on levels 2..kmax-1 (for exact bin-parity checks). */
int falwa_b200_compute_reference_states(const double* pv, const double* uu, const double* pt,
                                        const double* stat_host, int nlon, int nlat, int kmax, int jd,
                                        int npart, int maxits, double a, double om, double dz, double eps,
                                        double h, double dphi, double dlambda, double r, double cp,
                                        double rjac, double* qref, double* u, double* tref,
                                        int* num_of_iter_host, int32_t* bins, void* stream);

/* replaces falwa.compute_qref_and_fawa_first  (compute_qref_and_fawa_first.f90:1-176; oopinterface.py:1812)
 * qref,ubar,tbar,fawa,ckref: (nd,kmax); tjk: (jd-2,kmax-1); sjk: (jd-2,jd-2,kmax-1).
 * bins_pv / bins_vort: optional int32 (imax,jmax,kmax) bin indices, may be NULL. */
int falwa_b200_compute_qref_and_fawa_first(const double* pv, const double* uu, const double* vort,
                                           const double* pt, const double* tn0_host, int imax, int jmax,
                                           int kmax, int nd, int nnd, int jb, int jd, double a, double omega,
                                           double dz, double h, double dphi, double dlambda, double rr,
                                           double cp, double* qref, double* ubar, double* tbar, double* fawa,
                                           double* ckref, double* tjk, double* sjk, int32_t* bins_pv,
                                           int32_t* bins_vort, void* stream);

/* replaces falwa.matrix_b4_inversion  (matrix_b4_inversion.f90:1-88; oopinterface.py:1840)
 * qjj,djj,cjj: (jd-2,jd-2); rj: (jd-2). */
int falwa_b200_matrix_b4_inversion(int k, int jmax, int kmax, int nd, int jb, int jd, const double* z_host,
                                   const double* statn_host, const double* qref, const double* ckref,
                                   const double* sjk, double a, double om, double dz, double h, double rr,
                                   double cp, double* qjj, double* djj, double* cjj, double* rj, void* stream);

/* replaces falwa.matrix_after_inversion  (matrix_after_inversion.f90:1-39; oopinterface.py:1863)
 * qjj is the INVERSE of the matrix returned by matrix_b4_inversion; sjk,tjk are updated in place. */
int falwa_b200_matrix_after_inversion(int k, int kmax, int jd, const double* qjj, const double* djj,
                                      const double* cjj, const double* rj, double* sjk, double* tjk,
                                      void* stream);

/* replaces falwa.upward_sweep  (upward_sweep.f90:1-92; oopinterface.py:1872)
 * tref,u: (jd,kmax); qref: (nd,kmax). */
int falwa_b200_upward_sweep(int jmax, int kmax, int nd, int jb, int jd, const double* sjk, const double* tjk,
                            const double* ckref, const double* tb_host, const double* qref_over_cor,
                            double* tref, double* qref, double* u, double a, double om, double dz, double h,
                            double rr, double cp, void* stream);

/* replaces falwa.compute_flux_dirinv_nshem  (compute_flux_dirinv.f90:1-182; oopinterface.py:430, 1094)
 * 3-D outputs (imax,nd,kmax); ep4 (imax,nd).  mask_count (optional device int64[2], may be NULL)
 * receives the number of (i,j,jj,k) pairs in the anticyclonic / cyclonic sums.
 * method: 0 = automatic (interval/prefix-sum LWA where qref is monotone, brute force otherwise),
 *         1 = force the O(nd^2) brute-force double loop of the reference. */
int falwa_b200_compute_flux_dirinv_nshem(const double* pv, const double* uu, const double* vv,
                                         const double* pt, const double* ncforce, const double* tn0_host,
                                         const double* qref, const double* uref, const double* tref, int imax,
                                         int jmax, int kmax, int nd, int jb, int jd, int is_nhem, double a,
                                         double om, double dz, double h, double rr, double cp, double prefac,
                                         double* astar1, double* astar2, double* ncforce3d, double* ua1,
                                         double* ua2, double* ep1, double* ep2, double* ep3, double* ep4,
                                         int64_t* mask_count, int method, void* stream);

/* replaces falwa.compute_lwa_only_nhn22  (compute_lwa_only_nhn22.f90:1-107; oopinterface.py:652, 675)
 * astarbaro,ubaro: (imax,nd); astar1,astar2: (imax,nd,kmax). */
int falwa_b200_compute_lwa_only_nhn22(const double* pv, const double* uu, const double* qref, int imax,
                                      int jmax, int kmax, int nd, int jb, int is_nhem, double a, double om,
                                      double dz, double h, double rr, double cp, double prefac,
                                      double* astarbaro, double* ubaro, double* astar1, double* astar2,
                                      int method, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FALWA_B200_H */
