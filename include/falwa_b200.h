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
 * method: 0 = automatic: exact windowed sums (lwa_window.cu; fields 16-byte aligned, nd <= 384, even imax), else the
 *             interval scan (nd <= 416), else the double loop;
 *         1 = the O(nd^2) double loop of the reference (same summation order as the reference);
 *         2 = interval scan (flux_scan.cu);  3 = windowed sums (falls back like 0 where unsupported);
 *         4 = prefix sums over threshold intervals (lwa_prefix.cu: 64-bit fixed-point difference arrays, cost independent
 *             of how far the PV contours are displaced; it does not carry ncforce — ncforce3d then comes from the
 *             windowed sums; same support conditions as 3).
 * All methods form exactly the reference's sets (equal pair counts); sums agree within rounding. */
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

/* ======================================================================================
 * Section B — fused, batched, device-resident pipeline (what QGFieldNH18 / QGFieldNHN22 /
 * QGDataset of hn2016_falwa_b200 call).  A plan holds the grid-dependent tables on the device.
 * All 3-D fields are [batch][level][lat south->north][lon]; reference-state profiles are
 * [batch][kmax][nlat]; the southern hemisphere is handled by index arithmetic and sign factors
 * instead of the reference's flipped/negated copies (oopinterface.py:764-792, :1794-1806).
 * ====================================================================================== */
typedef struct falwa_plan falwa_plan;

/* kind: 0 = QGFieldNH18 (SOR reference state, jb forced to 0), 1 = QGFieldNHN22 (direct inversion).
 * Host tables supplied by the caller so that they are bit-identical to the reference's NumPy values:
 *   plev_to_height_host[nlev] = -H log(plev/1000)        (oopinterface.py:120)
 *   theta_factor_host[nlev]   = exp(R/cp * z_p / H)       (oopinterface.py:127-128)
 *   clat_host[nlat]           = |cos(deg2rad(ylat))|      (oopinterface.py:342)
 *   sinlat_host[nlat]         = sin(deg2rad(ylat))        (data_storage.py:175-176)
 * npart <= 0 means nlat.  prefactor: oopinterface.py:263-270. */
int falwa_b200_plan_create(falwa_plan** plan, int kind, int nlon, int nlat, int nlev, int kmax, int jb,
                           int npart, int maxit, const double* plev_to_height_host,
                           const double* theta_factor_host, const double* clat_host,
                           const double* sinlat_host, double a, double omega, double dz, double h, double rr,
                           double cp, double tol, double rjac, double prefactor);
int falwa_b200_plan_destroy(falwa_plan* plan);

/* Hemispheric cos-weighted mean potential temperature on the input pressure levels — the input of the
 * host-side static-stability spline (replaces the NumPy reductions of oopinterface.py:250-256).
 * t_in: temperature [batch][nlev][nlat][nlon]; means: [batch][2][nlev] (0 = south, 1 = north). */
int falwa_b200_theta_hemispheric_means(const falwa_plan* plan, int batch, const double* t_in, double* means,
                                       void* stream);

/* interpolate_fields (oopinterface.py:467-535): vertical interpolation of u, v, theta (theta = T * factor)
 * onto the uniform pseudo-height grid, absolute vorticity and QGPV.
 * profiles_host [batch][4][kmax] = t0_s, t0_n, stat_s, stat_n evaluated at the height grid (for NH18 pass
 * the hemispheric averages in both slots).  If already_on_height_grid != 0 the interpolation is skipped
 * and u, v, theta must already hold the fields on the height grid.
 * zm [batch][3][kmax][nlat] and ext (uint64 [batch][2][kmax][2]) are optional outputs (both or neither): the zonal
 * means of qgpv, u, theta and the per-level extrema of qgpv and avort, accumulated while the fields are produced so
 * that falwa_b200_reference_states need not read the 3-D fields again. */
int falwa_b200_interpolate_fields(const falwa_plan* plan, int batch, const double* u_in, const double* v_in,
                                  const double* t_in, int already_on_height_grid, const double* profiles_host,
                                  double* u, double* v, double* theta, double* avort, double* qgpv,
                                  double* zm, void* ext, void* stream);

/* compute_reference_states (oopinterface.py:550-615, :1604-1674, :1774-1888), both hemispheres, merged
 * into global-latitude storage like data_storage.py:32-62 (north first, south second).
 *   qref_st: qref as the reference stores it (qref/f); qref_cu: qref in s^-1 (data_storage.py:170-179).
 *   num_of_iter_host [batch][2]: NH18 SOR iteration counts (north, south); 0 for NHN22.
 * north_only != 0 computes the northern hemisphere only.  zm / ext: the statistics written by
 * falwa_b200_interpolate_fields, or NULL (recomputed from the fields).  Synchronises the stream before returning. */
int falwa_b200_reference_states(const falwa_plan* plan, int batch, const double* qgpv, const double* u,
                                const double* theta, const double* avort, const double* profiles_host,
                                int north_only, double* qref_st, double* qref_cu, double* uref_st,
                                double* ptref_st, int* num_of_iter_host, const double* zm, const void* ext,
                                void* stream);

/* compute_lwa_and_barotropic_fluxes (oopinterface.py:716-982), both hemispheres.
 *   lwa   [batch][kmax][nlat][nlon]
 *   out2d [batch][FALWA_OUT2D_COUNT][nlat][nlon], slots below.
 * ncforce may be NULL (zeros).  method as for falwa_b200_compute_flux_dirinv_nshem.  0 = automatic: a sampling kernel
 * measures, per (snapshot, hemisphere), how many rows the PV contours are displaced on average; weakly displaced ones
 * go to the windowed sums, strongly meandering ones to the prefix sums (with ncforce: always the windowed sums).  Both
 * kernels also carry the vertical sums of astar1, astar2, ua1, ua2 and write the 3-D LWA in the same pass. */
enum {
    FALWA_OUT2D_ASTAR1_BARO = 0, FALWA_OUT2D_ASTAR2_BARO = 1, FALWA_OUT2D_LWA_BARO = 2,
    FALWA_OUT2D_ADV_FLUX_F1 = 3, FALWA_OUT2D_ADV_FLUX_F2 = 4, FALWA_OUT2D_ADV_FLUX_F3 = 5,
    FALWA_OUT2D_EP2_BARO = 6, FALWA_OUT2D_EP3_BARO = 7, FALWA_OUT2D_MERIDIONAL_HEAT_FLUX = 8,
    FALWA_OUT2D_NCFORCE_BARO = 9, FALWA_OUT2D_U_BARO = 10, FALWA_OUT2D_CONVERGENCE_ZONAL_ADVECTIVE_FLUX = 11,
    FALWA_OUT2D_DIVERGENCE_EDDY_MOMENTUM_FLUX = 12, FALWA_OUT2D_FLUX_VECTOR_LAMBDA_BARO = 13,
    FALWA_OUT2D_FLUX_VECTOR_PHI_BARO = 14, FALWA_OUT2D_COUNT = 15
};
int falwa_b200_lwa_and_barotropic_fluxes(const falwa_plan* plan, int batch, const double* qgpv, const double* u,
                                         const double* v, const double* theta, const double* ncforce,
                                         const double* profiles_host, const double* qref_cu,
                                         const double* uref_st, const double* ptref_st, int north_only,
                                         double* lwa, double* out2d, int method, void* stream);

/* compute_layerwise_lwa_fluxes of a batch (oopinterface.py:1063-1156, stretching term :1021-1061; utilities.py:
 * 305-344).  out3d [batch][10][kmax][nlat][nlon]: astar1, astar2, ncforce, ua1, ua2, ep1, ep2, ep3, stretch_term,
 * lwa (= astar1 + astar2) in the global frame (southern hemisphere: ep2 / ep3 swapped and negated, ncforce negated; the equator row holds the
 * southern value).  Other arguments as for falwa_b200_lwa_and_barotropic_fluxes. */
int falwa_b200_layerwise_lwa_fluxes(const falwa_plan* plan, int batch, const double* qgpv, const double* u,
                                    const double* v, const double* theta, const double* ncforce,
                                    const double* profiles_host, const double* qref_cu, const double* uref_st,
                                    const double* ptref_st, int north_only, double* out3d, int method, void* stream);

/* compute_ncforce_from_heating_rate of a batch (oopinterface.py:985-1019): heating_rate [batch][nlev][nlat][nlon] on the
 * pressure levels (interpolated like the fields) or, with on_height_grid != 0, [batch][kmax][nlat][nlon];
 * ncforce [batch][kmax][nlat][nlon]. */
int falwa_b200_ncforce_from_heating_rate(const falwa_plan* plan, int batch, const double* heating_rate,
                                         int on_height_grid, const double* profiles_host, double* ncforce,
                                         void* stream);

/* ======================================================================================
 * Section C — 2-D (barotropic) path: BarotropicField.equivalent_latitudes / .lwa for a batch of fields
 * (replaces src/falwa/barotropic_field.py:104-125 -> basis.eqvlat_fawa(output_fawa=False), basis.lwa;
 * basis.py:11-64, :137-205).  vort [batch][nlat][nlon]; area [nlat][nlon]; ylat_host (degrees, ascending) and
 * dmu_host = a*cos(lat)*dphi, [nlat] each.  qref [batch][nlat]; lwa [batch][nlat][nlon], or [batch][2][nlat][nlon]
 * (anticyclonic, cyclonic) when partitioned != 0; lwa may be NULL (equivalent latitude only).
 * ====================================================================================== */
int falwa_b200_barotropic_eqvlat_lwa(int batch, int nlon, int nlat, int n_points, const double* vort,
                                     const double* area, const double* ylat_host, const double* dmu_host,
                                     double planet_radius, double* qref, double* lwa, int partitioned, void* stream);

/* The two halves separately, as the hemispheric callers of src/falwa/wrapper.py:89-574 use them.
 * falwa_b200_eqvlat_2d: mode 0 = basis.eqvlat_fawa (basis.py:11-64): qref [batch][nlat]; aux = fawa [batch][nlat-1]
 *   or NULL (output_fawa=False).  mode 1 = basis.eqvlat_vgrad (basis.py:67-134): aux = brac [batch][nlat] together
 *   with vgrad [batch][nlat][nlon], or both NULL.
 * falwa_b200_lwa_2d: basis.lwa (basis.py:137-205) with a prescribed Q(y): qref [batch][nlat] (device), dmu_host
 *   [nlat]; ncforce and bigsigma [batch][nlat][nlon] both given or both NULL; lwa as above. */
int falwa_b200_eqvlat_2d(int mode, int batch, int nlon, int nlat, int n_points, const double* vort,
                         const double* area, const double* vgrad, const double* ylat_host, double planet_radius,
                         double* qref, double* aux, void* stream);
int falwa_b200_lwa_2d(int batch, int nlon, int nlat, const double* vort, const double* qref, const double* dmu_host,
                      const double* ncforce, double* lwa, double* bigsigma, int partitioned, void* stream);

/* ======================================================================================
 * Section D — ingest of archived fields (the data format in front of the path): raw values as stored in the file
 * -> float64 [batch][nlev][nlat][nlon], latitude ascending, pressure descending.  Replaces, on the device, xarray's
 * CF decoding (value = raw * scale_factor + add_offset in float64, _FillValue -> NaN) and the np.flip reordering of
 * src/falwa/xarrayinterface.py:384-405 (recipe for the packed test snapshot: tests/test_output_results.py:26-45).
 * src_type: 0 float64, 1 float32, 2 int16; big_endian != 0: the stored values are big-endian (NetCDF-3 classic /
 * 64-bit-offset files).  Pass scale = 1, offset = 0 for unpacked values.  src, dst: device pointers, same shape,
 * not overlapping.
 * ====================================================================================== */
int falwa_b200_ingest_field(int src_type, int batch, int nlev, int nlat, int nlon, const void* src, int big_endian,
                            double scale, double offset, int has_fill, double fill_value, int flip_lat, int flip_lev,
                            double* dst, void* stream);

/* Diagnostics used by bench.py (not on the data path): number of kernels launched by this library so far;
 * optional per-kernel timing with CUDA events on the launching stream.  profile_read synchronises the device
 * and writes one line "kernel_name launches total_ms" per kernel into buf, then clears the records. */
long long falwa_b200_launch_count(void);
int falwa_b200_profile_enable(int on);
int falwa_b200_profile_read(char* buf, int buflen);

/* Host-only test hook: eigen-decomposition of a symmetric tridiagonal matrix (used by the NHN22 solve). */
int falwa_b200_debug_symtridiag_eig(int m, double* d, const double* e, double* z);

/* FP64 pipe micro-benchmark (tools/fp64_peak.py): `blocks` CTAs x 256 threads x 8 * iters dependent-chain DFMAs;
 * elapsed milliseconds of one launch in *ms_out.  sink_dev: one device double (never written in practice). */
int falwa_b200_debug_dfma_rate(int blocks, int iters, double* sink_dev, float* ms_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FALWA_B200_H */
