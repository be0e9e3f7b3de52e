/*
 * hpmpc_blasfeo_compat.h -- the two BLASFEO container types HPMPC's "libstr" entry points take, and those entry points on the
 * Riccati / box-IPM hot path.
 *
 * BLASFEO (github.com/giaf/blasfeo) is an external dependency of the reference that is neither vendored nor version-pinned
 * (Makefile.rule:47-48); HPMPC itself touches only .pA / .pa / .memsize (mpc_solvers/d_tree_ip2_res_hard_libstr.c:264,
 * mpc_solvers/c99/d_aux_ip_hard_libstr.c:58, lqcp_solvers/d_back_ric_rec_libstr.c:109).  A caller that includes the real
 * blasfeo_common.h gets the real definitions (this header then only declares the functions); a caller without BLASFEO gets the
 * layout below, which is the one of the BLASFEO generation HPMPC's `blasfeo_`-prefixed calls belong to (0.1.x, 2018; panel-major
 * "high-performance" target: bs = 4, element (i,j) of a matrix at pA[(i/4)*4*cn + i%4 + 4*j], cn = padded number of columns).
 * PARITY UNPINNED for anything specific to this flavour (no BLASFEO in the build image, no test of the reference pins it): the
 * arithmetic is that of the lib4 twins, which IS pinned; the conventions followed are cited per function.
 */
#ifndef HPMPC_BLASFEO_COMPAT_H
#define HPMPC_BLASFEO_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(BLASFEO_COMMON_H_) && !defined(HPMPC_B200_HAVE_BLASFEO)
struct blasfeo_dmat
	{
	int m;          /* rows */
	int n;          /* cols */
	int pm;         /* packed number of rows    */
	int cn;         /* packed number of columns */
	double *pA;     /* pm*cn doubles, panel-major (bs = 4) */
	double *dA;     /* inverse diagonal (min(m,n) doubles) */
	int use_dA;
	int memsize;    /* bytes */
	};
struct blasfeo_dvec
	{
	int m;          /* size */
	int pm;         /* packed size */
	double *pa;
	int memsize;    /* bytes */
	};
#endif

/* include/lqcp_solvers.h:50  (lqcp_solvers/d_back_ric_rec_libstr.c:39) */
int d_back_ric_rec_work_space_size_bytes_libstr(int N, int *nx, int *nu, int *nb, int *ng);
/* include/lqcp_solvers.h:52  (lqcp_solvers/d_back_ric_rec_libstr.c:76): hsb[n] edge n, hspi[n+1] / hsPb[n+1] node-indexed (:145,:196),
 * hsQx[n] / hsqx[n] = [bounds (nb) | general (ng)] (:102,:110); the factor goes to hsL[n].pA in this library's own layout */
void d_back_ric_rec_sv_libstr(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int update_b, struct blasfeo_dmat *hsBAbt,
                              struct blasfeo_dvec *hsb, int update_q, struct blasfeo_dmat *hsRSQrq, struct blasfeo_dvec *hsrq,
                              struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsQx, struct blasfeo_dvec *hsqx, struct blasfeo_dvec *hsux,
                              int compute_pi, struct blasfeo_dvec *hspi, int compute_Pb, struct blasfeo_dvec *hsPb,
                              struct blasfeo_dmat *hsL, void *work_space);
/* include/lqcp_solvers.h:54  (lqcp_solvers/d_back_ric_rec_libstr.c:229) */
void d_back_ric_rec_trf_libstr(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, struct blasfeo_dmat *hsBAbt,
                               struct blasfeo_dmat *hsRSQrq, struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsQx,
                               struct blasfeo_dmat *hsL, void *work);
/* include/lqcp_solvers.h:56  (lqcp_solvers/d_back_ric_rec_libstr.c:309) */
void d_back_ric_rec_trs_libstr(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, struct blasfeo_dmat *hsBAbt, struct blasfeo_dvec *hsb,
                               struct blasfeo_dvec *hsrq, struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsqx, struct blasfeo_dvec *hsux,
                               int compute_pi, struct blasfeo_dvec *hspi, int compute_Pb, struct blasfeo_dvec *hsPb,
                               struct blasfeo_dmat *hsL, void *work);
/* include/mpc_solvers.h:52  (mpc_solvers/d_ip2_res_hard_libstr.c:43) */
int d_ip2_res_mpc_hard_work_space_size_bytes_libstr(int N, int *nx, int *nu, int *nb, int *ng);
/* include/mpc_solvers.h:53  (mpc_solvers/d_ip2_res_hard_libstr.c:92): hsd / hslam / hst = [lb(nb) lg(ng) ub(nb) ug(ng)] UNPADDED
 * (interfaces/c/fortran_order_interface_libstr.c:408-415, :751-755), hspi[n+1] = multiplier of the dynamics into stage n+1 */
int d_ip2_res_mpc_hard_libstr(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat, int N,
                              int *nx, int *nu, int *nb, int **idxb, int *ng, struct blasfeo_dmat *hsBAbt, struct blasfeo_dmat *hsRSQrq,
                              struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsd, struct blasfeo_dvec *hsux, int compute_mult,
                              struct blasfeo_dvec *hspi, struct blasfeo_dvec *hslam, struct blasfeo_dvec *hst, void *work_memory);

#ifdef __cplusplus
}
#endif
#endif
