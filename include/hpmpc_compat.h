/*
 * hpmpc_compat.h -- the HPMPC entry points on the Riccati / box-IPM hot path, exported with the
 * reference's exact names and argument lists so existing C callers link against libhpmpc_b200.so unchanged.
 *
 * Each prototype cites the reference declaration it replaces (paths relative to the HPMPC tree).
 * Semantics, argument meaning and return codes follow the reference; differences are listed in
 * INTEGRATION.md (ng must be 0, N2 is ignored, work buffers are not used, caller matrices are never
 * modified, `memory` holds the factor in this library's own layout).
 */
#ifndef HPMPC_B200_COMPAT_H
#define HPMPC_B200_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- backward Riccati recursion, lib4 storage : include/lqcp_solvers.h:37-45 ---- */
/* include/lqcp_solvers.h:37  (lqcp_solvers/d_back_ric_rec.c:43) */
int d_back_ric_rec_sv_tv_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng);
/* include/lqcp_solvers.h:39  (lqcp_solvers/d_back_ric_rec.c:79) */
int d_back_ric_rec_sv_tv_memory_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng);
/* include/lqcp_solvers.h:41  (lqcp_solvers/d_back_ric_rec.c:112) factor + solve */
void d_back_ric_rec_sv_tv_res(int N, int *nx, int *nu, int *nb, int **idxb, int *ng,
                              int update_b, double **hpBAbt, double **b,
                              int update_q, double **hpQ, double **q, double **bd,
                              double **hpDCt, double **Qx, double **qx,
                              double **hux, int compute_pi, double **hpi, int compute_Pb, double **hPb,
                              double *memory, double *work);
/* include/lqcp_solvers.h:43  (lqcp_solvers/d_back_ric_rec.c:403) factor */
void d_back_ric_rec_trf_tv_res(int N, int *nx, int *nu, int *nb, int **idxb, int *ng,
                               double **hpBAbt, double **hpQ, double **hpDCt, double **Qx, double **bd,
                               double *memory, double *work);
/* include/lqcp_solvers.h:45  (lqcp_solvers/d_back_ric_rec.c:564) solve with the stored factor */
void d_back_ric_rec_trs_tv_res(int N, int *nx, int *nu, int *nb, int **idxb, int *ng,
                               double **hpBAbt, double **hb, double **hq, double **hpDCt, double **qx,
                               double **hux, int compute_pi, double **hpi, int compute_Pb, double **hPb,
                               double *memory, double *work);

/* ---- box-constrained IPM, lib4 storage : include/mpc_solvers.h:41-42 ---- */
/* include/mpc_solvers.h:41  (mpc_solvers/d_ip2_res_hard.c:57) */
int d_ip2_res_mpc_hard_tv_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng);
/* include/mpc_solvers.h:42  (mpc_solvers/d_ip2_res_hard.c:116); returns 0 / 1 / 2 / -1 (:1331-1343) */
int d_ip2_res_mpc_hard_tv(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat,
                          int N, int *nx, int *nu_N, int *nb, int **idxb, int *ng,
                          double **pBAbt, double **pQ, double **pDCt, double **d, double **ux,
                          int compute_mult, double **pi, double **lam, double **t, double *double_work_memory);

/* include/mpc_solvers.h:46  (mpc_solvers/d_ip2_res_hard.c:1922): the last KKT system of the preceding d_ip2_res_mpc_hard_tv call
 * (same sizes, same double_work_memory) solved again for new b, q, d; ux, pi, lam, t receive the result */
void d_kkt_solve_new_rhs_res_mpc_hard_tv(int N, int *nx, int *nu_N, int *nb, int **idxb, int *ng, double **pBAbt, double **b,
                                         double **pQ, double **q, double **pDCt, double **d, double **ux, int compute_mult,
                                         double **pi, double **lam, double **t, double *double_work_memory);

/* include/mpc_solvers.h:45  (mpc_solvers/d_ip2_res_hard.c:1348): k_max Newton steps from the iterate (ux0, pi0, lam0, t0) */
int d_ip2_res_mpc_hard_tv_single_newton_step(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat,
                                             int N, int *nx, int *nu_N, int *nb, int **idxb, int *ng, double **pBAbt, double **pQ, double **pDCt,
                                             double **d, double **ux, int compute_mult, double **pi, double **lam, double **t,
                                             double *double_work_memory, double **ux0, double **pi0, double **lam0, double **t0);
/* include/mpc_solvers.h:47  (mpc_solvers/c99/d_res_ip_res_hard.c:39): residuals res_q, res_b, res_d, res_m and mu of a given point */
void d_res_res_mpc_hard_tv(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hpQ, double **hq,
                           double **hux, double **hpDCt, double **hd, double **hpi, double **hlam, double **ht, double *work,
                           double **hrq, double **hrb, double **hrd, double **hrm, double *mu);
/* include/mpc_solvers.h:36  (mpc_solvers/d_res_ip_hard.c:38): the exit residuals of the high-level wrappers */
void d_res_mpc_hard_tv(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hpQ, double **hq,
                       double **hux, double **hpDCt, double **hd, double **hpi, double **hlam, double **ht, double **hrq, double **hrb,
                       double **hrd, double *mu);

/* ---- high-level interface, dense stage-wise arrays : include/c_interface.h:59-67 ---- */
/* include/c_interface.h:59  (interfaces/c/c_interface_work_space.c:70) */
int hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2);
/* partial condensing, reference include/lqcp_solvers.h:86-95 (lqcp_solvers/d_part_cond.c:694, :739, :860, :926, :1072, :1103): panel-major
 * (lib4) arguments as in the reference; the condensing and the expansion run on the device as a batch of one (batched forms:
 * hpmpc_b200.h).  `memory` receives the condensed matrices, `work` is not used.  General constraints only at stage N. */
void d_part_cond_compute_problem_size(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int *nx2, int *nu2, int *nb2, int *ng2);
int d_part_cond_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int *nx2, int *nu2, int *nb2, int *ng2);
int d_part_cond_memory_space_size_bytes(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int *nx2, int *nu2, int *nb2, int *ng2);
void d_part_cond(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **hpBAbt, double **hpRSQrq, double **hpDCt, double **hd, int N2, int *nx2, int *nu2, int *nb2, int **hidxb2, int *ng2, double **hpBAbt2, double **hpRSQrq2, double **hpDCt2, double **hd2, void *memory, void *work);
int d_part_expand_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng);
void d_part_expand_solution(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **hpBAbt, double **hb, double **hpRSQrq, double **hrq, double **hpDCt, double **hux, double **hpi, double **hlam, double **ht, int N2, int *nx2, int *nu2, int *nb2, int **hidxb2, int *ng2, double **hux2, double **hpi2, double **hlam2, double **ht2, void *work);
/* include/c_interface.h:62  (interfaces/c/c_order_interface.c:53) row-major matrices */
int c_order_d_ip_ocp_hard_tv(int *kk, int k_max, double mu0, double mu_tol,
                             int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int warm_start,
                             double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r,
                             double **lb, double **ub, double **C, double **D, double **lg, double **ug,
                             double **x, double **u, double **pi, double **lam,
                             double *inf_norm_res, void *work0, double *stat);
/* include/c_interface.h:65  (interfaces/c/fortran_order_interface.c:53) column-major matrices */
int fortran_order_d_ip_ocp_hard_tv(int *kk, int k_max, double mu0, double mu_tol,
                                   int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int warm_start,
                                   double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r,
                                   double **lb, double **ub, double **C, double **D, double **lg, double **ug,
                                   double **x, double **u, double **pi, double **lam,
                                   double *inf_norm_res, void *work0, double *stat);

/* include/c_interface.h:63,67  (interfaces/c/c_order_interface.c:1082, fortran_order_interface.c:1082): the last KKT system of the
 * preceding {c,fortran}_order_d_ip_ocp_hard_tv call ON THE SAME work0 solved again for new b, q, r and bounds */
void c_order_d_solve_kkt_new_rhs_ocp_hard_tv(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **A, double **B, double **b,
                                             double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub,
                                             double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
                                             double **lam, double *inf_norm_res, double *work0);
void fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **A, double **B, double **b,
                                                   double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub,
                                                   double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
                                                   double **lam, double *inf_norm_res, double *work0);
/* include/c_interface.h:66  (interfaces/c/fortran_order_interface.c:695) */
int fortran_order_d_ip_ocp_hard_tv_single_newton_step(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu_N, int *nb,
                                                      int **hidxb, int *ng, int N2, int warm_start, double **A, double **B, double **b,
                                                      double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub,
                                                      double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
                                                      double **lam, double **t, double *inf_norm_res, void *work0, double *stat,
                                                      double **ux0, double **pi0, double **lam0, double **t0);

#ifdef __cplusplus
}
#endif
#endif
