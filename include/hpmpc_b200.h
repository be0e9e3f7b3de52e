/*
 * hpmpc_b200.h -- batched C ABI of the B200-native Riccati / box-IPM engine.
 *
 * The reference (HPMPC) solves ONE optimal-control problem per call on one CPU thread:
 *     d_back_ric_rec_{sv,trf,trs}_tv_res      reference include/lqcp_solvers.h:37-45
 *     d_ip2_res_mpc_hard_tv                   reference include/mpc_solvers.h:41-42
 *     {c,fortran}_order_d_ip_ocp_hard_tv      reference include/c_interface.h:59-67
 * Those symbols are exported unchanged by this library (see hpmpc_compat.h) as a batch of one.
 * The entry points below are the same operations with a leading instance dimension: n_inst independent
 * problems that share one size pattern (N, nx[], nu[], nb[], idxb[][]) and differ in their numbers.
 *
 * All pointers are plain C; "d_" arguments are device pointers on the handle's device, "h_" arguments are
 * host pointers (pinned memory makes the copies asynchronous).  `stream` is a cudaStream_t passed as void*
 * (NULL = default stream).  Every function returns 0 on success, a negative value on error
 * (and prints the reason to stderr); nothing falls back to the CPU.
 *
 * Native packed instance layout: see hpmpc_b200/csrc/layout.h; offsets are queried with
 * hpmpc_b200_ocp_stage_offsets().  Outputs per instance:
 *     ux  [ux_stride]   u_n then x_n for n = 0..N        (reference hux[n], d_back_ric_rec.c:341)
 *     pi  [pi_stride]   multiplier of x_{n+1} = ..., n = 0..N-1   (reference hpi[n], edge-indexed)
 *     lam [lam_stride]  per stage [lb(nb) ub(nb) lg(ng) ug(ng)]  (reference lib4 ordering, c_order_interface.c:662-681)
 *     t   [lam_stride]  slacks, same ordering
 *     info[6+5*k_max]   kk, status(0 converged /1 k_max /2 alpha_min /-1), ||rq||inf, ||rb||inf, ||rd||inf, mu,
 *                       then the reference's stat table (sigma, alpha_aff, mu_aff, alpha, mu) per iteration
 */
#ifndef HPMPC_B200_H
#define HPMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hpmpc_b200_ocp hpmpc_b200_ocp;

typedef struct hpmpc_b200_sizes
	{
	long long in_stride, ux_stride, pi_stride, lam_stride, L_stride, ipm_work_stride;
	int N, nzM, nxM, nbtot;
	int grid, warps_per_cta, n_slots, smem_per_cta;
	int fast_variant;          /* >= 0 when a size-specialised kernel serves this pattern */
	int ipm_grid, ipm_warps_per_cta;   /* launch shape of the IPM kernel; one wave = ipm_grid*ipm_warps_per_cta instances */
	int ipm_fast_variant;      /* >= 0 when the IPM uses the size-specialised sweeps */
	} hpmpc_b200_sizes;

/* nu has N entries (nu[N] is taken as 0, like the reference high-level API, c_order_interface.c:78-81);
 * nb / hidxb may be NULL for an unconstrained problem; hidxb[n][j] indexes [u_n ; x_n] */
int  hpmpc_b200_ocp_create(hpmpc_b200_ocp **out, int N, const int *nx, const int *nu, const int *nb,
                           int *const *hidxb, int device);
/* the same with general (polytopic) constraints lg <= D_n u + C_n x <= ug, ng[n] of them at stage n (NULL = none): SURVEY 8f row f1,
 * reference lqcp_solvers/d_back_ric_rec_libstr.c:105-113,164-171, mpc_solvers/c99/d_aux_ip_hard_libstr.c:125,329,
 * mpc_solvers/d_res_ip_res_hard_libstr.c:120-144.  Patterns with ng > 0 run on the any-size kernels. */
int  hpmpc_b200_ocp_create_gen(hpmpc_b200_ocp **out, int N, const int *nx, const int *nu, const int *nb,
                               int *const *hidxb, const int *ng, int device);
/* the same for a UNIFORM pattern (nx[0] = 0, nx[1..N] = nx, nu[0..N-1] = nu, bounds only) whose (nx, nu) is not one of the compiled
 * shapes (8,3), (12,5), (24,11) (and (4,2) without bounds): the pattern is embedded in the smallest compiled shape (NX, NU) >= (nx, nu)
 * with decoupled dummy inputs and states, so the size-specialised kernels serve it.  The handle describes the PADDED frame (sizes,
 * stage offsets; per stage ux = [u (NU slots) x (NX slots)], pi NX slots, real entries first); hpmpc_b200_pack_instance and
 * hpmpc_b200_unpack_solution take / return the caller's real sizes.  Any other pattern gets a plain handle. */
int  hpmpc_b200_ocp_create_padded(hpmpc_b200_ocp **out, int N, const int *nx, const int *nu, const int *nb,
                                  int *const *hidxb, int device);
int  hpmpc_b200_ocp_padded_shape(const hpmpc_b200_ocp *p, int *NX, int *NU);      /* 1 when padded, else 0 */
/* CONCURRENCY: a handle owns one set of scratch slots and one work-queue counter, so it runs ONE call at a time.  The library
 * enforces this on the device: every entry point makes its stream wait for the handle's previous call (an event), so calls issued
 * to different streams are serialised, never raced.  Use one handle per stream for concurrent solves.  Host threads must not
 * call into the same handle simultaneously. */
void hpmpc_b200_ocp_destroy(hpmpc_b200_ocp *p);
/* ctas_per_sm <= 0 or warps_per_cta <= 0 selects the default for the problem size */
int  hpmpc_b200_ocp_set_launch(hpmpc_b200_ocp *p, int ctas_per_sm, int warps_per_cta);
void hpmpc_b200_ocp_sizes(const hpmpc_b200_ocp *p, hpmpc_b200_sizes *out);
void hpmpc_b200_ocp_stage_offsets(const hpmpc_b200_ocp *p, int n, int *off_BAbt, int *off_RSQ, int *off_d,
                                  int *off_ux, int *off_pi, int *off_lam, int *off_L);

/* host-side packing of one instance from the reference's stage-wise arrays (c_order = 0: column-major as
 * fortran_order_d_ip_ocp_hard_tv takes them; 1: row-major as c_order_d_ip_ocp_hard_tv).  lb/ub may be NULL. */
int  hpmpc_b200_pack_instance(const hpmpc_b200_ocp *p, int c_order, double *const *A, double *const *B, double *const *b,
                              double *const *Q, double *const *S, double *const *R, double *const *q, double *const *r,
                              double *const *lb, double *const *ub, double *block);
/* general constraints of one instance into its block (after hpmpc_b200_pack_instance): C[n] ng x nx, D[n] ng x nu (n < N) */
int  hpmpc_b200_pack_general(const hpmpc_b200_ocp *p, int c_order, double *const *C, double *const *D, double *const *lg,
                             double *const *ug, double *block);
void hpmpc_b200_ocp_general_offsets(const hpmpc_b200_ocp *p, int n, int *ng, int *off_DCt, int *off_dg, int *off_c);
void hpmpc_b200_unpack_solution(const hpmpc_b200_ocp *p, const double *ux, const double *pi, const double *lam,
                                double **x, double **u, double **pi_out, double **lam_out);

/* ---- data resident in HBM ---- */
int hpmpc_b200_d_back_ric_rec_sv_batch (hpmpc_b200_ocp *p, long long n_inst, const double *d_in,
                                        double *d_ux, double *d_pi, double *d_Pb /* may be NULL */, void *stream);
int hpmpc_b200_d_back_ric_rec_trf_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, double *d_L, void *stream);
/* the reference's Qx / qx arguments (d_back_ric_rec_sv_tv_res, lqcp_solvers/d_back_ric_rec.c:112): per-constraint terms added to
 * the Hessian diagonal / gradient row (through [D C]' for general constraints); nbtot doubles per instance in the flat
 * constraint order (stage after stage: box entries, then general ones); either pointer may be NULL.  These run on the any-size
 * kernels; trf / trs with updates need hpmpc_b200_ocp_generic_factor_layout() on shapes that have size-specialised sweeps. */
int hpmpc_b200_d_back_ric_rec_sv_upd_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_Qx,
                                           const double *d_qx, double *d_ux, double *d_pi, double *d_Pb, void *stream);
int hpmpc_b200_d_back_ric_rec_trf_upd_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_Qx,
                                            double *d_L, void *stream);
int hpmpc_b200_d_back_ric_rec_trs_upd_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_L,
                                            const double *d_qx, double *d_ux, double *d_pi, void *stream);
void hpmpc_b200_ocp_generic_factor_layout(hpmpc_b200_ocp *p);
int hpmpc_b200_d_back_ric_rec_trs_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_L,
                                        double *d_ux, double *d_pi, void *stream);
int hpmpc_b200_d_ip2_res_mpc_hard_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0,
                                        double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi,
                                        double *d_lam, double *d_t, double *d_info, void *stream);

/* ---- residuals of a given point (row a6): d_res_res_mpc_hard_tv, mpc_solvers/c99/d_res_ip_res_hard.c:39 ----
 * res_q [ux_stride], res_b [pi_stride], res_d and res_m [lam_stride, in the layout of lam], mu [1] per instance; d_rm may be NULL */
int hpmpc_b200_d_res_res_mpc_hard_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_ux,
                                        const double *d_pi, const double *d_lam, const double *d_t, double *d_rq, double *d_rb,
                                        double *d_rd, double *d_rm, double *d_mu, void *stream);

/* ---- a fixed number of Newton steps from a given iterate ----
 * reference: d_ip2_res_mpc_hard_tv_single_newton_step, mpc_solvers/d_ip2_res_hard.c:1348 (high level include/c_interface.h:66).
 * d_ux, d_pi, d_lam, d_t: in = (ux0, pi0, lam0, t0), out = the updated iterate; k_max residual-based steps, centering term mu0.
 * d_info as for the IPM (status is 1 after k_max steps, like the reference's return value). */
int hpmpc_b200_d_ip2_res_mpc_hard_single_newton_step_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max,
                                                           double mu0, double alpha_min, double *d_ux, double *d_pi,
                                                           double *d_lam, double *d_t, double *d_info, void *stream);

/* ---- re-solve with a new right-hand side on the IPM's last factorisation ----
 * reference: d_kkt_solve_new_rhs_res_mpc_hard_tv, mpc_solvers/d_ip2_res_hard.c:1922 (high level: include/c_interface.h:63,67).
 * The reference keeps the factor, t_inv and the backed-up iterate in the caller's work memory between the two calls; here the
 * IPM call writes them to d_kkt (hpmpc_b200_kkt_state_stride() doubles per instance), and the second call reads them there.
 * d_in of the second call holds the new b, [r q] and bounds in the same packed layout (its matrices must be those of the
 * first call).  d_info of the second call: 6 doubles per instance, [1] = 0, or -10 when the IPM ran no phase-2 iteration. */
long long hpmpc_b200_kkt_state_stride(const hpmpc_b200_ocp *p);
int hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0,
                                            double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi,
                                            double *d_lam, double *d_t, double *d_info, double *d_kkt, void *stream);
int hpmpc_b200_d_kkt_solve_new_rhs_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, double *d_kkt,
                                         double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, void *stream);

/* ---- SHARED DYNAMICS (a separately labelled mode, not the headline workload): all instances have the same matrices and their own
 * vectors b, q, r -- fleets of identical systems, sweeps over initial states / references, and the way the reference's own test
 * programs call the solver (test_problems/test_d_ip_hard.c: one pBAbt aliased over all stages).  Factor once, then a batched
 * d_back_ric_rec_trs_tv_res (lqcp_solvers/d_back_ric_rec.c:564) whose matrices sit in shared memory: an instance costs
 * hpmpc_b200_shared_vec_stride() doubles in ([r q] of every stage in the ux layout, then b of every stage in the pi layout) and its
 * ux, pi out.  d_in_shared is ONE packed instance block (its vectors are ignored). ---- */
long long hpmpc_b200_shared_vec_stride(const hpmpc_b200_ocp *p);
long long hpmpc_b200_shared_factor_doubles(const hpmpc_b200_ocp *p);
int hpmpc_b200_d_back_ric_rec_trf_shared(hpmpc_b200_ocp *p, const double *d_in_shared, double *d_L_shared, void *stream);
int hpmpc_b200_d_back_ric_rec_trs_shared_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in_shared, const double *d_L_shared,
                                               const double *d_vec, double *d_ux, double *d_pi, void *stream);
/* stage-major vectors (uniform (12,5) / (8,3) patterns only, -2 otherwise): the part of a vector that belongs to stage n (offset
 * o_n in the layout above, K_n doubles) is one array [n_inst][K_n] at o_n * n_inst; b parts start at ux_stride * n_inst */
int hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major(hpmpc_b200_ocp *p, long long n_inst, const double *d_in_shared, const double *d_L_shared,
                                               const double *d_vec, double *d_ux, double *d_pi, void *stream);
int hpmpc_b200_d_back_ric_rec_sv_shared_batch_host(hpmpc_b200_ocp *p, long long n_inst, const double *h_in_shared, const double *h_vec,
                                                   double *h_ux, double *h_pi);

/* ---- partial condensing (reference lqcp_solvers/d_part_cond.c, used by {c,fortran}_order_d_ip_ocp_hard_tv when N2 < N,
 * interfaces/c/fortran_order_interface.c:389-528).  The horizon is cut into N2 blocks (the first N - N2*(N/N2) blocks hold N/N2+1 stages);
 * inside a block the states are eliminated, so the batch is solved as an N2-stage problem with stage inputs [u_{T-1} .. u_0] (newest
 * first) and the bounds on eliminated states as general constraints.  A pcond handle owns two size patterns: the full one (pack
 * inputs / read outputs with it) and the condensed one.  Bounds only before stage N (like the reference); a condensed stage must keep
 * nu+nx+1 <= 64.  Everything runs on the device: condense -> IPM -> expand on one stream. ---- */
typedef struct hpmpc_b200_pcond hpmpc_b200_pcond;
/* d_part_cond_compute_problem_size (d_part_cond.c:694) plus the bound positions of the condensed stages; nu has N+1 entries here */
int hpmpc_b200_part_cond_compute_problem_size(int N, const int *nx, const int *nu, const int *nb, int *const *hidxb, const int *ng, int N2,
                                              int *nx2, int *nu2, int *nb2, int *ng2, int **hidxb2 /* may be NULL */);
int  hpmpc_b200_pcond_create(hpmpc_b200_pcond **out, int N, const int *nx, const int *nu /* N entries */, const int *nb,
                             int *const *hidxb, const int *ng /* NULL, or zero before stage N */, int N2, int device);
void hpmpc_b200_pcond_destroy(hpmpc_b200_pcond *h);
hpmpc_b200_ocp *hpmpc_b200_pcond_full(hpmpc_b200_pcond *h);     /* owned by h */
hpmpc_b200_ocp *hpmpc_b200_pcond_cond(hpmpc_b200_pcond *h);     /* owned by h */
/* d_part_cond (d_part_cond.c:926): packed full batch -> packed condensed batch (the condensed handle's in_stride per instance) */
int hpmpc_b200_d_part_cond_batch(hpmpc_b200_pcond *h, long long n_inst, const double *d_in_full, double *d_in_cond, void *stream);
/* d_part_expand_solution (d_part_cond.c:1103): solution of the condensed batch -> solution of the full batch */
int hpmpc_b200_d_part_expand_solution_batch(hpmpc_b200_pcond *h, long long n_inst, const double *d_in_full, const double *d_ux2,
                                            const double *d_pi2, const double *d_lam2, const double *d_t2, double *d_ux, double *d_pi,
                                            double *d_lam, double *d_t, void *stream);
/* condense, IPM (cold start) on the condensed batch, expand; info[0..1] and the stat table are those of the condensed solve,
 * info[2..5] the exit norms of the FULL problem at the expanded solution (as the reference's wrapper computes them) */
int hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(hpmpc_b200_pcond *h, long long n_inst, const double *d_in_full, int k_max, double mu0,
                                                  double mu_tol, double alpha_min, double *d_ux, double *d_pi, double *d_lam,
                                                  double *d_t, double *d_info, void *stream);

/* ---- data in host memory: copies in, solves, copies out (chunked so copies overlap the kernels) ---- */
int hpmpc_b200_d_back_ric_rec_sv_batch_host(hpmpc_b200_ocp *p, long long n_inst, const double *h_in,
                                            double *h_ux, double *h_pi);
int hpmpc_b200_d_ip2_res_mpc_hard_batch_host(hpmpc_b200_ocp *p, long long n_inst, const double *h_in, int k_max, double mu0,
                                             double mu_tol, double alpha_min, int warm_start, double *h_ux, double *h_pi,
                                             double *h_lam, double *h_t, double *h_info);

/* MEASUREMENT TOOL (bench.py, tools/): the memory traffic of the sv kernel without its arithmetic -- same bulk copies, L2 hints and
 * output stores; d_ux / d_pi receive meaningless values.  Config-2 shape only (-2 otherwise). */
int hpmpc_b200_sv_traffic_probe(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, double *d_ux, double *d_pi, void *stream);
/* measured FP64 FMA throughput of the device in TFLOP/s (roofline denominator) */
double hpmpc_b200_fp64_peak_tflops(int device);
const char *hpmpc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
