/*
 * compat.c -- HPMPC's own C symbols, exported by libhpmpc_b200.so so that existing callers link unchanged.
 *
 * Each function below has the reference's exact signature (cited per function) and runs the problem as a
 * batch of ONE through the CUDA engine: unpack the caller's arrays into the native packed block, copy in,
 * launch, copy out.  This path exists for link compatibility and for the parity tests; throughput comes
 * from the batched entry points of hpmpc_b200.h.  General constraints (ng > 0: hpDCt / C, D, lg, ug) are supported on the any-size
 * kernels.  Nothing here computes the solution on the CPU, and a
 * missing / failing GPU is reported on stderr and turned into an error return (or abort() for the void
 * reference signatures, mirroring the reference's own printf+exit(1) on bad sizes, c_order_interface.c:103-110).
 *
 * Storage conventions accepted (reference include/block_size.h, auxiliary/d_aux_lib4.c:1310):
 *   low-level symbols : "lib4" panel-major matrices, bs = 4, column padding ncl = 2:
 *                       element (i,j) of a matrix with sda padded columns is p[(i/4)*4*sda + i%4 + 4*j];
 *                       bound-like vectors d, lam, t are [lower(pnb) upper(pnb)], pnb = nb rounded up to 4
 *   high-level symbols: dense row-major (c_order_) or column-major (fortran_order_) stage arrays
 * Deliberate differences (documented in INTEGRATION.md): the caller's matrices are never modified
 * (lib4 writes q / b / diagonal updates into them and restores them later, d_ip2_res_hard.c:721-732);
 * `work` is not used; `memory` holds the factor in the layout of layout.h; N2 < N selects partial condensing (pcond.c).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <cuda_runtime_api.h>
#include "layout.h"
#include "../../include/hpmpc_b200.h"
#include "../../include/hpmpc_compat.h"
#include "../../include/hpmpc_blasfeo_compat.h"

#define BS 4
#define NCL 2
#define PM(p, sda, i, j) ((p)[((i)/BS)*BS*(sda) + (i)%BS + BS*(j)])
#define RUP(x, m) (((x)+(m)-1)/(m)*(m))

/* ---- one cached single-instance context, rebuilt when the size pattern changes ---- */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static struct
	{
	hpmpc_b200_ocp *ocp;
	int N, *nx, *nu, *nb, *ng, **idxb;
	hpmpc_b200_sizes sz;
	int k_max_alloc;
	double *h_in, *h_ux, *h_pi, *h_Pb, *h_lam, *h_t, *h_info, *h_Qx, *h_qx;
	double *d_in, *d_ux, *d_pi, *d_Pb, *d_L, *d_lam, *d_t, *d_info, *d_Qx, *d_qx;
	/* KKT state of the last d_ip2_res_mpc_hard_tv call (what the reference keeps in the caller's work memory) and the
	 * work-memory pointer that call was given: d_kkt_solve_new_rhs_res_mpc_hard_tv must be handed the same one */
	double *d_kkt; const void *kkt_key;
	} G;

static void fatal(const char *what)
	{
	fprintf(stderr, "hpmpc_b200: %s -- no CPU fallback exists, aborting\n", what);
	abort();
	}

static int same_pattern(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *ng)
	{
	int n, j;
	if(!G.ocp || G.N!=N) return 0;
	for(n=0; n<=N; n++)
		{
		int nun = n<N ? nu[n] : 0, nbn = nb ? nb[n] : 0, ngn = ng ? ng[n] : 0;
		if(G.nx[n]!=nx[n] || G.nu[n]!=nun || G.nb[n]!=nbn || G.ng[n]!=ngn) return 0;
		for(j=0; j<nbn; j++) if(G.idxb[n][j]!=idxb[n][j]) return 0;
		}
	return 1;
	}

static void ctx_free(void)
	{
	int n;
	if(!G.ocp) return;
	hpmpc_b200_ocp_destroy(G.ocp);
	for(n=0; n<=G.N; n++) free(G.idxb[n]);
	free(G.idxb); free(G.nx); free(G.nu); free(G.nb); free(G.ng);
	free(G.h_in); free(G.h_ux); free(G.h_pi); free(G.h_Pb); free(G.h_lam); free(G.h_t); free(G.h_info); free(G.h_Qx); free(G.h_qx);
	cudaFree(G.d_Qx); cudaFree(G.d_qx);
	cudaFree(G.d_in); cudaFree(G.d_ux); cudaFree(G.d_pi); cudaFree(G.d_Pb); cudaFree(G.d_L); cudaFree(G.d_lam); cudaFree(G.d_t); cudaFree(G.d_info); cudaFree(G.d_kkt);
	memset(&G, 0, sizeof(G));
	}

static int ctx_get(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *ng, int k_max)
	{
	int n, j, dev = 0;
	if(!same_pattern(N, nx, nu, nb, idxb, ng))
		{
		ctx_free();
		const char *e = getenv("HPMPC_B200_DEVICE");
		if(e) dev = atoi(e); else if(cudaGetDevice(&dev)!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: no CUDA device available\n"); return -1; }
		if(hpmpc_b200_ocp_create_gen(&G.ocp, N, nx, nu, nb, idxb, ng, dev)) { G.ocp = NULL; return -1; }
		{ extern void hpmpc_b200_internal_generic_trf(hpmpc_b200_ocp *p); hpmpc_b200_internal_generic_trf(G.ocp); }
		G.N = N;
		G.nx = malloc((N+1)*sizeof(int)); G.nu = malloc((N+1)*sizeof(int)); G.nb = malloc((N+1)*sizeof(int)); G.ng = malloc((N+1)*sizeof(int));
		G.idxb = calloc(N+1, sizeof(int*));
		for(n=0; n<=N; n++)
			{
			G.nx[n] = nx[n]; G.nu[n] = n<N ? nu[n] : 0; G.nb[n] = nb ? nb[n] : 0; G.ng[n] = ng ? ng[n] : 0;
			G.idxb[n] = malloc((G.nb[n]+1)*sizeof(int));
			for(j=0; j<G.nb[n]; j++) G.idxb[n][j] = idxb[n][j];
			}
		hpmpc_b200_ocp_sizes(G.ocp, &G.sz);
		size_t lam = (size_t)(G.sz.lam_stride>0 ? G.sz.lam_stride : 2);
		G.h_in = calloc(G.sz.in_stride, sizeof(double)); G.h_ux = calloc(G.sz.ux_stride, sizeof(double));
		G.h_pi = calloc(G.sz.pi_stride+2, sizeof(double)); G.h_Pb = calloc(G.sz.pi_stride+2, sizeof(double));
		G.h_lam = calloc(lam, sizeof(double)); G.h_t = calloc(lam, sizeof(double));
		G.h_Qx = calloc(lam, sizeof(double)); G.h_qx = calloc(lam, sizeof(double));
		if(cudaMalloc((void**)&G.d_in, sizeof(double)*G.sz.in_stride)!=cudaSuccess
		|| cudaMalloc((void**)&G.d_ux, sizeof(double)*G.sz.ux_stride)!=cudaSuccess
		|| cudaMalloc((void**)&G.d_pi, sizeof(double)*(G.sz.pi_stride+2))!=cudaSuccess
		|| cudaMalloc((void**)&G.d_Pb, sizeof(double)*(G.sz.pi_stride+2))!=cudaSuccess
		|| cudaMalloc((void**)&G.d_L, sizeof(double)*G.sz.L_stride)!=cudaSuccess
		|| cudaMalloc((void**)&G.d_Qx, sizeof(double)*lam)!=cudaSuccess
		|| cudaMalloc((void**)&G.d_qx, sizeof(double)*lam)!=cudaSuccess
		|| cudaMalloc((void**)&G.d_lam, sizeof(double)*lam)!=cudaSuccess
		|| cudaMalloc((void**)&G.d_t, sizeof(double)*lam)!=cudaSuccess)
			{ fprintf(stderr, "hpmpc_b200: device allocation failed\n"); return -1; }
		}
	if(k_max>G.k_max_alloc)
		{
		free(G.h_info); if(G.d_info) cudaFree(G.d_info);
		G.h_info = calloc(HB_IPM_INFO_HEAD+5*k_max, sizeof(double));
		if(cudaMalloc((void**)&G.d_info, sizeof(double)*(HB_IPM_INFO_HEAD+5*k_max))!=cudaSuccess) return -1;
		G.k_max_alloc = k_max;
		}
	return 0;
	}

static int h2d(double *d, const double *h, size_t n) { return cudaMemcpy(d, h, sizeof(double)*n, cudaMemcpyHostToDevice)!=cudaSuccess; }
static int d2h(double *h, const double *d, size_t n) { return cudaMemcpy(h, d, sizeof(double)*n, cudaMemcpyDeviceToHost)!=cudaSuccess; }

/* panel-major problem data -> native block (G.h_in) */
/* hpDCt[n]: [D C]' of the stage, (nu+nx) x ng panel-major with cng = ng rounded up to ncl padded columns; hd[n] then continues
 * with lg (png) and ug (png) behind the two padded bound blocks (interfaces/c/c_order_interface.c:276-283, :371-378) */
/* panel-major stage matrices -> one packed instance block of handle `ocp` (ngv: the handle's ng[], may be NULL) */
static void pack_from_pmat_to(hpmpc_b200_ocp *ocp, const hpmpc_b200_sizes *sz, double *h_in, const int *ngv, int N, const int *nx, const int *nu,
		const int *nb, double **hpBAbt, double **hpRSQrq, double **hd, double **hpDCt)
	{
	int n, i, j;
	memset(h_in, 0, sizeof(double)*sz->in_stride);
	for(n=0; n<=N; n++)
		{
		int oB, oH, oD;
		hpmpc_b200_ocp_stage_offsets(ocp, n, &oB, &oH, &oD, NULL, NULL, NULL, NULL);
		int nun = n<N ? nu[n] : 0, nux = nun+nx[n];
		if(n<N)
			{
			int nx1 = nx[n+1], cnx1 = RUP(nx1, NCL);
			double *M = h_in + oB;
			for(i=0; i<=nux; i++) for(j=0; j<nx1; j++) M[i*nx1+j] = PM(hpBAbt[n], cnx1, i, j);
			}
		int cnux = RUP(nux, NCL);
		double *H = h_in + oH;
		for(i=0; i<nux; i++) for(j=0; j<=i; j++) H[HB_TRI(i)+j] = PM(hpRSQrq[n], cnux, i, j);
		for(j=0; j<nux; j++) H[HB_TRI(nux)+j] = PM(hpRSQrq[n], cnux, nux, j);
		if(hd && nb && nb[n]>0)
			{
			int pnb = RUP(nb[n], BS);
			for(j=0; j<nb[n]; j++) { h_in[oD+j] = hd[n][j]; h_in[oD+nb[n]+j] = hd[n][pnb+j]; }
			}
		if(ngv && ngv[n]>0 && hpDCt)
			{
			int ng = ngv[n], cng = RUP(ng, NCL), png = RUP(ng, BS), pnb = (nb && nb[n]>0) ? RUP(nb[n], BS) : 0, oG, oDg;
			hpmpc_b200_ocp_general_offsets(ocp, n, NULL, &oG, &oDg, NULL);
			for(i=0; i<nux; i++) for(j=0; j<ng; j++) h_in[oG+i*ng+j] = PM(hpDCt[n], cng, i, j);
			if(hd) for(j=0; j<ng; j++) { h_in[oDg+j] = hd[n][2*pnb+j]; h_in[oDg+ng+j] = hd[n][2*pnb+png+j]; }
			}
		}
	}
static void pack_from_pmat_g(int N, const int *nx, const int *nu, const int *nb, double **hpBAbt, double **hpRSQrq, double **hd, double **hpDCt)
	{
	pack_from_pmat_to(G.ocp, &G.sz, G.h_in, G.ng, N, nx, nu, nb, hpBAbt, hpRSQrq, hd, hpDCt);
	}

static int stage_updates(int N, const int *nb, double **Qx, double **qx, const double **dQx, const double **dqx)
	{
	int n, j, any = 0;
	*dQx = NULL; *dqx = NULL;
	for(n=0; n<=N; n++) if((nb && nb[n]>0) || G.ng[n]>0) any = 1;
	if(!any || (!Qx && !qx)) return 0;
	for(n=0; n<=N; n++)
		{
		int oc, nbn = nb ? nb[n] : 0, pnb = nbn>0 ? RUP(nbn, BS) : 0;
		hpmpc_b200_ocp_general_offsets(G.ocp, n, NULL, NULL, NULL, &oc);
		for(j=0; j<nbn; j++) { if(Qx) G.h_Qx[oc+j] = Qx[n][j]; if(qx) G.h_qx[oc+j] = qx[n][j]; }
		for(j=0; j<G.ng[n]; j++) { if(Qx) G.h_Qx[oc+nbn+j] = Qx[n][pnb+j]; if(qx) G.h_qx[oc+nbn+j] = qx[n][pnb+j]; }
		}
	if(Qx) { if(cudaMemcpy(G.d_Qx, G.h_Qx, sizeof(double)*G.sz.nbtot, cudaMemcpyHostToDevice)!=cudaSuccess) return -1; *dQx = G.d_Qx; }
	if(qx) { if(cudaMemcpy(G.d_qx, G.h_qx, sizeof(double)*G.sz.nbtot, cudaMemcpyHostToDevice)!=cudaSuccess) return -1; *dqx = G.d_qx; }
	return 0;
	}

/* optional vector overrides of the Riccati entry points, folded into the packed block */
static void fold_updates(int N, const int *nx, const int *nu, const int *nb, int **idxb, int update_b, double **b,
		int update_q, double **q, double **bd, double **Qx, double **qx)
	{
	int n, j;
	for(n=0; n<=N; n++)
		{
		int oB, oH;
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, &oB, &oH, NULL, NULL, NULL, NULL, NULL);
		int nun = n<N ? nu[n] : 0, nux = nun+nx[n];
		double *H = G.h_in + oH;
		if(update_b && n<N) for(j=0; j<nx[n+1]; j++) G.h_in[oB+nux*nx[n+1]+j] = b[n][j];
		if(update_q) for(j=0; j<nux; j++) H[HB_TRI(nux)+j] = q[n][j];
		/* ddiaadin_libsp (d_back_ric_rec.c:199): the bounded diagonal entries are bd + Qx, i.e. bd replaces whatever the
		 * matrix holds; Qx itself (and qx) are applied on the device */
		if(nb && nb[n]>0 && Qx && bd)
			for(j=0; j<nb[n]; j++)
				{
				int id = idxb[n][j];
				H[HB_TRI(id)+id] = bd[n][j];
				}
		(void)qx;
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* reference include/lqcp_solvers.h:37-45                                                             */
/* ------------------------------------------------------------------------------------------------ */
int d_back_ric_rec_sv_tv_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)ng;
	return 64;     /* the engine keeps its scratch on the device */
	}

int d_back_ric_rec_sv_tv_memory_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng)
	{
	(void)nb; (void)ng;
	long long d = 0; int n;
	for(n=0; n<=N; n++) { int nux = (n<N ? nu[n] : 0) + nx[n]; d += HB_EVEN(HB_TRI(nux)+2*nux); }
	return (int)((d*sizeof(double)+63)/64*64);
	}

void d_back_ric_rec_sv_tv_res(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, int update_b, double **hpBAbt, double **b,
		int update_q, double **hpRSQrq, double **q, double **bd, double **hpDCt, double **Qx, double **qx, double **hux,
		int compute_pi, double **hpi, int compute_Pb, double **hPb, double *memory, double *work)
	{
	(void)work;
	int n, i;
	const double *dQx, *dqx;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu, nb, idxb, ng, 0)) fatal("d_back_ric_rec_sv_tv_res: GPU context unavailable");
	pack_from_pmat_g(N, nx, nu, nb, hpBAbt, hpRSQrq, NULL, hpDCt);
	fold_updates(N, nx, nu, nb, idxb, update_b, b, update_q, q, bd, Qx, qx);
	if(h2d(G.d_in, G.h_in, G.sz.in_stride) || stage_updates(N, nb, Qx, qx, &dQx, &dqx)) fatal("copy to device failed");
	/* factor kept per instance (d_L) so that a later trs call can reuse it through `memory` */
	if(hpmpc_b200_d_back_ric_rec_sv_upd_batch(G.ocp, 1, G.d_in, dQx, dqx, G.d_ux, G.d_pi, G.d_Pb, NULL)) fatal("sv launch failed");
	if(cudaDeviceSynchronize()!=cudaSuccess) fatal("sv kernel failed");
	if(d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride) || d2h(G.h_Pb, G.d_Pb, G.sz.pi_stride)) fatal("copy from device failed");
	if(memory)
		{
		/* slot 0 of the engine's stash holds this instance's factor */
		extern int hpmpc_b200_internal_copy_stash(hpmpc_b200_ocp *p, double *h_dst);
		if(hpmpc_b200_internal_copy_stash(G.ocp, memory)) fatal("copy of the factor failed");
		}
	for(n=0; n<=N; n++)
		{
		int oU, oP, nun = n<N ? nu[n] : 0;
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, NULL, NULL);
		for(i=0; i<nun+nx[n]; i++) hux[n][i] = G.h_ux[oU+i];
		if(n<N && compute_pi) for(i=0; i<nx[n+1]; i++) hpi[n][i] = G.h_pi[oP+i];
		if(n<N && compute_Pb) for(i=0; i<nx[n+1]; i++) hPb[n][i] = G.h_Pb[oP+i];
		}
	pthread_mutex_unlock(&g_lock);
	}

void d_back_ric_rec_trf_tv_res(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hpRSQrq,
		double **hpDCt, double **Qx, double **bd, double *memory, double *work)
	{
	(void)work;
	const double *dQx, *dqx;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu, nb, idxb, ng, 0)) fatal("d_back_ric_rec_trf_tv_res: GPU context unavailable");
	pack_from_pmat_g(N, nx, nu, nb, hpBAbt, hpRSQrq, NULL, hpDCt);
	fold_updates(N, nx, nu, nb, idxb, 0, NULL, 0, NULL, bd, Qx, NULL);
	if(h2d(G.d_in, G.h_in, G.sz.in_stride) || stage_updates(N, nb, Qx, NULL, &dQx, &dqx)) fatal("copy to device failed");
	if(hpmpc_b200_d_back_ric_rec_trf_upd_batch(G.ocp, 1, G.d_in, dQx, G.d_L, NULL)) fatal("trf launch failed");
	if(cudaDeviceSynchronize()!=cudaSuccess) fatal("trf kernel failed");
	if(d2h(memory, G.d_L, G.sz.L_stride)) fatal("copy from device failed");
	pthread_mutex_unlock(&g_lock);
	}

void d_back_ric_rec_trs_tv_res(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hq,
		double **hpDCt, double **qx, double **hux, int compute_pi, double **hpi, int compute_Pb, double **hPb, double *memory, double *work)
	{
	(void)work; (void)compute_Pb; (void)hPb;
	int n, i, j;
	const double *dQx, *dqx;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu, nb, idxb, ng, 0)) fatal("d_back_ric_rec_trs_tv_res: GPU context unavailable");
	/* only [B A]', b and the gradient are read by the solve; the Hessian part of the block is unused */
	memset(G.h_in, 0, sizeof(double)*G.sz.in_stride);
	for(n=0; n<=N; n++)
		{
		int oB, oH, nun = n<N ? nu[n] : 0, nux = nun+nx[n];
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, &oB, &oH, NULL, NULL, NULL, NULL, NULL);
		if(n<N)
			{
			int nx1 = nx[n+1], cnx1 = RUP(nx1, NCL);
			for(i=0; i<nux; i++) for(j=0; j<nx1; j++) G.h_in[oB+i*nx1+j] = PM(hpBAbt[n], cnx1, i, j);
			for(j=0; j<nx1; j++) G.h_in[oB+nux*nx1+j] = hb[n][j];
			}
		for(j=0; j<nux; j++) G.h_in[oH+HB_TRI(nux)+j] = hq[n][j];
		if(G.ng[n]>0 && hpDCt)
			{
			int ngn = G.ng[n], cng = RUP(ngn, NCL), oG;
			hpmpc_b200_ocp_general_offsets(G.ocp, n, NULL, &oG, NULL, NULL);
			for(i=0; i<nux; i++) for(j=0; j<ngn; j++) G.h_in[oG+i*ngn+j] = PM(hpDCt[n], cng, i, j);
			}
		}
	if(h2d(G.d_in, G.h_in, G.sz.in_stride) || h2d(G.d_L, memory, G.sz.L_stride) || stage_updates(N, nb, NULL, qx, &dQx, &dqx)) fatal("copy to device failed");
	if(hpmpc_b200_d_back_ric_rec_trs_upd_batch(G.ocp, 1, G.d_in, G.d_L, dqx, G.d_ux, G.d_pi, NULL)) fatal("trs launch failed");
	if(cudaDeviceSynchronize()!=cudaSuccess) fatal("trs kernel failed");
	if(d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride)) fatal("copy from device failed");
	for(n=0; n<=N; n++)
		{
		int oU, oP, nun = n<N ? nu[n] : 0;
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, NULL, NULL);
		for(i=0; i<nun+nx[n]; i++) hux[n][i] = G.h_ux[oU+i];
		if(n<N && compute_pi) for(i=0; i<nx[n+1]; i++) hpi[n][i] = G.h_pi[oP+i];
		}
	pthread_mutex_unlock(&g_lock);
	}

/* ------------------------------------------------------------------------------------------------ */
/* reference include/mpc_solvers.h:41-42                                                              */
/* ------------------------------------------------------------------------------------------------ */
int d_ip2_res_mpc_hard_tv_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)ng;
	return 64;
	}

static int run_ipm_single(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat, const void *keep_kkt)
	{
	int i;
	if(h2d(G.d_in, G.h_in, G.sz.in_stride)) return -1;
	if(warm_start && h2d(G.d_ux, G.h_ux, G.sz.ux_stride)) return -1;
	if(cudaMemset(G.d_info, 0, sizeof(double)*(HB_IPM_INFO_HEAD+5*k_max))!=cudaSuccess) return -1;
	G.kkt_key = NULL;
	if(keep_kkt && G.sz.nbtot>0)
		{
		if(!G.d_kkt && cudaMalloc((void**)&G.d_kkt, sizeof(double)*(size_t)hpmpc_b200_kkt_state_stride(G.ocp))!=cudaSuccess) return -1;
		if(hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch(G.ocp, 1, G.d_in, k_max, mu0, mu_tol, alpha_min, warm_start,
				G.d_ux, G.d_pi, G.d_lam, G.d_t, G.d_info, G.d_kkt, NULL)) return -1;
		G.kkt_key = keep_kkt;
		}
	else if(hpmpc_b200_d_ip2_res_mpc_hard_batch(G.ocp, 1, G.d_in, k_max, mu0, mu_tol, alpha_min, warm_start,
			G.d_ux, G.d_pi, G.d_lam, G.d_t, G.d_info, NULL)) return -1;
	if(cudaDeviceSynchronize()!=cudaSuccess) return -1;
	if(d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride) || d2h(G.h_info, G.d_info, HB_IPM_INFO_HEAD+5*k_max)) return -1;
	if(G.sz.lam_stride>0 && (d2h(G.h_lam, G.d_lam, G.sz.lam_stride) || d2h(G.h_t, G.d_t, G.sz.lam_stride))) return -1;
	*kk = (int)G.h_info[0];
	if(stat) for(i=0; i<5*(*kk); i++) stat[i] = G.h_info[HB_IPM_INFO_HEAD+i];
	return 0;
	}

int d_ip2_res_mpc_hard_tv(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat, int N,
		int *nx, int *nu_N, int *nb, int **idxb, int *ng, double **pBAbt, double **pQ, double **pDCt, double **d, double **ux,
		int compute_mult, double **pi, double **lam, double **t, double *double_work_memory)
	{
	(void)compute_mult;
	int n, i, status;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu_N, nb, idxb, ng, k_max)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: d_ip2_res_mpc_hard_tv: GPU context unavailable\n"); return -1; }
	pack_from_pmat_g(N, nx, nu_N, nb, pBAbt, pQ, d, pDCt);
	if(warm_start)
		for(n=0; n<=N; n++)
			{
			int oU, nun = n<N ? nu_N[n] : 0;
			hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, NULL, NULL, NULL);
			for(i=0; i<nun+nx[n]; i++) G.h_ux[oU+i] = ux[n][i];
			}
	if(run_ipm_single(kk, k_max, mu0, mu_tol, alpha_min, warm_start, stat, double_work_memory)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: d_ip2_res_mpc_hard_tv: GPU execution failed\n"); return -1; }
	status = (int)G.h_info[1];
	for(n=0; n<=N; n++)
		{
		int oU, oP, oLm, nun = n<N ? nu_N[n] : 0, nbn = nb[n], pnb = RUP(nbn, BS);
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, &oLm, NULL);
		for(i=0; i<nun+nx[n]; i++) ux[n][i] = G.h_ux[oU+i];
		if(n<N) for(i=0; i<nx[n+1]; i++) pi[n][i] = G.h_pi[oP+i];
		for(i=0; i<nbn; i++)
			{
			lam[n][i] = G.h_lam[oLm+i]; lam[n][pnb+i] = G.h_lam[oLm+nbn+i];
			t[n][i] = G.h_t[oLm+i]; t[n][pnb+i] = G.h_t[oLm+nbn+i];
			}
		/* general constraints behind the two padded bound blocks: [lg (png) ug (png)] */
		for(i=0; i<G.ng[n]; i++)
			{
			const int ngn = G.ng[n], png = RUP(ngn, BS);
			lam[n][2*pnb+i] = G.h_lam[oLm+2*nbn+i]; lam[n][2*pnb+png+i] = G.h_lam[oLm+2*nbn+ngn+i];
			t[n][2*pnb+i] = G.h_t[oLm+2*nbn+i]; t[n][2*pnb+png+i] = G.h_t[oLm+2*nbn+ngn+i];
			}
		}
	pthread_mutex_unlock(&g_lock);
	return status;
	}

/* reference include/mpc_solvers.h:46 (mpc_solvers/d_ip2_res_hard.c:1922): the last KKT system of the preceding
 * d_ip2_res_mpc_hard_tv call solved again for new b, q, d.  The reference finds the factor, t_inv and the backed-up iterate in
 * double_work_memory; here they stay on the device, keyed by that pointer: the call must follow a d_ip2_res_mpc_hard_tv call
 * with the same sizes and the same double_work_memory (anything else is reported and aborts, there is nothing to solve with). */
void d_kkt_solve_new_rhs_res_mpc_hard_tv(int N, int *nx, int *nu_N, int *nb, int **idxb, int *ng, double **pBAbt, double **b,
		double **pQ, double **q, double **pDCt, double **d, double **ux, int compute_mult, double **pi, double **lam, double **t,
		double *double_work_memory)
	{
	(void)compute_mult;
	int n, i;
	pthread_mutex_lock(&g_lock);
	/* an unconstrained problem has no IPM state: the reference's re-solve degenerates to one Riccati solve with the new b, q
	 * (d_ip2_res_hard.c:2225 with empty constraint sets) */
	{
	int nctot = 0;
	for(n=0; n<=N; n++) nctot += (nb ? nb[n] : 0) + (ng ? ng[n] : 0);
	if(nctot==0)
		{
		if(ctx_get(N, nx, nu_N, nb, idxb, ng, 1)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: d_kkt_solve_new_rhs_res_mpc_hard_tv: GPU context unavailable, outputs untouched\n"); return; }
		pack_from_pmat_g(N, nx, nu_N, nb, pBAbt, pQ, NULL, NULL);
		for(n=0; n<=N; n++)
			{
			int oB, oH, nun = n<N ? nu_N[n] : 0, nux = nun+nx[n];
			hpmpc_b200_ocp_stage_offsets(G.ocp, n, &oB, &oH, NULL, NULL, NULL, NULL, NULL);
			if(n<N) for(i=0; i<nx[n+1]; i++) G.h_in[oB + (size_t)nux*nx[n+1] + i] = b[n][i];
			for(i=0; i<nux; i++) G.h_in[oH + HB_TRI(nux) + i] = q[n][i];
			}
		if(h2d(G.d_in, G.h_in, G.sz.in_stride) || hpmpc_b200_d_back_ric_rec_sv_batch(G.ocp, 1, G.d_in, G.d_ux, G.d_pi, NULL, NULL)
		|| cudaDeviceSynchronize()!=cudaSuccess || d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride))
			{ pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: d_kkt_solve_new_rhs_res_mpc_hard_tv: GPU execution failed, outputs untouched\n"); return; }
		for(n=0; n<=N; n++)
			{
			int oU, oP, nun = n<N ? nu_N[n] : 0;
			hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, NULL, NULL);
			for(i=0; i<nun+nx[n]; i++) ux[n][i] = G.h_ux[oU+i];
			if(n<N) for(i=0; i<nx[n+1]; i++) pi[n][i] = G.h_pi[oP+i];
			}
		pthread_mutex_unlock(&g_lock);
		return;
		}
	}
	/* nothing to solve with: say why and leave the caller's arrays untouched (the reference would read whatever its work memory
	 * holds); a drop-in library does not kill the host process for this */
	if(!same_pattern(N, nx, nu_N, nb, idxb, ng) || G.kkt_key==NULL || G.kkt_key!=(const void*)double_work_memory)
		{
		pthread_mutex_unlock(&g_lock);
		fprintf(stderr, "hpmpc_b200: d_kkt_solve_new_rhs_res_mpc_hard_tv: no KKT state for this work memory -- call d_ip2_res_mpc_hard_tv with the same sizes and double_work_memory first; outputs untouched\n");
		return;
		}
	pack_from_pmat_g(N, nx, nu_N, nb, pBAbt, pQ, d, pDCt);
	for(n=0; n<=N; n++)
		{
		int oB, oH, nun = n<N ? nu_N[n] : 0, nux = nun+nx[n];
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, &oB, &oH, NULL, NULL, NULL, NULL, NULL);
		if(n<N) for(i=0; i<nx[n+1]; i++) G.h_in[oB + (size_t)nux*nx[n+1] + i] = b[n][i];
		for(i=0; i<nux; i++) G.h_in[oH + HB_TRI(nux) + i] = q[n][i];
		}
	if(h2d(G.d_in, G.h_in, G.sz.in_stride)
	|| hpmpc_b200_d_kkt_solve_new_rhs_batch(G.ocp, 1, G.d_in, G.d_kkt, G.d_ux, G.d_pi, G.d_lam, G.d_t, G.d_info, NULL)
	|| cudaDeviceSynchronize()!=cudaSuccess
	|| d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride) || d2h(G.h_info, G.d_info, HB_IPM_INFO_HEAD)
	|| d2h(G.h_lam, G.d_lam, G.sz.lam_stride) || d2h(G.h_t, G.d_t, G.sz.lam_stride))
		{ pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: d_kkt_solve_new_rhs_res_mpc_hard_tv: GPU execution failed, outputs untouched\n"); return; }
	if(G.h_info[1]!=0.0)
		{
		pthread_mutex_unlock(&g_lock);
		fprintf(stderr, "hpmpc_b200: d_kkt_solve_new_rhs_res_mpc_hard_tv: the preceding IPM call ran no phase-2 iteration (it converged in phase 1), there is no factor to reuse; outputs untouched\n");
		return;
		}
	for(n=0; n<=N; n++)
		{
		int oU, oP, oLm, nun = n<N ? nu_N[n] : 0, nbn = nb[n], pnb = RUP(nbn, BS);
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, &oLm, NULL);
		for(i=0; i<nun+nx[n]; i++) ux[n][i] = G.h_ux[oU+i];
		if(n<N) for(i=0; i<nx[n+1]; i++) pi[n][i] = G.h_pi[oP+i];
		for(i=0; i<nbn; i++)
			{
			lam[n][i] = G.h_lam[oLm+i]; lam[n][pnb+i] = G.h_lam[oLm+nbn+i];
			t[n][i] = G.h_t[oLm+i]; t[n][pnb+i] = G.h_t[oLm+nbn+i];
			}
		/* general constraints behind the two padded bound blocks: [lg (png) ug (png)] */
		for(i=0; i<G.ng[n]; i++)
			{
			const int ngn = G.ng[n], png = RUP(ngn, BS);
			lam[n][2*pnb+i] = G.h_lam[oLm+2*nbn+i]; lam[n][2*pnb+png+i] = G.h_lam[oLm+2*nbn+ngn+i];
			t[n][2*pnb+i] = G.h_t[oLm+2*nbn+i]; t[n][2*pnb+png+i] = G.h_t[oLm+2*nbn+ngn+i];
			}
		}
	pthread_mutex_unlock(&g_lock);
	}

/* reference include/mpc_solvers.h:45 (mpc_solvers/d_ip2_res_hard.c:1348): k_max residual-based Newton steps from the iterate
 * (ux0, pi0, lam0, t0) -- ux0[n] = [u_n ; x_n], lam0[n] / t0[n] = [lb(nb) ub(nb)] UNPADDED (c99/d_aux_ip_hard_lib4.c:178-186) --
 * with the centering term fixed at mu0.  Returns 1 after k_max steps like the reference (:1911). */
int d_ip2_res_mpc_hard_tv_single_newton_step(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat,
		int N, int *nx, int *nu_N, int *nb, int **idxb, int *ng, double **pBAbt, double **pQ, double **pDCt, double **d, double **ux,
		int compute_mult, double **pi, double **lam, double **t, double *double_work_memory, double **ux0, double **pi0, double **lam0,
		double **t0)
	{
	(void)mu_tol; (void)warm_start; (void)compute_mult; (void)double_work_memory;
	int n, i, status;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu_N, nb, idxb, ng, k_max)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: single_newton_step: GPU context unavailable\n"); return -1; }
	pack_from_pmat_g(N, nx, nu_N, nb, pBAbt, pQ, d, pDCt);
	if(G.sz.nbtot==0)
		{
		/* no constraints: one Riccati solve, kk = 0, return 0 (:1586-1604) */
		int rc = run_ipm_single(kk, k_max, mu0, 0.0, alpha_min, 0, stat, NULL);
		if(rc) { pthread_mutex_unlock(&g_lock); return -1; }
		}
	else
		{
		for(n=0; n<=N; n++)
			{
			int oU, oP, oLm, nun = n<N ? nu_N[n] : 0, nbn = nb[n];
			hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, &oLm, NULL);
			for(i=0; i<nun+nx[n]; i++) G.h_ux[oU+i] = ux0[n][i];
			if(n<N) for(i=0; i<nx[n+1]; i++) G.h_pi[oP+i] = pi0[n][i];
			for(i=0; i<2*nbn; i++) { G.h_lam[oLm+i] = lam0[n][i]; G.h_t[oLm+i] = t0[n][i]; }
			}
		if(h2d(G.d_in, G.h_in, G.sz.in_stride) || h2d(G.d_ux, G.h_ux, G.sz.ux_stride) || h2d(G.d_pi, G.h_pi, G.sz.pi_stride)
		|| h2d(G.d_lam, G.h_lam, G.sz.lam_stride) || h2d(G.d_t, G.h_t, G.sz.lam_stride)
		|| cudaMemset(G.d_info, 0, sizeof(double)*(HB_IPM_INFO_HEAD+5*k_max))!=cudaSuccess
		|| hpmpc_b200_d_ip2_res_mpc_hard_single_newton_step_batch(G.ocp, 1, G.d_in, k_max, mu0, alpha_min, G.d_ux, G.d_pi, G.d_lam, G.d_t, G.d_info, NULL)
		|| cudaDeviceSynchronize()!=cudaSuccess
		|| d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride) || d2h(G.h_info, G.d_info, HB_IPM_INFO_HEAD+5*k_max)
		|| d2h(G.h_lam, G.d_lam, G.sz.lam_stride) || d2h(G.h_t, G.d_t, G.sz.lam_stride))
			{ pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: single_newton_step: GPU execution failed\n"); return -1; }
		*kk = (int)G.h_info[0];
		if(stat) for(i=0; i<5*(*kk); i++) stat[i] = G.h_info[HB_IPM_INFO_HEAD+i];
		}
	status = (int)G.h_info[1];
	for(n=0; n<=N; n++)
		{
		int oU, oP, oLm, nun = n<N ? nu_N[n] : 0, nbn = nb[n], pnb = RUP(nbn, BS);
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, &oLm, NULL);
		for(i=0; i<nun+nx[n]; i++) ux[n][i] = G.h_ux[oU+i];
		if(n<N) for(i=0; i<nx[n+1]; i++) pi[n][i] = G.h_pi[oP+i];
		for(i=0; i<nbn; i++)
			{
			lam[n][i] = G.h_lam[oLm+i]; lam[n][pnb+i] = G.h_lam[oLm+nbn+i];
			t[n][i] = G.h_t[oLm+i]; t[n][pnb+i] = G.h_t[oLm+nbn+i];
			}
		}
	pthread_mutex_unlock(&g_lock);
	return status;
	}

/* reference include/mpc_solvers.h:47 (mpc_solvers/c99/d_res_ip_res_hard.c:39) and :36 (mpc_solvers/d_res_ip_hard.c:38): the KKT
 * residuals of a given point.  hb / hq are the vectors b_n, [r;q]_n (the reference reads them instead of the last rows of the
 * matrices); hd, hlam, ht, hrd, hrm in the lib4 padded layout [lb(pnb) ub(pnb) lg(png) ug(png)].  `flip_upper`: d_res_mpc_hard_tv
 * returns the upper-bound residual with the opposite sign (d_res_ip_hard.c:80,289-295) and no res_m. */
static void residuals_single(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hpQ, double **hq,
		double **hux, double **hpDCt, double **hd, double **hpi, double **hlam, double **ht, double **hrq, double **hrb, double **hrd,
		double **hrm, double *mu, int flip_upper, const char *who)
	{
	int n, i;
	double *d_rq = NULL, *d_rb = NULL, *d_rd = NULL, *d_rm = NULL, *d_mu = NULL;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu, nb, idxb, ng, 1)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: %s: GPU context unavailable, outputs untouched\n", who); return; }
	pack_from_pmat_g(N, nx, nu, nb, hpBAbt, hpQ, hd, hpDCt);
	const size_t lam = (size_t)(G.sz.lam_stride>0 ? G.sz.lam_stride : 2);
	double *h_rd = calloc(lam, sizeof(double)), *h_rm = calloc(lam, sizeof(double)), h_mu = 0.0;
	for(n=0; n<=N; n++)
		{
		int oB, oH, oU, oP, oLm, nun = n<N ? nu[n] : 0, nux = nun+nx[n], nbn = nb ? nb[n] : 0, pnb = nbn>0 ? RUP(nbn, BS) : 0, ngn = G.ng[n], png = RUP(ngn, BS);
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, &oB, &oH, NULL, &oU, &oP, &oLm, NULL);
		if(n<N && hb) for(i=0; i<nx[n+1]; i++) G.h_in[oB + (size_t)nux*nx[n+1] + i] = hb[n][i];
		if(hq) for(i=0; i<nux; i++) G.h_in[oH + HB_TRI(nux) + i] = hq[n][i];
		for(i=0; i<nux; i++) G.h_ux[oU+i] = hux[n][i];
		if(n<N) for(i=0; i<nx[n+1]; i++) G.h_pi[oP+i] = hpi[n][i];
		for(i=0; i<nbn; i++) { G.h_lam[oLm+i] = hlam[n][i]; G.h_lam[oLm+nbn+i] = hlam[n][pnb+i]; G.h_t[oLm+i] = ht[n][i]; G.h_t[oLm+nbn+i] = ht[n][pnb+i]; }
		for(i=0; i<ngn; i++)
			{
			G.h_lam[oLm+2*nbn+i] = hlam[n][2*pnb+i]; G.h_lam[oLm+2*nbn+ngn+i] = hlam[n][2*pnb+png+i];
			G.h_t[oLm+2*nbn+i] = ht[n][2*pnb+i]; G.h_t[oLm+2*nbn+ngn+i] = ht[n][2*pnb+png+i];
			}
		}
	int bad = cudaMalloc((void**)&d_rq, sizeof(double)*G.sz.ux_stride)!=cudaSuccess || cudaMalloc((void**)&d_rb, sizeof(double)*(G.sz.pi_stride+2))!=cudaSuccess
		|| cudaMalloc((void**)&d_rd, sizeof(double)*lam)!=cudaSuccess || cudaMalloc((void**)&d_rm, sizeof(double)*lam)!=cudaSuccess
		|| cudaMalloc((void**)&d_mu, sizeof(double))!=cudaSuccess
		|| h2d(G.d_in, G.h_in, G.sz.in_stride) || h2d(G.d_ux, G.h_ux, G.sz.ux_stride) || h2d(G.d_pi, G.h_pi, G.sz.pi_stride)
		|| (G.sz.lam_stride>0 && (h2d(G.d_lam, G.h_lam, G.sz.lam_stride) || h2d(G.d_t, G.h_t, G.sz.lam_stride)))
		|| hpmpc_b200_d_res_res_mpc_hard_batch(G.ocp, 1, G.d_in, G.d_ux, G.d_pi, G.d_lam, G.d_t, d_rq, d_rb, d_rd, d_rm, d_mu, NULL)
		|| cudaDeviceSynchronize()!=cudaSuccess
		|| d2h(G.h_ux, d_rq, G.sz.ux_stride) || d2h(G.h_pi, d_rb, G.sz.pi_stride)
		|| (G.sz.lam_stride>0 && (d2h(h_rd, d_rd, G.sz.lam_stride) || d2h(h_rm, d_rm, G.sz.lam_stride))) || d2h(&h_mu, d_mu, 1);
	cudaFree(d_rq); cudaFree(d_rb); cudaFree(d_rd); cudaFree(d_rm); cudaFree(d_mu);
	if(bad) { free(h_rd); free(h_rm); pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: %s: GPU execution failed, outputs untouched\n", who); return; }
	for(n=0; n<=N; n++)
		{
		int oU, oP, oLm, nun = n<N ? nu[n] : 0, nbn = nb ? nb[n] : 0, pnb = nbn>0 ? RUP(nbn, BS) : 0, ngn = G.ng[n], png = RUP(ngn, BS);
		const double su = flip_upper ? -1.0 : 1.0;
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, &oLm, NULL);
		for(i=0; i<nun+nx[n]; i++) hrq[n][i] = G.h_ux[oU+i];
		if(n<N) for(i=0; i<nx[n+1]; i++) hrb[n][i] = G.h_pi[oP+i];
		for(i=0; i<nbn; i++)
			{
			hrd[n][i] = h_rd[oLm+i]; hrd[n][pnb+i] = su*h_rd[oLm+nbn+i];
			if(hrm) { hrm[n][i] = h_rm[oLm+i]; hrm[n][pnb+i] = h_rm[oLm+nbn+i]; }
			}
		for(i=0; i<ngn; i++)
			{
			hrd[n][2*pnb+i] = h_rd[oLm+2*nbn+i]; hrd[n][2*pnb+png+i] = su*h_rd[oLm+2*nbn+ngn+i];
			if(hrm) { hrm[n][2*pnb+i] = h_rm[oLm+2*nbn+i]; hrm[n][2*pnb+png+i] = h_rm[oLm+2*nbn+ngn+i]; }
			}
		}
	*mu = h_mu;
	free(h_rd); free(h_rm);
	pthread_mutex_unlock(&g_lock);
	}

void d_res_res_mpc_hard_tv(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hpQ, double **hq,
		double **hux, double **hpDCt, double **hd, double **hpi, double **hlam, double **ht, double *work, double **hrq, double **hrb,
		double **hrd, double **hrm, double *mu)
	{
	(void)work;
	residuals_single(N, nx, nu, nb, idxb, ng, hpBAbt, hb, hpQ, hq, hux, hpDCt, hd, hpi, hlam, ht, hrq, hrb, hrd, hrm, mu, 0, "d_res_res_mpc_hard_tv");
	}

void d_res_mpc_hard_tv(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hpQ, double **hq,
		double **hux, double **hpDCt, double **hd, double **hpi, double **hlam, double **ht, double **hrq, double **hrb, double **hrd,
		double *mu)
	{
	residuals_single(N, nx, nu, nb, idxb, ng, hpBAbt, hb, hpQ, hq, hux, hpDCt, hd, hpi, hlam, ht, hrq, hrb, hrd, NULL, mu, 1, "d_res_mpc_hard_tv");
	}

/* ------------------------------------------------------------------------------------------------ */
/* reference include/c_interface.h:59-67                                                              */
/* ------------------------------------------------------------------------------------------------ */
int hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)hidxb; (void)ng; (void)N2;
	return 64;     /* callers may keep passing their (larger) buffer; it is not touched */
	}

/* reference include/lqcp_solvers.h:86 (lqcp_solvers/d_part_cond.c:694): sizes of the partially condensed problem; host arithmetic only */
void d_part_cond_compute_problem_size(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int *nx2, int *nu2, int *nb2, int *ng2)
	{
	if(hpmpc_b200_part_cond_compute_problem_size(N, nx, nu, nb, hidxb, ng, N2, nx2, nu2, nb2, ng2, NULL))
		fatal("d_part_cond_compute_problem_size: general constraints before stage N or N2 outside 1..N (the reference exits here too, d_part_cond.c:962-968)");
	}

/* ---- partial condensing (N2 < N): a second cached context, a pcond handle (pcond.c) keyed by the size pattern and N2 ---- */
static struct
	{
	hpmpc_b200_pcond *h;
	int N, N2, *nx, *nu, *nb, *ng, **idxb;
	hpmpc_b200_sizes sz;
	int k_max_alloc;
	double *h_in, *h_ux, *h_pi, *h_lam, *h_info;
	double *d_in, *d_ux, *d_pi, *d_lam, *d_t, *d_info;
	} GP;

static void pc_ctx_free(void)
	{
	int n;
	if(!GP.h) return;
	hpmpc_b200_pcond_destroy(GP.h);
	for(n=0; n<=GP.N; n++) free(GP.idxb[n]);
	free(GP.idxb); free(GP.nx); free(GP.nu); free(GP.nb); free(GP.ng);
	free(GP.h_in); free(GP.h_ux); free(GP.h_pi); free(GP.h_lam); free(GP.h_info);
	cudaFree(GP.d_in); cudaFree(GP.d_ux); cudaFree(GP.d_pi); cudaFree(GP.d_lam); cudaFree(GP.d_t); cudaFree(GP.d_info);
	memset(&GP, 0, sizeof(GP));
	}

static int pc_ctx_get(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *ng, int N2, int k_max)
	{
	int n, j, same = GP.h!=NULL && GP.N==N && GP.N2==N2, dev = 0;
	for(n=0; same && n<=N; n++)
		{
		const int nun = n<N ? nu[n] : 0, nbn = nb ? nb[n] : 0, ngn = ng ? ng[n] : 0;
		if(GP.nx[n]!=nx[n] || GP.nu[n]!=nun || GP.nb[n]!=nbn || GP.ng[n]!=ngn) same = 0;
		for(j=0; same && j<nbn; j++) if(GP.idxb[n][j]!=idxb[n][j]) same = 0;
		}
	if(!same)
		{
		pc_ctx_free();
		const char *e = getenv("HPMPC_B200_DEVICE");
		if(e) dev = atoi(e); else if(cudaGetDevice(&dev)!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: no CUDA device available\n"); return -1; }
		if(hpmpc_b200_pcond_create(&GP.h, N, nx, nu, nb, idxb, ng, N2, dev)) { GP.h = NULL; return -1; }
		GP.N = N; GP.N2 = N2;
		GP.nx = malloc((N+1)*sizeof(int)); GP.nu = malloc((N+1)*sizeof(int)); GP.nb = malloc((N+1)*sizeof(int)); GP.ng = malloc((N+1)*sizeof(int));
		GP.idxb = calloc(N+1, sizeof(int*));
		for(n=0; n<=N; n++)
			{
			GP.nx[n] = nx[n]; GP.nu[n] = n<N ? nu[n] : 0; GP.nb[n] = nb ? nb[n] : 0; GP.ng[n] = ng ? ng[n] : 0;
			GP.idxb[n] = malloc((GP.nb[n]+1)*sizeof(int));
			for(j=0; j<GP.nb[n]; j++) GP.idxb[n][j] = idxb[n][j];
			}
		hpmpc_b200_ocp_sizes(hpmpc_b200_pcond_full(GP.h), &GP.sz);
		const size_t lam = (size_t)(GP.sz.lam_stride>0 ? GP.sz.lam_stride : 2);
		GP.h_in = calloc(GP.sz.in_stride, sizeof(double)); GP.h_ux = calloc(GP.sz.ux_stride, sizeof(double));
		GP.h_pi = calloc(GP.sz.pi_stride+2, sizeof(double)); GP.h_lam = calloc(lam, sizeof(double));
		if(cudaMalloc((void**)&GP.d_in, sizeof(double)*GP.sz.in_stride)!=cudaSuccess
		|| cudaMalloc((void**)&GP.d_ux, sizeof(double)*GP.sz.ux_stride)!=cudaSuccess
		|| cudaMalloc((void**)&GP.d_pi, sizeof(double)*(GP.sz.pi_stride+2))!=cudaSuccess
		|| cudaMalloc((void**)&GP.d_lam, sizeof(double)*lam)!=cudaSuccess
		|| cudaMalloc((void**)&GP.d_t, sizeof(double)*lam)!=cudaSuccess)
			{ fprintf(stderr, "hpmpc_b200: device allocation failed\n"); return -1; }
		}
	if(k_max>GP.k_max_alloc)
		{
		free(GP.h_info); if(GP.d_info) cudaFree(GP.d_info);
		GP.h_info = calloc(HB_IPM_INFO_HEAD+5*k_max, sizeof(double));
		if(cudaMalloc((void**)&GP.d_info, sizeof(double)*(HB_IPM_INFO_HEAD+5*k_max))!=cudaSuccess) return -1;
		GP.k_max_alloc = k_max;
		}
	return 0;
	}

/* the N2 < N branch of the reference's wrappers (interfaces/c/fortran_order_interface.c:389-528): condense, IPM on the condensed
 * problem from a cold start (the reference never fills the condensed initial guess, :493-507), expand */
static int high_level_part_cond(int c_order, int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu, int *nb, int **hidxb,
		int *ng, int N2, double **A, double **B, double **b, double **Q, double **S, double **R, double **q,
		double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug,
		double **x, double **u, double **pi, double **lam, double *inf_norm_res, double *stat)
	{
	int n, i, j, status;
	const double alpha_min = 1e-8;
	if(pc_ctx_get(N, nx, nu, nb, hidxb, ng, N2, k_max)) { fprintf(stderr, "hpmpc_b200: partial condensing: GPU context unavailable\n"); return -1; }
	hpmpc_b200_ocp *full = hpmpc_b200_pcond_full(GP.h);
	hpmpc_b200_pack_instance(full, c_order, A, B, b, Q, S, R, q, r, lb, ub, GP.h_in);
	if(ng && ng[N]>0 && C && D && lg && ug) hpmpc_b200_pack_general(full, c_order, C, D, lg, ug, GP.h_in);
	if(h2d(GP.d_in, GP.h_in, GP.sz.in_stride)
	|| cudaMemset(GP.d_info, 0, sizeof(double)*(HB_IPM_INFO_HEAD+5*k_max))!=cudaSuccess
	|| hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(GP.h, 1, GP.d_in, k_max, mu0, mu_tol, alpha_min, GP.d_ux, GP.d_pi, GP.d_lam, GP.d_t, GP.d_info, NULL)
	|| cudaDeviceSynchronize()!=cudaSuccess
	|| d2h(GP.h_ux, GP.d_ux, GP.sz.ux_stride) || d2h(GP.h_pi, GP.d_pi, GP.sz.pi_stride) || d2h(GP.h_info, GP.d_info, HB_IPM_INFO_HEAD+5*k_max)
	|| (GP.sz.lam_stride>0 && d2h(GP.h_lam, GP.d_lam, GP.sz.lam_stride)))
		{ fprintf(stderr, "hpmpc_b200: partial condensing: GPU execution failed\n"); return -1; }
	*kk = (int)GP.h_info[0];
	status = (int)GP.h_info[1];
	if(stat) for(i=0; i<5*(*kk); i++) stat[i] = GP.h_info[HB_IPM_INFO_HEAD+i];
	hpmpc_b200_unpack_solution(full, GP.h_ux, GP.h_pi, GP.h_lam, x, u, pi, lam);
	for(n=0; n<N; n++)
		for(j=0; j<nb[n] && hidxb[n][j]<nu[n]; j++)
			if(lb[n][j]==ub[n][j]) u[n][hidxb[n][j]] = lb[n][j];
	for(i=0; i<4; i++) inf_norm_res[i] = GP.h_info[2+i];
	return status;
	}

/* ---- the lib4 routines themselves (reference include/lqcp_solvers.h:88-95, lqcp_solvers/d_part_cond.c:739, :860, :926, :1072, :1103):
 *      panel-major in and out, a batch of one through the device kernels.  `memory` receives the condensed matrices in the
 *      reference's own arrangement (all [B A b]', then all RSQrq, all [D C]', all d, then the idxb arrays); `work` is not used ---- */
static void pc_cond_dims(int N2, const int *nx2, const int *nu2, const int *nb2, const int *ng2, int k, int *pnz, int *pnux, int *pnb, int *png,
		int *cnx1, int *cnux, int *cng)
	{
	(void)N2;
	*pnz = RUP(nu2[k]+nx2[k]+1, BS); *pnux = RUP(nu2[k]+nx2[k], BS); *pnb = RUP(nb2[k], BS); *png = RUP(ng2[k], BS);
	*cnx1 = RUP(nx2[k+1], NCL); *cnux = RUP(nu2[k]+nx2[k], NCL); *cng = RUP(ng2[k], NCL);
	}

int d_part_cond_memory_space_size_bytes(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int *nx2, int *nu2, int *nb2, int *ng2)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)hidxb; (void)ng;
	int k, pnz, pnux, pnb, png, cnx1, cnux, cng;
	size_t d = 0, it = 0;
	for(k=0; k<N2; k++)
		{
		pc_cond_dims(N2, nx2, nu2, nb2, ng2, k, &pnz, &pnux, &pnb, &png, &cnx1, &cnux, &cng);
		d += (size_t)pnz*cnx1 + (size_t)pnz*cnux + (size_t)pnux*cng + 2*pnb + 2*png;
		it += nb2[k];
		}
	return (int)(d*sizeof(double) + it*sizeof(int) + 64);
	}
int d_part_cond_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2, int *nx2, int *nu2, int *nb2, int *ng2)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)hidxb; (void)ng; (void)N2; (void)nx2; (void)nu2; (void)nb2; (void)ng2;
	return 64;
	}
int d_part_expand_work_space_size_bytes(int N, int *nx, int *nu, int *nb, int *ng)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)ng;
	return 64;
	}

void d_part_cond(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **hpBAbt, double **hpRSQrq, double **hpDCt, double **hd,
		int N2, int *nx2, int *nu2, int *nb2, int **hidxb2, int *ng2, double **hpBAbt2, double **hpRSQrq2, double **hpDCt2, double **hd2,
		void *memory, void *work)
	{
	(void)work;
	int k, i, j;
	if(N2>=N)
		{
		/* nothing to condense: the condensed problem IS the problem (d_part_cond.c:937-962) */
		for(k=0; k<=N; k++)
			{
			nx2[k] = nx[k]; nu2[k] = nu[k]; nb2[k] = nb[k]; hidxb2[k] = hidxb[k]; ng2[k] = ng[k];
			if(k<N) hpBAbt2[k] = hpBAbt[k];
			hpRSQrq2[k] = hpRSQrq[k]; hpDCt2[k] = hpDCt[k]; hd2[k] = hd[k];
			}
		return;
		}
	pthread_mutex_lock(&g_lock);
	if(pc_ctx_get(N, nx, nu, nb, hidxb, ng, N2, 1)) { pthread_mutex_unlock(&g_lock); fatal("d_part_cond: GPU context unavailable"); }
	hpmpc_b200_ocp *full = hpmpc_b200_pcond_full(GP.h), *cond = hpmpc_b200_pcond_cond(GP.h);
	hpmpc_b200_sizes szc;
	hpmpc_b200_ocp_sizes(cond, &szc);
	pack_from_pmat_to(full, &GP.sz, GP.h_in, GP.ng, N, nx, nu, nb, hpBAbt, hpRSQrq, hd, hpDCt);
	double *d_in2 = NULL, *h2 = calloc(szc.in_stride, sizeof(double));
	int bad = h2==NULL || cudaMalloc((void**)&d_in2, sizeof(double)*szc.in_stride)!=cudaSuccess
		|| h2d(GP.d_in, GP.h_in, GP.sz.in_stride)
		|| hpmpc_b200_d_part_cond_batch(GP.h, 1, GP.d_in, d_in2, NULL)
		|| cudaDeviceSynchronize()!=cudaSuccess
		|| d2h(h2, d_in2, szc.in_stride);
	cudaFree(d_in2);
	if(bad) { free(h2); pthread_mutex_unlock(&g_lock); fatal("d_part_cond: GPU execution failed"); }
	/* the bound positions (d_cond_DCtd assigns them while it condenses) */
	{
	int *t0 = malloc((N2+1)*sizeof(int)), *t1 = malloc((N2+1)*sizeof(int)), *t2 = malloc((N2+1)*sizeof(int)), *t3 = malloc((N2+1)*sizeof(int));
	int **idx = calloc(N2+1, sizeof(int*)), nbt = 1;
	for(k=0; k<=N; k++) nbt += nb[k];
	for(k=0; k<=N2; k++) idx[k] = calloc(nbt, sizeof(int));
	hpmpc_b200_part_cond_compute_problem_size(N, nx, nu, nb, hidxb, ng, N2, t0, t1, t2, t3, idx);
	/* memory: matrices per kind, then the int arrays */
	double *ptr = (double*)memory;
	int pnz, pnux, pnb, png, cnx1, cnux, cng;
	for(k=0; k<N2; k++) { pc_cond_dims(N2, nx2, nu2, nb2, ng2, k, &pnz, &pnux, &pnb, &png, &cnx1, &cnux, &cng); hpBAbt2[k] = ptr; ptr += (size_t)pnz*cnx1; }
	for(k=0; k<N2; k++) { pc_cond_dims(N2, nx2, nu2, nb2, ng2, k, &pnz, &pnux, &pnb, &png, &cnx1, &cnux, &cng); hpRSQrq2[k] = ptr; ptr += (size_t)pnz*cnux; }
	for(k=0; k<N2; k++) { pc_cond_dims(N2, nx2, nu2, nb2, ng2, k, &pnz, &pnux, &pnb, &png, &cnx1, &cnux, &cng); hpDCt2[k] = ptr; ptr += (size_t)pnux*cng; }
	for(k=0; k<N2; k++) { pc_cond_dims(N2, nx2, nu2, nb2, ng2, k, &pnz, &pnux, &pnb, &png, &cnx1, &cnux, &cng); hd2[k] = ptr; ptr += 2*pnb + 2*png; }
	int *ip = (int*)ptr;
	for(k=0; k<N2; k++) { hidxb2[k] = ip; ip += nb2[k]; for(j=0; j<nb2[k]; j++) hidxb2[k][j] = idx[k][j]; }
	for(k=0; k<=N2; k++) free(idx[k]);
	free(idx); free(t0); free(t1); free(t2); free(t3);
	}
	for(k=0; k<N2; k++)
		{
		int oB, oH, oD, oG, oDg, ngk, pnz, pnux, pnb, png, cnx1, cnux, cng;
		const int nux2 = nu2[k]+nx2[k], nx1 = nx2[k+1];
		pc_cond_dims(N2, nx2, nu2, nb2, ng2, k, &pnz, &pnux, &pnb, &png, &cnx1, &cnux, &cng);
		hpmpc_b200_ocp_stage_offsets(cond, k, &oB, &oH, &oD, NULL, NULL, NULL, NULL);
		hpmpc_b200_ocp_general_offsets(cond, k, &ngk, &oG, &oDg, NULL);
		memset(hpBAbt2[k], 0, sizeof(double)*(size_t)pnz*cnx1); memset(hpRSQrq2[k], 0, sizeof(double)*(size_t)pnz*cnux);
		memset(hpDCt2[k], 0, sizeof(double)*(size_t)pnux*cng); memset(hd2[k], 0, sizeof(double)*(2*pnb+2*png));
		for(i=0; i<=nux2; i++) for(j=0; j<nx1; j++) PM(hpBAbt2[k], cnx1, i, j) = h2[oB+i*nx1+j];
		for(i=0; i<nux2; i++) for(j=0; j<=i; j++) PM(hpRSQrq2[k], cnux, i, j) = h2[oH+HB_TRI(i)+j];
		for(j=0; j<nux2; j++) PM(hpRSQrq2[k], cnux, nux2, j) = h2[oH+HB_TRI(nux2)+j];
		for(i=0; i<nux2; i++) for(j=0; j<ng2[k]; j++) PM(hpDCt2[k], cng, i, j) = h2[oG+i*ng2[k]+j];
		for(j=0; j<nb2[k]; j++) { hd2[k][j] = h2[oD+j]; hd2[k][pnb+j] = h2[oD+nb2[k]+j]; }
		for(j=0; j<ng2[k]; j++) { hd2[k][2*pnb+j] = h2[oDg+j]; hd2[k][2*pnb+png+j] = h2[oDg+ng2[k]+j]; }
		}
	/* the last stage is shared with the full problem (d_part_cond.c:1058-1062) */
	hpRSQrq2[N2] = hpRSQrq[N]; hpDCt2[N2] = hpDCt[N]; hd2[N2] = hd[N]; hidxb2[N2] = hidxb[N];
	free(h2);
	pthread_mutex_unlock(&g_lock);
	}

void d_part_expand_solution(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **hpBAbt, double **hb, double **hpRSQrq, double **hrq,
		double **hpDCt, double **hux, double **hpi, double **hlam, double **ht, int N2, int *nx2, int *nu2, int *nb2, int **hidxb2, int *ng2,
		double **hux2, double **hpi2, double **hlam2, double **ht2, void *work)
	{
	(void)work; (void)hidxb2;
	int n, k, i;
	if(N2>=N)
		{
		for(n=0; n<=N; n++)
			{
			const int nbn = nb[n], pnb = RUP(nbn, BS), ngn = ng[n], png = RUP(ngn, BS);
			if(hux[n]!=hux2[n]) for(i=0; i<nu[n]+nx[n]; i++) hux[n][i] = hux2[n][i];
			if(n<N && hpi[n]!=hpi2[n]) for(i=0; i<nx[n+1]; i++) hpi[n][i] = hpi2[n][i];
			if(hlam[n]!=hlam2[n]) for(i=0; i<2*pnb+2*png; i++) { hlam[n][i] = hlam2[n][i]; ht[n][i] = ht2[n][i]; }
			}
		return;
		}
	pthread_mutex_lock(&g_lock);
	if(pc_ctx_get(N, nx, nu, nb, hidxb, ng, N2, 1)) { pthread_mutex_unlock(&g_lock); fatal("d_part_expand_solution: GPU context unavailable"); }
	hpmpc_b200_ocp *full = hpmpc_b200_pcond_full(GP.h), *cond = hpmpc_b200_pcond_cond(GP.h);
	hpmpc_b200_sizes szc;
	hpmpc_b200_ocp_sizes(cond, &szc);
	pack_from_pmat_to(full, &GP.sz, GP.h_in, GP.ng, N, nx, nu, nb, hpBAbt, hpRSQrq, NULL, hpDCt);
	for(n=0; n<=N; n++)
		{
		int oB, oH, nux = nu[n]+nx[n];
		hpmpc_b200_ocp_stage_offsets(full, n, &oB, &oH, NULL, NULL, NULL, NULL, NULL);
		if(n<N && hb) for(i=0; i<nx[n+1]; i++) GP.h_in[oB + (size_t)nux*nx[n+1] + i] = hb[n][i];
		if(hrq) for(i=0; i<nux; i++) GP.h_in[oH + HB_TRI(nux) + i] = hrq[n][i];
		}
	const size_t lamc = (size_t)(szc.lam_stride>0 ? szc.lam_stride : 2), lamf = (size_t)(GP.sz.lam_stride>0 ? GP.sz.lam_stride : 2);
	double *h_ux2 = calloc(szc.ux_stride, sizeof(double)), *h_pi2 = calloc(szc.pi_stride+2, sizeof(double));
	double *h_lam2 = calloc(lamc, sizeof(double)), *h_t2 = calloc(lamc, sizeof(double)), *h_t = calloc(lamf, sizeof(double));
	for(k=0; k<=N2; k++)
		{
		int oU, oP, oL;
		const int nbk = nb2[k], pnb = RUP(nbk, BS), ngk = ng2[k], png = RUP(ngk, BS);
		hpmpc_b200_ocp_stage_offsets(cond, k, NULL, NULL, NULL, &oU, &oP, &oL, NULL);
		for(i=0; i<nu2[k]+nx2[k]; i++) h_ux2[oU+i] = hux2[k][i];
		if(k<N2) for(i=0; i<nx2[k+1]; i++) h_pi2[oP+i] = hpi2[k][i];
		for(i=0; i<nbk; i++) { h_lam2[oL+i] = hlam2[k][i]; h_lam2[oL+nbk+i] = hlam2[k][pnb+i]; h_t2[oL+i] = ht2[k][i]; h_t2[oL+nbk+i] = ht2[k][pnb+i]; }
		for(i=0; i<ngk; i++)
			{
			h_lam2[oL+2*nbk+i] = hlam2[k][2*pnb+i]; h_lam2[oL+2*nbk+ngk+i] = hlam2[k][2*pnb+png+i];
			h_t2[oL+2*nbk+i] = ht2[k][2*pnb+i]; h_t2[oL+2*nbk+ngk+i] = ht2[k][2*pnb+png+i];
			}
		}
	double *d2 = NULL;
	const size_t n2 = (size_t)szc.ux_stride + szc.pi_stride + 2 + 2*lamc;
	int bad = !h_ux2 || !h_pi2 || !h_lam2 || !h_t2 || !h_t || cudaMalloc((void**)&d2, sizeof(double)*n2)!=cudaSuccess;
	double *d_ux2 = d2, *d_pi2 = d2 ? d_ux2 + szc.ux_stride : NULL, *d_lam2 = d2 ? d_pi2 + szc.pi_stride + 2 : NULL, *d_t2 = d2 ? d_lam2 + lamc : NULL;
	bad = bad || h2d(GP.d_in, GP.h_in, GP.sz.in_stride) || h2d(d_ux2, h_ux2, szc.ux_stride) || h2d(d_pi2, h_pi2, szc.pi_stride)
		|| h2d(d_lam2, h_lam2, lamc) || h2d(d_t2, h_t2, lamc)
		|| hpmpc_b200_d_part_expand_solution_batch(GP.h, 1, GP.d_in, d_ux2, d_pi2, d_lam2, d_t2, GP.d_ux, GP.d_pi, GP.d_lam, GP.d_t, NULL)
		|| cudaDeviceSynchronize()!=cudaSuccess
		|| d2h(GP.h_ux, GP.d_ux, GP.sz.ux_stride) || d2h(GP.h_pi, GP.d_pi, GP.sz.pi_stride)
		|| (GP.sz.lam_stride>0 && (d2h(GP.h_lam, GP.d_lam, GP.sz.lam_stride) || d2h(h_t, GP.d_t, GP.sz.lam_stride)));
	cudaFree(d2);
	if(!bad)
		for(n=0; n<=N; n++)
			{
			int oU, oP, oL;
			const int nbn = nb[n], pnb = RUP(nbn, BS), ngn = ng[n], png = RUP(ngn, BS);
			hpmpc_b200_ocp_stage_offsets(full, n, NULL, NULL, NULL, &oU, &oP, &oL, NULL);
			for(i=0; i<nu[n]+nx[n]; i++) hux[n][i] = GP.h_ux[oU+i];
			if(n<N) for(i=0; i<nx[n+1]; i++) hpi[n][i] = GP.h_pi[oP+i];
			for(i=0; i<nbn; i++) { hlam[n][i] = GP.h_lam[oL+i]; hlam[n][pnb+i] = GP.h_lam[oL+nbn+i]; ht[n][i] = h_t[oL+i]; ht[n][pnb+i] = h_t[oL+nbn+i]; }
			for(i=0; i<ngn; i++)
				{
				hlam[n][2*pnb+i] = GP.h_lam[oL+2*nbn+i]; hlam[n][2*pnb+png+i] = GP.h_lam[oL+2*nbn+ngn+i];
				ht[n][2*pnb+i] = h_t[oL+2*nbn+i]; ht[n][2*pnb+png+i] = h_t[oL+2*nbn+ngn+i];
				}
			}
	free(h_ux2); free(h_pi2); free(h_lam2); free(h_t2); free(h_t);
	pthread_mutex_unlock(&g_lock);
	if(bad) fatal("d_part_expand_solution: GPU execution failed");
	}

static int high_level(int c_order, int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu, int *nb, int **hidxb,
		int *ng, int N2, int warm_start, double **A, double **B, double **b, double **Q, double **S, double **R, double **q,
		double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug,
		double **x, double **u, double **pi, double **lam, double *inf_norm_res, double *stat, const void *work0)
	{
	int n, i, j, l, status;
	const double alpha_min = 1e-8;       /* c_order_interface.c:141 */
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu, nb, hidxb, ng, k_max)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: GPU context unavailable\n"); return -1; }
	hpmpc_b200_pack_instance(G.ocp, c_order, A, B, b, Q, S, R, q, r, lb, ub, G.h_in);
	/* general constraints: [D C]' per stage.  (The reference's c_order wrapper transposes C[N] once more than the other stages,
	 * interfaces/c/c_order_interface.c:283 against :279-280 and fortran_order_interface.c:283 -- visible only for a non-symmetric
	 * C[N]; here both orders mean the same matrix, see INTEGRATION.md.) */
	if(G.sz.lam_stride>0 && C && D && lg && ug) hpmpc_b200_pack_general(G.ocp, c_order, C, D, lg, ug, G.h_in);
	/* mu0 estimate when the caller passes mu0 <= 0: signed maximum over the cost entries
	 * (c_order_interface.c:318-331 / fortran_order_interface.c:318-331) */
	if(mu0<=0)
		{
		for(n=0; n<=N; n++)
			{
			int nun = n<N ? nu[n] : 0;
			if(n<N)
				{
				for(j=0; j<nun*nun; j++) mu0 = fmax(mu0, R[n][j]);
				for(j=0; j<nx[n]*nun; j++) mu0 = fmax(mu0, S[n][j]);
				for(j=0; j<nun; j++) mu0 = fmax(mu0, r[n][j]);
				}
			for(j=0; j<nx[n]; j++) for(l=0; l<nx[n]; l++) mu0 = fmax(mu0, Q[n][j*nx[n]+l]);
			for(j=0; j<nx[n]; j++) mu0 = fmax(mu0, q[n][j]);
			}
		}
	/* partial condensing: 1 <= N2 < N and no general constraints before stage N (the reference falls back to N2 = N otherwise,
	 * interfaces/c/fortran_order_interface.c:85-97) */
	{
	int cond = N2>=1 && N2<N && nb!=NULL;
	if(ng) for(n=0; n<N; n++) if(ng[n]>0) cond = 0;
	if(cond)
		{
		status = high_level_part_cond(c_order, kk, k_max, mu0, mu_tol, N, nx, nu, nb, hidxb, ng, N2, A, B, b, Q, S, R, q, r, lb, ub,
				C, D, lg, ug, x, u, pi, lam, inf_norm_res, stat);
		pthread_mutex_unlock(&g_lock);
		return status;
		}
	}
	if(warm_start)
		for(n=0; n<=N; n++)
			{
			int oU, nun = n<N ? nu[n] : 0;
			hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, NULL, NULL, NULL);
			for(i=0; i<nun; i++) G.h_ux[oU+i] = u[n][i];
			for(i=0; i<nx[n]; i++) G.h_ux[oU+nun+i] = x[n][i];
			}
	/* the KKT state stays on the device, keyed by the caller's work0 (the reference keeps it IN work0): a later
	 * {c,fortran}_order_d_solve_kkt_new_rhs_ocp_hard_tv with the same work0 re-solves with it */
	if(run_ipm_single(kk, k_max, mu0, mu_tol, alpha_min, warm_start, stat, work0)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: GPU execution failed\n"); return -1; }
	status = (int)G.h_info[1];
	hpmpc_b200_unpack_solution(G.ocp, G.h_ux, G.h_pi, G.h_lam, x, u, pi, lam);
	/* inputs fixed by lb == ub are returned exactly on the bound (c_order_interface.c:599-608) */
	for(n=0; n<N; n++)
		for(j=0; j<nb[n] && hidxb[n][j]<nu[n]; j++)
			if(lb[n][j]==ub[n][j]) u[n][hidxb[n][j]] = lb[n][j];
	for(i=0; i<4; i++) inf_norm_res[i] = G.h_info[2+i];
	pthread_mutex_unlock(&g_lock);
	return status;
	}

int c_order_d_ip_ocp_hard_tv(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu, int *nb, int **hidxb, int *ng,
		int N2, int warm_start, double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r,
		double **lb, double **ub, double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
		double **lam, double *inf_norm_res, void *work0, double *stat)
	{
	return high_level(1, kk, k_max, mu0, mu_tol, N, nx, nu, nb, hidxb, ng, N2, warm_start, A, B, b, Q, S, R, q, r, lb, ub, C, D, lg, ug, x, u, pi, lam, inf_norm_res, stat, work0);
	}

int fortran_order_d_ip_ocp_hard_tv(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu, int *nb, int **hidxb, int *ng,
		int N2, int warm_start, double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r,
		double **lb, double **ub, double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
		double **lam, double *inf_norm_res, void *work0, double *stat)
	{
	return high_level(0, kk, k_max, mu0, mu_tol, N, nx, nu, nb, hidxb, ng, N2, warm_start, A, B, b, Q, S, R, q, r, lb, ub, C, D, lg, ug, x, u, pi, lam, inf_norm_res, stat, work0);
	}

/* reference include/c_interface.h:66 (interfaces/c/fortran_order_interface.c:695): k_max Newton steps from (ux0, pi0, lam0, t0) on
 * column-major stage arrays; x, u, pi, lam, t receive the new iterate, inf_norm_res its residual norms. */
int fortran_order_d_ip_ocp_hard_tv_single_newton_step(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu_N, int *nb,
		int **hidxb, int *ng, int N2, int warm_start, double **A, double **B, double **b, double **Q, double **S, double **R, double **q,
		double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
		double **lam, double **t, double *inf_norm_res, void *work0, double *stat, double **ux0, double **pi0, double **lam0, double **t0)
	{
	(void)mu_tol; (void)N2; (void)warm_start; (void)work0; (void)C; (void)D; (void)lg; (void)ug;
	int n, i, j, status;
	const double alpha_min = 1e-8;
	pthread_mutex_lock(&g_lock);
	if(ctx_get(N, nx, nu_N, nb, hidxb, ng, k_max)) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: GPU context unavailable\n"); return -1; }
	if(G.sz.nbtot==0 || k_max<1) { pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: single_newton_step needs bounds and k_max >= 1\n"); return -1; }
	hpmpc_b200_pack_instance(G.ocp, 0, A, B, b, Q, S, R, q, r, lb, ub, G.h_in);
	for(n=0; n<=N; n++)
		{
		int oU, oP, oLm, nun = n<N ? nu_N[n] : 0, nbn = nb[n];
		hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, &oU, &oP, &oLm, NULL);
		for(i=0; i<nun+nx[n]; i++) G.h_ux[oU+i] = ux0[n][i];
		if(n<N) for(i=0; i<nx[n+1]; i++) G.h_pi[oP+i] = pi0[n][i];
		for(i=0; i<2*nbn; i++) { G.h_lam[oLm+i] = lam0[n][i]; G.h_t[oLm+i] = t0[n][i]; }
		}
	if(h2d(G.d_in, G.h_in, G.sz.in_stride) || h2d(G.d_ux, G.h_ux, G.sz.ux_stride) || h2d(G.d_pi, G.h_pi, G.sz.pi_stride)
	|| h2d(G.d_lam, G.h_lam, G.sz.lam_stride) || h2d(G.d_t, G.h_t, G.sz.lam_stride)
	|| cudaMemset(G.d_info, 0, sizeof(double)*(HB_IPM_INFO_HEAD+5*k_max))!=cudaSuccess
	|| hpmpc_b200_d_ip2_res_mpc_hard_single_newton_step_batch(G.ocp, 1, G.d_in, k_max, mu0, alpha_min, G.d_ux, G.d_pi, G.d_lam, G.d_t, G.d_info, NULL)
	|| cudaDeviceSynchronize()!=cudaSuccess
	|| d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride) || d2h(G.h_info, G.d_info, HB_IPM_INFO_HEAD+5*k_max)
	|| d2h(G.h_lam, G.d_lam, G.sz.lam_stride) || d2h(G.h_t, G.d_t, G.sz.lam_stride))
		{ pthread_mutex_unlock(&g_lock); fprintf(stderr, "hpmpc_b200: single_newton_step: GPU execution failed\n"); return -1; }
	*kk = (int)G.h_info[0];
	if(stat) for(i=0; i<5*(*kk); i++) stat[i] = G.h_info[HB_IPM_INFO_HEAD+i];
	status = (int)G.h_info[1];
	hpmpc_b200_unpack_solution(G.ocp, G.h_ux, G.h_pi, G.h_lam, x, u, pi, lam);
	for(n=0; n<=N; n++) { int oLm; hpmpc_b200_ocp_stage_offsets(G.ocp, n, NULL, NULL, NULL, NULL, NULL, &oLm, NULL); for(i=0; i<2*nb[n]; i++) t[n][i] = G.h_t[oLm+i]; }
	for(n=0; n<N; n++)
		for(j=0; j<nb[n] && hidxb[n][j]<nu_N[n]; j++)
			if(lb[n][j]==ub[n][j]) u[n][hidxb[n][j]] = lb[n][j];
	for(i=0; i<4; i++) inf_norm_res[i] = G.h_info[2+i];
	pthread_mutex_unlock(&g_lock);
	return status;
	}

/* ------------------------------------------------------------------------------------------------ */
/* "libstr" twins (row a8): the same entry points on BLASFEO containers (include/hpmpc_blasfeo_compat.h).       */
/* Each call re-expresses its arguments in the lib4 conventions (host-side format conversion only) and      */
/* goes through the lib4 symbol above -- one code path to the kernels.                                     */
/* ------------------------------------------------------------------------------------------------ */
#define PMS(A, i, j) ((A)->pA[((i)/BS)*BS*(A)->cn + (i)%BS + BS*(j)])

/* blasfeo_dmat (m x n used) -> lib4 panel-major buffer with sda = n rounded up to ncl */
static double *ls_to_lib4(const struct blasfeo_dmat *A, int m, int n)
	{
	int i, j, sda = RUP(n>0 ? n : 1, NCL), pm = RUP(m>0 ? m : 1, BS);
	double *p = calloc((size_t)pm*sda + 8, sizeof(double));
	for(i=0; i<m; i++) for(j=0; j<n; j++) PM(p, sda, i, j) = PMS(A, i, j);
	return p;
	}

typedef struct { double **BAbt, **RSQ, **DCt, **Qx, **qx, **bd, **b, **q, **ux, **pi, **Pb; int N; } ls_args;

static void ls_free(ls_args *a)
	{
	int n;
	for(n=0; n<=a->N; n++)
		{
		if(a->BAbt) free(a->BAbt[n]);
		if(a->RSQ) free(a->RSQ[n]);
		if(a->DCt) free(a->DCt[n]);
		if(a->Qx) free(a->Qx[n]);
		if(a->qx) free(a->qx[n]);
		if(a->bd) free(a->bd[n]);
		}
	free(a->BAbt); free(a->RSQ); free(a->DCt); free(a->Qx); free(a->qx); free(a->bd); free(a->b); free(a->q); free(a->ux); free(a->pi); free(a->Pb);
	}

/* matrices, per-constraint vectors ([bounds | general] -> [bounds (pnb) | general]) and the node-indexed pi / Pb as edge arrays */
static void ls_convert(ls_args *a, int N, int *nx, int *nu, int *nb, int **idxb, int *ng, struct blasfeo_dmat *hsBAbt, struct blasfeo_dmat *hsRSQrq,
		struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsQx, struct blasfeo_dvec *hsqx, struct blasfeo_dvec *hsb, struct blasfeo_dvec *hsrq,
		struct blasfeo_dvec *hsux, struct blasfeo_dvec *hspi, struct blasfeo_dvec *hsPb)
	{
	int n, j;
	memset(a, 0, sizeof(*a));
	a->N = N;
	a->BAbt = calloc(N+1, sizeof(double*)); a->RSQ = calloc(N+1, sizeof(double*)); a->DCt = calloc(N+1, sizeof(double*));
	a->Qx = calloc(N+1, sizeof(double*)); a->qx = calloc(N+1, sizeof(double*)); a->bd = calloc(N+1, sizeof(double*));
	a->b = calloc(N+1, sizeof(double*)); a->q = calloc(N+1, sizeof(double*)); a->ux = calloc(N+1, sizeof(double*));
	a->pi = calloc(N+1, sizeof(double*)); a->Pb = calloc(N+1, sizeof(double*));
	for(n=0; n<=N; n++)
		{
		int nun = n<N ? nu[n] : 0, nux = nun+nx[n], nbn = nb ? nb[n] : 0, ngn = ng ? ng[n] : 0, pnb = nbn>0 ? RUP(nbn, BS) : 0;
		if(n<N && hsBAbt) a->BAbt[n] = ls_to_lib4(&hsBAbt[n], nux+1, nx[n+1]);
		if(hsRSQrq) a->RSQ[n] = ls_to_lib4(&hsRSQrq[n], nux+1, nux);
		if(ngn>0 && hsDCt) a->DCt[n] = ls_to_lib4(&hsDCt[n], nux, ngn);
		a->Qx[n] = calloc(pnb+ngn+4, sizeof(double)); a->qx[n] = calloc(pnb+ngn+4, sizeof(double)); a->bd[n] = calloc(pnb+4, sizeof(double));
		for(j=0; j<nbn; j++)
			{
			if(hsQx) a->Qx[n][j] = hsQx[n].pa[j];
			if(hsqx) a->qx[n][j] = hsqx[n].pa[j];
			if(hsRSQrq) a->bd[n][j] = PMS(&hsRSQrq[n], idxb[n][j], idxb[n][j]);     /* libstr ADDS Qx to the diagonal (:102): bd = the diagonal itself */
			}
		for(j=0; j<ngn; j++) { if(hsQx) a->Qx[n][pnb+j] = hsQx[n].pa[nbn+j]; if(hsqx) a->qx[n][pnb+j] = hsqx[n].pa[nbn+j]; }
		if(n<N && hsb) a->b[n] = hsb[n].pa;
		if(hsrq) a->q[n] = hsrq[n].pa;
		if(hsux) a->ux[n] = hsux[n].pa;
		if(n<N && hspi) a->pi[n] = hspi[n+1].pa;
		if(n<N && hsPb) a->Pb[n] = hsPb[n+1].pa;
		}
	}

int d_back_ric_rec_work_space_size_bytes_libstr(int N, int *nx, int *nu, int *nb, int *ng)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)ng;
	return 64;
	}

/* the factor lives in hsL[n].pA in this library's layout: packed trapezoid + inverse diagonal, HB_EVEN(tri(nux)+2 nux) doubles,
 * which fits the (nux+1) x nux panel-major matrix the caller allocated for every nux >= 1 */
static double *ls_gather_L(int N, int *nx, int *nu, struct blasfeo_dmat *hsL, int to_struct, double *mem)
	{
	int n; size_t off = 0;
	for(n=0; n<=N; n++)
		{
		int nux = (n<N ? nu[n] : 0) + nx[n]; size_t len = HB_EVEN(HB_TRI(nux)+2*nux);
		if(to_struct) memcpy(hsL[n].pA, mem+off, sizeof(double)*len); else memcpy(mem+off, hsL[n].pA, sizeof(double)*len);
		off += len;
		}
	return mem;
	}

void d_back_ric_rec_sv_libstr(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int update_b, struct blasfeo_dmat *hsBAbt,
		struct blasfeo_dvec *hsb, int update_q, struct blasfeo_dmat *hsRSQrq, struct blasfeo_dvec *hsrq, struct blasfeo_dmat *hsDCt,
		struct blasfeo_dvec *hsQx, struct blasfeo_dvec *hsqx, struct blasfeo_dvec *hsux, int compute_pi, struct blasfeo_dvec *hspi,
		int compute_Pb, struct blasfeo_dvec *hsPb, struct blasfeo_dmat *hsL, void *work_space)
	{
	(void)work_space;
	ls_args a;
	int any = 0, n;
	for(n=0; n<=N; n++) if((nb && nb[n]>0) || (ng && ng[n]>0)) any = 1;
	ls_convert(&a, N, nx, nu, nb, hidxb, ng, hsBAbt, hsRSQrq, hsDCt, any ? hsQx : NULL, any ? hsqx : NULL, hsb, hsrq, hsux, hspi, hsPb);
	double *mem = calloc((size_t)d_back_ric_rec_sv_tv_memory_space_size_bytes(N, nx, nu, nb, ng)/sizeof(double) + 8, sizeof(double));
	d_back_ric_rec_sv_tv_res(N, nx, nu, nb, hidxb, ng, update_b, a.BAbt, a.b, update_q, a.RSQ, a.q, a.bd, a.DCt, any ? a.Qx : NULL, any ? a.qx : NULL,
			a.ux, compute_pi, a.pi, compute_Pb, a.Pb, mem, NULL);
	if(hsL) ls_gather_L(N, nx, nu, hsL, 1, mem);
	free(mem);
	ls_free(&a);
	}

void d_back_ric_rec_trf_libstr(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, struct blasfeo_dmat *hsBAbt,
		struct blasfeo_dmat *hsRSQrq, struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsQx, struct blasfeo_dmat *hsL, void *work)
	{
	(void)work;
	ls_args a;
	int any = 0, n;
	for(n=0; n<=N; n++) if((nb && nb[n]>0) || (ng && ng[n]>0)) any = 1;
	ls_convert(&a, N, nx, nu, nb, hidxb, ng, hsBAbt, hsRSQrq, hsDCt, any ? hsQx : NULL, NULL, NULL, NULL, NULL, NULL, NULL);
	double *mem = calloc((size_t)d_back_ric_rec_sv_tv_memory_space_size_bytes(N, nx, nu, nb, ng)/sizeof(double) + 8, sizeof(double));
	d_back_ric_rec_trf_tv_res(N, nx, nu, nb, hidxb, ng, a.BAbt, a.RSQ, a.DCt, any ? a.Qx : NULL, a.bd, mem, NULL);
	ls_gather_L(N, nx, nu, hsL, 1, mem);
	free(mem);
	ls_free(&a);
	}

void d_back_ric_rec_trs_libstr(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, struct blasfeo_dmat *hsBAbt, struct blasfeo_dvec *hsb,
		struct blasfeo_dvec *hsrq, struct blasfeo_dmat *hsDCt, struct blasfeo_dvec *hsqx, struct blasfeo_dvec *hsux, int compute_pi,
		struct blasfeo_dvec *hspi, int compute_Pb, struct blasfeo_dvec *hsPb, struct blasfeo_dmat *hsL, void *work)
	{
	(void)work;
	ls_args a;
	int any = 0, n;
	for(n=0; n<=N; n++) if((nb && nb[n]>0) || (ng && ng[n]>0)) any = 1;
	ls_convert(&a, N, nx, nu, nb, idxb, ng, hsBAbt, NULL, hsDCt, NULL, any ? hsqx : NULL, hsb, hsrq, hsux, hspi, hsPb);
	double *mem = calloc((size_t)d_back_ric_rec_sv_tv_memory_space_size_bytes(N, nx, nu, nb, ng)/sizeof(double) + 8, sizeof(double));
	ls_gather_L(N, nx, nu, hsL, 0, mem);
	d_back_ric_rec_trs_tv_res(N, nx, nu, nb, idxb, ng, a.BAbt, a.b, a.q, a.DCt, any ? a.qx : NULL, a.ux, compute_pi, a.pi, compute_Pb, a.Pb, mem, NULL);
	free(mem);
	ls_free(&a);
	}

int d_ip2_res_mpc_hard_work_space_size_bytes_libstr(int N, int *nx, int *nu, int *nb, int *ng)
	{
	(void)N; (void)nx; (void)nu; (void)nb; (void)ng;
	return 64;
	}

int d_ip2_res_mpc_hard_libstr(int *kk, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start, double *stat, int N,
		int *nx, int *nu, int *nb, int **idxb, int *ng, struct blasfeo_dmat *hsBAbt, struct blasfeo_dmat *hsRSQrq, struct blasfeo_dmat *hsDCt,
		struct blasfeo_dvec *hsd, struct blasfeo_dvec *hsux, int compute_mult, struct blasfeo_dvec *hspi, struct blasfeo_dvec *hslam,
		struct blasfeo_dvec *hst, void *work_memory)
	{
	ls_args a;
	int n, j, status;
	ls_convert(&a, N, nx, nu, nb, idxb, ng, hsBAbt, hsRSQrq, hsDCt, NULL, NULL, NULL, NULL, hsux, hspi, NULL);
	/* [lb(nb) lg(ng) ub(nb) ug(ng)] unpadded <-> lib4 [lb(pnb) ub(pnb) lg(png) ug(png)] */
	double **d = calloc(N+1, sizeof(double*)), **lam = calloc(N+1, sizeof(double*)), **t = calloc(N+1, sizeof(double*));
	for(n=0; n<=N; n++)
		{
		int nbn = nb ? nb[n] : 0, ngn = ng ? ng[n] : 0, pnb = nbn>0 ? RUP(nbn, BS) : 0, png = ngn>0 ? RUP(ngn, BS) : 0;
		d[n] = calloc(2*pnb+2*png+4, sizeof(double)); lam[n] = calloc(2*pnb+2*png+4, sizeof(double)); t[n] = calloc(2*pnb+2*png+4, sizeof(double));
		for(j=0; j<nbn; j++) { d[n][j] = hsd[n].pa[j]; d[n][pnb+j] = hsd[n].pa[nbn+ngn+j]; }
		for(j=0; j<ngn; j++) { d[n][2*pnb+j] = hsd[n].pa[nbn+j]; d[n][2*pnb+png+j] = hsd[n].pa[2*nbn+ngn+j]; }
		}
	status = d_ip2_res_mpc_hard_tv(kk, k_max, mu0, mu_tol, alpha_min, warm_start, stat, N, nx, nu, nb, idxb, ng, a.BAbt, a.RSQ, a.DCt, d, a.ux,
			compute_mult, a.pi, lam, t, (double*)work_memory);
	for(n=0; n<=N; n++)
		{
		int nbn = nb ? nb[n] : 0, ngn = ng ? ng[n] : 0, pnb = nbn>0 ? RUP(nbn, BS) : 0, png = ngn>0 ? RUP(ngn, BS) : 0;
		for(j=0; j<nbn; j++)
			{
			hslam[n].pa[j] = lam[n][j]; hslam[n].pa[nbn+ngn+j] = lam[n][pnb+j];
			hst[n].pa[j] = t[n][j]; hst[n].pa[nbn+ngn+j] = t[n][pnb+j];
			}
		for(j=0; j<ngn; j++)
			{
			hslam[n].pa[nbn+j] = lam[n][2*pnb+j]; hslam[n].pa[2*nbn+ngn+j] = lam[n][2*pnb+png+j];
			hst[n].pa[nbn+j] = t[n][2*pnb+j]; hst[n].pa[2*nbn+ngn+j] = t[n][2*pnb+png+j];
			}
		free(d[n]); free(lam[n]); free(t[n]);
		}
	free(d); free(lam); free(t);
	ls_free(&a);
	return status;
	}

/* ------------------------------------------------------------------------------------------------ */
/* reference include/c_interface.h:63,67 (interfaces/c/{c,fortran}_order_interface.c:1082): the last KKT system of the         */
/* preceding {c,fortran}_order_d_ip_ocp_hard_tv call on the same work0, solved again for new b, q, r and bounds.             */
/* In the reference this pair is unusable as shipped: the second routine lays work0 out differently from the first           */
/* (fortran_order_interface.c:1193 puts the IPM work space right behind the matrices, :459 puts it last), so it reads         */
/* factors and backups from the wrong place (tools/repro_highlevel_kkt_new_rhs.py shows it on the compiled reference).        */
/* Here the pair does what its interface promises: the result equals the low-level pair d_ip2_res_mpc_hard_tv +               */
/* d_kkt_solve_new_rhs_res_mpc_hard_tv, which IS pinned on the reference (tests/test_kkt_new_rhs.py).                          */
/* ------------------------------------------------------------------------------------------------ */
static void high_level_new_rhs(int c_order, int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **A, double **B, double **b,
		double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug,
		double **x, double **u, double **pi, double **lam, double *inf_norm_res, const void *work0, const char *who)
	{
	int n, i, j;
	pthread_mutex_lock(&g_lock);
	if(!same_pattern(N, nx, nu, nb, hidxb, ng) || G.kkt_key==NULL || G.kkt_key!=work0 || G.sz.nbtot==0)
		{
		pthread_mutex_unlock(&g_lock);
		fprintf(stderr, "hpmpc_b200: %s: no KKT state for this work0 -- call the IPM with the same sizes and work0 first; outputs untouched\n", who);
		return;
		}
	hpmpc_b200_pack_instance(G.ocp, c_order, A, B, b, Q, S, R, q, r, lb, ub, G.h_in);
	if(C && D && lg && ug) hpmpc_b200_pack_general(G.ocp, c_order, C, D, lg, ug, G.h_in);
	double *d_rq = NULL, *d_rb = NULL, *d_rd = NULL, *d_mu = NULL;
	const size_t lamn = (size_t)G.sz.lam_stride;
	double *h_rq = calloc(G.sz.ux_stride, sizeof(double)), *h_rb = calloc(G.sz.pi_stride+2, sizeof(double)), *h_rd = calloc(lamn, sizeof(double)), h_mu = 0.0;
	int bad = h2d(G.d_in, G.h_in, G.sz.in_stride)
		|| hpmpc_b200_d_kkt_solve_new_rhs_batch(G.ocp, 1, G.d_in, G.d_kkt, G.d_ux, G.d_pi, G.d_lam, G.d_t, G.d_info, NULL)
		|| cudaMalloc((void**)&d_rq, sizeof(double)*G.sz.ux_stride)!=cudaSuccess || cudaMalloc((void**)&d_rb, sizeof(double)*(G.sz.pi_stride+2))!=cudaSuccess
		|| cudaMalloc((void**)&d_rd, sizeof(double)*lamn)!=cudaSuccess || cudaMalloc((void**)&d_mu, sizeof(double))!=cudaSuccess
		|| hpmpc_b200_d_res_res_mpc_hard_batch(G.ocp, 1, G.d_in, G.d_ux, G.d_pi, G.d_lam, G.d_t, d_rq, d_rb, d_rd, NULL, d_mu, NULL)
		|| cudaDeviceSynchronize()!=cudaSuccess
		|| d2h(G.h_ux, G.d_ux, G.sz.ux_stride) || d2h(G.h_pi, G.d_pi, G.sz.pi_stride) || d2h(G.h_info, G.d_info, HB_IPM_INFO_HEAD)
		|| d2h(G.h_lam, G.d_lam, G.sz.lam_stride) || d2h(h_rq, d_rq, G.sz.ux_stride) || d2h(h_rb, d_rb, G.sz.pi_stride) || d2h(h_rd, d_rd, lamn) || d2h(&h_mu, d_mu, 1);
	cudaFree(d_rq); cudaFree(d_rb); cudaFree(d_rd); cudaFree(d_mu);
	if(bad || G.h_info[1]!=0.0)
		{
		free(h_rq); free(h_rb); free(h_rd);
		pthread_mutex_unlock(&g_lock);
		fprintf(stderr, "hpmpc_b200: %s: %s; outputs untouched\n", who, bad ? "GPU execution failed" : "the preceding IPM call ran no phase-2 iteration, there is no factor to reuse");
		return;
		}
	hpmpc_b200_unpack_solution(G.ocp, G.h_ux, G.h_pi, G.h_lam, x, u, pi, lam);
	for(n=0; n<N; n++)
		for(j=0; j<nb[n] && hidxb[n][j]<nu[n]; j++)
			if(lb[n][j]==ub[n][j]) u[n][hidxb[n][j]] = lb[n][j];
	if(inf_norm_res)
		{
		double m0 = 0.0, m1 = 0.0, m2 = 0.0;
		for(i=0; i<G.sz.ux_stride; i++) m0 = fmax(m0, fabs(h_rq[i]));
		for(i=0; i<G.sz.pi_stride; i++) m1 = fmax(m1, fabs(h_rb[i]));
		for(i=0; i<(int)lamn; i++) m2 = fmax(m2, fabs(h_rd[i]));
		inf_norm_res[0] = m0; inf_norm_res[1] = m1; inf_norm_res[2] = m2; inf_norm_res[3] = h_mu;
		}
	free(h_rq); free(h_rb); free(h_rd);
	pthread_mutex_unlock(&g_lock);
	}

void c_order_d_solve_kkt_new_rhs_ocp_hard_tv(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **A, double **B, double **b,
		double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug,
		double **x, double **u, double **pi, double **lam, double *inf_norm_res, double *work0)
	{
	high_level_new_rhs(1, N, nx, nu, nb, hidxb, ng, A, B, b, Q, S, R, q, r, lb, ub, C, D, lg, ug, x, u, pi, lam, inf_norm_res, work0, "c_order_d_solve_kkt_new_rhs_ocp_hard_tv");
	}

void fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, double **A, double **B, double **b,
		double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug,
		double **x, double **u, double **pi, double **lam, double *inf_norm_res, double *work0)
	{
	high_level_new_rhs(0, N, nx, nu, nb, hidxb, ng, A, B, b, Q, S, R, q, r, lb, ub, C, D, lg, ug, x, u, pi, lam, inf_norm_res, work0, "fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv");
	}
