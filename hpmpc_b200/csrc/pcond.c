/*
 * pcond.c -- host side of partial condensing: size bookkeeping, the pair of size patterns (full / condensed)
 * and the three-step batched solve  condense -> IPM on the condensed batch -> expand, all on one stream with no host round trip.
 *
 * Reference: lqcp_solvers/d_part_cond.c (d_part_cond_compute_problem_size :694, d_part_cond :926, d_part_expand_solution :1103)
 * and its caller interfaces/c/fortran_order_interface.c:389-528.  Like the reference this handles bounds only on the stages
 * before N (general constraints there make the reference exit, d_part_cond.c:962-968; stage N may have them); unlike the lib4 reference it is not limited
 * to nu <= 4 (the lib4 routine returns a non-stationary point for nu > 4 and blocks of 3 or more stages).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "layout.h"
#include "pcond_launch.h"
#include "../../include/hpmpc_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if(e_!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

const hb_dims *hpmpc_b200_internal_dims(const hpmpc_b200_ocp *p);
const hb_stage *hpmpc_b200_internal_stages(const hpmpc_b200_ocp *p);
int hpmpc_b200_internal_device(const hpmpc_b200_ocp *p, int *sms);

struct hpmpc_b200_pcond
	{
	int N, N2, device, sms;
	hpmpc_b200_ocp *full, *cond;
	hpmpc_b200_sizes szF, szC;
	hb_pc_block *h_blk, *d_blk;
	long long scratch_stride;
	int grid, warps;
	double *scratch;
	double *buf; size_t buf_bytes;          /* condensed batch + its solution + residual vectors of the fused call */
	};

static int block_len(int N, int N2, int k) { int N1 = N/N2, R1 = N - N2*N1; return k<R1 ? N1+1 : N1; }

/* d_part_cond_compute_problem_size (d_part_cond.c:694-735) plus the bound positions d_cond_DCtd assigns (:637-679).
 * nx2, nu2, nb2, ng2: [N2+1]; hidxb2 may be NULL, else hidxb2[k] must hold nb2[k] ints (upper bound: the block's bounds). */
int hpmpc_b200_part_cond_compute_problem_size(int N, const int *nx, const int *nu, const int *nb, int *const *hidxb, const int *ng, int N2,
		int *nx2, int *nu2, int *nb2, int *ng2, int **hidxb2)
	{
	int k, j, l, n0 = 0;
	if(N2<1 || N2>N) return -1;
	if(ng) for(k=0; k<N; k++) if(ng[k]>0) { fprintf(stderr, "hpmpc_b200: partial condensing handles bounds only before stage N (ng[n<N] = 0)\n"); return -5; }
	for(k=0; k<N2; k++)
		{
		const int T = block_len(N, N2, k);
		int nu_tmp = 0, ib = 0;
		nx2[k] = nx[n0]; nu2[k] = 0; ng2[k] = 0;
		for(j=0; j<T; j++) nu2[k] += nu[n0+j];
		for(j=T-1; j>=1; j--)
			{
			nu_tmp += nu[n0+j];
			for(l=0; l<(nb ? nb[n0+j] : 0); l++)
				{
				if(hidxb[n0+j][l]<nu[n0+j]) { if(hidxb2) hidxb2[k][ib] = nu_tmp - nu[n0+j] + hidxb[n0+j][l]; ib++; }
				else ng2[k]++;
				}
			}
		nu_tmp += nu[n0];
		for(l=0; l<(nb ? nb[n0] : 0); l++) { if(hidxb2) hidxb2[k][ib] = nu_tmp - nu[n0] + hidxb[n0][l]; ib++; }
		nb2[k] = ib;
		n0 += T;
		}
	nx2[N2] = nx[N]; nu2[N2] = 0; nb2[N2] = nb ? nb[N] : 0; ng2[N2] = ng ? ng[N] : 0;
	if(hidxb2) for(l=0; l<nb2[N2]; l++) hidxb2[N2][l] = hidxb[N][l];
	return 0;
	}

void hpmpc_b200_pcond_destroy(hpmpc_b200_pcond *h)
	{
	if(!h) return;
	if(h->device>=0) cudaSetDevice(h->device);
	if(h->full) hpmpc_b200_ocp_destroy(h->full);
	if(h->cond) hpmpc_b200_ocp_destroy(h->cond);
	if(h->d_blk) cudaFree(h->d_blk);
	if(h->scratch) cudaFree(h->scratch);
	if(h->buf) cudaFree(h->buf);
	free(h->h_blk);
	free(h);
	}

/* nu has N entries (nu[N] = 0 as in the high-level API); ng may be NULL, only ng[N] may be non-zero; 1 <= N2 < N */
int hpmpc_b200_pcond_create(hpmpc_b200_pcond **out, int N, const int *nx, const int *nu, const int *nb, int *const *hidxb, const int *ng,
		int N2, int device)
	{
	int k, n, rc = -1;
	*out = NULL;
	if(N<2 || N2<1 || N2>=N) { fprintf(stderr, "hpmpc_b200: partial condensing needs 1 <= N2 < N\n"); return -2; }
	hpmpc_b200_pcond *h = calloc(1, sizeof(*h));
	int *nuF = calloc(N+1, sizeof(int));
	int *nx2 = calloc(N2+1, sizeof(int)), *nu2 = calloc(N2+1, sizeof(int)), *nb2 = calloc(N2+1, sizeof(int)), *ng2 = calloc(N2+1, sizeof(int));
	int **idxb2 = calloc(N2+1, sizeof(int*));
	int nbt = 1;
	if(!h || !nuF || !nx2 || !nu2 || !nb2 || !ng2 || !idxb2) goto done;
	h->N = N; h->N2 = N2; h->device = device;
	for(n=0; n<N; n++) nuF[n] = nu[n];
	for(n=0; n<=N; n++) nbt += nb ? nb[n] : 0;
	for(k=0; k<=N2; k++) if(!(idxb2[k] = calloc(nbt, sizeof(int)))) goto done;
	if((rc = hpmpc_b200_part_cond_compute_problem_size(N, nx, nuF, nb, hidxb, ng, N2, nx2, nu2, nb2, ng2, idxb2))) goto done;
	rc = -1;
	for(k=0; k<=N2; k++)
		if(nx2[k]+nu2[k]+1>64)
			{
			fprintf(stderr, "hpmpc_b200: partial condensing: block %d has nu+nx+1 = %d > 64 (the any-size kernels' limit); use a larger N2\n", k, nx2[k]+nu2[k]+1);
			rc = -3; goto done;
			}
	if(hpmpc_b200_ocp_create_gen(&h->full, N, nx, nu, nb, hidxb, ng, device)) goto done;
	if(hpmpc_b200_ocp_create_gen(&h->cond, N2, nx2, nu2, nb2, idxb2, ng2, device)) goto done;
	hpmpc_b200_ocp_sizes(h->full, &h->szF); hpmpc_b200_ocp_sizes(h->cond, &h->szC);
	h->h_blk = calloc(N2, sizeof(hb_pc_block));
	if(!h->h_blk) goto done;
	for(k=0, n=0; k<N2; k++) { h->h_blk[k].n0 = n; h->h_blk[k].T = block_len(N, N2, k); n += h->h_blk[k].T; }
	h->scratch_stride = hb_pcond_scratch_doubles(hpmpc_b200_internal_stages(h->full), N, h->h_blk, N2);
	if(device>=0)
		{
		if(cudaSetDevice(device)!=cudaSuccess) goto done;
		hpmpc_b200_internal_device(h->full, &h->sms);
		h->warps = 4; h->grid = 8*h->sms;
		if(cudaMalloc((void**)&h->d_blk, sizeof(hb_pc_block)*N2)!=cudaSuccess
		|| cudaMemcpy(h->d_blk, h->h_blk, sizeof(hb_pc_block)*N2, cudaMemcpyHostToDevice)!=cudaSuccess
		|| cudaMalloc((void**)&h->scratch, sizeof(double)*(size_t)h->grid*h->warps*h->scratch_stride)!=cudaSuccess)
			{ fprintf(stderr, "hpmpc_b200: partial condensing: device allocation failed\n"); goto done; }
		}
	rc = 0;
done:
	if(idxb2) for(k=0; k<=N2; k++) free(idxb2[k]);
	free(idxb2); free(nuF); free(nx2); free(nu2); free(nb2); free(ng2);
	if(rc) { hpmpc_b200_pcond_destroy(h); return rc; }
	*out = h;
	return 0;
	}

hpmpc_b200_ocp *hpmpc_b200_pcond_full(hpmpc_b200_pcond *h) { return h->full; }
hpmpc_b200_ocp *hpmpc_b200_pcond_cond(hpmpc_b200_pcond *h) { return h->cond; }

static int grid_for(const hpmpc_b200_pcond *h, long long items)
	{
	long long need = (items + h->warps - 1)/h->warps;
	return (int)(need<h->grid ? (need<1 ? 1 : need) : h->grid);
	}

/* d_part_cond for a batch: d_in_full [n_inst x szF.in_stride] -> d_in_cond [n_inst x szC.in_stride] */
int hpmpc_b200_d_part_cond_batch(hpmpc_b200_pcond *h, long long n_inst, const double *d_in_full, double *d_in_cond, void *stream)
	{
	if(n_inst<=0) return 0;
	if(h->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot condense; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(h->device));
	return hb_launch_pcond(hpmpc_b200_internal_dims(h->full), hpmpc_b200_internal_dims(h->cond), h->d_blk, h->N2, n_inst, d_in_full, d_in_cond,
			h->scratch, h->scratch_stride, grid_for(h, n_inst*h->N2), h->warps, stream);
	}

/* d_part_expand_solution for a batch: (ux2, pi2, lam2, t2) in the condensed handle's output layouts -> the full handle's */
int hpmpc_b200_d_part_expand_solution_batch(hpmpc_b200_pcond *h, long long n_inst, const double *d_in_full, const double *d_ux2,
		const double *d_pi2, const double *d_lam2, const double *d_t2, double *d_ux, double *d_pi, double *d_lam, double *d_t, void *stream)
	{
	if(n_inst<=0) return 0;
	if(h->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot expand; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(h->device));
	return hb_launch_pexpand(hpmpc_b200_internal_dims(h->full), hpmpc_b200_internal_dims(h->cond), h->d_blk, h->N2, n_inst, d_in_full,
			d_ux2, d_pi2, d_lam2, d_t2, d_ux, d_pi, d_lam, d_t, grid_for(h, n_inst*h->N2), h->warps, stream);
	}

/* the N2 < N branch of {c,fortran}_order_d_ip_ocp_hard_tv for a batch (interfaces/c/fortran_order_interface.c:389-528, :616-656):
 * condense, solve the condensed batch with the IPM (cold start, as the reference does), expand, and put the exit norms of the FULL
 * problem into info[2..5].  Outputs in the full handle's layouts; d_info [n_inst x (6+5*k_max)]: kk, status and the stat table are
 * those of the condensed solve. */
int hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(hpmpc_b200_pcond *h, long long n_inst, const double *d_in_full, int k_max, double mu0,
		double mu_tol, double alpha_min, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, void *stream)
	{
	if(n_inst<=0) return 0;
	if(h->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(h->device));
	const long long lamC = h->szC.lam_stride>0 ? h->szC.lam_stride : 2, lamF = h->szF.lam_stride>0 ? h->szF.lam_stride : 2;
	const size_t per = (size_t)h->szC.in_stride + h->szC.ux_stride + h->szC.pi_stride + 2*lamC
		+ h->szF.ux_stride + h->szF.pi_stride + lamF + 2;
	const size_t need = sizeof(double)*per*(size_t)n_inst;
	if(need>h->buf_bytes)
		{
		/* the previous call's work may still be using the old buffer */
		CK(cudaDeviceSynchronize());
		if(h->buf) CK(cudaFree(h->buf));
		h->buf = NULL; h->buf_bytes = 0;
		CK(cudaMalloc((void**)&h->buf, need));
		h->buf_bytes = need;
		}
	double *in2 = h->buf, *ux2 = in2 + (size_t)n_inst*h->szC.in_stride, *pi2 = ux2 + (size_t)n_inst*h->szC.ux_stride;
	double *lam2 = pi2 + (size_t)n_inst*h->szC.pi_stride, *t2 = lam2 + (size_t)n_inst*lamC;
	double *rq = t2 + (size_t)n_inst*lamC, *rb = rq + (size_t)n_inst*h->szF.ux_stride, *rd = rb + (size_t)n_inst*h->szF.pi_stride;
	double *mu = rd + (size_t)n_inst*lamF;
	int rc;
	if((rc = hpmpc_b200_d_part_cond_batch(h, n_inst, d_in_full, in2, stream))) return rc;
	if((rc = hpmpc_b200_d_ip2_res_mpc_hard_batch(h->cond, n_inst, in2, k_max, mu0, mu_tol, alpha_min, 0, ux2, pi2, lam2, t2, d_info, stream))) return rc;
	if((rc = hpmpc_b200_d_part_expand_solution_batch(h, n_inst, d_in_full, ux2, pi2, lam2, t2, d_ux, d_pi, d_lam, d_t, stream))) return rc;
	if((rc = hpmpc_b200_d_res_res_mpc_hard_batch(h->full, n_inst, d_in_full, d_ux, d_pi, d_lam, d_t, rq, rb, rd, NULL, mu, stream))) return rc;
	return hb_launch_res_norms(hpmpc_b200_internal_dims(h->full), n_inst, rq, rb, rd, mu, h->szF.lam_stride, d_info,
			HB_IPM_INFO_HEAD + 5*(long long)k_max, stream);
	}
