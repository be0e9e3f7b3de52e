/*
 * ocp.c -- host side (plain C) of the batched engine: size pattern -> device layout, packing from the
 * reference's stage-wise arrays, scratch-slot management, the device- and host-buffer entry points of
 * include/hpmpc_b200.h.  No solver arithmetic happens here; it all runs in ric_kernels.cu.
 *
 * Packing restates what the reference's wrapper does before calling its solver
 * (interfaces/c/fortran_order_interface.c:262-380 / c_order_interface.c:262-380): B', A', b into one
 * [B A b]' block, [R S'; S Q] and [r q] into one Hessian block, [lb ub] into one bound vector -- but
 * into the packed-trapezoid layout of layout.h instead of panel-major.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "layout.h"
#include "../../include/hpmpc_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if(e_!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

struct hpmpc_b200_ocp
	{
	int device;
	int N;
	int *nx, *nu, *nb, *ng;  /* [N+1] */
	int *r_nx, *r_nu;        /* [N+1] the caller's sizes when the pattern is embedded in a larger compiled shape (hpmpc_b200_ocp_create_padded), else NULL */
	int **idxb;              /* [N+1] */
	hb_stage *st;            /* host copy [N+1] */
	int *h_idxb, *h_cux;     /* flat [nbtot] */
	hb_dims dims;            /* st / idxb / c_ux are DEVICE pointers */
	long long lam_stride;
	/* launch shape: generic kernels */
	int sms, grid, warps, n_slots, smem_cta;
	/* size-specialised sv kernel (ric_fast.cuh), -1 when the pattern has no compiled variant */
	int fast_id, f_ipw, f_smem_warp, f_grid, f_warps, f_smem_cta;
	long long f_stash_inst;
	/* size-specialised IPM sweeps (ric_ipm_fast.cuh), -1 when none */
	int ipm_fast_id, i_smem_warp, i_grid, i_warps;
	int tf_id, tf_smem_warp, tf_grid, tf_warps;   /* size-specialised trf / trs (same sweeps as the fast IPM), -1 when the shape has none */
	long long tf_L_stride;                        /* their factor: column-packed, (N+1)*LBUF doubles per instance (>= dims.L_stride) */
	long long i_L_doubles, ipm_ws;
	/* scratch (device), grown on demand */
	double *scratch; size_t scratch_bytes;
	int *counter;
	/* one call in flight per handle: every device entry point makes its stream wait for the previous call's work (the calls share
	 * the scratch slots and the queue counter), so calls on different streams are serialised instead of racing */
	cudaEvent_t ev_busy; int busy_valid;
	/* staging for the host-buffer entry points */
	cudaStream_t s_copy[2], s_comp;
	cudaEvent_t ev_in[2], ev_done[2];
	double *stage_in[2], *stage_out[2]; size_t stage_in_bytes, stage_out_bytes;
	cudaEvent_t ev[2];
	};

static int ensure_scratch(hpmpc_b200_ocp *p, size_t bytes)
	{
	if(bytes<=p->scratch_bytes) return 0;
	if(p->scratch) CK(cudaFree(p->scratch));
	p->scratch = NULL; p->scratch_bytes = 0;
	CK(cudaMalloc((void**)&p->scratch, bytes));
	p->scratch_bytes = bytes;
	return 0;
	}

static void default_launch(hpmpc_b200_ocp *p, int ctas_per_sm, int warps)
	{
	int smem_warp = hb_smem_bytes_per_warp(&p->dims);
	if(warps<=0 && ctas_per_sm<=0)
		{
		/* as many resident warps per SM as shared memory allows (<= 16): 4-warp CTAs unless another CTA size fits more warps
		 * (large stages: e.g. 38.8 KB per warp at BASELINE config 4 -> one 5-warp CTA instead of one 4-warp CTA) */
		static const int cand[8] = { 4, 5, 6, 7, 8, 3, 2, 1 };
		int best = 0, k;
		for(k=0; k<8; k++)
			{
			int w = cand[k], c = (227*1024)/(w*smem_warp+1024);
			if(c>16/w) c = 16/w;
			if(c>=1 && w*c>best) { best = w*c; warps = w; ctas_per_sm = c; }
			}
		if(best==0) { warps = 1; ctas_per_sm = 1; }
		}
	if(warps<=0)
		{
		warps = 4;
		while(warps>1 && warps*smem_warp>200*1024) warps--;
		}
	int smem_cta = warps*smem_warp;
	if(ctas_per_sm<=0)
		{
		ctas_per_sm = (220*1024)/(smem_cta+1024);
		if(ctas_per_sm<1) ctas_per_sm = 1;
		int cap = 16/warps; if(cap<1) cap = 1;          /* <= 16 resident warps per SM by default */
		if(ctas_per_sm>cap) ctas_per_sm = cap;
		}
	p->warps = warps;
	p->grid = p->sms*ctas_per_sm;
	p->n_slots = p->grid*p->warps;
	p->smem_cta = smem_cta;
	}

static void fast_launch(hpmpc_b200_ocp *p, int ctas_per_sm, int warps)
	{
	if(p->fast_id<0) return;
	if(warps<=0) warps = 8;
	while(warps>1 && warps*p->f_smem_warp>113*1024) warps--;
	int smem_cta = warps*p->f_smem_warp;
	if(ctas_per_sm<=0)
		{
		ctas_per_sm = (228*1024)/(smem_cta+1024);
		if(ctas_per_sm<1) ctas_per_sm = 1;
		int cap = 16/warps; if(cap<1) cap = 1;           /* <= 16 resident warps per SM (128 registers per thread) */
		if(ctas_per_sm>cap) ctas_per_sm = cap;
		}
	p->f_warps = warps; p->f_grid = p->sms*ctas_per_sm; p->f_smem_cta = smem_cta;
	}

/* launch shape of the IPM kernel: the generic shape, or for the size-specialised sweeps as many 4-warp CTAs as fit */
static void ipm_launch(hpmpc_b200_ocp *p)
	{
	p->i_warps = p->warps; p->i_grid = p->grid;
	if(p->ipm_fast_id<0) return;
	int warps = 4;
	while(warps>1 && warps*p->i_smem_warp>113*1024) warps--;
	int per_sm = (228*1024)/(warps*p->i_smem_warp+1024);
	if(per_sm<1) per_sm = 1;
	if(per_sm*warps>16) per_sm = 16/warps;
	p->i_warps = warps; p->i_grid = p->sms*per_sm;
	}

static void trf_launch(hpmpc_b200_ocp *p)
	{
	if(p->tf_id<0) return;
	int warps = 4;
	while(warps>1 && warps*p->tf_smem_warp>113*1024) warps--;
	int per_sm = (228*1024)/(warps*p->tf_smem_warp+1024);
	if(per_sm<1) per_sm = 1;
	if(per_sm*warps>8) per_sm = 8/warps>0 ? 8/warps : 1;     /* ~250 registers per thread */
	p->tf_warps = warps; p->tf_grid = p->sms*per_sm;
	}

static void free_host_side(hpmpc_b200_ocp *p)
	{
	int n;
	if(p->idxb) for(n=0; n<=p->N; n++) free(p->idxb[n]);
	free(p->idxb); free(p->nx); free(p->nu); free(p->nb); free(p->ng); free(p->st); free(p->h_idxb); free(p->h_cux); free(p->r_nx); free(p->r_nu);
	}

/* all stage matrices first, then the constraint data of all stages: the distance between the matrices of consecutive stages
 * does not depend on nb / ng, which is what the size-specialised kernels' affine stage addressing relies on */
int hpmpc_b200_ocp_create_gen(hpmpc_b200_ocp **out, int N, const int *nx, const int *nu, const int *nb, int *const *hidxb,
		const int *ng, int device)
	{
	int n, j, rc = -2;
	*out = NULL;
	if(N<1) { fprintf(stderr, "hpmpc_b200: N must be >= 1\n"); return -2; }
	hpmpc_b200_ocp *p = (hpmpc_b200_ocp*)calloc(1, sizeof(*p));
	if(!p) return -1;
	p->device = -1;                      /* until the device side exists: the error path frees host memory only */
	p->N = N;
	p->nx = malloc((N+1)*sizeof(int)); p->nu = malloc((N+1)*sizeof(int)); p->nb = malloc((N+1)*sizeof(int)); p->ng = malloc((N+1)*sizeof(int));
	p->idxb = calloc(N+1, sizeof(int*));
	p->st = calloc(N+1, sizeof(hb_stage));
	if(!p->nx || !p->nu || !p->nb || !p->ng || !p->idxb || !p->st) { rc = -1; goto fail; }
	int nctot = 0, ngtot = 0, nzM = 1, nxM = 1;
	for(n=0; n<=N; n++)
		{
		p->nx[n] = nx[n]; p->nu[n] = n<N ? nu[n] : 0; p->nb[n] = nb ? nb[n] : 0; p->ng[n] = ng ? ng[n] : 0;
		if(p->nx[n]<0 || p->nu[n]<0 || p->nb[n]<0 || p->ng[n]<0) { fprintf(stderr, "hpmpc_b200: stage %d: negative size\n", n); goto fail; }
		if(p->nb[n]>p->nu[n]+p->nx[n])
			{
			/* the reference prints this and exit(1)s (c_order_interface.c:103-110); we return an error */
			fprintf(stderr, "hpmpc_b200: stage %d: nb=%d larger than nu+nx=%d\n", n, p->nb[n], p->nu[n]+p->nx[n]);
			goto fail;
			}
		if(p->nb[n]>0 && (hidxb==NULL || hidxb[n]==NULL)) { fprintf(stderr, "hpmpc_b200: stage %d: nb=%d but no idxb\n", n, p->nb[n]); goto fail; }
		nctot += p->nb[n]+p->ng[n]; ngtot += p->ng[n];
		if(p->nu[n]+p->nx[n]+1>nzM) nzM = p->nu[n]+p->nx[n]+1;
		if(p->nx[n]>nxM) nxM = p->nx[n];
		}
	p->h_idxb = malloc((nctot+1)*sizeof(int)); p->h_cux = malloc((nctot+1)*sizeof(int));
	if(!p->h_idxb || !p->h_cux) { rc = -1; goto fail; }
	long long o_in = 0, o_ux = 0, o_pi = 0, o_L = 0; int o_c = 0;
	for(n=0; n<=N; n++)
		{
		hb_stage *s = &p->st[n];
		int nux = p->nu[n]+p->nx[n];
		s->nx = p->nx[n]; s->nu = p->nu[n]; s->nb = p->nb[n]; s->ng = p->ng[n]; s->nx1 = n<N ? p->nx[n+1] : 0;
		s->off_BAbt = (int)o_in; o_in += HB_EVEN((nux+1)*s->nx1);
		s->off_RSQ = (int)o_in;  o_in += HB_EVEN(HB_TRI(nux)+nux);
		s->off_ux = (int)o_ux;   o_ux += nux;
		s->off_pi = (int)o_pi;   o_pi += s->nx1;
		s->off_c = o_c;
		s->off_L = (int)o_L;     o_L += HB_EVEN(HB_TRI(nux)+2*nux);
		p->idxb[n] = malloc((s->nb+1)*sizeof(int));
		if(!p->idxb[n]) { rc = -1; goto fail; }
		for(j=0; j<s->nb; j++)
			{
			int id = hidxb[n][j];
			if(id<0 || id>=nux) { fprintf(stderr, "hpmpc_b200: stage %d: idxb[%d]=%d out of range\n", n, j, id); goto fail; }
			p->idxb[n][j] = id;
			p->h_idxb[o_c+j] = id;
			p->h_cux[o_c+j] = s->off_ux + id;
			}
		for(j=0; j<s->ng; j++) { p->h_idxb[o_c+s->nb+j] = -1; p->h_cux[o_c+s->nb+j] = -1; }     /* general: value = (DCt' ux)_j */
		o_c += s->nb+s->ng;
		}
	for(n=0; n<=N; n++)
		{
		hb_stage *s = &p->st[n];
		int nux = p->nu[n]+p->nx[n];
		s->off_d = (int)o_in;    o_in += HB_EVEN(2*s->nb);
		s->off_DCt = (int)o_in;  o_in += HB_EVEN(nux*s->ng);
		s->off_dg = (int)o_in;   o_in += HB_EVEN(2*s->ng);
		}
	p->dims.N = N; p->dims.nzM = nzM; p->dims.nxM = nxM; p->dims.nbtot = nctot; p->dims.ngtot = ngtot;
	p->dims.in_stride = o_in; p->dims.ux_stride = HB_EVEN(o_ux); p->dims.pi_stride = HB_EVEN(o_pi); p->dims.L_stride = o_L;
	p->lam_stride = 2*(long long)nctot;
	if(nzM>64) { fprintf(stderr, "hpmpc_b200: nu+nx+1 = %d > 64 is not supported yet\n", nzM); goto fail; }

	/* size-specialised kernels: chosen from (N, nx, nu) and only when the stage matrices sit at an affine stride (true by
	 * construction of the layout above; checked because the kernels compute off(n) = off(1) + (n-1)*stride) and no general
	 * constraints are present */
	int affine = 1;
	for(n=2; n<N; n++) if(p->st[n].off_BAbt != p->st[1].off_BAbt + (n-1)*(p->st[2].off_BAbt-p->st[1].off_BAbt)) affine = 0;
	const int fast_ok = affine && ngtot==0;
	p->fast_id = (getenv("HPMPC_B200_NO_FAST") || !fast_ok) ? -1 : hb_fast_variant(N, p->nx, p->nu);
	if(p->fast_id>=0) hb_fast_info(p->fast_id, N, &p->f_ipw, &p->f_smem_warp, &p->f_stash_inst);
	p->ipm_fast_id = fast_ok ? hb_ipm_fast_variant(N, p->nx, p->nu, nctot) : -1;
	p->i_L_doubles = p->dims.L_stride;
	if(p->ipm_fast_id>=0) hb_ipm_fast_info(p->ipm_fast_id, N, &p->i_smem_warp, &p->i_L_doubles);
	p->tf_id = fast_ok ? hb_ric_shape_variant(N, p->nx, p->nu) : -1;
	if(p->tf_id>=0)
		{
		/* the factor of trf / trs is kept in the sweeps' column-packed form: the L_stride reported to callers makes room for it */
		long long Ld = 0;
		hb_ipm_fast_info(p->tf_id, N, &p->tf_smem_warp, &Ld);
		p->tf_L_stride = Ld>p->dims.L_stride ? Ld : p->dims.L_stride;
		}
	p->ipm_ws = hb_ipm_work_doubles2(&p->dims, p->i_L_doubles);
	if(device<0)
		{
		/* host-only handle: layout queries and packing work, every compute entry point refuses to run */
		p->sms = 148;
		default_launch(p, 0, 0);
		fast_launch(p, 0, 0);
		ipm_launch(p);
		trf_launch(p);
		*out = p;
		return 0;
		}
	rc = -1;
	if(cudaSetDevice(device)!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: no CUDA device %d\n", device); goto fail; }
	p->sms = hb_device_sm_count(device);
	if(p->sms<=0) { fprintf(stderr, "hpmpc_b200: no CUDA device %d\n", device); goto fail; }
	p->device = device;                  /* from here on the error path is hpmpc_b200_ocp_destroy (frees whatever exists) */
	{
	hb_stage *d_st = NULL; int *d_idxb = NULL, *d_cux = NULL;
	if(cudaMalloc((void**)&d_st, (N+1)*sizeof(hb_stage))!=cudaSuccess) goto fail_dev;
	p->dims.st = d_st;
	if(cudaMalloc((void**)&d_idxb, (nctot+1)*sizeof(int))!=cudaSuccess) goto fail_dev;
	p->dims.idxb = d_idxb;
	if(cudaMalloc((void**)&d_cux, (nctot+1)*sizeof(int))!=cudaSuccess) goto fail_dev;
	p->dims.c_ux = d_cux;
	if(cudaMemcpy(d_st, p->st, (N+1)*sizeof(hb_stage), cudaMemcpyHostToDevice)!=cudaSuccess
	|| cudaMemcpy(d_idxb, p->h_idxb, (nctot+1)*sizeof(int), cudaMemcpyHostToDevice)!=cudaSuccess
	|| cudaMemcpy(d_cux, p->h_cux, (nctot+1)*sizeof(int), cudaMemcpyHostToDevice)!=cudaSuccess
	|| cudaMalloc((void**)&p->counter, 64)!=cudaSuccess
	|| cudaEventCreateWithFlags(&p->ev_busy, cudaEventDisableTiming)!=cudaSuccess) goto fail_dev;
	}
	default_launch(p, 0, 0);
	fast_launch(p, 0, 0);
	ipm_launch(p);
	trf_launch(p);
	*out = p;
	return 0;
fail_dev:
	fprintf(stderr, "hpmpc_b200: device allocation failed (%s)\n", cudaGetErrorString(cudaGetLastError()));
	hpmpc_b200_ocp_destroy(p);
	return -1;
fail:
	free_host_side(p);
	free(p);
	return rc;
	}

int hpmpc_b200_ocp_create(hpmpc_b200_ocp **out, int N, const int *nx, const int *nu, const int *nb, int *const *hidxb, int device)
	{
	return hpmpc_b200_ocp_create_gen(out, N, nx, nu, nb, hidxb, NULL, device);
	}

/* A uniform pattern (nx[0] = 0, nx[1..N] = nx, nu[0..N-1] = nu, bounds only) that has no size-specialised kernels of its own is
 * EMBEDDED in the smallest compiled shape (NX, NU) >= (nx, nu): the missing inputs and states become decoupled dummies (unit
 * cost, zero dynamics, no bounds), which leaves every real quantity of the recursion unchanged -- the same device the kernels use
 * for stage 0 (no x) and stage N (no u).  The handle then describes the PADDED frame: strides, stage offsets, ux = [u (NU) x (NX)]
 * per stage with the real entries first in each part, idxb shifted accordingly; hpmpc_b200_pack_instance / _unpack_solution take
 * and return the caller's real sizes.  A pattern that is not uniform, already compiled, or larger than every compiled shape gets a
 * plain handle (any-size kernels), as from hpmpc_b200_ocp_create.  Addresses the fast path's shape list. */
static const int pad_shapes[][2] = { {4, 2}, {8, 3}, {12, 5}, {24, 11} };
int hpmpc_b200_ocp_create_padded(hpmpc_b200_ocp **out, int N, const int *nx, const int *nu, const int *nb, int *const *hidxb, int device)
	{
	int n, j, k, uniform = (N>=3 && nx[0]==0), nbt = 0;
	for(n=1; uniform && n<=N; n++) if(nx[n]!=nx[1]) uniform = 0;
	for(n=0; uniform && n<N; n++) if(nu[n]!=nu[0]) uniform = 0;
	for(n=0; nb && n<=N; n++) nbt += nb[n];
	int NX = -1, NU = -1;
	if(uniform)
		for(k=0; k<4; k++)
			{
			if(nbt>0 && k==0) continue;                      /* the smallest shape has no size-specialised IPM sweeps */
			if(nx[1]<=pad_shapes[k][0] && nu[0]<=pad_shapes[k][1]) { NX = pad_shapes[k][0]; NU = pad_shapes[k][1]; break; }
			}
	if(NX<0 || (NX==nx[1] && NU==nu[0])) return hpmpc_b200_ocp_create_gen(out, N, nx, nu, nb, hidxb, NULL, device);
	int *nxp = malloc((N+1)*sizeof(int)), *nup = malloc((N+1)*sizeof(int));
	int **idp = calloc(N+1, sizeof(int*));
	int rc = -1;
	if(nxp && nup && idp)
		{
		rc = 0;
		for(n=0; n<=N; n++)
			{
			nxp[n] = n==0 ? 0 : NX; nup[n] = n<N ? NU : 0;
			const int nbn = nb ? nb[n] : 0, nun = n<N ? nu[n] : 0;
			idp[n] = malloc((nbn+1)*sizeof(int));
			if(!idp[n]) { rc = -1; break; }
			for(j=0; j<nbn; j++)
				{
				const int id = hidxb[n][j];
				if(id<0 || id>=nun+nx[n]) { fprintf(stderr, "hpmpc_b200: stage %d: idxb[%d]=%d out of range\n", n, j, id); rc = -2; break; }
				idp[n][j] = id<nun ? id : nup[n] + (id-nun);
				}
			if(rc) break;
			}
		if(rc==0) rc = hpmpc_b200_ocp_create_gen(out, N, nxp, nup, nb, idp, NULL, device);
		if(rc==0)
			{
			hpmpc_b200_ocp *p = *out;
			p->r_nx = malloc((N+1)*sizeof(int)); p->r_nu = malloc((N+1)*sizeof(int));
			if(!p->r_nx || !p->r_nu) { hpmpc_b200_ocp_destroy(p); *out = NULL; rc = -1; }
			else for(n=0; n<=N; n++) { p->r_nx[n] = nx[n]; p->r_nu[n] = n<N ? nu[n] : 0; }
			}
		}
	if(idp) for(n=0; n<=N; n++) free(idp[n]);
	free(idp); free(nxp); free(nup);
	return rc;
	}

/* 1 and the compiled shape the pattern is embedded in when the handle came from hpmpc_b200_ocp_create_padded and was padded */
int hpmpc_b200_ocp_padded_shape(const hpmpc_b200_ocp *p, int *NX, int *NU)
	{
	if(!p->r_nx) return 0;
	if(NX) *NX = p->nx[p->N];
	if(NU) *NU = p->nu[0];
	return 1;
	}

void hpmpc_b200_ocp_destroy(hpmpc_b200_ocp *p)
	{
	int k;
	if(!p) return;
	if(p->device<0)
		{
		free_host_side(p);
		free(p);
		return;
		}
	cudaSetDevice(p->device);
	for(k=0; k<2; k++)
		{
		if(p->stage_in[k]) cudaFree(p->stage_in[k]);
		if(p->stage_out[k]) cudaFree(p->stage_out[k]);
		if(p->s_copy[k]) cudaStreamDestroy(p->s_copy[k]);
		if(p->ev[k]) cudaEventDestroy(p->ev[k]);
		if(p->ev_in[k]) cudaEventDestroy(p->ev_in[k]);
		if(p->ev_done[k]) cudaEventDestroy(p->ev_done[k]);
		}
	if(p->s_comp) cudaStreamDestroy(p->s_comp);
	if(p->scratch) cudaFree(p->scratch);
	if(p->counter) cudaFree(p->counter);
	if(p->ev_busy) cudaEventDestroy(p->ev_busy);
	if(p->dims.st) cudaFree((void*)p->dims.st);
	if(p->dims.idxb) cudaFree((void*)p->dims.idxb);
	if(p->dims.c_ux) cudaFree((void*)p->dims.c_ux);
	free_host_side(p);
	free(p);
	}

int hpmpc_b200_ocp_set_launch(hpmpc_b200_ocp *p, int ctas_per_sm, int warps_per_cta)
	{
	/* every kernel of this library is compiled with __launch_bounds__(256): at most 8 warps per CTA */
	if(warps_per_cta>8)
		{
		fprintf(stderr, "hpmpc_b200: set_launch: %d warps per CTA requested, the kernels take at most 8 -- rejected\n", warps_per_cta);
		return -3;
		}
	if(p->fast_id>=0)
		{
		/* the launch shape of the size-specialised kernel is what matters for this pattern */
		fast_launch(p, ctas_per_sm, warps_per_cta);
		return 0;
		}
	default_launch(p, ctas_per_sm, warps_per_cta);
	ipm_launch(p);
	if(p->smem_cta>227*1024) { fprintf(stderr, "hpmpc_b200: %d warps need %d bytes of shared memory\n", p->warps, p->smem_cta); default_launch(p, 0, 0); return -2; }
	return 0;
	}

void hpmpc_b200_ocp_sizes(const hpmpc_b200_ocp *p, hpmpc_b200_sizes *o)
	{
	o->in_stride = p->dims.in_stride; o->ux_stride = p->dims.ux_stride; o->pi_stride = p->dims.pi_stride;
	o->lam_stride = p->lam_stride; o->L_stride = p->tf_id>=0 ? p->tf_L_stride : p->dims.L_stride; o->ipm_work_stride = p->ipm_ws;
	o->N = p->N; o->nzM = p->dims.nzM; o->nxM = p->dims.nxM; o->nbtot = p->dims.nbtot;
	o->grid = p->grid; o->warps_per_cta = p->warps; o->n_slots = p->n_slots; o->smem_per_cta = p->smem_cta;
	o->fast_variant = p->fast_id;
	o->ipm_grid = p->i_grid; o->ipm_warps_per_cta = p->i_warps; o->ipm_fast_variant = p->ipm_fast_id;
	if(p->fast_id>=0)
		{
		o->grid = p->f_grid; o->warps_per_cta = p->f_warps; o->n_slots = p->f_grid*p->f_warps*p->f_ipw; o->smem_per_cta = p->f_smem_cta;
		}
	}

void hpmpc_b200_ocp_stage_offsets(const hpmpc_b200_ocp *p, int n, int *off_BAbt, int *off_RSQ, int *off_d,
		int *off_ux, int *off_pi, int *off_lam, int *off_L)
	{
	const hb_stage *s = &p->st[n];
	if(off_BAbt) *off_BAbt = s->off_BAbt;
	if(off_RSQ) *off_RSQ = s->off_RSQ;
	if(off_d) *off_d = s->off_d;
	if(off_ux) *off_ux = s->off_ux;
	if(off_pi) *off_pi = s->off_pi;
	if(off_lam) *off_lam = 2*s->off_c;
	if(off_L) *off_L = s->off_L;
	}

void hpmpc_b200_ocp_general_offsets(const hpmpc_b200_ocp *p, int n, int *ng, int *off_DCt, int *off_dg, int *off_c)
	{
	const hb_stage *s = &p->st[n];
	if(ng) *ng = s->ng;
	if(off_DCt) *off_DCt = s->off_DCt;
	if(off_dg) *off_dg = s->off_dg;
	if(off_c) *off_c = s->off_c;
	}

/* element (i,j) of an m x n matrix with leading dimension ld given in column- or row-major order */
#define EL(M, i, j, rows, cols, c_order) ((c_order) ? (M)[(size_t)(i)*(cols)+(j)] : (M)[(i)+(size_t)(j)*(rows)])

int hpmpc_b200_pack_instance(const hpmpc_b200_ocp *p, int c_order, double *const *A, double *const *B, double *const *b,
		double *const *Q, double *const *S, double *const *R, double *const *q, double *const *r,
		double *const *lb, double *const *ub, double *blk)
	{
	int n, i, j;
	memset(blk, 0, sizeof(double)*p->dims.in_stride);
	if(p->r_nx)
		{
		/* embedded pattern: real entries into the padded frame, unit cost on the dummy inputs and states */
		for(n=0; n<=p->N; n++)
			{
			const hb_stage *s = &p->st[n];
			const int NXp = s->nx, NUp = s->nu, nuxp = NXp+NUp, nx1p = s->nx1;
			const int nx = p->r_nx[n], nu = p->r_nu[n], nx1 = n<p->N ? p->r_nx[n+1] : 0;
			if(n<p->N)
				{
				double *M = blk + s->off_BAbt;
				for(i=0; i<nu; i++) for(j=0; j<nx1; j++) M[i*nx1p+j] = EL(B[n], j, i, nx1, nu, c_order);
				for(i=0; i<nx; i++) for(j=0; j<nx1; j++) M[(NUp+i)*nx1p+j] = EL(A[n], j, i, nx1, nx, c_order);
				for(j=0; j<nx1; j++) M[nuxp*nx1p+j] = b[n][j];
				}
			double *H = blk + s->off_RSQ;
			for(i=0; i<nu; i++) for(j=0; j<=i; j++) H[HB_TRI(i)+j] = EL(R[n], i, j, nu, nu, c_order);
			for(i=nu; i<NUp; i++) H[HB_TRI(i)+i] = 1.0;
			for(i=0; i<nx; i++)
				{
				for(j=0; j<nu; j++) H[HB_TRI(NUp+i)+j] = EL(S[n], j, i, nu, nx, c_order);
				for(j=0; j<=i; j++) H[HB_TRI(NUp+i)+NUp+j] = EL(Q[n], i, j, nx, nx, c_order);
				}
			for(i=nx; i<NXp; i++) H[HB_TRI(NUp+i)+NUp+i] = 1.0;
			for(j=0; j<nu; j++) H[HB_TRI(nuxp)+j] = r[n][j];
			for(j=0; j<nx; j++) H[HB_TRI(nuxp)+NUp+j] = q[n][j];
			double *d = blk + s->off_d;
			for(j=0; j<s->nb; j++) { d[j] = lb[n][j]; d[s->nb+j] = ub[n][j]; }
			}
		return 0;
		}
	for(n=0; n<=p->N; n++)
		{
		const hb_stage *s = &p->st[n];
		int nx = s->nx, nu = s->nu, nux = nx+nu, nx1 = s->nx1;
		if(n<p->N)
			{
			double *M = blk + s->off_BAbt;                     /* (nux+1) x nx1 row-major */
			for(i=0; i<nu; i++) for(j=0; j<nx1; j++) M[i*nx1+j] = EL(B[n], j, i, nx1, nu, c_order);
			for(i=0; i<nx; i++) for(j=0; j<nx1; j++) M[(nu+i)*nx1+j] = EL(A[n], j, i, nx1, nx, c_order);
			for(j=0; j<nx1; j++) M[nux*nx1+j] = b[n][j];
			}
		double *H = blk + s->off_RSQ;                          /* packed lower trapezoid */
		for(i=0; i<nu; i++) for(j=0; j<=i; j++) H[HB_TRI(i)+j] = EL(R[n], i, j, nu, nu, c_order);
		for(i=0; i<nx; i++)
			{
			for(j=0; j<nu; j++) H[HB_TRI(nu+i)+j] = EL(S[n], j, i, nu, nx, c_order);      /* S is nu x nx */
			for(j=0; j<=i; j++) H[HB_TRI(nu+i)+nu+j] = EL(Q[n], i, j, nx, nx, c_order);
			}
		for(j=0; j<nu; j++) H[HB_TRI(nux)+j] = r[n][j];
		for(j=0; j<nx; j++) H[HB_TRI(nux)+nu+j] = q[n][j];
		double *d = blk + s->off_d;
		for(j=0; j<s->nb; j++) { d[j] = lb[n][j]; d[s->nb+j] = ub[n][j]; }
		}
	return 0;
	}

/* general constraints lg <= D u + C x <= ug of one instance into its block: C[n] ng x nx, D[n] ng x nu (n < N), row- or
 * column-major like the other matrices; stored as [D C]'_n (interfaces/c/fortran_order_interface.c:276-283) */
int hpmpc_b200_pack_general(const hpmpc_b200_ocp *p, int c_order, double *const *C, double *const *D, double *const *lg,
		double *const *ug, double *blk)
	{
	int n, i, j;
	for(n=0; n<=p->N; n++)
		{
		const hb_stage *s = &p->st[n];
		const int nx = s->nx, nu = s->nu, ng = s->ng;
		if(ng==0) continue;
		double *G = blk + s->off_DCt;                          /* (nu+nx) x ng row-major */
		for(i=0; i<nu; i++) for(j=0; j<ng; j++) G[i*ng+j] = EL(D[n], j, i, ng, nu, c_order);
		for(i=0; i<nx; i++) for(j=0; j<ng; j++) G[(nu+i)*ng+j] = EL(C[n], j, i, ng, nx, c_order);
		for(j=0; j<ng; j++) { blk[s->off_dg+j] = lg[n][j]; blk[s->off_dg+ng+j] = ug[n][j]; }
		}
	return 0;
	}

void hpmpc_b200_unpack_solution(const hpmpc_b200_ocp *p, const double *ux, const double *pi, const double *lam,
		double **x, double **u, double **pi_out, double **lam_out)
	{
	int n, i;
	for(n=0; n<=p->N; n++)
		{
		const hb_stage *s = &p->st[n];
		/* an embedded pattern returns the caller's real entries: they come first in the u- and in the x-part of the padded frame */
		const int nu = p->r_nu ? p->r_nu[n] : s->nu, nx = p->r_nx ? p->r_nx[n] : s->nx, nx1 = p->r_nx ? (n<p->N ? p->r_nx[n+1] : 0) : s->nx1;
		if(u && n<p->N) for(i=0; i<nu; i++) u[n][i] = ux[s->off_ux+i];
		if(x) for(i=0; i<nx; i++) x[n][i] = ux[s->off_ux+s->nu+i];
		if(pi_out && pi && n<p->N) for(i=0; i<nx1; i++) pi_out[n][i] = pi[s->off_pi+i];
		if(lam_out && lam) for(i=0; i<2*(s->nb+s->ng); i++) lam_out[n][i] = lam[2*s->off_c+i];
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* device-pointer entry points                                                                       */
/* ------------------------------------------------------------------------------------------------ */
/* one call in flight per handle (see ev_busy): the stream of a new call first waits for the previous call's work */
static int call_begin(hpmpc_b200_ocp *p, void *stream)
	{
	if(p->busy_valid) CK(cudaStreamWaitEvent((cudaStream_t)stream, p->ev_busy, 0));
	return 0;
	}
static int call_end(hpmpc_b200_ocp *p, void *stream, int rc)
	{
	if(rc) return rc;
	CK(cudaEventRecord(p->ev_busy, (cudaStream_t)stream));
	p->busy_valid = 1;
	return 0;
	}

static int grid_for(const hpmpc_b200_ocp *p, long long n_inst)
	{
	long long need = (n_inst + p->warps - 1)/p->warps;
	return (int)(need<p->grid ? (need<1 ? 1 : need) : p->grid);
	}

int hpmpc_b200_d_back_ric_rec_sv_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in,
		double *d_ux, double *d_pi, double *d_Pb, void *stream)
	{
	return hpmpc_b200_d_back_ric_rec_sv_upd_batch(p, n_inst, d_in, NULL, NULL, d_ux, d_pi, d_Pb, stream);
	}

/* the same with the reference's Qx / qx arguments (d_back_ric_rec.c:112): per-constraint terms added to the Hessian diagonal and
 * to the gradient row -- for general constraints through [D C]' -- nbtot doubles per instance each, either may be NULL */
int hpmpc_b200_d_back_ric_rec_sv_upd_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_Qx, const double *d_qx,
		double *d_ux, double *d_pi, double *d_Pb, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	if(p->fast_id>=0 && d_Pb==NULL && d_Qx==NULL && d_qx==NULL)
		{
		long long groups = (n_inst + p->f_ipw - 1)/p->f_ipw, need = (groups + p->f_warps - 1)/p->f_warps;
		int grid = (int)(need<p->f_grid ? need : p->f_grid);
		if(ensure_scratch(p, sizeof(double)*(size_t)p->f_grid*p->f_warps*p->f_ipw*p->f_stash_inst)) return -1;
		return call_end(p, stream, hb_launch_ric_sv_fast(p->fast_id, &p->dims, n_inst, d_in, d_ux, d_pi, p->scratch, grid, p->f_warps, stream));
		}
	if(ensure_scratch(p, sizeof(double)*(size_t)p->n_slots*p->dims.L_stride)) return -1;
	return call_end(p, stream, hb_launch_ric_sv(&p->dims, n_inst, d_in, d_ux, d_pi, d_Pb, p->scratch, p->n_slots, grid_for(p, n_inst), p->warps, stream, d_Qx, d_qx));
	}

/* MEASUREMENT TOOL: the memory traffic of hpmpc_b200_d_back_ric_rec_sv_batch (same bulk copies, same L2 hints, same output stores)
 * without the arithmetic -- the ceiling the memory system sets for the two-sweep access pattern.  d_ux / d_pi receive
 * meaningless values.  Only for the config-2 shape (nx = 12, nu = 5, x0 eliminated); -2 otherwise. */
int hpmpc_b200_sv_traffic_probe(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, double *d_ux, double *d_pi, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0 || p->fast_id!=0) return -2;
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	long long groups = (n_inst + p->f_ipw - 1)/p->f_ipw, need = (groups + p->f_warps - 1)/p->f_warps;
	int grid = (int)(need<p->f_grid ? need : p->f_grid);
	if(ensure_scratch(p, sizeof(double)*(size_t)p->f_grid*p->f_warps*p->f_ipw*p->f_stash_inst)) return -1;
	return call_end(p, stream, hb_launch_sv_traffic(p->fast_id, &p->dims, n_inst, d_in, d_ux, d_pi, p->scratch, grid, p->f_warps, stream));
	}

int hpmpc_b200_d_back_ric_rec_trf_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, double *d_L, void *stream)
	{
	return hpmpc_b200_d_back_ric_rec_trf_upd_batch(p, n_inst, d_in, NULL, d_L, stream);
	}

int hpmpc_b200_d_back_ric_rec_trf_upd_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_Qx, double *d_L, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	if(p->tf_id>=0 && d_Qx!=NULL)
		{ fprintf(stderr, "hpmpc_b200: trf with Qx needs the generic factor layout: call hpmpc_b200_ocp_generic_factor_layout() on the handle first\n"); return -5; }
	if(call_begin(p, stream)) return -1;
	if(p->tf_id>=0)
		{
		long long need = (n_inst + p->tf_warps - 1)/p->tf_warps;
		return call_end(p, stream, hb_launch_ric_trf_trs_fast(p->tf_id, 0, &p->dims, n_inst, d_in, d_L, p->tf_L_stride, NULL, NULL, NULL,
				(int)(need<p->tf_grid ? need : p->tf_grid), p->tf_warps, stream));
		}
	return call_end(p, stream, hb_launch_ric_trf(&p->dims, n_inst, d_in, d_L, grid_for(p, n_inst), p->warps, stream, d_Qx));
	}

int hpmpc_b200_d_back_ric_rec_trs_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_L,
		double *d_ux, double *d_pi, void *stream)
	{
	return hpmpc_b200_d_back_ric_rec_trs_upd_batch(p, n_inst, d_in, d_L, NULL, d_ux, d_pi, stream);
	}

int hpmpc_b200_d_back_ric_rec_trs_upd_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_L, const double *d_qx,
		double *d_ux, double *d_pi, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	if(p->tf_id>=0 && d_qx!=NULL)
		{ fprintf(stderr, "hpmpc_b200: trs with qx needs the generic factor layout: call hpmpc_b200_ocp_generic_factor_layout() on the handle first\n"); return -5; }
	if(call_begin(p, stream)) return -1;
	if(p->tf_id>=0)
		{
		long long need = (n_inst + p->tf_warps - 1)/p->tf_warps;
		if(ensure_scratch(p, sizeof(double)*(size_t)p->tf_grid*p->tf_warps*(p->dims.ux_stride+p->dims.pi_stride))) return -1;
		return call_end(p, stream, hb_launch_ric_trf_trs_fast(p->tf_id, 1, &p->dims, n_inst, d_in, (double*)d_L, p->tf_L_stride, d_ux, d_pi, p->scratch,
				(int)(need<p->tf_grid ? need : p->tf_grid), p->tf_warps, stream));
		}
	if(ensure_scratch(p, sizeof(double)*(size_t)p->n_slots*(p->dims.ux_stride+2*p->dims.pi_stride))) return -1;
	return call_end(p, stream, hb_launch_ric_trs(&p->dims, n_inst, d_in, d_L, d_ux, d_pi, p->scratch, p->n_slots, grid_for(p, n_inst), p->warps, stream, d_qx));
	}

/* Waves: one launch per `slots` instances (one instance per resident warp).  Warps of a wave stay in step, so the SM's
 * instruction cache serves all of them from the same sweep; a single persistent launch in which warps pull instances from a
 * queue drifts apart and was measured 2.3x slower on config 3 (the fused kernel is several hundred KB of SASS).
 * HPMPC_B200_IPM_CHUNK=0 restores the single launch, any other value sets the wave size. */
/* KKT state of one instance: the IPM kernel's work slot (factor, vectors, t_inv) followed by the backup of the iterate the
 * factor belongs to (ux, pi, lam_lo, lam_up, t_lo, t_up) and a flag */
static long long kkt_stride(const hpmpc_b200_ocp *p)
	{
	return HB_EVEN(p->ipm_ws) + p->dims.ux_stride + p->dims.pi_stride + 4*(long long)HB_EVEN(p->dims.nbtot) + 2;
	}
long long hpmpc_b200_kkt_state_stride(const hpmpc_b200_ocp *p) { return kkt_stride(p); }

static int ipm_waves(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, double *scratch, int *counter,
		long long lam_len, void *stream, double *d_kkt)
	{
	const long long ws = p->ipm_ws;
	const long long ks = kkt_stride(p);
	const int slots = p->i_grid*p->i_warps;
	long long chunk = slots;
	{ const char *e = getenv("HPMPC_B200_IPM_CHUNK"); if(e) chunk = atoll(e); }
	if(chunk<=0 || chunk>n_inst) chunk = n_inst;
	const long long info_len = HB_IPM_INFO_HEAD + 5*(long long)k_max;
	long long done;
	for(done=0; done<n_inst; done+=chunk)
		{
		long long m = n_inst-done<chunk ? n_inst-done : chunk;
		long long need = (m + p->i_warps - 1)/p->i_warps;
		int grid = (int)(need<p->i_grid ? (need<1 ? 1 : need) : p->i_grid);
		int rc = hb_launch_ipm_kkt(&p->dims, m, d_in + done*p->dims.in_stride, k_max, mu0, mu_tol, alpha_min, warm_start,
				d_ux + done*p->dims.ux_stride, d_pi + done*p->dims.pi_stride, d_lam + done*lam_len, d_t + done*lam_len,
				d_info + done*info_len, scratch, ws, slots, grid, p->i_warps, counter, p->ipm_fast_id, stream,
				d_kkt ? d_kkt + done*ks : NULL, ks);
		if(rc) return rc;
		}
	return 0;
	}

/* Large batches: the multi-kernel driver with active-set compaction (cipm_kernels.cu).  Every instance gets its own state block,
 * so the batch is cut into chunks whose state fits `cap` bytes (HPMPC_B200_IPM_STATE_GB, default 16 GiB); small batches (less than
 * two waves of the fused kernel), k_max < 1 and unconstrained patterns stay on the fused one-kernel path.  HPMPC_B200_IPM_FUSED=1
 * forces the fused path, =0 forces the multi-kernel one.  Both give the same bits (tests/test_cipm.py). */
int hbt_wanted(const hb_dims *d);
static int ipm_multi_wanted(const hpmpc_b200_ocp *p, long long n_inst, int k_max, int warm_start)
	{
	const char *e = getenv("HPMPC_B200_IPM_FUSED");
	if(p->dims.nbtot<=0 || k_max<1 || warm_start==2) return 0;
	if(e) return atoi(e)==0;
	/* measured on B200 (profiles/r02_ipm_multi_kernel.txt): config 3 (size-specialised sweeps) 67.4 -> 82.9 K solves/s; config 4
	 * (run-time-size sweeps, one warp per instance in every kernel) 71.2 -> 64.4 K; with the four-warps-per-instance factorisation
	 * kernel (ric_team.cuh), which only the multi-kernel driver has, the any-size patterns go there too */
	if(p->ipm_fast_id<0) return hbt_wanted(&p->dims) && n_inst >= 64;
	return n_inst >= 2LL*p->i_grid*p->i_warps;
	}

static int ipm_multi(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, void *stream)
	{
	const long long ws = HB_EVEN(p->ipm_ws);
	double cap_gb = 16.0;
	{ const char *e = getenv("HPMPC_B200_IPM_STATE_GB"); if(e && atof(e)>0) cap_gb = atof(e); }
	long long chunk = (long long)(cap_gb*1073741824.0/(8.0*(double)ws));
	if(chunk<1) chunk = 1;
	if(chunk>n_inst) chunk = n_inst;
	const size_t work_b = sizeof(double)*(size_t)chunk*ws;
	const size_t aux_b = (size_t)hb_cipm_aux_bytes(chunk);
	if(ensure_scratch(p, work_b + aux_b + 256)) return -1;
	const long long info_len = HB_IPM_INFO_HEAD + 5*(long long)k_max;
	long long done;
	for(done=0; done<n_inst; done+=chunk)
		{
		long long m = n_inst-done<chunk ? n_inst-done : chunk;
		int rc = hb_launch_cipm(&p->dims, m, d_in + done*p->dims.in_stride, k_max, mu0, mu_tol, alpha_min, warm_start,
				d_ux + done*p->dims.ux_stride, d_pi + done*p->dims.pi_stride, d_lam + done*p->lam_stride, d_t + done*p->lam_stride,
				d_info + done*info_len, p->scratch, ws, (char*)p->scratch + work_b, p->i_grid, p->i_warps, p->sms, p->ipm_fast_id, stream);
		if(rc) return rc;
		}
	return 0;
	}

int hpmpc_b200_d_ip2_res_mpc_hard_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t,
		double *d_info, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	if(ipm_multi_wanted(p, n_inst, k_max, warm_start))
		return call_end(p, stream, ipm_multi(p, n_inst, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info, stream));
	if(ensure_scratch(p, sizeof(double)*(size_t)p->i_grid*p->i_warps*p->ipm_ws)) return -1;
	return call_end(p, stream, ipm_waves(p, n_inst, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info, p->scratch, p->counter,
			p->lam_stride, stream, NULL));
	}

/* residuals of a given point: d_res_res_mpc_hard_tv (mpc_solvers/c99/d_res_ip_res_hard.c:39; row a6).  d_rq [ux_stride],
 * d_rb [pi_stride], d_rd / d_rm [lam_stride] in the layout of lam, d_mu [1] per instance; d_rm may be NULL */
int hpmpc_b200_d_res_res_mpc_hard_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, const double *d_ux, const double *d_pi,
		const double *d_lam, const double *d_t, double *d_rq, double *d_rb, double *d_rd, double *d_rm, double *d_mu, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	if(ensure_scratch(p, sizeof(double)*(size_t)p->n_slots*hb_res_work_doubles(&p->dims))) return -1;
	return call_end(p, stream, hb_launch_res(&p->dims, n_inst, d_in, d_ux, d_pi, d_lam, d_t, d_rq, d_rb, d_rd, d_rm, d_mu, p->scratch,
			grid_for(p, n_inst), p->warps, stream));
	}

/* a fixed number of residual-based Newton steps from a caller-supplied iterate: d_ip2_res_mpc_hard_tv_single_newton_step
 * (mpc_solvers/d_ip2_res_hard.c:1348).  On entry d_ux, d_pi, d_lam, d_t hold (ux0, pi0, lam0, t0) in the output
 * layouts; they are updated in place.  The centering term is mu0 in every step, as in the reference.  Bounds only (the reference
 * rejects ng > 0 here, c99/d_aux_ip_hard_lib4.c:205-210). */
int hpmpc_b200_d_ip2_res_mpc_hard_single_newton_step_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0,
		double alpha_min, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	if(p->dims.ngtot>0) { fprintf(stderr, "hpmpc_b200: single Newton step: general constraints are not supported (neither does the reference)\n"); return -5; }
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	if(ensure_scratch(p, sizeof(double)*(size_t)p->i_grid*p->i_warps*p->ipm_ws)) return -1;
	return call_end(p, stream, ipm_waves(p, n_inst, d_in, k_max, mu0, 0.0, alpha_min, 2, d_ux, d_pi, d_lam, d_t, d_info, p->scratch, p->counter,
			p->lam_stride, stream, NULL));
	}

/* the same solve, every instance leaving its KKT state in d_kkt (hpmpc_b200_kkt_state_stride() doubles each) */
int hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t,
		double *d_info, double *d_kkt, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	if(d_kkt==NULL) return -2;
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	if(ensure_scratch(p, sizeof(double)*(size_t)p->i_grid*p->i_warps*p->ipm_ws)) return -1;
	return call_end(p, stream, ipm_waves(p, n_inst, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info, p->scratch, p->counter,
			p->lam_stride, stream, d_kkt));
	}

/* the last KKT system of that solve again, for the b, [r q] and bounds held in d_in (same matrices): one solve with the stored
 * factor per instance (reference d_kkt_solve_new_rhs_res_mpc_hard_tv, mpc_solvers/d_ip2_res_hard.c:1922).
 * d_info: 6 doubles per instance, [1] = 0 or -10 (the IPM left no phase-2 factor), [2..4] residual norms at the backup, [5] mu */
int hpmpc_b200_d_kkt_solve_new_rhs_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in, double *d_kkt,
		double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	if(d_kkt==NULL) return -2;
	CK(cudaSetDevice(p->device));
	long long need = (n_inst + p->i_warps - 1)/p->i_warps;
	int grid = (int)(need<p->i_grid ? (need<1 ? 1 : need) : p->i_grid);
	if(call_begin(p, stream)) return -1;
	return call_end(p, stream, hb_launch_kkt_new_rhs(&p->dims, n_inst, d_in, d_kkt, kkt_stride(p), d_ux, d_pi, d_lam, d_t, d_info, grid, p->i_warps,
			p->counter, p->ipm_fast_id, stream));
	}

/* ------------------------------------------------------------------------------------------------ */
/* host-buffer entry points: two staging buffers, copy-in / solve / copy-out of chunk k overlaps k+1  */
/* ------------------------------------------------------------------------------------------------ */
static int ensure_staging(hpmpc_b200_ocp *p, size_t in_bytes, size_t out_bytes)
	{
	int k;
	for(k=0; k<2; k++)
		{
		if(!p->s_copy[k]) CK(cudaStreamCreateWithFlags(&p->s_copy[k], cudaStreamNonBlocking));
		if(!p->ev[k]) CK(cudaEventCreateWithFlags(&p->ev[k], cudaEventDisableTiming));
		if(!p->ev_in[k]) CK(cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming));
		if(!p->ev_done[k]) CK(cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming));
		}
	if(!p->s_comp) CK(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
	if(in_bytes>p->stage_in_bytes)
		{
		for(k=0; k<2; k++) { if(p->stage_in[k]) CK(cudaFree(p->stage_in[k])); p->stage_in[k] = NULL; CK(cudaMalloc((void**)&p->stage_in[k], in_bytes)); }
		p->stage_in_bytes = in_bytes;
		}
	if(out_bytes>p->stage_out_bytes)
		{
		for(k=0; k<2; k++) { if(p->stage_out[k]) CK(cudaFree(p->stage_out[k])); p->stage_out[k] = NULL; CK(cudaMalloc((void**)&p->stage_out[k], out_bytes)); }
		p->stage_out_bytes = out_bytes;
		}
	return 0;
	}

static long long chunk_size(const hpmpc_b200_ocp *p, long long n_inst)
	{
	/* a chunk should fill the persistent grid a few times over, and stay below ~512 MiB of input */
	long long c = 8LL*p->n_slots;
	long long cap = (512LL<<20)/(long long)(sizeof(double)*p->dims.in_stride);
	if(cap<1) cap = 1;
	if(c>cap) c = cap;
	if(c>n_inst) c = n_inst;
	return c;
	}

int hpmpc_b200_d_back_ric_rec_sv_batch_host(hpmpc_b200_ocp *p, long long n_inst, const double *h_in, double *h_ux, double *h_pi)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	const long long cs = chunk_size(p, n_inst);
	const size_t in_b = sizeof(double)*(size_t)cs*p->dims.in_stride;
	const size_t ux_b = sizeof(double)*(size_t)cs*p->dims.ux_stride, pi_b = sizeof(double)*(size_t)cs*p->dims.pi_stride;
	if(ensure_staging(p, in_b, ux_b+pi_b)) return -1;
	if(call_begin(p, p->s_copy[0]) || call_begin(p, p->s_copy[1])) return -1;
	/* each stream owns its own slice of the scratch slots: both chunks can be in flight */
	{
	size_t gen = sizeof(double)*(size_t)2*p->n_slots*p->dims.L_stride;
	size_t fst = p->fast_id>=0 ? sizeof(double)*(size_t)2*p->f_grid*p->f_warps*p->f_ipw*p->f_stash_inst : 0;
	if(ensure_scratch(p, gen>fst ? gen : fst)) return -1;
	}
	long long done; int k = 0;
	for(done=0; done<n_inst; done+=cs, k^=1)
		{
		long long m = n_inst-done<cs ? n_inst-done : cs;
		cudaStream_t st = p->s_copy[k];
		double *d_in = p->stage_in[k], *d_ux = p->stage_out[k], *d_pi = d_ux + (size_t)cs*p->dims.ux_stride;
		CK(cudaMemcpyAsync(d_in, h_in + (size_t)done*p->dims.in_stride, sizeof(double)*(size_t)m*p->dims.in_stride, cudaMemcpyHostToDevice, st));
		if(p->fast_id>=0)
			{
			long long groups = (m + p->f_ipw - 1)/p->f_ipw, need = (groups + p->f_warps - 1)/p->f_warps;
			int grid = (int)(need<p->f_grid ? need : p->f_grid);
			if(hb_launch_ric_sv_fast(p->fast_id, &p->dims, m, d_in, d_ux, d_pi,
					p->scratch + (size_t)k*p->f_grid*p->f_warps*p->f_ipw*p->f_stash_inst, grid, p->f_warps, st)) return -1;
			}
		else if(hb_launch_ric_sv(&p->dims, m, d_in, d_ux, d_pi, NULL, p->scratch + (size_t)k*p->n_slots*p->dims.L_stride,
				p->n_slots, grid_for(p, m), p->warps, st, NULL, NULL)) return -1;
		CK(cudaMemcpyAsync(h_ux + (size_t)done*p->dims.ux_stride, d_ux, sizeof(double)*(size_t)m*p->dims.ux_stride, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(h_pi + (size_t)done*p->dims.pi_stride, d_pi, sizeof(double)*(size_t)m*p->dims.pi_stride, cudaMemcpyDeviceToHost, st));
		}
	CK(cudaStreamSynchronize(p->s_copy[0]));
	CK(cudaStreamSynchronize(p->s_copy[1]));
	p->busy_valid = 0;                   /* everything this handle issued has completed */
	return 0;
	}

int hpmpc_b200_d_ip2_res_mpc_hard_batch_host(hpmpc_b200_ocp *p, long long n_inst, const double *h_in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *h_ux, double *h_pi, double *h_lam, double *h_t, double *h_info)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	/* one chunk = one wave of the IPM kernel (one instance per resident warp), or two waves where the multi-kernel driver
	 * (cipm_kernels.cu, the faster path from two waves on) serves the pattern; never more than ~1.5 GiB of staging per buffer */
	const long long wave = (long long)p->i_grid*p->i_warps;
	long long cs = wave;
	{
	long long cap = (1536LL<<20)/(long long)(sizeof(double)*p->dims.in_stride);
	if(cap<1) cap = 1;
	if(ipm_multi_wanted(p, 2*wave, k_max, warm_start) && n_inst>=2*wave && cap>=2*wave) cs = 2*wave;
	if(cs>cap) cs = cap;
	if(cs>n_inst) cs = n_inst;
	}
	const long long info_len = HB_IPM_INFO_HEAD + 5*(long long)k_max, lam_len = p->lam_stride>0 ? p->lam_stride : 2;
	const size_t in_b = sizeof(double)*(size_t)cs*p->dims.in_stride;
	const size_t out_d = (size_t)cs*(p->dims.ux_stride + p->dims.pi_stride + 2*lam_len + info_len);
	if(ensure_staging(p, in_b, sizeof(double)*out_d)) return -1;
	if(call_begin(p, p->s_comp)) return -1;
	const long long ws = p->ipm_ws;
	const int slots = p->i_grid*p->i_warps;
	{
	/* scratch for both drivers up front: growing it between chunks would synchronise the device */
	size_t need = sizeof(double)*(size_t)slots*ws;
	if(cs>=2*wave)
		{
		const size_t multi_b = sizeof(double)*(size_t)cs*HB_EVEN(ws) + (size_t)hb_cipm_aux_bytes(cs) + 256;
		if(multi_b>need) need = multi_b;
		}
	if(ensure_scratch(p, need)) return -1;
	}
	/* copies alternate between two staging buffers / copy streams; the waves of all chunks run back to back on ONE compute
	 * stream (two IPM kernels sharing the SMs would drift apart in the instruction cache, see ipm_waves) */
	long long done; int k = 0;
	for(done=0; done<n_inst; done+=cs, k^=1)
		{
		long long m = n_inst-done<cs ? n_inst-done : cs;
		cudaStream_t st = p->s_copy[k];
		double *d_in = p->stage_in[k];
		double *d_ux = p->stage_out[k], *d_pi = d_ux + (size_t)cs*p->dims.ux_stride, *d_lam = d_pi + (size_t)cs*p->dims.pi_stride;
		double *d_t = d_lam + (size_t)cs*lam_len, *d_info = d_t + (size_t)cs*lam_len;
		CK(cudaMemcpyAsync(d_in, h_in + (size_t)done*p->dims.in_stride, sizeof(double)*(size_t)m*p->dims.in_stride, cudaMemcpyHostToDevice, st));
		if(warm_start) CK(cudaMemcpyAsync(d_ux, h_ux + (size_t)done*p->dims.ux_stride, sizeof(double)*(size_t)m*p->dims.ux_stride, cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(d_info, 0, sizeof(double)*(size_t)m*info_len, st));
		CK(cudaEventRecord(p->ev_in[k], st));
		CK(cudaStreamWaitEvent(p->s_comp, p->ev_in[k], 0));
		if(m>=2*wave && ipm_multi_wanted(p, m, k_max, warm_start))
			{
			if(ipm_multi(p, m, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info, p->s_comp)) return -1;
			}
		else if(ipm_waves(p, m, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info,
				p->scratch, p->counter, lam_len, p->s_comp, NULL)) return -1;
		CK(cudaEventRecord(p->ev_done[k], p->s_comp));
		CK(cudaStreamWaitEvent(st, p->ev_done[k], 0));
		CK(cudaMemcpyAsync(h_ux + (size_t)done*p->dims.ux_stride, d_ux, sizeof(double)*(size_t)m*p->dims.ux_stride, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(h_pi + (size_t)done*p->dims.pi_stride, d_pi, sizeof(double)*(size_t)m*p->dims.pi_stride, cudaMemcpyDeviceToHost, st));
		if(p->lam_stride>0)
			{
			CK(cudaMemcpyAsync(h_lam + (size_t)done*p->lam_stride, d_lam, sizeof(double)*(size_t)m*p->lam_stride, cudaMemcpyDeviceToHost, st));
			if(h_t) CK(cudaMemcpyAsync(h_t + (size_t)done*p->lam_stride, d_t, sizeof(double)*(size_t)m*p->lam_stride, cudaMemcpyDeviceToHost, st));
			}
		CK(cudaMemcpyAsync(h_info + (size_t)done*info_len, d_info, sizeof(double)*(size_t)m*info_len, cudaMemcpyDeviceToHost, st));
		}
	CK(cudaStreamSynchronize(p->s_copy[0]));
	CK(cudaStreamSynchronize(p->s_copy[1]));
	p->busy_valid = 0;                   /* everything this handle issued has completed */
	return 0;
	}

/* ------------------------------------------------------------------------------------------------ */
/* shared dynamics: one set of matrices for the whole batch, per-instance vectors (ric_kernels.cu: hb_ric_trs_shared_kernel)     */
/* ------------------------------------------------------------------------------------------------ */
long long hpmpc_b200_shared_vec_stride(const hpmpc_b200_ocp *p) { return p->dims.ux_stride + p->dims.pi_stride; }
long long hpmpc_b200_shared_factor_doubles(const hpmpc_b200_ocp *p) { return p->dims.L_stride; }

/* factorise the ONE shared block (d_in_shared: a packed instance block whose vectors are ignored) into d_L_shared */
int hpmpc_b200_d_back_ric_rec_trf_shared(hpmpc_b200_ocp *p, const double *d_in_shared, double *d_L_shared, void *stream)
	{
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	if(call_begin(p, stream)) return -1;
	return call_end(p, stream, hb_launch_ric_trf(&p->dims, 1, d_in_shared, d_L_shared, 1, 1, stream, NULL));
	}

static int trs_shared_shape(hpmpc_b200_ocp *p, long long n_inst, int *grid, int *warps, int *smem, int *resident)
	{
	*warps = 8;
	*smem = (int)hb_trs_shared_smem_bytes(&p->dims, p->st, *warps, resident);
	if(*smem>227*1024) { *warps = 4; *smem = (int)hb_trs_shared_smem_bytes(&p->dims, p->st, *warps, resident); }
	if(*smem>227*1024) { fprintf(stderr, "hpmpc_b200: shared-dynamics solve: stage too large for shared memory\n"); return -2; }
	int per_sm = (227*1024)/(*smem+1024); if(per_sm<1) per_sm = 1; if(per_sm>2) per_sm = 2;
	long long need = (n_inst + *warps - 1)/(*warps);
	*grid = (int)(need<(long long)p->sms*per_sm ? (need<1 ? 1 : need) : (long long)p->sms*per_sm);
	return 0;
	}

/* every instance: solve with the shared factor for its own vectors.  d_vec: n_inst x hpmpc_b200_shared_vec_stride() doubles,
 * per instance [r q] of every stage in the ux layout (ux_stride), then b of every stage in the pi layout (pi_stride). */
int hpmpc_b200_d_back_ric_rec_trs_shared_batch(hpmpc_b200_ocp *p, long long n_inst, const double *d_in_shared, const double *d_L_shared,
		const double *d_vec, double *d_ux, double *d_pi, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	int grid, warps, smem, resident;
	if(trs_shared_shape(p, n_inst, &grid, &warps, &smem, &resident)) return -2;
	if(call_begin(p, stream)) return -1;
	if(ensure_scratch(p, sizeof(double)*(size_t)grid*warps*p->dims.pi_stride)) return -1;
	return call_end(p, stream, hb_launch_ric_trs_shared(&p->dims, n_inst, d_in_shared, d_L_shared, d_vec, d_ux, d_pi, p->scratch, grid, warps, smem, resident, stream));
	}

/* the same with STAGE-MAJOR vectors (thread-per-instance shapes only, -2 otherwise): the part of a vector that belongs to stage n --
 * offset o_n in the instance-major layout, K_n doubles -- is one array [n_inst][K_n] at o_n * n_inst; d_vec: [r q] parts
 * (ux offsets, nu_n + nx_n doubles), then from ux_stride * n_inst on the b parts (pi offsets, nx_{n+1} doubles); d_ux, d_pi alike.
 * A warp's 32 instances then touch one contiguous run per stage, which is what DRAM wants. */
int hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major(hpmpc_b200_ocp *p, long long n_inst, const double *d_in_shared, const double *d_L_shared,
		const double *d_vec, double *d_ux, double *d_pi, void *stream)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	int grid, warps, smem, resident;
	if(trs_shared_shape(p, n_inst, &grid, &warps, &smem, &resident)) return -2;
	if(resident<16) { fprintf(stderr, "hpmpc_b200: stage-major shared-dynamics solve: no thread-per-instance kernel for this size pattern\n"); return -2; }
	if(call_begin(p, stream)) return -1;
	if(ensure_scratch(p, sizeof(double)*(size_t)grid*warps*p->dims.pi_stride)) return -1;
	return call_end(p, stream, hb_launch_ric_trs_shared(&p->dims, n_inst, d_in_shared, d_L_shared, d_vec, d_ux, d_pi, p->scratch, grid, warps, smem, resident + 64, stream));
	}

/* the same from host buffers: the shared block goes over once and is factorised once, then the vectors stream through in chunks
 * (H2D of chunk k+1 overlaps the solve / D2H of chunk k) */
int hpmpc_b200_d_back_ric_rec_sv_shared_batch_host(hpmpc_b200_ocp *p, long long n_inst, const double *h_in_shared, const double *h_vec,
		double *h_ux, double *h_pi)
	{
	if(n_inst<=0) return 0;
	if(p->device<0) { fprintf(stderr, "hpmpc_b200: host-only handle (device < 0) cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(p->device));
	const long long vs = p->dims.ux_stride + p->dims.pi_stride;
	/* chunk: 128 MB of vectors (the solve of a chunk is far shorter than its copies, so the pipeline's fill and drain -- one
	 * chunk's H2D and one chunk's D2H -- is what small chunks save); HPMPC_B200_SHARED_CHUNK_MB overrides */
	long long chunk_mb = 128;
	{ const char *e = getenv("HPMPC_B200_SHARED_CHUNK_MB"); if(e && atoll(e)>0) chunk_mb = atoll(e); }
	long long cs = (chunk_mb<<20)/(long long)(sizeof(double)*vs);
	if(cs<1) cs = 1;
	if(cs>n_inst) cs = n_inst;
	int grid, warps, smem, resident, k = 0;
	if(trs_shared_shape(p, cs, &grid, &warps, &smem, &resident)) return -2;
	const size_t in_b = sizeof(double)*(size_t)cs*vs, out_b = sizeof(double)*(size_t)cs*(p->dims.ux_stride + p->dims.pi_stride);
	if(ensure_staging(p, in_b, out_b)) return -1;
	if(call_begin(p, p->s_copy[0]) || call_begin(p, p->s_copy[1])) return -1;
	/* scratch: [shared block | shared factor | Pb of stream 0 | Pb of stream 1] */
	const size_t pb = (size_t)grid*warps*p->dims.pi_stride;
	if(ensure_scratch(p, sizeof(double)*((size_t)p->dims.in_stride + p->dims.L_stride + 2*pb + 64))) return -1;
	double *d_sh = p->scratch, *d_L = d_sh + p->dims.in_stride, *d_pb = d_L + HB_EVEN(p->dims.L_stride);
	CK(cudaMemcpyAsync(d_sh, h_in_shared, sizeof(double)*p->dims.in_stride, cudaMemcpyHostToDevice, p->s_copy[0]));
	if(hb_launch_ric_trf(&p->dims, 1, d_sh, d_L, 1, 1, p->s_copy[0], NULL)) return -1;
	CK(cudaEventRecord(p->ev[0], p->s_copy[0]));
	CK(cudaStreamWaitEvent(p->s_copy[1], p->ev[0], 0));
	long long done;
	for(done=0; done<n_inst; done+=cs, k^=1)
		{
		long long m = n_inst-done<cs ? n_inst-done : cs;
		cudaStream_t st = p->s_copy[k];
		double *d_vec = p->stage_in[k], *d_ux = p->stage_out[k], *d_pi = d_ux + (size_t)cs*p->dims.ux_stride;
		int g2 = grid; { long long need = (m + warps - 1)/warps; if(need<g2) g2 = (int)(need<1 ? 1 : need); }
		CK(cudaMemcpyAsync(d_vec, h_vec + (size_t)done*vs, sizeof(double)*(size_t)m*vs, cudaMemcpyHostToDevice, st));
		if(hb_launch_ric_trs_shared(&p->dims, m, d_sh, d_L, d_vec, d_ux, d_pi, d_pb + (size_t)k*pb, g2, warps, smem, resident, st)) return -1;
		CK(cudaMemcpyAsync(h_ux + (size_t)done*p->dims.ux_stride, d_ux, sizeof(double)*(size_t)m*p->dims.ux_stride, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(h_pi + (size_t)done*p->dims.pi_stride, d_pi, sizeof(double)*(size_t)m*p->dims.pi_stride, cudaMemcpyDeviceToHost, st));
		}
	CK(cudaStreamSynchronize(p->s_copy[0]));
	CK(cudaStreamSynchronize(p->s_copy[1]));
	p->busy_valid = 0;
	return 0;
	}

double hpmpc_b200_fp64_peak_tflops(int device)
	{
	if(cudaSetDevice(device)!=cudaSuccess) return -1.0;
	return hb_fp64_peak_probe(device, 1<<16, NULL);
	}

const char *hpmpc_b200_version(void) { return "hpmpc_b200 0.1 (sm_100a)"; }

/* used by compat.c only: the legacy symbols pass the factor of sv / trf to trs through the caller's `memory`, whose size and
 * layout are the generic ones (d_back_ric_rec_sv_tv_memory_space_size_bytes) -- keep trf / trs on the generic kernels there */
void hpmpc_b200_internal_generic_trf(hpmpc_b200_ocp *p) { p->tf_id = -1; }
/* public name of the same switch: trf / trs keep the factor in the generic packed-trapezoid layout (layout.h) even for a shape
 * that has size-specialised sweeps -- needed for the _upd_ variants (Qx / qx) and by callers that read the factor */
void hpmpc_b200_ocp_generic_factor_layout(hpmpc_b200_ocp *p) { p->tf_id = -1; }

/* used by pcond.c only: the device-side descriptor and the host copy of the stage table of a handle */
const hb_dims *hpmpc_b200_internal_dims(const hpmpc_b200_ocp *p) { return &p->dims; }
const hb_stage *hpmpc_b200_internal_stages(const hpmpc_b200_ocp *p) { return p->st; }
int hpmpc_b200_internal_device(const hpmpc_b200_ocp *p, int *sms) { if(sms) *sms = p->sms; return p->device; }

/* used by compat.c only: the factor of a batch-of-one sv call sits in scratch slot 0 */
int hpmpc_b200_internal_copy_stash(hpmpc_b200_ocp *p, double *h_dst)
	{
	CK(cudaMemcpy(h_dst, p->scratch, sizeof(double)*(size_t)p->dims.L_stride, cudaMemcpyDeviceToHost));
	return 0;
	}
