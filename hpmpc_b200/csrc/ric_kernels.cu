/*
 * ric_kernels.cu -- sm_100a kernels of the batched Riccati / box-IPM engine and their C launchers.
 *
 * Kernels (one warp per OCP instance, persistent grid, per-warp scratch "slot" for the factor stash):
 *   hb_ric_sv_kernel   factor + solve            <- d_back_ric_rec_sv_tv_res   (lqcp_solvers/d_back_ric_rec.c:112)
 *   hb_ric_trf_kernel  factor, L kept per inst.  <- d_back_ric_rec_trf_tv_res  (:403)
 *   hb_ric_trs_kernel  solve with stored L       <- d_back_ric_rec_trs_tv_res  (:564)
 *   hb_ipm_kernel      whole two-phase Mehrotra IPM on device, no host round trip per iteration
 *                                                <- d_ip2_res_mpc_hard_tv      (mpc_solvers/d_ip2_res_hard.c:116)
 *                      element-wise steps        <- mpc_solvers/c99/d_aux_ip_hard_lib4.c (lines cited inline)
 *                      residuals                 <- mpc_solvers/c99/d_res_ip_res_hard.c:39
 *   hb_fp64_probe      DFMA peak microbenchmark (roofline denominator, not part of the solver)
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "layout.h"
#include "ric_generic.cuh"
#include "ric_fast.cuh"
#include "ric_blk.cuh"
#include "ric_tree.cuh"
#include "ric_ipm_fast.cuh"

/* ------------------------------------------------------------------------------------------------ */
/* sweeps                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ hb_ctx hb_make_ctx(const hb_dims &d, double *smem_warp, int lane)
	{
	hb_ctx c;
	c.lane = lane;
	c.ldW = d.nxM | 1;
	int lsz = HB_EVEN(HB_TRI(d.nzM) + 2*d.nzM);
	c.bufA = smem_warp;
	c.bufB = c.bufA + lsz;
	c.sW = c.bufB + lsz;
	c.sV = c.sW + HB_EVEN(d.nzM*c.ldW);
	return c;
	}

__host__ __device__ inline int hb_smem_doubles_per_warp(int nzM, int nxM)
	{
	int lsz = HB_EVEN(HB_TRI(nzM) + 2*nzM);
	return 2*lsz + HB_EVEN(nzM*(nxM|1)) + 192;
	}

/* backward sweep n = N..0 ; factor of every stage is written to Lst (global) */
template<bool GRAD>
__device__ void hb_backward(const hb_ctx &c, const hb_dims &d, const double *in_inst, double *Lst,
		const double *bvec, const double *rqvec, const double *Qx, const double *qx, double *Pb)
	{
	double *cur = c.bufA, *prev = c.bufB;
	for(int n=d.N; n>=0; n--)
		{
		const hb_stage s = d.st[n];
		const int nu1 = (n<d.N) ? d.st[n+1].nu : 0;
		hb_stage_factor<GRAD>(c, s, nu1, in_inst, bvec, rqvec, Qx, qx, d.idxb, Pb, cur, prev);
		hb_copy(c, Lst + s.off_L, cur, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		double *t = cur; cur = prev; prev = t;
		__syncwarp();
		}
	}

/* forward sweep n = 0..N-1 */
__device__ void hb_forward(const hb_ctx &c, const hb_dims &d, const double *in_inst, const double *Lst,
		const double *lrow, const double *bvec, bool trs, double *ux, double *pi, bool compute_pi)
	{
	double *a = c.bufA, *b = c.bufB;
	{
	const hb_stage s0 = d.st[0];
	hb_copy(c, a, Lst + s0.off_L, HB_TRI(s0.nu+s0.nx) + 2*(s0.nu+s0.nx));
	}
	for(int n=0; n<d.N; n++)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		hb_copy(c, b, Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx));
		hb_load_BAbt(c, s, in_inst);
		__syncwarp();
		hb_stage_forward(c, s, s1, n, a, b, lrow, bvec, trs, ux, pi, compute_pi);
		double *t = a; a = b; b = t;
		}
	}

/* solve-only backward vector sweep; w is kept in ux */
__device__ void hb_trs_backward(const hb_ctx &c, const hb_dims &d, const double *in_inst, const double *Lst,
		const double *bvec, const double *rqvec, const double *qx, double *ux, double *Pb, bool compute_Pb)
	{
	const int lane = c.lane;
	{
	const hb_stage s = d.st[d.N];
	const int nux = s.nu+s.nx;
	for(int i=lane; i<nux; i+=32) ux[s.off_ux+i] = rqvec[s.off_ux+i];
	__syncwarp();
	if(qx!=nullptr) for(int j=lane; j<s.nb; j+=32) ux[s.off_ux+d.idxb[s.off_c+j]] += qx[s.off_c+j];
	__syncwarp();
	}
	for(int n=d.N-1; n>=0; n--)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		hb_copy(c, c.bufA, Lst + s.off_L, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		if(compute_Pb) hb_copy(c, c.bufB, Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx));
		hb_load_BAbt(c, s, in_inst);
		__syncwarp();
		hb_trs_stage_back(c, s, s1, n, c.bufA, c.bufB, bvec, rqvec, qx, d.idxb, ux, Pb, compute_Pb);
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* Riccati kernels                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
extern __shared__ double hb_smem[];

__global__ void hb_ric_sv_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ Pb, double *__restrict__ stash)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	double *Lst = stash + gw*d.L_stride;
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *in_inst = in + inst*d.in_stride;
		hb_backward<true>(c, d, in_inst, Lst, nullptr, nullptr, nullptr, nullptr, Pb!=nullptr ? Pb + inst*d.pi_stride : nullptr);
		__syncwarp();
		hb_forward(c, d, in_inst, Lst, nullptr, nullptr, false, ux + inst*d.ux_stride, pi + inst*d.pi_stride, true);
		__syncwarp();
		}
	}

__global__ void hb_ric_trf_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, double *__restrict__ L)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		hb_backward<false>(c, d, in + inst*d.in_stride, L + inst*d.L_stride, nullptr, nullptr, nullptr, nullptr, nullptr);
		__syncwarp();
		}
	}

/* solve with the stored factor; b and [r q] are taken from the instance block (new right-hand sides are
 * supplied by packing them into a copy of the block) */
__global__ void hb_ric_trs_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, const double *__restrict__ L,
		double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ work)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	/* per-slot work: rq (ux layout), b (pi layout), Pb (pi layout) */
	double *rq = work + gw*(d.ux_stride + 2*d.pi_stride), *bv = rq + d.ux_stride, *Pb = bv + d.pi_stride;
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *in_inst = in + inst*d.in_stride;
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			const int nux = s.nu+s.nx;
			for(int i=lane; i<nux; i+=32) rq[s.off_ux+i] = in_inst[s.off_RSQ+HB_TRI(nux)+i];
			for(int j=lane; j<s.nx1; j+=32) bv[s.off_pi+j] = in_inst[s.off_BAbt+nux*s.nx1+j];
			}
		__syncwarp();
		double *uxi = ux + inst*d.ux_stride;
		hb_trs_backward(c, d, in_inst, L + inst*d.L_stride, bv, rq, nullptr, uxi, Pb, true);
		hb_forward(c, d, in_inst, L + inst*d.L_stride, uxi, bv, true, uxi, pi + inst*d.pi_stride, true);
		__syncwarp();
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* IPM                                                                                               */
/* ------------------------------------------------------------------------------------------------ */
enum { CV_LB=0, CV_UB, CV_LAM_LO, CV_LAM_UP, CV_T_LO, CV_T_UP, CV_DLAM_LO, CV_DLAM_UP, CV_DT_LO, CV_DT_UP,
       CV_TINV_LO, CV_TINV_UP, CV_LAMT_LO, CV_LAMT_UP, CV_QXD /* "Qx": Hessian diagonal term */,
       CV_QXG /* "qx": gradient term */, CV_RD_LO, CV_RD_UP, CV_RM_LO, CV_RM_UP, CV_COUNT };

struct hb_ipm_ws
	{
	double *L;                               /* factor stash */
	double *dux, *res_q, *rq0;               /* ux layout */
	double *dpi, *Pb, *res_b, *b0;           /* pi layout */
	double *cv;                              /* CV_COUNT x nbp */
	int nbp;
	__device__ __forceinline__ double *v(int k) const { return cv + (size_t)k*nbp; }
	};

__host__ __device__ inline long long hb_ipm_work_doubles_(const hb_dims &d)
	{
	long long nbp = HB_EVEN(d.nbtot);
	return d.L_stride + 3*d.ux_stride + 4*d.pi_stride + (long long)CV_COUNT*nbp;
	}

__device__ __forceinline__ double hb_warp_min(double v)
	{
	for(int o=16; o>0; o>>=1) v = fmin(v, __shfl_xor_sync(HB_FULL, v, o));
	return v;
	}
__device__ __forceinline__ double hb_warp_max(double v)
	{
	for(int o=16; o>0; o>>=1) v = fmax(v, __shfl_xor_sync(HB_FULL, v, o));
	return v;
	}
/* fixed-order (deterministic) warp sum */
__device__ __forceinline__ double hb_warp_sum(double v)
	{
	for(int o=16; o>0; o>>=1) v += __shfl_xor_sync(HB_FULL, v, o);
	return v;
	}

/* bound part of the residuals: res_d, res_m and their sum (mpc_solvers/c99/d_res_ip_res_hard.c:39-319) */
__device__ __forceinline__ void hb_ipm_residuals_bounds(int lane, const hb_dims &d, const hb_ipm_ws &w, const double *ux, double &mu2, double &nd)
	{
	const double *lam_lo = w.v(CV_LAM_LO), *lam_up = w.v(CV_LAM_UP), *t_lo = w.v(CV_T_LO), *t_up = w.v(CV_T_UP);
	mu2 = 0.0; nd = 0.0;
	for(int cc=lane; cc<d.nbtot; cc+=32)
		{
		double u = ux[d.c_ux[cc]];
		double rdl = w.v(CV_LB)[cc] - u + t_lo[cc];
		double rdu = w.v(CV_UB)[cc] - u - t_up[cc];
		double rml = lam_lo[cc]*t_lo[cc], rmu = lam_up[cc]*t_up[cc];
		w.v(CV_RD_LO)[cc] = rdl; w.v(CV_RD_UP)[cc] = rdu;
		w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
		mu2 += rml + rmu;
		nd = fmax(nd, fmax(fabs(rdl), fabs(rdu)));
		}
	mu2 = hb_warp_sum(mu2);
	}

/* res_q, res_b, res_d, res_m and mu (mpc_solvers/c99/d_res_ip_res_hard.c:39-319); also returns the three
 * infinity norms used by the high-level wrapper on exit (interfaces/c/fortran_order_interface.c:616-652) */
__device__ void hb_ipm_residuals(const hb_ctx &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
		const double *ux, const double *pi, double *mu, double *norms)
	{
	const int lane = c.lane;
	double nq = 0.0, nb_ = 0.0, nd = 0.0, mu2 = 0.0;
	const double *lam_lo = w.v(CV_LAM_LO), *lam_up = w.v(CV_LAM_UP), *t_lo = w.v(CV_T_LO), *t_up = w.v(CV_T_UP);
	for(int cc=lane; cc<d.nbtot; cc+=32)
		{
		double u = ux[d.c_ux[cc]];
		double rdl = w.v(CV_LB)[cc] - u + t_lo[cc];
		double rdu = w.v(CV_UB)[cc] - u - t_up[cc];
		double rml = lam_lo[cc]*t_lo[cc], rmu = lam_up[cc]*t_up[cc];
		w.v(CV_RD_LO)[cc] = rdl; w.v(CV_RD_UP)[cc] = rdu;
		w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
		mu2 += rml + rmu;
		nd = fmax(nd, fmax(fabs(rdl), fabs(rdu)));
		}
	mu2 = hb_warp_sum(mu2);
	double *xs = c.sV;            /* ux_n */
	double *ps = c.sV + 64;       /* pi_n */
	for(int n=0; n<=d.N; n++)
		{
		const hb_stage s = d.st[n];
		const int nu = s.nu, nx = s.nx, nux = nu+nx, nx1 = s.nx1;
		double *H = c.bufA;
		hb_copy(c, H, in_inst + s.off_RSQ, HB_TRI(nux));
		if(nx1>0) hb_load_BAbt(c, s, in_inst);
		for(int i=lane; i<nux; i+=32) xs[i] = ux[s.off_ux+i];
		for(int j=lane; j<nx1; j+=32) ps[j] = pi[s.off_pi+j];
		/* rq = rq0 - pi_{n-1} (x part) + (lam_up - lam_lo)[idxb] */
		for(int i=lane; i<nux; i+=32)
			{
			double v = w.rq0[s.off_ux+i];
			if(n>0 && i>=nu) v -= pi[d.st[n-1].off_pi + (i-nu)];
			w.res_q[s.off_ux+i] = v;
			}
		__syncwarp();
		for(int j=lane; j<s.nb; j+=32)
			w.res_q[s.off_ux+d.idxb[s.off_c+j]] += -lam_lo[s.off_c+j] + lam_up[s.off_c+j];
		__syncwarp();
		for(int i=lane; i<nux; i+=32)
			{
			double acc = w.res_q[s.off_ux+i];
			const double *hi = H + HB_TRI(i);
			for(int j=0; j<=i; j++) acc += hi[j]*xs[j];
			for(int j=i+1; j<nux; j++) acc += H[HB_TRI(j)+i]*xs[j];
			const double *wr = c.sW + i*c.ldW;
			for(int j=0; j<nx1; j++) acc += wr[j]*ps[j];
			w.res_q[s.off_ux+i] = acc;
			nq = fmax(nq, fabs(acc));
			}
		if(nx1>0)
			{
			const hb_stage s1 = d.st[n+1];
			for(int j=lane; j<nx1; j+=32)
				{
				double acc = w.b0[s.off_pi+j] - ux[s1.off_ux+s1.nu+j];
				for(int i=0; i<nux; i++) acc += c.sW[i*c.ldW+j]*xs[i];
				w.res_b[s.off_pi+j] = acc;
				nb_ = fmax(nb_, fabs(acc));
				}
			}
		__syncwarp();
		}
	if(d.nbtot>0) *mu = mu2/(2.0*d.nbtot);
	if(norms!=nullptr)
		{
		norms[0] = hb_warp_max(nq); norms[1] = hb_warp_max(nb_); norms[2] = hb_warp_max(nd);
		}
	}

/* step length + dt, dlam.  RES = false: phase 1 (c99/d_aux_ip_hard_lib4.c:489-614) ; true: phase 2 (:1180-1313) */
template<bool RES>
__device__ __forceinline__ double hb_ipm_alpha(int lane_, const hb_dims &d, const hb_ipm_ws &w, const double *dux)
	{
	double alpha = 1.0;
	for(int cc=lane_; cc<d.nbtot; cc+=32)
		{
		double du = dux[d.c_ux[cc]];
		double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc], tl = w.v(CV_T_LO)[cc], tu = w.v(CV_T_UP)[cc];
		double dtl, dtu, dll, dlu;
		if(!RES)
			{
			dtl =  du - w.v(CV_LB)[cc] - tl;
			dtu = -du + w.v(CV_UB)[cc] - tu;
			dll = w.v(CV_DLAM_LO)[cc] - (w.v(CV_LAMT_LO)[cc]*dtl + ll);
			dlu = w.v(CV_DLAM_UP)[cc] - (w.v(CV_LAMT_UP)[cc]*dtu + lu);
			}
		else
			{
			dtl =  du - w.v(CV_RD_LO)[cc];
			dtu = -du + w.v(CV_RD_UP)[cc];
			dll = -w.v(CV_TINV_LO)[cc]*(ll*dtl + w.v(CV_RM_LO)[cc]);
			dlu = -w.v(CV_TINV_UP)[cc]*(lu*dtu + w.v(CV_RM_UP)[cc]);
			}
		w.v(CV_DT_LO)[cc] = dtl; w.v(CV_DT_UP)[cc] = dtu;
		w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
		if(-alpha*dll>ll) alpha = -ll/dll;
		if(-alpha*dlu>lu) alpha = -lu/dlu;
		if(-alpha*dtl>tl) alpha = -tl/dtl;
		if(-alpha*dtu>tu) alpha = -tu/dtu;
		}
	return hb_warp_min(alpha);
	}

/* mu_aff = mu_scal * sum (lam + a dlam)(t + a dt)   (c99/d_aux_ip_hard_lib4.c:715-770, :1453-1508) */
__device__ __forceinline__ double hb_ipm_mu_aff(int lane_, const hb_dims &d, const hb_ipm_ws &w, double alpha, double mu_scal)
	{
	double mu = 0.0;
	for(int cc=lane_; cc<d.nbtot; cc+=32)
		mu += (w.v(CV_LAM_LO)[cc] + alpha*w.v(CV_DLAM_LO)[cc])*(w.v(CV_T_LO)[cc] + alpha*w.v(CV_DT_LO)[cc])
		    + (w.v(CV_LAM_UP)[cc] + alpha*w.v(CV_DLAM_UP)[cc])*(w.v(CV_T_UP)[cc] + alpha*w.v(CV_DT_UP)[cc]);
	return hb_warp_sum(mu)*mu_scal;
	}

/* The IPM kernel is written once; the five sweeps over the horizon come from a policy: the run-time-size routines of
 * ric_generic.cuh, or the size-specialised, bulk-copy-pipelined ones of ric_ipm_fast.cuh. */
struct hb_sweeps_generic
	{
	typedef hb_ctx ctx_t;
	__device__ static __forceinline__ int smem_doubles(const hb_dims &d) { return hb_smem_doubles_per_warp(d.nzM, d.nxM); }
	__device__ static __forceinline__ long long L_doubles(const hb_dims &d) { return d.L_stride; }
	__device__ static __forceinline__ void init(ctx_t &c, const hb_dims &d, double *smem_warp, int lane) { c = hb_make_ctx(d, smem_warp, lane); }
	__device__ static __forceinline__ void backward(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *Qx, const double *qx)
		{ hb_backward<true>(c, d, in_inst, w.L, bv, rqv!=nullptr ? rqv : w.rq0, Qx, qx, w.Pb); }
	__device__ static __forceinline__ void forward_sv(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, double *ux, double *pi)
		{ hb_forward(c, d, in_inst, w.L, nullptr, bv, false, ux, pi, true); }
	__device__ static __forceinline__ void trs(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *qx)
		{
		hb_trs_backward(c, d, in_inst, w.L, bv, rqv, qx, w.dux, w.Pb, false);
		hb_forward(c, d, in_inst, w.L, w.dux, bv, true, w.dux, w.dpi, true);
		}
	__device__ static __forceinline__ void residuals(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *ux, const double *pi, double *mu, double *norms)
		{ hb_ipm_residuals(c, d, in_inst, w, ux, pi, mu, norms); }
	};

template<class C>
struct hb_sweeps_fast
	{
	typedef hbi_ctx<C> ctx_t;
	__device__ static __forceinline__ int smem_doubles(const hb_dims &) { return hbi_cfg<C>::PER_WARP; }
	__device__ static __forceinline__ long long L_doubles(const hb_dims &d) { return (long long)(d.N+1)*C::LBUF; }
	__device__ static __forceinline__ void init(ctx_t &c, const hb_dims &d, double *smem_warp, int lane) { c.init(smem_warp, lane, d); }
	__device__ static __forceinline__ void backward(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *Qx, const double *qx)
		{ hbi_backward<C>(c, d, in_inst, w.L, bv, rqv, Qx, qx, w.Pb); }
	__device__ static __forceinline__ void forward_sv(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, double *ux, double *pi)
		{ hbi_forward<C, false>(c, in_inst, w.L, bv, nullptr, ux, pi); }
	__device__ static __forceinline__ void trs(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *qx)
		{
		hbi_trs_backward<C>(c, d, in_inst, w.L, rqv, qx, w.Pb, w.dux);
		__syncwarp();
		hbi_forward<C, true>(c, in_inst, w.L, bv, w.dux, w.dux, w.dpi);
		}
	__device__ static __forceinline__ void residuals(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *ux, const double *pi, double *mu, double *norms)
		{
		double mu2, nd, nq = 0.0, nb_ = 0.0;
		hb_ipm_residuals_bounds(c.lane, d, w, ux, mu2, nd);
		__syncwarp();
		hbi_residuals<C>(c, d, in_inst, w.rq0, w.b0, w.v(CV_LAM_LO), w.v(CV_LAM_UP), ux, pi, w.res_q, w.res_b, nq, nb_);
		if(d.nbtot>0) *mu = mu2/(2.0*d.nbtot);
		if(norms!=nullptr) { norms[0] = hb_warp_max(nq); norms[1] = hb_warp_max(nb_); norms[2] = hb_warp_max(nd); }
		}
	};

template<class S>
__global__ void __launch_bounds__(256) hb_ipm_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *__restrict__ ux_all, double *__restrict__ pi_all,
		double *__restrict__ lam_all, double *__restrict__ t_all, double *__restrict__ info_all,
		double *__restrict__ work, long long work_stride, int *counter)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp;
	typename S::ctx_t c;
	S::init(c, d, hb_smem + (size_t)warp*S::smem_doubles(d), lane);
	hb_ipm_ws w;
	{
	double *p = work + gw*work_stride;
	w.L = p; p += S::L_doubles(d);
	w.dux = p; p += d.ux_stride; w.res_q = p; p += d.ux_stride; w.rq0 = p; p += d.ux_stride;
	w.dpi = p; p += d.pi_stride; w.Pb = p; p += d.pi_stride; w.res_b = p; p += d.pi_stride; w.b0 = p; p += d.pi_stride;
	w.cv = p; w.nbp = HB_EVEN(d.nbtot);
	}
	const int info_len = HB_IPM_INFO_HEAD + 5*k_max;
	const double thr0 = 0.1;

	for(;;)
		{
		/* dynamic instance queue: a warp that converges early simply takes the next instance, so the
		 * active set stays compact without a separate compaction pass */
		long long inst = 0;
		if(lane==0) inst = atomicAdd(counter, 1);
		inst = __shfl_sync(HB_FULL, inst, 0);
		if(inst>=n_inst) break;

		const double *in_inst = in + inst*d.in_stride;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		double *info = info_all + inst*info_len;
		double *stat = info + HB_IPM_INFO_HEAD;

		/* vectors taken from the instance block: rq0 = [r q], b0 = b, bounds */
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			const int nux = s.nu+s.nx;
			for(int i=lane; i<nux; i+=32) w.rq0[s.off_ux+i] = in_inst[s.off_RSQ+HB_TRI(nux)+i];
			for(int j=lane; j<s.nx1; j+=32) w.b0[s.off_pi+j] = in_inst[s.off_BAbt+nux*s.nx1+j];
			for(int j=lane; j<s.nb; j+=32)
				{
				w.v(CV_LB)[s.off_c+j] = in_inst[s.off_d+j];
				w.v(CV_UB)[s.off_c+j] = in_inst[s.off_d+s.nb+j];
				}
			}
		__syncwarp();

		int kk = 0, status = -1;
		double mu = 0.0, norms[3] = {0.0, 0.0, 0.0};

		if(d.nbtot==0)
			{
			/* no constraints: one Riccati solve (d_ip2_res_hard.c:430-450) */
			S::backward(c, d, in_inst, w, nullptr, nullptr, nullptr, nullptr);
			__syncwarp();
			S::forward_sv(c, d, in_inst, w, nullptr, ux, pi);
			__syncwarp();
			S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
			status = 0;
			}
		else
			{
			const double mu_scal = 1.0/(2.0*d.nbtot);
			double sigma = 0.0, alpha = 1.0, mu_aff;
			/* init (c99/d_aux_ip_hard_lib4.c:43-149) */
			if(!warm_start) for(long long i=lane; i<d.ux_stride; i+=32) ux[i] = 0.0;
			for(long long i=lane; i<d.pi_stride; i+=32) pi[i] = 0.0;
			__syncwarp();
			for(int cc=lane; cc<d.nbtot; cc+=32)
				{
				const int iu = d.c_ux[cc];
				double lb = w.v(CV_LB)[cc], ub = w.v(CV_UB)[cc], u = ux[iu];
				double tl = -lb + u, tu = ub - u;
				if(tl<thr0)
					{
					if(tu<thr0) { ux[iu] = (-ub + lb)*0.5; tl = thr0; tu = thr0; }
					else { tl = thr0; ux[iu] = lb + thr0; }
					}
				else if(tu<thr0) { tu = thr0; ux[iu] = ub - thr0; }
				w.v(CV_T_LO)[cc] = tl; w.v(CV_T_UP)[cc] = tu;
				w.v(CV_LAM_LO)[cc] = mu0/tl; w.v(CV_LAM_UP)[cc] = mu0/tu;
				}
			__syncwarp();
			mu = mu0;
			const double mu_tol_low = mu_tol<1e-5 ? 1e-5 : mu_tol;

			/* ---------- phase 1 (d_ip2_res_hard.c:503-718) ---------- */
			while(kk<k_max && mu>mu_tol_low && alpha>=alpha_min)
				{
				/* update_hessian, sigma_mu = 0 (c99/d_aux_ip_hard_lib4.c:217-383) */
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double til = 1.0/w.v(CV_T_LO)[cc], tiu = 1.0/w.v(CV_T_UP)[cc];
					double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc];
					double ltl = ll*til, ltu = lu*tiu;
					double dll = til*0.0, dlu = tiu*0.0;
					w.v(CV_TINV_LO)[cc] = til; w.v(CV_TINV_UP)[cc] = tiu;
					w.v(CV_LAMT_LO)[cc] = ltl; w.v(CV_LAMT_UP)[cc] = ltu;
					w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
					w.v(CV_QXD)[cc] = ltl + ltu;
					w.v(CV_QXG)[cc] = lu - ltu*w.v(CV_UB)[cc] + dlu - ll - ltl*w.v(CV_LB)[cc] - dll;
					}
				__syncwarp();
				HBF_STAMP(300);
				S::backward(c, d, in_inst, w, nullptr, nullptr, w.v(CV_QXD), w.v(CV_QXG));
				__syncwarp();
				HBF_STAMP(301);
				S::forward_sv(c, d, in_inst, w, nullptr, w.dux, w.dpi);
				HBF_STAMP(302);
				__syncwarp();
				alpha = hb_ipm_alpha<false>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
				alpha *= 0.995;
				mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
				if(lane==0) stat[5*kk+2] = mu_aff;
				sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
				{
				/* update_gradient (c99/d_aux_ip_hard_lib4.c:387-485) */
				const double sm = sigma*mu;
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double dll = w.v(CV_TINV_LO)[cc]*(sm - w.v(CV_DLAM_LO)[cc]*w.v(CV_DT_LO)[cc]);
					double dlu = w.v(CV_TINV_UP)[cc]*(sm - w.v(CV_DLAM_UP)[cc]*w.v(CV_DT_UP)[cc]);
					w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
					w.v(CV_QXG)[cc] += dlu - dll;
					}
				}
				__syncwarp();
				HBF_STAMP(303);
				S::trs(c, d, in_inst, w, w.b0, w.rq0, w.v(CV_QXG));
				HBF_STAMP(304);
				__syncwarp();
				alpha = hb_ipm_alpha<false>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
				alpha *= 0.995;
				/* update_var (c99/d_aux_ip_hard_lib4.c:618-711) */
				for(long long i=lane; i<d.ux_stride; i+=32) ux[i] += alpha*(w.dux[i] - ux[i]);
				for(long long i=lane; i<d.pi_stride; i+=32) pi[i] += alpha*(w.dpi[i] - pi[i]);
				double ms = 0.0;
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double ll = w.v(CV_LAM_LO)[cc] + alpha*w.v(CV_DLAM_LO)[cc];
					double lu = w.v(CV_LAM_UP)[cc] + alpha*w.v(CV_DLAM_UP)[cc];
					double tl = w.v(CV_T_LO)[cc] + alpha*w.v(CV_DT_LO)[cc];
					double tu = w.v(CV_T_UP)[cc] + alpha*w.v(CV_DT_UP)[cc];
					w.v(CV_LAM_LO)[cc] = ll; w.v(CV_LAM_UP)[cc] = lu; w.v(CV_T_LO)[cc] = tl; w.v(CV_T_UP)[cc] = tu;
					ms += ll*tl + lu*tu;
					}
				mu = hb_warp_sum(ms)*mu_scal;
				if(lane==0) stat[5*kk+4] = mu;
				kk++;
				__syncwarp();
				}

			/* ---------- phase 2 (d_ip2_res_hard.c:756-1273) ---------- */
			S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
			__syncwarp();
			while(kk<k_max && mu>mu_tol && alpha>=alpha_min)
				{
				/* update_hessian_gradient_res (c99/d_aux_ip_hard_lib4.c:954-1078) */
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double til = 1.0/w.v(CV_T_LO)[cc], tiu = 1.0/w.v(CV_T_UP)[cc];
					double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc];
					w.v(CV_TINV_LO)[cc] = til; w.v(CV_TINV_UP)[cc] = tiu;
					w.v(CV_QXD)[cc] = til*ll + tiu*lu;
					w.v(CV_QXG)[cc] = til*(w.v(CV_RM_LO)[cc] - ll*w.v(CV_RD_LO)[cc]) - tiu*(w.v(CV_RM_UP)[cc] + lu*w.v(CV_RD_UP)[cc]);
					}
				__syncwarp();
				HBF_STAMP(300);
				S::backward(c, d, in_inst, w, w.res_b, w.res_q, w.v(CV_QXD), w.v(CV_QXG));
				__syncwarp();
				HBF_STAMP(301);
				S::forward_sv(c, d, in_inst, w, w.res_b, w.dux, w.dpi);
				HBF_STAMP(302);
				__syncwarp();
				alpha = hb_ipm_alpha<true>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
				alpha *= 0.995;
				mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
				if(lane==0) stat[5*kk+2] = mu_aff;
				sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
				{
				/* centering correction + update_gradient_res (c99/d_aux_ip_hard_lib4.c:1512-1546, :1550-1639) */
				const double sm = sigma*mu;
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double rml = w.v(CV_RM_LO)[cc] + (w.v(CV_DT_LO)[cc]*w.v(CV_DLAM_LO)[cc] - sm);
					double rmu = w.v(CV_RM_UP)[cc] + (w.v(CV_DT_UP)[cc]*w.v(CV_DLAM_UP)[cc] - sm);
					w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
					w.v(CV_QXG)[cc] = w.v(CV_TINV_LO)[cc]*(rml - w.v(CV_LAM_LO)[cc]*w.v(CV_RD_LO)[cc])
					                  - w.v(CV_TINV_UP)[cc]*(rmu + w.v(CV_LAM_UP)[cc]*w.v(CV_RD_UP)[cc]);
					}
				}
				__syncwarp();
				HBF_STAMP(303);
				S::trs(c, d, in_inst, w, w.res_b, w.res_q, w.v(CV_QXG));
				HBF_STAMP(304);
				__syncwarp();
				alpha = hb_ipm_alpha<true>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
				alpha *= 0.995;
				/* backup_update_var_res (c99/d_aux_ip_hard_lib4.c:1382-1449) */
				for(long long i=lane; i<d.ux_stride; i+=32) ux[i] += alpha*w.dux[i];
				for(long long i=lane; i<d.pi_stride; i+=32) pi[i] += alpha*w.dpi[i];
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					w.v(CV_LAM_LO)[cc] += alpha*w.v(CV_DLAM_LO)[cc]; w.v(CV_LAM_UP)[cc] += alpha*w.v(CV_DLAM_UP)[cc];
					w.v(CV_T_LO)[cc] += alpha*w.v(CV_DT_LO)[cc]; w.v(CV_T_UP)[cc] += alpha*w.v(CV_DT_UP)[cc];
					}
				__syncwarp();
				HBF_STAMP(305);
				S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
				HBF_STAMP(306);
				if(lane==0) stat[5*kk+4] = mu;
				kk++;
				__syncwarp();
				}
			if(mu<=mu_tol) status = 0;
			else if(kk>=k_max) status = 1;
			else if(alpha<alpha_min) status = 2;
			else status = -1;
			}

		/* results: lam, t as [lower(nb) upper(nb)] per stage (interfaces/c/fortran_order_interface.c:662-671) */
		double *lam = lam_all + inst*2*(long long)d.nbtot, *tt = t_all + inst*2*(long long)d.nbtot;
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			for(int j=lane; j<s.nb; j+=32)
				{
				lam[2*s.off_c+j] = w.v(CV_LAM_LO)[s.off_c+j]; lam[2*s.off_c+s.nb+j] = w.v(CV_LAM_UP)[s.off_c+j];
				tt[2*s.off_c+j] = w.v(CV_T_LO)[s.off_c+j]; tt[2*s.off_c+s.nb+j] = w.v(CV_T_UP)[s.off_c+j];
				}
			}
		if(lane==0)
			{
			info[0] = (double)kk; info[1] = (double)status;
			info[2] = norms[0]; info[3] = norms[1]; info[4] = norms[2]; info[5] = mu;
			}
		__syncwarp();
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* scenario tree                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
__global__ void hb_tree_kernel(hb_tdims d, long long n_trees, const double *__restrict__ in, double *__restrict__ ux_all,
		double *__restrict__ pi_all, double *__restrict__ L_all, int mode, int seg_lo, int seg_hi)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c;
	{
	double *smem_warp = hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM);
	int lsz = HB_EVEN(HB_TRI(d.nzM) + 2*d.nzM);
	c.lane = lane; c.ldW = d.nxM | 1;
	c.bufA = smem_warp; c.bufB = c.bufA + lsz; c.sW = c.bufB + lsz; c.sV = c.sW + HB_EVEN(d.nzM*c.ldW);
	}
	const int nseg = seg_hi - seg_lo;
	const long long n_items = n_trees*nseg;
	for(long long item=gw; item<n_items; item+=tw)
		{
		const long long t = item/nseg;
		const int seg = seg_lo + (int)(item - t*nseg);
		const double *in_tree = in + t*d.in_stride;
		double *Lt = L_all + t*d.L_stride, *ux = ux_all + t*d.ux_stride, *pi = pi_all + t*d.pi_stride;
		const int s0 = d.seg_start[seg], s1 = d.seg_start[seg+1];
		if(mode==0 || mode==2)
			for(int q=s1-1; q>=s0; q--)
				{
				hb_tree_node_factor(c, d.tn, d.seg_nodes[q], in_tree, Lt, c.bufA, c.bufB);
				__syncwarp();
				}
		if(mode==1 || mode==2)
			for(int q=s0; q<s1; q++)
				{
				hb_tree_node_forward(c, d.tn, d.seg_nodes[q], in_tree, Lt, ux, pi, c.bufA, c.bufB);
				__syncwarp();
				}
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* FP64 peak probe                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
__global__ void hb_fp64_probe(double *out, int iters)
	{
	double a0 = threadIdx.x*1e-9, a1 = a0+1, a2 = a0+2, a3 = a0+3, a4 = a0+4, a5 = a0+5, a6 = a0+6, a7 = a0+7;
	const double x = 1.0000001, y = 1e-9;
	for(int i=0; i<iters; i++)
		{
		a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
		a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
		}
	out[blockIdx.x*blockDim.x+threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7;
	}

/* ------------------------------------------------------------------------------------------------ */
/* launchers (C ABI)                                                                                 */
/* ------------------------------------------------------------------------------------------------ */
#define HB_CK(x) do { cudaError_t e_ = (x); if(e_!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

extern "C" int hb_smem_bytes_per_warp(const hb_dims *d)
	{
	return (int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	}

extern "C" int hb_smem_bytes_per_warp_sz(int nzM, int nxM) { return (int)sizeof(double)*hb_smem_doubles_per_warp(nzM, nxM); }

extern "C" long long hb_ipm_work_doubles(const hb_dims *d) { return hb_ipm_work_doubles_(*d); }

extern "C" int hb_device_sm_count(int device)
	{
	int n = 0;
	if(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device)!=cudaSuccess) return -1;
	return n;
	}

template<typename K>
static int hb_prep(K kernel, int smem)
	{
	if(smem>227*1024) { fprintf(stderr, "hpmpc_b200: stage too large for shared memory (%d bytes)\n", smem); return -1; }
	HB_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	return 0;
	}

extern "C" int hb_launch_ric_sv(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi, double *Pb,
		double *stash, int n_slots, int grid, int warps, void *stream)
	{
	if(d->nzM>64) { fprintf(stderr, "hpmpc_b200: nu+nx+1 > 64 not supported\n"); return -2; }
	if(grid*warps>n_slots) return -3;
	int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_ric_sv_kernel, smem)) return -1;
	hb_ric_sv_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, ux, pi, Pb, stash);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_ric_trf(const hb_dims *d, long long n_inst, const double *in, double *L, int grid, int warps, void *stream)
	{
	if(d->nzM>64) return -2;
	int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_ric_trf_kernel, smem)) return -1;
	hb_ric_trf_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, L);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_ric_trs(const hb_dims *d, long long n_inst, const double *in, const double *L, double *ux, double *pi,
		double *work, int n_slots, int grid, int warps, void *stream)
	{
	if(d->nzM>64) return -2;
	if(grid*warps>n_slots) return -3;
	int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_ric_trs_kernel, smem)) return -1;
	hb_ric_trs_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, L, ux, pi, work);
	HB_CK(cudaGetLastError());
	return 0;
	}

/* size-specialised IPM sweeps (ric_ipm_fast.cuh): one warp per instance, x0 eliminated, uniform (nx, nu) */
typedef hbf_cfg<24, 11, 32> hbi_v0;    /* BASELINE config 3 */
typedef hbf_cfg<12, 5, 32> hbi_v1;     /* config-2 sizes with bounds */
typedef hbf_cfg<8, 3, 32> hbi_v2;      /* the reference's own IPM test size (test_d_ip_hard.c) */
#define HBI_NVAR 3
static const int hbi_shapes[HBI_NVAR][2] = { {24, 11}, {12, 5}, {8, 3} };

extern "C" int hb_ipm_fast_variant(int N, const int *nx, const int *nu, int nbtot)
	{
	if(getenv("HPMPC_B200_NO_FAST_IPM")!=NULL || nbtot<=0) return -1;
	for(int id=0; id<HBI_NVAR; id++)
		{
		int ok = (nx[0]==0) && N>=3;
		for(int n=0; n<N && ok; n++) ok = (nu[n]==hbi_shapes[id][1]) && (n==0 || nx[n]==hbi_shapes[id][0]);
		ok = ok && nx[N]==hbi_shapes[id][0];
		if(ok) return id;
		}
	return -1;
	}

template<class C> static void hbi_info(int N, int *smem_warp, long long *L_doubles)
	{ *smem_warp = (int)sizeof(double)*hbi_cfg<C>::PER_WARP; *L_doubles = (long long)(N+1)*C::LBUF; }

extern "C" int hb_ipm_fast_info(int id, int N, int *smem_warp, long long *L_doubles)
	{
	switch(id)
		{
		case 0: hbi_info<hbi_v0>(N, smem_warp, L_doubles); return 0;
		case 1: hbi_info<hbi_v1>(N, smem_warp, L_doubles); return 0;
		case 2: hbi_info<hbi_v2>(N, smem_warp, L_doubles); return 0;
		}
	return -1;
	}

/* doubles of per-slot work area; L_doubles = size of the factor stash of the variant in use */
extern "C" long long hb_ipm_work_doubles2(const hb_dims *d, long long L_doubles)
	{
	return hb_ipm_work_doubles_(*d) - d->L_stride + L_doubles;
	}

template<class S> static int hb_launch_ipm_t(int smem, const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int grid, int warps, int *counter, cudaStream_t st)
	{
	if(hb_prep(hb_ipm_kernel<S>, smem)) return -1;
	if(getenv("HPMPC_B200_VERBOSE"))
		{
		int nb = 0; cudaFuncAttributes fa;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hb_ipm_kernel<S>, warps*32, smem);
		cudaFuncGetAttributes(&fa, hb_ipm_kernel<S>);
		fprintf(stderr, "hpmpc_b200: ipm kernel: grid %d x %d threads, %d B dynamic + %zu B static smem, %d regs, %zu B local, %d CTAs/SM\n",
			grid, warps*32, smem, fa.sharedSizeBytes, fa.numRegs, fa.localSizeBytes, nb);
		}
	HB_CK(cudaMemsetAsync(counter, 0, sizeof(int), st));
	hb_ipm_kernel<S><<<grid, warps*32, smem, st>>>(*d, n_inst, in, k_max, mu0, mu_tol, alpha_min, warm_start,
			ux, pi, lam, t, info, work, work_stride, counter);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_ipm(const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int n_slots, int grid, int warps, int *counter, int fast_id, void *stream)
	{
	if(d->nzM>64) return -2;
	if(grid*warps>n_slots || warps>8) return -3;
	cudaStream_t st = (cudaStream_t)stream;
#define HB_IPM_ARGS d, n_inst, in, k_max, mu0, mu_tol, alpha_min, warm_start, ux, pi, lam, t, info, work, work_stride, grid, warps, counter, st
	switch(fast_id)
		{
		case 0: return hb_launch_ipm_t<hb_sweeps_fast<hbi_v0> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v0>::PER_WARP, HB_IPM_ARGS);
		case 1: return hb_launch_ipm_t<hb_sweeps_fast<hbi_v1> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v1>::PER_WARP, HB_IPM_ARGS);
		case 2: return hb_launch_ipm_t<hb_sweeps_fast<hbi_v2> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v2>::PER_WARP, HB_IPM_ARGS);
		}
	return hb_launch_ipm_t<hb_sweeps_generic>(warps*hb_smem_bytes_per_warp(d), HB_IPM_ARGS);
#undef HB_IPM_ARGS
	}

extern "C" int hb_launch_tree(const hb_tdims *d, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int grid, int warps, void *stream)
	{
	if(d->nzM>64) { fprintf(stderr, "hpmpc_b200: tree: nu+nx+1 > 64 not supported\n"); return -2; }
	if(seg_hi<=seg_lo || n_trees<=0) return 0;
	int smem = warps*(int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	if(hb_prep(hb_tree_kernel, smem)) return -1;
	hb_tree_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" double hb_fp64_peak_probe(int device, int iters, void *stream)
	{
	int sms = hb_device_sm_count(device);
	if(sms<=0) return -1.0;
	const int threads = 512, blocks = sms*4;
	double *out = nullptr;
	if(cudaMalloc(&out, sizeof(double)*threads*blocks)!=cudaSuccess) return -1.0;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaStream_t st = (cudaStream_t)stream;
	hb_fp64_probe<<<blocks, threads, 0, st>>>(out, iters);
	cudaEventRecord(e0, st);
	hb_fp64_probe<<<blocks, threads, 0, st>>>(out, iters);
	cudaEventRecord(e1, st);
	cudaEventSynchronize(e1);
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(out);
	double flops = 2.0*8.0*(double)iters*threads*blocks;
	return flops/(ms*1e-3)/1e12;
	}

/* ------------------------------------------------------------------------------------------------ */
/* size-specialised variants (ric_blk.cuh: register-blocked; ric_fast.cuh: one row per lane)          */
/* ------------------------------------------------------------------------------------------------ */
#ifndef HBK_V0_G
#define HBK_V0_G 8
#define HBK_V0_R 2
#endif
typedef hbk_cfg<12, 5, HBK_V0_G, HBK_V0_R> hbk_v0;   /* BASELINE config 2: four instances per warp, two rows per lane */
typedef hbk_cfg<8, 3, 4, 3> hbk_v1;    /* the reference's own test size (test_d_ip_hard.c): eight instances per warp */
typedef hbk_cfg<4, 2, 4, 2> hbk_v2;    /* eight instances per warp */
typedef hbf_cfg<24, 11, 32> hbf_v3;    /* BASELINE config 3 shape: one instance per warp, 4 column-owned rows */   /* (hbk_cfg<24,11,16,2> was tried: 1.7 KB of local-memory stack, 9x slower) */
typedef hbf_cfg<12, 5, 16> hbf_v0;     /* one-row-per-lane predecessors, kept for A/B runs (HPMPC_B200_FAST_GEN=1) */
typedef hbf_cfg<8, 3, 16> hbf_v1;
typedef hbf_cfg<4, 2, 8> hbf_v2;
#define HBF_NVAR 7
static const int hbf_shapes[HBF_NVAR][2] = { {12, 5}, {8, 3}, {4, 2}, {24, 11}, {12, 5}, {8, 3}, {4, 2} };

/* a pattern qualifies when x0 is eliminated (nx[0] = 0) and every other stage has the variant's (nx, nu) */
extern "C" int hb_fast_variant(int N, const int *nx, const int *nu)
	{
	const char *gen = getenv("HPMPC_B200_FAST_GEN");
	const int first = (gen!=NULL && gen[0]=='1') ? 3 : 0;
	for(int id=first; id<HBF_NVAR; id++)
		{
		int ok = (nx[0]==0) && N>=3;
		for(int n=0; n<N && ok; n++) ok = (nu[n]==hbf_shapes[id][1]) && (n==0 || nx[n]==hbf_shapes[id][0]);
		ok = ok && nx[N]==hbf_shapes[id][0];
		if(ok) return id;
		}
	return -1;
	}

template<class C> static void hbf_info(int N, int *ipw, int *smem_warp, long long *stash_per_inst)
	{
	*ipw = C::IPW; *smem_warp = (int)sizeof(double)*C::PER_WARP; *stash_per_inst = (long long)(N+1)*C::LBUF;
	}
#ifdef HBK_EXPERIMENTAL_V2
static int hbk_use_v2() { const char *e = getenv("HPMPC_B200_BLK_V2"); return e!=NULL && e[0]=='1'; }
#else
static int hbk_use_v2() { return 0; }
template<class C> struct hbk2_cfg { static constexpr int PER_WARP = C::PER_WARP; };
#define hbk2_ric_sv_kernel hbk_ric_sv_kernel
#endif
template<class C> static void hbk_info(int N, int *ipw, int *smem_warp, long long *stash_per_inst)
	{
	*ipw = C::IPW; *smem_warp = (int)sizeof(double)*(hbk_use_v2() ? hbk2_cfg<C>::PER_WARP : C::PER_WARP); *stash_per_inst = (long long)(N+1)*C::SB;
	}

extern "C" int hb_fast_info(int id, int N, int *ipw, int *smem_warp, long long *stash_per_inst)
	{
	switch(id)
		{
		case 0: hbk_info<hbk_v0>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 1: hbk_info<hbk_v1>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 2: hbk_info<hbk_v2>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 3: hbf_info<hbf_v3>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 4: hbf_info<hbf_v0>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 5: hbf_info<hbf_v1>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 6: hbf_info<hbf_v2>(N, ipw, smem_warp, stash_per_inst); return 0;
		}
	return -1;
	}

template<class C> static int hbf_launch(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, cudaStream_t st)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbf_ric_sv_kernel<C>, smem)) return -1;
	hbf_ric_sv_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_inst, in, ux, pi, stash);
	HB_CK(cudaGetLastError());
	return 0;
	}
/* The factor stash is scratch that every warp slot rewrites for each instance it solves (written in the backward sweep,
 * read back in the forward sweep).  A persisting-L2 access window over it keeps a fraction of its lines resident, so that
 * fraction of the stash traffic never reaches HBM (HPMPC_B200_L2_PERSIST=0 disables; value = hit ratio in percent). */
static int hb_stash_window(cudaLaunchAttribute *attr, const void *stash, size_t stash_bytes)
	{
	static int inited = 0, max_persist = 0, max_window = 0, pct = -1;
	if(!inited)
		{
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
		const char *e = getenv("HPMPC_B200_L2_PERSIST");
		pct = e ? atoi(e) : 0;          /* opt-in: measured neutral (16-40 MiB set aside) to harmful (79 MiB) on B200 */
		{ const char *m = getenv("HPMPC_B200_L2_MB"); if(m && atoi(m)>0 && ((size_t)atoi(m)<<20)<(size_t)max_persist) max_persist = atoi(m)<<20; }
		if(pct!=0 && max_persist>0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
		if(getenv("HPMPC_B200_VERBOSE")) fprintf(stderr, "hpmpc_b200: persisting L2 max %d MiB, window max %d MiB\n", max_persist>>20, max_window>>20);
		inited = 1;
		}
	if(pct==0 || max_persist<=0 || max_window<=0 || stash_bytes==0) return 0;
	size_t win = stash_bytes<(size_t)max_window ? stash_bytes : (size_t)max_window;
	double ratio = pct>0 ? pct/100.0 : 0.9*(double)max_persist/(double)win;
	if(ratio>1.0) ratio = 1.0;
	attr->id = cudaLaunchAttributeAccessPolicyWindow;
	attr->val.accessPolicyWindow.base_ptr = (void*)stash;
	attr->val.accessPolicyWindow.num_bytes = win;
	attr->val.accessPolicyWindow.hitRatio = (float)ratio;
	attr->val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	attr->val.accessPolicyWindow.missProp = getenv("HPMPC_B200_L2_MISS_NORMAL") ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
	return 1;
	}

template<class C> static int hbk_launch(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, cudaStream_t st)
	{
	const int v2 = hbk_use_v2();
	int smem = warps*(int)sizeof(double)*(v2 ? hbk2_cfg<C>::PER_WARP : C::PER_WARP);
	if(v2 ? hb_prep(hbk2_ric_sv_kernel<C>, smem) : hb_prep(hbk_ric_sv_kernel<C>, smem)) return -1;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = dim3(grid); cfg.blockDim = dim3(warps*32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	cfg.attrs = attr;
	cfg.numAttrs = hb_stash_window(&attr[0], stash, sizeof(double)*(size_t)grid*warps*C::IPW*(size_t)(d->N+1)*C::SB);
	if(v2) { HB_CK(cudaLaunchKernelEx(&cfg, hbk2_ric_sv_kernel<C>, *d, n_inst, in, ux, pi, stash)); }
	else { HB_CK(cudaLaunchKernelEx(&cfg, hbk_ric_sv_kernel<C>, *d, n_inst, in, ux, pi, stash)); }
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_ric_sv_fast(int id, const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, void *stream)
	{
	cudaStream_t st = (cudaStream_t)stream;
	switch(id)
		{
		case 0: return hbk_launch<hbk_v0>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 1: return hbk_launch<hbk_v1>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 2: return hbk_launch<hbk_v2>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 3: return hbf_launch<hbf_v3>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 4: return hbf_launch<hbf_v0>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 5: return hbf_launch<hbf_v1>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 6: return hbf_launch<hbf_v2>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		}
	return -2;
	}

/* ------------------------------------------------------------------------------------------------ */
/* size-specialised tails of scenario trees                                                          */
/* ------------------------------------------------------------------------------------------------ */
extern "C" int hb_tail_variant(int nx, int nu)
	{
	for(int id=0; id<3; id++) if(hbf_shapes[id][0]==nx && hbf_shapes[id][1]==nu) return id;
	return -1;
	}

template<class C> static void hbk_tail_info_t(int *ipw, int *smem_warp, int *image_doubles)
	{ *ipw = C::IPW; *smem_warp = (int)sizeof(double)*C::PER_WARP; *image_doubles = C::SB; }

extern "C" int hb_tail_info(int id, int *ipw, int *smem_warp, int *image_doubles)
	{
	switch(id)
		{
		case 0: hbk_tail_info_t<hbk_v0>(ipw, smem_warp, image_doubles); return 0;
		case 1: hbk_tail_info_t<hbk_v1>(ipw, smem_warp, image_doubles); return 0;
		case 2: hbk_tail_info_t<hbk_v2>(ipw, smem_warp, image_doubles); return 0;
		}
	return -1;
	}

template<class C> static int hbk_tail_launch(const hb_tdims *d, const hb_tail_tab *tab, long long n_trees, const double *in, double *ux,
		double *pi, double *L, int mode, int tail_lo, int tail_hi, int grid, int warps, cudaStream_t st)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbk_tail_kernel<C>, smem)) return -1;
	hbk_tail_kernel<C><<<grid, warps*32, smem, st>>>(*tab, n_trees, d->in_stride, d->ux_stride, d->pi_stride, d->L_stride, in, ux, pi, L,
			mode, tail_lo, tail_hi);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_tail(int id, const hb_tdims *d, const hb_tail_tab *tab, long long n_trees, const double *in, double *ux, double *pi,
		double *L, int mode, int tail_lo, int tail_hi, int grid, int warps, void *stream)
	{
	if(tail_hi<=tail_lo || n_trees<=0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	switch(id)
		{
		case 0: return hbk_tail_launch<hbk_v0>(d, tab, n_trees, in, ux, pi, L, mode, tail_lo, tail_hi, grid, warps, st);
		case 1: return hbk_tail_launch<hbk_v1>(d, tab, n_trees, in, ux, pi, L, mode, tail_lo, tail_hi, grid, warps, st);
		case 2: return hbk_tail_launch<hbk_v2>(d, tab, n_trees, in, ux, pi, L, mode, tail_lo, tail_hi, grid, warps, st);
		}
	return -2;
	}

template<class C> static int hbk_top_launch(const hb_tdims *d, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int first, int grid, int warps, cudaStream_t st)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbk_top_kernel<C>, smem)) return -1;
	hbk_top_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_top(int id, const hb_tdims *d, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int first, int grid, int warps, void *stream)
	{
	if(seg_hi<=seg_lo || n_trees<=0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	switch(id)
		{
		case 0: return hbk_top_launch<hbk_v0>(d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, grid, warps, st);
		case 1: return hbk_top_launch<hbk_v1>(d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, grid, warps, st);
		case 2: return hbk_top_launch<hbk_v2>(d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, grid, warps, st);
		}
	return -2;
	}

#ifdef HBF_TIMING
extern "C" int hb_debug_timing(long long *d_buf)
	{
	int zero = 0;
	static int gen = 1000;
	gen++;
	HB_CK(cudaMemcpyToSymbol(hbf_dbg_gen, &gen, sizeof(int)));
	HB_CK(cudaMemcpyToSymbol(hbf_dbg, &d_buf, sizeof(d_buf)));
	HB_CK(cudaMemcpyToSymbol(hbf_dbg_n, &zero, sizeof(int)));
	return 0;
	}
#endif
