/*
 * ipm_sweeps.cuh -- what the IPM kernels share: the chain residual routine, the vectors taken from / returned to an instance
 * block, and the sweep POLICIES the IPM is written against (generic run-time-size sweeps, size-specialised pipelined sweeps).
 * Included by ipm_kernels.cu (the fused one-kernel IPM, the KKT re-solve) and cipm_kernels.cu (the multi-kernel IPM driver with
 * active-set compaction).  Reference lines are cited at each function.
 */
#pragma once
#include "layout.h"
#include "ric_sweeps.cuh"
#include "ric_fast.cuh"
#include "ric_ipm_fast.cuh"

extern "C" int hb_smem_bytes_per_warp(const hb_dims *d);

#include "ipm_elem.cuh"

/* res_q, res_b, res_d, res_m and mu (mpc_solvers/c99/d_res_ip_res_hard.c:39-319); also returns the three
 * infinity norms used by the high-level wrapper on exit (interfaces/c/fortran_order_interface.c:616-652) */
static __device__ void hb_ipm_residuals(const hb_ctx &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
		const double *ux, const double *pi, double *mu, double *norms)
	{
	const int lane = c.lane;
	double nq = 0.0, nb_ = 0.0, nd = 0.0, mu2 = 0.0;
	const double *lam_lo = w.v(CV_LAM_LO), *lam_up = w.v(CV_LAM_UP), *t_lo = w.v(CV_T_LO), *t_up = w.v(CV_T_UP);
	hb_gen_values(lane, d, in_inst, w, ux);
	for(int cc=lane; cc<d.nbtot; cc+=32)
		{
		double u = hb_cval(d, w, ux, cc);
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
		hb_g2s(lane, H, in_inst + s.off_RSQ, HB_TRI(nux));
		if(nx1>0) hb_load_BAbt_async(c, s, in_inst);
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
		if(s.ng>0)
			{
			/* general constraints: + [D C]' (lam_ug - lam_lg)  (mpc_solvers/c99/d_res_ip_res_hard.c:120-144 twin of
			 * d_res_ip_res_hard_libstr.c:120-144) */
			const double *G = in_inst + s.off_DCt;
			const int cg = s.off_c + s.nb;
			for(int i=lane; i<nux; i+=32)
				{
				double acc = w.res_q[s.off_ux+i];
				for(int j=0; j<s.ng; j++) acc += G[i*s.ng+j]*(lam_up[cg+j] - lam_lo[cg+j]);
				w.res_q[s.off_ux+i] = acc;
				}
			__syncwarp();
			}
		hb_g2s_wait();
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

/* vectors taken from the instance block of a chain: rq0 = [r q], b0 = b, bounds */
__device__ __forceinline__ void hb_ipm_extract_chain(int lane, const hb_dims &d, const double *__restrict__ in_inst, const hb_ipm_ws &w)
	{
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
		for(int j=lane; j<s.ng; j+=32)
			{
			w.v(CV_LB)[s.off_c+s.nb+j] = in_inst[s.off_dg+j];
			w.v(CV_UB)[s.off_c+s.nb+j] = in_inst[s.off_dg+s.ng+j];
			}
		}
	}

/* results: lam, t as [lower(nb) upper(nb)] per stage (interfaces/c/fortran_order_interface.c:662-671) */
__device__ __forceinline__ void hb_ipm_emit_chain(int lane, const hb_dims &d, const hb_ipm_ws &w, double *lam, double *tt)
	{
	for(int n=0; n<=d.N; n++)
		{
		const hb_stage s = d.st[n];
		for(int j=lane; j<s.nb; j+=32)
			{
			lam[2*s.off_c+j] = w.v(CV_LAM_LO)[s.off_c+j]; lam[2*s.off_c+s.nb+j] = w.v(CV_LAM_UP)[s.off_c+j];
			tt[2*s.off_c+j] = w.v(CV_T_LO)[s.off_c+j]; tt[2*s.off_c+s.nb+j] = w.v(CV_T_UP)[s.off_c+j];
			}
		/* general constraints follow the bounds: [lb ub lg ug] per stage, the lib4 order (interfaces/c/c_order_interface.c:662-681) */
		for(int j=lane; j<s.ng; j+=32)
			{
			const int o = 2*s.off_c + 2*s.nb, cg = s.off_c + s.nb + j;
			lam[o+j] = w.v(CV_LAM_LO)[cg]; lam[o+s.ng+j] = w.v(CV_LAM_UP)[cg];
			tt[o+j] = w.v(CV_T_LO)[cg]; tt[o+s.ng+j] = w.v(CV_T_UP)[cg];
			}
		}
	}

/* the inverse of hb_ipm_emit_chain: lam, t of a caller-supplied iterate into the work vectors (single Newton step,
 * d_init_var_mpc_hard_tv_single_newton, mpc_solvers/c99/d_aux_ip_hard_lib4.c:153-213; bounds only, as in the reference) */
__device__ __forceinline__ void hb_ipm_load_chain(int lane, const hb_dims &d, const hb_ipm_ws &w, const double *lam, const double *tt)
	{
	for(int n=0; n<=d.N; n++)
		{
		const hb_stage s = d.st[n];
		for(int j=lane; j<s.nb; j+=32)
			{
			w.v(CV_LAM_LO)[s.off_c+j] = lam[2*s.off_c+j]; w.v(CV_LAM_UP)[s.off_c+j] = lam[2*s.off_c+s.nb+j];
			w.v(CV_T_LO)[s.off_c+j] = tt[2*s.off_c+j]; w.v(CV_T_UP)[s.off_c+j] = tt[2*s.off_c+s.nb+j];
			}
		for(int j=lane; j<s.ng; j+=32)
			{
			const int o = 2*s.off_c + 2*s.nb, cg = s.off_c + s.nb + j;
			w.v(CV_LAM_LO)[cg] = lam[o+j]; w.v(CV_LAM_UP)[cg] = lam[o+s.ng+j];
			w.v(CV_T_LO)[cg] = tt[o+j]; w.v(CV_T_UP)[cg] = tt[o+s.ng+j];
			}
		}
	}

/* The IPM kernel is written once; the sweeps over the horizon (or the tree) come from a policy: the run-time-size routines
 * of ric_generic.cuh, the size-specialised, bulk-copy-pipelined ones of ric_ipm_fast.cuh, or the tree ones of
 * ric_tree_ipm.cuh. */
struct hb_sweeps_generic
	{
	typedef hb_ctx ctx_t;
	static constexpr bool has_kkt = true;
	__device__ static __forceinline__ void extract(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w) { hb_ipm_extract_chain(c.lane, d, in_inst, w); }
	__device__ static __forceinline__ void emit(ctx_t &c, const hb_dims &d, const hb_ipm_ws &w, double *lam, double *tt) { hb_ipm_emit_chain(c.lane, d, w, lam, tt); }
	__device__ static __forceinline__ void load(ctx_t &c, const hb_dims &d, const hb_ipm_ws &w, const double *lam, const double *tt) { hb_ipm_load_chain(c.lane, d, w, lam, tt); }
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
	/* the same with Pb recomputed from bv (d_back_ric_rec_trs_tv_res with compute_Pb = 1): b is new, not the one of the last sv */
	__device__ static __forceinline__ void trs_newb(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *qx)
		{
		hb_trs_backward(c, d, in_inst, w.L, bv, rqv, qx, w.dux, w.Pb, true);
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
	static constexpr bool has_kkt = true;
	__device__ static __forceinline__ void extract(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w) { hb_ipm_extract_chain(c.lane, d, in_inst, w); }
	__device__ static __forceinline__ void emit(ctx_t &c, const hb_dims &d, const hb_ipm_ws &w, double *lam, double *tt) { hb_ipm_emit_chain(c.lane, d, w, lam, tt); }
	__device__ static __forceinline__ void load(ctx_t &c, const hb_dims &d, const hb_ipm_ws &w, const double *lam, const double *tt) { hb_ipm_load_chain(c.lane, d, w, lam, tt); }
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
	__device__ static __forceinline__ void trs_newb(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *qx)
		{
		hbi_trs_backward_newb<C>(c, d, in_inst, w.L, rqv, qx, bv, w.dux);
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


/* doubles of one work slot of the IPM kernel for the sweeps S: [factor | 3 ux-like | 4 pi-like | constraint vectors] */
template<class S> __device__ __forceinline__ long long hb_ipm_slot_doubles(const hb_dims &d)
	{ return S::L_doubles(d) + 3*d.ux_stride + 4*d.pi_stride + (long long)CV_COUNT*HB_EVEN(d.nbtot); }

template<class S> __device__ __forceinline__ hb_ipm_ws hb_ipm_make_ws(const hb_dims &d, double *p)
	{
	hb_ipm_ws w;
	w.L = p; p += S::L_doubles(d);
	w.dux = p; p += d.ux_stride; w.res_q = p; p += d.ux_stride; w.rq0 = p; p += d.ux_stride;
	w.dpi = p; p += d.pi_stride; w.Pb = p; p += d.pi_stride; w.res_b = p; p += d.pi_stride; w.b0 = p; p += d.pi_stride;
	w.cv = p; w.nbp = HB_EVEN(d.nbtot);
	return w;
	}

/* the same sweeps on the single-slot context (ric_ipm_fast.cuh: hbi_ctx1): solve-only and residual sweeps only, half the shared
 * memory per warp -- what the multi-kernel driver launches at 16 warps per SM (cipm_kernels.cu) */
template<class C>
struct hb_sweeps_fast1 : hb_sweeps_fast<C>
	{
	typedef hbi_ctx1<C> ctx_t;
	__device__ static __forceinline__ int smem_doubles(const hb_dims &) { return hbi_cfg1<C>::PER_WARP; }
	__device__ static __forceinline__ void init(ctx_t &c, const hb_dims &d, double *smem_warp, int lane) { c.init(smem_warp, lane, d); }
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

/* size-specialised IPM sweeps (ric_ipm_fast.cuh): one warp per instance, x0 eliminated, uniform (nx, nu) */
typedef hbf_cfg<24, 11, 32> hbi_v0;    /* BASELINE config 3 */
typedef hbf_cfg<12, 5, 32> hbi_v1;     /* config-2 sizes with bounds */
typedef hbf_cfg<8, 3, 32> hbi_v2;      /* the reference's own IPM test size (test_d_ip_hard.c) */
#define HBI_NVAR 3
