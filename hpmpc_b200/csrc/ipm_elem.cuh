/*
 * ipm_elem.cuh -- per-constraint (element-wise) pieces of the box IPM shared by the fused kernel (ipm_kernels.cu) and the
 * multi-kernel scenario-tree driver (tree_ipm_kernels.cu): work-area layout, warp reductions, step length, mu_aff,
 * the bound part of the residuals.  Reference lines cited at each function (mpc_solvers/c99/d_aux_ip_hard_lib4.c).
 */
#pragma once
#include "layout.h"
#include "ric_generic.cuh"

enum { CV_LB=0, CV_UB, CV_LAM_LO, CV_LAM_UP, CV_T_LO, CV_T_UP, CV_DLAM_LO, CV_DLAM_UP, CV_DT_LO, CV_DT_UP,
       CV_TINV_LO, CV_TINV_UP, CV_LAMT_LO, CV_LAMT_UP, CV_QXD /* "Qx": Hessian diagonal term */,
       CV_QXG /* "qx": gradient term */, CV_RD_LO, CV_RD_UP, CV_RM_LO, CV_RM_UP,
       CV_VAL /* values [D C] ux of the general constraints (box constraints read ux[c_ux] directly) */, CV_COUNT };

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

/* Value of constraint cc for the vector v (ux layout): the bounded variable itself for a box constraint, ([D C] v)_j for a
 * general one -- the latter precomputed into CV_VAL by hb_gen_values (the reference's dgemv_t on pDCt,
 * mpc_solvers/c99/d_aux_ip_hard_lib4.c:135,570,...; libstr twin c99/d_aux_ip_hard_libstr.c:125,329) */
__device__ __forceinline__ double hb_cval(const hb_dims &d, const hb_ipm_ws &w, const double *v, int cc)
	{
	const int iu = d.c_ux[cc];
	return iu>=0 ? v[iu] : w.v(CV_VAL)[cc];
	}
/* CV_VAL <- [D C]_n v_n for every general constraint; thread tid of `stride` cooperating threads; nothing to do when ngtot = 0 */
__device__ __forceinline__ void hb_gen_values_part(int tid, int stride, const hb_dims &d, const double *__restrict__ in_inst, const hb_ipm_ws &w, const double *v)
	{
	if(d.ngtot==0 || d.st==nullptr) return;
	for(int n=0; n<=d.N; n++)
		{
		const hb_stage s = d.st[n];
		if(s.ng==0) continue;
		const int nux = s.nu+s.nx;
		const double *G = in_inst + s.off_DCt;
		for(int j=tid; j<s.ng; j+=stride)
			{
			double acc = 0.0;
			for(int i=0; i<nux; i++) acc += G[i*s.ng+j]*v[s.off_ux+i];
			w.v(CV_VAL)[s.off_c+s.nb+j] = acc;
			}
		}
	}
__device__ __forceinline__ void hb_gen_values(int lane, const hb_dims &d, const double *__restrict__ in_inst, const hb_ipm_ws &w, const double *v)
	{
	if(d.ngtot==0) return;
	__syncwarp();
	hb_gen_values_part(lane, 32, d, in_inst, w, v);
	__syncwarp();
	}

/* bound part of the residuals: res_d, res_m and their sum (mpc_solvers/c99/d_res_ip_res_hard.c:39-319) */
/* _part: thread tid of `stride` cooperating threads, partial results (the caller reduces) */
__device__ __forceinline__ void hb_ipm_residuals_bounds_part(int tid, int stride, const hb_dims &d, const hb_ipm_ws &w, const double *ux, double &mu2, double &nd)
	{
	const double *lam_lo = w.v(CV_LAM_LO), *lam_up = w.v(CV_LAM_UP), *t_lo = w.v(CV_T_LO), *t_up = w.v(CV_T_UP);
	mu2 = 0.0; nd = 0.0;
	for(int cc=tid; cc<d.nbtot; cc+=stride)
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
	}
__device__ __forceinline__ void hb_ipm_residuals_bounds(int lane, const hb_dims &d, const hb_ipm_ws &w, const double *ux, double &mu2, double &nd)
	{
	hb_ipm_residuals_bounds_part(lane, 32, d, w, ux, mu2, nd);
	mu2 = hb_warp_sum(mu2);
	}

/* step length + dt, dlam.  RES = false: phase 1 (c99/d_aux_ip_hard_lib4.c:489-614) ; true: phase 2 (:1180-1313) */
template<bool RES>
__device__ __forceinline__ double hb_ipm_alpha_part(int tid, int stride, const hb_dims &d, const hb_ipm_ws &w, const double *dux)
	{
	double alpha = 1.0;
	for(int cc=tid; cc<d.nbtot; cc+=stride)
		{
		double du = hb_cval(d, w, dux, cc);
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
	return alpha;
	}
template<bool RES>
__device__ __forceinline__ double hb_ipm_alpha(int lane_, const hb_dims &d, const hb_ipm_ws &w, const double *dux)
	{
	return hb_warp_min(hb_ipm_alpha_part<RES>(lane_, 32, d, w, dux));
	}

/* mu_aff = mu_scal * sum (lam + a dlam)(t + a dt)   (c99/d_aux_ip_hard_lib4.c:715-770, :1453-1508) */
__device__ __forceinline__ double hb_ipm_mu_aff_part(int tid, int stride, const hb_dims &d, const hb_ipm_ws &w, double alpha)
	{
	double mu = 0.0;
	for(int cc=tid; cc<d.nbtot; cc+=stride)
		mu += (w.v(CV_LAM_LO)[cc] + alpha*w.v(CV_DLAM_LO)[cc])*(w.v(CV_T_LO)[cc] + alpha*w.v(CV_DT_LO)[cc])
		    + (w.v(CV_LAM_UP)[cc] + alpha*w.v(CV_DLAM_UP)[cc])*(w.v(CV_T_UP)[cc] + alpha*w.v(CV_DT_UP)[cc]);
	return mu;
	}
__device__ __forceinline__ double hb_ipm_mu_aff(int lane_, const hb_dims &d, const hb_ipm_ws &w, double alpha, double mu_scal)
	{
	return hb_warp_sum(hb_ipm_mu_aff_part(lane_, 32, d, w, alpha))*mu_scal;
	}

