/*
 * ric_generic.cuh -- warp-per-instance Riccati building blocks, run-time stage sizes (nu+nx+1 <= 64).
 *
 * One warp owns one OCP instance and walks its horizon.  Stage matrices live in shared memory:
 *   cur / prev : packed lower-trapezoid L_n / L_{n+1}   (layout.h)
 *   sW         : [B A b]' of the stage, row-major with odd leading dimension, overwritten in place by
 *                W = [B A b]' * Lxx_{n+1}
 * Lane i owns row i (and i+32) of W / H / L; the other operand of every product is read as a
 * shared-memory broadcast, so the inner loops are 1 DFMA per (broadcast LDS + own-row LDS).
 *
 * What each function restates (reference paths relative to /root/reference):
 *   hb_stage_factor   lqcp_solvers/d_back_ric_rec_libstr.c:125-181 (sv) / :229-305 (trf); lib4 twin
 *                     lqcp_solvers/d_back_ric_rec.c:236-333: trmm, Pb, gradient-row add, syrk+potrf
 *   hb_chol           kernel/c99/kernel_dpotrf_c99_lib4.c:553-640 pivot rule (>1e-15 else zero column,
 *                     inverse diagonal kept and multiplied)
 *   hb_stage_forward  lqcp_solvers/d_back_ric_rec.c:341-397 (sv) / :737-789 (trs)
 *   hb_trs_stage_back lqcp_solvers/d_back_ric_rec.c:628-732
 */
#pragma once
#include "layout.h"

#define HB_FULL 0xffffffffu

struct hb_ctx
	{
	int lane;
	int ldW;
	double *bufA, *bufB;     /* packed L buffers, tri(nzM)+2*nzM each */
	double *sW;              /* nzM x ldW */
	double *sV;              /* 4*nzM scratch vectors */
	};

__device__ __forceinline__ int hb_Lsize(int nux) { return HB_EVEN(HB_TRI(nux) + 2*nux); }

/* left-looking Cholesky of the packed m x nux trapezoid in smem; dinv written after the trapezoid */
__device__ __forceinline__ void hb_chol(const hb_ctx &c, double *L, int m, int nux)
	{
	double *dinv = L + HB_TRI(nux) + nux;
	const int lane = c.lane;
	for(int j=0; j<nux; j++)
		{
		const double *lj = L + HB_TRI(j);
		for(int i=j+lane; i<m; i+=32)
			{
			double *li = L + HB_TRI(i);
			double v = li[j];
			for(int k=0; k<j; k++) v -= li[k]*lj[k];
			li[j] = v;
			}
		__syncwarp();
		double piv = L[HB_TRI(j)+j];
		double d, inv;
		if(piv>1e-15) { d = sqrt(piv); inv = 1.0/d; }
		else { d = 0.0; inv = 0.0; }
		__syncwarp();
		for(int i=j+lane; i<m; i+=32)
			{
			double *li = L + HB_TRI(i);
			li[j] = (i==j) ? d : li[j]*inv;
			}
		if(lane==0) dinv[j] = inv;
		__syncwarp();
		}
	}

/* one backward stage: cur <- chol_mn( RSQrq_n (+Qx,qx) + W W' ),  W = [B A b]'_n Lxx_{n+1}  */
template<bool GRAD>
__device__ __forceinline__ void hb_stage_factor(const hb_ctx &c, const hb_stage &s, int nu1,
		const double *__restrict__ in_inst, const double *bvec, const double *rqvec,
		const double *Qx, const double *qx, const int *__restrict__ idxb, double *Pb,
		double *cur, const double *prev)
	{
	const int lane = c.lane;
	const int nu = s.nu, nux = s.nu + s.nx, nx1 = s.nx1;
	const int m = GRAD ? nux+1 : nux;
	const int ntri = HB_TRI(nux) + (GRAD ? nux : 0);
	const double *g = in_inst + s.off_RSQ;
	for(int e=lane; e<ntri; e+=32) cur[e] = g[e];
	__syncwarp();
	if(GRAD && rqvec!=nullptr)
		{
		for(int e=lane; e<nux; e+=32) cur[HB_TRI(nux)+e] = rqvec[s.off_ux+e];
		__syncwarp();
		}
	if(Qx!=nullptr && s.nb>0)
		{
		for(int j=lane; j<s.nb; j+=32)
			{
			int id = idxb[s.off_c+j];
			cur[HB_TRI(id)+id] += Qx[s.off_c+j];
			if(GRAD && qx!=nullptr) cur[HB_TRI(nux)+id] += qx[s.off_c+j];
			}
		__syncwarp();
		}
	if(Qx!=nullptr && s.ng>0)
		{
		/* general constraints: H += [D C]' diag(Qx_g) [D C], gradient row += [D C]' qx_g  (the reference appends
		 * DCt*diag(Qx_g) to W and DCt to its transpose before the syrk: lqcp_solvers/d_back_ric_rec.c:293-315,
		 * readable twin d_back_ric_rec_libstr.c:105-113,164-171) */
		const int ng = s.ng;
		const double *G = in_inst + s.off_DCt;
		const double *Qg = Qx + s.off_c + s.nb;
		for(int i=lane; i<m; i+=32)
			{
			double *hi = cur + HB_TRI(i);
			if(i<nux)
				{
				const double *gi = G + i*ng;
				for(int k=0; k<=i; k++)
					{
					const double *gk = G + k*ng;
					double acc = 0.0;
					for(int j=0; j<ng; j++) acc += gi[j]*Qg[j]*gk[j];
					hi[k] += acc;
					}
				}
			else if(qx!=nullptr)
				{
				const double *qg = qx + s.off_c + s.nb;
				for(int k=0; k<nux; k++)
					{
					const double *gk = G + k*ng;
					double acc = 0.0;
					for(int j=0; j<ng; j++) acc += qg[j]*gk[j];
					hi[k] += acc;
					}
				}
			}
		__syncwarp();
		}
	(void)nu;
	if(nx1>0)
		{
		double *sW = c.sW; const int ldW = c.ldW;
		const double *gb = in_inst + s.off_BAbt;
		for(int e=lane; e<m*nx1; e+=32)
			{
			int i = e/nx1, j = e - i*nx1;
			double v = gb[e];
			if(GRAD && bvec!=nullptr && i==nux) v = bvec[s.off_pi+j];
			sW[i*ldW+j] = v;
			}
		__syncwarp();
		/* W = [B A b]' * Lxx_{n+1}, in place, row per lane */
		for(int i=lane; i<m; i+=32)
			{
			double *w = sW + i*ldW;
			for(int j=0; j<nx1; j++)
				{
				double acc = 0.0;
				for(int k=j; k<nx1; k++) acc += w[k]*prev[HB_TRI(nu1+k)+nu1+j];
				w[j] = acc;
				}
			}
		__syncwarp();
		if(GRAD)
			{
			const double *wl = sW + nux*ldW;
			if(Pb!=nullptr)
				for(int i=lane; i<nx1; i+=32)
					{
					double acc = 0.0;
					for(int k=0; k<=i; k++) acc += prev[HB_TRI(nu1+i)+nu1+k]*wl[k];
					Pb[s.off_pi+i] = acc;
					}
			__syncwarp();
			for(int j=lane; j<nx1; j+=32) sW[nux*ldW+j] += prev[HB_TRI(nu1+nx1)+nu1+j];
			__syncwarp();
			}
		/* H += W W' (lower), row per lane */
		for(int i=lane; i<m; i+=32)
			{
			const double *wi = sW + i*ldW;
			double *hi = cur + HB_TRI(i);
			int kmax = i<nux ? i : nux-1;
			for(int k=0; k<=kmax; k++)
				{
				const double *wk = sW + k*ldW;
				double acc = 0.0;
				for(int mm=0; mm<nx1; mm++) acc += wi[mm]*wk[mm];
				hi[k] += acc;
				}
			}
		__syncwarp();
		}
	hb_chol(c, cur, m, nux);
	}

/* copy a packed factor between smem and the stash (global) */
__device__ __forceinline__ void hb_copy(const hb_ctx &c, double *dst, const double *src, int n)
	{
	for(int e=c.lane; e<n; e+=32) dst[e] = src[e];
	}

/* load [B A b]'_n into sW (all nux+1 rows) */
__device__ __forceinline__ void hb_load_BAbt(const hb_ctx &c, const hb_stage &s, const double *__restrict__ in_inst)
	{
	const int nux = s.nu+s.nx, nx1 = s.nx1;
	const double *gb = in_inst + s.off_BAbt;
	for(int e=c.lane; e<(nux+1)*nx1; e+=32)
		{
		int i = e/nx1, j = e - i*nx1;
		c.sW[i*c.ldW+j] = gb[e];
		}
	}

/* one forward stage n: solves for u_n (all of ux_0 when n==0), then x_{n+1} and pi_n.
 *   Ln (smem) = L_n, Ln1 (smem) = L_{n+1}, sW holds [B A b]'_n
 *   lrow : rhs vector in ux layout (trs) or nullptr -> gradient row of L_n (sv)
 *   bvec : b in pi layout or nullptr -> row nux of sW
 *   trs  : pi_n starts from the x-part of the eliminated rhs of stage n+1 (already sitting in ux)  */
__device__ __forceinline__ void hb_stage_forward(const hb_ctx &c, const hb_stage &s, const hb_stage &s1, int n,
		const double *Ln, const double *Ln1, const double *lrow, const double *bvec, bool trs,
		double *ux, double *pi, bool compute_pi)
	{
	const int lane = c.lane;
	const int nu = s.nu, nux = s.nu+s.nx, nx1 = s.nx1, nu1 = s1.nu, nux1 = s1.nu+s1.nx;
	const int ks = (n==0) ? nux : nu;
	const double *dinv = Ln + HB_TRI(nux) + nux;
	double *v = c.sV, *xs = c.sV + 64, *tmp = c.sV + 128;
	for(int i=lane; i<nux; i+=32)
		v[i] = (i<ks) ? -(lrow!=nullptr ? lrow[s.off_ux+i] : Ln[HB_TRI(nux)+i]) : ux[s.off_ux+i];
	__syncwarp();
	/* v[:ks] -= L[ks:nux,:ks]' v[ks:nux] */
	for(int i=lane; i<ks; i+=32)
		{
		double acc = v[i];
		for(int j=ks; j<nux; j++) acc -= Ln[HB_TRI(j)+i]*v[j];
		v[i] = acc;
		}
	__syncwarp();
	/* v[:ks] = L[:ks,:ks]^{-T} v[:ks]  (column oriented) */
	for(int j=ks-1; j>=0; j--)
		{
		if(lane==(j&31)) v[j] *= dinv[j];
		__syncwarp();
		double vj = v[j];
		for(int i=lane; i<j; i+=32) v[i] -= Ln[HB_TRI(j)+i]*vj;
		__syncwarp();
		}
	for(int i=lane; i<ks; i+=32) ux[s.off_ux+i] = v[i];
	/* x_{n+1} = b + [B A] ux */
	for(int j=lane; j<nx1; j+=32)
		{
		double acc = bvec!=nullptr ? bvec[s.off_pi+j] : c.sW[nux*c.ldW+j];
		for(int i=0; i<nux; i++) acc += c.sW[i*c.ldW+j]*v[i];
		if(trs && compute_pi) pi[s.off_pi+j] = ux[s1.off_ux+nu1+j];
		ux[s1.off_ux+nu1+j] = acc;
		xs[j] = acc;
		}
	__syncwarp();
	if(compute_pi)
		{
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = trs ? 0.0 : Ln1[HB_TRI(nux1)+nu1+i];
			for(int k=i; k<nx1; k++) acc += Ln1[HB_TRI(nu1+k)+nu1+i]*xs[k];
			tmp[i] = acc;
			}
		__syncwarp();
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = trs ? pi[s.off_pi+i] : 0.0;
			for(int k=0; k<=i; k++) acc += Ln1[HB_TRI(nu1+i)+nu1+k]*tmp[k];
			pi[s.off_pi+i] = acc;
			}
		}
	__syncwarp();
	}

/* one backward stage of the solve-only sweep (n < N):  w_n = L-eliminated( rq_n (+qx) + [B A]'(Pb_n + w_{n+1,x}) )
 *   Ln (smem) = L_n ; Ln1 (smem) = L_{n+1} (only read when compute_Pb) ; sW holds [B A b]'_n ; w lives in ux */
__device__ __forceinline__ void hb_trs_stage_back(const hb_ctx &c, const hb_stage &s, const hb_stage &s1, int n,
		const double *Ln, const double *Ln1, const double *bvec, const double *rqvec, const double *qx,
		const int *__restrict__ idxb, double *ux, double *Pb, bool compute_Pb, const double *__restrict__ in_inst = nullptr)
	{
	const int lane = c.lane;
	const int nu = s.nu, nux = s.nu+s.nx, nx1 = s.nx1, nu1 = s1.nu;
	const int ks = (n==0) ? nux : nu;
	const double *dinv = Ln + HB_TRI(nux) + nux;
	double *v = c.sV, *tmp = c.sV + 64, *t2 = c.sV + 128;
	if(compute_Pb)
		{
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = 0.0;
			for(int k=i; k<nx1; k++) acc += Ln1[HB_TRI(nu1+k)+nu1+i]*bvec[s.off_pi+k];
			t2[i] = acc;
			}
		__syncwarp();
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = 0.0;
			for(int k=0; k<=i; k++) acc += Ln1[HB_TRI(nu1+i)+nu1+k]*t2[k];
			Pb[s.off_pi+i] = acc;
			}
		__syncwarp();
		}
	for(int i=lane; i<nux; i+=32) v[i] = rqvec[s.off_ux+i];
	for(int j=lane; j<nx1; j+=32) tmp[j] = Pb[s.off_pi+j] + ux[s1.off_ux+nu1+j];
	__syncwarp();
	if(qx!=nullptr && s.nb>0)
		{
		for(int j=lane; j<s.nb; j+=32) v[idxb[s.off_c+j]] += qx[s.off_c+j];
		__syncwarp();
		}
	if(qx!=nullptr && s.ng>0 && in_inst!=nullptr)
		{
		/* general constraints: + [D C]' qx_g  (lqcp_solvers/d_back_ric_rec.c:668-672) */
		const double *G = in_inst + s.off_DCt, *qg = qx + s.off_c + s.nb;
		for(int i=lane; i<nux; i+=32)
			{
			double acc = v[i];
			for(int j=0; j<s.ng; j++) acc += G[i*s.ng+j]*qg[j];
			v[i] = acc;
			}
		__syncwarp();
		}
	for(int i=lane; i<nux; i+=32)
		{
		double acc = v[i];
		const double *w = c.sW + i*c.ldW;
		for(int j=0; j<nx1; j++) acc += w[j]*tmp[j];
		v[i] = acc;
		}
	__syncwarp();
	/* forward substitution with the first ks columns of L_n */
	for(int j=0; j<ks; j++)
		{
		if(lane==(j&31)) v[j] *= dinv[j];
		__syncwarp();
		double vj = v[j];
		for(int i=j+1+lane; i<nux; i+=32) v[i] -= Ln[HB_TRI(i)+j]*vj;
		__syncwarp();
		}
	for(int i=lane; i<nux; i+=32) ux[s.off_ux+i] = v[i];
	__syncwarp();
	(void)nu;
	}
