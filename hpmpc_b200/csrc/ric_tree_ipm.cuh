/*
 * ric_tree_ipm.cuh -- the sweeps of the box-constrained IPM over a scenario tree (one warp per tree, any node sizes).
 *
 * Restates (reference paths relative to /root/reference):
 *   mpc_solvers/d_tree_ip2_res_hard_libstr.c:80       the tree IPM = the chain IPM (d_ip2_res_hard.c:116) with the tree Riccati
 *   lqcp_solvers/d_tree_back_ric_rec_libstr.c:79-200  backward node: diag/gradient updates at idxb, W_k = [B A b]'_k Lxx_k summed
 *                                                     over the kids, Pb_k = Lxx_k (Lxx_k' b_k) per edge
 *   lqcp_solvers/d_tree_back_ric_rec_libstr.c:625     solve with the stored factors for a new right-hand side (trs)
 *   mpc_solvers/d_tree_res_ip_res_hard_libstr.c:66    residuals: node n sees -pi of its own edge and + [B A]_k pi_k of every kid
 *
 * Same arithmetic per node as ric_generic.cuh (chain); the node table hb_tnode carries the topology and the bound offsets.
 * Vectors: "ux layout" = node-indexed off_ux, "pi layout" = edge-into-node off_pi; constraints flat at off_c.
 */
#pragma once
#include "layout.h"
#include "ric_generic.cuh"

/* [B A b]'_k (edge into kid k, rows of the dad: m of them) -> sW ; the b row replaced by bv when given */
__device__ __forceinline__ void hb_tipm_load_edge(const hb_ctx &c, const hb_tnode &k, const double *__restrict__ in_tree, int m, int brow,
		const double *bv)
	{
	const int nx1 = k.nx;
	const double *gb = in_tree + k.off_BAbt;
	for(int e=c.lane; e<m*nx1; e+=32)
		{
		const int i = e/nx1, j = e - i*nx1;
		double v = gb[e];
		if(bv!=nullptr && i==brow) v = bv[k.off_pi+j];
		c.sW[i*c.ldW+j] = v;
		}
	}

/* backward step of node nn: cur <- chol_mn( RSQrq_n (+Qx, qx, rq) + sum_k W_k W_k' ), factor -> Lt, Pb per kid edge */
__device__ __forceinline__ void hb_tipm_node_factor(const hb_ctx &c, const hb_tnode *__restrict__ tn, int nn,
		const double *__restrict__ in_tree, double *Lt, const double *bv, const double *rqv, const double *Qx, const double *qx,
		const int *__restrict__ idxb, double *Pb, double *cur, double *prev)
	{
	const int lane = c.lane;
	const hb_tnode s = tn[nn];
	const int nux = s.nu + s.nx, m = nux+1;
	const double *g = in_tree + s.off_RSQ;
	for(int e=lane; e<HB_TRI(nux)+nux; e+=32) cur[e] = g[e];
	__syncwarp();
	if(rqv!=nullptr)
		{
		for(int e=lane; e<nux; e+=32) cur[HB_TRI(nux)+e] = rqv[s.off_ux+e];
		__syncwarp();
		}
	if(Qx!=nullptr && s.nb>0)
		{
		for(int j=lane; j<s.nb; j+=32)
			{
			const int id = idxb[s.off_c+j];
			cur[HB_TRI(id)+id] += Qx[s.off_c+j];
			if(qx!=nullptr) cur[HB_TRI(nux)+id] += qx[s.off_c+j];
			}
		__syncwarp();
		}
	double *sW = c.sW; const int ldW = c.ldW;
	for(int kc=0; kc<s.nkids; kc++)
		{
		const hb_tnode k = tn[s.first_kid+kc];
		const int nx1 = k.nx, nu1 = k.nu, nux1 = nu1+nx1;
		hb_copy(c, prev, Lt + k.off_L, HB_TRI(nux1) + 2*nux1);
		hb_tipm_load_edge(c, k, in_tree, m, nux, bv);
		__syncwarp();
		for(int i=lane; i<m; i+=32)
			{
			double *w = sW + i*ldW;
			for(int j=0; j<nx1; j++)
				{
				double acc = 0.0;
				for(int kk=j; kk<nx1; kk++) acc += w[kk]*prev[HB_TRI(nu1+kk)+nu1+j];
				w[j] = acc;
				}
			}
		__syncwarp();
		if(Pb!=nullptr)
			{
			const double *wl = sW + nux*ldW;
			for(int i=lane; i<nx1; i+=32)
				{
				double acc = 0.0;
				for(int kk=0; kk<=i; kk++) acc += prev[HB_TRI(nu1+i)+nu1+kk]*wl[kk];
				Pb[k.off_pi+i] = acc;
				}
			__syncwarp();
			}
		for(int j=lane; j<nx1; j+=32) sW[nux*ldW+j] += prev[HB_TRI(nux1)+nu1+j];
		__syncwarp();
		for(int i=lane; i<m; i+=32)
			{
			const double *wi = sW + i*ldW;
			double *hi = cur + HB_TRI(i);
			const int kmax = i<nux ? i : nux-1;
			for(int kk=0; kk<=kmax; kk++)
				{
				const double *wk = sW + kk*ldW;
				double acc = 0.0;
				for(int mm=0; mm<nx1; mm++) acc += wi[mm]*wk[mm];
				hi[kk] += acc;
				}
			}
		__syncwarp();
		}
	hb_chol(c, cur, m, nux);
	hb_copy(c, Lt + s.off_L, cur, HB_TRI(nux) + 2*nux);
	__syncwarp();
	}

/* forward step of node nn: its inputs (all of ux at the root), then x and pi of every kid.
 *   lrow : eliminated right-hand side in ux layout (trs; may alias ux) or nullptr -> gradient row of L_n (sv)
 *   trs  : pi_k starts from the x-part of the eliminated rhs of the kid, still sitting in ux */
__device__ __forceinline__ void hb_tipm_node_forward(const hb_ctx &c, const hb_tnode *__restrict__ tn, int nn,
		const double *__restrict__ in_tree, const double *Lt, const double *lrow, const double *bv, bool trs,
		double *ux, double *pi, double *La, double *Lb)
	{
	const int lane = c.lane;
	const hb_tnode s = tn[nn];
	const int nu = s.nu, nux = s.nu + s.nx;
	const int ks = (s.dad<0) ? nux : nu;
	hb_copy(c, La, Lt + s.off_L, HB_TRI(nux) + 2*nux);
	__syncwarp();
	const double *dinv = La + HB_TRI(nux) + nux;
	double *v = c.sV, *xs = c.sV + 64, *tmp = c.sV + 128;
	for(int i=lane; i<nux; i+=32)
		v[i] = (i<ks) ? -(lrow!=nullptr ? lrow[s.off_ux+i] : La[HB_TRI(nux)+i]) : ux[s.off_ux+i];
	__syncwarp();
	for(int i=lane; i<ks; i+=32)
		{
		double acc = v[i];
		for(int j=ks; j<nux; j++) acc -= La[HB_TRI(j)+i]*v[j];
		v[i] = acc;
		}
	__syncwarp();
	for(int j=ks-1; j>=0; j--)
		{
		if(lane==(j&31)) v[j] *= dinv[j];
		__syncwarp();
		const double vj = v[j];
		for(int i=lane; i<j; i+=32) v[i] -= La[HB_TRI(j)+i]*vj;
		__syncwarp();
		}
	for(int i=lane; i<ks; i+=32) ux[s.off_ux+i] = v[i];
	for(int kc=0; kc<s.nkids; kc++)
		{
		const hb_tnode k = tn[s.first_kid+kc];
		const int nx1 = k.nx, nu1 = k.nu, nux1 = nu1+nx1;
		hb_copy(c, Lb, Lt + k.off_L, HB_TRI(nux1) + 2*nux1);
		hb_tipm_load_edge(c, k, in_tree, nux+1, nux, bv);
		__syncwarp();
		for(int j=lane; j<nx1; j+=32)
			{
			double acc = c.sW[nux*c.ldW+j];
			for(int i=0; i<nux; i++) acc += c.sW[i*c.ldW+j]*v[i];
			if(trs) pi[k.off_pi+j] = ux[k.off_ux+nu1+j];
			ux[k.off_ux+nu1+j] = acc;
			xs[j] = acc;
			}
		__syncwarp();
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = trs ? 0.0 : Lb[HB_TRI(nux1)+nu1+i];
			for(int kk=i; kk<nx1; kk++) acc += Lb[HB_TRI(nu1+kk)+nu1+i]*xs[kk];
			tmp[i] = acc;
			}
		__syncwarp();
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = trs ? pi[k.off_pi+i] : 0.0;
			for(int kk=0; kk<=i; kk++) acc += Lb[HB_TRI(nu1+i)+nu1+kk]*tmp[kk];
			pi[k.off_pi+i] = acc;
			}
		__syncwarp();
		}
	}

/* backward step of the solve-only sweep: w_n = L-eliminated( rq_n (+qx) + sum_k [B A]_k'(Pb_k + w_k,x) ) -> ux ; leaves keep the
 * raw right-hand side (the chain's last stage) */
__device__ __forceinline__ void hb_tipm_node_trs_back(const hb_ctx &c, const hb_tnode *__restrict__ tn, int nn,
		const double *__restrict__ in_tree, const double *Lt, const double *rqv, const double *qx, const int *__restrict__ idxb,
		double *ux, const double *Pb, double *La)
	{
	const int lane = c.lane;
	const hb_tnode s = tn[nn];
	const int nu = s.nu, nux = s.nu + s.nx;
	const int ks = (s.dad<0) ? nux : nu;
	double *v = c.sV, *tmp = c.sV + 64;
	for(int i=lane; i<nux; i+=32) v[i] = rqv[s.off_ux+i];
	__syncwarp();
	if(qx!=nullptr && s.nb>0)
		{
		for(int j=lane; j<s.nb; j+=32) v[idxb[s.off_c+j]] += qx[s.off_c+j];
		__syncwarp();
		}
	if(s.nkids>0)
		{
		for(int kc=0; kc<s.nkids; kc++)
			{
			const hb_tnode k = tn[s.first_kid+kc];
			const int nx1 = k.nx, nu1 = k.nu;
			hb_tipm_load_edge(c, k, in_tree, nux, -1, nullptr);
			for(int j=lane; j<nx1; j+=32) tmp[j] = Pb[k.off_pi+j] + ux[k.off_ux+nu1+j];
			__syncwarp();
			for(int i=lane; i<nux; i+=32)
				{
				double acc = v[i];
				const double *w = c.sW + i*c.ldW;
				for(int j=0; j<nx1; j++) acc += w[j]*tmp[j];
				v[i] = acc;
				}
			__syncwarp();
			}
		hb_copy(c, La, Lt + s.off_L, HB_TRI(nux) + 2*nux);
		__syncwarp();
		const double *dinv = La + HB_TRI(nux) + nux;
		for(int j=0; j<ks; j++)
			{
			if(lane==(j&31)) v[j] *= dinv[j];
			__syncwarp();
			const double vj = v[j];
			for(int i=j+1+lane; i<nux; i+=32) v[i] -= La[HB_TRI(i)+j]*vj;
			__syncwarp();
			}
		}
	for(int i=lane; i<nux; i+=32) ux[s.off_ux+i] = v[i];
	__syncwarp();
	}

/* res_q of node nn and res_b of the edges into its kids; lamd? no: lam_lo / lam_up flat at off_c.  Updates the running maxima. */
__device__ __forceinline__ void hb_tipm_node_residuals(const hb_ctx &c, const hb_tnode *__restrict__ tn, int nn,
		const double *__restrict__ in_tree, const double *rq0, const double *b0, const double *lam_lo, const double *lam_up,
		const int *__restrict__ idxb, const double *ux, const double *pi, double *res_q, double *res_b, double &nq, double &nb_)
	{
	const int lane = c.lane;
	const hb_tnode s = tn[nn];
	const int nu = s.nu, nx = s.nx, nux = nu+nx;
	double *H = c.bufA, *xs = c.sV, *ps = c.sV + 64;
	hb_copy(c, H, in_tree + s.off_RSQ, HB_TRI(nux));
	for(int i=lane; i<nux; i+=32)
		{
		xs[i] = ux[s.off_ux+i];
		double v = rq0[s.off_ux+i];
		if(s.dad>=0 && i>=nu) v -= pi[s.off_pi + (i-nu)];
		res_q[s.off_ux+i] = v;
		}
	__syncwarp();
	for(int j=lane; j<s.nb; j+=32) res_q[s.off_ux+idxb[s.off_c+j]] += -lam_lo[s.off_c+j] + lam_up[s.off_c+j];
	__syncwarp();
	for(int i=lane; i<nux; i+=32)
		{
		double acc = res_q[s.off_ux+i];
		const double *hi = H + HB_TRI(i);
		for(int j=0; j<=i; j++) acc += hi[j]*xs[j];
		for(int j=i+1; j<nux; j++) acc += H[HB_TRI(j)+i]*xs[j];
		res_q[s.off_ux+i] = acc;
		}
	__syncwarp();
	for(int kc=0; kc<s.nkids; kc++)
		{
		const hb_tnode k = tn[s.first_kid+kc];
		const int nx1 = k.nx, nu1 = k.nu;
		hb_tipm_load_edge(c, k, in_tree, nux, -1, nullptr);
		for(int j=lane; j<nx1; j+=32) ps[j] = pi[k.off_pi+j];
		__syncwarp();
		for(int i=lane; i<nux; i+=32)
			{
			double acc = res_q[s.off_ux+i];
			const double *wr = c.sW + i*c.ldW;
			for(int j=0; j<nx1; j++) acc += wr[j]*ps[j];
			res_q[s.off_ux+i] = acc;
			}
		for(int j=lane; j<nx1; j+=32)
			{
			double acc = b0[k.off_pi+j] - ux[k.off_ux+nu1+j];
			for(int i=0; i<nux; i++) acc += c.sW[i*c.ldW+j]*xs[i];
			res_b[k.off_pi+j] = acc;
			nb_ = fmax(nb_, fabs(acc));
			}
		__syncwarp();
		}
	for(int i=lane; i<nux; i+=32) nq = fmax(nq, fabs(res_q[s.off_ux+i]));
	__syncwarp();
	}
