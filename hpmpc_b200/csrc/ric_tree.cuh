/*
 * ric_tree.cuh -- Riccati factor + solve over a scenario tree (one warp per (tree, segment) work item).
 *
 * Restates lqcp_solvers/d_tree_back_ric_rec_libstr.c (reference paths relative to /root/reference):
 *   backward node with kids   :79-156   W = [W_kid0 | W_kid1 | ...], W_k = [B A b]'_k Lxx_k, last row += l_x,k,
 *                                       L = chol_mn(RSQrq + sum_k W_k W_k')
 *   leaf                      :160-200  L = chol_mn(RSQrq)
 *   forward                   :204-260  root solves all of ux_0, other nodes only u; every kid gets
 *                                       x_k = b_k + [B A]_k ux_dad and pi_k = Lxx_k (Lxx_k' x_k + l_x,k)
 *   driver                    :524-583  nodes in BFS order, edges indexed by the kid
 *
 * The reference walks the nodes one by one.  Here a tree is cut into segments (host: tree.c): the "top" (all nodes above
 * the level from which every node has at most one kid, i.e. the robust horizon) and one "tail" chain per node of that
 * level.  Tails are independent of each other, so phase 0 (backward over the tails) and phase 2 (forward over the tails)
 * run one warp per (tree, tail) -- and can be split over GPUs, the only exchange being the factors of the tail roots
 * needed by phase 1 (backward + forward over the top, one warp per tree).
 *
 * Device layout of one tree instance: for every node n in BFS order  [ [B A b]'_in(n) : (nux_dad+1) x nx_n row-major, n > 0 ]
 * [ RSQrq_n : packed lower trapezoid, layout.h ].  Factors live in a node-indexed stash (global memory, per tree).
 */
#pragma once
#include "layout.h"
#include "ric_generic.cuh"

/* backward step of one node: cur (smem) <- chol_mn( RSQrq_n + sum over kids W_k W_k' ) ; Lt = this tree's factor stash */
__device__ __forceinline__ void hb_tree_node_factor(const hb_ctx &c, const hb_tnode *__restrict__ tn, int nn,
		const double *__restrict__ in_tree, double *Lt, double *cur, double *prev)
	{
	const int lane = c.lane;
	const hb_tnode s = tn[nn];
	const int nux = s.nu + s.nx, m = nux+1;
	const double *g = in_tree + s.off_RSQ;
	for(int e=lane; e<HB_TRI(nux)+nux; e+=32) cur[e] = g[e];
	__syncwarp();
	for(int kc=0; kc<s.nkids; kc++)
		{
		const hb_tnode k = tn[s.first_kid+kc];
		const int nx1 = k.nx, nu1 = k.nu, nux1 = nu1+nx1;
		hb_copy(c, prev, Lt + k.off_L, HB_TRI(nux1) + 2*nux1);
		double *sW = c.sW; const int ldW = c.ldW;
		const double *gb = in_tree + k.off_BAbt;
		for(int e=lane; e<m*nx1; e+=32)
			{
			int i = e/nx1, j = e - i*nx1;
			sW[i*ldW+j] = gb[e];
			}
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
		for(int j=lane; j<nx1; j+=32) sW[nux*ldW+j] += prev[HB_TRI(nux1)+nu1+j];
		__syncwarp();
		for(int i=lane; i<m; i+=32)
			{
			const double *wi = sW + i*ldW;
			double *hi = cur + HB_TRI(i);
			int kmax = i<nux ? i : nux-1;
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

/* forward step of one node: u_n (all of ux_n at the root), then x and pi of every kid */
__device__ __forceinline__ void hb_tree_node_forward(const hb_ctx &c, const hb_tnode *__restrict__ tn, int nn,
		const double *__restrict__ in_tree, const double *Lt, double *ux, double *pi, double *La, double *Lb)
	{
	const int lane = c.lane;
	const hb_tnode s = tn[nn];
	const int nu = s.nu, nux = s.nu + s.nx;
	const int ks = (s.dad<0) ? nux : nu;
	hb_copy(c, La, Lt + s.off_L, HB_TRI(nux) + 2*nux);
	__syncwarp();
	const double *dinv = La + HB_TRI(nux) + nux;
	double *v = c.sV, *xs = c.sV + 64, *tmp = c.sV + 128;
	for(int i=lane; i<nux; i+=32) v[i] = (i<ks) ? -La[HB_TRI(nux)+i] : ux[s.off_ux+i];
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
		double vj = v[j];
		for(int i=lane; i<j; i+=32) v[i] -= La[HB_TRI(j)+i]*vj;
		__syncwarp();
		}
	for(int i=lane; i<ks; i+=32) ux[s.off_ux+i] = v[i];
	for(int kc=0; kc<s.nkids; kc++)
		{
		const hb_tnode k = tn[s.first_kid+kc];
		const int nx1 = k.nx, nu1 = k.nu, nux1 = nu1+nx1;
		hb_copy(c, Lb, Lt + k.off_L, HB_TRI(nux1) + 2*nux1);
		const double *gb = in_tree + k.off_BAbt;
		for(int e=lane; e<(nux+1)*nx1; e+=32)
			{
			int i = e/nx1, j = e - i*nx1;
			c.sW[i*c.ldW+j] = gb[e];
			}
		__syncwarp();
		for(int j=lane; j<nx1; j+=32)
			{
			double acc = c.sW[nux*c.ldW+j];
			for(int i=0; i<nux; i++) acc += c.sW[i*c.ldW+j]*v[i];
			ux[k.off_ux+nu1+j] = acc;
			xs[j] = acc;
			}
		__syncwarp();
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = Lb[HB_TRI(nux1)+nu1+i];
			for(int kk=i; kk<nx1; kk++) acc += Lb[HB_TRI(nu1+kk)+nu1+i]*xs[kk];
			tmp[i] = acc;
			}
		__syncwarp();
		for(int i=lane; i<nx1; i+=32)
			{
			double acc = 0.0;
			for(int kk=0; kk<=i; kk++) acc += Lb[HB_TRI(nu1+i)+nu1+kk]*tmp[kk];
			pi[k.off_pi+i] = acc;
			}
		__syncwarp();
		}
	}
