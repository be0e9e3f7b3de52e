/*
 * ric_sweeps.cuh -- the generic (any size pattern) Riccati sweeps on one warp per instance, shared by the Riccati kernels
 * (ric_kernels.cu) and the generic policy of the IPM kernel (ipm_kernels.cu).  Reference lines are cited at each sweep.
 */
#pragma once
#include "layout.h"
#include "ric_generic.cuh"

extern __shared__ double hb_smem[];

/* ------------------------------------------------------------------------------------------------ */
/* sweeps                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ hb_ctx hb_make_ctx(const hb_dims &d, double *smem_warp, int lane)
	{
	hb_ctx c;
	c.lane = lane;
	c.ldW = HB_LDW(d.nxM);
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
	return 2*lsz + HB_EVEN(nzM*HB_LDW(nxM)) + 192;
	}

/* the any-size sweeps fetch a stage with ordinary loads, a few in flight per lane: pull the NEXT stage's data towards the SM (L2)
 * while the current stage is being worked on, one 128-byte line per lane and step */
__device__ __forceinline__ void hb_prefetch_l2(const double *p, int n, int lane)
	{
	for(int o=16*lane; o<n; o+=512) asm volatile("prefetch.global.L2 [%0];" :: "l"(p + o));
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
		if(n>0)
			{
			const hb_stage sp = d.st[n-1];                                   /* [B A b]' and RSQrq of a stage are contiguous */
			hb_prefetch_l2(in_inst + sp.off_BAbt, HB_EVEN((sp.nu+sp.nx+1)*sp.nx1) + HB_TRI(sp.nu+sp.nx) + sp.nu+sp.nx, c.lane);
			}
		hb_stage_factor<GRAD>(c, s, nu1, in_inst, bvec, rqvec, Qx, qx, d.idxb, Pb, cur, prev);
		hb_copy(c, Lst + s.off_L, cur, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		double *t = cur; cur = prev; prev = t;
		__syncwarp();
		}
	}

/* forward sweep n = 0..N-1 */
static __device__ void hb_forward(const hb_ctx &c, const hb_dims &d, const double *in_inst, const double *Lst,
		const double *lrow, const double *bvec, bool trs, double *ux, double *pi, bool compute_pi)
	{
	double *a = c.bufA, *b = c.bufB;
	{
	const hb_stage s0 = d.st[0];
	hb_g2s(c.lane, a, Lst + s0.off_L, HB_TRI(s0.nu+s0.nx) + 2*(s0.nu+s0.nx));
	}
	for(int n=0; n<d.N; n++)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		if(n+2<=d.N)
			{
			const hb_stage s2 = d.st[n+2];
			hb_prefetch_l2(Lst + s2.off_L, HB_TRI(s2.nu+s2.nx) + 2*(s2.nu+s2.nx), c.lane);
			if(n+1<d.N) hb_prefetch_l2(in_inst + s1.off_BAbt, (s1.nu+s1.nx+1)*s1.nx1, c.lane);
			}
		hb_g2s(c.lane, b, Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx));
		hb_load_BAbt_async(c, s, in_inst);
		hb_g2s_wait();
		__syncwarp();
		hb_stage_forward(c, s, s1, n, a, b, lrow, bvec, trs, ux, pi, compute_pi);
		double *t = a; a = b; b = t;
		}
	}

/* solve-only backward vector sweep; w is kept in ux */
static __device__ void hb_trs_backward(const hb_ctx &c, const hb_dims &d, const double *in_inst, const double *Lst,
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
	if(qx!=nullptr && s.ng>0)
		{
		const double *G = in_inst + s.off_DCt, *qg = qx + s.off_c + s.nb;
		for(int i=lane; i<nux; i+=32)
			{
			double acc = ux[s.off_ux+i];
			for(int j=0; j<s.ng; j++) acc += G[i*s.ng+j]*qg[j];
			ux[s.off_ux+i] = acc;
			}
		__syncwarp();
		}
	}
	for(int n=d.N-1; n>=0; n--)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		if(n>0)
			{
			const hb_stage sp = d.st[n-1];
			hb_prefetch_l2(Lst + sp.off_L, HB_TRI(sp.nu+sp.nx) + 2*(sp.nu+sp.nx), lane);
			hb_prefetch_l2(in_inst + sp.off_BAbt, (sp.nu+sp.nx+1)*sp.nx1, lane);
			}
		hb_g2s(lane, c.bufA, Lst + s.off_L, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		if(compute_Pb) hb_g2s(lane, c.bufB, Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx));
		hb_load_BAbt_async(c, s, in_inst);
		hb_g2s_wait();
		__syncwarp();
		hb_trs_stage_back(c, s, s1, n, c.bufA, c.bufB, bvec, rqvec, qx, d.idxb, ux, Pb, compute_Pb, in_inst);
		}
	}

