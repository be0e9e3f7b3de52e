/*
 * ric_team.cuh -- any-size backward (factorisation) sweep with FOUR warps per instance.
 *
 * The one-warp-per-instance sweeps of ric_sweeps.cuh keep [B A b]' and two packed factors of an instance in shared memory
 * (38 KB at nu+nx+1 = 49), so an SM holds five or six instances -- and with one warp each that is one warp per scheduler: every
 * dependent instruction pays its full latency (ncu, config 4: 5.7 cycles per warp instruction, FP64 pipe 10 % busy).  Here the
 * same shared-memory footprint is worked on by a CTA of 128 threads: warp w owns rows 16w .. 16w+15 of W / H / L (lane = (row
 * slot, half) as in ric_generic.cuh, one row per lane), so an SM runs 20-24 warps and an instance's stage takes a quarter of
 * the dependent steps.  The forward sweep of the same instance follows on warp 0 (hb_forward) or in a kernel of its own
 * (multi-kernel IPM driver).
 *
 * Restates the same reference lines as hb_stage_factor (lqcp_solvers/d_back_ric_rec.c:236-333; readable twin
 * d_back_ric_rec_libstr.c:125-181 / :229-305); the stage routine is kept operation-for-operation equal to the one-warp
 * version's register-tiled path (same sums in the same order), so both give the same bits.
 */
#pragma once
#include "ric_sweeps.cuh"

#define HBT_WARPS 4
#define HBT_THREADS (32*HBT_WARPS)

__device__ __forceinline__ void hbt_sync() { __syncthreads(); }

/* -DHBT_TIMING: block 0 accumulates clock64() per phase of the backward sweep, lane 0 of every warp (tools/phase_timing_team.py) */
#ifdef HBT_TIMING
__device__ long long hbt_tm_g[4*12];
#define HBT_T0 long long hbt_t0_ = clock64();
#define HBT_T(idx) do { if(blockIdx.x==0 && (tid&31)==0) { long long t_ = clock64(); atomicAdd((unsigned long long*)&hbt_tm_g[12*(tid>>5)+(idx)], (unsigned long long)(t_-hbt_t0_)); hbt_t0_ = t_; } } while(0)
#else
#define HBT_T0
#define HBT_T(idx)
#endif

/* one 4-column panel (W part, then L part)  of cur <- chol_mn(cur) on the 16 rows of this warp (row0 .. row0+15); see hbg_syrk_chol.
 * Dblk: 16 doubles of shared scratch for the diagonal block */
__device__ __forceinline__ void hbt_panel_acc(double *cur, const double *sW, int ld, int lane, int row0, int m, int nux, int nx1, int jb,
		double (&acc)[4], int &irow_out, bool &valid_out)
	{
	const int rt = lane>>1, h = lane&1;
	int i = row0 + rt;
	const bool valid = i<m;
	i = valid ? i : m-1;
	const int wrow = i*ld, lrow = HB_TRI(i);
	int wj[4], lj[4];
#pragma unroll
	for(int cc=0; cc<4; cc++) { int j = jb+cc; j = j<nux ? j : nux-1; wj[cc] = j*ld; lj[cc] = HB_TRI(j); }
	acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; acc[3] = 0.0;
	double hw[4];
#pragma unroll 4
	for(int k=h; k<nx1; k+=2)
		{
		const double b0 = sW[wj[0]+k], b1 = sW[wj[1]+k], b2 = sW[wj[2]+k], b3 = sW[wj[3]+k];
		const double a = sW[wrow+k];
		acc[0] = fma(a, b0, acc[0]); acc[1] = fma(a, b1, acc[1]);
		acc[2] = fma(a, b2, acc[2]); acc[3] = fma(a, b3, acc[3]);
		}
#pragma unroll
	for(int cc=0; cc<4; cc++)
		{
		double v = acc[cc] + __shfl_xor_sync(HB_FULL, acc[cc], 1);
		if((jb+cc<nux) && (jb+cc<=i)) v += cur[lrow+jb+cc];
		hw[cc] = v;
		acc[cc] = 0.0;
		}
#pragma unroll 4
	for(int k=h; k<jb; k+=2)
		{
		const double b0 = cur[lj[0]+k], b1 = cur[lj[1]+k], b2 = cur[lj[2]+k], b3 = cur[lj[3]+k];
		const double a = -cur[lrow+k];
		acc[0] = fma(a, b0, acc[0]); acc[1] = fma(a, b1, acc[1]);
		acc[2] = fma(a, b2, acc[2]); acc[3] = fma(a, b3, acc[3]);
		}
#pragma unroll
	for(int cc=0; cc<4; cc++) acc[cc] = (acc[cc] + __shfl_xor_sync(HB_FULL, acc[cc], 1)) + hw[cc];
	irow_out = i; valid_out = valid;
	}

/* the whole factorisation of one stage by the team: cur <- chol_mn(cur + W W'), W in sW (m x nx1), Dblk 16 doubles */
__device__ __forceinline__ void hbt_syrk_chol(double *cur, const double *sW, int ld, int tid, int m, int nux, int nx1, double *Dblk
#ifdef HBT_TIMING
		, long long &hbt_t0_
#endif
		)
	{
	const int warp = tid>>5, lane = tid&31, h = lane&1;
	const int row0 = 16*warp;
	double *dinv = cur + HB_TRI(nux) + nux;
	for(int jb=0; jb<nux; jb+=4)
		{
		const bool active = (row0+15>=jb) && (row0<m);                 /* warp-uniform */
		double acc[4]; int i = 0; bool valid = false;
		HBT_T(8);
		if(active)
			{
			hbt_panel_acc(cur, sW, ld, lane, row0, m, nux, nx1, jb, acc, i, valid);
			if(valid && h==0 && i>=jb && i<jb+4)
				{
#pragma unroll
				for(int cc=0; cc<4; cc++) if(jb+cc<=i) Dblk[4*(i-jb)+cc] = acc[cc];
				}
			}
		HBT_T(4);
		hbt_sync();
		HBT_T(5);
		if(active)
			{
			double D[4][4], dd[4], iv[4];
#pragma unroll
			for(int cc=0; cc<4; cc++)
#pragma unroll
				for(int c2=0; c2<=cc; c2++)
					D[cc][c2] = (jb+cc<nux) ? Dblk[4*cc+c2] : (c2==cc ? 1.0 : 0.0);
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
				const double piv = D[cc][cc];
				const double inv = (piv>1e-15) ? hbg_rsqrt(piv) : 0.0;
				dd[cc] = piv*inv; iv[cc] = inv;
#pragma unroll
				for(int c2=cc+1; c2<4; c2++) D[c2][cc] *= inv;
#pragma unroll
				for(int c2=cc+1; c2<4; c2++)
#pragma unroll
					for(int c3=cc+1; c3<=c2; c3++) D[c2][c3] = fma(-D[c2][cc], D[c3][cc], D[c2][c3]);
				}
			double x[4];
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
				double v = acc[cc];
#pragma unroll
				for(int c2=0; c2<cc; c2++) v = fma(-x[c2], D[cc][c2], v);
				x[cc] = v*iv[cc];
				}
			if(valid)
				{
				const int lrow = HB_TRI(i);
#pragma unroll
				for(int cc=0; cc<4; cc++)
					if((cc>>1)==h && jb+cc<nux && jb+cc<=i)
						cur[lrow+jb+cc] = (i==jb+cc) ? dd[cc] : x[cc];
				}
			if(lane==0 && (jb>>4)==warp)
				{
				dinv[jb] = iv[0];
				if(jb+1<nux) dinv[jb+1] = iv[1];
				if(jb+2<nux) dinv[jb+2] = iv[2];
				if(jb+3<nux) dinv[jb+3] = iv[3];
				}
			}
		HBT_T(6);
		hbt_sync();
		HBT_T(7);
		}
	}

/* ---------------------------------------------------------------------------------------------------------------- */
/* FP64 tensor-core path (mma.sync.m8n8k4.f64).  ncu on the FMA version of this sweep (profiles/r02_ncu_cipm_team_sv.txt): issue  */
/* slots 47 % and the shared-memory pipe 48 % busy with the FP64 pipe at 16 % -- a one-row-per-lane panel spends 10 LDS and 8 DFMA  */
/* (plus their address arithmetic) on 256 multiply-adds; one DMMA does them from two fragment loads.  DMMA runs on the same pipe at */
/* the same rate (tools/microbench_dmma.cu: 4.0 SM-cycles per warp-DMMA = 64 FMA/clk/SM, 26 cycles dependent latency), so the gain  */
/* is in issue slots and shared-memory wavefronts, which is what bounds this kernel.                                              */
/* Fragments (PTX ISA, m8n8k4): A row-major 8x4, lane l holds A[l>>2][l&3]; B col-major 4x8, lane l holds B[l&3][l>>2]; C 8x8, lane l   */
/* holds C[l>>2][2(l&3)], C[l>>2][2(l&3)+1].                                                                                       */
/* ---------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void hbt_dmma(double &c0, double &c1, double a, double b)
	{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
	}

/* W <- W * Lxx_{n+1} in place by the team on 8 x 8 tiles.  A warp owns the row tiles rt = warp, warp+4 and computes ALL their column
 * tiles (tile (rt, ct) sums over k >= 8ct) before it stores any: it reads and writes its own rows only, so the warps do not meet
 * until W is complete.  One pass over k serves every column tile at or before k (a B fragment per column tile, an A fragment per row
 * tile; up to 16 independent DMMA chains).  The k range is padded to a multiple of 4: columns nx1 .. nx1p-1 of W hold zeros
 * (hbt_stage_factor), the matching rows of Lxx are taken as zero */
template<int NCT>
__device__ __noinline__ void hbt_trmm_mma_t(double *sW, int ld, int tid, int m, int nx1, const double *prev, int nu1)
	{
	const int warp = tid>>5, lane = tid&31, g = lane>>2, q = lane&3;
	const int n_rt = (m+7)>>3, n_ct = (nx1+7)>>3, nx1p = (nx1+3)&~3;
	if(warp<n_rt)
		{
		const bool two = warp+4<n_rt;
		int r0 = 8*warp + g, r1 = 8*(warp+4) + g;
		const int ra0 = r0<m ? r0 : m-1, ra1 = r1<m ? r1 : m-1;
		const double *pa0 = sW + ra0*ld + q, *pa1 = sW + ra1*ld + q;
		double acc0[NCT][2], acc1[NCT][2];
#pragma unroll
		for(int ct=0; ct<NCT; ct++) { acc0[ct][0] = 0.0; acc0[ct][1] = 0.0; acc1[ct][0] = 0.0; acc1[ct][1] = 0.0; }
		for(int k0=0; k0<nx1p; k0+=4)
			{
			const int kr = k0 + q;
			const double a0 = pa0[k0], a1 = pa1[k0];
			const double *pb = prev + HB_TRI(nu1+kr) + nu1 + g;
			const bool kin = kr<nx1;
#pragma unroll
			for(int ct=0; ct<NCT; ct++)
				if(8*ct<=k0 && ct<n_ct)                                      /* warp-uniform */
					{
					const int col = 8*ct + g;
					const double b = (kin && kr>=col) ? pb[8*ct] : 0.0;          /* kr >= col and kr < nx1 imply col < nx1 */
					hbt_dmma(acc0[ct][0], acc0[ct][1], a0, b);
					if(two) hbt_dmma(acc1[ct][0], acc1[ct][1], a1, b);
					}
			}
		__syncwarp();
#pragma unroll
		for(int ct=0; ct<NCT; ct++)
			if(ct<n_ct)
				{
				const int cc = 8*ct + 2*q;
				if(r0<m)
					{
					if(cc<nx1) sW[r0*ld+cc] = acc0[ct][0];
					if(cc+1<nx1) sW[r0*ld+cc+1] = acc0[ct][1];
					}
				if(two && r1<m)
					{
					if(cc<nx1) sW[r1*ld+cc] = acc1[ct][0];
					if(cc+1<nx1) sW[r1*ld+cc+1] = acc1[ct][1];
					}
				}
		}
	hbt_sync();
	}

__device__ __forceinline__ void hbt_trmm_mma(double *sW, int ld, int tid, int m, int nx1, const double *prev, int nu1)
	{
	if(nx1<=16) hbt_trmm_mma_t<2>(sW, ld, tid, m, nx1, prev, nu1);
	else if(nx1<=32) hbt_trmm_mma_t<4>(sW, ld, tid, m, nx1, prev, nu1);
	else if(nx1<=48) hbt_trmm_mma_t<6>(sW, ld, tid, m, nx1, prev, nu1);
	else hbt_trmm_mma_t<8>(sW, ld, tid, m, nx1, prev, nu1);
	}

/* cur <- chol_mn(cur + W W') by the team, left-looking by panels of EIGHT columns: the panel's 8 x 8 row tiles (rows at or below the
 * panel) are dealt to the warps and accumulated by DMMA over the columns of W and over the finished columns of L, added to the
 * entries of H in place; warp 0 factorises the 8 x 8 diagonal block (pivot rule of kernel/c99/kernel_dpotrf_c99_lib4.c:553-640 as in
 * hb_chol) into scratch S (80 doubles), then the rows are handed to one thread each, which scales its row and stores it */
__device__ __forceinline__ void hbt_syrk_chol_mma(double *cur, const double *sW, int ld, int tid, int m, int nux, int nx1, double *S
#ifdef HBT_TIMING
		, long long &hbt_t0_
#endif
		)
	{
	const int warp = tid>>5, lane = tid&31, g = lane>>2, q = lane&3;
	const int n_rt = (m+7)>>3, nx1p = (nx1+3)&~3;
	double *dinv = cur + HB_TRI(nux) + nux;
	for(int jb=0; jb<nux; jb+=8)
		{
		HBT_T(8);
		const int ncol = nux-jb<8 ? nux-jb : 8;
		int rb = jb + g; rb = rb<nux ? rb : nux-1;                              /* row of W / L behind column jb+g of the panel */
		const double *pbw = sW + rb*ld + q, *pbl = cur + HB_TRI(rb) + q;
		for(int t=(jb>>3)+warp; t<n_rt; t+=4)
			{
			int ra = 8*t + g; ra = ra<m ? ra : m-1;
			const double *paw = sW + ra*ld + q, *pal = cur + HB_TRI(ra) + q;
			double c0 = 0.0, c1 = 0.0;
#pragma unroll 2
			for(int k0=0; k0<nx1p; k0+=4) hbt_dmma(c0, c1, paw[k0], pbw[k0]);
#pragma unroll 2
			for(int k0=0; k0<jb; k0+=4) hbt_dmma(c0, c1, -pal[k0], pbl[k0]);
			const int r = 8*t + g, cc = jb + 2*q;
			if(r<m)
				{
				/* the tile on the diagonal also leaves its entries in scratch: the rows of the diagonal block are overwritten by their
				 * threads while threads of other warps still need the block */
				double *h = cur + HB_TRI(r);
				const bool dg = (t==(jb>>3));
				if(cc<nux && cc<=r) { const double nv = h[cc] + c0; h[cc] = nv; if(dg) S[8*g+2*q] = nv; }
				if(cc+1<nux && cc+1<=r) { const double nv = h[cc+1] + c1; h[cc+1] = nv; if(dg) S[8*g+2*q+1] = nv; }
				}
			}
		HBT_T(4);
		hbt_sync();
		HBT_T(5);
		/* one thread per row of the panel (rows jb .. m-1).  The 8 x 8 diagonal block [D11 0; D21 D22] is taken in two 4 x 4 steps, each
		 * factorised by every thread for itself in registers (no exchange, four dependent pivots): columns jb..jb+3 of all rows first;
		 * then, with L21 (rows jb+4..jb+7 of those columns) read back, D22 - L21 L21' and columns jb+4..jb+7 */
		const int r = jb + tid;
		const bool mine = r<m;
		double v[8];
		if(mine)
			{
			const double *h = cur + HB_TRI(r) + jb;
#pragma unroll
			for(int cc=0; cc<8; cc++) v[cc] = (cc<ncol && jb+cc<=r) ? h[cc] : 0.0;
			double D[4][4], iv[4];
#pragma unroll
			for(int cc=0; cc<4; cc++)
#pragma unroll
				for(int c2=0; c2<=cc; c2++)
					D[cc][c2] = (cc<ncol) ? S[8*cc+c2] : (c2==cc ? 1.0 : 0.0);
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
				const double piv = D[cc][cc];
				const double inv = (piv>1e-15) ? hbg_rsqrt(piv) : 0.0;
				D[cc][cc] = piv*inv; iv[cc] = inv;
#pragma unroll
				for(int c2=cc+1; c2<4; c2++) D[c2][cc] *= inv;
#pragma unroll
				for(int c2=cc+1; c2<4; c2++)
#pragma unroll
					for(int c3=cc+1; c3<=c2; c3++) D[c2][c3] = fma(-D[c2][cc], D[c3][cc], D[c2][c3]);
				}
			double *hs = cur + HB_TRI(r) + jb;
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
				double y = v[cc];
#pragma unroll
				for(int c2=0; c2<cc; c2++) y = fma(-v[c2], D[cc][c2], y);
				v[cc] = y*iv[cc];
				if(cc<ncol && jb+cc<=r) hs[cc] = (r==jb+cc) ? D[cc][cc] : v[cc];
				}
			if(tid==0)
				{
#pragma unroll
				for(int cc=0; cc<4; cc++) if(cc<ncol) dinv[jb+cc] = iv[cc];
				}
			}
		hbt_sync();
		if(mine && ncol>4 && r>=jb+4)
			{
			double L21[4][4], D2[4][4], iv[4];
#pragma unroll
			for(int cc=0; cc<4; cc++)
#pragma unroll
				for(int c2=0; c2<4; c2++)
					{
					L21[cc][c2] = (cc+4<ncol) ? cur[HB_TRI(jb+4+cc)+jb+c2] : 0.0;
					if(c2<=cc) D2[cc][c2] = (cc+4<ncol) ? S[8*(cc+4)+4+c2] : (c2==cc ? 1.0 : 0.0);
					}
			/* D22 - L21 L21' (lower), own entries - x[0..3] L21' */
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
#pragma unroll
				for(int c2=0; c2<=cc; c2++)
#pragma unroll
					for(int k=0; k<4; k++) D2[cc][c2] = fma(-L21[cc][k], L21[c2][k], D2[cc][c2]);
#pragma unroll
				for(int k=0; k<4; k++) v[4+cc] = fma(-v[k], L21[cc][k], v[4+cc]);
				}
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
				const double piv = D2[cc][cc];
				const double inv = (piv>1e-15) ? hbg_rsqrt(piv) : 0.0;
				D2[cc][cc] = piv*inv; iv[cc] = inv;
#pragma unroll
				for(int c2=cc+1; c2<4; c2++) D2[c2][cc] *= inv;
#pragma unroll
				for(int c2=cc+1; c2<4; c2++)
#pragma unroll
					for(int c3=cc+1; c3<=c2; c3++) D2[c2][c3] = fma(-D2[c2][cc], D2[c3][cc], D2[c2][c3]);
				}
			double *hs = cur + HB_TRI(r) + jb + 4;
#pragma unroll
			for(int cc=0; cc<4; cc++)
				{
				double y = v[4+cc];
#pragma unroll
				for(int c2=0; c2<cc; c2++) y = fma(-v[4+c2], D2[cc][c2], y);
				v[4+cc] = y*iv[cc];
				if(cc+4<ncol && jb+4+cc<=r) hs[cc] = (r==jb+4+cc) ? D2[cc][cc] : v[4+cc];
				}
			if(tid==4)
				{
#pragma unroll
				for(int cc=0; cc<4; cc++) if(cc+4<ncol) dinv[jb+4+cc] = iv[cc];
				}
			}
		HBT_T(6);
		hbt_sync();
		HBT_T(7);
		}
	}

/* one backward stage by the team (the arguments of hb_stage_factor; c.lane is not used, tid is) */
template<bool GRAD>
__device__ __forceinline__ void hbt_stage_factor(const hb_ctx &c, int tid, const hb_stage &s, int nu1,
		const double *__restrict__ in_inst, const double *bvec, const double *rqvec,
		const double *Qx, const double *qx, const int *__restrict__ idxb, double *Pb,
		double *cur, const double *prev)
	{
	const int warp = tid>>5, lane = tid&31;
	const int nux = s.nu + s.nx, nx1 = s.nx1;
	const int m = GRAD ? nux+1 : nux;
	const int ntri = HB_TRI(nux) + (GRAD ? nux : 0);
	double *sW = c.sW; const int ldW = c.ldW;
	HBT_T0
	{
	const double *g = in_inst + s.off_RSQ;
	for(int e=tid; e<ntri; e+=HBT_THREADS) hb_cp8(cur + e, g + e);
	if(nx1>0)
		{
		const double *gb = in_inst + s.off_BAbt;
		const int tot = m*nx1;
		int i = tid/nx1, j = tid - i*nx1;
		const int di = HBT_THREADS/nx1, dj = HBT_THREADS - di*nx1;         /* one step of 128 elements in (row, column) */
		for(int e=tid; e<tot; e+=HBT_THREADS)
			{
			hb_cp8(sW + i*ldW + j, gb + e);
			i += di; j += dj;
			if(j>=nx1) { j -= nx1; i++; }
			}
		}
	hb_g2s_wait();
	}
	hbt_sync();
	HBT_T(0);
#ifndef HBT_FMA_PANELS
	{
	/* zero columns nx1 .. nx1p-1 of W: the DMMA k range is a multiple of 4 */
	const int npad = ((nx1+3)&~3) - nx1;
	if(npad>0) for(int e=tid; e<m*npad; e+=HBT_THREADS) { const int i = e/npad, j = e - i*npad; sW[i*ldW + nx1 + j] = 0.0; }
	}
#endif
	if(GRAD && rqvec!=nullptr)
		for(int e=tid; e<nux; e+=HBT_THREADS) cur[HB_TRI(nux)+e] = rqvec[s.off_ux+e];
	if(GRAD && bvec!=nullptr && nx1>0)
		for(int j=tid; j<nx1; j+=HBT_THREADS) sW[nux*ldW+j] = bvec[s.off_pi+j];
	hbt_sync();
	if(Qx!=nullptr && s.nb>0)
		{
		for(int j=tid; j<s.nb; j+=HBT_THREADS)
			{
			int id = idxb[s.off_c+j];
			cur[HB_TRI(id)+id] += Qx[s.off_c+j];
			if(GRAD && qx!=nullptr) cur[HB_TRI(nux)+id] += qx[s.off_c+j];
			}
		}
	if(Qx!=nullptr && s.ng>0)
		{
		/* general constraints: H += [D C]' diag(Qx_g) [D C] (see hb_stage_factor) on 8 x 8 DMMA tiles of the lower triangle, the
		 * fragments taken from the instance block (row-major nux x ng, L2-resident after the stage's loads); the gradient row
		 * += [D C]' qx_g with one thread per column */
		hbt_sync();
		const int ng = s.ng, ngp = (ng+3)&~3;
		const double *G = in_inst + s.off_DCt;
		const double *Qg = Qx + s.off_c + s.nb;
		{
		const int lane = tid&31, g = lane>>2, q = lane&3;
		const int n_t = (nux+7)>>3, ntile = n_t*(n_t+1)/2;
		for(int idx=warp; idx<ntile; idx+=HBT_WARPS)
			{
			int t = 0, rem = idx;                                           /* tile idx -> (row tile t, column tile p <= t) */
			while(rem>t) { rem -= t+1; t++; }
			const int p = rem;
			const int ra = 8*t + g, rb = 8*p + g;
			const double *pa = G + (ra<nux ? ra : nux-1)*ng, *pb = G + (rb<nux ? rb : nux-1)*ng;
			double c0 = 0.0, c1 = 0.0;
			for(int k0=0; k0<ngp; k0+=4)
				{
				const int kj = k0 + q;
				const bool in = kj<ng;
				const double a = in ? pa[kj] : 0.0;
				const double b = in ? pb[kj]*Qg[kj] : 0.0;
				hbt_dmma(c0, c1, a, b);
				}
			const int cc = 8*p + 2*q;
			if(ra<nux)
				{
				double *h = cur + HB_TRI(ra);
				if(cc<=ra) h[cc] += c0;
				if(cc+1<=ra) h[cc+1] += c1;
				}
			}
		}
		if(GRAD && qx!=nullptr)
			{
			const double *qg = qx + s.off_c + s.nb;
			double *hi = cur + HB_TRI(nux);
			for(int k=tid; k<nux; k+=HBT_THREADS)
				{
				const double *gk = G + k*ng;
				double acc = 0.0;
				for(int j=0; j<ng; j++) acc += qg[j]*gk[j];
				hi[k] += acc;
				}
			}
		}
	hbt_sync();
	HBT_T(1);
	if(nx1>0)
		{
		/* W = [B A b]' Lxx_{n+1}: eight column tiles per round, one per (warp, h), all rows of a tile in one lane set */
#ifdef HBT_FMA_PANELS
		hbg_trmm_any<true>(sW, ldW, tid, m, nx1, prev, nu1);
#else
		hbt_trmm_mma(sW, ldW, tid, m, nx1, prev, nu1);
#endif
		HBT_T(2);
		if(GRAD)
			{
			const double *wl = sW + nux*ldW;
			if(Pb!=nullptr)
				for(int i=tid; i<nx1; i+=HBT_THREADS)
					{
					double acc = 0.0;
					for(int k=0; k<=i; k++) acc += prev[HB_TRI(nu1+i)+nu1+k]*wl[k];
					Pb[s.off_pi+i] = acc;
					}
			hbt_sync();
			for(int j=tid; j<nx1; j+=HBT_THREADS) sW[nux*ldW+j] += prev[HB_TRI(nu1+nx1)+nu1+j];
			hbt_sync();
			}
		}
	HBT_T(3);
	/* a separate W W' pass on R-row tiles, panels dealt to the warps (hbt_syrk), was measured and is slower than accumulating the W part
	 * inside the panel loop (342 K against ~110 K cycles per factorisation at config 4, tools/phase_timing_team.py) */
#ifdef HBT_FMA_PANELS
#ifdef HBT_TIMING
	hbt_syrk_chol(cur, sW, ldW, tid, m, nux, nx1, c.sV, hbt_t0_);
#else
	hbt_syrk_chol(cur, sW, ldW, tid, m, nux, nx1, c.sV);
#endif
#else
#ifdef HBT_TIMING
	hbt_syrk_chol_mma(cur, sW, ldW, tid, m, nux, nx1, c.sV, hbt_t0_);
#else
	hbt_syrk_chol_mma(cur, sW, ldW, tid, m, nux, nx1, c.sV);
#endif
#endif
	}

/* backward sweep n = N..0 by the team; the factor of every stage goes to Lst (global) */
template<bool GRAD>
__device__ void hbt_backward(const hb_ctx &c, int tid, const hb_dims &d, const double *in_inst, double *Lst,
		const double *bvec, const double *rqvec, const double *Qx, const double *qx, double *Pb)
	{
	double *cur = c.bufA, *prev = c.bufB;
	for(int n=d.N; n>=0; n--)
		{
		const hb_stage s = d.st[n];
		const int nu1 = (n<d.N) ? d.st[n+1].nu : 0;
		if(n>0 && tid<32)
			{
			const hb_stage sp = d.st[n-1];
			hb_prefetch_l2(in_inst + sp.off_BAbt, HB_EVEN((sp.nu+sp.nx+1)*sp.nx1) + HB_TRI(sp.nu+sp.nx) + sp.nu+sp.nx, tid);
			}
		hbt_stage_factor<GRAD>(c, tid, s, nu1, in_inst, bvec, rqvec, Qx, qx, d.idxb, Pb, cur, prev);
		const int nL = HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx);
		double *dst = Lst + s.off_L;
		for(int e=tid; e<nL; e+=HBT_THREADS) dst[e] = cur[e];
		double *t = cur; cur = prev; prev = t;
		hbt_sync();
		}
	}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Solve sweeps by the team.  A stage of the forward / solve-only backward sweep is a handful of matrix-vector products with  */
/* 40-50 terms per entry; on one warp each is a dependent chain of that length, run twice when there are more than 32 entries.  */
/* Here warp q takes the terms k = k0 + q, k0 + q + 4, ... of every entry, the four partial sums meet in shared memory         */
/* (P: 4 x 64 doubles behind the one-warp layout).  The sums are therefore ordered differently from the one-warp sweeps:        */
/* results agree to rounding (tests/test_team.py: 1e-9 and identical IPM iteration counts), not bit for bit.                    */
/* ---------------------------------------------------------------------------------------------------------------- */
#define HBT_P_DOUBLES 256

__host__ __device__ inline int hbt_smem_doubles(int nzM, int nxM) { return hb_smem_doubles_per_warp(nzM, nxM) + HBT_P_DOUBLES; }

/* out(o, sum_{k0(o) <= k < k1(o)} a(o,k)*b(o,k)) for o < n_out <= 64; ends with a team barrier */
template<class K0, class K1, class A, class B, class FIN>
__device__ __forceinline__ void hbt_dot(int tid, int n_out, double *P, K0 k0, K1 k1, A a, B b, FIN fin)
	{
	const int q = tid>>5, lane = tid&31;
	for(int o=lane; o<n_out; o+=32)
		{
		double acc = 0.0;
		const int ke = k1(o);
#pragma unroll 4
		for(int k=k0(o)+q; k<ke; k+=4) acc = fma(a(o, k), b(o, k), acc);
		P[64*q+o] = acc;
		}
	hbt_sync();
	for(int o=tid; o<n_out; o+=HBT_THREADS) fin(o, (P[o] + P[64+o]) + (P[128+o] + P[192+o]));
	hbt_sync();
	}

/* pi of a stage from x_{n+1} (in c.sV + 64, left there by hbt_stage_forward) and the factor of stage n+1 (nx1, nu1, nux1: its sizes):
 * pi = (trs ? pi : 0) + Lxx (Lxx' x + (trs ? 0 : gradient row)) */
__device__ __forceinline__ void hbt_stage_pi(const hb_ctx &c, int tid, double *P, int nx1, int nu1, int nux1, const double *Ln1, bool trs, double *pi_n)
	{
	double *xs = c.sV + 64, *tmp = c.sV + 128;
	hbt_dot(tid, nx1, P, [&](int o) { return o; }, [&](int) { return nx1; },
			[&](int o, int k) { return Ln1[HB_TRI(nu1+k)+nu1+o]; }, [&](int, int k) { return xs[k]; },
			[&](int o, double sum) { tmp[o] = (trs ? 0.0 : Ln1[HB_TRI(nux1)+nu1+o]) + sum; });
	hbt_dot(tid, nx1, P, [&](int) { return 0; }, [&](int o) { return o+1; },
			[&](int o, int k) { return Ln1[HB_TRI(nu1+o)+nu1+k]; }, [&](int, int k) { return tmp[k]; },
			[&](int o, double sum) { pi_n[o] = (trs ? pi_n[o] : 0.0) + sum; });
	}

/* one forward stage by the team: the arguments and the algebra of hb_stage_forward (Ln1 == nullptr: pi is left to the caller) (lqcp_solvers/d_back_ric_rec.c:341-397 sv,
 * :737-789 trs); the triangular solve with the ks x ks corner stays on warp 0 */
__device__ __forceinline__ void hbt_stage_forward(const hb_ctx &c, int tid, double *P, const hb_stage &s, const hb_stage &s1, int n,
		const double *Ln, const double *Ln1, const double *lrow, const double *bvec, bool trs,
		double *ux, double *pi, bool compute_pi)
	{
	const int lane = tid&31;
	const int nu = s.nu, nux = s.nu+s.nx, nx1 = s.nx1, nu1 = s1.nu, nux1 = s1.nu+s1.nx;
	const int ks = (n==0) ? nux : nu;
	const int ld = c.ldW;
	const double *dinv = Ln + HB_TRI(nux) + nux;
	const double *sW = c.sW;
	double *v = c.sV, *xs = c.sV + 64;
	for(int i=tid; i<nux; i+=HBT_THREADS)
		v[i] = (i<ks) ? -(lrow!=nullptr ? lrow[s.off_ux+i] : Ln[HB_TRI(nux)+i]) : ux[s.off_ux+i];
	hbt_sync();
	if(ks<nux)
		hbt_dot(tid, ks, P, [&](int) { return ks; }, [&](int) { return nux; },
				[&](int o, int k) { return Ln[HB_TRI(k)+o]; }, [&](int, int k) { return v[k]; },
				[&](int o, double sum) { v[o] -= sum; });
	if(tid<32)
		for(int j=ks-1; j>=0; j--)
			{
			if(lane==(j&31)) v[j] *= dinv[j];
			__syncwarp();
			const double vj = v[j];
			for(int i=lane; i<j; i+=32) v[i] -= Ln[HB_TRI(j)+i]*vj;
			__syncwarp();
			}
	hbt_sync();
	for(int i=tid; i<ks; i+=HBT_THREADS) ux[s.off_ux+i] = v[i];
	hbt_dot(tid, nx1, P, [&](int) { return 0; }, [&](int) { return nux; },
			[&](int o, int k) { return sW[k*ld+o]; }, [&](int, int k) { return v[k]; },
			[&](int o, double sum)
				{
				const double acc = (bvec!=nullptr ? bvec[s.off_pi+o] : sW[nux*ld+o]) + sum;
				if(trs && compute_pi) pi[s.off_pi+o] = ux[s1.off_ux+nu1+o];
				ux[s1.off_ux+nu1+o] = acc;
				xs[o] = acc;
				});
	if(compute_pi && Ln1!=nullptr) hbt_stage_pi(c, tid, P, nx1, nu1, nux1, Ln1, trs, pi + s.off_pi);
	}

/* stage data of the team's solve sweeps: packed factor(s) and [B A b]' by all threads (LDGSTS) */
__device__ __forceinline__ void hbt_g2s(int tid, double *dst_smem, const double *__restrict__ src, int n)
	{
	for(int e=tid; e<n; e+=HBT_THREADS) hb_cp8(dst_smem + e, src + e);
	}
__device__ __forceinline__ void hbt_load_BAbt(const hb_ctx &c, int tid, const hb_stage &s, const double *__restrict__ in_inst)
	{
	const int nx1 = s.nx1, tot = (s.nu+s.nx+1)*nx1;
	if(nx1<=0) return;
	const double *gb = in_inst + s.off_BAbt;
	int i = tid/nx1, j = tid - i*nx1;
	const int di = HBT_THREADS/nx1, dj = HBT_THREADS - di*nx1;
	for(int e=tid; e<tot; e+=HBT_THREADS)
		{
		hb_cp8(c.sW + i*c.ldW + j, gb + e);
		i += di; j += dj;
		if(j>=nx1) { j -= nx1; i++; }
		}
	}

/* forward sweep n = 0..N-1 by the team (hb_forward) */
static __device__ void hbt_forward(const hb_ctx &c, int tid, double *P, const hb_dims &d, const double *in_inst, const double *Lst,
		const double *lrow, const double *bvec, bool trs, double *ux, double *pi, bool compute_pi)
	{
	double *a = c.bufA, *b = c.bufB;
	{
	const hb_stage s0 = d.st[0];
	hbt_g2s(tid, a, Lst + s0.off_L, HB_TRI(s0.nu+s0.nx) + 2*(s0.nu+s0.nx));
	}
	for(int n=0; n<d.N; n++)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		if(n+2<=d.N && tid<32)
			{
			const hb_stage s2 = d.st[n+2];
			hb_prefetch_l2(Lst + s2.off_L, HB_TRI(s2.nu+s2.nx) + 2*(s2.nu+s2.nx), tid);
			if(n+1<d.N) hb_prefetch_l2(in_inst + s1.off_BAbt, (s1.nu+s1.nx+1)*s1.nx1, tid);
			}
		hbt_g2s(tid, b, Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx));
		hbt_load_BAbt(c, tid, s, in_inst);
		hb_g2s_wait();
		hbt_sync();
		hbt_stage_forward(c, tid, P, s, s1, n, a, b, lrow, bvec, trs, ux, pi, compute_pi);
		double *t = a; a = b; b = t;
		}
	}

/* forward sweep with ONE factor buffer: stage n needs L_n only when pi_{n-1} = f(L_n, x_n) is taken at stage n instead of stage n-1
 * (and pi_{N-1} after the loop, behind a last load of L_N).  With [B A b]' that is 30 KB of stage data per instance instead of 41:
 * seven CTAs per SM instead of five for the kernels that do not factorise.  Same operations in the same order per entry as
 * hbt_forward */
static __device__ void hbt_forward1(const hb_ctx &c, int tid, double *P, const hb_dims &d, const double *in_inst, const double *Lst,
		const double *lrow, const double *bvec, bool trs, double *ux, double *pi, bool compute_pi)
	{
	double *a = c.bufA;
	for(int n=0; n<d.N; n++)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		if(tid<32)
			{
			hb_prefetch_l2(Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx), tid);
			if(n+1<d.N) hb_prefetch_l2(in_inst + s1.off_BAbt, (s1.nu+s1.nx+1)*s1.nx1, tid);
			}
		hbt_g2s(tid, a, Lst + s.off_L, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		hbt_load_BAbt(c, tid, s, in_inst);
		hb_g2s_wait();
		hbt_sync();
		if(n>0 && compute_pi) hbt_stage_pi(c, tid, P, s.nx, s.nu, s.nu+s.nx, a, trs, pi + d.st[n-1].off_pi);
		hbt_stage_forward(c, tid, P, s, s1, n, a, nullptr, lrow, bvec, trs, ux, pi, compute_pi);
		}
	if(compute_pi && d.N>0)
		{
		const hb_stage s = d.st[d.N];
		hbt_g2s(tid, a, Lst + s.off_L, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		hb_g2s_wait();
		hbt_sync();
		hbt_stage_pi(c, tid, P, s.nx, s.nu, s.nu+s.nx, a, trs, pi + d.st[d.N-1].off_pi);
		}
	}

/* the layout of the kernels that keep one factor buffer: [L | W | vectors | P] */
__host__ __device__ inline int hbt_smem1_doubles(int nzM, int nxM)
	{
	return HB_EVEN(HB_TRI(nzM) + 2*nzM) + HB_EVEN(nzM*HB_LDW(nxM)) + 192 + HBT_P_DOUBLES;
	}
__device__ __forceinline__ hb_ctx hbt_make_ctx1(const hb_dims &d, double *smem, int lane, double *&P)
	{
	hb_ctx c;
	c.lane = lane;
	c.ldW = HB_LDW(d.nxM);
	c.bufA = smem;
	c.bufB = smem;                                                     /* not used by these kernels */
	c.sW = smem + HB_EVEN(HB_TRI(d.nzM) + 2*d.nzM);
	c.sV = c.sW + HB_EVEN(d.nzM*c.ldW);
	P = c.sV + 192;
	return c;
	}

/* one backward stage of the solve-only sweep by the team (hb_trs_stage_back; lqcp_solvers/d_back_ric_rec.c:628-732) */
__device__ __forceinline__ void hbt_trs_stage_back(const hb_ctx &c, int tid, double *P, const hb_stage &s, const hb_stage &s1, int n,
		const double *Ln, const double *Ln1, const double *bvec, const double *rqvec, const double *qx,
		const int *__restrict__ idxb, double *ux, double *Pb, bool compute_Pb, const double *__restrict__ in_inst)
	{
	const int lane = tid&31;
	const int nu = s.nu, nux = s.nu+s.nx, nx1 = s.nx1, nu1 = s1.nu;
	const int ks = (n==0) ? nux : nu;
	const int ld = c.ldW;
	const double *dinv = Ln + HB_TRI(nux) + nux;
	const double *sW = c.sW;
	double *v = c.sV, *tmp = c.sV + 64, *t2 = c.sV + 128;
	if(compute_Pb)
		{
		hbt_dot(tid, nx1, P, [&](int o) { return o; }, [&](int) { return nx1; },
				[&](int o, int k) { return Ln1[HB_TRI(nu1+k)+nu1+o]; }, [&](int, int k) { return bvec[s.off_pi+k]; },
				[&](int o, double sum) { t2[o] = sum; });
		hbt_dot(tid, nx1, P, [&](int) { return 0; }, [&](int o) { return o+1; },
				[&](int o, int k) { return Ln1[HB_TRI(nu1+o)+nu1+k]; }, [&](int, int k) { return t2[k]; },
				[&](int o, double sum) { Pb[s.off_pi+o] = sum; });
		}
	for(int i=tid; i<nux; i+=HBT_THREADS) v[i] = rqvec[s.off_ux+i];
	for(int j=tid; j<nx1; j+=HBT_THREADS) tmp[j] = Pb[s.off_pi+j] + ux[s1.off_ux+nu1+j];
	hbt_sync();
	if(qx!=nullptr && s.nb>0)
		{
		for(int j=tid; j<s.nb; j+=HBT_THREADS) v[idxb[s.off_c+j]] += qx[s.off_c+j];
		hbt_sync();
		}
	if(qx!=nullptr && s.ng>0 && in_inst!=nullptr)
		{
		const double *G = in_inst + s.off_DCt, *qg = qx + s.off_c + s.nb;
		for(int i=tid; i<nux; i+=HBT_THREADS)
			{
			double acc = v[i];
			for(int j=0; j<s.ng; j++) acc += G[i*s.ng+j]*qg[j];
			v[i] = acc;
			}
		hbt_sync();
		}
	hbt_dot(tid, nux, P, [&](int) { return 0; }, [&](int) { return nx1; },
			[&](int o, int k) { return sW[o*ld+k]; }, [&](int, int k) { return tmp[k]; },
			[&](int o, double sum) { v[o] += sum; });
	if(tid<32)
		for(int j=0; j<ks; j++)
			{
			if(lane==(j&31)) v[j] *= dinv[j];
			__syncwarp();
			const double vj = v[j];
			for(int i=j+1+lane; i<nux; i+=32) v[i] -= Ln[HB_TRI(i)+j]*vj;
			__syncwarp();
			}
	hbt_sync();
	for(int i=tid; i<nux; i+=HBT_THREADS) ux[s.off_ux+i] = v[i];
	hbt_sync();
	(void)nu;
	}

/* solve-only backward vector sweep by the team; w is kept in ux (hb_trs_backward) */
static __device__ void hbt_trs_backward(const hb_ctx &c, int tid, double *P, const hb_dims &d, const double *in_inst, const double *Lst,
		const double *bvec, const double *rqvec, const double *qx, double *ux, double *Pb, bool compute_Pb)
	{
	{
	const hb_stage s = d.st[d.N];
	const int nux = s.nu+s.nx;
	for(int i=tid; i<nux; i+=HBT_THREADS) ux[s.off_ux+i] = rqvec[s.off_ux+i];
	hbt_sync();
	if(qx!=nullptr) for(int j=tid; j<s.nb; j+=HBT_THREADS) ux[s.off_ux+d.idxb[s.off_c+j]] += qx[s.off_c+j];
	hbt_sync();
	if(qx!=nullptr && s.ng>0)
		{
		const double *G = in_inst + s.off_DCt, *qg = qx + s.off_c + s.nb;
		for(int i=tid; i<nux; i+=HBT_THREADS)
			{
			double acc = ux[s.off_ux+i];
			for(int j=0; j<s.ng; j++) acc += G[i*s.ng+j]*qg[j];
			ux[s.off_ux+i] = acc;
			}
		hbt_sync();
		}
	}
	for(int n=d.N-1; n>=0; n--)
		{
		const hb_stage s = d.st[n];
		const hb_stage s1 = d.st[n+1];
		if(n>0 && tid<32)
			{
			const hb_stage sp = d.st[n-1];
			hb_prefetch_l2(Lst + sp.off_L, HB_TRI(sp.nu+sp.nx) + 2*(sp.nu+sp.nx), tid);
			hb_prefetch_l2(in_inst + sp.off_BAbt, (sp.nu+sp.nx+1)*sp.nx1, tid);
			}
		hbt_g2s(tid, c.bufA, Lst + s.off_L, HB_TRI(s.nu+s.nx) + 2*(s.nu+s.nx));
		if(compute_Pb) hbt_g2s(tid, c.bufB, Lst + s1.off_L, HB_TRI(s1.nu+s1.nx) + 2*(s1.nu+s1.nx));
		hbt_load_BAbt(c, tid, s, in_inst);
		hb_g2s_wait();
		hbt_sync();
		hbt_trs_stage_back(c, tid, P, s, s1, n, c.bufA, c.bufB, bvec, rqvec, qx, d.idxb, ux, Pb, compute_Pb, in_inst);
		}
	}
