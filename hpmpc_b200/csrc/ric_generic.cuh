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
/* leading dimension of the shared-memory W / [B A b]' rows: odd (row-per-lane accesses are conflict-free) and at least nx rounded up
 * to a multiple of 4 (the FP64 tensor-core path of ric_team.cuh pads the k range of a stage with zero columns) */
#define HB_LDW(nxM) (((((nxM)+3)&~3))|1)

struct hb_ctx
	{
	int lane;
	int ldW;
	double *bufA, *bufB;     /* packed L buffers, tri(nzM)+2*nzM each */
	double *sW;              /* nzM x ldW */
	double *sV;              /* 4*nzM scratch vectors */
	};

__device__ __forceinline__ int hb_Lsize(int nux) { return HB_EVEN(HB_TRI(nux) + 2*nux); }

/* global -> shared staging without a register round trip (LDGSTS): a lane puts ALL its 8-byte copies of a stage in flight at once
 * (the plain `dst[e] = src[e]` loops keep ~4 loads per lane in flight, which at one warp per scheduler is ~1 KB per warp against
 * a ~700-cycle round trip); hb_g2s_wait() + __syncwarp() before the data is used */
__device__ __forceinline__ void hb_cp8(double *dst_smem, const double *src)
	{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
	}
__device__ __forceinline__ void hb_g2s_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void hb_g2s(int lane, double *dst_smem, const double *__restrict__ src, int n)
	{
	for(int e=lane; e<n; e+=32) hb_cp8(dst_smem + e, src + e);
	}
/* rows of nx1 doubles to rows of ld doubles */
__device__ __forceinline__ void hb_g2s_rows(int lane, double *dst_smem, int ld, const double *__restrict__ src, int nrow, int nx1)
	{
	if(nx1<=0) return;
	int i = lane/nx1, j = lane - i*nx1;
	const int di = 32/nx1, dj = 32 - di*nx1;                              /* one step of 32 elements in (row, column) */
	const int tot = nrow*nx1;
	for(int e=lane; e<tot; e+=32)
		{
		hb_cp8(dst_smem + i*ld + j, src + e);
		i += di; j += dj;
		if(j>=nx1) { j -= nx1; i++; }
		}
	}

/* 1/sqrt(p): hardware seed (MUFU.RSQ64H) and two coupled Goldschmidt steps, as in the size-specialised kernels (ric_blk.cuh) */
__device__ __forceinline__ double hbg_rsqrt(double p)
	{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
	const double hp = 0.5*p;
	double g = hp*y;
	double r = fma(-g, y, 0.5);
	g = fma(g, r, g); y = fma(y, r, y);
	r = fma(-g, y, 0.5);
	y = fma(y, r, y);
	return y;
	}

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

/* ---------------------------------------------------------------------------------------------------------------- */
/* Register-tiled stage routines (round 2).  The row-per-lane loops above and below need two shared-memory loads per DFMA  */
/* and leave the lanes beyond the row count idle; here the warp is 16 row slots x 2: lane (rt, h) owns rows rt, rt+16,     */
/* rt+32, rt+48 (R of them, R = ceil(m/16)) and four columns at a time, i.e. an R x 4 register tile fed by R + 4 loads per  */
/* k.  Rows 16 apart keep the row operand conflict-free both in the odd-stride W rows and in the packed triangle           */
/* (i(i+1)/2 mod 16 is a permutation of 0..15 over 16 consecutive i).                                                       */
/* ---------------------------------------------------------------------------------------------------------------- */

/* W <- W * Lxx_{n+1} in place (m x nx1, Lxx lower nx1 x nx1 inside the packed prev); h picks the column tile of a round.
 * W[i][j] = sum_{k>=j} W[i][k] Lxx[k][j]: a round reads columns >= its own tiles only, and a lane's reads of its partner's
 * tile happen before the barrier that precedes the writes */
/* TEAM: called by all 128 threads of a four-warp team (ric_team.cuh): a round is eight column tiles, one per (warp, h) */
template<bool TEAM> __device__ __forceinline__ void hbg_sync() { if(TEAM) __syncthreads(); else __syncwarp(); }
template<int R, bool TEAM>
__device__ __noinline__ void hbg_trmm(double *sW, int ld, int tid, int m, int nx1, const double *prev, int nu1)
	{
	const int lane = tid&31, rt = lane>>1, h = lane&1;
	const int slot = TEAM ? 2*(tid>>5) + h : h, nslot = TEAM ? 8 : 2;
	int row[R];
#pragma unroll
	for(int r=0; r<R; r++) { int i = rt + 16*r; row[r] = (i<m ? i : m-1)*ld; }
	const int ntile = (nx1+3)>>2;
	for(int t0=0; t0<ntile; t0+=nslot)
		{
		const int t = t0 + slot, j0 = 4*t;
		double acc[R][4];
#pragma unroll
		for(int r=0; r<R; r++) { acc[r][0] = 0.0; acc[r][1] = 0.0; acc[r][2] = 0.0; acc[r][3] = 0.0; }
		if(t<ntile)
			{
#pragma unroll
			for(int kk=0; kk<4; kk++)
				{
				const int k = j0 + kk;
				if(k<nx1)
					{
					const double *lk = prev + HB_TRI(nu1+k) + nu1 + j0;
					double a[R];
#pragma unroll
					for(int r=0; r<R; r++) a[r] = sW[row[r]+k];
#pragma unroll
					for(int cc=0; cc<=kk; cc++)
						{
						const double b = lk[cc];
#pragma unroll
						for(int r=0; r<R; r++) acc[r][cc] = fma(a[r], b, acc[r][cc]);
						}
					}
				}
#pragma unroll 2
			for(int k=j0+4; k<nx1; k++)
				{
				const double *lk = prev + HB_TRI(nu1+k) + nu1 + j0;
				const double b0 = lk[0], b1 = lk[1], b2 = lk[2], b3 = lk[3];
#pragma unroll
				for(int r=0; r<R; r++)
					{
					const double a = sW[row[r]+k];
					acc[r][0] = fma(a, b0, acc[r][0]); acc[r][1] = fma(a, b1, acc[r][1]);
					acc[r][2] = fma(a, b2, acc[r][2]); acc[r][3] = fma(a, b3, acc[r][3]);
					}
				}
			}
		hbg_sync<TEAM>();
		if(t<ntile)
			{
#pragma unroll
			for(int r=0; r<R; r++)
				if(rt+16*r<m)
					{
#pragma unroll
					for(int cc=0; cc<4; cc++) if(j0+cc<nx1) sW[row[r]+j0+cc] = acc[r][cc];
					}
			}
		hbg_sync<TEAM>();
		}
	}

/* cur <- chol_mn(cur + W W') by panels of four columns, left-looking: the panel's entries are accumulated in registers over
 * the nx1 columns of W and over the columns of L already done (h splits both sums by the parity of k, one shuffle joins them),
 * the 4 x 4 diagonal block is factorised by every lane on its own, the panel is scaled and stored.  Pivot rule of
 * kernel/c99/kernel_dpotrf_c99_lib4.c:553-640 as in hb_chol.  m = nux (+1 with the gradient row, which is row nux of W and of
 * the packed cur) */
template<int R>
__device__ __noinline__ void hbg_syrk_chol(double *cur, const double *sW, int ld, int lane, int m, int nux, int nx1)
	{
	const int rt = lane>>1, h = lane&1;
	int irow[R], wrow[R], lrow[R];
#pragma unroll
	for(int r=0; r<R; r++) { int i = rt + 16*r; i = i<m ? i : m-1; irow[r] = i; wrow[r] = i*ld; lrow[r] = HB_TRI(i); }
	double *dinv = cur + HB_TRI(nux) + nux;
	for(int jb=0; jb<nux; jb+=4)
		{
		int wj[4], lj[4];
#pragma unroll
		for(int cc=0; cc<4; cc++) { int j = jb+cc; j = j<nux ? j : nux-1; wj[cc] = j*ld; lj[cc] = HB_TRI(j); }
		const int rlo = jb>>4;
		double acc[R][4];
#pragma unroll
		for(int r=0; r<R; r++) { acc[r][0] = 0.0; acc[r][1] = 0.0; acc[r][2] = 0.0; acc[r][3] = 0.0; }
#pragma unroll 2
		for(int k=h; k<nx1; k+=2)
			{
			const double b0 = sW[wj[0]+k], b1 = sW[wj[1]+k], b2 = sW[wj[2]+k], b3 = sW[wj[3]+k];
#pragma unroll
			for(int r=0; r<R; r++)
				if(r>=rlo)
					{
					const double a = sW[wrow[r]+k];
					acc[r][0] = fma(a, b0, acc[r][0]); acc[r][1] = fma(a, b1, acc[r][1]);
					acc[r][2] = fma(a, b2, acc[r][2]); acc[r][3] = fma(a, b3, acc[r][3]);
					}
			}
		/* the W part: its two halves joined and added to the entries of H (the team version does this in a pass of its own and
		 * stores the sum, hbt_syrk: same operations in the same order) */
		double hw[R][4];
#pragma unroll
		for(int r=0; r<R; r++)
			if(r>=rlo)
				{
#pragma unroll
				for(int cc=0; cc<4; cc++)
					{
					double v = acc[r][cc] + __shfl_xor_sync(HB_FULL, acc[r][cc], 1);
					if((jb+cc<nux) && (jb+cc<=irow[r])) v += cur[lrow[r]+jb+cc];
					hw[r][cc] = v;
					acc[r][cc] = 0.0;
					}
				}
#pragma unroll 2
		for(int k=h; k<jb; k+=2)
			{
			const double b0 = cur[lj[0]+k], b1 = cur[lj[1]+k], b2 = cur[lj[2]+k], b3 = cur[lj[3]+k];
#pragma unroll
			for(int r=0; r<R; r++)
				if(r>=rlo)
					{
					const double a = -cur[lrow[r]+k];
					acc[r][0] = fma(a, b0, acc[r][0]); acc[r][1] = fma(a, b1, acc[r][1]);
					acc[r][2] = fma(a, b2, acc[r][2]); acc[r][3] = fma(a, b3, acc[r][3]);
					}
			}
		/* join the two halves of the L part, add, park the panel so that the diagonal block can be read by every lane */
#pragma unroll
		for(int r=0; r<R; r++)
			if(r>=rlo)
				{
#pragma unroll
				for(int cc=0; cc<4; cc++)
					acc[r][cc] = (acc[r][cc] + __shfl_xor_sync(HB_FULL, acc[r][cc], 1)) + hw[r][cc];
				}
		__syncwarp();
#pragma unroll
		for(int r=0; r<R; r++)
			if(r>=rlo && h==0 && irow[r]<jb+4)                                  /* only the rows of the diagonal block are read back */
				{
#pragma unroll
				for(int cc=0; cc<4; cc++) if((jb+cc<nux) && (jb+cc<=irow[r])) cur[lrow[r]+jb+cc] = acc[r][cc];
				}
		__syncwarp();
		double D[4][4], dd[4], iv[4];
#pragma unroll
		for(int cc=0; cc<4; cc++)
#pragma unroll
			for(int c2=0; c2<=cc; c2++)
				D[cc][c2] = (jb+cc<nux) ? cur[lj[cc]+jb+c2] : (c2==cc ? 1.0 : 0.0);
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
		__syncwarp();
#pragma unroll
		for(int r=0; r<R; r++)
			if(r>=rlo)
				{
				double x[4];
#pragma unroll
				for(int cc=0; cc<4; cc++)
					{
					double v = acc[r][cc];
#pragma unroll
					for(int c2=0; c2<cc; c2++) v = fma(-x[c2], D[cc][c2], v);
					x[cc] = v*iv[cc];
					}
				if(rt+16*r<m)
					{
#pragma unroll
					for(int cc=0; cc<4; cc++)
						if((cc>>1)==h && jb+cc<nux && jb+cc<=irow[r])
							cur[lrow[r]+jb+cc] = (irow[r]==jb+cc) ? dd[cc] : x[cc];
					}
				}
		if(lane==0)
			{
			dinv[jb] = iv[0];
			if(jb+1<nux) dinv[jb+1] = iv[1];
			if(jb+2<nux) dinv[jb+2] = iv[2];
			if(jb+3<nux) dinv[jb+3] = iv[3];
			}
		__syncwarp();
		}
	}

template<bool TEAM>
__device__ __forceinline__ void hbg_trmm_any(double *sW, int ld, int tid, int m, int nx1, const double *prev, int nu1)
	{
	if(m<=16) hbg_trmm<1, TEAM>(sW, ld, tid, m, nx1, prev, nu1);
	else if(m<=32) hbg_trmm<2, TEAM>(sW, ld, tid, m, nx1, prev, nu1);
	else if(m<=48) hbg_trmm<3, TEAM>(sW, ld, tid, m, nx1, prev, nu1);
	else hbg_trmm<4, TEAM>(sW, ld, tid, m, nx1, prev, nu1);
	}
__device__ __forceinline__ void hbg_syrk_chol_any(const hb_ctx &c, double *cur, int m, int nux, int nx1)
	{
	if(m<=16) hbg_syrk_chol<1>(cur, c.sW, c.ldW, c.lane, m, nux, nx1);
	else if(m<=32) hbg_syrk_chol<2>(cur, c.sW, c.ldW, c.lane, m, nux, nx1);
	else if(m<=48) hbg_syrk_chol<3>(cur, c.sW, c.ldW, c.lane, m, nux, nx1);
	else hbg_syrk_chol<4>(cur, c.sW, c.ldW, c.lane, m, nux, nx1);
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
	hb_g2s(lane, cur, g, ntri);
	if(nx1>0) hb_g2s_rows(lane, c.sW, c.ldW, in_inst + s.off_BAbt, m, nx1);
	hb_g2s_wait();
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
		if(GRAD && bvec!=nullptr)
			{
			for(int j=lane; j<nx1; j+=32) sW[nux*ldW+j] = bvec[s.off_pi+j];
			__syncwarp();
			}
#ifdef HB_GENERIC_ROWWISE
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
#else
		hbg_trmm_any<false>(c.sW, c.ldW, c.lane, m, nx1, prev, nu1);
#endif
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
#ifdef HB_GENERIC_ROWWISE
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
#endif
		}
#ifdef HB_GENERIC_ROWWISE
	hb_chol(c, cur, m, nux);
#else
	hbg_syrk_chol_any(c, cur, m, nux, nx1);           /* H + W W' is formed panel by panel inside the factorisation */
#endif
	}

/* copy a packed factor between smem and the stash (global) */
__device__ __forceinline__ void hb_copy(const hb_ctx &c, double *dst, const double *src, int n)
	{
	for(int e=c.lane; e<n; e+=32) dst[e] = src[e];
	}

/* the same, asynchronous: hb_g2s_wait() + __syncwarp() before use */
__device__ __forceinline__ void hb_load_BAbt_async(const hb_ctx &c, const hb_stage &s, const double *__restrict__ in_inst)
	{
	hb_g2s_rows(c.lane, c.sW, c.ldW, in_inst + s.off_BAbt, s.nu+s.nx+1, s.nx1);
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
