/*
 * pcond_kernels.cu -- partial condensing on the device: one warp per (OCP instance, block of stages).
 *
 * The horizon is cut into N2 blocks; inside a block the states are eliminated, so the block becomes ONE stage of a shorter problem
 * whose inputs are [u_{T-1} .. u_1 u_0] (newest first) and whose state is the state of the block's first stage.  Bounds on the
 * eliminated states become general constraints of the condensed stage.  The kernels read the full problem in the native packed
 * layout (layout.h) and write the condensed problem in the same layout of the condensed size pattern, so the condensed batch goes
 * straight into the IPM (ipm_kernels.cu, general-constraint path) without a host round trip; the expansion kernel maps the
 * condensed solution back (inputs copied, states simulated, lam / t by position, pi by the backward stationarity recursion).
 *
 * Restates (reference paths relative to /root/reference):
 *   hb_pcond_kernel     lqcp_solvers/d_part_cond.c:214-309 (d_cond_BAbt: Gamma_j = [B_j' ; Gamma_{j-1} A_j'] + b_j)
 *                                                  :312-574 (d_cond_RSQrq: pL_s = RSQrq_s + (BAbt_s Lx)(BAbt_s Lx)', Lx = chol of the
 *                                                            state block of pL_{s+1} with the gradient row; D / M / m placement)
 *                                                  :579-689 (d_cond_DCtd: which bounds stay bounds, rows of [D C]', shifted bounds)
 *                                                  :926-1066 (d_part_cond: block loop, last stage shared)
 *   hb_pexpand_kernel   lqcp_solvers/d_part_cond.c:1103-1306 (d_part_expand_solution)
 *   hb_res_norms_kernel interfaces/c/fortran_order_interface.c:616-656 (exit norms over the full problem)
 */
#include <cuda_runtime.h>
#include "pcond_launch.h"
#include "ric_generic.cuh"
#include "launch_util.cuh"

__device__ __forceinline__ int pc_even(int x) { return (x+1)&~1; }

/* Gamma_j of the block: (inputs of stages 0..j + nx0 + 1) x nx_{j+1}, row-major; j < T-1 lives in scratch, j = T-1 IS [B A b]' of
 * the condensed stage */
struct pc_gamma { double *p; int rows; };
__device__ __forceinline__ pc_gamma pc_gam(const hb_dims &dF, int n0, int T, int j, double *G, double *BAbt2)
	{
	int off = 0, nuc = 0;
	const int nx0 = dF.st[n0].nx;
	for(int i=0; i<j; i++) { nuc += dF.st[n0+i].nu; off += (nuc+nx0+1)*dF.st[n0+i].nx1; }
	nuc += dF.st[n0+j].nu;
	pc_gamma g; g.rows = nuc+nx0+1; g.p = (j==T-1) ? BAbt2 : G + off;
	return g;
	}

__global__ void __launch_bounds__(128) hb_pcond_kernel(hb_dims dF, hb_dims dC, const hb_pc_block *__restrict__ blk, int N2, long long n_inst,
		const double *__restrict__ in_full, double *__restrict__ in_cond, double *__restrict__ scratch, long long sstride, int szG)
	{
	const int lane = threadIdx.x & 31, wpb = blockDim.x>>5;
	const long long warp = (long long)blockIdx.x*wpb + (threadIdx.x>>5), nwarps = (long long)gridDim.x*wpb;
	double *G = scratch + warp*sstride;
	double *pL = G + szG;
	double *Lx = pL + pc_even(HB_TRI(dF.nzM)+dF.nzM);
	double *W = Lx + pc_even(HB_TRI(dF.nxM+1)+2*dF.nxM+2);
	hb_ctx cx; cx.lane = lane; cx.ldW = 0; cx.bufA = cx.bufB = cx.sW = cx.sV = nullptr;
	for(long long item=warp; item<n_inst*N2; item+=nwarps)
		{
		const long long inst = item/N2; const int k = (int)(item - inst*N2);
		const double *inF = in_full + inst*dF.in_stride;
		double *inC = in_cond + inst*dC.in_stride;
		const int n0 = blk[k].n0, T = blk[k].T;
		const hb_stage sc = dC.st[k];
		const int nx0 = sc.nx, nux2 = sc.nu+sc.nx;
		double *BAbt2 = inC + sc.off_BAbt, *H2 = inC + sc.off_RSQ;
		/* ---- Gamma_0 .. Gamma_{T-1} ---- */
		for(int j=0; j<T; j++)
			{
			const hb_stage s = dF.st[n0+j];
			const double *M = inF + s.off_BAbt;
			const pc_gamma g = pc_gam(dF, n0, T, j, G, BAbt2);
			if(j==0)
				{ for(int i=lane; i<g.rows*s.nx1; i+=32) g.p[i] = M[i]; }
			else
				{
				const pc_gamma gp = pc_gam(dF, n0, T, j-1, G, BAbt2);
				for(int idx=lane; idx<g.rows*s.nx1; idx+=32)
					{
					const int i = idx/s.nx1, c = idx - i*s.nx1;
					double v;
					if(i<s.nu) v = M[idx];
					else
						{
						const double *gr = gp.p + (i-s.nu)*s.nx;
						v = 0.0;
						for(int m=0; m<s.nx; m++) v = fma(gr[m], M[(s.nu+m)*s.nx1+c], v);
						if(i==g.rows-1) v += M[(s.nu+s.nx)*s.nx1+c];
						}
					g.p[idx] = v;
					}
				}
			__syncwarp();
			}
		/* ---- Hessian / gradient of the block ---- */
		for(int i=lane; i<HB_TRI(nux2)+nux2; i+=32) H2[i] = 0.0;
		if(sc.ng>0) for(int i=lane; i<nux2*sc.ng; i+=32) inC[sc.off_DCt+i] = 0.0;
		__syncwarp();
		{
		int off = 0;
		for(int sI=T-1; sI>=0; sI--)
			{
			const hb_stage s = dF.st[n0+sI];
			const int nu = s.nu, nx = s.nx, nux = nu+nx, nz = nux+1;
			const double *H = inF + s.off_RSQ;
			if(sI==T-1)
				{ for(int i=lane; i<HB_TRI(nux)+nux; i+=32) pL[i] = H[i]; }
			else
				{
				const hb_stage s1 = dF.st[n0+sI+1];
				const int nus = s1.nu, nxs = s1.nx;
				for(int i=0; i<=nxs; i++)
					for(int c=lane; c<(i<nxs ? i+1 : nxs); c+=32) Lx[HB_TRI(i)+c] = pL[HB_TRI(nus+i)+nus+c];
				__syncwarp();
				hb_chol(cx, Lx, nxs+1, nxs);
				const double *M = inF + s.off_BAbt;
				for(int idx=lane; idx<nz*nxs; idx+=32)
					{
					const int i = idx/nxs, c = idx - i*nxs;
					double v = 0.0;
					for(int m=c; m<nxs; m++) v = fma(M[i*nxs+m], Lx[HB_TRI(m)+c], v);
					if(i==nux) v += Lx[HB_TRI(nxs)+c];
					W[idx] = v;
					}
				__syncwarp();
				for(int i=0; i<nz; i++)
					for(int c=lane; c<(i<nux ? i+1 : nux); c+=32)
						{
						double v = 0.0;
						for(int m=0; m<nxs; m++) v = fma(W[i*nxs+m], W[c*nxs+m], v);
						pL[HB_TRI(i)+c] = H[HB_TRI(i)+c] + v;
						}
				}
			__syncwarp();
			if(sI==0)
				{
				for(int i=0; i<nz; i++)
					for(int c=lane; c<(i<nux ? i+1 : nux); c+=32) H2[HB_TRI(off+i)+off+c] = pL[HB_TRI(i)+c];
				break;
				}
			/* D */
			for(int i=0; i<nu; i++)
				for(int c=lane; c<=i; c+=32) H2[HB_TRI(off+i)+off+c] = pL[HB_TRI(i)+c];
			/* M : Gamma_{s-1} times the state rows of the u-columns (its last row is the b-part of the gradient) */
			const pc_gamma gp = pc_gam(dF, n0, T, sI-1, G, BAbt2);
			for(int idx=lane; idx<gp.rows*nu; idx+=32)
				{
				const int i = idx/nu, c = idx - i*nu;
				double v = 0.0;
				for(int m=0; m<nx; m++) v = fma(gp.p[i*nx+m], pL[HB_TRI(nu+m)+c], v);
				H2[HB_TRI(off+nu+i)+off+c] = v;
				}
			__syncwarp();
			/* m */
			for(int c=lane; c<nu; c+=32) H2[HB_TRI(nux2)+off+c] += pL[HB_TRI(nux)+c];
			__syncwarp();
			off += nu;
			}
		}
		/* ---- constraints: stages T-1 .. 1 (input bounds stay bounds, state bounds become rows of [D C]), then stage 0 ---- */
		{
		int nu_tmp = 0, ib = 0, ig = 0;
		double *d2 = inC + sc.off_d, *dg2 = inC + sc.off_dg, *DCt2 = inC + sc.off_DCt;
		for(int sI=T-1; sI>=1; sI--)
			{
			const hb_stage s = dF.st[n0+sI];
			const pc_gamma gp = pc_gam(dF, n0, T, sI-1, G, BAbt2);
			const int brow = gp.rows-1;
			nu_tmp += s.nu;
			for(int l=0; l<s.nb; l++)
				{
				const int id = dF.idxb[s.off_c+l];
				const double lo = inF[s.off_d+l], up = inF[s.off_d+s.nb+l];
				if(id<s.nu)
					{ if(lane==0) { d2[ib] = lo; d2[sc.nb+ib] = up; } ib++; }
				else
					{
					const double *gcol = gp.p + (id-s.nu);           /* column id-nu of Gamma_{s-1}, stride nx */
					const double gb = gcol[brow*s.nx];
					if(lane==0) { dg2[ig] = lo - gb; dg2[sc.ng+ig] = up - gb; }
					for(int i=lane; i<brow; i+=32) DCt2[(nu_tmp+i)*sc.ng+ig] = gcol[i*s.nx];
					ig++;
					}
				}
			}
		const hb_stage s = dF.st[n0];
		for(int l=lane; l<s.nb; l+=32) { d2[ib+l] = inF[s.off_d+l]; d2[sc.nb+ib+l] = inF[s.off_d+s.nb+l]; }
		}
		/* ---- last stage: shared with the full problem ---- */
		if(k==N2-1)
			{
			const hb_stage s = dF.st[dF.N], s2 = dC.st[N2];
			for(int i=lane; i<HB_TRI(s.nx)+s.nx; i+=32) inC[s2.off_RSQ+i] = inF[s.off_RSQ+i];
			for(int i=lane; i<2*s.nb; i+=32) inC[s2.off_d+i] = inF[s.off_d+i];
			for(int i=lane; i<s.nx*s.ng; i+=32) inC[s2.off_DCt+i] = inF[s.off_DCt+i];
			for(int i=lane; i<2*s.ng; i+=32) inC[s2.off_dg+i] = inF[s.off_dg+i];
			}
		__syncwarp();
		}
	}

__global__ void __launch_bounds__(128) hb_pexpand_kernel(hb_dims dF, hb_dims dC, const hb_pc_block *__restrict__ blk, int N2, long long n_inst,
		const double *__restrict__ in_full, const double *__restrict__ ux2, const double *__restrict__ pi2, const double *__restrict__ lam2,
		const double *__restrict__ t2, double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ lam, double *__restrict__ t)
	{
	const int lane = threadIdx.x & 31, wpb = blockDim.x>>5;
	const long long warp = (long long)blockIdx.x*wpb + (threadIdx.x>>5), nwarps = (long long)gridDim.x*wpb;
	for(long long item=warp; item<n_inst*N2; item+=nwarps)
		{
		const long long inst = item/N2; const int k = (int)(item - inst*N2);
		const double *inF = in_full + inst*dF.in_stride;
		const int n0 = blk[k].n0, T = blk[k].T;
		const hb_stage sc = dC.st[k];
		const double *u2 = ux2 + inst*dC.ux_stride + sc.off_ux;
		const double *l2 = lam2 + inst*2*dC.nbtot + 2*sc.off_c, *tt2 = t2 + inst*2*dC.nbtot + 2*sc.off_c;
		double *uxI = ux + inst*dF.ux_stride, *piI = pi + inst*dF.pi_stride;
		double *lamI = lam + inst*2*dF.nbtot, *tI = t + inst*2*dF.nbtot;
		/* inputs, lam, t by position */
		{
		int nu_tmp = 0, ib = 0, ig = 0;
		for(int j=T-1; j>=0; j--)
			{
			const hb_stage s = dF.st[n0+j];
			const int ncopy = (j==0) ? s.nu+s.nx : s.nu;
			for(int l=lane; l<ncopy; l+=32) uxI[s.off_ux+l] = u2[nu_tmp+l];
			nu_tmp += s.nu;
			int nbb = s.nb;
			if(j>0) { nbb = 0; for(int l=0; l<s.nb; l++) if(dF.idxb[s.off_c+l]<s.nu) nbb++; }
			double *lo = lamI + 2*s.off_c, *to = tI + 2*s.off_c;
			for(int l=lane; l<nbb; l+=32)
				{ lo[l] = l2[ib+l]; lo[s.nb+l] = l2[sc.nb+ib+l]; to[l] = tt2[ib+l]; to[s.nb+l] = tt2[sc.nb+ib+l]; }
			for(int l=nbb+lane; l<s.nb; l+=32)
				{
				const int g = ig + l - nbb;
				lo[l] = l2[2*sc.nb+g]; lo[s.nb+l] = l2[2*sc.nb+sc.ng+g]; to[l] = tt2[2*sc.nb+g]; to[s.nb+l] = tt2[2*sc.nb+sc.ng+g];
				}
			ib += nbb; ig += s.nb-nbb;
			}
		}
		if(k==N2-1)
			{
			const hb_stage s = dF.st[dF.N], s2 = dC.st[N2];
			for(int l=lane; l<s.nx; l+=32) uxI[s.off_ux+l] = ux2[inst*dC.ux_stride + s2.off_ux + l];
			for(int l=lane; l<2*(s.nb+s.ng); l+=32)
				{ lamI[2*s.off_c+l] = lam2[inst*2*dC.nbtot + 2*s2.off_c + l]; tI[2*s.off_c+l] = t2[inst*2*dC.nbtot + 2*s2.off_c + l]; }
			}
		__syncwarp();
		/* states inside the block: x_{j+1} = b_j + [B A]_j [u_j ; x_j] */
		for(int j=0; j<T-1; j++)
			{
			const hb_stage s = dF.st[n0+j], s1 = dF.st[n0+j+1];
			const double *M = inF + s.off_BAbt;
			const int nux = s.nu+s.nx;
			for(int l=lane; l<s.nx1; l+=32)
				{
				double a = M[nux*s.nx1+l];
				for(int i=0; i<nux; i++) a = fma(M[i*s.nx1+l], uxI[s.off_ux+i], a);
				uxI[s1.off_ux+s1.nu+l] = a;
				}
			__syncwarp();
			}
		/* multipliers of the dynamics: the block's last edge is the condensed one, the others follow backwards from the x-rows of
		 * q + (bound multipliers) + [S Q] [u ; x] + A' pi */
		{
		const hb_stage sl = dF.st[n0+T-1];
		for(int l=lane; l<sl.nx1; l+=32) piI[sl.off_pi+l] = pi2[inst*dC.pi_stride + sc.off_pi + l];
		__syncwarp();
		for(int j=T-1; j>=1; j--)
			{
			const hb_stage s = dF.st[n0+j], sm = dF.st[n0+j-1];
			const double *M = inF + s.off_BAbt, *H = inF + s.off_RSQ;
			const int nux = s.nu+s.nx;
			const double *lo = lamI + 2*s.off_c;
			for(int l=lane; l<s.nx; l+=32)
				{
				const int r = s.nu+l;
				double a = H[HB_TRI(nux)+r];
				for(int b=0; b<s.nb; b++) if(dF.idxb[s.off_c+b]==r) a += -lo[b] + lo[s.nb+b];
				double h = 0.0;
				for(int i=0; i<nux; i++) h = fma(i<=r ? H[HB_TRI(r)+i] : H[HB_TRI(i)+r], uxI[s.off_ux+i], h);
				a += h;
				double g = 0.0;
				for(int i=0; i<s.nx1; i++) g = fma(M[r*s.nx1+i], piI[s.off_pi+i], g);
				a += g;
				piI[sm.off_pi+l] = a;
				}
			__syncwarp();
			}
		}
		}
	}

/* one warp per instance */
__global__ void __launch_bounds__(128) hb_res_norms_kernel(hb_dims dF, long long n_inst, const double *__restrict__ rq, const double *__restrict__ rb,
		const double *__restrict__ rd, const double *__restrict__ mu, long long lam_stride, double *__restrict__ info, long long info_stride)
	{
	const int lane = threadIdx.x & 31;
	const long long inst = (long long)blockIdx.x*(blockDim.x>>5) + (threadIdx.x>>5);
	if(inst>=n_inst) return;
	double a = 0.0, b = 0.0, c = 0.0;
	const int n_ux = dF.st[dF.N].off_ux + dF.st[dF.N].nx, n_pi = dF.st[dF.N].off_pi;
	for(int i=lane; i<n_ux; i+=32) a = fmax(a, fabs(rq[inst*dF.ux_stride+i]));
	for(int i=lane; i<n_pi; i+=32) b = fmax(b, fabs(rb[inst*dF.pi_stride+i]));
	for(int i=lane; i<2*dF.nbtot; i+=32) c = fmax(c, fabs(rd[inst*lam_stride+i]));
	for(int o=16; o>0; o>>=1)
		{
		a = fmax(a, __shfl_xor_sync(HB_FULL, a, o)); b = fmax(b, __shfl_xor_sync(HB_FULL, b, o)); c = fmax(c, __shfl_xor_sync(HB_FULL, c, o));
		}
	if(lane==0) { double *o = info + inst*info_stride; o[2] = a; o[3] = b; o[4] = c; o[5] = mu[inst]; }
	}

static int pc_szG(const hb_stage *stF, const hb_pc_block *blk, int N2)
	{
	int best = 2;
	for(int k=0; k<N2; k++)
		{
		int off = 0, nuc = 0;
		const int nx0 = stF[blk[k].n0].nx;
		for(int j=0; j<blk[k].T-1; j++) { nuc += stF[blk[k].n0+j].nu; off += (nuc+nx0+1)*stF[blk[k].n0+j].nx1; }
		if(off>best) best = off;
		}
	return (best+1)&~1;
	}

extern "C" long long hb_pcond_scratch_doubles(const hb_stage *stF, int N, const hb_pc_block *blk, int N2)
	{
	int nzM = 1, nxM = 1;
	for(int n=0; n<=N; n++) { if(stF[n].nu+stF[n].nx+1>nzM) nzM = stF[n].nu+stF[n].nx+1; if(stF[n].nx>nxM) nxM = stF[n].nx; }
	const int ev = 1;
	return (long long)pc_szG(stF, blk, N2) + ((HB_TRI(nzM)+nzM+ev)&~1) + ((HB_TRI(nxM+1)+2*nxM+2+ev)&~1) + (((long long)nzM*nxM+ev)&~1);
	}

/* szG travels in blk[0].off_G (host copy) so the signature stays small */
extern "C" int hb_launch_pcond(const hb_dims *dF, const hb_dims *dC, const hb_pc_block *blk, int N2, long long n_inst, const double *in_full,
		double *in_cond, double *scratch, long long scratch_stride, int grid, int warps, void *stream)
	{
	/* scratch_stride = szG + the fixed parts: recover szG */
	const int nzM = dF->nzM, nxM = dF->nxM;
	const long long fixed = ((HB_TRI(nzM)+nzM+1)&~1) + ((HB_TRI(nxM+1)+2*nxM+2+1)&~1) + (((long long)nzM*nxM+1)&~1);
	const int szG = (int)(scratch_stride - fixed);
	if(szG<0) return -1;
	hb_pcond_kernel<<<grid, warps*32, 0, (cudaStream_t)stream>>>(*dF, *dC, blk, N2, n_inst, in_full, in_cond, scratch, scratch_stride, szG);
	return cudaGetLastError()==cudaSuccess ? 0 : -1;
	}

extern "C" int hb_launch_pexpand(const hb_dims *dF, const hb_dims *dC, const hb_pc_block *blk, int N2, long long n_inst, const double *in_full,
		const double *ux2, const double *pi2, const double *lam2, const double *t2, double *ux, double *pi, double *lam, double *t,
		int grid, int warps, void *stream)
	{
	hb_pexpand_kernel<<<grid, warps*32, 0, (cudaStream_t)stream>>>(*dF, *dC, blk, N2, n_inst, in_full, ux2, pi2, lam2, t2, ux, pi, lam, t);
	return cudaGetLastError()==cudaSuccess ? 0 : -1;
	}

extern "C" int hb_launch_res_norms(const hb_dims *dF, long long n_inst, const double *rq, const double *rb, const double *rd, const double *mu,
		long long lam_stride, double *info, long long info_stride, void *stream)
	{
	const int warps = 4;
	const long long grid = (n_inst+warps-1)/warps;
	hb_res_norms_kernel<<<(unsigned)grid, warps*32, 0, (cudaStream_t)stream>>>(*dF, n_inst, rq, rb, rd, mu, lam_stride, info, info_stride);
	return cudaGetLastError()==cudaSuccess ? 0 : -1;
	}
