/*
 * ric_shared_tpi.cuh -- SHARED DYNAMICS, one THREAD per instance (size-specialised: nu = NU at every stage n < N, nu_N = 0,
 * nx = NX at every stage n >= 1, nx_0 = 0 or NX).
 *
 * With one set of matrices for the whole batch the solve-only recursion (lqcp_solvers/d_back_ric_rec.c:564-789,
 * d_back_ric_rec_trs_tv_res) is a sequence of small matrix-vector products in which every instance multiplies the SAME matrix:
 * the factor and [B A]' of the whole horizon sit in the CTA's shared memory, every lane of a warp reads the same matrix element
 * (a broadcast, one wavefront, no bank conflicts) and multiplies it into ITS instance's vector, which lives in registers --
 * no shuffles, no __syncwarp, 32 independent instances per warp instead of one.  The warp-per-instance kernel
 * (hb_ric_trs_shared_kernel, ric_kernels.cu) keeps every other shape.
 *
 * The order of the operations inside a stage is the one of hb_trs_stage_back / hb_stage_forward (ric_generic.cuh), so the two
 * kernels agree to the last bit wherever the compiler contracts the same multiply-adds.
 *
 * Per instance and stage the thread reads r/q (nu+nx) and b (nx) once, writes w (nu+nx) in the backward sweep, and in the
 * forward sweep reads w back, writes u, x and pi: 8 x (2(nu+nx) + nx + (nu+nx) + (nu+nx) + nx) bytes -- the vectors' own
 * sizes; nothing else crosses HBM.  A thread's accesses to one stage are a contiguous run of (nu+nx) or nx doubles, so the
 * sectors a warp touches are used completely (through L1) although neighbouring lanes are an instance stride apart.
 *
 * Measured alternative (not kept): the vectors staged through a per-warp 32 x 29 shared-memory buffer with warp-coalesced copies,
 * one 16-warp CTA per SM.  Global LSU wavefronts fell from 73 M to 19 M per 65 536 instances, shared-memory wavefronts rose from
 * 44 M to 98 M (row stores + loads, 18 M bank conflicts of the copies), the LSU data pipe stayed at 60-63 % of its peak and the
 * rate at 262 144 instances was the same (83.7 vs 83.8 M solves/s; +10 % at 65 536) -- profiles/r02_ncu_shared_tpi.txt.
 * Second measured alternative (not kept): two instances per lane (each broadcast load serves 64 instances; 255 registers, 8 warps
 * per SM): 85.6 vs 83.8 M solves/s.  Three kernels with different LSU / shared-memory / occupancy profiles tie, so the common
 * factor is what limits them: the vectors' layout in HBM -- per stage a warp touches 32 runs of 96-136 bytes that are an instance
 * stride (7 KB) apart, and DRAM delivers 2.0 TB/s on that pattern.  A stage-major layout of the vectors (one 4 KB run per warp and
 * stage) is the next step; it changes this mode's buffer format.
 */
#pragma once
#include "layout.h"

/* every element of the x-block of a factor is used twice (Lxx' x, then Lxx t); without a fence between the two passes the
 * compiler keeps all of them in registers from the first use to the second (255 registers and 4 KB of spills) */
#define TPI_FENCE() asm volatile("" ::: "memory")

template<int NX, int NU1>
__device__ __forceinline__ void tpi_LLt(const double *__restrict__ Ln1, const double (&x)[NX], double (&y)[NX], const double (&y0)[NX])
	{
	/* y = y0 + Lxx (Lxx' x), Lxx = the x-block of the factor of stage n+1 (rows nu1.., columns nu1..) */
	double t[NX];
	#pragma unroll
	for(int i=0; i<NX; i++)
		{
		double acc = 0.0;
		#pragma unroll
		for(int k=i; k<NX; k++) acc += Ln1[HB_TRI(NU1+k)+NU1+i]*x[k];
		t[i] = acc;
		}
	TPI_FENCE();
	#pragma unroll
	for(int i=0; i<NX; i++)
		{
		double acc = y0[i];
		#pragma unroll
		for(int k=0; k<=i; k++) acc += Ln1[HB_TRI(NU1+i)+NU1+k]*t[k];
		y[i] = acc;
		}
	}

/* one backward stage (n < N).  NUX = nu_n + nx_n, KS = columns eliminated (nu_n, or NUX at n = 0), NU1 = nu_{n+1}.
 * wx: in = x-part of w_{n+1}, out = x-part of w_n (when the stage has one).  rq / b are read before the first fence and w_out is
 * written after the last one, so w_out may be the memory rq came from */
template<int NX, int NU, int NUX, int KS, int NU1>
__device__ __forceinline__ void tpi_back(const double *__restrict__ Ln, const double *__restrict__ Ln1, const double *__restrict__ W,
		const double *rq, const double *b, double *w_out, double (&wx)[NX])
	{
	double bb[NX], zero[NX], Pb[NX], tmp[NX], v[NUX];
	#pragma unroll
	for(int k=0; k<NX; k++) { bb[k] = b[k]; zero[k] = 0.0; }
	tpi_LLt<NX, NU1>(Ln1, bb, Pb, zero);
	TPI_FENCE();
	#pragma unroll
	for(int j=0; j<NX; j++) tmp[j] = Pb[j] + wx[j];
	#pragma unroll
	for(int i=0; i<NUX; i++)
		{
		double acc = rq[i];
		#pragma unroll
		for(int j=0; j<NX; j++) acc += W[i*NX+j]*tmp[j];
		v[i] = acc;
		}
	TPI_FENCE();
	const double *dinv = Ln + HB_TRI(NUX) + NUX;
	#pragma unroll
	for(int j=0; j<KS; j++)
		{
		v[j] *= dinv[j];
		#pragma unroll
		for(int i=j+1; i<NUX; i++) v[i] -= Ln[HB_TRI(i)+j]*v[j];
		}
	#pragma unroll
	for(int i=0; i<NUX; i++) w_out[i] = v[i];
	if constexpr (NUX==NU+NX)
		{
		#pragma unroll
		for(int j=0; j<NX; j++) wx[j] = v[NU+j];
		}
	}

/* one forward stage (n < N).  xs: in = x_n (when the stage has one), out = x_{n+1}.
 * ux_n: w_n in, u_n (and at n = 0 x_0) out; ux_n1x: x-part of w_{n+1} in, x_{n+1} out */
template<int NX, int NU, int NUX, int KS, int NU1>
__device__ __forceinline__ void tpi_fwd(const double *__restrict__ Ln, const double *__restrict__ Ln1, const double *__restrict__ W,
		const double *b, double *ux_n, double *ux_n1x, double *pi_n, double (&xs)[NX])
	{
	double v[NUX], pin[NX], pout[NX];
	#pragma unroll
	for(int i=0; i<KS; i++) v[i] = -ux_n[i];
	#pragma unroll
	for(int i=KS; i<NUX; i++) v[i] = xs[i-KS];
	#pragma unroll
	for(int i=0; i<KS; i++)
		{
		double acc = v[i];
		#pragma unroll
		for(int j=KS; j<NUX; j++) acc -= Ln[HB_TRI(j)+i]*v[j];
		v[i] = acc;
		}
	TPI_FENCE();
	const double *dinv = Ln + HB_TRI(NUX) + NUX;
	#pragma unroll
	for(int j=KS-1; j>=0; j--)
		{
		v[j] *= dinv[j];
		#pragma unroll
		for(int i=0; i<j; i++) v[i] -= Ln[HB_TRI(j)+i]*v[j];
		}
	#pragma unroll
	for(int i=0; i<KS; i++) ux_n[i] = v[i];
	TPI_FENCE();
	/* (two elements of a [B A]' row per 128-bit shared-memory load were measured 7 % SLOWER: the same instruction count after
	 * the extra moves, and a broadcast LDS.128 is two wavefronts) */
	#pragma unroll
	for(int j=0; j<NX; j++)
		{
		double acc = b[j];
		#pragma unroll
		for(int i=0; i<NUX; i++) acc += W[i*NX+j]*v[i];
		pin[j] = ux_n1x[j];
		ux_n1x[j] = acc;
		xs[j] = acc;
		}
	TPI_FENCE();
	tpi_LLt<NX, NU1>(Ln1, xs, pout, pin);
	#pragma unroll
	for(int i=0; i<NX; i++) pi_n[i] = pout[i];
	}

/* shared memory: [factor of the horizon, L_stride doubles | [B A]' of stage 0 .. N-1, rows of exactly nx doubles].
 * SM: STAGE-MAJOR vectors -- the part of a vector that belongs to stage n (offset o_n, K_n doubles per instance) is one array
 * [n_inst][K_n] at o_n * n_inst, so the 32 instances of a warp touch ONE run of 32 K_n doubles per stage instead of 32 runs an
 * instance stride apart (the b part of vec starts at ux_stride * n_inst). */
template<int NX, int NU, bool SM>
__global__ void __launch_bounds__(256, 2) hb_ric_trs_shared_tpi_kernel(hb_dims d, long long n_inst, const double *__restrict__ in_shared,
		const double *__restrict__ L_shared, const double *__restrict__ vec, double *__restrict__ ux_all, double *__restrict__ pi_all,
		unsigned long long *__restrict__ queue, int nx0)
	{
	extern __shared__ double tpi_smem[];
	const int N = d.N, lane = threadIdx.x&31;
	constexpr int NZ = NU+NX;
	const int nux0 = NU + nx0;
	const int szB0 = HB_EVEN(nux0*NX), szB = HB_EVEN(NZ*NX);
	double *Lsm = tpi_smem, *Bsm = Lsm + HB_EVEN(d.L_stride);
	for(long long e=threadIdx.x; e<d.L_stride; e+=blockDim.x) Lsm[e] = L_shared[e];
	for(int n=0; n<N; n++)
		{
		const double *gb = in_shared + d.st[n].off_BAbt;
		double *dst = Bsm + (n==0 ? 0 : szB0 + (n-1)*szB);
		const int cnt = (n==0 ? nux0 : NZ)*NX;
		for(int e=threadIdx.x; e<cnt; e+=blockDim.x) dst[e] = gb[e];
		}
	__syncthreads();
	const long long vs = d.ux_stride + d.pi_stride, us = d.ux_stride, ps = d.pi_stride;
	const hb_stage *__restrict__ st = d.st;
	for(;;)
		{
		unsigned long long blk = 0;
		if(lane==0) blk = atomicAdd(queue, 1ULL);
		blk = __shfl_sync(0xffffffffu, blk, 0);
		if((long long)(blk*32ULL)>=n_inst) break;
		const long long inst = (long long)(blk*32ULL) + lane;
		if(inst>=n_inst) continue;                      /* nothing below talks to the other lanes */
		/* the instance's part of stage s in [r q], b, ux, pi */
		const double *const rq0 = vec + inst*vs, *const bv0 = rq0 + us;          /* instance-major bases */
		double *const ux0 = ux_all + inst*us, *const pi0 = pi_all + inst*ps;
		const double *const bsm = vec + us*n_inst + inst*NX;                      /* stage-major: + off_pi * n_inst */
		double *const psm = pi_all + inst*NX;
		auto RQ = [&](const hb_stage &s) { return SM ? vec + (long long)s.off_ux*n_inst + inst*(s.nu+s.nx) : rq0 + s.off_ux; };
		auto BV = [&](const hb_stage &s) { return SM ? bsm + (long long)s.off_pi*n_inst : bv0 + s.off_pi; };
		auto UX = [&](const hb_stage &s) { return SM ? ux_all + (long long)s.off_ux*n_inst + inst*(s.nu+s.nx) : ux0 + s.off_ux; };
		auto PI = [&](const hb_stage &s) { return SM ? psm + (long long)s.off_pi*n_inst : pi0 + s.off_pi; };
		double wx[NX];
		{
		const hb_stage s = st[N];
		const double *q = RQ(s); double *u = UX(s);
		#pragma unroll
		for(int j=0; j<NX; j++) { wx[j] = q[j]; u[j] = wx[j]; }
		}
		/* backward: stage N-1 looks at a stage N without inputs, stage 0 eliminates everything it has */
		{
		const hb_stage s = st[N-1];
		tpi_back<NX, NU, NZ, NU, 0>(Lsm + s.off_L, Lsm + st[N].off_L, Bsm + szB0 + (N-2)*szB, RQ(s), BV(s), UX(s), wx);
		}
		for(int n=N-2; n>=1; n--)
			{
			const hb_stage s = st[n];
			tpi_back<NX, NU, NZ, NU, NU>(Lsm + s.off_L, Lsm + st[n+1].off_L, Bsm + szB0 + (n-1)*szB, RQ(s), BV(s), UX(s), wx);
			}
		{
		const hb_stage s = st[0];
		if(nx0==0) tpi_back<NX, NU, NU, NU, NU>(Lsm + s.off_L, Lsm + st[1].off_L, Bsm, RQ(s), BV(s), UX(s), wx);
		else       tpi_back<NX, NU, NZ, NZ, NU>(Lsm + s.off_L, Lsm + st[1].off_L, Bsm, RQ(s), BV(s), UX(s), wx);
		}
		/* forward */
		double xs[NX];
		{
		const hb_stage s = st[0], s1 = st[1];
		if(nx0==0) tpi_fwd<NX, NU, NU, NU, NU>(Lsm + s.off_L, Lsm + s1.off_L, Bsm, BV(s), UX(s), UX(s1) + NU, PI(s), xs);
		else       tpi_fwd<NX, NU, NZ, NZ, NU>(Lsm + s.off_L, Lsm + s1.off_L, Bsm, BV(s), UX(s), UX(s1) + NU, PI(s), xs);
		}
		for(int n=1; n<N-1; n++)
			{
			const hb_stage s = st[n], s1 = st[n+1];
			tpi_fwd<NX, NU, NZ, NU, NU>(Lsm + s.off_L, Lsm + s1.off_L, Bsm + szB0 + (n-1)*szB, BV(s), UX(s), UX(s1) + NU, PI(s), xs);
			}
		{
		const hb_stage s = st[N-1], s1 = st[N];
		tpi_fwd<NX, NU, NZ, NU, 0>(Lsm + s.off_L, Lsm + s1.off_L, Bsm + szB0 + (N-2)*szB, BV(s), UX(s), UX(s1), PI(s), xs);
		}
		}
	}

/* which instantiation serves the size pattern (-1: none -- the warp-per-instance kernel does it) */
static int hb_tpi_variant(const hb_dims *d, const hb_stage *st_host, int *nx0)
	{
	const int N = d->N;
	if(N<3 || st_host==NULL) return -1;
	const int NU = st_host[0].nu, NX = st_host[1].nx;
	if(st_host[N].nu!=0 || st_host[N].nx!=NX) return -1;
	if(st_host[0].nx!=0 && st_host[0].nx!=NX) return -1;
	for(int n=0; n<N; n++) if(st_host[n].nu!=NU || (n>0 && st_host[n].nx!=NX)) return -1;
	*nx0 = st_host[0].nx;
	if(NX==12 && NU==5) return 0;
	if(NX==8 && NU==3) return 1;
	return -1;
	}
static long long hb_tpi_smem_bytes(const hb_dims *d, const hb_stage *st_host)
	{
	const int N = d->N, NU = st_host[0].nu, NX = st_host[1].nx, nux0 = NU + st_host[0].nx;
	return 8LL*(HB_EVEN(d->L_stride) + HB_EVEN(nux0*NX) + (long long)(N-1)*HB_EVEN((NU+NX)*NX));
	}

