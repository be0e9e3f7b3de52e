/*
 * ric_kernels.cu -- sm_100a kernels of the batched Riccati / box-IPM engine and their C launchers.
 *
 * Kernels (one warp per OCP instance, persistent grid, per-warp scratch "slot" for the factor stash):
 *   hb_ric_sv_kernel   factor + solve            <- d_back_ric_rec_sv_tv_res   (lqcp_solvers/d_back_ric_rec.c:112)
 *   hb_ric_trf_kernel  factor, L kept per inst.  <- d_back_ric_rec_trf_tv_res  (:403)
 *   hb_ric_trs_kernel  solve with stored L       <- d_back_ric_rec_trs_tv_res  (:564)
 *   hb_ipm_kernel      whole two-phase Mehrotra IPM on device, no host round trip per iteration
 *                                                <- d_ip2_res_mpc_hard_tv      (mpc_solvers/d_ip2_res_hard.c:116)
 *                      element-wise steps        <- mpc_solvers/c99/d_aux_ip_hard_lib4.c (lines cited inline)
 *                      residuals                 <- mpc_solvers/c99/d_res_ip_res_hard.c:39
 *   hb_fp64_probe      DFMA peak microbenchmark (roofline denominator, not part of the solver)
 */
#include "launch_util.cuh"
#include "layout.h"
#include "ric_sweeps.cuh"
#include "ric_team.cuh"
#include "ric_tree.cuh"
#include "ric_shared_tpi.cuh"

/* ------------------------------------------------------------------------------------------------ */
/* Riccati kernels                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
extern __shared__ double hb_smem[];
/* Qx, qx (may be NULL): the IPM's per-constraint updates of the Hessian diagonal / gradient row, nbtot doubles per instance in
 * the flat constraint order (stage after stage: box entries, then general ones) -- the reference's Qx / qx arguments
 * (lqcp_solvers/d_back_ric_rec.c:112); with general constraints they weight [D C]' (:293-315) */
__global__ void hb_ric_sv_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ Pb, double *__restrict__ stash,
		const double *__restrict__ Qx, const double *__restrict__ qx)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	double *Lst = stash + gw*d.L_stride;
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *in_inst = in + inst*d.in_stride;
		hb_backward<true>(c, d, in_inst, Lst, nullptr, nullptr, Qx!=nullptr ? Qx + inst*d.nbtot : nullptr,
				qx!=nullptr ? qx + inst*d.nbtot : nullptr, Pb!=nullptr ? Pb + inst*d.pi_stride : nullptr);
		__syncwarp();
		hb_forward(c, d, in_inst, Lst, nullptr, nullptr, false, ux + inst*d.ux_stride, pi + inst*d.pi_stride, true);
		__syncwarp();
		}
	}

__global__ void hb_ric_trf_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, double *__restrict__ L,
		const double *__restrict__ Qx)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		hb_backward<false>(c, d, in + inst*d.in_stride, L + inst*d.L_stride, nullptr, nullptr, Qx!=nullptr ? Qx + inst*d.nbtot : nullptr, nullptr, nullptr);
		__syncwarp();
		}
	}

/* the same two kernels with four warps per instance (ric_team.cuh): one CTA of 128 threads works on one instance at a time, the
 * factor stash is one slot per CTA; the forward sweep is warp 0's */
__global__ void __launch_bounds__(HBT_THREADS, 5) hbt_ric_sv_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ Pb, double *__restrict__ stash,
		const double *__restrict__ Qx, const double *__restrict__ qx)
	{
	const int tid = threadIdx.x;
	hb_ctx c = hb_make_ctx(d, hb_smem, tid&31);
	double *P = hb_smem + hb_smem_doubles_per_warp(d.nzM, d.nxM);
	double *Lst = stash + (long long)blockIdx.x*d.L_stride;
	for(long long inst=blockIdx.x; inst<n_inst; inst+=gridDim.x)
		{
		const double *in_inst = in + inst*d.in_stride;
		hbt_backward<true>(c, tid, d, in_inst, Lst, nullptr, nullptr, Qx!=nullptr ? Qx + inst*d.nbtot : nullptr,
				qx!=nullptr ? qx + inst*d.nbtot : nullptr, Pb!=nullptr ? Pb + inst*d.pi_stride : nullptr);
		hbt_forward(c, tid, P, d, in_inst, Lst, nullptr, nullptr, false, ux + inst*d.ux_stride, pi + inst*d.pi_stride, true);
		hbt_sync();
		}
	}

/* solve with the stored factor by the team (hb_ric_trs_kernel) */
__global__ void __launch_bounds__(HBT_THREADS) hbt_ric_trs_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, const double *__restrict__ L,
		double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ work, const double *__restrict__ qx)
	{
	const int tid = threadIdx.x;
	hb_ctx c = hb_make_ctx(d, hb_smem, tid&31);
	double *P = hb_smem + hb_smem_doubles_per_warp(d.nzM, d.nxM);
	double *rq = work + (long long)blockIdx.x*(d.ux_stride + 2*d.pi_stride), *bv = rq + d.ux_stride, *Pb = bv + d.pi_stride;
	for(long long inst=blockIdx.x; inst<n_inst; inst+=gridDim.x)
		{
		const double *in_inst = in + inst*d.in_stride;
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			const int nux = s.nu+s.nx;
			for(int i=tid; i<nux; i+=HBT_THREADS) rq[s.off_ux+i] = in_inst[s.off_RSQ+HB_TRI(nux)+i];
			for(int j=tid; j<s.nx1; j+=HBT_THREADS) bv[s.off_pi+j] = in_inst[s.off_BAbt+nux*s.nx1+j];
			}
		hbt_sync();
		double *uxi = ux + inst*d.ux_stride;
		hbt_trs_backward(c, tid, P, d, in_inst, L + inst*d.L_stride, bv, rq, qx!=nullptr ? qx + inst*d.nbtot : nullptr, uxi, Pb, true);
		hbt_forward(c, tid, P, d, in_inst, L + inst*d.L_stride, uxi, bv, true, uxi, pi + inst*d.pi_stride, true);
		hbt_sync();
		}
	}

__global__ void __launch_bounds__(HBT_THREADS, 5) hbt_ric_trf_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, double *__restrict__ L,
		const double *__restrict__ Qx)
	{
	const int tid = threadIdx.x;
	hb_ctx c = hb_make_ctx(d, hb_smem, tid&31);
	for(long long inst=blockIdx.x; inst<n_inst; inst+=gridDim.x)
		hbt_backward<false>(c, tid, d, in + inst*d.in_stride, L + inst*d.L_stride, nullptr, nullptr, Qx!=nullptr ? Qx + inst*d.nbtot : nullptr, nullptr, nullptr);
	}

/* solve with the stored factor; b and [r q] are taken from the instance block (new right-hand sides are
 * supplied by packing them into a copy of the block) */
__global__ void hb_ric_trs_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, const double *__restrict__ L,
		double *__restrict__ ux, double *__restrict__ pi, double *__restrict__ work, const double *__restrict__ qx)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	/* per-slot work: rq (ux layout), b (pi layout), Pb (pi layout) */
	double *rq = work + gw*(d.ux_stride + 2*d.pi_stride), *bv = rq + d.ux_stride, *Pb = bv + d.pi_stride;
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *in_inst = in + inst*d.in_stride;
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			const int nux = s.nu+s.nx;
			for(int i=lane; i<nux; i+=32) rq[s.off_ux+i] = in_inst[s.off_RSQ+HB_TRI(nux)+i];
			for(int j=lane; j<s.nx1; j+=32) bv[s.off_pi+j] = in_inst[s.off_BAbt+nux*s.nx1+j];
			}
		__syncwarp();
		double *uxi = ux + inst*d.ux_stride;
		hb_trs_backward(c, d, in_inst, L + inst*d.L_stride, bv, rq, qx!=nullptr ? qx + inst*d.nbtot : nullptr, uxi, Pb, true);
		hb_forward(c, d, in_inst, L + inst*d.L_stride, uxi, bv, true, uxi, pi + inst*d.pi_stride, true);
		__syncwarp();
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* SHARED DYNAMICS: every instance of the batch has the same matrices (A, B, Q, S, R at every stage) and its own vectors (b, q, r) --
 * a fleet of identical systems, a parameter sweep over initial states / references; also how the reference's own test programs
 * call the solver (one pBAbt pointer aliased over all stages, test_problems/test_d_ip_hard.c).  The factorisation is then the
 * same for all of them: it is done once (hb_ric_trf_kernel on the one shared block), and this kernel is the batched
 * d_back_ric_rec_trs_tv_res (lqcp_solvers/d_back_ric_rec.c:564) on that one factor.  The factor and [B A]' of the WHOLE horizon
 * are loaded into the CTA's shared memory once; after that an instance costs its vectors in and its solution out, nothing else
 * crosses HBM.  vec: per instance [r q] of every stage (ux layout, ux_stride doubles) followed by b of every stage (pi layout).
 * ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) hb_ric_trs_shared_kernel(hb_dims d, long long n_inst, const double *__restrict__ in_shared,
		const double *__restrict__ L_shared, const double *__restrict__ vec, double *__restrict__ ux_all, double *__restrict__ pi_all,
		double *__restrict__ work, int resident)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	const int ldW = d.nxM | 1, N = d.N;
	/* shared memory: [per-warp scratch sV 192 | per-warp stage buffers when not resident] ... then the resident horizon */
	hb_ctx c;
	c.lane = lane; c.ldW = ldW;
	double *base = hb_smem;
	c.sV = base + (size_t)warp*192;
	double *Lsm = base + (size_t)nw*192;
	double *Bsm = Lsm + d.L_stride;
	int *offB = reinterpret_cast<int*>(Bsm + 0);            /* set below: table of the stages' offsets inside Bsm */
	if(resident)
		{
		/* table first (N+1 ints, padded to doubles), then the matrices */
		double *tab_end = Bsm + HB_EVEN((N+2+1)/2);
		offB = reinterpret_cast<int*>(Bsm);
		Bsm = tab_end;
		if(threadIdx.x==0)
			{
			int o = 0;
			for(int n=0; n<N; n++) { offB[n] = o; o += HB_EVEN((d.st[n].nu + d.st[n].nx + 1)*ldW); }
			offB[N] = o;
			}
		for(long long e=threadIdx.x; e<d.L_stride; e+=blockDim.x) Lsm[e] = L_shared[e];
		__syncthreads();
		for(int n=0; n<N; n++)
			{
			const hb_stage s = d.st[n];
			const int nux = s.nu+s.nx, nx1 = s.nx1;
			const double *gb = in_shared + s.off_BAbt;
			double *dst = Bsm + offB[n];
			for(int e=threadIdx.x; e<(nux+1)*nx1; e+=blockDim.x) { int i = e/nx1, j = e - i*nx1; dst[i*ldW+j] = gb[e]; }
			}
		__syncthreads();
		}
	else
		{
		/* the horizon does not fit: stage matrices are fetched per stage into per-warp buffers (they still come from L2) */
		const int lsz = HB_EVEN(HB_TRI(d.nzM) + 2*d.nzM);
		double *wb = base + (size_t)nw*192 + (size_t)warp*(2*lsz + HB_EVEN(d.nzM*ldW));
		c.bufA = wb; c.bufB = wb + lsz; c.sW = wb + 2*lsz;
		}
	double *Pb = work + gw*d.pi_stride;
	const long long vs = d.ux_stride + d.pi_stride;
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *rq = vec + inst*vs, *bv = rq + d.ux_stride;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		if(!resident)
			{
			hb_trs_backward(c, d, in_shared, L_shared, bv, rq, nullptr, ux, Pb, true);
			hb_forward(c, d, in_shared, L_shared, ux, bv, true, ux, pi, true);
			__syncwarp();
			continue;
			}
		{
		const hb_stage s = d.st[N];
		for(int i=lane; i<s.nu+s.nx; i+=32) ux[s.off_ux+i] = rq[s.off_ux+i];
		__syncwarp();
		}
		for(int n=N-1; n>=0; n--)
			{
			const hb_stage s = d.st[n], s1 = d.st[n+1];
			c.sW = Bsm + offB[n];
			hb_trs_stage_back(c, s, s1, n, Lsm + s.off_L, Lsm + s1.off_L, bv, rq, nullptr, d.idxb, ux, Pb, true);
			}
		for(int n=0; n<N; n++)
			{
			const hb_stage s = d.st[n], s1 = d.st[n+1];
			c.sW = Bsm + offB[n];
			hb_stage_forward(c, s, s1, n, Lsm + s.off_L, Lsm + s1.off_L, ux, bv, true, ux, pi, true);
			}
		__syncwarp();
		}
	}

extern "C" long long hb_trs_shared_smem_bytes(const hb_dims *d, const hb_stage *st_host, int warps, int *resident)
	{
	/* size patterns with a thread-per-instance kernel (ric_shared_tpi.cuh): *resident = 16 + 2 variant + (nx_0 != 0);
	 * HPMPC_B200_SHARED_GENERIC=1 keeps the warp-per-instance kernel */
	{
	int nx0 = 0;
	const int var = getenv("HPMPC_B200_SHARED_GENERIC")!=NULL ? -1 : hb_tpi_variant(d, st_host, &nx0);
	if(var>=0 && hb_tpi_smem_bytes(d, st_host)<=113*1024) { *resident = 16 + 2*var + (nx0!=0); return hb_tpi_smem_bytes(d, st_host); }
	}
	const int ldW = d->nxM | 1;
	long long B = HB_EVEN((d->N+2+1)/2);
	for(int n=0; n<d->N; n++) B += HB_EVEN((st_host[n].nu + st_host[n].nx + 1)*ldW);
	long long res = 8LL*((long long)warps*192 + d->L_stride + B);
	if(res<=220*1024) { *resident = 1; return res; }
	const int lsz = HB_EVEN(HB_TRI(d->nzM) + 2*d->nzM);
	*resident = 0;
	return 8LL*((long long)warps*192 + (long long)warps*(2*lsz + HB_EVEN(d->nzM*ldW)));
	}

extern "C" int hb_launch_ric_trs_shared(const hb_dims *d, long long n_inst, const double *in_shared, const double *L_shared, const double *vec,
		double *ux, double *pi, double *work, int grid, int warps, int smem, int resident, void *stream)
	{
	if(d->nzM>64) return -2;
	if(warps>8) return -3;
	if(resident>=16)
		{
		/* one thread per instance, warps take blocks of 32 instances from a queue (its counter: the first 8 bytes of `work`) */
		/* + 64: stage-major vectors (hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major) */
		const int stage_major = resident>=64;
		if(stage_major) resident -= 64;
		const int var = (resident-16)>>1, nx0 = ((resident-16)&1) ? d->nxM : 0;
		/* one CTA per SM slot as soon as there are that many blocks of 32 instances: the queue then spreads the blocks evenly */
		const long long need = (n_inst + 31)/32;
		if(need<grid) grid = (int)need;
		unsigned long long *queue = reinterpret_cast<unsigned long long*>(work);
		HB_CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), (cudaStream_t)stream));
#define HB_TPI_LAUNCH(NX_, NU_, SM_) do { if(hb_prep(hb_ric_trs_shared_tpi_kernel<NX_, NU_, SM_>, smem)) return -1; \
		hb_ric_trs_shared_tpi_kernel<NX_, NU_, SM_><<<grid, 256, smem, (cudaStream_t)stream>>>(*d, n_inst, in_shared, L_shared, vec, ux, pi, queue, nx0); } while(0)
		if(var==0) { if(stage_major) HB_TPI_LAUNCH(12, 5, true); else HB_TPI_LAUNCH(12, 5, false); }
		else       { if(stage_major) HB_TPI_LAUNCH(8, 3, true);  else HB_TPI_LAUNCH(8, 3, false); }
#undef HB_TPI_LAUNCH
		HB_CK(cudaGetLastError());
		return 0;
		}
	if(hb_prep(hb_ric_trs_shared_kernel, smem)) return -1;
	hb_ric_trs_shared_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in_shared, L_shared, vec, ux, pi, work, resident);
	HB_CK(cudaGetLastError());
	return 0;
	}

/* ------------------------------------------------------------------------------------------------ */
/* scenario tree                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
__global__ void hb_tree_kernel(hb_tdims d, long long n_trees, const double *__restrict__ in, double *__restrict__ ux_all,
		double *__restrict__ pi_all, double *__restrict__ L_all, int mode, int seg_lo, int seg_hi, const double *__restrict__ skip)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c;
	{
	double *smem_warp = hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM);
	int lsz = HB_EVEN(HB_TRI(d.nzM) + 2*d.nzM);
	c.lane = lane; c.ldW = HB_LDW(d.nxM);
	c.bufA = smem_warp; c.bufB = c.bufA + lsz; c.sW = c.bufB + lsz; c.sV = c.sW + HB_EVEN(d.nzM*c.ldW);
	}
	if(skip!=nullptr && skip[n_trees*8]==0.0) return;          /* tree IPM driver: every tree has finished */
	const int nseg = seg_hi - seg_lo;
	const long long n_items = n_trees*nseg;
	for(long long item=gw; item<n_items; item+=tw)
		{
		const long long t = item/nseg;
		if(skip!=nullptr && (int)skip[t*8+3]==5 /* TS_DONE: a finished tree of the IPM driver */) continue;
		const int seg = seg_lo + (int)(item - t*nseg);
		const double *in_tree = in + t*d.in_stride;
		double *Lt = L_all + t*d.L_stride, *ux = ux_all + t*d.ux_stride, *pi = pi_all + t*d.pi_stride;
		const int s0 = d.seg_start[seg], s1 = d.seg_start[seg+1];
		if(mode==0 || mode==2)
			for(int q=s1-1; q>=s0; q--)
				{
				hb_tree_node_factor(c, d.tn, d.seg_nodes[q], in_tree, Lt, c.bufA, c.bufB);
				__syncwarp();
				}
		if(mode==1 || mode==2)
			for(int q=s0; q<s1; q++)
				{
				hb_tree_node_forward(c, d.tn, d.seg_nodes[q], in_tree, Lt, ux, pi, c.bufA, c.bufB);
				__syncwarp();
				}
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* FP64 peak probe                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
__global__ void hb_fp64_probe(double *out, int iters)
	{
	double a0 = threadIdx.x*1e-9, a1 = a0+1, a2 = a0+2, a3 = a0+3, a4 = a0+4, a5 = a0+5, a6 = a0+6, a7 = a0+7;
	const double x = 1.0000001, y = 1e-9;
	for(int i=0; i<iters; i++)
		{
		a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
		a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
		}
	out[blockIdx.x*blockDim.x+threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7;
	}

/* ------------------------------------------------------------------------------------------------ */
/* launchers (C ABI)                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

extern "C" int hb_smem_bytes_per_warp(const hb_dims *d)
	{
	return (int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	}

extern "C" int hb_smem_bytes_per_warp_sz(int nzM, int nxM) { return (int)sizeof(double)*hb_smem_doubles_per_warp(nzM, nxM); }


extern "C" int hb_device_sm_count(int device)
	{
	int n = 0;
	if(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device)!=cudaSuccess) return -1;
	return n;
	}

#ifdef HBT_TIMING
/* phase clocks of block 0 (lane 0 of each of the four warps), then reset */
extern "C" int hbt_timing_read(long long *out48)
	{
	HB_CK(cudaDeviceSynchronize());
	HB_CK(cudaMemcpyFromSymbol(out48, hbt_tm_g, sizeof(long long)*48));
	long long z[48]; memset(z, 0, sizeof(z));
	HB_CK(cudaMemcpyToSymbol(hbt_tm_g, z, sizeof(z)));
	return 0;
	}
#endif

extern "C" int hbt_smem_bytes(const hb_dims *d) { return (int)sizeof(double)*hbt_smem_doubles(d->nzM, d->nxM); }

/* launch shape of the four-warps-per-instance kernels: CTAs of one instance, as many per SM as the shared memory holds.
 * HPMPC_B200_TEAM=0 keeps the one-warp-per-instance kernels (A/B runs) */
/* which kernel set serves an any-size pattern: the team pays off when a stage has more rows than one warp's 16 row slots; small
 * stages (nu+nx+1 < 20) keep one warp per instance, of which an SM then holds many (measured: nx=10 nu=4 N=25 4.7 M solves/s on one
 * warp against 3.0 M on the team; nx=12 nu=10 ng=6 (23 rows) 104 K against 119 K IPM solves/s).  HPMPC_B200_TEAM=0 / =1 force one */
extern "C" int hbt_wanted(const hb_dims *d)
	{
	const char *e = getenv("HPMPC_B200_TEAM");
	if(e && e[0]=='0') return 0;
	if(e && e[0]=='1') return 1;
	return d->nzM>=20;
	}

static int hbt_grid(const hb_dims *d, long long n_inst, int max_ctas)
	{
	if(!hbt_wanted(d)) return 0;
	int dev = 0, sms = 0;
	if(cudaGetDevice(&dev)!=cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)!=cudaSuccess) return 0;
	int per_sm = 233472/(hbt_smem_bytes(d) + 1024);
	if(per_sm>16) per_sm = 16;
	if(per_sm<1) return 0;
	long long g = (long long)sms*per_sm;
	if(g>n_inst) g = n_inst;
	if(max_ctas>0 && g>max_ctas) g = max_ctas;
	return (int)(g<1 ? 1 : g);
	}

extern "C" int hb_launch_ric_sv(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi, double *Pb,
		double *stash, int n_slots, int grid, int warps, void *stream, const double *Qx, const double *qx)
	{
	if(d->nzM>64) { fprintf(stderr, "hpmpc_b200: nu+nx+1 > 64 not supported\n"); return -2; }
	if(grid*warps>n_slots) return -3;
	{
	const int tg = hbt_grid(d, n_inst, n_slots);
	if(tg>0)
		{
		const int tsmem = hbt_smem_bytes(d);
		if(hb_prep(hbt_ric_sv_kernel, tsmem)) return -1;
		hbt_ric_sv_kernel<<<tg, HBT_THREADS, tsmem, (cudaStream_t)stream>>>(*d, n_inst, in, ux, pi, Pb, stash, Qx, qx);
		HB_CK(cudaGetLastError());
		return 0;
		}
	}
	int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_ric_sv_kernel, smem)) return -1;
	hb_ric_sv_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, ux, pi, Pb, stash, Qx, qx);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_ric_trf(const hb_dims *d, long long n_inst, const double *in, double *L, int grid, int warps, void *stream,
		const double *Qx)
	{
	if(d->nzM>64) return -2;
	{
	const int tg = hbt_grid(d, n_inst, 0);
	if(tg>0)
		{
		const int tsmem = hbt_smem_bytes(d);
		if(hb_prep(hbt_ric_trf_kernel, tsmem)) return -1;
		hbt_ric_trf_kernel<<<tg, HBT_THREADS, tsmem, (cudaStream_t)stream>>>(*d, n_inst, in, L, Qx);
		HB_CK(cudaGetLastError());
		return 0;
		}
	}
	int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_ric_trf_kernel, smem)) return -1;
	hb_ric_trf_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, L, Qx);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_ric_trs(const hb_dims *d, long long n_inst, const double *in, const double *L, double *ux, double *pi,
		double *work, int n_slots, int grid, int warps, void *stream, const double *qx)
	{
	if(d->nzM>64) return -2;
	if(grid*warps>n_slots) return -3;
	{
	const int tg = hbt_grid(d, n_inst, n_slots);
	if(tg>0)
		{
		const int tsmem = hbt_smem_bytes(d);
		if(hb_prep(hbt_ric_trs_kernel, tsmem)) return -1;
		hbt_ric_trs_kernel<<<tg, HBT_THREADS, tsmem, (cudaStream_t)stream>>>(*d, n_inst, in, L, ux, pi, work, qx);
		HB_CK(cudaGetLastError());
		return 0;
		}
	}
	int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_ric_trs_kernel, smem)) return -1;
	hb_ric_trs_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, L, ux, pi, work, qx);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_tree(const hb_tdims *d, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int grid, int warps, void *stream, const double *skip)
	{
	if(d->nzM>64) { fprintf(stderr, "hpmpc_b200: tree: nu+nx+1 > 64 not supported\n"); return -2; }
	if(seg_hi<=seg_lo || n_trees<=0) return 0;
	int smem = warps*(int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	if(hb_prep(hb_tree_kernel, smem)) return -1;
	hb_tree_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, skip);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" double hb_fp64_peak_probe(int device, int iters, void *stream)
	{
	int sms = hb_device_sm_count(device);
	if(sms<=0) return -1.0;
	const int threads = 512, blocks = sms*4;
	double *out = nullptr;
	if(cudaMalloc(&out, sizeof(double)*threads*blocks)!=cudaSuccess) return -1.0;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaStream_t st = (cudaStream_t)stream;
	hb_fp64_probe<<<blocks, threads, 0, st>>>(out, iters);
	cudaEventRecord(e0, st);
	hb_fp64_probe<<<blocks, threads, 0, st>>>(out, iters);
	cudaEventRecord(e1, st);
	cudaEventSynchronize(e1);
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(out);
	double flops = 2.0*8.0*(double)iters*threads*blocks;
	return flops/(ms*1e-3)/1e12;
	}

