/*
 * ric_fast.cuh -- size-specialised Riccati factor+solve for sm_100a: G lanes per OCP instance (32/G instances per warp),
 * compile-time (NX, NU), H / W tiles in registers, stage inputs prefetched by 1-D bulk async copies
 * (cp.async.bulk -> UBLKCP) into shared memory behind an mbarrier, factor stash written with bulk stores.
 *
 * Mapping (frame = NZ x NUX trapezoid, NZ = NU+NX+1):
 *   rows r < RO = min(NZ,G)         : "row-owned": lane r keeps row r of H / L and of W in registers
 *   rows RO..NZ-1 (E "extra" rows)  : "column-owned": entry (row, c) lives on lane c for c < CO = min(NUX,G)
 *   columns CO..NUX-1 (NCC corner)  : the small E x NCC corner block is replicated on every lane
 * so nx=12, nu=5 (18 x 17) runs on 16 lanes with all of them busy, two instances per warp.
 * The second operand of every product is a shared-memory broadcast (one wavefront per LDS.128 for all instances of
 * the warp), which is what keeps the kernel on the FP64 pipe instead of the LDS pipe.
 *
 * Stages whose u- or x-block is absent (stage N: nu = 0; stage 0 after x0 elimination: nx = 0) are embedded in the
 * same frame with an identity diagonal and zero coupling, which leaves every quantity the recursion uses unchanged.
 *
 * Restates (reference paths relative to /root/reference):
 *   backward stage   lqcp_solvers/d_back_ric_rec.c:236-333  (dtrmm_nt_u, gradient-row add, dsyrk_dpotrf)
 *   pivot rule       kernel/c99/kernel_dpotrf_c99_lib4.c:553-573
 *   forward stage    lqcp_solvers/d_back_ric_rec.c:341-397  (dtrsv_t, dgemv_t, dtrmv_u_n / dtrmv_u_t for pi)
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "layout.h"

#define HBF_FULL 0xffffffffu

template<int NX_, int NU_, int G_>
struct hbf_cfg
	{
	static constexpr int NX = NX_, NU = NU_, G = G_;
	static constexpr int NUX = NX+NU, NZ = NUX+1;
	static constexpr int RO = NZ<G ? NZ : G;
	static constexpr int E = NZ-RO;
	static constexpr int CO = NUX<G ? NUX : G;
	static constexpr int NCC = NUX-CO;
	static constexpr int IPW = 32/G;
	__host__ __device__ static constexpr int even(int x) { return (x+1)&~1; }
	/* packed column buffer of L: column c holds rows c..NZ-1 and starts on an even offset */
	__host__ __device__ static constexpr int colOff(int c) { int o = 0; for(int j=0; j<c; j++) o += even(NZ-j); return o; }
	static constexpr int DINV = colOff(NUX);
	static constexpr int LBUF = colOff(NUX) + even(NUX);
	static constexpr int LDW = ((even(NX)/2)%2==0) ? even(NX)+2 : even(NX);
	static constexpr int INB = even(NZ*NX) + even(HB_TRI(NUX)+NUX);          /* one stage of inputs: [B A b]' then RSQrq */
	static constexpr int VEC = 4*even(NZ);
	static constexpr int PER_INST = 2*INB + 3*LBUF + NZ*LDW + VEC;           /* doubles of smem per instance */
	static constexpr int PER_WARP = IPW*PER_INST + 8;                        /* + 4 mbarriers (8 doubles) */
	};

/* ---- PTX helpers: mbarrier + bulk async copy (TMA 1-D) ---- */
__device__ __forceinline__ uint32_t hbf_saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hbf_mbar_init(uint64_t *bar, int count)
	{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(hbf_saddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void hbf_mbar_expect(uint64_t *bar, uint32_t bytes)
	{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(hbf_saddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void hbf_mbar_wait(uint64_t *bar, uint32_t parity)
	{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}" :: "r"(hbf_saddr(bar)), "r"(parity) : "memory");
	}
__device__ __forceinline__ void hbf_bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
	{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(hbf_saddr(sdst)), "l"(gsrc), "r"(bytes), "r"(hbf_saddr(bar)) : "memory");
	}
__device__ __forceinline__ void hbf_bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes)
	{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(hbf_saddr(ssrc)), "r"(bytes) : "memory");
	}
__device__ __forceinline__ void hbf_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void hbf_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template<int N> __device__ __forceinline__ void hbf_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void hbf_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* per-instance view of the stage sizes inside the frame */
struct hbf_stage_view
	{
	int hu, hx;          /* u block / x block present */
	int nu_n, nux_n;     /* actual sizes */
	};

template<class C>
__device__ __forceinline__ int hbf_arow(const hbf_stage_view &v, int f)
	{
	/* frame row f -> actual row of the stage's matrices, or -1 when the row is a phantom */
	if(f<C::NU) return v.hu ? f : -1;
	if(f<C::NUX) return v.hx ? v.nu_n + f - C::NU : -1;
	return v.nux_n;
	}

/* ------------------------------------------------------------------------------------------------ */
/* one backward stage for the G-lane group: Lc <- chol_mn( RSQrq_n + W W' ), W = [B A b]'_n Lxx_{n+1} */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ __forceinline__ void hbf_stage_backward(int l, const hbf_stage_view v, bool last, const double *__restrict__ sB,
		const double *__restrict__ sQ, const double *__restrict__ Lp, double *__restrict__ Lc, double *__restrict__ sW)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, NZ = C::NZ, RO = C::RO, E = C::E, CO = C::CO, NCC = C::NCC, G = C::G, LDW = C::LDW;
	double Hrow[CO];
	double Hext[E>0 ? E : 1];
	double Hcor[E>0 ? E : 1][NCC>0 ? NCC : 1];
	const int ar_own = hbf_arow<C>(v, l);          /* actual row of this lane's frame row (l < RO) */
	int ar_ext[E>0 ? E : 1];
	#pragma unroll
	for(int e=0; e<E; e++) ar_ext[e] = hbf_arow<C>(v, RO+e);
	const int ac_own = (l<NUX) ? hbf_arow<C>(v, l) : -1;   /* actual column index of frame column l */

	/* ---- H <- RSQrq (identity on phantom diagonals) ---- */
	#pragma unroll
	for(int k=0; k<CO; k++)
		{
		const int ak = hbf_arow<C>(v, k);
		double h = (k==l) ? 1.0 : 0.0;
		if(l<RO && k<=l && ar_own>=0 && ak>=0) h = sQ[HB_TRI(ar_own)+ak];
		Hrow[k] = h;
		}
	#pragma unroll
	for(int e=0; e<E; e++)
		{
		double h = 0.0;
		if(l<CO && ar_ext[e]>=0 && ac_own>=0) h = sQ[HB_TRI(ar_ext[e])+ac_own];
		Hext[e] = h;
		#pragma unroll
		for(int cc=0; cc<NCC; cc++)
			{
			/* corner column CO+cc is the diagonal column of extra row RO+cc */
			const int acol = hbf_arow<C>(v, CO+cc);
			double hc = (e==cc) ? 1.0 : 0.0;
			if(e>=cc && ar_ext[e]>=0 && acol>=0) hc = sQ[HB_TRI(ar_ext[e])+acol];
			Hcor[e][cc] = hc;
			}
		}

	if(!last)
		{
		/* ---- W = [B A b]' Lxx' : own row ---- */
		double w[NX];
		{
		double a[NX];
		#pragma unroll
		for(int k=0; k<NX; k+=2)
			{
			double2 t = make_double2(0.0, 0.0);
			if(l<RO && ar_own>=0) t = *reinterpret_cast<const double2*>(sB + ar_own*NX + k);
			a[k] = t.x; if(k+1<NX) a[k+1] = t.y;
			}
		#pragma unroll
		for(int j=0; j<NX; j++)
			{
			double acc = 0.0;
			const double *col = Lp + C::colOff(NU+j);          /* rows NU+j.. of column NU+j : (k-j) offsets */
			#pragma unroll
			for(int k=j; k<NX; k+=2)
				{
				double2 t = *reinterpret_cast<const double2*>(col + (k-j));
				acc = fma(a[k], t.x, acc);
				if(k+1<NX) acc = fma(a[k+1], t.y, acc);
				}
			w[j] = acc;
			}
		if(E==0 && l==NUX)          /* gradient row is row-owned: add l_x' */
			{
			#pragma unroll
			for(int j=0; j<NX; j++) w[j] += Lp[C::colOff(NU+j) + (NZ-1-(NU+j))];
			}
		}
		/* ---- extra rows: lane j owns column j of them ---- */
		double wext[E>0 ? E : 1];
		if(E>0)
			{
			#pragma unroll
			for(int e=0; e<E; e++) wext[e] = 0.0;
			const int j = l;
			int coff = 0;
			/* colOff(NU+j) for the run-time j: small table in registers via unrolled select */
			#pragma unroll
			for(int jj=0; jj<NX; jj++) if(jj==j) coff = C::colOff(NU+jj);
			#pragma unroll
			for(int k=0; k<NX; k++)
				{
				double lkj = 0.0;
				if(j<NX && k>=j) lkj = Lp[coff + (k-j)];
				#pragma unroll
				for(int e=0; e<E; e++)
					{
					double b = (ar_ext[e]>=0) ? sB[ar_ext[e]*NX + k] : 0.0;
					wext[e] = fma(b, lkj, wext[e]);
					}
				}
			if(j<NX) wext[E-1] += Lp[coff + (NZ-1-(NU+j))];      /* gradient row: + l_x' */
			}
		/* ---- W -> smem ---- */
		if(l<RO)
			{
			#pragma unroll
			for(int j=0; j<NX; j+=2)
				*reinterpret_cast<double2*>(sW + l*LDW + j) = make_double2(w[j], (j+1<NX) ? w[j+1] : 0.0);
			}
		if(E>0 && l<NX)
			{
			#pragma unroll
			for(int e=0; e<E; e++) sW[(RO+e)*LDW + l] = wext[e];
			}
		__syncwarp();
		/* ---- H += W W' ---- */
		#pragma unroll
		for(int k=0; k<CO; k++)
			{
			double acc = Hrow[k];
			#pragma unroll
			for(int m=0; m<NX; m+=2)
				{
				double2 t = *reinterpret_cast<const double2*>(sW + k*LDW + m);
				acc = fma(w[m], t.x, acc);
				if(m+1<NX) acc = fma(w[m+1], t.y, acc);
				}
			Hrow[k] = acc;
			}
		if(E>0)
			{
			double wr[E>0 ? E : 1][NX];
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				double acc = Hext[e];
				#pragma unroll
				for(int m=0; m<NX; m+=2)
					{
					double2 t = *reinterpret_cast<const double2*>(sW + (RO+e)*LDW + m);
					wr[e][m] = t.x; if(m+1<NX) wr[e][m+1] = t.y;
					acc = fma(w[m], t.x, acc);
					if(m+1<NX) acc = fma(w[m+1], t.y, acc);
					}
				Hext[e] = acc;
				}
			#pragma unroll
			for(int e=0; e<E; e++)
				#pragma unroll
				for(int cc=0; cc<NCC; cc++)
					if(e>=cc)
						{
						double acc = Hcor[e][cc];
						#pragma unroll
						for(int m=0; m<NX; m++) acc = fma(wr[e][m], wr[cc][m], acc);
						Hcor[e][cc] = acc;
						}
			}
		}

	/* ---- Cholesky, right-looking; column c is finished on lane c and broadcast through Lc ---- */
	#pragma unroll
	for(int c=0; c<CO; c++)
		{
		const double p = __shfl_sync(HBF_FULL, Hrow[c], c, G);
		const double inv = (p>1e-15) ? rsqrt(p) : 0.0;
		const double lc = Hrow[c]*inv;                      /* lane c: p*inv = sqrt(p) ; lanes > c: L[l][c] */
		double *col = Lc + C::colOff(c);
		if(l>=c && l<RO) col[l-c] = lc;
		if(l==c)
			{
			#pragma unroll
			for(int e=0; e<E; e++) col[RO+e-c] = Hext[e]*inv;
			Lc[C::DINV+c] = inv;
			}
		__syncwarp();
		/* broadcast reads of column c: the pairs (c,c+1), (c+2,c+3), ... are 16-byte aligned */
		#pragma unroll
		for(int q=0; c+2*q<CO; q++)
			{
			const double2 t = *reinterpret_cast<const double2*>(col + 2*q);
			const int k0 = c+2*q, k1 = k0+1;
			if(q>0 && k0<CO) Hrow[k0] = fma(-lc, t.x, Hrow[k0]);
			if(k1<CO) Hrow[k1] = fma(-lc, t.y, Hrow[k1]);
			}
		if(E>0)
			{
			double le[E>0 ? E : 1];
			#pragma unroll
			for(int e=0; e<E; e++) le[e] = col[RO+e-c];
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				Hext[e] = fma(-le[e], lc, Hext[e]);          /* meaningful on lanes > c (column l of the extra row) */
				#pragma unroll
				for(int cc=0; cc<NCC; cc++) if(e>=cc) Hcor[e][cc] = fma(-le[e], le[cc], Hcor[e][cc]);
				}
			}
		}
	/* ---- corner columns (replicated) ---- */
	#pragma unroll
	for(int cc=0; cc<NCC; cc++)
		{
		const double p = Hcor[cc][cc];
		const double inv = (p>1e-15) ? rsqrt(p) : 0.0;
		double *col = Lc + C::colOff(CO+cc);
		double lcol[E>0 ? E : 1];
		#pragma unroll
		for(int e=cc; e<E; e++) lcol[e] = Hcor[e][cc]*inv;
		if(l==0)
			{
			#pragma unroll
			for(int e=cc; e<E; e++) col[e-cc] = lcol[e];
			Lc[C::DINV+CO+cc] = inv;
			}
		#pragma unroll
		for(int e=cc+1; e<E; e++)
			#pragma unroll
			for(int c2=cc+1; c2<NCC; c2++)
				if(e>=c2) Hcor[e][c2] = fma(-lcol[e], lcol[c2], Hcor[e][c2]);
		}
	__syncwarp();
	}

/* ------------------------------------------------------------------------------------------------ */
/* one forward stage: u_n from x_n, then x_{n+1} and pi_n.  x is carried in `xreg` (lane j < NX holds x[j]). */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ __forceinline__ void hbf_stage_forward(int l, const hbf_stage_view v, const double *__restrict__ sB,
		const double *__restrict__ Ln, const double *__restrict__ Ln1, double *__restrict__ vec, double &xreg,
		double *__restrict__ g_u, double *__restrict__ g_x1, double *__restrict__ g_pi, bool active)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, NZ = C::NZ, G = C::G;
	static_assert(NX%2==0, "NX must be even (16-byte rows)");
	static_assert(NU<=G && NX<=G, "forward sweep keeps one column per lane");
	double *ux = vec, *x1 = vec + C::even(NZ), *tmp = vec + 2*C::even(NZ);
	/* x_n -> smem (phantom x block at stage 0 is zero) */
	if(l<NX) ux[NU+l] = v.hx ? xreg : 0.0;
	__syncwarp();
	/* t_c = l_u[c] + sum_k Lxu[k][c] x[k] */
	double t = 0.0;
	int coff = 0;
	#pragma unroll
	for(int cc=0; cc<NU; cc++) if(cc==l) coff = C::colOff(cc);
	if(l<NU)
		{
		const double *col = Ln + coff - l;                 /* col[k] = L[k][l] */
		t = col[NZ-1];
		#pragma unroll
		for(int k=0; k<NX; k++) t = fma(col[NU+k], ux[NU+k], t);
		t = -t;
		}
	/* back substitution with Luu' */
	const double di = (l<NU) ? Ln[C::DINV+l] : 0.0;
	#pragma unroll
	for(int j=NU-1; j>=0; j--)
		{
		const double vj = __shfl_sync(HBF_FULL, t*di, j, G);
		if(l==j) t = vj;
		else if(l<j) t = fma(-Ln[coff - l + j], vj, t);
		}
	if(l<NU) { ux[l] = t; if(active) g_u[l] = t; }
	__syncwarp();
	/* x_{n+1} = b + B u + A x */
	double xn = 0.0;
	if(l<NX)
		{
		xn = sB[v.nux_n*NX + l];
		#pragma unroll
		for(int i=0; i<NU; i++) xn = fma(sB[i*NX+l], ux[i], xn);
		if(v.hx)
			{
			#pragma unroll
			for(int i=0; i<NX; i++) xn = fma(sB[(v.nu_n+i)*NX+l], ux[NU+i], xn);
			}
		x1[l] = xn;
		if(active) g_x1[l] = xn;
		}
	xreg = xn;
	__syncwarp();
	/* pi = Lxx' (Lxx'^T x' + l_x')  on L_{n+1} */
	int coff1 = 0;
	#pragma unroll
	for(int jj=0; jj<NX; jj++) if(jj==l) coff1 = C::colOff(NU+jj);
	if(l<NX)
		{
		const double *col = Ln1 + coff1 - (NU+l);          /* col[r] = L1[r][NU+l] */
		double acc = col[NZ-1];
		#pragma unroll
		for(int k=0; k<NX; k++) if(k>=l) acc = fma(col[NU+k], x1[k], acc);
		tmp[l] = acc;
		}
	__syncwarp();
	if(l<NX)
		{
		double acc = 0.0;
		#pragma unroll
		for(int cc=0; cc<NX; cc++)
			if(cc<=l) acc = fma(Ln1[C::colOff(NU+cc) + (l-cc)], tmp[cc], acc);
		if(active) g_pi[l] = acc;
		}
	__syncwarp();
	}

/* ------------------------------------------------------------------------------------------------ */
/* kernel: persistent warps, IPW instances per warp                                                  */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__global__ void __launch_bounds__(256) hbf_ric_sv_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ stash)
	{
	constexpr int G = C::G, IPW = C::IPW, NX = C::NX, NU = C::NU, INB = C::INB, LBUF = C::LBUF;
	extern __shared__ __align__(16) double hbf_smem[];
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const int g = lane/G, l = lane%G;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	double *wbase = hbf_smem + (size_t)warp*C::PER_WARP;
	uint64_t *bars = reinterpret_cast<uint64_t*>(wbase);          /* [0..1] backward stage inputs, [2..3] forward loads */
	double *ibase = wbase + 8 + (size_t)g*C::PER_INST;
	double *inb[2] = { ibase, ibase + INB };
	double *Lb[3] = { ibase + 2*INB, ibase + 2*INB + LBUF, ibase + 2*INB + 2*LBUF };
	double *sW = ibase + 2*INB + 3*LBUF;
	double *vec = sW + C::NZ*C::LDW;
	if(lane==0)
		{
		for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	__syncwarp();
	uint32_t ph_in = 0, ph_F = 0;                                  /* parity bits of the barriers */
	const int N = d.N;
	const long long stash_stride = (long long)(N+1)*LBUF;
	const long long n_groups = (n_inst + IPW - 1)/IPW;
	double *stash_w = stash + gw*IPW*stash_stride;                 /* this warp's IPW stash slots */

	for(long long grp=gw; grp<n_groups; grp+=tw)
		{
		long long inst = grp*IPW + g;
		const bool active = inst<n_inst;
		if(!active) inst = n_inst-1;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;

		/* lane 0 moves the data of every instance of the warp; one mbarrier per buffer slot */
		auto issue_backward = [&](int n, int slot)            /* [B A b]'_n | RSQrq_n -> inb[slot] */
			{
			if(lane==0)
				{
				const hb_stage s = d.st[n];
				const int nux = s.nu+s.nx;
				const uint32_t bytes = 8u*(uint32_t)(HB_EVEN((nux+1)*s.nx1) + HB_EVEN(HB_TRI(nux)+nux));
				hbf_mbar_expect(&bars[slot], bytes*IPW);
				for(int gg=0; gg<IPW; gg++)
					{
					long long ii = grp*IPW + gg; if(ii>=n_inst) ii = n_inst-1;
					hbf_bulk_g2s(wbase + 8 + (size_t)gg*C::PER_INST + slot*INB, in + ii*d.in_stride + s.off_BAbt, bytes, &bars[slot]);
					}
				}
			};
		auto issue_forward = [&](int n, int islot, int lslot)  /* [B A b]'_n -> inb[islot] ; L_{n+1} -> Lb[lslot] */
			{
			if(lane==0)
				{
				const hb_stage s = d.st[n];
				const int nux = s.nu+s.nx;
				const uint32_t bytes = 8u*(uint32_t)HB_EVEN((nux+1)*s.nx1);
				hbf_mbar_expect(&bars[2+islot], (bytes + 8u*LBUF)*IPW);
				for(int gg=0; gg<IPW; gg++)
					{
					long long ii = grp*IPW + gg; if(ii>=n_inst) ii = n_inst-1;
					double *ib = wbase + 8 + (size_t)gg*C::PER_INST;
					hbf_bulk_g2s(ib + islot*INB, in + ii*d.in_stride + s.off_BAbt, bytes, &bars[2+islot]);
					hbf_bulk_g2s(ib + 2*INB + lslot*LBUF, stash_w + gg*stash_stride + (long long)(n+1)*LBUF, 8u*LBUF, &bars[2+islot]);
					}
				}
			};

		/* ---------------- backward sweep: stage n uses inb[slot], writes Lb[slot], reads Lb[slot^1] ---------------- */
		issue_backward(N, 0);
		int slot = 0;
		for(int n=N; n>=0; n--, slot^=1)
			{
			if(n>0) issue_backward(n-1, slot^1);
			const hb_stage s = d.st[n];
			hbf_stage_view v; v.hu = (s.nu==NU); v.hx = (s.nx==NX); v.nu_n = s.nu; v.nux_n = s.nu+s.nx;
			/* the factor of stage n+2 was stored from Lb[slot]: that bulk store must have finished reading smem */
			if(lane==0) hbf_bulk_wait_read<1>();
			hbf_mbar_wait(&bars[slot], (ph_in>>slot)&1); ph_in ^= (1u<<slot);
			__syncwarp();
			const double *sB = inb[slot];
			const double *sQ = sB + HB_EVEN((v.nux_n+1)*s.nx1);
			hbf_stage_backward<C>(l, v, n==N, sB, sQ, Lb[slot^1], Lb[slot], sW);
			hbf_fence_async();
			__syncwarp();
			if(lane==0)
				{
				for(int gg=0; gg<IPW; gg++)
					hbf_bulk_s2g(stash_w + gg*stash_stride + (long long)n*LBUF, wbase + 8 + (size_t)gg*C::PER_INST + 2*INB + slot*LBUF, 8u*LBUF);
				hbf_bulk_commit();
				}
			}
		/* stage 0 ran with slot0 = slot^1: inb[slot0] = [B A b]'_0, Lb[slot0] = L_0, Lb[slot0^1] = L_1, Lb[2] free */
		if(lane==0) hbf_bulk_wait_all<0>();        /* every factor is in the stash before any is read back */
		__syncwarp();

		/* ---------------- forward sweep: ring of three factor buffers, inputs double-buffered ---------------- */
		int r0 = slot^1, r1 = slot, r2 = 2;        /* L_n, L_{n+1}, prefetch target for L_{n+2} */
		int is = slot^1;                           /* inb slot holding [B A b]'_n */
		double xreg = 0.0;
		for(int n=0; n<N; n++)
			{
			if(n+1<N) issue_forward(n+1, is^1, r2);            /* [B A b]'_{n+1} and L_{n+2} for the next stage */
			const hb_stage s = d.st[n];
			const hb_stage s1 = d.st[n+1];
			hbf_stage_view v; v.hu = (s.nu==NU); v.hx = (s.nx==NX); v.nu_n = s.nu; v.nux_n = s.nu+s.nx;
			if(n>0) { hbf_mbar_wait(&bars[2+is], (ph_F>>is)&1); ph_F ^= (1u<<is); }
			hbf_stage_forward<C>(l, v, inb[is], Lb[r0], Lb[r1], vec, xreg, ux + s.off_ux, ux + s1.off_ux + s1.nu, pi + s.off_pi, active);
			const int t = r0; r0 = r1; r1 = r2; r2 = t;
			is ^= 1;
			}
		__syncwarp();
		}
	}
