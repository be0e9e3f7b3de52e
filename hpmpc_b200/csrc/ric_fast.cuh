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
#ifdef HBF_TIMING
/* debug build only (make dbg): warp 0 of the grid records clock64() at phase boundaries */
__device__ long long *hbf_dbg = nullptr;
__device__ int hbf_dbg_n = 0;
__device__ int hbf_dbg_gen = 0;
/* cheap stamp: the running index lives in shared memory (a global counter costs an L2 round trip per stamp) */
__device__ __forceinline__ void hbf_stamp_(long long tag)
	{
	__shared__ int cnt_init, cnt;
	if(blockIdx.x==0 && threadIdx.x==0 && hbf_dbg!=nullptr)
		{
		const int gen = hbf_dbg_gen;
		if(cnt_init!=gen) { cnt_init = gen; cnt = 0; }
		if(cnt<3990) { const long long t = clock64(); hbf_dbg[2*cnt] = tag; hbf_dbg[2*cnt+1] = t; cnt++; }
		}
	}
#define HBF_STAMP(tag) hbf_stamp_(tag)
#else
#define HBF_STAMP(tag) do { } while(0)
#endif
#ifndef HBF_RSQRT_ITERS
#define HBF_RSQRT_ITERS 2
#endif

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
	static constexpr int BAB = even(NZ*NX);                                  /* [B A b]' of a full stage */
	static constexpr int INB = BAB + even(HB_TRI(NUX)+NUX);                  /* one stage of inputs: [B A b]' then RSQrq */
	static constexpr int WSZ = even(NZ*LDW);
	__host__ __device__ static constexpr int max3(int a, int b, int c) { return a>b ? (a>c ? a : c) : (b>c ? b : c); }
	/* one buffer serves as stage input, then as W (backward); as two [B A b]' slots (forward) */
	static constexpr int IOB = max3(INB, WSZ, 2*BAB);
	static constexpr int XS = even(NX);
	static constexpr int VEC = even(NU) + 3*XS;                              /* u, x (two slots), tmp */
	static constexpr int PER_INST = IOB + 2*LBUF + VEC;                      /* doubles of smem per instance */
	static constexpr int PER_WARP = IPW*PER_INST + 8;                        /* + mbarriers (8 doubles) */
	static constexpr int MINB = (CO+2*NX>48) ? 1 : 2;                        /* 2 CTAs of 8 warps per SM = 128 registers per thread */
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
/* same with an L2 eviction-priority hint (createpolicy) */
__device__ __forceinline__ uint64_t hbf_policy_evict_first()
	{ uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t hbf_policy_evict_last()
	{ uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void hbf_bulk_g2s_hint(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t policy)
	{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		:: "r"(hbf_saddr(sdst)), "l"(gsrc), "r"(bytes), "r"(hbf_saddr(bar)), "l"(policy) : "memory");
	}
__device__ __forceinline__ void hbf_st_hint(double *p, double v, uint64_t policy)
	{ asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(p), "d"(v), "l"(policy) : "memory"); }
__device__ __forceinline__ void hbf_bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes)
	{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(hbf_saddr(ssrc)), "r"(bytes) : "memory");
	}
__device__ __forceinline__ void hbf_bulk_s2g_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t policy)
	{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
		:: "l"(gdst), "r"(hbf_saddr(ssrc)), "r"(bytes), "l"(policy) : "memory");
	}
__device__ __forceinline__ uint64_t hbf_policy_evict_normal()
	{ uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
/* drop one 128-byte line from L2 without writing it back (scratch that has just been consumed and will be rewritten) */
__device__ __forceinline__ void hbf_discard_line(const void *g)
	{ asm volatile("discard.global.L2 [%0], 128;" :: "l"(g) : "memory"); }
__device__ __forceinline__ void hbf_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void hbf_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template<int N> __device__ __forceinline__ void hbf_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void hbf_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* 1/sqrt(p) in FP64: hardware seed (MUFU.RSQ64H, ~20 bits) + two Newton steps; ~2 ulp, no slow-path branch */
__device__ __forceinline__ double hbf_rsqrt(double p)
	{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
	const double h = 0.5*p;
	double t = h*y, e = fma(-t, y, 0.5);
	y = fma(y, e, y);
	t = h*y; e = fma(-t, y, 0.5);
	y = fma(y, e, y);
#if HBF_RSQRT_ITERS>2
	t = h*y; e = fma(-t, y, 0.5);
	y = fma(y, e, y);
#endif
	return y;
	}

/* stage kinds inside the (NU, NX) frame */
enum { HBF_FIRST = 0 /* nu = NU, nx = 0 */, HBF_MID = 1, HBF_LAST = 2 /* nu = 0, nx = NX */ };

template<class C>
__device__ __forceinline__ int hbf_arow(int kind, int f)
	{
	/* frame row f -> actual row of the stage's matrices, or -1 when the row is a phantom */
	if(kind==HBF_MID) return f;
	if(f<C::NU) return kind==HBF_FIRST ? f : -1;
	if(f<C::NUX) return kind==HBF_LAST ? f - C::NU : -1;
	return kind==HBF_FIRST ? C::NU : C::NX;
	}

template<class C>
struct hbf_tile
	{
	double Hrow[C::CO];                                   /* own row, columns < own index */
	double hd;                                            /* own diagonal */
	double Hext[C::E>0 ? C::E : 1];                       /* column-owned entries of the extra rows */
	double Hcor[C::E>0 ? C::E : 1][C::NCC>0 ? C::NCC : 1];/* replicated corner */
	};

/* ------------------------------------------------------------------------------------------------ */
/* backward stage, part 1: H <- RSQrq_n + W W',  W = [B A b]'_n Lxx_{n+1}.  `io` holds the stage inputs   */
/* on entry and W afterwards (dead on return).                                                        */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ __forceinline__ void hbf_back_assemble(int l, int kind, double *__restrict__ io, int rsq_off,
		const double *__restrict__ Lp, hbf_tile<C> &T)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, NZ = C::NZ, RO = C::RO, E = C::E, CO = C::CO, NCC = C::NCC, LDW = C::LDW;
	const double *sB = io, *sQ = io + rsq_off;
	if(kind==HBF_MID)
		{
		const double *row = sQ + HB_TRI(l);
		#pragma unroll
		for(int k=0; k<CO; k++) T.Hrow[k] = (k<l && l<RO) ? row[k] : 0.0;
		T.hd = (l<CO) ? row[l] : 1.0;
		#pragma unroll
		for(int e=0; e<E; e++)
			{
			T.Hext[e] = (l<CO) ? sQ[HB_TRI(RO+e)+l] : 0.0;
			#pragma unroll
			for(int cc=0; cc<NCC; cc++) T.Hcor[e][cc] = (e>=cc) ? sQ[HB_TRI(RO+e)+CO+cc] : 0.0;
			}
		}
	else
		{
		const int ar = (l<RO) ? hbf_arow<C>(kind, l) : -1;
		#pragma unroll
		for(int k=0; k<CO; k++)
			{
			const int ak = hbf_arow<C>(kind, k);
			T.Hrow[k] = (k<l && ar>=0 && ak>=0) ? sQ[HB_TRI(ar)+ak] : 0.0;
			}
		T.hd = (l<CO && ar>=0) ? sQ[HB_TRI(ar)+ar] : 1.0;
		#pragma unroll
		for(int e=0; e<E; e++)
			{
			const int ae = hbf_arow<C>(kind, RO+e);
			const int ac = (l<CO) ? hbf_arow<C>(kind, l) : -1;
			T.Hext[e] = (ae>=0 && ac>=0) ? sQ[HB_TRI(ae)+ac] : 0.0;
			#pragma unroll
			for(int cc=0; cc<NCC; cc++)
				{
				const int acol = hbf_arow<C>(kind, CO+cc);
				double hc = (e==cc) ? 1.0 : 0.0;
				if(e>=cc && ae>=0 && acol>=0) hc = sQ[HB_TRI(ae)+acol];
				T.Hcor[e][cc] = hc;
				}
			}
		}
	if(kind==HBF_LAST) { __syncwarp(); return; }

	/* ---- W = [B A b]' Lxx' ---- */
	double w[NX];
	{
	const int ar = (l<RO) ? hbf_arow<C>(kind, l) : -1;
	double a[NX];
	#pragma unroll
	for(int k=0; k<NX; k+=2)
		{
		double2 t = make_double2(0.0, 0.0);
		if(ar>=0) t = *reinterpret_cast<const double2*>(sB + ar*NX + k);
		a[k] = t.x; a[k+1] = t.y;
		}
	#pragma unroll
	for(int j=0; j<NX; j++)
		{
		const double *col = Lp + C::colOff(NU+j);              /* col[k-j] = Lxx'[k][j] */
		double acc0 = 0.0, acc1 = 0.0;
		#pragma unroll
		for(int k=j; k<NX; k+=2)
			{
			const double2 t = *reinterpret_cast<const double2*>(col + (k-j));
			acc0 = fma(a[k], t.x, acc0);
			if(k+1<NX) acc1 = fma(a[k+1], t.y, acc1);
			}
		w[j] = acc0 + acc1;
		}
	if(E==0 && l==NUX)              /* gradient row is row-owned: add l_x' */
		{
		#pragma unroll
		for(int j=0; j<NX; j++) w[j] += Lp[C::colOff(NU+j) + (NZ-1-(NU+j))];
		}
	}
	double wext[E>0 ? E : 1];
	if(E>0)
		{
		/* extra rows: lane j owns column j of them */
		int ae[E>0 ? E : 1];
		#pragma unroll
		for(int e=0; e<E; e++) { wext[e] = 0.0; ae[e] = hbf_arow<C>(kind, RO+e); }
		const int j = l;
		int coff = 0;
		#pragma unroll
		for(int jj=0; jj<NX; jj++) if(jj==j) coff = C::colOff(NU+jj);
		#pragma unroll
		for(int k=0; k<NX; k++)
			{
			const double lkj = (j<NX && k>=j) ? Lp[coff + (k-j)] : 0.0;
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				const double b = (ae[e]>=0) ? sB[ae[e]*NX + k] : 0.0;
				wext[e] = fma(b, lkj, wext[e]);
				}
			}
		if(j<NX) wext[E-1] += Lp[coff + (NZ-1-(NU+j))];       /* gradient row: + l_x' */
		}
	__syncwarp();                       /* every lane has read its inputs: the buffer becomes W */
	double *sW = io;
	if(l<RO)
		{
		#pragma unroll
		for(int j=0; j<NX; j+=2) *reinterpret_cast<double2*>(sW + l*LDW + j) = make_double2(w[j], w[j+1]);
		}
	if(E>0 && l<NX)
		{
		#pragma unroll
		for(int e=0; e<E; e++) sW[(RO+e)*LDW + l] = wext[e];
		}
	__syncwarp();
	/* ---- H += W W' ---- */
	{
	double acc = T.hd;
	#pragma unroll
	for(int m=0; m<NX; m++) acc = fma(w[m], w[m], acc);
	T.hd = acc;
	}
	/* m outer, k inner: CO independent accumulator chains keep the FP64 pipe busy */
	constexpr int KS = (RO>CO) ? CO : CO-1;     /* off-diagonal columns that some row-owned row uses */
	#pragma unroll
	for(int m=0; m<NX; m+=2)
		{
		#pragma unroll
		for(int k=0; k<KS; k++)
			{
			const double2 t = *reinterpret_cast<const double2*>(sW + k*LDW + m);
			T.Hrow[k] = fma(w[m], t.x, T.Hrow[k]);
			T.Hrow[k] = fma(w[m+1], t.y, T.Hrow[k]);
			}
		}
	if(E>0)
		{
		double wr[E>0 ? E : 1][NX];
		#pragma unroll
		for(int e=0; e<E; e++)
			{
			double acc0 = T.Hext[e], acc1 = 0.0;
			#pragma unroll
			for(int m=0; m<NX; m+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(sW + (RO+e)*LDW + m);
				wr[e][m] = t.x; wr[e][m+1] = t.y;
				acc0 = fma(w[m], t.x, acc0);
				acc1 = fma(w[m+1], t.y, acc1);
				}
			T.Hext[e] = acc0 + acc1;
			}
		#pragma unroll
		for(int e=0; e<E; e++)
			#pragma unroll
			for(int cc=0; cc<NCC; cc++)
				if(e>=cc)
					{
					double acc = T.Hcor[e][cc];
					#pragma unroll
					for(int m=0; m<NX; m++) acc = fma(wr[e][m], wr[cc][m], acc);
					T.Hcor[e][cc] = acc;
					}
		}
	__syncwarp();                       /* W is dead: the caller may refill the buffer */
	}

/* ------------------------------------------------------------------------------------------------ */
/* backward stage, part 2: right-looking Cholesky with a look-ahead diagonal; writes the packed column */
/* buffer Lc (column c is finished on lane c and broadcast through shared memory)                      */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ __forceinline__ void hbf_back_factor(int l, hbf_tile<C> &T, double *__restrict__ Lc)
	{
	constexpr int RO = C::RO, E = C::E, CO = C::CO, NCC = C::NCC, G = C::G;
	#pragma unroll
	for(int c=0; c<CO; c++)
		{
		/* lane c's diagonal is final here; every lane evaluates the same instruction stream */
		const double rs = (T.hd>1e-15) ? hbf_rsqrt(T.hd) : 0.0;
		const double inv = __shfl_sync(HBF_FULL, rs, c, G);
		const double lc = (l==c) ? T.hd*inv : T.Hrow[c]*inv;      /* lane c: sqrt(p) ; lanes > c: L[l][c] */
		if(l>c) T.hd = fma(-lc, lc, T.hd);                        /* look-ahead: the next pivot does not wait for smem */
		double *col = Lc + C::colOff(c);
		if(l>=c && l<RO) col[l-c] = lc;
		if(l==c)
			{
			#pragma unroll
			for(int e=0; e<E; e++) col[RO+e-c] = T.Hext[e]*inv;
			Lc[C::DINV+c] = inv;
			}
		__syncwarp();
		#pragma unroll
		for(int q=0; c+2*q<CO; q++)
			{
			const double2 t = *reinterpret_cast<const double2*>(col + 2*q);
			const int k0 = c+2*q, k1 = k0+1;
			if(q>0 && k0<CO) T.Hrow[k0] = fma(-lc, t.x, T.Hrow[k0]);
			if(k1<CO) T.Hrow[k1] = fma(-lc, t.y, T.Hrow[k1]);
			}
		if(E>0)
			{
			double le[E>0 ? E : 1];
			#pragma unroll
			for(int e=0; e<E; e++) le[e] = col[RO+e-c];
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				T.Hext[e] = fma(-le[e], lc, T.Hext[e]);            /* meaningful on lanes > c */
				#pragma unroll
				for(int cc=0; cc<NCC; cc++) if(e>=cc) T.Hcor[e][cc] = fma(-le[e], le[cc], T.Hcor[e][cc]);
				}
			}
		}
	#pragma unroll
	for(int cc=0; cc<NCC; cc++)
		{
		const double p = T.Hcor[cc][cc];
		const double inv = (p>1e-15) ? hbf_rsqrt(p) : 0.0;
		double *col = Lc + C::colOff(CO+cc);
		double lcol[E>0 ? E : 1];
		#pragma unroll
		for(int e=cc; e<E; e++) lcol[e] = T.Hcor[e][cc]*inv;
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
				if(e>=c2) T.Hcor[e][c2] = fma(-lcol[e], lcol[c2], T.Hcor[e][c2]);
		}
	__syncwarp();
	}

/* ------------------------------------------------------------------------------------------------ */
/* forward stage n: (1) pi_{n-1} from x_n and L_n, (2) u_n, (3) x_{n+1}.                              */
/*   xs holds x_n on entry (lane j < NX also has it in xreg), xo receives x_{n+1}                      */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ __forceinline__ void hbf_stage_forward(int l, int kind, bool do_pi, const double *__restrict__ sB,
		const double *__restrict__ Ln, double *__restrict__ us, const double *__restrict__ xs, double *__restrict__ xo,
		double *__restrict__ tmp, double &xreg, double *__restrict__ g_u, double *__restrict__ g_x1,
		double *__restrict__ g_pi, bool active)
	{
	constexpr int NX = C::NX, NU = C::NU, NZ = C::NZ, G = C::G;
	static_assert(NX%2==0, "NX must be even (16-byte rows)");
	static_assert(NU<=G && NX<=G, "forward sweep keeps one column per lane");
	/* ---- phase A: t = l_u + Lxu' x (lanes < NU) and tmp = l_x + Lxx' x (lanes < NX), independent chains ---- */
	int coff = 0, coff1 = 0;
	#pragma unroll
	for(int cc=0; cc<NU; cc++) if(cc==l) coff = C::colOff(cc);
	#pragma unroll
	for(int jj=0; jj<NX; jj++) if(jj==l) coff1 = C::colOff(NU+jj);
	double t = 0.0;
	if(l<NU)
		{
		const double *col = Ln + coff - l;                     /* col[k] = L[k][l] */
		double t0 = col[NZ-1], t1 = 0.0;
		#pragma unroll
		for(int k=0; k<NX; k+=2) { t0 = fma(col[NU+k], xs[k], t0); t1 = fma(col[NU+k+1], xs[k+1], t1); }
		t = -(t0+t1);
		}
	if(do_pi && l<NX)
		{
		const double *col = Ln + coff1 - (NU+l);               /* col[r] = L[r][NU+l] */
		double a0 = col[NZ-1], a1 = 0.0;
		#pragma unroll
		for(int k=0; k<NX; k+=2)
			{
			if(k>=l) a0 = fma(col[NU+k], xs[k], a0);
			if(k+1>=l) a1 = fma(col[NU+k+1], xs[k+1], a1);
			}
		tmp[l] = a0+a1;
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
	if(l<NU) { us[l] = t; if(active) g_u[l] = t; }
	__syncwarp();
	/* ---- phase B: x_{n+1} = b + B u + A x  and  pi_{n-1} = Lxx tmp, independent chains ---- */
	double xn = 0.0;
	if(l<NX)
		{
		const int brow = (kind==HBF_FIRST) ? NU : C::NUX;
		double x0 = sB[brow*NX + l], x1 = 0.0, x2 = 0.0;
		#pragma unroll
		for(int i=0; i<NU; i++) x0 = fma(sB[i*NX+l], us[i], x0);
		if(kind!=HBF_FIRST)
			{
			#pragma unroll
			for(int i=0; i<NX; i+=2) { x1 = fma(sB[(NU+i)*NX+l], xs[i], x1); x2 = fma(sB[(NU+i+1)*NX+l], xs[i+1], x2); }
			}
		xn = x0 + (x1+x2);
		xo[l] = xn;
		if(active) g_x1[l] = xn;
		if(do_pi)
			{
			double p0 = 0.0, p1 = 0.0;
			#pragma unroll
			for(int cc=0; cc<NX; cc+=2)
				{
				if(cc<=l) p0 = fma(Ln[C::colOff(NU+cc) + (l-cc)], tmp[cc], p0);
				if(cc+1<=l) p1 = fma(Ln[C::colOff(NU+cc+1) + (l-cc-1)], tmp[cc+1], p1);
				}
			if(active) g_pi[l] = p0+p1;
			}
		}
	xreg = xn;
	__syncwarp();
	}

/* pi_{N-1} from x_N and L_N */
template<class C>
__device__ __forceinline__ void hbf_final_pi(int l, const double *__restrict__ Ln, const double *__restrict__ xs,
		double *__restrict__ tmp, double *__restrict__ g_pi, bool active)
	{
	constexpr int NX = C::NX, NU = C::NU, NZ = C::NZ;
	int coff1 = 0;
	#pragma unroll
	for(int jj=0; jj<NX; jj++) if(jj==l) coff1 = C::colOff(NU+jj);
	if(l<NX)
		{
		const double *col = Ln + coff1 - (NU+l);
		double a0 = col[NZ-1];
		#pragma unroll
		for(int k=0; k<NX; k++) if(k>=l) a0 = fma(col[NU+k], xs[k], a0);
		tmp[l] = a0;
		}
	__syncwarp();
	if(l<NX)
		{
		double p0 = 0.0;
		#pragma unroll
		for(int cc=0; cc<NX; cc++) if(cc<=l) p0 = fma(Ln[C::colOff(NU+cc) + (l-cc)], tmp[cc], p0);
		if(active) g_pi[l] = p0;
		}
	__syncwarp();
	}

/* ------------------------------------------------------------------------------------------------ */
/* kernel: persistent warps, IPW instances per warp                                                  */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__global__ void __launch_bounds__(256, C::MINB) hbf_ric_sv_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ stash)
	{
	constexpr int G = C::G, IPW = C::IPW, NX = C::NX, NU = C::NU, NUX = C::NUX, LBUF = C::LBUF, IOB = C::IOB, BAB = C::BAB;
	extern __shared__ __align__(16) double hbf_smem[];
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const int g = lane/G, l = lane%G;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	double *wbase = hbf_smem + (size_t)warp*C::PER_WARP;
	uint64_t *bars = reinterpret_cast<uint64_t*>(wbase);      /* [0] stage inputs / [B A b]' slot 0, [1] [B A b]' slot 1, [2..3] factors */
	double *ibase = wbase + 8 + (size_t)g*C::PER_INST;
	double *io = ibase;
	double *Lb0 = ibase + IOB, *Lb1 = Lb0 + LBUF;
	double *us = Lb1 + LBUF, *xs0 = us + C::even(NU), *xs1 = xs0 + C::XS, *tmp = xs1 + C::XS;
	if(lane==0)
		{
		for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	__syncwarp();
	uint32_t phase = 0;                                           /* parity bits of the 4 barriers */
	const int N = d.N;
	/* uniform pattern: stage 0 (nu = NU, nx = 0), stages 1..N-1 full, stage N (nu = 0): offsets are arithmetic */
	const int o_in1 = d.st[1].off_BAbt, s_in = d.st[2].off_BAbt - d.st[1].off_BAbt, o_inN = d.st[N].off_BAbt;
	constexpr uint32_t bytes_first = 8u*(uint32_t)(C::even((NU+1)*NX) + C::even(HB_TRI(NU)+NU));
	constexpr uint32_t bytes_mid = 8u*(uint32_t)C::INB;
	constexpr uint32_t bytes_last = 8u*(uint32_t)C::even(HB_TRI(NX)+NX);
	const long long stash_stride = (long long)(N+1)*LBUF;
	const long long n_groups = (n_inst + IPW - 1)/IPW;
	double *stash_w = stash + gw*IPW*stash_stride;                 /* this warp's IPW stash slots */

	for(long long grp=gw; grp<n_groups; grp+=tw)
		{
		long long inst = grp*IPW + g;
		const bool active = inst<n_inst;
		if(!active) inst = n_inst-1;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;

		/* lane 0 moves the data of every instance of the warp; one mbarrier per buffer */
		auto issue_backward = [&](int n)                          /* [B A b]'_n | RSQrq_n -> io */
			{
			if(lane==0)
				{
				const int off = (n==0) ? 0 : (n==N ? o_inN : o_in1 + (n-1)*s_in);
				const uint32_t bytes = (n==0) ? bytes_first : (n==N ? bytes_last : bytes_mid);
				hbf_mbar_expect(&bars[0], bytes*IPW);
				#pragma unroll
				for(int gg=0; gg<IPW; gg++)
					{
					long long ii = grp*IPW + gg; if(ii>=n_inst) ii = n_inst-1;
					hbf_bulk_g2s(wbase + 8 + (size_t)gg*C::PER_INST, in + ii*d.in_stride + off, bytes, &bars[0]);
					}
				}
			};
		auto issue_BAbt = [&](int n, int slot)                    /* forward: [B A b]'_n -> io + slot*BAB */
			{
			if(lane==0)
				{
				const int off = (n==0) ? 0 : o_in1 + (n-1)*s_in;
				const uint32_t bytes = (n==0) ? 8u*(uint32_t)C::even((NU+1)*NX) : 8u*(uint32_t)BAB;
				hbf_mbar_expect(&bars[slot], bytes*IPW);
				#pragma unroll
				for(int gg=0; gg<IPW; gg++)
					{
					long long ii = grp*IPW + gg; if(ii>=n_inst) ii = n_inst-1;
					hbf_bulk_g2s(wbase + 8 + (size_t)gg*C::PER_INST + slot*BAB, in + ii*d.in_stride + off, bytes, &bars[slot]);
					}
				}
			};
		auto issue_L = [&](int n, int slot)                       /* forward: L_n -> Lb[slot] */
			{
			if(lane==0)
				{
				hbf_mbar_expect(&bars[2+slot], 8u*LBUF*IPW);
				#pragma unroll
				for(int gg=0; gg<IPW; gg++)
					hbf_bulk_g2s(wbase + 8 + (size_t)gg*C::PER_INST + IOB + slot*LBUF, stash_w + gg*stash_stride + (long long)n*LBUF, 8u*LBUF, &bars[2+slot]);
				}
			};
		auto wait_bar = [&](int b) { hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b); };

		/* ---------------- backward sweep: stage n writes Lb[n&1], reads Lb[(n+1)&1] ---------------- */
		issue_backward(N);
		for(int n=N; n>=0; n--)
			{
			const int kind = (n==0) ? HBF_FIRST : (n==N ? HBF_LAST : HBF_MID);
			const int rsq_off = (n==0) ? C::even((NU+1)*NX) : (n==N ? 0 : BAB);
			double *Lc = (n&1) ? Lb1 : Lb0;
			const double *Lp = (n&1) ? Lb0 : Lb1;
			hbf_tile<C> T;
			HBF_STAMP(100);
			wait_bar(0);
			HBF_STAMP(101);
			hbf_back_assemble<C>(l, kind, io, rsq_off, Lp, T);
			HBF_STAMP(102);
			if(n>0) issue_backward(n-1);                          /* lands while the factorization runs */
			/* the factor of stage n+2 was stored from this buffer: that bulk store must have finished reading smem */
			if(lane==0) hbf_bulk_wait_read<1>();
			__syncwarp();
			HBF_STAMP(103);
			hbf_back_factor<C>(l, T, Lc);
			HBF_STAMP(104);
			hbf_fence_async();
			__syncwarp();
			if(lane==0)
				{
				#pragma unroll
				for(int gg=0; gg<IPW; gg++)
					hbf_bulk_s2g(stash_w + gg*stash_stride + (long long)n*LBUF, wbase + 8 + (size_t)gg*C::PER_INST + IOB + (n&1)*LBUF, 8u*LBUF);
				hbf_bulk_commit();
				}
			}
		/* now: Lb0 = L_0, Lb1 = L_1 ; io is free */
		if(lane==0) hbf_bulk_wait_all<0>();        /* every factor is in the stash before any is read back */
		__syncwarp();

		/* ---------------- forward sweep: L double-buffered by stage parity, [B A b]' in two slots of io ---------------- */
		issue_BAbt(0, 0);
		if(N>1) issue_BAbt(1, 1);
		double xreg = 0.0;
		if(l<NX) xs0[l] = 0.0;                     /* x_0 is absent (eliminated) */
		__syncwarp();
		for(int n=0; n<N; n++)
			{
			const int kind = (n==0) ? HBF_FIRST : HBF_MID;
			const double *Ln = (n&1) ? Lb1 : Lb0;
			const double *xs = (n&1) ? xs1 : xs0;
			double *xo = (n&1) ? xs0 : xs1;
			HBF_STAMP(200);
			if(n>=2) wait_bar(2+(n&1));                           /* L_n (n = 0, 1 are still resident from the backward sweep) */
			wait_bar(n&1);                                        /* [B A b]'_n */
			HBF_STAMP(201);
			const int o_ux = (n==0) ? 0 : NU + (n-1)*NUX, o_ux1 = NU + n*NUX + ((n+1<N) ? NU : 0);
			hbf_stage_forward<C>(l, kind, n>0, io + (n&1)*BAB, Ln, us, xs, xo, tmp, xreg,
					ux + o_ux, ux + o_ux1, pi + (n-1)*NX, active);
			HBF_STAMP(202);
			/* both buffers of parity n are free now: prefetch stage n+2 */
			if(n+2<=N) issue_L(n+2, n&1);
			if(n+2<N) issue_BAbt(n+2, n&1);
			}
		/* pi_{N-1} needs L_N (parity N) and x_N */
		if(N>=2) wait_bar(2+(N&1));
		hbf_final_pi<C>(l, (N&1) ? Lb1 : Lb0, (N&1) ? xs1 : xs0, tmp, pi + (N-1)*NX, active);
		}
	}
