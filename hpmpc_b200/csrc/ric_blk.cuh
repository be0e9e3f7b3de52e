/*
 * ric_blk.cuh -- register-blocked, size-specialised Riccati factor+solve for sm_100a.
 *
 * G lanes work on one OCP instance (32/G instances per warp) and every lane owns R rows of the stage's
 * (NU+NX+1) x (NU+NX) trapezoid, so one shared-memory broadcast operand feeds R FP64 FMAs and the per-column
 * overhead of the Cholesky (reciprocal square root, shuffles, barriers) is shared by 32/G instances.
 * ncu on the one-row-per-lane kernel (ric_fast.cuh) showed the shared-memory data pipe at 74-86 % and only
 * 37 % of the issued instructions being FP64 math; this mapping halves the LDS count per FMA (R = 2) and the
 * overhead per instance (G = 8 instead of 16 for nx = 12, nu = 5).
 *
 * Frame = NZ x NUX trapezoid, NZ = NU+NX+1, virtual row v = l + s*G (lane l, slot s < R):
 *   rows v < RO = min(NZ, G*R)       "row-owned": H[s][k] (k < v) and the diagonal hd[s] live in registers of lane l
 *   rows RO..NZ-1 (E of them)        "column-owned": entry (row, c) lives on lane c%G, slot c/G  (c < CO = min(NUX, G*R))
 *   columns CO..NUX-1 (NCC)          the E x NCC corner is replicated on every lane
 *
 * What leaves the backward sweep for the forward sweep ("stash", written with one bulk store per stage) is not the
 * whole factor but the part the forward sweep needs, in the form it needs it:
 *   K = -Luu^-T Lxu'  (NU x NX),  k = -Luu^-T l_u,  the packed columns of Lxx with l_x appended
 * so the forward stage is three small matrix-vector products and no triangular solve:
 *   u_n = k + K x_n ,  x_{n+1} = b + B u + A x ,  pi_n = Lxx (Lxx' x_n + l_x)
 *
 * Restates (reference paths relative to /root/reference):
 *   backward stage   lqcp_solvers/d_back_ric_rec.c:236-333      (dtrmm_nt_u, gradient-row add, dsyrk_dpotrf)
 *   pivot rule       kernel/c99/kernel_dpotrf_c99_lib4.c:553-573 (pivot <= 1e-15 -> zero column)
 *   forward stage    lqcp_solvers/d_back_ric_rec.c:341-397      (dtrsv_t + dgemv_t, dtrmv_u_n / dtrmv_u_t for pi)
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "layout.h"
#include "ric_fast.cuh"          /* PTX helpers (mbarrier, bulk copies), stage kinds */

/* state value of a finished tree in the tree IPM driver's per-tree record (tree_ipm_kernels.cu: TS_DONE, record[3]) */
#define HBK_TS_DONE 5

template<int NX_, int NU_, int G_, int R_>
struct hbk_cfg
	{
	static constexpr int NX = NX_, NU = NU_, G = G_, R = R_;
	static constexpr int NUX = NX+NU, NZ = NUX+1;
	static constexpr int GR = G*R;
	static constexpr int RO = NZ<GR ? NZ : GR;
	static constexpr int E = NZ-RO;
	static constexpr int CO = NUX<GR ? NUX : GR;
	static constexpr int NCC = NUX-CO;
	static constexpr int IPW = 32/G;
	__host__ __device__ static constexpr int even(int x) { return (x+1)&~1; }
	/* u-columns of L (c < NU): column c holds rows c..NZ-1 */
	__host__ __device__ static constexpr int uOff(int c) { int o = 0; for(int j=0; j<c; j++) o += even(NZ-j); return o; }
	/* x-columns of L (j < NX, frame column NU+j): rows NU+j..NZ-1, i.e. Lxx[j..NX-1][j] then l_x[j] */
	__host__ __device__ static constexpr int xOff(int j) { int o = 0; for(int i=0; i<j; i++) o += even(NX+1-i); return o; }
	static constexpr int UDINV = uOff(NU);                                   /* inverse diagonal of the u-columns */
	static constexpr int LU = uOff(NU) + even(NU);                           /* scratch: u-columns + their inverse diagonal */
	/* stash image of one stage */
	static constexpr int SK = 0;                                             /* K, NU x NX row-major */
	static constexpr int Sk = even(NU*NX);                                   /* k */
	static constexpr int SX = Sk + even(NU);                                 /* x-columns */
	static constexpr int SB = SX + xOff(NX);
#ifdef HBK_BULK_STASH
	static constexpr int SBG = (SB+15)&~15;                                  /* image stride in the sv kernel's stash: whole 128-byte lines */
#else
	static constexpr int SBG = SB;
#endif
	static constexpr int LDW = ((even(NX)/2)%2==0) ? even(NX)+2 : even(NX);
	static constexpr int BAB = even(NZ*NX);
	static constexpr int RSQ = even(HB_TRI(NUX)+NUX);
	static constexpr int INB = BAB + RSQ;
	static constexpr int WSZ = even(NZ*LDW);
	__host__ __device__ static constexpr int max3(int a, int b, int c) { return a>b ? (a>c ? a : c) : (b>c ? b : c); }
	static constexpr int IOB = max3(INB, WSZ, 2*BAB);
	static constexpr int XS = even(NX);
	static constexpr int VEC = even(NU) + 3*XS;
	static constexpr int PER_INST = IOB + LU + 2*SB + VEC;
#ifndef HBK_NO_SKEW
	/* Bank skew between the instances of a warp.  The instances' buffers start at (0, 8, 4, 12, 2, 10, 6, 14) doubles mod 16:
	 * a broadcast LDS.128 (one 16-byte chunk per instance) then touches disjoint banks for all instances of the warp, and the
	 * one-row-per-lane LDS.64 of the packed triangles (row offsets i(i+1)/2, a complete residue system mod 16 that uses half of
	 * the 16 double-slots) of the two instances that share a half-warp interleave exactly (offset 8) instead of colliding
	 * (ncu r01: 23 % of the shared-memory wavefronts were bank conflicts with the uniform stride PER_INST = 6 mod 16). */
	__host__ __device__ static constexpr int skew_target(int g) { return ((g&1)<<3) | ((g&2)<<1) | ((g&4)>>1); }
	/* start of instance g's buffers: the first offset behind instance g-1 that has the wanted residue (regions never overlap) */
	__host__ __device__ static constexpr int inst_off_c(int g)
		{
		int off = 0;
		for(int k=1; k<=g; k++) { off += PER_INST; off += (skew_target(k) - off) & 15; }
		return off;
		}
	__host__ __device__ static constexpr int inst_off(int g)
		{
		int off = 0;
		for(int k=1; k<IPW; k++) if(g==k) off = inst_off_c(k);
		return off;
		}
	static constexpr int PER_WARP = inst_off_c(IPW-1) + PER_INST + 8;
#else
	__host__ __device__ static constexpr int inst_off(int g) { return g*PER_INST; }
	static constexpr int PER_WARP = IPW*PER_INST + 8;
#endif
	static constexpr int KS = (RO>CO) ? CO : CO-1;                           /* off-diagonal columns some row-owned row uses */
	static_assert(NX%2==0, "NX must be even (16-byte rows)");
	static_assert(NU<=GR && NX<GR, "one u-column / x-column (and the gradient) per lane slot");
	};

/* 1/sqrt(p) with the reference's pivot rule (p <= 1e-15 -> 0): hardware seed (MUFU.RSQ64H, ~22 bits) and two coupled
 * (Goldschmidt) steps on h ~ 1/sqrt(p), g ~ sqrt(p)/2; dependent depth 5 FP64 operations after the seed */
__device__ __forceinline__ double hbk_rsqrt(double p)
	{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
	const double hp = 0.5*p;
	double g = hp*y;
	double r = fma(-g, y, 0.5);
	g = fma(g, r, g); y = fma(y, r, y);
	r = fma(-g, y, 0.5);
	y = fma(y, r, y);
	return (p>1e-15) ? y : 0.0;
	}

template<class C, int KIND>
__device__ __forceinline__ constexpr int hbk_arow(int f)
	{
	if(KIND==HBF_MID) return f;
	if(f<C::NU) return KIND==HBF_FIRST ? f : -1;
	if(f<C::NUX) return KIND==HBF_LAST ? f - C::NU : -1;
	return KIND==HBF_FIRST ? C::NU : C::NX;
	}

template<class C>
struct hbk_tile
	{
	double H[C::R][C::CO];                                /* H[s][k]: row l+s*G, column k (meaningful for k < row) */
	double hd[C::R];                                      /* own diagonals */
	double X[C::E>0 ? C::E : 1][C::R];                    /* column-owned entries of the extra rows */
	double Z[C::E>0 ? C::E : 1][C::NCC>0 ? C::NCC : 1];   /* replicated corner */
	};

/* per-thread constants: offsets of the columns this lane owns */
template<class C>
struct hbk_lane
	{
	int l;
	int uo[C::R];        /* uOff(c), c = l+s*G (u-column), 0 when c >= NU */
	int xo[C::R];        /* xOff(j), j = l+s*G (x-column), 0 when j >= NX */
	__device__ __forceinline__ void init(int l_)
		{
		l = l_;
		#pragma unroll
		for(int s=0; s<C::R; s++)
			{
			uo[s] = 0; xo[s] = 0;
			#pragma unroll
			for(int c=0; c<C::NU; c++) if(c==l+s*C::G) uo[s] = C::uOff(c);
			#pragma unroll
			for(int j=0; j<C::NX; j++) if(j==l+s*C::G) xo[s] = C::xOff(j);
			}
		}
	};

/* pointer to column c of the factor being built: u-columns in the scratch, x-columns in the stash image */
template<class C>
__device__ __forceinline__ double *hbk_col(int c, double *LUs, double *Xc)
	{
	return c<C::NU ? LUs + C::uOff(c) : Xc + C::xOff(c-C::NU);
	}

/* ------------------------------------------------------------------------------------------------ */
/* backward stage, part 1: H <- RSQrq_n + W W',  W = [B A b]'_n Lxx_{n+1}  (io: inputs in, W scratch) */
/* ------------------------------------------------------------------------------------------------ */
struct hbk_nop { __device__ __forceinline__ void operator()() const {} };

/* sBW: [B A b]' of the stage on entry, W (leading dimension LDW) afterwards ; sQ: RSQrq of the stage ; xc: x-columns of L_{n+1} ;
 * pre_h() is called after W has been computed and before RSQrq is read (the place to wait for a separately fetched RSQrq) */
/* ACC: add the W W' of a further kid (scenario-tree nodes with several kids) to an H that is already in the tile */
template<class C, int KIND, int LDW, class F, bool ACC = false>
__device__ __forceinline__ void hbk_back_assemble(const hbk_lane<C> &ln, double *__restrict__ sBW, const double *__restrict__ sQ,
		const double *__restrict__ xc, hbk_tile<C> &T, F pre_h)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, G = C::G, R = C::R, RO = C::RO, E = C::E, CO = C::CO, NCC = C::NCC;
	const int l = ln.l;
	const double *sB = sBW;
	double w[R][NX];
	double wx[E>0 ? E : 1][R];
	if(KIND!=HBF_LAST)
		{
		/* xc: x-columns of L_{n+1}: col j at xOff(j), [k-j] = Lxx[k][j], [NX-j] = l_x[j] */
		/* ---- W = [B A b]' Lxx' : own rows ---- */
		double a[R][NX];
		#pragma unroll
		for(int s=0; s<R; s++)
			{
			const int v = l + s*G;
			int ar = -1;
			if(KIND==HBF_MID) ar = (v<RO) ? v : -1;
			else ar = (v<NU) ? v : ((v==NUX && v<RO) ? NU : -1);              /* FIRST: B' rows and the b row */
			#pragma unroll
			for(int k=0; k<NX; k+=2)
				{
				double2 t = make_double2(0.0, 0.0);
				if(ar>=0) t = *reinterpret_cast<const double2*>(sB + ar*NX + k);
				a[s][k] = t.x; a[s][k+1] = t.y;
				}
			}
		#pragma unroll
		for(int j=0; j<NX; j++)
			{
			const double *col = xc + C::xOff(j);
			double acc[R][2];
			#pragma unroll
			for(int s=0; s<R; s++) { acc[s][0] = 0.0; acc[s][1] = 0.0; }
			#pragma unroll
			for(int k=j; k<NX; k+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(col + (k-j));
				#pragma unroll
				for(int s=0; s<R; s++)
					{
					acc[s][0] = fma(a[s][k], t.x, acc[s][0]);
					if(k+1<NX) acc[s][1] = fma(a[s][k+1], t.y, acc[s][1]);
					}
				}
			#pragma unroll
			for(int s=0; s<R; s++) w[s][j] = acc[s][0] + acc[s][1];
			if(E==0)                    /* gradient row is row-owned: add l_x' */
				{
				const double lx = col[NX-j];
				#pragma unroll
				for(int s=0; s<R; s++) if(l+s*G==NUX) w[s][j] += lx;
				}
			}
		/* ---- extra rows: lane owns columns j = l+s*G of them ---- */
		if(E>0)
			{
			#pragma unroll
			for(int s=0; s<R; s++)
				{
				const int j = l + s*G;
				#pragma unroll
				for(int e=0; e<E; e++) wx[e][s] = 0.0;
				if(s*G<NX)
					{
					const double *col = xc + ln.xo[s] - (j<NX ? j : 0);      /* col[k] = Lxx[k][j], k >= j */
					#pragma unroll
					for(int k=s*G; k<NX; k++)
						{
						const double lkj = (j<NX && k>=j) ? col[k] : 0.0;
						#pragma unroll
						for(int e=0; e<E; e++)
							{
							constexpr int dummy = 0; (void)dummy;
							const int ae = hbk_arow<C, KIND>(RO+e);
							const double b = (ae>=0) ? sB[ae*NX + k] : 0.0;
							wx[e][s] = fma(b, lkj, wx[e][s]);
							}
						}
					if(j<NX) wx[E-1][s] += col[NX];                          /* gradient row: + l_x[j] */
					}
				}
			}
		}
	HBF_STAMP(110);
	if(!ACC)
	{
	pre_h();
	/* ---- H <- RSQrq (before W is stored: with LDW > NX the W rows run over the RSQ part of the buffer) ---- */
	#pragma unroll
	for(int s=0; s<R; s++)
		{
		const int v = l + s*G;
		int ar = -1;
		if(v<RO)
			{
			if(KIND==HBF_MID) ar = v;
			else if(KIND==HBF_FIRST) ar = (v<NU) ? v : (v==NUX ? NU : -1);
			else ar = (v<NU) ? -1 : v-NU;                                    /* LAST: rows NU.. map to 0.. ; gradient row NUX -> NX */
			}
		const double *row = sQ + HB_TRI(ar>=0 ? ar : 0);
		#pragma unroll
		for(int k=0; k<CO; k++)
			{
			if(k < (s+1)*G-1 && k<C::KS)
				{
				constexpr int dummy = 0; (void)dummy;
				const int ak = hbk_arow<C, KIND>(k);
				T.H[s][k] = (k<v && ar>=0 && ak>=0) ? row[ak>=0 ? ak : 0] : 0.0;
				}
			else T.H[s][k] = 0.0;
			}
		T.hd[s] = (v<CO && ar>=0) ? row[ar>=0 ? ar : 0] : 1.0;
		#pragma unroll
		for(int e=0; e<E; e++)
			{
			const int ae = hbk_arow<C, KIND>(RO+e);
			int ac = -1;
			if(v<CO)
				{
				if(KIND==HBF_MID) ac = v;
				else if(KIND==HBF_FIRST) ac = (v<NU) ? v : -1;
				else ac = (v<NU) ? -1 : v-NU;
				}
			T.X[e][s] = (ae>=0 && ac>=0) ? sQ[HB_TRI(ae>=0 ? ae : 0) + (ac>=0 ? ac : 0)] : 0.0;
			}
		}
	#pragma unroll
	for(int e=0; e<E; e++)
		#pragma unroll
		for(int cc=0; cc<NCC; cc++)
			{
			const int ae = hbk_arow<C, KIND>(RO+e), acol = hbk_arow<C, KIND>(CO+cc);
			double hc = (e==cc) ? 1.0 : 0.0;
			if(e>=cc && ae>=0 && acol>=0) hc = sQ[HB_TRI(ae>=0 ? ae : 0) + (acol>=0 ? acol : 0)];
			T.Z[e][cc] = hc;
			}
	}
	if(KIND==HBF_LAST) { __syncwarp(); return; }
	{
	__syncwarp();                       /* every lane has read its inputs: the buffer becomes W */
		double *sW = sBW;
		#pragma unroll
		for(int s=0; s<R; s++)
			{
			const int v = l + s*G;
			if(v<RO)
				{
				#pragma unroll
				for(int j=0; j<NX; j+=2) *reinterpret_cast<double2*>(sW + v*LDW + j) = make_double2(w[s][j], w[s][j+1]);
				}
			if(E>0 && v<NX)
				{
				#pragma unroll
				for(int e=0; e<E; e++) sW[(RO+e)*LDW + v] = wx[e][s];
				}
			}
	}
	__syncwarp();
	HBF_STAMP(111);
	/* ---- H += W W' ---- */
	const double *sW = sBW;
	#pragma unroll
	for(int s=0; s<R; s++)
		{
		double acc = T.hd[s];
		#pragma unroll
		for(int m=0; m<NX; m++) acc = fma(w[s][m], w[s][m], acc);
		T.hd[s] = acc;
		}
	#pragma unroll
	for(int m=0; m<NX; m+=2)
		{
		#pragma unroll
		for(int k=0; k<C::KS; k++)
			{
			const double2 t = *reinterpret_cast<const double2*>(sW + k*LDW + m);
			#pragma unroll
			for(int s=0; s<R; s++)
				if(k < (s+1)*G-1)
					{
					T.H[s][k] = fma(w[s][m], t.x, T.H[s][k]);
					T.H[s][k] = fma(w[s][m+1], t.y, T.H[s][k]);
					}
			}
		}
	if(E>0)
		{
		double wr[E>0 ? E : 1][NX];
		#pragma unroll
		for(int e=0; e<E; e++)
			{
			double acc[R][2];
			#pragma unroll
			for(int s=0; s<R; s++) { acc[s][0] = T.X[e][s]; acc[s][1] = 0.0; }
			#pragma unroll
			for(int m=0; m<NX; m+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(sW + (RO+e)*LDW + m);
				wr[e][m] = t.x; wr[e][m+1] = t.y;
				#pragma unroll
				for(int s=0; s<R; s++)
					{
					acc[s][0] = fma(w[s][m], t.x, acc[s][0]);
					acc[s][1] = fma(w[s][m+1], t.y, acc[s][1]);
					}
				}
			#pragma unroll
			for(int s=0; s<R; s++) T.X[e][s] = acc[s][0] + acc[s][1];
			}
		#pragma unroll
		for(int e=0; e<E; e++)
			#pragma unroll
			for(int cc=0; cc<NCC; cc++)
				if(e>=cc)
					{
					double acc = T.Z[e][cc];
					#pragma unroll
					for(int m=0; m<NX; m++) acc = fma(wr[e][m], wr[cc][m], acc);
					T.Z[e][cc] = acc;
					}
		}
	__syncwarp();                       /* W is dead: the caller may refill the buffer */
	}

/* ------------------------------------------------------------------------------------------------ */
/* backward stage, part 2: right-looking Cholesky, look-ahead diagonal; column c is finished on lane  */
/* c%G (slot c/G) and broadcast through shared memory.  Then K, k for the forward sweep.               */
/* ------------------------------------------------------------------------------------------------ */
/* Xc: where the x-columns go in shared memory ; gS: stash image of the stage in global memory ; after_u() is called once the
 * u-columns (and K, k) are done and the scratch LUs is dead */
/* GIMG = false: gS is the image buffer in shared memory that Xc points into (Xc = gS + SX): the x-columns are already there,
 * only K and k are added, and the caller moves the finished image to global memory with one bulk store */
template<class C, int KIND, class F, bool GIMG = true>
__device__ __forceinline__ void hbk_back_factor(const hbk_lane<C> &ln, hbk_tile<C> &T, double *__restrict__ LUs, double *__restrict__ Xc,
		double *__restrict__ gS, F after_u)
	{
	constexpr int NX = C::NX, NU = C::NU, G = C::G, R = C::R, RO = C::RO, E = C::E, CO = C::CO, NCC = C::NCC;
	const int l = ln.l;
	/* software pipeline: the reciprocal square root of pivot c+1 is started as soon as the look-ahead update has
	 * finished its diagonal, so its dependent chain overlaps the broadcast and trailing update of column c */
	double rs = hbk_rsqrt(T.hd[0]);
	#pragma unroll
	for(int c=0; c<CO; c++)
		{
		constexpr int dummy = 0; (void)dummy;
		const int so = c/G, lo = c%G;
		const double inv = __shfl_sync(HBF_FULL, rs, lo, G);
		double lc[R];
		#pragma unroll
		for(int s=0; s<R; s++) lc[s] = 0.0;
		#pragma unroll
		for(int s=so; s<R; s++)
			{
			const int v = l + s*G;
			lc[s] = (s==so && v==c) ? T.hd[s]*inv : T.H[s][c<C::KS ? c : 0]*inv;     /* owner: sqrt(p) ; rows > c: L[v][c] */
			if(s>so || v>c) T.hd[s] = fma(-lc[s], lc[s], T.hd[s]);                  /* look-ahead */
			}
		if(c+1<CO) rs = hbk_rsqrt(T.hd[(c+1)/G]);
		double *col = hbk_col<C>(c, LUs, Xc);
		#pragma unroll
		for(int s=so; s<R; s++)
			{
			const int v = l + s*G;
			if((s>so || v>=c) && v<RO)
				{
				col[v-c] = lc[s];
				if(GIMG && c>=NU) gS[C::SX + C::xOff(c>=NU ? c-NU : 0) + (v-c)] = lc[s];      /* x-columns go to the stash as they are finished */
				}
			}
		if(l==lo)
			{
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				const double le = T.X[e][so]*inv;
				col[RO+e-c] = le;
				if(GIMG && c>=NU) gS[C::SX + C::xOff(c>=NU ? c-NU : 0) + (RO+e-c)] = le;
				}
			if(c<NU) LUs[C::UDINV+c] = inv;
			}
		__syncwarp();
		#pragma unroll
		for(int q=0; c+2*q<C::KS; q++)
			{
			const double2 t = *reinterpret_cast<const double2*>(col + 2*q);
			const int k0 = c+2*q, k1 = k0+1;
			#pragma unroll
			for(int s=so; s<R; s++)
				{
				if(q>0 && k0<C::KS && k0 < (s+1)*G-1) T.H[s][k0] = fma(-lc[s], t.x, T.H[s][k0]);
				if(k1<C::KS && k1 < (s+1)*G-1) T.H[s][k1] = fma(-lc[s], t.y, T.H[s][k1]);
				}
			}
		if(E>0)
			{
			double le[E>0 ? E : 1];
			#pragma unroll
			for(int e=0; e<E; e++) le[e] = col[RO+e-c];
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				#pragma unroll
				for(int s=so; s<R; s++) T.X[e][s] = fma(-le[e], lc[s], T.X[e][s]);    /* meaningful for columns > c */
				#pragma unroll
				for(int cc=0; cc<NCC; cc++) if(e>=cc) T.Z[e][cc] = fma(-le[e], le[cc], T.Z[e][cc]);
				}
			}
		if(c==NU-1 && KIND!=HBF_LAST)
			{
			/* the u-columns are complete: K' rows, y Luu = -L[NU+x][0:NU]  (x = NX is the gradient row: y = k').
			 * Independent of the remaining columns, so it fills their pivot latency. */
			#pragma unroll
			for(int s=0; s<R; s++)
				{
				const int x = l + s*G;
				if(s*G<=NX)
					{
					double y[NU];
					#pragma unroll
					for(int c1=0; c1<NU; c1++) y[c1] = (x<=NX) ? -LUs[C::uOff(c1) + (NU+(x<=NX ? x : 0)-c1)] : 0.0;
					#pragma unroll
					for(int c1=NU-1; c1>=0; c1--)
						{
						y[c1] *= LUs[C::UDINV+c1];
						#pragma unroll
						for(int c2=0; c2<c1; c2++) y[c2] = fma(-y[c1], LUs[C::uOff(c2) + (c1-c2)], y[c2]);
						}
					if(x<NX)
						{
						#pragma unroll
						for(int c1=0; c1<NU; c1++) gS[C::SK + c1*NX + x] = y[c1];
						}
					else if(x==NX)
						{
						#pragma unroll
						for(int c1=0; c1<NU; c1++) gS[C::Sk + c1] = y[c1];
						}
					}
				}
			}
		if(c==NU-1) { __syncwarp(); after_u(); }
		}
	#pragma unroll
	for(int cc=0; cc<NCC; cc++)
		{
		const double p = T.Z[cc][cc];
		const double rsc = hbk_rsqrt(p);
		double *col = hbk_col<C>(CO+cc, LUs, Xc);
		double lcol[E>0 ? E : 1];
		lcol[cc] = p*rsc;
		#pragma unroll
		for(int e=cc+1; e<E; e++) lcol[e] = T.Z[e][cc]*rsc;
		if(l==0)
			{
			#pragma unroll
			for(int e=cc; e<E; e++)
				{
				col[e-cc] = lcol[e];
				if(GIMG && CO+cc>=NU) gS[C::SX + C::xOff(CO+cc>=NU ? CO+cc-NU : 0) + (e-cc)] = lcol[e];
				}
			if(CO+cc<NU) LUs[C::UDINV+CO+cc] = rsc;
			}
		#pragma unroll
		for(int e=cc+1; e<E; e++)
			#pragma unroll
			for(int c2=cc+1; c2<NCC; c2++)
				if(e>=c2) T.Z[e][c2] = fma(-lcol[e], lcol[c2], T.Z[e][c2]);
		}
	__syncwarp();
	HBF_STAMP(120);
	}

/* ------------------------------------------------------------------------------------------------ */
/* forward stage n: u_n = k + K x_n ; x_{n+1} = b + B u + A x ; pi_{n-1} = Lxx (Lxx' x_n + l_x)       */
/* ------------------------------------------------------------------------------------------------ */
template<class C, int KIND>
__device__ __forceinline__ void hbk_stage_forward(const hbk_lane<C> &ln, const double *__restrict__ sB, const double *__restrict__ Sn,
		double *__restrict__ us, const double *__restrict__ xs, double *__restrict__ xo, double *__restrict__ tmp,
		double *__restrict__ g_u, double *__restrict__ g_x1, double *__restrict__ g_pi, bool active)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, G = C::G, R = C::R;
	const int l = ln.l;
	const double *xc = Sn + C::SX;
	/* ---- phase A ---- */
	#pragma unroll
	for(int s=0; s<R; s++)
		{
		const int i = l + s*G;
		if(s*G<NU && i<NU)
			{
			double t0 = Sn[C::Sk + i], t1 = 0.0;
			if(KIND!=HBF_FIRST)
				{
				const double *Kr = Sn + C::SK + i*NX;
				#pragma unroll
				for(int k=0; k<NX; k+=2)
					{
					const double2 kk = *reinterpret_cast<const double2*>(Kr + k);
					const double2 xx = *reinterpret_cast<const double2*>(xs + k);
					t0 = fma(kk.x, xx.x, t0); t1 = fma(kk.y, xx.y, t1);
					}
				}
			const double u = t0 + t1;
			us[i] = u;
			if(active) g_u[i] = u;
			}
		if(KIND!=HBF_FIRST && s*G<NX && i<NX)
			{
			const double *col = xc + ln.xo[s] - i;                          /* col[k] = Lxx[k][i] */
			double a0 = col[NX], a1 = 0.0;
			#pragma unroll
			for(int k=s*G; k<NX; k+=2)
				{
				if(k>=i) a0 = fma(col[k], xs[k], a0);
				if(k+1>=i && k+1<NX) a1 = fma(col[k+1], xs[k+1], a1);
				}
			tmp[i] = a0 + a1;
			}
		}
	__syncwarp();
	/* ---- phase B ---- */
	#pragma unroll
	for(int s=0; s<R; s++)
		{
		const int j = l + s*G;
		if(s*G<NX && j<NX)
			{
			constexpr int brow = (KIND==HBF_FIRST) ? NU : NUX;
			double x0 = sB[brow*NX + j], x1 = 0.0, x2 = 0.0;
			#pragma unroll
			for(int i=0; i<NU; i++) x0 = fma(sB[i*NX+j], us[i], x0);
			if(KIND!=HBF_FIRST)
				{
				#pragma unroll
				for(int i=0; i<NX; i+=2) { x1 = fma(sB[(NU+i)*NX+j], xs[i], x1); x2 = fma(sB[(NU+i+1)*NX+j], xs[i+1], x2); }
				}
			const double xn = x0 + (x1+x2);
			xo[j] = xn;
			if(active) g_x1[j] = xn;
			if(KIND!=HBF_FIRST)
				{
				double p0 = 0.0, p1 = 0.0;
				#pragma unroll
				for(int cc=0; cc<NX; cc+=2)
					{
					if(cc<(s+1)*G && cc<=j) p0 = fma(xc[C::xOff(cc) + (j-cc)], tmp[cc], p0);
					if(cc+1<(s+1)*G && cc+1<=j) p1 = fma(xc[C::xOff(cc+1) + (j-cc-1)], tmp[cc+1], p1);
					}
				if(active) g_pi[j] = p0+p1;
				}
			}
		}
	__syncwarp();
	}

/* pi_{N-1} from x_N and the x-columns of L_N */
template<class C>
__device__ __forceinline__ void hbk_final_pi(const hbk_lane<C> &ln, const double *__restrict__ Sn, const double *__restrict__ xs,
		double *__restrict__ tmp, double *__restrict__ g_pi, bool active)
	{
	constexpr int NX = C::NX, G = C::G, R = C::R;
	const int l = ln.l;
	const double *xc = Sn + C::SX;
	#pragma unroll
	for(int s=0; s<R; s++)
		{
		const int i = l + s*G;
		if(s*G<NX && i<NX)
			{
			const double *col = xc + ln.xo[s] - i;
			double a0 = col[NX];
			#pragma unroll
			for(int k=s*G; k<NX; k++) if(k>=i) a0 = fma(col[k], xs[k], a0);
			tmp[i] = a0;
			}
		}
	__syncwarp();
	#pragma unroll
	for(int s=0; s<R; s++)
		{
		const int j = l + s*G;
		if(s*G<NX && j<NX)
			{
			double p0 = 0.0;
			#pragma unroll
			for(int cc=0; cc<NX; cc++) if(cc<(s+1)*G && cc<=j) p0 = fma(xc[C::xOff(cc) + (j-cc)], tmp[cc], p0);
			if(active) g_pi[j] = p0;
			}
		}
	__syncwarp();
	}

/* ------------------------------------------------------------------------------------------------ */
/* kernel: persistent warps, IPW instances per warp                                                  */
/* ------------------------------------------------------------------------------------------------ */
/* L2 management of the factor stash (flags, set by the launcher):
 *   HBK_F_IN_FIRST    stage inputs of the backward sweep are fetched evict-first (single use: RSQrq; [B A b]' is re-read a whole
 *                     sweep later, by which time it has left L2 anyway)
 *   HBK_F_RE_FIRST    the forward sweep's re-read of [B A b]' is fetched evict-first (dead afterwards)
 *   HBK_F_KEEP        (bulk-store build) images of stages n < keep_n0 are written evict-last, the others evict-first: the images
 *                     with the shortest write-to-read distance are the ones worth keeping on chip
 *   HBK_F_DISCARD     an image is dropped from L2 (discard.global.L2, no write-back) as soon as the forward sweep has fetched it:
 *                     the slot is rewritten by the warp's next instance, its old content is dead
 * HBK_BULK_STASH (compile time): the image of a stage is completed in shared memory (K, k next to the x-columns) and moved to
 * the stash with ONE bulk store per instance and stage instead of 162 STG.64; images 0 and 1 never leave shared memory. */
#define HBK_F_IN_FIRST 1
#define HBK_F_RE_FIRST 2
#define HBK_F_KEEP 4
#define HBK_F_DISCARD 8
template<class C>
__global__ void __launch_bounds__(256, 1) hbk_ric_sv_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ stash, int flags, int keep_n0)
	{
	constexpr int G = C::G, IPW = C::IPW, NX = C::NX, NU = C::NU, NUX = C::NUX, SB = C::SB, SBG = C::SBG, IOB = C::IOB, BAB = C::BAB, LU = C::LU;
	extern __shared__ __align__(16) double hbf_smem[];
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const int g = lane/G;
	hbk_lane<C> ln; ln.init(lane%G);
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	double *wbase = hbf_smem + (size_t)warp*C::PER_WARP;
	uint64_t *bars = reinterpret_cast<uint64_t*>(wbase);      /* [0] stage inputs / [B A b]' slot 0, [1] [B A b]' slot 1, [2..3] stash images */
	double *ibase = wbase + 8 + C::inst_off(g);
	double *io = ibase;
	double *LUs = ibase + IOB;
	double *S0 = LUs + LU, *S1 = S0 + SB;
	double *us = S1 + SB, *xs0 = us + C::even(NU), *xs1 = xs0 + C::XS, *tmp = xs1 + C::XS;
	if(lane==0)
		{
		for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	__syncwarp();
	uint32_t phase = 0;
	const int N = d.N;
	const int o_in1 = d.st[1].off_BAbt, s_in = d.st[2].off_BAbt - d.st[1].off_BAbt, o_inN = d.st[N].off_BAbt;
	constexpr uint32_t bytes_first = 8u*(uint32_t)(C::even((NU+1)*NX) + C::even(HB_TRI(NU)+NU));
	constexpr uint32_t bytes_mid = 8u*(uint32_t)C::INB;
	constexpr uint32_t bytes_last = 8u*(uint32_t)C::even(HB_TRI(NX)+NX);
	const long long stash_stride = (long long)(N+1)*SBG;
	const long long n_groups = (n_inst + IPW - 1)/IPW;
	double *stash_w = stash + gw*IPW*stash_stride;
	const uint64_t pol_first = hbf_policy_evict_first();
#ifdef HBK_BULK_STASH
	const uint64_t pol_last = hbf_policy_evict_last(), pol_norm = hbf_policy_evict_normal();
#endif
	const bool in_first = (flags&HBK_F_IN_FIRST)!=0, re_first = (flags&HBK_F_RE_FIRST)!=0, discard = (flags&HBK_F_DISCARD)!=0;

	for(long long grp=gw; grp<n_groups; grp+=tw)
		{
		long long inst = grp*IPW + g;
		const bool active = inst<n_inst;
		if(!active) inst = n_inst-1;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;

		/* lanes 0..IPW-1 each move the data of one instance of the warp (one bulk copy per lane and buffer);
		 * lane 0 posts the byte count, one mbarrier per buffer */
		const int mg = lane<IPW ? lane : 0;
		long long my_i = grp*IPW + mg; if(my_i>=n_inst) my_i = n_inst-1;
		const double *my_in = in + my_i*d.in_stride;
		double *my_sm = wbase + 8 + C::inst_off(mg);
		double *my_st = stash_w + mg*stash_stride;
		auto issue_backward = [&](int n)                          /* [B A b]'_n | RSQrq_n -> io */
			{
			const int off = (n==0) ? 0 : (n==N ? o_inN : o_in1 + (n-1)*s_in);
			const uint32_t bytes = (n==0) ? bytes_first : (n==N ? bytes_last : bytes_mid);
			if(lane==0) hbf_mbar_expect(&bars[0], bytes*IPW);
			if(lane<IPW)
				{
				if(in_first) hbf_bulk_g2s_hint(my_sm, my_in + off, bytes, &bars[0], pol_first);
				else hbf_bulk_g2s(my_sm, my_in + off, bytes, &bars[0]);
				}
			};
		auto issue_BAbt = [&](int n, int slot)                    /* forward: [B A b]'_n -> io + slot*BAB */
			{
			const int off = (n==0) ? 0 : o_in1 + (n-1)*s_in;
			const uint32_t bytes = (n==0) ? 8u*(uint32_t)C::even((NU+1)*NX) : 8u*(uint32_t)BAB;
			if(lane==0) hbf_mbar_expect(&bars[slot], bytes*IPW);
			if(lane<IPW)
				{
				if(re_first) hbf_bulk_g2s_hint(my_sm + slot*BAB, my_in + off, bytes, &bars[slot], pol_first);
				else hbf_bulk_g2s(my_sm + slot*BAB, my_in + off, bytes, &bars[slot]);
				}
			};
		auto issue_S = [&](int n, int slot)                       /* forward: stash image of stage n -> S[slot] */
			{
			if(lane==0) hbf_mbar_expect(&bars[2+slot], 8u*SB*IPW);
			if(lane<IPW) hbf_bulk_g2s(my_sm + IOB + LU + slot*SB, my_st + (long long)n*SBG, 8u*SB, &bars[2+slot]);
			};
		auto wait_bar = [&](int b) { hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b); };
		double *gst = stash_w + g*stash_stride;                 /* this instance's stash slot */
#ifdef HBK_BULK_STASH
		/* the finished image of stage n (in S[n&1]) -> stash, one bulk store per instance; the buffer may be rewritten once the
		 * store has read it (wait_img_read before the factorisation of stage n-2) */
		auto store_img = [&](int n)
			{
			hbf_fence_async();                                    /* generic-proxy writes of the image -> visible to the bulk copy */
			__syncwarp();
			if(lane<IPW)
				{
				const uint64_t pol = (flags&HBK_F_KEEP) ? (n<keep_n0 ? pol_last : pol_first) : pol_norm;
				hbf_bulk_s2g_hint(my_st + (long long)n*SBG, my_sm + IOB + LU + (n&1)*SB, 8u*SB, pol);
				hbf_bulk_commit();
				}
			};
		auto wait_img_read = [&]() { if(lane<IPW) hbf_bulk_wait_read<1>(); __syncwarp(); };
#endif

		/* ---------------- backward sweep: stage n builds S[n&1], reads the x-columns in S[(n+1)&1] ---------------- */
		issue_backward(N);
		{
		hbk_tile<C> T;
		wait_bar(0);
		hbk_back_assemble<C, HBF_LAST, C::LDW>(ln, io, io, nullptr, T, hbk_nop());
		issue_backward(N-1);
#ifdef HBK_BULK_STASH
		hbk_back_factor<C, HBF_LAST, hbk_nop, false>(ln, T, LUs, ((N&1) ? S1 : S0) + C::SX, (N&1) ? S1 : S0, hbk_nop());
		store_img(N);
#else
		hbk_back_factor<C, HBF_LAST>(ln, T, LUs, ((N&1) ? S1 : S0) + C::SX, gst + (long long)N*SBG, hbk_nop());
#endif
		}
		for(int n=N-1; n>0; n--)
			{
			double *Sc = (n&1) ? S1 : S0;
			const double *Sp = (n&1) ? S0 : S1;
			hbk_tile<C> T;
			HBF_STAMP(100);
			wait_bar(0);
			HBF_STAMP(101);
			hbk_back_assemble<C, HBF_MID, C::LDW>(ln, io, io + BAB, Sp + C::SX, T, hbk_nop());
			HBF_STAMP(102);
			issue_backward(n-1);                                  /* lands while the factorization runs */
			HBF_STAMP(103);
#ifdef HBK_BULK_STASH
			wait_img_read();                                      /* the store of image n+2 has left this buffer */
			hbk_back_factor<C, HBF_MID, hbk_nop, false>(ln, T, LUs, Sc + C::SX, Sc, hbk_nop());
			if(n>=2) store_img(n);                                /* images 0 and 1 stay in shared memory for the forward sweep */
#else
			hbk_back_factor<C, HBF_MID>(ln, T, LUs, Sc + C::SX, gst + (long long)n*SBG, hbk_nop());
#endif
			HBF_STAMP(104);
			}
		{
		hbk_tile<C> T;
		wait_bar(0);
		hbk_back_assemble<C, HBF_FIRST, C::LDW>(ln, io, io + C::even((NU+1)*NX), S1 + C::SX, T, hbk_nop());
#ifdef HBK_BULK_STASH
		wait_img_read();
		hbk_back_factor<C, HBF_FIRST, hbk_nop, false>(ln, T, LUs, S0 + C::SX, S0, hbk_nop());
		if(lane<IPW) hbf_bulk_wait_all<0>();                      /* every image is in the stash before any is fetched back */
		__syncwarp();
#else
		hbk_back_factor<C, HBF_FIRST>(ln, T, LUs, S0 + C::SX, gst, hbk_nop());
#endif
		}
#ifndef HBK_BULK_STASH
		/* the images were written through the generic proxy and are read back by bulk copies (async proxy) */
		asm volatile("fence.proxy.async;" ::: "memory");
		__syncwarp();
#endif

		/* an image that has been fetched is dead: drop its lines from L2 before they are written back (the slot is rewritten by
		 * the next instance of this warp).  Needs whole 128-byte lines per image (SBG). */
		auto drop_img = [&](int n)
			{
			if(discard && (SBG%16)==0)
				{
				const double *img = gst + (long long)n*SBG;
				#pragma unroll
				for(int q=ln.l; q<SBG/16; q+=G) hbf_discard_line(img + 16*q);
				}
			};

		/* ---------------- forward sweep ---------------- */
#ifdef HBK_BULK_STASH
		issue_BAbt(0, 0);
		if(N>1) issue_BAbt(1, 1);
		wait_bar(0);
#else
		issue_S(0, 0);
		issue_BAbt(0, 0);
		issue_S(1, 1);
		if(N>1) issue_BAbt(1, 1);
		wait_bar(2);
		wait_bar(0);
		drop_img(0);
#endif
		hbk_stage_forward<C, HBF_FIRST>(ln, io, S0, us, xs0, xs1, tmp, ux, ux + NU + ((1<N) ? NU : 0), pi, active);
		if(2<=N) issue_S(2, 0);
		if(2<N) issue_BAbt(2, 0);
		for(int n=1; n<N; n++)
			{
			const double *Sn = (n&1) ? S1 : S0;
			const double *xs = (n&1) ? xs1 : xs0;
			double *xo = (n&1) ? xs0 : xs1;
			HBF_STAMP(200);
#ifdef HBK_BULK_STASH
			if(n>=2) { wait_bar(2+(n&1)); drop_img(n); }          /* image 1 is still in shared memory */
#else
			wait_bar(2+(n&1));                                    /* image n */
			drop_img(n);
#endif
			wait_bar(n&1);                                        /* [B A b]'_n */
			HBF_STAMP(201);
			const int o_ux = NU + (n-1)*NUX, o_ux1 = NU + n*NUX + ((n+1<N) ? NU : 0);
			hbk_stage_forward<C, HBF_MID>(ln, io + (n&1)*BAB, Sn, us, xs, xo, tmp, ux + o_ux, ux + o_ux1, pi + (n-1)*NX, active);
			HBF_STAMP(202);
			if(n+2<=N) issue_S(n+2, n&1);
			if(n+2<N) issue_BAbt(n+2, n&1);
			}
#ifdef HBK_BULK_STASH
		if(N>=2) { wait_bar(2+(N&1)); drop_img(N); }
#else
		wait_bar(2+(N&1));
		drop_img(N);
#endif
		hbk_final_pi<C>(ln, (N&1) ? S1 : S0, (N&1) ? xs1 : xs0, tmp, pi + (N-1)*NX, active);
		}
	}

#ifdef HBK_BULK_STASH
/* ------------------------------------------------------------------------------------------------ */
/* Traffic-equivalent probe of hbk_ric_sv_kernel: the same persistent loop, the same bulk copies in the  */
/* same order with the same L2 hints (stage inputs, stash images out and back, [B A b]' re-read), the   */
/* same output stores -- and NO arithmetic.  Its rate is the ceiling the memory system sets for this     */
/* two-sweep access pattern at this occupancy; the sv kernel's distance from it is what the stage         */
/* arithmetic (shared-memory pipe, pivot chains) costs.  Measurement tool, not part of the solver.        */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__global__ void __launch_bounds__(256, 1) hbk_traffic_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ stash, int flags, int keep_n0)
	{
	constexpr int G = C::G, IPW = C::IPW, NX = C::NX, NU = C::NU, NUX = C::NUX, SB = C::SB, SBG = C::SBG, IOB = C::IOB, BAB = C::BAB, LU = C::LU;
	extern __shared__ __align__(16) double hbf_smem[];
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const int g = lane/G, l = lane%G;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	double *wbase = hbf_smem + (size_t)warp*C::PER_WARP;
	uint64_t *bars = reinterpret_cast<uint64_t*>(wbase);
	double *ibase = wbase + 8 + C::inst_off(g);
	double *S0 = ibase + IOB + LU;
	if(lane==0)
		{
		for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	__syncwarp();
	uint32_t phase = 0;
	const int N = d.N;
	const int o_in1 = d.st[1].off_BAbt, s_in = d.st[2].off_BAbt - d.st[1].off_BAbt, o_inN = d.st[N].off_BAbt;
	constexpr uint32_t bytes_first = 8u*(uint32_t)(C::even((NU+1)*NX) + C::even(HB_TRI(NU)+NU));
	constexpr uint32_t bytes_mid = 8u*(uint32_t)C::INB;
	constexpr uint32_t bytes_last = 8u*(uint32_t)C::even(HB_TRI(NX)+NX);
	const long long stash_stride = (long long)(N+1)*SBG;
	const long long n_groups = (n_inst + IPW - 1)/IPW;
	double *stash_w = stash + gw*IPW*stash_stride;
	const uint64_t pol_first = hbf_policy_evict_first(), pol_last = hbf_policy_evict_last(), pol_norm = hbf_policy_evict_normal();
	const bool in_first = (flags&HBK_F_IN_FIRST)!=0, re_first = (flags&HBK_F_RE_FIRST)!=0, discard = (flags&HBK_F_DISCARD)!=0;
	for(long long grp=gw; grp<n_groups; grp+=tw)
		{
		long long inst = grp*IPW + g;
		const bool active = inst<n_inst;
		if(!active) inst = n_inst-1;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		const int mg = lane<IPW ? lane : 0;
		long long my_i = grp*IPW + mg; if(my_i>=n_inst) my_i = n_inst-1;
		const double *my_in = in + my_i*d.in_stride;
		double *my_sm = wbase + 8 + C::inst_off(mg);
		double *my_st = stash_w + mg*stash_stride;
		double *gst = stash_w + g*stash_stride;
		auto issue_backward = [&](int n)
			{
			const int off = (n==0) ? 0 : (n==N ? o_inN : o_in1 + (n-1)*s_in);
			const uint32_t bytes = (n==0) ? bytes_first : (n==N ? bytes_last : bytes_mid);
			if(lane==0) hbf_mbar_expect(&bars[0], bytes*IPW);
			if(lane<IPW) { if(in_first) hbf_bulk_g2s_hint(my_sm, my_in + off, bytes, &bars[0], pol_first); else hbf_bulk_g2s(my_sm, my_in + off, bytes, &bars[0]); }
			};
		auto issue_BAbt = [&](int n, int slot)
			{
			const int off = (n==0) ? 0 : o_in1 + (n-1)*s_in;
			const uint32_t bytes = (n==0) ? 8u*(uint32_t)C::even((NU+1)*NX) : 8u*(uint32_t)BAB;
			if(lane==0) hbf_mbar_expect(&bars[slot], bytes*IPW);
			if(lane<IPW) { if(re_first) hbf_bulk_g2s_hint(my_sm + slot*BAB, my_in + off, bytes, &bars[slot], pol_first); else hbf_bulk_g2s(my_sm + slot*BAB, my_in + off, bytes, &bars[slot]); }
			};
		auto issue_S = [&](int n, int slot)
			{
			if(lane==0) hbf_mbar_expect(&bars[2+slot], 8u*SB*IPW);
			if(lane<IPW) hbf_bulk_g2s(my_sm + IOB + LU + slot*SB, my_st + (long long)n*SBG, 8u*SB, &bars[2+slot]);
			};
		auto wait_bar = [&](int b) { hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b); };
		auto store_img = [&](int n)
			{
			hbf_fence_async();
			__syncwarp();
			if(lane<IPW)
				{
				const uint64_t pol = (flags&HBK_F_KEEP) ? (n<keep_n0 ? pol_last : pol_first) : pol_norm;
				hbf_bulk_s2g_hint(my_st + (long long)n*SBG, my_sm + IOB + LU + (n&1)*SB, 8u*SB, pol);
				hbf_bulk_commit();
				}
			};
		auto drop_img = [&](int n)
			{
			if(discard && (SBG%16)==0)
				for(int q=l; q<SBG/16; q+=G) hbf_discard_line(gst + (long long)n*SBG + 16*q);
			};
		/* the outputs of one stage: u (NU), x (NX), pi (NX), written by the lanes of the instance like the solver does */
		auto write_out = [&](int n)
			{
			if(!active) return;
			const double v = S0[l];                               /* something that depends on the fetched data */
			const int o_ux = (n==0) ? 0 : NU + (n-1)*NUX;
			for(int i=l; i<NUX; i+=G) ux[o_ux + i] = v;
			if(n>0) for(int i=l; i<NX; i+=G) pi[(n-1)*NX + i] = v;
			};
		issue_backward(N);
		for(int n=N; n>=0; n--)
			{
			wait_bar(0);
			__syncwarp();
			if(n>0) issue_backward(n-1);
			if(lane<IPW) hbf_bulk_wait_read<1>();
			__syncwarp();
			if(n>=2) store_img(n);
			}
		if(lane<IPW) hbf_bulk_wait_all<0>();
		__syncwarp();
		issue_BAbt(0, 0);
		if(N>1) issue_BAbt(1, 1);
		wait_bar(0);
		write_out(0);
		__syncwarp();
		if(2<=N) issue_S(2, 0);
		if(2<N) issue_BAbt(2, 0);
		for(int n=1; n<N; n++)
			{
			if(n>=2) { wait_bar(2+(n&1)); drop_img(n); }
			wait_bar(n&1);
			write_out(n);
			__syncwarp();
			if(n+2<=N) issue_S(n+2, n&1);
			if(n+2<N) issue_BAbt(n+2, n&1);
			}
		if(N>=2) { wait_bar(2+(N&1)); drop_img(N); }
		write_out(N);
		__syncwarp();
		}
	}
#endif

/* ------------------------------------------------------------------------------------------------ */
/* tails of a scenario tree (ric_tree.cuh): the same register-blocked stages on chains whose first node */
/* has a given state (it comes from the top of the tree) and whose data sit in the tree's node-indexed  */
/* layout.  mode 0: backward over the tail, images -> the nodes' factor slots, and the root's Lxx, l_x   */
/* also in the generic packed form the top kernel reads; mode 1: forward from the root's x.              */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__global__ void __launch_bounds__(256, 1) hbk_tail_kernel(hb_tail_tab tab, long long n_trees, long long in_stride, long long ux_stride,
		long long pi_stride, long long L_stride, const double *__restrict__ in, double *__restrict__ ux_all, double *__restrict__ pi_all,
		double *__restrict__ L_all, int mode, int tail_lo, int tail_hi, const double *__restrict__ skip)
	{
	constexpr int G = C::G, IPW = C::IPW, NX = C::NX, NU = C::NU, SB = C::SB, IOB = C::IOB, BAB = C::BAB, LU = C::LU;
	extern __shared__ __align__(16) double hbf_smem[];
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const int g = lane/G;
	hbk_lane<C> ln; ln.init(lane%G);
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	double *wbase = hbf_smem + (size_t)warp*C::PER_WARP;
	uint64_t *bars = reinterpret_cast<uint64_t*>(wbase);
	double *ibase = wbase + 8 + C::inst_off(g);
	double *io = ibase;
	double *LUs = ibase + IOB;
	double *S0 = LUs + LU, *S1 = S0 + SB;
	double *us = S1 + SB, *xs0 = us + C::even(NU), *xs1 = xs0 + C::XS, *tmp = xs1 + C::XS;
	if(lane==0)
		{
		for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	__syncwarp();
	uint32_t phase = 0;
	if(skip!=nullptr && skip[n_trees*8]==0.0) return;          /* tree IPM driver: every tree has finished */
	const int len = tab.len, ntl = tail_hi - tail_lo;
	const long long n_items = n_trees*ntl, n_groups = (n_items + IPW - 1)/IPW;
	constexpr uint32_t bytes_Q = 8u*(uint32_t)C::RSQ, bytes_Qlast = 8u*(uint32_t)C::even(HB_TRI(NX)+NX), bytes_B = 8u*(uint32_t)BAB;

	for(long long grp=gw; grp<n_groups; grp+=tw)
		{
		long long item = grp*IPW + g;
		const bool active = item<n_items;
		if(!active) item = n_items-1;
		const long long t = item/ntl;
		/* tree IPM driver: trees that have finished are skipped (their state record says so); a group is skipped as a whole */
		if(skip!=nullptr && __all_sync(HBF_FULL, (int)skip[t*8+3]==HBK_TS_DONE)) continue;
		const int j = tail_lo + (int)(item - t*ntl);
		double *ux = ux_all + t*ux_stride, *pi = pi_all + t*pi_stride, *Lt = L_all + t*L_stride;
		const int mg = lane<IPW ? lane : 0;
		long long my_item = grp*IPW + mg; if(my_item>=n_items) my_item = n_items-1;
		const long long my_t = my_item/ntl;
		const int my_j = tail_lo + (int)(my_item - my_t*ntl);
		const double *my_in = in + my_t*in_stride;
		const double *my_L = L_all + my_t*L_stride;
		double *my_sm = wbase + 8 + C::inst_off(mg);
		auto wait_bar = [&](int b) { hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b); };

		if(mode==0)
			{
			/* stage m needs RSQrq of node m and [B A b]' of the edge m -> m+1, which the tree layout keeps with node m+1 */
			auto issue = [&](int m)
				{
				const bool last = (m==len-1);
				if(lane==0) hbf_mbar_expect(&bars[0], (last ? bytes_Qlast : bytes_B + bytes_Q)*IPW);
				if(lane<IPW)
					{
					if(!last) hbf_bulk_g2s(my_sm, my_in + tab.posB[m+1] + my_j*tab.strB[m+1], bytes_B, &bars[0]);
					hbf_bulk_g2s(my_sm + (last ? 0 : BAB), my_in + tab.posQ[m] + my_j*tab.strB[m], last ? bytes_Qlast : bytes_Q, &bars[0]);
					}
				};
			issue(len-1);
			{
			hbk_tile<C> T;
			wait_bar(0);
			hbk_back_assemble<C, HBF_LAST, C::LDW>(ln, io, io, nullptr, T, hbk_nop());
			issue(len-2);
			hbk_back_factor<C, HBF_LAST>(ln, T, LUs, (((len-1)&1) ? S1 : S0) + C::SX, Lt + tab.posI[len-1] + j*tab.strL[len-1], hbk_nop());
			}
			for(int m=len-2; m>=0; m--)
				{
				double *Sc = (m&1) ? S1 : S0;
				const double *Sp = (m&1) ? S0 : S1;
				hbk_tile<C> T;
				wait_bar(0);
				hbk_back_assemble<C, HBF_MID, C::LDW>(ln, io, io + BAB, Sp + C::SX, T, hbk_nop());
				if(m>0) issue(m-1);
				hbk_back_factor<C, HBF_MID>(ln, T, LUs, Sc + C::SX, Lt + tab.posI[m] + j*tab.strL[m], hbk_nop());
				}
			/* the root's Lxx and l_x in the generic packed form (row r = NU+x: element (r, NU+c) at r(r+1)/2 + NU + c) for the top */
			if(active)
				{
				double *gen = Lt + tab.posL[0] + j*tab.strL[0];
				const double *xc = S0 + C::SX;
				#pragma unroll
				for(int s=0; s<C::R; s++)
					{
					const int c = ln.l + s*G;
					if(s*G<NX && c<NX)
						{
						const double *col = xc + ln.xo[s] - c;                   /* col[x] = Lxx[x][c], col[NX] = l_x[c] */
						for(int x=c; x<=NX; x++) gen[HB_TRI(NU+x) + NU + c] = col[x];
						}
					}
				}
			__syncwarp();
			}
		else
			{
			auto issue_B = [&](int m, int slot)                   /* [B A b]' of the edge m -> m+1 */
				{
				if(lane==0) hbf_mbar_expect(&bars[slot], bytes_B*IPW);
				if(lane<IPW) hbf_bulk_g2s(my_sm + slot*BAB, my_in + tab.posB[m+1] + my_j*tab.strB[m+1], bytes_B, &bars[slot]);
				};
			auto issue_S = [&](int m, int slot)
				{
				if(lane==0) hbf_mbar_expect(&bars[2+slot], 8u*SB*IPW);
				if(lane<IPW) hbf_bulk_g2s(my_sm + IOB + LU + slot*SB, my_L + tab.posI[m] + my_j*tab.strL[m], 8u*SB, &bars[2+slot]);
				};
			asm volatile("fence.proxy.async;" ::: "memory");
			issue_S(0, 0); issue_B(0, 0);
			issue_S(1, 1); if(len>2) issue_B(1, 1);
			if(ln.l<NX) xs0[ln.l] = ux[tab.posU[0] + j*tab.strU[0] + NU + ln.l];
			if(C::R>1 && ln.l+G<NX) xs0[ln.l+G] = ux[tab.posU[0] + j*tab.strU[0] + NU + ln.l + G];
			__syncwarp();
			for(int m=0; m<len-1; m++)
				{
				const double *Sn = (m&1) ? S1 : S0;
				const double *xs = (m&1) ? xs1 : xs0;
				double *xo = (m&1) ? xs0 : xs1;
				wait_bar(2+(m&1));
				wait_bar(m&1);
				hbk_stage_forward<C, HBF_MID>(ln, io + (m&1)*BAB, Sn, us, xs, xo, tmp, ux + tab.posU[m] + j*tab.strU[m],
						ux + tab.posU[m+1] + j*tab.strU[m+1] + ((m+1<len-1) ? NU : 0), pi + tab.posP[m] + j*tab.strP[m], active);
				if(m+2<=len-1) issue_S(m+2, m&1);
				if(m+2<len-1) issue_B(m+2, m&1);
				}
			wait_bar(2+((len-1)&1));
			hbk_final_pi<C>(ln, ((len-1)&1) ? S1 : S0, ((len-1)&1) ? xs1 : xs0, tmp, pi + tab.posP[len-1] + j*tab.strP[len-1], active);
			}
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* top of a scenario tree: one (tree, node) work item per G lanes, all nodes of one level per launch.     */
/* mode 0: H = RSQrq + sum over kids W_k W_k' (d_tree_back_ric_rec_libstr.c:79-156), factor, image -> the */
/* node's slot; mode 1: u from the node's image, x of every kid (:204-260); pi of a node is computed when   */
/* the node itself is visited, from its own image (as in the chain kernels).                              */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__global__ void __launch_bounds__(256, 1) hbk_top_kernel(hb_tdims d, long long n_trees, const double *__restrict__ in,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ L_all, int mode, int seg_lo, int seg_hi, int first,
		const double *__restrict__ skip)
	{
	constexpr int G = C::G, IPW = C::IPW, NX = C::NX, NU = C::NU, NUX = C::NUX, SB = C::SB, IOB = C::IOB, BAB = C::BAB, LU = C::LU;
	constexpr int GEN = HB_EVEN(HB_TRI(NUX)+2*NUX), GEN0 = HB_EVEN(HB_TRI(NU)+2*NU), XC = C::xOff(NX);
	extern __shared__ __align__(16) double hbf_smem[];
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const int g = lane/G;
	hbk_lane<C> ln; ln.init(lane%G);
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	double *wbase = hbf_smem + (size_t)warp*C::PER_WARP;
	uint64_t *bars = reinterpret_cast<uint64_t*>(wbase);      /* [0] [B A b]' of a kid, [1] RSQrq / own image, [2] x-columns of a kid */
	double *ibase = wbase + 8 + C::inst_off(g);
	double *io = ibase, *LUs = ibase + IOB, *S0 = LUs + LU, *S1 = S0 + SB;
	double *us = S1 + SB, *xs0 = us + C::even(NU), *xs1 = xs0 + C::XS, *tmp = xs1 + C::XS;
	if(lane==0)
		{
		for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	__syncwarp();
	uint32_t phase = 0;
	if(skip!=nullptr && skip[n_trees*8]==0.0) return;
	const int nseg = seg_hi - seg_lo;
	const long long n_items = n_trees*nseg, n_groups = (n_items + IPW - 1)/IPW;
	const uint32_t bytes_B = 8u*(uint32_t)(first ? C::even((NU+1)*NX) : BAB);
	const uint32_t bytes_Q = 8u*(uint32_t)(first ? C::even(HB_TRI(NU)+NU) : C::RSQ);
	const int gen_own = first ? GEN0 : GEN;
	asm volatile("fence.proxy.async;" ::: "memory");

	for(long long grp=gw; grp<n_groups; grp+=tw)
		{
		long long item = grp*IPW + g;
		const bool active = item<n_items;
		if(!active) item = n_items-1;
		const long long t = item/nseg;
		if(skip!=nullptr && __all_sync(HBF_FULL, (int)skip[t*8+3]==HBK_TS_DONE)) continue;
		const hb_tnode nd = d.tn[d.seg_nodes[d.seg_start[seg_lo + (int)(item - t*nseg)]]];
		double *ux = ux_all + t*d.ux_stride, *pi = pi_all + t*d.pi_stride, *Lt = L_all + t*d.L_stride;
		const int mg = lane<IPW ? lane : 0;
		long long my_item = grp*IPW + mg; if(my_item>=n_items) my_item = n_items-1;
		const long long my_t = my_item/nseg;
		const hb_tnode my_nd = d.tn[d.seg_nodes[d.seg_start[seg_lo + (int)(my_item - my_t*nseg)]]];
		const double *my_in = in + my_t*d.in_stride;
		const double *my_L = L_all + my_t*d.L_stride;
		double *my_sm = wbase + 8 + C::inst_off(mg);
		auto wait_bar = [&](int b) { hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b); };
		const int nkids = nd.nkids;                              /* the same for every node of a level */

		if(mode==0)
			{
			hbk_tile<C> T;
			if(lane==0) hbf_mbar_expect(&bars[1], bytes_Q*IPW);
			if(lane<IPW) hbf_bulk_g2s(my_sm + BAB, my_in + my_nd.off_RSQ, bytes_Q, &bars[1]);
			for(int kc=0; kc<nkids; kc++)
				{
				if(lane==0) { hbf_mbar_expect(&bars[0], bytes_B*IPW); hbf_mbar_expect(&bars[2], 8u*XC*IPW); }
				if(lane<IPW)
					{
					const hb_tnode kd = d.tn[my_nd.first_kid+kc];
					hbf_bulk_g2s(my_sm, my_in + kd.off_BAbt, bytes_B, &bars[0]);
					hbf_bulk_g2s(my_sm + IOB + LU + SB + C::SX, my_L + kd.off_L + GEN + C::SX, 8u*XC, &bars[2]);
					}
				wait_bar(0); wait_bar(2);
				if(kc==0)
					{
					if(first) hbk_back_assemble<C, HBF_FIRST, C::LDW>(ln, io, io + BAB, S1 + C::SX, T, [&]() { wait_bar(1); });
					else      hbk_back_assemble<C, HBF_MID, C::LDW>(ln, io, io + BAB, S1 + C::SX, T, [&]() { wait_bar(1); });
					}
				else
					{
					if(first) hbk_back_assemble<C, HBF_FIRST, C::LDW, hbk_nop, true>(ln, io, io + BAB, S1 + C::SX, T, hbk_nop());
					else      hbk_back_assemble<C, HBF_MID, C::LDW, hbk_nop, true>(ln, io, io + BAB, S1 + C::SX, T, hbk_nop());
					}
				__syncwarp();
				}
			if(first) hbk_back_factor<C, HBF_FIRST>(ln, T, LUs, S0 + C::SX, Lt + nd.off_L + gen_own, hbk_nop());
			else      hbk_back_factor<C, HBF_MID>(ln, T, LUs, S0 + C::SX, Lt + nd.off_L + gen_own, hbk_nop());
			__syncwarp();
			}
		else
			{
			if(lane==0) hbf_mbar_expect(&bars[1], 8u*SB*IPW);
			if(lane<IPW) hbf_bulk_g2s(my_sm + IOB + LU, my_L + my_nd.off_L + gen_own, 8u*SB, &bars[1]);
			if(!first)
				{
				if(ln.l<NX) xs0[ln.l] = ux[nd.off_ux + NU + ln.l];
				if(C::R>1 && ln.l+G<NX) xs0[ln.l+G] = ux[nd.off_ux + NU + ln.l + G];
				}
			__syncwarp();
			wait_bar(1);
			for(int kc=0; kc<nkids; kc++)
				{
				if(lane==0) hbf_mbar_expect(&bars[0], bytes_B*IPW);
				if(lane<IPW) hbf_bulk_g2s(my_sm, my_in + d.tn[my_nd.first_kid+kc].off_BAbt, bytes_B, &bars[0]);
				wait_bar(0);
				const hb_tnode kd = d.tn[nd.first_kid+kc];
				/* u and pi of the node are recomputed for every kid (a few dozen FMAs) to reuse the chain routine unchanged */
				if(first) hbk_stage_forward<C, HBF_FIRST>(ln, io, S0, us, xs0, xs1, tmp, ux + nd.off_ux, ux + kd.off_ux + NU, pi + nd.off_pi, active);
				else      hbk_stage_forward<C, HBF_MID>(ln, io, S0, us, xs0, xs1, tmp, ux + nd.off_ux, ux + kd.off_ux + NU, pi + nd.off_pi, active);
				}
			}
		}
	}
