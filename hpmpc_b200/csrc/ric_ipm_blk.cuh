/*
 * ric_ipm_blk.cuh -- the IPM's factorisation sweep with the register-blocked tile of ric_blk.cuh: G = 16 lanes per OCP instance,
 * two instances per warp, R = 2 rows of the stage's (NU+NX+1) x (NU+NX) trapezoid per lane.
 *
 * Why: in the one-row-per-lane sweep (ric_ipm_fast.cuh: hbi_backward, G = 32) every shared-memory broadcast operand feeds ONE FP64
 * FMA per lane, the shared-memory pipe is the limiter (ncu r01: 74-86 % of its peak) and a stage of BASELINE config 3 (36 x 35)
 * costs 47 K cycles per warp.  With two rows per lane a broadcast feeds two FMAs, and the per-column overhead of the Cholesky
 * (reciprocal square root, shuffle, barrier) is shared by the two instances of the warp.  The plain transfer of ric_blk.cuh's
 * assembly to this size spills (H 128 + own W rows 96 + ... > 255 registers; tried in round 1: 1.7 KB of stack, 9x slower), so
 * here the own W rows are NOT kept in registers: W is stored to shared memory first (leading dimension NX, exactly over the
 * stage's [B A b]' so the Hessian behind it is untouched), H is loaded afterwards, and the rank-NX update re-reads the own rows
 * as one LDS.128 per row and column pair.
 *
 * The factor leaves the sweep in the packed-column layout of ric_fast.cuh (hbf_cfg::colOff, DINV), so the solve-only sweeps, the
 * forward sweep and the residual sweep of ric_ipm_fast.cuh work on it unchanged.
 *
 * Restates the same reference lines as hbi_backward: lqcp_solvers/d_back_ric_rec.c:184-333 (update_b / update_q, diag += Qx,
 * gradient row += qx, Pb :273-283, dtrmm_nt_u, dsyrk_dpotrf), pivot rule kernel/c99/kernel_dpotrf_c99_lib4.c:553-573.
 */
#pragma once
#include "ric_blk.cuh"
#include "ric_ipm_fast.cuh"

template<class C>                       /* C = hbf_cfg<NX, NU, 32>: the layout the other sweeps use */
struct hbi2_cfg
	{
	typedef hbk_cfg<C::NX, C::NU, 16, 2> K;
	static constexpr int PW = hbi_cfg<C>::PER_WARP;                    /* one instance's region: the layout of hbi_ctx */
	static constexpr int PER_WARP = 2*PW;
	/* factorisation-only kernel (the forward sweep runs elsewhere): stage inputs / W, ONE factor buffer, the x-columns of the
	 * previous stage's factor, Lxx'b -- 21 instead of 27 KB per instance, five warps per SM */
	static constexpr int XC = K::xOff(C::NX);
	/* W with leading dimension NX+2: the one-row-per-lane LDS.128 / STS.128 of the own rows are conflict-free (192-byte rows put
	 * four lanes of a quarter-warp on the same banks); W then runs past [B A b]', so the Hessian is fetched behind it */
	static constexpr int LDWS = ((C::even(C::NX)/2)%2==0) ? C::even(C::NX)+2 : C::even(C::NX);
	static constexpr int QOFF = C::even(K::NZ*LDWS);
	static constexpr int IOS = QOFF + C::even(HB_TRI(C::NUX)+C::NUX);
	static constexpr int PWS = IOS + C::LBUF + XC + C::XS;
	static constexpr int PER_WARP_SLIM = 8 + 2*PWS;
	static_assert(K::uOff(C::NU)==C::colOff(C::NU) && K::uOff(C::NU)+K::xOff(C::NX-1)==C::colOff(C::NUX-1), "column layouts must agree");
	static_assert(C::BAB>=K::NZ*K::NX, "W (leading dimension NX) must fit over [B A b]'");
	static_assert(K::RO==K::GR && K::E>0, "every lane slot owns a row; the gradient row is column-owned");
	};

__device__ __forceinline__ void hbi2_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

/* ---- H <- RSQrq_n + W W',  W = [B A b]'_n Lxx_{n+1}.  sBW: [B A b]' in, W (ld NX) out; sQ: RSQrq, NOT overlapping W ---- */
template<class K, int KIND, int LDW>
__device__ __forceinline__ void hbi2_assemble(const hbk_lane<K> &ln, double *__restrict__ sBW, const double *__restrict__ sQ,
		const double *__restrict__ xc, hbk_tile<K> &T, double *__restrict__ lxb /* out: Lxx' b (NX), needs E > 0 */)
	{
	constexpr int NX = K::NX, NU = K::NU, NUX = K::NUX, G = K::G, R = K::R, RO = K::RO, E = K::E, CO = K::CO, NCC = K::NCC;
	const int l = ln.l;
	if(KIND!=HBF_LAST)
		{
		const double *sB = sBW;
		/* own rows of [B A b]' into registers; they are the only readers of those rows */
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
		/* extra rows first (every lane reads them): the lane owns columns j = l+s*G of them */
		double wx[E>0 ? E : 1][R];
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
					/* two accumulators per entry and the extra rows two columns at a time: eight independent chains */
					double wy[E>0 ? E : 1];
					#pragma unroll
					for(int e=0; e<E; e++) wy[e] = 0.0;
					#pragma unroll
					for(int k=s*G; k<NX; k+=2)
						{
						const double l0 = (j<NX && k>=j) ? col[k] : 0.0, l1 = (j<NX && k+1>=j) ? col[k+1] : 0.0;
						#pragma unroll
						for(int e=0; e<E; e++)
							{
							const int ae = hbk_arow<K, KIND>(RO+e);
							double2 b = make_double2(0.0, 0.0);
							if(ae>=0) b = *reinterpret_cast<const double2*>(sB + ae*NX + k);
							wx[e][s] = fma(b.x, l0, wx[e][s]);
							wy[e] = fma(b.y, l1, wy[e]);
							}
						}
					#pragma unroll
					for(int e=0; e<E; e++) wx[e][s] += wy[e];
					if(j<NX)
						{
						lxb[j] = wx[E-1][s];                                 /* Lxx' b: the first half of Pb = Lxx (Lxx' b) */
						wx[E-1][s] += col[NX];                               /* gradient row: + l_x[j] */
						}
					}
				}
			}
		__syncwarp();                       /* every lane has read the shared rows: the buffer becomes W (ld NX), column by column */
		if(E>0)
			{
			#pragma unroll
			for(int s=0; s<R; s++)
				{
				const int v = l + s*G;
				if(v<NX)
					{
					#pragma unroll
					for(int e=0; e<E; e++) sBW[(RO+e)*LDW + v] = wx[e][s];
					}
				}
			}
		#pragma unroll
		for(int j=0; j<NX; j+=2)
			{
			double wj[R][2];
			#pragma unroll
			for(int jj=0; jj<2; jj++)
				{
				const double *col = xc + K::xOff(j+jj);
				double acc[R][2];
				#pragma unroll
				for(int s=0; s<R; s++) { acc[s][0] = 0.0; acc[s][1] = 0.0; }
				#pragma unroll
				for(int k=j+jj; k<NX; k+=2)
					{
					const double2 t = *reinterpret_cast<const double2*>(col + (k-j-jj));
					#pragma unroll
					for(int s=0; s<R; s++)
						{
						acc[s][0] = fma(a[s][k], t.x, acc[s][0]);
						if(k+1<NX) acc[s][1] = fma(a[s][k+1], t.y, acc[s][1]);
						}
					}
				#pragma unroll
				for(int s=0; s<R; s++) wj[s][jj] = acc[s][0] + acc[s][1];
				if(E==0)
					{
					const double lx = col[NX-j-jj];
					#pragma unroll
					for(int s=0; s<R; s++) if(l+s*G==NUX) wj[s][jj] += lx;
					}
				}
			#pragma unroll
			for(int s=0; s<R; s++) *reinterpret_cast<double2*>(sBW + (l+s*G)*LDW + j) = make_double2(wj[s][0], wj[s][1]);
			}
		}
	/* ---- H <- RSQrq ---- */
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
			if(k < (s+1)*G-1 && k<K::KS)
				{
				const int ak = hbk_arow<K, KIND>(k);
				T.H[s][k] = (k<v && ar>=0 && ak>=0) ? row[ak>=0 ? ak : 0] : 0.0;
				}
			else T.H[s][k] = 0.0;
			}
		T.hd[s] = (v<CO && ar>=0) ? row[ar>=0 ? ar : 0] : 1.0;
		#pragma unroll
		for(int e=0; e<E; e++)
			{
			const int ae = hbk_arow<K, KIND>(RO+e);
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
			const int ae = hbk_arow<K, KIND>(RO+e), acol = hbk_arow<K, KIND>(CO+cc);
			double hc = (e==cc) ? 1.0 : 0.0;
			if(e>=cc && ae>=0 && acol>=0) hc = sQ[HB_TRI(ae>=0 ? ae : 0) + (acol>=0 ? acol : 0)];
			T.Z[e][cc] = hc;
			}
	__syncwarp();
	if(KIND==HBF_LAST) return;
	/* ---- H += W W' : the own rows come back from shared memory, two columns of W at a time ---- */
	const double *sW = sBW;
	/* a real loop (not unrolled): the body is ~250 instructions, and the warps of an SM have no second warp per scheduler to
	 * hide instruction-cache misses behind (ncu on the unrolled version: `no_instruction` was the top stall, icc hit rate 77 %) */
	#pragma unroll 1
	for(int m=0; m<NX; m+=2)
		{
		double2 wo[R];
		#pragma unroll
		for(int s=0; s<R; s++)
			{
			wo[s] = *reinterpret_cast<const double2*>(sW + (l+s*G)*LDW + m);
			T.hd[s] = fma(wo[s].x, wo[s].x, T.hd[s]);
			T.hd[s] = fma(wo[s].y, wo[s].y, T.hd[s]);
			}
		#pragma unroll
		for(int k4=0; k4<K::KS; k4+=4)
			{
			double2 t[4];
			#pragma unroll
			for(int q=0; q<4; q++) if(k4+q<K::KS) t[q] = *reinterpret_cast<const double2*>(sW + (k4+q)*LDW + m);
			#pragma unroll
			for(int q=0; q<4; q++)
				#pragma unroll
				for(int s=0; s<R; s++)
					if(k4+q<K::KS && k4+q < (s+1)*G-1) T.H[s][k4+q] = fma(wo[s].x, t[q].x, T.H[s][k4+q]);
			#pragma unroll
			for(int q=0; q<4; q++)
				#pragma unroll
				for(int s=0; s<R; s++)
					if(k4+q<K::KS && k4+q < (s+1)*G-1) T.H[s][k4+q] = fma(wo[s].y, t[q].y, T.H[s][k4+q]);
			}
		if(E>0)
			{
			double2 te[E>0 ? E : 1];
			#pragma unroll
			for(int e=0; e<E; e++) te[e] = *reinterpret_cast<const double2*>(sW + (RO+e)*LDW + m);
			#pragma unroll
			for(int e=0; e<E; e++)
				{
				#pragma unroll
				for(int s=0; s<R; s++)
					{
					T.X[e][s] = fma(wo[s].x, te[e].x, T.X[e][s]);
					T.X[e][s] = fma(wo[s].y, te[e].y, T.X[e][s]);
					}
				#pragma unroll
				for(int cc=0; cc<NCC; cc++)
					if(e>=cc)
						{
						T.Z[e][cc] = fma(te[e].x, te[cc].x, T.Z[e][cc]);
						T.Z[e][cc] = fma(te[e].y, te[cc].y, T.Z[e][cc]);
						}
				}
			}
		}
	__syncwarp();                       /* W is dead: the caller may refill the buffer */
	}

/* ---- right-looking Cholesky with look-ahead diagonal (ric_blk.cuh: hbk_back_factor without the K-form epilogue); the factor
 *      goes to Lc in the packed-column layout of ric_fast.cuh: column c at colOff(c), all inverse diagonals at DINV ---- */
template<class C, class K>
__device__ __forceinline__ void hbi2_factor(const hbk_lane<K> &ln, hbk_tile<K> &T, double *__restrict__ Lc, const int (&co)[2])
	{
	constexpr int NU = K::NU, G = K::G, R = K::R, RO = K::RO, E = K::E, CO = K::CO, NCC = K::NCC;
	const int l = ln.l;
	double *Xc = Lc + K::uOff(NU);
	double rs = hbk_rsqrt(T.hd[0]);
	#pragma unroll
	for(int c=0; c<CO; c++)
		{
		const int so = c/G, lo = c%G;
		const double inv = __shfl_sync(HBF_FULL, rs, lo, G);
		double lc[R];
		#pragma unroll
		for(int s=0; s<R; s++) lc[s] = 0.0;
		#pragma unroll
		for(int s=so; s<R; s++)
			{
			const int v = l + s*G;
			lc[s] = (s==so && v==c) ? T.hd[s]*inv : T.H[s][c<K::KS ? c : 0]*inv;
			if(s>so || v>c) T.hd[s] = fma(-lc[s], lc[s], T.hd[s]);
			}
		if(c+1<CO) rs = hbk_rsqrt(T.hd[(c+1)/G]);
		double *col = hbk_col<K>(c, Lc, Xc);
		#pragma unroll
		for(int s=so; s<R; s++)
			{
			const int v = l + s*G;
			if((s>so || v>=c) && v<RO) col[v-c] = lc[s];
			}
		if(l==lo)
			{
			#pragma unroll
			for(int e=0; e<E; e++) col[RO+e-c] = T.X[e][so]*inv;
			Lc[C::DINV+c] = inv;
			}
		__syncwarp();
		#pragma unroll
		for(int q=0; c+2*q<K::KS; q++)
			{
			const double2 t = *reinterpret_cast<const double2*>(col + 2*q);
			const int k0 = c+2*q, k1 = k0+1;
			#pragma unroll
			for(int s=so; s<R; s++)
				{
				if(q>0 && k0<K::KS && k0 < (s+1)*G-1) T.H[s][k0] = fma(-lc[s], t.x, T.H[s][k0]);
				if(k1<K::KS && k1 < (s+1)*G-1) T.H[s][k1] = fma(-lc[s], t.y, T.H[s][k1]);
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
				for(int s=so; s<R; s++) T.X[e][s] = fma(-le[e], lc[s], T.X[e][s]);
				}
			}
		}
	/* the replicated E x NCC corner: Z[e][cc] -= sum over the row-owned columns of L[RO+e][c] L[RO+cc][c].  Every lane sums its own
	 * two columns (the values it wrote itself), then a butterfly over the instance's lanes -- instead of 9 FMAs on every lane in
	 * every column step */
	if(E>0 && NCC>0)
		{
		double z[E>0 ? E : 1][NCC>0 ? NCC : 1];
		#pragma unroll
		for(int e=0; e<E; e++)
			#pragma unroll
			for(int cc=0; cc<NCC; cc++) z[e][cc] = 0.0;
		#pragma unroll
		for(int s=0; s<R; s++)
			{
			const int cown = l + s*G;
			if(cown<CO)
				{
				const double *col = Lc + co[s] + (RO-cown);                   /* co[s] = colOff(cown) */
				double le[E>0 ? E : 1];
				#pragma unroll
				for(int e=0; e<E; e++) le[e] = col[e];
				#pragma unroll
				for(int e=0; e<E; e++)
					#pragma unroll
					for(int cc=0; cc<NCC; cc++) if(e>=cc) z[e][cc] = fma(le[e], le[cc], z[e][cc]);
				}
			}
		#pragma unroll
		for(int e=0; e<E; e++)
			#pragma unroll
			for(int cc=0; cc<NCC; cc++)
				if(e>=cc)
					{
					double v = z[e][cc];
					#pragma unroll
					for(int o=G/2; o>0; o>>=1) v += __shfl_xor_sync(HBF_FULL, v, o, G);
					T.Z[e][cc] -= v;
					}
		}
	#pragma unroll
	for(int cc=0; cc<NCC; cc++)
		{
		const double p = T.Z[cc][cc];
		const double rsc = hbk_rsqrt(p);
		double *col = hbk_col<K>(CO+cc, Lc, Xc);
		double lcol[E>0 ? E : 1];
		lcol[cc] = p*rsc;
		#pragma unroll
		for(int e=cc+1; e<E; e++) lcol[e] = T.Z[e][cc]*rsc;
		if(l==0)
			{
			#pragma unroll
			for(int e=cc; e<E; e++) col[e-cc] = lcol[e];
			Lc[C::DINV+CO+cc] = rsc;
			}
		#pragma unroll
		for(int e=cc+1; e<E; e++)
			#pragma unroll
			for(int c2=cc+1; c2<NCC; c2++)
				if(e>=c2) T.Z[e][c2] = fma(-lcol[e], lcol[c2], T.Z[e][c2]);
		}
	__syncwarp();
	}

/* ---- the sweep: two instances per warp.  wbase: the warp's shared memory, two regions of hbi2_cfg::PW doubles in the layout of
 *      hbi_ctx (so the forward sweep of ric_ipm_fast.cuh can follow on either region); every pointer argument is the lane's own
 *      instance's (lanes 0-15: instance 0, lanes 16-31: instance 1), null rqv / bv per instance allowed ---- */
template<class C, bool SLIM = false>
struct hbi2_ctx
	{
	static constexpr int RS = SLIM ? hbi2_cfg<C>::PWS : hbi2_cfg<C>::PW;    /* distance between the two instances' regions */
	int lane, l, g, N;
	uint64_t *bar;
	uint32_t ph;
	double *wbase, *io, *Lb0, *Lb1, *tmp, *xp;
	int o_in1, s_in, o_inN;
	__device__ __forceinline__ void init(double *wbase_, int lane_, const hb_dims &d)
		{
		lane = lane_; l = lane_&15; g = lane_>>4; N = d.N; wbase = wbase_;
		io = wbase_ + 8 + (size_t)g*RS;
		if(SLIM) { Lb0 = io + hbi2_cfg<C>::IOS; Lb1 = Lb0; xp = Lb0 + C::LBUF; tmp = xp + hbi2_cfg<C>::XC; }
		else { Lb0 = io + C::IOB; Lb1 = Lb0 + C::LBUF; xp = nullptr; tmp = Lb1 + C::LBUF + C::even(C::NU) + 2*C::XS; }
		bar = reinterpret_cast<uint64_t*>(wbase_) + 4;            /* slots 0..3 belong to the hbi_ctx of region 0 */
		ph = 0;
		o_in1 = d.st[1].off_BAbt; s_in = d.st[2].off_BAbt - d.st[1].off_BAbt; o_inN = d.st[N].off_BAbt;
		if(lane==0) { hbf_mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
		__syncwarp();
		}
	__device__ __forceinline__ int off_in(int n) const { return (n==0) ? 0 : (n==N ? o_inN : o_in1 + (n-1)*s_in); }
	__device__ __forceinline__ int off_ux(int n) const { return (n==0) ? 0 : C::NU + (n-1)*C::NUX; }
	};

template<class C, bool SLIM>
__device__ void hbi2_backward(hbi2_ctx<C, SLIM> &c, const hb_dims &d, const double *__restrict__ in_inst, double *__restrict__ Lst,
		const double *bv, const double *rqv, const double *Qx, const double *qx, double *Pb)
	{
	typedef typename hbi2_cfg<C>::K K;
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, LBUF = C::LBUF, BAB = C::BAB, PW = hbi2_ctx<C, SLIM>::RS;
	constexpr int LDW = SLIM ? hbi2_cfg<C>::LDWS : NX, QO = SLIM ? hbi2_cfg<C>::QOFF : BAB;      /* W's leading dimension, where RSQrq sits in io */
	const int lane = c.lane, l = c.l, N = c.N;
	hbk_lane<K> ln; ln.init(l);
	int co[2] = {0, 0};                                                  /* offsets of the two factor columns this lane finishes */
	#pragma unroll
	for(int s=0; s<2; s++)
		#pragma unroll
		for(int c2=0; c2<K::CO; c2++) if(c2==l+16*s) co[s] = C::colOff(c2);
	/* lane 0 moves the data of both instances */
	const double *in_o = reinterpret_cast<const double*>(__shfl_sync(HBF_FULL, reinterpret_cast<unsigned long long>(in_inst), 16));
	double *Lst_o = reinterpret_cast<double*>(__shfl_sync(HBF_FULL, reinterpret_cast<unsigned long long>(Lst), 16));
	double *io0 = c.wbase + 8, *io1 = c.wbase + PW + 8;
	auto issue = [&](int n)
		{
		if(lane!=0) return;
		const int off = c.off_in(n);
		if(n==0)
			{
			const uint32_t bB = hbi_bytes_BAbt<C>(true), bQ = hbi_bytes_RSQ<C>(HBF_FIRST);
			hbf_mbar_expect(c.bar, 2u*(bB+bQ));
			hbf_bulk_g2s(io0, in_inst + off, bB, c.bar); hbf_bulk_g2s(io0 + QO, in_inst + off + bB/8, bQ, c.bar);
			hbf_bulk_g2s(io1, in_o + off, bB, c.bar); hbf_bulk_g2s(io1 + QO, in_o + off + bB/8, bQ, c.bar);
			}
		else if(n==N)
			{
			const uint32_t bQ = hbi_bytes_RSQ<C>(HBF_LAST);
			hbf_mbar_expect(c.bar, 2u*bQ);
			hbf_bulk_g2s(io0 + QO, in_inst + off, bQ, c.bar); hbf_bulk_g2s(io1 + QO, in_o + off, bQ, c.bar);
			}
		else if(SLIM)
			{
			const uint32_t bB = 8u*(uint32_t)BAB, bQ = 8u*(uint32_t)(C::INB-BAB);
			hbf_mbar_expect(c.bar, 2u*(bB+bQ));
			hbf_bulk_g2s(io0, in_inst + off, bB, c.bar); hbf_bulk_g2s(io0 + QO, in_inst + off + BAB, bQ, c.bar);
			hbf_bulk_g2s(io1, in_o + off, bB, c.bar); hbf_bulk_g2s(io1 + QO, in_o + off + BAB, bQ, c.bar);
			}
		else
			{
			hbf_mbar_expect(c.bar, 2u*8u*(uint32_t)C::INB);
			hbf_bulk_g2s(io0, in_inst + off, 8u*C::INB, c.bar); hbf_bulk_g2s(io1, in_o + off, 8u*C::INB, c.bar);
			}
		};
	double *sQ = c.io + QO;
	issue(N);
	for(int n=N; n>=0; n--)
		{
		const int nux = (n==0) ? NU : (n==N ? NX : NUX), brow = (n==0) ? NU : NUX;
		double *Lc = (n&1) ? c.Lb1 : c.Lb0;
		const double *Lp = (n&1) ? c.Lb0 : c.Lb1;
		const double *xc = SLIM ? c.xp : Lp + C::colOff(NU);             /* x-columns of the factor of stage n+1 */
		hbf_mbar_wait(c.bar, c.ph&1); c.ph ^= 1u;
		/* ---- hooks: new right-hand sides, barrier terms of the IPM (all global loads first: one round trip, not two) ---- */
		{
		const hb_stage s = d.st[n];
		int id[3]; double vQ[3], vq[3];
		#pragma unroll
		for(int t=0; t<3; t++)
			{
			const int j = l + 16*t;
			id[t] = -1; vQ[t] = 0.0; vq[t] = 0.0;
			if(Qx!=nullptr && j<s.nb)
				{
				id[t] = d.idxb[s.off_c+j]; vQ[t] = Qx[s.off_c+j];
				if(qx!=nullptr) vq[t] = qx[s.off_c+j];
				}
			}
		if(rqv!=nullptr) for(int i=l; i<nux; i+=16) sQ[HB_TRI(nux)+i] = rqv[c.off_ux(n)+i];
		if(bv!=nullptr && n<N) for(int j=l; j<NX; j+=16) c.io[brow*NX+j] = bv[n*NX+j];
		if(n>0)
			{
			/* the next stage's hook vectors into L1 now: with one warp per scheduler nothing else hides their latency later */
			const hb_stage s1 = d.st[n-1];
			const int o1 = 16*l;                                             /* one 128-byte line per lane */
			if(Qx!=nullptr && o1<s1.nb) hbi2_prefetch_l1(Qx + s1.off_c + o1);
			if(qx!=nullptr && o1<s1.nb) hbi2_prefetch_l1(qx + s1.off_c + o1);
			if(rqv!=nullptr && o1<NUX) hbi2_prefetch_l1(rqv + c.off_ux(n-1) + o1);
			if(bv!=nullptr && o1<NX) hbi2_prefetch_l1(bv + (n-1)*NX + o1);
			}
		__syncwarp();
		#pragma unroll
		for(int t=0; t<3; t++)
			if(id[t]>=0)
				{
				sQ[HB_TRI(id[t])+id[t]] += vQ[t];
				if(qx!=nullptr) sQ[HB_TRI(nux)+id[t]] += vq[t];
				}
		}
		__syncwarp();
		hbk_tile<K> T;
		if(n==N) hbi2_assemble<K, HBF_LAST, LDW>(ln, c.io, sQ, xc, T, c.tmp);
		else if(n==0) hbi2_assemble<K, HBF_FIRST, LDW>(ln, c.io, sQ, xc, T, c.tmp);
		else hbi2_assemble<K, HBF_MID, LDW>(ln, c.io, sQ, xc, T, c.tmp);
		/* ---- Pb_n = Lxx (Lxx' b) with Lxx of stage n+1; Lxx' b came out of the assembly (the b-row of W before l_x is added) ---- */
		if(n<N && Pb!=nullptr)
			{
			#pragma unroll
			for(int s=0; s<2; s++)
				{
				const int r = l + 16*s;
				if(r<NX)
					{
					double p0 = 0.0, p1 = 0.0;
					#pragma unroll
					for(int cc=0; cc<NX; cc+=2)
						{
						if(cc<=r) p0 = fma(xc[K::xOff(cc) + (r-cc)], c.tmp[cc], p0);
						if(cc+1<=r) p1 = fma(xc[K::xOff(cc+1) + (r-cc-1)], c.tmp[cc+1], p1);
						}
					Pb[n*NX+r] = p0 + p1;
					}
				}
			}
		if(n>0) issue(n-1);
		/* the factor buffer about to be overwritten must have left for HBM: the store before last (two buffers), the last one (one) */
		if(lane==0) { if(SLIM) hbf_bulk_wait_read<0>(); else hbf_bulk_wait_read<1>(); }
		__syncwarp();
		hbi2_factor<C, K>(ln, T, Lc, co);
		hbf_fence_async();
		__syncwarp();
		if(lane==0)
			{
			const int lo = (int)(Lc - c.io);                                 /* same offset in both regions */
			hbf_bulk_s2g(Lst + (long long)n*LBUF, io0 + lo, 8u*LBUF);
			hbf_bulk_s2g(Lst_o + (long long)n*LBUF, io1 + lo, 8u*LBUF);
			hbf_bulk_commit();
			}
		if(SLIM && n>0)
			{
			/* the next stage needs the x-columns only: keep a copy, the factor buffer is reused */
			const double *src = Lc + C::colOff(NU);
			for(int i=2*l; i<hbi2_cfg<C>::XC; i+=32) *reinterpret_cast<double2*>(c.xp + i) = *reinterpret_cast<const double2*>(src + i);
			__syncwarp();
			}
		}
	if(lane==0) hbf_bulk_wait_all<0>();
	__syncwarp();
	}
