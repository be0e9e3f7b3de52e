/*
 * ric_ipm_fast.cuh -- size-specialised sweeps for the fused box-IPM kernel (one warp per OCP instance, G = 32 lanes,
 * compile-time NX, NU; x0 eliminated: stage 0 has nx = 0, stage N has nu = 0).
 *
 * The generic IPM kernel (ric_kernels.cu: hb_ipm_kernel) spends its time copying every stage's factor and [B A b]' from
 * HBM into shared memory with a handful of loads in flight per lane (latency bound) and in shared-memory-bound inner
 * loops.  Here every sweep over the horizon is a software pipeline: the data of the next stage(s) is fetched by 1-D bulk
 * async copies (UBLKCP) behind mbarriers while the current stage is computed, the factorisation uses the register-tile
 * routines of ric_fast.cuh, and the factor is kept in HBM in that file's packed-column layout.
 *
 * What each sweep restates (reference paths relative to /root/reference):
 *   hbi_backward      lqcp_solvers/d_back_ric_rec.c:184-333   sv with update_b / update_q, diag += Qx, gradient row += qx,
 *                                                              Pb (:273-283), factor kept for trs
 *   hbi_forward       lqcp_solvers/d_back_ric_rec.c:341-397 (sv), :737-789 (trs)
 *   hbi_trs_backward  lqcp_solvers/d_back_ric_rec.c:628-732   backward vector sweep of the solve-only routine
 *   hbi_residuals     mpc_solvers/c99/d_res_ip_res_hard.c:39-319 ; exit norms mpc_solvers/d_res_ip_hard.c:38
 */
#pragma once
#include "ric_fast.cuh"

template<class C>
struct hbi_cfg
	{
	static constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, NZ = C::NZ;
	static constexpr int VZ = C::even(NZ);
	static constexpr int PER_WARP = 8 + C::IOB + 2*C::LBUF + C::even(NU) + 3*C::XS + 3*VZ;
	static_assert(C::G==32, "the fused IPM sweeps use one warp per instance");
	static_assert(C::even(HB_TRI(NUX)+NUX)<=C::LBUF, "a Hessian block must fit a factor buffer (residual sweep)");
	};

template<class C>
struct hbi_ctx
	{
	static constexpr int SB = C::BAB;                   /* distance between the two [B A b]' slots of io */
	int lane, N;
	uint64_t *bars;
	double *io, *Lb0, *Lb1, *us, *xs0, *xs1, *tmp, *va, *vb, *vc;
	uint32_t phase;
	int o_in1, s_in, o_inN;
	__device__ __forceinline__ void init(double *wbase, int lane_, const hb_dims &d)
		{
		lane = lane_; N = d.N;
		bars = reinterpret_cast<uint64_t*>(wbase);
		io = wbase + 8; Lb0 = io + C::IOB; Lb1 = Lb0 + C::LBUF;
		us = Lb1 + C::LBUF; xs0 = us + C::even(C::NU); xs1 = xs0 + C::XS; tmp = xs1 + C::XS;
		va = tmp + C::XS; vb = va + hbi_cfg<C>::VZ; vc = vb + hbi_cfg<C>::VZ;
		phase = 0;
		o_in1 = d.st[1].off_BAbt; s_in = d.st[2].off_BAbt - d.st[1].off_BAbt; o_inN = d.st[N].off_BAbt;
		if(lane==0)
			{
			for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
		__syncwarp();
		}
	__device__ __forceinline__ void wait(int b) { hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b); }
	__device__ __forceinline__ void kick(int) {}            /* the copies are already in flight (see hbi_ctx1) */
	__device__ __forceinline__ void load(int b, double *dst, const double *src, uint32_t bytes)
		{
		if(lane==0) { hbf_mbar_expect(&bars[b], bytes); hbf_bulk_g2s(dst, src, bytes, &bars[b]); }
		}
	__device__ __forceinline__ int off_in(int n) const { return (n==0) ? 0 : (n==N ? o_inN : o_in1 + (n-1)*s_in); }
	__device__ __forceinline__ int off_ux(int n) const { return (n==0) ? 0 : C::NU + (n-1)*C::NUX; }
	/* generic -> async proxy ordering before a buffer this warp has written or read is refilled by a bulk copy */
	__device__ __forceinline__ void sync() { __syncwarp(); }
	};

/* The same interface with ONE slot per kind of stage data: half the shared memory (14 instead of 27 KB per warp at config 3), so
 * the solve-only and residual sweeps of the multi-kernel driver run 16 warps per SM.  load() only records the request; the bulk
 * copy is issued by wait(), i.e. when the previous stage's data is dead -- nothing is prefetched inside a warp, the other warps of
 * the SM cover the latency (the sweeps keep their two-stage-ahead issue order, which is what makes two records per kind enough). */
template<class C>
struct hbi_cfg1
	{
	static constexpr int VZ = C::even(C::NZ);
	static constexpr int PER_WARP = 8 + C::BAB + C::LBUF + C::even(C::NU) + 3*C::XS + 3*VZ;
	static_assert(C::even(HB_TRI(C::NUX)+C::NUX)<=C::LBUF, "a Hessian block must fit the factor buffer (residual sweep)");
	};
template<class C>
struct hbi_ctx1
	{
	static constexpr int SB = 0;
	int lane, N;
	uint64_t *bars;
	double *io, *Lb0, *Lb1, *us, *xs0, *xs1, *tmp, *va, *vb, *vc;
	uint32_t phase;
	int o_in1, s_in, o_inN;
	const double *p_src0, *p_src1, *p_src2, *p_src3;    /* recorded requests, one per barrier */
	uint32_t p_b0, p_b1, p_b2, p_b3, issued;
	__device__ __forceinline__ void init(double *wbase, int lane_, const hb_dims &d)
		{
		lane = lane_; N = d.N;
		bars = reinterpret_cast<uint64_t*>(wbase);
		io = wbase + 8; Lb0 = io + C::BAB; Lb1 = Lb0;
		us = Lb0 + C::LBUF; xs0 = us + C::even(C::NU); xs1 = xs0 + C::XS; tmp = xs1 + C::XS;
		va = tmp + C::XS; vb = va + hbi_cfg1<C>::VZ; vc = vb + hbi_cfg1<C>::VZ;
		phase = 0;
		o_in1 = d.st[1].off_BAbt; s_in = d.st[2].off_BAbt - d.st[1].off_BAbt; o_inN = d.st[N].off_BAbt;
		p_src0 = p_src1 = p_src2 = p_src3 = nullptr; p_b0 = p_b1 = p_b2 = p_b3 = 0; issued = 0;
		if(lane==0)
			{
			for(int b=0; b<4; b++) hbf_mbar_init(&bars[b], 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
		__syncwarp();
		}
	__device__ __forceinline__ void load(int b, double *, const double *src, uint32_t bytes)
		{
		if(b==0) { p_src0 = src; p_b0 = bytes; } else if(b==1) { p_src1 = src; p_b1 = bytes; }
		else if(b==2) { p_src2 = src; p_b2 = bytes; } else { p_src3 = src; p_b3 = bytes; }
		}
	/* issue the recorded copy of barrier b now (the data it replaces must be dead): a sweep calls this at the top of a stage, before
	 * it fetches the stage's vectors from global memory, so that the two latencies overlap; wait() issues it if nobody did */
	__device__ __forceinline__ void kick(int b)
		{
		if((issued>>b)&1u) return;
		const double *src = (b==0) ? p_src0 : (b==1 ? p_src1 : (b==2 ? p_src2 : p_src3));
		const uint32_t bytes = (b==0) ? p_b0 : (b==1 ? p_b1 : (b==2 ? p_b2 : p_b3));
		__syncwarp();                                   /* every lane is done with the data this copy replaces */
		if(lane==0) { hbf_mbar_expect(&bars[b], bytes); hbf_bulk_g2s(b<2 ? io : Lb0, src, bytes, &bars[b]); }
		issued |= (1u<<b);
		}
	__device__ __forceinline__ void wait(int b)
		{
		kick(b);
		hbf_mbar_wait(&bars[b], (phase>>b)&1); phase ^= (1u<<b);
		issued &= ~(1u<<b);
		}
	__device__ __forceinline__ int off_in(int n) const { return (n==0) ? 0 : (n==N ? o_inN : o_in1 + (n-1)*s_in); }
	__device__ __forceinline__ int off_ux(int n) const { return (n==0) ? 0 : C::NU + (n-1)*C::NUX; }
	__device__ __forceinline__ void sync() { __syncwarp(); }
	};

/* one 128-byte line per lane of a vector of n doubles into L1 (the sweeps fetch the NEXT stage's vectors a stage ahead: in the
 * single-slot sweeps a warp has nothing else in flight while it waits for them) */
__device__ __forceinline__ void hbi_prefetch_vec(const double *p, int n, int lane)
	{
	if(p!=nullptr && 16*lane<n) asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 16*lane));
	}

template<class C> __device__ __forceinline__ constexpr uint32_t hbi_bytes_BAbt(bool first)
	{ return 8u*(uint32_t)(first ? C::even((C::NU+1)*C::NX) : C::BAB); }
template<class C> __device__ __forceinline__ constexpr uint32_t hbi_bytes_RSQ(int kind)
	{ return 8u*(uint32_t)(kind==HBF_FIRST ? C::even(HB_TRI(C::NU)+C::NU) : (kind==HBF_LAST ? C::even(HB_TRI(C::NX)+C::NX) : C::even(HB_TRI(C::NUX)+C::NUX))); }

/* ------------------------------------------------------------------------------------------------ */
/* factorisation sweep n = N..0 ; factor of stage n -> Lst + n*LBUF (packed columns), Pb edge-indexed  */
/*   rqv : replaces the gradient row of RSQrq (ux layout) when non-null ; bv : replaces b (pi layout)  */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ void hbi_backward(hbi_ctx<C> &c, const hb_dims &d, const double *__restrict__ in_inst, double *__restrict__ Lst,
		const double *bv, const double *rqv, const double *Qx, const double *qx, double *Pb)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, LBUF = C::LBUF, BAB = C::BAB;
	const int lane = c.lane, N = c.N;
	auto issue = [&](int n)
		{
		const uint32_t bytes = (n==0) ? hbi_bytes_BAbt<C>(true) + hbi_bytes_RSQ<C>(HBF_FIRST) : (n==N ? hbi_bytes_RSQ<C>(HBF_LAST) : 8u*(uint32_t)C::INB);
		c.load(0, c.io, in_inst + c.off_in(n), bytes);
		};
	issue(N);
	for(int n=N; n>=0; n--)
		{
		const int kind = (n==0) ? HBF_FIRST : (n==N ? HBF_LAST : HBF_MID);
		const int rsq_off = (n==0) ? C::even((NU+1)*NX) : (n==N ? 0 : BAB);
		const int nux = (n==0) ? NU : (n==N ? NX : NUX), brow = (n==0) ? NU : NUX;
		double *Lc = (n&1) ? c.Lb1 : c.Lb0;
		const double *Lp = (n&1) ? c.Lb0 : c.Lb1;
		double *sQ = c.io + rsq_off;
		HBF_STAMP(310);
		c.wait(0);
		HBF_STAMP(311);
		/* ---- hooks: new right-hand sides, barrier terms of the IPM ---- */
		if(rqv!=nullptr) for(int i=lane; i<nux; i+=32) sQ[HB_TRI(nux)+i] = rqv[c.off_ux(n)+i];
		if(bv!=nullptr && n<N) for(int j=lane; j<NX; j+=32) c.io[brow*NX+j] = bv[n*NX+j];
		__syncwarp();
		{
		const hb_stage s = d.st[n];
		if(Qx!=nullptr)
			for(int j=lane; j<s.nb; j+=32)
				{
				const int id = d.idxb[s.off_c+j];
				sQ[HB_TRI(id)+id] += Qx[s.off_c+j];
				if(qx!=nullptr) sQ[HB_TRI(nux)+id] += qx[s.off_c+j];
				}
		}
		__syncwarp();
		HBF_STAMP(312);
		/* ---- Pb_n = Lxx (Lxx' b)  with Lxx of stage n+1 ---- */
		if(n<N && Pb!=nullptr)
			{
			if(lane<NX)
				{
				int coff = 0;
				#pragma unroll
				for(int jj=0; jj<NX; jj++) if(jj==lane) coff = C::colOff(NU+jj);
				const double *col = Lp + coff - lane;               /* col[m] = Lxx[m][lane] */
				double a0 = 0.0;
				#pragma unroll
				for(int m=0; m<NX; m++) if(m>=lane) a0 = fma(col[m], c.io[brow*NX+m], a0);
				c.tmp[lane] = a0;
				}
			__syncwarp();
			if(lane<NX)
				{
				double p0 = 0.0;
				#pragma unroll
				for(int cc=0; cc<NX; cc++) if(cc<=lane) p0 = fma(Lp[C::colOff(NU+cc) + (lane-cc)], c.tmp[cc], p0);
				Pb[n*NX+lane] = p0;
				}
			__syncwarp();
			}
		HBF_STAMP(313);
		hbf_tile<C> T;
		hbf_back_assemble<C>(lane, kind, c.io, rsq_off, Lp, T);
		HBF_STAMP(314);
		if(n>0) issue(n-1);
		if(lane==0) hbf_bulk_wait_read<1>();
		__syncwarp();
		HBF_STAMP(315);
		hbf_back_factor<C>(lane, T, Lc);
		HBF_STAMP(316);
		hbf_fence_async();
		__syncwarp();
		if(lane==0)
			{
			hbf_bulk_s2g(Lst + (long long)n*LBUF, Lc, 8u*LBUF);
			hbf_bulk_commit();
			}
		}
	if(lane==0) hbf_bulk_wait_all<0>();
	__syncwarp();
	}

/* ------------------------------------------------------------------------------------------------ */
/* forward sweep n = 0..N-1 (+ pi of the last edge).  TRS: w (ux layout, may alias ux) holds the       */
/* eliminated right-hand side of hbi_trs_backward: l_u := w_n[:nu], p := x-part of w_n                 */
/* ------------------------------------------------------------------------------------------------ */
template<class C, bool TRS, class X>
__device__ void hbi_forward(X &c, const double *__restrict__ in_inst, const double *__restrict__ Lst,
		const double *bv, const double *w, double *ux, double *pi)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, NZ = C::NZ, LBUF = C::LBUF;
	const int l = c.lane, N = c.N;
	auto issue_B = [&](int n) { c.load(n&1, c.io + (n&1)*X::SB, in_inst + c.off_in(n), hbi_bytes_BAbt<C>(n==0)); };
	auto issue_L = [&](int n) { c.load(2+(n&1), (n&1) ? c.Lb1 : c.Lb0, Lst + (long long)n*LBUF, 8u*LBUF); };
	issue_L(0); issue_B(0);
	issue_L(1); if(N>1) issue_B(1);
	int coff = 0, coff1 = 0;
	#pragma unroll
	for(int cc=0; cc<NU; cc++) if(cc==l) coff = C::colOff(cc);
	#pragma unroll
	for(int jj=0; jj<NX; jj++) if(jj==l) coff1 = C::colOff(NU+jj);
	if(l<NX) c.xs0[l] = 0.0;
	double preg = 0.0;                                      /* TRS: x-part of w_n, read before x_n overwrites it */
	__syncwarp();
	for(int n=0; n<N; n++)
		{
		const bool first = (n==0);
		const double *Ln = (n&1) ? c.Lb1 : c.Lb0;
		double *sB = c.io + (n&1)*X::SB;
		const double *xs = (n&1) ? c.xs1 : c.xs0;
		double *xo = (n&1) ? c.xs0 : c.xs1;
		const int brow = first ? NU : NUX;
		c.kick(2+(n&1)); c.kick(n&1);
		if(n+1<N) { hbi_prefetch_vec(bv!=nullptr ? bv + (n+1)*NX : nullptr, NX, l); if(TRS) hbi_prefetch_vec(w + c.off_ux(n+1), NUX, l); }
		const double bvl = (bv!=nullptr && l<NX) ? bv[n*NX+l] : 0.0;            /* in flight while the stage's matrices arrive */
		c.wait(2+(n&1)); c.wait(n&1);
		if(bv!=nullptr) { if(l<NX) sB[brow*NX+l] = bvl; __syncwarp(); }
		const int o_ux = c.off_ux(n), o_ux1 = c.off_ux(n+1) + ((n+1<N) ? NU : 0);
		/* ---- phase A ---- */
		double t = 0.0;
		if(l<NU)
			{
			const double *col = Ln + coff - l;
			double t0 = TRS ? w[o_ux+l] : col[NZ-1], t1 = 0.0;
			if(!first)
				{
				#pragma unroll
				for(int k=0; k<NX; k+=2) { t0 = fma(col[NU+k], xs[k], t0); t1 = fma(col[NU+k+1], xs[k+1], t1); }
				}
			t = -(t0+t1);
			}
		if(!first && l<NX)
			{
			const double *col = Ln + coff1 - (NU+l);
			double a0 = TRS ? 0.0 : col[NZ-1], a1 = 0.0;
			#pragma unroll
			for(int k=0; k<NX; k+=2)
				{
				if(k>=l) a0 = fma(col[NU+k], xs[k], a0);
				if(k+1>=l) a1 = fma(col[NU+k+1], xs[k+1], a1);
				}
			c.tmp[l] = a0+a1;
			}
		const double di = (l<NU) ? Ln[C::DINV+l] : 0.0;
		#pragma unroll
		for(int j=NU-1; j>=0; j--)
			{
			const double vj = __shfl_sync(HBF_FULL, t*di, j);
			if(l==j) t = vj;
			else if(l<j) t = fma(-Ln[coff - l + j], vj, t);
			}
		if(l<NU) { c.us[l] = t; ux[o_ux+l] = t; }
		__syncwarp();
		/* ---- phase B ---- */
		if(l<NX)
			{
			double x0 = sB[brow*NX + l], x1 = 0.0, x2 = 0.0;
			#pragma unroll
			for(int i=0; i<NU; i++) x0 = fma(sB[i*NX+l], c.us[i], x0);
			if(!first)
				{
				#pragma unroll
				for(int i=0; i<NX; i+=2) { x1 = fma(sB[(NU+i)*NX+l], xs[i], x1); x2 = fma(sB[(NU+i+1)*NX+l], xs[i+1], x2); }
				}
			const double xn = x0 + (x1+x2);
			xo[l] = xn;
			if(!first)
				{
				double p0 = TRS ? preg : 0.0, p1 = 0.0;
				#pragma unroll
				for(int cc=0; cc<NX; cc+=2)
					{
					if(cc<=l) p0 = fma(Ln[C::colOff(NU+cc) + (l-cc)], c.tmp[cc], p0);
					if(cc+1<=l) p1 = fma(Ln[C::colOff(NU+cc+1) + (l-cc-1)], c.tmp[cc+1], p1);
					}
				pi[(n-1)*NX+l] = p0+p1;
				}
			if(TRS) preg = w[o_ux1+l];
			ux[o_ux1+l] = xn;
			}
		__syncwarp();
		if(n+2<=N) issue_L(n+2);
		if(n+2<N) issue_B(n+2);
		}
	/* pi_{N-1} from x_N and L_N */
	c.wait(2+(N&1));
	{
	const double *Ln = (N&1) ? c.Lb1 : c.Lb0;
	const double *xs = (N&1) ? c.xs1 : c.xs0;
	if(l<NX)
		{
		const double *col = Ln + coff1 - (NU+l);
		double a0 = TRS ? 0.0 : col[NZ-1];
		#pragma unroll
		for(int k=0; k<NX; k++) if(k>=l) a0 = fma(col[NU+k], xs[k], a0);
		c.tmp[l] = a0;
		}
	__syncwarp();
	if(l<NX)
		{
		double p0 = TRS ? preg : 0.0;
		#pragma unroll
		for(int cc=0; cc<NX; cc++) if(cc<=l) p0 = fma(Ln[C::colOff(NU+cc) + (l-cc)], c.tmp[cc], p0);
		pi[(N-1)*NX+l] = p0;
		}
	__syncwarp();
	}
	}

/* ------------------------------------------------------------------------------------------------ */
/* solve-only backward vector sweep: w_n = L-eliminated( rq_n (+qx) + [B A]'(Pb_n + w_{n+1,x}) ) -> wv  */
/* ------------------------------------------------------------------------------------------------ */
template<class C, class X>
__device__ void hbi_trs_backward(X &c, const hb_dims &d, const double *__restrict__ in_inst, const double *__restrict__ Lst,
		const double *rqv, const double *qx, const double *Pb, double *wv)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, LBUF = C::LBUF;
	const int l = c.lane, N = c.N;
	auto issue_B = [&](int n) { c.load(n&1, c.io + (n&1)*X::SB, in_inst + c.off_in(n), hbi_bytes_BAbt<C>(n==0)); };
	auto issue_L = [&](int n) { c.load(2+(n&1), (n&1) ? c.Lb1 : c.Lb0, Lst + (long long)n*LBUF, 8u*LBUF); };
	issue_L(N-1); issue_B(N-1);
	if(N>1) { issue_L(N-2); issue_B(N-2); }
	{
	const hb_stage s = d.st[N];
	const int o = c.off_ux(N);
	for(int i=l; i<NX; i+=32) wv[o+i] = rqv[o+i];
	__syncwarp();
	if(qx!=nullptr) for(int j=l; j<s.nb; j+=32) wv[o+d.idxb[s.off_c+j]] += qx[s.off_c+j];
	__syncwarp();
	/* the x-part of w_{n+1} travels to the next stage through shared memory (vc), not through the global array it is stored in */
	if(l<NX) c.vc[l] = wv[o+l];
	__syncwarp();
	}
	for(int n=N-1; n>=0; n--)
		{
		const bool first = (n==0);
		const double *Ln = (n&1) ? c.Lb1 : c.Lb0;
		const double *sB = c.io + (n&1)*X::SB;
		const int nux = first ? NU : NUX, o = c.off_ux(n), o1 = c.off_ux(n+1) + ((n+1<N) ? NU : 0);
		const hb_stage s = d.st[n];
		c.kick(2+(n&1)); c.kick(n&1);
		if(n>0)
			{
			const hb_stage sp = d.st[n-1];
			hbi_prefetch_vec(rqv + c.off_ux(n-1), NUX, l); hbi_prefetch_vec(Pb + (n-1)*NX, NX, l);
			hbi_prefetch_vec(qx!=nullptr ? qx + sp.off_c : nullptr, sp.nb, l);
			}
		for(int i=l; i<nux; i+=32) c.va[i] = rqv[o+i];
		if(l<NX) c.vb[l] = Pb[n*NX+l] + c.vc[l];
		__syncwarp();
		if(qx!=nullptr) for(int j=l; j<s.nb; j+=32) c.va[d.idxb[s.off_c+j]] += qx[s.off_c+j];
		__syncwarp();
		c.wait(2+(n&1)); c.wait(n&1);
		/* v = va + [B A]' vb : lane l owns rows l and l+32 */
		double v0 = 0.0, v1 = 0.0;
		if(l<nux)
			{
			double a0 = c.va[l], a1 = 0.0;
			#pragma unroll
			for(int j=0; j<NX; j+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(sB + l*NX + j);
				a0 = fma(t.x, c.vb[j], a0); a1 = fma(t.y, c.vb[j+1], a1);
				}
			v0 = a0+a1;
			}
		if(NUX>32 && l+32<nux)
			{
			double a0 = c.va[l+32], a1 = 0.0;
			#pragma unroll
			for(int j=0; j<NX; j+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(sB + (l+32)*NX + j);
				a0 = fma(t.x, c.vb[j], a0); a1 = fma(t.y, c.vb[j+1], a1);
				}
			v1 = a0+a1;
			}
		/* forward substitution with the first NU columns of L_n */
		const double di = (l<NU) ? Ln[C::DINV+l] : 0.0;
		#pragma unroll
		for(int j=0; j<NU; j++)
			{
			const double vj = __shfl_sync(HBF_FULL, v0*di, j);
			const double *col = Ln + C::colOff(j) - j;            /* col[r] = L[r][j] in frame rows */
			if(l==j) v0 = vj;
			else if(l>j && l<nux) v0 = fma(-col[l], vj, v0);
			if(NUX>32 && l+32<nux) v1 = fma(-col[l+32], vj, v1);
			}
		if(l<nux) wv[o+l] = v0;
		if(NUX>32 && l+32<nux) wv[o+l+32] = v1;
		if(!first)
			{
			/* x-part of w_n = entries NU.. of this stage's vector */
			if(l>=NU && l<nux) c.vc[l-NU] = v0;
			if(NUX>32 && l+32<nux) c.vc[l+32-NU] = v1;
			}
		__syncwarp();
		if(n-2>=0) { issue_L(n-2); issue_B(n-2); }
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* the same sweep for a NEW b (d_back_ric_rec_trs_tv_res with compute_Pb = 1): Pb_n is formed from the factor of stage n+1   */
/* while that factor is on chip for its own elimination step, instead of a separate sweep over all the factors              */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ void hbi_trs_backward_newb(hbi_ctx<C> &c, const hb_dims &d, const double *__restrict__ in_inst, const double *__restrict__ Lst,
		const double *rqv, const double *qx, const double *bv, double *wv)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, LBUF = C::LBUF, BAB = C::BAB;
	const int l = c.lane, N = c.N;
	auto issue_B = [&](int n) { c.load(n&1, c.io + (n&1)*BAB, in_inst + c.off_in(n), hbi_bytes_BAbt<C>(n==0)); };
	auto issue_L = [&](int n) { c.load(2+(n&1), (n&1) ? c.Lb1 : c.Lb0, Lst + (long long)n*LBUF, 8u*LBUF); };
	int coff = 0;
	#pragma unroll
	for(int jj=0; jj<NX; jj++) if(jj==l) coff = C::colOff(NU+jj);
	/* Pb_n = Lxx_{n+1} (Lxx_{n+1}' b_n) with the factor of stage n+1 resident in Lp; lane l < NX returns component l */
	auto Pb_from = [&](const double *Lp, int n) -> double
		{
		if(l<NX) c.va[l] = bv[n*NX+l];
		__syncwarp();
		if(l<NX)
			{
			const double *col = Lp + coff - l;                     /* col[m] = Lxx[m][l] */
			double a0 = 0.0;
			#pragma unroll
			for(int m=0; m<NX; m++) if(m>=l) a0 = fma(col[m], c.va[m], a0);
			c.tmp[l] = a0;
			}
		__syncwarp();
		double p0 = 0.0;
		if(l<NX)
			{
			#pragma unroll
			for(int cc=0; cc<NX; cc++) if(cc<=l) p0 = fma(Lp[C::colOff(NU+cc) + (l-cc)], c.tmp[cc], p0);
			}
		__syncwarp();
		return p0;
		};
	/* the factor of stage N is needed once, for Pb_{N-1}; its buffer is the one stage N-2 goes to next */
	issue_L(N); issue_L(N-1); issue_B(N-1);
	c.wait(2+(N&1));
	double pbn = Pb_from((N&1) ? c.Lb1 : c.Lb0, N-1);
	if(N>1) { issue_L(N-2); issue_B(N-2); }
	{
	const hb_stage s = d.st[N];
	const int o = c.off_ux(N);
	for(int i=l; i<NX; i+=32) wv[o+i] = rqv[o+i];
	__syncwarp();
	if(qx!=nullptr) for(int j=l; j<s.nb; j+=32) wv[o+d.idxb[s.off_c+j]] += qx[s.off_c+j];
	__syncwarp();
	}
	for(int n=N-1; n>=0; n--)
		{
		const bool first = (n==0);
		const double *Ln = (n&1) ? c.Lb1 : c.Lb0;
		const double *sB = c.io + (n&1)*BAB;
		const int nux = first ? NU : NUX, o = c.off_ux(n), o1 = c.off_ux(n+1) + ((n+1<N) ? NU : 0);
		const hb_stage s = d.st[n];
		for(int i=l; i<nux; i+=32) c.va[i] = rqv[o+i];
		if(l<NX) c.vb[l] = pbn + wv[o1+l];
		__syncwarp();
		if(qx!=nullptr) for(int j=l; j<s.nb; j+=32) c.va[d.idxb[s.off_c+j]] += qx[s.off_c+j];
		__syncwarp();
		c.wait(2+(n&1)); c.wait(n&1);
		/* v = va + [B A]' vb : lane l owns rows l and l+32 */
		double v0 = 0.0, v1 = 0.0;
		if(l<nux)
			{
			double a0 = c.va[l], a1 = 0.0;
			#pragma unroll
			for(int j=0; j<NX; j+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(sB + l*NX + j);
				a0 = fma(t.x, c.vb[j], a0); a1 = fma(t.y, c.vb[j+1], a1);
				}
			v0 = a0+a1;
			}
		if(NUX>32 && l+32<nux)
			{
			double a0 = c.va[l+32], a1 = 0.0;
			#pragma unroll
			for(int j=0; j<NX; j+=2)
				{
				const double2 t = *reinterpret_cast<const double2*>(sB + (l+32)*NX + j);
				a0 = fma(t.x, c.vb[j], a0); a1 = fma(t.y, c.vb[j+1], a1);
				}
			v1 = a0+a1;
			}
		/* forward substitution with the first NU columns of L_n */
		const double di = (l<NU) ? Ln[C::DINV+l] : 0.0;
		#pragma unroll
		for(int j=0; j<NU; j++)
			{
			const double vj = __shfl_sync(HBF_FULL, v0*di, j);
			const double *col = Ln + C::colOff(j) - j;            /* col[r] = L[r][j] in frame rows */
			if(l==j) v0 = vj;
			else if(l>j && l<nux) v0 = fma(-col[l], vj, v0);
			if(NUX>32 && l+32<nux) v1 = fma(-col[l+32], vj, v1);
			}
		if(l<nux) wv[o+l] = v0;
		if(NUX>32 && l+32<nux) wv[o+l+32] = v1;
		__syncwarp();
		if(n>0) pbn = Pb_from(Ln, n-1);          /* while the factor of stage n is still on chip */
		if(n-2>=0) { issue_L(n-2); issue_B(n-2); }
		}
	}


/* ------------------------------------------------------------------------------------------------ */
/* Pb_n = Lxx_{n+1} (Lxx_{n+1}' b_n) for every edge, b taken from the instance block: what a solve with stored  */
/* factors needs first when b is new (d_back_ric_rec_trs_tv_res with compute_Pb = 1, d_back_ric_rec.c:564)       */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__device__ void hbi_Pb_sweep(hbi_ctx<C> &c, const double *__restrict__ in_inst, const double *__restrict__ Lst, double *Pb,
		const double *bv = nullptr /* b_n as vectors (edge n at n*NX) instead of the block's rows */)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX, LBUF = C::LBUF;
	const int l = c.lane, N = c.N;
	auto issue_L = [&](int n) { c.load(2+(n&1), (n&1) ? c.Lb1 : c.Lb0, Lst + (long long)n*LBUF, 8u*LBUF); };
	issue_L(1); if(N>1) issue_L(2);
	int coff = 0;
	#pragma unroll
	for(int jj=0; jj<NX; jj++) if(jj==l) coff = C::colOff(NU+jj);
	for(int n=0; n<N; n++)
		{
		const int m1 = n+1;
		const double *Lp = (m1&1) ? c.Lb1 : c.Lb0;
		const int brow = (n==0) ? NU : NUX;
		if(l<NX) c.va[l] = bv!=nullptr ? bv[n*NX+l] : in_inst[c.off_in(n) + brow*NX + l];
		__syncwarp();
		c.wait(2+(m1&1));
		if(l<NX)
			{
			const double *col = Lp + coff - l;                     /* col[m] = Lxx[m][l] */
			double a0 = 0.0;
			#pragma unroll
			for(int m=0; m<NX; m++) if(m>=l) a0 = fma(col[m], c.va[m], a0);
			c.tmp[l] = a0;
			}
		__syncwarp();
		if(l<NX)
			{
			double p0 = 0.0;
			#pragma unroll
			for(int cc=0; cc<NX; cc++) if(cc<=l) p0 = fma(Lp[C::colOff(NU+cc) + (l-cc)], c.tmp[cc], p0);
			Pb[n*NX+l] = p0;
			}
		__syncwarp();
		if(m1+2<=N) issue_L(m1+2);
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* residuals of the KKT system (res_q, res_b) with the stage data streamed through shared memory.     */
/*   rq0, b0 : original gradient / b (ux / pi layout) ; lamd = lam_up - lam_lo per bound (nbtot)       */
/*   returns max |res_q|, max |res_b| over the lanes' entries (caller reduces)                          */
/* ------------------------------------------------------------------------------------------------ */
template<class C, class X>
__device__ void hbi_residuals(X &c, const hb_dims &d, const double *__restrict__ in_inst, const double *rq0, const double *b0,
		const double *lam_lo, const double *lam_up, const double *ux, const double *pi, double *res_q, double *res_b,
		double &nq, double &nb_)
	{
	constexpr int NX = C::NX, NU = C::NU, NUX = C::NUX;
	const int l = c.lane, N = c.N;
	auto issue = [&](int n)
		{
		const int kind = (n==0) ? HBF_FIRST : (n==N ? HBF_LAST : HBF_MID);
		const uint32_t bB = (n<N) ? hbi_bytes_BAbt<C>(n==0) : 0u;
		if(n<N) c.load(n&1, c.io + (n&1)*X::SB, in_inst + c.off_in(n), bB);
		c.load(2+(n&1), (n&1) ? c.Lb1 : c.Lb0, in_inst + c.off_in(n) + bB/8, hbi_bytes_RSQ<C>(kind));
		};
	issue(0); issue(1);
	for(int n=0; n<=N; n++)
		{
		const int nu = (n==N) ? 0 : NU, nux = (n==0) ? NU : (n==N ? NX : NUX);
		const int o = c.off_ux(n), o1 = c.off_ux(n+1) + ((n+1<N) ? NU : 0);
		const double *sB = c.io + (n&1)*X::SB;
		const double *H = (n&1) ? c.Lb1 : c.Lb0;
		const hb_stage s = d.st[n];
		c.kick(2+(n&1)); if(n<N) c.kick(n&1);
		if(n<N)
			{
			const hb_stage sn = d.st[n+1];
			hbi_prefetch_vec(ux + c.off_ux(n+1), NUX, l); hbi_prefetch_vec(rq0 + c.off_ux(n+1), NUX, l); hbi_prefetch_vec(pi + n*NX, 2*NX, l);
			hbi_prefetch_vec(lam_lo + sn.off_c, sn.nb, l); hbi_prefetch_vec(lam_up + sn.off_c, sn.nb, l); hbi_prefetch_vec(b0 + n*NX, NX, l);
			}
		for(int i=l; i<nux; i+=32)
			{
			c.va[i] = ux[o+i];
			double v = rq0[o+i];
			if(n>0 && i>=nu) v -= pi[(n-1)*NX + (i-nu)];
			c.vc[i] = v;
			}
		if(n<N && l<NX) c.vb[l] = pi[n*NX+l];
		__syncwarp();
		for(int j=l; j<s.nb; j+=32) c.vc[d.idxb[s.off_c+j]] += -lam_lo[s.off_c+j] + lam_up[s.off_c+j];
		__syncwarp();
		c.wait(2+(n&1)); if(n<N) c.wait(n&1);
		#pragma unroll
		for(int slot=0; slot<(NUX>32 ? 2 : 1); slot++)
			{
			const int i = l + 32*slot;
			if(i<nux)
				{
				double a0 = c.vc[i], a1 = 0.0;
				const double *hi = H + HB_TRI(i);
				for(int j=0; j<=i; j++) a0 = fma(hi[j], c.va[j], a0);
				for(int j=i+1; j<nux; j++) a1 = fma(H[HB_TRI(j)+i], c.va[j], a1);
				if(n<N)
					{
					#pragma unroll
					for(int j=0; j<NX; j+=2)
						{
						const double2 t = *reinterpret_cast<const double2*>(sB + i*NX + j);
						a0 = fma(t.x, c.vb[j], a0); a1 = fma(t.y, c.vb[j+1], a1);
						}
					}
				const double r = a0+a1;
				res_q[o+i] = r;
				nq = fmax(nq, fabs(r));
				}
			}
		if(n<N && l<NX)
			{
			double a0 = b0[n*NX+l] - ux[o1+l], a1 = 0.0;
			for(int i=0; i+1<nux; i+=2) { a0 = fma(sB[i*NX+l], c.va[i], a0); a1 = fma(sB[(i+1)*NX+l], c.va[i+1], a1); }
			if(nux&1) a0 = fma(sB[(nux-1)*NX+l], c.va[nux-1], a0);
			const double r = a0+a1;
			res_b[n*NX+l] = r;
			nb_ = fmax(nb_, fabs(r));
			}
		__syncwarp();
		if(n+2<=N) issue(n+2);
		}
	}
