/*
 * cipm_kernels.cu -- the box / general-constraint IPM for chains as a MULTI-KERNEL driver with active-set compaction.
 *
 * The fused kernel of ipm_kernels.cu runs a whole IPM in one warp: one launch per wave, 254 registers for every phase, several
 * hundred KB of SASS that the warps of an SM walk through out of step, and a wave that ends with its slowest instance.  Here the
 * same iteration (mpc_solvers/d_ip2_res_hard.c:116-1345) is cut at its sweeps over the horizon:
 *
 *   hb_cipm_step_kernel   everything element-wise between two sweeps, as a per-instance state machine (init, the two step-length /
 *                         centering computations, the variable update, the loop tests, the barrier terms of the next system);
 *                         no shared memory, 8 warps per CTA; at the end of a round it COMPACTS: instances that are not finished
 *                         are appended to the next round's active list, so every warp of the following kernels has work
 *   hb_cipm_sv_kernel     factor + solve with the IPM hooks      (S::backward + S::forward_sv)   the only 254-register phase;
 *                         BASELINE config 3: hb_cipm_sv2_kernel, two instances per warp on the register-blocked tile (ric_ipm_blk.cuh)
 *   hb_cipm_trs_kernel    solve with the stored factor             (S::trs)
 *   hb_cipm_res_kernel    residuals, mu, exit norms                (S::residuals)
 *
 * A round is sv, step, trs, step, res, step: every active instance advances by one IPM iteration, whatever its phase.  The host
 * enqueues k_max rounds without ever reading anything back; a kernel that finds the active list empty returns at once.
 * Every instance works in its own state block in HBM (the fused kernel's work slot: factor, iterate vectors, constraint vectors).
 * The device functions are the fused kernel's own (ipm_sweeps.cuh, ipm_elem.cuh), executed in the same order per instance, so the
 * results -- iteration counts included -- are bit-identical to the fused path (tests/test_cipm.py).
 */
#include "launch_util.cuh"
#include "ipm_sweeps.cuh"
#include "ric_ipm_blk.cuh"
#include "ric_team.cuh"

enum { CS_INIT=0, CS_P1_SV, CS_P1_A, CS_P1_TRS, CS_P1_B, CS_P2_SV, CS_P2_A, CS_P2_TRS, CS_P2_B,
       CS_RES_ENTER /* residuals wanted before phase 2 starts */, CS_RES_ITER /* ... at the end of a phase-2 iteration */,
       CS_RES_ENTER_DONE, CS_RES_ITER_DONE, CS_DONE };

/* per-instance record: doubles [0] mu [1] alpha [2] sigma [3..5] exit norms ; ints [0] state [1] kk */
#define CIPM_D 6
#define CIPM_I 2

struct hb_cipm_args
	{
	hb_dims d;
	long long n_inst;
	const double *in;
	int k_max; double mu0, mu_tol, alpha_min; int warm_start;
	double *ux, *pi, *lam, *t, *info;
	double *work; long long work_stride;
	double *sd; int *si;                    /* state records */
	const int *act; const int *n_act;       /* active list of this launch (act == nullptr: all instances 0..n_inst-1) */
	int *act_next; int *n_act_next;         /* step kernel, end of a round: the list it builds (nullptr: none) */
	int *wq;                                /* factorisation kernel: work-queue counter (pairs of list entries), zeroed before the launch */
	int team_fwd;                           /* any-size team kernel <0>: 1 = the forward sweep follows in the same kernel, 0 = factorisation only */
	};

/* ---------------------------------------------------------------------------------------------------------------- */
template<class S>
__global__ void __launch_bounds__(256) hb_cipm_step_kernel(hb_cipm_args a)
	{
	const hb_dims &d = a.d;
	const int lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + (threadIdx.x>>5), tw = (long long)gridDim.x*nw;
	const long long n_items = a.act ? (long long)*a.n_act : a.n_inst;
	typename S::ctx_t c;
	c.lane = lane;                                /* extract / emit only look at the lane */
	const int info_len = HB_IPM_INFO_HEAD + 5*a.k_max;
	const double thr0 = 0.1, mu0 = a.mu0, mu_tol = a.mu_tol, alpha_min = a.alpha_min;
	const double mu_tol_low = mu_tol<1e-5 ? 1e-5 : mu_tol;
	const double mu_scal = d.nbtot>0 ? 1.0/(2.0*d.nbtot) : 0.0;
	const int k_max = a.k_max;
	for(long long it=gw; it<n_items; it+=tw)
		{
		const long long inst = a.act ? a.act[it] : it;
		int *si = a.si + inst*CIPM_I; double *sd = a.sd + inst*CIPM_D;
		int st = si[0], kk = si[1];
		double mu = sd[0], alpha = sd[1], sigma = sd[2], mu_aff;
		const hb_ipm_ws w = hb_ipm_make_ws<S>(d, a.work + inst*a.work_stride);
		const double *in_inst = a.in + inst*d.in_stride;
		double *ux = a.ux + inst*d.ux_stride, *pi = a.pi + inst*d.pi_stride;
		double *info = a.info + inst*info_len, *stat = info + HB_IPM_INFO_HEAD;
		bool top1 = false, top2 = false;
		__syncwarp();
		if(st==CS_INIT)
			{
			S::extract(c, d, in_inst, w);
			__syncwarp();
			/* init (c99/d_aux_ip_hard_lib4.c:43-149) */
			if(!a.warm_start) for(long long i=lane; i<d.ux_stride; i+=32) ux[i] = 0.0;
			for(long long i=lane; i<d.pi_stride; i+=32) pi[i] = 0.0;
			__syncwarp();
			for(int cc=lane; cc<d.nbtot; cc+=32)
				{
				const int iu = d.c_ux[cc];
				if(iu<0) continue;
				double lb = w.v(CV_LB)[cc], ub = w.v(CV_UB)[cc], u = ux[iu];
				double tl = -lb + u, tu = ub - u;
				if(tl<thr0)
					{
					if(tu<thr0) { ux[iu] = (-ub + lb)*0.5; tl = thr0; tu = thr0; }
					else { tl = thr0; ux[iu] = lb + thr0; }
					}
				else if(tu<thr0) { tu = thr0; ux[iu] = ub - thr0; }
				w.v(CV_T_LO)[cc] = tl; w.v(CV_T_UP)[cc] = tu;
				w.v(CV_LAM_LO)[cc] = mu0/tl; w.v(CV_LAM_UP)[cc] = mu0/tu;
				}
			__syncwarp();
			if(d.ngtot>0)
				{
				hb_gen_values(lane, d, in_inst, w, ux);
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					if(d.c_ux[cc]>=0) continue;
					const double v = w.v(CV_VAL)[cc];
					const double tl = fmax(thr0, v - w.v(CV_LB)[cc]), tu = fmax(thr0, -v + w.v(CV_UB)[cc]);
					w.v(CV_T_LO)[cc] = tl; w.v(CV_T_UP)[cc] = tu;
					w.v(CV_LAM_LO)[cc] = mu0/tl; w.v(CV_LAM_UP)[cc] = mu0/tu;
					}
				__syncwarp();
				}
			mu = mu0; alpha = 1.0; sigma = 0.0; kk = 0;
			top1 = true;
			}
		else if(st==CS_P1_A || st==CS_P2_A)
			{
			/* after the predictor solve: affine step length, mu_aff, sigma, the corrector's gradient (d_ip2_res_hard.c:560-640, :1010-1100) */
			const bool p2 = (st==CS_P2_A);
			hb_gen_values(lane, d, in_inst, w, w.dux);
			alpha = p2 ? hb_ipm_alpha<true>(lane, d, w, w.dux) : hb_ipm_alpha<false>(lane, d, w, w.dux);
			__syncwarp();
			if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
			alpha *= 0.995;
			mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
			if(lane==0) stat[5*kk+2] = mu_aff;
			sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
			const double sm = sigma*mu;
			if(!p2)
				{
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double dll = w.v(CV_TINV_LO)[cc]*(sm - w.v(CV_DLAM_LO)[cc]*w.v(CV_DT_LO)[cc]);
					double dlu = w.v(CV_TINV_UP)[cc]*(sm - w.v(CV_DLAM_UP)[cc]*w.v(CV_DT_UP)[cc]);
					w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
					w.v(CV_QXG)[cc] += dlu - dll;
					}
				}
			else
				{
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double rml = w.v(CV_RM_LO)[cc] + (w.v(CV_DT_LO)[cc]*w.v(CV_DLAM_LO)[cc] - sm);
					double rmu = w.v(CV_RM_UP)[cc] + (w.v(CV_DT_UP)[cc]*w.v(CV_DLAM_UP)[cc] - sm);
					w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
					w.v(CV_QXG)[cc] = w.v(CV_TINV_LO)[cc]*(rml - w.v(CV_LAM_LO)[cc]*w.v(CV_RD_LO)[cc])
					                  - w.v(CV_TINV_UP)[cc]*(rmu + w.v(CV_LAM_UP)[cc]*w.v(CV_RD_UP)[cc]);
					}
				}
			st = p2 ? CS_P2_TRS : CS_P1_TRS;
			}
		else if(st==CS_P1_B)
			{
			/* after the corrector solve: step length, update_var, mu (c99/d_aux_ip_hard_lib4.c:489-711) */
			hb_gen_values(lane, d, in_inst, w, w.dux);
			alpha = hb_ipm_alpha<false>(lane, d, w, w.dux);
			__syncwarp();
			if(lane==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
			alpha *= 0.995;
			for(long long i=lane; i<d.ux_stride; i+=32) ux[i] += alpha*(w.dux[i] - ux[i]);
			for(long long i=lane; i<d.pi_stride; i+=32) pi[i] += alpha*(w.dpi[i] - pi[i]);
			double ms = 0.0;
			for(int cc=lane; cc<d.nbtot; cc+=32)
				{
				double ll = w.v(CV_LAM_LO)[cc] + alpha*w.v(CV_DLAM_LO)[cc];
				double lu = w.v(CV_LAM_UP)[cc] + alpha*w.v(CV_DLAM_UP)[cc];
				double tl = w.v(CV_T_LO)[cc] + alpha*w.v(CV_DT_LO)[cc];
				double tu = w.v(CV_T_UP)[cc] + alpha*w.v(CV_DT_UP)[cc];
				w.v(CV_LAM_LO)[cc] = ll; w.v(CV_LAM_UP)[cc] = lu; w.v(CV_T_LO)[cc] = tl; w.v(CV_T_UP)[cc] = tu;
				ms += ll*tl + lu*tu;
				}
			mu = hb_warp_sum(ms)*mu_scal;
			if(lane==0) stat[5*kk+4] = mu;
			kk++;
			top1 = true;
			}
		else if(st==CS_P2_B)
			{
			/* phase 2: step length and update (c99/d_aux_ip_hard_lib4.c:1180-1449); the residuals decide what comes next */
			hb_gen_values(lane, d, in_inst, w, w.dux);
			alpha = hb_ipm_alpha<true>(lane, d, w, w.dux);
			__syncwarp();
			if(lane==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
			alpha *= 0.995;
			for(long long i=lane; i<d.ux_stride; i+=32) ux[i] += alpha*w.dux[i];
			for(long long i=lane; i<d.pi_stride; i+=32) pi[i] += alpha*w.dpi[i];
			for(int cc=lane; cc<d.nbtot; cc+=32)
				{
				w.v(CV_LAM_LO)[cc] += alpha*w.v(CV_DLAM_LO)[cc]; w.v(CV_LAM_UP)[cc] += alpha*w.v(CV_DLAM_UP)[cc];
				w.v(CV_T_LO)[cc] += alpha*w.v(CV_DT_LO)[cc]; w.v(CV_T_UP)[cc] += alpha*w.v(CV_DT_UP)[cc];
				}
			st = CS_RES_ITER;
			}
		else if(st==CS_RES_ENTER_DONE) top2 = true;
		else if(st==CS_RES_ITER_DONE)
			{
			if(lane==0) stat[5*kk+4] = mu;
			kk++;
			top2 = true;
			}
		__syncwarp();
		if(top1)
			{
			/* top of the phase-1 loop (d_ip2_res_hard.c:503) */
			if(kk<k_max && mu>mu_tol_low && alpha>=alpha_min)
				{
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double til = 1.0/w.v(CV_T_LO)[cc], tiu = 1.0/w.v(CV_T_UP)[cc];
					double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc];
					double ltl = ll*til, ltu = lu*tiu;
					double dll = til*0.0, dlu = tiu*0.0;
					w.v(CV_TINV_LO)[cc] = til; w.v(CV_TINV_UP)[cc] = tiu;
					w.v(CV_LAMT_LO)[cc] = ltl; w.v(CV_LAMT_UP)[cc] = ltu;
					w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
					w.v(CV_QXD)[cc] = ltl + ltu;
					w.v(CV_QXG)[cc] = lu - ltu*w.v(CV_UB)[cc] + dlu - ll - ltl*w.v(CV_LB)[cc] - dll;
					}
				st = CS_P1_SV;
				}
			else st = CS_RES_ENTER;
			}
		if(top2)
			{
			/* top of the phase-2 loop (d_ip2_res_hard.c:783) */
			if(kk<k_max && mu>mu_tol && alpha>=alpha_min)
				{
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double til = 1.0/w.v(CV_T_LO)[cc], tiu = 1.0/w.v(CV_T_UP)[cc];
					double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc];
					w.v(CV_TINV_LO)[cc] = til; w.v(CV_TINV_UP)[cc] = tiu;
					w.v(CV_QXD)[cc] = til*ll + tiu*lu;
					w.v(CV_QXG)[cc] = til*(w.v(CV_RM_LO)[cc] - ll*w.v(CV_RD_LO)[cc]) - tiu*(w.v(CV_RM_UP)[cc] + lu*w.v(CV_RD_UP)[cc]);
					}
				st = CS_P2_SV;
				}
			else
				{
				int status;
				if(mu<=mu_tol) status = 0;
				else if(kk>=k_max) status = 1;
				else if(alpha<alpha_min) status = 2;
				else status = -1;
				S::emit(c, d, w, a.lam + inst*2*(long long)d.nbtot, a.t + inst*2*(long long)d.nbtot);
				if(lane==0)
					{
					info[0] = (double)kk; info[1] = (double)status;
					info[2] = sd[3]; info[3] = sd[4]; info[4] = sd[5]; info[5] = mu;
					}
				st = CS_DONE;
				}
			}
		__syncwarp();
		if(lane==0)
			{
			si[0] = st; si[1] = kk;
			sd[0] = mu; sd[1] = alpha; sd[2] = sigma;
			if(a.act_next!=nullptr && st!=CS_DONE) a.act_next[atomicAdd(a.n_act_next, 1)] = (int)inst;      /* compaction */
			}
		__syncwarp();
		}
	}

/* ---------------------------------------------------------------------------------------------------------------- */
/* which: 0 = factor + solve (predictor), 1 = solve with the stored factor (corrector), 2 = residuals                 */
template<class S, int WHICH>
__global__ void __launch_bounds__(256) hb_cipm_sweep_kernel(hb_cipm_args a)
	{
	const hb_dims &d = a.d;
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	const long long n_items = a.act ? (long long)*a.n_act : a.n_inst;
	if(n_items==0) return;
	typename S::ctx_t c;
	S::init(c, d, hb_smem + (size_t)warp*S::smem_doubles(d), lane);
	for(long long it=gw; it<n_items; it+=tw)
		{
		const long long inst = a.act ? a.act[it] : it;
		int *si = a.si + inst*CIPM_I; double *sd = a.sd + inst*CIPM_D;
		const int st = si[0];
		const hb_ipm_ws w = hb_ipm_make_ws<S>(d, a.work + inst*a.work_stride);
		const double *in_inst = a.in + inst*d.in_stride;
		__syncwarp();
		if constexpr(WHICH==0)
			{
			if(st!=CS_P1_SV && st!=CS_P2_SV) continue;
			const bool p2 = (st==CS_P2_SV);
			S::backward(c, d, in_inst, w, p2 ? w.res_b : nullptr, p2 ? w.res_q : nullptr, w.v(CV_QXD), w.v(CV_QXG));
			__syncwarp();
			S::forward_sv(c, d, in_inst, w, p2 ? w.res_b : nullptr, w.dux, w.dpi);
			__syncwarp();
			if(lane==0) si[0] = p2 ? CS_P2_A : CS_P1_A;
			}
		else if constexpr(WHICH==3)
			{
			/* forward sweep of the predictor, after a factorisation-only hb_cipm_sv2_kernel */
			if(st!=CS_P1_SV && st!=CS_P2_SV) continue;
			const bool p2 = (st==CS_P2_SV);
			S::forward_sv(c, d, in_inst, w, p2 ? w.res_b : nullptr, w.dux, w.dpi);
			__syncwarp();
			if(lane==0) si[0] = p2 ? CS_P2_A : CS_P1_A;
			}
		else if constexpr(WHICH==1)
			{
			if(st!=CS_P1_TRS && st!=CS_P2_TRS) continue;
			const bool p2 = (st==CS_P2_TRS);
			S::trs(c, d, in_inst, w, p2 ? w.res_b : w.b0, p2 ? w.res_q : w.rq0, w.v(CV_QXG));
			__syncwarp();
			if(lane==0) si[0] = p2 ? CS_P2_B : CS_P1_B;
			}
		else
			{
			if(st!=CS_RES_ENTER && st!=CS_RES_ITER) continue;
			double mu = sd[0], norms[3] = {0.0, 0.0, 0.0};
			S::residuals(c, d, in_inst, w, a.ux + inst*d.ux_stride, a.pi + inst*d.pi_stride, &mu, norms);
			__syncwarp();
			if(lane==0)
				{
				sd[0] = mu; sd[3] = norms[0]; sd[4] = norms[1]; sd[5] = norms[2];
				si[0] = (st==CS_RES_ENTER) ? CS_RES_ENTER_DONE : CS_RES_ITER_DONE;
				}
			}
		__syncwarp();
		}
	}

/* ---------------------------------------------------------------------------------------------------------------- */
/* factor + solve (predictor) with the register-blocked factorisation sweep of ric_ipm_blk.cuh: two instances per warp.  The    */
/* forward sweep that follows is the one-instance-per-warp routine, run for the two instances in turn on their own regions.     */
template<class C, bool FWD>
__global__ void __launch_bounds__(128, 1) hb_cipm_sv2_kernel(hb_cipm_args a)
	{
	typedef hb_sweeps_fast<C> S;
	const hb_dims &d = a.d;
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5, g = lane>>4;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	const long long n_items = a.act ? (long long)*a.n_act : a.n_inst;
	if(n_items==0) return;
	/* FWD: two regions in the layout of hbi_ctx, so that the forward sweep can follow; else the slim factorisation-only layout */
	double *wbase = hb_smem + (size_t)warp*(FWD ? hbi2_cfg<C>::PER_WARP : hbi2_cfg<C>::PER_WARP_SLIM);
	hbi_ctx<C> c0, c1;
	if constexpr(FWD) { c0.init(wbase, lane, d); c1.init(wbase + hbi2_cfg<C>::PW, lane, d); }
	hbi2_ctx<C, !FWD> c2;
	c2.init(wbase, lane, d);
	/* pairs are handed out by an atomic counter (no idle warps behind a slow pair at the end of the list) */
	(void)gw; (void)tw;
	for(;;)
		{
		long long it = 0;
		if(lane==0) it = (long long)atomicAdd(a.wq, 2);
		it = __shfl_sync(0xffffffffu, it, 0);
		if(it>=n_items) break;
		const long long i1 = it+1<n_items ? it+1 : it;
		long long inst0 = a.act ? a.act[it] : it, inst1 = a.act ? a.act[i1] : i1;
		int st0 = a.si[inst0*CIPM_I], st1 = a.si[inst1*CIPM_I];
		const bool v0 = (st0==CS_P1_SV || st0==CS_P2_SV), v1 = (st1==CS_P1_SV || st1==CS_P2_SV) && inst1!=inst0;
		if(!v0 && !v1) continue;
		/* a pair with one idle member works twice on the other one (same data, same results, same addresses) */
		if(!v0) { inst0 = inst1; st0 = st1; }
		if(!v1) { inst1 = inst0; st1 = st0; }
		__syncwarp();
		{
		const long long inst = g ? inst1 : inst0;
		const bool p2 = ((g ? st1 : st0)==CS_P2_SV);
		const hb_ipm_ws w = hb_ipm_make_ws<S>(d, a.work + inst*a.work_stride);
		hbi2_backward<C, !FWD>(c2, d, a.in + inst*d.in_stride, w.L, p2 ? w.res_b : nullptr, p2 ? w.res_q : nullptr, w.v(CV_QXD), w.v(CV_QXG), w.Pb);
		}
		__syncwarp();
		if constexpr(FWD)
			{
			{
			const bool p2 = (st0==CS_P2_SV);
			const hb_ipm_ws w = hb_ipm_make_ws<S>(d, a.work + inst0*a.work_stride);
			S::forward_sv(c0, d, a.in + inst0*d.in_stride, w, p2 ? w.res_b : nullptr, w.dux, w.dpi);
			}
			if(inst1!=inst0)
				{
				const bool p2 = (st1==CS_P2_SV);
				const hb_ipm_ws w = hb_ipm_make_ws<S>(d, a.work + inst1*a.work_stride);
				__syncwarp();
				S::forward_sv(c1, d, a.in + inst1*d.in_stride, w, p2 ? w.res_b : nullptr, w.dux, w.dpi);
				}
			__syncwarp();
			if(lane==0)
				{
				a.si[inst0*CIPM_I] = (st0==CS_P2_SV) ? CS_P2_A : CS_P1_A;
				if(inst1!=inst0) a.si[inst1*CIPM_I] = (st1==CS_P2_SV) ? CS_P2_A : CS_P1_A;
				}
			}
		/* !FWD: the state stays *_SV; the forward kernel (hb_cipm_sweep_kernel<S1, 3>, 16 warps per SM) picks the instance up */
		__syncwarp();
		}
	}

/* which kernel runs the factor + solve sweep of a round */
template<class S> struct hb_cipm_sv
	{
	static int prep(int) { return 0; }
	static bool use(const hb_dims &) { return false; }
	static bool has_trs() { return false; }
	static void launch(const hb_cipm_args &, int, cudaStream_t, bool) {}
	static void launch_trs(const hb_cipm_args &, int, cudaStream_t) {}
	static void launch_res(const hb_cipm_args &, int, cudaStream_t) {}
	};
template<> struct hb_cipm_sv<hb_sweeps_fast<hbi_v0> >
	{
	typedef hbi_v0 C;
	static constexpr int WARPS = 4, WARPS_SLIM = 4;      /* a fifth warp fits the slim layout but adds nothing (measured: 99.0 vs 98.6 K solves/s) */
	static int smem() { return WARPS*(int)sizeof(double)*hbi2_cfg<C>::PER_WARP; }
	static int smem_slim() { return WARPS_SLIM*(int)sizeof(double)*hbi2_cfg<C>::PER_WARP_SLIM; }
	static int prep(int) { return hb_prep(hb_cipm_sv2_kernel<C, true>, smem()) || hb_prep(hb_cipm_sv2_kernel<C, false>, smem_slim()); }
	/* HPMPC_B200_IPM_SV2=0 keeps the one-instance-per-warp sweep (A/B runs) */
	static bool use(const hb_dims &) { const char *e = getenv("HPMPC_B200_IPM_SV2"); return !(e && e[0]=='0'); }
	static bool has_trs() { return false; }
	static void launch_trs(const hb_cipm_args &, int, cudaStream_t) {}
	static void launch_res(const hb_cipm_args &, int, cudaStream_t) {}
	static void launch(const hb_cipm_args &a, int sms, cudaStream_t st, bool fwd)
		{
		const int w = fwd ? WARPS : WARPS_SLIM;
		long long need = (a.n_inst + 2*w - 1)/(2*w);
		const int grid = (int)(need<sms ? (need<1 ? 1 : need) : sms);
		if(fwd) hb_cipm_sv2_kernel<C, true><<<grid, w*32, smem(), st>>>(a);
		else hb_cipm_sv2_kernel<C, false><<<grid, w*32, smem_slim(), st>>>(a);
		}
	};

/* residuals by the four-warp team: hb_ipm_residuals (mpc_solvers/c99/d_res_ip_res_hard.c:39-319, exit norms of
 * interfaces/c/fortran_order_interface.c:616-652) with the two matrix-vector products of a stage split over the warps (hbt_dot) and
 * the constraint loops over all 128 threads; mu and the norms are reduced warp by warp, then over the four warps in a fixed order */
static __device__ void hbt_ipm_residuals(const hb_ctx &c, int tid, double *P, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
		const double *ux, const double *pi, double *mu, double *norms)
	{
	const int ld = c.ldW;
	double nq = 0.0, nb_ = 0.0, nd = 0.0, mu2 = 0.0;
	const double *lam_lo = w.v(CV_LAM_LO), *lam_up = w.v(CV_LAM_UP);
	if(d.ngtot>0)
		{
		hb_gen_values_part(tid, HBT_THREADS, d, in_inst, w, ux);
		hbt_sync();
		}
	hb_ipm_residuals_bounds_part(tid, HBT_THREADS, d, w, ux, mu2, nd);
	double *xs = c.sV, *ps = c.sV + 64;
	double *H = c.bufA;
	const double *sW = c.sW;
	for(int n=0; n<=d.N; n++)
		{
		const hb_stage s = d.st[n];
		const int nu = s.nu, nux = s.nu+s.nx, nx1 = s.nx1;
		hbt_g2s(tid, H, in_inst + s.off_RSQ, HB_TRI(nux));
		hbt_load_BAbt(c, tid, s, in_inst);
		for(int i=tid; i<nux; i+=HBT_THREADS) xs[i] = ux[s.off_ux+i];
		for(int j=tid; j<nx1; j+=HBT_THREADS) ps[j] = pi[s.off_pi+j];
		for(int i=tid; i<nux; i+=HBT_THREADS)
			{
			double v = w.rq0[s.off_ux+i];
			if(n>0 && i>=nu) v -= pi[d.st[n-1].off_pi + (i-nu)];
			w.res_q[s.off_ux+i] = v;
			}
		hbt_sync();
		for(int j=tid; j<s.nb; j+=HBT_THREADS)
			w.res_q[s.off_ux+d.idxb[s.off_c+j]] += -lam_lo[s.off_c+j] + lam_up[s.off_c+j];
		if(s.ng>0)
			{
			hbt_sync();
			const double *G = in_inst + s.off_DCt;
			const int cg = s.off_c + s.nb;
			for(int i=tid; i<nux; i+=HBT_THREADS)
				{
				double acc = w.res_q[s.off_ux+i];
				for(int j=0; j<s.ng; j++) acc += G[i*s.ng+j]*(lam_up[cg+j] - lam_lo[cg+j]);
				w.res_q[s.off_ux+i] = acc;
				}
			}
		hb_g2s_wait();
		hbt_sync();
		hbt_dot(tid, nux, P, [&](int) { return 0; }, [&](int) { return nux + nx1; },
				[&](int o, int k) { return k<nux ? (k<=o ? H[HB_TRI(o)+k] : H[HB_TRI(k)+o]) : sW[o*ld + (k-nux)]; },
				[&](int, int k) { return k<nux ? xs[k] : ps[k-nux]; },
				[&](int o, double sum)
					{
					const double acc = w.res_q[s.off_ux+o] + sum;
					w.res_q[s.off_ux+o] = acc;
					nq = fmax(nq, fabs(acc));
					});
		if(nx1>0)
			{
			const hb_stage s1 = d.st[n+1];
			hbt_dot(tid, nx1, P, [&](int) { return 0; }, [&](int) { return nux; },
					[&](int o, int k) { return sW[k*ld+o]; }, [&](int, int k) { return xs[k]; },
					[&](int o, double sum)
						{
						const double acc = (w.b0[s.off_pi+o] - ux[s1.off_ux+s1.nu+o]) + sum;
						w.res_b[s.off_pi+o] = acc;
						nb_ = fmax(nb_, fabs(acc));
						});
			}
		}
	/* warp results, then the four warps in order */
	mu2 = hb_warp_sum(mu2); nq = hb_warp_max(nq); nb_ = hb_warp_max(nb_); nd = hb_warp_max(nd);
	if((tid&31)==0) { P[4*(tid>>5)] = mu2; P[4*(tid>>5)+1] = nq; P[4*(tid>>5)+2] = nb_; P[4*(tid>>5)+3] = nd; }
	hbt_sync();
	if(d.nbtot>0) *mu = ((P[0] + P[4]) + (P[8] + P[12]))/(2.0*d.nbtot);
	norms[0] = fmax(fmax(P[1], P[5]), fmax(P[9], P[13]));
	norms[1] = fmax(fmax(P[2], P[6]), fmax(P[10], P[14]));
	norms[2] = fmax(fmax(P[3], P[7]), fmax(P[11], P[15]));
	hbt_sync();
	}

/* any-size patterns: the sweeps with four warps per instance (ric_team.cuh), one CTA per instance.  WHICH 0: factor + solve
 * (predictor) or factorisation only, instances handed out by an atomic counter; 1: solve with the stored factor (corrector);
 * 2: residuals; 3: the predictor's forward sweep behind a factorisation-only <0> */
extern "C" int hbt_smem_bytes(const hb_dims *d);
extern "C" int hbt_wanted(const hb_dims *d);
template<int WHICH>
__global__ void __launch_bounds__(HBT_THREADS, WHICH==0 ? 5 : 7) hb_cipm_team_kernel(hb_cipm_args a)
	{
	typedef hb_sweeps_generic S;
	const hb_dims &d = a.d;
	const int tid = threadIdx.x;
	const long long n_items = a.act ? (long long)*a.n_act : a.n_inst;
	if(n_items==0) return;
	/* the kernels that do not factorise keep one factor buffer (30 KB instead of 41 KB of stage data: more CTAs per SM) */
	double *P = hb_smem + hb_smem_doubles_per_warp(d.nzM, d.nxM);
	hb_ctx c = (WHICH==0) ? hb_make_ctx(d, hb_smem, tid&31) : hbt_make_ctx1(d, hb_smem, tid&31, P);
	__shared__ int s_it;
	long long it = (long long)blockIdx.x - gridDim.x;
	for(;;)
		{
		if(WHICH==0)
			{
			if(tid==0) s_it = atomicAdd(a.wq, 1);
			__syncthreads();
			it = s_it;
			__syncthreads();
			}
		else it += gridDim.x;
		if(it>=n_items) break;
		const long long inst = a.act ? a.act[it] : it;
		const int st = a.si[inst*CIPM_I];
		const hb_ipm_ws w = hb_ipm_make_ws<S>(d, a.work + inst*a.work_stride);
		const double *in_inst = a.in + inst*d.in_stride;
		if(WHICH==0)
			{
			if(st!=CS_P1_SV && st!=CS_P2_SV) continue;
			const bool p2 = (st==CS_P2_SV);
			hbt_backward<true>(c, tid, d, in_inst, w.L, p2 ? w.res_b : nullptr, p2 ? w.res_q : w.rq0, w.v(CV_QXD), w.v(CV_QXG), w.Pb);
			if(a.team_fwd)
				{
				hbt_forward(c, tid, P, d, in_inst, w.L, nullptr, p2 ? w.res_b : nullptr, false, w.dux, w.dpi, true);
				if(tid==0) a.si[inst*CIPM_I] = p2 ? CS_P2_A : CS_P1_A;
				}
			/* else: the state stays *_SV and hb_cipm_team_kernel<3> (one factor buffer, seven CTAs per SM) picks the instance up */
			}
		else if(WHICH==3)
			{
			if(st!=CS_P1_SV && st!=CS_P2_SV) continue;
			const bool p2 = (st==CS_P2_SV);
			hbt_forward1(c, tid, P, d, in_inst, w.L, nullptr, p2 ? w.res_b : nullptr, false, w.dux, w.dpi, true);
			if(tid==0) a.si[inst*CIPM_I] = p2 ? CS_P2_A : CS_P1_A;
			}
		else if(WHICH==1)
			{
			if(st!=CS_P1_TRS && st!=CS_P2_TRS) continue;
			const bool p2 = (st==CS_P2_TRS);
			const double *bv = p2 ? w.res_b : w.b0, *rqv = p2 ? w.res_q : w.rq0;
			hbt_trs_backward(c, tid, P, d, in_inst, w.L, bv, rqv, w.v(CV_QXG), w.dux, w.Pb, false);
			hbt_forward1(c, tid, P, d, in_inst, w.L, w.dux, bv, true, w.dux, w.dpi, true);
			if(tid==0) a.si[inst*CIPM_I] = p2 ? CS_P2_B : CS_P1_B;
			}
		else
			{
			if(st!=CS_RES_ENTER && st!=CS_RES_ITER) continue;
			double *sd = a.sd + inst*CIPM_D;
			double mu = sd[0], norms[3] = {0.0, 0.0, 0.0};
			hbt_ipm_residuals(c, tid, P, d, in_inst, w, a.ux + inst*d.ux_stride, a.pi + inst*d.pi_stride, &mu, norms);
			if(tid==0)
				{
				sd[0] = mu; sd[3] = norms[0]; sd[4] = norms[1]; sd[5] = norms[2];
				a.si[inst*CIPM_I] = (st==CS_RES_ENTER) ? CS_RES_ENTER_DONE : CS_RES_ITER_DONE;
				}
			}
		__syncthreads();
		}
	}
template<> struct hb_cipm_sv<hb_sweeps_generic>
	{
	/* the kernels have a static shared word of their own (the queue ticket), so the dynamic maximum is set below the architectural one */
	static int prep(int)
		{
		HB_CK(cudaFuncSetAttribute(hb_cipm_team_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226*1024));
		HB_CK(cudaFuncSetAttribute(hb_cipm_team_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226*1024));
		HB_CK(cudaFuncSetAttribute(hb_cipm_team_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226*1024));
		HB_CK(cudaFuncSetAttribute(hb_cipm_team_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226*1024));
		return 0;
		}
	/* HPMPC_B200_TEAM=0 keeps the one-warp-per-instance sweeps (A/B runs, bit-identical to the fused kernel); small stages stay there
	 * by default (hbt_wanted, ric_kernels.cu) */
	static bool use(const hb_dims &d) { return hbt_wanted(&d)!=0; }
	static bool has_trs() { return true; }
	static int grid(const hb_cipm_args &a, int sms, int smem)
		{
		int per_sm = 232448/(smem + 1024);
		if(per_sm>16) per_sm = 16;
		if(per_sm<1) per_sm = 1;
		long long g = (long long)sms*per_sm;
		if(g>a.n_inst) g = a.n_inst;
		return (int)(g<1 ? 1 : g);
		}
	/* HPMPC_B200_TEAM_FWD=1 keeps the predictor's forward sweep inside the factorisation kernel (A/B runs) */
	static void launch(const hb_cipm_args &a0, int sms, cudaStream_t st, bool)
		{
		hb_cipm_args a = a0;
		const char *e = getenv("HPMPC_B200_TEAM_FWD");
		a.team_fwd = (e && e[0]=='1') ? 1 : 0;
		const int smem = hbt_smem_bytes(&a.d);
		hb_cipm_team_kernel<0><<<grid(a, sms, smem), HBT_THREADS, smem, st>>>(a);
		if(!a.team_fwd)
			{
			const int smem1 = (int)sizeof(double)*hbt_smem1_doubles(a.d.nzM, a.d.nxM);
			hb_cipm_team_kernel<3><<<grid(a, sms, smem1), HBT_THREADS, smem1, st>>>(a);
			}
		}
	static void launch_trs(const hb_cipm_args &a, int sms, cudaStream_t st)
		{
		const int smem = (int)sizeof(double)*hbt_smem1_doubles(a.d.nzM, a.d.nxM);
		hb_cipm_team_kernel<1><<<grid(a, sms, smem), HBT_THREADS, smem, st>>>(a);
		}
	static void launch_res(const hb_cipm_args &a, int sms, cudaStream_t st)
		{
		const int smem = (int)sizeof(double)*hbt_smem1_doubles(a.d.nzM, a.d.nxM);
		hb_cipm_team_kernel<2><<<grid(a, sms, smem), HBT_THREADS, smem, st>>>(a);
		}
	};

/* which sweep set runs the solve-only and residual kernels, and at what launch shape: the size-specialised sweeps have a
 * single-slot twin (half the shared memory) that runs 16 warps per SM */
template<class S> struct hb_cipm_light
	{
	typedef S type;
	static bool use() { return false; }
	static int smem(int) { return 0; }
	};
template<class C> struct hb_cipm_light<hb_sweeps_fast<C> >
	{
	typedef hb_sweeps_fast1<C> type;
	static constexpr int WARPS = 8;
	/* HPMPC_B200_IPM_LIGHT=0 keeps the double-buffered sweeps at 8 warps per SM (A/B runs) */
	static bool use() { const char *e = getenv("HPMPC_B200_IPM_LIGHT"); return !(e && e[0]=='0') && 2*(smem(WARPS)+1024)<=233472; }
	static int smem(int warps) { return warps*(int)sizeof(double)*hbi_cfg1<C>::PER_WARP; }
	};

/* ---------------------------------------------------------------------------------------------------------------- */
template<class S> static int hb_cipm_run(int smem_sweep, const hb_cipm_args &base, int *lists, int *counters, int grid_sweep, int warps_sweep,
		int sms, cudaStream_t st)
	{
	/* lists: 2 x n_inst ints ; counters: 2 ints */
	hb_cipm_args a = base;
	const long long n = a.n_inst;
	const int step_warps = 8;
	long long need = (n + step_warps - 1)/step_warps;
	const int grid_step = (int)(need < (long long)sms*8 ? (need<1 ? 1 : need) : (long long)sms*8);
	if(hb_prep(hb_cipm_sweep_kernel<S, 0>, smem_sweep) || hb_prep(hb_cipm_sweep_kernel<S, 1>, smem_sweep) || hb_prep(hb_cipm_sweep_kernel<S, 2>, smem_sweep)) return -1;
	const bool sv2 = hb_cipm_sv<S>::use(a.d);
	if(sv2 && hb_cipm_sv<S>::prep(0)) return -1;
	typedef typename hb_cipm_light<S>::type S1;
	const bool light = hb_cipm_light<S>::use();
	int smem_l = smem_sweep, warps_l = warps_sweep, grid_l = grid_sweep;
	if(light)
		{
		warps_l = 8; smem_l = hb_cipm_light<S>::smem(warps_l);
		long long need_l = (n + warps_l - 1)/warps_l;
		grid_l = (int)(need_l < 2LL*sms ? (need_l<1 ? 1 : need_l) : 2LL*sms);
		if(hb_prep(hb_cipm_sweep_kernel<S1, 1>, smem_l) || hb_prep(hb_cipm_sweep_kernel<S1, 2>, smem_l) || hb_prep(hb_cipm_sweep_kernel<S1, 3>, smem_l)) return -1;
		}
	/* the predictor's forward sweep leaves the two-instances-per-warp kernel (4 warps per SM) for the 16-warp launch shape */
	const bool split_fw = sv2 && light;
	int *act[2] = { lists, lists + n }, *cnt[2] = { counters, counters + 1 };
	a.wq = counters + 2;
	HB_CK(cudaMemsetAsync(a.si, 0, sizeof(int)*CIPM_I*(size_t)n, st));          /* every instance starts in CS_INIT */
	HB_CK(cudaMemsetAsync(counters, 0, 2*sizeof(int), st));
	/* init: all instances, builds list 0 */
	a.act = nullptr; a.n_act = nullptr; a.act_next = act[0]; a.n_act_next = cnt[0];
	hb_cipm_step_kernel<S><<<grid_step, step_warps*32, 0, st>>>(a);
	/* instances whose first loop test already fails (mu0 below the phase-1 threshold) want residuals before anything else */
	a.act = act[0]; a.n_act = cnt[0]; a.act_next = nullptr; a.n_act_next = nullptr;
	const bool team_res = sv2 && hb_cipm_sv<S>::has_trs();
	if(team_res) hb_cipm_sv<S>::launch_res(a, sms, st);
	else if(light) hb_cipm_sweep_kernel<S1, 2><<<grid_l, warps_l*32, smem_l, st>>>(a);
	else hb_cipm_sweep_kernel<S, 2><<<grid_sweep, warps_sweep*32, smem_sweep, st>>>(a);
	hb_cipm_step_kernel<S><<<grid_step, step_warps*32, 0, st>>>(a);
	for(int r=0; r<a.k_max; r++)
		{
		const int cur = r&1, nxt = cur^1;
		a.act = act[cur]; a.n_act = cnt[cur]; a.act_next = nullptr; a.n_act_next = nullptr;
		if(sv2) { HB_CK(cudaMemsetAsync(a.wq, 0, sizeof(int), st)); hb_cipm_sv<S>::launch(a, sms, st, !split_fw); }
		else hb_cipm_sweep_kernel<S, 0><<<grid_sweep, warps_sweep*32, smem_sweep, st>>>(a);
		if(split_fw) hb_cipm_sweep_kernel<S1, 3><<<grid_l, warps_l*32, smem_l, st>>>(a);
		hb_cipm_step_kernel<S><<<grid_step, step_warps*32, 0, st>>>(a);
		if(sv2 && hb_cipm_sv<S>::has_trs()) hb_cipm_sv<S>::launch_trs(a, sms, st);
		else if(light) hb_cipm_sweep_kernel<S1, 1><<<grid_l, warps_l*32, smem_l, st>>>(a);
		else hb_cipm_sweep_kernel<S, 1><<<grid_sweep, warps_sweep*32, smem_sweep, st>>>(a);
		hb_cipm_step_kernel<S><<<grid_step, step_warps*32, 0, st>>>(a);
		if(team_res) hb_cipm_sv<S>::launch_res(a, sms, st);
		else if(light) hb_cipm_sweep_kernel<S1, 2><<<grid_l, warps_l*32, smem_l, st>>>(a);
		else hb_cipm_sweep_kernel<S, 2><<<grid_sweep, warps_sweep*32, smem_sweep, st>>>(a);
		HB_CK(cudaMemsetAsync(cnt[nxt], 0, sizeof(int), st));
		a.act_next = act[nxt]; a.n_act_next = cnt[nxt];
		hb_cipm_step_kernel<S><<<grid_step, step_warps*32, 0, st>>>(a);
		}
	HB_CK(cudaGetLastError());
	return 0;
	}

/* work: n_inst state blocks of work_stride doubles ; aux: n_inst*(CIPM_D doubles) + ints (2 lists, records, 2 counters) */
extern "C" long long hb_cipm_aux_bytes(long long n_inst)
	{
	return (long long)sizeof(double)*CIPM_D*n_inst + (long long)sizeof(int)*(2*n_inst + CIPM_I*n_inst + 8) + 64;
	}

extern "C" int hb_launch_cipm(const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *ux, double *pi, double *lam, double *t, double *info, double *work, long long work_stride, void *aux,
		int grid, int warps, int sms, int fast_id, void *stream)
	{
	if(d->nzM>64) return -2;
	if(warps>8 || d->nbtot<=0 || n_inst>2000000000LL) return -3;
	cudaStream_t st = (cudaStream_t)stream;
	hb_cipm_args a;
	a.d = *d; a.n_inst = n_inst; a.in = in; a.k_max = k_max; a.mu0 = mu0; a.mu_tol = mu_tol; a.alpha_min = alpha_min; a.warm_start = warm_start;
	a.ux = ux; a.pi = pi; a.lam = lam; a.t = t; a.info = info; a.work = work; a.work_stride = work_stride;
	a.sd = (double*)aux;
	int *ip = (int*)((char*)aux + sizeof(double)*CIPM_D*(size_t)n_inst);
	a.si = ip; ip += CIPM_I*n_inst;
	int *lists = ip; ip += 2*n_inst;
	int *counters = ip;
	a.act = nullptr; a.n_act = nullptr; a.act_next = nullptr; a.n_act_next = nullptr; a.team_fwd = 1;
	long long need = (n_inst + warps - 1)/warps;
	if(need<grid) grid = (int)(need<1 ? 1 : need);
	switch(fast_id)
		{
		case 0: return hb_cipm_run<hb_sweeps_fast<hbi_v0> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v0>::PER_WARP, a, lists, counters, grid, warps, sms, st);
		case 1: return hb_cipm_run<hb_sweeps_fast<hbi_v1> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v1>::PER_WARP, a, lists, counters, grid, warps, sms, st);
		case 2: return hb_cipm_run<hb_sweeps_fast<hbi_v2> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v2>::PER_WARP, a, lists, counters, grid, warps, sms, st);
		}
	return hb_cipm_run<hb_sweeps_generic>(warps*hb_smem_bytes_per_warp(d), a, lists, counters, grid, warps, sms, st);
	}
