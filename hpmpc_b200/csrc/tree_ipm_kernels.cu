/*
 * tree_ipm_kernels.cu -- the scenario-tree box IPM as a multi-kernel driver around the size-specialised tree Riccati.
 *
 * hb_ipm_kernel<hb_sweeps_tree> (ipm_kernels.cu) runs a whole tree IPM in one warp with run-time-size node routines.  Here the
 * same iteration (mpc_solvers/d_tree_ip2_res_hard_libstr.c:80 = d_ip2_res_hard.c:116 with the tree Riccati) is cut at its
 * linear solves: hb_tipm_step_kernel executes, one warp per tree, everything between two solves -- it is a per-tree state machine,
 * so trees in different phases share a launch -- and writes the system of the next solve into a private copy of the tree's
 * packed block (Hessian diagonal + Qx at the bounded entries, gradient row := right-hand side + qx, b row := res_b in phase 2);
 * the solves themselves are the batched tree Riccati kernels of ric_blk.cuh (tree.c launches them on that copy).  The corrector
 * system has the same matrix as the predictor's; it is solved by factor+solve again (same factors bit for bit) instead of the
 * reference's solve-only sweep, which costs one more factorisation per iteration at the fast kernels' speed.
 * Element-wise formulas: ipm_elem.cuh / the fused kernel's own (lines of mpc_solvers/c99/d_aux_ip_hard_lib4.c cited there).
 */
#include "launch_util.cuh"
#include "layout.h"
#include "ric_sweeps.cuh"
#include "ipm_elem.cuh"
#include "ric_tree_ipm.cuh"

enum { TS_INIT=0, TS_P1_PRED, TS_P1_CORR, TS_P2_PRED, TS_P2_CORR, TS_DONE, TS_RES_ENTER /* residuals wanted before phase 2 starts */,
       TS_RES_ITER /* ... at the end of a phase-2 iteration */ };
#define TIPM_THREADS 256

/* block-wide reductions in a fixed order (deterministic); every thread gets the result */
struct tipm_sum { __device__ static double op(double a, double b) { return a+b; } __device__ static double w(double v) { return hb_warp_sum(v); } };
struct tipm_min { __device__ static double op(double a, double b) { return fmin(a, b); } __device__ static double w(double v) { return hb_warp_min(v); } };
struct tipm_max { __device__ static double op(double a, double b) { return fmax(a, b); } __device__ static double w(double v) { return hb_warp_max(v); } };
template<class R> __device__ __forceinline__ double tipm_reduce(double v, double *red)
	{
	v = R::w(v);
	if((threadIdx.x&31)==0) red[threadIdx.x>>5] = v;
	__syncthreads();
	double r = red[0];
	for(int i=1; i<(int)(blockDim.x>>5); i++) r = R::op(r, red[i]);
	__syncthreads();
	return r;
	}

__device__ __forceinline__ hb_ipm_ws tipm_ws(const hb_dims &d, long long tree, double *work, long long work_stride, double *dux_all, double *dpi_all)
	{
	hb_ipm_ws w;
	double *p = work + tree*work_stride;
	w.L = nullptr; w.Pb = nullptr;
	w.dux = dux_all + tree*d.ux_stride; w.dpi = dpi_all + tree*d.pi_stride;
	w.res_q = p; p += d.ux_stride; w.rq0 = p; p += d.ux_stride;
	w.res_b = p; p += d.pi_stride; w.b0 = p; p += d.pi_stride;
	w.cv = p; w.nbp = HB_EVEN(d.nbtot);
	return w;
	}

/* write the right-hand sides of the next solve into the tree's private block */
__device__ __forceinline__ void hb_tipm_put_grad(int tid, int nthr, const hb_dims &d, const hb_tipm_maps &m, const hb_ipm_ws &w, const double *rqv, double *im)
	{
	for(int i=tid; i<m.n_ux; i+=nthr) im[m.g_ux[i]] = rqv[i];
	__syncthreads();
	for(int cc=tid; cc<d.nbtot; cc+=nthr) im[m.c_grad[cc]] += w.v(CV_QXG)[cc];
	}
__device__ __forceinline__ void hb_tipm_put_diag(int tid, int nthr, const hb_dims &d, const hb_tipm_maps &m, const hb_ipm_ws &w, const double *in_t, double *im)
	{
	for(int cc=tid; cc<d.nbtot; cc+=nthr) im[m.c_diag[cc]] = in_t[m.c_diag[cc]] + w.v(CV_QXD)[cc];
	}

/* res_q, res_b of the trees that asked for residuals: one warp per (tree, node); the infinity norms go to the tree's state
 * record through atomicMax on the bit pattern (non-negative doubles order like unsigned integers) */
__global__ void __launch_bounds__(128) hb_tipm_res_kernel(hb_dims d, long long n_trees, const double *__restrict__ in,
		const double *__restrict__ ux_all, const double *__restrict__ pi_all, double *__restrict__ dux_all, double *__restrict__ dpi_all,
		double *__restrict__ work, long long work_stride, double *__restrict__ state_all)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	const long long Nn = d.N+1, items = n_trees*Nn, tw = (long long)gridDim.x*nw;
	if(state_all[n_trees*8]==0.0) return;            /* gate: no tree of the batch is unfinished (written by hb_tipm_gate_kernel) */
	for(long long it=(long long)blockIdx.x*nw + warp; it<items; it+=tw)
		{
		const long long tree = it/Nn; const int n = (int)(it - tree*Nn);
		double *state = state_all + tree*8;
		const int st = (int)state[3];
		if(st!=TS_RES_ENTER && st!=TS_RES_ITER) continue;
		const hb_ipm_ws w = tipm_ws(d, tree, work, work_stride, dux_all, dpi_all);
		double nq = 0.0, nb_ = 0.0;
		hb_tipm_node_residuals(c, d.tn, n, in + tree*d.in_stride, w.rq0, w.b0, w.v(CV_LAM_LO), w.v(CV_LAM_UP), d.idxb,
				ux_all + tree*d.ux_stride, pi_all + tree*d.pi_stride, w.res_q, w.res_b, nq, nb_);
		nq = hb_warp_max(nq); nb_ = hb_warp_max(nb_);
		if(lane==0)
			{
			atomicMax(reinterpret_cast<unsigned long long*>(state+5), (unsigned long long)__double_as_longlong(nq));
			atomicMax(reinterpret_cast<unsigned long long*>(state+6), (unsigned long long)__double_as_longlong(nb_));
			}
		__syncwarp();
		}
	}

/* one CTA per tree.  part 0: everything after a solve up to the next solve or up to the point where residuals are wanted;
 * part 1: trees whose residuals have just been computed: mu, then the top of the phase-2 loop.
 * counters[0] = trees not finished, counters[1] = trees waiting for residuals */
__global__ void __launch_bounds__(TIPM_THREADS, 4) hb_tipm_step_kernel(hb_dims d, hb_tipm_maps m, int part, long long n_trees, const double *__restrict__ in,
		double *__restrict__ in_mod, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ dux_all, double *__restrict__ dpi_all,
		double *__restrict__ lam_all, double *__restrict__ t_all, double *__restrict__ info_all,
		double *__restrict__ work, long long work_stride, double *__restrict__ state_all, int *counters)
	{
	__shared__ double red[8];
	const int tid = threadIdx.x, nthr = blockDim.x;
	const long long tree = blockIdx.x;
	if(state_all[n_trees*8]==0.0) return;            /* gate: every tree has finished */
	double *state = state_all + tree*8;
	int st = (int)state[3];
	if(st==TS_DONE) return;
	if(part==1 && st!=TS_RES_ENTER && st!=TS_RES_ITER) { if(tid==0) atomicAdd(&counters[0], 1); return; }
	const hb_ipm_ws w = tipm_ws(d, tree, work, work_stride, dux_all, dpi_all);
	const double *in_t = in + tree*d.in_stride;
	double *im = in_mod + tree*d.in_stride;
	double *ux = ux_all + tree*d.ux_stride, *pi = pi_all + tree*d.pi_stride;
	const int info_len = HB_IPM_INFO_HEAD + 5*k_max;
	double *info = info_all + tree*info_len, *stat = info + HB_IPM_INFO_HEAD;
	const double thr0 = 0.1;
	const double mu_scal = 1.0/(2.0*d.nbtot);
	const double mu_tol_low = mu_tol<1e-5 ? 1e-5 : mu_tol;
	double mu = state[0], alpha = state[1], sigma = state[2], mu_aff;
	int kk = (int)state[4];
	double norms[3] = {state[5], state[6], state[7]};
	bool top1 = false, top2 = false, want_res = false;
	__syncthreads();                         /* everybody has read the state record before thread 0 rewrites it */

	if(part==1)
		{
		/* bound part of the residuals and mu (mpc_solvers/c99/d_res_ip_res_hard.c:39-319); res_q, res_b came from hb_tipm_res_kernel */
		double mu2, nd;
		hb_ipm_residuals_bounds_part(tid, nthr, d, w, ux, mu2, nd);
		mu2 = tipm_reduce<tipm_sum>(mu2, red);
		norms[2] = tipm_reduce<tipm_max>(nd, red);
		mu = mu2/(2.0*d.nbtot);
		if(st==TS_RES_ITER) { if(tid==0) stat[5*kk+4] = mu; kk++; }
		top2 = true;
		}
	else if(st==TS_INIT)
		{
		for(int i=tid; i<m.n_ux; i+=nthr) w.rq0[i] = in_t[m.g_ux[i]];
		for(int j=tid; j<m.n_pi; j+=nthr) w.b0[j] = in_t[m.b_pi[j]];
		for(int n=tid; n<=d.N; n+=nthr)
			{
			const hb_tnode s = d.tn[n];
			for(int j=0; j<s.nb; j++)
				{
				w.v(CV_LB)[s.off_c+j] = in_t[s.off_d+j];
				w.v(CV_UB)[s.off_c+j] = in_t[s.off_d+s.nb+j];
				}
			}
		/* init (c99/d_aux_ip_hard_lib4.c:43-149) */
		if(!warm_start) for(long long i=tid; i<d.ux_stride; i+=nthr) ux[i] = 0.0;
		for(long long i=tid; i<d.pi_stride; i+=nthr) pi[i] = 0.0;
		__syncthreads();
		for(int cc=tid; cc<d.nbtot; cc+=nthr)
			{
			const int iu = d.c_ux[cc];
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
		__syncthreads();
		mu = mu0; alpha = 1.0; sigma = 0.0; kk = 0;
		top1 = true;
		}
	else if(st==TS_P1_PRED || st==TS_P2_PRED)
		{
		const bool p2 = (st==TS_P2_PRED);
		alpha = p2 ? hb_ipm_alpha_part<true>(tid, nthr, d, w, w.dux) : hb_ipm_alpha_part<false>(tid, nthr, d, w, w.dux);
		alpha = tipm_reduce<tipm_min>(alpha, red);
		if(tid==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
		alpha *= 0.995;
		mu_aff = tipm_reduce<tipm_sum>(hb_ipm_mu_aff_part(tid, nthr, d, w, alpha), red)*mu_scal;
		if(tid==0) stat[5*kk+2] = mu_aff;
		sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
		const double sm = sigma*mu;
		if(!p2)
			{
			/* update_gradient (c99/d_aux_ip_hard_lib4.c:387-485) */
			for(int cc=tid; cc<d.nbtot; cc+=nthr)
				{
				double dll = w.v(CV_TINV_LO)[cc]*(sm - w.v(CV_DLAM_LO)[cc]*w.v(CV_DT_LO)[cc]);
				double dlu = w.v(CV_TINV_UP)[cc]*(sm - w.v(CV_DLAM_UP)[cc]*w.v(CV_DT_UP)[cc]);
				w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
				w.v(CV_QXG)[cc] += dlu - dll;
				}
			}
		else
			{
			/* centering correction + update_gradient_res (c99/d_aux_ip_hard_lib4.c:1512-1546, :1550-1639) */
			for(int cc=tid; cc<d.nbtot; cc+=nthr)
				{
				double rml = w.v(CV_RM_LO)[cc] + (w.v(CV_DT_LO)[cc]*w.v(CV_DLAM_LO)[cc] - sm);
				double rmu = w.v(CV_RM_UP)[cc] + (w.v(CV_DT_UP)[cc]*w.v(CV_DLAM_UP)[cc] - sm);
				w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
				w.v(CV_QXG)[cc] = w.v(CV_TINV_LO)[cc]*(rml - w.v(CV_LAM_LO)[cc]*w.v(CV_RD_LO)[cc])
				                  - w.v(CV_TINV_UP)[cc]*(rmu + w.v(CV_LAM_UP)[cc]*w.v(CV_RD_UP)[cc]);
				}
			}
		__syncthreads();
		hb_tipm_put_grad(tid, nthr, d, m, w, p2 ? w.res_q : w.rq0, im);
		st = p2 ? TS_P2_CORR : TS_P1_CORR;
		}
	else if(st==TS_P1_CORR)
		{
		alpha = tipm_reduce<tipm_min>(hb_ipm_alpha_part<false>(tid, nthr, d, w, w.dux), red);
		if(tid==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
		alpha *= 0.995;
		/* update_var (c99/d_aux_ip_hard_lib4.c:618-711) */
		for(long long i=tid; i<d.ux_stride; i+=nthr) ux[i] += alpha*(w.dux[i] - ux[i]);
		for(long long i=tid; i<d.pi_stride; i+=nthr) pi[i] += alpha*(w.dpi[i] - pi[i]);
		double ms = 0.0;
		for(int cc=tid; cc<d.nbtot; cc+=nthr)
			{
			double ll = w.v(CV_LAM_LO)[cc] + alpha*w.v(CV_DLAM_LO)[cc];
			double lu = w.v(CV_LAM_UP)[cc] + alpha*w.v(CV_DLAM_UP)[cc];
			double tl = w.v(CV_T_LO)[cc] + alpha*w.v(CV_DT_LO)[cc];
			double tu = w.v(CV_T_UP)[cc] + alpha*w.v(CV_DT_UP)[cc];
			w.v(CV_LAM_LO)[cc] = ll; w.v(CV_LAM_UP)[cc] = lu; w.v(CV_T_LO)[cc] = tl; w.v(CV_T_UP)[cc] = tu;
			ms += ll*tl + lu*tu;
			}
		mu = tipm_reduce<tipm_sum>(ms, red)*mu_scal;
		if(tid==0) stat[5*kk+4] = mu;
		kk++;
		top1 = true;
		}
	else if(st==TS_P2_CORR)
		{
		alpha = tipm_reduce<tipm_min>(hb_ipm_alpha_part<true>(tid, nthr, d, w, w.dux), red);
		if(tid==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
		alpha *= 0.995;
		/* backup_update_var_res (c99/d_aux_ip_hard_lib4.c:1382-1449) */
		for(long long i=tid; i<d.ux_stride; i+=nthr) ux[i] += alpha*w.dux[i];
		for(long long i=tid; i<d.pi_stride; i+=nthr) pi[i] += alpha*w.dpi[i];
		for(int cc=tid; cc<d.nbtot; cc+=nthr)
			{
			w.v(CV_LAM_LO)[cc] += alpha*w.v(CV_DLAM_LO)[cc]; w.v(CV_LAM_UP)[cc] += alpha*w.v(CV_DLAM_UP)[cc];
			w.v(CV_T_LO)[cc] += alpha*w.v(CV_DT_LO)[cc]; w.v(CV_T_UP)[cc] += alpha*w.v(CV_DT_UP)[cc];
			}
		st = TS_RES_ITER; want_res = true;
		}

	if(top1)
		{
		/* top of the phase-1 loop (d_ip2_res_hard.c:503) */
		if(kk<k_max && mu>mu_tol_low && alpha>=alpha_min)
			{
			/* update_hessian, sigma_mu = 0 (c99/d_aux_ip_hard_lib4.c:217-383) */
			for(int cc=tid; cc<d.nbtot; cc+=nthr)
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
			__syncthreads();
			hb_tipm_put_diag(tid, nthr, d, m, w, in_t, im);
			hb_tipm_put_grad(tid, nthr, d, m, w, w.rq0, im);
			st = TS_P1_PRED;
			}
		else { st = TS_RES_ENTER; want_res = true; }
		}
	if(top2)
		{
		/* top of the phase-2 loop (d_ip2_res_hard.c:756) */
		if(kk<k_max && mu>mu_tol && alpha>=alpha_min)
			{
			/* update_hessian_gradient_res (c99/d_aux_ip_hard_lib4.c:954-1078) */
			for(int cc=tid; cc<d.nbtot; cc+=nthr)
				{
				double til = 1.0/w.v(CV_T_LO)[cc], tiu = 1.0/w.v(CV_T_UP)[cc];
				double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc];
				w.v(CV_TINV_LO)[cc] = til; w.v(CV_TINV_UP)[cc] = tiu;
				w.v(CV_QXD)[cc] = til*ll + tiu*lu;
				w.v(CV_QXG)[cc] = til*(w.v(CV_RM_LO)[cc] - ll*w.v(CV_RD_LO)[cc]) - tiu*(w.v(CV_RM_UP)[cc] + lu*w.v(CV_RD_UP)[cc]);
				}
			__syncthreads();
			hb_tipm_put_diag(tid, nthr, d, m, w, in_t, im);
			hb_tipm_put_grad(tid, nthr, d, m, w, w.res_q, im);
			for(int j=tid; j<m.n_pi; j+=nthr) im[m.b_pi[j]] = w.res_b[j];
			st = TS_P2_PRED;
			}
		else
			{
			int status;
			if(mu<=mu_tol) status = 0;
			else if(kk>=k_max) status = 1;
			else if(alpha<alpha_min) status = 2;
			else status = -1;
			/* results: lam, t as [lower(nb) upper(nb)] per node */
			double *lam = lam_all + tree*2*(long long)d.nbtot, *tt = t_all + tree*2*(long long)d.nbtot;
			for(int n=tid; n<=d.N; n+=nthr)
				{
				const hb_tnode s = d.tn[n];
				for(int j=0; j<s.nb; j++)
					{
					lam[2*s.off_c+j] = w.v(CV_LAM_LO)[s.off_c+j]; lam[2*s.off_c+s.nb+j] = w.v(CV_LAM_UP)[s.off_c+j];
					tt[2*s.off_c+j] = w.v(CV_T_LO)[s.off_c+j]; tt[2*s.off_c+s.nb+j] = w.v(CV_T_UP)[s.off_c+j];
					}
				}
			if(tid==0)
				{
				info[0] = (double)kk; info[1] = (double)status;
				info[2] = norms[0]; info[3] = norms[1]; info[4] = norms[2]; info[5] = mu;
				}
			st = TS_DONE;
			}
		}
	if(tid==0)
		{
		state[0] = mu; state[1] = alpha; state[2] = sigma; state[3] = (double)st; state[4] = (double)kk;
		if(want_res) { state[5] = 0.0; state[6] = 0.0; }       /* the residual kernel accumulates the maxima here */
		state[7] = norms[2];
		if(st!=TS_DONE) atomicAdd(&counters[0], 1);
		if(want_res) atomicAdd(&counters[1], 1);
		}
	}

extern "C" long long hb_tipm_work_doubles(const hb_dims *d)
	{
	return 2*d->ux_stride + 2*d->pi_stride + (long long)CV_COUNT*HB_EVEN(d->nbtot);
	}

/* the gate record behind the last tree's state record: the number of unfinished trees (init > 0: that value; else counters[0]) */
__global__ void hb_tipm_gate_kernel(const int *counters, double *gate, int init)
	{
	*gate = init>0 ? (double)init : (double)counters[0];
	}
extern "C" int hb_launch_tipm_gate(const int *counters, double *gate, int init, void *stream)
	{
	hb_tipm_gate_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counters, gate, init);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_tipm_step(const hb_dims *d, const hb_tipm_maps *m, int part, long long n_trees, const double *in, double *in_mod, int k_max,
		double mu0, double mu_tol, double alpha_min, int warm_start, double *ux, double *pi, double *dux, double *dpi, double *lam,
		double *t, double *info, double *work, long long work_stride, double *state, int *counters, void *stream)
	{
	if(d->tn==NULL || d->nbtot<=0) return -4;
	hb_tipm_step_kernel<<<(int)n_trees, TIPM_THREADS, 0, (cudaStream_t)stream>>>(*d, *m, part, n_trees, in, in_mod, k_max, mu0, mu_tol, alpha_min,
			warm_start, ux, pi, dux, dpi, lam, t, info, work, work_stride, state, counters);
	HB_CK(cudaGetLastError());
	return 0;
	}

/* res_q, res_b of every tree waiting for residuals (between part 0 and part 1 of a step) */
extern "C" int hb_launch_tipm_res(const hb_dims *d, long long n_trees, const double *in, const double *ux, const double *pi,
		double *dux, double *dpi, double *work, long long work_stride, double *state, int sms, void *stream)
	{
	const int warps = 4;
	const int smem = warps*(int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	if(hb_prep(hb_tipm_res_kernel, smem)) return -1;
	long long need = (n_trees*(d->N+1) + warps - 1)/warps, cap = (long long)sms*8;
	hb_tipm_res_kernel<<<(int)(need<cap ? need : cap), warps*32, smem, (cudaStream_t)stream>>>(*d, n_trees, in, ux, pi, dux, dpi, work, work_stride, state);
	HB_CK(cudaGetLastError());
	return 0;
	}

/* ------------------------------------------------------------------------------------------------ */
/* factor only / solve with stored factors over a tree (d_tree_back_ric_rec_trf_libstr, _trs_libstr:   */
/* lqcp_solvers/d_tree_back_ric_rec_libstr.c:591, 625): one warp per tree, the node routines of        */
/* ric_tree_ipm.cuh without the IPM hooks.  mode 0: factors -> L_all.  mode 1: right-hand sides b, [r q] */
/* taken from the block, Pb_k = Lxx_k (Lxx_k' b_k) per edge, backward vector sweep, forward sweep.      */
/* work: per warp slot  rq (ux layout) | b (pi layout) | Pb (pi layout)                                */
/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(128) hb_tree_trf_trs_kernel(hb_dims d, long long n_trees, const double *__restrict__ in,
		double *__restrict__ L_all, double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ work, int mode)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	double *rq = work + gw*(d.ux_stride + 2*d.pi_stride), *bv = rq + d.ux_stride, *Pb = bv + d.pi_stride;
	for(long long tree=gw; tree<n_trees; tree+=tw)
		{
		const double *in_t = in + tree*d.in_stride;
		double *Lt = L_all + tree*d.L_stride;
		if(mode==0)
			{
			for(int n=d.N; n>=0; n--)
				hb_tipm_node_factor(c, d.tn, n, in_t, Lt, nullptr, nullptr, nullptr, nullptr, d.idxb, nullptr, c.bufA, c.bufB);
			continue;
			}
		double *ux = ux_all + tree*d.ux_stride, *pi = pi_all + tree*d.pi_stride;
		for(int n=0; n<=d.N; n++)
			{
			const hb_tnode s = d.tn[n];
			const int nux = s.nu+s.nx, nu = s.nu;
			for(int i=lane; i<nux; i+=32) rq[s.off_ux+i] = in_t[s.off_RSQ+HB_TRI(nux)+i];
			if(s.dad>=0)
				{
				const hb_tnode dd = d.tn[s.dad];
				const int nuxd = dd.nu+dd.nx, nx = s.nx;
				double *t2 = c.sV, *bs = c.sV + 64;
				hb_copy(c, c.bufA, Lt + s.off_L, HB_TRI(nux) + 2*nux);
				for(int j=lane; j<nx; j+=32) { const double v = in_t[s.off_BAbt+nuxd*nx+j]; bv[s.off_pi+j] = v; bs[j] = v; }
				__syncwarp();
				/* Pb = Lxx (Lxx' b) */
				for(int i=lane; i<nx; i+=32)
					{
					double acc = 0.0;
					for(int k=i; k<nx; k++) acc += c.bufA[HB_TRI(nu+k)+nu+i]*bs[k];
					t2[i] = acc;
					}
				__syncwarp();
				for(int i=lane; i<nx; i+=32)
					{
					double acc = 0.0;
					for(int k=0; k<=i; k++) acc += c.bufA[HB_TRI(nu+i)+nu+k]*t2[k];
					Pb[s.off_pi+i] = acc;
					}
				__syncwarp();
				}
			}
		__syncwarp();
		for(int n=d.N; n>=0; n--) hb_tipm_node_trs_back(c, d.tn, n, in_t, Lt, rq, nullptr, d.idxb, ux, Pb, c.bufA);
		for(int n=0; n<=d.N; n++) hb_tipm_node_forward(c, d.tn, n, in_t, Lt, ux, bv, true, ux, pi, c.bufA, c.bufB);
		__syncwarp();
		}
	}

extern "C" int hb_launch_tree_trf_trs(const hb_dims *d, long long n_trees, const double *in, double *L, double *ux, double *pi,
		double *work, int n_slots, int mode, int grid, int warps, void *stream)
	{
	if(d->tn==NULL) return -4;
	if(grid*warps>n_slots || warps>4) return -3;
	const int smem = warps*(int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	if(hb_prep(hb_tree_trf_trs_kernel, smem)) return -1;
	hb_tree_trf_trs_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_trees, in, L, ux, pi, work, mode);
	HB_CK(cudaGetLastError());
	return 0;
	}
