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

enum { TS_INIT=0, TS_P1_PRED, TS_P1_CORR, TS_P2_PRED, TS_P2_CORR, TS_DONE };

/* write the right-hand sides of the next solve into the tree's private block */
__device__ __forceinline__ void hb_tipm_put_grad(int lane, const hb_dims &d, const hb_tipm_maps &m, const hb_ipm_ws &w, const double *rqv, double *im)
	{
	for(int i=lane; i<m.n_ux; i+=32) im[m.g_ux[i]] = rqv[i];
	__syncwarp();
	for(int cc=lane; cc<d.nbtot; cc+=32) im[m.c_grad[cc]] += w.v(CV_QXG)[cc];
	}
__device__ __forceinline__ void hb_tipm_put_diag(int lane, const hb_dims &d, const hb_tipm_maps &m, const hb_ipm_ws &w, const double *in_t, double *im)
	{
	for(int cc=lane; cc<d.nbtot; cc+=32) im[m.c_diag[cc]] = in_t[m.c_diag[cc]] + w.v(CV_QXD)[cc];
	}

__device__ __forceinline__ void hb_tipm_residuals(const hb_ctx &c, const hb_dims &d, const double *in_t, const hb_ipm_ws &w,
		const double *ux, const double *pi, double *mu, double *norms)
	{
	double mu2, nd, nq = 0.0, nb_ = 0.0;
	hb_ipm_residuals_bounds(c.lane, d, w, ux, mu2, nd);
	__syncwarp();
	for(int n=0; n<=d.N; n++)
		hb_tipm_node_residuals(c, d.tn, n, in_t, w.rq0, w.b0, w.v(CV_LAM_LO), w.v(CV_LAM_UP), d.idxb, ux, pi, w.res_q, w.res_b, nq, nb_);
	if(d.nbtot>0) *mu = mu2/(2.0*d.nbtot);
	norms[0] = hb_warp_max(nq); norms[1] = hb_warp_max(nb_); norms[2] = hb_warp_max(nd);
	}

__global__ void __launch_bounds__(128) hb_tipm_step_kernel(hb_dims d, hb_tipm_maps m, long long n_trees, const double *__restrict__ in,
		double *__restrict__ in_mod, int k_max, double mu0, double mu_tol, double alpha_min, int warm_start,
		double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ dux_all, double *__restrict__ dpi_all,
		double *__restrict__ lam_all, double *__restrict__ t_all, double *__restrict__ info_all,
		double *__restrict__ work, long long work_stride, double *__restrict__ state_all, int *n_active)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long tree = (long long)blockIdx.x*nw + warp;
	if(tree>=n_trees) return;
	double *state = state_all + tree*8;
	int st = (int)state[3];
	if(st==TS_DONE) return;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	hb_ipm_ws w;
	{
	double *p = work + tree*work_stride;
	w.L = nullptr; w.Pb = nullptr;
	w.dux = dux_all + tree*d.ux_stride; w.dpi = dpi_all + tree*d.pi_stride;
	w.res_q = p; p += d.ux_stride; w.rq0 = p; p += d.ux_stride;
	w.res_b = p; p += d.pi_stride; w.b0 = p; p += d.pi_stride;
	w.cv = p; w.nbp = HB_EVEN(d.nbtot);
	}
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
	bool top1 = false, top2 = false;

	if(st==TS_INIT)
		{
		for(int i=lane; i<m.n_ux; i+=32) w.rq0[i] = in_t[m.g_ux[i]];
		for(int j=lane; j<m.n_pi; j+=32) w.b0[j] = in_t[m.b_pi[j]];
		for(int n=0; n<=d.N; n++)
			{
			const hb_tnode s = d.tn[n];
			for(int j=lane; j<s.nb; j+=32)
				{
				w.v(CV_LB)[s.off_c+j] = in_t[s.off_d+j];
				w.v(CV_UB)[s.off_c+j] = in_t[s.off_d+s.nb+j];
				}
			}
		/* init (c99/d_aux_ip_hard_lib4.c:43-149) */
		if(!warm_start) for(long long i=lane; i<d.ux_stride; i+=32) ux[i] = 0.0;
		for(long long i=lane; i<d.pi_stride; i+=32) pi[i] = 0.0;
		__syncwarp();
		for(int cc=lane; cc<d.nbtot; cc+=32)
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
		__syncwarp();
		mu = mu0; alpha = 1.0; sigma = 0.0; kk = 0;
		top1 = true;
		}
	else if(st==TS_P1_PRED)
		{
		alpha = hb_ipm_alpha<false>(lane, d, w, w.dux);
		__syncwarp();
		if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
		alpha *= 0.995;
		mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
		if(lane==0) stat[5*kk+2] = mu_aff;
		sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
		const double sm = sigma*mu;
		for(int cc=lane; cc<d.nbtot; cc+=32)
			{
			double dll = w.v(CV_TINV_LO)[cc]*(sm - w.v(CV_DLAM_LO)[cc]*w.v(CV_DT_LO)[cc]);
			double dlu = w.v(CV_TINV_UP)[cc]*(sm - w.v(CV_DLAM_UP)[cc]*w.v(CV_DT_UP)[cc]);
			w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
			w.v(CV_QXG)[cc] += dlu - dll;
			}
		__syncwarp();
		hb_tipm_put_grad(lane, d, m, w, w.rq0, im);
		st = TS_P1_CORR;
		}
	else if(st==TS_P1_CORR)
		{
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
		__syncwarp();
		top1 = true;
		}
	else if(st==TS_P2_PRED)
		{
		alpha = hb_ipm_alpha<true>(lane, d, w, w.dux);
		__syncwarp();
		if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
		alpha *= 0.995;
		mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
		if(lane==0) stat[5*kk+2] = mu_aff;
		sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
		const double sm = sigma*mu;
		for(int cc=lane; cc<d.nbtot; cc+=32)
			{
			double rml = w.v(CV_RM_LO)[cc] + (w.v(CV_DT_LO)[cc]*w.v(CV_DLAM_LO)[cc] - sm);
			double rmu = w.v(CV_RM_UP)[cc] + (w.v(CV_DT_UP)[cc]*w.v(CV_DLAM_UP)[cc] - sm);
			w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
			w.v(CV_QXG)[cc] = w.v(CV_TINV_LO)[cc]*(rml - w.v(CV_LAM_LO)[cc]*w.v(CV_RD_LO)[cc])
			                  - w.v(CV_TINV_UP)[cc]*(rmu + w.v(CV_LAM_UP)[cc]*w.v(CV_RD_UP)[cc]);
			}
		__syncwarp();
		hb_tipm_put_grad(lane, d, m, w, w.res_q, im);
		st = TS_P2_CORR;
		}
	else if(st==TS_P2_CORR)
		{
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
		__syncwarp();
		hb_tipm_residuals(c, d, in_t, w, ux, pi, &mu, norms);
		if(lane==0) stat[5*kk+4] = mu;
		kk++;
		__syncwarp();
		top2 = true;
		}

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
			__syncwarp();
			hb_tipm_put_diag(lane, d, m, w, in_t, im);
			hb_tipm_put_grad(lane, d, m, w, w.rq0, im);
			st = TS_P1_PRED;
			}
		else
			{
			hb_tipm_residuals(c, d, in_t, w, ux, pi, &mu, norms);
			__syncwarp();
			top2 = true;
			}
		}
	if(top2)
		{
		/* top of the phase-2 loop (d_ip2_res_hard.c:756) */
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
			__syncwarp();
			hb_tipm_put_diag(lane, d, m, w, in_t, im);
			hb_tipm_put_grad(lane, d, m, w, w.res_q, im);
			for(int j=lane; j<m.n_pi; j+=32) im[m.b_pi[j]] = w.res_b[j];
			st = TS_P2_PRED;
			}
		else
			{
			int status;
			if(mu<=mu_tol) status = 0;
			else if(kk>=k_max) status = 1;
			else if(alpha<alpha_min) status = 2;
			else status = -1;
			double *lam = lam_all + tree*2*(long long)d.nbtot, *tt = t_all + tree*2*(long long)d.nbtot;
			for(int n=0; n<=d.N; n++)
				{
				const hb_tnode s = d.tn[n];
				for(int j=lane; j<s.nb; j+=32)
					{
					lam[2*s.off_c+j] = w.v(CV_LAM_LO)[s.off_c+j]; lam[2*s.off_c+s.nb+j] = w.v(CV_LAM_UP)[s.off_c+j];
					tt[2*s.off_c+j] = w.v(CV_T_LO)[s.off_c+j]; tt[2*s.off_c+s.nb+j] = w.v(CV_T_UP)[s.off_c+j];
					}
				}
			if(lane==0)
				{
				info[0] = (double)kk; info[1] = (double)status;
				info[2] = norms[0]; info[3] = norms[1]; info[4] = norms[2]; info[5] = mu;
				}
			st = TS_DONE;
			}
		}
	__syncwarp();
	if(lane==0)
		{
		state[0] = mu; state[1] = alpha; state[2] = sigma; state[3] = (double)st; state[4] = (double)kk;
		state[5] = norms[0]; state[6] = norms[1]; state[7] = norms[2];
		if(st!=TS_DONE) atomicAdd(n_active, 1);
		}
	}

extern "C" long long hb_tipm_work_doubles(const hb_dims *d)
	{
	return 2*d->ux_stride + 2*d->pi_stride + (long long)CV_COUNT*HB_EVEN(d->nbtot);
	}

extern "C" int hb_launch_tipm_step(const hb_dims *d, const hb_tipm_maps *m, long long n_trees, const double *in, double *in_mod, int k_max,
		double mu0, double mu_tol, double alpha_min, int warm_start, double *ux, double *pi, double *dux, double *dpi, double *lam,
		double *t, double *info, double *work, long long work_stride, double *state, int *n_active, void *stream)
	{
	if(d->tn==NULL || d->nbtot<=0) return -4;
	const int warps = 4;
	const int smem = warps*(int)sizeof(double)*hb_smem_doubles_per_warp(d->nzM, d->nxM);
	if(hb_prep(hb_tipm_step_kernel, smem)) return -1;
	const int grid = (int)((n_trees + warps - 1)/warps);
	hb_tipm_step_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, *m, n_trees, in, in_mod, k_max, mu0, mu_tol, alpha_min, warm_start,
			ux, pi, dux, dpi, lam, t, info, work, work_stride, state, n_active);
	HB_CK(cudaGetLastError());
	return 0;
	}
