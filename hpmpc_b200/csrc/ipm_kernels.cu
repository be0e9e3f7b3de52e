/*
 * ipm_kernels.cu -- the fused two-phase Mehrotra IPM kernel (policy-templated on the sweeps) and its C launchers.
 *   hb_ipm_kernel      whole IPM on device, no host round trip per iteration
 *                                                <- d_ip2_res_mpc_hard_tv      (mpc_solvers/d_ip2_res_hard.c:116)
 *                      element-wise steps        <- mpc_solvers/c99/d_aux_ip_hard_lib4.c (lines cited inline)
 *                      residuals                 <- mpc_solvers/c99/d_res_ip_res_hard.c:39
 * A translation unit of its own so that it compiles in parallel with the Riccati kernels (the size-specialised
 * instantiation is ~0.5 MB of SASS).
 */
#include "launch_util.cuh"
#include "ipm_sweeps.cuh"
#include "ric_tree_ipm.cuh"

/* scenario tree: d.tn is the node table (BFS order), d.N = Nn-1; nodes play the role of the stages */
struct hb_sweeps_tree
	{
	typedef hb_ctx ctx_t;
	static constexpr bool has_kkt = false;     /* the re-solve is built for chains */
	__device__ static __forceinline__ int smem_doubles(const hb_dims &d) { return hb_smem_doubles_per_warp(d.nzM, d.nxM); }
	__device__ static __forceinline__ long long L_doubles(const hb_dims &d) { return d.L_stride; }
	__device__ static __forceinline__ void init(ctx_t &c, const hb_dims &d, double *smem_warp, int lane) { c = hb_make_ctx(d, smem_warp, lane); }
	__device__ static __forceinline__ void extract(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w)
		{
		const int lane = c.lane;
		for(int n=0; n<=d.N; n++)
			{
			const hb_tnode s = d.tn[n];
			const int nux = s.nu+s.nx;
			for(int i=lane; i<nux; i+=32) w.rq0[s.off_ux+i] = in_inst[s.off_RSQ+HB_TRI(nux)+i];
			if(s.dad>=0)
				{
				const hb_tnode dd = d.tn[s.dad];
				const int nuxd = dd.nu+dd.nx;
				for(int j=lane; j<s.nx; j+=32) w.b0[s.off_pi+j] = in_inst[s.off_BAbt+nuxd*s.nx+j];
				}
			for(int j=lane; j<s.nb; j+=32)
				{
				w.v(CV_LB)[s.off_c+j] = in_inst[s.off_d+j];
				w.v(CV_UB)[s.off_c+j] = in_inst[s.off_d+s.nb+j];
				}
			}
		}
	__device__ static __forceinline__ void load(ctx_t &c, const hb_dims &d, const hb_ipm_ws &w, const double *lam, const double *tt)
		{
		for(int n=0; n<=d.N; n++)
			{
			const hb_tnode s = d.tn[n];
			for(int j=c.lane; j<s.nb; j+=32)
				{
				w.v(CV_LAM_LO)[s.off_c+j] = lam[2*s.off_c+j]; w.v(CV_LAM_UP)[s.off_c+j] = lam[2*s.off_c+s.nb+j];
				w.v(CV_T_LO)[s.off_c+j] = tt[2*s.off_c+j]; w.v(CV_T_UP)[s.off_c+j] = tt[2*s.off_c+s.nb+j];
				}
			}
		}
	__device__ static __forceinline__ void emit(ctx_t &c, const hb_dims &d, const hb_ipm_ws &w, double *lam, double *tt)
		{
		for(int n=0; n<=d.N; n++)
			{
			const hb_tnode s = d.tn[n];
			for(int j=c.lane; j<s.nb; j+=32)
				{
				lam[2*s.off_c+j] = w.v(CV_LAM_LO)[s.off_c+j]; lam[2*s.off_c+s.nb+j] = w.v(CV_LAM_UP)[s.off_c+j];
				tt[2*s.off_c+j] = w.v(CV_T_LO)[s.off_c+j]; tt[2*s.off_c+s.nb+j] = w.v(CV_T_UP)[s.off_c+j];
				}
			}
		}
	__device__ static __forceinline__ void backward(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *Qx, const double *qx)
		{
		for(int n=d.N; n>=0; n--)
			hb_tipm_node_factor(c, d.tn, n, in_inst, w.L, bv, rqv!=nullptr ? rqv : w.rq0, Qx, qx, d.idxb, w.Pb, c.bufA, c.bufB);
		}
	__device__ static __forceinline__ void forward_sv(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, double *ux, double *pi)
		{
		for(int n=0; n<=d.N; n++) hb_tipm_node_forward(c, d.tn, n, in_inst, w.L, nullptr, bv, false, ux, pi, c.bufA, c.bufB);
		}
	__device__ static __forceinline__ void trs(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *bv, const double *rqv, const double *qx)
		{
		for(int n=d.N; n>=0; n--) hb_tipm_node_trs_back(c, d.tn, n, in_inst, w.L, rqv, qx, d.idxb, w.dux, w.Pb, c.bufA);
		for(int n=0; n<=d.N; n++) hb_tipm_node_forward(c, d.tn, n, in_inst, w.L, w.dux, bv, true, w.dux, w.dpi, c.bufA, c.bufB);
		}
	__device__ static __forceinline__ void residuals(ctx_t &c, const hb_dims &d, const double *in_inst, const hb_ipm_ws &w,
			const double *ux, const double *pi, double *mu, double *norms)
		{
		double mu2, nd, nq = 0.0, nb_ = 0.0;
		hb_ipm_residuals_bounds(c.lane, d, w, ux, mu2, nd);
		__syncwarp();
		for(int n=0; n<=d.N; n++)
			hb_tipm_node_residuals(c, d.tn, n, in_inst, w.rq0, w.b0, w.v(CV_LAM_LO), w.v(CV_LAM_UP), d.idxb, ux, pi, w.res_q, w.res_b, nq, nb_);
		if(d.nbtot>0) *mu = mu2/(2.0*d.nbtot);
		if(norms!=nullptr) { norms[0] = hb_warp_max(nq); norms[1] = hb_warp_max(nb_); norms[2] = hb_warp_max(nd); }
		}
	};


/* ------------------------------------------------------------------------------------------------ */
/* stand-alone residuals of a given point (ux, pi, lam, t): d_res_res_mpc_hard_tv                      */
/* (mpc_solvers/c99/d_res_ip_res_hard.c:39) -- res_q, res_b, res_d, res_m and mu; one warp per instance. */
/* res_d / res_m come out in the layout of lam ([lb ub lg ug] per stage).                               */
/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) hb_res_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		const double *__restrict__ ux_all, const double *__restrict__ pi_all, const double *__restrict__ lam_all,
		const double *__restrict__ t_all, double *__restrict__ rq_all, double *__restrict__ rb_all, double *__restrict__ rd_all,
		double *__restrict__ rm_all, double *__restrict__ mu_all, double *__restrict__ work, long long work_stride)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hb_ctx c = hb_make_ctx(d, hb_smem + (size_t)warp*hb_smem_doubles_per_warp(d.nzM, d.nxM), lane);
	hb_ipm_ws w;
	{
	double *p = work + gw*work_stride;
	w.L = p;                         /* no factor here: the vectors start at the slot */
	w.dux = p; p += d.ux_stride; w.res_q = p; p += d.ux_stride; w.rq0 = p; p += d.ux_stride;
	w.dpi = p; p += d.pi_stride; w.Pb = p; p += d.pi_stride; w.res_b = p; p += d.pi_stride; w.b0 = p; p += d.pi_stride;
	w.cv = p; w.nbp = HB_EVEN(d.nbtot);
	}
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *in_inst = in + inst*d.in_stride;
		const double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		hb_ipm_extract_chain(lane, d, in_inst, w);
		hb_ipm_load_chain(lane, d, w, lam_all + inst*2*(long long)d.nbtot, t_all + inst*2*(long long)d.nbtot);
		__syncwarp();
		double mu = 0.0, norms[3];
		hb_ipm_residuals(c, d, in_inst, w, ux, pi, &mu, norms);
		__syncwarp();
		for(long long i=lane; i<d.ux_stride; i+=32) rq_all[inst*d.ux_stride+i] = w.res_q[i];
		for(long long i=lane; i<d.pi_stride; i+=32) rb_all[inst*d.pi_stride+i] = w.res_b[i];
		double *rd = rd_all + inst*2*(long long)d.nbtot, *rm = rm_all ? rm_all + inst*2*(long long)d.nbtot : nullptr;
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			for(int j=lane; j<s.nb; j+=32)
				{
				rd[2*s.off_c+j] = w.v(CV_RD_LO)[s.off_c+j]; rd[2*s.off_c+s.nb+j] = w.v(CV_RD_UP)[s.off_c+j];
				if(rm) { rm[2*s.off_c+j] = w.v(CV_RM_LO)[s.off_c+j]; rm[2*s.off_c+s.nb+j] = w.v(CV_RM_UP)[s.off_c+j]; }
				}
			for(int j=lane; j<s.ng; j+=32)
				{
				const int o = 2*s.off_c + 2*s.nb, cg = s.off_c + s.nb + j;
				rd[o+j] = w.v(CV_RD_LO)[cg]; rd[o+s.ng+j] = w.v(CV_RD_UP)[cg];
				if(rm) { rm[o+j] = w.v(CV_RM_LO)[cg]; rm[o+s.ng+j] = w.v(CV_RM_UP)[cg]; }
				}
			}
		if(lane==0) mu_all[inst] = mu;
		__syncwarp();
		}
	}

extern "C" long long hb_res_work_doubles(const hb_dims *d) { return 3*d->ux_stride + 4*d->pi_stride + (long long)CV_COUNT*HB_EVEN(d->nbtot); }
extern "C" int hb_launch_res(const hb_dims *d, long long n_inst, const double *in, const double *ux, const double *pi, const double *lam,
		const double *t, double *rq, double *rb, double *rd, double *rm, double *mu, double *work, int grid, int warps, void *stream)
	{
	if(d->nzM>64) return -2;
	if(warps>8) return -3;
	const int smem = warps*hb_smem_bytes_per_warp(d);
	if(hb_prep(hb_res_kernel, smem)) return -1;
	hb_res_kernel<<<grid, warps*32, smem, (cudaStream_t)stream>>>(*d, n_inst, in, ux, pi, lam, t, rq, rb, rd, rm, mu, work, hb_res_work_doubles(d));
	HB_CK(cudaGetLastError());
	return 0;
	}

/* KKT = true: every instance also leaves its KKT state in kkt + inst*kkt_stride (section on the re-solve below); the plain IPM is
 * compiled without any of it */
template<class S, bool KKT>
__global__ void __launch_bounds__(256) hb_ipm_kernel(hb_dims d, long long n_inst, const double *__restrict__ in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *__restrict__ ux_all, double *__restrict__ pi_all,
		double *__restrict__ lam_all, double *__restrict__ t_all, double *__restrict__ info_all,
		double *__restrict__ work, long long work_stride, int *counter, double *__restrict__ kkt, long long kkt_stride)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp;
	typename S::ctx_t c;
	S::init(c, d, hb_smem + (size_t)warp*S::smem_doubles(d), lane);
	const long long slot_doubles = hb_ipm_slot_doubles<S>(d);
	/* the warp's own slot -- or, when the KKT state is kept, the instance's state block itself (set per instance below):
	 * factor, t_inv and vectors are then already where the re-solve expects them when the instance is done */
	hb_ipm_ws w = hb_ipm_make_ws<S>(d, work + gw*work_stride);
	const int info_len = HB_IPM_INFO_HEAD + 5*k_max;
	const double thr0 = 0.1;
	/* warm_start = 2: single-Newton-step mode (d_ip2_res_mpc_hard_tv_single_newton_step, mpc_solvers/d_ip2_res_hard.c:1348): the
	 * iterate (ux, pi, lam, t) the caller left in the output arrays is taken as is, phase 1 is skipped, every iteration is a
	 * residual-based step with the centering term fixed at mu0 (:1749), and the loop runs k_max iterations (:1652) */
	const bool newton = (warm_start==2);

	for(;;)
		{
		/* dynamic instance queue: a warp that converges early simply takes the next instance, so the
		 * active set stays compact without a separate compaction pass */
		long long inst = 0;
		if(lane==0) inst = atomicAdd(counter, 1);
		inst = __shfl_sync(HB_FULL, inst, 0);
		if(inst>=n_inst) break;

		const double *in_inst = in + inst*d.in_stride;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		double *info = info_all + inst*info_len;
		double *stat = info + HB_IPM_INFO_HEAD;
		if(KKT) w = hb_ipm_make_ws<S>(d, kkt + inst*kkt_stride);

		/* vectors taken from the instance block: rq0 = [r q], b0 = b, bounds */
		S::extract(c, d, in_inst, w);
		__syncwarp();

		int kk = 0, status = -1, n_ph2 = 0;
		double mu = 0.0, norms[3] = {0.0, 0.0, 0.0};

		if(d.nbtot==0)
			{
			/* no constraints: one Riccati solve (d_ip2_res_hard.c:430-450) */
			S::backward(c, d, in_inst, w, nullptr, nullptr, nullptr, nullptr);
			__syncwarp();
			S::forward_sv(c, d, in_inst, w, nullptr, ux, pi);
			__syncwarp();
			S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
			status = 0;
			}
		else
			{
			const double mu_scal = 1.0/(2.0*d.nbtot);
			double sigma = 0.0, alpha = 1.0, mu_aff;
			/* init (c99/d_aux_ip_hard_lib4.c:43-149) */
			if(newton) S::load(c, d, w, lam_all + inst*2*(long long)d.nbtot, t_all + inst*2*(long long)d.nbtot);
			if(!warm_start) for(long long i=lane; i<d.ux_stride; i+=32) ux[i] = 0.0;
			if(!newton) for(long long i=lane; i<d.pi_stride; i+=32) pi[i] = 0.0;
			__syncwarp();
			for(int cc=lane; cc<(newton ? 0 : d.nbtot); cc+=32)
				{
				const int iu = d.c_ux[cc];
				if(iu<0) continue;                                    /* general constraint: below */
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
			if(d.ngtot>0 && !newton)
				{
				/* general constraints, from the ux the bounds have just moved: t = max(thr0, +-([D C] ux - d)), no projection
				 * (c99/d_aux_ip_hard_lib4.c:121-147) */
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
			mu = mu0;
			const double mu_tol_low = mu_tol<1e-5 ? 1e-5 : mu_tol;

			/* ---------- phase 1 (d_ip2_res_hard.c:503-718) ---------- */
			while(!newton && kk<k_max && mu>mu_tol_low && alpha>=alpha_min)
				{
				/* update_hessian, sigma_mu = 0 (c99/d_aux_ip_hard_lib4.c:217-383) */
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
				HBF_STAMP(300);
				S::backward(c, d, in_inst, w, nullptr, nullptr, w.v(CV_QXD), w.v(CV_QXG));
				__syncwarp();
				HBF_STAMP(301);
				S::forward_sv(c, d, in_inst, w, nullptr, w.dux, w.dpi);
				HBF_STAMP(302);
				__syncwarp();
				hb_gen_values(lane, d, in_inst, w, w.dux);
				alpha = hb_ipm_alpha<false>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
				alpha *= 0.995;
				mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
				if(lane==0) stat[5*kk+2] = mu_aff;
				sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
				{
				/* update_gradient (c99/d_aux_ip_hard_lib4.c:387-485) */
				const double sm = sigma*mu;
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double dll = w.v(CV_TINV_LO)[cc]*(sm - w.v(CV_DLAM_LO)[cc]*w.v(CV_DT_LO)[cc]);
					double dlu = w.v(CV_TINV_UP)[cc]*(sm - w.v(CV_DLAM_UP)[cc]*w.v(CV_DT_UP)[cc]);
					w.v(CV_DLAM_LO)[cc] = dll; w.v(CV_DLAM_UP)[cc] = dlu;
					w.v(CV_QXG)[cc] += dlu - dll;
					}
				}
				__syncwarp();
				HBF_STAMP(303);
				S::trs(c, d, in_inst, w, w.b0, w.rq0, w.v(CV_QXG));
				HBF_STAMP(304);
				__syncwarp();
				hb_gen_values(lane, d, in_inst, w, w.dux);
				alpha = hb_ipm_alpha<false>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
				alpha *= 0.995;
				/* update_var (c99/d_aux_ip_hard_lib4.c:618-711) */
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
				}

			/* ---------- phase 2 (d_ip2_res_hard.c:756-1273) ---------- */
			S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
			__syncwarp();
			while(kk<k_max && (newton || (mu>mu_tol && alpha>=alpha_min)))
				{
				/* update_hessian_gradient_res (c99/d_aux_ip_hard_lib4.c:954-1078) */
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double til = 1.0/w.v(CV_T_LO)[cc], tiu = 1.0/w.v(CV_T_UP)[cc];
					double ll = w.v(CV_LAM_LO)[cc], lu = w.v(CV_LAM_UP)[cc];
					w.v(CV_TINV_LO)[cc] = til; w.v(CV_TINV_UP)[cc] = tiu;
					w.v(CV_QXD)[cc] = til*ll + tiu*lu;
					w.v(CV_QXG)[cc] = til*(w.v(CV_RM_LO)[cc] - ll*w.v(CV_RD_LO)[cc]) - tiu*(w.v(CV_RM_UP)[cc] + lu*w.v(CV_RD_UP)[cc]);
					}
				__syncwarp();
				HBF_STAMP(300);
				/* single-Newton-step mode: the reference solves the predictor with the ORIGINAL b and q (update_b = 0, update_q = 1
				 * with q, d_ip2_res_hard.c:1736) and only the corrector with the residuals (:1788); restated as it is */
				S::backward(c, d, in_inst, w, newton ? nullptr : w.res_b, newton ? nullptr : w.res_q, w.v(CV_QXD), w.v(CV_QXG));
				__syncwarp();
				HBF_STAMP(301);
				S::forward_sv(c, d, in_inst, w, newton ? nullptr : w.res_b, w.dux, w.dpi);
				HBF_STAMP(302);
				__syncwarp();
				hb_gen_values(lane, d, in_inst, w, w.dux);
				alpha = hb_ipm_alpha<true>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+1] = alpha; }
				alpha *= 0.995;
				mu_aff = hb_ipm_mu_aff(lane, d, w, alpha, mu_scal);
				if(lane==0) stat[5*kk+2] = mu_aff;
				if(!newton) { sigma = mu_aff/mu; sigma = sigma*sigma*sigma; }
				{
				/* centering correction + update_gradient_res (c99/d_aux_ip_hard_lib4.c:1512-1546, :1550-1639) */
				const double sm = newton ? mu0 : sigma*mu;
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					double rml = w.v(CV_RM_LO)[cc] + (w.v(CV_DT_LO)[cc]*w.v(CV_DLAM_LO)[cc] - sm);
					double rmu = w.v(CV_RM_UP)[cc] + (w.v(CV_DT_UP)[cc]*w.v(CV_DLAM_UP)[cc] - sm);
					w.v(CV_RM_LO)[cc] = rml; w.v(CV_RM_UP)[cc] = rmu;
					w.v(CV_QXG)[cc] = w.v(CV_TINV_LO)[cc]*(rml - w.v(CV_LAM_LO)[cc]*w.v(CV_RD_LO)[cc])
					                  - w.v(CV_TINV_UP)[cc]*(rmu + w.v(CV_LAM_UP)[cc]*w.v(CV_RD_UP)[cc]);
					}
				}
				__syncwarp();
				HBF_STAMP(303);
				S::trs(c, d, in_inst, w, w.res_b, w.res_q, w.v(CV_QXG));
				HBF_STAMP(304);
				__syncwarp();
				hb_gen_values(lane, d, in_inst, w, w.dux);
				alpha = hb_ipm_alpha<true>(lane, d, w, w.dux);
				__syncwarp();
				if(lane==0) { stat[5*kk] = sigma; stat[5*kk+3] = alpha; }
				alpha *= 0.995;
				/* backup_update_var_res (c99/d_aux_ip_hard_lib4.c:1382-1449): the backup is what a later solve with a new
				 * right-hand side starts from (d_ip2_res_hard.c:2138-2173); kept only when the caller asked for the KKT state */
				if(KKT)
					{
					double *kb = kkt + inst*kkt_stride + HB_EVEN(slot_doubles);
					for(long long i=lane; i<d.ux_stride; i+=32) kb[i] = ux[i];
					kb += d.ux_stride;
					for(long long i=lane; i<d.pi_stride; i+=32) kb[i] = pi[i];
					kb += d.pi_stride;
					for(int cc=lane; cc<d.nbtot; cc+=32)
						{
						kb[cc] = w.v(CV_LAM_LO)[cc]; kb[w.nbp+cc] = w.v(CV_LAM_UP)[cc];
						kb[2*w.nbp+cc] = w.v(CV_T_LO)[cc]; kb[3*w.nbp+cc] = w.v(CV_T_UP)[cc];
						}
					n_ph2++;
					}
				for(long long i=lane; i<d.ux_stride; i+=32) ux[i] += alpha*w.dux[i];
				for(long long i=lane; i<d.pi_stride; i+=32) pi[i] += alpha*w.dpi[i];
				for(int cc=lane; cc<d.nbtot; cc+=32)
					{
					w.v(CV_LAM_LO)[cc] += alpha*w.v(CV_DLAM_LO)[cc]; w.v(CV_LAM_UP)[cc] += alpha*w.v(CV_DLAM_UP)[cc];
					w.v(CV_T_LO)[cc] += alpha*w.v(CV_DT_LO)[cc]; w.v(CV_T_UP)[cc] += alpha*w.v(CV_DT_UP)[cc];
					}
				__syncwarp();
				HBF_STAMP(305);
				S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
				HBF_STAMP(306);
				if(lane==0) stat[5*kk+4] = mu;
				kk++;
				__syncwarp();
				}
			if(newton) status = (kk>=k_max) ? 1 : (alpha<alpha_min ? 2 : -1);       /* d_ip2_res_hard.c:1911-1918 */
			else if(mu<=mu_tol) status = 0;
			else if(kk>=k_max) status = 1;
			else if(alpha<alpha_min) status = 2;
			else status = -1;
			}

		/* results: lam, t as [lower(nb) upper(nb)] per stage (interfaces/c/fortran_order_interface.c:662-671) */
		double *lam = lam_all + inst*2*(long long)d.nbtot, *tt = t_all + inst*2*(long long)d.nbtot;
		S::emit(c, d, w, lam, tt);
		if(KKT)
			{
			/* the work slot IS the state block: the factor of the last iteration and t_inv are in place; only the flag is left */
			if(lane==0) kkt[inst*kkt_stride + HB_EVEN(slot_doubles) + d.ux_stride + d.pi_stride + 4*w.nbp] = (double)n_ph2;
			}
		if(lane==0)
			{
			info[0] = (double)kk; info[1] = (double)status;
			info[2] = norms[0]; info[3] = norms[1]; info[4] = norms[2]; info[5] = mu;
			}
		__syncwarp();
		}
	}

/* ------------------------------------------------------------------------------------------------ */
/* the IPM's last KKT system solved again for a new right-hand side:             */
/* d_kkt_solve_new_rhs_res_mpc_hard_tv, mpc_solvers/d_ip2_res_hard.c:1922.  `in` carries the new b, */
/* [r q] and bounds (its matrices are the ones the factor was made from); `kkt` is the state an IPM */
/* call left behind.  One warp per instance, working inside the instance's own state block:        */
/* the factor, t_inv and the backup are read, the other vectors of the block are scratch.           */
/* ------------------------------------------------------------------------------------------------ */
template<class S>
__global__ void __launch_bounds__(256) hb_kkt_new_rhs_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ kkt, long long kkt_stride, double *__restrict__ ux_all, double *__restrict__ pi_all,
		double *__restrict__ lam_all, double *__restrict__ t_all, double *__restrict__ info_all, int *counter)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31;
	typename S::ctx_t c;
	S::init(c, d, hb_smem + (size_t)warp*S::smem_doubles(d), lane);
	const long long slot_doubles = hb_ipm_slot_doubles<S>(d);
	for(;;)
		{
		long long inst = 0;
		if(lane==0) inst = atomicAdd(counter, 1);
		inst = __shfl_sync(HB_FULL, inst, 0);
		if(inst>=n_inst) break;
		hb_ipm_ws w;
		double *p = kkt + inst*kkt_stride;
		w.L = p; p += S::L_doubles(d);
		w.dux = p; p += d.ux_stride; w.res_q = p; p += d.ux_stride; w.rq0 = p; p += d.ux_stride;
		w.dpi = p; p += d.pi_stride; w.Pb = p; p += d.pi_stride; w.res_b = p; p += d.pi_stride; w.b0 = p; p += d.pi_stride;
		w.cv = p; w.nbp = HB_EVEN(d.nbtot);
		const double *kb = kkt + inst*kkt_stride + HB_EVEN(slot_doubles);
		const double *in_inst = in + inst*d.in_stride;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		double *info = info_all + inst*HB_IPM_INFO_HEAD;
		double mu = 0.0, norms[3] = {0.0, 0.0, 0.0};
		const bool valid = kb[d.ux_stride + d.pi_stride + 4*w.nbp] > 0.0 && d.nbtot>0;
		if(valid)
			{
			/* new b, [r q], bounds ; iterate := backup (d_ip2_res_hard.c:2138-2173) */
			S::extract(c, d, in_inst, w);
			for(long long i=lane; i<d.ux_stride; i+=32) ux[i] = kb[i];
			for(long long i=lane; i<d.pi_stride; i+=32) pi[i] = kb[d.ux_stride+i];
			{
			const double *kc = kb + d.ux_stride + d.pi_stride;
			for(int cc=lane; cc<d.nbtot; cc+=32)
				{
				w.v(CV_LAM_LO)[cc] = kc[cc]; w.v(CV_LAM_UP)[cc] = kc[w.nbp+cc];
				w.v(CV_T_LO)[cc] = kc[2*w.nbp+cc]; w.v(CV_T_UP)[cc] = kc[3*w.nbp+cc];
				}
			}
			__syncwarp();
			/* residuals at the backup with the new vectors (:2192), gradient from the stored t_inv (:2216) */
			S::residuals(c, d, in_inst, w, ux, pi, &mu, norms);
			__syncwarp();
			for(int cc=lane; cc<d.nbtot; cc+=32)
				w.v(CV_QXG)[cc] = w.v(CV_TINV_LO)[cc]*(w.v(CV_RM_LO)[cc] - w.v(CV_LAM_LO)[cc]*w.v(CV_RD_LO)[cc])
				                  - w.v(CV_TINV_UP)[cc]*(w.v(CV_RM_UP)[cc] + w.v(CV_LAM_UP)[cc]*w.v(CV_RD_UP)[cc]);
			__syncwarp();
			/* one solve with the stored factor, Pb from the new b (:2225) */
			S::trs_newb(c, d, in_inst, w, w.res_b, w.res_q, w.v(CV_QXG));
			__syncwarp();
			/* dt, dlam (:2236) and the full step (:2239) */
			hb_gen_values(lane, d, in_inst, w, w.dux);
			(void)hb_ipm_alpha<true>(lane, d, w, w.dux);
			__syncwarp();
			for(long long i=lane; i<d.ux_stride; i+=32) ux[i] += 1.0*w.dux[i];
			for(long long i=lane; i<d.pi_stride; i+=32) pi[i] += 1.0*w.dpi[i];
			for(int cc=lane; cc<d.nbtot; cc+=32)
				{
				w.v(CV_LAM_LO)[cc] += 1.0*w.v(CV_DLAM_LO)[cc]; w.v(CV_LAM_UP)[cc] += 1.0*w.v(CV_DLAM_UP)[cc];
				w.v(CV_T_LO)[cc] += 1.0*w.v(CV_DT_LO)[cc]; w.v(CV_T_UP)[cc] += 1.0*w.v(CV_DT_UP)[cc];
				}
			__syncwarp();
			S::emit(c, d, w, lam_all + inst*2*(long long)d.nbtot, t_all + inst*2*(long long)d.nbtot);
			}
		if(lane==0)
			{
			/* status 0, or -10: the IPM call left no phase-2 factor behind (the reference reads an unset backup then) */
			info[0] = 0.0; info[1] = valid ? 0.0 : -10.0;
			info[2] = norms[0]; info[3] = norms[1]; info[4] = norms[2]; info[5] = mu;
			}
		__syncwarp();
		}
	}

extern "C" long long hb_ipm_work_doubles(const hb_dims *d) { return hb_ipm_work_doubles_(*d); }

static const int hbi_shapes[HBI_NVAR][2] = { {24, 11}, {12, 5}, {8, 3} };

extern "C" int hb_ipm_fast_variant(int N, const int *nx, const int *nu, int nbtot)
	{
	if(getenv("HPMPC_B200_NO_FAST_IPM")!=NULL || nbtot<=0) return -1;
	for(int id=0; id<HBI_NVAR; id++)
		{
		int ok = (nx[0]==0) && N>=3;
		for(int n=0; n<N && ok; n++) ok = (nu[n]==hbi_shapes[id][1]) && (n==0 || nx[n]==hbi_shapes[id][0]);
		ok = ok && nx[N]==hbi_shapes[id][0];
		if(ok) return id;
		}
	return -1;
	}

template<class C> static void hbi_info(int N, int *smem_warp, long long *L_doubles)
	{ *smem_warp = (int)sizeof(double)*hbi_cfg<C>::PER_WARP; *L_doubles = (long long)(N+1)*C::LBUF; }

extern "C" int hb_ipm_fast_info(int id, int N, int *smem_warp, long long *L_doubles)
	{
	switch(id)
		{
		case 0: hbi_info<hbi_v0>(N, smem_warp, L_doubles); return 0;
		case 1: hbi_info<hbi_v1>(N, smem_warp, L_doubles); return 0;
		case 2: hbi_info<hbi_v2>(N, smem_warp, L_doubles); return 0;
		}
	return -1;
	}

/* ------------------------------------------------------------------------------------------------ */
/* factor only / solve with stored factors on the size-specialised sweeps (d_back_ric_rec_trf_tv_res,  */
/* d_back_ric_rec_trs_tv_res: lqcp_solvers/d_back_ric_rec.c:403, 564), one warp per instance.          */
/* The factor is kept in the sweeps' own column-packed form, (N+1)*LBUF doubles per instance.          */
/* ------------------------------------------------------------------------------------------------ */
template<class C>
__global__ void __launch_bounds__(256) hbi_ric_trf_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		double *__restrict__ L_all, long long L_stride)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hbi_ctx<C> c;
	c.init(hb_smem + (size_t)warp*hbi_cfg<C>::PER_WARP, lane, d);
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		hbi_backward<C>(c, d, in + inst*d.in_stride, L_all + inst*L_stride, nullptr, nullptr, nullptr, nullptr, nullptr);
		__syncwarp();
		}
	}

template<class C>
__global__ void __launch_bounds__(256) hbi_ric_trs_kernel(hb_dims d, long long n_inst, const double *__restrict__ in,
		const double *__restrict__ L_all, long long L_stride, double *__restrict__ ux_all, double *__restrict__ pi_all, double *__restrict__ work)
	{
	const int warp = threadIdx.x>>5, lane = threadIdx.x&31, nw = blockDim.x>>5;
	const long long gw = (long long)blockIdx.x*nw + warp, tw = (long long)gridDim.x*nw;
	hbi_ctx<C> c;
	c.init(hb_smem + (size_t)warp*hbi_cfg<C>::PER_WARP, lane, d);
	double *rq = work + gw*(d.ux_stride + d.pi_stride), *Pb = rq + d.ux_stride;
	for(long long inst=gw; inst<n_inst; inst+=tw)
		{
		const double *in_inst = in + inst*d.in_stride;
		const double *Lst = L_all + inst*L_stride;
		double *ux = ux_all + inst*d.ux_stride, *pi = pi_all + inst*d.pi_stride;
		for(int n=0; n<=d.N; n++)
			{
			const hb_stage s = d.st[n];
			const int nux = s.nu+s.nx;
			for(int i=lane; i<nux; i+=32) rq[s.off_ux+i] = in_inst[s.off_RSQ+HB_TRI(nux)+i];
			}
		__syncwarp();
		hbi_Pb_sweep<C>(c, in_inst, Lst, Pb);
		__syncwarp();
		hbi_trs_backward<C>(c, d, in_inst, Lst, rq, nullptr, Pb, ux);
		__syncwarp();
		hbi_forward<C, true>(c, in_inst, Lst, nullptr, ux, ux, pi);
		__syncwarp();
		}
	}

/* shape-only test (no bounds needed): which size-specialised sweep set serves this size pattern, -1 for none */
extern "C" int hb_ric_shape_variant(int N, const int *nx, const int *nu)
	{
	if(getenv("HPMPC_B200_NO_FAST_IPM")!=NULL || getenv("HPMPC_B200_NO_FAST_TRF")!=NULL) return -1;
	for(int id=0; id<HBI_NVAR; id++)
		{
		int ok = (nx[0]==0) && N>=3;
		for(int n=0; n<N && ok; n++) ok = (nu[n]==hbi_shapes[id][1]) && (n==0 || nx[n]==hbi_shapes[id][0]);
		ok = ok && nx[N]==hbi_shapes[id][0];
		if(ok) return id;
		}
	return -1;
	}

template<class C> static int hbi_trf_launch(const hb_dims *d, long long n_inst, const double *in, double *L, long long L_stride,
		int grid, int warps, cudaStream_t st)
	{
	const int smem = warps*(int)sizeof(double)*hbi_cfg<C>::PER_WARP;
	if(hb_prep(hbi_ric_trf_kernel<C>, smem)) return -1;
	hbi_ric_trf_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_inst, in, L, L_stride);
	HB_CK(cudaGetLastError());
	return 0;
	}
template<class C> static int hbi_trs_launch(const hb_dims *d, long long n_inst, const double *in, const double *L, long long L_stride,
		double *ux, double *pi, double *work, int grid, int warps, cudaStream_t st)
	{
	const int smem = warps*(int)sizeof(double)*hbi_cfg<C>::PER_WARP;
	if(hb_prep(hbi_ric_trs_kernel<C>, smem)) return -1;
	hbi_ric_trs_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_inst, in, L, L_stride, ux, pi, work);
	HB_CK(cudaGetLastError());
	return 0;
	}

/* work (trs only): (ux_stride + pi_stride) doubles per warp of the grid */
extern "C" int hb_launch_ric_trf_trs_fast(int id, int mode, const hb_dims *d, long long n_inst, const double *in, double *L, long long L_stride,
		double *ux, double *pi, double *work, int grid, int warps, void *stream)
	{
	cudaStream_t st = (cudaStream_t)stream;
	if(warps>8) return -3;
	switch(id)
		{
		case 0: return mode==0 ? hbi_trf_launch<hbi_v0>(d, n_inst, in, L, L_stride, grid, warps, st) : hbi_trs_launch<hbi_v0>(d, n_inst, in, L, L_stride, ux, pi, work, grid, warps, st);
		case 1: return mode==0 ? hbi_trf_launch<hbi_v1>(d, n_inst, in, L, L_stride, grid, warps, st) : hbi_trs_launch<hbi_v1>(d, n_inst, in, L, L_stride, ux, pi, work, grid, warps, st);
		case 2: return mode==0 ? hbi_trf_launch<hbi_v2>(d, n_inst, in, L, L_stride, grid, warps, st) : hbi_trs_launch<hbi_v2>(d, n_inst, in, L, L_stride, ux, pi, work, grid, warps, st);
		}
	return -2;
	}

/* doubles of per-slot work area; L_doubles = size of the factor stash of the variant in use */
extern "C" long long hb_ipm_work_doubles2(const hb_dims *d, long long L_doubles)
	{
	return hb_ipm_work_doubles_(*d) - d->L_stride + L_doubles;
	}

template<class S> static int hb_launch_ipm_t(int smem, const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int grid, int warps, int *counter, cudaStream_t st, double *kkt, long long kkt_stride)
	{
	if(kkt!=NULL)
		{
		if constexpr (S::has_kkt)
			{
			if(hb_prep(hb_ipm_kernel<S, true>, smem)) return -1;
			HB_CK(cudaMemsetAsync(counter, 0, sizeof(int), st));
			hb_ipm_kernel<S, true><<<grid, warps*32, smem, st>>>(*d, n_inst, in, k_max, mu0, mu_tol, alpha_min, warm_start,
					ux, pi, lam, t, info, work, work_stride, counter, kkt, kkt_stride);
			HB_CK(cudaGetLastError());
			return 0;
			}
		else return -4;
		}
	if(hb_prep(hb_ipm_kernel<S, false>, smem)) return -1;
	if(getenv("HPMPC_B200_VERBOSE"))
		{
		int nb = 0; cudaFuncAttributes fa;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hb_ipm_kernel<S, false>, warps*32, smem);
		cudaFuncGetAttributes(&fa, hb_ipm_kernel<S, false>);
		fprintf(stderr, "hpmpc_b200: ipm kernel: grid %d x %d threads, %d B dynamic + %zu B static smem, %d regs, %zu B local, %d CTAs/SM\n",
			grid, warps*32, smem, fa.sharedSizeBytes, fa.numRegs, fa.localSizeBytes, nb);
		}
	HB_CK(cudaMemsetAsync(counter, 0, sizeof(int), st));
	hb_ipm_kernel<S, false><<<grid, warps*32, smem, st>>>(*d, n_inst, in, k_max, mu0, mu_tol, alpha_min, warm_start,
			ux, pi, lam, t, info, work, work_stride, counter, nullptr, 0);
	HB_CK(cudaGetLastError());
	return 0;
	}

template<class S> static int hb_launch_kkt_t(int smem, const hb_dims *d, long long n_inst, const double *in, double *kkt, long long kkt_stride,
		double *ux, double *pi, double *lam, double *t, double *info, int grid, int warps, int *counter, cudaStream_t st)
	{
	if(hb_prep(hb_kkt_new_rhs_kernel<S>, smem)) return -1;
	HB_CK(cudaMemsetAsync(counter, 0, sizeof(int), st));
	hb_kkt_new_rhs_kernel<S><<<grid, warps*32, smem, st>>>(*d, n_inst, in, kkt, kkt_stride, ux, pi, lam, t, info, counter);
	HB_CK(cudaGetLastError());
	return 0;
	}

/* new right-hand side on the state left by hb_launch_ipm_kkt (chains only) */
extern "C" int hb_launch_kkt_new_rhs(const hb_dims *d, long long n_inst, const double *in, double *kkt, long long kkt_stride,
		double *ux, double *pi, double *lam, double *t, double *info, int grid, int warps, int *counter, int fast_id, void *stream)
	{
	if(d->nzM>64) return -2;
	if(warps>8) return -3;
	cudaStream_t st = (cudaStream_t)stream;
#define HB_KKT_ARGS d, n_inst, in, kkt, kkt_stride, ux, pi, lam, t, info, grid, warps, counter, st
	switch(fast_id)
		{
		case 0: return hb_launch_kkt_t<hb_sweeps_fast<hbi_v0> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v0>::PER_WARP, HB_KKT_ARGS);
		case 1: return hb_launch_kkt_t<hb_sweeps_fast<hbi_v1> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v1>::PER_WARP, HB_KKT_ARGS);
		case 2: return hb_launch_kkt_t<hb_sweeps_fast<hbi_v2> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v2>::PER_WARP, HB_KKT_ARGS);
		}
	return hb_launch_kkt_t<hb_sweeps_generic>(warps*hb_smem_bytes_per_warp(d), HB_KKT_ARGS);
#undef HB_KKT_ARGS
	}

extern "C" int hb_launch_ipm_kkt(const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int n_slots, int grid, int warps, int *counter, int fast_id, void *stream,
		double *kkt, long long kkt_stride);
extern "C" int hb_launch_ipm(const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int n_slots, int grid, int warps, int *counter, int fast_id, void *stream)
	{
	return hb_launch_ipm_kkt(d, n_inst, in, k_max, mu0, mu_tol, alpha_min, warm_start, ux, pi, lam, t, info, work, work_stride, n_slots,
			grid, warps, counter, fast_id, stream, NULL, 0);
	}

/* kkt != NULL: every instance also leaves its KKT state (factor, t_inv, backup of the last iterate) in kkt + inst*kkt_stride */
extern "C" int hb_launch_ipm_kkt(const hb_dims *d, long long n_inst, const double *in, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int n_slots, int grid, int warps, int *counter, int fast_id, void *stream,
		double *kkt, long long kkt_stride)
	{
	if(d->nzM>64) return -2;
	if(grid*warps>n_slots || warps>8) return -3;
	cudaStream_t st = (cudaStream_t)stream;
#define HB_IPM_ARGS d, n_inst, in, k_max, mu0, mu_tol, alpha_min, warm_start, ux, pi, lam, t, info, work, work_stride, grid, warps, counter, st, kkt, kkt_stride
	switch(fast_id)
		{
		case HB_IPM_TREE: return d->tn==NULL ? -4 : hb_launch_ipm_t<hb_sweeps_tree>(warps*hb_smem_bytes_per_warp(d), HB_IPM_ARGS);
		case 0: return hb_launch_ipm_t<hb_sweeps_fast<hbi_v0> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v0>::PER_WARP, HB_IPM_ARGS);
		case 1: return hb_launch_ipm_t<hb_sweeps_fast<hbi_v1> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v1>::PER_WARP, HB_IPM_ARGS);
		case 2: return hb_launch_ipm_t<hb_sweeps_fast<hbi_v2> >(warps*(int)sizeof(double)*hbi_cfg<hbi_v2>::PER_WARP, HB_IPM_ARGS);
		}
	return hb_launch_ipm_t<hb_sweeps_generic>(warps*hb_smem_bytes_per_warp(d), HB_IPM_ARGS);
#undef HB_IPM_ARGS
	}

