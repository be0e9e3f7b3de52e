/*
 * oracle/ref_harness.c -- TEST / BENCH INFRASTRUCTURE: times a library that exports HPMPC's own symbols
 * (oracle/_ref/libhpmpc_ref_{avx2,c99}.so = the unmodified reference) on the host cores, one private
 * workspace per thread (the reference is re-entrant: it never allocates, SURVEY.md section 8b "Threading").
 *
 * Follows the reference's own timing programs: pre-packed panel-major data, nrep calls of
 * d_back_ric_rec_sv_tv_res (test_problems/test_d_ric_mpc.c:540-560), flush-to-zero on
 * (test_problems/test_d_ric_libstr.c:162), and fortran_order_d_ip_ocp_hard_tv for the IPM
 * (test_problems/test_d_ip_hard_libstr.c:964).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#if defined(__x86_64__)
#include <xmmintrin.h>
#endif

typedef void (*ric_sv_fn)(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, int update_b, double **hpBAbt, double **b,
		int update_q, double **hpQ, double **q, double **bd, double **hpDCt, double **Qx, double **qx, double **hux,
		int compute_pi, double **hpi, int compute_Pb, double **hPb, double *memory, double *work);
typedef void (*ric_trf_fn)(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hpQ, double **hpDCt, double **Qx,
		double **bd, double *memory, double *work);
typedef void (*ric_trs_fn)(int N, int *nx, int *nu, int *nb, int **idxb, int *ng, double **hpBAbt, double **hb, double **hq, double **hpDCt,
		double **qx, double **hux, int compute_pi, double **hpi, int compute_Pb, double **hPb, double *memory, double *work);
typedef int (*size_fn)(int N, int *nx, int *nu, int *nb, int *ng);
typedef int (*ipm_fn)(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu, int *nb, int **hidxb, int *ng,
		int N2, int warm_start, double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r,
		double **lb, double **ub, double **C, double **D, double **lg, double **ug, double **x, double **u, double **pi,
		double **lam, double *inf_norm_res, void *work0, double *stat);
typedef int (*ipm_size_fn)(int N, int *nx, int *nu, int *nb, int **hidxb, int *ng, int N2);

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9*t.tv_nsec; }
static void *amalloc(size_t bytes) { void *p = NULL; if(posix_memalign(&p, 64, bytes+64)) return NULL; memset(p, 0, bytes+64); return p; }

typedef struct
	{
	int tid, n_threads, n_pass, N, *nx, *nu, *nb, **idxb, k_max, mode;
	long n_inst, inst_stride;
	const long *off;           /* sv: off_BAbt[N], off_RSQ[N+1] ; ipm: 10 x (N+1) */
	double *data;
	double mu0, mu_tol;
	void *fn, *fsz1, *fsz2, *fn2;
	const double *vec; long vec_stride, vec_b_off;      /* mode 2 (shared dynamics): per instance [r q] (ux layout) then b (pi layout) */
	double *ux_out, *pi_out; long ux_stride, pi_stride;
	double *lam_out; long lam_stride;          /* IPM: lam as the wrapper returns it, [lb ub] per stage, stage after stage */
	double *res_out;                           /* IPM: inf_norm_res[4] per instance */
	int *kk_out, *status_out;
	pthread_barrier_t *bar;
	double t0, t1;
	} job;

static void *worker(void *arg)
	{
	job *J = (job*)arg;
	int N = J->N, n, i, pass;
	long inst;
#if defined(__x86_64__)
	_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
#endif
	int *ng = calloc(N+1, sizeof(int));
	long lo = J->n_inst*J->tid/J->n_threads, hi = J->n_inst*(J->tid+1)/J->n_threads;
	double **p[18];
	for(i=0; i<18; i++) p[i] = calloc(N+2, sizeof(double*));
	if(J->mode==0)
		{
		int *nb0 = calloc(N+1, sizeof(int));
		int wsz = ((size_fn)J->fsz1)(N, J->nx, J->nu, nb0, ng), msz = ((size_fn)J->fsz2)(N, J->nx, J->nu, nb0, ng);
		double *work = amalloc(wsz), *mem = amalloc(msz);
		for(n=0; n<=N; n++)
			{
			int pnz = (J->nx[n]+J->nu[n]+1+3)/4*4, pnx1 = n<N ? (J->nx[n+1]+3)/4*4 : 4;
			p[2][n] = amalloc(sizeof(double)*pnz); p[3][n] = amalloc(sizeof(double)*pnx1); p[4][n] = amalloc(sizeof(double)*pnx1);
			p[5][n] = amalloc(64);
			}
		pthread_barrier_wait(J->bar);
		J->t0 = now();
		for(pass=0; pass<J->n_pass; pass++)
			for(inst=lo; inst<hi; inst++)
				{
				double *base = J->data + inst*J->inst_stride;
				for(n=0; n<N; n++) p[0][n] = base + J->off[n];
				for(n=0; n<=N; n++) p[1][n] = base + J->off[N+n];
				((ric_sv_fn)J->fn)(N, J->nx, J->nu, nb0, J->idxb, ng, 0, p[0], p[5], 0, p[1], p[5], p[5], p[5], p[5], p[5],
						p[2], 1, p[3], 1, p[4], mem, work);
				if(J->ux_out && pass==0)
					{
					long o = 0;
					for(n=0; n<=N; n++) { for(i=0; i<J->nx[n]+J->nu[n]; i++) J->ux_out[inst*J->ux_stride+o+i] = p[2][n][i]; o += J->nx[n]+J->nu[n]; }
					o = 0;
					for(n=0; n<N; n++) { for(i=0; i<J->nx[n+1]; i++) J->pi_out[inst*J->pi_stride+o+i] = p[3][n][i]; o += J->nx[n+1]; }
					}
				}
		J->t1 = now();
		}
	else if(J->mode==2)
		{
		/* shared dynamics: the matrices of instance 0 for everybody, factorised once per thread (outside the timed region would
		 * flatter the CPU: it is inside, once per pass), then d_back_ric_rec_trs_tv_res per instance with its own b, q */
		int *nb0 = calloc(N+1, sizeof(int));
		int wsz = ((size_fn)J->fsz1)(N, J->nx, J->nu, nb0, ng), msz = ((size_fn)J->fsz2)(N, J->nx, J->nu, nb0, ng);
		double *work = amalloc(wsz), *mem = amalloc(msz);
		for(n=0; n<=N; n++)
			{
			int pnz = (J->nx[n]+J->nu[n]+1+3)/4*4, pnx1 = n<N ? (J->nx[n+1]+3)/4*4 : 4;
			p[2][n] = amalloc(sizeof(double)*pnz); p[3][n] = amalloc(sizeof(double)*pnx1); p[4][n] = amalloc(sizeof(double)*pnx1);
			p[5][n] = amalloc(64); p[6][n] = amalloc(sizeof(double)*pnx1); p[7][n] = amalloc(sizeof(double)*pnz);
			}
		for(n=0; n<N; n++) p[0][n] = J->data + J->off[n];
		for(n=0; n<=N; n++) p[1][n] = J->data + J->off[N+n];
		pthread_barrier_wait(J->bar);
		J->t0 = now();
		for(pass=0; pass<J->n_pass; pass++)
			{
			((ric_trf_fn)J->fn)(N, J->nx, J->nu, nb0, J->idxb, ng, p[0], p[1], p[5], p[5], p[5], mem, work);
			for(inst=lo; inst<hi; inst++)
				{
				const double *v = J->vec + inst*J->vec_stride;
				long o = 0, ob = J->vec_b_off;
				for(n=0; n<=N; n++)
					{
					int nux = J->nx[n]+J->nu[n];
					for(i=0; i<nux; i++) p[7][n][i] = v[o+i];
					o += nux;
					if(n<N) { for(i=0; i<J->nx[n+1]; i++) p[6][n][i] = v[ob+i]; ob += J->nx[n+1]; }
					}
				((ric_trs_fn)J->fn2)(N, J->nx, J->nu, nb0, J->idxb, ng, p[0], p[6], p[7], p[5], p[5], p[2], 1, p[3], 1, p[4], mem, work);
				if(J->ux_out && pass==0)
					{
					long oo = 0;
					for(n=0; n<=N; n++) { for(i=0; i<J->nx[n]+J->nu[n]; i++) J->ux_out[inst*J->ux_stride+oo+i] = p[2][n][i]; oo += J->nx[n]+J->nu[n]; }
					}
				}
			}
		J->t1 = now();
		}
	else
		{
		int wsz = ((ipm_size_fn)J->fsz1)(N, J->nx, J->nu, J->nb, J->idxb, ng, N);
		void *work = amalloc(wsz);
		double *stat = amalloc(sizeof(double)*5*(J->k_max+1)), res[8];
		for(n=0; n<=N; n++)
			{
			p[14][n] = amalloc(sizeof(double)*(J->nx[n]+1)); p[15][n] = amalloc(sizeof(double)*(J->nu[n]+1));
			p[16][n] = amalloc(sizeof(double)*((n<N ? J->nx[n+1] : 0)+1)); p[17][n] = amalloc(sizeof(double)*(2*J->nb[n]+1));
			p[10][n] = p[11][n] = p[12][n] = p[13][n] = amalloc(64);
			}
		pthread_barrier_wait(J->bar);
		J->t0 = now();
		for(pass=0; pass<J->n_pass; pass++)
			for(inst=lo; inst<hi; inst++)
				{
				double *base = J->data + inst*J->inst_stride;
				for(i=0; i<10; i++) for(n=0; n<=N; n++) p[i][n] = base + J->off[i*(N+1)+n];
				int kk = 0;
				int st = ((ipm_fn)J->fn)(&kk, J->k_max, J->mu0, J->mu_tol, N, J->nx, J->nu, J->nb, J->idxb, ng, N, 0,
						p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13],
						p[14], p[15], p[16], p[17], res, work, stat);
				if(pass==0)
					{
					if(J->kk_out) { J->kk_out[inst] = kk; J->status_out[inst] = st; }
					if(J->ux_out)
						{
						long o = 0;
						for(n=0; n<=N; n++)
							{
							for(i=0; i<J->nu[n]; i++) J->ux_out[inst*J->ux_stride+o+i] = p[15][n][i];
							for(i=0; i<J->nx[n]; i++) J->ux_out[inst*J->ux_stride+o+J->nu[n]+i] = p[14][n][i];
							o += J->nx[n]+J->nu[n];
							}
						}
					if(J->pi_out)
						{
						long o = 0;
						for(n=0; n<N; n++) { for(i=0; i<J->nx[n+1]; i++) J->pi_out[inst*J->pi_stride+o+i] = p[16][n][i]; o += J->nx[n+1]; }
						}
					if(J->lam_out)
						{
						long o = 0;
						for(n=0; n<=N; n++) { for(i=0; i<2*J->nb[n]; i++) J->lam_out[inst*J->lam_stride+o+i] = p[17][n][i]; o += 2*J->nb[n]; }
						}
					if(J->res_out) for(i=0; i<4; i++) J->res_out[inst*4+i] = res[i];
					}
				}
		J->t1 = now();
		}
	return NULL;
	}

static const double *g_vec; static long g_vec_stride, g_vec_b_off;
static double run(const char *libpath, int mode, int n_threads, long n_inst, int n_pass, int N, int *nx, int *nu, int *nb,
		int *idxb_flat, int k_max, double mu0, double mu_tol, double *data, long inst_stride, const long *off,
		double *ux_out, long ux_stride, double *pi_out, long pi_stride, int *kk_out, int *status_out,
		double *lam_out, long lam_stride, double *res_out)
	{
	void *h = dlopen(libpath, RTLD_NOW|RTLD_LOCAL);
	if(!h) { fprintf(stderr, "ref_harness: cannot load %s: %s\n", libpath, dlerror()); return -1.0; }
	void *fn = dlsym(h, mode==0 ? "d_back_ric_rec_sv_tv_res" : (mode==2 ? "d_back_ric_rec_trf_tv_res" : "fortran_order_d_ip_ocp_hard_tv"));
	void *fn2 = dlsym(h, "d_back_ric_rec_trs_tv_res");
	void *f1 = dlsym(h, mode!=1 ? "d_back_ric_rec_sv_tv_work_space_size_bytes" : "hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes");
	void *f2 = dlsym(h, "d_back_ric_rec_sv_tv_memory_space_size_bytes");
	if(!fn || !f1 || !f2) { fprintf(stderr, "ref_harness: symbols missing in %s\n", libpath); return -1.0; }
	int n, t, o = 0;
	int **idxb = calloc(N+1, sizeof(int*));
	int *nbz = calloc(N+1, sizeof(int));
	for(n=0; n<=N; n++) { idxb[n] = idxb_flat ? idxb_flat + o : nbz; o += nb ? nb[n] : 0; }
	if(n_threads<1) n_threads = 1;
	if(n_threads>n_inst) n_threads = (int)n_inst;
	pthread_barrier_t bar;
	pthread_barrier_init(&bar, NULL, n_threads);
	job *J = calloc(n_threads, sizeof(job));
	pthread_t *th = calloc(n_threads, sizeof(pthread_t));
	for(t=0; t<n_threads; t++)
		{
		J[t].tid = t; J[t].n_threads = n_threads; J[t].n_pass = n_pass; J[t].N = N; J[t].nx = nx; J[t].nu = nu; J[t].nb = nb ? nb : nbz;
		J[t].idxb = idxb; J[t].k_max = k_max; J[t].mode = mode; J[t].n_inst = n_inst; J[t].inst_stride = inst_stride; J[t].off = off;
		J[t].data = data; J[t].mu0 = mu0; J[t].mu_tol = mu_tol; J[t].fn = fn; J[t].fsz1 = f1; J[t].fsz2 = f2; J[t].fn2 = fn2;
		J[t].vec = g_vec; J[t].vec_stride = g_vec_stride; J[t].vec_b_off = g_vec_b_off;
		J[t].ux_out = ux_out; J[t].pi_out = pi_out; J[t].ux_stride = ux_stride; J[t].pi_stride = pi_stride;
		J[t].kk_out = kk_out; J[t].status_out = status_out; J[t].bar = &bar;
		J[t].lam_out = lam_out; J[t].lam_stride = lam_stride; J[t].res_out = res_out;
		pthread_create(&th[t], NULL, worker, &J[t]);
		}
	double t0 = 1e300, t1 = 0.0;
	for(t=0; t<n_threads; t++)
		{
		pthread_join(th[t], NULL);
		if(J[t].t0<t0) t0 = J[t].t0;
		if(J[t].t1>t1) t1 = J[t].t1;
		}
	pthread_barrier_destroy(&bar);
	free(J); free(th); free(idxb); free(nbz);
	return t1-t0;        /* wall time from the first thread starting to the last thread finishing */
	}

/* panel-major instances: off = [off_BAbt[0..N-1], off_RSQ[0..N]] (doubles from the start of an instance) */
double ref_harness_ric_sv(const char *libpath, int n_threads, long n_inst, int n_pass, int N, int *nx, int *nu,
		double *pm_data, long inst_stride, const long *off, double *ux_out, long ux_stride, double *pi_out, long pi_stride)
	{
	return run(libpath, 0, n_threads, n_inst, n_pass, N, nx, nu, NULL, NULL, 0, 0.0, 0.0, pm_data, inst_stride, off,
			ux_out, ux_stride, pi_out, pi_stride, NULL, NULL, NULL, 0, NULL);
	}

/* column-major stage arrays: off = 10 x (N+1) table for A,B,b,Q,S,R,q,r,lb,ub */
double ref_harness_ipm(const char *libpath, int n_threads, long n_inst, int n_pass, int N, int *nx, int *nu, int *nb, int *idxb_flat,
		int k_max, double mu0, double mu_tol, double *data, long inst_stride, const long *off,
		int *kk_out, int *status_out, double *ux_out, long ux_stride)
	{
	return run(libpath, 1, n_threads, n_inst, n_pass, N, nx, nu, nb, idxb_flat, k_max, mu0, mu_tol, data, inst_stride, off,
			ux_out, ux_stride, NULL, 0, kk_out, status_out, NULL, 0, NULL);
	}

/* the same, also returning pi, lam and inf_norm_res of every instance (full-batch parity runs, tests/test_parity_hardening.py) */
double ref_harness_ipm_full(const char *libpath, int n_threads, long n_inst, int N, int *nx, int *nu, int *nb, int *idxb_flat,
		int k_max, double mu0, double mu_tol, double *data, long inst_stride, const long *off,
		int *kk_out, int *status_out, double *ux_out, long ux_stride, double *pi_out, long pi_stride, double *lam_out, long lam_stride,
		double *res_out)
	{
	return run(libpath, 1, n_threads, n_inst, 1, N, nx, nu, nb, idxb_flat, k_max, mu0, mu_tol, data, inst_stride, off,
			ux_out, ux_stride, pi_out, pi_stride, kk_out, status_out, lam_out, lam_stride, res_out);
	}

/* shared dynamics: pm_data = the ONE panel-major problem, vec = n_inst x vec_stride doubles ([r q] of every stage, then b of every
 * stage starting at vec_b_off); times (factor once per thread and pass) + one d_back_ric_rec_trs_tv_res per instance */
double ref_harness_ric_trs_shared(const char *libpath, int n_threads, long n_inst, int n_pass, int N, int *nx, int *nu,
		double *pm_data, const long *off, const double *vec, long vec_stride, long vec_b_off, double *ux_out, long ux_stride)
	{
	g_vec = vec; g_vec_stride = vec_stride; g_vec_b_off = vec_b_off;
	return run(libpath, 2, n_threads, n_inst, n_pass, N, nx, nu, NULL, NULL, 0, 0.0, 0.0, pm_data, 0, off,
			ux_out, ux_stride, NULL, 0, NULL, NULL, NULL, 0, NULL);
	}
