/*
 * oracle/ric_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded, dense column-major restatement of the HPMPC hot path:
 * backward Riccati recursion (factor / solve / factor+solve) and the two-phase Mehrotra
 * predictor-corrector IPM for box-constrained OCPs.  Nothing under hpmpc_b200/ may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do,
 * and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here against
 * the real reference (oracle/_ref/libhpmpc_ref_c99.so, built from /root/reference by oracle/Makefile)
 * on seeded mass-spring problems, and tests/golden/ holds vectors produced by that reference build
 * (tests/golden/make_golden.py).
 *
 * Reference lines followed (paths relative to /root/reference):
 *   chol_mn            kernel/c99/kernel_dpotrf_c99_lib4.c:553-640 (pivot rule >1e-15, inverse diagonal
 *                      multiplied), blas/blas_d_lib4.c:3622 (dsyrk_dpotrf_lib: m x n trapezoid)
 *   orc_ric_backward   lqcp_solvers/d_back_ric_rec.c:184-333 (sv), :470-556 (trf);
 *                      readable twin lqcp_solvers/d_back_ric_rec_libstr.c:89-181
 *   orc_ric_forward    lqcp_solvers/d_back_ric_rec.c:341-397 ; _libstr.c:187-221
 *   orc_ric_trs        lqcp_solvers/d_back_ric_rec.c:564-791 ; _libstr.c:309-424
 *   IPM element-wise   mpc_solvers/c99/d_aux_ip_hard_lib4.c:43 (init), :217 (update_hessian), :387
 *                      (update_gradient), :489 (compute_alpha), :618 (update_var), :715 (compute_mu),
 *                      :954 (update_hessian_gradient_res), :1180 (compute_alpha_res), :1382
 *                      (backup_update_var_res), :1453 (compute_mu_res), :1512 (centering), :1550
 *   residuals          mpc_solvers/c99/d_res_ip_res_hard.c:39 ; exit residuals mpc_solvers/d_res_ip_hard.c:38
 *   IPM driver         mpc_solvers/d_ip2_res_hard.c:116-1345
 *   high-level wrapper interfaces/c/fortran_order_interface.c:53-688 (packing, mu0 estimate, u=lb if lb==ub,
 *                      inf_norm_res, lam ordering [lb ub])
 *
 * General constraints lg <= D u + C x <= ug (ng > 0): the reference treats them exactly like bounds
 * with the bounded variable replaced by the product [D C] ux (mpc_solvers/c99/d_aux_ip_hard_lib4.c:121-147, :302-383, :556-607;
 * lqcp_solvers/d_back_ric_rec.c:293-315); here a stage owns nt = nb + ng constraints, box entries first.
 * Scope: N2 >= N runs the IPM on the problem as given; N2 < N condenses it first (orc_part_cond / orc_part_expand below).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------------------------- */
/* problem container: everything dense, column-major                                            */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
	int N;
	int *nx, *nu, *nb;      /* [N+1], nu[N] = 0 ; nb[n] = ALL constraints of the stage: nbx[n] bounds, then ng[n] general ones */
	int *nbx, *ng;          /* [N+1] */
	double **DCt;           /* [N+1] (nux) x ng, ld = nux : [D C]' (NULL entries when ng = 0) */
	int **idxb;             /* [N+1][nbx]       */
	double **BAbt;          /* [N]   (nux+1) x nx1, ld = nux+1 ; rows: B' , A' , b'   */
	double **RSQrq;         /* [N+1] (nux+1) x nux, ld = nux+1 ; [R S';S Q] lower + last row [r' q'] */
	double **d;             /* [N+1] [lower(nb) ; upper(nb)], lower = [lb ; lg], upper = [ub ; ug] */
	/* factorization memory */
	double **L;             /* [N+1] (nux+1) x nux, ld = nux+1 */
	double **dinv;          /* [N+1] nux : inverse diagonal of L */
	int nzM, nxM;
	double *W;              /* scratch (nzM x nxM) */
	double *tmp;            /* scratch 2*nzM */
	/* scenario tree (NULL for a chain): stages are the nodes in BFS order, dad[0] = -1, and "edge n" (BAbt[n], b, pi[n],
	 * Pb[n], n = 0..N-1) is the edge INTO node n+1, whose rows are those of dad[n+1] instead of stage n */
	int *dad;
	int *fk, *nk;           /* first kid and number of kids of every stage/node (kids are contiguous in BFS order; chain: n+1, 1) */
	/* what the IPM leaves behind for a later solve with a new right-hand side (the reference keeps the same in its work
	 * space: ux_bkp, pi_bkp, t_bkp, lam_bkp, t_inv and the factor, mpc_solvers/d_ip2_res_hard.c:2019-2133) */
	double **k_ux, **k_pi, **k_lam, **k_t, **k_tinv;
	int k_valid;            /* 1 once a phase-2 iteration has run */
} orc_prob;

static int nux_(const orc_prob *P, int n) { return P->nu[n] + P->nx[n]; }
/* value of constraint j of stage n for the vector v (ux layout): the bounded variable, or row j-nbx of [D C] times v */
static double cval(const orc_prob *P, int n, const double *v, int j)
	{
	if(j<P->nbx[n]) return v[P->idxb[n][j]];
	int nux = nux_(P, n), i; const double *g = P->DCt[n] + (size_t)nux*(j-P->nbx[n]);
	double s = 0.0;
	for(i=0; i<nux; i++) s += g[i]*v[i];
	return s;
	}
/* v += a * (gradient of constraint j of stage n) */
static void cscatter(const orc_prob *P, int n, double *v, int j, double a)
	{
	if(j<P->nbx[n]) { v[P->idxb[n][j]] += a; return; }
	int nux = nux_(P, n), i; const double *g = P->DCt[n] + (size_t)nux*(j-P->nbx[n]);
	for(i=0; i<nux; i++) v[i] += a*g[i];
	}
static int dad_(const orc_prob *P, int k) { return P->dad ? P->dad[k] : k-1; }

orc_prob *orc_prob_create_gen(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *dad, const int *ng);
orc_prob *orc_prob_create_tree(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *dad)
	{
	return orc_prob_create_gen(N, nx, nu, nb, idxb, dad, NULL);
	}
orc_prob *orc_prob_create(int N, const int *nx, const int *nu, const int *nb, int *const *idxb)
	{
	return orc_prob_create_gen(N, nx, nu, nb, idxb, NULL, NULL);
	}

/* nb[n] = number of bounds of stage n, ng[n] (may be NULL) = number of general constraints */
orc_prob *orc_prob_create_gen(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *dad, const int *ng)
	{
	orc_prob *P = calloc(1, sizeof(orc_prob));
	int n, j;
	P->N = N;
	if(dad) { P->dad = malloc((N+1)*sizeof(int)); for(n=0; n<=N; n++) P->dad[n] = dad[n]; }
	P->fk = malloc((N+1)*sizeof(int)); P->nk = calloc(N+1, sizeof(int));
	for(n=0; n<=N; n++) P->fk[n] = N+1;
	for(n=1; n<=N; n++)
		{
		int dd = dad ? dad[n] : n-1;
		if(P->nk[dd]==0) P->fk[dd] = n;
		if(n!=P->fk[dd]+P->nk[dd]) { fprintf(stderr, "oracle: kids of node %d are not contiguous\n", dd); abort(); }
		P->nk[dd]++;
		}
	P->nx = malloc((N+1)*sizeof(int)); P->nu = malloc((N+1)*sizeof(int)); P->nb = malloc((N+1)*sizeof(int));
	P->nbx = malloc((N+1)*sizeof(int)); P->ng = malloc((N+1)*sizeof(int)); P->DCt = calloc(N+1, sizeof(double*));
	for(n=0; n<=N; n++)
		{
		P->nx[n] = nx[n]; P->nu[n] = (n<N || dad) ? nu[n] : 0;
		P->nbx[n] = nb ? nb[n] : 0; P->ng[n] = ng ? ng[n] : 0; P->nb[n] = P->nbx[n] + P->ng[n];
		}
	P->idxb = calloc(N+1, sizeof(int*));
	for(n=0; n<=N; n++)
		{
		P->idxb[n] = malloc((P->nbx[n]+1)*sizeof(int));
		for(j=0; j<P->nbx[n]; j++) P->idxb[n][j] = idxb[n][j];
		P->DCt[n] = calloc((size_t)(P->nu[n]+P->nx[n]+1)*(P->ng[n]+1), sizeof(double));
		}
	P->BAbt = calloc(N+1, sizeof(double*)); P->RSQrq = calloc(N+1, sizeof(double*)); P->d = calloc(N+1, sizeof(double*));
	P->L = calloc(N+1, sizeof(double*)); P->dinv = calloc(N+1, sizeof(double*));
	P->nzM = 1; P->nxM = 1;
	for(n=0; n<=N; n++)
		{
		int nux = nux_(P, n), nz = nux+1;
		if(nz>P->nzM) P->nzM = nz;
		if(P->nx[n]>P->nxM) P->nxM = P->nx[n];
		if(n<N) P->BAbt[n] = calloc((size_t)(nux_(P, dad_(P, n+1))+1)*(P->nx[n+1]+1), sizeof(double));
		P->RSQrq[n] = calloc((size_t)nz*(nux+1), sizeof(double));
		P->L[n] = calloc((size_t)nz*(nux+1), sizeof(double));
		P->dinv[n] = calloc(nux+1, sizeof(double));
		P->d[n] = calloc(2*P->nb[n]+1, sizeof(double));
		}
	P->W = calloc((size_t)P->nzM*P->nxM, sizeof(double));
	P->tmp = calloc(4*P->nzM, sizeof(double));
	P->k_ux = calloc(N+1, sizeof(double*)); P->k_pi = calloc(N+1, sizeof(double*)); P->k_lam = calloc(N+1, sizeof(double*));
	P->k_t = calloc(N+1, sizeof(double*)); P->k_tinv = calloc(N+1, sizeof(double*));
	for(n=0; n<=N; n++)
		{
		P->k_ux[n] = calloc(nux_(P, n)+1, sizeof(double)); P->k_pi[n] = calloc(P->nxM+1, sizeof(double));
		P->k_lam[n] = calloc(2*P->nb[n]+1, sizeof(double)); P->k_t[n] = calloc(2*P->nb[n]+1, sizeof(double));
		P->k_tinv[n] = calloc(2*P->nb[n]+1, sizeof(double));
		}
	return P;
	}

void orc_prob_free(orc_prob *P)
	{
	int n;
	for(n=0; n<=P->N; n++)
		{
		free(P->idxb[n]); free(P->BAbt[n]); free(P->RSQrq[n]); free(P->L[n]); free(P->dinv[n]); free(P->d[n]); free(P->DCt[n]);
		free(P->k_ux[n]); free(P->k_pi[n]); free(P->k_lam[n]); free(P->k_t[n]); free(P->k_tinv[n]);
		}
	free(P->k_ux); free(P->k_pi); free(P->k_lam); free(P->k_t); free(P->k_tinv);
	free(P->idxb); free(P->BAbt); free(P->RSQrq); free(P->L); free(P->dinv); free(P->d);
	free(P->nx); free(P->nu); free(P->nb); free(P->nbx); free(P->ng); free(P->DCt); free(P->W); free(P->tmp); free(P->dad); free(P->fk); free(P->nk); free(P);
	}

/* fill from the stage-wise column-major ("fortran order") arrays of the high-level API:
 * A[n] nx1 x nx, B[n] nx1 x nu, S[n] nu x nx  (interfaces/c/fortran_order_interface.c:262-311) */
void orc_prob_set(orc_prob *P, double *const *A, double *const *B, double *const *b,
		double *const *Q, double *const *S, double *const *R, double *const *q, double *const *r,
		double *const *lb, double *const *ub)
	{
	int n, i, j;
	for(n=0; n<=P->N; n++)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, nz = nux+1;
		if(n<P->N)
			{
			int nx1 = P->nx[n+1];
			double *M = P->BAbt[n];
			for(j=0; j<nx1; j++)
				{
				for(i=0; i<nu; i++) M[i+nz*j] = B[n][j+nx1*i];
				for(i=0; i<nx; i++) M[nu+i+nz*j] = A[n][j+nx1*i];
				M[nux+nz*j] = b[n][j];
				}
			}
		double *H = P->RSQrq[n];
		memset(H, 0, sizeof(double)*nz*(nux+1));
		for(j=0; j<nu; j++) for(i=0; i<nu; i++) H[i+nz*j] = R[n][i+nu*j];
		for(j=0; j<nu; j++) for(i=0; i<nx; i++) H[nu+i+nz*j] = S[n][j+nu*i];       /* S' below R */
		for(j=0; j<nx; j++) for(i=0; i<nx; i++) H[nu+i+nz*(nu+j)] = Q[n][i+nx*j];
		for(j=0; j<nu; j++) H[nux+nz*j] = r[n][j];
		for(j=0; j<nx; j++) H[nux+nz*(nu+j)] = q[n][j];
		for(j=0; j<P->nbx[n]; j++) { P->d[n][j] = lb[n][j]; P->d[n][P->nb[n]+j] = ub[n][j]; }
		}
	}

/* general constraints from the column-major arrays of the high-level API: C[n] ng x nx, D[n] ng x nu (n < N)
 * (interfaces/c/fortran_order_interface.c:276-283, :371-378) */
void orc_prob_set_general(orc_prob *P, double *const *C, double *const *D, double *const *lg, double *const *ug)
	{
	int n, i, j;
	for(n=0; n<=P->N; n++)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, ng = P->ng[n], nbx = P->nbx[n], nt = P->nb[n];
		for(j=0; j<ng; j++)
			{
			for(i=0; i<nu; i++) P->DCt[n][i+nux*j] = D[n][j+ng*i];
			for(i=0; i<nx; i++) P->DCt[n][nu+i+nux*j] = C[n][j+ng*i];
			P->d[n][nbx+j] = lg[n][j]; P->d[n][nt+nbx+j] = ug[n][j];
			}
		}
	}

/* ------------------------------------------------------------------------------------------- */
/* m x n "trapezoidal" Cholesky: top n x n factorized, remaining rows solved against it         */
/* ------------------------------------------------------------------------------------------- */
static void chol_mn(int m, int n, double *A, int lda, double *dinv)
	{
	int i, j, k;
	for(j=0; j<n; j++)
		{
		double c = A[j+lda*j];
		for(k=0; k<j; k++) c -= A[j+lda*k]*A[j+lda*k];
		double inv;
		if(c>1e-15) { c = sqrt(c); inv = 1.0/c; }
		else { c = 0.0; inv = 0.0; }
		A[j+lda*j] = c;
		dinv[j] = inv;
		for(i=j+1; i<m; i++)
			{
			double v = A[i+lda*j];
			for(k=0; k<j; k++) v -= A[i+lda*k]*A[j+lda*k];
			A[i+lda*j] = v*inv;
			}
		}
	}

/* ------------------------------------------------------------------------------------------- */
/* backward recursion.  with_grad = 1: sv (gradient row carried) ; 0: trf                       */
/*   bvec  : if non-NULL, b_n replaces the last row of BAbt_n      (update_b)                    */
/*   rqvec : if non-NULL, rq_n replaces the last row of RSQrq_n    (update_q)                    */
/*   Qx,qx : if non-NULL, diag[idxb] += Qx , lastrow[idxb] += qx   (IPM box terms)               */
/*   Pb    : if non-NULL, Pb[n] = P_{n+1} b_n  (edge-indexed like lib4)                          */
/* ------------------------------------------------------------------------------------------- */
void orc_ric_backward(orc_prob *P, int with_grad, double *const *bvec, double *const *rqvec,
		double *const *Qx, double *const *qx, double **Pb)
	{
	int N = P->N, n, i, j, k;
	for(n=N; n>=0; n--)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, nz = nux+1;
		int m = with_grad ? nz : nux;
		double *L = P->L[n], *H = P->RSQrq[n];
		/* L <- lower part of H (+ box terms) */
		for(j=0; j<nux; j++) for(i=j; i<m; i++) L[i+nz*j] = H[i+nz*j];
		if(with_grad && rqvec) for(j=0; j<nux; j++) L[nux+nz*j] = rqvec[n][j];
		if(P->nb[n]>0 && Qx)
			for(j=0; j<P->nb[n]; j++)
				{
				if(j<P->nbx[n])
					{
					int id = P->idxb[n][j];
					L[id+nz*id] += Qx[n][j];
					if(with_grad && qx) L[nux+nz*id] += qx[n][j];
					}
				else
					{
					/* general constraint: + Qx g g' on the Hessian, + qx g' on the gradient row (d_back_ric_rec.c:293-315) */
					const double *g = P->DCt[n] + (size_t)nux*(j-P->nbx[n]);
					int c2;
					for(c2=0; c2<nux; c2++)
						{
						for(i=c2; i<nux; i++) L[i+nz*c2] += Qx[n][j]*g[i]*g[c2];
						if(with_grad && qx) L[nux+nz*c2] += qx[n][j]*g[c2];
						}
					}
				}
		/* every kid c of this node (a chain has the one kid n+1; lqcp_solvers/d_tree_back_ric_rec_libstr.c:79-156 sums
		 * W_c W_c' over the kids); e = c-1 is the edge into the kid */
		for(int c=P->fk[n]; c<P->fk[n]+P->nk[n]; c++)
			{
			const int e = c-1;
			int nx1 = P->nx[c], nu1 = P->nu[c], nz1 = nx1+nu1+1;
			double *Ln = P->L[c], *M = P->BAbt[e], *W = P->W;
			/* W = BAbt * Lxx_c   (m x nx1) */
			for(i=0; i<m; i++)
				for(j=0; j<nx1; j++)
					{
					double s = 0.0;
					for(k=j; k<nx1; k++)
						{
						double a = (i==nux && bvec) ? bvec[e][k] : M[i+nz*k];
						s += a*Ln[nu1+k+nz1*(nu1+j)];
						}
					W[i+nz*j] = s;
					}
			if(with_grad)
				{
				if(Pb) /* Pb = Lxx * (last row of W)' */
					for(i=0; i<nx1; i++)
						{
						double s = 0.0;
						for(k=0; k<=i; k++) s += Ln[nu1+i+nz1*(nu1+k)]*W[nux+nz*k];
						Pb[e][i] = s;
						}
				for(j=0; j<nx1; j++) W[nux+nz*j] += Ln[nu1+nx1+nz1*(nu1+j)];   /* + l_x */
				}
			/* L += W W' (lower, m rows) */
			for(j=0; j<nux; j++)
				for(i=j; i<m; i++)
					{
					double s = 0.0;
					for(k=0; k<nx1; k++) s += W[i+nz*k]*W[j+nz*k];
					L[i+nz*j] += s;
					}
			}
		chol_mn(m, nux, L, nz, P->dinv[n]);
		}
	}

/* x-part helpers on stage n: Lxx_n is L[nu:nu+nx, nu:nu+nx], l_x is L[nux, nu:] */
static void pi_from_x(const orc_prob *P, int n, const double *x, const double *p, double *pi, double *tmp)
	{
	/* pi = p + Lxx (Lxx' x) */
	int nx = P->nx[n], nu = P->nu[n], nz = nx+nu+1, i, k;
	const double *L = P->L[n];
	for(i=0; i<nx; i++)
		{
		double s = 0.0;
		for(k=i; k<nx; k++) s += L[nu+k+nz*(nu+i)]*x[k];
		tmp[i] = s;
		}
	for(i=0; i<nx; i++)
		{
		double s = p ? p[i] : 0.0;
		for(k=0; k<=i; k++) s += L[nu+i+nz*(nu+k)]*tmp[k];
		pi[i] = s;
		}
	}

/* forward substitution shared by sv and trs.
 *   lrow[n] : the "gradient" for stage n (sv: last row of L_n ; trs: eliminated rhs w_n), length nux
 *   bsrc    : b_n (length nx1) for the state recursion
 *   p[n]    : x-part of the eliminated rhs of stage n+1 (NULL in sv, where l_x of L_{n+1} is used)   */
static void ric_forward(orc_prob *P, double *const *lrow, double *const *bsrc, double *const *p,
		double **ux, int compute_pi, double **pi)
	{
	int N = P->N, n, i, j;
	/* node by node (for a chain: stage by stage): first the state and multiplier of the edge into the node, from the
	 * finished ux of its dad, then the node's own inputs */
	for(n=0; n<=N; n++)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, nz = nux+1;
		if(n>0)
			{
			const int e = n-1, dd = dad_(P, n);
			const int nuxd = nux_(P, dd), nzd = nuxd+1;
			/* x_n = b + BAbt[:nuxd]' ux_dad */
			const double *M = P->BAbt[e], *vd = ux[dd];
			double *xn = ux[n]+nu;
			for(j=0; j<nx; j++)
				{
				double s = bsrc ? bsrc[e][j] : M[nuxd+nzd*j];
				for(i=0; i<nuxd; i++) s += M[i+nzd*j]*vd[i];
				xn[j] = s;
				}
			if(compute_pi)
				{
				if(p) pi_from_x(P, n, xn, p[e], pi[e], P->tmp);
				else
					{
					/* sv: p = Lxx l_x  folded as  Lxx (Lxx' x + l_x) */
					int k;
					const double *L1 = P->L[n];
					double *tmp = P->tmp;
					for(i=0; i<nx; i++)
						{
						double s = L1[nu+nx+nz*(nu+i)];
						for(k=i; k<nx; k++) s += L1[nu+k+nz*(nu+i)]*xn[k];
						tmp[i] = s;
						}
					for(i=0; i<nx; i++)
						{
						double s = 0.0;
						for(k=0; k<=i; k++) s += L1[nu+i+nz*(nu+k)]*tmp[k];
						pi[e][i] = s;
						}
					}
				}
			}
		if(!P->dad && n==N) break;            /* the last stage of a chain has no inputs */
		int ks = (n==0) ? nux : nu;           /* stage 0 solves for all of ux_0 */
		const double *L = P->L[n], *dinv = P->dinv[n];
		double *v = ux[n];
		for(i=0; i<ks; i++) v[i] = -lrow[n][i];
		/* v[:ks] = L[:ks,:ks]^{-T} ( v[:ks] - L[ks:nux,:ks]' v[ks:nux] ) */
		for(i=ks-1; i>=0; i--)
			{
			double s = v[i];
			for(j=i+1; j<nux; j++) s -= L[j+nz*i]*v[j];
			v[i] = s*dinv[i];
			}
		}
	}

/* factor + solve (d_back_ric_rec_sv_tv_res).  pi[n], Pb[n] edge-indexed n=0..N-1. */
void orc_ric_sv(orc_prob *P, double *const *bvec, double *const *rqvec, double *const *Qx, double *const *qx,
		double **ux, int compute_pi, double **pi, double **Pb)
	{
	int n;
	orc_ric_backward(P, 1, bvec, rqvec, Qx, qx, Pb);
	/* gradient rows as vectors */
	double **lrow = calloc(P->N+1, sizeof(double*));
	for(n=0; n<=P->N; n++)
		{
		int nux = nux_(P,n), nz = nux+1, j;
		lrow[n] = malloc((nux+1)*sizeof(double));
		for(j=0; j<nux; j++) lrow[n][j] = P->L[n][nux+nz*j];
		}
	ric_forward(P, lrow, bvec, NULL, ux, compute_pi, pi);
	for(n=0; n<=P->N; n++) free(lrow[n]);
	free(lrow);
	}

void orc_ric_trf(orc_prob *P, double *const *Qx)
	{
	orc_ric_backward(P, 0, NULL, NULL, Qx, NULL, NULL);
	}

/* solve with stored factorization (d_back_ric_rec_trs_tv_res).  bvec[n] (nx1), rqvec[n] (nux) required. */
void orc_ric_trs(orc_prob *P, double *const *bvec, double *const *rqvec, double *const *qx,
		double **ux, int compute_pi, double **pi, int compute_Pb, double **Pb)
	{
	int N = P->N, n, i, j, k;
	double **w = calloc(N+1, sizeof(double*));
	double **p = calloc(N+1, sizeof(double*));
	for(n=0; n<=N; n++) { w[n] = calloc(nux_(P,n)+1, sizeof(double)); p[n] = calloc(P->nxM+1, sizeof(double)); }
	/* backward vector sweep */
	for(n=N; n>=0; n--)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, nz = nux+1;
		for(i=0; i<nux; i++) w[n][i] = rqvec[n][i];
		if(P->nb[n]>0 && qx) for(j=0; j<P->nb[n]; j++) cscatter(P, n, w[n], j, qx[n][j]);
		int has_kid = 0;
		for(int c=P->fk[n]; c<P->fk[n]+P->nk[n]; c++)
			{
			const int e = c-1;
			has_kid = 1;
			int nx1 = P->nx[c], nu1 = P->nu[c];
			double *tmp = P->tmp;
			if(compute_Pb)
				{
				double *t2 = P->tmp+P->nzM;
				int nz1 = nx1+nu1+1;
				const double *L1 = P->L[c];
				for(i=0; i<nx1; i++)
					{
					double s = 0.0;
					for(k=i; k<nx1; k++) s += L1[nu1+k+nz1*(nu1+i)]*bvec[e][k];
					t2[i] = s;
					}
				for(i=0; i<nx1; i++)
					{
					double s = 0.0;
					for(k=0; k<=i; k++) s += L1[nu1+i+nz1*(nu1+k)]*t2[k];
					Pb[e][i] = s;
					}
				}
			for(j=0; j<nx1; j++) tmp[j] = Pb[e][j] + w[c][nu1+j];
			const double *M = P->BAbt[e];
			for(i=0; i<nux; i++)
				{
				double s = w[n][i];
				for(j=0; j<nx1; j++) s += M[i+nz*j]*tmp[j];
				w[n][i] = s;
				}
			}
		if(has_kid)
			{
			int ks = (n==0) ? nux : nu;
			const double *L = P->L[n], *dinv = P->dinv[n];
			for(i=0; i<ks; i++)
				{
				double s = w[n][i];
				for(j=0; j<i; j++) s -= L[i+nz*j]*w[n][j];
				w[n][i] = s*dinv[i];
				}
			for(i=ks; i<nux; i++)
				{
				double s = w[n][i];
				for(j=0; j<ks; j++) s -= L[i+nz*j]*w[n][j];
				w[n][i] = s;
				}
			}
		}
	/* p[n] = x-part of w_{n+1} */
	for(n=0; n<N; n++) for(i=0; i<P->nx[n+1]; i++) p[n][i] = w[n+1][P->nu[n+1]+i];
	ric_forward(P, w, bvec, p, ux, compute_pi, pi);
	for(n=0; n<=N; n++) { free(w[n]); free(p[n]); }
	free(w); free(p);
	}

/* ------------------------------------------------------------------------------------------- */
/* IPM                                                                                          */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
	double **ux, **pi, **lam, **t;            /* iterate: lam,t = [lower(nb) ; upper(nb)] */
	double **dux, **dpi, **dlam, **dt, **tinv, **lamt, **Qx, **qx, **Pb;
	double **b, **rq;                          /* copies of b_n and [r;q]_n as vectors */
	double **res_q, **res_b, **res_d, **res_m;
} orc_ipm_ws;

static double **vecs(int n, const int *len) { double **v = malloc(n*sizeof(double*)); for(int i=0;i<n;i++) v[i]=calloc(len[i]+1,sizeof(double)); return v; }
static void vecs_free(double **v, int n) { for(int i=0;i<n;i++) free(v[i]); free(v); }

/* res_q,res_b,res_d,res_m, mu   (mpc_solvers/c99/d_res_ip_res_hard.c:39) */
static void ipm_residuals(const orc_prob *P, const orc_ipm_ws *w, double *mu)
	{
	int N = P->N, n, i, j;
	double mu2 = 0.0; int nb_tot = 0;
	for(n=0; n<=N; n++)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, nz = nux+1, nb = P->nb[n];
		double *rq = w->res_q[n];
		for(i=0; i<nux; i++) rq[i] = w->rq[n][i];
		if(n>0) for(i=0; i<nx; i++) rq[nu+i] -= w->pi[n-1][i];
		nb_tot += nb;
		for(j=0; j<nb; j++)
			{
			const double v = cval(P, n, w->ux[n], j);
			cscatter(P, n, rq, j, -w->lam[n][j] + w->lam[n][nb+j]);
			w->res_d[n][j]    = P->d[n][j]    - v + w->t[n][j];
			w->res_d[n][nb+j] = P->d[n][nb+j] - v - w->t[n][nb+j];
			w->res_m[n][j]    = w->lam[n][j]*w->t[n][j];
			w->res_m[n][nb+j] = w->lam[n][nb+j]*w->t[n][nb+j];
			mu2 += w->res_m[n][j] + w->res_m[n][nb+j];
			}
		/* rq += H ux  (H symmetric, lower stored) */
		const double *H = P->RSQrq[n];
		for(i=0; i<nux; i++)
			{
			double s = 0.0;
			for(j=0; j<nux; j++) s += (i>=j ? H[i+nz*j] : H[j+nz*i])*w->ux[n][j];
			rq[i] += s;
			}
		/* every edge out of this node (chain: the one into n+1; tree: mpc_solvers/d_tree_res_ip_res_hard_libstr.c:66) */
		for(int c=P->fk[n]; c<P->fk[n]+P->nk[n]; c++)
			{
			const int e = c-1;
			int nx1 = P->nx[c], nu1 = P->nu[c];
			const double *M = P->BAbt[e];
			for(j=0; j<nx1; j++)
				{
				double s = w->b[e][j] - w->ux[c][nu1+j];
				for(i=0; i<nux; i++) s += M[i+nz*j]*w->ux[n][i];
				w->res_b[e][j] = s;
				}
			for(i=0; i<nux; i++)
				{
				double s = 0.0;
				for(j=0; j<nx1; j++) s += M[i+nz*j]*w->pi[e][j];
				rq[i] += s;
				}
			}
		}
	if(nb_tot!=0) *mu = mu2/(2.0*nb_tot);
	}

/* the IPM proper (mpc_solvers/d_ip2_res_hard.c:116).  ux/pi/lam/t in w are in/out.
 * newton = 1: d_ip2_res_mpc_hard_tv_single_newton_step (:1348): the iterate is taken as given (no initialisation, no phase 1),
 * k_max residual-based steps with the centering term fixed at mu0 (:1749), sigma stays 0, return value :1911-1918. */
static int ipm_impl(orc_prob *P, int *kk, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *stat, double **ux, double **pi, double **lam, double **t, int newton);
int orc_ip2_res_mpc_hard(orc_prob *P, int *kk, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *stat, double **ux, double **pi, double **lam, double **t)
	{
	return ipm_impl(P, kk, k_max, mu0, mu_tol, alpha_min, warm_start, stat, ux, pi, lam, t, 0);
	}
int orc_ip2_res_mpc_hard_single_newton_step(orc_prob *P, int *kk, int k_max, double mu0, double alpha_min, double *stat,
		double **ux, double **pi, double **lam, double **t)
	{
	return ipm_impl(P, kk, k_max, mu0, 0.0, alpha_min, 1, stat, ux, pi, lam, t, 1);
	}
static int ipm_impl(orc_prob *P, int *kk, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *stat, double **ux, double **pi, double **lam, double **t, int newton)
	{
	int N = P->N, n, i, j;
	int *lnux = malloc((N+1)*sizeof(int)), *lnx1 = malloc((N+1)*sizeof(int)), *l2nb = malloc((N+1)*sizeof(int)), *lnb = malloc((N+1)*sizeof(int));
	for(n=0; n<=N; n++) { lnux[n] = nux_(P,n)+1; lnx1[n] = n<N ? P->nx[n+1] : 0; l2nb[n] = 2*P->nb[n]; lnb[n] = P->nb[n]; }
	orc_ipm_ws W, *w = &W;
	w->ux = ux; w->pi = pi; w->lam = lam; w->t = t;
	w->dux = vecs(N+1, lnux); w->dpi = vecs(N+1, lnx1); w->dlam = vecs(N+1, l2nb); w->dt = vecs(N+1, l2nb);
	w->tinv = vecs(N+1, l2nb); w->lamt = vecs(N+1, l2nb); w->Qx = vecs(N+1, lnb); w->qx = vecs(N+1, lnb);
	w->Pb = vecs(N+1, lnx1); w->b = vecs(N+1, lnx1); w->rq = vecs(N+1, lnux);
	w->res_q = vecs(N+1, lnux); w->res_b = vecs(N+1, lnx1); w->res_d = vecs(N+1, l2nb); w->res_m = vecs(N+1, l2nb);
	for(n=0; n<=N; n++)
		{
		int nux = nux_(P,n), nz = nux+1;
		for(j=0; j<nux; j++) w->rq[n][j] = P->RSQrq[n][nux+nz*j];
		if(n<N) { int nuxd = nux_(P, dad_(P, n+1)); for(j=0; j<P->nx[n+1]; j++) w->b[n][j] = P->BAbt[n][nuxd+(nuxd+1)*j]; }
		}

	int status = -1;
	double mu_scal = 0.0;
	for(n=0; n<=N; n++) mu_scal += 2*P->nb[n];
	if(mu_scal==0.0)
		{
		orc_ric_sv(P, NULL, NULL, NULL, NULL, ux, 1, pi, w->Pb);
		*kk = 0;
		status = 0;
		goto done;
		}
	mu_scal = 1.0/mu_scal;
	double sigma = 0.0, alpha, mu, mu_aff = 0.0;
	const double thr0 = 0.1;

	/* init (c99/d_aux_ip_hard_lib4.c:43-149) */
	if(warm_start==0) for(n=0; n<=N; n++) for(i=0; i<nux_(P,n); i++) ux[n][i] = 0.0;
	for(n=0; n<=N && !newton; n++)
		{
		int nb = P->nb[n];
		for(j=0; j<P->nbx[n]; j++)
			{
			int id = P->idxb[n][j];
			t[n][j]    = -P->d[n][j]    + ux[n][id];
			t[n][nb+j] =  P->d[n][nb+j] - ux[n][id];
			if(t[n][j]<thr0)
				{
				if(t[n][nb+j]<thr0)
					{
					ux[n][id] = (-P->d[n][nb+j] + P->d[n][j])*0.5;
					t[n][j] = thr0; t[n][nb+j] = thr0;
					}
				else { t[n][j] = thr0; ux[n][id] = P->d[n][j] + thr0; }
				}
			else if(t[n][nb+j]<thr0) { t[n][nb+j] = thr0; ux[n][id] = P->d[n][nb+j] - thr0; }
			lam[n][j] = mu0/t[n][j];
			lam[n][nb+j] = mu0/t[n][nb+j];
			}
		}
	for(n=0; n<N && !newton; n++) for(i=0; i<P->nx[n+1]; i++) pi[n][i] = 0.0;
	/* general constraints, from the ux the bounds have just moved; no projection (c99/d_aux_ip_hard_lib4.c:121-147) */
	for(n=0; n<=N && !newton; n++)
		{
		int nb = P->nb[n];
		for(j=P->nbx[n]; j<nb; j++)
			{
			const double v = cval(P, n, ux[n], j);
			t[n][j] = fmax(thr0, v - P->d[n][j]);
			t[n][nb+j] = fmax(thr0, -v + P->d[n][nb+j]);
			lam[n][j] = mu0/t[n][j]; lam[n][nb+j] = mu0/t[n][nb+j];
			}
		}

	mu = mu0; *kk = 0; alpha = 1.0;
	double mu_tol_low = mu_tol<1e-5 ? 1e-5 : mu_tol;

	/* ---------------- phase 1: no residuals (d_ip2_res_hard.c:503-718) ---------------- */
	while(!newton && *kk<k_max && mu>mu_tol_low && alpha>=alpha_min)
		{
		for(n=0; n<=N; n++)
			{
			int nb = P->nb[n];
			for(j=0; j<nb; j++)
				{
				w->tinv[n][j] = 1.0/t[n][j]; w->tinv[n][nb+j] = 1.0/t[n][nb+j];
				w->lamt[n][j] = lam[n][j]*w->tinv[n][j]; w->lamt[n][nb+j] = lam[n][nb+j]*w->tinv[n][nb+j];
				w->dlam[n][j] = w->tinv[n][j]*0.0; w->dlam[n][nb+j] = w->tinv[n][nb+j]*0.0;
				w->Qx[n][j] = w->lamt[n][j] + w->lamt[n][nb+j];
				w->qx[n][j] = lam[n][nb+j] - w->lamt[n][nb+j]*P->d[n][nb+j] + w->dlam[n][nb+j]
				            - lam[n][j] - w->lamt[n][j]*P->d[n][j] - w->dlam[n][j];
				}
			}
		orc_ric_sv(P, NULL, w->rq, w->Qx, w->qx, w->dux, 1, w->dpi, w->Pb);
		for(int pass=0; pass<2; pass++)
			{
			alpha = 1.0;
			for(n=0; n<=N; n++)
				{
				int nb = P->nb[n];
				for(j=0; j<nb; j++)
					{
					const double dv = cval(P, n, w->dux[n], j);
					w->dt[n][j]    =  dv - P->d[n][j]    - t[n][j];
					w->dt[n][nb+j] = -dv + P->d[n][nb+j] - t[n][nb+j];
					w->dlam[n][j]    -= w->lamt[n][j]*w->dt[n][j] + lam[n][j];
					w->dlam[n][nb+j] -= w->lamt[n][nb+j]*w->dt[n][nb+j] + lam[n][nb+j];
					if(-alpha*w->dlam[n][j]>lam[n][j]) alpha = -lam[n][j]/w->dlam[n][j];
					if(-alpha*w->dlam[n][nb+j]>lam[n][nb+j]) alpha = -lam[n][nb+j]/w->dlam[n][nb+j];
					if(-alpha*w->dt[n][j]>t[n][j]) alpha = -t[n][j]/w->dt[n][j];
					if(-alpha*w->dt[n][nb+j]>t[n][nb+j]) alpha = -t[n][nb+j]/w->dt[n][nb+j];
					}
				}
			if(pass==0)
				{
				stat[5*(*kk)] = sigma; stat[5*(*kk)+1] = alpha;
				alpha *= 0.995;
				mu_aff = 0.0;
				for(n=0; n<=N; n++)
					{
					int nb = P->nb[n];
					for(j=0; j<nb; j++)
						mu_aff += (lam[n][j] + alpha*w->dlam[n][j])*(t[n][j] + alpha*w->dt[n][j])
						        + (lam[n][nb+j] + alpha*w->dlam[n][nb+j])*(t[n][nb+j] + alpha*w->dt[n][nb+j]);
					}
				mu_aff *= mu_scal;
				stat[5*(*kk)+2] = mu_aff;
				sigma = mu_aff/mu; sigma = sigma*sigma*sigma;
				double sm = sigma*mu;
				for(n=0; n<=N; n++)
					{
					int nb = P->nb[n];
					for(j=0; j<nb; j++)
						{
						w->dlam[n][j]    = w->tinv[n][j]*(sm - w->dlam[n][j]*w->dt[n][j]);
						w->dlam[n][nb+j] = w->tinv[n][nb+j]*(sm - w->dlam[n][nb+j]*w->dt[n][nb+j]);
						w->qx[n][j] += w->dlam[n][nb+j] - w->dlam[n][j];
						}
					}
				orc_ric_trs(P, w->b, w->rq, w->qx, w->dux, 1, w->dpi, 0, w->Pb);
				}
			}
		stat[5*(*kk)] = sigma; stat[5*(*kk)+3] = alpha;
		alpha *= 0.995;
		mu = 0.0;
		for(n=0; n<=N; n++)
			{
			int nux = nux_(P,n), nb = P->nb[n];
			for(i=0; i<nux; i++) ux[n][i] += alpha*(w->dux[n][i] - ux[n][i]);
			if(n<N) for(i=0; i<P->nx[n+1]; i++) pi[n][i] += alpha*(w->dpi[n][i] - pi[n][i]);
			for(j=0; j<nb; j++)
				{
				lam[n][j] += alpha*w->dlam[n][j]; lam[n][nb+j] += alpha*w->dlam[n][nb+j];
				t[n][j] += alpha*w->dt[n][j]; t[n][nb+j] += alpha*w->dt[n][nb+j];
				mu += lam[n][j]*t[n][j] + lam[n][nb+j]*t[n][nb+j];
				}
			}
		mu *= mu_scal;
		stat[5*(*kk)+4] = mu;
		(*kk)++;
		}

	/* ---------------- phase 2: with residuals (d_ip2_res_hard.c:756-1273) ---------------- */
	ipm_residuals(P, w, &mu);
	while(*kk<k_max && (newton || (mu>mu_tol && alpha>=alpha_min)))
		{
		for(n=0; n<=N; n++)
			{
			int nb = P->nb[n];
			for(j=0; j<nb; j++)
				{
				w->tinv[n][j] = 1.0/t[n][j]; w->tinv[n][nb+j] = 1.0/t[n][nb+j];
				w->Qx[n][j] = w->tinv[n][j]*lam[n][j] + w->tinv[n][nb+j]*lam[n][nb+j];
				w->qx[n][j] = w->tinv[n][j]*(w->res_m[n][j] - lam[n][j]*w->res_d[n][j])
				            - w->tinv[n][nb+j]*(w->res_m[n][nb+j] + lam[n][nb+j]*w->res_d[n][nb+j]);
				}
			}
		/* the single-Newton-step routine of the reference solves its predictor system with the ORIGINAL b and q (update_b = 0,
		 * update_q = 1 with q, mpc_solvers/d_ip2_res_hard.c:1736) and only the corrector with the residuals (:1788, reusing the
		 * predictor's Pb); restated as it is */
		if(newton) orc_ric_sv(P, NULL, w->rq, w->Qx, w->qx, w->dux, 1, w->dpi, w->Pb);
		else orc_ric_sv(P, w->res_b, w->res_q, w->Qx, w->qx, w->dux, 1, w->dpi, w->Pb);
		for(int pass=0; pass<2; pass++)
			{
			alpha = 1.0;
			for(n=0; n<=N; n++)
				{
				int nb = P->nb[n];
				for(j=0; j<nb; j++)
					{
					const double dv = cval(P, n, w->dux[n], j);
					w->dt[n][j]    =  dv - w->res_d[n][j];
					w->dt[n][nb+j] = -dv + w->res_d[n][nb+j];
					w->dlam[n][j]    = -w->tinv[n][j]*(lam[n][j]*w->dt[n][j] + w->res_m[n][j]);
					w->dlam[n][nb+j] = -w->tinv[n][nb+j]*(lam[n][nb+j]*w->dt[n][nb+j] + w->res_m[n][nb+j]);
					if(-alpha*w->dlam[n][j]>lam[n][j]) alpha = -lam[n][j]/w->dlam[n][j];
					if(-alpha*w->dlam[n][nb+j]>lam[n][nb+j]) alpha = -lam[n][nb+j]/w->dlam[n][nb+j];
					if(-alpha*w->dt[n][j]>t[n][j]) alpha = -t[n][j]/w->dt[n][j];
					if(-alpha*w->dt[n][nb+j]>t[n][nb+j]) alpha = -t[n][nb+j]/w->dt[n][nb+j];
					}
				}
			if(pass==0)
				{
				stat[5*(*kk)] = sigma; stat[5*(*kk)+1] = alpha;
				alpha *= 0.995;
				mu_aff = 0.0;
				for(n=0; n<=N; n++)
					{
					int nb = P->nb[n];
					for(j=0; j<nb; j++)
						mu_aff += (lam[n][j] + alpha*w->dlam[n][j])*(t[n][j] + alpha*w->dt[n][j])
						        + (lam[n][nb+j] + alpha*w->dlam[n][nb+j])*(t[n][nb+j] + alpha*w->dt[n][nb+j]);
					}
				mu_aff *= mu_scal;
				stat[5*(*kk)+2] = mu_aff;
				if(!newton) { sigma = mu_aff/mu; sigma = sigma*sigma*sigma; }
				double sm = newton ? mu0 : sigma*mu;
				for(n=0; n<=N; n++)
					{
					int nb = P->nb[n];
					for(j=0; j<nb; j++)
						{
						w->res_m[n][j]    += w->dt[n][j]*w->dlam[n][j] - sm;
						w->res_m[n][nb+j] += w->dt[n][nb+j]*w->dlam[n][nb+j] - sm;
						w->qx[n][j] = w->tinv[n][j]*(w->res_m[n][j] - lam[n][j]*w->res_d[n][j])
						            - w->tinv[n][nb+j]*(w->res_m[n][nb+j] + lam[n][nb+j]*w->res_d[n][nb+j]);
						}
					}
				orc_ric_trs(P, w->res_b, w->res_q, w->qx, w->dux, 1, w->dpi, 0, w->Pb);
				}
			}
		stat[5*(*kk)] = sigma; stat[5*(*kk)+3] = alpha;
		alpha *= 0.995;
		/* backup of the iterate the factor belongs to (d_backup_update_var_res_mpc_hard_tv, c99/d_aux_ip_hard_lib4.c:1382) */
		for(n=0; n<=N; n++)
			{
			int nux = nux_(P,n), nb = P->nb[n];
			for(i=0; i<nux; i++) P->k_ux[n][i] = ux[n][i];
			if(n<N) for(i=0; i<P->nx[n+1]; i++) P->k_pi[n][i] = pi[n][i];
			for(j=0; j<2*nb; j++) { P->k_lam[n][j] = lam[n][j]; P->k_t[n][j] = t[n][j]; P->k_tinv[n][j] = w->tinv[n][j]; }
			}
		P->k_valid = 1;
		for(n=0; n<=N; n++)
			{
			int nux = nux_(P,n), nb = P->nb[n];
			for(i=0; i<nux; i++) ux[n][i] += alpha*w->dux[n][i];
			if(n<N) for(i=0; i<P->nx[n+1]; i++) pi[n][i] += alpha*w->dpi[n][i];
			for(j=0; j<2*nb; j++) { lam[n][j] += alpha*w->dlam[n][j]; t[n][j] += alpha*w->dt[n][j]; }
			}
		ipm_residuals(P, w, &mu);
		stat[5*(*kk)+4] = mu;
		(*kk)++;
		}

	if(newton) status = (*kk>=k_max) ? 1 : (alpha<alpha_min ? 2 : -1);
	else if(mu<=mu_tol) status = 0;
	else if(*kk>=k_max) status = 1;
	else if(alpha<alpha_min) status = 2;
	else status = -1;

done:
	vecs_free(w->dux, N+1); vecs_free(w->dpi, N+1); vecs_free(w->dlam, N+1); vecs_free(w->dt, N+1);
	vecs_free(w->tinv, N+1); vecs_free(w->lamt, N+1); vecs_free(w->Qx, N+1); vecs_free(w->qx, N+1);
	vecs_free(w->Pb, N+1); vecs_free(w->b, N+1); vecs_free(w->rq, N+1);
	vecs_free(w->res_q, N+1); vecs_free(w->res_b, N+1); vecs_free(w->res_d, N+1); vecs_free(w->res_m, N+1);
	free(lnux); free(lnx1); free(l2nb); free(lnb);
	return status;
	}

/* solve the KKT system of the IPM's last iteration again for a new right-hand side (new b_n, [r;q]_n, same matrices and
 * bounds): d_kkt_solve_new_rhs_res_mpc_hard_tv, mpc_solvers/d_ip2_res_hard.c:1922.  Start from the backed-up iterate (:2138-2173),
 * residuals there with the new vectors (:2192), qx from the stored t_inv (d_update_gradient_res, :2216), one solve with the
 * stored factor and compute_Pb = 1 (:2225), dt / dlam (d_compute_dt_dlam_res, :2236), full step alpha = 1 (:2239).
 * bnew[n] (nx_{n+1}), rqnew[n] (nux_n).  Needs P as left by orc_ip2_res_mpc_hard with at least one phase-2 iteration. */
int orc_kkt_solve_new_rhs(orc_prob *P, double *const *bnew, double *const *rqnew, double **ux, double **pi, double **lam, double **t)
	{
	int N = P->N, n, i, j;
	if(!P->k_valid) return -1;
	int *lnux = malloc((N+1)*sizeof(int)), *lnx1 = malloc((N+1)*sizeof(int)), *l2nb = malloc((N+1)*sizeof(int)), *lnb = malloc((N+1)*sizeof(int));
	for(n=0; n<=N; n++) { lnux[n] = nux_(P,n)+1; lnx1[n] = n<N ? P->nx[n+1] : 0; l2nb[n] = 2*P->nb[n]; lnb[n] = P->nb[n]; }
	orc_ipm_ws W, *w = &W;
	memset(w, 0, sizeof(W));
	w->ux = ux; w->pi = pi; w->lam = lam; w->t = t;
	w->dux = vecs(N+1, lnux); w->dpi = vecs(N+1, lnx1); w->qx = vecs(N+1, lnb); w->Pb = vecs(N+1, lnx1);
	w->b = vecs(N+1, lnx1); w->rq = vecs(N+1, lnux);
	w->res_q = vecs(N+1, lnux); w->res_b = vecs(N+1, lnx1); w->res_d = vecs(N+1, l2nb); w->res_m = vecs(N+1, l2nb);
	for(n=0; n<=N; n++)
		{
		int nux = nux_(P,n), nb = P->nb[n];
		for(i=0; i<nux; i++) { w->rq[n][i] = rqnew[n][i]; ux[n][i] = P->k_ux[n][i]; }
		if(n<N) for(i=0; i<P->nx[n+1]; i++) { w->b[n][i] = bnew[n][i]; pi[n][i] = P->k_pi[n][i]; }
		for(j=0; j<2*nb; j++) { lam[n][j] = P->k_lam[n][j]; t[n][j] = P->k_t[n][j]; }
		}
	double mu = 0.0;
	ipm_residuals(P, w, &mu);
	for(n=0; n<=N; n++)
		{
		int nb = P->nb[n]; const double *ti = P->k_tinv[n];
		for(j=0; j<nb; j++)
			w->qx[n][j] = ti[j]*(w->res_m[n][j] - lam[n][j]*w->res_d[n][j]) - ti[nb+j]*(w->res_m[n][nb+j] + lam[n][nb+j]*w->res_d[n][nb+j]);
		}
	orc_ric_trs(P, w->res_b, w->res_q, w->qx, w->dux, 1, w->dpi, 1, w->Pb);
	for(n=0; n<=N; n++)
		{
		int nux = nux_(P,n), nb = P->nb[n]; const double *ti = P->k_tinv[n];
		for(j=0; j<nb; j++)
			{
			const double dv = cval(P, n, w->dux[n], j);
			double dtl =  dv - w->res_d[n][j], dtu = -dv + w->res_d[n][nb+j];
			double dll = -ti[j]*(lam[n][j]*dtl + w->res_m[n][j]), dlu = -ti[nb+j]*(lam[n][nb+j]*dtu + w->res_m[n][nb+j]);
			lam[n][j] += 1.0*dll; lam[n][nb+j] += 1.0*dlu; t[n][j] += 1.0*dtl; t[n][nb+j] += 1.0*dtu;
			}
		for(i=0; i<nux; i++) ux[n][i] += 1.0*w->dux[n][i];
		if(n<N) for(i=0; i<P->nx[n+1]; i++) pi[n][i] += 1.0*w->dpi[n][i];
		}
	vecs_free(w->dux, N+1); vecs_free(w->dpi, N+1); vecs_free(w->qx, N+1); vecs_free(w->Pb, N+1); vecs_free(w->b, N+1); vecs_free(w->rq, N+1);
	vecs_free(w->res_q, N+1); vecs_free(w->res_b, N+1); vecs_free(w->res_d, N+1); vecs_free(w->res_m, N+1);
	free(lnux); free(lnx1); free(l2nb); free(lnb);
	return 0;
	}

/* exit residual norms (mpc_solvers/d_res_ip_hard.c:38 + interfaces/c/fortran_order_interface.c:612-652) */
void orc_exit_residuals(const orc_prob *P, double *const *ux, double *const *pi, double *const *lam, double *const *t,
		double *inf_norm_res)
	{
	int N = P->N, n, i, j;
	double nq = 0.0, nb_ = 0.0, nd = 0.0, mu = 0.0; int nb_tot = 0;
	for(n=0; n<=N; n++)
		{
		int nx = P->nx[n], nu = P->nu[n], nux = nx+nu, nz = nux+1, nb = P->nb[n];
		const double *H = P->RSQrq[n];
		double *rq = calloc(nux+1, sizeof(double));
		for(i=0; i<nux; i++) rq[i] = H[nux+nz*i];
		if(n>0) for(i=0; i<nx; i++) rq[nu+i] -= pi[n-1][i];
		nb_tot += nb;
		for(j=0; j<nb; j++)
			{
			const double v = cval(P, n, ux[n], j);
			cscatter(P, n, rq, j, -lam[n][j] + lam[n][nb+j]);
			mu += lam[n][j]*t[n][j] + lam[n][nb+j]*t[n][nb+j];
			nd = fmax(nd, fabs(v - P->d[n][j] - t[n][j]));
			nd = fmax(nd, fabs(-v + P->d[n][nb+j] - t[n][nb+j]));
			}
		for(i=0; i<nux; i++)
			{
			double s = 0.0;
			for(j=0; j<nux; j++) s += (i>=j ? H[i+nz*j] : H[j+nz*i])*ux[n][j];
			rq[i] += s;
			}
		if(n<N)
			{
			int nx1 = P->nx[n+1], nu1 = P->nu[n+1];
			const double *M = P->BAbt[n];
			for(j=0; j<nx1; j++)
				{
				double s = M[nux+nz*j] - ux[n+1][nu1+j];
				for(i=0; i<nux; i++) s += M[i+nz*j]*ux[n][i];
				nb_ = fmax(nb_, fabs(s));
				}
			for(i=0; i<nux; i++)
				{
				double s = 0.0;
				for(j=0; j<nx1; j++) s += M[i+nz*j]*pi[n][j];
				rq[i] += s;
				}
			}
		for(i=0; i<nux; i++) nq = fmax(nq, fabs(rq[i]));
		free(rq);
		}
	if(nb_tot) mu /= 2.0*nb_tot;
	inf_norm_res[0] = nq; inf_norm_res[1] = nb_; inf_norm_res[2] = nd; inf_norm_res[3] = mu;
	}

/* ------------------------------------------------------------------------------------------- */
/* partial condensing (lqcp_solvers/d_part_cond.c): the horizon is cut into N2 blocks, the states inside a block are    */
/* eliminated, and every block becomes one stage of a shorter problem with inputs [u_{T-1} .. u_1 u_0] (newest first)     */
/* and the state of the block's first stage.  State bounds inside a block become general constraints.                    */
/* ------------------------------------------------------------------------------------------- */
/* block sizes: the first R1 = N - N2*(N/N2) blocks hold N/N2+1 stages, the others N/N2 (d_part_cond.c:694-735) */
static int pc_block_len(int N, int N2, int k) { int N1 = N/N2, R1 = N - N2*N1; return k<R1 ? N1+1 : N1; }

/* sizes of the condensed problem and the positions of its bounds (d_part_cond.c:694-735 for the counts, :637-679 for idxb2).
 * nx2, nu2, nb2, ng2: [N2+1]; idxb2[k] must hold nb2[k] ints.  Returns 0, or -1 when a stage before N has general constraints
 * (the reference stops there, :962-968). */
int orc_part_cond_sizes(int N, const int *nx, const int *nu, const int *nb, int *const *idxb, const int *ng, int N2,
		int *nx2, int *nu2, int *nb2, int *ng2, int **idxb2)
	{
	int k, j, l, n0 = 0;
	if(ng) for(k=0; k<N; k++) if(ng[k]>0) return -1;
	for(k=0; k<N2; k++)
		{
		const int T = pc_block_len(N, N2, k);
		nx2[k] = nx[n0]; nu2[k] = 0; nb2[k] = 0; ng2[k] = 0;
		for(j=0; j<T; j++) nu2[k] += nu[n0+j];
		/* constraints in the order the reference emits them: stages T-1 .. 1 (inputs stay bounds, states become rows of
		 * [D C]), then every bound of the block's first stage */
		int nu_tmp = 0, ib = 0;
		for(j=T-1; j>=1; j--)
			{
			nu_tmp += nu[n0+j];
			for(l=0; l<nb[n0+j]; l++)
				{
				if(idxb[n0+j][l]<nu[n0+j]) { if(idxb2) idxb2[k][ib] = nu_tmp - nu[n0+j] + idxb[n0+j][l]; ib++; }
				else ng2[k]++;
				}
			}
		nu_tmp += nu[n0];
		for(l=0; l<nb[n0]; l++) { if(idxb2) idxb2[k][ib] = nu_tmp - nu[n0] + idxb[n0][l]; ib++; }
		nb2[k] = ib;
		n0 += T;
		}
	nx2[N2] = nx[N]; nu2[N2] = 0; nb2[N2] = nb[N]; ng2[N2] = ng ? ng[N] : 0;
	if(idxb2) for(l=0; l<nb[N]; l++) idxb2[N2][l] = idxb[N][l];
	return 0;
	}

/* one block: stages n0 .. n0+T-1 of P -> stage k of P2 */
static void pc_cond_block(const orc_prob *P, int n0, int T, orc_prob *P2, int k)
	{
	const int nx0 = P->nx[n0];
	int j, i, c, m, l;
	/* ---- Gamma_j = [ B_j' ; Gamma_{j-1} A_j' ] (+ b_j on the last row): the states x_{j+1} as an affine function of
	 *      [u_j .. u_0, x_0, 1]   (d_cond_BAbt, d_part_cond.c:214-309) ---- */
	double **G = malloc(T*sizeof(double*));
	int *gr = malloc(T*sizeof(int));                          /* rows of Gamma_j = inputs so far + nx0 + 1 */
	int nuc = 0;
	for(j=0; j<T; j++)
		{
		const int s = n0+j, nu = P->nu[s], nx = P->nx[s], nx1 = P->nx[s+1], nz = nu+nx+1;
		nuc += nu; gr[j] = nuc + nx0 + 1;
		G[j] = calloc((size_t)gr[j]*(nx1+1), sizeof(double));
		const double *M = P->BAbt[s];
		if(j==0) { for(c=0; c<nx1; c++) for(i=0; i<nz; i++) G[0][i+gr[0]*c] = M[i+nz*c]; continue; }
		for(c=0; c<nx1; c++)
			{
			for(i=0; i<nu; i++) G[j][i+gr[j]*c] = M[i+nz*c];
			for(i=0; i<gr[j-1]; i++)
				{
				double a = 0.0;
				for(m=0; m<nx; m++) a += G[j-1][i+gr[j-1]*m]*M[nu+m+nz*c];
				G[j][nu+i+gr[j]*c] = a;
				}
			G[j][gr[j]-1+gr[j]*c] += M[nu+nx+nz*c];
			}
		}
	{
	const int nz2 = gr[T-1], nx1 = P->nx[n0+T];
	for(c=0; c<nx1; c++) for(i=0; i<nz2; i++) P2->BAbt[k][i+nz2*c] = G[T-1][i+nz2*c];
	}
	/* ---- Hessian and gradient of the block (d_cond_RSQrq, d_part_cond.c:312-574): backward recursion
	 *      pL_s = RSQrq_s + ([B A b]'_s Lx)([B A b]'_s Lx)' with Lx the Cholesky factor of the state block of pL_{s+1} (gradient
	 *      row carried), no elimination of u; the u-columns of pL_s go to the block's Hessian, the coupling with everything
	 *      earlier through Gamma_{s-1} ---- */
	{
	const int nu2 = P2->nu[k], nux2 = nu2+nx0, nz2 = nux2+1;
	double *H2 = P2->RSQrq[k];
	const int nzM = P->nzM;
	double *pL = calloc((size_t)nzM*nzM, sizeof(double)), *Lx = calloc((size_t)nzM*nzM, sizeof(double));
	double *W = calloc((size_t)nzM*nzM, sizeof(double)), *dl = calloc(nzM, sizeof(double));
	int off = 0;                                              /* nu3: inputs of the stages behind s */
	for(int s=T-1; s>=0; s--)
		{
		const int st = n0+s, nu = P->nu[st], nx = P->nx[st], nux = nu+nx, nz = nux+1;
		const double *H = P->RSQrq[st];
		if(s==T-1)
			{ for(c=0; c<nux; c++) for(i=c; i<nz; i++) pL[i+nz*c] = H[i+nz*c]; }
		else
			{
			const int nus = P->nu[st+1], nxs = P->nx[st+1], nzs = nus+nxs+1;    /* pL still holds stage st+1, ld nzs */
			for(c=0; c<nxs; c++) for(i=c; i<=nxs; i++) Lx[i+(nxs+1)*c] = pL[nus+i+nzs*(nus+c)];
			chol_mn(nxs+1, nxs, Lx, nxs+1, dl);
			const double *M = P->BAbt[st];
			for(i=0; i<nz; i++)
				for(c=0; c<nxs; c++)
					{
					double a = 0.0;
					for(m=c; m<nxs; m++) a += M[i+nz*m]*Lx[m+(nxs+1)*c];
					W[i+nz*c] = a;
					}
			for(c=0; c<nxs; c++) W[nux+nz*c] += Lx[nxs+(nxs+1)*c];
			for(c=0; c<nux; c++)
				for(i=c; i<nz; i++)
					{
					double a = 0.0;
					for(m=0; m<nxs; m++) a += W[i+nz*m]*W[c+nz*m];
					pL[i+nz*c] = H[i+nz*c] + a;
					}
			}
		if(s==0)
			{ for(c=0; c<nux; c++) for(i=c; i<nz; i++) H2[off+i+nz2*(off+c)] = pL[i+nz*c]; break; }
		/* D */
		for(c=0; c<nu; c++) for(i=c; i<nu; i++) H2[off+i+nz2*(off+c)] = pL[i+nz*c];
		/* M : Gamma_{s-1} times the state rows of the u-columns; its last row is the b-part of the gradient */
		for(c=0; c<nu; c++)
			for(i=0; i<gr[s-1]; i++)
				{
				double a = 0.0;
				for(m=0; m<nx; m++) a += G[s-1][i+gr[s-1]*m]*pL[nu+m+nz*c];
				H2[off+nu+i+nz2*(off+c)] = a;
				}
		/* m */
		for(c=0; c<nu; c++) H2[nux2+nz2*(off+c)] += pL[nux+nz*c];
		off += nu;
		}
	free(pL); free(Lx); free(W); free(dl);
	}
	/* ---- constraints (d_cond_DCtd, d_part_cond.c:579-689) ---- */
	{
	const int nux2 = P2->nu[k]+nx0, nbx2 = P2->nbx[k], nt2 = P2->nb[k];
	int nu_tmp = 0, ib = 0, ig = 0;
	for(int s=T-1; s>=1; s--)
		{
		const int st = n0+s, nu = P->nu[st];
		const int brow = gr[s-1]-1;                             /* the b-row of Gamma_{s-1} */
		nu_tmp += nu;
		for(l=0; l<P->nbx[st]; l++)
			{
			const int id = P->idxb[st][l];
			if(id<nu)
				{ P2->d[k][ib] = P->d[st][l]; P2->d[k][nt2+ib] = P->d[st][P->nb[st]+l]; ib++; }
			else
				{
				const double *g = G[s-1] + (size_t)gr[s-1]*(id-nu);
				P2->d[k][nbx2+ig] = P->d[st][l] - g[brow]; P2->d[k][nt2+nbx2+ig] = P->d[st][P->nb[st]+l] - g[brow];
				for(i=0; i<brow; i++) P2->DCt[k][nu_tmp+i+(size_t)nux2*ig] = g[i];
				ig++;
				}
			}
		}
	for(l=0; l<P->nbx[n0]; l++) { P2->d[k][ib] = P->d[n0][l]; P2->d[k][nt2+ib] = P->d[n0][P->nb[n0]+l]; ib++; }
	}
	for(j=0; j<T; j++) free(G[j]);
	free(G); free(gr);
	}

/* the condensed problem of P with N2 blocks (d_part_cond, d_part_cond.c:926-1066); NULL when P is not condensable */
orc_prob *orc_part_cond(const orc_prob *P, int N2)
	{
	const int N = P->N;
	int k, n0 = 0;
	if(N2<1 || N2>=N || P->dad) return NULL;
	int *nx2 = calloc(N2+1, sizeof(int)), *nu2 = calloc(N2+1, sizeof(int)), *nb2 = calloc(N2+1, sizeof(int)), *ng2 = calloc(N2+1, sizeof(int));
	int **idxb2 = calloc(N2+1, sizeof(int*));
	int nbt = 1;
	for(k=0; k<=N; k++) nbt += P->nbx[k];
	for(k=0; k<=N2; k++) idxb2[k] = calloc(nbt, sizeof(int));
	orc_prob *P2 = NULL;
	if(orc_part_cond_sizes(N, P->nx, P->nu, P->nbx, P->idxb, P->ng, N2, nx2, nu2, nb2, ng2, idxb2)==0)
		{
		P2 = orc_prob_create_gen(N2, nx2, nu2, nb2, idxb2, NULL, ng2);
		for(k=0; k<N2; k++) { const int T = pc_block_len(N, N2, k); pc_cond_block(P, n0, T, P2, k); n0 += T; }
		/* last stage: the same data (d_part_cond.c:1058-1062) */
		const int nx = P->nx[N], nz = nx+1, nt = P->nb[N];
		memcpy(P2->RSQrq[N2], P->RSQrq[N], sizeof(double)*nz*(nx+1));
		memcpy(P2->d[N2], P->d[N], sizeof(double)*2*nt);
		memcpy(P2->DCt[N2], P->DCt[N], sizeof(double)*(size_t)nx*P->ng[N]);
		}
	for(k=0; k<=N2; k++) free(idxb2[k]);
	free(idxb2); free(nx2); free(nu2); free(nb2); free(ng2);
	return P2;
	}

/* solution of the full problem from the condensed one (d_part_expand_solution, d_part_cond.c:1103-1306): inputs are copied,
 * the states inside a block are simulated, lam / t are copied back by position (a block's bounds on inputs first, then the
 * state bounds that became general constraints), pi inside a block by the backward recursion of the stationarity condition */
void orc_part_expand(const orc_prob *P, const orc_prob *P2, double *const *ux2, double *const *pi2, double *const *lam2, double *const *t2,
		double **ux, double **pi, double **lam, double **t)
	{
	const int N = P->N, N2 = P2->N;
	int k, j, l, i, n0 = 0;
	double *w = calloc(P->nzM+1, sizeof(double));
	for(k=0; k<N2; k++)
		{
		const int T = pc_block_len(N, N2, k);
		const int nt2 = P2->nb[k], nbx2 = P2->nbx[k];
		int nu_tmp = 0, ib = 0, ig = 0;
		for(j=T-1; j>=1; j--)
			{
			const int st = n0+j, nu = P->nu[st], nt = P->nb[st];
			for(l=0; l<nu; l++) ux[st][l] = ux2[k][nu_tmp+l];
			nu_tmp += nu;
			int nbb = 0;
			for(l=0; l<P->nbx[st]; l++) if(P->idxb[st][l]<nu) nbb++;
			for(l=0; l<nbb; l++, ib++)
				{ lam[st][l] = lam2[k][ib]; lam[st][nt+l] = lam2[k][nt2+ib]; t[st][l] = t2[k][ib]; t[st][nt+l] = t2[k][nt2+ib]; }
			for(l=nbb; l<P->nbx[st]; l++, ig++)
				{ lam[st][l] = lam2[k][nbx2+ig]; lam[st][nt+l] = lam2[k][nt2+nbx2+ig]; t[st][l] = t2[k][nbx2+ig]; t[st][nt+l] = t2[k][nt2+nbx2+ig]; }
			}
		{
		const int nt = P->nb[n0];
		for(l=0; l<P->nu[n0]+P->nx[n0]; l++) ux[n0][l] = ux2[k][nu_tmp+l];
		for(l=0; l<P->nbx[n0]; l++, ib++)
			{ lam[n0][l] = lam2[k][ib]; lam[n0][nt+l] = lam2[k][nt2+ib]; t[n0][l] = t2[k][ib]; t[n0][nt+l] = t2[k][nt2+ib]; }
		}
		n0 += T;
		}
	for(l=0; l<P->nx[N]; l++) ux[N][l] = ux2[N2][l];
	for(l=0; l<2*P->nb[N]; l++) { lam[N][l] = lam2[N2][l]; t[N][l] = t2[N2][l]; }
	/* states inside the blocks */
	n0 = 0;
	for(k=0; k<N2; k++)
		{
		const int T = pc_block_len(N, N2, k);
		for(j=0; j<T-1; j++)
			{
			const int st = n0+j, nux = nux_(P, st), nz = nux+1, nx1 = P->nx[st+1], nu1 = P->nu[st+1];
			for(l=0; l<nx1; l++)
				{
				double a = P->BAbt[st][nux+nz*l];
				for(i=0; i<nux; i++) a += P->BAbt[st][i+nz*l]*ux[st][i];
				ux[st+1][nu1+l] = a;
				}
			}
		n0 += T;
		}
	/* multipliers of the dynamics */
	n0 = 0;
	for(k=0; k<N2; k++)
		{
		const int T = pc_block_len(N, N2, k);
		for(l=0; l<P->nx[n0+T]; l++) pi[n0+T-1][l] = pi2[k][l];
		for(j=T-1; j>=1; j--)
			{
			const int st = n0+j, nu = P->nu[st], nux = nux_(P, st), nz = nux+1, nx1 = P->nx[st+1], nt = P->nb[st];
			const double *H = P->RSQrq[st];
			for(l=0; l<nux; l++) w[l] = H[nux+nz*l];
			for(l=0; l<nt; l++) cscatter(P, st, w, l, -lam[st][l] + lam[st][nt+l]);
			for(l=0; l<nux; l++)
				{
				double a = 0.0;
				for(i=0; i<nux; i++) a += (i>=l ? H[i+nz*l] : H[l+nz*i])*ux[st][i];
				w[l] += a;
				}
			for(l=0; l<nux; l++)
				{
				double a = 0.0;
				for(i=0; i<nx1; i++) a += P->BAbt[st][l+nz*i]*pi[st][i];
				w[l] += a;
				}
			for(l=0; l<P->nx[st]; l++) pi[st-1][l] = w[nu+l];
			}
		n0 += T;
		}
	free(w);
	}

/* ------------------------------------------------------------------------------------------- */
/* drivers with the reference's own high-level signature (column-major stage-wise arrays)       */
/* ------------------------------------------------------------------------------------------- */
static double **alloc_ux(const orc_prob *P) { int N=P->N; double **v = malloc((N+1)*sizeof(double*)); for(int n=0;n<=N;n++) v[n]=calloc(nux_(P,n)+1,sizeof(double)); return v; }

/* same argument list as fortran_order_d_ip_ocp_hard_tv (include/c_interface.h:65 of the reference) */
int orc_fortran_order_d_ip_ocp_hard_tv(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu_N, int *nb,
		int **hidxb, int *ng, int N2, int warm_start, double **A, double **B, double **b, double **Q, double **S,
		double **R, double **q, double **r, double **lb, double **ub, double **C, double **D, double **lg, double **ug,
		double **x, double **u, double **pi, double **lam, double *inf_norm_res, void *work0, double *stat)
	{
	(void)work0;
	int n, i, j, l, any_g = 0;
	if(ng) for(n=0; n<=N; n++) if(ng[n]>0) any_g = 1;
	orc_prob *P = orc_prob_create_gen(N, nx, nu_N, nb, hidxb, NULL, any_g ? ng : NULL);
	orc_prob_set(P, A, B, b, Q, S, R, q, r, lb, ub);
	if(any_g) orc_prob_set_general(P, C, D, lg, ug);
	/* mu0 estimate: signed max over cost entries (fortran_order_interface.c:318-331) */
	if(mu0<=0)
		{
		for(n=0; n<N; n++)
			{
			int nu = P->nu[n];
			for(j=0; j<nu; j++) for(l=0; l<nu; l++) mu0 = fmax(mu0, R[n][j*nu+l]);
			for(j=0; j<nx[n]*nu; j++) mu0 = fmax(mu0, S[n][j]);
			for(j=0; j<nx[n]; j++) for(l=0; l<nx[n]; l++) mu0 = fmax(mu0, Q[n][j*nx[n]+l]);
			for(j=0; j<nu; j++) mu0 = fmax(mu0, r[n][j]);
			for(j=0; j<nx[n]; j++) mu0 = fmax(mu0, q[n][j]);
			}
		n = N;
		for(j=0; j<nx[n]; j++) for(l=0; l<nx[n]; l++) mu0 = fmax(mu0, Q[n][j*nx[n]+l]);
		for(j=0; j<nx[n]; j++) mu0 = fmax(mu0, q[n][j]);
		}
	double **hux = alloc_ux(P);
	double **hpi = malloc((N+1)*sizeof(double*)), **hlam = malloc((N+1)*sizeof(double*)), **ht = malloc((N+1)*sizeof(double*));
	for(n=0; n<=N; n++) { hpi[n] = calloc(P->nxM+1, sizeof(double)); hlam[n] = calloc(2*P->nb[n]+1, sizeof(double)); ht[n] = calloc(2*P->nb[n]+1, sizeof(double)); }
	if(warm_start)
		{
		for(n=0; n<N; n++) for(i=0; i<P->nu[n]; i++) hux[n][i] = u[n][i];
		for(n=0; n<=N; n++) for(i=0; i<nx[n]; i++) hux[n][P->nu[n]+i] = x[n][i];
		}
	int status;
	orc_prob *P2 = (N2>=1 && N2<N) ? orc_part_cond(P, N2) : NULL;
	if(P2!=NULL)
		{
		/* the IPM runs on the condensed problem, cold start (the reference leaves the condensed initial guess unset,
		 * interfaces/c/fortran_order_interface.c:493-507); the full solution is expanded from it (:511-528) */
		double **ux2 = alloc_ux(P2);
		double **pi2 = malloc((N2+1)*sizeof(double*)), **lam2 = malloc((N2+1)*sizeof(double*)), **t2 = malloc((N2+1)*sizeof(double*));
		for(n=0; n<=N2; n++) { pi2[n] = calloc(P2->nxM+1, sizeof(double)); lam2[n] = calloc(2*P2->nb[n]+1, sizeof(double)); t2[n] = calloc(2*P2->nb[n]+1, sizeof(double)); }
		status = orc_ip2_res_mpc_hard(P2, kk, k_max, mu0, mu_tol, 1e-8, 0, stat, ux2, pi2, lam2, t2);
		orc_part_expand(P, P2, ux2, pi2, lam2, t2, hux, hpi, hlam, ht);
		for(n=0; n<=N2; n++) { free(ux2[n]); free(pi2[n]); free(lam2[n]); free(t2[n]); }
		free(ux2); free(pi2); free(lam2); free(t2);
		orc_prob_free(P2);
		}
	else
		status = orc_ip2_res_mpc_hard(P, kk, k_max, mu0, mu_tol, 1e-8, warm_start, stat, hux, hpi, hlam, ht);
	for(n=0; n<N; n++) for(i=0; i<P->nu[n]; i++) u[n][i] = hux[n][i];
	for(n=0; n<=N; n++) for(i=0; i<nx[n]; i++) x[n][i] = hux[n][P->nu[n]+i];
	for(n=0; n<N; n++)
		for(j=0; j<nb[n] && hidxb[n][j]<P->nu[n]; j++)
			if(lb[n][j]==ub[n][j]) u[n][hidxb[n][j]] = lb[n][j];
	orc_exit_residuals(P, hux, hpi, hlam, ht, inf_norm_res);
	for(n=0; n<N; n++) for(i=0; i<nx[n+1]; i++) pi[n][i] = hpi[n][i];
	/* lam as [lb ub lg ug] per stage, the lib4 order (interfaces/c/fortran_order_interface.c:662-681) */
	for(n=0; n<=N; n++)
		{
		const int nbx = P->nbx[n], g = P->ng[n], nt = P->nb[n];
		for(j=0; j<nbx; j++) { lam[n][j] = hlam[n][j]; lam[n][nbx+j] = hlam[n][nt+j]; }
		for(j=0; j<g; j++) { lam[n][2*nbx+j] = hlam[n][nbx+j]; lam[n][2*nbx+g+j] = hlam[n][nt+nbx+j]; }
		}
	for(n=0; n<=N; n++) { free(hux[n]); free(hpi[n]); free(hlam[n]); free(ht[n]); }
	free(hux); free(hpi); free(hlam); free(ht);
	orc_prob_free(P);
	return status;
	}

/* fortran_order_d_ip_ocp_hard_tv followed by fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv on the same work space
 * (interfaces/c/fortran_order_interface.c:53 and :1082): the IPM runs on (b, q, r, lb, ub); the last KKT system is then
 * solved again for (b2, q2, r2, lb2, ub2) with the matrices and the factor left by the IPM (:1333).  x, u, pi, lam, t receive the
 * result of the SECOND call (lam, t as [lower(nb) upper(nb)] per stage).  Returns the IPM status, or -10 when the IPM left
 * no phase-2 factor behind (the reference reads an uninitialised backup in that case). */
int orc_fortran_order_d_ip_then_kkt_new_rhs(int *kk, int k_max, double mu0, double mu_tol, int N, int *nx, int *nu_N, int *nb, int **hidxb,
		double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub,
		double **b2, double **q2, double **r2, double **lb2, double **ub2,
		double **x, double **u, double **pi, double **lam, double **t)
	{
	int n, i, j;
	orc_prob *P = orc_prob_create(N, nx, nu_N, nb, hidxb);
	orc_prob_set(P, A, B, b, Q, S, R, q, r, lb, ub);
	double **hux = alloc_ux(P);
	double **hpi = malloc((N+1)*sizeof(double*)), **hlam = malloc((N+1)*sizeof(double*)), **ht = malloc((N+1)*sizeof(double*));
	double **bn = malloc((N+1)*sizeof(double*)), **rqn = malloc((N+1)*sizeof(double*));
	for(n=0; n<=N; n++)
		{
		hpi[n] = calloc(P->nxM+1, sizeof(double)); hlam[n] = calloc(2*P->nb[n]+1, sizeof(double)); ht[n] = calloc(2*P->nb[n]+1, sizeof(double));
		bn[n] = calloc(P->nxM+1, sizeof(double)); rqn[n] = calloc(nux_(P,n)+1, sizeof(double));
		}
	double *stat = calloc(5*k_max+5, sizeof(double));
	int status = orc_ip2_res_mpc_hard(P, kk, k_max, mu0, mu_tol, 1e-8, 0, stat, hux, hpi, hlam, ht);
	for(n=0; n<=N; n++)
		{
		int nu = P->nu[n];
		if(n<N) { for(i=0; i<nx[n+1]; i++) bn[n][i] = b2[n][i]; for(i=0; i<nu; i++) rqn[n][i] = r2[n][i]; }
		for(i=0; i<nx[n]; i++) rqn[n][nu+i] = q2[n][i];
		for(j=0; j<nb[n]; j++) { P->d[n][j] = lb2[n][j]; P->d[n][nb[n]+j] = ub2[n][j]; }
		}
	if(orc_kkt_solve_new_rhs(P, bn, rqn, hux, hpi, hlam, ht)) status = -10;
	for(n=0; n<N; n++) for(i=0; i<P->nu[n]; i++) u[n][i] = hux[n][i];
	for(n=0; n<=N; n++) for(i=0; i<nx[n]; i++) x[n][i] = hux[n][P->nu[n]+i];
	for(n=0; n<N; n++) for(i=0; i<nx[n+1]; i++) pi[n][i] = hpi[n][i];
	for(n=0; n<=N; n++) for(j=0; j<2*nb[n]; j++) { lam[n][j] = hlam[n][j]; t[n][j] = ht[n][j]; }
	for(n=0; n<=N; n++) { free(hux[n]); free(hpi[n]); free(hlam[n]); free(ht[n]); free(bn[n]); free(rqn[n]); }
	free(hux); free(hpi); free(hlam); free(ht); free(bn); free(rqn); free(stat);
	orc_prob_free(P);
	return status;
	}

/* unconstrained LQCP, factor+solve, same stage-wise arrays.  mode: 0 = sv ; 1 = trf followed by trs */
void orc_fortran_order_d_ric(int mode, int N, int *nx, int *nu_N, double **A, double **B, double **b, double **Q, double **S,
		double **R, double **q, double **r, double **x, double **u, double **pi)
	{
	int n, i;
	int *nb = calloc(N+1, sizeof(int)); int **idxb = calloc(N+1, sizeof(int*));
	for(n=0; n<=N; n++) idxb[n] = calloc(1, sizeof(int));
	orc_prob *P = orc_prob_create(N, nx, nu_N, nb, idxb);
	orc_prob_set(P, A, B, b, Q, S, R, q, r, b /*unused*/, b /*unused*/);
	double **hux = alloc_ux(P);
	double **hpi = malloc((N+1)*sizeof(double*)), **hPb = malloc((N+1)*sizeof(double*));
	for(n=0; n<=N; n++) { hpi[n] = calloc(P->nxM+1, sizeof(double)); hPb[n] = calloc(P->nxM+1, sizeof(double)); }
	if(mode==0)
		orc_ric_sv(P, NULL, NULL, NULL, NULL, hux, 1, hpi, hPb);
	else
		{
		double **bv = malloc((N+1)*sizeof(double*)), **rq = malloc((N+1)*sizeof(double*));
		for(n=0; n<=N; n++)
			{
			int nux = nux_(P,n), nz = nux+1, j;
			rq[n] = calloc(nux+1, sizeof(double)); bv[n] = calloc(P->nxM+1, sizeof(double));
			for(j=0; j<nux; j++) rq[n][j] = P->RSQrq[n][nux+nz*j];
			if(n<N) for(j=0; j<nx[n+1]; j++) bv[n][j] = P->BAbt[n][nux+nz*j];
			}
		orc_ric_trf(P, NULL);
		orc_ric_trs(P, bv, rq, NULL, hux, 1, hpi, 1, hPb);
		for(n=0; n<=N; n++) { free(bv[n]); free(rq[n]); }
		free(bv); free(rq);
		}
	for(n=0; n<N; n++) for(i=0; i<P->nu[n]; i++) u[n][i] = hux[n][i];
	for(n=0; n<=N; n++) for(i=0; i<nx[n]; i++) x[n][i] = hux[n][P->nu[n]+i];
	for(n=0; n<N; n++) for(i=0; i<nx[n+1]; i++) pi[n][i] = hpi[n][i];
	for(n=0; n<=N; n++) { free(hux[n]); free(hpi[n]); free(hPb[n]); free(idxb[n]); }
	free(hux); free(hpi); free(hPb); free(nb); free(idxb);
	orc_prob_free(P);
	}

/* ------------------------------------------------------------------------------------------- */
/* scenario tree: Riccati factor + solve over a tree of nodes (TEST INFRASTRUCTURE)             */
/*                                                                                             */
/* Follows lqcp_solvers/d_tree_back_ric_rec_libstr.c:                                          */
/*   :524-583  d_tree_back_ric_rec_sv_libstr: nodes in BFS order, backward nn = Nn-1..0,         */
/*             forward nn = 0..Nn-1; edge data (BAbt) indexed by kid-1                           */
/*   :79-156   d_back_ric_sv_back_1_libstr: W = [W_kid0 | W_kid1 | ...], W_k = BAbt_k Lxx_k,     */
/*             last row += l_x,k ; L = chol_mn(RSQrq + W W')                                     */
/*   :160-200  d_back_ric_sv_back_N_libstr: leaf, L = chol_mn(RSQrq)                            */
/*   :204-260  d_back_ric_sv_forw_0 / forw_1: root solves all of ux_0, the others only u;       */
/*             every kid gets x_k = b_k + BAbt_k' ux_dad and pi_k = Lxx_k (Lxx_k' x_k + l_x,k)   */
/* BLASFEO (needed by that file) is not in this image, so parity is pinned indirectly: for a   */
/* path graph this function is checked against orc_ric_sv / the lib4 reference, and for real   */
/* trees against the reference solving the stacked chain problem of                            */
/* test_problems/test_d_tree_ric_libstr.c:797-1018 (tests/test_tree.py).                        */
/*                                                                                             */
/* nodes 0..Nn-1, dad[0] = -1, the kids of a node are first_kid[n] .. first_kid[n]+nkids[n]-1  */
/* BAbt[k] (k >= 1): (nux_dad+1) x nx_k column-major, ld = nux_dad+1 (rows B', A', b')          */
/* RSQrq[n]: (nux_n+1) x nux_n column-major, ld = nux_n+1                                      */
/* out: ux[n] (nu_n + nx_n), pi[k] (nx_k, k >= 1; pi[0] untouched)                              */
/* ------------------------------------------------------------------------------------------- */
void orc_tree_ric_sv(int Nn, const int *dad, const int *first_kid, const int *nkids, const int *nx, const int *nu,
		double *const *BAbt, double *const *RSQrq, double **ux, double **pi)
	{
	int nn, i, j, k, c;
	double **L = calloc(Nn, sizeof(double*)), **dinv = calloc(Nn, sizeof(double*));
	int nzM = 1, nxM = 1;
	for(nn=0; nn<Nn; nn++)
		{
		int nz = nu[nn]+nx[nn]+1;
		L[nn] = calloc((size_t)nz*nz, sizeof(double)); dinv[nn] = calloc(nz, sizeof(double));
		if(nz>nzM) nzM = nz;
		if(nx[nn]>nxM) nxM = nx[nn];
		}
	double *W = calloc((size_t)nzM*nxM, sizeof(double)), *tmp = calloc(nxM+1, sizeof(double));
	/* backward */
	for(nn=Nn-1; nn>=0; nn--)
		{
		int nux = nu[nn]+nx[nn], nz = nux+1;
		double *Ln = L[nn];
		for(j=0; j<nux; j++) for(i=j; i<nz; i++) Ln[i+nz*j] = RSQrq[nn][i+nz*j];
		for(c=0; c<nkids[nn]; c++)
			{
			int kid = first_kid[nn]+c, nx1 = nx[kid], nu1 = nu[kid], nz1 = nx1+nu1+1;
			const double *Lk = L[kid], *M = BAbt[kid];
			for(i=0; i<nz; i++)
				for(j=0; j<nx1; j++)
					{
					double s = 0.0;
					for(k=j; k<nx1; k++) s += M[i+nz*k]*Lk[nu1+k+nz1*(nu1+j)];
					W[i+nz*j] = s;
					}
			for(j=0; j<nx1; j++) W[nux+nz*j] += Lk[nu1+nx1+nz1*(nu1+j)];
			for(j=0; j<nux; j++)
				for(i=j; i<nz; i++)
					{
					double s = 0.0;
					for(k=0; k<nx1; k++) s += W[i+nz*k]*W[j+nz*k];
					Ln[i+nz*j] += s;
					}
			}
		chol_mn(nz, nux, Ln, nz, dinv[nn]);
		}
	/* forward */
	for(nn=0; nn<Nn; nn++)
		{
		int nux = nu[nn]+nx[nn], nz = nux+1;
		int ks = (dad[nn]<0) ? nux : nu[nn];
		const double *Ln = L[nn];
		double *v = ux[nn];
		for(i=0; i<ks; i++) v[i] = -Ln[nux+nz*i];
		for(i=ks-1; i>=0; i--)
			{
			double s = v[i];
			for(j=i+1; j<nux; j++) s -= Ln[j+nz*i]*v[j];
			v[i] = s*dinv[nn][i];
			}
		for(c=0; c<nkids[nn]; c++)
			{
			int kid = first_kid[nn]+c, nx1 = nx[kid], nu1 = nu[kid], nz1 = nx1+nu1+1;
			const double *Lk = L[kid], *M = BAbt[kid];
			double *xk = ux[kid]+nu1;
			for(j=0; j<nx1; j++)
				{
				double s = M[nux+nz*j];
				for(i=0; i<nux; i++) s += M[i+nz*j]*v[i];
				xk[j] = s;
				}
			for(i=0; i<nx1; i++)
				{
				double s = Lk[nu1+nx1+nz1*(nu1+i)];
				for(k=i; k<nx1; k++) s += Lk[nu1+k+nz1*(nu1+i)]*xk[k];
				tmp[i] = s;
				}
			for(i=0; i<nx1; i++)
				{
				double s = 0.0;
				for(k=0; k<=i; k++) s += Lk[nu1+i+nz1*(nu1+k)]*tmp[k];
				pi[kid][i] = s;
				}
			}
		}
	for(nn=0; nn<Nn; nn++) { free(L[nn]); free(dinv[nn]); }
	free(L); free(dinv); free(W); free(tmp);
	}

/* ------------------------------------------------------------------------------------------- */
/* box-constrained IPM over a scenario tree: the IPM above with the tree Riccati               */
/* (mpc_solvers/d_tree_ip2_res_hard_libstr.c:80: "IPM identical to the chain one with N = Nn-1",   */
/* tree residuals mpc_solvers/d_tree_res_ip_res_hard_libstr.c:66).  The reference's tree path    */
/* needs BLASFEO (absent): this is pinned by tests/test_tree_ipm.py against the chain IPM (itself */
/* pinned on the compiled reference) solving the level-stacked problem, and on md = 1 trees.      */
/*                                                                                             */
/* nodes in BFS order; matrices as in orc_tree_ric_sv; d[n] = [lb(nb_n) ; ub(nb_n)]             */
/* in/out: ux[n], lam[n], t[n] (2 nb_n each) ; pi[k], k >= 1 ; stat 5*k_max ; returns the status  */
/* ------------------------------------------------------------------------------------------- */
int orc_tree_ip2_res_mpc_hard(int Nn, const int *dad, const int *nx, const int *nu, const int *nb, int *const *idxb,
		double *const *BAbt, double *const *RSQrq, double *const *d, int *kk, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *stat, double **ux, double **pi, double **lam, double **t)
	{
	int n, N = Nn-1;
	orc_prob *P = orc_prob_create_tree(N, nx, nu, nb, idxb, dad);
	for(n=0; n<=N; n++)
		{
		int nux = nx[n]+nu[n], nz = nux+1;
		memcpy(P->RSQrq[n], RSQrq[n], sizeof(double)*nz*(nux>0 ? nux : 0));
		memcpy(P->d[n], d[n], sizeof(double)*2*P->nb[n]);
		if(n>0) { int nzd = nux_(P, dad[n])+1; memcpy(P->BAbt[n-1], BAbt[n], sizeof(double)*nzd*nx[n]); }
		}
	int status = orc_ip2_res_mpc_hard(P, kk, k_max, mu0, mu_tol, alpha_min, warm_start, stat, ux, pi+1, lam, t);
	orc_prob_free(P);
	return status;
	}

/* factorize, then solve with the stored factors for the (b, rq) held in the matrices: the tree counterparts of
 * d_back_ric_rec_trf / _trs (lqcp_solvers/d_tree_back_ric_rec_libstr.c:591,625) through the dad-aware chain routines above.
 * Same arguments as orc_tree_ric_sv (first_kid / nkids are implied by dad). */
void orc_tree_ric_trf_trs(int Nn, const int *dad, const int *nx, const int *nu, double *const *BAbt, double *const *RSQrq,
		double **ux, double **pi)
	{
	int n, j, N = Nn-1;
	orc_prob *P = orc_prob_create_tree(N, nx, nu, NULL, NULL, dad);
	double **b = malloc((N+1)*sizeof(double*)), **rq = malloc((N+1)*sizeof(double*)), **Pb = malloc((N+1)*sizeof(double*));
	for(n=0; n<=N; n++)
		{
		int nux = nx[n]+nu[n], nz = nux+1;
		memcpy(P->RSQrq[n], RSQrq[n], sizeof(double)*nz*(nux>0 ? nux : 0));
		rq[n] = calloc(nux+1, sizeof(double));
		for(j=0; j<nux; j++) rq[n][j] = RSQrq[n][nux+nz*j];
		b[n] = calloc(P->nxM+1, sizeof(double)); Pb[n] = calloc(P->nxM+1, sizeof(double));
		if(n>0)
			{
			int nuxd = nux_(P, dad[n]), nzd = nuxd+1;
			memcpy(P->BAbt[n-1], BAbt[n], sizeof(double)*nzd*nx[n]);
			for(j=0; j<nx[n]; j++) b[n-1][j] = BAbt[n][nuxd+nzd*j];
			}
		}
	orc_ric_trf(P, NULL);
	orc_ric_trs(P, b, rq, NULL, ux, 1, pi+1, 1, Pb);
	for(n=0; n<=N; n++) { free(b[n]); free(rq[n]); free(Pb[n]); }
	free(b); free(rq); free(Pb);
	orc_prob_free(P);
	}

/* fortran_order_d_ip_ocp_hard_tv_single_newton_step (interfaces/c/fortran_order_interface.c:695): k_max Newton steps from the iterate
 * (ux0, pi0, lam0, t0); ux0[n] = [u_n ; x_n], lam0 / t0 = [lb(nb) ub(nb)] per stage.  x, u, pi, lam, t receive the new iterate. */
int orc_fortran_order_single_newton_step(int *kk, int k_max, double mu0, int N, int *nx, int *nu_N, int *nb, int **hidxb,
		double **A, double **B, double **b, double **Q, double **S, double **R, double **q, double **r, double **lb, double **ub,
		double **x, double **u, double **pi, double **lam, double **t, double *inf_norm_res, double *stat,
		double **ux0, double **pi0, double **lam0, double **t0)
	{
	int n, i, j;
	orc_prob *P = orc_prob_create(N, nx, nu_N, nb, hidxb);
	orc_prob_set(P, A, B, b, Q, S, R, q, r, lb, ub);
	double **hux = alloc_ux(P);
	double **hpi = malloc((N+1)*sizeof(double*)), **hlam = malloc((N+1)*sizeof(double*)), **ht = malloc((N+1)*sizeof(double*));
	for(n=0; n<=N; n++)
		{
		hpi[n] = calloc(P->nxM+1, sizeof(double)); hlam[n] = calloc(2*P->nb[n]+1, sizeof(double)); ht[n] = calloc(2*P->nb[n]+1, sizeof(double));
		for(i=0; i<nux_(P,n); i++) hux[n][i] = ux0[n][i];
		if(n<N) for(i=0; i<nx[n+1]; i++) hpi[n][i] = pi0[n][i];
		for(j=0; j<2*P->nb[n]; j++) { hlam[n][j] = lam0[n][j]; ht[n][j] = t0[n][j]; }
		}
	int status = orc_ip2_res_mpc_hard_single_newton_step(P, kk, k_max, mu0, 1e-8, stat, hux, hpi, hlam, ht);
	for(n=0; n<N; n++) for(i=0; i<P->nu[n]; i++) u[n][i] = hux[n][i];
	for(n=0; n<=N; n++) for(i=0; i<nx[n]; i++) x[n][i] = hux[n][P->nu[n]+i];
	orc_exit_residuals(P, hux, hpi, hlam, ht, inf_norm_res);
	for(n=0; n<N; n++) for(i=0; i<nx[n+1]; i++) pi[n][i] = hpi[n][i];
	for(n=0; n<=N; n++) for(j=0; j<2*nb[n]; j++) { lam[n][j] = hlam[n][j]; t[n][j] = ht[n][j]; }
	for(n=0; n<=N; n++) { free(hux[n]); free(hpi[n]); free(hlam[n]); free(ht[n]); }
	free(hux); free(hpi); free(hlam); free(ht);
	orc_prob_free(P);
	return status;
	}
