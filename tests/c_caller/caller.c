/*
 * tests/c_caller/caller.c -- a plain C program written against HPMPC's public C interface, linked against
 * libhpmpc_b200.so instead of libhpmpc.a.  TEST CODE.
 *
 * Built two ways by tests/test_c_caller.py:
 *   -DWITH_REFERENCE_HEADERS -I/root/reference/include (build container only): the reference's own c_interface.h /
 *       mpc_solvers.h / lqcp_solvers.h are included NEXT TO include/hpmpc_compat.h, so the compiler rejects any prototype of ours
 *       that differs from the reference's ("conflicting types");
 *   with include/hpmpc_compat.h alone (GPU box, where /root/reference does not exist): linked and RUN.
 * It reads one problem from a flat binary file (written by the test), calls fortran_order_d_ip_ocp_hard_tv the way
 * test_problems/test_d_ip_hard_libstr.c:964 does, and writes kk, status, inf_norm_res and the solution to a second file.
 */
#include <stdio.h>
#include <stdlib.h>
#ifdef WITH_REFERENCE_HEADERS
#include "c_interface.h"
#include "mpc_solvers.h"
#include "lqcp_solvers.h"
#endif
#include "hpmpc_compat.h"

static double *rd(FILE *f, int n) { double *p = calloc(n>0 ? n : 1, sizeof(double)); if(n>0 && fread(p, sizeof(double), n, f)!=(size_t)n) exit(3); return p; }

int main(int argc, char **argv)
	{
	if(argc<3) { fprintf(stderr, "usage: caller problem.bin result.bin\n"); return 2; }
	FILE *f = fopen(argv[1], "rb");
	if(!f) return 2;
	int hdr[4];
	if(fread(hdr, sizeof(int), 4, f)!=4) return 3;
	int N = hdr[0], k_max = hdr[1], n, j;
	int *nx = malloc((N+1)*sizeof(int)), *nu = malloc((N+1)*sizeof(int)), *nb = malloc((N+1)*sizeof(int)), *ng = calloc(N+1, sizeof(int));
	if(fread(nx, sizeof(int), N+1, f)!=(size_t)(N+1) || fread(nu, sizeof(int), N+1, f)!=(size_t)(N+1) || fread(nb, sizeof(int), N+1, f)!=(size_t)(N+1)) return 3;
	int **idxb = malloc((N+1)*sizeof(int*));
	for(n=0; n<=N; n++) { idxb[n] = malloc((nb[n]+1)*sizeof(int)); if(nb[n]>0 && fread(idxb[n], sizeof(int), nb[n], f)!=(size_t)nb[n]) return 3; }
	double **A = malloc((N+1)*sizeof(double*)), **B = malloc((N+1)*sizeof(double*)), **b = malloc((N+1)*sizeof(double*)), **Q = malloc((N+1)*sizeof(double*)),
		**S = malloc((N+1)*sizeof(double*)), **R = malloc((N+1)*sizeof(double*)), **q = malloc((N+1)*sizeof(double*)), **r = malloc((N+1)*sizeof(double*)),
		**lb = malloc((N+1)*sizeof(double*)), **ub = malloc((N+1)*sizeof(double*)), **x = malloc((N+1)*sizeof(double*)), **u = malloc((N+1)*sizeof(double*)),
		**pi = malloc((N+1)*sizeof(double*)), **lam = malloc((N+1)*sizeof(double*)), **e = malloc((N+1)*sizeof(double*));
	for(n=0; n<=N; n++)
		{
		int nx1 = n<N ? nx[n+1] : 0, nun = n<N ? nu[n] : 0;
		A[n] = rd(f, nx1*nx[n]); B[n] = rd(f, nx1*nun); b[n] = rd(f, nx1);
		Q[n] = rd(f, nx[n]*nx[n]); S[n] = rd(f, nun*nx[n]); R[n] = rd(f, nun*nun); q[n] = rd(f, nx[n]); r[n] = rd(f, nun);
		lb[n] = rd(f, nb[n]); ub[n] = rd(f, nb[n]);
		x[n] = calloc(nx[n]+1, sizeof(double)); u[n] = calloc(nun+1, sizeof(double)); pi[n] = calloc(nx1+1, sizeof(double));
		lam[n] = calloc(2*nb[n]+1, sizeof(double)); e[n] = calloc(1, sizeof(double));
		}
	fclose(f);
	int wsz = hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes(N, nx, nu, nb, idxb, ng, N);
	void *work = malloc(wsz+64);
	double inf_norm_res[4] = {0, 0, 0, 0}, *stat = calloc(5*k_max+5, sizeof(double));
	int kk = -1;
	int status = fortran_order_d_ip_ocp_hard_tv(&kk, k_max, 2.0, 1e-8, N, nx, nu, nb, idxb, ng, N, 0, A, B, b, Q, S, R, q, r, lb, ub, e, e, e, e,
			x, u, pi, lam, inf_norm_res, work, stat);
	f = fopen(argv[2], "wb");
	if(!f) return 2;
	int out[2] = { kk, status };
	fwrite(out, sizeof(int), 2, f);
	fwrite(inf_norm_res, sizeof(double), 4, f);
	for(n=0; n<N; n++) fwrite(u[n], sizeof(double), nu[n], f);
	for(n=0; n<=N; n++) fwrite(x[n], sizeof(double), nx[n], f);
	for(n=0; n<N; n++) fwrite(pi[n], sizeof(double), nx[n+1], f);
	for(n=0; n<=N; n++) fwrite(lam[n], sizeof(double), 2*nb[n], f);
	fclose(f);
	printf("kk = %d, status = %d, mu = %e\n", kk, status, inf_norm_res[3]);
	(void)j;
	return 0;
	}
