/*
 * pcond_launch.h -- launchers of the partial-condensing kernels (pcond_kernels.cu), called from pcond.c.
 * All pointers are device pointers; `dF` / `dC` describe the full and the condensed problem (layout.h), their st / idxb
 * members are device pointers.
 */
#ifndef HPMPC_B200_PCOND_LAUNCH_H
#define HPMPC_B200_PCOND_LAUNCH_H
#include "layout.h"
#ifdef __cplusplus
extern "C" {
#endif

/* one block of consecutive stages that becomes one stage of the condensed problem */
typedef struct hb_pc_block
	{
	int n0, T;                 /* first stage and number of stages of the block */
	int off_G;                 /* scratch: offsets (doubles) of Gamma_j, j = 0..T-2, are off_G + the running sum of rows_j*nx_{j+1} */
	int pad;
	} hb_pc_block;

/* doubles of per-warp scratch the condensing kernel needs for this pair of problems */
long long hb_pcond_scratch_doubles(const hb_stage *stF, int N, const hb_pc_block *blk, int N2);
int hb_launch_pcond(const hb_dims *dF, const hb_dims *dC, const hb_pc_block *blk, int N2, long long n_inst, const double *in_full,
		double *in_cond, double *scratch, long long scratch_stride, int grid, int warps, void *stream);
int hb_launch_pexpand(const hb_dims *dF, const hb_dims *dC, const hb_pc_block *blk, int N2, long long n_inst, const double *in_full,
		const double *ux2, const double *pi2, const double *lam2, const double *t2, double *ux, double *pi, double *lam, double *t,
		int grid, int warps, void *stream);
/* info[2..5] of every instance <- max |rq|, max |rb|, max |rd|, mu  (the exit norms of the high-level wrappers) */
int hb_launch_res_norms(const hb_dims *dF, long long n_inst, const double *rq, const double *rb, const double *rd, const double *mu,
		long long lam_stride, double *info, long long info_stride, void *stream);

#ifdef __cplusplus
}
#endif
#endif
