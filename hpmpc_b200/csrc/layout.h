/*
 * layout.h -- internal POD descriptors shared by the C host shim and the sm_100a kernels.
 *
 * Device data layout ("native packed layout", one contiguous block of doubles per OCP instance):
 *
 *   for n = 0..N :   [ BAbt_n ]  (nu_n+nx_n+1) x nx_{n+1}, row-major, ld = nx_{n+1}      (n < N only)
 *                                rows 0..nu-1 = B_n' , rows nu..nux-1 = A_n' , row nux = b_n'
 *                    [ RSQrq_n ] packed lower triangle by rows of the (nux+1) x nux trapezoid
 *                                [R S';S Q ; r' q'] : row i holds columns 0..min(i,nux-1),
 *                                element (i,k) at i*(i+1)/2 + k ; row nux (gradient) has nux entries
 *   then, behind the matrices of all stages (so that the stride between the stage matrices does not depend on nb / ng):
 *   for n = 0..N :   [ d_n ]     lb (nb) then ub (nb)
 *                    [ DCt_n ]   [D C]'_n, (nu_n+nx_n) x ng_n row-major          (general constraints lg <= D u + C x <= ug)
 *                    [ dg_n ]    lg (ng) then ug (ng)
 *   every sub-block starts on a 16-byte boundary (sizes padded to an even number of doubles) so a
 *   stage can be moved with 16-byte vector loads or a 1-D bulk (TMA) copy.
 *
 * This replaces the reference's panel-major (bs=4) hpBAbt / hpRSQrq / hd arrays
 * (reference auxiliary/d_aux_lib4.c:1310, interfaces/c/fortran_order_interface.c:176-380): only the
 * unique Hessian entries are stored (SURVEY.md section 8d "algorithmic bytes").
 *
 * The Riccati factor L_n uses the same packed-trapezoid layout as RSQrq_n (the triangular row
 * offsets i*(i+1)/2 are a complete residue system mod 16, so one-row-per-lane accesses in shared memory
 * are bank-conflict free), followed by the nux inverse diagonal entries.
 */
#ifndef HPMPC_B200_LAYOUT_H
#define HPMPC_B200_LAYOUT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_stage
	{
	int nx, nu, nb, nx1;        /* nx1 = nx[n+1], 0 at n = N */
	int off_BAbt, off_RSQ, off_d;   /* into the instance input block (doubles) */
	int off_ux;                 /* into ux-like vectors (sum of nux over earlier stages) */
	int off_pi;                 /* into pi-like vectors, edge n -> length nx1 */
	int off_c;                  /* constraint offset: sum of nb over earlier stages */
	int off_L;                  /* into the factor stash: packed L_n then dinv_n */
	int ng;                     /* general constraints lg <= D u + C x <= ug of the stage (0: none) */
	int off_DCt;                /* [D C]' of the stage: (nu+nx) x ng row-major, in the instance input block */
	int off_dg;                 /* lg (ng) then ug (ng), in the instance input block */
	int pad0, pad1;
	} hb_stage;

typedef struct hb_dims
	{
	int N;
	int nzM;                    /* max over stages of nu+nx+1 */
	int nxM;                    /* max over stages of nx      */
	int nbtot;                  /* constraints per instance: sum of nb + ng ; stage n owns [off_c, off_c+nb+ng): box first, then general */
	int ngtot;                  /* sum of ng (0: box constraints only, the fast paths) */
	long long in_stride;        /* doubles per instance: inputs        */
	long long ux_stride;        /* doubles per instance: ux            */
	long long pi_stride;        /* doubles per instance: pi / Pb / b   */
	long long L_stride;         /* doubles per instance: factor stash  */
	const hb_stage *st;         /* [N+1]  (device pointer in kernel launches) */
	const int *idxb;            /* [nbtot] bound index within its stage's ux */
	const int *c_ux;            /* [nbtot] flat index into the instance's ux vector */
	const struct hb_tnode *tn;  /* scenario-tree IPM only (ric_tree_ipm.cuh): node table [N+1], st is unused then; else NULL */
	} hb_dims;

/* scenario tree (ric_tree.cuh): one entry per node, BFS order; the edge data [B A b]' belongs to the kid */
typedef struct hb_tnode
	{
	int nx, nu, nkids, first_kid, dad;
	int off_BAbt, off_RSQ;      /* into the tree's input block (doubles); off_BAbt is the edge INTO this node */
	int off_ux, off_pi, off_L;  /* node-indexed outputs / factor stash */
	int nb, off_c, off_d;       /* box bounds of the node: count, first flat constraint, [lb(nb) ub(nb)] in the input block */
	int pad0, pad1, pad2;
	} hb_tnode;

typedef struct hb_tdims
	{
	int Nn, nzM, nxM, n_seg;    /* segments: one per top node (ordered by level), then one per tail */
	long long in_stride, ux_stride, pi_stride, L_stride;
	const hb_tnode *tn;         /* [Nn]            (device pointer in kernel launches) */
	const int *seg_start;       /* [n_seg+1]       first entry of each segment in seg_nodes */
	const int *seg_nodes;       /* node indices, top-down within a segment */
	} hb_tdims;

/* flat maps of the multi-kernel tree IPM (tree_ipm_kernels.cu): where the entries of the right-hand sides and of the bounded
 * Hessian diagonal sit in a tree's packed block (device pointers in kernel launches) */
typedef struct hb_tipm_maps
	{
	const int *g_ux;            /* [n_ux]  ux index -> gradient-row entry of its node's RSQrq */
	const int *b_pi;            /* [n_pi]  pi index -> b-row entry of the edge's [B A b]'     */
	const int *c_diag, *c_grad; /* [nbtot] constraint -> Hessian diagonal / gradient-row entry of the bounded variable */
	int n_ux, n_pi;
	} hb_tipm_maps;

/* tails of a scenario tree solved by the size-specialised chain kernel (ric_blk.cuh: hbk_tail_kernel): all tails have the same
 * length and node sizes, and the offsets of the node at position m of tail j are affine in j: pos*[m] + j*str*[m] */
#define HB_TAIL_MAXLEN 64
typedef struct hb_tail_tab
	{
	int len;                           /* nodes per tail; position len-1 is the leaf (nu = 0) */
	int posB[HB_TAIL_MAXLEN], strB[HB_TAIL_MAXLEN];     /* [B A b]' of the edge INTO the node */
	int posQ[HB_TAIL_MAXLEN];                           /* RSQrq of the node (same stride as posB) */
	int posU[HB_TAIL_MAXLEN], strU[HB_TAIL_MAXLEN];     /* ux */
	int posP[HB_TAIL_MAXLEN], strP[HB_TAIL_MAXLEN];     /* pi (multiplier of the edge into the node) */
	int posL[HB_TAIL_MAXLEN], posI[HB_TAIL_MAXLEN], strL[HB_TAIL_MAXLEN];   /* factor slot: generic block at posL, stash image at posI */
	} hb_tail_tab;

/* packed trapezoid helpers */
#define HB_TRI(i) (((i)*((i)+1))>>1)
#define HB_EVEN(x) (((x)+1)&~1)

/* per-instance IPM result record (doubles) : kk, status, inf_norm_res[4], then stat[5*k_max] */
#define HB_IPM_INFO_HEAD 6

/* launchers implemented in ric_kernels.cu ; all pointers are device pointers, stream is a cudaStream_t */
int hb_launch_ric_sv(const hb_dims *dims, long long n_inst, const double *in, double *ux, double *pi, double *Pb,
		double *stash, int n_slots, int grid, int warps, void *stream, const double *Qx, const double *qx);
int hb_launch_ric_trf(const hb_dims *dims, long long n_inst, const double *in, double *L, int grid, int warps, void *stream, const double *Qx);
int hb_launch_ric_trs(const hb_dims *dims, long long n_inst, const double *in, const double *L, double *ux, double *pi,
		double *work, int n_slots, int grid, int warps, void *stream, const double *qx);
long long hb_trs_shared_smem_bytes(const hb_dims *dims, const hb_stage *st_host, int warps, int *resident);
int hb_launch_ric_trs_shared(const hb_dims *dims, long long n_inst, const double *in_shared, const double *L_shared, const double *vec,
		double *ux, double *pi, double *work, int grid, int warps, int smem, int resident, void *stream);
int hb_launch_ipm(const hb_dims *dims, long long n_inst, const double *in, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int n_slots, int grid, int warps, int *counter, int fast_id, void *stream);
int hb_launch_ipm_kkt(const hb_dims *dims, long long n_inst, const double *in, int k_max, double mu0, double mu_tol,
		double alpha_min, int warm_start, double *ux, double *pi, double *lam, double *t, double *info,
		double *work, long long work_stride, int n_slots, int grid, int warps, int *counter, int fast_id, void *stream,
		double *kkt, long long kkt_stride);
int hb_launch_kkt_new_rhs(const hb_dims *dims, long long n_inst, const double *in, double *kkt, long long kkt_stride,
		double *ux, double *pi, double *lam, double *t, double *info, int grid, int warps, int *counter, int fast_id, void *stream);
/* fast_id -2 : dims->tn describes a scenario tree (one warp per tree, generic node sizes) */
#define HB_IPM_TREE (-2)
int hb_ipm_fast_variant(int N, const int *nx, const int *nu, int nbtot);
int hb_ric_shape_variant(int N, const int *nx, const int *nu);      /* same sweep sets for trf / trs, bounds not required */
int hb_launch_ric_trf_trs_fast(int id, int mode /* 0 factor, 1 solve */, const hb_dims *dims, long long n_inst, const double *in, double *L,
		long long L_stride, double *ux, double *pi, double *work, int grid, int warps, void *stream);
int hb_ipm_fast_info(int id, int N, int *smem_warp, long long *L_doubles);
long long hb_ipm_work_doubles2(const hb_dims *dims, long long L_doubles);
int hb_fast_variant(int N, const int *nx, const int *nu);
int hb_fast_info(int id, int N, int *ipw, int *smem_warp, long long *stash_per_inst);
int hb_launch_ric_sv_fast(int id, const hb_dims *dims, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, void *stream);
long long hb_cipm_aux_bytes(long long n_inst);
int hb_launch_cipm(const hb_dims *dims, long long n_inst, const double *in, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *ux, double *pi, double *lam, double *t, double *info, double *work, long long work_stride, void *aux,
		int grid, int warps, int sms, int fast_id, void *stream);
long long hb_res_work_doubles(const hb_dims *dims);
int hb_launch_res(const hb_dims *dims, long long n_inst, const double *in, const double *ux, const double *pi, const double *lam,
		const double *t, double *rq, double *rb, double *rd, double *rm, double *mu, double *work, int grid, int warps, void *stream);
int hb_launch_sv_traffic(int id, const hb_dims *dims, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, void *stream);
int hb_launch_tree(const hb_tdims *dims, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode /* 0 backward, 1 forward, 2 backward then forward */, int seg_lo, int seg_hi, int grid, int warps, void *stream, const double *skip);
int hb_tail_variant(int nx, int nu);                     /* -1 when no size-specialised tail kernel exists */
int hb_tail_info(int id, int *ipw, int *smem_warp, int *image_doubles);
int hb_launch_tail(int id, const hb_tdims *dims, const hb_tail_tab *tab, long long n_trees, const double *in, double *ux, double *pi,
		double *L, int mode, int tail_lo, int tail_hi, int grid, int warps, void *stream, const double *skip);
int hb_launch_top(int id, const hb_tdims *dims, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int first, int grid, int warps, void *stream, const double *skip);
long long hb_tipm_work_doubles(const hb_dims *dims);
int hb_launch_tipm_step(const hb_dims *dims, const hb_tipm_maps *maps, int part, long long n_trees, const double *in, double *in_mod, int k_max,
		double mu0, double mu_tol, double alpha_min, int warm_start, double *ux, double *pi, double *dux, double *dpi, double *lam,
		double *t, double *info, double *work, long long work_stride, double *state, int *counters, void *stream);
int hb_launch_tipm_gate(const int *counters, double *gate, int init, void *stream);
int hb_launch_tipm_res(const hb_dims *dims, long long n_trees, const double *in, const double *ux, const double *pi,
		double *dux, double *dpi, double *work, long long work_stride, double *state, int sms, void *stream);
int hb_launch_tree_trf_trs(const hb_dims *dims, long long n_trees, const double *in, double *L, double *ux, double *pi,
		double *work, int n_slots, int mode /* 0 factor, 1 solve with stored factors */, int grid, int warps, void *stream);
long long hb_ipm_work_doubles(const hb_dims *dims);
int hb_smem_bytes_per_warp(const hb_dims *dims);
int hb_smem_bytes_per_warp_sz(int nzM, int nxM);
int hb_device_sm_count(int device);
double hb_fp64_peak_probe(int device, int iters, void *stream);   /* measured DFMA TFLOP/s */

#ifdef __cplusplus
}
#endif
#endif
