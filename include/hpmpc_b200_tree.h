/*
 * hpmpc_b200_tree.h -- batched Riccati factor+solve over scenario trees (BASELINE config 5).
 *
 * Replaces, for a batch of trees that share one topology and size pattern,
 *     d_tree_back_ric_rec_sv_libstr        reference include/lqcp_solvers.h (lqcp_solvers/d_tree_back_ric_rec_libstr.c:524)
 * whose own signature takes BLASFEO matrix structs (BLASFEO is an external, un-vendored dependency of the reference,
 * Makefile.rule:47-48), so the batched entry points below take the tree (`struct node`, reference include/tree.h:34-44,
 * same fields in the same order) plus plain arrays in the packed layout described by hpmpc_b200_tree_node_offsets().
 *
 * A tree is cut at the first level from which every node has at most one kid (the robust horizon) into the "top" and
 * one "tail" chain per node of that level.  Tails are independent: phase 0 (backward over tails) and phase 2 (forward
 * over tails) can run on different GPUs for different tail ranges; phase 1 (backward + forward over the top) needs the
 * factor blocks of all tail roots, which is the one exchange of the multi-GPU path (hpmpc_b200_tree_tail_root()).
 *
 * Every function returns 0 on success, negative on error; nothing falls back to the CPU.
 */
#ifndef HPMPC_B200_TREE_H
#define HPMPC_B200_TREE_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef TREE_MPC
#define TREE_MPC
struct node          /* reference include/tree.h:34-44 */
	{
	int *kids;
	int idx;
	int dad;
	int nkids;
	int stage;
	int real;
	int idxkid;
	};
#endif

typedef struct hpmpc_b200_tree hpmpc_b200_tree;

typedef struct hpmpc_b200_tree_sizes
	{
	long long in_stride, ux_stride, pi_stride, L_stride;   /* doubles per tree */
	int Nn, nzM, nxM;
	int n_tails, n_top_nodes, cut_stage;
	int n_shard_nodes;          /* nodes of the deepest top level (cut_stage-1): the roots of the subtrees that shard over GPUs */
	} hpmpc_b200_tree_sizes;

/* nodes in BFS order (kids of a node contiguous, as the reference's setup_tree builds them,
 * test_problems/test_d_tree_ip_hard_libstr.c:93-176); nx, nu are node-indexed, nu of a leaf is taken as given */
int  hpmpc_b200_tree_create(hpmpc_b200_tree **out, int Nn, const struct node *tree, const int *nx, const int *nu, int device);
void hpmpc_b200_tree_destroy(hpmpc_b200_tree *t);
void hpmpc_b200_tree_sizes_get(const hpmpc_b200_tree *t, hpmpc_b200_tree_sizes *out);
/* offsets (doubles) of node n: [B A b]' of the edge into n ((nux_dad+1) x nx_n row-major; -1 for the root), RSQrq_n
 * (packed lower trapezoid by rows, gradient row last), ux_n, pi_n (multiplier of the edge into n), L_n in the stash */
void hpmpc_b200_tree_node_offsets(const hpmpc_b200_tree *t, int n, int *off_BAbt, int *off_RSQ, int *off_ux, int *off_pi, int *off_L);
/* tail j = 0..n_tails-1: its root node and where that node's factor block sits in the per-tree stash */
void hpmpc_b200_tree_tail_root(const hpmpc_b200_tree *t, int tail, int *node, int *off_L, int *len_L);
/* subtree k = 0..n_shard_nodes-1: its root node (level cut_stage-1), that node's factor block in the per-tree stash, and the
 * range of tails [tail_lo, tail_hi) below it */
void hpmpc_b200_tree_shard_node(const hpmpc_b200_tree *t, int k, int *node, int *off_L, int *len_L, int *tail_lo, int *tail_hi);
/* host-side packing of one tree from node/edge-indexed column-major arrays: A[k] nx_k x nx_dad, B[k] nx_k x nu_dad,
 * b[k] nx_k for k >= 1 (entry 0 unused); Q[n] nx x nx, S[n] nu x nx, R[n] nu x nu, q[n], r[n] for every node */
int  hpmpc_b200_tree_pack_instance(const hpmpc_b200_tree *t, double *const *A, double *const *B, double *const *b,
                                   double *const *Q, double *const *S, double *const *R, double *const *q, double *const *r, double *block);

/* ---- box-constrained IPM over the tree:  d_tree_ip2_res_mpc_hard_libstr (reference include/mpc_solvers.h,
 * mpc_solvers/d_tree_ip2_res_hard_libstr.c:80), ng = 0 ----
 * nb[n] bounds at node n on the entries idxb[n][0..nb[n]) of [u_n ; x_n]; with nb == NULL this is hpmpc_b200_tree_create.
 * The bounds [lb(nb) ub(nb)] of a node follow its RSQrq in the packed block (hpmpc_b200_tree_bound_offsets: off_d);
 * constraints are numbered node by node (off_c), and d_lam / d_t hold per tree, per node, [lower(nb) upper(nb)] at 2*off_c. */
int  hpmpc_b200_tree_create_box(hpmpc_b200_tree **out, int Nn, const struct node *tree, const int *nx, const int *nu,
                                const int *nb, int *const *idxb, int device);
void hpmpc_b200_tree_bound_offsets(const hpmpc_b200_tree *t, int n, int *nb, int *off_c, int *off_d);
int  hpmpc_b200_tree_pack_bounds(const hpmpc_b200_tree *t, double *const *lb, double *const *ub, double *block);   /* after pack_instance */
/* same arguments and return record as hpmpc_b200_d_ip2_res_mpc_hard_batch (hpmpc_b200.h): d_info holds 6 + 5*k_max doubles per
 * tree (kk, status, |res_q|, |res_b|, |res_d|, mu, then sigma/alpha_aff/mu_aff/alpha/mu per iteration); status 0 converged,
 * 1 kk >= k_max, 2 alpha < alpha_min (mpc_solvers/d_ip2_res_hard.c:1331-1343).  d_ux, d_pi in the node-indexed layouts above. */
int  hpmpc_b200_d_tree_ip2_res_mpc_hard_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in, int k_max, double mu0,
                                              double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi,
                                              double *d_lam, double *d_t, double *d_info, void *stream);

/* factor only / solve with stored factors:  d_tree_back_ric_rec_trf_libstr, d_tree_back_ric_rec_trs_libstr
 * (lqcp_solvers/d_tree_back_ric_rec_libstr.c:591, 625).  _trf writes the node factors to d_L (L_stride doubles per tree); _trs
 * solves with them for the b and [r q] held in d_in, so a caller re-solves by re-packing only the right-hand sides. */
int hpmpc_b200_d_tree_back_ric_rec_trf_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in, double *d_L, void *stream);
int hpmpc_b200_d_tree_back_ric_rec_trs_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in, const double *d_L,
                                             double *d_ux, double *d_pi, void *stream);

/* kernels launched through this handle so far (measurement: bench.py's gpu_launches) */
long long hpmpc_b200_tree_launch_count(const hpmpc_b200_tree *t);

/* whole solve on one GPU (phases 0, 1, 2 back to back) */
int hpmpc_b200_d_tree_back_ric_rec_sv_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in,
                                            double *d_ux, double *d_pi, double *d_L, void *stream);
/* one phase; lo/hi are tail indices for phases 0 and 2, subtree indices for phases 3 and 5, ignored otherwise:
 *   0 = backward over tails [lo, hi)          2 = forward over tails [lo, hi)
 *   1 = the whole top (backward + forward)    -- tail sharding: 0, exchange tail-root blocks, 1, 2
 *   3 = backward over subtree roots [lo, hi)  4 = levels above the subtree roots (backward + forward)  5 = forward over subtree roots [lo, hi)
 *                                             -- subtree sharding: 0, 3, exchange subtree-root blocks, 4, 5, 2 */
int hpmpc_b200_d_tree_back_ric_rec_sv_phase(hpmpc_b200_tree *t, long long n_trees, int phase, int tail_lo, int tail_hi,
                                            const double *d_in, double *d_ux, double *d_pi, double *d_L, void *stream);

/* ---- several GPUs: subtrees shard over the ranks, one all-gather of subtree-root factor blocks per solve,
 * issued by the library with NCCL on the caller's stream (reference recursion: lqcp_solvers/d_tree_back_ric_rec_libstr.c:524-583).
 * One process per GPU.  Rank 0 calls hpmpc_b200_comm_unique_id (128 bytes) and hands the id to the other ranks by whatever means the
 * application has (MPI, a file, torch.distributed, ...); every rank then creates its communicator.  hpmpc_b200_comm_wrap adopts an
 * ncclComm_t the application already owns.  The tree must have n_shard_nodes divisible by the number of ranks. */
typedef struct hpmpc_b200_comm hpmpc_b200_comm;
int  hpmpc_b200_comm_unique_id(void *id, int id_bytes);
int  hpmpc_b200_comm_create(hpmpc_b200_comm **out, int world, int rank, const void *id, int device);
int  hpmpc_b200_comm_wrap(hpmpc_b200_comm **out, void *nccl_comm, int world, int rank);
void hpmpc_b200_comm_destroy(hpmpc_b200_comm *c);
/* every rank holds all trees' data and calls this with the same arguments; on return it has ux / pi of the nodes of its subtrees
 * and of the levels above them.  Non-blocking: everything is enqueued on `stream`. 
 * CONCURRENCY: the IPM, _trs_batch and _sv_batch_mg use scratch owned by the handle; the library makes the stream of such a call wait for
 * the handle's previous one (an event), so calls issued to different streams are serialised, never raced. */
int  hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg(hpmpc_b200_tree *t, hpmpc_b200_comm *c, long long n_trees, const double *d_in,
                                                double *d_ux, double *d_pi, double *d_L, void *stream);

#ifdef __cplusplus
}
#endif
#endif
