/*
 * tree.c -- host side (plain C) of the scenario-tree Riccati path: topology -> node table, segments (top + tails),
 * device layout, packing, phase launches.  The arithmetic runs in ric_tree.cuh / ric_kernels.cu.
 * Reference: include/tree.h:34-44 (struct node), lqcp_solvers/d_tree_back_ric_rec_libstr.c:524-583 (node order, edges
 * indexed by the kid), test_problems/test_d_tree_ip_hard_libstr.c:93-176 (how callers build the tree).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include "layout.h"
#include "../../include/hpmpc_b200_tree.h"

#define CK(x) do { cudaError_t e_ = (x); if(e_!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

struct hpmpc_b200_tree
	{
	int device, Nn, cut_stage, n_tails, n_top;
	hb_tnode *tn;            /* host copy */
	int *stage;
	int *seg_start, *seg_nodes, n_seg;   /* segments: every top node on its own (ordered by level), then one segment per tail */
	int *lvl_seg;            /* [cut+1] first segment of each top level */
	int *slot;               /* [Nn] size of each node's factor slot in the stash (doubles) */
	int *tail_root;
	hb_tdims dims;           /* device pointers inside */
	int sms;
	int tail_fast_id, t_ipw, t_smem_warp;   /* size-specialised tail kernel, -1 when the tails do not qualify */
	int top_fast;                           /* the top runs on the size-specialised stage routines too */
	hb_tail_tab tab;
	/* box-constrained IPM over the tree (hpmpc_b200_tree_create_box): flat constraint tables and the per-warp work slots */
	int nbtot;
	int *idxb, *c_ux;        /* host copies [nbtot] */
	hb_dims ipm_dims;        /* device pointers inside; st unused, tn = the node table */
	double *ipm_ws; long long ipm_ws_stride; int ipm_slots; int *ipm_counter;
	/* multi-kernel IPM around the size-specialised tree Riccati (tree_ipm_kernels.cu): flat maps and per-tree buffers */
	hb_tipm_maps maps;       /* device pointers inside */
	long long f_trees;       /* trees the buffers below are sized for */
	double *f_in_mod, *f_dux, *f_dpi, *f_L, *f_ws, *f_state; long long f_ws_stride;
	int *f_nact, *h_nact;    /* device / pinned host: trees still iterating */
	long long n_launches;    /* kernels launched through this handle so far (bench.py reports it) */
	double *trs_ws; int trs_slots;          /* right-hand sides and Pb of the solve-only path, per warp slot */
	const double *skip_state;               /* tree IPM driver: per-tree state records; the Riccati kernels skip finished trees */
	double *mg_send, *mg_recv; size_t mg_send_bytes, mg_recv_bytes;   /* staging of the multi-GPU exchange (tree_mg below) */
	/* one call in flight per handle for the entry points that use handle-owned scratch (IPM, trs, multi-GPU exchange): the stream
	 * of a new call first waits for the previous call's work, so calls on different streams are serialised instead of racing */
	cudaEvent_t ev_busy; int busy_valid, busy_made;
	};

static int tcall_begin(hpmpc_b200_tree *t, void *stream)
	{
	if(!t->busy_made) { CK(cudaEventCreateWithFlags(&t->ev_busy, cudaEventDisableTiming)); t->busy_made = 1; }
	if(t->busy_valid) CK(cudaStreamWaitEvent((cudaStream_t)stream, t->ev_busy, 0));
	return 0;
	}
static int tcall_end(hpmpc_b200_tree *t, void *stream, int rc)
	{
	if(rc) return rc;
	CK(cudaEventRecord(t->ev_busy, (cudaStream_t)stream));
	t->busy_valid = 1;
	return 0;
	}

int hpmpc_b200_tree_create(hpmpc_b200_tree **out, int Nn, const struct node *tree, const int *nx, const int *nu, int device)
	{
	return hpmpc_b200_tree_create_box(out, Nn, tree, nx, nu, NULL, NULL, device);
	}

static int tree_create_impl(hpmpc_b200_tree **out, hpmpc_b200_tree **partial, int Nn, const struct node *tree, const int *nx, const int *nu,
		const int *nb, int *const *idxb, int device);

/* every error return of the construction goes through here: whatever was allocated so far is released */
int hpmpc_b200_tree_create_box(hpmpc_b200_tree **out, int Nn, const struct node *tree, const int *nx, const int *nu,
		const int *nb, int *const *idxb, int device)
	{
	hpmpc_b200_tree *partial = NULL;
	int rc = tree_create_impl(out, &partial, Nn, tree, nx, nu, nb, idxb, device);
	if(rc) { hpmpc_b200_tree_destroy(partial); *out = NULL; }
	return rc;
	}

static int tree_create_impl(hpmpc_b200_tree **out, hpmpc_b200_tree **partial, int Nn, const struct node *tree, const int *nx, const int *nu,
		const int *nb, int *const *idxb, int device)
	{
	int n, j;
	*out = NULL;
	if(Nn<1) return -2;
	hpmpc_b200_tree *t = calloc(1, sizeof(*t));
	if(!t) return -1;
	*partial = t;
	t->device = device; t->Nn = Nn;
	t->tn = calloc(Nn, sizeof(hb_tnode));
	t->stage = calloc(Nn, sizeof(int));
	t->slot = calloc(Nn, sizeof(int));
	if(!t->tn || !t->stage || !t->slot) return -1;
	int nzM = 1, nxM = 1, max_stage = 0;
	for(n=0; n<Nn; n++)
		{
		hb_tnode *s = &t->tn[n];
		s->nx = nx[n]; s->nu = nu[n]; s->nkids = tree[n].nkids; s->dad = (n==0) ? -1 : tree[n].dad;
		if(n>0 && (s->dad<0 || s->dad>=n)) { fprintf(stderr, "hpmpc_b200: tree: node %d: dad = %d must precede the node (BFS order)\n", n, s->dad); return -2; }
		s->first_kid = s->nkids>0 ? tree[n].kids[0] : -1;
		for(j=0; j<s->nkids; j++)
			if(tree[n].kids[j]!=s->first_kid+j || s->first_kid<=n || s->first_kid+j>=Nn)
				{ fprintf(stderr, "hpmpc_b200: tree: node %d: kids must be contiguous and follow their dad (BFS order)\n", n); return -2; }
		t->stage[n] = (n==0) ? 0 : t->stage[s->dad]+1;
		if(t->stage[n]>max_stage) max_stage = t->stage[n];
		if(s->nu+s->nx+1>nzM) nzM = s->nu+s->nx+1;
		if(s->nx>nxM) nxM = s->nx;
		s->nb = nb ? nb[n] : 0;
		if(s->nb<0 || s->nb>s->nu+s->nx) { fprintf(stderr, "hpmpc_b200: tree: node %d: nb = %d out of range\n", n, s->nb); return -2; }
		s->off_c = t->nbtot; t->nbtot += s->nb;
		}
	if(nzM>64) { fprintf(stderr, "hpmpc_b200: tree: nu+nx+1 = %d > 64 is not supported\n", nzM); return -2; }
	/* cut: smallest stage c such that every node of stage >= c has at most one kid */
	int cut = 0;
	for(n=0; n<Nn; n++) if(t->tn[n].nkids>1 && t->stage[n]+1>cut) cut = t->stage[n]+1;
	t->cut_stage = cut;
	/* do all tails look alike (same length, uniform (nx, nu), leaf without inputs) and does a size-specialised kernel exist? */
	int tail_len = 0, tnx = -1, tnu = -1, uniform = 1, img = 0;
	for(n=0; n<Nn && uniform; n++)
		if(t->stage[n]>=cut)
			{
			const hb_tnode *s = &t->tn[n];
			if(s->nkids>0) { if(tnx<0) { tnx = s->nx; tnu = s->nu; } if(s->nx!=tnx || s->nu!=tnu) uniform = 0; }
			else if(s->nu!=0 || (tnx>=0 && s->nx!=tnx)) uniform = 0;
			}
	tail_len = max_stage - cut + 1;
	for(n=0; n<Nn && uniform; n++) if(t->tn[n].nkids==0 && t->stage[n]!=max_stage) uniform = 0;
	t->tail_fast_id = -1;
	if(uniform && tnx>0 && tail_len>=3 && tail_len<=HB_TAIL_MAXLEN && getenv("HPMPC_B200_NO_FAST")==NULL)
		{
		t->tail_fast_id = hb_tail_variant(tnx, tnu);
		if(t->tail_fast_id>=0) hb_tail_info(t->tail_fast_id, &t->t_ipw, &t->t_smem_warp, &img);
		}
	/* the top qualifies when the root has no state (x0 eliminated), every other top node has the tails' (nx, nu) and the
	 * nodes of a level all have the same number of kids */
	t->top_fast = (t->tail_fast_id>=0 && cut>0 && getenv("HPMPC_B200_NO_FAST_TOP")==NULL);
	for(n=0; n<Nn && t->top_fast; n++)
		if(t->stage[n]<cut)
			{
			const hb_tnode *s = &t->tn[n];
			if(n==0) { if(s->nx!=0 || s->nu!=tnu) t->top_fast = 0; }
			else if(s->nx!=tnx || s->nu!=tnu) t->top_fast = 0;
			for(j=n+1; j<Nn && t->stage[j]==t->stage[n]; j++) if(t->tn[j].nkids!=s->nkids) t->top_fast = 0;
			}
	long long o_in = 0, o_ux = 0, o_pi = 0, o_L = 0;
	int n_top = 0, n_tails = 0;
	for(n=0; n<Nn; n++)
		{
		hb_tnode *s = &t->tn[n];
		int nux = s->nu+s->nx;
		if(n>0)
			{
			int nuxd = t->tn[s->dad].nu + t->tn[s->dad].nx;
			s->off_BAbt = (int)o_in; o_in += HB_EVEN((nuxd+1)*s->nx);
			}
		else s->off_BAbt = -1;
		s->off_RSQ = (int)o_in; o_in += HB_EVEN(HB_TRI(nux)+nux);
		s->off_d = (int)o_in; o_in += HB_EVEN(2*s->nb);           /* [lb(nb) ub(nb)] */
		s->off_ux = (int)o_ux; o_ux += nux;
		s->off_pi = (int)o_pi; o_pi += s->nx;
		s->off_L = (int)o_L;
		{
		/* factor slot: the generic packed block; for tail nodes of the size-specialised kernel also room for its stash image
		 * (behind the generic block at the tail root, which the top reads in the generic form) */
		int gen = HB_EVEN(HB_TRI(nux)+2*nux), slot = gen;
		if(t->tail_fast_id>=0 && t->stage[n]>=cut) slot = (t->stage[n]==cut) ? gen + img : (gen>img ? gen : img);
		if(t->top_fast && t->stage[n]<cut) slot = gen + img;
		o_L += slot;
		t->slot[n] = slot;
		}
		if(t->stage[n]<cut) n_top++;
		if(t->stage[n]==cut) n_tails++;
		}
	t->n_top = n_top; t->n_tails = n_tails; t->n_seg = n_top+n_tails;
	t->seg_start = calloc(t->n_seg+1, sizeof(int)); t->seg_nodes = calloc(Nn, sizeof(int)); t->tail_root = calloc(n_tails+1, sizeof(int));
	t->lvl_seg = calloc(cut+2, sizeof(int));
	int pos = 0, seg = 0, lv;
	t->seg_start[0] = 0;
	/* top: nodes are in BFS order, so the nodes of a level are contiguous; one single-node segment each (level-parallel launches) */
	for(lv=0; lv<cut; lv++)
		{
		t->lvl_seg[lv] = seg;
		for(n=0; n<Nn; n++) if(t->stage[n]==lv) { t->seg_nodes[pos++] = n; t->seg_start[++seg] = pos; }
		}
	t->lvl_seg[cut] = seg;
	for(n=0; n<Nn; n++)
		if(t->stage[n]==cut)
			{
			int m = n;
			t->tail_root[seg-n_top] = n;
			for(;;)
				{
				t->seg_nodes[pos++] = m;
				if(t->tn[m].nkids==0) break;
				m = t->tn[m].first_kid;
				}
			t->seg_start[++seg] = pos;
			}
	if(pos!=Nn) { fprintf(stderr, "hpmpc_b200: tree: inconsistent topology (%d of %d nodes reached)\n", pos, Nn); return -2; }
	if(t->tail_fast_id>=0)
		{
		/* offsets must be affine in the tail index; otherwise fall back to the generic kernel */
		int m, j, ok = 1;
		hb_tail_tab *tb = &t->tab;
		tb->len = tail_len;
		for(m=0; m<tail_len && ok; m++)
			{
			const hb_tnode *a = &t->tn[t->seg_nodes[t->seg_start[n_top]+m]];
			const hb_tnode *b = n_tails>1 ? &t->tn[t->seg_nodes[t->seg_start[n_top+1]+m]] : a;
			int gen = HB_EVEN(HB_TRI(a->nu+a->nx)+2*(a->nu+a->nx));
			if(t->seg_start[n_top+1]-t->seg_start[n_top]!=tail_len) { ok = 0; break; }
			tb->posB[m] = a->off_BAbt; tb->strB[m] = b->off_BAbt - a->off_BAbt; tb->posQ[m] = a->off_RSQ;
			tb->posU[m] = a->off_ux; tb->strU[m] = b->off_ux - a->off_ux;
			tb->posP[m] = a->off_pi; tb->strP[m] = b->off_pi - a->off_pi;
			tb->posL[m] = a->off_L; tb->strL[m] = b->off_L - a->off_L; tb->posI[m] = a->off_L + (m==0 ? gen : 0);
			for(j=0; j<n_tails && ok; j++)
				{
				const hb_tnode *c = &t->tn[t->seg_nodes[t->seg_start[n_top+j]+m]];
				if(t->seg_start[n_top+j+1]-t->seg_start[n_top+j]!=tail_len) { ok = 0; break; }
				ok = c->off_BAbt==tb->posB[m]+j*tb->strB[m] && c->off_RSQ==tb->posQ[m]+j*tb->strB[m] && c->off_ux==tb->posU[m]+j*tb->strU[m]
					&& c->off_pi==tb->posP[m]+j*tb->strP[m] && c->off_L==tb->posL[m]+j*tb->strL[m];
				}
			}
		if(!ok) t->tail_fast_id = -1;          /* not affine: the generic tail kernel serves this tree */
		}
	t->dims.Nn = Nn; t->dims.nzM = nzM; t->dims.nxM = nxM; t->dims.n_seg = t->n_seg;
	t->dims.in_stride = o_in; t->dims.ux_stride = HB_EVEN(o_ux); t->dims.pi_stride = HB_EVEN(o_pi); t->dims.L_stride = o_L;
	t->sms = 148;
	t->idxb = calloc(t->nbtot+1, sizeof(int)); t->c_ux = calloc(t->nbtot+1, sizeof(int));
	for(n=0; n<Nn; n++)
		for(j=0; j<t->tn[n].nb; j++)
			{
			int id = idxb[n][j];
			if(id<0 || id>=t->tn[n].nu+t->tn[n].nx) { fprintf(stderr, "hpmpc_b200: tree: node %d: idxb[%d] = %d out of range\n", n, j, id); return -2; }
			t->idxb[t->tn[n].off_c+j] = id; t->c_ux[t->tn[n].off_c+j] = t->tn[n].off_ux + id;
			}
	memset(&t->ipm_dims, 0, sizeof(t->ipm_dims));
	t->ipm_dims.N = Nn-1; t->ipm_dims.nzM = nzM; t->ipm_dims.nxM = nxM; t->ipm_dims.nbtot = t->nbtot;
	t->ipm_dims.in_stride = t->dims.in_stride; t->ipm_dims.ux_stride = t->dims.ux_stride; t->ipm_dims.pi_stride = t->dims.pi_stride;
	t->ipm_dims.L_stride = HB_EVEN(t->dims.L_stride);
	if(device>=0)
		{
		CK(cudaSetDevice(device));
		hb_tnode *d_tn; int *d_ss, *d_sn;
		CK(cudaMalloc((void**)&d_tn, Nn*sizeof(hb_tnode)));
		CK(cudaMalloc((void**)&d_ss, (t->n_seg+1)*sizeof(int)));
		CK(cudaMalloc((void**)&d_sn, Nn*sizeof(int)));
		CK(cudaMemcpy(d_tn, t->tn, Nn*sizeof(hb_tnode), cudaMemcpyHostToDevice));
		CK(cudaMemcpy(d_ss, t->seg_start, (t->n_seg+1)*sizeof(int), cudaMemcpyHostToDevice));
		CK(cudaMemcpy(d_sn, t->seg_nodes, Nn*sizeof(int), cudaMemcpyHostToDevice));
		t->dims.tn = d_tn; t->dims.seg_start = d_ss; t->dims.seg_nodes = d_sn;
		{
		int *d_idxb, *d_cux;
		CK(cudaMalloc((void**)&d_idxb, (t->nbtot+1)*sizeof(int)));
		CK(cudaMalloc((void**)&d_cux, (t->nbtot+1)*sizeof(int)));
		CK(cudaMemcpy(d_idxb, t->idxb, (t->nbtot+1)*sizeof(int), cudaMemcpyHostToDevice));
		CK(cudaMemcpy(d_cux, t->c_ux, (t->nbtot+1)*sizeof(int), cudaMemcpyHostToDevice));
		t->ipm_dims.idxb = d_idxb; t->ipm_dims.c_ux = d_cux; t->ipm_dims.tn = d_tn;
		}
		if(t->nbtot>0)
			{
			/* where the right-hand sides and the bounded diagonal entries sit in a packed block */
			int n_ux = (int)o_ux, n_pi = (int)o_pi, i;
			int *g_ux = calloc(n_ux+1, sizeof(int)), *b_pi = calloc(n_pi+1, sizeof(int));
			int *c_diag = calloc(t->nbtot+1, sizeof(int)), *c_grad = calloc(t->nbtot+1, sizeof(int));
			for(n=0; n<Nn; n++)
				{
				const hb_tnode *s = &t->tn[n];
				int nux = s->nu+s->nx;
				for(i=0; i<nux; i++) g_ux[s->off_ux+i] = s->off_RSQ + HB_TRI(nux) + i;
				if(n>0) { int nuxd = t->tn[s->dad].nu + t->tn[s->dad].nx; for(i=0; i<s->nx; i++) b_pi[s->off_pi+i] = s->off_BAbt + nuxd*s->nx + i; }
				for(j=0; j<s->nb; j++)
					{
					int id = t->idxb[s->off_c+j];
					c_diag[s->off_c+j] = s->off_RSQ + HB_TRI(id) + id; c_grad[s->off_c+j] = s->off_RSQ + HB_TRI(nux) + id;
					}
				}
			int *dv[4]; const int *hv[4] = { g_ux, b_pi, c_diag, c_grad }; const int len[4] = { n_ux+1, n_pi+1, t->nbtot+1, t->nbtot+1 };
			for(i=0; i<4; i++)
				{
				CK(cudaMalloc((void**)&dv[i], len[i]*sizeof(int)));
				CK(cudaMemcpy(dv[i], hv[i], len[i]*sizeof(int), cudaMemcpyHostToDevice));
				}
			t->maps.g_ux = dv[0]; t->maps.b_pi = dv[1]; t->maps.c_diag = dv[2]; t->maps.c_grad = dv[3];
			t->maps.n_ux = n_ux; t->maps.n_pi = n_pi;
			free(g_ux); free(b_pi); free(c_diag); free(c_grad);
			}
		t->sms = hb_device_sm_count(device);
		if(t->sms<=0) return -1;
		}
	*out = t;
	return 0;
	}

void hpmpc_b200_tree_destroy(hpmpc_b200_tree *t)
	{
	if(!t) return;
	if(t->device>=0)
		{
		cudaSetDevice(t->device);
		cudaFree((void*)t->dims.tn); cudaFree((void*)t->dims.seg_start); cudaFree((void*)t->dims.seg_nodes);
		cudaFree((void*)t->ipm_dims.idxb); cudaFree((void*)t->ipm_dims.c_ux); cudaFree(t->ipm_ws); cudaFree(t->ipm_counter);
		cudaFree((void*)t->maps.g_ux); cudaFree((void*)t->maps.b_pi); cudaFree((void*)t->maps.c_diag); cudaFree((void*)t->maps.c_grad);
		cudaFree(t->f_in_mod); cudaFree(t->f_dux); cudaFree(t->f_dpi); cudaFree(t->f_L); cudaFree(t->f_ws); cudaFree(t->f_state);
		cudaFree(t->f_nact); if(t->h_nact) cudaFreeHost(t->h_nact); cudaFree(t->trs_ws);
		cudaFree(t->mg_send); cudaFree(t->mg_recv);
		if(t->busy_made) cudaEventDestroy(t->ev_busy);
		}
	free(t->idxb); free(t->c_ux);
	free(t->tn); free(t->stage); free(t->seg_start); free(t->seg_nodes); free(t->lvl_seg); free(t->slot); free(t->tail_root); free(t);
	}

void hpmpc_b200_tree_sizes_get(const hpmpc_b200_tree *t, hpmpc_b200_tree_sizes *o)
	{
	o->in_stride = t->dims.in_stride; o->ux_stride = t->dims.ux_stride; o->pi_stride = t->dims.pi_stride; o->L_stride = t->dims.L_stride;
	o->Nn = t->Nn; o->nzM = t->dims.nzM; o->nxM = t->dims.nxM;
	o->n_tails = t->n_tails; o->n_top_nodes = t->n_top; o->cut_stage = t->cut_stage;
	o->n_shard_nodes = t->cut_stage>0 ? t->lvl_seg[t->cut_stage] - t->lvl_seg[t->cut_stage-1] : 0;
	}

void hpmpc_b200_tree_node_offsets(const hpmpc_b200_tree *t, int n, int *off_BAbt, int *off_RSQ, int *off_ux, int *off_pi, int *off_L)
	{
	const hb_tnode *s = &t->tn[n];
	if(off_BAbt) *off_BAbt = s->off_BAbt;
	if(off_RSQ) *off_RSQ = s->off_RSQ;
	if(off_ux) *off_ux = s->off_ux;
	if(off_pi) *off_pi = s->off_pi;
	if(off_L) *off_L = s->off_L;
	}

void hpmpc_b200_tree_tail_root(const hpmpc_b200_tree *t, int tail, int *node, int *off_L, int *len_L)
	{
	int n = t->tail_root[tail];
	if(node) *node = n;
	if(off_L) *off_L = t->tn[n].off_L;
	if(len_L) *len_L = t->slot[n];           /* the whole slot: generic block and, when present, the stash image */
	}

void hpmpc_b200_tree_shard_node(const hpmpc_b200_tree *t, int k, int *node, int *off_L, int *len_L, int *tail_lo, int *tail_hi)
	{
	int n = t->seg_nodes[t->seg_start[t->lvl_seg[t->cut_stage-1]+k]];
	if(node) *node = n;
	if(off_L) *off_L = t->tn[n].off_L;
	if(len_L) *len_L = t->slot[n];           /* the whole slot: generic block and, when present, the stash image */
	/* kids of a node are contiguous and tail roots are numbered in BFS order */
	if(tail_lo) *tail_lo = t->tn[n].first_kid - t->tail_root[0];
	if(tail_hi) *tail_hi = t->tn[n].first_kid + t->tn[n].nkids - t->tail_root[0];
	}

int hpmpc_b200_tree_pack_instance(const hpmpc_b200_tree *t, double *const *A, double *const *B, double *const *b,
		double *const *Q, double *const *S, double *const *R, double *const *q, double *const *r, double *blk)
	{
	int n, i, j;
	memset(blk, 0, sizeof(double)*t->dims.in_stride);
	for(n=0; n<t->Nn; n++)
		{
		const hb_tnode *s = &t->tn[n];
		int nx = s->nx, nu = s->nu, nux = nx+nu;
		if(n>0)
			{
			const hb_tnode *d = &t->tn[s->dad];
			int nxd = d->nx, nud = d->nu, nuxd = nxd+nud;
			double *M = blk + s->off_BAbt;                       /* (nuxd+1) x nx row-major */
			for(i=0; i<nud; i++) for(j=0; j<nx; j++) M[i*nx+j] = B[n][j+(size_t)nx*i];
			for(i=0; i<nxd; i++) for(j=0; j<nx; j++) M[(nud+i)*nx+j] = A[n][j+(size_t)nx*i];
			for(j=0; j<nx; j++) M[nuxd*nx+j] = b[n][j];
			}
		double *H = blk + s->off_RSQ;
		for(i=0; i<nu; i++) for(j=0; j<=i; j++) H[HB_TRI(i)+j] = R[n][i+(size_t)nu*j];
		for(i=0; i<nx; i++)
			{
			for(j=0; j<nu; j++) H[HB_TRI(nu+i)+j] = S[n][j+(size_t)nu*i];
			for(j=0; j<=i; j++) H[HB_TRI(nu+i)+nu+j] = Q[n][i+(size_t)nx*j];
			}
		for(j=0; j<nu; j++) H[HB_TRI(nux)+j] = r[n][j];
		for(j=0; j<nx; j++) H[HB_TRI(nux)+nu+j] = q[n][j];
		}
	return 0;
	}

int hpmpc_b200_tree_pack_bounds(const hpmpc_b200_tree *t, double *const *lb, double *const *ub, double *blk)
	{
	int n, j;
	for(n=0; n<t->Nn; n++)
		{
		const hb_tnode *s = &t->tn[n];
		for(j=0; j<s->nb; j++) { blk[s->off_d+j] = lb[n][j]; blk[s->off_d+s->nb+j] = ub[n][j]; }
		}
	return 0;
	}

long long hpmpc_b200_tree_launch_count(const hpmpc_b200_tree *t) { return t->n_launches; }

void hpmpc_b200_tree_bound_offsets(const hpmpc_b200_tree *t, int n, int *nb, int *off_c, int *off_d)
	{
	if(nb) *nb = t->tn[n].nb;
	if(off_c) *off_c = t->tn[n].off_c;
	if(off_d) *off_d = t->tn[n].off_d;
	}

static void launch_shape(const hpmpc_b200_tree *t, long long items, int *grid, int *warps)
	{
	int w = 4;
	int smem_warp = hb_smem_bytes_per_warp_sz(t->dims.nzM, t->dims.nxM);
	while(w>1 && w*smem_warp>100*1024) w--;
	int per_sm = (220*1024)/(w*smem_warp+1024); if(per_sm<1) per_sm = 1; if(per_sm*w>16) per_sm = 16/w>0 ? 16/w : 1;
	long long need = (items + w - 1)/w, cap = (long long)t->sms*per_sm;
	*warps = w; *grid = (int)(need<cap ? (need<1 ? 1 : need) : cap);
	}

/* launch shape of the size-specialised kernels: `items` work items, t_ipw of them per warp */
static void fast_shape(const hpmpc_b200_tree *t, long long items, int *grid, int *warps)
	{
	long long groups = (items + t->t_ipw - 1)/t->t_ipw;
	int w = 8;
	while(w>1 && w*t->t_smem_warp>113*1024) w--;
	int per_sm = (228*1024)/(w*t->t_smem_warp+1024); if(per_sm<1) per_sm = 1; if(per_sm*w>16) per_sm = 16/w;
	long long need = (groups + w - 1)/w, cap = (long long)t->sms*per_sm;
	*warps = w; *grid = (int)(need<cap ? (need<1 ? 1 : need) : cap);
	}

int hpmpc_b200_d_tree_back_ric_rec_sv_phase(hpmpc_b200_tree *t, long long n_trees, int phase, int tail_lo, int tail_hi,
		const double *d_in, double *d_ux, double *d_pi, double *d_L, void *stream)
	{
	if(n_trees<=0) return 0;
	if(t->device<0) { fprintf(stderr, "hpmpc_b200: host-only tree handle cannot solve; there is no CPU fallback\n"); return -4; }
	if(phase==0 || phase==2) { if(tail_lo<0 || tail_hi>t->n_tails || tail_lo>tail_hi) return -2; }
	if(phase==3 || phase==5) { if(t->cut_stage<1 || tail_lo<0 || tail_lo>tail_hi || tail_hi>t->lvl_seg[t->cut_stage]-t->lvl_seg[t->cut_stage-1]) return -2; }
	CK(cudaSetDevice(t->device));
	int grid, warps;
	if(phase==1 || phase==3 || phase==4 || phase==5)
		{
		/* the top, level by level: all nodes of a level are independent (one warp per (tree, node)).  Phases 3 / 5 restrict
		 * the deepest top level to the subtree roots [lo, hi); phase 4 is everything above that level. */
		int lv, rc;
		const int deep = t->cut_stage-1;
		if(t->n_top==0) return 0;
		for(lv=deep; lv>=0; lv--)
			{
			int a = t->lvl_seg[lv], b = t->lvl_seg[lv+1];
			if(phase==5 || (phase==4 && lv==deep) || (phase==3 && lv<deep)) continue;
			if(phase==3) { b = a + tail_hi; a = a + tail_lo; }
			if(b<=a) continue;
			if(t->top_fast)
				{
				fast_shape(t, n_trees*(b-a), &grid, &warps);
				t->n_launches++;
				if((rc = hb_launch_top(t->tail_fast_id, &t->dims, n_trees, d_in, d_ux, d_pi, d_L, 0, a, b, lv==0, grid, warps, stream, t->skip_state))) return rc;
				continue;
				}
			launch_shape(t, n_trees*(b-a), &grid, &warps);
			t->n_launches++;
			if((rc = hb_launch_tree(&t->dims, n_trees, d_in, d_ux, d_pi, d_L, 0, a, b, grid, warps, stream, t->skip_state))) return rc;
			}
		for(lv=0; lv<=deep; lv++)
			{
			int a = t->lvl_seg[lv], b = t->lvl_seg[lv+1];
			if(phase==3 || (phase==4 && lv==deep) || (phase==5 && lv<deep)) continue;
			if(phase==5) { b = a + tail_hi; a = a + tail_lo; }
			if(b<=a) continue;
			if(t->top_fast)
				{
				fast_shape(t, n_trees*(b-a), &grid, &warps);
				t->n_launches++;
				if((rc = hb_launch_top(t->tail_fast_id, &t->dims, n_trees, d_in, d_ux, d_pi, d_L, 1, a, b, lv==0, grid, warps, stream, t->skip_state))) return rc;
				continue;
				}
			launch_shape(t, n_trees*(b-a), &grid, &warps);
			t->n_launches++;
			if((rc = hb_launch_tree(&t->dims, n_trees, d_in, d_ux, d_pi, d_L, 1, a, b, grid, warps, stream, t->skip_state))) return rc;
			}
		return 0;
		}
	if(t->tail_fast_id>=0)
		{
		fast_shape(t, n_trees*(tail_hi-tail_lo), &grid, &warps);
		t->n_launches++;
		return hb_launch_tail(t->tail_fast_id, &t->dims, &t->tab, n_trees, d_in, d_ux, d_pi, d_L, phase==0 ? 0 : 1, tail_lo, tail_hi, grid, warps, stream, t->skip_state);
		}
	launch_shape(t, n_trees*(tail_hi-tail_lo), &grid, &warps);
	t->n_launches++;
	return hb_launch_tree(&t->dims, n_trees, d_in, d_ux, d_pi, d_L, phase==0 ? 0 : 1, t->n_top+tail_lo, t->n_top+tail_hi, grid, warps, stream, t->skip_state);
	}

int hpmpc_b200_d_tree_back_ric_rec_sv_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in,
		double *d_ux, double *d_pi, double *d_L, void *stream)
	{
	int rc;
	if((rc = hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 0, 0, t->n_tails, d_in, d_ux, d_pi, d_L, stream))) return rc;
	if((rc = hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 1, 0, 0, d_in, d_ux, d_pi, d_L, stream))) return rc;
	return hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 2, 0, t->n_tails, d_in, d_ux, d_pi, d_L, stream);
	}

/* multi-kernel IPM: per-tree state machine between the solves (hb_tipm_step_kernel), the solves by the size-specialised tree
 * Riccati on a private copy of the blocks.  Blocking: the host reads the number of trees still iterating after every step. */
static int tree_ipm_multi(hpmpc_b200_tree *t, long long n_trees, const double *d_in, int k_max, double mu0, double mu_tol, double alpha_min,
		int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info, void *stream)
	{
	cudaStream_t st = (cudaStream_t)stream;
	int rc, round;
	if(t->f_trees<n_trees)
		{
		CK(cudaStreamSynchronize(st));
		cudaFree(t->f_in_mod); cudaFree(t->f_dux); cudaFree(t->f_dpi); cudaFree(t->f_L); cudaFree(t->f_ws); cudaFree(t->f_state);
		t->f_in_mod = t->f_dux = t->f_dpi = t->f_L = t->f_ws = t->f_state = NULL; t->f_trees = 0;
		t->f_ws_stride = HB_EVEN(hb_tipm_work_doubles(&t->ipm_dims));
		CK(cudaMalloc((void**)&t->f_in_mod, sizeof(double)*(size_t)n_trees*t->dims.in_stride));
		CK(cudaMalloc((void**)&t->f_dux, sizeof(double)*(size_t)n_trees*t->dims.ux_stride));
		CK(cudaMalloc((void**)&t->f_dpi, sizeof(double)*(size_t)n_trees*t->dims.pi_stride));
		CK(cudaMalloc((void**)&t->f_L, sizeof(double)*(size_t)n_trees*t->dims.L_stride));
		CK(cudaMalloc((void**)&t->f_ws, sizeof(double)*(size_t)n_trees*t->f_ws_stride));
		CK(cudaMalloc((void**)&t->f_state, sizeof(double)*(size_t)(n_trees+1)*8));       /* + the gate record */
		/* stride padding of the step vectors is never written by the solver kernels; keep it at zero so that the element-wise
		 * updates over whole strides leave the padding of ux / pi at zero too */
		CK(cudaMemsetAsync(t->f_dux, 0, sizeof(double)*(size_t)n_trees*t->dims.ux_stride, st));
		CK(cudaMemsetAsync(t->f_dpi, 0, sizeof(double)*(size_t)n_trees*t->dims.pi_stride, st));
		if(t->f_nact==NULL) { CK(cudaMalloc((void**)&t->f_nact, 2*sizeof(int))); CK(cudaMallocHost((void**)&t->h_nact, 2*sizeof(int))); }
		t->f_trees = n_trees;
		}
	CK(cudaMemcpyAsync(t->f_in_mod, d_in, sizeof(double)*(size_t)n_trees*t->dims.in_stride, cudaMemcpyDeviceToDevice, st));
	CK(cudaMemsetAsync(t->f_state, 0, sizeof(double)*(size_t)n_trees*8, st));
	if((rc = hb_launch_tipm_gate(t->f_nact, t->f_state + (size_t)n_trees*8, (int)(n_trees>2000000000LL ? 2000000000LL : n_trees), stream))) return rc;
	/* A fixed launch schedule, nothing read back: 2 k_max + 2 rounds of [step, residuals, step, tree Riccati]; every kernel looks at
	 * the per-tree state record and passes over trees that are not in a state it serves -- finished trees included, which the
	 * Riccati kernels skip group-wise -- so rounds behind the last active tree cost a few microseconds each.  HPMPC_B200_TREE_IPM_SYNC=1
	 * restores the adaptive loop that reads the number of unfinished trees after every step (blocking, stops at the last round). */
	const int adaptive = getenv("HPMPC_B200_TREE_IPM_SYNC")!=NULL;
	for(round=0; round<2*k_max+2; round++)
		{
		int part;
		for(part=0; part<2; part++)
			{
			CK(cudaMemsetAsync(t->f_nact, 0, 2*sizeof(int), st));
			t->n_launches++;
			if((rc = hb_launch_tipm_step(&t->ipm_dims, &t->maps, part, n_trees, d_in, t->f_in_mod, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi,
					t->f_dux, t->f_dpi, d_lam, d_t, d_info, t->f_ws, t->f_ws_stride, t->f_state, t->f_nact, stream))) return rc;
			if(adaptive)
				{
				CK(cudaMemcpyAsync(t->h_nact, t->f_nact, 2*sizeof(int), cudaMemcpyDeviceToHost, st));
				CK(cudaStreamSynchronize(st));
				if(part==1 || t->h_nact[1]==0) break;
				}
			else if(part==1) break;
			/* trees that want their residuals (phase switch or end of a phase-2 iteration): node-parallel, then part 1 */
			t->n_launches++;
			if((rc = hb_launch_tipm_res(&t->ipm_dims, n_trees, d_in, d_ux, d_pi, t->f_dux, t->f_dpi, t->f_ws, t->f_ws_stride, t->f_state, t->sms, stream))) return rc;
			}
		if(adaptive && t->h_nact[0]==0) return 0;
		/* the gate for everything that follows: the number of trees the last step left unfinished */
		if((rc = hb_launch_tipm_gate(t->f_nact, t->f_state + (size_t)n_trees*8, 0, stream))) return rc;
		t->skip_state = t->f_state;
		rc = hpmpc_b200_d_tree_back_ric_rec_sv_batch(t, n_trees, t->f_in_mod, t->f_dux, t->f_dpi, t->f_L, stream);
		t->skip_state = NULL;
		if(rc) return rc;
		}
	if(!adaptive) return 0;
	fprintf(stderr, "hpmpc_b200: tree IPM: state machine did not terminate\n");
	return -5;
	}

/* box-constrained IPM over a batch of trees: the whole iteration runs in one kernel, one warp per tree taking trees from a
 * queue (hb_ipm_kernel with the tree sweeps of ric_tree_ipm.cuh); per-warp work slots live in the handle */
static int tree_ipm_impl(hpmpc_b200_tree *t, long long n_trees, const double *d_in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info,
		void *stream);
int hpmpc_b200_d_tree_ip2_res_mpc_hard_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info,
		void *stream)
	{
	if(n_trees<=0) return 0;
	if(t->device<0) { fprintf(stderr, "hpmpc_b200: host-only tree handle cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(t->device));
	if(tcall_begin(t, stream)) return -1;
	return tcall_end(t, stream, tree_ipm_impl(t, n_trees, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info, stream));
	}
static int tree_ipm_impl(hpmpc_b200_tree *t, long long n_trees, const double *d_in, int k_max, double mu0,
		double mu_tol, double alpha_min, int warm_start, double *d_ux, double *d_pi, double *d_lam, double *d_t, double *d_info,
		void *stream)
	{
	if(n_trees<=0) return 0;
	if(t->device<0) { fprintf(stderr, "hpmpc_b200: host-only tree handle cannot solve; there is no CPU fallback\n"); return -4; }
	if(k_max<1) return -2;
	CK(cudaSetDevice(t->device));
	/* uniform tails with a size-specialised Riccati: iterate around the fast tree solver (HPMPC_B200_TREE_IPM_FUSED=1 keeps the
	 * single-kernel path below, which also serves every other tree) */
	if(t->tail_fast_id>=0 && t->nbtot>0 && getenv("HPMPC_B200_TREE_IPM_FUSED")==NULL)
		{
		/* the driver keeps a private copy of the blocks, the factor images and the work vectors of every tree it iterates on
		 * (about 8 MB per tree at BASELINE config 5): large batches go through in chunks whose buffers stay below ~24 GB */
		const double per_tree = 8.0*((double)t->dims.in_stride + t->dims.L_stride + t->dims.ux_stride + t->dims.pi_stride
				+ (double)hb_tipm_work_doubles(&t->ipm_dims) + 8);
		long long chunk = (long long)(24e9/per_tree), done;
		{ const char *e = getenv("HPMPC_B200_TREE_IPM_CHUNK"); if(e && atoll(e)>0) chunk = atoll(e); }
		if(chunk<1) chunk = 1;
		const long long info_len = HB_IPM_INFO_HEAD + 5*(long long)k_max;
		for(done=0; done<n_trees; done+=chunk)
			{
			long long m = n_trees-done<chunk ? n_trees-done : chunk;
			int rc = tree_ipm_multi(t, m, d_in + done*t->dims.in_stride, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux + done*t->dims.ux_stride,
					d_pi + done*t->dims.pi_stride, d_lam + done*2*(long long)t->nbtot, d_t + done*2*(long long)t->nbtot, d_info + done*info_len, stream);
			if(rc) return rc;
			}
		return 0;
		}
	int grid, warps;
	launch_shape(t, n_trees, &grid, &warps);
	const long long stride = HB_EVEN(hb_ipm_work_doubles(&t->ipm_dims));
	if(t->ipm_ws==NULL || t->ipm_slots<grid*warps || t->ipm_ws_stride!=stride)
		{
		CK(cudaStreamSynchronize((cudaStream_t)stream));
		cudaFree(t->ipm_ws); t->ipm_ws = NULL;
		t->ipm_slots = grid*warps; t->ipm_ws_stride = stride;
		CK(cudaMalloc((void**)&t->ipm_ws, sizeof(double)*(size_t)stride*t->ipm_slots));
		if(t->ipm_counter==NULL) CK(cudaMalloc((void**)&t->ipm_counter, sizeof(int)));
		}
	t->n_launches++;
	return hb_launch_ipm(&t->ipm_dims, n_trees, d_in, k_max, mu0, mu_tol, alpha_min, warm_start, d_ux, d_pi, d_lam, d_t, d_info,
			t->ipm_ws, stride, t->ipm_slots, grid, warps, t->ipm_counter, HB_IPM_TREE, stream);
	}

/* factor only: d_L receives the node factors (generic packed blocks at hpmpc_b200_tree_node_offsets' off_L) */
int hpmpc_b200_d_tree_back_ric_rec_trf_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in, double *d_L, void *stream)
	{
	if(n_trees<=0) return 0;
	if(t->device<0) { fprintf(stderr, "hpmpc_b200: host-only tree handle cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(t->device));
	int grid, warps;
	launch_shape(t, n_trees, &grid, &warps);
	t->n_launches++;
	return hb_launch_tree_trf_trs(&t->ipm_dims, n_trees, d_in, d_L, NULL, NULL, NULL, grid*warps, 0, grid, warps, stream);
	}

/* solve with the factors of a previous _trf_batch for the b and [r q] held in d_in (new right-hand sides, same matrices) */
int hpmpc_b200_d_tree_back_ric_rec_trs_batch(hpmpc_b200_tree *t, long long n_trees, const double *d_in, const double *d_L,
		double *d_ux, double *d_pi, void *stream)
	{
	if(n_trees<=0) return 0;
	if(t->device<0) { fprintf(stderr, "hpmpc_b200: host-only tree handle cannot solve; there is no CPU fallback\n"); return -4; }
	CK(cudaSetDevice(t->device));
	int grid, warps;
	launch_shape(t, n_trees, &grid, &warps);
	if(tcall_begin(t, stream)) return -1;
	if(t->trs_ws==NULL || t->trs_slots<grid*warps)
		{
		CK(cudaStreamSynchronize((cudaStream_t)stream));
		cudaFree(t->trs_ws); t->trs_ws = NULL;
		t->trs_slots = grid*warps;
		CK(cudaMalloc((void**)&t->trs_ws, sizeof(double)*(size_t)t->trs_slots*(t->dims.ux_stride + 2*t->dims.pi_stride)));
		}
	t->n_launches++;
	return tcall_end(t, stream, hb_launch_tree_trf_trs(&t->ipm_dims, n_trees, d_in, (double*)d_L, d_ux, d_pi, t->trs_ws, t->trs_slots, 1, grid, warps, stream));
	}

/* ------------------------------------------------------------------------------------------------ */
/* Scenario tree on several GPUs: subtrees shard over the ranks, ONE exchange per solve -- an all-gather    */
/* of the subtree roots' factor blocks over NVLink with NCCL -- issued by this library on the caller's stream.                    */
/* Reference recursion: lqcp_solvers/d_tree_back_ric_rec_libstr.c:524-583 (the node loop the phases below cut at the level of    */
/* the subtree roots).  NCCL is reached through dlopen("libnccl.so.2"): the library has no link-time dependency on it, and in a   */
/* process that already loaded NCCL (torch) the same copy is used.                                                               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { char internal[128]; } hb_nccl_uid;
typedef void *hb_nccl_comm;
static struct
	{
	void *lib;
	int (*GetUniqueId)(hb_nccl_uid *);
	int (*CommInitRank)(hb_nccl_comm *, int, hb_nccl_uid, int);
	int (*CommDestroy)(hb_nccl_comm);
	int (*AllGather)(const void *, void *, size_t, int, hb_nccl_comm, cudaStream_t);
	const char *(*GetErrorString)(int);
	} NC;

static int nccl_load(void)
	{
	if(NC.lib) return 0;
	const char *names[3] = { getenv("HPMPC_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
	int k;
	for(k=0; k<3 && !NC.lib; k++) if(names[k]) NC.lib = dlopen(names[k], RTLD_NOW|RTLD_GLOBAL);
	if(!NC.lib) { fprintf(stderr, "hpmpc_b200: NCCL not found (%s)\n", dlerror()); return -1; }
	NC.GetUniqueId = (int (*)(hb_nccl_uid*))dlsym(NC.lib, "ncclGetUniqueId");
	NC.CommInitRank = (int (*)(hb_nccl_comm*, int, hb_nccl_uid, int))dlsym(NC.lib, "ncclCommInitRank");
	NC.CommDestroy = (int (*)(hb_nccl_comm))dlsym(NC.lib, "ncclCommDestroy");
	NC.AllGather = (int (*)(const void*, void*, size_t, int, hb_nccl_comm, cudaStream_t))dlsym(NC.lib, "ncclAllGather");
	NC.GetErrorString = (const char *(*)(int))dlsym(NC.lib, "ncclGetErrorString");
	if(!NC.GetUniqueId || !NC.CommInitRank || !NC.CommDestroy || !NC.AllGather) { fprintf(stderr, "hpmpc_b200: NCCL symbols missing\n"); NC.lib = NULL; return -1; }
	return 0;
	}
#define NCK(x) do { int r_ = (x); if(r_!=0) { fprintf(stderr, "hpmpc_b200: NCCL error %s at %s:%d\n", NC.GetErrorString ? NC.GetErrorString(r_) : "?", __FILE__, __LINE__); return -1; } } while(0)

struct hpmpc_b200_comm { hb_nccl_comm comm; int world, rank, owned; };

int hpmpc_b200_comm_unique_id(void *id, int id_bytes)
	{
	if(id_bytes<(int)sizeof(hb_nccl_uid) || nccl_load()) return -1;
	NCK(NC.GetUniqueId((hb_nccl_uid*)id));
	return 0;
	}

int hpmpc_b200_comm_create(hpmpc_b200_comm **out, int world, int rank, const void *id, int device)
	{
	*out = NULL;
	if(nccl_load()) return -1;
	CK(cudaSetDevice(device));
	hpmpc_b200_comm *c = calloc(1, sizeof(*c));
	if(!c) return -1;
	hb_nccl_uid uid; memcpy(&uid, id, sizeof(uid));
	int r = NC.CommInitRank(&c->comm, world, uid, rank);
	if(r!=0) { fprintf(stderr, "hpmpc_b200: ncclCommInitRank failed: %s\n", NC.GetErrorString ? NC.GetErrorString(r) : "?"); free(c); return -1; }
	c->world = world; c->rank = rank; c->owned = 1;
	*out = c;
	return 0;
	}

/* an ncclComm_t the caller already has (it must come from the NCCL copy this process has loaded) */
int hpmpc_b200_comm_wrap(hpmpc_b200_comm **out, void *nccl_comm, int world, int rank)
	{
	*out = NULL;
	if(nccl_load()) return -1;
	hpmpc_b200_comm *c = calloc(1, sizeof(*c));
	if(!c) return -1;
	c->comm = nccl_comm; c->world = world; c->rank = rank; c->owned = 0;
	*out = c;
	return 0;
	}

void hpmpc_b200_comm_destroy(hpmpc_b200_comm *c)
	{
	if(!c) return;
	if(c->owned && NC.CommDestroy) NC.CommDestroy(c->comm);
	free(c);
	}

/* Every rank holds the data of all n_trees trees and calls this with the same arguments; on return every rank has ux / pi of ITS
 * subtrees' nodes and of the nodes above them (the latter computed redundantly on every rank).  Subtrees k*world/.. are split
 * evenly: n_shard_nodes must be a multiple of the communicator size.  Nothing blocks the host. */
int hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg(hpmpc_b200_tree *t, hpmpc_b200_comm *c, long long n_trees, const double *d_in,
		double *d_ux, double *d_pi, double *d_L, void *stream)
	{
	if(n_trees<=0) return 0;
	if(!c || c->world<=1) return hpmpc_b200_d_tree_back_ric_rec_sv_batch(t, n_trees, d_in, d_ux, d_pi, d_L, stream);
	if(t->device<0) { fprintf(stderr, "hpmpc_b200: host-only tree handle cannot solve; there is no CPU fallback\n"); return -4; }
	if(t->cut_stage<1) { fprintf(stderr, "hpmpc_b200: tree_mg: the tree has no level to shard below\n"); return -6; }
	const int ns = t->lvl_seg[t->cut_stage] - t->lvl_seg[t->cut_stage-1];
	if(ns%c->world!=0) { fprintf(stderr, "hpmpc_b200: tree_mg: %d subtrees do not split evenly over %d ranks\n", ns, c->world); return -6; }
	const int per = ns/c->world, lo = c->rank*per, hi = lo+per;
	int node, o0, ln, tlo, thi, k, rc;
	hpmpc_b200_tree_shard_node(t, 0, &node, &o0, &ln, NULL, NULL);
	for(k=0; k<ns; k++)
		{
		int ok, lk;
		hpmpc_b200_tree_shard_node(t, k, &node, &ok, &lk, NULL, NULL);
		if(ok!=o0+k*ln || lk!=ln) { fprintf(stderr, "hpmpc_b200: tree_mg: subtree-root factor blocks are not contiguous\n"); return -6; }
		}
	hpmpc_b200_tree_shard_node(t, lo, NULL, NULL, NULL, &tlo, NULL);
	hpmpc_b200_tree_shard_node(t, hi-1, NULL, NULL, NULL, NULL, &thi);
	CK(cudaSetDevice(t->device));
	cudaStream_t st = (cudaStream_t)stream;
	if(tcall_begin(t, stream)) return -1;
	const size_t seg = (size_t)per*ln;                         /* doubles per tree this rank contributes */
	const size_t send_b = sizeof(double)*seg*(size_t)n_trees, recv_b = send_b*(size_t)c->world;
	if(send_b>t->mg_send_bytes) { CK(cudaStreamSynchronize(st)); cudaFree(t->mg_send); t->mg_send = NULL; CK(cudaMalloc((void**)&t->mg_send, send_b)); t->mg_send_bytes = send_b; }
	if(recv_b>t->mg_recv_bytes) { CK(cudaStreamSynchronize(st)); cudaFree(t->mg_recv); t->mg_recv = NULL; CK(cudaMalloc((void**)&t->mg_recv, recv_b)); t->mg_recv_bytes = recv_b; }
	/* backward: own tails, own subtree roots */
	if((rc = hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 0, tlo, thi, d_in, d_ux, d_pi, d_L, stream))) return rc;
	if((rc = hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 3, lo, hi, d_in, d_ux, d_pi, d_L, stream))) return rc;
	/* the one exchange: this rank's root blocks of every tree -> contiguous, all-gather over NVLink, scatter back into the stashes */
	CK(cudaMemcpy2DAsync(t->mg_send, sizeof(double)*seg, d_L + o0 + (size_t)lo*ln, sizeof(double)*t->dims.L_stride, sizeof(double)*seg, (size_t)n_trees,
			cudaMemcpyDeviceToDevice, st));
	NCK(NC.AllGather(t->mg_send, t->mg_recv, seg*(size_t)n_trees, 8 /* ncclFloat64 */, c->comm, st));
	for(k=0; k<c->world; k++)
		{
		if(k==c->rank) continue;
		CK(cudaMemcpy2DAsync(d_L + o0 + (size_t)k*per*ln, sizeof(double)*t->dims.L_stride, t->mg_recv + (size_t)k*seg*(size_t)n_trees, sizeof(double)*seg,
				sizeof(double)*seg, (size_t)n_trees, cudaMemcpyDeviceToDevice, st));
		}
	/* the levels above the subtree roots (redundantly), then down the own subtrees */
	if((rc = hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 4, 0, 0, d_in, d_ux, d_pi, d_L, stream))) return rc;
	if((rc = hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 5, lo, hi, d_in, d_ux, d_pi, d_L, stream))) return rc;
	return tcall_end(t, stream, hpmpc_b200_d_tree_back_ric_rec_sv_phase(t, n_trees, 2, tlo, thi, d_in, d_ux, d_pi, d_L, stream));
	}
