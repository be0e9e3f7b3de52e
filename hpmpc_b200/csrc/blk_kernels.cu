/*
 * blk_kernels.cu -- launchers of the size-specialised Riccati kernels: register-blocked (ric_blk.cuh: sv, tree tails,
 * tree top) and one-row-per-lane (ric_fast.cuh).  A translation unit of its own so that it compiles in parallel.
 */
#include "launch_util.cuh"
#include "layout.h"
#include "ric_fast.cuh"
#include "ric_blk.cuh"
#ifndef HBK_DEFAULT_FLAGS
/* measured on B200 (profiles/r02_sv_traffic.txt): evict-first on the single-use streams, evict-last + discard on the stash cuts
 * the DRAM traffic of the sv kernel from 231 to 199 KB per solve (time unchanged: the kernel is shared-memory-pipe bound) */
#ifdef HBK_BULK_STASH
#define HBK_DEFAULT_FLAGS (HBK_F_IN_FIRST|HBK_F_RE_FIRST|HBK_F_KEEP|HBK_F_DISCARD)
#define HBK_DEFAULT_KEEP 1000
#else
#define HBK_DEFAULT_FLAGS (HBK_F_IN_FIRST|HBK_F_RE_FIRST)
#define HBK_DEFAULT_KEEP 0
#endif
#endif

/* ------------------------------------------------------------------------------------------------ */
/* size-specialised variants (ric_blk.cuh: register-blocked; ric_fast.cuh: one row per lane)          */
/* ------------------------------------------------------------------------------------------------ */
#ifndef HBK_V0_G
#define HBK_V0_G 8
#define HBK_V0_R 2
#endif
typedef hbk_cfg<12, 5, HBK_V0_G, HBK_V0_R> hbk_v0;   /* BASELINE config 2: four instances per warp, two rows per lane */
typedef hbk_cfg<8, 3, 4, 3> hbk_v1;    /* the reference's own test size (test_d_ip_hard.c): eight instances per warp */
typedef hbk_cfg<4, 2, 4, 2> hbk_v2;    /* eight instances per warp */
typedef hbf_cfg<24, 11, 32> hbf_v3;    /* BASELINE config 3 shape: one instance per warp, 4 column-owned rows */   /* (hbk_cfg<24,11,16,2> was tried: 1.7 KB of local-memory stack, 9x slower) */
typedef hbf_cfg<12, 5, 16> hbf_v0;     /* one-row-per-lane predecessors, kept for A/B runs (HPMPC_B200_FAST_GEN=1) */
typedef hbf_cfg<8, 3, 16> hbf_v1;
typedef hbf_cfg<4, 2, 8> hbf_v2;
#define HBF_NVAR 7
static const int hbf_shapes[HBF_NVAR][2] = { {12, 5}, {8, 3}, {4, 2}, {24, 11}, {12, 5}, {8, 3}, {4, 2} };

/* a pattern qualifies when x0 is eliminated (nx[0] = 0) and every other stage has the variant's (nx, nu) */
extern "C" int hb_fast_variant(int N, const int *nx, const int *nu)
	{
	const char *gen = getenv("HPMPC_B200_FAST_GEN");
	const int first = (gen!=NULL && gen[0]=='1') ? 3 : 0;
	for(int id=first; id<HBF_NVAR; id++)
		{
		int ok = (nx[0]==0) && N>=3;
		for(int n=0; n<N && ok; n++) ok = (nu[n]==hbf_shapes[id][1]) && (n==0 || nx[n]==hbf_shapes[id][0]);
		ok = ok && nx[N]==hbf_shapes[id][0];
		if(ok) return id;
		}
	return -1;
	}

template<class C> static void hbf_info(int N, int *ipw, int *smem_warp, long long *stash_per_inst)
	{
	*ipw = C::IPW; *smem_warp = (int)sizeof(double)*C::PER_WARP; *stash_per_inst = (long long)(N+1)*C::LBUF;
	}
template<class C> static void hbk_info(int N, int *ipw, int *smem_warp, long long *stash_per_inst)
	{
	*ipw = C::IPW; *smem_warp = (int)sizeof(double)*C::PER_WARP; *stash_per_inst = (long long)(N+1)*C::SBG;
	}

extern "C" int hb_fast_info(int id, int N, int *ipw, int *smem_warp, long long *stash_per_inst)
	{
	switch(id)
		{
		case 0: hbk_info<hbk_v0>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 1: hbk_info<hbk_v1>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 2: hbk_info<hbk_v2>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 3: hbf_info<hbf_v3>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 4: hbf_info<hbf_v0>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 5: hbf_info<hbf_v1>(N, ipw, smem_warp, stash_per_inst); return 0;
		case 6: hbf_info<hbf_v2>(N, ipw, smem_warp, stash_per_inst); return 0;
		}
	return -1;
	}

template<class C> static int hbf_launch(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, cudaStream_t st)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbf_ric_sv_kernel<C>, smem)) return -1;
	hbf_ric_sv_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_inst, in, ux, pi, stash);
	HB_CK(cudaGetLastError());
	return 0;
	}
/* The factor stash is scratch that every warp slot rewrites for each instance it solves (written in the backward sweep,
 * read back in the forward sweep).  A persisting-L2 access window over it keeps a fraction of its lines resident, so that
 * fraction of the stash traffic never reaches HBM (HPMPC_B200_L2_PERSIST=0 disables; value = hit ratio in percent). */
static int hb_stash_window(cudaLaunchAttribute *attr, const void *stash, size_t stash_bytes)
	{
	static int inited = 0, max_persist = 0, max_window = 0, pct = -1;
	if(!inited)
		{
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
		const char *e = getenv("HPMPC_B200_L2_PERSIST");
		pct = e ? atoi(e) : 0;          /* opt-in: measured neutral (16-40 MiB set aside) to harmful (79 MiB) on B200 */
		{ const char *m = getenv("HPMPC_B200_L2_MB"); if(m && atoi(m)>0 && ((size_t)atoi(m)<<20)<(size_t)max_persist) max_persist = atoi(m)<<20; }
		if(pct!=0 && max_persist>0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
		if(getenv("HPMPC_B200_VERBOSE")) fprintf(stderr, "hpmpc_b200: persisting L2 max %d MiB, window max %d MiB\n", max_persist>>20, max_window>>20);
		inited = 1;
		}
	if(pct==0 || max_persist<=0 || max_window<=0 || stash_bytes==0) return 0;
	size_t win = stash_bytes<(size_t)max_window ? stash_bytes : (size_t)max_window;
	double ratio = pct>0 ? pct/100.0 : 0.9*(double)max_persist/(double)win;
	if(ratio>1.0) ratio = 1.0;
	attr->id = cudaLaunchAttributeAccessPolicyWindow;
	attr->val.accessPolicyWindow.base_ptr = (void*)stash;
	attr->val.accessPolicyWindow.num_bytes = win;
	attr->val.accessPolicyWindow.hitRatio = (float)ratio;
	attr->val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	attr->val.accessPolicyWindow.missProp = getenv("HPMPC_B200_L2_MISS_NORMAL") ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
	return 1;
	}

template<class C> static int hbk_launch(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, cudaStream_t st)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbk_ric_sv_kernel<C>, smem)) return -1;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = dim3(grid); cfg.blockDim = dim3(warps*32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	cfg.attrs = attr;
	cfg.numAttrs = hb_stash_window(&attr[0], stash, sizeof(double)*(size_t)grid*warps*C::IPW*(size_t)(d->N+1)*C::SBG);
	/* L2 management of the stash (ric_blk.cuh: HBK_F_*); HPMPC_B200_SV_FLAGS / HPMPC_B200_SV_KEEP override the defaults for A/B runs */
	static int flags = -1, keep = -1;
	if(flags<0)
		{
		const char *e = getenv("HPMPC_B200_SV_FLAGS"), *k = getenv("HPMPC_B200_SV_KEEP");
		keep = k ? atoi(k) : HBK_DEFAULT_KEEP;
		flags = e ? atoi(e) : HBK_DEFAULT_FLAGS;
		}
	HB_CK(cudaLaunchKernelEx(&cfg, hbk_ric_sv_kernel<C>, *d, n_inst, in, ux, pi, stash, flags, keep));
	HB_CK(cudaGetLastError());
	return 0;
	}

/* traffic-equivalent probe of the sv kernel (ric_blk.cuh: hbk_traffic_kernel); same launch shape, same flags */
template<class C> static int hbk_traffic_launch(const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, cudaStream_t st)
	{
#ifdef HBK_BULK_STASH
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbk_traffic_kernel<C>, smem)) return -1;
	const char *e = getenv("HPMPC_B200_SV_FLAGS"), *k = getenv("HPMPC_B200_SV_KEEP");
	hbk_traffic_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_inst, in, ux, pi, stash, e ? atoi(e) : HBK_DEFAULT_FLAGS, k ? atoi(k) : HBK_DEFAULT_KEEP);
	HB_CK(cudaGetLastError());
	return 0;
#else
	return -2;
#endif
	}
extern "C" int hb_launch_sv_traffic(int id, const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, void *stream)
	{
	if(id==0) return hbk_traffic_launch<hbk_v0>(d, n_inst, in, ux, pi, stash, grid, warps, (cudaStream_t)stream);
	return -2;
	}

extern "C" int hb_launch_ric_sv_fast(int id, const hb_dims *d, long long n_inst, const double *in, double *ux, double *pi,
		double *stash, int grid, int warps, void *stream)
	{
	cudaStream_t st = (cudaStream_t)stream;
	switch(id)
		{
		case 0: return hbk_launch<hbk_v0>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 1: return hbk_launch<hbk_v1>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 2: return hbk_launch<hbk_v2>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 3: return hbf_launch<hbf_v3>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 4: return hbf_launch<hbf_v0>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 5: return hbf_launch<hbf_v1>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		case 6: return hbf_launch<hbf_v2>(d, n_inst, in, ux, pi, stash, grid, warps, st);
		}
	return -2;
	}

/* ------------------------------------------------------------------------------------------------ */
/* size-specialised tails of scenario trees                                                          */
/* ------------------------------------------------------------------------------------------------ */
extern "C" int hb_tail_variant(int nx, int nu)
	{
	for(int id=0; id<3; id++) if(hbf_shapes[id][0]==nx && hbf_shapes[id][1]==nu) return id;
	return -1;
	}

template<class C> static void hbk_tail_info_t(int *ipw, int *smem_warp, int *image_doubles)
	{ *ipw = C::IPW; *smem_warp = (int)sizeof(double)*C::PER_WARP; *image_doubles = C::SB; }

extern "C" int hb_tail_info(int id, int *ipw, int *smem_warp, int *image_doubles)
	{
	switch(id)
		{
		case 0: hbk_tail_info_t<hbk_v0>(ipw, smem_warp, image_doubles); return 0;
		case 1: hbk_tail_info_t<hbk_v1>(ipw, smem_warp, image_doubles); return 0;
		case 2: hbk_tail_info_t<hbk_v2>(ipw, smem_warp, image_doubles); return 0;
		}
	return -1;
	}

template<class C> static int hbk_tail_launch(const hb_tdims *d, const hb_tail_tab *tab, long long n_trees, const double *in, double *ux,
		double *pi, double *L, int mode, int tail_lo, int tail_hi, int grid, int warps, cudaStream_t st, const double *skip)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbk_tail_kernel<C>, smem)) return -1;
	hbk_tail_kernel<C><<<grid, warps*32, smem, st>>>(*tab, n_trees, d->in_stride, d->ux_stride, d->pi_stride, d->L_stride, in, ux, pi, L,
			mode, tail_lo, tail_hi, skip);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_tail(int id, const hb_tdims *d, const hb_tail_tab *tab, long long n_trees, const double *in, double *ux, double *pi,
		double *L, int mode, int tail_lo, int tail_hi, int grid, int warps, void *stream, const double *skip)
	{
	if(tail_hi<=tail_lo || n_trees<=0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	switch(id)
		{
		case 0: return hbk_tail_launch<hbk_v0>(d, tab, n_trees, in, ux, pi, L, mode, tail_lo, tail_hi, grid, warps, st, skip);
		case 1: return hbk_tail_launch<hbk_v1>(d, tab, n_trees, in, ux, pi, L, mode, tail_lo, tail_hi, grid, warps, st, skip);
		case 2: return hbk_tail_launch<hbk_v2>(d, tab, n_trees, in, ux, pi, L, mode, tail_lo, tail_hi, grid, warps, st, skip);
		}
	return -2;
	}

template<class C> static int hbk_top_launch(const hb_tdims *d, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int first, int grid, int warps, cudaStream_t st, const double *skip)
	{
	int smem = warps*(int)sizeof(double)*C::PER_WARP;
	if(hb_prep(hbk_top_kernel<C>, smem)) return -1;
	hbk_top_kernel<C><<<grid, warps*32, smem, st>>>(*d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, skip);
	HB_CK(cudaGetLastError());
	return 0;
	}

extern "C" int hb_launch_top(int id, const hb_tdims *d, long long n_trees, const double *in, double *ux, double *pi, double *L,
		int mode, int seg_lo, int seg_hi, int first, int grid, int warps, void *stream, const double *skip)
	{
	if(seg_hi<=seg_lo || n_trees<=0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	switch(id)
		{
		case 0: return hbk_top_launch<hbk_v0>(d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, grid, warps, st, skip);
		case 1: return hbk_top_launch<hbk_v1>(d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, grid, warps, st, skip);
		case 2: return hbk_top_launch<hbk_v2>(d, n_trees, in, ux, pi, L, mode, seg_lo, seg_hi, first, grid, warps, st, skip);
		}
	return -2;
	}

#ifdef HBF_TIMING
extern "C" int hb_debug_timing(long long *d_buf)
	{
	int zero = 0;
	static int gen = 1000;
	gen++;
	HB_CK(cudaMemcpyToSymbol(hbf_dbg_gen, &gen, sizeof(int)));
	HB_CK(cudaMemcpyToSymbol(hbf_dbg, &d_buf, sizeof(d_buf)));
	HB_CK(cudaMemcpyToSymbol(hbf_dbg_n, &zero, sizeof(int)));
	return 0;
	}
#endif

