/* launch_util.cuh -- error check and dynamic-shared-memory opt-in shared by the launcher translation units */
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HB_CK(x) do { cudaError_t e_ = (x); if(e_!=cudaSuccess) { fprintf(stderr, "hpmpc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

template<typename K>
static int hb_prep(K kernel, int smem)
	{
	if(smem>227*1024) { fprintf(stderr, "hpmpc_b200: stage too large for shared memory (%d bytes)\n", smem); return -1; }
	/* always the architectural maximum, not this launch's size: the attribute is per kernel and device, and concurrent host
	 * threads launching the same kernel with different CTA shapes must not lower it under each other */
	HB_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
	return 0;
	}

