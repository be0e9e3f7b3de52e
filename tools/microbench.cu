// microbench.cu -- shared-memory broadcast patterns / FP64 latencies on sm_100a (design input for the Riccati kernels)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if(e!=cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while(0)

// pattern: G lanes per group; group g reads from byte offset g*stride (+ i*step per iteration)
template<int W>  // W = 8 (LDS.64) or 16 (LDS.128)
__global__ void lds_kernel(int G, int stride, int step, int iters, long long *out, double *sink)
	{
	extern __shared__ __align__(16) unsigned char sm[];
	const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
	for(int i=threadIdx.x; i<48*1024/8; i+=blockDim.x) reinterpret_cast<double*>(sm)[i] = i;
	__syncthreads();
	const int g = lane/G;
	uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + g*stride + warp*0;
	double acc0 = 0, acc1 = 0;
	long long t0 = clock64();
	for(int it=0; it<iters; it++)
		{
		#pragma unroll
		for(int u=0; u<16; u++)
			{
			uint32_t a = base + ((u*step) & 8191);
			if(W==8)
				{ double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); acc0 += v; }
			else
				{ double v, w; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(w) : "r"(a)); acc0 += v; acc1 += w; }
			}
		}
	long long t1 = clock64();
	if(lane==0) out[blockIdx.x*(blockDim.x>>5)+warp] = t1-t0;
	if(acc0+acc1==123.456) sink[0] = acc0;
	}

// same but without the DADD dependency cost dominating: use integer xor of the raw bits
template<int W>
__global__ void lds_kernel2(int G, int stride, int step, int iters, long long *out, unsigned *sink)
	{
	extern __shared__ __align__(16) unsigned char sm[];
	const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
	for(int i=threadIdx.x; i<48*1024/4; i+=blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i;
	__syncthreads();
	const int g = lane/G;
	uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + g*stride;
	unsigned acc = 0;
	long long t0 = clock64();
	for(int it=0; it<iters; it++)
		{
		#pragma unroll
		for(int u=0; u<16; u++)
			{
			uint32_t a = base + ((u*step) & 8191);
			if(W==8)
				{ unsigned x, y; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); acc ^= x ^ y; }
			else
				{ unsigned x, y, z, w; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a)); acc ^= x ^ y ^ z ^ w; }
			}
		}
	long long t1 = clock64();
	if(lane==0) out[blockIdx.x*(blockDim.x>>5)+warp] = t1-t0;
	if(acc==0x12345) sink[0] = acc;
	}

__global__ void dfma_lat(int iters, long long *out, double *sink, double a, double b)
	{
	double x = a;
	long long t0 = clock64();
	for(int i=0; i<iters; i++)
		{
		#pragma unroll
		for(int u=0; u<16; u++) x = fma(x, b, a);
		}
	long long t1 = clock64();
	if(threadIdx.x==0) out[0] = t1-t0;
	if(x==123.456) sink[0] = x;
	}
template<int ILP>
__global__ void dfma_tp(int iters, long long *out, double *sink, double a, double b)
	{
	double x[ILP];
	#pragma unroll
	for(int k=0; k<ILP; k++) x[k] = a+k;
	__syncthreads();
	long long t0 = clock64();
	for(int i=0; i<iters; i++)
		{
		#pragma unroll
		for(int u=0; u<8; u++)
			#pragma unroll
			for(int k=0; k<ILP; k++) x[k] = fma(x[k], b, a);
		}
	long long t1 = clock64();
	if((threadIdx.x&31)==0) out[blockIdx.x*(blockDim.x>>5)+(threadIdx.x>>5)] = t1-t0;
	double s = 0;
	#pragma unroll
	for(int k=0; k<ILP; k++) s += x[k];
	if(s==123.456) sink[0] = s;
	}
__global__ void rsq_lat(int iters, long long *out, double *sink, double a)
	{
	double x = a;
	long long t0 = clock64();
	for(int i=0; i<iters; i++)
		{
		#pragma unroll
		for(int u=0; u<8; u++)
			{
			double y;
			asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
			x = y + a;
			}
		}
	long long t1 = clock64();
	if(threadIdx.x==0) out[0] = t1-t0;
	if(x==123.456) sink[0] = x;
	}
__global__ void shfl_lat(int iters, long long *out, double *sink, double a)
	{
	double x = a + threadIdx.x;
	long long t0 = clock64();
	for(int i=0; i<iters; i++)
		{
		#pragma unroll
		for(int u=0; u<8; u++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x+1)&31) + a;
		}
	long long t1 = clock64();
	if(threadIdx.x==0) out[0] = t1-t0;
	if(x==123.456) sink[0] = x;
	}
__global__ void lds_lat(int iters, long long *out, unsigned *sink)
	{
	__shared__ unsigned s[1024];
	for(int i=threadIdx.x; i<1024; i+=blockDim.x) s[i] = (i*7+3)&1023;
	__syncthreads();
	unsigned x = threadIdx.x;
	long long t0 = clock64();
	for(int i=0; i<iters; i++)
		{
		#pragma unroll
		for(int u=0; u<8; u++) x = s[x];
		}
	long long t1 = clock64();
	if(threadIdx.x==0) out[0] = t1-t0;
	if(x==0x12345) sink[0] = x;
	}
// STS + syncwarp + LDS round trip (the Cholesky column broadcast)
__global__ void sts_lds_lat(int iters, long long *out, double *sink, double a)
	{
	__shared__ double s[64];
	double x = a + threadIdx.x;
	long long t0 = clock64();
	for(int i=0; i<iters; i++)
		{
		#pragma unroll
		for(int u=0; u<8; u++)
			{
			s[threadIdx.x] = x;
			__syncwarp();
			x = s[(threadIdx.x+1)&31] + a;
			__syncwarp();
			}
		}
	long long t1 = clock64();
	if(threadIdx.x==0) out[0] = t1-t0;
	if(x==123.456) sink[0] = x;
	}

int main()
	{
	long long *d_out, h_out[4096]; double *d_sink;
	CK(cudaMalloc(&d_out, sizeof(h_out))); CK(cudaMalloc(&d_sink, 64));
	int iters = 2000;
	CK(cudaFuncSetAttribute(lds_kernel2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64*1024));
	CK(cudaFuncSetAttribute(lds_kernel2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64*1024));
	struct pat { const char *name; int W, G, stride, step; };
	pat pats[] = {
		{"LDS.64  uniform", 8, 32, 0, 8},
		{"LDS.128 uniform", 16, 32, 0, 16},
		{"LDS.64  2x16 private(stride 6960)", 8, 16, 6960, 8},
		{"LDS.128 2x16 private(stride 6960)", 16, 16, 6960, 16},
		{"LDS.128 2x16 private(stride 6928+16)", 16, 16, 6928+16, 16},
		{"LDS.64  4x8 private(stride 4112)", 8, 8, 4112, 8},
		{"LDS.128 4x8 private(stride 4112)", 16, 8, 4112, 16},
		{"LDS.64  8x4 private(stride 4112)", 8, 4, 4112, 8},
		{"LDS.128 8x4 private(stride 4112)", 16, 4, 4112, 16},
		{"LDS.64  8x4 private(stride 4104)", 8, 4, 4104, 8},
		{"LDS.64  8x4 interleaved(stride 8, step 64)", 8, 4, 8, 64},
		{"LDS.128 8x4 interleaved(stride 16, step 128)", 16, 4, 16, 128},
		{"LDS.64  16x2 interleaved(stride 8, step 128)", 8, 2, 8, 128},
		{"LDS.128 16x2 interleaved(stride 16, step 256)", 16, 2, 16, 256},
		{"LDS.64  32x1 contiguous", 8, 1, 8, 256},
		{"LDS.128 32x1 contiguous", 16, 1, 16, 512},
		{"LDS.64  4x8 interleaved(stride 8, step 32)", 8, 8, 8, 32},
		{"LDS.128 4x8 interleaved(stride 16, step 64)", 16, 8, 16, 64},
		{"LDS.128 2x16 interleaved(stride 16, step 32)", 16, 16, 16, 32},
		{"LDS.64  2x16 interleaved(stride 8, step 16)", 8, 16, 8, 16},
	};
	for(auto &p : pats)
		{
		for(int nw : {4, 16})
			{
			int blocks = 148;
			if(p.W==8) lds_kernel2<8><<<blocks, nw*32, 60*1024>>>(p.G, p.stride, p.step, iters, d_out, (unsigned*)d_sink);
			else       lds_kernel2<16><<<blocks, nw*32, 60*1024>>>(p.G, p.stride, p.step, iters, d_out, (unsigned*)d_sink);
			CK(cudaDeviceSynchronize());
			CK(cudaMemcpy(h_out, d_out, blocks*nw*sizeof(long long), cudaMemcpyDeviceToHost));
			double mx = 0; for(int i=0; i<blocks*nw; i++) if(h_out[i]>mx) mx = h_out[i];
			printf("%-48s warps/SM %2d : %.2f SM-cycles per LDS warp-instr\n", p.name, nw, mx/((double)iters*16*nw));
			}
		}
	dfma_lat<<<1, 32>>>(iters, d_out, d_sink, 1.0, 0.5); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost)); printf("DFMA dependent latency: %.2f cycles\n", h_out[0]/((double)iters*16));
	for(int nw : {4, 8, 16})
		{
		dfma_tp<8><<<148, nw*32>>>(iters, d_out, d_sink, 1.0, 0.5); CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(h_out, d_out, 148*nw*8, cudaMemcpyDeviceToHost));
		double mx = 0; for(int i=0; i<148*nw; i++) if(h_out[i]>mx) mx = h_out[i];
		printf("DFMA throughput ILP8 warps/SM %2d: %.3f SM-cycles per warp-DFMA (%.1f lanes/clk/SM)\n", nw, mx/((double)iters*64*nw), 32.0/(mx/((double)iters*64*nw)));
		}
	for(int nw : {4})
		{
		dfma_tp<2><<<148, nw*32>>>(iters, d_out, d_sink, 1.0, 0.5); CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(h_out, d_out, 148*nw*8, cudaMemcpyDeviceToHost));
		double mx = 0; for(int i=0; i<148*nw; i++) if(h_out[i]>mx) mx = h_out[i];
		printf("DFMA ILP2 1 warp/SMSP: %.3f cycles per DFMA per warp\n", mx/((double)iters*16));
		dfma_tp<4><<<148, nw*32>>>(iters, d_out, d_sink, 1.0, 0.5); CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(h_out, d_out, 148*nw*8, cudaMemcpyDeviceToHost));
		mx = 0; for(int i=0; i<148*nw; i++) if(h_out[i]>mx) mx = h_out[i];
		printf("DFMA ILP4 1 warp/SMSP: %.3f cycles per DFMA per warp\n", mx/((double)iters*32));
		}
	rsq_lat<<<1, 32>>>(iters, d_out, d_sink, 1.5); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost)); printf("MUFU.RSQ64H + DADD dependent latency: %.2f cycles\n", h_out[0]/((double)iters*8));
	shfl_lat<<<1, 32>>>(iters, d_out, d_sink, 1.5); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost)); printf("SHFL f64 (2x SHFL.32) dependent latency: %.2f cycles\n", h_out[0]/((double)iters*8));
	lds_lat<<<1, 32>>>(iters, d_out, (unsigned*)d_sink); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost)); printf("LDS.32 pointer-chase latency: %.2f cycles\n", h_out[0]/((double)iters*8));
	sts_lds_lat<<<1, 32>>>(iters, d_out, d_sink, 1.5); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h_out, d_out, 8, cudaMemcpyDeviceToHost)); printf("STS.64 + syncwarp + LDS.64 + DADD + syncwarp round trip: %.2f cycles\n", h_out[0]/((double)iters*8));
	return 0;
	}
