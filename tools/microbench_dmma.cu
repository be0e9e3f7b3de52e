// DMMA (mma.sync.m8n8k4.f64) throughput and latency on sm_100a next to the DFMA pipe: is the FP64 tensor path worth it for the
// any-size stage contractions?  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_dmma.bin tools/microbench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if(e!=cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while(0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
	{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
	}

template<int ILP>
__global__ void dmma_tp(int iters, long long *out, double *sink, double a, double b)
	{
	double c[ILP][2];
#pragma unroll
	for(int i=0; i<ILP; i++) { c[i][0] = threadIdx.x*1e-9 + i; c[i][1] = i; }
	long long t0 = clock64();
	for(int it=0; it<iters; it++)
		{
#pragma unroll
		for(int i=0; i<ILP; i++) dmma(c[i][0], c[i][1], a, b);
		}
	long long t1 = clock64();
	double s = 0;
#pragma unroll
	for(int i=0; i<ILP; i++) s += c[i][0] + c[i][1];
	sink[blockIdx.x*blockDim.x+threadIdx.x] = s;
	if(threadIdx.x==0) out[blockIdx.x] = t1 - t0;
	}

int main()
	{
	long long *d_out; double *sink; long long h[1024];
	CK(cudaMalloc(&d_out, 8*1024)); CK(cudaMalloc(&sink, 8*1024*1024));
	int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	const int iters = 20000;
	dmma_tp<1><<<1, 32>>>(iters, d_out, sink, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
	printf("DMMA m8n8k4 dependent latency: %.2f cycles\n", h[0]/(double)iters);
	for(int nw : {4, 8, 16})
		{
		dmma_tp<8><<<sms, 32*nw>>>(iters, d_out, sink, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(h, d_out, 8*sms, cudaMemcpyDeviceToHost));
		long long mx = 0; for(int i=0; i<sms; i++) if(h[i]>mx) mx = h[i];
		double cyc = mx/((double)iters*8*nw);
		printf("DMMA throughput ILP8 warps/SM %2d: %.3f SM-cycles per warp-DMMA = %.1f FMA/clk/SM (DFMA pipe: 64)\n", nw, cyc, 256.0/cyc);
		}
	dmma_tp<2><<<sms, 128>>>(iters, d_out, sink, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(h, d_out, 8*sms, cudaMemcpyDeviceToHost));
	printf("DMMA ILP2 1 warp/SMSP: %.2f cycles per DMMA per warp\n", h[0]/((double)iters*2));
	return 0;
	}
