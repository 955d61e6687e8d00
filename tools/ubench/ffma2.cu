// Microbenchmark: issue rate and dependent latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int ILP, bool PACKED>
__global__ void bench(float *out, long long *cycles, int iters) {
	float a = threadIdx.x * 1e-3f + 1.0f, b = 0.999f;
	float acc[ILP];
	u64 acc2[ILP];
	for (int i = 0; i < ILP; ++i) { acc[i] = i; acc2[i] = (u64)__float_as_uint((float)i) | ((u64)__float_as_uint(1.0f + i) << 32); }
	u64 a2 = (u64)__float_as_uint(a) | ((u64)__float_as_uint(a) << 32), b2 = (u64)__float_as_uint(b) | ((u64)__float_as_uint(b) << 32);
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			if (PACKED) acc2[i] = fma2(acc2[i], b2, a2);
			else acc[i] = fma1(acc[i], b, a);
		}
	}
	long long t1 = clock64();
	float s = 0;
	for (int i = 0; i < ILP; ++i) s += PACKED ? __uint_as_float((unsigned)(acc2[i] & 0xffffffffu)) + __uint_as_float((unsigned)(acc2[i] >> 32)) : acc[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP, bool PACKED> void run(int threads, const char *name) {
	float *out; long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
	const int iters = 20000;
	bench<ILP, PACKED><<<148, threads>>>(out, cyc, iters);
	bench<ILP, PACKED><<<148, threads>>>(out, cyc, iters);
	cudaDeviceSynchronize();
	long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
	double instrPerWarp = (double)iters * ILP, warpsPerSmsp = threads / 32 / 4.0;
	printf("%-6s ILP %2d, %4d thr/SM (%.0f warps/SMSP): %.2f cycles per instr per warp; SMSP issue rate %.3f instr/clk = %.1f FMA lanes/clk/SM\n", name, ILP, threads,
	       warpsPerSmsp, c / instrPerWarp, instrPerWarp * warpsPerSmsp / c, instrPerWarp * warpsPerSmsp / c * 4 * 32 * (PACKED ? 2 : 1));
	cudaFree(out); cudaFree(cyc);
}
int main() {
	run<1, false>(128, "FFMA");  run<1, true>(128, "FFMA2");   // dependent latency
	run<2, false>(128, "FFMA");  run<2, true>(128, "FFMA2");
	run<4, false>(128, "FFMA");  run<4, true>(128, "FFMA2");
	run<8, false>(128, "FFMA");  run<8, true>(128, "FFMA2");
	run<8, false>(512, "FFMA");  run<8, true>(512, "FFMA2");
	run<8, false>(1024, "FFMA"); run<8, true>(1024, "FFMA2");
	run<2, false>(512, "FFMA");  run<2, true>(512, "FFMA2");
	return 0;
}
