// fp64_probe.cu -- dependent-issue latency of the FP64 instructions the spring / gravity kernels are made of (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe tools/fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP>
__global__ void k(double* out, long long* cyc, double a, double b) {
	double x = a + threadIdx.x * 1e-9, y = b;
	long long t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < N; i++) {
		if (OP == 0) x = fma(x, y, 1e-3);
		if (OP == 1) x = x + y;
		if (OP == 2) x = x * y;
		if (OP == 3) x = __dsqrt_rn(x + 1.5);
		if (OP == 4) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
		if (OP == 5) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
		if (OP == 6) x = 1.0 / (x + 1.5);
		if (OP == 7) x = __shfl_xor_sync(0xffffffffu, x, 1) + y;
		if (OP == 8) x = fmaf((float) x, 1.0001f, 1e-3f);
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int warps) {
	double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
	k<OP><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999);
	k<OP><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999);
	long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("%-28s warps/CTA=%2d : %.1f cycles per dependent op (per warp)\n", name, warps, (double) h / N);
	cudaFree(out); cudaFree(cyc);
}
int main() {
	for (int w : {1, 4, 8, 16, 32}) {
		if (w == 1) { run<0>("DFMA", 1); run<1>("DADD", 1); run<2>("DMUL", 1); run<3>("DADD+dsqrt_rn", 1); run<4>("MUFU.RCP64H+DADD", 1); run<5>("MUFU.RSQ64H+DADD", 1); run<6>("DADD + 1.0/x (IEEE div)", 1); run<7>("SHFL(double)+DADD", 1); run<8>("cvt+FFMA (fp32 ref)", 1); }
		if (w == 4) { run<0>("DFMA", 4); run<3>("DADD+dsqrt_rn", 4); }
		if (w == 8) { run<0>("DFMA", 8); run<3>("DADD+dsqrt_rn", 8); }
		if (w == 16) { run<0>("DFMA", 16); run<3>("DADD+dsqrt_rn", 16); }
		if (w == 32) { run<0>("DFMA", 32); run<3>("DADD+dsqrt_rn", 32); }
	}
	return 0;
}
