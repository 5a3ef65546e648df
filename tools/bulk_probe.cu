// bulk_probe.cu -- how fast can persistent CTAs stream HBM through a cp.async.bulk + mbarrier ring?  (B200)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe tools/bulk_probe.cu && ./bulk_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
	asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int STAGES>
__global__ void k_ring(const unsigned char* __restrict__ src, size_t total, int chunk, double* sink, int pieces) {
	extern __shared__ __align__(128) unsigned char sm[];
	__shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
	const int tid = threadIdx.x, nwarp = blockDim.x / 32 - 1;
	if (tid == 0) { for (int q = 0; q < STAGES; q++) { mbar_init(&full[q], 1); mbar_init(&empty[q], nwarp); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	__syncthreads();
	const size_t nchunks = total / chunk;
	if (tid >= nwarp * 32) {
		if (tid != nwarp * 32) return;
		int it = 0;
		for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, it++) {
			const int st = it % STAGES;
			if (it >= STAGES) mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1);
			mbar_expect_tx(&full[st], chunk);
			const int piece = chunk / pieces;
			for (int p = 0; p < pieces; p++) bulk_g2s(sm + (size_t) st * chunk + p * piece, src + c * chunk + p * piece, piece, &full[st]);
		}
		return;
	}
	int it = 0; double acc = 0;
	for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, it++) {
		const int st = it % STAGES;
		mbar_wait(&full[st], (it / STAGES) & 1);
		acc += ((const double*) (sm + (size_t) st * chunk))[tid];
		__syncwarp();
		if ((tid & 31) == 0) mbar_arrive(&empty[st]);
	}
	if (acc == 123.456) sink[0] = acc;
}
// plain coalesced LDG.128 grid-stride read for comparison
__global__ void k_ldg(const double2* __restrict__ src, size_t n, double* sink) {
	double a = 0;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) { double2 v = src[i]; a += v.x + v.y; }
	if (a == 123.456) sink[0] = a;
}
template <int STAGES> void run(const unsigned char* src, size_t total, int chunk, int ctas_per_sm, int pieces, double* sink) {
	size_t smem = (size_t) STAGES * chunk;
	cudaFuncSetAttribute(k_ring<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_ring<STAGES><<<148 * ctas_per_sm, 288, smem>>>(src, total, chunk, sink, pieces);
	cudaEventRecord(e0);
	k_ring<STAGES><<<148 * ctas_per_sm, 288, smem>>>(src, total, chunk, sink, pieces);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	cudaError_t err = cudaGetLastError();
	printf("ring stages=%d chunk=%6d B x%d pieces, %d CTA/SM (%3zu KB in flight/SM): %7.1f us  %6.0f GB/s %s\n", STAGES, chunk, pieces, ctas_per_sm, smem * ctas_per_sm / 1024, ms * 1e3, total / ms / 1e6, err ? cudaGetErrorString(err) : "");
}
int main() {
	for (size_t total : { (size_t) 48 << 20, (size_t) 1 << 30 }) {
		unsigned char* src; double* sink; cudaMalloc(&src, total); cudaMalloc(&sink, 8); cudaMemset(src, 0, total);
		unsigned char* flush; cudaMalloc(&flush, 256 << 20);
		printf("== %zu MB\n", total >> 20);
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		for (int rep = 0; rep < 2; rep++) {
			cudaMemset(flush, 1, 256 << 20);
			cudaEventRecord(e0); k_ldg<<<148 * 8, 256>>>((const double2*) src, total / 16, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			printf("plain LDG.128 grid-stride: %7.1f us %6.0f GB/s\n", ms * 1e3, total / ms / 1e6);
		}
		for (int cps : {1, 2, 3}) {
			cudaMemset(flush, 1, 256 << 20); run<2>(src, total, 16384, cps, 1, sink);
			cudaMemset(flush, 1, 256 << 20); run<3>(src, total, 16384, cps, 1, sink);
			cudaMemset(flush, 1, 256 << 20); run<4>(src, total, 16384, cps, 1, sink);
			cudaMemset(flush, 1, 256 << 20); run<3>(src, total, 16384, cps, 4, sink);
			cudaMemset(flush, 1, 256 << 20); run<2>(src, total, 32768, cps, 1, sink);
			cudaMemset(flush, 1, 256 << 20); run<3>(src, total, 32768, cps == 3 ? 2 : cps, 8, sink);
			cudaMemset(flush, 1, 256 << 20); run<4>(src, total, 8192, cps, 1, sink);
		}
		cudaFree(src);
	}
	return 0;
}
