// xfer_probe.cu -- which mechanism moves `struct reb_particle` rows (128 B stride) between the host array and the device
// SoA fastest?  Stand-alone measurement tool (not part of libass_b200.so); its result picked the transfer path in state.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xfer_probe tools/xfer_probe.cu && ./xfer_probe [N]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

struct __align__(32) vec4d { double x, y, z, w; };
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_unpack(const unsigned char* __restrict__ rows, size_t stride, int n, vec4d* posm, vec4d* vel, vec4d* acc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double* r = (const double*) (rows + (size_t) i * stride);
	vec4d p, v, a;
	p.x = r[0]; p.y = r[1]; p.z = r[2]; p.w = r[9];
	v.x = r[3]; v.y = r[4]; v.z = r[5]; v.w = 0;
	a.x = r[6]; a.y = r[7]; a.z = r[8]; a.w = 0;
	posm[i] = p; vel[i] = v; acc[i] = a;
}
// one thread per row, three 32-byte sector loads straight from mapped host memory
__global__ void k_unpack_zc32(const unsigned char* __restrict__ rows, size_t stride, int n, vec4d* posm, vec4d* vel, vec4d* acc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double2* r = (const double2*) (rows + (size_t) i * stride);  // malloc'd rows are 16-byte aligned only
	double2 q0 = r[0], q1 = r[1], q2 = r[2], q3 = r[3], q4 = r[4];
	vec4d p, v, a;
	p.x = q0.x; p.y = q0.y; p.z = q1.x; p.w = q4.y;
	v.x = q1.y; v.y = q2.x; v.z = q2.y; v.w = 0;
	a.x = q3.x; a.y = q3.y; a.z = q4.x; a.w = 0;
	posm[i] = p; vel[i] = v; acc[i] = a;
}
// 16 lanes per row, one double per lane (coalesced 96 B runs), transposed through shared memory
__global__ void k_unpack_zc8(const unsigned char* __restrict__ rows, size_t stride, int n, vec4d* posm, vec4d* vel, vec4d* acc) {
	__shared__ double sm[16][17];
	const int row0 = blockIdx.x * 16;
	const int lr = threadIdx.x >> 4, c = threadIdx.x & 15;
	const int i = row0 + lr;
	if (i < n && c < 12) sm[lr][c] = ((const double*) (rows + (size_t) i * stride))[c];
	__syncthreads();
	if (threadIdx.x < 16) {
		const int j = row0 + threadIdx.x;
		if (j < n) {
			const double* r = sm[threadIdx.x];
			vec4d p, v, a;
			p.x = r[0]; p.y = r[1]; p.z = r[2]; p.w = r[9];
			v.x = r[3]; v.y = r[4]; v.z = r[5]; v.w = 0;
			a.x = r[6]; a.y = r[7]; a.z = r[8]; a.w = 0;
			posm[j] = p; vel[j] = v; acc[j] = a;
		}
	}
}
__global__ void k_pack_dense(double* __restrict__ out, int n, const vec4d* posm, const vec4d* vel, const vec4d* acc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	vec4d p = posm[i], v = vel[i], a = acc[i];
	double* r = out + (size_t) i * 9;
	r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = v.x; r[4] = v.y; r[5] = v.z; r[6] = a.x; r[7] = a.y; r[8] = a.z;
}
// one thread per row writes 9 doubles into the mapped host row: 32 + 32 + 8 bytes
__global__ void k_pack_zc_row(unsigned char* __restrict__ rows, size_t stride, int n, const vec4d* posm, const vec4d* vel, const vec4d* acc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	vec4d p = posm[i], v = vel[i], a = acc[i];
	double2* r = (double2*) (rows + (size_t) i * stride);
	r[0] = make_double2(p.x, p.y); r[1] = make_double2(p.z, v.x); r[2] = make_double2(v.y, v.z); r[3] = make_double2(a.x, a.y);
	((double*) r)[8] = a.z;
}
// 16 lanes per row, one double per lane, coalesced 72-byte runs
__global__ void k_pack_zc8(unsigned char* __restrict__ rows, size_t stride, int n, const vec4d* posm, const vec4d* vel, const vec4d* acc) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int i = t >> 4, c = t & 15;
	if (i >= n || c >= 9) return;
	const double* src = (c < 3) ? (const double*) (posm + i) + c : (c < 6) ? (const double*) (vel + i) + (c - 3) : (const double*) (acc + i) + (c - 6);
	((double*) (rows + (size_t) i * stride))[c] = *src;
}

template <class F>
static double timeit(F f, int reps = 10) {
	f();
	CK(cudaDeviceSynchronize());
	auto t0 = std::chrono::steady_clock::now();
	for (int r = 0; r < reps; r++) f();
	CK(cudaDeviceSynchronize());
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main(int argc, char** argv) {
	int n = argc > 1 ? atoi(argv[1]) : 96536;
	const size_t stride = 128, bytes = (size_t) n * stride;
	unsigned char* pageable = (unsigned char*) malloc(bytes + 4096);
	unsigned char* reg = (unsigned char*) malloc(bytes + 4096);
	memset(pageable, 1, bytes); memset(reg, 1, bytes);
	auto t0 = std::chrono::steady_clock::now();
	CK(cudaHostRegister(reg, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
	printf("cudaHostRegister(%zu bytes, malloc'd, unaligned ok): %.3f ms\n", bytes, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	unsigned char* reg_d = nullptr;
	CK(cudaHostGetDevicePointer((void**) &reg_d, reg, 0));
	printf("device alias %s host pointer\n", reg_d == reg ? "==" : "!=");
	double* bounce = nullptr;
	CK(cudaMallocHost((void**) &bounce, (size_t) n * 10 * 8));
	unsigned char* stage; vec4d *posm, *vel, *acc;
	CK(cudaMalloc(&stage, bytes)); CK(cudaMalloc(&posm, 32 * (size_t) n)); CK(cudaMalloc(&vel, 32 * (size_t) n)); CK(cudaMalloc(&acc, 32 * (size_t) n));
	cudaStream_t st; CK(cudaStreamCreate(&st));
	const int g = (n + 255) / 256;
	printf("N = %d rows of %zu B (%.1f MB)\n", n, stride, bytes / 1e6);
	printf("UP  pageable memcpy + unpack      : %.3f ms\n", timeit([&] { cudaMemcpyAsync(stage, pageable, bytes, cudaMemcpyHostToDevice, st); k_unpack<<<g, 256, 0, st>>>(stage, stride, n, posm, vel, acc); cudaStreamSynchronize(st); }));
	printf("UP  registered memcpy + unpack    : %.3f ms\n", timeit([&] { cudaMemcpyAsync(stage, reg, bytes, cudaMemcpyHostToDevice, st); k_unpack<<<g, 256, 0, st>>>(stage, stride, n, posm, vel, acc); cudaStreamSynchronize(st); }));
	printf("UP  zero-copy 5x16B per row       : %.3f ms\n", timeit([&] { k_unpack_zc32<<<g, 256, 0, st>>>(reg_d, stride, n, posm, vel, acc); cudaStreamSynchronize(st); }));
	printf("UP  zero-copy 16 lanes per row    : %.3f ms\n", timeit([&] { k_unpack_zc8<<<(n + 15) / 16, 256, 0, st>>>(reg_d, stride, n, posm, vel, acc); cudaStreamSynchronize(st); }));
	printf("DN  pack + memcpy bounce + scatter: %.3f ms\n", timeit([&] {
		k_pack_dense<<<g, 256, 0, st>>>((double*) stage, n, posm, vel, acc);
		cudaMemcpyAsync(bounce, stage, (size_t) n * 72, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
		for (int i = 0; i < n; i++) memcpy(pageable + (size_t) i * stride, bounce + (size_t) i * 9, 72);
	}));
	printf("DN  pack + memcpy2D into rows     : %.3f ms\n", timeit([&] {
		k_pack_dense<<<g, 256, 0, st>>>((double*) stage, n, posm, vel, acc);
		cudaMemcpy2DAsync(reg, stride, stage, 72, 72, n, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
	}));
	printf("DN  zero-copy row writes 4x16+8   : %.3f ms\n", timeit([&] { k_pack_zc_row<<<g, 256, 0, st>>>(reg_d, stride, n, posm, vel, acc); cudaStreamSynchronize(st); }));
	printf("DN  zero-copy 16 lanes per row    : %.3f ms\n", timeit([&] { k_pack_zc8<<<(n * 16 + 255) / 256, 256, 0, st>>>(reg_d, stride, n, posm, vel, acc); cudaStreamSynchronize(st); }));
	// full rows both ways by DMA (upper bound of the copy engines on this link)
	printf("DMA H2D %.1f MB registered          : %.3f ms\n", bytes / 1e6, timeit([&] { cudaMemcpyAsync(stage, reg, bytes, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); }));
	printf("DMA D2H %.1f MB registered          : %.3f ms\n", bytes / 1e6, timeit([&] { cudaMemcpyAsync(reg, stage, bytes, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }));
	// check the zero-copy round trip
	k_unpack_zc8<<<(n + 15) / 16, 256, 0, st>>>(reg_d, stride, n, posm, vel, acc);
	memset(pageable, 0, bytes);
	k_pack_zc8<<<(n * 16 + 255) / 256, 256, 0, st>>>(reg_d, stride, n, posm, vel, acc);
	CK(cudaStreamSynchronize(st));
	CK(cudaHostUnregister(reg));
	printf("done\n");
	return 0;
}
