// probe.cu -- roofline denominators measured on the box: a DFMA-only kernel (MEASURED_PEAKS.json carries no FP64
// figure) and a plain device copy.  Not on the simulation path.
#include "common.cuh"

namespace {

// 8 independent DFMA chains per thread, fully unrolled; 2 flop per DFMA.
__global__ void __launch_bounds__(256) k_dfma(double* __restrict__ out, int iters, double a, double b) {
	double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
			r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
		}
	}
	const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
	if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains alive
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (; i < n; i += stride) out[i] = in[i];
}

}  // namespace

extern "C" int ass_probe_fp64_peak(int iters, double* tflops, double* ms_out) {
	REQUIRE(tflops, ASS_EINVAL, "null out");
	REQUIRE(iters > 0, ASS_EINVAL, "iters must be positive");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int blocks = ass_sm_count() * 8, threads = 256;
	double* d = nullptr;
	CU_TRY(cudaMalloc(&d, sizeof(double) * (size_t) blocks * threads));
	cudaEvent_t e0, e1;
	CU_TRY(cudaEventCreate(&e0));
	CU_TRY(cudaEventCreate(&e1));
	k_dfma<<<blocks, threads>>>(d, 64, 0.999999, 1e-9);  // warm-up
	g_launch_count++;
	float best = 1e30f;
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(e0);
		k_dfma<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
		g_launch_count++;
		cudaEventRecord(e1);
		cudaError_t e = cudaEventSynchronize(e1);
		if (e != cudaSuccess) {
			ass_set_error("fp64 probe: %s", cudaGetErrorString(e));
			cudaFree(d);
			return ASS_ECUDA;
		}
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	const double flop = 2.0 * 64.0 * (double) iters * (double) blocks * threads;
	*tflops = flop / (best * 1e-3) / 1e12;
	if (ms_out) *ms_out = best;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	return ASS_OK;
}

extern "C" int ass_probe_copy_bw(long long bytes, double* gbs) {
	REQUIRE(gbs, ASS_EINVAL, "null out");
	REQUIRE(bytes >= (1 << 20), ASS_EINVAL, "use at least 1 MiB");
	int rc = ass_ensure_device();
	if (rc) return rc;
	double2 *a = nullptr, *b = nullptr;
	const size_t n = (size_t) bytes / sizeof(double2);
	if (cudaMalloc(&a, n * sizeof(double2)) != cudaSuccess || cudaMalloc(&b, n * sizeof(double2)) != cudaSuccess) {
		cudaGetLastError();
		cudaFree(a);
		ass_set_error("copy probe: out of device memory");
		return ASS_ENOMEM;
	}
	cudaMemset(a, 0, n * sizeof(double2));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 6; rep++) {
		cudaEventRecord(e0);
		k_copy<<<ass_sm_count() * 16, 256>>>(a, b, n);
		g_launch_count++;
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep > 0 && ms < best) best = ms;
	}
	*gbs = 2.0 * (double) (n * sizeof(double2)) / (best * 1e-3) / 1e9;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(a);
	cudaFree(b);
	return ASS_OK;
}

// ---- self-check of the branch-free square root used by the spring kernels (spring_math.cuh) ----
#include "spring_math.cuh"

namespace {
__device__ __forceinline__ uint64_t splitmix(uint64_t& x) {
	x += 0x9E3779B97F4A7C15ULL;
	uint64_t z = x;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
// every thread tries `per_thread` arguments: random mantissa, exponent uniform in [2^e_lo, 2^e_hi)
__global__ void k_check_sqrt(long long per_thread, int e_lo, int e_hi, uint64_t seed, unsigned long long* mismatches, double* first_bad) {
	uint64_t st = seed + 0x1234567ULL * ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x);
	unsigned long long bad = 0;
	for (long long q = 0; q < per_thread; q++) {
		const uint64_t r = splitmix(st);
		const int ex = e_lo + (int) ((r >> 52) % (uint64_t) (e_hi - e_lo));
		const uint64_t bits = ((uint64_t) (ex + 1023) << 52) | (r & 0x000FFFFFFFFFFFFFULL);
		const double a = __longlong_as_double((long long) bits);
		if (__double_as_longlong(sqrt_rn_normal(a)) != __double_as_longlong(__dsqrt_rn(a))) {
			if (bad == 0 && atomicAdd(mismatches, 0ULL) == 0) *first_bad = a;
			bad++;
		}
	}
	if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int ass_selfcheck_sqrt(long long samples, int exp_lo, int exp_hi, long long* mismatches, double* first_bad) {
	REQUIRE(mismatches && first_bad, ASS_EINVAL, "null out");
	REQUIRE(samples > 0 && exp_hi > exp_lo && exp_lo > -960 && exp_hi < 1020, ASS_EINVAL, "bad range");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int blocks = ass_sm_count() * 8, threads = 256;
	const long long per_thread = (samples + (long long) blocks * threads - 1) / ((long long) blocks * threads);
	unsigned long long* d_bad = nullptr;
	double* d_first = nullptr;
	CU_TRY(cudaMalloc(&d_bad, sizeof(unsigned long long)));
	CU_TRY(cudaMalloc(&d_first, sizeof(double)));
	CU_TRY(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
	CU_TRY(cudaMemset(d_first, 0, sizeof(double)));
	k_check_sqrt<<<blocks, threads>>>(per_thread, exp_lo, exp_hi, 0xC0FFEEULL, d_bad, d_first);
	LAUNCH_CHECK();
	unsigned long long bad = 0;
	CU_TRY(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
	CU_TRY(cudaMemcpy(first_bad, d_first, sizeof(double), cudaMemcpyDeviceToHost));
	cudaFree(d_bad);
	cudaFree(d_first);
	*mismatches = (long long) bad;
	return ASS_OK;
}

// Pair-once gravity between two particle ranges of a simulation, timed on the device (L2 flushed before every repetition):
// the pieces a rank of a sharded body launches per step (own triangle, rectangles against runs of slabs, half of the
// antipodal rectangle), measurable on ONE GPU.  Writes the simulation's acceleration array (measurement only).
extern "C" int ass_probe_gravity_pair(ass_sim* s, int rlo, int rhi, int slo, int shi, int reps, double* ms_avg) {
	REQUIRE(s && ms_avg && reps > 0, ASS_EINVAL, "bad argument");
	{
		int rc0 = ass_ensure_device();
		if (rc0) return rc0;
	}
	REQUIRE(s->have_state && !s->shard.active, ASS_ESTATE, "needs an unsharded simulation with particles");
	REQUIRE(0 <= rlo && rlo < rhi && rhi <= s->n && 0 <= slo && slo < shi && shi <= s->n, ASS_EINVAL, "bad range");
	REQUIRE((rlo == slo && rhi == shi) || rhi <= slo || shi <= rlo, ASS_EINVAL, "ranges must be identical or disjoint");
	vec4d* part = nullptr;
	CU_TRY(cudaMalloc(&part, sizeof(vec4d) * (size_t) s->n));
	int rc = ass_launch_gravity_sym_pair(s, rlo, rhi, slo, shi, s->acc, false, part, false);  // builds the plan
	double total = 0.0;
	for (int rep = 0; rep < reps && !rc; rep++) {
		double ms = 0.0;
		if ((rc = ass_l2_flush(s)) || (rc = ass_mark_begin(s))) break;
		if ((rc = ass_launch_gravity_sym_pair(s, rlo, rhi, slo, shi, s->acc, false, part, false))) break;
		if ((rc = ass_mark_end(s, &ms))) break;
		total += ms;
	}
	cudaStreamSynchronize(s->stream);
	cudaFree(part);
	if (rc) return rc;
	*ms_avg = total / reps;
	return ASS_OK;
}

// Several range pairs per repetition, as a rank of a sharded body runs them: one after the other on the simulation's
// stream (concurrent == 0), or each pair kernel on a stream of its own with the reductions afterwards (concurrent != 0).
// ranges: n_pieces x {rlo, rhi, slo, shi}.
extern "C" int ass_probe_gravity_pieces(ass_sim* s, int n_pieces, const int* ranges, int concurrent, int reps, double* ms_avg) {
	REQUIRE(s && ranges && ms_avg && reps > 0 && n_pieces > 0 && n_pieces <= 4, ASS_EINVAL, "bad argument");
	{
		int rc0 = ass_ensure_device();
		if (rc0) return rc0;
	}
	REQUIRE(s->have_state && !s->shard.active, ASS_ESTATE, "needs an unsharded simulation with particles");
	vec4d* part = nullptr;
	CU_TRY(cudaMalloc(&part, sizeof(vec4d) * (size_t) s->n));
	cudaStream_t st[4] = {};
	cudaEvent_t go = nullptr, done[4] = {};
	cudaEventCreateWithFlags(&go, cudaEventDisableTiming);
	for (int q = 0; q < n_pieces; q++) { cudaStreamCreateWithFlags(&st[q], cudaStreamNonBlocking); cudaEventCreateWithFlags(&done[q], cudaEventDisableTiming); }
	int rc = ASS_OK;
	for (int q = 0; q < n_pieces && !rc; q++) rc = ass_gravity_sym_prepare(s, ranges[4 * q], ranges[4 * q + 1], ranges[4 * q + 2], ranges[4 * q + 3]);
	double total = 0.0;
	for (int rep = -1; rep < reps && !rc; rep++) {
		double ms = 0.0;
		if ((rc = ass_l2_flush(s)) || (rc = ass_mark_begin(s))) break;
		if (concurrent) {
			cudaEventRecord(go, s->stream);
			for (int q = 0; q < n_pieces && !rc; q++) {
				const int* r = ranges + 4 * q;
				cudaStreamWaitEvent(st[q], go, 0);
				rc = ass_launch_gravity_sym_main(s, r[0], r[1], r[2], r[3], st[q]);
				cudaEventRecord(done[q], st[q]);
			}
			for (int q = 0; q < n_pieces && !rc; q++) {
				const int* r = ranges + 4 * q;
				cudaStreamWaitEvent(s->stream, done[q], 0);
				rc = ass_launch_gravity_sym_finish(s, r[0], r[1], r[2], r[3], s->acc, q > 0, part, false);
			}
		} else {
			for (int q = 0; q < n_pieces && !rc; q++) {
				const int* r = ranges + 4 * q;
				rc = ass_launch_gravity_sym_pair(s, r[0], r[1], r[2], r[3], s->acc, q > 0, part, false);
			}
		}
		if (rc || (rc = ass_mark_end(s, &ms))) break;
		if (rep >= 0) total += ms;
	}
	cudaStreamSynchronize(s->stream);
	for (int q = 0; q < n_pieces; q++) { cudaStreamSynchronize(st[q]); cudaStreamDestroy(st[q]); cudaEventDestroy(done[q]); }
	cudaEventDestroy(go);
	cudaFree(part);
	if (rc) return rc;
	*ms_avg = total / reps;
	return ASS_OK;
}
