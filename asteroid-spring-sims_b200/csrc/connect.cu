// connect.cu -- connect_springs_dist (reference src_spring/springs.cpp:110-143) and mindist (src_spring/shapes.cpp:814-833)
// as a uniform-grid cell-list search on the device.
//
// The reference tests all N^2/2 pairs and, for every accepted pair, scans the whole spring list for a duplicate
// (O(N^2 + NS^2); 11.8 s at N = 9.3k).  Here: counting-sort the particles into cells of edge >= max_dist,
// let every particle scan its 27 neighbouring cells, count, prefix-sum, fill, and sort each particle's short partner
// list.  The output is the reference's list bit for bit:
//   - pair test   sqrt(fma(dz,dz, fma(dx,dx, dy*dy))) < max_dist, strict, IEEE sqrt  (Vector::len)
//   - order       lexicographic (i, j), i < j                                         (springs.cpp:122-126)
//   - rs0         that same length                                                    (springs.cpp:84)
//   - duplicates  pairs already present in the caller's list are skipped              (springs.cpp:70-79)
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace {

struct Grid {
	double x0, y0, z0, inv_h, h;
	int nx, ny, nz;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// fmin / fmax first: they map NaN to the bound and keep +-inf inside the int range, so a non-finite coordinate lands in
// a border cell instead of in an undefined double -> int conversion
__device__ __forceinline__ int cell_axis(double x, double x0, double inv_h, int n) {
	return (int) fmin(fmax(floor((x - x0) * inv_h), 0.0), (double) (n - 1));
}
__device__ __forceinline__ void cell_of(const Grid g, double x, double y, double z, int& cx, int& cy, int& cz) {
	cx = cell_axis(x, g.x0, g.inv_h, g.nx);
	cy = cell_axis(y, g.y0, g.inv_h, g.ny);
	cz = cell_axis(z, g.z0, g.inv_h, g.nz);
}

// ---- bounding box ----
__global__ void k_bbox(const vec4d* __restrict__ posm, int lo, int hi, double* __restrict__ part /* [grid][6] */) {
	double mn[3] = { 1e308, 1e308, 1e308 }, mx[3] = { -1e308, -1e308, -1e308 };
	for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
		const vec4d p = posm[i];
		// a particle with a non-finite coordinate (a blown-up run) has no partner within any cutoff -- every length to it
		// is inf or NaN, which is never < max_dist -- so it stays out of the box; cell_of() clamps it into a border cell
		if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
		mn[0] = fmin(mn[0], p.x); mn[1] = fmin(mn[1], p.y); mn[2] = fmin(mn[2], p.z);
		mx[0] = fmax(mx[0], p.x); mx[1] = fmax(mx[1], p.y); mx[2] = fmax(mx[2], p.z);
	}
	__shared__ double sm[8][6];
	for (int m = 16; m > 0; m >>= 1)
		for (int d = 0; d < 3; d++) {
			mn[d] = fmin(mn[d], shfl_down_d(mn[d], m));
			mx[d] = fmax(mx[d], shfl_down_d(mx[d], m));
		}
	if ((threadIdx.x & 31) == 0)
		for (int d = 0; d < 3; d++) { sm[threadIdx.x >> 5][d] = mn[d]; sm[threadIdx.x >> 5][3 + d] = mx[d]; }
	__syncthreads();
	if (threadIdx.x < 6) {
		double a = sm[0][threadIdx.x];
		for (int w = 1; w < (int) (blockDim.x >> 5); w++) a = threadIdx.x < 3 ? fmin(a, sm[w][threadIdx.x]) : fmax(a, sm[w][threadIdx.x]);
		part[blockIdx.x * 6 + threadIdx.x] = a;
	}
}

// ---- counting sort into cells ----
__global__ void k_cell_count(const vec4d* __restrict__ posm, int lo, int hi, Grid g, int* __restrict__ cell_of_p, int* __restrict__ count) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const vec4d p = posm[i];
	int cx, cy, cz;
	cell_of(g, p.x, p.y, p.z, cx, cy, cz);
	const int c = (cz * g.ny + cy) * g.nx + cx;
	cell_of_p[i - lo] = c;
	atomicAdd(&count[c], 1);
}
__global__ void k_cell_scatter(const vec4d* __restrict__ posm, int lo, int hi, const int* __restrict__ cell_of_p, const int* __restrict__ start,
		int* __restrict__ fill, vec4d* __restrict__ sorted) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const int c = cell_of_p[i - lo];
	const int at = start[c] + atomicAdd(&fill[c], 1);
	vec4d p = posm[i];
	p.w = (double) i;  // carry the particle index with the position (exact for i < 2^53)
	sorted[at] = p;
}

// ---- exclusive scan of ints (three passes, 1024 items per CTA) ----
constexpr int SCAN_B = 256, SCAN_ITEMS = 4;
__global__ void k_scan_local(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ block_sum) {
	__shared__ int sm[SCAN_B];
	const int base = blockIdx.x * SCAN_B * SCAN_ITEMS + threadIdx.x * SCAN_ITEMS;
	int v[SCAN_ITEMS], t = 0;
	for (int q = 0; q < SCAN_ITEMS; q++) { v[q] = (base + q < n) ? in[base + q] : 0; t += v[q]; }
	sm[threadIdx.x] = t;
	__syncthreads();
	for (int off = 1; off < SCAN_B; off <<= 1) {
		int a = (threadIdx.x >= off) ? sm[threadIdx.x - off] : 0;
		__syncthreads();
		sm[threadIdx.x] += a;
		__syncthreads();
	}
	int run = sm[threadIdx.x] - t;
	for (int q = 0; q < SCAN_ITEMS; q++) { if (base + q < n) out[base + q] = run; run += v[q]; }
	if (threadIdx.x == SCAN_B - 1) block_sum[blockIdx.x] = sm[threadIdx.x];
}
__global__ void k_scan_blocks(int* __restrict__ block_sum, int nb, int* __restrict__ total) {
	// single thread: nb <= n/1024 (a few thousand)
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		int run = 0;
		for (int b = 0; b < nb; b++) { int v = block_sum[b]; block_sum[b] = run; run += v; }
		*total = run;
	}
}
__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ block_sum) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] += block_sum[i / (SCAN_B * SCAN_ITEMS)];
}

// ---- neighbour search ----
// MODE 0: count partners j > i within max_dist;  MODE 1: write them at list[off[i]...];  MODE 2: min distance
template <int MODE>
__global__ void k_neighbours(const vec4d* __restrict__ posm, int lo, int hi, Grid g, const int* __restrict__ start, const vec4d* __restrict__ sorted,
		double max_dist, int* __restrict__ count, const int* __restrict__ off, int* __restrict__ list, double* __restrict__ best) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	double my_best = 1.7976931348623157e308;
	int c = 0;
	if (i < hi) {
		const vec4d p = posm[i];
		int cx, cy, cz;
		cell_of(g, p.x, p.y, p.z, cx, cy, cz);
		const int base = (MODE == 1) ? off[i - lo] : 0;
		for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); z++)
			for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); y++) {
				// the three x-cells are contiguous in the sorted array
				const int c0 = (z * g.ny + y) * g.nx + max(cx - 1, 0);
				const int c1 = (z * g.ny + y) * g.nx + min(cx + 1, g.nx - 1);
				for (int q = start[c0]; q < start[c1 + 1]; q++) {
					const vec4d s = sorted[q];
					const int j = (int) s.w;
					if (j <= i) continue;
					const double d = ref_len(p.x - s.x, p.y - s.y, p.z - s.z);
					if (MODE == 2) {
						my_best = fmin(my_best, d);
					} else if (d < max_dist) {
						if (MODE == 1) list[base + c] = j;
						c++;
					}
				}
			}
		if (MODE == 0) count[i - lo] = c;
		if (MODE == 1) {  // ascending j: insertion sort of this particle's short segment
			for (int a = 1; a < c; a++) {
				const int v = list[base + a];
				int b = a - 1;
				while (b >= 0 && list[base + b] > v) { list[base + b + 1] = list[base + b]; b--; }
				list[base + b + 1] = v;
			}
		}
	}
	if (MODE == 2) {
		for (int m = 16; m > 0; m >>= 1) my_best = fmin(my_best, shfl_down_d(my_best, m));
		// positive doubles order like their bit patterns
		if ((threadIdx.x & 31) == 0) atomicMin((unsigned long long*) best, (unsigned long long) __double_as_longlong(my_best));
	}
}

__global__ void k_emit_springs(const vec4d* __restrict__ posm, int lo, int hi, const int* __restrict__ off, const int* __restrict__ count,
		const int* __restrict__ list, double k, double gamma, ass_spring* __restrict__ out) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const vec4d p = posm[i];
	const int base = off[i - lo], c = count[i - lo];
	for (int a = 0; a < c; a++) {
		const int j = list[base + a];
		const vec4d q = posm[j];
		ass_spring s;
		s.k = k; s.gamma = gamma;
		s.rs0 = ref_len(p.x - q.x, p.y - q.y, p.z - q.z);  // spring_r(...).len(), springs.cpp:84
		s.particle_1 = i; s.particle_2 = j;
		out[base + a] = s;
	}
}

struct Scratch {
	int *cell_of_p = nullptr, *count = nullptr, *start = nullptr, *fill = nullptr, *bsum = nullptr, *total = nullptr;
	int *ncount = nullptr, *noff = nullptr, *list = nullptr;
	vec4d* sorted = nullptr;
	double* part = nullptr;
	ass_spring* springs = nullptr;
	~Scratch() {
		cudaFree(cell_of_p); cudaFree(count); cudaFree(start); cudaFree(fill); cudaFree(bsum); cudaFree(total);
		cudaFree(ncount); cudaFree(noff); cudaFree(list); cudaFree(sorted); cudaFree(part); cudaFree(springs);
	}
};

int scan_exclusive(ass_sim* s, const int* in, int* out, int n, int* bsum, int* total_dev) {
	const int nb = (n + SCAN_B * SCAN_ITEMS - 1) / (SCAN_B * SCAN_ITEMS);
	k_scan_local<<<nb, SCAN_B, 0, s->stream>>>(in, out, n, bsum);
	LAUNCH_CHECK();
	k_scan_blocks<<<1, 32, 0, s->stream>>>(bsum, nb, total_dev);
	LAUNCH_CHECK();
	k_scan_add<<<(n + 255) / 256, 256, 0, s->stream>>>(out, n, bsum);
	LAUNCH_CHECK();
	return ASS_OK;
}

#define CU_ALLOC(ptr, bytes)                                                       \
	do {                                                                           \
		if (cudaMalloc(&(ptr), (bytes)) != cudaSuccess) {                          \
			cudaGetLastError();                                                    \
			ass_set_error("connect: out of device memory (%zu bytes)", (size_t) (bytes)); \
			return ASS_ENOMEM;                                                     \
		}                                                                          \
	} while (0)

// build the cell grid for particles [lo,hi) with cell edge >= h; fills sc.start / sc.sorted
int build_grid(ass_sim* s, int lo, int hi, double h, Scratch& sc, Grid& g) {
	const int cnt = hi - lo;
	const int rb = std::max(1, std::min((cnt + 255) / 256, ass_sm_count() * 4));
	CU_ALLOC(sc.part, sizeof(double) * 6 * rb);
	k_bbox<<<rb, 256, 0, s->stream>>>(s->posm, lo, hi, sc.part);
	LAUNCH_CHECK();
	std::vector<double> hp((size_t) 6 * rb);
	CU_TRY(cudaMemcpyAsync(hp.data(), sc.part, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	double mn[3] = { 1e308, 1e308, 1e308 }, mx[3] = { -1e308, -1e308, -1e308 };
	for (int b = 0; b < rb; b++)
		for (int d = 0; d < 3; d++) {
			mn[d] = std::min(mn[d], hp[(size_t) b * 6 + d]);
			mx[d] = std::max(mx[d], hp[(size_t) b * 6 + 3 + d]);
		}
	// inflate the edge a hair so that |x_i - x_j| < cutoff can never straddle two cells after rounding, and coarsen
	// sparse clouds: a coarser grid finds the same pairs
	h *= (1.0 + 1e-9);
	for (int d = 0; d < 3; d++)
		if (!(mx[d] >= mn[d])) mn[d] = mx[d] = 0.0;  // no finite particle at all: one cell
	const double max_cells = 8.0 * cnt + 4096.0;
	for (int tries = 0;; tries++) {
		const double cells = (floor((mx[0] - mn[0]) / h) + 1) * (floor((mx[1] - mn[1]) / h) + 1) * (floor((mx[2] - mn[2]) / h) + 1);
		if (cells <= max_cells) break;
		if (tries > 2000 || !std::isfinite(h)) {  // cannot happen for a finite box; never spin
			h = std::max(std::max(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]) * 2.0 + 1.0;
			break;
		}
		h *= 1.5;
	}
	g.x0 = mn[0]; g.y0 = mn[1]; g.z0 = mn[2];
	g.h = h; g.inv_h = 1.0 / h;
	g.nx = (int) floor((mx[0] - mn[0]) / h) + 1;
	g.ny = (int) floor((mx[1] - mn[1]) / h) + 1;
	g.nz = (int) floor((mx[2] - mn[2]) / h) + 1;
	const size_t nc = (size_t) g.nx * g.ny * g.nz;
	const int nb = (int) ((nc + 1 + SCAN_B * SCAN_ITEMS - 1) / (SCAN_B * SCAN_ITEMS));
	CU_ALLOC(sc.cell_of_p, sizeof(int) * cnt);
	CU_ALLOC(sc.count, sizeof(int) * (nc + 1));
	CU_ALLOC(sc.start, sizeof(int) * (nc + 1));
	CU_ALLOC(sc.fill, sizeof(int) * nc);
	CU_ALLOC(sc.bsum, sizeof(int) * std::max(nb, (cnt + SCAN_B * SCAN_ITEMS - 1) / (SCAN_B * SCAN_ITEMS) + 1));
	CU_ALLOC(sc.total, sizeof(int));
	CU_ALLOC(sc.sorted, sizeof(vec4d) * cnt);
	CU_TRY(cudaMemsetAsync(sc.count, 0, sizeof(int) * (nc + 1), s->stream));
	CU_TRY(cudaMemsetAsync(sc.fill, 0, sizeof(int) * nc, s->stream));
	k_cell_count<<<(cnt + 255) / 256, 256, 0, s->stream>>>(s->posm, lo, hi, g, sc.cell_of_p, sc.count);
	LAUNCH_CHECK();
	int rc = scan_exclusive(s, sc.count, sc.start, (int) nc + 1, sc.bsum, sc.total);
	if (rc) return rc;
	k_cell_scatter<<<(cnt + 255) / 256, 256, 0, s->stream>>>(s->posm, lo, hi, sc.cell_of_p, sc.start, sc.fill, sc.sorted);
	LAUNCH_CHECK();
	return ASS_OK;
}

}  // namespace

extern "C" int ass_connect_springs_dist(ass_sim* s, double max_dist, int lo, int hi, double k, double gamma, const ass_spring* existing,
		int n_existing, ass_spring** out, int* n_out) {
	REQUIRE(s && out && n_out, ASS_EINVAL, "null argument");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	REQUIRE(n_existing >= 0 && (n_existing == 0 || existing), ASS_EINVAL, "bad existing-spring list");
	*out = nullptr;
	*n_out = 0;
	if (hi <= lo) return ASS_OK;  // springs.cpp:116-117
	REQUIRE(lo >= 0 && hi <= s->n, ASS_EINVAL, "bad particle range");
	if (!(max_dist > 0.0)) return ASS_OK;  // a length is never < a non-positive (or NaN) cutoff
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int cnt = hi - lo;
	Scratch sc;
	Grid g;
	ass_tic(s, ASS_K_CONNECT);
	if ((rc = build_grid(s, lo, hi, max_dist, sc, g))) return rc;
	CU_ALLOC(sc.ncount, sizeof(int) * cnt);
	CU_ALLOC(sc.noff, sizeof(int) * cnt);
	k_neighbours<0><<<(cnt + 127) / 128, 128, 0, s->stream>>>(s->posm, lo, hi, g, sc.start, sc.sorted, max_dist, sc.ncount, nullptr, nullptr, nullptr);
	LAUNCH_CHECK();
	if ((rc = scan_exclusive(s, sc.ncount, sc.noff, cnt, sc.bsum, sc.total))) return rc;
	int total = 0;
	CU_TRY(cudaMemcpyAsync(&total, sc.total, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	if (total == 0) { ass_toc(s, ASS_K_CONNECT); return ASS_OK; }
	CU_ALLOC(sc.list, sizeof(int) * (size_t) total);
	CU_ALLOC(sc.springs, sizeof(ass_spring) * (size_t) total);
	k_neighbours<1><<<(cnt + 127) / 128, 128, 0, s->stream>>>(s->posm, lo, hi, g, sc.start, sc.sorted, max_dist, nullptr, sc.noff, sc.list, nullptr);
	LAUNCH_CHECK();
	k_emit_springs<<<(cnt + 127) / 128, 128, 0, s->stream>>>(s->posm, lo, hi, sc.noff, sc.ncount, sc.list, k, gamma, sc.springs);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_CONNECT);
	ass_spring* h = (ass_spring*) malloc(sizeof(ass_spring) * (size_t) total);
	REQUIRE(h, ASS_ENOMEM, "out of host memory");
	cudaError_t e = cudaMemcpyAsync(h, sc.springs, sizeof(ass_spring) * (size_t) total, cudaMemcpyDeviceToHost, s->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
	if (e != cudaSuccess) {
		free(h);
		ass_set_error("connect: copy back -> %s", cudaGetErrorString(e));
		return ASS_ECUDA;
	}
	int kept = total;
	if (n_existing > 0) {
		// add_spring returns the existing index instead of adding (springs.cpp:70-79)
		std::unordered_set<uint64_t> have;
		have.reserve((size_t) n_existing * 2);
		// ... matching (particle_1, particle_2) == (low, high) AS STORED: a list entry written the other way round does not
		// count as a duplicate there, so it does not here
		for (int q = 0; q < n_existing; q++) have.insert(((uint64_t) (uint32_t) existing[q].particle_1 << 32) | (uint32_t) existing[q].particle_2);
		kept = 0;
		for (int q = 0; q < total; q++) {
			const uint64_t key = ((uint64_t) (uint32_t) h[q].particle_1 << 32) | (uint32_t) h[q].particle_2;
			if (have.count(key)) continue;
			h[kept++] = h[q];
		}
	}
	*out = h;
	*n_out = kept;
	return ASS_OK;
}

extern "C" int ass_mindist(ass_sim* s, int lo, int hi, double* out) {
	REQUIRE(s && out, ASS_EINVAL, "null argument");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	REQUIRE(lo >= 0 && hi <= s->n && lo <= hi, ASS_EINVAL, "bad particle range");
	*out = 1.7976931348623157e308;  // DBL_MAX, shapes.cpp:819
	const int cnt = hi - lo;
	if (cnt < 2) return ASS_OK;
	int rc = ass_ensure_device();
	if (rc) return rc;
	// min is exact and order-free: any search that visits the closest pair returns the reference's double.  Visit
	// pairs in adjacent cells; if the best found is below the cell edge no closer pair can hide elsewhere; otherwise
	// coarsen and retry.
	// first guess: build_grid coarsens a vanishing edge up to <= 8 cells per particle, i.e. about half the mean spacing
	double h = 1e-300;
	for (int attempt = 0; attempt < 48; attempt++) {
		Scratch sc;
		Grid g;
		if ((rc = build_grid(s, lo, hi, h, sc, g))) return rc;
		double* d_best = s->d_red + 32;
		const double init = 1.7976931348623157e308;
		CU_TRY(cudaMemcpyAsync(d_best, &init, sizeof(double), cudaMemcpyHostToDevice, s->stream));
		ass_tic(s, ASS_K_CONNECT);
		k_neighbours<2><<<(cnt + 127) / 128, 128, 0, s->stream>>>(s->posm, lo, hi, g, sc.start, sc.sorted, 0.0, nullptr, nullptr, nullptr, d_best);
		LAUNCH_CHECK();
		ass_toc(s, ASS_K_CONNECT);
		double best = 0.0;
		CU_TRY(cudaMemcpyAsync(&best, d_best, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
		CU_TRY(cudaStreamSynchronize(s->stream));
		if (best < g.h * (1.0 - 1e-6)) { *out = best; return ASS_OK; }
		if (g.nx == 1 && g.ny == 1 && g.nz == 1) { *out = best; return ASS_OK; }  // one cell = all pairs seen
		h = g.h * 2.0;
	}
	ass_set_error("ass_mindist: grid search did not converge");
	return ASS_ESTATE;
}
