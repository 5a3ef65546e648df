// shard.cu -- one large body split over the GPUs of a node by contiguous slabs of particles (BASELINE C4).
//
// Rank g owns particles [bounds[g], bounds[g+1]): it is the only rank that integrates them and sums the springs incident
// on them.  Every rank keeps arrays for all N particles, because gravity needs every source and a spring owner needs
// its neighbours' positions and velocities.  The reference has no counterpart (its MPI path is tree-only and never used
// with springs); the nearest analogue is one all-gather of x and v plus one reduce-scatter of partial accelerations.
//
// Gravity is divided by PAIRS OF SLABS so that every unordered particle pair is still evaluated once (gravity_sym.cu)
// and every rank does the same amount of work.  With P ranks and D = (P-1)/2, rank g evaluates
//     slab g x slab g                     (triangle)                         -> its own accelerations
//     slab g x slab (g+d) mod P, d = 1..D (rectangle)                        -> its own, and a partial for slab g+d
//     half of slab g x slab (g+P/2) mod P (P even only: the two ranks of an antipodal pair split its rectangle -- the
//                                          lower rank takes its own lower half against the whole other slab, the higher
//                                          rank its whole slab against the other's upper half)
// and trailing point masses (params.n_perts) are handled one-sidedly by whoever owns the targets.  The partials for
// foreign slabs are left in an exported buffer; their owners pull and add them in a fixed order.
//
// Exchange.  posm, vel and the partial-acceleration buffer live in ONE exportable allocation (the "window").  Peers map
// it (CUDA IPC across processes, or the raw pointer inside one process); every transfer is a PULL by the copy engines on
// a side stream: the peers' slabs while the SMs already run the own-slab triangle, the partials before the springs.
// (Reading remote tiles from inside the gravity kernel would be the literal "fused" form, but peer loads bypass the
// local L2: every tile would re-fetch remote particles over NVLink, hundreds of times the bytes of one gather.)
//
// A step is four calls separated by three barriers that the launcher owns (torch.distributed; launcher.py):
//   ass_shard_begin     part1 on the own slab unless the previous finish already pre-drifted it; sync.
//   --- barrier ---     every slab is final
//   ass_shard_exchange  start the pulls + launch the own-slab triangle; returns when the pulls have landed.
//   --- barrier ---     nobody is still reading my slab
//   ass_shard_gravity   the rectangles, the antipodal half, the point masses; sync (my partials are complete).
//   --- barrier ---     every partial is complete
//   ass_shard_finish    pull + add the partials for the own slab, [zero_accel], spring_forces on the own slab,
//                       part2 (+ next part1) on the own slab; sync.
// With params.n_active restricting the sources the pair-once division does not apply and the ranks fall back to
// one-sided sums over all sources (own slab first, the rest after the pulls).
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"

static int n_sources(const ass_sim* s) { return (s->prm.n_active == -1) ? s->n : std::max(0, std::min(s->prm.n_active, s->n)); }

extern "C" int ass_shard_setup(ass_sim* s, int rank, int nranks, const int* bounds) {
	REQUIRE(s && bounds, ASS_EINVAL, "null argument");
	REQUIRE(s->have_state, ASS_ESTATE, "upload the full body on every rank before sharding");
	REQUIRE(!s->shard.active, ASS_ESTATE, "already sharded");
	REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, ASS_EINVAL, "bad rank");
	REQUIRE(bounds[0] == 0 && bounds[nranks] == s->n, ASS_EINVAL, "bounds must run from 0 to N");
	for (int r = 0; r < nranks; r++) REQUIRE(bounds[r] <= bounds[r + 1], ASS_EINVAL, "bounds must be non-decreasing");
	int rc = ass_ensure_device();
	if (rc) return rc;
	ShardInfo& sh = s->shard;
	const size_t n = (size_t) s->n;
	CU_TRY(cudaMalloc(&sh.window, sizeof(vec4d) * 3 * std::max<size_t>(n, 1)));  // posm | vel | partial accelerations
	CU_TRY(cudaMemsetAsync(sh.window + 2 * n, 0, sizeof(vec4d) * n, s->stream));
	CU_TRY(cudaMemcpyAsync(sh.window, s->posm, sizeof(vec4d) * n, cudaMemcpyDeviceToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(sh.window + n, s->vel, sizeof(vec4d) * n, cudaMemcpyDeviceToDevice, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	cudaFree(s->posm); cudaFree(s->vel); cudaFree(s->posm_alt); cudaFree(s->vel_alt);
	s->posm = sh.window;
	s->vel = sh.window + n;
	s->posm_alt = s->vel_alt = nullptr;  // the fused ping-pong loop is not used by sharded simulations
	s->cap = s->n;
	CU_TRY(cudaStreamCreateWithFlags(&sh.copy_stream, cudaStreamNonBlocking));
	CU_TRY(cudaEventCreateWithFlags(&sh.pulled, cudaEventDisableTiming));
	sh.rank = rank; sh.nranks = nranks;
	sh.bounds.assign(bounds, bounds + nranks + 1);
	sh.lo = bounds[rank]; sh.hi = bounds[rank + 1];
	sh.peer_window.assign(nranks, nullptr);
	sh.peer_is_ipc.assign(nranks, 0);
	sh.active = true;
	s->predrifted = false;
	return ASS_OK;
}

extern "C" int ass_shard_export(ass_sim* s, unsigned char handle[ASS_IPC_HANDLE_BYTES]) {
	REQUIRE(s && handle, ASS_EINVAL, "null argument");
	REQUIRE(s->shard.active, ASS_ESTATE, "call ass_shard_setup first");
	static_assert(sizeof(cudaIpcMemHandle_t) == ASS_IPC_HANDLE_BYTES, "IPC handle size");
	cudaIpcMemHandle_t h;
	CU_TRY(cudaIpcGetMemHandle(&h, s->shard.window));
	memcpy(handle, &h, sizeof(h));
	return ASS_OK;
}

extern "C" int ass_shard_import(ass_sim* s, int peer, const unsigned char handle[ASS_IPC_HANDLE_BYTES]) {
	REQUIRE(s && handle, ASS_EINVAL, "null argument");
	REQUIRE(s->shard.active, ASS_ESTATE, "call ass_shard_setup first");
	REQUIRE(peer >= 0 && peer < s->shard.nranks && peer != s->shard.rank, ASS_EINVAL, "bad peer rank");
	int rc = ass_ensure_device();
	if (rc) return rc;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	void* p = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("ass_shard_import: cudaIpcOpenMemHandle(peer %d) -> %s", peer, cudaGetErrorString(e));
		return ASS_ECOMM;
	}
	s->shard.peer_window[peer] = (vec4d*) p;
	s->shard.peer_is_ipc[peer] = 1;
	return ASS_OK;
}

// Same-process variant (several ass_sim objects in one process, e.g. one host thread per GPU, or tests on one GPU).
extern "C" int ass_shard_import_local(ass_sim* s, int peer, ass_sim* peer_sim) {
	REQUIRE(s && peer_sim, ASS_EINVAL, "null argument");
	REQUIRE(s->shard.active && peer_sim->shard.active, ASS_ESTATE, "call ass_shard_setup on both first");
	REQUIRE(peer >= 0 && peer < s->shard.nranks && peer != s->shard.rank, ASS_EINVAL, "bad peer rank");
	REQUIRE(peer_sim->n == s->n, ASS_EINVAL, "peer holds a different body");
	s->shard.peer_window[peer] = peer_sim->shard.window;
	s->shard.peer_is_ipc[peer] = 0;
	return ASS_OK;
}

extern "C" int ass_shard_begin(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->shard.active, ASS_ESTATE, "not sharded");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const double dt = s->prm.dt;
	if (!s->predrifted) {
		if ((rc = ass_launch_part1_range(s, dt, s->shard.lo, s->shard.hi))) return rc;
	}
	s->predrifted = false;
	s->prm.t += dt / 2.;  // integrator_leapfrog.c:50
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}

// does the pair-once division apply?  (every particle a source; gravity on and not zeroed again by zero_accel)
static bool gravity_on(const ass_sim* s) { return s->prm.gravity == 1 && s->prm.additional_forces != 2; }
static bool pairwise(const ass_sim* s) { return gravity_on(s) && n_sources(s) == s->n; }
// body part of rank r's slab: trailing point masses [N - n_perts, N) are kept out of the pair-once kernels
static void body_slab(const ass_sim* s, int r, int* lo, int* hi) {
	const int nb = s->n - std::max(0, std::min(s->prm.n_perts, s->n));
	*lo = std::min(s->shard.bounds[r], nb);
	*hi = std::min(s->shard.bounds[r + 1], nb);
}

extern "C" int ass_shard_exchange(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const size_t n = (size_t) s->n;
	// 1. pull every peer slab (posm and vel) with the copy engines
	for (int r = 0; r < sh.nranks; r++) {
		if (r == sh.rank) continue;
		const size_t lo = (size_t) sh.bounds[r], cnt = (size_t) (sh.bounds[r + 1] - sh.bounds[r]);
		if (!cnt) continue;
		REQUIRE(sh.peer_window[r], ASS_ESTATE, "a peer window has not been imported");
		CU_TRY(cudaMemcpyAsync(sh.window + lo, sh.peer_window[r] + lo, sizeof(vec4d) * cnt, cudaMemcpyDeviceToDevice, sh.copy_stream));
		CU_TRY(cudaMemcpyAsync(sh.window + n + lo, sh.peer_window[r] + n + lo, sizeof(vec4d) * cnt, cudaMemcpyDeviceToDevice, sh.copy_stream));
	}
	CU_TRY(cudaEventRecord(sh.pulled, sh.copy_stream));
	// 2. meanwhile: the part of gravity that needs the own slab only
	sh.rows_local = 0;
	if (pairwise(s)) {
		int lo, hi;
		body_slab(s, sh.rank, &lo, &hi);
		// a = triangle of the own slab; slabs too small for the pair-once tiles still go through it (forced), so that
		// every rank evaluates exactly the pairs the division assigns to it
		if (hi > lo) {
			if ((rc = ass_launch_gravity_sym(s, lo, hi))) return rc;
		}
		// own particles outside the body part (point masses) start from zero
		if ((rc = ass_launch_zero_accel_range(s, std::max(sh.lo, hi), sh.hi))) return rc;
	} else if (gravity_on(s)) {
		const int ns = n_sources(s);
		const int a = std::min(sh.lo, ns), b = std::min(sh.hi, ns);
		const int rows = ass_gravity_rows(sh.hi - sh.lo, b - a) + ass_gravity_rows(sh.hi - sh.lo, a) + ass_gravity_rows(sh.hi - sh.lo, ns - b);
		if ((rc = ass_buf_reserve(s->grav_partial, sizeof(vec4d) * (size_t) std::max(rows, 1) * n))) return rc;
		if ((rc = ass_launch_gravity_partial(s, s->posm, a, b, sh.lo, sh.hi, 0, &sh.rows_local))) return rc;
	}
	// 3. the caller may enter its barrier once the peers' data has landed here
	CU_TRY(cudaStreamSynchronize(sh.copy_stream));
	return ASS_OK;
}

// where an antipodal pair cuts the lower rank's slab
static int half_point(int lo, int hi) { return lo + (hi - lo) / 2; }

// one-sided: acc[t_lo, t_hi) += pull of sources [s_lo, s_hi)
static int one_sided_add(ass_sim* s, int s_lo, int s_hi, int t_lo, int t_hi) {
	if (t_hi <= t_lo || s_hi <= s_lo) return ASS_OK;
	int rc = ass_buf_reserve(s->grav_partial, sizeof(vec4d) * (size_t) std::max(ass_gravity_rows(t_hi - t_lo, s_hi - s_lo), 1) * (size_t) s->n);
	if (rc) return rc;
	int used = 0;
	if ((rc = ass_launch_gravity_partial(s, s->posm, s_lo, s_hi, t_lo, t_hi, 0, &used))) return rc;
	return ass_launch_gravity_reduce(s, t_lo, t_hi, used, true);
}

extern "C" int ass_shard_gravity(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	int rc = ass_ensure_device();
	if (rc) return rc;
	CU_TRY(cudaStreamWaitEvent(s->stream, sh.pulled, 0));
	if (pairwise(s)) {
		const int P = sh.nranks, g = sh.rank, D = (P - 1) / 2;
		const int n = s->n, nb = n - std::max(0, std::min(s->prm.n_perts, n));
		vec4d* fpart = sh.window + 2 * (size_t) n;
		int lo, hi;
		body_slab(s, g, &lo, &hi);
		for (int d = 1; d <= D; d++) {  // rectangles: both sides of every pair
			int plo, phi;
			body_slab(s, (g + d) % P, &plo, &phi);
			if ((rc = ass_launch_gravity_sym_pair(s, lo, hi, plo, phi, s->acc, true, fpart, false))) return rc;
		}
		if (P % 2 == 0) {  // antipodal slab: the pair's rectangle is split between its two ranks
			const int h = (g + P / 2) % P;
			int plo, phi;
			body_slab(s, h, &plo, &phi);
			if (g < h) rc = ass_launch_gravity_sym_pair(s, lo, half_point(lo, hi), plo, phi, s->acc, true, fpart, false);
			else rc = ass_launch_gravity_sym_pair(s, lo, hi, half_point(plo, phi), phi, s->acc, true, fpart, false);
			if (rc) return rc;
		}
		// point masses: they pull on my body particles; whoever owns them sums everything onto them
		if ((rc = one_sided_add(s, nb, n, lo, hi))) return rc;
		if ((rc = one_sided_add(s, 0, n, std::max(sh.lo, nb), sh.hi))) return rc;
	}
	CU_TRY(cudaStreamSynchronize(s->stream));  // my partials for other slabs are complete
	return ASS_OK;
}

__global__ void k_add_partial(vec4d* __restrict__ acc, const vec4d* __restrict__ part, int lo, int hi) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d a = acc[i];
	const vec4d b = part[i - lo];
	a.x += b.x; a.y += b.y; a.z += b.z;
	acc[i] = a;
}

extern "C" int ass_shard_finish(ass_sim* s, int pre_drift_next) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const ass_params& P = s->prm;
	const double dt = P.dt;
	CU_TRY(cudaStreamWaitEvent(s->stream, sh.pulled, 0));
	if (pairwise(s)) {
		// the partials other ranks computed for my slab, pulled and added in a fixed order (d = 1..D)
		const int R = sh.nranks, g = sh.rank, D = (R - 1) / 2;
		int lo, hi;
		body_slab(s, g, &lo, &hi);
		const size_t cnt = (size_t) std::max(hi - lo, 0), n = (size_t) s->n;
		// (rank, first particle) of every contribution, in the order they are added
		std::vector<std::pair<int, int>> from;
		for (int d = 1; d <= D; d++) from.push_back({ ((g - d) % R + R) % R, lo });
		if (R % 2 == 0) {
			const int h = (g + R / 2) % R;
			from.push_back({ h, g < h ? half_point(lo, hi) : lo });  // the higher rank covered only my upper half
		}
		if (cnt && !from.empty()) {
			if ((rc = ass_buf_reserve(s->grav_partial, sizeof(vec4d) * cnt * from.size()))) return rc;
			vec4d* land = (vec4d*) s->grav_partial.p;
			for (size_t q = 0; q < from.size(); q++) {
				const int r = from[q].first, first = from[q].second;
				if (hi <= first) continue;
				REQUIRE(sh.peer_window[r], ASS_ESTATE, "a peer window has not been imported");
				CU_TRY(cudaMemcpyAsync(land + q * cnt, sh.peer_window[r] + 2 * n + (size_t) first, sizeof(vec4d) * (size_t) (hi - first), cudaMemcpyDeviceToDevice, s->stream));
			}
			ass_tic(s, ASS_K_GRAVITY);
			for (size_t q = 0; q < from.size(); q++) {
				const int first = from[q].second;
				if (hi <= first) continue;
				k_add_partial<<<(unsigned) ((hi - first + 255) / 256), 256, 0, s->stream>>>(s->acc, land + q * cnt, first, hi);
				LAUNCH_CHECK();
			}
			ass_toc(s, ASS_K_GRAVITY);
		}
	} else if (gravity_on(s)) {
		const int ns = n_sources(s);
		const int a = std::min(sh.lo, ns), b = std::min(sh.hi, ns);
		int rows = sh.rows_local, used = 0;
		if ((rc = ass_launch_gravity_partial(s, s->posm, 0, a, sh.lo, sh.hi, rows, &used))) return rc;
		rows += used;
		if ((rc = ass_launch_gravity_partial(s, s->posm, b, ns, sh.lo, sh.hi, rows, &used))) return rc;
		rows += used;
		if ((rc = ass_launch_gravity_reduce(s, sh.lo, sh.hi, rows))) return rc;
	}
	if (P.additional_forces != 0 && s->ns > 0) {
		if ((rc = ass_launch_spring_forces_range(s, P.additional_forces == 2, sh.lo, sh.hi))) return rc;
	} else if (P.additional_forces == 2) {
		if ((rc = ass_launch_zero_accel_range(s, sh.lo, sh.hi))) return rc;
	}
	if ((rc = ass_launch_part2_range(s, dt, pre_drift_next != 0, sh.lo, sh.hi))) return rc;
	s->predrifted = pre_drift_next != 0;
	s->prm.t += dt / 2.;  // integrator_leapfrog.c:65-66
	s->prm.dt_last_done = dt;
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}

// Bring every slab home (positions, velocities) so that ass_download_particles returns the whole body; accelerations of
// foreign slabs are pulled by the caller rank-by-rank if it wants them (each rank only ever computes its own).
extern "C" int ass_shard_gather(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	REQUIRE(!s->predrifted, ASS_ESTATE, "finish the step without pre_drift_next before gathering");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const size_t n = (size_t) s->n;
	for (int r = 0; r < sh.nranks; r++) {
		if (r == sh.rank) continue;
		const size_t lo = (size_t) sh.bounds[r], cnt = (size_t) (sh.bounds[r + 1] - sh.bounds[r]);
		if (!cnt) continue;
		REQUIRE(sh.peer_window[r], ASS_ESTATE, "a peer window has not been imported");
		CU_TRY(cudaMemcpyAsync(sh.window + lo, sh.peer_window[r] + lo, sizeof(vec4d) * cnt, cudaMemcpyDeviceToDevice, sh.copy_stream));
		CU_TRY(cudaMemcpyAsync(sh.window + n + lo, sh.peer_window[r] + n + lo, sizeof(vec4d) * cnt, cudaMemcpyDeviceToDevice, sh.copy_stream));
	}
	CU_TRY(cudaStreamSynchronize(sh.copy_stream));
	return ASS_OK;
}

void ass_shard_release(ass_sim* s) {
	ShardInfo& sh = s->shard;
	if (!sh.active) return;
	for (int r = 0; r < (int) sh.peer_window.size(); r++)
		if (sh.peer_window[r] && sh.peer_is_ipc[r]) cudaIpcCloseMemHandle(sh.peer_window[r]);
	if (sh.copy_stream) cudaStreamDestroy(sh.copy_stream);
	if (sh.pulled) cudaEventDestroy(sh.pulled);
	// posm / vel alias the window: free it once and clear the aliases
	cudaFree(sh.window);
	s->posm = s->vel = nullptr;
	sh.window = nullptr;
	sh.active = false;
}
