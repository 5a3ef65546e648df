// shard.cu -- one large body split over the GPUs of a node by contiguous slabs of particles (BASELINE C4).
//
// Rank g owns particles [bounds[g], bounds[g+1]): it is the only rank that integrates them, evaluates gravity onto them
// and sums the springs incident on them.  Every rank keeps arrays for all N particles, because a gravity target needs
// every source and a spring owner needs its neighbours' positions and velocities.  The reference has no counterpart
// (its MPI path is tree-only and never used with springs); the nearest analogue is one all-gather of x and v per step.
//
// Exchange.  posm and vel live in ONE exportable allocation (the "window").  Peers map it (CUDA IPC across processes,
// or the raw pointer inside one process) and each step every rank PULLS the other slabs with the copy engines on a
// side stream while its SMs already evaluate gravity against the sources of its own slab; the rest of the sources
// follow once the pull has landed.  So the all-gather is hidden behind 1/P of the gravity work.  (Reading remote tiles
// from inside the gravity kernel would be the literal "fused" form, but peer loads bypass the local L2: every target
// block would re-fetch every remote slab over NVLink, ~tblocks x the bytes of one gather.)
//
// A step is three calls separated by two barriers that the launcher owns (torch.distributed, or nothing in-process):
//   ass_shard_begin     part1 on the own slab unless the previous finish already pre-drifted it; sync.
//   --- barrier ---     every slab is final
//   ass_shard_exchange  start the pulls + launch gravity(own-slab sources); returns when the pulls have landed.
//   --- barrier ---     nobody is still reading my slab
//   ass_shard_finish    gravity(remote sources), fixed-order row reduction, [zero_accel], spring_forces on the own slab,
//                       part2 (+ next part1) on the own slab; sync.
#include <string.h>

#include <algorithm>

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
	CU_TRY(cudaMalloc(&sh.window, sizeof(vec4d) * 2 * std::max<size_t>(n, 1)));
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
	// 2. meanwhile: gravity onto the own slab from the sources of the own slab
	sh.rows_local = 0;
	if (s->prm.gravity == 1 && s->prm.additional_forces != 2) {
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

extern "C" int ass_shard_finish(ass_sim* s, int pre_drift_next) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const ass_params& P = s->prm;
	const double dt = P.dt;
	CU_TRY(cudaStreamWaitEvent(s->stream, sh.pulled, 0));
	if (P.gravity == 1 && P.additional_forces != 2) {
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
