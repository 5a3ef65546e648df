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
// foreign slabs are left in the exported window; their owners pull and add them in a fixed order.
//
// THE HOST IS NOT IN THE STEP.  ass_shard_step queues nsteps whole steps on the rank's two streams and returns; ranks
// meet on the DEVICE through 64-bit epoch flags that live in every rank's window and are written by the producers
// straight into their peers' memory (st.release.sys over NVLink) and polled locally by the consumers (ld.acquire.sys):
//     POS[r]  = e   rank r's slab carries the (drifted) positions and velocities of step e
//     PART[r] = e   rank r's partial accelerations of step e for the other slabs are complete
// The window is ONE exportable allocation  [posm0 | vel0 | posm1 | vel1 | partials | flags]  mapped by every peer (CUDA
// IPC across processes, the raw pointer inside one process).  Everything that belongs to step e lives in buffer e & 1
// and the kick of step e writes the own slab of buffer (e+1) & 1 (the ping-pong the unsharded fused loop uses too), so a
// rank that runs ahead never overwrites what a slower peer still reads: buffer b is next written during step e+1, which
// cannot start anywhere before every rank has published POS(e+1), i.e. has finished step e and with it every read of
// buffer b's previous contents.  The partial buffer is single: it is rewritten by the rectangles of step e+1, which wait
// for POS(e+1) of the slab's owner, set after that owner pulled the partials of step e.
// One step of rank g, epoch e, buffer b = e & 1:
//     main stream                                            copy stream
//     [part1 on the own slab, if not pre-drifted]
//     POS[g] := e  (into every window)
//     triangle of the own slab                               wait POS[q] >= e for all q; pull slab q of posm_b, vel_b
//     wait for the pulls                                       from q's window (copy engines: one all-gather)
//     rectangles, antipodal half, point masses
//     PART[g] := e (into every window)
//     wait PART[q] >= e for the ranks that worked for me; pull their partials, add them in a fixed order
//     mirror; springs + kick + drift (+ next part1) for the own slab -> buffer b ^ 1
// With params.n_active restricting the sources the pair-once division does not apply and the ranks fall back to
// one-sided sums over all sources (own slab first, the rest after the pulls); the PART phase is then empty.
// (Reading remote tiles from inside the gravity kernel would be the literal "fused" form, but peer loads bypass the
// local L2: every tile would re-fetch remote particles over NVLink, hundreds of times the bytes of one gather.)
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

constexpr unsigned long long SPIN_LIMIT_NS = 30ULL * 1000000000ULL;  // a peer that never arrives: give up, flag the error

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

struct FlagPtrs {
	unsigned long long* w[ASS_SHARD_MAX_RANKS];  // every rank's flag block (own included), as mapped here
};

// thread q publishes  flags_of_rank_q[phase][me] = epoch.  Everything this stream did before is complete (stream order)
// and made visible system-wide by the fence before the releasing store.
__global__ void k_flag_set(FlagPtrs fp, int nranks, int phase, int me, unsigned long long epoch) {
	const int q = threadIdx.x;
	if (q >= nranks || !fp.w[q]) return;
	__threadfence_system();
	st_release_sys(fp.w[q] + phase * ASS_SHARD_MAX_RANKS + me, epoch);
}

// thread q waits until  my_flags[phase][q] >= epoch  for every q in `mask`; polls local memory only
__global__ void k_flag_wait(const unsigned long long* flags, unsigned mask, int phase, unsigned long long epoch, int* err) {
	const int q = threadIdx.x;
	if (!((mask >> q) & 1u)) return;
	const unsigned long long* f = flags + phase * ASS_SHARD_MAX_RANKS + q;
	const unsigned long long t0 = global_ns();
	while (ld_acquire_sys(f) < epoch) {
		__nanosleep(100);
		if (global_ns() - t0 > SPIN_LIMIT_NS) { atomicExch(err, 1 + phase); return; }
	}
}

// Slab transfers between ranks of ONE process on ONE device (tests, several ass_sim per GPU) are done by a kernel: the
// copy engines of a device serve their queue in order, and a copy of rank A that waits (stream order) behind A's flag
// wait would hold up the copy of rank B that A is waiting for.  Ranks in different processes own their GPU's copy
// engines, whose queue then only ever holds work that depends on the PEERS' progress: no cycle.
// publish one flag, then wait for others, in one launch (the PART phase: a rank that has finished its rectangles says so
// and at once starts waiting for the ranks that worked for it)
__global__ void k_flag_set_wait(FlagPtrs fp, int nranks, int phase, int me, unsigned long long epoch, const unsigned long long* flags, unsigned mask, int* err) {
	const int q = threadIdx.x;
	if (q < nranks && fp.w[q]) {
		__threadfence_system();
		st_release_sys(fp.w[q] + phase * ASS_SHARD_MAX_RANKS + me, epoch);
	}
	if (!((mask >> q) & 1u)) return;
	const unsigned long long* f = flags + phase * ASS_SHARD_MAX_RANKS + q;
	const unsigned long long t0 = global_ns();
	while (ld_acquire_sys(f) < epoch) {
		__nanosleep(100);
		if (global_ns() - t0 > SPIN_LIMIT_NS) { atomicExch(err, 1 + phase); return; }
	}
}

__global__ void k_copy_vec4d(vec4d* __restrict__ dst, const vec4d* __restrict__ src, size_t n) {
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// The reduce-scatter half of the exchange, fused with its use: acc[i] += the partial accelerations the peers computed for
// particle i, read straight out of the peers' windows (peer loads over NVLink: 32 B per particle and source, no staging
// copy, no per-source launch) and added in the fixed order of `src`.  src[q] is indexed by absolute particle number and
// covers [first[q], hi).
struct PartialSources {
	const vec4d* src[ASS_SHARD_MAX_RANKS];
	int first[ASS_SHARD_MAX_RANKS];
	int n;
};
__global__ void k_pull_add_partials(vec4d* __restrict__ acc, PartialSources ps, int lo, int hi) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d a = acc[i];
	for (int q = 0; q < ps.n; q++) {
		if (i < ps.first[q]) continue;
		const double2 b0 = __ldcg((const double2*) (ps.src[q] + i)), b1 = __ldcg((const double2*) (ps.src[q] + i) + 1);
		a.x += b0.x; a.y += b0.y; a.z += b1.x;
	}
	acc[i] = a;
}

// reb_integrator_leapfrog_part2 (src/integrator_leapfrog.c:52-67) for bodies without springs: out of buffer b into b ^ 1
__global__ void k_part2_pingpong(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const vec4d* __restrict__ acc,
		vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, int lo, int hi, double dt, int pre_drift_next) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const double h = __dmul_rn(0.5, dt);
	vec4d p = posm[i];
	vec4d v = vel[i];
	const vec4d a = acc[i];
	v.x = __dadd_rn(v.x, __dmul_rn(dt, a.x));
	v.y = __dadd_rn(v.y, __dmul_rn(dt, a.y));
	v.z = __dadd_rn(v.z, __dmul_rn(dt, a.z));
	p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
	p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
	p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
	if (pre_drift_next) {
		p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
		p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
		p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
	}
	vel_out[i] = v;
	posm_out[i] = p;
}

int n_sources(const ass_sim* s) { return (s->prm.n_active == -1) ? s->n : std::max(0, std::min(s->prm.n_active, s->n)); }
// does the pair-once division apply?  (every particle a source; gravity on and not zeroed again by zero_accel)
bool gravity_on(const ass_sim* s) { return s->prm.gravity == 1 && s->prm.additional_forces != 2; }
bool pairwise(const ass_sim* s) { return gravity_on(s) && n_sources(s) == s->n; }
// body part of rank r's slab: trailing point masses [N - n_perts, N) are kept out of the pair-once kernels
// where an antipodal pair cuts the lower rank's slab
int half_point(int lo, int hi) { return lo + (hi - lo) / 2; }
// slab r of a body of nb particles cut at `bounds` (trailing point masses are not part of the body)
void slab_of(const int* bounds, int nb, int r, int* lo, int* hi) {
	*lo = std::min(bounds[r], nb);
	*hi = std::min(bounds[r + 1], nb);
}
void body_slab(const ass_sim* s, int r, int* lo, int* hi) {
	slab_of(s->shard.bounds.data(), s->n - std::max(0, std::min(s->prm.n_perts, s->n)), r, lo, hi);
}

// layout of the window, in vec4d units
size_t win_posm(const ass_sim* s, int b) { return (size_t) b * 2 * (size_t) s->n; }
size_t win_vel(const ass_sim* s, int b) { return ((size_t) b * 2 + 1) * (size_t) s->n; }
size_t win_part(const ass_sim* s) { return 4 * (size_t) s->n; }
size_t win_flags(const ass_sim* s) { return 5 * (size_t) std::max(s->n, 1); }
constexpr size_t FLAG_VEC4 = (ASS_SHARD_PHASES * ASS_SHARD_MAX_RANKS * sizeof(unsigned long long) + 64 + sizeof(vec4d) - 1) / sizeof(vec4d);
unsigned long long* flags_of(const ass_sim* s, vec4d* window) { return (unsigned long long*) (window + win_flags(s)); }
int* err_of(const ass_sim* s, vec4d* window) { return (int*) (flags_of(s, window) + ASS_SHARD_PHASES * ASS_SHARD_MAX_RANKS); }

// The slabs rank g meets as whole rectangles, (g+1 .. g+D) mod P, as contiguous particle runs: one run, or two when the
// sequence wraps past the last slab.  One launch per run instead of one per slab: 3 x fewer launches and reductions at
// P = 8, and a run of 3 slabs has enough tiles to fill the machine with tiles cut in 2 instead of 8 (a CTA that meets
// only 128 streamed particles spends a tenth of its time loading and storing its 1024 residents).
int rect_runs_of(int P, int g, const int* bounds, int nb, int (&run_lo)[2], int (&run_hi)[2]) {
	const int D = (P - 1) / 2;
	if (D <= 0) return 0;
	int lo, hi, n = 0;
	const int first = g + 1, last = g + D;  // slab numbers before the wrap
	if (last < P) {
		slab_of(bounds, nb, first, &run_lo[0], &hi); slab_of(bounds, nb, last, &lo, &run_hi[0]);
		n = 1;
	} else {
		if (first < P) { slab_of(bounds, nb, first, &run_lo[n], &hi); slab_of(bounds, nb, P - 1, &lo, &run_hi[n]); n++; }
		slab_of(bounds, nb, first < P ? 0 : first - P, &run_lo[n], &hi); slab_of(bounds, nb, last - P, &lo, &run_hi[n]); n++;
	}
	return n;
}
int rect_runs(const ass_sim* s, int (&run_lo)[2], int (&run_hi)[2]) {
	return rect_runs_of(s->shard.nranks, s->shard.rank, s->shard.bounds.data(), s->n - std::max(0, std::min(s->prm.n_perts, s->n)), run_lo, run_hi);
}

// Rank g's pair-once pieces {resident lo, hi, streamed lo, hi}: [0] the triangle of its own slab, then the rectangles against
// the runs of the next D slabs, then (even P) its half of the antipodal rectangle.  Empty pieces are kept (callers skip them).
int shard_pieces_of(int P, int g, const int* bounds, int nb, int (&pc)[4][4]) {
	int lo, hi, npc = 0;
	slab_of(bounds, nb, g, &lo, &hi);
	pc[npc][0] = lo; pc[npc][1] = hi; pc[npc][2] = lo; pc[npc][3] = hi; npc++;
	int run_lo[2], run_hi[2];
	const int runs = rect_runs_of(P, g, bounds, nb, run_lo, run_hi);
	for (int q = 0; q < runs; q++) { pc[npc][0] = lo; pc[npc][1] = hi; pc[npc][2] = run_lo[q]; pc[npc][3] = run_hi[q]; npc++; }
	if (P % 2 == 0 && P > 1) {  // antipodal slab: the pair's rectangle is split between its two ranks
		const int h = (g + P / 2) % P;
		int plo, phi;
		slab_of(bounds, nb, h, &plo, &phi);
		if (g < h) { pc[npc][0] = lo; pc[npc][1] = half_point(lo, hi); pc[npc][2] = plo; pc[npc][3] = phi; }
		else { pc[npc][0] = lo; pc[npc][1] = hi; pc[npc][2] = half_point(plo, phi); pc[npc][3] = phi; }
		npc++;
	}
	return npc;
}

// (rank, first particle) of every partial the peers compute for rank g's slab, in the order they are added
void partial_sources_of(int R, int g, const int* bounds, int nb, std::vector<std::pair<int, int>>& from) {
	const int D = (R - 1) / 2;
	int lo, hi;
	slab_of(bounds, nb, g, &lo, &hi);
	from.clear();
	for (int d = 1; d <= D; d++) from.push_back({ ((g - d) % R + R) % R, lo });
	if (R % 2 == 0 && R > 1) {
		const int h = (g + R / 2) % R;
		from.push_back({ h, g < h ? half_point(lo, hi) : lo });  // the higher rank covered only my upper half
	}
}
void partial_sources(const ass_sim* s, std::vector<std::pair<int, int>>& from) {
	partial_sources_of(s->shard.nranks, s->shard.rank, s->shard.bounds.data(), s->n - std::max(0, std::min(s->prm.n_perts, s->n)), from);
}

// dst (mine) <- src (in peer r's window): cnt vec4d
int pull(ass_sim* s, cudaStream_t st, int r, vec4d* dst, const vec4d* src, size_t cnt) {
	if (!cnt) return ASS_OK;
	if (s->shard.peer_is_ipc[r]) {
		CU_TRY(cudaMemcpyAsync(dst, src, sizeof(vec4d) * cnt, cudaMemcpyDeviceToDevice, st));
	} else {
		const unsigned grid = (unsigned) std::min<size_t>((cnt + 255) / 256, (size_t) ass_sm_count() * 4);
		k_copy_vec4d<<<grid, 256, 0, st>>>(dst, src, cnt);
		LAUNCH_CHECK();
	}
	return ASS_OK;
}

int launch_flag_set(ass_sim* s, cudaStream_t st, int phase, unsigned long long epoch) {
	ShardInfo& sh = s->shard;
	FlagPtrs fp;
	memset(&fp, 0, sizeof(fp));
	for (int q = 0; q < sh.nranks; q++) {
		vec4d* w = (q == sh.rank) ? sh.window : sh.peer_window[q];
		REQUIRE(w, ASS_ESTATE, "a peer window has not been imported");
		fp.w[q] = flags_of(s, w);
	}
	k_flag_set<<<1, 32, 0, st>>>(fp, sh.nranks, phase, sh.rank, epoch);
	LAUNCH_CHECK();
	return ASS_OK;
}

int launch_flag_wait(ass_sim* s, cudaStream_t st, unsigned mask, int phase, unsigned long long epoch) {
	if (!mask) return ASS_OK;
	k_flag_wait<<<1, 32, 0, st>>>(flags_of(s, s->shard.window), mask, phase, epoch, err_of(s, s->shard.window));
	LAUNCH_CHECK();
	return ASS_OK;
}

// Everything a step may have to allocate or synchronise for -- pair-once plans (tile tables, mass check), scratch rows,
// the springs' equal-mass verdict -- BEFORE anything of the step is queued: a step in flight waits on peers, and a host
// that blocks on its own stream at that point would deadlock ranks that share the host thread (tests) and stall the
// others.
int shard_prepare(ass_sim* s) {
	ShardInfo& sh = s->shard;
	int rc;
	const int n = s->n;
	size_t scratch_rows_bytes = 0;
	if (pairwise(s)) {
		const int P = sh.nranks, g = sh.rank;
		const int nb = n - std::max(0, std::min(s->prm.n_perts, n));
		int lo, hi;
		body_slab(s, g, &lo, &hi);
		int pc[4][4];
		const int npc = shard_pieces_of(P, g, sh.bounds.data(), nb, pc);
		for (int q = 0; q < npc; q++)
			if (pc[q][1] > pc[q][0] && pc[q][3] > pc[q][2] && (rc = ass_gravity_sym_prepare(s, pc[q][0], pc[q][1], pc[q][2], pc[q][3]))) return rc;
		const int rows = std::max(ass_gravity_rows(hi - lo, n - nb), ass_gravity_rows(sh.hi - std::max(sh.lo, nb), n));
		scratch_rows_bytes = sizeof(vec4d) * (size_t) std::max(rows, 1) * (size_t) n;
	} else if (gravity_on(s)) {
		const int ns = n_sources(s);
		const int a = std::min(sh.lo, ns), b = std::min(sh.hi, ns);
		const int rows = ass_gravity_rows(sh.hi - sh.lo, b - a) + ass_gravity_rows(sh.hi - sh.lo, a) + ass_gravity_rows(sh.hi - sh.lo, ns - b);
		scratch_rows_bytes = sizeof(vec4d) * (size_t) std::max(rows, 1) * (size_t) n;
	}
	if (scratch_rows_bytes && (rc = ass_buf_reserve(s->grav_partial, scratch_rows_bytes))) return rc;
	if (s->prm.additional_forces != 0 && s->ns > 0 && (rc = ass_springs_prepare(s))) return rc;
	return ASS_OK;
}

// one-sided: acc[t_lo, t_hi) += pull of sources [s_lo, s_hi); scratch rows reserved by shard_prepare
int one_sided_add(ass_sim* s, int s_lo, int s_hi, int t_lo, int t_hi) {
	if (t_hi <= t_lo || s_hi <= s_lo) return ASS_OK;
	int used = 0, rc;
	if ((rc = ass_launch_gravity_partial(s, s->posm, s_lo, s_hi, t_lo, t_hi, 0, &used))) return rc;
	return ass_launch_gravity_reduce(s, t_lo, t_hi, used, true);
}

// posm / vel follow the current buffer, the alt pointers the other one
void point_at(ass_sim* s, int b) {
	ShardInfo& sh = s->shard;
	sh.cur = b;
	s->posm = sh.window + win_posm(s, b);
	s->vel = sh.window + win_vel(s, b);
	s->posm_alt = sh.window + win_posm(s, b ^ 1);
	s->vel_alt = sh.window + win_vel(s, b ^ 1);
}

int check_device_error(ass_sim* s) {
	int e = 0;
	CU_TRY(cudaMemcpy(&e, err_of(s, s->shard.window), sizeof(int), cudaMemcpyDeviceToHost));
	if (e) {
		ass_set_error("sharded step: rank %d waited more than %llu s for a peer's %s flag (a rank died, or the ranks did not call ass_shard_step with the same step counts)",
				s->shard.rank, SPIN_LIMIT_NS / 1000000000ULL, e == 1 + ASS_SHARD_POS ? "positions" : e == 1 + ASS_SHARD_PART ? "partials" : "gather");
		return ASS_ECOMM;
	}
	return ASS_OK;
}

}  // namespace

// Host-only self-check of how gravity is divided by pairs of slabs (no device is touched): over all ranks, every unordered
// pair of body particles must be evaluated by exactly one piece, and the partials a rank adds must be exactly the ones its
// peers compute for its slab.  Checked on segments (the slabs cut at the antipodal half points).  stats: [0] pieces with
// work, [1] segments, [2] violations (must be 0), [3] largest / [4] smallest number of pairs a rank evaluates.
extern "C" int ass_selfcheck_shard_division(int nranks, const int* bounds, int n_body, long long* stats) {
	REQUIRE(stats && bounds && nranks >= 1 && nranks <= ASS_SHARD_MAX_RANKS && n_body >= 0, ASS_EINVAL, "bad argument");
	for (int r = 0; r < nranks; r++) REQUIRE(bounds[r] <= bounds[r + 1], ASS_EINVAL, "slab bounds must not decrease");
	REQUIRE(bounds[0] == 0 && bounds[nranks] >= n_body, ASS_EINVAL, "slabs must cover the body");
	const int P = nranks, nb = n_body;
	std::vector<int> cutp;
	for (int r = 0; r <= P; r++) cutp.push_back(std::min(bounds[r], nb));
	for (int r = 0; r < P; r++) {
		int lo, hi;
		slab_of(bounds, nb, r, &lo, &hi);
		cutp.push_back(half_point(lo, hi));
	}
	std::sort(cutp.begin(), cutp.end());
	cutp.erase(std::unique(cutp.begin(), cutp.end()), cutp.end());
	const int nseg = (int) cutp.size() - 1;
	auto seg_of = [&](int x) { return (int) (std::lower_bound(cutp.begin(), cutp.end(), x) - cutp.begin()); };
	long long viol = 0, pieces = 0, most = 0, least = -1;
	std::vector<int> cover((size_t) std::max(nseg, 0) * std::max(nseg, 0), 0);
	for (int g = 0; g < P; g++) {
		int pc[4][4];
		const int npc = shard_pieces_of(P, g, bounds, nb, pc);
		long long pairs = 0;
		for (int q = 0; q < npc; q++) {
			const int rlo = pc[q][0], rhi = pc[q][1], slo = pc[q][2], shi = pc[q][3];
			if (rhi <= rlo || shi <= slo) continue;
			pieces++;
			const bool tri = rlo == slo && rhi == shi;
			if (!tri && !(rhi <= slo || shi <= rlo)) viol++;  // a rectangle's ranges must be disjoint
			if (cutp[(size_t) seg_of(rlo)] != rlo || cutp[(size_t) seg_of(rhi)] != rhi || cutp[(size_t) seg_of(slo)] != slo || cutp[(size_t) seg_of(shi)] != shi) { viol++; continue; }
			pairs += tri ? (long long) (rhi - rlo) * (rhi - rlo - 1) / 2 : (long long) (rhi - rlo) * (shi - slo);
			for (int a = seg_of(rlo); a < seg_of(rhi); a++)
				for (int b = seg_of(slo); b < seg_of(shi); b++) {
					if (tri && b < a) continue;
					cover[(size_t) std::min(a, b) * nseg + std::max(a, b)]++;
				}
			// whoever owns a slab the streamed range reaches into must expect this rank's partial, from the right particle on
			if (!tri)
				for (int h = 0; h < P; h++) {
					int hlo, hhi;
					slab_of(bounds, nb, h, &hlo, &hhi);
					const int first = std::max(hlo, slo), last = std::min(hhi, shi);
					if (last <= first) continue;
					std::vector<std::pair<int, int>> from;
					partial_sources_of(P, h, bounds, nb, from);
					int hits = 0;
					for (auto& f : from)
						if (f.first == g) { hits++; if (f.second != first || last != hhi) viol++; }
					if (hits != 1) viol++;
				}
		}
		most = std::max(most, pairs);
		least = least < 0 ? pairs : std::min(least, pairs);
		// and nobody is expected who computes nothing for this slab
		std::vector<std::pair<int, int>> from;
		partial_sources_of(P, g, bounds, nb, from);
		int glo, ghi;
		slab_of(bounds, nb, g, &glo, &ghi);
		for (auto& f : from) {
			if (ghi <= f.second) continue;  // (nothing of the slab left from there on: the step skips it too)
			// the source's piece that reaches into this slab must start where this rank starts adding -- unless the source has
			// nothing to compute there (an empty slab or half slab of its own): its partials are the zeros the window was
			// created with, and adding them changes nothing
			int pc2[4][4];
			const int n2 = shard_pieces_of(P, f.first, bounds, nb, pc2);
			int with_work = 0, matching = 0;
			for (int q = 1; q < n2; q++) {
				if (pc2[q][1] <= pc2[q][0] || std::min(pc2[q][3], ghi) <= std::max(pc2[q][2], glo)) continue;
				with_work++;
				if (std::max(pc2[q][2], glo) == f.second && std::min(pc2[q][3], ghi) == ghi) matching++;
			}
			if (with_work != matching || with_work > 1) viol++;
		}
	}
	for (int a = 0; a < nseg; a++)
		for (int b = a; b < nseg; b++) {
			const bool real = cutp[(size_t) a + 1] > cutp[(size_t) a] && cutp[(size_t) b + 1] > cutp[(size_t) b] && !(a == b && cutp[(size_t) a + 1] - cutp[(size_t) a] < 2);
			const int c = cover[(size_t) a * nseg + b];
			if (real ? c != 1 : c > 1) viol++;
		}
	stats[0] = pieces; stats[1] = nseg; stats[2] = viol; stats[3] = most; stats[4] = std::max(least, 0LL);
	return ASS_OK;
}

extern "C" int ass_shard_setup(ass_sim* s, int rank, int nranks, const int* bounds) {
	REQUIRE(s && bounds, ASS_EINVAL, "null argument");
	REQUIRE(s->have_state, ASS_ESTATE, "upload the full body on every rank before sharding");
	REQUIRE(!s->shard.active, ASS_ESTATE, "already sharded");
	REQUIRE(nranks >= 1 && nranks <= ASS_SHARD_MAX_RANKS && rank >= 0 && rank < nranks, ASS_EINVAL, "bad rank (at most 16 ranks)");
	REQUIRE(bounds[0] == 0 && bounds[nranks] == s->n, ASS_EINVAL, "bounds must run from 0 to N");
	for (int r = 0; r < nranks; r++) REQUIRE(bounds[r] <= bounds[r + 1], ASS_EINVAL, "bounds must be non-decreasing");
	int rc = ass_ensure_device();
	if (rc) return rc;
	ShardInfo& sh = s->shard;
	const size_t n = (size_t) s->n;
	const size_t total = win_flags(s) + FLAG_VEC4;
	CU_TRY(cudaMalloc(&sh.window, sizeof(vec4d) * total));
	CU_TRY(cudaMemsetAsync(sh.window, 0, sizeof(vec4d) * total, s->stream));
	for (int b = 0; b < 2; b++) {
		CU_TRY(cudaMemcpyAsync(sh.window + win_posm(s, b), s->posm, sizeof(vec4d) * n, cudaMemcpyDeviceToDevice, s->stream));
		CU_TRY(cudaMemcpyAsync(sh.window + win_vel(s, b), s->vel, sizeof(vec4d) * n, cudaMemcpyDeviceToDevice, s->stream));
	}
	CU_TRY(cudaStreamSynchronize(s->stream));
	cudaFree(s->posm); cudaFree(s->vel); cudaFree(s->posm_alt); cudaFree(s->vel_alt);
	s->cap = s->n;
	CU_TRY(cudaStreamCreateWithFlags(&sh.copy_stream, cudaStreamNonBlocking));
	CU_TRY(cudaEventCreateWithFlags(&sh.pulled, cudaEventDisableTiming));
	CU_TRY(cudaEventCreateWithFlags(&sh.drifted, cudaEventDisableTiming));
	for (int q = 0; q < 3; q++) {
		CU_TRY(cudaStreamCreateWithFlags(&sh.grav_stream[q], cudaStreamNonBlocking));
		CU_TRY(cudaEventCreateWithFlags(&sh.grav_done[q], cudaEventDisableTiming));
	}
	sh.rank = rank; sh.nranks = nranks;
	sh.bounds.assign(bounds, bounds + nranks + 1);
	sh.lo = bounds[rank]; sh.hi = bounds[rank + 1];
	sh.peer_window.assign(nranks, nullptr);
	sh.peer_is_ipc.assign(nranks, 0);
	sh.epoch = 0;
	sh.gather_epoch = 0;
	sh.active = true;
	point_at(s, 0);
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
	// With lazy module loading (the CUDA 12 default) the FIRST launch of a kernel loads it, and loading synchronises the
	// context: it would wait for every rank's flag waits already in flight -- which wait for ranks the same host thread
	// has not queued yet.  Ranks in separate processes are immune (a load only ever waits for the peers' own progress).
	const char* ml = getenv("CUDA_MODULE_LOADING");
	REQUIRE(ml && strcmp(ml, "EAGER") == 0, ASS_ESTATE,
	        "same-process ranks need CUDA_MODULE_LOADING=EAGER in the environment before CUDA starts (a lazily loaded kernel's first launch synchronises the context and would deadlock against a peer's flag wait)");
	s->shard.peer_window[peer] = peer_sim->shard.window;
	s->shard.peer_is_ipc[peer] = 0;
	return ASS_OK;
}

// reb_step x nsteps for the own slab; queued, not awaited (ass_sync / ass_shard_gather / a download wait for it).
// Collective: every rank calls it with the same nsteps and pre_drift_last.  pre_drift_last != 0 leaves the positions with
// the following step's first half-drift already applied (only stepping may follow).
extern "C" int ass_shard_step(ass_sim* s, int nsteps, int pre_drift_last) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	REQUIRE(nsteps >= 0, ASS_EINVAL, "negative step count");
	int rc = ass_ensure_device();
	if (rc) return rc;
	for (int q = 0; q < sh.nranks; q++) REQUIRE(q == sh.rank || sh.peer_window[q], ASS_ESTATE, "a peer window has not been imported");
	if (nsteps == 0) return ASS_OK;
	if ((rc = shard_prepare(s))) return rc;
	const int P = sh.nranks, g = sh.rank, D = (P - 1) / 2;
	const int n = s->n;
	const unsigned all_others = ((1u << P) - 1u) & ~(1u << g);
	const bool have_springs = s->prm.additional_forces != 0 && s->ns > 0;
	std::vector<std::pair<int, int>> from;
	for (int step = 0; step < nsteps; step++) {
		const double dt = s->prm.dt;
		const bool pre = (step < nsteps - 1) || pre_drift_last != 0;
		const unsigned long long e = ++sh.epoch;
		const int b = sh.cur;
		if (sh.gather_pending) {  // a peer may still be pulling my slab for the last gather: it must not move yet
			if ((rc = launch_flag_wait(s, s->stream, all_others, ASS_SHARD_GATHER, sh.gather_pending))) return rc;
			sh.gather_pending = 0;
		}
		// ---- drifted positions of the own slab, published ----
		if (!s->predrifted && (rc = ass_launch_part1_range(s, dt, sh.lo, sh.hi))) return rc;
		s->predrifted = false;
		s->prm.t += dt / 2.;  // integrator_leapfrog.c:50
		if ((rc = launch_flag_set(s, s->stream, ASS_SHARD_POS, e))) return rc;
		CU_TRY(cudaEventRecord(sh.drifted, s->stream));  // own slab drifted, last step's reductions done: the gravity streams may go
		// ---- copy stream: the other slabs of this step, as soon as their owners have published them ----
		if ((rc = launch_flag_wait(s, sh.copy_stream, all_others, ASS_SHARD_POS, e))) return rc;
		for (int r = 0; r < P; r++) {
			if (r == g) continue;
			const size_t lo = (size_t) sh.bounds[r], cnt = (size_t) (sh.bounds[r + 1] - sh.bounds[r]);
			if (!cnt) continue;
			if ((rc = pull(s, sh.copy_stream, r, sh.window + win_posm(s, b) + lo, sh.peer_window[r] + win_posm(s, b) + lo, cnt))) return rc;
			if ((rc = pull(s, sh.copy_stream, r, sh.window + win_vel(s, b) + lo, sh.peer_window[r] + win_vel(s, b) + lo, cnt))) return rc;
		}
		CU_TRY(cudaEventRecord(sh.pulled, sh.copy_stream));
		// ---- meanwhile: the part of gravity that needs the own slab only ----
		int rows_local = 0;
		if (pairwise(s)) {
			int lo, hi;
			body_slab(s, g, &lo, &hi);
			// a = triangle of the own slab; slabs too small for the pair-once tiles still go through it (forced), so that
			// every rank evaluates exactly the pairs the division assigns to it
			if (hi > lo && (rc = ass_launch_gravity_sym(s, lo, hi))) return rc;
			// own particles outside the body part (point masses) start from zero
			if ((rc = ass_launch_zero_accel_range(s, std::max(sh.lo, hi), sh.hi))) return rc;
		} else if (gravity_on(s)) {
			const int ns = n_sources(s);
			const int a = std::min(sh.lo, ns), c = std::min(sh.hi, ns);
			if ((rc = ass_launch_gravity_partial(s, s->posm, a, c, sh.lo, sh.hi, 0, &rows_local))) return rc;
		}
		CU_TRY(cudaStreamWaitEvent(s->stream, sh.pulled, 0));
		// ---- gravity against the other slabs ----
		if (pairwise(s)) {
			const int nb = n - std::max(0, std::min(s->prm.n_perts, n));
			vec4d* fpart = sh.window + win_part(s);
			int lo, hi;
			body_slab(s, g, &lo, &hi);
			// rectangles (both sides of every pair) and, for an even rank count, this rank's half of the antipodal rectangle.
			// Their pair kernels run on streams of their own, next to the triangle still running on the main stream: each of
			// these launches ends in a partly filled round of CTAs (a 12 288-particle slab is 2.1 / 2.9 / 1.9 rounds), and the
			// block scheduler tops a draining kernel up with the next one's CTAs (measured on one GPU, tools/
			// prof_gravity_pairs.py: 889 -> 815 us per rank and step at C3 / 8).  The fixed-order reductions follow on the
			// main stream.  ASS_SHARD_SERIAL_GRAVITY=1: one after the other on the main stream.
			int all[4][4], pc[3][4];
			const int nall = shard_pieces_of(P, g, sh.bounds.data(), nb, all);  // [0] is the triangle, queued above
			const int npc = nall - 1;
			for (int q = 0; q < npc; q++)
				for (int k = 0; k < 4; k++) pc[q][k] = all[q + 1][k];
			const char* serial = getenv("ASS_SHARD_SERIAL_GRAVITY");
			if (serial && *serial && *serial != '0') {
				for (int q = 0; q < npc; q++)
					if ((rc = ass_launch_gravity_sym_pair(s, pc[q][0], pc[q][1], pc[q][2], pc[q][3], s->acc, true, fpart, false))) return rc;
			} else {
				for (int q = 0; q < npc; q++) {
					if (pc[q][1] <= pc[q][0] || pc[q][3] <= pc[q][2]) continue;
					CU_TRY(cudaStreamWaitEvent(sh.grav_stream[q], sh.drifted, 0));
					CU_TRY(cudaStreamWaitEvent(sh.grav_stream[q], sh.pulled, 0));
					if ((rc = ass_launch_gravity_sym_main(s, pc[q][0], pc[q][1], pc[q][2], pc[q][3], sh.grav_stream[q]))) return rc;
					CU_TRY(cudaEventRecord(sh.grav_done[q], sh.grav_stream[q]));
				}
				ass_tic(s, ASS_K_GRAVITY);
				for (int q = 0; q < npc; q++) {
					if (pc[q][1] <= pc[q][0] || pc[q][3] <= pc[q][2]) continue;
					CU_TRY(cudaStreamWaitEvent(s->stream, sh.grav_done[q], 0));
					if ((rc = ass_launch_gravity_sym_finish(s, pc[q][0], pc[q][1], pc[q][2], pc[q][3], s->acc, true, fpart, false))) return rc;
				}
				ass_toc(s, ASS_K_GRAVITY);
			}
			// point masses: they pull on my body particles; whoever owns them sums everything onto them
			if ((rc = one_sided_add(s, nb, n, lo, hi))) return rc;
			if ((rc = one_sided_add(s, 0, n, std::max(sh.lo, nb), sh.hi))) return rc;
			// my partials are complete; the ones other ranks computed for my slab are read out of their windows and added in
			// a fixed order (d = 1..D, then the antipode)
			partial_sources(s, from);
			unsigned need = 0;
			if (hi > lo)
				for (auto& f : from) need |= 1u << f.first;
			{
				FlagPtrs fp;
				memset(&fp, 0, sizeof(fp));
				for (int q = 0; q < P; q++) fp.w[q] = flags_of(s, q == g ? sh.window : sh.peer_window[q]);
				k_flag_set_wait<<<1, 32, 0, s->stream>>>(fp, P, ASS_SHARD_PART, g, e, flags_of(s, sh.window), need, err_of(s, sh.window));
				LAUNCH_CHECK();
			}
			if (need) {
				PartialSources ps;
				memset(&ps, 0, sizeof(ps));
				for (size_t q = 0; q < from.size(); q++) {
					if (hi <= from[q].second) continue;
					ps.src[ps.n] = sh.peer_window[from[q].first] + win_part(s);
					ps.first[ps.n] = from[q].second;
					ps.n++;
				}
				if (ps.n) {
					ass_tic(s, ASS_K_XFER);
					k_pull_add_partials<<<(unsigned) ((hi - lo + 255) / 256), 256, 0, s->stream>>>(s->acc, ps, lo, hi);
					LAUNCH_CHECK();
					ass_toc(s, ASS_K_XFER);
				}
			}
		} else if (gravity_on(s)) {
			const int ns = n_sources(s);
			const int a = std::min(sh.lo, ns), c = std::min(sh.hi, ns);
			int rows = rows_local, used = 0;
			if ((rc = ass_launch_gravity_partial(s, s->posm, 0, a, sh.lo, sh.hi, rows, &used))) return rc;
			rows += used;
			if ((rc = ass_launch_gravity_partial(s, s->posm, c, ns, sh.lo, sh.hi, rows, &used))) return rc;
			rows += used;
			if ((rc = ass_launch_gravity_reduce(s, sh.lo, sh.hi, rows))) return rc;
		}
		// ---- springs, kick, drift(s) for the own slab: out of buffer b into buffer b ^ 1 ----
		if (have_springs) {
			if ((rc = ass_launch_spring_kick_drift_range(s, s->prm.additional_forces == 2, dt, pre, sh.lo, sh.hi))) return rc;
		} else {
			if (s->prm.additional_forces == 2 && (rc = ass_launch_zero_accel_range(s, sh.lo, sh.hi))) return rc;
			if (sh.hi > sh.lo) {
				ass_tic(s, ASS_K_LEAPFROG);
				k_part2_pingpong<<<(sh.hi - sh.lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, s->acc, s->posm_alt, s->vel_alt, sh.lo, sh.hi, dt, pre ? 1 : 0);
				LAUNCH_CHECK();
				ass_toc(s, ASS_K_LEAPFROG);
			}
		}
		point_at(s, b ^ 1);
		s->predrifted = pre;
		s->prm.t += dt / 2.;  // integrator_leapfrog.c:65-66
		s->prm.dt_last_done = dt;
	}
	return ASS_OK;
}

// Bring every slab home (positions, velocities) so that ass_download_particles returns the whole body; accelerations of
// foreign slabs are pulled by the caller rank-by-rank if it wants them (each rank only ever computes its own).
// Collective, queued like a step: GATHER[g] = 2c - 1 says "my slab is final", the pulls wait for that from every
// peer, GATHER[g] = 2c says "I have pulled everybody's"; the next ass_shard_step starts by waiting for 2c from every
// peer, because its first kernel moves the own slab in place.  A download (or ass_sync) waits for the pulls.
extern "C" int ass_shard_gather(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	ShardInfo& sh = s->shard;
	REQUIRE(sh.active, ASS_ESTATE, "not sharded");
	REQUIRE(!s->predrifted, ASS_ESTATE, "finish the steps without pre_drift_last before gathering");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int P = sh.nranks, g = sh.rank, b = sh.cur;
	const unsigned all_others = ((1u << P) - 1u) & ~(1u << g);
	for (int q = 0; q < P; q++) REQUIRE(q == g || sh.peer_window[q], ASS_ESTATE, "a peer window has not been imported");
	const unsigned long long c = ++sh.gather_epoch;
	if ((rc = launch_flag_set(s, s->stream, ASS_SHARD_GATHER, 2 * c - 1))) return rc;
	if ((rc = launch_flag_wait(s, s->stream, all_others, ASS_SHARD_GATHER, 2 * c - 1))) return rc;
	ass_tic(s, ASS_K_XFER);
	for (int r = 0; r < P; r++) {
		if (r == g) continue;
		const size_t lo = (size_t) sh.bounds[r], cnt = (size_t) (sh.bounds[r + 1] - sh.bounds[r]);
		if (!cnt) continue;
		if ((rc = pull(s, s->stream, r, sh.window + win_posm(s, b) + lo, sh.peer_window[r] + win_posm(s, b) + lo, cnt))) return rc;
		if ((rc = pull(s, s->stream, r, sh.window + win_vel(s, b) + lo, sh.peer_window[r] + win_vel(s, b) + lo, cnt))) return rc;
	}
	ass_toc(s, ASS_K_XFER);
	if ((rc = launch_flag_set(s, s->stream, ASS_SHARD_GATHER, 2 * c))) return rc;
	sh.gather_pending = 2 * c;
	return ASS_OK;
}

// ass_sync for a sharded simulation: both streams, then the device-side error flag of the waits
int ass_shard_sync(ass_sim* s) {
	ShardInfo& sh = s->shard;
	CU_TRY(cudaStreamSynchronize(s->stream));
	CU_TRY(cudaStreamSynchronize(sh.copy_stream));
	return check_device_error(s);
}

void ass_shard_release(ass_sim* s) {
	ShardInfo& sh = s->shard;
	if (!sh.active) return;
	for (int r = 0; r < (int) sh.peer_window.size(); r++)
		if (sh.peer_window[r] && sh.peer_is_ipc[r]) cudaIpcCloseMemHandle(sh.peer_window[r]);
	if (sh.copy_stream) { cudaStreamSynchronize(sh.copy_stream); cudaStreamDestroy(sh.copy_stream); }
	if (sh.pulled) cudaEventDestroy(sh.pulled);
	if (sh.drifted) cudaEventDestroy(sh.drifted);
	for (int q = 0; q < 3; q++) {
		if (sh.grav_stream[q]) { cudaStreamSynchronize(sh.grav_stream[q]); cudaStreamDestroy(sh.grav_stream[q]); }
		if (sh.grav_done[q]) cudaEventDestroy(sh.grav_done[q]);
		sh.grav_stream[q] = nullptr; sh.grav_done[q] = nullptr;
	}
	sh.copy_stream = nullptr; sh.pulled = nullptr; sh.drifted = nullptr;
	// posm / vel and their alt twins alias the window: free it once and clear the aliases
	cudaFree(sh.window);
	s->posm = s->vel = s->posm_alt = s->vel_alt = nullptr;
	sh.window = nullptr;
	sh.active = false;
}
