// ensemble.cu -- B independent small bodies (spin-rate / material sweeps, BASELINE config 5: 4096 x ~1k nodes).
//
// The members never interact, so there is nothing to exchange: one CTA owns one body for a whole batch of steps
// ("body in shared memory, springs streamed").
//   * node positions and velocities (56 B/node: three arrays of 16-byte pairs + the mass) live in shared memory for the
//     whole batch, in the member's spatial RANK order -- the same ranking, rows and lane-transposed half-edge records
//     (state.cu: ass_build_spring_layout) as the single-body path, so a member sums in the same order and gets the same
//     bits; accelerations go through global memory (written once, read once per step, L2-resident);
//   * the half-edge records (2 x 16 B per spring, ~470 KB per body with padding) are read every step straight from L2
//     by the warps that need them, 512 contiguous bytes per warp and step, one pair of steps ahead of the arithmetic
//     (spring_math.cuh: lane_spring_sum) -- nothing is staged and the spring phase has no CTA-wide barrier inside it;
//   * per step:  part1 | [gravity: all pairs out of shared memory] | springs | kick + drift (+ next part1) | [center_sim].
// The C-ABI keeps the caller's node order: upload and download go through the permutation `perm`.
// Across GPUs the members are dealt round-robin by the host launcher; no collective is involved.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "spring_math.cuh"

namespace {

#ifndef ASS_ENS_THREADS
#define ASS_ENS_THREADS 512
#endif
constexpr int ET = ASS_ENS_THREADS;  // threads per CTA: one body owns a whole SM (16 warps x 128 registers)
constexpr int GPW = 32 / SPRING_LANES;  // nodes per warp

struct BodyDesc {
	int node_off, n, grp_off, swarp_off;  // first node / lane group / warp-offset entry of the member in the concatenated arrays
	int n_swarps, n_perts, gravity, af;
	double dt, negG, soft2;
	int heartbeat, pal_idx;  // pal_idx: the palette entry all of the member's springs share, -1 if they differ
};

// A member's state and accelerations may have been written by another SM earlier in the same launch (the previous batch
// of its steps): read them past the (non-coherent) L1.
__device__ __forceinline__ vec4d ldcg4(const vec4d* p) {
	const double2 a = __ldcg((const double2*) p), b = __ldcg((const double2*) p + 1);
	vec4d r;
	r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
	return r;
}

// Work is handed out in ITEMS = (member b, batch j of `item_steps` consecutive steps), numbered j * n_bodies + b and
// pulled from one atomic counter by persistent CTAs (one per SM).  With one CTA per member for the whole run, 512
// members on 148 SMs take 4 rounds of which the last is half empty (86 % at BASELINE C5 on 8 GPUs); with items the
// machine stays full and a member costs one extra trip of its state through global memory per batch.
// done[b] = number of batches of member b that are complete: an item waits (briefly, and only near a batch boundary)
// until its predecessor has published the member's state.
// All per-node arrays (posm_g, vel_g, acc_g, ord) are in the member's rank order.
template <bool UNI, bool PAL1>
__global__ void __launch_bounds__(ET, 1) k_ensemble(const BodyDesc* __restrict__ desc, vec4d* __restrict__ posm_g, vec4d* __restrict__ vel_g,
		vec4d* __restrict__ acc_g, const int* __restrict__ ord, const int2* __restrict__ grp, const int* __restrict__ sord, const int* __restrict__ warp_off,
		const HalfEdge* __restrict__ heT, const double2* __restrict__ palf, int nsteps_total, int with_hb, int max_n, int n_bodies, int item_steps,
		int* __restrict__ queue, int* __restrict__ done) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ double red[ET / 32][4];
	__shared__ double com[4];
	__shared__ int s_item, s_next;  // s_next: the next slice of the spring phase (warps pull slices: no warp waits for a longer share)
	const int tid = threadIdx.x;
	const int n_batches = (nsteps_total + item_steps - 1) / item_steps;
	while (true) {
	if (tid == 0) s_item = atomicAdd(queue, 1);
	__syncthreads();
	const int item = s_item;
	__syncthreads();
	if (item >= n_bodies * n_batches) break;
	const int batch = item / n_bodies, body = item - batch * n_bodies;
	const int nsteps = min(item_steps, nsteps_total - batch * item_steps);
	if (tid == 0) {  // the member's previous batch must have published its state
		while (atomicAdd(done + body, 0) < batch) __nanosleep(200);
		__threadfence();
	}
	__syncthreads();
	const BodyDesc B = desc[body];
	const int n = B.n;
	// carve: xy | zv | vxy (16-byte entries) | m (8-byte), each max_n long
	double2* s_xy = (double2*) smem_raw;
	double2* s_zv = s_xy + max_n;
	double2* s_vxy = s_zv + max_n;
	double* s_m = (double*) (s_vxy + max_n);
	vec4d* acc = acc_g + B.node_off;
	auto ldp = [&](int i) { const double2 a = s_xy[i]; vec4d p; p.x = a.x; p.y = a.y; p.z = s_zv[i].x; p.w = s_m[i]; return p; };
	auto ldv = [&](int i) { const double2 a = s_vxy[i]; vec4d v; v.x = a.x; v.y = a.y; v.z = s_zv[i].y; v.w = 0.0; return v; };
	auto stp = [&](int i, const vec4d p) { s_xy[i] = make_double2(p.x, p.y); s_zv[i].x = p.z; s_m[i] = p.w; };
	auto stv = [&](int i, const vec4d v) { s_vxy[i] = make_double2(v.x, v.y); s_zv[i].y = v.z; };
	const SharedNodes nodes{ s_xy, s_zv, s_vxy, s_m };
	for (int i = tid; i < n; i += ET) {
		stp(i, ldcg4(posm_g + B.node_off + i));
		stv(i, ldcg4(vel_g + B.node_off + i));
	}
	if (tid == 0) s_next = 0;
	__syncthreads();
	const double dt = B.dt, h = __dmul_rn(0.5, dt);
	const bool hb = with_hb && B.heartbeat == 1;
	const bool have_springs = B.af != 0 && warp_off[B.swarp_off + B.n_swarps] > warp_off[B.swarp_off];
	const bool grav = B.gravity == 1 && B.af != 2;
	const int lane = tid & 31;
	double2 nkg1 = make_double2(0.0, 0.0);
	if (PAL1) nkg1 = __ldg(palf + B.pal_idx);  // every member's springs share one (k, gamma): the pair travels in registers
	bool predrifted = false;
	for (int step = 0; step < nsteps; step++) {
		const bool last = step == nsteps - 1;
		if (!predrifted) {  // reb_integrator_leapfrog_part1, integrator_leapfrog.c:44-48
			for (int i = tid; i < n; i += ET) {
				vec4d p = ldp(i);
				const vec4d v = ldv(i);
				p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
				p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
				p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
				stp(i, p);
			}
			__syncthreads();
		}
		if (grav) {  // REB_GRAVITY_BASIC, gravity.c:72-103, all pairs out of shared memory
			for (int i = tid; i < n; i += ET) {
				const vec4d pi = ldp(i);
				double ax = 0.0, ay = 0.0, az = 0.0;
				for (int j = 0; j < n; j++) {
					const vec4d pj = ldp(j);
					const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
					double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, B.soft2)));
					r2 = (j == i) ? 1.0 : r2;
					const double s = rsqrt_cubed(r2) * pj.w;
					ax = fma(s, dx, ax); ay = fma(s, dy, ay); az = fma(s, dz, az);
				}
				vec4d a;
				a.x = B.negG * ax; a.y = B.negG * ay; a.z = B.negG * az; a.w = 0.0;
				acc[i] = a;
			}
			__syncthreads();  // also orders the global acc writes before the spring phase's reads (same CTA)
		}
		if (have_springs) {  // spring_forces (+ zero_accel when af == 2): warps pull slices of 32 / SPRING_LANES nodes from a counter
			while (true) {
				int sw = 0;
				if (lane == 0) sw = atomicAdd(&s_next, 1);
				sw = __shfl_sync(0xffffffffu, sw, 0);
				if (sw >= B.n_swarps) break;
				sw = __ldg(sord + B.grp_off / GPW + sw);  // slices go out longest first: the tail the barrier waits for is the shortest ones
				const int2 gi = __ldg(grp + B.grp_off + sw * GPW + lane / SPRING_LANES);  // (rank, degree); (-1, 0): padding group
				const int off = __ldg(warp_off + B.swarp_off + sw);
				const int P = (__ldg(warp_off + B.swarp_off + sw + 1) - off) >> 6;
				const bool live = gi.x >= 0;
				const int k = live ? gi.x : 0;
				const int nt = (gi.y - (lane & (SPRING_LANES - 1)) + SPRING_LANES - 1) / SPRING_LANES;
				const NodeState ni = nodes.get<false>(k);
				const Force3 f = lane_spring_sum<UNI, PAL1, false>((const int4*) heT + off + lane, P, nt, ni, nodes, palf, nkg1);
				if (live && (lane & (SPRING_LANES - 1)) == 0) {
					const double inv_m = (gi.y > 0) ? fast_rcp(ni.m) : 0.0;
					vec4d a;
					if (B.af == 2) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
					else a = ldcg4(acc + k);
					a.x = fma(f.x, inv_m, a.x);
					a.y = fma(f.y, inv_m, a.y);
					a.z = fma(f.z, inv_m, a.z);
					acc[k] = a;
				}
			}
			__syncthreads();  // nobody reads positions any more: the kick may move them
			if (tid == 0) s_next = 0;  // for the next step's spring phase (two barriers away)
		} else {
			if (B.af == 2)
				for (int i = tid; i < n; i += ET) { vec4d z; z.x = 0; z.y = 0; z.z = 0; z.w = 0; acc[i] = z; }
			__syncthreads();
		}
		const bool pre = !last && !hb;
		for (int i = tid; i < n; i += ET) {  // reb_integrator_leapfrog_part2, integrator_leapfrog.c:57-63
			vec4d p = ldp(i), v = ldv(i);
			// written by this CTA earlier in this step unless no force term is active (then: by an earlier batch)
			const vec4d a = (have_springs || grav || B.af == 2) ? acc[i] : ldcg4(acc + i);
			v.x = __dadd_rn(v.x, __dmul_rn(dt, a.x));
			v.y = __dadd_rn(v.y, __dmul_rn(dt, a.y));
			v.z = __dadd_rn(v.z, __dmul_rn(dt, a.z));
			p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
			p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
			p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
			if (pre) {
				p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
				p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
				p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
			}
			stv(i, v);
			stp(i, p);
		}
		predrifted = pre;
		__syncthreads();
		if (hb) {  // center_sim(0, n - n_perts): CoM of the resolved body, shift everything (physics.cpp:98-109)
			const int hi = n - B.n_perts;  // in the caller's numbering: ord[] maps a rank back to it
			double v[4] = { 0.0, 0.0, 0.0, 0.0 };
			for (int i = tid; i < n; i += ET) {
				if (B.n_perts > 0 && __ldg(ord + B.node_off + i) >= hi) continue;
				const vec4d p = ldp(i);
				v[0] += p.w * p.x; v[1] += p.w * p.y; v[2] += p.w * p.z; v[3] += p.w;
			}
#pragma unroll
			for (int q = 0; q < 4; q++)
				for (int m = 16; m > 0; m >>= 1) v[q] += shfl_down_d(v[q], m);
			if ((tid & 31) == 0)
				for (int q = 0; q < 4; q++) red[tid >> 5][q] = v[q];
			__syncthreads();
			if (tid < 4) {
				double a = 0.0;
				for (int w = 0; w < ET / 32; w++) a += red[w][tid];
				com[tid] = a;
			}
			__syncthreads();
			const double cx = com[0] / com[3], cy = com[1] / com[3], cz = com[2] / com[3];
			for (int i = tid; i < n; i += ET) {
				vec4d p = ldp(i);
				p.x = __dadd_rn(p.x, -cx); p.y = __dadd_rn(p.y, -cy); p.z = __dadd_rn(p.z, -cz);
				if (!last) {
					const vec4d w = ldv(i);
					p.x = __dadd_rn(p.x, __dmul_rn(h, w.x));
					p.y = __dadd_rn(p.y, __dmul_rn(h, w.y));
					p.z = __dadd_rn(p.z, __dmul_rn(h, w.z));
				}
				stp(i, p);
			}
			predrifted = !last;
			__syncthreads();
		}
	}
	for (int i = tid; i < n; i += ET) {
		posm_g[B.node_off + i] = ldp(i);
		vel_g[B.node_off + i] = ldv(i);
	}
	__threadfence();  // state and accelerations of this batch are visible before the member is released
	__syncthreads();
	if (tid == 0) atomicExch(done + body, batch + 1);
	}  // next item
}

// rows in the caller's node order <-> device arrays in rank order: perm[i] = device index of the caller's node i
__global__ void k_unpack(const unsigned char* __restrict__ rows, size_t stride, int n, const int* __restrict__ perm, vec4d* __restrict__ posm,
		vec4d* __restrict__ vel, vec4d* __restrict__ acc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double* r = (const double*) (rows + (size_t) i * stride);
	vec4d p, v, a;
	p.x = r[0]; p.y = r[1]; p.z = r[2]; p.w = r[9];
	v.x = r[3]; v.y = r[4]; v.z = r[5]; v.w = 0.0;
	a.x = r[6]; a.y = r[7]; a.z = r[8]; a.w = 0.0;
	const int k = perm[i];
	posm[k] = p; vel[k] = v; acc[k] = a;
}
__global__ void k_pack(double* __restrict__ dense, int n, const int* __restrict__ perm, const vec4d* __restrict__ posm, const vec4d* __restrict__ vel,
		const vec4d* __restrict__ acc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double* r = dense + (size_t) i * 10;
	const int k = perm[i];
	const vec4d p = posm[k], v = vel[k], a = acc[k];
	r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = v.x; r[4] = v.y; r[5] = v.z; r[6] = a.x; r[7] = a.y; r[8] = a.z; r[9] = p.w;
}

}  // namespace

struct ass_ensemble {
	int n_bodies = 0, total_nodes = 0, max_n = 0;
	long long total_he = 0;
	std::vector<ass_params> prm;
	std::vector<BodyDesc> desc;
	BodyDesc* d_desc = nullptr;
	vec4d *posm = nullptr, *vel = nullptr, *acc = nullptr;  // rank order within each member
	int* perm = nullptr;      // caller's global node index -> device index (node_off + rank)
	int* ord = nullptr;       // device index -> caller's node index inside its member
	int2* grp = nullptr;      // lane groups of all members (common.cuh: ass_sim::grp), ranks local to the member
	int* warp_off = nullptr;  // per member n_swarps + 1 offsets into heT
	int* sord = nullptr;      // per member n_swarps slice numbers in hand-out order (by decreasing length)
	HalfEdge* heT = nullptr;  // lane-transposed records of all members; nbr = rank inside the member
	double2* palf = nullptr;  // distinct (-k, -max(gamma, 0)) pairs over all members
	double* d_dense = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	size_t smem = 0;
	bool pal1 = false;     // inside every member all springs have the same (k, gamma): the pair travels in registers
	bool uni_mass = true;  // within every member, all nodes carry the same mass (m_red = m / 2, spring_math.cuh)
	int* d_queue = nullptr;  // [0] the item counter, [1 .. n_bodies] batches completed per member (k_ensemble)
};

static void ens_free(ass_ensemble* e) {
	if (!e) return;
	cudaFree(e->d_queue); cudaFree(e->d_desc); cudaFree(e->posm); cudaFree(e->vel); cudaFree(e->acc); cudaFree(e->perm); cudaFree(e->ord);
	cudaFree(e->grp); cudaFree(e->sord); cudaFree(e->warp_off); cudaFree(e->heT); cudaFree(e->palf); cudaFree(e->d_dense);
	if (e->e0) cudaEventDestroy(e->e0);
	if (e->e1) cudaEventDestroy(e->e1);
	if (e->stream) cudaStreamDestroy(e->stream);
	cudaGetLastError();
	delete e;
}

static void fill_desc(ass_ensemble* e) {
	for (int b = 0; b < e->n_bodies; b++) {
		BodyDesc& d = e->desc[b];
		const ass_params& p = e->prm[b];
		d.n_perts = p.n_perts; d.dt = p.dt; d.negG = -p.G; d.soft2 = p.softening * p.softening;
		d.gravity = p.gravity; d.af = p.additional_forces; d.heartbeat = p.heartbeat;
	}
}

extern "C" int ass_ensemble_create(ass_ensemble** out, int n_bodies, const int* node_off, const int* spring_off, const void* base,
		size_t stride, const ass_spring* springs, const ass_params* params) {
	REQUIRE(out && node_off && spring_off && base && params, ASS_EINVAL, "null argument");
	REQUIRE(n_bodies > 0, ASS_EINVAL, "need at least one body");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "bad row stride");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int total_nodes = node_off[n_bodies];
	const int total_springs = spring_off[n_bodies];
	REQUIRE(node_off[0] == 0 && spring_off[0] == 0, ASS_EINVAL, "offsets must start at 0");
	REQUIRE(total_springs == 0 || springs, ASS_EINVAL, "null spring array");
	ass_ensemble* e = new ass_ensemble();
	e->n_bodies = n_bodies;
	e->total_nodes = total_nodes;
	e->prm.assign(params, params + n_bodies);
	e->desc.resize(n_bodies);
	// per member: the single-body path's spring layout (state.cu), concatenated; ranks stay local to the member
	std::vector<int> perm((size_t) std::max(total_nodes, 0)), ord((size_t) std::max(total_nodes, 0));
	std::vector<int2> grp;
	std::vector<int> warp_off;
	std::vector<int> sord;
	std::vector<HalfEdge> heT;
	// palette of distinct (k, gamma) bit patterns over the whole ensemble (a sweep has a handful), stored as the force
	// kernels want it (spring_math.cuh)
	std::vector<double2> palf;
	bool all_pal1 = true;  // inside every member, all springs share one palette entry (a material sweep: k differs between members)
	std::unordered_map<std::string, int> seen;
	auto pal_index = [&](const ass_spring& v) {
		std::string key((const char*) &v.k, 8);
		key.append((const char*) &v.gamma, 8);
		auto it = seen.find(key);
		if (it == seen.end()) {
			it = seen.emplace(key, (int) palf.size()).first;
			palf.push_back(make_double2(-v.k, v.gamma > 0.0 ? -v.gamma : -0.0));
		}
		return it->second;
	};
	for (int b = 0; b < n_bodies; b++) {
		const int n = node_off[b + 1] - node_off[b], ns = spring_off[b + 1] - spring_off[b];
		if (n <= 0 || ns < 0) { ens_free(e); ass_set_error("ass_ensemble_create: body %d is empty or offsets decrease", b); return ASS_EINVAL; }
		const ass_params& p = params[b];
		if ((p.gravity != 0 && p.gravity != 1) || p.additional_forces < 0 || p.additional_forces > 2 || (p.heartbeat != 0 && p.heartbeat != 1)) {
			ens_free(e); ass_set_error("ass_ensemble_create: bad params for body %d", b); return ASS_EINVAL;
		}
		e->max_n = std::max(e->max_n, n);
	}
	// The members' layouts are independent: built a chunk at a time by a few host threads (a layout is ~10 ms of host work
	// per 1k-node member), then appended in member order.
	constexpr int CHUNK = 64;
	const int n_threads = (int) std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	std::vector<SpringLayout> lays(CHUNK);
	std::vector<std::vector<int>> pidxs(CHUNK);
	std::vector<int> rcs(CHUNK), bads(CHUNK);
	std::vector<char> unis(CHUNK);
	for (int b0 = 0; b0 < n_bodies; b0 += CHUNK) {
		const int nb_chunk = std::min(CHUNK, n_bodies - b0);
		for (int j = 0; j < nb_chunk; j++) {  // palette indices: one palette for the whole ensemble, filled in member order
			const int b = b0 + j, ns = spring_off[b + 1] - spring_off[b];
			const ass_spring* sp = springs + spring_off[b];
			pidxs[(size_t) j].resize((size_t) ns);
			for (int q = 0; q < ns; q++) pidxs[(size_t) j][(size_t) q] = pal_index(sp[q]);
		}
		std::atomic<int> next(0);
		auto work = [&]() {
			for (int j = next.fetch_add(1); j < nb_chunk; j = next.fetch_add(1)) {
				const int b = b0 + j, n = node_off[b + 1] - node_off[b], ns = spring_off[b + 1] - spring_off[b];
				std::vector<vec4d> pos((size_t) n);
				bool uni = true;
				for (int i = 0; i < n; i++) {
					const double* rr = (const double*) ((const unsigned char*) base + (size_t) (node_off[b] + i) * stride);
					pos[(size_t) i].x = rr[0]; pos[(size_t) i].y = rr[1]; pos[(size_t) i].z = rr[2]; pos[(size_t) i].w = rr[9];
					if (memcmp(&pos[(size_t) i].w, &pos[0].w, sizeof(double)) != 0) uni = false;
				}
				unis[(size_t) j] = uni ? 1 : 0;
				bads[(size_t) j] = -1;
				rcs[(size_t) j] = ass_build_spring_layout(pos.data(), n, springs + spring_off[b], ns, pidxs[(size_t) j].data(), lays[(size_t) j], &bads[(size_t) j]);
			}
		};
		{
			std::vector<std::thread> pool;
			for (int t = 1; t < std::min(n_threads, nb_chunk); t++) pool.emplace_back(work);
			work();
			for (auto& th : pool) th.join();
		}
		for (int j = 0; j < nb_chunk; j++) {
			const int b = b0 + j, n = node_off[b + 1] - node_off[b], ns = spring_off[b + 1] - spring_off[b];
			const ass_spring* sp = springs + spring_off[b];
			const SpringLayout& lay = lays[(size_t) j];
			const std::vector<int>& pidx = pidxs[(size_t) j];
			if (!unis[(size_t) j]) e->uni_mass = false;
			if (rcs[(size_t) j] == ASS_EINVAL) {
				const int bad = bads[(size_t) j];
				ass_set_error("ass_ensemble_create: body %d spring %d joins (%d,%d): indices are local to the body and must differ", b, bad,
				              sp[bad].particle_1, sp[bad].particle_2);
				ens_free(e);
				return ASS_EINVAL;
			}
			const long long rec_base = (long long) heT.size();
			if (rcs[(size_t) j] || rec_base + (long long) lay.heT.size() > 0x7fffffffLL) {
				ass_set_error("ass_ensemble_create: the members' half-edge records overflow a 32-bit index at body %d", b);
				ens_free(e);
				return ASS_EINVAL;
			}
			BodyDesc& d = e->desc[b];
			d.node_off = node_off[b]; d.n = n;
			d.grp_off = (int) grp.size(); d.swarp_off = (int) warp_off.size(); d.n_swarps = (int) lay.warp_off.size() - 1;
			d.pal_idx = ns > 0 ? pidx[0] : 0;
			for (int q = 1; q < ns; q++)
				if (pidx[(size_t) q] != pidx[0]) { d.pal_idx = -1; break; }
			if (d.pal_idx < 0) all_pal1 = false;
			for (int i = 0; i < n; i++) {
				perm[(size_t) node_off[b] + i] = node_off[b] + lay.rank[(size_t) i];
				ord[(size_t) node_off[b] + lay.rank[(size_t) i]] = i;
			}
			grp.insert(grp.end(), lay.grp.begin(), lay.grp.end());
			{
				const int nsw = (int) lay.warp_off.size() - 1;
				std::vector<int> o((size_t) nsw);
				for (int w = 0; w < nsw; w++) o[(size_t) w] = w;
				std::stable_sort(o.begin(), o.end(), [&](int a, int b) { return lay.warp_off[(size_t) a + 1] - lay.warp_off[a] > lay.warp_off[(size_t) b + 1] - lay.warp_off[b]; });
				sord.insert(sord.end(), o.begin(), o.end());
			}
			for (int w : lay.warp_off) warp_off.push_back((int) (rec_base + w));
			heT.insert(heT.end(), lay.heT.begin(), lay.heT.end());
			e->total_he += 2LL * ns;
		}
	}
	fill_desc(e);
	e->pal1 = all_pal1;
	e->smem = (3 * sizeof(double2) + sizeof(double)) * ((size_t) e->max_n + 1) + 64;
	if (e->smem > 227 * 1024 - 2048) {
		ens_free(e);
		ass_set_error("ass_ensemble_create: a body of %d nodes does not fit in shared memory (limit ~4000 nodes); use ass_sim for large bodies", e->max_n);
		return ASS_EINVAL;
	}
	const size_t nb = (size_t) total_nodes * stride;
	void* d_rows = nullptr;
	bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e->e0) == cudaSuccess &&
	          cudaEventCreate(&e->e1) == cudaSuccess && cudaMalloc(&e->d_desc, sizeof(BodyDesc) * n_bodies) == cudaSuccess &&
	          cudaMalloc(&e->posm, sizeof(vec4d) * total_nodes) == cudaSuccess && cudaMalloc(&e->vel, sizeof(vec4d) * total_nodes) == cudaSuccess &&
	          cudaMalloc(&e->acc, sizeof(vec4d) * total_nodes) == cudaSuccess && cudaMalloc(&e->perm, sizeof(int) * total_nodes) == cudaSuccess &&
	          cudaMalloc(&e->ord, sizeof(int) * total_nodes) == cudaSuccess && cudaMalloc(&e->grp, sizeof(int2) * std::max<size_t>(grp.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->warp_off, sizeof(int) * std::max<size_t>(warp_off.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->sord, sizeof(int) * std::max<size_t>(sord.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->heT, sizeof(HalfEdge) * std::max<size_t>(heT.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->palf, sizeof(double2) * std::max<size_t>(palf.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->d_dense, sizeof(double) * 10 * (size_t) total_nodes) == cudaSuccess && cudaMalloc(&d_rows, nb) == cudaSuccess;
	if (!ok) {
		cudaFree(d_rows);
		ass_set_error("ass_ensemble_create: %s", cudaGetErrorString(cudaGetLastError()));
		ens_free(e);
		return ASS_ENOMEM;
	}
	cudaMemcpyAsync(d_rows, base, nb, cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->ord, ord.data(), sizeof(int) * ord.size(), cudaMemcpyHostToDevice, e->stream);
	if (!grp.empty()) cudaMemcpyAsync(e->grp, grp.data(), sizeof(int2) * grp.size(), cudaMemcpyHostToDevice, e->stream);
	if (!warp_off.empty()) cudaMemcpyAsync(e->warp_off, warp_off.data(), sizeof(int) * warp_off.size(), cudaMemcpyHostToDevice, e->stream);
	if (!sord.empty()) cudaMemcpyAsync(e->sord, sord.data(), sizeof(int) * sord.size(), cudaMemcpyHostToDevice, e->stream);
	if (!heT.empty()) cudaMemcpyAsync(e->heT, heT.data(), sizeof(HalfEdge) * heT.size(), cudaMemcpyHostToDevice, e->stream);
	if (!palf.empty()) cudaMemcpyAsync(e->palf, palf.data(), sizeof(double2) * palf.size(), cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->d_desc, e->desc.data(), sizeof(BodyDesc) * n_bodies, cudaMemcpyHostToDevice, e->stream);
	k_unpack<<<(total_nodes + 255) / 256, 256, 0, e->stream>>>((const unsigned char*) d_rows, stride, total_nodes, e->perm, e->posm, e->vel, e->acc);
	g_launch_count++;
	cudaError_t err = cudaStreamSynchronize(e->stream);
	cudaFree(d_rows);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
	if (err != cudaSuccess) {
		ass_set_error("ass_ensemble_create: %s", cudaGetErrorString(err));
		ens_free(e);
		return ASS_ECUDA;
	}
	*out = e;
	return ASS_OK;
}

extern "C" int ass_ensemble_set_params(ass_ensemble* e, const ass_params* params) {
	REQUIRE(e && params, ASS_EINVAL, "null argument");
	int rc = ass_ensure_device();
	if (rc) return rc;
	e->prm.assign(params, params + e->n_bodies);
	fill_desc(e);
	CU_TRY(cudaMemcpyAsync(e->d_desc, e->desc.data(), sizeof(BodyDesc) * e->n_bodies, cudaMemcpyHostToDevice, e->stream));
	CU_TRY(cudaStreamSynchronize(e->stream));
	return ASS_OK;
}

extern "C" int ass_ensemble_step(ass_ensemble* e, int nsteps, int with_heartbeat, double* device_ms) {
	REQUIRE(e, ASS_EINVAL, "null ensemble");
	REQUIRE(nsteps >= 0, ASS_EINVAL, "negative step count");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (device_ms) *device_ms = 0.0;
	if (nsteps == 0) return ASS_OK;
	if (device_ms) CU_TRY(cudaEventRecord(e->e0, e->stream));
	// persistent CTAs pulling (member, batch of steps) items; one batch when the members fit the machine in one round
	const int grid = std::min(e->n_bodies, ass_sm_count());
	// Items cost ~0.6 % per batch boundary per 10 steps (state through global memory, ring drain): cut the run into
	// batches only when whole-run items would leave the last round of members badly filled.
	int item_steps = nsteps;
	const int rounds = (e->n_bodies + grid - 1) / grid;
	if ((double) e->n_bodies / ((double) rounds * grid) < 0.95) item_steps = std::max(1, std::min(nsteps, std::max(5, std::min(25, (nsteps + 7) / 8))));
	if (const char* env = getenv("ASS_ENS_ITEM_STEPS")) item_steps = std::max(1, std::min(nsteps, atoi(env)));
	if (!e->d_queue) CU_TRY(cudaMalloc(&e->d_queue, sizeof(int) * ((size_t) e->n_bodies + 1)));
	CU_TRY(cudaMemsetAsync(e->d_queue, 0, sizeof(int) * ((size_t) e->n_bodies + 1), e->stream));
#define ASS_ENS(U, P1)                                                                                                                       \
	k_ensemble<U, P1><<<grid, ET, e->smem, e->stream>>>(e->d_desc, e->posm, e->vel, e->acc, e->ord, e->grp, e->sord, e->warp_off, e->heT, e->palf, nsteps, \
			with_heartbeat, e->max_n, e->n_bodies, item_steps, e->d_queue, e->d_queue + 1)
	if (e->uni_mass) { if (e->pal1) ASS_ENS(true, true); else ASS_ENS(true, false); }
	else { if (e->pal1) ASS_ENS(false, true); else ASS_ENS(false, false); }
#undef ASS_ENS
	LAUNCH_CHECK();
	for (auto& p : e->prm)
		for (int s = 0; s < nsteps; s++) {  // the reference's own rounding of t (integrator_leapfrog.c:50,65)
			p.t += p.dt / 2.;
			p.t += p.dt / 2.;
			p.dt_last_done = p.dt;
		}
	if (device_ms) {
		CU_TRY(cudaEventRecord(e->e1, e->stream));
		CU_TRY(cudaEventSynchronize(e->e1));
		float ms = 0.f;
		CU_TRY(cudaEventElapsedTime(&ms, e->e0, e->e1));
		*device_ms = ms;
	}
	return ASS_OK;
}

extern "C" int ass_ensemble_download(ass_ensemble* e, void* base, size_t stride, unsigned fields) {
	REQUIRE(e && base, ASS_EINVAL, "null argument");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "bad row stride");
	REQUIRE((fields & ~ASS_F_ALL) == 0 && fields != 0, ASS_EINVAL, "bad field mask");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int n = e->total_nodes;
	k_pack<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_dense, n, e->perm, e->posm, e->vel, e->acc);
	LAUNCH_CHECK();
	std::vector<double> h((size_t) n * 10);
	CU_TRY(cudaMemcpyAsync(h.data(), e->d_dense, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, e->stream));
	CU_TRY(cudaStreamSynchronize(e->stream));
	for (int i = 0; i < n; i++) {
		double* r = (double*) ((unsigned char*) base + (size_t) i * stride);
		const double* q = h.data() + (size_t) i * 10;
		if (fields & ASS_F_POS) { r[0] = q[0]; r[1] = q[1]; r[2] = q[2]; }
		if (fields & ASS_F_VEL) { r[3] = q[3]; r[4] = q[4]; r[5] = q[5]; }
		if (fields & ASS_F_ACC) { r[6] = q[6]; r[7] = q[7]; r[8] = q[8]; }
		if (fields & ASS_F_MASS) r[9] = q[9];
	}
	return ASS_OK;
}

extern "C" int ass_ensemble_get_params(ass_ensemble* e, ass_params* params) {
	REQUIRE(e && params, ASS_EINVAL, "null argument");
	for (int b = 0; b < e->n_bodies; b++) params[b] = e->prm[b];
	return ASS_OK;
}
extern "C" int ass_ensemble_sync(ass_ensemble* e) {
	REQUIRE(e, ASS_EINVAL, "null ensemble");
	CU_TRY(cudaStreamSynchronize(e->stream));
	return ASS_OK;
}
extern "C" int ass_ensemble_destroy(ass_ensemble* e) {
	if (!e) return ASS_OK;
	ass_ensure_device();
	if (e->stream) cudaStreamSynchronize(e->stream);
	ens_free(e);
	return ASS_OK;
}
