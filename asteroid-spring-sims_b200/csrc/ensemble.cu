// ensemble.cu -- B independent small bodies (spin-rate / material sweeps, BASELINE config 5: 4096 x ~1k nodes).
//
// The members never interact, so there is nothing to exchange: one CTA owns one body for a whole batch of steps
// ("body in shared memory, springs streamed").
//   * node positions and velocities (64 B/node, 68 KB for 1065 nodes) and the CSR row offsets live in shared memory for
//     the whole batch; accelerations go through global memory (written once, read once per step, L2-resident);
//   * the half-edge records (2 x 16 B per spring, ~445 KB per body) are streamed every step by the bulk-copy engine
//     (cp.async.bulk -> mbarrier, SASS UBLKCP): nodes are processed in waves of 64 (one node per 8-lane group), the
//     records of a wave are one contiguous run of the node-sorted array, and the run of wave w+2 is in flight while wave
//     w is being evaluated (two-deep ring).  So inside the spring phase every operand comes from shared memory and the
//     dependent chain "row offset -> record -> endpoint gather" costs shared-memory latencies, not HBM ones;
//   * per step:  part1 | [gravity: all pairs out of shared memory] | springs (same arithmetic and lane order as springs.cu,
//     so a member gets the bits the single-body path would give it) | kick + drift (+ next part1) | [center_sim].
// Across GPUs the members are dealt round-robin by the host launcher; no collective is involved.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "spring_math.cuh"

namespace {

#ifndef ASS_ENS_THREADS
#define ASS_ENS_THREADS 1024
#endif
constexpr int ET = ASS_ENS_THREADS;            // threads per CTA: one body owns a whole SM
constexpr int EW = ET / SPRING_LANES;          // nodes per wave (128)

struct BodyDesc {
	int node_off, n, row_off, n_perts;
	double dt, negG, soft2;
	int gravity, af, heartbeat, pad;
};

// A member's state and accelerations may have been written by another SM earlier in the same launch (the previous batch
// of its steps): read them past the (non-coherent) L1.
__device__ __forceinline__ vec4d ldcg4(const vec4d* p) {
	const double2 a = __ldcg((const double2*) p), b = __ldcg((const double2*) p + 1);
	vec4d r;
	r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
	return r;
}

// endpoint state out of shared memory (see the layout note in k_ensemble): three 16-byte reads per endpoint, plus the
// mass only when the member's masses differ (with equal masses the arithmetic never looks at the neighbour's)
template <bool UNI>
struct SharedNodes {
	const double2 *xy, *zv, *vxy;
	const double* m;
	__device__ __forceinline__ void get(int i, vec4d& p, vec4d& v) const {
		const double2 a = xy[i], b = zv[i], c = vxy[i];
		p.x = a.x; p.y = a.y; p.z = b.x; p.w = UNI ? 0.0 : m[i];
		v.x = c.x; v.y = c.y; v.z = b.y; v.w = 0.0;
	}
};

// Work is handed out in ITEMS = (member b, batch j of `item_steps` consecutive steps), numbered j * n_bodies + b and
// pulled from one atomic counter by persistent CTAs (one per SM).  With one CTA per member for the whole run, 512
// members on 148 SMs take 4 rounds of which the last is half empty (86 % at BASELINE C5 on 8 GPUs); with items the
// machine stays full and a member costs one extra 2 x 68 KB trip of its state through global memory per batch.
// done[b] = number of batches of member b that are complete: an item waits (briefly, and only near a batch boundary)
// until its predecessor has published the member's state.
template <bool UNI>
__global__ void __launch_bounds__(ET, 1) k_ensemble(const BodyDesc* __restrict__ desc, vec4d* __restrict__ posm_g, vec4d* __restrict__ vel_g,
		vec4d* __restrict__ acc_g, const int* __restrict__ row_ptr, const HalfEdge* __restrict__ he, const double2* __restrict__ pal, int nsteps_total,
		int with_hb, int max_n, int ring_records, int n_bodies, int item_steps, int* __restrict__ queue, int* __restrict__ done) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ double red[ET / 32][4];
	__shared__ double com[4];
	__shared__ __align__(8) uint64_t full[2];
	__shared__ int s_item;
	const int tid = threadIdx.x;
	if (tid == 0) {
		mbar_init(&full[0], 1);
		mbar_init(&full[1], 1);
		mbar_fence_init();
	}
	__syncthreads();
	const int n_batches = (nsteps_total + item_steps - 1) / item_steps;
	int c_base = 0;  // waves this CTA has streamed so far: ring slot and mbarrier parity carry across items
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
	// carve: ring[2][ring_records] | xy | zv | vxy (16-byte entries) | m (8-byte) | srow[max_n + 1]
	// Node state is kept as arrays of 16-byte pairs -- (x, y), (z, vz), (vx, vy) -- rather than 32-byte vectors: a random
	// gather of 32-byte entries starts on only four distinct bank groups (8-way conflicts); 16-byte entries start on
	// eight and tile all 32 banks.  Three reads per endpoint instead of four: the shared-memory pipe is what bounds this
	// kernel.  The mass has its own array.
	HalfEdge* ring = (HalfEdge*) smem_raw;
	double2* s_xy = (double2*) (ring + 2 * (size_t) ring_records);
	double2* s_zv = s_xy + max_n;
	double2* s_vxy = s_zv + max_n;
	double* s_m = (double*) (s_vxy + max_n);
	int* srow = (int*) (s_m + max_n + (max_n & 1));
	vec4d* acc = acc_g + B.node_off;
	auto ldp = [&](int i) { const double2 a = s_xy[i]; vec4d p; p.x = a.x; p.y = a.y; p.z = s_zv[i].x; p.w = s_m[i]; return p; };
	auto ldv = [&](int i) { const double2 a = s_vxy[i]; vec4d v; v.x = a.x; v.y = a.y; v.z = s_zv[i].y; v.w = 0.0; return v; };
	auto stp = [&](int i, const vec4d p) { s_xy[i] = make_double2(p.x, p.y); s_zv[i].x = p.z; s_m[i] = p.w; };
	auto stv = [&](int i, const vec4d v) { s_vxy[i] = make_double2(v.x, v.y); s_zv[i].y = v.z; };
	const SharedNodes<UNI> nodes{ s_xy, s_zv, s_vxy, s_m };
	for (int i = tid; i < n; i += ET) {
		stp(i, ldcg4(posm_g + B.node_off + i));
		stv(i, ldcg4(vel_g + B.node_off + i));
	}
	for (int i = tid; i <= n; i += ET) srow[i] = row_ptr[B.row_off + i];
	__syncthreads();
	const double dt = B.dt, h = __dmul_rn(0.5, dt);
	const bool hb = with_hb && B.heartbeat == 1;
	const bool have_springs = B.af != 0 && srow[n] > srow[0];
	const bool grav = B.gravity == 1 && B.af != 2;
	const int lane = tid & (SPRING_LANES - 1), group = tid / SPRING_LANES;
	const int nw = (n + EW - 1) / EW;
	const int total_waves = have_springs ? nsteps * nw : 0;
	// records of wave c (counted over all steps of the item) -> ring buffer (c_base + c) & 1
	auto issue = [&](int c) {  // thread 0 only
		const int w = c % nw, slot = (c_base + c) & 1;
		const int e0 = srow[w * EW], e1 = srow[min(n, (w + 1) * EW)];
		const uint32_t bytes = (uint32_t) (e1 - e0) * (uint32_t) sizeof(HalfEdge);
		mbar_expect_tx(&full[slot], bytes);
		if (bytes) bulk_g2s(ring + (size_t) slot * ring_records, he + e0, bytes, &full[slot]);
	};
	if (tid == 0) {
		if (total_waves > 0) issue(0);
		if (total_waves > 1) issue(1);
	}
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
		if (have_springs) {  // spring_forces (+ zero_accel when af == 2), one wave of 64 nodes at a time
			for (int w = 0; w < nw; w++) {
				const int c = step * nw + w, cg = c_base + c;
				mbar_wait(&full[cg & 1], (uint32_t) ((cg >> 1) & 1));
				const int gid = w * EW + group;
				const bool live = gid < n;
				const int node = live ? gid : n - 1;
				const vec4d pi = ldp(node), vi = ldv(node);
				const int beg = srow[node], end = live ? srow[node + 1] : beg;
				// the ring buffer holds records [srow[w*EW], ...): index it with global record numbers
				const HalfEdge* wave_he = ring + (size_t) (cg & 1) * ring_records - srow[w * EW];
				const Force3 f = node_spring_sum<false, UNI>(wave_he, pal, beg, end, lane, pi, vi, nodes);
				if (live && lane == 0) {
					const double inv_m = (end > beg) ? fast_rcp(pi.w) : 0.0;
					vec4d a;
					if (B.af == 2) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
					else a = ldcg4(acc + node);
					a.x = fma(f.x, inv_m, a.x);
					a.y = fma(f.y, inv_m, a.y);
					a.z = fma(f.z, inv_m, a.z);
					acc[node] = a;
				}
				__syncthreads();  // everyone is done with ring buffer c & 1 (and, after the last wave, with sp / sv)
				if (tid == 0 && c + 2 < total_waves) issue(c + 2);
			}
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
			const int hi = n - B.n_perts;
			double v[4] = { 0.0, 0.0, 0.0, 0.0 };
			for (int i = tid; i < hi; i += ET) {
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
	c_base += total_waves;
	__threadfence();  // state and accelerations of this batch are visible before the member is released
	__syncthreads();
	if (tid == 0) atomicExch(done + body, batch + 1);
	}  // next item
}

__global__ void k_unpack(const unsigned char* __restrict__ rows, size_t stride, int n, vec4d* __restrict__ posm, vec4d* __restrict__ vel,
		vec4d* __restrict__ acc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double* r = (const double*) (rows + (size_t) i * stride);
	vec4d p, v, a;
	p.x = r[0]; p.y = r[1]; p.z = r[2]; p.w = r[9];
	v.x = r[3]; v.y = r[4]; v.z = r[5]; v.w = 0.0;
	a.x = r[6]; a.y = r[7]; a.z = r[8]; a.w = 0.0;
	posm[i] = p; vel[i] = v; acc[i] = a;
}
__global__ void k_pack(double* __restrict__ dense, int n, const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const vec4d* __restrict__ acc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double* r = dense + (size_t) i * 10;
	const vec4d p = posm[i], v = vel[i], a = acc[i];
	r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = v.x; r[4] = v.y; r[5] = v.z; r[6] = a.x; r[7] = a.y; r[8] = a.z; r[9] = p.w;
}

}  // namespace

struct ass_ensemble {
	int n_bodies = 0, total_nodes = 0, max_n = 0;
	long long total_he = 0;
	std::vector<ass_params> prm;
	std::vector<BodyDesc> desc;
	BodyDesc* d_desc = nullptr;
	vec4d *posm = nullptr, *vel = nullptr, *acc = nullptr;
	int* row_ptr = nullptr;
	HalfEdge* he = nullptr;
	double2* pal = nullptr;  // distinct (k, gamma) pairs over all members
	double* d_dense = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	size_t smem = 0;
	int ring_records = 0;  // capacity of one ring buffer: the longest run of half-edges of one wave of nodes
	bool uni_mass = true;  // within every member, all nodes carry the same mass (m_red = m / 2, spring_math.cuh)
	int* d_queue = nullptr;  // [0] the item counter, [1 .. n_bodies] batches completed per member (k_ensemble)
};

static void ens_free(ass_ensemble* e) {
	if (!e) return;
	cudaFree(e->d_queue); cudaFree(e->d_desc); cudaFree(e->posm); cudaFree(e->vel); cudaFree(e->acc); cudaFree(e->row_ptr); cudaFree(e->he); cudaFree(e->pal); cudaFree(e->d_dense);
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
		d.gravity = p.gravity; d.af = p.additional_forces; d.heartbeat = p.heartbeat; d.pad = 0;
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
	// CSR per body: row entries are offsets into the concatenated half-edge array; indices stay local to the body
	std::vector<int> row((size_t) total_nodes + n_bodies, 0);
	std::vector<HalfEdge> he((size_t) 2 * total_springs);
	// palette of distinct (k, gamma) bit patterns over the whole ensemble (a sweep has a handful)
	std::vector<double2> pal;
	std::unordered_map<std::string, int> seen;
	auto pal_index = [&](const ass_spring& v) {
		std::string key((const char*) &v.k, 8);
		key.append((const char*) &v.gamma, 8);
		auto it = seen.find(key);
		if (it == seen.end()) {
			it = seen.emplace(key, (int) pal.size()).first;
			pal.push_back(make_double2(v.k, v.gamma));
		}
		return it->second;
	};
	size_t he_base = 0;
	for (int b = 0; b < n_bodies; b++) {
		const int n = node_off[b + 1] - node_off[b], ns = spring_off[b + 1] - spring_off[b];
		if (n <= 0 || ns < 0) { ens_free(e); ass_set_error("ass_ensemble_create: body %d is empty or offsets decrease", b); return ASS_EINVAL; }
		const ass_params& p = params[b];
		if ((p.gravity != 0 && p.gravity != 1) || p.additional_forces < 0 || p.additional_forces > 2 || (p.heartbeat != 0 && p.heartbeat != 1)) {
			ens_free(e); ass_set_error("ass_ensemble_create: bad params for body %d", b); return ASS_EINVAL;
		}
		e->max_n = std::max(e->max_n, n);
		BodyDesc& d = e->desc[b];
		d.node_off = node_off[b]; d.n = n; d.row_off = node_off[b] + b;
		int* r = row.data() + d.row_off;  // n + 1 entries
		const ass_spring* sp = springs + spring_off[b];
		std::vector<int> deg((size_t) n + 1, 0);
		for (int q = 0; q < ns; q++) {
			const int i = sp[q].particle_1, j = sp[q].particle_2;
			if (i < 0 || j < 0 || i >= n || j >= n || i == j) {
				ens_free(e);
				ass_set_error("ass_ensemble_create: body %d spring %d joins (%d,%d): indices are local to the body and must differ", b, q, i, j);
				return ASS_EINVAL;
			}
			deg[(size_t) i + 1]++; deg[(size_t) j + 1]++;
		}
		for (int i = 0; i < n; i++) deg[(size_t) i + 1] += deg[i];
		for (int i = 0; i <= n; i++) r[i] = (int) (he_base + deg[i]);
		for (int w0 = 0; w0 < n; w0 += EW) e->ring_records = std::max(e->ring_records, deg[std::min(n, w0 + EW)] - deg[w0]);
		// Within a node's row the records are ordered like the single-body path orders them (state.cu: by the
		// neighbour's spatial rank, ties by spring number), so a member sums in the same order and gets the same bits.
		std::vector<vec4d> pos((size_t) n);
		for (int i = 0; i < n; i++) {
			const double* rr = (const double*) ((const unsigned char*) base + (size_t) (node_off[b] + i) * stride);
			pos[(size_t) i].x = rr[0]; pos[(size_t) i].y = rr[1]; pos[(size_t) i].z = rr[2]; pos[(size_t) i].w = rr[9];
			if (memcmp(&pos[(size_t) i].w, &pos[0].w, sizeof(double)) != 0) e->uni_mass = false;
		}
		std::vector<int> order, rank((size_t) n);
		ass_morton_order(pos.data(), pos.size(), order);
		for (int k = 0; k < n; k++) rank[(size_t) order[(size_t) k]] = k;
		std::vector<std::pair<std::pair<int, int>, int>> ent((size_t) 2 * ns);  // ((neighbour rank, spring), neighbour)
		std::vector<int> fill(deg.begin(), deg.end() - 1);
		for (int q = 0; q < ns; q++) {
			const int ends[2] = { sp[q].particle_1, sp[q].particle_2 };
			for (int side = 0; side < 2; side++) ent[(size_t) fill[ends[side]]++] = { { rank[(size_t) ends[1 - side]], 2 * q + side }, ends[1 - side] };
		}
		for (int i = 0; i < n; i++) std::sort(ent.begin() + deg[(size_t) i], ent.begin() + deg[(size_t) i + 1]);
		for (size_t at = 0; at < ent.size(); at++) {
			const int q = ent[at].first.second >> 1;
			HalfEdge hrec;
			hrec.rs0 = sp[q].rs0; hrec.nbr = ent[at].second; hrec.pal = pal_index(sp[q]);
			he[he_base + at] = hrec;
		}
		he_base += (size_t) 2 * ns;
	}
	e->total_he = (long long) he_base;
	fill_desc(e);
	e->ring_records = std::max(16, (e->ring_records + 7) / 8 * 8);
	e->smem = sizeof(HalfEdge) * 2 * (size_t) e->ring_records + (3 * sizeof(double2) + sizeof(double)) * ((size_t) e->max_n + 1) + sizeof(int) * ((size_t) e->max_n + 1) + 64;
	if (e->smem > 227 * 1024 - 2048) {
		ens_free(e);
		ass_set_error("ass_ensemble_create: a body of %d nodes (%d half-edges per 64-node wave) does not fit in shared memory (limit ~3000 nodes); use ass_sim for large bodies",
		              e->max_n, e->ring_records);
		return ASS_EINVAL;
	}
	const size_t nb = (size_t) total_nodes * stride;
	void* d_rows = nullptr;
	bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e->e0) == cudaSuccess &&
	          cudaEventCreate(&e->e1) == cudaSuccess && cudaMalloc(&e->d_desc, sizeof(BodyDesc) * n_bodies) == cudaSuccess &&
	          cudaMalloc(&e->posm, sizeof(vec4d) * total_nodes) == cudaSuccess && cudaMalloc(&e->vel, sizeof(vec4d) * total_nodes) == cudaSuccess &&
	          cudaMalloc(&e->acc, sizeof(vec4d) * total_nodes) == cudaSuccess && cudaMalloc(&e->row_ptr, sizeof(int) * row.size()) == cudaSuccess &&
	          cudaMalloc(&e->he, sizeof(HalfEdge) * std::max<size_t>(he.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->pal, sizeof(double2) * std::max<size_t>(pal.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->d_dense, sizeof(double) * 10 * (size_t) total_nodes) == cudaSuccess && cudaMalloc(&d_rows, nb) == cudaSuccess;
	if (!ok) {
		cudaFree(d_rows);
		ass_set_error("ass_ensemble_create: %s", cudaGetErrorString(cudaGetLastError()));
		ens_free(e);
		return ASS_ENOMEM;
	}
	cudaMemcpyAsync(d_rows, base, nb, cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->row_ptr, row.data(), sizeof(int) * row.size(), cudaMemcpyHostToDevice, e->stream);
	if (!he.empty()) cudaMemcpyAsync(e->he, he.data(), sizeof(HalfEdge) * he.size(), cudaMemcpyHostToDevice, e->stream);
	if (!pal.empty()) cudaMemcpyAsync(e->pal, pal.data(), sizeof(double2) * pal.size(), cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->d_desc, e->desc.data(), sizeof(BodyDesc) * n_bodies, cudaMemcpyHostToDevice, e->stream);
	k_unpack<<<(total_nodes + 255) / 256, 256, 0, e->stream>>>((const unsigned char*) d_rows, stride, total_nodes, e->posm, e->vel, e->acc);
	g_launch_count++;
	cudaError_t err = cudaStreamSynchronize(e->stream);
	cudaFree(d_rows);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
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
	if (e->uni_mass)
		k_ensemble<true><<<grid, ET, e->smem, e->stream>>>(e->d_desc, e->posm, e->vel, e->acc, e->row_ptr, e->he, e->pal, nsteps, with_heartbeat, e->max_n,
				e->ring_records, e->n_bodies, item_steps, e->d_queue, e->d_queue + 1);
	else
		k_ensemble<false><<<grid, ET, e->smem, e->stream>>>(e->d_desc, e->posm, e->vel, e->acc, e->row_ptr, e->he, e->pal, nsteps, with_heartbeat, e->max_n,
				e->ring_records, e->n_bodies, item_steps, e->d_queue, e->d_queue + 1);
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
	k_pack<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_dense, n, e->posm, e->vel, e->acc);
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
