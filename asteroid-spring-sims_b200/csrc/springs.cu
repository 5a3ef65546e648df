// springs.cu -- spring_forces (reference src_spring/springs.cpp:288-346) as a node-sorted CSR segmented reduction.
//
// The reference walks the spring list once and scatter-adds +F/m_i, -F/m_j into both endpoints.  Here every node owns
// the list of its incident half-edges (built by ass_set_springs) and a group of SPRING_LANES threads sums them: no
// atomics, fixed summation order, so the result is deterministic run to run.  Rows are laid out in spatial (Morton)
// rank order and endpoint state is gathered from the rank-ordered mirrors ps / vs (common.cuh): each lane streams
// 16-byte HalfEdge records (coalesced) and gathers the other endpoint's position / velocity sector; the neighbours of
// the four nodes a warp works on are close in rank, so one request touches a few adjacent lines.
// The arithmetic is in spring_math.cuh.
//
// Algorithmic traffic: 32 B x NS + 80 B x N per evaluation.  Real traffic of this layout: 32 B x NS of
// half-edge stream (two 16-byte records per spring) + 64 B gathered per half-edge (L1/L2-resident) + ~160 B x N.
#include <stdlib.h>

#include <algorithm>

#include "spring_math.cuh"

namespace {

#ifndef ASS_SPRING_BLOCK
#define ASS_SPRING_BLOCK 256
#endif
#ifndef ASS_SPRING_MINBLOCKS
#define ASS_SPRING_MINBLOCKS 1
#endif
constexpr int BLOCK = ASS_SPRING_BLOCK;  // BLOCK / SPRING_LANES nodes per CTA

// ps[k] = posm[order[k]], vs[k] = vel[order[k]]
__global__ void k_mirror(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const int* __restrict__ order, int n,
		vec4d* __restrict__ ps, vec4d* __restrict__ vs) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const int node = order[k];
	ps[k] = posm[node];
	vs[k] = vel[node];
}

// acc (+)= F / m for one node, then optionally the kick and drift(s) of reb_integrator_leapfrog_part2 (and the next
// step's part1) -- the epilogue shared by both kernels below; executed by one lane.
template <bool KICK>
__device__ __forceinline__ void finish_node(const Force3 F, const bool has_springs, vec4d a, const vec4d pi, const vec4d vi, const int node, const int k,
		vec4d* __restrict__ acc, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, vec4d* __restrict__ ps_out, vec4d* __restrict__ vs_out,
		const double dt, const int pre_drift_next) {
	// a node without springs (e.g. a point-mass perturber, possibly massless) is left untouched
	const double inv_m = has_springs ? fast_rcp(pi.w) : 0.0;
	a.x = fma(F.x, inv_m, a.x);
	a.y = fma(F.y, inv_m, a.y);
	a.z = fma(F.z, inv_m, a.z);
	acc[node] = a;
	if (KICK) {
		// reb_integrator_leapfrog_part2, src/integrator_leapfrog.c:57-63, rounded exactly as the ISO-C build does
		const double h = __dmul_rn(0.5, dt);
		vec4d v = vi, p = pi;
		v.x = __dadd_rn(v.x, __dmul_rn(dt, a.x));
		v.y = __dadd_rn(v.y, __dmul_rn(dt, a.y));
		v.z = __dadd_rn(v.z, __dmul_rn(dt, a.z));
		p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
		p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
		p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
		if (pre_drift_next) {  // the following step's part1 (src/integrator_leapfrog.c:45-47)
			p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
			p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
			p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
			ps_out[k] = p;  // these ARE the positions the next force evaluation sees: keep the mirror current
			vs_out[k] = v;
		}
		vel_out[node] = v;
		posm_out[node] = p;
	}
}

constexpr int NODES_PER_WARP = 32 / SPRING_LANES;

// The plain kernel: one group of SPRING_LANES lanes per node, everything read from global memory.
// ZERO_FIRST  false: acc += springs (spring_forces);  true: acc = springs (zero_accel + spring_forces)
// KICK        additionally v += dt*a ; x += (dt/2)*v [; x += (dt/2)*v again = next step's part1] into the alt buffers
// BY_NODE     false: group g works on RANK lo + g (rows and mirrors stream in order; whole body, used when the streaming
//                    kernel's ring does not fit in shared memory)
//             true:  group g works on NODE lo + g, wherever its row lies (a slab of nodes: multi-GPU shards)
template <bool ZERO_FIRST, bool KICK, bool BY_NODE, bool UNI>
__global__ void __launch_bounds__(BLOCK, ASS_SPRING_MINBLOCKS) k_spring_forces(const vec4d* __restrict__ ps, const vec4d* __restrict__ vs,
		const int* __restrict__ order, const int* __restrict__ rank, vec4d* __restrict__ acc, const int* __restrict__ row_ptr,
		const HalfEdge* __restrict__ he, const double2* __restrict__ pal, int lo, int hi, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, vec4d* __restrict__ ps_out, vec4d* __restrict__ vs_out, double dt,
		int pre_drift_next) {
	const int gid = lo + (blockIdx.x * BLOCK + threadIdx.x) / SPRING_LANES;
	const int lane = threadIdx.x & (SPRING_LANES - 1);
	const bool live = gid < hi;
	const int g = live ? gid : hi - 1;
	const int k = BY_NODE ? rank[g] : g;
	const vec4d pi = ps[k];
	const vec4d vi = vs[k];
	const int beg = row_ptr[k], end = live ? row_ptr[k + 1] : beg;
	const GlobalNodes nodes{ ps, vs };
	const Force3 F = node_spring_sum<true, UNI>(he, pal, beg, end, lane, pi, vi, nodes);
	if (!live || lane != 0) return;
	const int node = BY_NODE ? g : order[k];
	vec4d a;
	if (ZERO_FIRST) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
	else a = acc[node];
	finish_node<KICK>(F, end > beg, a, pi, vi, node, k, acc, posm_out, vel_out, ps_out, vs_out, dt, pre_drift_next);
}

// ---------------------------------------------------------------------------------------------------
// The streaming kernel (whole body, rank order).
//
// The plain kernel above starts every node with a chain of dependent DRAM round trips -- row offsets, then the first
// records, then the endpoint gathers, then (late) the node's old acceleration -- and a CTA lives only as long as that
// chain; measured at C3 (profiles/): 46 us per evaluation, every unit under 50 % busy, warps waiting.  Here CTAs are
// persistent, one per SM, and thread 0 keeps the bulk-copy engine ST_STAGES - 1 waves ahead of the sixteen warps: everything a wave of 64 consecutive ranks needs that is CONTIGUOUS in rank order -- its half-edge records,
// row offsets, own position / velocity, node ids -- arrives in shared memory through cp.async.bulk (SASS UBLKCP)
// completing on an mbarrier (a measured 7 TB/s stream, tools/bulk_probe.cu, that costs the SMs nothing).  The
// consumers' only global-memory waits are the endpoint gathers, which are pipelined one trip ahead (spring_math.cuh)
// and, neighbours being close in rank, touch ~3 cache lines per warp-wide request instead of 32; the old
// accelerations are requested at the top of the wave.  Consumer warps hand a stage back through a second mbarrier and
// never meet at a CTA-wide barrier: a warp with short rows runs ahead of one with long rows.
// (Gathering the wave's distinct neighbours into a shared-memory table first was tried and is slower: eight-byte-pair
// reads at random addresses cost the shared-memory pipe ~2.5 cycles per half-edge.)
// ---------------------------------------------------------------------------------------------------
constexpr int ST_WAVE = ASS_SPRING_WAVE;                 // ranks per wave
constexpr int ST_THREADS = ST_WAVE * SPRING_LANES;       // 512 threads (16 warps, 128 registers each): one group of lanes per rank
#ifndef ASS_STREAM_STAGES
#define ASS_STREAM_STAGES 4
#endif
constexpr int ST_STAGES = ASS_STREAM_STAGES;

struct StageLayout {
	// byte offsets inside one stage; every part starts 32-byte aligned
	int he, ps, vs, rows, ord, bytes;
	__host__ __device__ explicit StageLayout(int wave_he_max) {
		he = 0;
		ps = he + ((wave_he_max * (int) sizeof(HalfEdge) + 31) & ~31);
		vs = ps + ST_WAVE * (int) sizeof(vec4d);
		rows = vs + ST_WAVE * (int) sizeof(vec4d);
		ord = rows + (((ST_WAVE + 4) * (int) sizeof(int) + 31) & ~31);
		bytes = ord + ST_WAVE * (int) sizeof(int);
	}
};

template <bool ZERO_FIRST, bool KICK, bool UNI>
__global__ void __launch_bounds__(ST_THREADS, 1) k_spring_stream(const vec4d* __restrict__ ps, const vec4d* __restrict__ vs,
		const int* __restrict__ order, vec4d* __restrict__ acc, const int* __restrict__ row_ptr, const int* __restrict__ wave_off,
		const int* __restrict__ cta_wave, const HalfEdge* __restrict__ he, const double2* __restrict__ pal, int n, int wave_he_max,
		vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, vec4d* __restrict__ ps_out, vec4d* __restrict__ vs_out, double dt,
		int pre_drift_next) {
	extern __shared__ __align__(128) unsigned char st_smem[];
	__shared__ __align__(8) uint64_t full[ST_STAGES], empty[ST_STAGES];
	__shared__ vec4d a_old[ST_WAVE];  // landing zone of the old accelerations (slot g is private to group g's lane 0)
	const StageLayout L(wave_he_max);
	const int tid = threadIdx.x;
	if (tid == 0) {
		for (int q = 0; q < ST_STAGES; q++) {
			mbar_init(&full[q], 1);
			mbar_init(&empty[q], ST_THREADS / 32);
		}
		mbar_fence_init();
	}
	__syncthreads();
	// thread 0 only: queue the copies of this CTA's `jt`-th wave into stage jt % ST_STAGES (waiting, from the second lap
	// on, until every warp has handed that stage back)
	// this CTA's waves: a contiguous run [w_beg, w_end), cut on the host so that every CTA owns the same number of
	// half-edges (+ a per-node constant), not the same number of waves
	const int w_beg = cta_wave[blockIdx.x], w_end = cta_wave[blockIdx.x + 1];
	auto produce = [&](int jt) {
		const int w = w_beg + jt;
		if (w >= w_end) return;
		const int st = jt % ST_STAGES;
		if (jt >= ST_STAGES) mbar_wait(&empty[st], (uint32_t) (((jt / ST_STAGES) & 1) ^ 1));
		unsigned char* stage = st_smem + (size_t) st * L.bytes;
		const int k0 = w * ST_WAVE, cnt = min(ST_WAVE, n - k0);
		const int e0 = __ldg(wave_off + w), e1 = __ldg(wave_off + w + 1);
		const uint32_t he_bytes = (uint32_t) (e1 - e0) * (uint32_t) sizeof(HalfEdge);
		const uint32_t pv_bytes = (uint32_t) cnt * (uint32_t) sizeof(vec4d);
		const uint32_t row_bytes = (ST_WAVE + 4) * (uint32_t) sizeof(int), ord_bytes = ST_WAVE * (uint32_t) sizeof(int);
		mbar_expect_tx(&full[st], he_bytes + 2 * pv_bytes + row_bytes + ord_bytes);
		if (he_bytes) bulk_g2s(stage + L.he, he + e0, he_bytes, &full[st]);
		bulk_g2s(stage + L.ps, ps + k0, pv_bytes, &full[st]);
		bulk_g2s(stage + L.vs, vs + k0, pv_bytes, &full[st]);
		bulk_g2s(stage + L.rows, row_ptr + k0, row_bytes, &full[st]);
		bulk_g2s(stage + L.ord, order + k0, ord_bytes, &full[st]);
	};
	if (tid == 0)
		for (int jt = 0; jt < ST_STAGES - 1; jt++) produce(jt);
	// ---- consumers: one group of SPRING_LANES lanes per rank of the wave ----
	// The trip pipeline of spring_math.cuh (operands of trip t+1 requested before the arithmetic of trip t) runs ACROSS
	// waves: during the warp's last trip of a wave every lane requests the first trip of its node in the NEXT wave, so
	// the only gather latency a warp ever exposes is that of its very first trip.
	const int g = tid / SPRING_LANES, lane = tid & (SPRING_LANES - 1);
	const GlobalNodes nodes{ ps, vs };
	if (w_beg >= w_end) return;
	// state of the node this lane is working on
	int e, end;
	bool v0, v1;
	const HalfEdge* wave_he;
	TripOperands cur;
	auto open_wave = [&](int it) {  // rows and records of wave w_beg + it are (or become) available: position the lane on its node
		const int st = it % ST_STAGES;
		mbar_wait(&full[st], (uint32_t) ((it / ST_STAGES) & 1));
		const unsigned char* stage = st_smem + (size_t) st * L.bytes;
		const int* rows = (const int*) (stage + L.rows);
		const bool live = (w_beg + it) * ST_WAVE + g < n;
		wave_he = (const HalfEdge*) (stage + L.he) - rows[0];  // indexed with global record numbers
		e = rows[g] + lane;
		end = live ? rows[g + 1] : rows[g];
		v0 = e < end; v1 = e + SPRING_LANES < end;
	};
	open_wave(0);
	if (v0) cur = trip_request(wave_he, pal, e, v1, nodes);
	for (int it = 0; w_beg + it < w_end; it++) {
		const int st = it % ST_STAGES;
		if (tid == 0) produce(it + ST_STAGES - 1);  // refills the stage every warp left one wave ago
		const unsigned char* stage = st_smem + (size_t) st * L.bytes;
		const int k = (w_beg + it) * ST_WAVE + g;
		const bool live = k < n;
		const int node = ((const int*) (stage + L.ord))[g];
		const bool has_springs = ((const int*) (stage + L.rows))[g + 1] > ((const int*) (stage + L.rows))[g];
		// the node's old acceleration: requested now (asynchronously, into shared memory: no register is held and
		// nothing waits), needed after the sum
		if (!ZERO_FIRST && live && lane == 0) {
			cp_async16(&a_old[g], acc + node);
			cp_async16((char*) &a_old[g] + 16, (const char*) (acc + node) + 16);
		}
		const vec4d pi = ((const vec4d*) (stage + L.ps))[live ? g : 0];
		const vec4d vi = ((const vec4d*) (stage + L.vs))[live ? g : 0];
		const bool more_waves = w_beg + it + 1 < w_end;
		double ax = 0.0, ay = 0.0, az = 0.0;
		while (true) {
			// request trip t+1 of this node, or -- on the warp's last trip of the wave -- the first trip of the next wave's
			const int en = e + 2 * SPRING_LANES;
			const bool w0 = v0 && en < end, w1 = v0 && en + SPRING_LANES < end;
			TripOperands nxt = cur;
			const bool warp_last = !__any_sync(0xffffffffu, w0);
			if (w0) nxt = trip_request(wave_he, pal, en, w1, nodes);
			// arithmetic of trip t (the operands were requested one trip ago)
			const TripOperands now = cur;
			const bool n0 = v0, n1 = v1;
			if (warp_last && more_waves) {  // every lane of the warp is on its last trip (or done): look into the next wave
				open_wave(it + 1);
				if (v0) nxt = trip_request(wave_he, pal, e, v1, nodes);
			} else {
				v0 = w0; v1 = w1; e = en;
			}
			if (n0) trip_accumulate<UNI>(now, n1, pi, vi, ax, ay, az);
			cur = nxt;
			if (warp_last) break;
		}
#pragma unroll
		for (int m = SPRING_LANES / 2; m > 0; m >>= 1) {
			ax += shfl_xor_d(ax, m, SPRING_LANES);
			ay += shfl_xor_d(ay, m, SPRING_LANES);
			az += shfl_xor_d(az, m, SPRING_LANES);
		}
		__syncwarp();
		if ((tid & 31) == 0) mbar_arrive(&empty[st]);  // this warp has read everything it needs from the stage
		if (live && lane == 0) {
			vec4d a;
			a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0;
			if (!ZERO_FIRST) {
				cp_async_wait_all();
				a = a_old[g];
			}
			Force3 F;
			F.x = ax; F.y = ay; F.z = az;
			finish_node<KICK>(F, has_springs, a, pi, vi, node, k, acc, posm_out, vel_out, ps_out, vs_out, dt, pre_drift_next);
		}
	}
}

// *differs = 1 if two nodes that carry springs have masses with different bits (rows are by rank, masses by node)
__global__ void k_sprung_mass_differs(const vec4d* __restrict__ posm, const int* __restrict__ order, const int* __restrict__ row_ptr, int n,
		int* __restrict__ differs) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n || row_ptr[k + 1] == row_ptr[k]) return;
	// the reference mass: that of the lowest-ranked node with springs, found by every thread of the first block that has one
	int k0 = 0;
	while (k0 < n && row_ptr[k0 + 1] == row_ptr[k0]) k0++;
	if (__double_as_longlong(posm[order[k]].w) != __double_as_longlong(posm[order[k0]].w)) *differs = 1;
}

}  // namespace

// Do all nodes that carry springs have the same mass (m_red = m / 2, spring_math.cuh)?  Cached per (masses, topology).
static int springs_uniform_mass(ass_sim* s, bool* uni) {
	if (s->uni_mass_epoch != s->mass_epoch || s->uni_spring_epoch != s->spring_epoch) {
		int* d_flag = (int*) s->d_red + 32;  // tail of the simulation's small reduction scratch
		CU_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), s->stream));
		if (s->n > 0 && s->ns > 0) {
			k_sprung_mass_differs<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->posm, s->order, s->row_ptr, s->n, d_flag);
			LAUNCH_CHECK();
		}
		int differs = 1;
		CU_TRY(cudaMemcpyAsync(&differs, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CU_TRY(cudaStreamSynchronize(s->stream));
		const char* off = getenv("ASS_SPRINGS_NO_UNIFORM");
		s->uni_mass = !differs && !(off && *off && *off != '0');
		s->uni_mass_epoch = s->mass_epoch;
		s->uni_spring_epoch = s->spring_epoch;
	}
	*uni = s->uni_mass;
	return ASS_OK;
}

// Cut the waves into `grid` contiguous runs of equal weight (half-edges + a constant per node for the row epilogue): a
// CTA's share of the work, not of the wave count, is what must be equal -- surface waves carry half the springs of
// interior ones.  Cached per (grid, spring list).
static int ensure_cta_waves(ass_sim* s, int grid) {
	if (s->cta_wave && s->cta_wave_grid == grid && s->cta_wave_epoch == s->spring_epoch) return ASS_OK;
	const int nw = s->n_waves;
	std::vector<double> prefix((size_t) nw + 1, 0.0);
	for (int w = 0; w < nw; w++)
		prefix[(size_t) w + 1] = prefix[w] + (double) (s->h_wave_off[(size_t) w + 1] - s->h_wave_off[w]) + 6.0 * ASS_SPRING_WAVE;
	std::vector<int> cut((size_t) grid + 1, nw);
	cut[0] = 0;
	int w = 0;
	for (int c = 1; c < grid; c++) {
		const double target = prefix[nw] * (double) c / (double) grid;
		while (w < nw && prefix[(size_t) w + 1] <= target) w++;
		// the boundary wave goes to whichever side leaves the smaller error
		if (w < nw && target - prefix[w] > prefix[(size_t) w + 1] - target) w++;
		cut[c] = std::max(w, cut[(size_t) c - 1]);
	}
	if ((size_t) grid + 1 > s->cta_wave_cap) {
		cudaFree(s->cta_wave);
		s->cta_wave = nullptr; s->cta_wave_cap = 0;
		if (cudaMalloc(&s->cta_wave, sizeof(int) * ((size_t) grid + 1)) != cudaSuccess) {
			cudaGetLastError();
			ass_set_error("out of device memory for %d wave cuts", grid + 1);
			return ASS_ENOMEM;
		}
		s->cta_wave_cap = (size_t) grid + 1;
	}
	CU_TRY(cudaMemcpyAsync(s->cta_wave, cut.data(), sizeof(int) * cut.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	s->cta_wave_grid = grid;
	s->cta_wave_epoch = s->spring_epoch;
	return ASS_OK;
}

// Launch the streaming kernel if one stage ring fits in shared memory; returns 1 if it did not (caller falls back).
template <bool ZERO_FIRST, bool KICK, bool UNI>
static int launch_stream_t(ass_sim* s, double dt, bool pre_drift_next) {
	const StageLayout L(s->wave_he_max);
	const size_t smem = (size_t) L.bytes * ST_STAGES;
	if (smem > 200 * 1024 || getenv("ASS_SPRINGS_NO_STREAM")) return 1;
	auto kern = k_spring_stream<ZERO_FIRST, KICK, UNI>;
	static int ctas_per_sm = 0;  // per instantiation
	static size_t smem_set = 0;
	if (smem > smem_set) {
		CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		smem_set = smem;
		ctas_per_sm = 0;
	}
	if (ctas_per_sm == 0) {
		CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, ST_THREADS, smem_set));
		if (ctas_per_sm < 1) return 1;
	}
	const int grid = std::min(s->n_waves, ass_sm_count() * ctas_per_sm);
	int rc = ensure_cta_waves(s, grid);
	if (rc) return rc;
	kern<<<grid, ST_THREADS, smem_set, s->stream>>>(s->ps, s->vs, s->order, s->acc, s->row_ptr, s->wave_off, s->cta_wave, s->he, s->pal, s->n,
			s->wave_he_max, s->posm_alt, s->vel_alt, s->ps_alt, s->vs_alt, dt, pre_drift_next ? 1 : 0);
	LAUNCH_CHECK();
	return ASS_OK;
}

template <bool ZERO_FIRST, bool KICK>
static int launch_stream(ass_sim* s, double dt, bool pre_drift_next, bool uni) {
	return uni ? launch_stream_t<ZERO_FIRST, KICK, true>(s, dt, pre_drift_next) : launch_stream_t<ZERO_FIRST, KICK, false>(s, dt, pre_drift_next);
}

// the plain kernel, all template axes chosen at run time
template <bool KICK, bool BY_NODE>
static void launch_plain(ass_sim* s, int grid, bool zero_first, bool uni, int lo, int hi, double dt, int pre) {
#define ASS_PLAIN(Z, U)                                                                                                                         \
	k_spring_forces<Z, KICK, BY_NODE, U><<<grid, BLOCK, 0, s->stream>>>(s->ps, s->vs, s->order, s->rank, s->acc, s->row_ptr, s->he, s->pal, lo, hi, \
			KICK ? s->posm_alt : nullptr, KICK ? s->vel_alt : nullptr, KICK ? s->ps_alt : nullptr, KICK ? s->vs_alt : nullptr, dt, pre)
	if (zero_first) { if (uni) ASS_PLAIN(true, true); else ASS_PLAIN(true, false); }
	else { if (uni) ASS_PLAIN(false, true); else ASS_PLAIN(false, false); }
#undef ASS_PLAIN
}

static int launch_mirror(ass_sim* s) {
	k_mirror<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, s->order, s->n, s->ps, s->vs);
	LAUNCH_CHECK();
	return ASS_OK;
}

// a (+)= springs for NODES [lo, hi)
int ass_launch_spring_forces_range(ass_sim* s, bool zero_first, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	bool uni;
	int rc = springs_uniform_mass(s, &uni);
	if (rc) return rc;
	const int groups_per_block = BLOCK / SPRING_LANES;
	const int grid = (hi - lo + groups_per_block - 1) / groups_per_block;
	ass_tic(s, ASS_K_SPRINGS);
	if ((rc = launch_mirror(s))) return rc;
	launch_plain<false, true>(s, grid, zero_first, uni, lo, hi, 0.0, 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	return ASS_OK;
}

// a (+)= springs for the whole body
int ass_launch_spring_forces(ass_sim* s, bool zero_first, bool mirror_fresh) {
	if (s->n == 0) return ASS_OK;
	bool uni;
	int rc = springs_uniform_mass(s, &uni);
	if (rc) return rc;
	ass_tic(s, ASS_K_SPRINGS);
	if (!mirror_fresh && (rc = launch_mirror(s))) return rc;
	rc = zero_first ? launch_stream<true, false>(s, 0.0, false, uni) : launch_stream<false, false>(s, 0.0, false, uni);
	if (rc < 0) return rc;
	if (rc == 1) {  // the ring does not fit in shared memory: plain kernel over ranks
		const int groups_per_block = BLOCK / SPRING_LANES;
		launch_plain<false, false>(s, (s->n + groups_per_block - 1) / groups_per_block, zero_first, uni, 0, s->n, 0.0, 0);
		LAUNCH_CHECK();
	}
	ass_toc(s, ASS_K_SPRINGS);
	return ASS_OK;
}

int ass_launch_spring_kick_drift(ass_sim* s, bool zero_first, double dt, bool pre_drift_next, bool mirror_fresh) {
	if (s->n == 0) return ASS_OK;
	bool uni;
	int rc = springs_uniform_mass(s, &uni);
	if (rc) return rc;
	ass_tic(s, ASS_K_SPRINGS);
	if (!mirror_fresh && (rc = launch_mirror(s))) return rc;
	rc = zero_first ? launch_stream<true, true>(s, dt, pre_drift_next, uni) : launch_stream<false, true>(s, dt, pre_drift_next, uni);
	if (rc < 0) return rc;
	if (rc == 1) {
		const int groups_per_block = BLOCK / SPRING_LANES;
		launch_plain<true, false>(s, (s->n + groups_per_block - 1) / groups_per_block, zero_first, uni, 0, s->n, dt, pre_drift_next ? 1 : 0);
		LAUNCH_CHECK();
	}
	ass_toc(s, ASS_K_SPRINGS);
	std::swap(s->posm, s->posm_alt);
	std::swap(s->vel, s->vel_alt);
	if (pre_drift_next) {  // the kernel wrote the next evaluation's mirror
		std::swap(s->ps, s->ps_alt);
		std::swap(s->vs, s->vs_alt);
	}
	return ASS_OK;
}

extern "C" int ass_spring_forces(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (s->ns == 0 || s->row_ptr == nullptr) return ASS_OK;
	if (s->shard.active) return ass_launch_spring_forces_range(s, false, s->shard.lo, s->shard.hi);
	return ass_launch_spring_forces(s, false, false);
}
