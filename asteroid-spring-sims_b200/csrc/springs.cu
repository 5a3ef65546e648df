// springs.cu -- spring_forces (reference src_spring/springs.cpp:288-346) as a node-sorted segmented reduction.
//
// The reference walks the spring list once and scatter-adds +F/m_i, -F/m_j into both endpoints.  Here every node owns
// the list of its incident half-edges (built by ass_set_springs) and a group of SPRING_LANES lanes sums them: no
// atomics, fixed summation order (spring_math.cuh: fold_lanes), so the result is deterministic run to run.  Rows are
// laid out in spatial (Morton) rank order and endpoint state is gathered from the rank-ordered mirror (common.cuh):
// the neighbours of the nodes a warp works on are close in rank, so one request touches a few adjacent cache lines.
// The arithmetic is in spring_math.cuh.
//
// Algorithmic traffic: 32 B x NS + 80 B x N per evaluation.  Real traffic of this layout: 32 B x NS of
// half-edge stream (two 16-byte records per spring, + ~6 % padding) + 48 B gathered per half-edge (L1/L2-resident)
// + ~170 B x N.
#include <stdlib.h>

#include <algorithm>

#include "spring_math.cuh"

namespace {

#ifndef ASS_SPRING_BLOCK
#define ASS_SPRING_BLOCK 256
#endif
constexpr int BLOCK = ASS_SPRING_BLOCK;  // BLOCK / SPRING_LANES nodes per CTA of the CSR kernel

// mir[k] = (posm, vel)[order[k]]
__global__ void k_mirror(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const int* __restrict__ order, int n, SpringMirror mir,
		double* __restrict__ mir_m) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const int node = order[k];
	const vec4d p = posm[node], v = vel[node];
	mir.xy[k] = make_double2(p.x, p.y);
	mir.zv[k] = make_double2(p.z, v.z);
	mir.vxy[k] = make_double2(v.x, v.y);
	mir_m[k] = p.w;
}

// acc (+)= F / m for one node, then optionally the kick and drift(s) of reb_integrator_leapfrog_part2 (and the next
// step's part1) -- the epilogue shared by both kernels below; executed by one lane.
template <bool KICK>
__device__ __forceinline__ void finish_node(const Force3 F, const bool has_springs, vec4d a, const NodeState& ni, const int node, const int k,
		vec4d* __restrict__ acc, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, const SpringMirror out, const double dt, const int pre_drift_next) {
	// a node without springs (e.g. a point-mass perturber, possibly massless) is left untouched
	const double inv_m = has_springs ? fast_rcp(ni.m) : 0.0;
	a.x = fma(F.x, inv_m, a.x);
	a.y = fma(F.y, inv_m, a.y);
	a.z = fma(F.z, inv_m, a.z);
	acc[node] = a;
	if (KICK) {
		// reb_integrator_leapfrog_part2, src/integrator_leapfrog.c:57-63, rounded exactly as the ISO-C build does
		const double h = __dmul_rn(0.5, dt);
		vec4d v, p;
		v.x = __dadd_rn(ni.vxy.x, __dmul_rn(dt, a.x));
		v.y = __dadd_rn(ni.vxy.y, __dmul_rn(dt, a.y));
		v.z = __dadd_rn(ni.zv.y, __dmul_rn(dt, a.z));
		v.w = 0.0;
		p.x = __dadd_rn(ni.xy.x, __dmul_rn(h, v.x));
		p.y = __dadd_rn(ni.xy.y, __dmul_rn(h, v.y));
		p.z = __dadd_rn(ni.zv.x, __dmul_rn(h, v.z));
		p.w = ni.m;
		if (pre_drift_next) {  // the following step's part1 (src/integrator_leapfrog.c:45-47)
			p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
			p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
			p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
			// these ARE the positions the next force evaluation sees: keep the mirror current
			out.xy[k] = make_double2(p.x, p.y);
			out.zv[k] = make_double2(p.z, v.z);
			out.vxy[k] = make_double2(v.x, v.y);
		}
		vel_out[node] = v;
		posm_out[node] = p;
	}
}

// The CSR kernel: one group of SPRING_LANES lanes per node, rows read from the CSR copy of the half-edge list.
// ZERO_FIRST  false: acc += springs (spring_forces);  true: acc = springs (zero_accel + spring_forces)
// KICK        additionally v += dt*a ; x += (dt/2)*v [; x += (dt/2)*v again = next step's part1] into the alt buffers
// BY_NODE     false: group g works on RANK lo + g (whole body; the reference kernel the lane kernel is tested against)
//             true:  group g works on NODE lo + g, wherever its row lies (a slab of nodes: multi-GPU shards)
template <bool ZERO_FIRST, bool KICK, bool BY_NODE, bool UNI>
__global__ void __launch_bounds__(BLOCK) k_spring_forces(const GlobalNodes nodes, const int* __restrict__ order, const int* __restrict__ rank,
		vec4d* __restrict__ acc, const int* __restrict__ row_ptr, const HalfEdge* __restrict__ he, const double2* __restrict__ palf, int lo, int hi,
		vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, const SpringMirror out, double dt, int pre_drift_next) {
	const int gid = lo + (blockIdx.x * BLOCK + threadIdx.x) / SPRING_LANES;
	const int lane = threadIdx.x & (SPRING_LANES - 1);
	const bool live = gid < hi;
	const int g = live ? gid : hi - 1;
	const int k = BY_NODE ? rank[g] : g;
	const NodeState ni = nodes.get<false>(k);
	const int beg = row_ptr[k], end = live ? row_ptr[k + 1] : beg;
	const Force3 F = node_spring_sum<true, UNI>(he, palf, beg, end, lane, ni, nodes);
	if (!live || lane != 0) return;
	const int node = BY_NODE ? g : order[k];
	vec4d a;
	if (ZERO_FIRST) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
	else a = acc[node];
	finish_node<KICK>(F, end > beg, a, ni, node, k, acc, posm_out, vel_out, out, dt, pre_drift_next);
}

// ---------------------------------------------------------------------------------------------------
// The lane kernel (whole body): reads the lane-transposed copy of the half-edge list (ass_sim::heT, common.cuh).
//
// A warp owns 32 / SPRING_LANES nodes of similar degree -- a SLICE -- and walks their rows in lock step: step t of lane
// l is ONE record at heT[warp_off + 32 t + l], so a step of the warp is 512 contiguous bytes and the trip count is
// the same for every lane (no votes, no divergence, no per-row bookkeeping).  CTAs are persistent, one per SM: CTA c owns
// the contiguous run of slices [ccut[c], ccut[c+1]) (cut on the host so that every CTA carries the same number of step
// pairs) and its sixteen warps take them round-robin, so at any moment the whole SM works on ~128 nodes that are
// neighbours in space and the L1 holds one neighbourhood, not sixteen.  A slice's records are one contiguous run of
// heT: each warp streams its records through a private ring in shared memory with the bulk-copy engine
// (cp.async.bulk -> mbarrier, SASS UBLKCP; LN_STAGES pairs of steps = 4 KB ahead, across its slices; producer and
// consumer are the same warp, so handing a stage back needs no second barrier), and its software
// pipeline -- records of pair p+1 out of the ring, their endpoint gathers issued, arithmetic of pair p (two interleaved
// FP64 chains per lane, spring_math.cuh) -- runs straight across slice boundaries: at a boundary the warp folds its
// partial sums, runs the node epilogue (acc, optionally kick + drift + mirror) and switches to the next slice's nodes,
// whose header was read one slice ago and whose state was prefetched one pair ago.  The only exposed memory latency
// of a warp is that of its very first pair.
// The loop is unrolled by two over explicit operand sets A / B so that the pipeline costs no register moves.
// Padding records (rows shorter than the slice's longest) point at the node itself and enter the sum with a zero
// coefficient.  The L1 is left to the endpoint gathers: 16-byte reads of the rank-ordered mirror -- a group's
// SPRING_LANES lanes read consecutive entries of a row sorted by neighbour rank, i.e. mostly one or two cache lines.
// ---------------------------------------------------------------------------------------------------
#ifndef ASS_LANES_THREADS
#define ASS_LANES_THREADS 512
#endif
#ifndef ASS_LANES_MINBLOCKS
#define ASS_LANES_MINBLOCKS 1
#endif
#ifndef ASS_LANES_STAGES
#define ASS_LANES_STAGES 4
#endif
constexpr int LN_THREADS = ASS_LANES_THREADS;
constexpr int LN_WARPS = LN_THREADS / 32;
constexpr int LN_STAGES = ASS_LANES_STAGES;  // power of two
constexpr int LN_PAIR_BYTES = 64 * (int) sizeof(HalfEdge);  // two steps of a warp
static_assert((LN_STAGES & (LN_STAGES - 1)) == 0, "LN_STAGES must be a power of two");
constexpr int GPW = 32 / SPRING_LANES;

template <bool ZERO_FIRST, bool KICK, bool UNI, bool PAL1>
__global__ void __launch_bounds__(LN_THREADS, ASS_LANES_MINBLOCKS) k_spring_lanes(const GlobalNodes nodes, const int* __restrict__ order,
		vec4d* __restrict__ acc, const int2* __restrict__ grp, const int* __restrict__ warp_off, const int* __restrict__ ccut, const int* __restrict__ wpairs,
		const HalfEdge* __restrict__ heT, const double2* __restrict__ palf, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out,
		const SpringMirror out, double dt, int pre_drift_next) {
	extern __shared__ __align__(128) unsigned char ring_all[];  // [LN_WARPS][LN_STAGES][LN_PAIR_BYTES]
	__shared__ __align__(8) uint64_t full_all[LN_WARPS][LN_STAGES];
	const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
	const int sub = lane & (SPRING_LANES - 1);
	unsigned char(*ring)[LN_PAIR_BYTES] = (unsigned char(*)[LN_PAIR_BYTES]) (ring_all + (size_t) wi * LN_STAGES * LN_PAIR_BYTES);
	uint64_t* full = full_all[wi];
	if (lane == 0) {
#pragma unroll
		for (int q = 0; q < LN_STAGES; q++) mbar_init(&full[q], 1);
		mbar_fence_init();
	}
	__syncwarp();
	// this CTA's slices: [c0, s1); warp wi takes c0 + wi, c0 + wi + LN_WARPS, ...
	const int c0 = ccut[blockIdx.x], s1 = ccut[blockIdx.x + 1];
	int s = c0 + wi;
	const int Ptot = wpairs[blockIdx.x * LN_WARPS + wi];  // pairs of steps over all of this warp's slices
	if (s >= s1) return;  // whole warps
	// lane 0 only: the producer's cursor runs up to LN_STAGES pairs ahead of the consumers, through the same slices
	int p_slice = s - LN_WARPS, p_left = 0;
	const unsigned char* p_src = nullptr;
	int p_q = 0;
	auto produce = [&]() {  // the next pair of the warp's stream -> ring stage p_q % LN_STAGES
		while (p_left == 0) {  // (a loop: slices without springs have no pairs; Ptot bounds the calls, so a slice exists)
			p_slice += LN_WARPS;
			const int off = __ldg(warp_off + p_slice);
			p_left = (__ldg(warp_off + p_slice + 1) - off) >> 6;
			p_src = (const unsigned char*) (heT + off);
		}
		uint64_t* bar = &full[p_q & (LN_STAGES - 1)];
		mbar_expect_tx(bar, LN_PAIR_BYTES);
		bulk_g2s(ring[p_q & (LN_STAGES - 1)], p_src, LN_PAIR_BYTES, bar);
		p_src += LN_PAIR_BYTES;
		p_left--;
		p_q++;
	};
	if (lane == 0)
		for (int q = 0; q < LN_STAGES && q < Ptot; q++) produce();
	double2 nkg1 = make_double2(0.0, 0.0);
	if (PAL1) nkg1 = __ldg(palf);
	// ---- the slice this warp is on ----
	int2 gi, gi_next;     // (rank, degree) of this lane's group; (-1, 0): padding group
	int P_s, P_next;      // pairs of steps of the slice
	int lp = 0, nt = 0, k = 0, node = 0;
	bool live = false;
	NodeState ni;
	double ax = 0.0, ay = 0.0, az = 0.0;
	auto read_header = [&](int sl, int2& g, int& P) {
		g = __ldg(grp + (size_t) sl * GPW + lane / SPRING_LANES);
		P = (__ldg(warp_off + sl + 1) - __ldg(warp_off + sl)) >> 6;
	};
	auto open_slice = [&]() {  // slice s: its header was read one slice ago
		gi = gi_next; P_s = P_next;
		if (s + LN_WARPS < s1) read_header(s + LN_WARPS, gi_next, P_next);
		live = gi.x >= 0;
		k = live ? gi.x : 0;
		nt = (gi.y - sub + SPRING_LANES - 1) / SPRING_LANES;  // entries sub, sub + SPRING_LANES, ... of the row are this lane's
		ni = nodes.get<false>(k);
		node = __ldg(order + k);
		if (!ZERO_FIRST) asm volatile("prefetch.global.L1 [%0];" ::"l"(acc + node));  // needed after the sum
		lp = 0;
		ax = 0.0; ay = 0.0; az = 0.0;
	};
	auto close_slice = [&]() {
		fold_lanes(ax, ay, az);
		if (live && sub == 0) {
			vec4d a;
			if (ZERO_FIRST) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
			else a = acc[node];
			Force3 F;
			F.x = ax; F.y = ay; F.z = az;
			finish_node<KICK>(F, gi.y > 0, a, ni, node, k, acc, posm_out, vel_out, out, dt, pre_drift_next);
		}
	};
	// after the arithmetic of a pair: step the slice; false when the warp's run is exhausted
	auto retire = [&]() -> bool {
		lp++;
		if (lp + 1 == P_s && s + LN_WARPS < s1 && gi_next.x >= 0) {  // one pair to go: bring the next slice's own state close
			asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes.xy + gi_next.x));
			asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes.zv + gi_next.x));
			asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes.vxy + gi_next.x));
		}
		while (lp == P_s) {  // (a loop: slices without springs have no pairs)
			close_slice();
			if ((s += LN_WARPS) >= s1) return false;
			open_slice();
		}
		return true;
	};
	auto fetch = [&](int q) {  // records of pair q out of the ring, their gathers issued, the stage refilled
		const int st = q & (LN_STAGES - 1);
		mbar_wait(&full[st], (uint32_t) ((q / LN_STAGES) & 1));
		const int4 r0 = *(const int4*) (ring[st] + lane * 16);
		const int4 r1 = *(const int4*) (ring[st] + 512 + lane * 16);
		PairOperands o;
		o.rs0[0] = __hiloint2double(r0.y, r0.x); o.rs0[1] = __hiloint2double(r1.y, r1.x);
		o.pal[0] = r0.w; o.pal[1] = r1.w;
		o.nj[0] = nodes.get<UNI>(r0.z);
		o.nj[1] = nodes.get<UNI>(r1.z);
		// the gathers above could not be issued before the whole warp's record reads had returned: the stage is free
		__syncwarp();
		if (lane == 0 && q + LN_STAGES < Ptot) produce();
		return o;
	};
	auto compute = [&](const PairOperands& o) {
		const bool valid[2] = { 2 * lp < nt, 2 * lp + 1 < nt };
		const double2 nkg[2] = { PAL1 ? nkg1 : __ldg(palf + o.pal[0]), PAL1 ? nkg1 : __ldg(palf + o.pal[1]) };
		half_edges_accumulate<2, UNI>(o.rs0, nkg, o.nj, valid, ni, ax, ay, az);
	};
	read_header(s, gi_next, P_next);
	open_slice();
	while (P_s == 0) {  // leading slices without springs
		close_slice();
		if ((s += LN_WARPS) >= s1) return;
		open_slice();
	}
	PairOperands A = fetch(0), B;
	for (int p = 0; p < Ptot; p += 2) {
		if (p + 1 < Ptot) B = fetch(p + 1);
		compute(A);
		if (!retire()) break;
		if (p + 2 < Ptot) A = fetch(p + 2);
		compute(B);
		if (!retire()) break;
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

static GlobalNodes mirror_nodes(const ass_sim* s) { return GlobalNodes{ s->mir.xy, s->mir.zv, s->mir.vxy, s->mir_m }; }

// the CSR kernel, all template axes chosen at run time
template <bool KICK, bool BY_NODE>
static void launch_csr(ass_sim* s, bool zero_first, bool uni, int lo, int hi, double dt, int pre) {
	const int groups_per_block = BLOCK / SPRING_LANES;
	const int grid = (hi - lo + groups_per_block - 1) / groups_per_block;
#define ASS_CSR(Z, U)                                                                                                                        \
	k_spring_forces<Z, KICK, BY_NODE, U><<<grid, BLOCK, 0, s->stream>>>(mirror_nodes(s), s->order, s->rank, s->acc, s->row_ptr, s->he, s->palf, lo, \
			hi, KICK ? s->posm_alt : nullptr, KICK ? s->vel_alt : nullptr, s->mir_alt, dt, pre)
	if (zero_first) { if (uni) ASS_CSR(true, true); else ASS_CSR(true, false); }
	else { if (uni) ASS_CSR(false, true); else ASS_CSR(false, false); }
#undef ASS_CSR
}

// Cut the slices into one contiguous run per CTA, of equal weight (pairs of steps + a constant per slice for the
// boundary work): a CTA's share of the work, not of the slice count, is what must be equal -- surface slices carry half
// the springs of interior ones.  wpairs[c * LN_WARPS + w] = pairs of steps of warp w's slices.  Cached per (grid, list).
static int ensure_cta_cuts(ass_sim* s, int grid) {
	if (s->ccut && s->ccut_grid == grid && s->ccut_epoch == s->spring_epoch) return ASS_OK;
	const int ns = s->n_swarps;
	auto pairs = [&](int w) { return (s->h_warp_off[(size_t) w + 1] - s->h_warp_off[w]) >> 6; };
	std::vector<double> prefix((size_t) ns + 1, 0.0);
	for (int w = 0; w < ns; w++) prefix[(size_t) w + 1] = prefix[w] + (double) pairs(w) + 0.75;
	std::vector<int> cut((size_t) grid + 1 + (size_t) grid * LN_WARPS, 0);  // [cuts | wpairs]
	int w = 0;
	for (int c = 1; c < grid; c++) {
		const double target = prefix[ns] * (double) c / (double) grid;
		while (w < ns && prefix[(size_t) w + 1] <= target) w++;
		// the boundary slice goes to whichever side leaves the smaller error
		if (w < ns && target - prefix[w] > prefix[(size_t) w + 1] - target) w++;
		cut[c] = std::max(w, cut[(size_t) c - 1]);
	}
	cut[grid] = ns;
	for (int c = 0; c < grid; c++)
		for (int sl = cut[c]; sl < cut[(size_t) c + 1]; sl++) cut[(size_t) grid + 1 + (size_t) c * LN_WARPS + (sl - cut[c]) % LN_WARPS] += pairs(sl);
	if (cut.size() > s->ccut_cap) {
		cudaFree(s->ccut);
		s->ccut = nullptr; s->ccut_cap = 0;
		if (cudaMalloc(&s->ccut, sizeof(int) * cut.size()) != cudaSuccess) {
			cudaGetLastError();
			ass_set_error("out of device memory for %zu slice cuts", cut.size());
			return ASS_ENOMEM;
		}
		s->ccut_cap = cut.size();
	}
	CU_TRY(cudaMemcpyAsync(s->ccut, cut.data(), sizeof(int) * cut.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	s->ccut_grid = grid;
	s->ccut_epoch = s->spring_epoch;
	return ASS_OK;
}

template <bool ZERO_FIRST, bool KICK, bool UNI, bool PAL1>
static int launch_lanes_t(ass_sim* s, double dt, int pre) {
	auto kern = k_spring_lanes<ZERO_FIRST, KICK, UNI, PAL1>;
	static int ctas_per_sm = 0;  // per instantiation
	constexpr size_t smem = (size_t) LN_WARPS * LN_STAGES * LN_PAIR_BYTES;
	if (ctas_per_sm == 0) {
		CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, LN_THREADS, smem));
		if (ctas_per_sm < 1) ctas_per_sm = 1;
	}
	const int grid = std::max(1, std::min((s->n_swarps + LN_WARPS - 1) / LN_WARPS, ass_sm_count() * ctas_per_sm));
	int rc = ensure_cta_cuts(s, grid);
	if (rc) return rc;
	kern<<<grid, LN_THREADS, smem, s->stream>>>(mirror_nodes(s), s->order, s->acc, s->grp, s->warp_off, s->ccut, s->ccut + grid + 1, s->heT, s->palf,
			KICK ? s->posm_alt : nullptr, KICK ? s->vel_alt : nullptr, s->mir_alt, dt, pre);
	return ASS_OK;
}

template <bool KICK>
static int launch_lanes(ass_sim* s, bool zero_first, bool uni, double dt, int pre) {
	const bool pal1 = s->n_pal == 1;
#define ASS_LANES_U(Z, U) (pal1 ? launch_lanes_t<Z, KICK, U, true>(s, dt, pre) : launch_lanes_t<Z, KICK, U, false>(s, dt, pre))
	if (zero_first) return uni ? ASS_LANES_U(true, true) : ASS_LANES_U(true, false);
	return uni ? ASS_LANES_U(false, true) : ASS_LANES_U(false, false);
#undef ASS_LANES_U
}

// ASS_SPRINGS_NO_STREAM=1 runs the whole-body evaluations on the CSR kernel instead (same bits: tests)
static bool use_csr_kernel() {
	const char* e = getenv("ASS_SPRINGS_NO_STREAM");
	return e && *e && *e != '0';
}

static int launch_mirror(ass_sim* s) {
	k_mirror<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, s->order, s->n, s->mir, s->mir_m);
	LAUNCH_CHECK();
	return ASS_OK;
}

// a (+)= springs for NODES [lo, hi)
int ass_launch_spring_forces_range(ass_sim* s, bool zero_first, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	bool uni;
	int rc = springs_uniform_mass(s, &uni);
	if (rc) return rc;
	ass_tic(s, ASS_K_SPRINGS);
	if ((rc = launch_mirror(s))) return rc;
	launch_csr<false, true>(s, zero_first, uni, lo, hi, 0.0, 0);
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
	if (use_csr_kernel()) launch_csr<false, false>(s, zero_first, uni, 0, s->n, 0.0, 0);
	else if ((rc = launch_lanes<false>(s, zero_first, uni, 0.0, 0))) return rc;
	LAUNCH_CHECK();
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
	if (use_csr_kernel()) launch_csr<true, false>(s, zero_first, uni, 0, s->n, dt, pre_drift_next ? 1 : 0);
	else if ((rc = launch_lanes<true>(s, zero_first, uni, dt, pre_drift_next ? 1 : 0))) return rc;
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	std::swap(s->posm, s->posm_alt);
	std::swap(s->vel, s->vel_alt);
	if (pre_drift_next) std::swap(s->mir, s->mir_alt);  // the kernel wrote the next evaluation's mirror
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
