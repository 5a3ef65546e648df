// springs.cu -- spring_forces (reference src_spring/springs.cpp:288-346) as a node-sorted segmented reduction.
//
// The reference walks the spring list once and scatter-adds +F/m_i, -F/m_j into both endpoints.  Here every node owns
// the list of its incident half-edges (built by ass_set_springs) and a group of SPRING_LANES lanes sums them: no
// atomics, fixed summation order (spring_math.cuh: fold_lanes), so the result is deterministic run to run.  Rows are
// laid out in spatial (Morton) rank order and endpoint state comes from the rank-ordered mirror (common.cuh): through a
// per-tile shared-memory table in the whole-body kernel (k_spring_tiles), through L1 in the CSR kernel that node slabs
// of a sharded body use (k_spring_forces) -- the neighbours of the nodes a warp works on are close in rank.
// The arithmetic is in spring_math.cuh.
//
// Algorithmic traffic: 32 B x NS + 80 B x N per evaluation.  Real DRAM traffic of the tiled kernel (ncu, C3):
// 34 B x NS of half-edge stream (two 16-byte records per spring, + 7 % padding) + ~150 B x N of mirror, acc and outputs;
// the 48 B gathered per half-edge come out of shared memory, the tables are filled from L2 (~3 x 56 B x N).
#include <stdlib.h>
#include <string.h>

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
		vec4d* __restrict__ acc, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, const SpringMirror out, const double dt, const int pre_drift_next,
		const bool write_state = true) {
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
		if (write_state) {  // (the persistent multi-step kernel keeps intermediate steps in the mirror only)
			vel_out[node] = v;
			posm_out[node] = p;
		}
	}
}

// The same epilogue spread over the node's lane group (SPRING_LANES >= 4): after fold_lanes every lane of the group holds
// the node's sum, so lane 0 / 1 / 2 takes the x / y / z component through kick and drift(s) and lane 3 the fourth word of
// every vector -- a third of the dependent FP64 instructions per slice, all lanes of a quarter warp storing one 32-byte
// vector together.  Operation for operation what finish_node does for the component: the same bits.
// a_c: the lane's component of the old acceleration (0 with zero_accel fused in; lane 3: the fourth word, carried through).
template <bool KICK>
__device__ __forceinline__ void finish_node_lanes(const Force3 F, const bool has_springs, const double a_c, const NodeState& ni, const int node,
		const int k, const int sub, vec4d* __restrict__ acc, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, const SpringMirror out,
		const double dt, const int pre_drift_next, const bool write_state) {
	const double inv_m = has_springs ? fast_rcp(ni.m) : 0.0;
	const double Fc = sub == 0 ? F.x : (sub == 1 ? F.y : (sub == 2 ? F.z : 0.0));
	const double a = sub == 3 ? a_c : fma(Fc, inv_m, a_c);
	((double*) (acc + node))[sub] = a;
	if (KICK) {
		const double h = __dmul_rn(0.5, dt);
		const double pc = sub == 0 ? ni.xy.x : (sub == 1 ? ni.xy.y : ni.zv.x);
		const double vc = sub == 0 ? ni.vxy.x : (sub == 1 ? ni.vxy.y : ni.zv.y);
		double v = __dadd_rn(vc, __dmul_rn(dt, a));
		double p = __dadd_rn(pc, __dmul_rn(h, v));
		if (pre_drift_next) {
			p = __dadd_rn(p, __dmul_rn(h, v));
			if (sub < 3) {
				double* const pd = sub == 0 ? &out.xy[k].x : (sub == 1 ? &out.xy[k].y : &out.zv[k].x);
				double* const vd = sub == 0 ? &out.vxy[k].x : (sub == 1 ? &out.vxy[k].y : &out.zv[k].y);
				*pd = p;
				*vd = v;
			}
		}
		if (write_state) {
			((double*) (vel_out + node))[sub] = sub == 3 ? 0.0 : v;
			((double*) (posm_out + node))[sub] = sub == 3 ? ni.m : p;
		}
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
// The tiled kernel (whole body; the default): reads the lane-transposed copy of the half-edge list (common.cuh).
//
// A warp owns 32 / SPRING_LANES nodes of similar degree -- a SLICE -- and walks their rows in lock step: step t of lane
// l is ONE record at warp_off + 32 t + l, so a step of the warp is 512 contiguous bytes and the trip count is the same
// for every lane (no votes, no divergence, no per-row bookkeeping; spring_math.cuh: lane_spring_sum_ring).
// Endpoint state comes out of SHARED memory.  Gathering it through L1 (an earlier version of this kernel) is bounded
// by latency: even at a 90 % L1 hit rate nearly every warp-wide request contains a miss; the ensemble kernel, which
// gathers from shared memory, spends half the time per half-edge.  So the body is treated like an ensemble of TILES.
// CTAs are persistent, one per SM; CTA c owns a contiguous run of slices (cut on the host so that every CTA carries the
// same number of step pairs), cut further into tiles whose nodes AND all their neighbours (~3 x as many: 14 MB of L2
// reads per evaluation at C3) fit in a shared-memory table.  Per tile the CTA copies that state from the rank-ordered
// mirror into the table with cp.async, in two dependent round trips (table entries and slice headers, then the state:
// tiles_pass), then its sixteen warps pull slices from a counter -- no warp waits for a longer share -- and run the
// arithmetic with table slots in place of ranks (heL: records renumbered on the host, slot = rank mod 8 so the
// bank-group-aware row order still holds).  A warp's records reach it through a ring of four 1 KB slots in shared memory
// that the warp fills itself with cp.async, two pairs of steps ahead of the arithmetic and across slice boundaries
// (spring_math.cuh: RecordStream): bytes in flight do not cost registers.  Padding records (rows shorter than the
// slice's longest) point at the node itself: d = 0, a zero contribution.  Same rows, same lanes, same order of summation
// as the CSR kernel: identical bits (tests).
// Shared memory of a CTA: [record rings, 64 KB][table of the current tile][its slice headers]; tiles are cut against
// what the rings leave (TILE_BYTES_LIMIT).
// What bounds it (profiles/r2_spring_kernel_designs.md): instruction issue -- every spring is evaluated at both
// endpoints, 36 FP64 instructions each, and an FP64 instruction holds the issue port for two cycles.
// (Built and measured slower: per-warp bulk-copy rings with slices dealt round-robin; records two pairs ahead in
// registers; requests into L2 32 slices ahead; 768 / 1024 threads with one chain per lane; 1 or 4 lanes per node.)
// ---------------------------------------------------------------------------------------------------
#ifndef ASS_TILE_THREADS
#define ASS_TILE_THREADS 512
#endif
#ifndef ASS_TILE_CHAINS
#define ASS_TILE_CHAINS 2
#endif
constexpr int TL_THREADS = ASS_TILE_THREADS;  // 512 with two interleaved chains per lane, or 1024 with one (half the registers)
constexpr int TL_CHAINS = ASS_TILE_CHAINS;
constexpr int TL_WARPS = TL_THREADS / 32;
#ifndef ASS_TILE_RING
#define ASS_TILE_RING 4
#endif
constexpr int TL_RING = ASS_TILE_RING;                    // pairs of steps (1 KB each) per warp in the record ring (spring_math.cuh: RecordStream)
constexpr int TL_RING_BYTES = TL_WARPS * TL_RING * 1024;  // 64 KB
constexpr int GPW = 32 / SPRING_LANES;
constexpr int TILE_SLOTS_LIMIT = 2944;  // the table's share of the SM's 227 KB (checked below, next to the launcher)
constexpr int TL_FILL = (TILE_SLOTS_LIMIT + TL_THREADS - 1) / TL_THREADS;  // table slots per thread, all in flight at once while the table is filled
struct TileDesc {
	int s0, s1, tab_off, n_slots;
};
static_assert(sizeof(TileDesc) == sizeof(int4), "TileDesc is uploaded as int4");
// What a CTA needs before it can request anything, as KERNEL PARAMETERS: with the descriptors in global memory a launch
// began with a chain of dependent cold misses (tile range -> tile descriptor -> table -> mirror), ~1 us each after an L2 flush.
constexpr int TL_MAX_CTAS = 160;
struct CtaHead {
	int tile0, ntiles;  // this CTA's tiles
	TileDesc first;     // descriptor of the first one
	int sl_end;         // one past the CTA's last slice (bounds the look-ahead requests)
	int pad;
};
struct TilesParams {
	CtaHead h[TL_MAX_CTAS];
};
static_assert(sizeof(TilesParams) <= 8192, "per-CTA heads travel as kernel parameters");

// Grid-wide barrier of the multi-step kernel (k_spring_steps), split in two: a CTA ARRIVES when its writes of a step are out,
// then requests everything of the next step that does not depend on other CTAs (table entries, slice headers, the first
// records of every warp), and only then WAITS for the others -- right before it copies their nodes' new state.
// bar: a zeroed counter; the wait after step s is for (s + 1) * gridDim.x arrivals.
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();  // this CTA's mirror / acc writes before its arrival
		atomicAdd(bar, 1u);
	}
}
__device__ __forceinline__ void grid_wait(const unsigned* bar, unsigned target) {
	if (threadIdx.x == 0) {
		unsigned v;
		do {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
		} while (v < target);
	}
	__syncthreads();
}

// One pass of a CTA over its tiles: the body of both kernels below.
template <bool ZERO_FIRST, bool KICK, bool UNI, bool PAL1>
__device__ __forceinline__ void tiles_pass(const CtaHead& H, unsigned char* tile_smem, int* s_next, const GlobalNodes mir, vec4d* __restrict__ acc,
		const TileDesc* __restrict__ tiles, const int* __restrict__ tab, const int* __restrict__ sl_ord, const int4* __restrict__ grpT,
		const int* __restrict__ warp_off, const HalfEdge* __restrict__ heL, const double2* __restrict__ palf, const double2 nkg1, vec4d* __restrict__ posm_out,
		vec4d* __restrict__ vel_out, const SpringMirror out, double dt, int pre_drift_next, bool write_state, const unsigned* bar = nullptr,
		unsigned bar_target = 0) {
	// [record rings: TL_RING KB per warp][table: xy | zv | vxy | m, one entry per slot of THIS tile][headers: groups | record offsets]
	double2* const s_xy = (double2*) (tile_smem + TL_RING_BYTES);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane & (SPRING_LANES - 1);
	RecordStream<TL_RING> rs;  // this warp's record stream (spring_math.cuh)
	rs.ring = (unsigned) __cvta_generic_to_shared(tile_smem) + (unsigned) (warp * TL_RING * 1024 + lane * 16);
	for (int t = H.tile0; t < H.tile0 + H.ntiles; t++) {
		const TileDesc T = (t == H.tile0) ? H.first : tiles[t];
		const int n_sl = T.s1 - T.s0;
		double2* const s_zv = s_xy + T.n_slots;
		double2* const s_vxy = s_zv + T.n_slots;
		double* const s_m = (double*) (s_vxy + T.n_slots);
		int4* const s_grp = (int4*) (s_m + T.n_slots);      // n_sl * GPW
		int* const s_off = (int*) (s_grp + (size_t) n_sl * GPW);  // n_sl + 1
		int* const s_ord = s_off + n_sl + 1;                      // n_sl: hand-out order of the slices (longest first)
		const SharedNodes nodes{ s_xy, s_zv, s_vxy, s_m };
		__syncthreads();  // everyone has left the previous tile (its table, headers and s_next)
		// ---- before the table exists: this warp's first slice is slice `warp`; its header and first records are requested
		//      now and arrive while the table is being filled ----
		int sw = warp;
		int4 gi = make_int4(-1, -1, 0, 0);
		int off = 0, off1 = 0;
		if (sw < n_sl) {
			gi = __ldg(grpT + (size_t) (T.s0 + sw) * GPW + lane / SPRING_LANES);
			off = __ldg(warp_off + T.s0 + sw);
			off1 = __ldg(warp_off + T.s0 + sw + 1);
		}
		// ---- the table, in TWO dependent round trips (after an L2 flush each costs ~1.5 us): (1) every table entry of this
		//      thread (slot -> rank) together with the slice headers; (2) the state of those ranks, copied global -> shared
		//      without passing through registers (cp.async: all slots of a thread in flight at once), together with the
		//      first records of this warp's slice ----
		int rk[TL_FILL];
#pragma unroll
		for (int q = 0; q < TL_FILL; q++) {
			const int i = tid + q * TL_THREADS;
			rk[q] = (i < T.n_slots) ? __ldg(tab + T.tab_off + i) : -2;
		}
		for (int i = tid; i < n_sl * GPW; i += TL_THREADS) cp_async_16(s_grp + i, grpT + (size_t) T.s0 * GPW + i);
		for (int i = tid; i <= n_sl; i += TL_THREADS) cp_async_4(s_off + i, warp_off + T.s0 + i);
		for (int i = tid; i < n_sl; i += TL_THREADS) cp_async_4(s_ord + i, sl_ord + T.s0 + i);
		int P = (off1 - off) >> 6;  // first use of the headers: after every table request
		rs.begin((const int4*) heL + off + lane, P);
		rs.top_up();
		// (multi-step kernel) everything above is this CTA's own; what follows reads what the other CTAs wrote last step
		if (bar && t == H.tile0) grid_wait(bar, bar_target);
#pragma unroll
		for (int q = 0; q < TL_FILL; q++) {
			const int i = tid + q * TL_THREADS;
			if (rk[q] >= 0) {
				cp_async_16(s_xy + i, mir.xy + rk[q]); cp_async_16(s_zv + i, mir.zv + rk[q]); cp_async_16(s_vxy + i, mir.vxy + rk[q]); cp_async_8(s_m + i, mir.m + rk[q]);
			} else if (rk[q] == -1) {  // a slot without a rank (classes are padded to the fullest): finite values, never a live operand
				s_xy[i] = make_double2(0.0, 0.0); s_zv[i] = make_double2(0.0, 0.0); s_vxy[i] = make_double2(0.0, 0.0); s_m[i] = 1.0;
			}
		}
		cp_async_wait_all();
		if (tid == 0) *s_next = TL_WARPS;
		__syncthreads();
		while (sw < n_sl) {
			// the slice after this one: taken from the counter now, so its header and first records are on their way
			// long before this slice's last pair
			int nx = 0;
			if (lane == 0) {
				nx = atomicAdd(s_next, 1);
				nx = nx < n_sl ? s_ord[nx] : n_sl;  // the counter hands out positions of the order, not slice numbers
			}
			nx = __shfl_sync(0xffffffffu, nx, 0);
			const bool live = gi.x >= 0;
			const int slot = live ? gi.x : 0;
			// the old acceleration is needed after the sum: requested now (SPRING_LANES >= 4: one word per lane, see finish_node_lanes)
			vec4d a_old;
			a_old.x = 0.0; a_old.y = 0.0; a_old.z = 0.0; a_old.w = 0.0;
			double a_old_c = 0.0;
			if (SPRING_LANES >= 4) {
				if (!ZERO_FIRST && live && sub < 4) a_old_c = ((const double*) (acc + gi.w))[sub];
			} else if (!ZERO_FIRST && live && sub == 0) a_old = acc[gi.w];
			const NodeState ni = nodes.get<false>(slot);
			int4 gi_n = make_int4(-1, -1, 0, 0);
			int P_n = 0;
			if (nx < n_sl) {
				gi_n = s_grp[nx * GPW + lane / SPRING_LANES];  // (slot, rank, degree, node); slot < 0: padding group
				const int off_n = s_off[nx];
				P_n = (s_off[nx + 1] - off_n) >> 6;
				rs.next_slice((const int4*) heL + off_n + lane, P_n);  // the stream walks on into it when this slice's records are out
			}
			const Force3 F = lane_spring_sum_ring<UNI, PAL1, TL_CHAINS>(rs, P, ni, nodes, palf, nkg1);
			if (SPRING_LANES >= 4) {
				if (live && sub < 4) finish_node_lanes<KICK>(F, gi.z > 0, a_old_c, ni, gi.w, gi.y, sub, acc, posm_out, vel_out, out, dt, pre_drift_next, write_state);
			} else if (live && sub == 0) finish_node<KICK>(F, gi.z > 0, a_old, ni, gi.w, gi.y, acc, posm_out, vel_out, out, dt, pre_drift_next, write_state);
			sw = nx; gi = gi_n; P = P_n;
		}
	}
}

template <bool ZERO_FIRST, bool KICK, bool UNI, bool PAL1>
__global__ void __launch_bounds__(TL_THREADS, 1) k_spring_tiles(const __grid_constant__ TilesParams heads, const GlobalNodes mir, vec4d* __restrict__ acc,
		const TileDesc* __restrict__ tiles, const int* __restrict__ tab, const int* __restrict__ sl_ord, const int4* __restrict__ grpT,
		const int* __restrict__ warp_off, const HalfEdge* __restrict__ heL, const double2* __restrict__ palf, vec4d* __restrict__ posm_out,
		vec4d* __restrict__ vel_out, const SpringMirror out, double dt, int pre_drift_next) {
	extern __shared__ __align__(128) unsigned char tile_smem[];
	__shared__ int s_next;
	double2 nkg1 = make_double2(0.0, 0.0);
	if (PAL1) nkg1 = __ldg(palf);
	tiles_pass<ZERO_FIRST, KICK, UNI, PAL1>(heads.h[blockIdx.x], tile_smem, &s_next, mir, acc, tiles, tab, sl_ord, grpT, warp_off, heL, palf, nkg1, posm_out, vel_out, out, dt,
			pre_drift_next, true);
}

// The persistent multi-step form for bodies without a force term between the spring evaluations (REB_GRAVITY_NONE, or
// zero_accel right after gravity: every shipped module): nsteps x [springs + kick + drift + next half-drift] in ONE
// launch, the CTAs meeting at a grid-wide barrier after every step.  The state of the intermediate steps lives in the two
// rank-ordered mirrors only (step s reads mirror s & 1 and writes the other one); posm / vel are written by the last step.
// A launch, the parameter fetch and the cold chain tile table -> mirror cost ~8 us per evaluation (12 us of a 1k-node
// body's step are this fixed part); here they are paid once per batch.  Launched cooperatively (all CTAs resident: the
// barrier spins).
template <bool ZERO_FIRST, bool UNI, bool PAL1>
__global__ void __launch_bounds__(TL_THREADS, 1) k_spring_steps(const __grid_constant__ TilesParams heads, const GlobalNodes mirA, const GlobalNodes mirB,
		vec4d* __restrict__ acc, const TileDesc* __restrict__ tiles, const int* __restrict__ tab, const int* __restrict__ sl_ord,
		const int4* __restrict__ grpT, const int* __restrict__ warp_off, const HalfEdge* __restrict__ heL, const double2* __restrict__ palf,
		vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, double dt, int nsteps, unsigned* __restrict__ bar) {
	extern __shared__ __align__(128) unsigned char tile_smem[];
	__shared__ int s_next;
	double2 nkg1 = make_double2(0.0, 0.0);
	if (PAL1) nkg1 = __ldg(palf);
	const SpringMirror outA{ (double2*) mirA.xy, (double2*) mirA.zv, (double2*) mirA.vxy }, outB{ (double2*) mirB.xy, (double2*) mirB.zv, (double2*) mirB.vxy };
	for (int step = 0; step < nsteps; step++) {
		const bool last = step == nsteps - 1;
		const bool even = (step & 1) == 0;
		// (a CTA without tiles still keeps in step: its arrivals must not run ahead of the others')
		if (heads.h[blockIdx.x].ntiles == 0 && step > 0) grid_wait(bar, (unsigned) step * gridDim.x);
		tiles_pass<ZERO_FIRST, true, UNI, PAL1>(heads.h[blockIdx.x], tile_smem, &s_next, even ? mirA : mirB, acc, tiles, tab, sl_ord, grpT, warp_off, heL, palf, nkg1,
				posm_out, vel_out, even ? outB : outA, dt, last ? 0 : 1, last, step > 0 ? bar : nullptr, (unsigned) step * gridDim.x);
		if (!last) grid_arrive(bar);
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
		int* d_flag = (int*) (s->d_red + 48);  // doubles [48..63] of the reduction scratch: nothing else uses them (diagnostics [0..20], mindist [32])
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

// shared memory per table slot of the tiled kernel: three 16-byte pairs and the mass
constexpr size_t TILE_SLOT_BYTES = 3 * sizeof(double2) + sizeof(double);
// and in bytes, slice headers included (16 GPW + 8 bytes per slice: groups, record offset, hand-out order): what is left of the SM's 227 KB next to the record rings
constexpr size_t TILE_SMEM_MAX = 227 * 1024 - 1024;  // dynamic shared memory the kernels opt in to
constexpr size_t TILE_BYTES_LIMIT = TILE_SMEM_MAX - TL_RING_BYTES;
static size_t tile_bytes(int slots, int slices) { return TILE_SLOT_BYTES * (size_t) slots + (size_t) slices * (GPW * sizeof(int4) + 2 * sizeof(int)) + 2 * sizeof(int); }
static_assert(TILE_BYTES_LIMIT >= 64 * 1024, "no room for a table next to the record rings");

template <class T>
static int tiles_reserve(T*& ptr, size_t& cap, size_t need, const char* what) {
	if (need <= cap && ptr) return ASS_OK;
	cudaFree(ptr);
	ptr = nullptr; cap = 0;
	const size_t c = need + need / 8 + 64;
	if (cudaMalloc(&ptr, sizeof(T) * c) != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("out of device memory for %zu %s", c, what);
		return ASS_ENOMEM;
	}
	cap = c;
	return ASS_OK;
}

// What the tiled kernel reads besides the mirror, as built on the host.
struct TilesHost {
	int grid = 0, slots_max = 0, slices_max = 0;
	size_t bytes_max = 0;  // table + headers of the largest tile
	std::vector<TileDesc> tiles;
	std::vector<int> cta_tile, tab;
	std::vector<int> sl_ord;  // per tile, at [s0, s1): local slice indices in hand-out order
	std::vector<HalfEdge> heL;
	std::vector<int4> grpT;
};

// Cut the slices into one run per CTA (equal weight: pairs of steps + a constant per slice for the boundary work) and
// every run into tiles whose node tables fit in `slots_limit` slots; renumber the records.  Pure host work,
// proportional to the half-edge count, done once per spring list.  Returns ASS_EINVAL if a single slice does not fit.
static int build_tiles_host(const SpringLayout& lay, int grid, int slots_limit, TilesHost& out, size_t bytes_limit = TILE_BYTES_LIMIT) {
	const int ns = (int) lay.warp_off.size() - 1;
	const int n = (int) lay.rank.size();
	out.grid = grid;
	auto pairs = [&](int w) { return (lay.warp_off[(size_t) w + 1] - lay.warp_off[w]) >> 6; };
	std::vector<double> prefix((size_t) ns + 1, 0.0);
	double slice_weight = 0.75;  // a slice's own work (header, own state, fold, epilogue) in pairs of steps
	if (const char* e = getenv("ASS_TILE_SLICE_WEIGHT")) slice_weight = atof(e);
	for (int w = 0; w < ns; w++) prefix[(size_t) w + 1] = prefix[w] + (double) pairs(w) + slice_weight;
	std::vector<int> cut((size_t) grid + 1, 0);
	{
		int w = 0;
		for (int c = 1; c < grid; c++) {
			const double target = prefix[ns] * (double) c / (double) grid;
			while (w < ns && prefix[(size_t) w + 1] <= target) w++;
			// the boundary slice goes to whichever side leaves the smaller error
			if (w < ns && target - prefix[w] > prefix[(size_t) w + 1] - target) w++;
			cut[c] = std::max(w, cut[(size_t) c - 1]);
		}
		cut[grid] = ns;
	}
	std::vector<TileDesc>& tiles = out.tiles;
	std::vector<int>& cta_tile = out.cta_tile;
	std::vector<int>& tab = out.tab;
	std::vector<HalfEdge>& heL = out.heL;
	std::vector<int4>& grpT = out.grpT;
	tiles.clear(); tab.clear();
	out.sl_ord.assign((size_t) ns, 0);
	cta_tile.assign((size_t) grid + 1, 0);
	heL = lay.heT;
	grpT.assign(lay.grp.size(), make_int4(-1, -1, 0, 0));
	std::vector<int> stamp((size_t) n, -1), slot_of((size_t) n, -1), members, added;
	for (int c = 0; c < grid; c++) {
		cta_tile[c] = (int) tiles.size();
		int sl = cut[c];
		while (sl < cut[(size_t) c + 1]) {
			// grow a tile slice by slice while its table (8 x the fullest class: slot = rank mod 8) fits
			const int tile_id = (int) tiles.size();
			int cnt[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			members.clear();
			const int s0 = sl;
			while (sl < cut[(size_t) c + 1]) {
				added.clear();
				auto touch = [&](int r) {
					if (stamp[(size_t) r] != tile_id) { stamp[(size_t) r] = tile_id; cnt[r & 7]++; added.push_back(r); }
				};
				for (int g = 0; g < GPW; g++) {
					const int2 gi = lay.grp[(size_t) sl * GPW + g];
					if (gi.x >= 0) touch(gi.x);
				}
				for (int at = lay.warp_off[sl]; at < lay.warp_off[(size_t) sl + 1]; at++) touch(lay.heT[(size_t) at].nbr);
				const int need = 8 * *std::max_element(cnt, cnt + 8);
				const bool too_big = need > TILE_SLOTS_LIMIT || tile_bytes(need, sl - s0 + 1) > bytes_limit;
				if ((need > slots_limit || too_big) && sl > s0) {  // does not fit any more: take the slice back, close the tile
					for (int r : added) { stamp[(size_t) r] = -1; cnt[r & 7]--; }
					break;
				}
				if (too_big) return ASS_EINVAL;  // (a lone slice above a test hook's smaller limit is accepted)
				members.insert(members.end(), added.begin(), added.end());
				sl++;
			}
			// table: ranks ascending inside their class; slot = 8 j + class
			std::sort(members.begin(), members.end());
			int fill[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			const int n_slots = 8 * *std::max_element(cnt, cnt + 8);
			TileDesc T;
			T.s0 = s0; T.s1 = sl; T.tab_off = (int) tab.size(); T.n_slots = n_slots;
			tab.resize(tab.size() + (size_t) n_slots, -1);
			for (int r : members) {
				const int slot = 8 * fill[r & 7]++ + (r & 7);
				slot_of[(size_t) r] = slot;
				tab[(size_t) T.tab_off + slot] = r;
			}
			for (int w = s0; w < sl; w++) {
				for (int g = 0; g < GPW; g++) {
					const int2 gi = lay.grp[(size_t) w * GPW + g];
					if (gi.x >= 0) grpT[(size_t) w * GPW + g] = make_int4(slot_of[(size_t) gi.x], gi.x, gi.y, lay.order[(size_t) gi.x]);
				}
				for (int at = lay.warp_off[w]; at < lay.warp_off[(size_t) w + 1]; at++) {
					const int g = ((at - lay.warp_off[w]) & 31) / SPRING_LANES;
					const bool live = lay.grp[(size_t) w * GPW + g].x >= 0;
					heL[(size_t) at].nbr = live ? slot_of[(size_t) lay.heT[(size_t) at].nbr] : 0;
				}
			}
			// hand-out order: warp w starts with slice w (its header is requested before the table exists); the others go out
			// longest first, so that the last ones pulled -- the tail every warp waits for at the end of the tile -- are the
			// shortest (boundary nodes).  The order of the slices does not touch any sum.
			{
				std::vector<int> rest;
				for (int w = s0; w < sl; w++) {
					if (w - s0 < TL_WARPS) out.sl_ord[(size_t) w] = w - s0;
					else rest.push_back(w - s0);
				}
				std::stable_sort(rest.begin(), rest.end(), [&](int a, int b) { return pairs(s0 + a) > pairs(s0 + b); });
				for (size_t i = 0; i < rest.size(); i++) out.sl_ord[(size_t) s0 + TL_WARPS + i] = rest[i];
			}
			out.slots_max = std::max(out.slots_max, n_slots);
			out.slices_max = std::max(out.slices_max, sl - s0);
			out.bytes_max = std::max(out.bytes_max, tile_bytes(n_slots, sl - s0));
			tiles.push_back(T);
		}
	}
	cta_tile[grid] = (int) tiles.size();
	return ASS_OK;
}

int ass_build_spring_tiles(ass_sim* s, const SpringLayout& lay) {
	s->tiles_grid = 0;
	const int ns = (int) lay.warp_off.size() - 1;
	if (ns <= 0 || lay.heT.empty()) return ASS_OK;
	// small bodies: at least ~4 slices per CTA, so that a few thousand nodes already spread over the whole machine
	int grid = std::max(1, std::min(std::min((ns + 3) / 4, ass_sm_count()), TL_MAX_CTAS));
	int slots_limit = TILE_SLOTS_LIMIT;
	// test hooks: fewer CTAs / smaller tables, so that a small body exercises several tiles per CTA
	if (const char* e = getenv("ASS_TILE_GRID")) grid = std::max(1, std::min(grid, atoi(e)));
	if (const char* e = getenv("ASS_TILE_SLOTS")) slots_limit = std::max(64, std::min(TILE_SLOTS_LIMIT, atoi(e) & ~7));
	TilesHost th;
	int rc = build_tiles_host(lay, grid, slots_limit, th);
	if (rc) return ASS_OK;  // a slice of GPW nodes touches more nodes than a table holds (hubs of degree > ~2900): no tiles, the CSR kernel serves
	if ((rc = tiles_reserve(s->heL, s->heL_cap, th.heL.size(), "renumbered half-edge records"))) return rc;
	if ((rc = tiles_reserve(s->grpT, s->grpT_cap, th.grpT.size(), "tile lane groups"))) return rc;
	if ((rc = tiles_reserve(s->tile_tab, s->tile_tab_cap, th.tab.size(), "tile table entries"))) return rc;
	if ((rc = tiles_reserve(s->tile_desc, s->tile_desc_cap, th.tiles.size(), "tile descriptors"))) return rc;
	if ((rc = tiles_reserve(s->tile_sl_ord, s->tile_sl_ord_cap, th.sl_ord.size(), "slice hand-out order"))) return rc;
	CU_TRY(cudaMemcpyAsync(s->tile_sl_ord, th.sl_ord.data(), sizeof(int) * th.sl_ord.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(s->heL, th.heL.data(), sizeof(HalfEdge) * th.heL.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(s->grpT, th.grpT.data(), sizeof(int4) * th.grpT.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(s->tile_tab, th.tab.data(), sizeof(int) * th.tab.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(s->tile_desc, th.tiles.data(), sizeof(TileDesc) * th.tiles.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	// per-CTA heads: they travel as kernel parameters (k_spring_tiles)
	s->tile_heads.assign((size_t) grid * (sizeof(CtaHead) / sizeof(int)), 0);
	for (int c = 0; c < grid; c++) {
		CtaHead h;
		memset(&h, 0, sizeof(h));
		h.tile0 = th.cta_tile[(size_t) c];
		h.ntiles = th.cta_tile[(size_t) c + 1] - h.tile0;
		if (h.ntiles > 0) {
			h.first = th.tiles[(size_t) h.tile0];
			h.sl_end = th.tiles[(size_t) h.tile0 + h.ntiles - 1].s1;
		}
		memcpy(s->tile_heads.data() + (size_t) c * (sizeof(CtaHead) / sizeof(int)), &h, sizeof(h));
	}
	s->tiles_grid = grid;
	s->tile_slots_max = th.slots_max;
	s->tile_slices_max = th.slices_max;
	s->tile_bytes_max = th.bytes_max;
	return ASS_OK;
}

extern "C" int ass_spring_lanes(void) { return SPRING_LANES; }

// Host-only self-check of everything ass_set_springs builds for the spring kernels (include/ass_b200.h): no device
// is touched, so the CPU test suite can drive it.
extern "C" int ass_selfcheck_spring_layout(const double* pos_xyzm, int n, const ass_spring* springs, int ns, int grid, int slots, long long* stats) {
	REQUIRE(n >= 0 && ns >= 0 && (n == 0 || pos_xyzm) && (ns == 0 || springs) && stats, ASS_EINVAL, "bad argument");
	REQUIRE(grid >= 1, ASS_EINVAL, "grid must be positive");
	std::vector<vec4d> pos((size_t) n);
	for (int i = 0; i < n; i++) { pos[(size_t) i].x = pos_xyzm[4 * i]; pos[(size_t) i].y = pos_xyzm[4 * i + 1]; pos[(size_t) i].z = pos_xyzm[4 * i + 2]; pos[(size_t) i].w = pos_xyzm[4 * i + 3]; }
	std::vector<int> pidx((size_t) ns);
	for (int q = 0; q < ns; q++) pidx[(size_t) q] = q & 3;  // any palette indices: they must travel with their spring
	SpringLayout lay;
	int bad = -1;
	int rc = ass_build_spring_layout(pos.data(), n, springs, ns, pidx.data(), lay, &bad);
	if (rc) {
		ass_set_error("ass_selfcheck_spring_layout: layout failed (rc %d, spring %d)", rc, bad);
		return rc;
	}
	const int n_sl = (int) lay.warp_off.size() - 1;
	TilesHost th;
	const int slots_limit = std::max(64, std::min(TILE_SLOTS_LIMIT, slots & ~7));
	if (n_sl > 0 && (rc = build_tiles_host(lay, std::min(grid, std::max(1, n_sl)), slots_limit, th))) {
		ass_set_error("ass_selfcheck_spring_layout: a slice does not fit in a table");
		return rc;
	}
	long long viol = 0, pad = 0, passes = 0, ideal = 0;
	// (1) every rank owns exactly one lane group; rows hold exactly the springs of their node, with their properties
	std::vector<int> seen((size_t) n, 0);
	for (const int2& g : lay.grp)
		if (g.x >= 0) { seen[(size_t) g.x]++; if (g.y != lay.row[(size_t) g.x + 1] - lay.row[g.x]) viol++; }
	for (int k = 0; k < n; k++) if (seen[(size_t) k] != 1) viol++;
	for (int q = 0; q < ns; q++)
		for (int side = 0; side < 2; side++) {
			const int own = lay.rank[(size_t) (side ? springs[q].particle_2 : springs[q].particle_1)];
			const int oth = lay.rank[(size_t) (side ? springs[q].particle_1 : springs[q].particle_2)];
			const int at = lay.slot[(size_t) 2 * q + side], at_t = lay.slotT[(size_t) 2 * q + side];
			if (at < lay.row[own] || at >= lay.row[(size_t) own + 1]) { viol++; continue; }
			const HalfEdge& h = lay.he[(size_t) at];
			const HalfEdge& ht = lay.heT[(size_t) at_t];
			if (h.nbr != oth || ht.nbr != oth || h.rs0 != springs[q].rs0 || ht.rs0 != springs[q].rs0 || h.pal != (q & 3) || ht.pal != (q & 3)) viol++;
		}
	// (2) the lane-transposed copy is the CSR copy, step by step; padding points at the node itself; step counts are even
	for (int w = 0; w < n_sl; w++) {
		const int steps = (lay.warp_off[(size_t) w + 1] - lay.warp_off[w]) / 32;
		if ((steps & 1) || (lay.warp_off[(size_t) w + 1] - lay.warp_off[w]) % 32) viol++;
		for (int t = 0; t < steps; t++) {
			for (int lane = 0; lane < 32; lane++) {
				const int g = lane / SPRING_LANES, e = t * SPRING_LANES + lane % SPRING_LANES;
				const int2 gi = lay.grp[(size_t) w * GPW + g];
				const HalfEdge& h = lay.heT[(size_t) lay.warp_off[w] + (size_t) t * 32 + lane];
				if (e < gi.y) {
					const HalfEdge& c = lay.he[(size_t) lay.row[gi.x] + e];
					if (c.nbr != h.nbr || c.rs0 != h.rs0 || c.pal != h.pal) viol++;
				} else {
					pad++;
					if (h.nbr != (gi.x >= 0 ? gi.x : 0)) viol++;
				}
			}
			// gather passes of this step: per quarter warp, the fullest bank group among the distinct entries
			for (int qw = 0; qw < 4; qw++) {
				int distinct[8], nd = 0, cnt[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
				bool real = false;
				for (int l8 = 0; l8 < 8; l8++) {
					const int lane = qw * 8 + l8;
					const int2 gi = lay.grp[(size_t) w * GPW + lane / SPRING_LANES];
					if (t * SPRING_LANES + lane % SPRING_LANES < gi.y) real = true;
					const int r = lay.heT[(size_t) lay.warp_off[w] + (size_t) t * 32 + lane].nbr;
					bool dup = false;
					for (int i = 0; i < nd; i++) dup = dup || distinct[i] == r;
					if (!dup) { distinct[nd++] = r; cnt[r & 7]++; }
				}
				if (real) { passes += *std::max_element(cnt, cnt + 8); ideal++; }
			}
		}
	}
	// (3) tiles: CTA runs and tiles partition the slices in order; a table holds every rank its tile touches, at a slot
	//     congruent to the rank mod 8; renumbered records and group headers lead back to the same ranks
	long long expect = 0;
	if (n_sl > 0) {
		if ((int) th.cta_tile.size() != th.grid + 1 || th.cta_tile[0] != 0 || th.cta_tile[(size_t) th.grid] != (int) th.tiles.size()) viol++;
		for (size_t t = 0; t < th.tiles.size(); t++) {
			const TileDesc& T = th.tiles[t];
			if (T.s0 != expect || T.s1 <= T.s0 || (T.n_slots & 7) || T.n_slots > th.slots_max || T.s1 - T.s0 > th.slices_max) viol++;
			if (T.n_slots > slots_limit && T.s1 - T.s0 > 1) viol++;
			{
				std::vector<char> seen_sl((size_t) (T.s1 - T.s0), 0);
				for (int w = T.s0; w < T.s1; w++) {
					const int l = th.sl_ord[(size_t) w];
					if (l < 0 || l >= T.s1 - T.s0 || seen_sl[(size_t) l]++) { viol++; break; }
					if (w - T.s0 < TL_WARPS && l != w - T.s0) viol++;
				}
			}
			if (tile_bytes(T.n_slots, T.s1 - T.s0) > TILE_BYTES_LIMIT || tile_bytes(T.n_slots, T.s1 - T.s0) > th.bytes_max) viol++;  // fits next to the record rings
			expect = T.s1;
			for (int i = 0; i < T.n_slots; i++) {
				const int r = th.tab[(size_t) T.tab_off + i];
				if (r >= 0 && (r & 7) != (i & 7)) viol++;
			}
			for (int w = T.s0; w < T.s1; w++) {
				for (int g = 0; g < GPW; g++) {
					const int2 gi = lay.grp[(size_t) w * GPW + g];
					const int4 gt = th.grpT[(size_t) w * GPW + g];
					if (gi.x < 0) { if (gt.x >= 0) viol++; continue; }
					if (gt.x < 0 || gt.x >= T.n_slots || th.tab[(size_t) T.tab_off + gt.x] != gi.x || gt.y != gi.x || gt.z != gi.y || gt.w != lay.order[(size_t) gi.x]) viol++;
				}
				for (int at = lay.warp_off[w]; at < lay.warp_off[(size_t) w + 1]; at++) {
					const int g = ((at - lay.warp_off[w]) & 31) / SPRING_LANES;
					const HalfEdge& hl = th.heL[(size_t) at];
					const HalfEdge& ht = lay.heT[(size_t) at];
					if (hl.rs0 != ht.rs0 || hl.pal != ht.pal) viol++;
					if (lay.grp[(size_t) w * GPW + g].x < 0) { if (hl.nbr != 0) viol++; continue; }
					if (hl.nbr < 0 || hl.nbr >= T.n_slots || th.tab[(size_t) T.tab_off + hl.nbr] != ht.nbr) viol++;
				}
			}
		}
		if (expect != n_sl) viol++;
	}
	stats[0] = n_sl; stats[1] = pad; stats[2] = (long long) th.tiles.size(); stats[3] = th.slots_max;
	stats[4] = passes; stats[5] = ideal; stats[6] = viol; stats[7] = (long long) lay.heT.size();
	return ASS_OK;
}

template <bool ZERO_FIRST, bool KICK, bool UNI, bool PAL1>
static int launch_tiles_t(ass_sim* s, double dt, int pre) {
	auto kern = k_spring_tiles<ZERO_FIRST, KICK, UNI, PAL1>;
	const size_t smem = TL_RING_BYTES + s->tile_bytes_max;
	// the opt-in above 48 KB is per function AND per device: remembered per device (a process may drive several), set to
	// the most a tile can need so that it never has to be raised again
	static bool opted[64] = {};
	int dev = 0;
	CU_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64 || !opted[dev]) {
		CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TILE_SMEM_MAX));
		if (dev >= 0 && dev < 64) opted[dev] = true;
	}
	TilesParams P;
	memcpy(&P, s->tile_heads.data(), std::min(sizeof(P), s->tile_heads.size() * sizeof(int)));
	kern<<<s->tiles_grid, TL_THREADS, smem, s->stream>>>(P, mirror_nodes(s), s->acc, (const TileDesc*) s->tile_desc, s->tile_tab, s->tile_sl_ord, s->grpT, s->warp_off, s->heL,
			s->palf, KICK ? s->posm_alt : nullptr, KICK ? s->vel_alt : nullptr, s->mir_alt, dt, pre);
	return ASS_OK;
}

template <bool KICK>
static int launch_tiles(ass_sim* s, bool zero_first, bool uni, double dt, int pre) {
	const bool pal1 = s->n_pal == 1;
#define ASS_TILES_U(Z, U) (pal1 ? launch_tiles_t<Z, KICK, U, true>(s, dt, pre) : launch_tiles_t<Z, KICK, U, false>(s, dt, pre))
	if (zero_first) return uni ? ASS_TILES_U(true, true) : ASS_TILES_U(true, false);
	return uni ? ASS_TILES_U(false, true) : ASS_TILES_U(false, false);
#undef ASS_TILES_U
}

// nsteps x [springs + kick + drift (+ the next step's part1)] in one cooperative launch (k_spring_steps)
template <bool ZERO_FIRST, bool UNI, bool PAL1>
static int launch_steps_t(ass_sim* s, double dt, int nsteps) {
	auto kern = k_spring_steps<ZERO_FIRST, UNI, PAL1>;
	size_t smem = TL_RING_BYTES + s->tile_bytes_max;
	static bool opted[64] = {};
	int dev = 0;
	CU_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64 || !opted[dev]) {
		CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TILE_SMEM_MAX));
		if (dev >= 0 && dev < 64) opted[dev] = true;
	}
	if (!s->d_step_bar) CU_TRY(cudaMalloc(&s->d_step_bar, sizeof(unsigned)));
	CU_TRY(cudaMemsetAsync(s->d_step_bar, 0, sizeof(unsigned), s->stream));
	TilesParams P;
	memcpy(&P, s->tile_heads.data(), std::min(sizeof(P), s->tile_heads.size() * sizeof(int)));
	GlobalNodes mirA = mirror_nodes(s);
	GlobalNodes mirB{ s->mir_alt.xy, s->mir_alt.zv, s->mir_alt.vxy, s->mir_m };
	vec4d* acc = s->acc;
	const TileDesc* tiles = (const TileDesc*) s->tile_desc;
	const int* tab = s->tile_tab;
	const int* sl_ord = s->tile_sl_ord;
	const int4* grpT = s->grpT;
	const int* warp_off = s->warp_off;
	const HalfEdge* heL = s->heL;
	const double2* palf = s->palf;
	vec4d *posm_out = s->posm_alt, *vel_out = s->vel_alt;
	unsigned* bar = s->d_step_bar;
	void* args[] = { &P, &mirA, &mirB, &acc, &tiles, &tab, &sl_ord, &grpT, &warp_off, &heL, &palf, &posm_out, &vel_out, &dt, &nsteps, &bar };
	// A device that cannot hold the whole grid at once right now (another context's kernels, MPS limits) refuses the launch
	// instead of dead-locking at the barrier: the caller then steps with one launch per step (return value 1).
	const cudaError_t e = cudaLaunchCooperativeKernel((const void*) kern, dim3((unsigned) s->tiles_grid), dim3(TL_THREADS), args, smem, s->stream);
	if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
		cudaGetLastError();
		return 1;
	}
	CU_TRY(e);
	return ASS_OK;
}

// ASS_SPRINGS_NO_STREAM=1 runs the whole-body evaluations on the CSR kernel instead (same bits: tests)
static bool use_csr_kernel(const ass_sim* s) {
	const char* e = getenv("ASS_SPRINGS_NO_STREAM");
	return s->tiles_grid <= 0 || (e && *e && *e != '0');
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

int ass_springs_prepare(ass_sim* s) {
	bool uni;
	return springs_uniform_mass(s, &uni);
}

// springs + kick + drift(s) for NODES [lo, hi) of a sharded body (shard.cu): gathers from the mirror of posm / vel (all
// nodes of the current buffer), writes the own nodes of posm_alt / vel_alt (the other buffer).  The caller swaps.
int ass_launch_spring_kick_drift_range(ass_sim* s, bool zero_first, double dt, bool pre_drift_next, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	bool uni;
	int rc = springs_uniform_mass(s, &uni);  // cached: shard_prepare asked before the step was queued
	if (rc) return rc;
	ass_tic(s, ASS_K_SPRINGS);
	if ((rc = launch_mirror(s))) return rc;
	launch_csr<true, true>(s, zero_first, uni, lo, hi, dt, pre_drift_next ? 1 : 0);
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
	if (use_csr_kernel(s)) launch_csr<false, false>(s, zero_first, uni, 0, s->n, 0.0, 0);
	else if ((rc = launch_tiles<false>(s, zero_first, uni, 0.0, 0))) return rc;
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
	if (use_csr_kernel(s)) launch_csr<true, false>(s, zero_first, uni, 0, s->n, dt, pre_drift_next ? 1 : 0);
	else if ((rc = launch_tiles<true>(s, zero_first, uni, dt, pre_drift_next ? 1 : 0))) return rc;
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	std::swap(s->posm, s->posm_alt);
	std::swap(s->vel, s->vel_alt);
	if (pre_drift_next) std::swap(s->mir, s->mir_alt);  // the kernel wrote the next evaluation's mirror
	return ASS_OK;
}

// Can a batch of steps run as ONE launch?  Needs the tiled kernel (every CTA resident at once: the grid is at most one CTA
// per SM) and a device that launches cooperatively.  ASS_SPRINGS_NO_MULTISTEP=1 keeps one launch per step (same bits: tests).
bool ass_spring_steps_applicable(const ass_sim* s) {
	const char* e = getenv("ASS_SPRINGS_NO_MULTISTEP");
	if (e && *e && *e != '0') return false;
	return s->n > 0 && s->ns > 0 && !use_csr_kernel(s) && s->tiles_grid <= ass_sm_count() && ass_cooperative_launch_ok();
}

// nsteps whole steps without gravity, heartbeat or any other work between the spring evaluations, starting from pre-drifted
// positions with a fresh mirror: step i = springs + kick + drift, and for all but the last the next step's part1.
// On return posm / vel / acc hold the end-of-batch state (not pre-drifted); the mirror is stale.  Returns 1 (and has done
// nothing) if the device refuses the cooperative launch: the caller falls back to one launch per step.
int ass_launch_spring_steps(ass_sim* s, bool zero_first, double dt, int nsteps) {
	if (s->n == 0 || nsteps <= 0) return ASS_OK;
	bool uni;
	int rc = springs_uniform_mass(s, &uni);
	if (rc) return rc;
	const bool pal1 = s->n_pal == 1;
	ass_tic(s, ASS_K_SPRINGS);
#define ASS_STEPS_U(Z, U) (pal1 ? launch_steps_t<Z, U, true>(s, dt, nsteps) : launch_steps_t<Z, U, false>(s, dt, nsteps))
	if (zero_first) rc = uni ? ASS_STEPS_U(true, true) : ASS_STEPS_U(true, false);
	else rc = uni ? ASS_STEPS_U(false, true) : ASS_STEPS_U(false, false);
#undef ASS_STEPS_U
	if (rc) {
		if (rc > 0) ass_toc(s, ASS_K_SPRINGS);  // refused: nothing ran (the timer's interval stays balanced)
		return rc;
	}
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	std::swap(s->posm, s->posm_alt);
	std::swap(s->vel, s->vel_alt);
	if ((nsteps - 1) & 1) std::swap(s->mir, s->mir_alt);  // (the mirror written last; stale either way: the last step writes none)
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
