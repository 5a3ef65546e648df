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
// A warp owns 32 / SPRING_LANES nodes of similar degree and walks their rows in lock step (spring_math.cuh:
// lane_spring_sum): step t of lane l is ONE record at heT[warp_off + 32 t + l], so the warp reads 512 contiguous bytes
// per step and the trip count is the same for every lane.  Nothing is staged in shared memory: the L1 is left to the
// endpoint gathers, which are 16-byte reads of the rank-ordered mirror -- a group's SPRING_LANES lanes read consecutive entries
// of a row sorted by neighbour rank, i.e. mostly one or two cache lines per group.
// Launch: plain grid, one warp per 32 / SPRING_LANES nodes, LN_THREADS threads per CTA; CTAs of different length
// overlap through the block scheduler.
// ---------------------------------------------------------------------------------------------------
#ifndef ASS_LANES_THREADS
#define ASS_LANES_THREADS 128
#endif
#ifndef ASS_LANES_MINBLOCKS
#define ASS_LANES_MINBLOCKS 4
#endif
constexpr int LN_THREADS = ASS_LANES_THREADS;
template <bool ZERO_FIRST, bool KICK, bool UNI, bool PAL1>
__global__ void __launch_bounds__(LN_THREADS, ASS_LANES_MINBLOCKS) k_spring_lanes(const GlobalNodes nodes, const int* __restrict__ order,
		vec4d* __restrict__ acc, const int2* __restrict__ grp, const int* __restrict__ warp_off, const HalfEdge* __restrict__ heT,
		const double2* __restrict__ palf, int n_swarps, vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, const SpringMirror out, double dt,
		int pre_drift_next) {
	const int gt = blockIdx.x * LN_THREADS + threadIdx.x;
	const int w = gt >> 5;
	if (w >= n_swarps) return;  // whole warps
	const int lane = threadIdx.x & 31;
	const int2 gi = grp[gt / SPRING_LANES];  // (rank, degree); (-1, 0): padding group
	const int off = warp_off[w];
	const int P = (warp_off[w + 1] - off) >> 6;  // pairs of steps (the warp's step count is even)
	const bool live = gi.x >= 0;
	const int k = live ? gi.x : 0;
	// entries l, l + SPRING_LANES, ... of the row are this lane's: nt of them
	const int nt = (gi.y - (lane & (SPRING_LANES - 1)) + SPRING_LANES - 1) / SPRING_LANES;
	const NodeState ni = nodes.get<false>(k);
	const int node = order[k];
	if (!ZERO_FIRST) asm volatile("prefetch.global.L1 [%0];" ::"l"(acc + node));  // needed after the sum
	double2 nkg1 = make_double2(0.0, 0.0);
	if (PAL1) nkg1 = __ldg(palf);
	const Force3 F = lane_spring_sum<UNI, PAL1, true>((const int4*) heT + off + lane, P, nt, ni, nodes, palf, nkg1);
	if (!live || (lane & (SPRING_LANES - 1)) != 0) return;
	vec4d a;
	if (ZERO_FIRST) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
	else a = acc[node];
	finish_node<KICK>(F, gi.y > 0, a, ni, node, k, acc, posm_out, vel_out, out, dt, pre_drift_next);
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

template <bool KICK>
static void launch_lanes(ass_sim* s, bool zero_first, bool uni, double dt, int pre) {
	const int grid = (s->n_swarps * 32 + LN_THREADS - 1) / LN_THREADS;
	const bool pals = s->n_pal == 1;
#define ASS_LANES(Z, U, PS)                                                                                                                   \
	k_spring_lanes<Z, KICK, U, PS><<<grid, LN_THREADS, 0, s->stream>>>(mirror_nodes(s), s->order, s->acc, s->grp, s->warp_off, s->heT, s->palf,    \
			s->n_swarps, KICK ? s->posm_alt : nullptr, KICK ? s->vel_alt : nullptr, s->mir_alt, dt, pre)
#define ASS_LANES_U(Z, U) do { if (pals) ASS_LANES(Z, U, true); else ASS_LANES(Z, U, false); } while (0)
	if (zero_first) { if (uni) ASS_LANES_U(true, true); else ASS_LANES_U(true, false); }
	else { if (uni) ASS_LANES_U(false, true); else ASS_LANES_U(false, false); }
#undef ASS_LANES_U
#undef ASS_LANES
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
	else launch_lanes<false>(s, zero_first, uni, 0.0, 0);
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
	else launch_lanes<true>(s, zero_first, uni, dt, pre_drift_next ? 1 : 0);
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
