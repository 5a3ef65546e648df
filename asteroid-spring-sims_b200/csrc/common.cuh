// common.cuh -- internal declarations shared by the sm_100a translation units of libass_b200.so.
// Nothing here is part of the C-ABI (include/ass_b200.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/ass_b200.h"

// ---------------------------------------------------------------------------------------------------
// Device data layout ("SoA of 32-byte vectors")
//   posm[i] = {x, y, z, m}      one 32 B sector per gather
//   vel[i]  = {vx, vy, vz, 0}
//   acc[i]  = {ax, ay, az, 0}
// A gravity source tile is a dense run of posm.  The spring kernels gather from a rank-ordered mirror of posm / vel in
// 16-byte pairs (ass_sim::mir).  HalfEdge is the node-sorted copy of `struct spring`, one per endpoint.
// ---------------------------------------------------------------------------------------------------
struct __align__(32) vec4d {
	double x, y, z, w;
};

// One endpoint's view of a spring, 16 bytes: the rest length, the other endpoint, and an index into the palette of
// distinct (k, gamma) pairs.  A body has a handful of distinct pairs (set_gamma makes gamma uniform, adjust_spring_props
// paints a few radial shells), so two half-edges cost 32 B per spring -- exactly sizeof(struct spring), the algorithmic
// figure of SURVEY 8d -- although every spring is stored at both of its endpoints.  Lossless for any number of pairs.
struct __align__(16) HalfEdge {
	double rs0;
	int nbr;  // the other endpoint: its spatial rank (see ass_sim::order); in heL: its slot in the tile's shared-memory table
	int pal;  // index into the (k, gamma) palette
};
static_assert(sizeof(HalfEdge) == 16, "HalfEdge must be 16 bytes");
static_assert(sizeof(ass_spring) == 32, "ass_spring must match struct spring (32 B)");

#define ASS_L_EPS 1e-6  // src_spring/springs.cpp:28
#define ASS_LEN2_FLOOR 1e-300  // added to every squared spring length: exact above 1e-284, keeps coincident endpoints finite (spring_math.cuh)
#ifndef ASS_SPRING_LANES
#define ASS_SPRING_LANES 2
#endif
constexpr int SPRING_LANES = ASS_SPRING_LANES;  // lanes cooperating on one node's half-edge list (1, 2, 4 or 8)
static_assert(SPRING_LANES == 1 || SPRING_LANES == 2 || SPRING_LANES == 4 || SPRING_LANES == 8, "SPRING_LANES must be 1, 2, 4 or 8");
#define ASS_MAX_MASKS 4
#define ASS_SORT_BLOCK 256  // ranks per block inside which nodes are dealt to lane groups by degree (ass_sim::grp)

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
void ass_set_error(const char* fmt, ...);
extern long long g_launch_count;

#define CU_TRY(expr)                                                                              \
	do {                                                                                          \
		cudaError_t _e = (expr);                                                                  \
		if (_e != cudaSuccess) {                                                                  \
			ass_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
			return ASS_ECUDA;                                                                     \
		}                                                                                         \
	} while (0)

#define REQUIRE(cond, code, msg)                              \
	do {                                                      \
		if (!(cond)) {                                        \
			ass_set_error("%s: %s", __func__, msg);           \
			return (code);                                    \
		}                                                     \
	} while (0)

// ---------------------------------------------------------------------------------------------------
// the simulation object
// ---------------------------------------------------------------------------------------------------
struct DevBuf {
	void* p = nullptr;
	size_t bytes = 0;
};

// device-side meeting points of a sharded body's ranks (shard.cu): 64-bit epochs, one row per phase, one slot per rank
#define ASS_SHARD_MAX_RANKS 16
#define ASS_SHARD_PHASES 3
#define ASS_SHARD_POS 0     // slab carries the drifted positions / velocities of step e
#define ASS_SHARD_PART 1    // partial accelerations of step e for the other slabs are complete
#define ASS_SHARD_GATHER 2  // 2c - 1: slab final for gather c;  2c: this rank has pulled everybody's

struct ShardInfo {
	int rank = 0, nranks = 1;
	int lo = 0, hi = 0;               // owned slab of targets / spring owners
	std::vector<int> bounds;          // nranks + 1 slab boundaries
	// one exportable allocation [posm0 | vel0 | posm1 | vel1 | partial accelerations | flags]; sim->posm / vel point at
	// buffer `cur`, sim->posm_alt / vel_alt at the other one
	vec4d* window = nullptr;
	int cur = 0;
	unsigned long long epoch = 0;         // steps started since ass_shard_setup
	unsigned long long gather_epoch = 0;  // gathers started
	unsigned long long gather_pending = 0;  // GATHER value every peer must reach before the own slab may move again
	std::vector<vec4d*> peer_window;  // peers' windows mapped here (nullptr for self)
	std::vector<char> peer_is_ipc;    // mapped through cudaIpcOpenMemHandle (needs closing) or a same-process pointer
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t pulled = nullptr;
	cudaEvent_t drifted = nullptr;                                // main stream: own slab drifted (and the previous step's reductions queued before it)
	cudaStream_t grav_stream[3] = { nullptr, nullptr, nullptr };  // pair kernels of the rectangles / the antipodal half, next to the triangle
	cudaEvent_t grav_done[3] = { nullptr, nullptr, nullptr };
	bool active = false;
};

// rank-ordered node state as the spring kernels gather it (see ass_sim::mir)
struct SpringMirror {
	double2 *xy = nullptr, *zv = nullptr, *vxy = nullptr;  // (x, y), (z, vz), (vx, vy)
};

struct ass_sim {
	int n = 0, cap = 0;     // particles
	int ns = 0;             // springs
	int n_he = 0;           // half-edges (2*ns)
	ass_params prm;
	cudaStream_t stream = nullptr;

	vec4d *posm = nullptr, *vel = nullptr, *acc = nullptr;
	vec4d *posm_alt = nullptr, *vel_alt = nullptr;  // ping-pong targets of the fused spring+kick+drift kernel
	long long mass_epoch = 0;  // bumped whenever masses may have changed on the device (uploads carrying ASS_F_MASS)
	long long spring_epoch = 0;  // bumped by ass_set_springs
	// "every node that carries springs has the same mass" (springs.cu), valid for (uni_mass_epoch, uni_spring_epoch)
	bool uni_mass = false;
	long long uni_mass_epoch = -1, uni_spring_epoch = -1;
	bool have_state = false;
	bool predrifted = false;  // positions already carry the next step's first half-drift (fused loop only)
	bool sum_sequential = true;  // CoM / CoV sums in the reference's particle order (bit-exact) or as a parallel tree
	bool sum_order_forced = false;  // ass_set_sum_order was called: otherwise the built-in per-step heartbeat uses the tree
	int gravity_kernel = 0;      // 0 auto, 1 one-sided tiles (gravity.cu), 2 pair-once (gravity_sym.cu)
	std::vector<struct SymPlan*> sym_plans;  // CTA tables + scratch rows of the pair-once kernel, one per (resident, streamed) range pair

	// springs: CSR over nodes in SPATIAL order.  ass_set_springs ranks the nodes along a Morton curve through their
	// positions (a solid body: neighbours stay neighbours); row k belongs to node order[k], HalfEdge::nbr is the
	// neighbour's rank, and the mirror (below) holds positions / velocities in rank order.  The spring kernels walk rows
	// in rank order and gather endpoint state from the mirror, so the lanes of one request touch a handful of adjacent
	// cache lines instead of 32 random ones (the generators emit nodes in random order).  Everything else keeps the
	// caller's node order.
	HalfEdge* he = nullptr;   // CSR copy (rows by rank, each row sorted by neighbour rank): diagnostics, node-slab kernel
	int* row_ptr = nullptr;   // n+1, by rank
	int* slot = nullptr;      // 2*ns: position in he of (spring s, endpoint e)
	int* order = nullptr;     // rank -> node
	int* rank = nullptr;      // node -> rank
	// The rank-ordered mirror of posm / vel, as three arrays of 16-byte pairs -- (x, y), (z, vz), (vx, vy) -- plus the
	// masses.  A random gather of 32-byte {x,y,z,m} entries can start on only four of the eight 16-byte bank groups of
	// the L1 data array and is processed in two passes (measured: 22 data wavefronts per warp-wide 256-bit gather);
	// 16-byte entries use all eight groups, and the endpoint costs 48 bytes instead of 64.
	SpringMirror mir, mir_alt;  // mir_alt: written by the fused kick (a node may not move while neighbours still read it)
	double2* mir_block = nullptr;  // the one allocation behind mir and mir_alt
	double* mir_m = nullptr;
	size_t sorted_cap = 0;
	// The lane-transposed copy of the half-edge list, read by the whole-body spring kernel (springs.cu: k_spring_tiles)
	// and, per member, by the ensemble kernel.  Groups of SPRING_LANES lanes own one node; 32 / SPRING_LANES groups make
	// a warp's SLICE; within blocks of ASS_SORT_BLOCK ranks the nodes are dealt to groups in order of decreasing degree,
	// so the groups of a slice have rows of similar length.  Slice w owns records [warp_off[w], warp_off[w+1]): step t of
	// lane l is record warp_off[w] + 32 t + l, so a warp reads 512 contiguous bytes per step; lane l of a group takes
	// entries l, l + SPRING_LANES, ... of its node's row; rows are padded (with records that are loaded and computed but
	// enter the sum with a zero coefficient) to the slice's longest, rounded up to an even number of steps.
	int* slotT = nullptr;     // 2*ns: position in the lane-transposed copy of (spring s, endpoint e)
	int* warp_off = nullptr;  // n_swarps + 1
	int n_swarps = 0;
	size_t swarp_cap = 0;
	// On the device the copy lives in the tiled kernel's form: every persistent CTA's run of slices is cut into TILES
	// whose nodes AND their neighbours fit in shared memory; tile_tab lists each tile's table (slot -> rank, -1: unused;
	// slot = rank mod 8, so the bank-group-aware row order survives), heL holds the records with HalfEdge::nbr renumbered
	// to table slots, grpT[g] = (slot, rank, degree, node) of lane group g, (-1, -1, 0, 0) for padding groups.  Built by
	// ass_set_springs for one CTA per SM.
	HalfEdge* heL = nullptr;
	int4* grpT = nullptr;
	int* tile_tab = nullptr;
	int* tile_sl_ord = nullptr;  // per tile: the order in which its slices are handed out (local indices; the longest first after the warps' fixed first ones)
	int4* tile_desc = nullptr;  // (first slice, end slice, offset into tile_tab, slots)
	std::vector<int> tile_heads;  // per CTA: (first tile, tile count, first tile's descriptor, last slice): kernel parameters
	int tiles_grid = 0, tile_slots_max = 0, tile_slices_max = 0;
	size_t tile_bytes_max = 0;  // shared memory of the largest tile (table + slice headers)
	size_t heL_cap = 0, grpT_cap = 0, tile_tab_cap = 0, tile_desc_cap = 0, tile_sl_ord_cap = 0;
	double2* pal = nullptr;   // palette of distinct (k, gamma) pairs
	double2* palf = nullptr;  // the same entries as the force kernels want them: (-k, -max(gamma, 0)); lives behind pal
	int n_pal = 0;
	size_t he_cap = 0, row_cap = 0, pal_cap = 0;
	int max_degree = 0;

	// node masks for the masked force terms (stress.cu: apply_impact's surface nodes, the cube module's top / bottom faces)
	unsigned char* masks = nullptr;  // ASS_MAX_MASKS x mask_cap
	size_t mask_cap = 0;

	// scratch
	DevBuf stage;             // raw host rows on the device
	void* h_pinned = nullptr; // pinned bounce buffer
	size_t h_pinned_bytes = 0;
	DevBuf grav_partial;      // [split][n] vec4d
	double* d_red = nullptr;  // reduction outputs (device)
	unsigned* d_step_bar = nullptr;  // arrival counter of the multi-step spring kernel's grid barrier
	double* h_red = nullptr;  // pinned mirror
	DevBuf red_partial;
	DevBuf flush;             // L2 flush target

	// timing
	bool timing = false;
	double k_ms[ASS_K_COUNT];
	long long k_launches[ASS_K_COUNT];
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
	std::vector<std::pair<int, int>> ev_open;  // (class, pool index) pending
	cudaEvent_t mark0 = nullptr, mark1 = nullptr;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> lap_ev;  // ass_lap_begin / ass_lap_end
	int lap_open = 0;
	bool lap_flush = false;  // ass_step_timed: flush L2 and lap every step of the fused loop

	ShardInfo shard;
};

void ass_shard_release(ass_sim* s);  // shard.cu
int ass_shard_sync(ass_sim* s);
// gravity_sym.cu
void ass_sym_plans_free(ass_sim* s);
bool ass_gravity_sym_applicable(int count);
int ass_launch_gravity_sym(ass_sim* s, int lo, int hi);
int ass_launch_gravity_sym_pair(ass_sim* s, int rlo, int rhi, int slo, int shi, vec4d* res_out, bool res_add, vec4d* str_out, bool str_add);
int ass_launch_gravity_sym_main(ass_sim* s, int rlo, int rhi, int slo, int shi, cudaStream_t st);  // the pair kernel alone, on any stream
int ass_launch_gravity_sym_finish(ass_sim* s, int rlo, int rhi, int slo, int shi, vec4d* res_out, bool res_add, vec4d* str_out, bool str_add);
// build (or refresh) the plan of a range pair now: whatever allocation and synchronisation it needs happens here, not in the launch
int ass_gravity_sym_prepare(ass_sim* s, int rlo, int rhi, int slo, int shi);
int ass_buf_reserve(DevBuf& b, size_t bytes);
// order[k] = node of rank k along a Morton curve through pos[0..n) (state.cu); deterministic, ties keep node order
void ass_morton_order(const vec4d* pos, size_t n, std::vector<int>& order);
// everything the spring kernels read, built on the host (state.cu); the fields mirror ass_sim's
struct SpringLayout {
	std::vector<int> order, rank, row, slot, slotT, warp_off;
	std::vector<HalfEdge> he, heT;
	std::vector<int2> grp;
	int max_degree = 0;
};
int ass_build_spring_layout(const vec4d* pos, int n, const ass_spring* sp, int ns, const int* pidx, SpringLayout& out, int* bad_spring);
int ass_build_spring_tiles(ass_sim* s, const SpringLayout& lay);  // springs.cu; queues uploads on s->stream and waits for them
int ass_ensure_device(void);
int ass_sm_count(void);
bool ass_cooperative_launch_ok(void);

// RAII-free timing bracket: tic before a launch, toc after it
void ass_tic(ass_sim* s, int klass);
void ass_toc(ass_sim* s, int klass);

#define LAUNCH_CHECK()                                                                        \
	do {                                                                                      \
		g_launch_count++;                                                                     \
		cudaError_t _e = cudaGetLastError();                                                  \
		if (_e != cudaSuccess) {                                                              \
			ass_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
			return ASS_ECUDA;                                                                 \
		}                                                                                     \
	} while (0)

// ---------------------------------------------------------------------------------------------------
// internal launchers (one per .cu file)
// ---------------------------------------------------------------------------------------------------
// springs.cu
// mirror_fresh: the mirror already holds posm / vel (the caller just produced them); otherwise they are re-gathered first
int ass_launch_spring_forces(ass_sim* s, bool zero_first, bool mirror_fresh);
int ass_launch_spring_forces_range(ass_sim* s, bool zero_first, int lo, int hi);
int ass_launch_spring_kick_drift(ass_sim* s, bool zero_first, double dt, bool pre_drift_next, bool mirror_fresh);
bool ass_spring_steps_applicable(const ass_sim* s);
int ass_launch_spring_steps(ass_sim* s, bool zero_first, double dt, int nsteps);  // a batch of gravity-free steps in one launch
// the same for NODES [lo, hi) of a sharded body: reads posm / vel (all nodes), writes posm_alt / vel_alt (own nodes); no swap
int ass_launch_spring_kick_drift_range(ass_sim* s, bool zero_first, double dt, bool pre_drift_next, int lo, int hi);
int ass_springs_prepare(ass_sim* s);  // the equal-mass verdict (may synchronise), taken before a step is queued
// gravity.cu
int ass_launch_gravity(ass_sim* s);
int ass_gravity_rows(int n_tgt, int n_src);
int ass_launch_gravity_partial(ass_sim* s, const vec4d* src, int s_lo, int s_hi, int t_lo, int t_hi, int row_base, int* rows_out);
int ass_launch_gravity_reduce(ass_sim* s, int t_lo, int t_hi, int rows, bool add = false);
// leapfrog.cu
int ass_launch_part1(ass_sim* s, double dt);
int ass_launch_part1_mirrored(ass_sim* s, double dt);  // part1 that also refreshes the rank-ordered mirror
int ass_launch_part2(ass_sim* s, double dt);
int ass_launch_part1_range(ass_sim* s, double dt, int lo, int hi);
int ass_launch_part2_range(ass_sim* s, double dt, bool pre_drift_next, int lo, int hi);
int ass_launch_zero_accel(ass_sim* s);
int ass_launch_zero_accel_range(ass_sim* s, int lo, int hi);
// reduce.cu
int ass_launch_com(ass_sim* s, int lo, int hi, bool velocity, double* d_out3_and_mass);
int ass_launch_shift(ass_sim* s, int lo, int hi, const double* d_shift3, bool velocity, double sign, double drift_half_dt);

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// Vector::len() of the reference as gcc compiles it (matrix_math.cpp:158-166 under GNU C++ -O3 with FMA):
// sqrt(fma(z,z, fma(x,x, y*y))), every operation correctly rounded.  The intrinsics pin both the nesting and the
// rounding so nvcc's own contraction cannot change the bits.
__device__ __forceinline__ double ref_len2(double x, double y, double z) {
	return __fma_rn(z, z, __fma_rn(x, x, __dmul_rn(y, y)));
}
__device__ __forceinline__ double ref_len(double x, double y, double z) {
	return __dsqrt_rn(ref_len2(x, y, z));
}
__device__ __forceinline__ double ref_dot(double lx, double ly, double lz, double rx, double ry, double rz) {
	return __fma_rn(lz, rz, __fma_rn(lx, rx, __dmul_rn(ly, ry)));
}

// 1/a to ~1 ulp without the IEEE-division slow path: hardware seed (MUFU.RCP64H, ~2^-20) + two Newton steps.
__device__ __forceinline__ double fast_rcp(double a) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
	double e = fma(-a, y, 1.0);
	y = fma(y, e, y);
	e = fma(-a, y, 1.0);
	y = fma(y, e, y);
	return y;
}

// a^(-3/2) in SIX FP64 instructions from the hardware seed y0 = MUFU.RSQ64H(a).  The seed carries 20 mantissa bits (the
// instruction produces the high word only), so c2 = y0*y0 is exact and eps = 1 - a*c2 costs one FMA with one rounding:
//   a^(-3/2) = y0^3 (1 - eps)^(-3/2) = y0^3 (1 + eps (3/2 + 15/8 eps)) + O(eps^3),   eps ~ 2^-21: 35/16 eps^3 ~ 2^-62.
// y0^3 does not wait for the correction: the dependent chain is 4 deep after the seed.
__device__ __forceinline__ double rsqrt_cubed(double a) {
	double y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
	const double c2 = y0 * y0;
	const double eps = fma(-a, c2, 1.0);
	const double c = c2 * y0;
	const double q = eps * fma(1.875, eps, 1.5);
	return fma(c, q, c);
}

// ---- per-thread asynchronous copies global -> shared (LDGSTS): no registers, any number in flight ----
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {  // past L1
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// the same with the destination as a 32-bit shared-memory address, groups, and a plain 16-byte shared-memory load
__device__ __forceinline__ void cp_async_16s(unsigned dst_smem, const void* src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int4 lds_128(unsigned src_smem) {
	int4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src_smem) : "memory");
	return v;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- bulk-copy engine (TMA, 1-D, no tensor map) + mbarrier: global -> shared staging with completion by byte count ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t"
		"}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// size and both addresses must be multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
			"l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double shfl_xor_d(double v, int mask, int width = 32) {
	return __shfl_xor_sync(0xffffffffu, v, mask, width);
}
__device__ __forceinline__ double shfl_down_d(double v, int delta, int width = 32) {
	return __shfl_down_sync(0xffffffffu, v, delta, width);
}
#endif
