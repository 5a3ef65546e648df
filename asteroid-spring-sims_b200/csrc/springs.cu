// springs.cu -- spring_forces (reference src_spring/springs.cpp:288-346) as a node-sorted CSR segmented reduction.
//
// The reference walks the spring list once and scatter-adds +F/m_i, -F/m_j into both endpoints.  Here every node owns
// the list of its incident half-edges (built by ass_set_springs) and a group of SPRING_LANES threads sums them: no
// atomics, fixed summation order, so the result is deterministic run to run.  Each lane streams 32-byte HalfEdge
// records (coalesced: SPRING_LANES consecutive sectors) and gathers the other endpoint's posm / vel sectors.
// The arithmetic is in spring_math.cuh.
//
// Algorithmic traffic: 32 B x NS + 80 B x N per evaluation.  Real traffic of this layout: 64 B x NS of
// half-edge stream (each spring is stored at both endpoints) + 64 B gathered per half-edge (L2-resident) + 96 B x N.
#include "spring_math.cuh"

namespace {

constexpr int BLOCK = 256;  // 32 nodes per CTA

// ZERO_FIRST  false: acc += springs (spring_forces);  true: acc = springs (zero_accel + spring_forces)
// KICK        additionally v += dt*a ; x += (dt/2)*v [; x += (dt/2)*v again = next step's part1] into the alt buffers
template <bool ZERO_FIRST, bool KICK>
__global__ void __launch_bounds__(BLOCK) k_spring_forces(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel,
		vec4d* __restrict__ acc, const int* __restrict__ row_ptr, const HalfEdge* __restrict__ he, const double2* __restrict__ pal, int lo, int hi,
		vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, double dt, int pre_drift_next) {
	const int gid = lo + (blockIdx.x * BLOCK + threadIdx.x) / SPRING_LANES;
	const int lane = threadIdx.x & (SPRING_LANES - 1);
	const bool live = gid < hi;
	const int node = live ? gid : hi - 1;
	const vec4d pi = posm[node];
	const vec4d vi = vel[node];
	const int beg = row_ptr[node], end = live ? row_ptr[node + 1] : beg;
	const GlobalNodes nodes{ posm, vel };
	const Force3 f = node_spring_sum(he, pal, beg, end, lane, pi, vi, nodes);
	if (!live || lane != 0) return;
	// a node without springs (e.g. a point-mass perturber, possibly massless) is left untouched
	const double inv_m = (end > beg) ? fast_rcp(pi.w) : 0.0;
	vec4d a;
	if (ZERO_FIRST) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
	else a = acc[node];
	a.x = fma(f.x, inv_m, a.x);
	a.y = fma(f.y, inv_m, a.y);
	a.z = fma(f.z, inv_m, a.z);
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
		}
		vel_out[node] = v;
		posm_out[node] = p;
	}
}

}  // namespace

// a (+)= springs for nodes [lo, hi)
int ass_launch_spring_forces_range(ass_sim* s, bool zero_first, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	const int groups_per_block = BLOCK / SPRING_LANES;
	const int grid = (hi - lo + groups_per_block - 1) / groups_per_block;
	ass_tic(s, ASS_K_SPRINGS);
	if (zero_first)
		k_spring_forces<true, false><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->pal,lo, hi, nullptr, nullptr, 0.0, 0);
	else
		k_spring_forces<false, false><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->pal,lo, hi, nullptr, nullptr, 0.0, 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	return ASS_OK;
}

int ass_launch_spring_forces(ass_sim* s, bool zero_first) { return ass_launch_spring_forces_range(s, zero_first, 0, s->n); }

int ass_launch_spring_kick_drift(ass_sim* s, bool zero_first, double dt, bool pre_drift_next) {
	if (s->n == 0) return ASS_OK;
	const int groups_per_block = BLOCK / SPRING_LANES;
	const int grid = (s->n + groups_per_block - 1) / groups_per_block;
	ass_tic(s, ASS_K_SPRINGS);
	if (zero_first)
		k_spring_forces<true, true><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->pal,0, s->n, s->posm_alt, s->vel_alt, dt, pre_drift_next ? 1 : 0);
	else
		k_spring_forces<false, true><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->pal,0, s->n, s->posm_alt, s->vel_alt, dt, pre_drift_next ? 1 : 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	std::swap(s->posm, s->posm_alt);
	std::swap(s->vel, s->vel_alt);
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
	return ass_launch_spring_forces(s, false);
}
