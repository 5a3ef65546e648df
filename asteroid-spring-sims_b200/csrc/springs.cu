// springs.cu -- spring_forces (reference src_spring/springs.cpp:288-346) as a node-sorted CSR segmented reduction.
//
// The reference walks the spring list once and scatter-adds +F/m_i, -F/m_j into both endpoints.  Here every node owns
// the list of its incident half-edges (built by ass_set_springs) and a group of LANES threads sums them: no atomics,
// fixed summation order, so the result is deterministic run to run.  Each lane streams 32-byte HalfEdge records
// (coalesced: LANES consecutive sectors) and gathers the other endpoint's posm / vel sectors.
//
// Arithmetic kept bit-identical to the reference where cancellation would amplify a difference:
//   len   = sqrt(fma(dz,dz, fma(dx,dx, dy*dy))) + L_EPS       (Vector::len as gcc compiles it)
//   c     = (-k) * (len - rs0)                                 (springs.cpp:312)
// (at rest len - rs0 = L_EPS +- 1 ulp(len): one ulp of difference in len would already be 1e-10 of the force).
// The remaining factors (l^ = dx/len, /m, the damping term) are evaluated with reciprocals: a few ulp per term.
//
// Algorithmic traffic: 32 B x NS + 80 B x N per evaluation.  Real traffic of this layout: 64 B x NS of
// half-edge stream (each spring is stored at both endpoints) + 64 B gathered per half-edge (L2-resident) + 96 B x N.
#include "common.cuh"

namespace {

constexpr int LANES = 8;      // threads per node
constexpr int BLOCK = 256;    // 32 nodes per CTA

struct Force3 {
	double x, y, z;
};

// contribution of one half-edge to its owner node: returns s * (x_self - x_nbr) with
// s = [ -k (len - rs0) - gamma m_red (dv . l^)/len ] / (len * m_self)
__device__ __forceinline__ Force3 half_edge_force(const HalfEdge h, const vec4d pi, const vec4d vi,
		const vec4d* __restrict__ posm, const vec4d* __restrict__ vel) {
	const vec4d pj = posm[h.nbr];
	const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
	const double len = __dadd_rn(ref_len(dx, dy, dz), ASS_L_EPS);
	const double inv_len = fast_rcp(len);
	double c = __dmul_rn(-h.k, __dsub_rn(len, h.rs0));
	if (h.gamma > 0.0) {
		const vec4d vj = vel[h.nbr];
		const double dvx = vi.x - vj.x, dvy = vi.y - vj.y, dvz = vi.z - vj.z;
		// strain_rate = dot(dv, l^)/len = dot(dv, dx)/len^2 ; m_red = m_i m_j / (m_i + m_j)
		const double sr = fma(dvz, dz, fma(dvx, dx, dvy * dy)) * inv_len * inv_len;
		const double m_red = pi.w * pj.w * fast_rcp(pi.w + pj.w);
		c = fma(-h.gamma * m_red, sr, c);
	}
	const double s = c * inv_len;
	Force3 f;
	f.x = s * dx; f.y = s * dy; f.z = s * dz;
	return f;
}

// MODE 0: acc += springs            (spring_forces)
// MODE 1: acc  = springs            (zero_accel + spring_forces)
// KICK : additionally v += dt*a ; x += (dt/2)*v [; x += (dt/2)*v again = next step's part1] into the alt buffers
template <bool ZERO_FIRST, bool KICK>
__global__ void __launch_bounds__(BLOCK) k_spring_forces(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel,
		vec4d* __restrict__ acc, const int* __restrict__ row_ptr, const HalfEdge* __restrict__ he, int n,
		vec4d* __restrict__ posm_out, vec4d* __restrict__ vel_out, double dt, int pre_drift_next) {
	const int gid = (blockIdx.x * BLOCK + threadIdx.x) / LANES;
	const int lane = threadIdx.x & (LANES - 1);
	const bool live = gid < n;
	const int node = live ? gid : n - 1;
	const vec4d pi = posm[node];
	const vec4d vi = vel[node];
	const int beg = row_ptr[node], end = live ? row_ptr[node + 1] : beg;
	double ax = 0.0, ay = 0.0, az = 0.0;
	int e = beg + lane;
	// two half-edges in flight per lane
	for (; e + LANES < end; e += 2 * LANES) {
		const HalfEdge h0 = he[e], h1 = he[e + LANES];
		const Force3 f0 = half_edge_force(h0, pi, vi, posm, vel);
		const Force3 f1 = half_edge_force(h1, pi, vi, posm, vel);
		ax += f0.x; ay += f0.y; az += f0.z;
		ax += f1.x; ay += f1.y; az += f1.z;
	}
	if (e < end) {
		const Force3 f0 = half_edge_force(he[e], pi, vi, posm, vel);
		ax += f0.x; ay += f0.y; az += f0.z;
	}
#pragma unroll
	for (int m = LANES / 2; m > 0; m >>= 1) {
		ax += shfl_xor_d(ax, m, LANES);
		ay += shfl_xor_d(ay, m, LANES);
		az += shfl_xor_d(az, m, LANES);
	}
	if (!live || lane != 0) return;
	// a node without springs (e.g. a point-mass perturber, possibly massless) is left untouched
	const double inv_m = (end > beg) ? fast_rcp(pi.w) : 0.0;
	vec4d a;
	if (ZERO_FIRST) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
	else a = acc[node];
	a.x = fma(ax, inv_m, a.x);
	a.y = fma(ay, inv_m, a.y);
	a.z = fma(az, inv_m, a.z);
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

int ass_launch_spring_forces(ass_sim* s, bool zero_first) {
	if (s->n == 0) return ASS_OK;
	const int groups_per_block = BLOCK / LANES;
	const int grid = (s->n + groups_per_block - 1) / groups_per_block;
	ass_tic(s, ASS_K_SPRINGS);
	if (zero_first)
		k_spring_forces<true, false><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->n, nullptr, nullptr, 0.0, 0);
	else
		k_spring_forces<false, false><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->n, nullptr, nullptr, 0.0, 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_SPRINGS);
	return ASS_OK;
}

int ass_launch_spring_kick_drift(ass_sim* s, bool zero_first, double dt, bool pre_drift_next) {
	if (s->n == 0) return ASS_OK;
	const int groups_per_block = BLOCK / LANES;
	const int grid = (s->n + groups_per_block - 1) / groups_per_block;
	ass_tic(s, ASS_K_SPRINGS);
	if (zero_first)
		k_spring_forces<true, true><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->n, s->posm_alt, s->vel_alt, dt, pre_drift_next ? 1 : 0);
	else
		k_spring_forces<false, true><<<grid, BLOCK, 0, s->stream>>>(s->posm, s->vel, s->acc, s->row_ptr, s->he, s->n, s->posm_alt, s->vel_alt, dt, pre_drift_next ? 1 : 0);
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
	return ass_launch_spring_forces(s, false);
}
