// orbit.cu -- the remaining per-step terms of the phobos / phobos_deimos modules (reference src_spring/orb.cpp), so
// that an unmodified module of that family stays resident on the device through its whole step:
//   quadrupole_accel (orb.cpp:311-366)   called from additional_forces after spring_forces, O(N)
//   drift_resolved   (orb.cpp:253-302)   called from the heartbeat: migration of the body about the central mass
//   drift_bin        (orb.cpp:198-250)   the same between two point masses
//   sum_mass         (orb.cpp:438-449)
// All HBM-bound and tiny next to gravity; what matters is that nothing here needs the host arrays.  Sums that the
// reference accumulates in particle order (the reaction on the quadrupole's owner) are fixed-order trees here.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int QB = 256;

// a_i += ((constant * r) / r5 * (5 cos^2 - diff)) / m_i, component-wise exactly as Vector's operators evaluate it
// (orb.cpp:347); block partials of the reaction sum a_i * (m_i / m_p) for the owner.
__global__ void __launch_bounds__(QB) k_quadrupole(const vec4d* __restrict__ posm, vec4d* __restrict__ acc, int n, int i_p, double constant,
		double hx, double hy, double hz, double* __restrict__ part) {
	__shared__ double red[3][QB / 32];
	const int i = blockIdx.x * QB + threadIdx.x;
	double rx = 0.0, ry = 0.0, rz = 0.0;
	if (i < n && i != i_p) {
		const vec4d pi = posm[i], pp = posm[i_p];
		const double x = pi.x - pp.x, y = pi.y - pp.y, z = pi.z - pp.z;
		const double len = ref_len(x, y, z);
		const double l2 = len * len;
		const double r5 = l2 * l2 * len + 0.5;  // "arbitrary softening of r^5", orb.cpp:340
		const double c = ref_dot(x, y, z, hx, hy, hz) / len;
		const double cos2 = c * c;
		const double ax = (constant * x) / r5 * (-(1.0 - 5.0 * cos2)) / pi.w;
		const double ay = (constant * y) / r5 * (-(1.0 - 5.0 * cos2)) / pi.w;
		const double az = (constant * z) / r5 * (-(3.0 - 5.0 * cos2)) / pi.w;
		vec4d a = acc[i];
		a.x += ax; a.y += ay; a.z += az;
		acc[i] = a;
		const double mfac = pi.w / pp.w;
		rx = ax * mfac; ry = ay * mfac; rz = az * mfac;
	}
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		rx += shfl_down_d(rx, m); ry += shfl_down_d(ry, m); rz += shfl_down_d(rz, m);
	}
	if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = rx; red[1][threadIdx.x >> 5] = ry; red[2][threadIdx.x >> 5] = rz; }
	__syncthreads();
	if (threadIdx.x < 3) {
		double t = 0.0;
		for (int w = 0; w < QB / 32; w++) t += red[threadIdx.x][w];
		part[(size_t) blockIdx.x * 3 + threadIdx.x] = t;
	}
}

__global__ void k_quadrupole_owner(const double* __restrict__ part, int nblk, vec4d* __restrict__ acc, int i_p) {
	__shared__ double tot[3];
	if (threadIdx.x < 3) {
		double t = 0.0;
		for (int b = 0; b < nblk; b++) t += part[(size_t) b * 3 + threadIdx.x];
		tot[threadIdx.x] = t;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		vec4d a = acc[i_p];
		a.x -= tot[0]; a.y -= tot[1]; a.z -= tot[2];
		acc[i_p] = a;
	}
}

// dv = timestep (v inv_tau_a / 2 + (v - v_cent r x (r x v)/|.|) inv_tau_e), orb.cpp:218-234 / 271-287
__device__ void drift_dv(const double x[3], const double v[3], double GM, double timestep, double ita, double ite, double dv[3]) {
	const double rad = ref_len(x[0], x[1], x[2]);
	const double v_cent = sqrt(GM / rad);
	// cross(), matrix_math.cpp:169-174 as gcc contracts it: fma(a, b, -(c d))
	double xv[3], rl[3];
	xv[0] = __fma_rn(x[1], v[2], -__dmul_rn(x[2], v[1]));
	xv[1] = __fma_rn(x[2], v[0], -__dmul_rn(x[0], v[2]));
	xv[2] = __fma_rn(x[0], v[1], -__dmul_rn(x[1], v[0]));
	rl[0] = __fma_rn(x[1], xv[2], -__dmul_rn(x[2], xv[1]));
	rl[1] = __fma_rn(x[2], xv[0], -__dmul_rn(x[0], xv[2]));
	rl[2] = __fma_rn(x[0], xv[1], -__dmul_rn(x[1], xv[0]));
	const double nrm = ref_len(rl[0], rl[1], rl[2]);
	for (int c = 0; c < 3; c++) {
		const double dir = __ddiv_rn(rl[c], nrm);
		const double delta_v = __dsub_rn(v[c], __dmul_rn(v_cent, dir));
		dv[c] = __dmul_rn(timestep, __dadd_rn(__ddiv_rn(__dmul_rn(v[c], ita), 2.0), __dmul_rn(delta_v, ite)));
	}
}

// moments: [0..2] CoM, [3] mass, [4..6] CoV, [7] mass.  Updates the particle's velocity and leaves the body's shift
// (-dv_2) in shift_out[0..2] for k_shift.
__global__ void k_drift_resolved(const vec4d* __restrict__ posm, vec4d* __restrict__ vel, int i_part, const double* __restrict__ moments, double G,
		double timestep, double ita, double ite, double* __restrict__ shift_out) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	const vec4d pp = posm[i_part];
	vec4d vp = vel[i_part];
	const double m1 = pp.w, m2 = moments[3];
	const double x[3] = { moments[0] - pp.x, moments[1] - pp.y, moments[2] - pp.z };
	const double v[3] = { moments[4] - vp.x, moments[5] - vp.y, moments[6] - vp.z };
	double dv[3];
	drift_dv(x, v, __dmul_rn(G, __dadd_rn(m1, m2)), timestep, ita, ite, dv);
	const double mt = __dadd_rn(m1, m2);
	vp.x = __dsub_rn(vp.x, __ddiv_rn(__dmul_rn(m1, dv[0]), mt));
	vp.y = __dsub_rn(vp.y, __ddiv_rn(__dmul_rn(m1, dv[1]), mt));
	vp.z = __dsub_rn(vp.z, __ddiv_rn(__dmul_rn(m1, dv[2]), mt));
	vel[i_part] = vp;
	for (int c = 0; c < 3; c++) shift_out[c] = -__ddiv_rn(__dmul_rn(m2, dv[c]), mt);
}

__global__ void k_drift_bin(const vec4d* __restrict__ posm, vec4d* __restrict__ vel, int p1, int p2, double G, double timestep, double ita, double ite) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	const vec4d a = posm[p1], b = posm[p2];
	vec4d va = vel[p1], vb = vel[p2];
	const double m1 = a.w, m2 = b.w;
	const double x[3] = { b.x - a.x, b.y - a.y, b.z - a.z };
	const double v[3] = { vb.x - va.x, vb.y - va.y, vb.z - va.z };
	double dv[3];
	drift_dv(x, v, __dmul_rn(G, __dadd_rn(m1, m2)), timestep, ita, ite, dv);
	const double mt = __dadd_rn(m1, m2);
	va.x = __dsub_rn(va.x, __ddiv_rn(__dmul_rn(m1, dv[0]), mt));
	va.y = __dsub_rn(va.y, __ddiv_rn(__dmul_rn(m1, dv[1]), mt));
	va.z = __dsub_rn(va.z, __ddiv_rn(__dmul_rn(m1, dv[2]), mt));
	vb.x = __dsub_rn(vb.x, __ddiv_rn(__dmul_rn(m2, dv[0]), mt));
	vb.y = __dsub_rn(vb.y, __ddiv_rn(__dmul_rn(m2, dv[1]), mt));
	vb.z = __dsub_rn(vb.z, __ddiv_rn(__dmul_rn(m2, dv[2]), mt));
	vel[p1] = va;
	vel[p2] = vb;
}

}  // namespace

#define ENTER_ORB(s)                                                                      \
	REQUIRE(s, ASS_EINVAL, "null sim");                                                   \
	REQUIRE((s)->have_state, ASS_ESTATE, "no particle state on the device");              \
	REQUIRE(!(s)->predrifted, ASS_ESTATE, "internal: positions are mid-step");            \
	REQUIRE(!(s)->shard.active, ASS_ESTATE, "not available on a sharded simulation");     \
	{                                                                                     \
		int _rc = ass_ensure_device();                                                    \
		if (_rc) return _rc;                                                              \
	}

extern "C" int ass_quadrupole_accel(ass_sim* s, double J2_p, double R_p, double phi_p, double theta_p, int i_p) {
	ENTER_ORB(s);
	if (i_p >= s->n) return ASS_OK;   // orb.cpp:315-316
	if (J2_p == 0.0) return ASS_OK;   // orb.cpp:319-320
	REQUIRE(i_p >= 0, ASS_EINVAL, "negative particle index");
	// the owner's mass lives on the device: fold G m_c into the constant there?  No: one 8-byte read keeps the formula's
	// rounding (constant = 1.5 * (G m_c) * J2 * R^2, orb.cpp:326-327) on the host, with the host's pow / sin / cos.
	double m_c = 0.0;
	CU_TRY(cudaMemcpyAsync(&m_c, &s->posm[i_p].w, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	const double Gmc = s->prm.G * m_c;
	const double constant = 1.5 * Gmc * J2_p * pow(R_p, 2.0);
	const double hx = sin(theta_p) * cos(phi_p), hy = sin(theta_p) * sin(phi_p), hz = cos(theta_p);
	const int grid = (s->n + QB - 1) / QB;
	int rc = ass_buf_reserve(s->red_partial, sizeof(double) * 3 * (size_t) grid);
	if (rc) return rc;
	ass_tic(s, ASS_K_REDUCE);
	k_quadrupole<<<grid, QB, 0, s->stream>>>(s->posm, s->acc, s->n, i_p, constant, hx, hy, hz, (double*) s->red_partial.p);
	LAUNCH_CHECK();
	k_quadrupole_owner<<<1, 32, 0, s->stream>>>((const double*) s->red_partial.p, grid, s->acc, i_p);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

extern "C" int ass_drift_resolved(ass_sim* s, double timestep, double inv_tau_a, double inv_tau_e, int i_part, int i_low, int i_high) {
	ENTER_ORB(s);
	REQUIRE(i_low >= 0 && i_high <= s->n && i_low < i_high, ASS_EINVAL, "bad particle range");
	REQUIRE(i_part >= 0 && i_part < s->n && (i_part < i_low || i_part >= i_high), ASS_EINVAL, "the point mass must lie outside the resolved body");
	int rc;
	double* d = s->d_red;  // [0..3] CoM + mass, [4..7] CoV + mass, [8..10] shift
	if ((rc = ass_launch_com(s, i_low, i_high, false, d))) return rc;
	if ((rc = ass_launch_com(s, i_low, i_high, true, d + 4))) return rc;
	ass_tic(s, ASS_K_REDUCE);
	k_drift_resolved<<<1, 32, 0, s->stream>>>(s->posm, s->vel, i_part, d, s->prm.G, timestep, inv_tau_a, inv_tau_e, d + 8);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ass_launch_shift(s, i_low, i_high, d + 8, true, 1.0, 0.0);  // move_resolved(zero, -dv_2), orb.cpp:301
}

extern "C" int ass_drift_bin(ass_sim* s, double timestep, double inv_tau_a, double inv_tau_e, int part_1, int part_2) {
	ENTER_ORB(s);
	REQUIRE(part_1 >= 0 && part_1 < s->n && part_2 >= 0 && part_2 < s->n && part_1 != part_2, ASS_EINVAL, "bad particle indices");
	ass_tic(s, ASS_K_REDUCE);
	k_drift_bin<<<1, 32, 0, s->stream>>>(s->posm, s->vel, part_1, part_2, s->prm.G, timestep, inv_tau_a, inv_tau_e);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

extern "C" int ass_sum_mass(ass_sim* s, int i_low, int i_high, double* out) {
	ENTER_ORB(s);
	REQUIRE(out, ASS_EINVAL, "null out");
	REQUIRE(i_low >= 0 && i_high <= s->n && i_low <= i_high, ASS_EINVAL, "bad particle range");
	if (i_low == i_high) { *out = 0.0; return ASS_OK; }
	int rc = ass_launch_com(s, i_low, i_high, false, s->d_red);
	if (rc) return rc;
	CU_TRY(cudaMemcpyAsync(s->h_red, s->d_red, sizeof(double) * 4, cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	*out = s->h_red[3];
	return ASS_OK;
}

