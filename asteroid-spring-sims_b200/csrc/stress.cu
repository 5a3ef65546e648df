// stress.cu -- the "next" rows of SURVEY 8f that sit beside spring_forces:
//   * the per-node stress tensor of update_stress (reference src_spring/stress.cpp:32-85): the same gather over a node's
//     springs as spring_forces, summing +-outer(F, L) instead of F;
//   * the masked per-step force terms that live in two modules' own additional_forces: apply_impact
//     (modules/bennu2/problem.cpp:366-460) and top_push / table_top (modules/cube/problem.cpp:205-264).
// None of it is on the timed step path; it exists so that a module running in resident mode never has to bring the body
// back to the host for these terms.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace {

// he_spring[pos] = 2 * spring + side for the half-edge stored at `pos` of the CSR copy (the inverse of ass_sim::slot)
__global__ void k_invert_slots(const int* __restrict__ slot, int n_he, int* __restrict__ he_spring) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q < n_he) he_spring[slot[q]] = q;
}

// One thread per node (by rank).  The reference adds a node's contributions in SPRING-INDEX order (stress.cpp:49-74
// walks the list once); rows here are in the spring kernels' bank-aware order, so the thread picks its row's entries by
// ascending spring index (rows hold ~30 entries: a selection scan is cheap at diagnostic cadence).  Every operation
// is the reference's, rounded once each: spring_i_force (springs.cpp:350-392: true divisions by len, no L_EPS), outer
// (matrix_math.cpp), += / -= per element, / vol at the end.  Result: the reference's bits.
__global__ void k_stress(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const int* __restrict__ order, const int* __restrict__ row_ptr,
		const HalfEdge* __restrict__ he, const int* __restrict__ he_spring, const double2* __restrict__ pal, int n, double vol,
		double* __restrict__ tensors, double* __restrict__ max_force, int* __restrict__ max_spring) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const int node = order[k];
	const int beg = row_ptr[k], end = row_ptr[k + 1];
	const vec4d ps = posm[node], vs = vel[node];
	double T[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
	double best_f = 0.0;
	int best_s = -1, last = -1;
	for (int cnt = beg; cnt < end; cnt++) {
		int pick = -1, id = 0x7fffffff;
		for (int e = beg; e < end; e++) {
			const int s = he_spring[e] >> 1;
			if (s > last && s < id) { id = s; pick = e; }
		}
		if (pick < 0) break;
		last = id;
		const int side = he_spring[pick] & 1;  // 0: this node is particle_1, 1: it is particle_2
		const HalfEdge h = he[pick];
		const int other = order[h.nbr];
		const vec4d po = posm[other], vo = vel[other];
		const vec4d pi = side ? po : ps, pj = side ? ps : po, vi = side ? vo : vs, vj = side ? vs : vo;
		const double2 kg = pal[h.pal];
		// spring_r = x_i - x_j (springs.cpp:220-232), len without L_EPS (:360)
		const double Lx = __dsub_rn(pi.x, pj.x), Ly = __dsub_rn(pi.y, pj.y), Lz = __dsub_rn(pi.z, pj.z);
		const double len = ref_len(Lx, Ly, Lz);
		const double hx = __ddiv_rn(Lx, len), hy = __ddiv_rn(Ly, len), hz = __ddiv_rn(Lz, len);
		const double c = __dmul_rn(-kg.x, __dsub_rn(len, h.rs0));  // -k * (len - len0)
		double Fx = __dmul_rn(hx, c), Fy = __dmul_rn(hy, c), Fz = __dmul_rn(hz, c);
		if (kg.y > 0.0) {  // :375-391
			const double dvx = __dsub_rn(vi.x, vj.x), dvy = __dsub_rn(vi.y, vj.y), dvz = __dsub_rn(vi.z, vj.z);
			const double strain_rate = __ddiv_rn(ref_dot(dvx, dvy, dvz, hx, hy, hz), len);
			const double m_red = __ddiv_rn(__dmul_rn(pi.w, pj.w), __dadd_rn(pi.w, pj.w));
			const double g = __dmul_rn(__dmul_rn(kg.y, m_red), strain_rate);
			Fx = __dsub_rn(Fx, __dmul_rn(hx, g)); Fy = __dsub_rn(Fy, __dmul_rn(hy, g)); Fz = __dsub_rn(Fz, __dmul_rn(hz, g));
		}
		const double F[3] = { Fx, Fy, Fz }, L[3] = { Lx, Ly, Lz };
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int b = 0; b < 3; b++) {
				const double o = __dmul_rn(F[a], L[b]);
				T[3 * a + b] = side ? __dsub_rn(T[3 * a + b], o) : __dadd_rn(T[3 * a + b], o);  // :61-62
			}
		const double F_mag = ref_len(Fx, Fy, Fz);
		if (F_mag > best_f) { best_f = F_mag; best_s = id; }  // :66-73, strict >: the first spring of the largest force
	}
#pragma unroll
	for (int q = 0; q < 9; q++) tensors[(size_t) node * 9 + q] = __ddiv_rn(T[q], vol);  // :81-84
	max_force[node] = best_f;
	max_spring[node] = best_s;
}

// ---- node masks ----
__global__ void k_mask_plane(const vec4d* __restrict__ posm, int lo, int hi, int axis, double thr, int greater, unsigned char* __restrict__ mask,
		int* __restrict__ count) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const vec4d p = posm[i];
	const double c = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
	const bool in = greater ? (c > thr) : (c < thr);
	mask[i] = in ? 1 : 0;
	if (in) atomicAdd(count, 1);
}

// apply_impact's first pass (modules/bennu2/problem.cpp:391-410): masked nodes within `dangle` of the pulse direction
__device__ __forceinline__ bool within_pulse(const vec4d p, double x0, double y0, double z0, double dangle, double* hx, double* hy, double* hz, double* len) {
	*len = ref_len(p.x, p.y, p.z);
	*hx = __ddiv_rn(p.x, *len); *hy = __ddiv_rn(p.y, *len); *hz = __ddiv_rn(p.z, *len);
	const double projection = ref_dot(x0, y0, z0, *hx, *hy, *hz);
	return acos(projection) < dangle;
}
__global__ void k_impulse_count(const vec4d* __restrict__ posm, const unsigned char* __restrict__ mask, int lo, int hi, double x0, double y0, double z0,
		double dangle, int* __restrict__ count) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi || !mask[i]) return;
	double hx, hy, hz, len;
	if (within_pulse(posm[i], x0, y0, z0, dangle, &hx, &hy, &hz, &len)) atomicAdd(count, 1);
}
// second pass (:415-447): a -= force * x_hat / (m * num_pushed * |x|) * envelope
__global__ void k_impulse_apply(const vec4d* __restrict__ posm, vec4d* __restrict__ acc, const unsigned char* __restrict__ mask, int lo, int hi, double x0,
		double y0, double z0, double dangle, double force, double envelope, const int* __restrict__ count) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi || !mask[i]) return;
	const vec4d p = posm[i];
	double hx, hy, hz, len;
	if (!within_pulse(p, x0, y0, z0, dangle, &hx, &hy, &hz, &len)) return;
	const double den = __dmul_rn(__dmul_rn(p.w, (double) *count), len);  // m * num_pushed * x.len()
	vec4d a = acc[i];
	a.x = __dsub_rn(a.x, __dmul_rn(__ddiv_rn(__dmul_rn(hx, force), den), envelope));
	a.y = __dsub_rn(a.y, __dmul_rn(__ddiv_rn(__dmul_rn(hy, force), den), envelope));
	a.z = __dsub_rn(a.z, __dmul_rn(__ddiv_rn(__dmul_rn(hz, force), den), envelope));
	acc[i] = a;
}

// top_push (modules/cube/problem.cpp:230-234): a[axis] -= per_node / m;  table_top (:259-263): a[axis] = 0
__global__ void k_plane_force(vec4d* __restrict__ acc, const vec4d* __restrict__ posm, const unsigned char* __restrict__ mask, int n, int axis, double per_node,
		int clamp) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !mask[i]) return;
	vec4d a = acc[i];
	double* c = axis == 0 ? &a.x : (axis == 1 ? &a.y : &a.z);
	*c = clamp ? 0.0 : __dsub_rn(*c, __ddiv_rn(per_node, posm[i].w));
	acc[i] = a;
}

}  // namespace

#define ENTER_S(s)                                                                    \
	REQUIRE(s, ASS_EINVAL, "null sim");                                               \
	REQUIRE((s)->have_state, ASS_ESTATE, "no particle state on the device");          \
	REQUIRE(!(s)->predrifted, ASS_ESTATE, "internal: positions are mid-step");        \
	{                                                                                 \
		int _rc = ass_ensure_device();                                                \
		if (_rc) return _rc;                                                          \
	}

extern "C" int ass_stress(ass_sim* s, double* tensors, double* max_force, int* max_force_spring) {
	ENTER_S(s);
	REQUIRE(tensors && max_force && max_force_spring, ASS_EINVAL, "null output");
	const int n = s->n;
	if (n == 0) return ASS_OK;
	const size_t need = sizeof(double) * 10 * (size_t) n + sizeof(int) * (size_t) n + sizeof(int) * (size_t) std::max(s->n_he, 1);
	int rc = ass_buf_reserve(s->stage, need);
	if (rc) return rc;
	double* d_t = (double*) s->stage.p;
	double* d_f = d_t + 9 * (size_t) n;
	int* d_s = (int*) (d_f + n);
	int* d_inv = d_s + n;
	const double vol = 4.0 * M_PI / (3.0 * (double) n);  // stress.cpp:80
	ass_tic(s, ASS_K_REDUCE);
	if (s->ns > 0 && s->row_ptr) {
		k_invert_slots<<<(s->n_he + 255) / 256, 256, 0, s->stream>>>(s->slot, s->n_he, d_inv);
		LAUNCH_CHECK();
		k_stress<<<(n + 127) / 128, 128, 0, s->stream>>>(s->posm, s->vel, s->order, s->row_ptr, s->he, d_inv, s->pal, n, vol, d_t, d_f, d_s);
		LAUNCH_CHECK();
	} else {
		CU_TRY(cudaMemsetAsync(d_t, 0, sizeof(double) * 10 * (size_t) n, s->stream));
		CU_TRY(cudaMemsetAsync(d_s, 0xff, sizeof(int) * (size_t) n, s->stream));  // -1
	}
	ass_toc(s, ASS_K_REDUCE);
	CU_TRY(cudaMemcpyAsync(tensors, d_t, sizeof(double) * 9 * (size_t) n, cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaMemcpyAsync(max_force, d_f, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaMemcpyAsync(max_force_spring, d_s, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}

static int ensure_masks(ass_sim* s) {
	if (s->mask_cap >= (size_t) s->n && s->masks) return ASS_OK;
	cudaFree(s->masks);
	s->masks = nullptr; s->mask_cap = 0;
	const size_t cap = (size_t) s->n + s->n / 8 + 64;
	if (cudaMalloc(&s->masks, ASS_MAX_MASKS * cap) != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("out of device memory for node masks");
		return ASS_ENOMEM;
	}
	CU_TRY(cudaMemsetAsync(s->masks, 0, ASS_MAX_MASKS * cap, s->stream));
	s->mask_cap = cap;
	return ASS_OK;
}
static unsigned char* mask_ptr(ass_sim* s, int id) { return s->masks + (size_t) id * s->mask_cap; }

extern "C" int ass_set_node_mask(ass_sim* s, int mask_id, const unsigned char* mask) {
	ENTER_S(s);
	REQUIRE(mask_id >= 0 && mask_id < ASS_MAX_MASKS, ASS_EINVAL, "mask id out of range");
	REQUIRE(mask || s->n == 0, ASS_EINVAL, "null mask");
	int rc = ensure_masks(s);
	if (rc) return rc;
	if (s->n) {
		CU_TRY(cudaMemcpyAsync(mask_ptr(s, mask_id), mask, (size_t) s->n, cudaMemcpyHostToDevice, s->stream));
		CU_TRY(cudaStreamSynchronize(s->stream));
	}
	return ASS_OK;
}

extern "C" int ass_mask_from_plane(ass_sim* s, int mask_id, int i_low, int i_high, int axis, double threshold, int greater, int* count) {
	ENTER_S(s);
	REQUIRE(mask_id >= 0 && mask_id < ASS_MAX_MASKS, ASS_EINVAL, "mask id out of range");
	REQUIRE(axis >= 0 && axis < 3, ASS_EINVAL, "axis must be 0, 1 or 2");
	REQUIRE(i_low >= 0 && i_high <= s->n && i_low <= i_high, ASS_EINVAL, "bad particle range");
	int rc = ensure_masks(s);
	if (rc) return rc;
	int* d_count = (int*) (s->d_red + 56);
	CU_TRY(cudaMemsetAsync(d_count, 0, sizeof(int), s->stream));
	CU_TRY(cudaMemsetAsync(mask_ptr(s, mask_id), 0, (size_t) s->n, s->stream));
	if (i_high > i_low) {
		k_mask_plane<<<(i_high - i_low + 255) / 256, 256, 0, s->stream>>>(s->posm, i_low, i_high, axis, threshold, greater ? 1 : 0, mask_ptr(s, mask_id), d_count);
		LAUNCH_CHECK();
	}
	int c = 0;
	CU_TRY(cudaMemcpyAsync(&c, d_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	if (count) *count = c;
	return ASS_OK;
}

extern "C" int ass_apply_radial_impulse(ass_sim* s, int mask_id, int i_low, int i_high, double lat, double longi, double dangle, double force, double envelope,
		int* num_pushed) {
	ENTER_S(s);
	REQUIRE(mask_id >= 0 && mask_id < ASS_MAX_MASKS && s->masks, ASS_EINVAL, "mask not set");
	REQUIRE(i_low >= 0 && i_high <= s->n && i_low <= i_high, ASS_EINVAL, "bad particle range");
	// direction of the pulse centre (modules/bennu2/problem.cpp:386), evaluated on the host as the module does
	const double x0 = cos(longi) * cos(lat), y0 = sin(longi) * cos(lat), z0 = sin(lat);
	int* d_count = (int*) (s->d_red + 56);
	CU_TRY(cudaMemsetAsync(d_count, 0, sizeof(int), s->stream));
	if (i_high > i_low) {
		const int grid = (i_high - i_low + 255) / 256;
		ass_tic(s, ASS_K_REDUCE);
		k_impulse_count<<<grid, 256, 0, s->stream>>>(s->posm, mask_ptr(s, mask_id), i_low, i_high, x0, y0, z0, dangle, d_count);
		LAUNCH_CHECK();
		k_impulse_apply<<<grid, 256, 0, s->stream>>>(s->posm, s->acc, mask_ptr(s, mask_id), i_low, i_high, x0, y0, z0, dangle, force, envelope, d_count);
		LAUNCH_CHECK();
		ass_toc(s, ASS_K_REDUCE);
	}
	if (num_pushed) {  // only a caller that wants the count (the module prints it) pays for the round trip
		CU_TRY(cudaMemcpyAsync(num_pushed, d_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CU_TRY(cudaStreamSynchronize(s->stream));
	}
	return ASS_OK;
}

extern "C" int ass_plane_push(ass_sim* s, int mask_id, int axis, double per_node_force) {
	ENTER_S(s);
	REQUIRE(mask_id >= 0 && mask_id < ASS_MAX_MASKS && s->masks, ASS_EINVAL, "mask not set");
	REQUIRE(axis >= 0 && axis < 3, ASS_EINVAL, "axis must be 0, 1 or 2");
	if (s->n == 0) return ASS_OK;
	ass_tic(s, ASS_K_REDUCE);
	k_plane_force<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->acc, s->posm, mask_ptr(s, mask_id), s->n, axis, per_node_force, 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

extern "C" int ass_plane_constraint(ass_sim* s, int mask_id, int axis) {
	ENTER_S(s);
	REQUIRE(mask_id >= 0 && mask_id < ASS_MAX_MASKS && s->masks, ASS_EINVAL, "mask not set");
	REQUIRE(axis >= 0 && axis < 3, ASS_EINVAL, "axis must be 0, 1 or 2");
	if (s->n == 0) return ASS_OK;
	ass_tic(s, ASS_K_REDUCE);
	k_plane_force<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->acc, s->posm, mask_ptr(s, mask_id), s->n, axis, 0.0, 1);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}
