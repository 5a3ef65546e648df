// reduce.cu -- per-step heartbeat helpers and print-cadence diagnostics (reference src_spring/physics.cpp).
//
// The reference accumulates these sums sequentially in particle order; a parallel sum cannot reproduce those bits,
// so the contract is "deterministic (fixed tree), equal to the reference to rounding" -- tests bound the difference
// at 1e-13 relative to the sum of magnitudes.  All kernels are two-stage: per-CTA partials in a fixed order, then a
// single CTA folds the partials, also in a fixed order.
#include "common.cuh"

namespace {

constexpr int RB = 256;     // threads per CTA
constexpr int MAXV = 16;    // values reduced together

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ out /* [gridDim.x][NV] */) {
	__shared__ double sm[RB / 32][MAXV];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int q = 0; q < NV; q++) {
#pragma unroll
		for (int m = 16; m > 0; m >>= 1) v[q] += shfl_down_d(v[q], m);
	}
	if (lane == 0) {
#pragma unroll
		for (int q = 0; q < NV; q++) sm[warp][q] = v[q];
	}
	__syncthreads();
	if (threadIdx.x < NV) {
		double a = 0.0;
		for (int w = 0; w < RB / 32; w++) a += sm[w][threadIdx.x];
		out[(size_t) blockIdx.x * NV + threadIdx.x] = a;
	}
}

// fold partials[nblk][NV] -> out[NV]  (one CTA)
template <int NV>
__global__ void k_fold(const double* __restrict__ part, int nblk, double* __restrict__ out) {
	double v[NV];
#pragma unroll
	for (int q = 0; q < NV; q++) v[q] = 0.0;
	for (int b = threadIdx.x; b < nblk; b += RB) {
#pragma unroll
		for (int q = 0; q < NV; q++) v[q] += part[(size_t) b * NV + q];
	}
	__shared__ double sm[RB / 32][MAXV];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int q = 0; q < NV; q++) {
#pragma unroll
		for (int m = 16; m > 0; m >>= 1) v[q] += shfl_down_d(v[q], m);
	}
	if (lane == 0) {
#pragma unroll
		for (int q = 0; q < NV; q++) sm[warp][q] = v[q];
	}
	__syncthreads();
	if (threadIdx.x < NV) {
		double a = 0.0;
		for (int w = 0; w < RB / 32; w++) a += sm[w][threadIdx.x];
		out[threadIdx.x] = a;
	}
}

// sum m*x (or m*v) and sum m over [lo,hi): compute_com / compute_cov / sum_mass  (physics.cpp:41-74, orb.cpp:438-449)
__global__ void __launch_bounds__(RB) k_moment(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, int lo, int hi, int velocity,
		double* __restrict__ part) {
	double v[4] = { 0.0, 0.0, 0.0, 0.0 };
	for (int i = lo + blockIdx.x * RB + threadIdx.x; i < hi; i += gridDim.x * RB) {
		const vec4d p = posm[i];
		if (velocity) {
			const vec4d w = vel[i];
			v[0] += p.w * w.x; v[1] += p.w * w.y; v[2] += p.w * w.z;
		} else {
			v[0] += p.w * p.x; v[1] += p.w * p.y; v[2] += p.w * p.z;
		}
		v[3] += p.w;
	}
	block_reduce_store<4>(v, part);
}

// The same moments in the REFERENCE's order: m_tot = sum_mass() then x_times_m += m * x for i ascending, each product
// and each add rounded separately (operator*(double, Vector) and operator+= are out-of-line calls in matrix_math.cpp,
// so gcc cannot fuse them), then x_times_m / m_tot.  A floating-point sum in a fixed sequential order cannot be
// parallelised without changing its bits, so one CTA stages 1024 products at a time into shared memory and four
// threads run the four dependent chains.  ~8 cycles per particle: fine at set-up (subtract_com, spin_body before
// connect_springs_dist, where one ulp of position would change rs0 bits); per-step callers can ask for the tree.
constexpr int SEQ_CHUNK = 1024;
__global__ void __launch_bounds__(256) k_moment_seq(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, int lo, int hi, int velocity,
		double* __restrict__ out) {
	__shared__ double buf[4][SEQ_CHUNK];
	double acc = 0.0;
	for (int base = lo; base < hi; base += SEQ_CHUNK) {
		const int cnt = min(SEQ_CHUNK, hi - base);
		for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
			const vec4d p = posm[base + k];
			double x = p.x, y = p.y, z = p.z;
			if (velocity) { const vec4d w = vel[base + k]; x = w.x; y = w.y; z = w.z; }
			buf[0][k] = __dmul_rn(x, p.w);
			buf[1][k] = __dmul_rn(y, p.w);
			buf[2][k] = __dmul_rn(z, p.w);
			buf[3][k] = p.w;
		}
		__syncthreads();
		if (threadIdx.x < 4) {
			const double* b = buf[threadIdx.x];
			for (int k = 0; k < cnt; k++) acc = __dadd_rn(acc, b[k]);
		}
		__syncthreads();
	}
	__shared__ double tot[4];
	if (threadIdx.x < 4) tot[threadIdx.x] = acc;
	__syncthreads();
	if (threadIdx.x < 3) out[threadIdx.x] = __ddiv_rn(tot[threadIdx.x], tot[3]);
	if (threadIdx.x == 3) out[3] = tot[3];
}

// mom_inertia (physics.cpp:116-134) summed in the reference's particle order, one rounding per operation:
//     inertia += m * pow(|dx|, 2) * I ;  inertia -= m * outer(dx, dx)        (dx = x - CoM)
// i.e. per particle  I_aa = (I_aa + m |dx|^2) - (dx_a dx_a) m  and  I_ab = I_ab - (dx_a dx_b) m.  Same staging as
// k_moment_seq: all threads prepare a chunk, six threads run the six dependent chains (xx yy zz xy xz yz).  For set-up:
// rotate_to_principal takes eigenvectors of this tensor, and their signs hang on its last bits.
__global__ void __launch_bounds__(256) k_inertia_seq(const vec4d* __restrict__ posm, int lo, int hi, const double* __restrict__ com, double* __restrict__ out6) {
	__shared__ double sub[6][SEQ_CHUNK / 2];  // what is subtracted: (dx_a dx_b) m
	__shared__ double add[SEQ_CHUNK / 2];     // what the diagonal gains first: m |dx|^2
	double acc = 0.0;
	const double cx = com[0], cy = com[1], cz = com[2];
	const int t = threadIdx.x;
	const int ia[6] = { 0, 1, 2, 0, 0, 1 }, ib[6] = { 0, 1, 2, 1, 2, 2 };
	for (int base = lo; base < hi; base += SEQ_CHUNK / 2) {
		const int cnt = min(SEQ_CHUNK / 2, hi - base);
		for (int k = t; k < cnt; k += blockDim.x) {
			const vec4d p = posm[base + k];
			const double d[3] = { __dsub_rn(p.x, cx), __dsub_rn(p.y, cy), __dsub_rn(p.z, cz) };
			const double len = ref_len(d[0], d[1], d[2]);
			add[k] = __dmul_rn(p.w, __dmul_rn(len, len));  // m * pow(len, 2.0): pow(x, 2) is the correctly rounded square
#pragma unroll
			for (int q = 0; q < 6; q++) sub[q][k] = __dmul_rn(__dmul_rn(d[ia[q]], d[ib[q]]), p.w);
		}
		__syncthreads();
		if (t < 6) {
			const double* sb = sub[t];
			if (t < 3) for (int k = 0; k < cnt; k++) acc = __dsub_rn(__dadd_rn(acc, add[k]), sb[k]);
			else for (int k = 0; k < cnt; k++) acc = __dsub_rn(acc, sb[k]);  // (+ m |dx|^2 * 0 leaves an off-diagonal sum unchanged)
		}
		__syncthreads();
	}
	if (t < 6) out6[t] = acc;
}

// out[0..2] = sums / mass ; out[3] = mass
__global__ void k_fold_com(const double* __restrict__ part, int nblk, double* __restrict__ out) {
	__shared__ double sm[4];
	if (threadIdx.x < 4) {
		double a = 0.0;
		for (int b = 0; b < nblk; b++) a += part[(size_t) b * 4 + threadIdx.x];
		sm[threadIdx.x] = a;
	}
	__syncthreads();
	if (threadIdx.x < 3) out[threadIdx.x] = sm[threadIdx.x] / sm[3];
	if (threadIdx.x == 3) out[3] = sm[3];
}

// x (or v) += sign * shift for i in [lo,hi); optionally followed by the next step's part1 drift on ALL of [lo,hi)
__global__ void k_shift(vec4d* __restrict__ posm, vec4d* __restrict__ vel, int lo, int hi, const double* __restrict__ shift, int velocity,
		double sign, double dt_drift) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const double sx = sign * shift[0], sy = sign * shift[1], sz = sign * shift[2];
	if (velocity) {
		vec4d v = vel[i];
		v.x = __dadd_rn(v.x, sx); v.y = __dadd_rn(v.y, sy); v.z = __dadd_rn(v.z, sz);
		vel[i] = v;
	} else {
		vec4d p = posm[i];
		p.x = __dadd_rn(p.x, sx); p.y = __dadd_rn(p.y, sy); p.z = __dadd_rn(p.z, sz);
		if (dt_drift != 0.0) {
			const double h = __dmul_rn(0.5, dt_drift);
			const vec4d v = vel[i];
			p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
			p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
			p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
		}
		posm[i] = p;
	}
}

// move_resolved: x += dx, v += dv (physics.cpp:387-402)
__global__ void k_move(vec4d* __restrict__ posm, vec4d* __restrict__ vel, int lo, int hi, double dx, double dy, double dz, double dvx,
		double dvy, double dvz) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d p = posm[i], v = vel[i];
	p.x = __dadd_rn(p.x, dx); p.y = __dadd_rn(p.y, dy); p.z = __dadd_rn(p.z, dz);
	v.x = __dadd_rn(v.x, dvx); v.y = __dadd_rn(v.y, dvy); v.z = __dadd_rn(v.z, dvz);
	posm[i] = p; vel[i] = v;
}

// spin_body: v = CoV + omega x (x - CoM)   (physics.cpp:184-210); com[0..2], cov[0..2] on the device
__global__ void k_spin(const vec4d* __restrict__ posm, vec4d* __restrict__ vel, int lo, int hi, const double* __restrict__ com,
		const double* __restrict__ cov, double ox, double oy, double oz) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const vec4d p = posm[i];
	const double dx = p.x - com[0], dy = p.y - com[1], dz = p.z - com[2];
	vec4d v;
	// cross() as gcc compiles it: fma(a, b, -(c*d))  (matrix_math.cpp:169-174)
	v.x = __dadd_rn(cov[0], __fma_rn(oy, dz, -__dmul_rn(oz, dy)));
	v.y = __dadd_rn(cov[1], __fma_rn(oz, dx, -__dmul_rn(ox, dz)));
	v.z = __dadd_rn(cov[2], __fma_rn(ox, dy, -__dmul_rn(oy, dx)));
	v.w = 0.0;
	vel[i] = v;
}

// x <- R (x - c) + c, v <- R (v - u) + u for a 3x3 matrix R (row-major): rotate_body (c = CoM, u = CoV of the range),
// rotate_origin and rotate_to_principal (c = u = 0), physics.cpp:249-385.  Matrix * Vector as gcc compiles
// matrix_math.cpp:442-450: res = a0*v0, then two fused multiply-adds, in index order.
__global__ void k_transform(vec4d* __restrict__ posm, vec4d* __restrict__ vel, int lo, int hi, const double* __restrict__ com,
		const double* __restrict__ cov, double r00, double r01, double r02, double r10, double r11, double r12, double r20, double r21, double r22) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const double cx = com ? com[0] : 0.0, cy = com ? com[1] : 0.0, cz = com ? com[2] : 0.0;
	const double ux = cov ? cov[0] : 0.0, uy = cov ? cov[1] : 0.0, uz = cov ? cov[2] : 0.0;
	vec4d p = posm[i], v = vel[i];
	const double x = com ? __dsub_rn(p.x, cx) : p.x, y = com ? __dsub_rn(p.y, cy) : p.y, z = com ? __dsub_rn(p.z, cz) : p.z;
	const double a = cov ? __dsub_rn(v.x, ux) : v.x, b = cov ? __dsub_rn(v.y, uy) : v.y, c = cov ? __dsub_rn(v.z, uz) : v.z;
	const double X = __fma_rn(r02, z, __fma_rn(r01, y, __dmul_rn(r00, x)));
	const double Y = __fma_rn(r12, z, __fma_rn(r11, y, __dmul_rn(r10, x)));
	const double Z = __fma_rn(r22, z, __fma_rn(r21, y, __dmul_rn(r20, x)));
	const double A = __fma_rn(r02, c, __fma_rn(r01, b, __dmul_rn(r00, a)));
	const double B = __fma_rn(r12, c, __fma_rn(r11, b, __dmul_rn(r10, a)));
	const double C = __fma_rn(r22, c, __fma_rn(r21, b, __dmul_rn(r20, a)));
	p.x = com ? __dadd_rn(X, cx) : X; p.y = com ? __dadd_rn(Y, cy) : Y; p.z = com ? __dadd_rn(Z, cz) : Z;
	v.x = cov ? __dadd_rn(A, ux) : A; v.y = cov ? __dadd_rn(B, uy) : B; v.z = cov ? __dadd_rn(C, uz) : C;
	posm[i] = p;
	vel[i] = v;
}

// node sums needing CoM and CoV: L (3), KE (1), inertia (6)  -> 10 values
__global__ void __launch_bounds__(RB) k_node_diag(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, int lo, int hi,
		const double* __restrict__ com, const double* __restrict__ cov, double* __restrict__ part) {
	double v[10];
#pragma unroll
	for (int q = 0; q < 10; q++) v[q] = 0.0;
	const double cx = com[0], cy = com[1], cz = com[2], ux = cov[0], uy = cov[1], uz = cov[2];
	for (int i = lo + blockIdx.x * RB + threadIdx.x; i < hi; i += gridDim.x * RB) {
		const vec4d p = posm[i], w = vel[i];
		const double m = p.w;
		const double dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
		const double dvx = w.x - ux, dvy = w.y - uy, dvz = w.z - uz;
		// measure_L, physics.cpp:136-162
		v[0] += m * (dy * dvz - dz * dvy);
		v[1] += m * (dz * dvx - dx * dvz);
		v[2] += m * (dx * dvy - dy * dvx);
		// compute_rot_kin, physics.cpp:508-528
		v[3] += 0.5 * m * (dvx * dvx + dvy * dvy + dvz * dvz);
		// mom_inertia, physics.cpp:116-134
		const double r2 = dx * dx + dy * dy + dz * dz;
		v[4] += m * (r2 - dx * dx);
		v[5] += m * (r2 - dy * dy);
		v[6] += m * (r2 - dz * dz);
		v[7] -= m * dx * dy;
		v[8] -= m * dx * dz;
		v[9] -= m * dy * dz;
	}
	block_reduce_store<10>(v, part);
}

// spring sums: SPE (physics.cpp:457-476) and damping power (physics.cpp:409-454).  Each spring appears as two
// half-edges; take the one whose neighbour ranks above its owner so every spring counts once.  Rows and HalfEdge::nbr
// are in spatial rank order (common.cuh); `order` maps a rank back to the node.
__global__ void __launch_bounds__(RB) k_spring_diag(const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const int* __restrict__ order,
		const int* __restrict__ row_ptr, const HalfEdge* __restrict__ he, const double2* __restrict__ pal, int n, double* __restrict__ part) {
	double v[2] = { 0.0, 0.0 };
	for (int k = blockIdx.x * RB + threadIdx.x; k < n; k += gridDim.x * RB) {
		const int i = order[k];
		const vec4d pi = posm[i], vi = vel[i];
		for (int e = row_ptr[k]; e < row_ptr[k + 1]; e++) {
			const HalfEdge h = he[e];
			const int kn = h.nbr;
			if (kn < k) continue;
			const int j = order[kn];
			const double2 kg = pal[h.pal];
			const vec4d pj = posm[j];
			const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
			const double len = __dadd_rn(ref_len(dx, dy, dz), ASS_L_EPS);
			const double ext = len - h.rs0;
			v[0] += kg.x * (ext * ext) / 2.0;
			if (kg.y != 0.0) {
				const vec4d vj = vel[j];
				const double sr = ((vi.x - vj.x) * dx + (vi.y - vj.y) * dy + (vi.z - vj.z) * dz) / (len * len);
				const double m_red = pi.w * pj.w / (pi.w + pj.w);
				v[1] += kg.y * m_red * (sr * sr);
			}
		}
	}
	block_reduce_store<2>(v, part);
}

// grav_potential_energy (physics.cpp:479-504): -G sum_{i<j} m_i m_j / |x_i - x_j|, unsoftened.  Tiled like the force
// kernel; every ordered pair (i != j) is visited and the total halved.
constexpr int PE_T = 128;
__global__ void __launch_bounds__(PE_T) k_gpe(const vec4d* __restrict__ posm, int lo, int hi, double* __restrict__ part) {
	__shared__ vec4d tile[PE_T];
	const int i = lo + blockIdx.x * PE_T + threadIdx.x;
	const bool live = i < hi;
	const vec4d pi = posm[live ? i : hi - 1];
	double e = 0.0;
	for (int jb = lo; jb < hi; jb += PE_T) {
		__syncthreads();
		const int j = jb + threadIdx.x;
		vec4d t;
		if (j < hi) t = posm[j];
		else { t.x = 0; t.y = 0; t.z = 0; t.w = 0; }
		tile[threadIdx.x] = t;
		__syncthreads();
		const int cnt = min(PE_T, hi - jb);
		for (int jj = 0; jj < cnt; jj++) {
			const vec4d pj = tile[jj];
			const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
			const double r2 = ref_len2(dx, dy, dz);
			const double inv = (jb + jj == i) ? 0.0 : rsqrt(r2);
			e = fma(pj.w, inv, e);
		}
	}
	double v[1] = { live ? pi.w * e : 0.0 };
	// block_reduce_store assumes RB threads; do it by hand for PE_T
	__shared__ double sm[PE_T / 32];
	for (int m = 16; m > 0; m >>= 1) v[0] += shfl_down_d(v[0], m);
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v[0];
	__syncthreads();
	if (threadIdx.x == 0) {
		double a = 0.0;
		for (int w = 0; w < PE_T / 32; w++) a += sm[w];
		part[blockIdx.x] = a;
	}
}

int reduce_grid(int count) {
	const int want = (count + RB - 1) / RB;
	return std::max(1, std::min(want, ass_sm_count() * 4));
}

}  // namespace

// d_out[0..2] = CoM (or CoV), d_out[3] = mass; stays on the device
int ass_launch_com(ass_sim* s, int lo, int hi, bool velocity, double* d_out) {
	if (s->sum_sequential) {
		ass_tic(s, ASS_K_REDUCE);
		k_moment_seq<<<1, 256, 0, s->stream>>>(s->posm, s->vel, lo, hi, velocity ? 1 : 0, d_out);
		LAUNCH_CHECK();
		ass_toc(s, ASS_K_REDUCE);
		return ASS_OK;
	}
	const int grid = reduce_grid(hi - lo);
	int rc = ass_buf_reserve(s->red_partial, sizeof(double) * MAXV * (size_t) std::max(grid, ass_sm_count() * 4));
	if (rc) return rc;
	ass_tic(s, ASS_K_REDUCE);
	k_moment<<<grid, RB, 0, s->stream>>>(s->posm, s->vel, lo, hi, velocity ? 1 : 0, (double*) s->red_partial.p);
	LAUNCH_CHECK();
	k_fold_com<<<1, 32, 0, s->stream>>>((const double*) s->red_partial.p, grid, d_out);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

int ass_launch_shift(ass_sim* s, int lo, int hi, const double* d_shift3, bool velocity, double sign, double drift_dt) {
	if (hi <= lo) return ASS_OK;
	ass_tic(s, ASS_K_REDUCE);
	k_shift<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, lo, hi, d_shift3, velocity ? 1 : 0, sign, drift_dt);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

#define ENTER_RANGE(s, lo, hi)                                                            \
	REQUIRE(s, ASS_EINVAL, "null sim");                                                   \
	REQUIRE((s)->have_state, ASS_ESTATE, "no particle state on the device");              \
	REQUIRE(!(s)->predrifted, ASS_ESTATE, "internal: positions are mid-step");            \
	REQUIRE((lo) >= 0 && (hi) <= (s)->n && (lo) <= (hi), ASS_EINVAL, "bad particle range"); \
	{                                                                                     \
		int _rc = ass_ensure_device();                                                    \
		if (_rc) return _rc;                                                              \
	}

static int fetch(ass_sim* s, int count) {
	CU_TRY(cudaMemcpyAsync(s->h_red, s->d_red, sizeof(double) * count, cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}

extern "C" int ass_set_sum_order(ass_sim* s, int sequential) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	s->sum_sequential = sequential != 0;
	s->sum_order_forced = true;
	return ASS_OK;
}

extern "C" int ass_compute_com(ass_sim* s, int lo, int hi, double out[3]) {
	ENTER_RANGE(s, lo, hi);
	REQUIRE(out, ASS_EINVAL, "null out");
	int rc = ass_launch_com(s, lo, hi, false, s->d_red);
	if (rc || (rc = fetch(s, 4))) return rc;
	out[0] = s->h_red[0]; out[1] = s->h_red[1]; out[2] = s->h_red[2];
	return ASS_OK;
}
extern "C" int ass_compute_cov(ass_sim* s, int lo, int hi, double out[3]) {
	ENTER_RANGE(s, lo, hi);
	REQUIRE(out, ASS_EINVAL, "null out");
	int rc = ass_launch_com(s, lo, hi, true, s->d_red);
	if (rc || (rc = fetch(s, 4))) return rc;
	out[0] = s->h_red[0]; out[1] = s->h_red[1]; out[2] = s->h_red[2];
	return ASS_OK;
}
extern "C" int ass_center_sim(ass_sim* s, int lo, int hi) {  // physics.cpp:98-109: CoM of [lo,hi), shift ALL particles
	ENTER_RANGE(s, lo, hi);
	int rc = ass_launch_com(s, lo, hi, false, s->d_red);
	if (rc) return rc;
	return ass_launch_shift(s, 0, s->n, s->d_red, false, -1.0, 0.0);
}
extern "C" int ass_subtract_com(ass_sim* s, int lo, int hi) {  // physics.cpp:77-84
	ENTER_RANGE(s, lo, hi);
	int rc = ass_launch_com(s, lo, hi, false, s->d_red);
	if (rc) return rc;
	return ass_launch_shift(s, lo, hi, s->d_red, false, -1.0, 0.0);
}
extern "C" int ass_subtract_cov(ass_sim* s, int lo, int hi) {  // physics.cpp:88-94
	ENTER_RANGE(s, lo, hi);
	int rc = ass_launch_com(s, lo, hi, true, s->d_red);
	if (rc) return rc;
	return ass_launch_shift(s, lo, hi, s->d_red, true, -1.0, 0.0);
}
extern "C" int ass_move_resolved(ass_sim* s, const double dx[3], const double dv[3], int lo, int hi) {
	ENTER_RANGE(s, lo, hi);
	REQUIRE(dx && dv, ASS_EINVAL, "null shift");
	if (hi == lo) return ASS_OK;
	ass_tic(s, ASS_K_REDUCE);
	k_move<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, lo, hi, dx[0], dx[1], dx[2], dv[0], dv[1], dv[2]);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}
extern "C" int ass_spin_body(ass_sim* s, int lo, int hi, const double omega[3]) {
	ENTER_RANGE(s, lo, hi);
	REQUIRE(omega, ASS_EINVAL, "null omega");
	if (hi == lo) return ASS_OK;
	// operate only if the spin is large enough (physics.cpp:194)
	if (!(sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]) > 1e-5)) return ASS_OK;
	int rc = ass_launch_com(s, lo, hi, false, s->d_red);
	if (rc || (rc = ass_launch_com(s, lo, hi, true, s->d_red + 4))) return rc;
	ass_tic(s, ASS_K_REDUCE);
	k_spin<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, lo, hi, s->d_red, s->d_red + 4, omega[0], omega[1], omega[2]);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

extern "C" int ass_transform_body(ass_sim* s, int lo, int hi, const double R[9], int about_com) {
	ENTER_RANGE(s, lo, hi);
	REQUIRE(R, ASS_EINVAL, "null matrix");
	if (hi == lo) return ASS_OK;
	int rc;
	if (about_com) {
		if ((rc = ass_launch_com(s, lo, hi, false, s->d_red)) || (rc = ass_launch_com(s, lo, hi, true, s->d_red + 4))) return rc;
	}
	ass_tic(s, ASS_K_REDUCE);
	k_transform<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, lo, hi, about_com ? s->d_red : nullptr, about_com ? s->d_red + 4 : nullptr, R[0],
			R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8]);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_REDUCE);
	return ASS_OK;
}

extern "C" int ass_diagnostics(ass_sim* s, int lo, int hi, int want_gpe, double out[20]) {
	ENTER_RANGE(s, lo, hi);
	REQUIRE(out, ASS_EINVAL, "null out");
	REQUIRE(hi > lo, ASS_EINVAL, "empty particle range");
	int rc;
	double* d = s->d_red;  // [0..3] com+mass, [4..7] cov+mass, [8..17] node sums, [18..19] spring sums, [20] gpe
	if ((rc = ass_launch_com(s, lo, hi, false, d))) return rc;
	if ((rc = ass_launch_com(s, lo, hi, true, d + 4))) return rc;
	const int grid = reduce_grid(hi - lo);
	const int gpe_blocks = (hi - lo + PE_T - 1) / PE_T;
	if ((rc = ass_buf_reserve(s->red_partial, sizeof(double) * (size_t) std::max(MAXV * std::max(grid, ass_sm_count() * 4), gpe_blocks)))) return rc;
	double* part = (double*) s->red_partial.p;
	ass_tic(s, ASS_K_REDUCE);
	k_node_diag<<<grid, RB, 0, s->stream>>>(s->posm, s->vel, lo, hi, d, d + 4, part);
	LAUNCH_CHECK();
	k_fold<10><<<1, RB, 0, s->stream>>>(part, grid, d + 8);
	LAUNCH_CHECK();
	if (s->sum_sequential) {  // set-up: the inertia tensor in the reference's order (its eigenvectors feed rotate_to_principal)
		k_inertia_seq<<<1, 256, 0, s->stream>>>(s->posm, lo, hi, d, d + 12);
		LAUNCH_CHECK();
	}
	if (s->ns > 0) {
		const int g2 = reduce_grid(s->n);
		k_spring_diag<<<g2, RB, 0, s->stream>>>(s->posm, s->vel, s->order, s->row_ptr, s->he, s->pal, s->n, part);
		LAUNCH_CHECK();
		k_fold<2><<<1, RB, 0, s->stream>>>(part, g2, d + 18);
		LAUNCH_CHECK();
	} else {
		CU_TRY(cudaMemsetAsync(d + 18, 0, 2 * sizeof(double), s->stream));
	}
	if (want_gpe) {
		k_gpe<<<gpe_blocks, PE_T, 0, s->stream>>>(s->posm, lo, hi, part);
		LAUNCH_CHECK();
		k_fold<1><<<1, RB, 0, s->stream>>>(part, gpe_blocks, d + 20);
		LAUNCH_CHECK();
	} else {
		CU_TRY(cudaMemsetAsync(d + 20, 0, sizeof(double), s->stream));
	}
	ass_toc(s, ASS_K_REDUCE);
	if ((rc = fetch(s, 21))) return rc;
	const double* h = s->h_red;
	out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
	out[3] = h[4]; out[4] = h[5]; out[5] = h[6];
	out[6] = h[8]; out[7] = h[9]; out[8] = h[10];
	out[9] = h[11];
	out[10] = h[18];
	out[11] = want_gpe ? -s->prm.G * 0.5 * h[20] : 0.0;
	out[12] = h[19];
	out[13] = h[12]; out[14] = h[13]; out[15] = h[14]; out[16] = h[15]; out[17] = h[16]; out[18] = h[17];
	out[19] = h[3];
	return ASS_OK;
}
