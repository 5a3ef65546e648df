// ensemble.cu -- B independent small bodies (spin-rate / material sweeps, BASELINE config 5: 4096 x ~1k nodes).
//
// The members never interact, so there is nothing to exchange: one CTA owns one body for a whole batch of steps.
// The body's node state (posm, vel, acc: 96 B/node, ~100 KB for 1065 nodes) lives in shared memory for the batch;
// only the half-edge records (2 x 32 B per spring, ~0.9 MB per body) are streamed from HBM/L2 each step, so a step
// costs one pass over the spring data and no launch.  Per step inside the kernel:
//     part1 | [gravity: all pairs in shared memory] | springs (same arithmetic and lane order as springs.cu, so a member
//     gets the bits the single-body path would give it) | kick + drift (+ next part1) | [center_sim]
// with __syncthreads between the phases.  Across GPUs the members are dealt round-robin by the host launcher; no
// collective is involved.
#include <algorithm>
#include <vector>

#include "spring_math.cuh"

namespace {

constexpr int ET = 512;  // threads per CTA (64 node groups)

struct BodyDesc {
	int node_off, n, row_off, n_perts;
	double dt, negG, soft2;
	int gravity, af, heartbeat, pad;
};

__global__ void __launch_bounds__(ET, 2) k_ensemble(const BodyDesc* __restrict__ desc, vec4d* __restrict__ posm_g, vec4d* __restrict__ vel_g,
		vec4d* __restrict__ acc_g, const int* __restrict__ row_ptr, const HalfEdge* __restrict__ he, int nsteps, int with_hb) {
	extern __shared__ __align__(32) unsigned char smem_raw[];
	__shared__ double red[ET / 32][4];
	__shared__ double com[4];
	const BodyDesc B = desc[blockIdx.x];
	const int n = B.n, tid = threadIdx.x;
	vec4d* sp = (vec4d*) smem_raw;
	vec4d* sv = sp + n;
	vec4d* sa = sv + n;
	for (int i = tid; i < n; i += ET) {
		sp[i] = posm_g[B.node_off + i];
		sv[i] = vel_g[B.node_off + i];
		sa[i] = acc_g[B.node_off + i];
	}
	__syncthreads();
	const int* row = row_ptr + B.row_off;
	const double dt = B.dt, h = __dmul_rn(0.5, dt);
	const bool hb = with_hb && B.heartbeat == 1;
	const bool have_springs = B.af != 0;
	const bool grav = B.gravity == 1 && B.af != 2;
	const int lane = tid & (SPRING_LANES - 1), group = tid / SPRING_LANES;
	bool predrifted = false;
	for (int step = 0; step < nsteps; step++) {
		const bool last = step == nsteps - 1;
		if (!predrifted) {  // reb_integrator_leapfrog_part1, integrator_leapfrog.c:44-48
			for (int i = tid; i < n; i += ET) {
				vec4d p = sp[i];
				const vec4d v = sv[i];
				p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
				p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
				p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
				sp[i] = p;
			}
			__syncthreads();
		}
		if (grav) {  // REB_GRAVITY_BASIC, gravity.c:72-103, all pairs out of shared memory
			for (int i = tid; i < n; i += ET) {
				const vec4d pi = sp[i];
				double ax = 0.0, ay = 0.0, az = 0.0;
				for (int j = 0; j < n; j++) {
					const vec4d pj = sp[j];
					const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
					double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, B.soft2)));
					r2 = (j == i) ? 1.0 : r2;
					const double y = fast_rsqrt(r2);
					const double s = (y * y) * (y * pj.w);
					ax = fma(s, dx, ax); ay = fma(s, dy, ay); az = fma(s, dz, az);
				}
				vec4d a;
				a.x = B.negG * ax; a.y = B.negG * ay; a.z = B.negG * az; a.w = 0.0;
				sa[i] = a;
			}
			__syncthreads();
		}
		if (have_springs) {  // spring_forces (+ zero_accel when af == 2)
			for (int base = 0; base < n; base += ET / SPRING_LANES) {
				const int gid = base + group;
				const bool live = gid < n;
				const int node = live ? gid : n - 1;
				const vec4d pi = sp[node], vi = sv[node];
				const int beg = row[node], end = live ? row[node + 1] : beg;
				const Force3 f = node_spring_sum(he, beg, end, lane, pi, vi, sp, sv);
				if (live && lane == 0) {
					const double inv_m = (end > beg) ? fast_rcp(pi.w) : 0.0;
					vec4d a;
					if (B.af == 2) { a.x = 0.0; a.y = 0.0; a.z = 0.0; a.w = 0.0; }
					else a = sa[node];
					a.x = fma(f.x, inv_m, a.x);
					a.y = fma(f.y, inv_m, a.y);
					a.z = fma(f.z, inv_m, a.z);
					sa[node] = a;
				}
			}
		} else if (B.af == 2) {
			for (int i = tid; i < n; i += ET) { vec4d z; z.x = 0; z.y = 0; z.z = 0; z.w = 0; sa[i] = z; }
		}
		__syncthreads();  // every neighbour read of sp/sv is done before anyone moves
		const bool pre = !last && !hb;
		for (int i = tid; i < n; i += ET) {  // reb_integrator_leapfrog_part2, integrator_leapfrog.c:57-63
			vec4d p = sp[i], v = sv[i];
			const vec4d a = sa[i];
			v.x = __dadd_rn(v.x, __dmul_rn(dt, a.x));
			v.y = __dadd_rn(v.y, __dmul_rn(dt, a.y));
			v.z = __dadd_rn(v.z, __dmul_rn(dt, a.z));
			p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
			p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
			p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
			if (pre) {
				p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
				p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
				p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
			}
			sv[i] = v;
			sp[i] = p;
		}
		predrifted = pre;
		__syncthreads();
		if (hb) {  // center_sim(0, n - n_perts): CoM of the resolved body, shift everything (physics.cpp:98-109)
			const int hi = n - B.n_perts;
			double v[4] = { 0.0, 0.0, 0.0, 0.0 };
			for (int i = tid; i < hi; i += ET) {
				const vec4d p = sp[i];
				v[0] += p.w * p.x; v[1] += p.w * p.y; v[2] += p.w * p.z; v[3] += p.w;
			}
#pragma unroll
			for (int q = 0; q < 4; q++)
				for (int m = 16; m > 0; m >>= 1) v[q] += shfl_down_d(v[q], m);
			if ((tid & 31) == 0)
				for (int q = 0; q < 4; q++) red[tid >> 5][q] = v[q];
			__syncthreads();
			if (tid < 4) {
				double a = 0.0;
				for (int w = 0; w < ET / 32; w++) a += red[w][tid];
				com[tid] = a;
			}
			__syncthreads();
			const double cx = com[0] / com[3], cy = com[1] / com[3], cz = com[2] / com[3];
			for (int i = tid; i < n; i += ET) {
				vec4d p = sp[i];
				p.x = __dadd_rn(p.x, -cx); p.y = __dadd_rn(p.y, -cy); p.z = __dadd_rn(p.z, -cz);
				if (!last) {
					const vec4d w = sv[i];
					p.x = __dadd_rn(p.x, __dmul_rn(h, w.x));
					p.y = __dadd_rn(p.y, __dmul_rn(h, w.y));
					p.z = __dadd_rn(p.z, __dmul_rn(h, w.z));
				}
				sp[i] = p;
			}
			predrifted = !last;
			__syncthreads();
		}
	}
	for (int i = tid; i < n; i += ET) {
		posm_g[B.node_off + i] = sp[i];
		vel_g[B.node_off + i] = sv[i];
		acc_g[B.node_off + i] = sa[i];
	}
}

__global__ void k_unpack(const unsigned char* __restrict__ rows, size_t stride, int n, vec4d* __restrict__ posm, vec4d* __restrict__ vel,
		vec4d* __restrict__ acc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double* r = (const double*) (rows + (size_t) i * stride);
	vec4d p, v, a;
	p.x = r[0]; p.y = r[1]; p.z = r[2]; p.w = r[9];
	v.x = r[3]; v.y = r[4]; v.z = r[5]; v.w = 0.0;
	a.x = r[6]; a.y = r[7]; a.z = r[8]; a.w = 0.0;
	posm[i] = p; vel[i] = v; acc[i] = a;
}
__global__ void k_pack(double* __restrict__ dense, int n, const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const vec4d* __restrict__ acc) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double* r = dense + (size_t) i * 10;
	const vec4d p = posm[i], v = vel[i], a = acc[i];
	r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = v.x; r[4] = v.y; r[5] = v.z; r[6] = a.x; r[7] = a.y; r[8] = a.z; r[9] = p.w;
}

}  // namespace

struct ass_ensemble {
	int n_bodies = 0, total_nodes = 0, max_n = 0;
	long long total_he = 0;
	std::vector<ass_params> prm;
	std::vector<BodyDesc> desc;
	BodyDesc* d_desc = nullptr;
	vec4d *posm = nullptr, *vel = nullptr, *acc = nullptr;
	int* row_ptr = nullptr;
	HalfEdge* he = nullptr;
	double* d_dense = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	size_t smem = 0;
};

static void ens_free(ass_ensemble* e) {
	if (!e) return;
	cudaFree(e->d_desc); cudaFree(e->posm); cudaFree(e->vel); cudaFree(e->acc); cudaFree(e->row_ptr); cudaFree(e->he); cudaFree(e->d_dense);
	if (e->e0) cudaEventDestroy(e->e0);
	if (e->e1) cudaEventDestroy(e->e1);
	if (e->stream) cudaStreamDestroy(e->stream);
	cudaGetLastError();
	delete e;
}

static void fill_desc(ass_ensemble* e) {
	for (int b = 0; b < e->n_bodies; b++) {
		BodyDesc& d = e->desc[b];
		const ass_params& p = e->prm[b];
		d.n_perts = p.n_perts; d.dt = p.dt; d.negG = -p.G; d.soft2 = p.softening * p.softening;
		d.gravity = p.gravity; d.af = p.additional_forces; d.heartbeat = p.heartbeat; d.pad = 0;
	}
}

extern "C" int ass_ensemble_create(ass_ensemble** out, int n_bodies, const int* node_off, const int* spring_off, const void* base,
		size_t stride, const ass_spring* springs, const ass_params* params) {
	REQUIRE(out && node_off && spring_off && base && params, ASS_EINVAL, "null argument");
	REQUIRE(n_bodies > 0, ASS_EINVAL, "need at least one body");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "bad row stride");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int total_nodes = node_off[n_bodies];
	const int total_springs = spring_off[n_bodies];
	REQUIRE(node_off[0] == 0 && spring_off[0] == 0, ASS_EINVAL, "offsets must start at 0");
	REQUIRE(total_springs == 0 || springs, ASS_EINVAL, "null spring array");
	ass_ensemble* e = new ass_ensemble();
	e->n_bodies = n_bodies;
	e->total_nodes = total_nodes;
	e->prm.assign(params, params + n_bodies);
	e->desc.resize(n_bodies);
	// CSR per body: row entries are offsets into the concatenated half-edge array; indices stay local to the body
	std::vector<int> row((size_t) total_nodes + n_bodies, 0);
	std::vector<HalfEdge> he((size_t) 2 * total_springs);
	size_t he_base = 0;
	for (int b = 0; b < n_bodies; b++) {
		const int n = node_off[b + 1] - node_off[b], ns = spring_off[b + 1] - spring_off[b];
		if (n <= 0 || ns < 0) { ens_free(e); ass_set_error("ass_ensemble_create: body %d is empty or offsets decrease", b); return ASS_EINVAL; }
		const ass_params& p = params[b];
		if ((p.gravity != 0 && p.gravity != 1) || p.additional_forces < 0 || p.additional_forces > 2 || (p.heartbeat != 0 && p.heartbeat != 1)) {
			ens_free(e); ass_set_error("ass_ensemble_create: bad params for body %d", b); return ASS_EINVAL;
		}
		e->max_n = std::max(e->max_n, n);
		BodyDesc& d = e->desc[b];
		d.node_off = node_off[b]; d.n = n; d.row_off = node_off[b] + b;
		int* r = row.data() + d.row_off;  // n + 1 entries
		const ass_spring* sp = springs + spring_off[b];
		std::vector<int> deg((size_t) n + 1, 0);
		for (int q = 0; q < ns; q++) {
			const int i = sp[q].particle_1, j = sp[q].particle_2;
			if (i < 0 || j < 0 || i >= n || j >= n || i == j) {
				ens_free(e);
				ass_set_error("ass_ensemble_create: body %d spring %d joins (%d,%d): indices are local to the body and must differ", b, q, i, j);
				return ASS_EINVAL;
			}
			deg[(size_t) i + 1]++; deg[(size_t) j + 1]++;
		}
		for (int i = 0; i < n; i++) deg[(size_t) i + 1] += deg[i];
		for (int i = 0; i <= n; i++) r[i] = (int) (he_base + deg[i]);
		std::vector<int> fill(deg.begin(), deg.end() - 1);
		for (int q = 0; q < ns; q++) {
			const int ends[2] = { sp[q].particle_1, sp[q].particle_2 };
			for (int side = 0; side < 2; side++) {
				HalfEdge hrec;
				hrec.k = sp[q].k; hrec.rs0 = sp[q].rs0; hrec.gamma = sp[q].gamma; hrec.nbr = ends[1 - side]; hrec.sid = q;
				he[he_base + fill[ends[side]]++] = hrec;
			}
		}
		he_base += (size_t) 2 * ns;
	}
	e->total_he = (long long) he_base;
	fill_desc(e);
	e->smem = sizeof(vec4d) * 3 * (size_t) e->max_n;
	if (e->smem > 227 * 1024 - 2048) {
		ens_free(e);
		ass_set_error("ass_ensemble_create: a body of %d nodes does not fit in shared memory (limit ~2400); use ass_sim for large bodies", e->max_n);
		return ASS_EINVAL;
	}
	const size_t nb = (size_t) total_nodes * stride;
	void* d_rows = nullptr;
	bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e->e0) == cudaSuccess &&
	          cudaEventCreate(&e->e1) == cudaSuccess && cudaMalloc(&e->d_desc, sizeof(BodyDesc) * n_bodies) == cudaSuccess &&
	          cudaMalloc(&e->posm, sizeof(vec4d) * total_nodes) == cudaSuccess && cudaMalloc(&e->vel, sizeof(vec4d) * total_nodes) == cudaSuccess &&
	          cudaMalloc(&e->acc, sizeof(vec4d) * total_nodes) == cudaSuccess && cudaMalloc(&e->row_ptr, sizeof(int) * row.size()) == cudaSuccess &&
	          cudaMalloc(&e->he, sizeof(HalfEdge) * std::max<size_t>(he.size(), 1)) == cudaSuccess &&
	          cudaMalloc(&e->d_dense, sizeof(double) * 10 * (size_t) total_nodes) == cudaSuccess && cudaMalloc(&d_rows, nb) == cudaSuccess;
	if (!ok) {
		cudaFree(d_rows);
		ass_set_error("ass_ensemble_create: %s", cudaGetErrorString(cudaGetLastError()));
		ens_free(e);
		return ASS_ENOMEM;
	}
	cudaMemcpyAsync(d_rows, base, nb, cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->row_ptr, row.data(), sizeof(int) * row.size(), cudaMemcpyHostToDevice, e->stream);
	if (!he.empty()) cudaMemcpyAsync(e->he, he.data(), sizeof(HalfEdge) * he.size(), cudaMemcpyHostToDevice, e->stream);
	cudaMemcpyAsync(e->d_desc, e->desc.data(), sizeof(BodyDesc) * n_bodies, cudaMemcpyHostToDevice, e->stream);
	k_unpack<<<(total_nodes + 255) / 256, 256, 0, e->stream>>>((const unsigned char*) d_rows, stride, total_nodes, e->posm, e->vel, e->acc);
	g_launch_count++;
	cudaError_t err = cudaStreamSynchronize(e->stream);
	cudaFree(d_rows);
	if (err == cudaSuccess) err = cudaFuncSetAttribute(k_ensemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->smem);
	if (err != cudaSuccess) {
		ass_set_error("ass_ensemble_create: %s", cudaGetErrorString(err));
		ens_free(e);
		return ASS_ECUDA;
	}
	*out = e;
	return ASS_OK;
}

extern "C" int ass_ensemble_set_params(ass_ensemble* e, const ass_params* params) {
	REQUIRE(e && params, ASS_EINVAL, "null argument");
	int rc = ass_ensure_device();
	if (rc) return rc;
	e->prm.assign(params, params + e->n_bodies);
	fill_desc(e);
	CU_TRY(cudaMemcpyAsync(e->d_desc, e->desc.data(), sizeof(BodyDesc) * e->n_bodies, cudaMemcpyHostToDevice, e->stream));
	CU_TRY(cudaStreamSynchronize(e->stream));
	return ASS_OK;
}

extern "C" int ass_ensemble_step(ass_ensemble* e, int nsteps, int with_heartbeat, double* device_ms) {
	REQUIRE(e, ASS_EINVAL, "null ensemble");
	REQUIRE(nsteps >= 0, ASS_EINVAL, "negative step count");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (device_ms) *device_ms = 0.0;
	if (nsteps == 0) return ASS_OK;
	if (device_ms) CU_TRY(cudaEventRecord(e->e0, e->stream));
	k_ensemble<<<e->n_bodies, ET, e->smem, e->stream>>>(e->d_desc, e->posm, e->vel, e->acc, e->row_ptr, e->he, nsteps, with_heartbeat);
	LAUNCH_CHECK();
	for (auto& p : e->prm)
		for (int s = 0; s < nsteps; s++) {  // the reference's own rounding of t (integrator_leapfrog.c:50,65)
			p.t += p.dt / 2.;
			p.t += p.dt / 2.;
			p.dt_last_done = p.dt;
		}
	if (device_ms) {
		CU_TRY(cudaEventRecord(e->e1, e->stream));
		CU_TRY(cudaEventSynchronize(e->e1));
		float ms = 0.f;
		CU_TRY(cudaEventElapsedTime(&ms, e->e0, e->e1));
		*device_ms = ms;
	}
	return ASS_OK;
}

extern "C" int ass_ensemble_download(ass_ensemble* e, void* base, size_t stride, unsigned fields) {
	REQUIRE(e && base, ASS_EINVAL, "null argument");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "bad row stride");
	REQUIRE((fields & ~ASS_F_ALL) == 0 && fields != 0, ASS_EINVAL, "bad field mask");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int n = e->total_nodes;
	k_pack<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_dense, n, e->posm, e->vel, e->acc);
	LAUNCH_CHECK();
	std::vector<double> h((size_t) n * 10);
	CU_TRY(cudaMemcpyAsync(h.data(), e->d_dense, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, e->stream));
	CU_TRY(cudaStreamSynchronize(e->stream));
	for (int i = 0; i < n; i++) {
		double* r = (double*) ((unsigned char*) base + (size_t) i * stride);
		const double* q = h.data() + (size_t) i * 10;
		if (fields & ASS_F_POS) { r[0] = q[0]; r[1] = q[1]; r[2] = q[2]; }
		if (fields & ASS_F_VEL) { r[3] = q[3]; r[4] = q[4]; r[5] = q[5]; }
		if (fields & ASS_F_ACC) { r[6] = q[6]; r[7] = q[7]; r[8] = q[8]; }
		if (fields & ASS_F_MASS) r[9] = q[9];
	}
	return ASS_OK;
}

extern "C" int ass_ensemble_get_params(ass_ensemble* e, ass_params* params) {
	REQUIRE(e && params, ASS_EINVAL, "null argument");
	for (int b = 0; b < e->n_bodies; b++) params[b] = e->prm[b];
	return ASS_OK;
}
extern "C" int ass_ensemble_sync(ass_ensemble* e) {
	REQUIRE(e, ASS_EINVAL, "null ensemble");
	CU_TRY(cudaStreamSynchronize(e->stream));
	return ASS_OK;
}
extern "C" int ass_ensemble_destroy(ass_ensemble* e) {
	if (!e) return ASS_OK;
	ass_ensure_device();
	if (e->stream) cudaStreamSynchronize(e->stream);
	ens_free(e);
	return ASS_OK;
}
