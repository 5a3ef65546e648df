// gravity.cu -- REBOUND's direct-sum self-gravity, REB_GRAVITY_BASIC (reference src/gravity.c:72-103), in FP64.
//
//   a_i = sum_{j < N_active, j != i}  -G m_j (x_i - x_j) / (|x_i - x_j|^2 + softening^2)^(3/2)        for i < N
//
// This is FP64 CUDA-core work (no dense contraction -> no tensor cores).  One thread owns TI targets in registers;
// sources stream through shared memory in tiles of 32-byte {x,y,z,m} vectors, staged by the bulk-copy engine
// (cp.async.bulk -> mbarrier; SASS: UBLKCP) into a two-deep ring so the copy of tile t+1 overlaps the math on tile t.
// Every lane of a warp reads the same source (shared-memory broadcast), so the inner loop is FP64-pipe bound:
//
//   per ordered pair: 3 DADD (d) + 3 DFMA (r^2 incl. softening) + 5 (rsqrt refinement from the MUFU.RSQ64H seed)
//                     + 3 DMUL (y^3 m_j) + 3 DFMA (accumulate)  = 17 FP64 instructions, 1 MUFU.
//
// SURVEY 8d counts 20 flop per ordered pair; at 17 FP64 issue slots per pair the ceiling of this formulation is
// 20/(2*17) = 59% of the DFMA peak.  The reference evaluates sqrt and a true division and applies -G/(r^3) * m_j
// per pair; we factor -G out of the sum and use rsqrt: agreement is ~1e-15 relative per term (tests bound the sum at
// 1e-12 of max|a|).
//
// Load balance: the grid is (target blocks) x (source splits); each split writes its own partial-sum row and a
// fixed-order pass adds the rows, so the result stays deterministic.  The split count is chosen so that the CTA count
// is a near-multiple of the SM count (all CTAs take the same time; see plan_split).  The same machinery lets a sharded
// simulation (shard.cu) evaluate "sources in my slab" and "sources in the other slabs" as separate launches that land
// in different rows and are reduced once.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int GB = 128;   // threads per CTA
constexpr int JT = 256;   // sources per tile (8 KB)

template <int TI>
struct Targets {
	double x[TI], y[TI], z[TI];
	double ax[TI], ay[TI], az[TI];
	int idx[TI];
};

// one source against TI targets; CHECK: skip j == i (only needed when softening == 0)
template <int TI, bool CHECK>
__device__ __forceinline__ void pair_update(Targets<TI>& t, const vec4d sj, int j, double soft2) {
#pragma unroll
	for (int k = 0; k < TI; k++) {
		const double dx = t.x[k] - sj.x, dy = t.y[k] - sj.y, dz = t.z[k] - sj.z;
		double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, soft2)));
		if (CHECK) r2 = (j == t.idx[k]) ? 1.0 : r2;  // d == 0 there, so the term is exactly 0
		const double s = rsqrt_cubed(r2) * sj.w;
		t.ax[k] = fma(s, dx, t.ax[k]);
		t.ay[k] = fma(s, dy, t.ay[k]);
		t.az[k] = fma(s, dz, t.az[k]);
	}
}

// grid.x = target blocks of GB*TI over [t_lo, t_hi); grid.y = source splits of split_len over [s_lo, s_hi).
// Row blockIdx.y of `out` (stride out_stride vectors) receives -G * (partial sum).
template <int TI, bool CHECK>
__global__ void __launch_bounds__(GB) k_gravity(const vec4d* __restrict__ src, int s_lo, int s_hi, const vec4d* __restrict__ tgt,
		int t_lo, int t_hi, int split_len, double soft2, double negG, vec4d* __restrict__ out, size_t out_stride) {
	__shared__ __align__(128) vec4d tile[2][JT];
	__shared__ __align__(8) uint64_t full[2];

	const int tid = threadIdx.x;
	const int i0 = t_lo + blockIdx.x * (GB * TI);
	const int j_beg = s_lo + blockIdx.y * split_len;
	const int j_end = min(s_hi, j_beg + split_len);
	const int n_tiles = (j_end - j_beg + JT - 1) / JT;

	Targets<TI> t;
#pragma unroll
	for (int k = 0; k < TI; k++) {
		const int i = i0 + tid + k * GB;
		t.idx[k] = i;
		const vec4d p = tgt[min(i, t_hi - 1)];
		t.x[k] = p.x; t.y[k] = p.y; t.z[k] = p.z;
		t.ax[k] = 0.0; t.ay[k] = 0.0; t.az[k] = 0.0;
	}

	if (tid == 0) {
		mbar_init(&full[0], 1);
		mbar_init(&full[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	auto issue = [&](int tl) {  // thread 0 only
		const int jb = j_beg + tl * JT;
		const int cnt = min(JT, j_end - jb);
		const uint32_t bytes = (uint32_t) cnt * (uint32_t) sizeof(vec4d);
		mbar_expect_tx(&full[tl & 1], bytes);
		bulk_g2s(&tile[tl & 1][0], src + jb, bytes, &full[tl & 1]);
	};
	if (tid == 0 && n_tiles > 0) issue(0);

	for (int tl = 0; tl < n_tiles; tl++) {
		const int buf = tl & 1;
		// the buffer we are about to refill was last read in iteration tl-1: everyone is past it after this barrier
		__syncthreads();
		if (tid == 0 && tl + 1 < n_tiles) issue(tl + 1);
		mbar_wait(&full[buf], (uint32_t) ((tl >> 1) & 1));
		const int jb = j_beg + tl * JT;
		const int cnt = min(JT, j_end - jb);
		// does this tile contain any of this CTA's own targets?
		const bool diag = CHECK && (jb < i0 + GB * TI) && (jb + cnt > i0);
		if (cnt == JT && !diag) {
#pragma unroll 4
			for (int jj = 0; jj < JT; jj++) pair_update<TI, false>(t, tile[buf][jj], 0, soft2);
		} else if (!diag) {
			for (int jj = 0; jj < cnt; jj++) pair_update<TI, false>(t, tile[buf][jj], 0, soft2);
		} else {
			for (int jj = 0; jj < cnt; jj++) pair_update<TI, CHECK>(t, tile[buf][jj], jb + jj, soft2);
		}
	}

	vec4d* o = out + (size_t) blockIdx.y * out_stride;
#pragma unroll
	for (int k = 0; k < TI; k++) {
		const int i = t.idx[k];
		if (i < t_hi) {
			vec4d a;
			a.x = negG * t.ax[k]; a.y = negG * t.ay[k]; a.z = negG * t.az[k]; a.w = 0.0;
			o[i] = a;
		}
	}
}

// acc[i] = sum over rows, fixed order
__global__ void k_gravity_reduce(const vec4d* __restrict__ part, size_t stride, int rows, int lo, int hi, vec4d* __restrict__ acc, int add) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d a = part[i];
	for (int s = 1; s < rows; s++) {
		const vec4d b = part[(size_t) s * stride + i];
		a.x += b.x; a.y += b.y; a.z += b.z;
	}
	if (add) { const vec4d p = acc[i]; a.x += p.x; a.y += p.y; a.z += p.z; }
	a.w = 0.0;
	acc[i] = a;
}

__global__ void k_zero_range(vec4d* __restrict__ acc, int lo, int hi) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d z; z.x = 0; z.y = 0; z.z = 0; z.w = 0;
	acc[i] = z;
}

int pick_ti(int n_tgt, int sms) {
	if (n_tgt >= sms * GB * 4 * 2) return 4;
	if (n_tgt >= sms * GB * 2) return 2;
	return 1;
}

// All CTAs do the same amount of work, and the block scheduler hands them out as SMs free up, so the makespan is
// ceil(ctas / sms) CTA-times.  Pick the number of source splits that wastes the least of the last round, preferring
// few splits (each costs one more partial-sum row).
void plan_split(int n_tgt, int n_src, int sms, int* ti_out, int* split_out, int* split_len_out) {
	const int ti = pick_ti(n_tgt, sms);
	const int tblocks = (n_tgt + GB * ti - 1) / (GB * ti);
	int best = 1;
	double best_score = -1.0;
	const int max_split = std::max(1, std::min(32, n_src / (4 * JT)));
	for (int s = 1; s <= max_split; s++) {
		const double ctas = (double) tblocks * s;
		const double rounds = ceil(ctas / sms);
		const double score = ctas / (rounds * sms) - 0.002 * s;
		if (score > best_score + 1e-12) { best_score = score; best = s; }
	}
	int split_len = (n_src + best - 1) / best;
	split_len = ((split_len + JT - 1) / JT) * JT;  // tile-aligned split boundaries
	*ti_out = ti;
	*split_len_out = split_len;
	*split_out = (n_src + split_len - 1) / split_len;
}

template <int TI>
void launch_ti(ass_sim* s, const vec4d* src, int s_lo, int s_hi, int t_lo, int t_hi, int split, int split_len, double soft2, double negG,
		vec4d* out, size_t stride) {
	const int tblocks = (t_hi - t_lo + GB * TI - 1) / (GB * TI);
	dim3 grid(tblocks, split);
	if (soft2 > 0.0)
		k_gravity<TI, false><<<grid, GB, 0, s->stream>>>(src, s_lo, s_hi, s->posm, t_lo, t_hi, split_len, soft2, negG, out, stride);
	else
		k_gravity<TI, true><<<grid, GB, 0, s->stream>>>(src, s_lo, s_hi, s->posm, t_lo, t_hi, split_len, soft2, negG, out, stride);
}

}  // namespace

// Number of partial rows a launch over (n_tgt targets, n_src sources) will write.
int ass_gravity_rows(int n_tgt, int n_src) {
	if (n_tgt <= 0 || n_src <= 0) return 0;
	int ti, split, len;
	plan_split(n_tgt, n_src, ass_sm_count(), &ti, &split, &len);
	return split;
}

// Partial sums of targets [t_lo,t_hi) over sources src[s_lo,s_hi) into rows [row_base, row_base + rows) of grav_partial
// (which the caller has reserved).  Returns the number of rows written through *rows_out.
int ass_launch_gravity_partial(ass_sim* s, const vec4d* src, int s_lo, int s_hi, int t_lo, int t_hi, int row_base, int* rows_out) {
	*rows_out = 0;
	if (t_hi <= t_lo || s_hi <= s_lo) return ASS_OK;
	int ti, split, len;
	plan_split(t_hi - t_lo, s_hi - s_lo, ass_sm_count(), &ti, &split, &len);
	const double soft2 = s->prm.softening * s->prm.softening;
	const double negG = -s->prm.G;
	vec4d* out = (vec4d*) s->grav_partial.p + (size_t) row_base * (size_t) s->n;
	ass_tic(s, ASS_K_GRAVITY);
	if (ti == 4) launch_ti<4>(s, src, s_lo, s_hi, t_lo, t_hi, split, len, soft2, negG, out, (size_t) s->n);
	else if (ti == 2) launch_ti<2>(s, src, s_lo, s_hi, t_lo, t_hi, split, len, soft2, negG, out, (size_t) s->n);
	else launch_ti<1>(s, src, s_lo, s_hi, t_lo, t_hi, split, len, soft2, negG, out, (size_t) s->n);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_GRAVITY);
	*rows_out = split;
	return ASS_OK;
}

// acc[t_lo, t_hi) (=|+=) the sum of the first `rows` partial rows
int ass_launch_gravity_reduce(ass_sim* s, int t_lo, int t_hi, int rows, bool add) {
	if (t_hi <= t_lo) return ASS_OK;
	if (rows <= 0 && add) return ASS_OK;
	ass_tic(s, ASS_K_GRAVITY);
	if (rows <= 0) k_zero_range<<<(t_hi - t_lo + 255) / 256, 256, 0, s->stream>>>(s->acc, t_lo, t_hi);
	else k_gravity_reduce<<<(t_hi - t_lo + 255) / 256, 256, 0, s->stream>>>((const vec4d*) s->grav_partial.p, (size_t) s->n, rows, t_lo, t_hi, s->acc, add ? 1 : 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_GRAVITY);
	return ASS_OK;
}

// a = gravity for every particle, all sources local (the unsharded path)
int ass_launch_gravity(ass_sim* s) {
	const int n = s->n;
	if (n == 0) return ASS_OK;
	const int n_src = (s->prm.n_active == -1) ? n : std::max(0, std::min(s->prm.n_active, n));
	// every particle both source and target: each unordered pair can be evaluated once (gravity_sym.cu).  Trailing
	// point masses (params.n_perts, e.g. Mars next to a body of equal-mass nodes) are taken out of the pair-once part so
	// that it keeps its equal-mass form: body x body pair-once, then body <- perturbers and perturbers <- everything
	// one-sided (N x n_perts pairs: nothing).
	if (n_src == n && s->gravity_kernel != 1 && (s->gravity_kernel == 2 || ass_gravity_sym_applicable(n))) {
		const int np = std::max(0, std::min(s->prm.n_perts, n)), nb = n - np;
		if (np == 0 || nb == 0) return ass_launch_gravity_sym(s, 0, n);
		int rc = ass_launch_gravity_sym(s, 0, nb);
		if (rc) return rc;
		const int rows = std::max(ass_gravity_rows(nb, np), ass_gravity_rows(np, n));
		if ((rc = ass_buf_reserve(s->grav_partial, sizeof(vec4d) * (size_t) std::max(rows, 1) * (size_t) n))) return rc;
		int used = 0;
		if ((rc = ass_launch_gravity_partial(s, s->posm, nb, n, 0, nb, 0, &used))) return rc;
		if ((rc = ass_launch_gravity_reduce(s, 0, nb, used, true))) return rc;
		if ((rc = ass_launch_gravity_partial(s, s->posm, 0, n, nb, n, 0, &used))) return rc;
		return ass_launch_gravity_reduce(s, nb, n, used, false);
	}
	const int rows = ass_gravity_rows(n, n_src);
	int rc = ass_buf_reserve(s->grav_partial, sizeof(vec4d) * (size_t) std::max(rows, 1) * (size_t) n);
	if (rc) return rc;
	int used = 0;
	if ((rc = ass_launch_gravity_partial(s, s->posm, 0, n_src, 0, n, 0, &used))) return rc;
	return ass_launch_gravity_reduce(s, 0, n, used);
}

extern "C" int ass_set_gravity_kernel(ass_sim* s, int which) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(which >= 0 && which <= 2, ASS_EINVAL, "0 = auto, 1 = one-sided tiles, 2 = pair-once");
	s->gravity_kernel = which;
	return ASS_OK;
}

extern "C" int ass_gravity_basic(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	REQUIRE(!s->shard.active, ASS_ESTATE, "sharded simulations evaluate gravity inside ass_shard_exchange / ass_shard_finish");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (s->prm.gravity == 0) return ASS_OK;  // REB_GRAVITY_NONE: "Do nothing." (gravity.c:68-70)
	return ass_launch_gravity(s);
}
