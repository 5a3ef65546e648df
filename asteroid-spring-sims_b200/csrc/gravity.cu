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
// Load balance: the grid is (target blocks) x (source splits); each split writes its own partial sum row and a
// fixed-order pass adds the rows, so the result stays deterministic.  The split count is chosen so that the CTA count
// is a near-multiple of the SM count (all CTAs take the same time; see choose_split).
#include "common.cuh"

namespace {

constexpr int GB = 128;   // threads per CTA
constexpr int JT = 256;   // sources per tile (8 KB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
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
// 1-D bulk copy global -> shared, completion signalled on the mbarrier (TMA engine, no tensor map needed)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
			"l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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
		const double y = fast_rsqrt(r2);
		const double s = (y * y) * (y * sj.w);
		t.ax[k] = fma(s, dx, t.ax[k]);
		t.ay[k] = fma(s, dy, t.ay[k]);
		t.az[k] = fma(s, dz, t.az[k]);
	}
}

// grid.x = target blocks of GB*TI, grid.y = source splits.
// src: source vectors (length >= j_end rounded up to... see padding note), tgt: target vectors.
template <int TI, bool CHECK>
__global__ void __launch_bounds__(GB) k_gravity(const vec4d* __restrict__ src, int n_src, const vec4d* __restrict__ tgt,
		int t_lo, int t_hi, int split_len, double soft2, double negG, vec4d* __restrict__ out, size_t out_stride) {
	__shared__ __align__(128) vec4d tile[2][JT];
	__shared__ __align__(8) uint64_t full[2];

	const int tid = threadIdx.x;
	const int i0 = t_lo + blockIdx.x * (GB * TI);
	const int j_beg = blockIdx.y * split_len;
	const int j_end = min(n_src, j_beg + split_len);
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

// acc[i] = sum over splits, fixed order
__global__ void k_gravity_reduce(const vec4d* __restrict__ part, size_t stride, int splits, int lo, int hi, vec4d* __restrict__ acc) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d a = part[i];
	for (int s = 1; s < splits; s++) {
		const vec4d b = part[(size_t) s * stride + i];
		a.x += b.x; a.y += b.y; a.z += b.z;
	}
	a.w = 0.0;
	acc[i] = a;
}

// zero the accelerations of particles that are not targets of this evaluation (gravity.c:76-80 zeroes all N)
__global__ void k_zero_range(vec4d* __restrict__ acc, int lo, int hi) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d z; z.x = 0; z.y = 0; z.z = 0; z.w = 0;
	acc[i] = z;
}

// All CTAs do the same amount of work, and the block scheduler hands them out as SMs free up, so the makespan is
// ceil(ctas / sms) CTA-times.  Pick the number of source splits that wastes the least of the last round, preferring
// few splits (each costs one more partial-sum row).
int choose_split(int target_blocks, int n_src, int sms) {
	int best = 1;
	double best_score = -1.0;
	const int max_split = std::max(1, std::min(32, n_src / (4 * JT)));
	for (int s = 1; s <= max_split; s++) {
		const double ctas = (double) target_blocks * s;
		const double rounds = ceil(ctas / sms);
		const double eff = ctas / (rounds * sms);
		const double score = eff - 0.002 * s;
		if (score > best_score + 1e-12) { best_score = score; best = s; }
	}
	return best;
}

template <int TI>
int launch_gravity_ti(ass_sim* s, const vec4d* src, int n_src, int t_lo, int t_hi, double soft2, double negG) {
	const int n_tgt = t_hi - t_lo;
	const int tblocks = (n_tgt + GB * TI - 1) / (GB * TI);
	int split = choose_split(tblocks, n_src, ass_sm_count());
	int split_len = (n_src + split - 1) / split;
	split_len = ((split_len + JT - 1) / JT) * JT;  // tile-aligned split boundaries keep every bulk copy 16 B aligned
	split = (n_src + split_len - 1) / split_len;
	vec4d* out = s->acc;
	size_t stride = 0;
	if (split > 1) {
		int rc = ass_buf_reserve(s->grav_partial, sizeof(vec4d) * (size_t) split * (size_t) s->n);
		if (rc) return rc;
		out = (vec4d*) s->grav_partial.p;
		stride = (size_t) s->n;
	}
	dim3 grid(tblocks, split);
	ass_tic(s, ASS_K_GRAVITY);
	if (soft2 > 0.0)
		k_gravity<TI, false><<<grid, GB, 0, s->stream>>>(src, n_src, s->posm, t_lo, t_hi, split_len, soft2, negG, out, stride);
	else
		k_gravity<TI, true><<<grid, GB, 0, s->stream>>>(src, n_src, s->posm, t_lo, t_hi, split_len, soft2, negG, out, stride);
	LAUNCH_CHECK();
	if (split > 1) {
		k_gravity_reduce<<<(n_tgt + 255) / 256, 256, 0, s->stream>>>(out, stride, split, t_lo, t_hi, s->acc);
		LAUNCH_CHECK();
	}
	ass_toc(s, ASS_K_GRAVITY);
	return ASS_OK;
}

}  // namespace

// a = gravity for targets [t_lo, t_hi) from sources src[0, n_src)
int ass_launch_gravity(ass_sim* s) {
	const int n = s->n;
	if (n == 0) return ASS_OK;
	const int n_src = (s->prm.n_active == -1) ? n : std::min(s->prm.n_active, n);
	const double soft2 = s->prm.softening * s->prm.softening;
	const double negG = -s->prm.G;
	int t_lo = 0, t_hi = n;
	const vec4d* src = s->posm;
	if (s->shard.active) {
		t_lo = s->shard.lo;
		t_hi = s->shard.hi;
		src = s->shard.window;  // filled by ass_shard_gather / read in place from peers (shard.cu)
	}
	if (n_src <= 0 || t_hi <= t_lo) {
		k_zero_range<<<(n + 255) / 256, 256, 0, s->stream>>>(s->acc, 0, n);
		LAUNCH_CHECK();
		return ASS_OK;
	}
	const int n_tgt = t_hi - t_lo;
	const int sms = ass_sm_count();
	if (n_tgt >= sms * GB * 4 * 2) return launch_gravity_ti<4>(s, src, n_src, t_lo, t_hi, soft2, negG);
	if (n_tgt >= sms * GB * 2) return launch_gravity_ti<2>(s, src, n_src, t_lo, t_hi, soft2, negG);
	return launch_gravity_ti<1>(s, src, n_src, t_lo, t_hi, soft2, negG);
}

extern "C" int ass_gravity_basic(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (s->prm.gravity == 0) return ASS_OK;  // REB_GRAVITY_NONE: "Do nothing." (gravity.c:68-70)
	return ass_launch_gravity(s);
}
