// gravity_sym.cu -- REB_GRAVITY_BASIC (reference src/gravity.c:72-103) with every unordered pair evaluated ONCE.
//
// The reference's double loop computes pair (i,j) and pair (j,i) separately; they share everything but the mass factor
// and the sign (Newton's third law).  Evaluating the pair once and applying it to both particles costs
//     3 DADD (d) + 3 DFMA (r^2) + 6 (y^3 = r^-3 from the rsqrt seed) + 2 DMUL (y^3 m_r, y^3 m_s) + 6 DFMA (both sides)
//   = 20 FP64 instructions per unordered pair = 10 per ordered pair, against 16 in the one-sided kernel (gravity.cu),
// i.e. against SURVEY 8d's 20 flop per ordered pair the ceiling moves from 62 % to 100 % of the DFMA peak.
// When every particle of the range carries the same mass -- the generators give every node M/N (shapes.cpp:262-267), so
// this is the normal case -- the two mass products drop out of the pair (m is applied once per particle by the
// reduction): 18 instructions per unordered pair.  The plan checks the masses on the device and picks the kernel.
//
// Layout of the computation (all deterministic, no atomics):
//   * particles are cut into blocks of P = 1024.  A CTA keeps ONE block "resident" in registers (4 particles per
//     thread: position, mass, accumulator) and streams other blocks past it through shared memory.
//   * a warp takes 32 streamed particles at a time and runs a 32-step systolic rotation: at step s lane l meets streamed
//     particle (l+s) mod 32 -- its position comes from shared memory (conflict-free, every lane a different particle),
//     its accumulator travels from lane to lane by shuffle -- and evaluates the 4 pairs with its residents.  After 32
//     steps every resident of the warp has met every streamed particle of the chunk, each pair exactly once.
//   * streamed-side sums are folded over the CTA's 8 warps in shared memory (fixed order) and written to a scratch row
//     owned by the tile (resident block, streamed block); resident-side sums go to a scratch row owned by the CTA.
//     k_gravity_sym_reduce adds, for every particle, its rows in a fixed order and applies -G.
//   * tiles: strictly-lower-triangle pairs of blocks are symmetric; the diagonal tile (a block against itself) is
//     evaluated one-sided with the self pair masked.  CTAs take runs of at most U tiles of one resident block; the
//     block scheduler balances them (thousands of equal-size CTAs).
// Scratch: N^2/(2P) x 24 B for the streamed side (110 MB at N = 1e5, 11 GB at N = 1e6 -- HBM is 180 GB and the rows are
// written and read once per evaluation: < 2 % of the kernel's time).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int SG_T = 256;            // threads per CTA
constexpr int SG_W = SG_T / 32;      // warps
// Resident particles per thread R and CTAs per SM come in two compiled variants, chosen per plan (get_plan):
//   R = 4, 2 CTAs / SM (128 registers): blocks of 1024 -- small problems (a ~1e4-particle body, the slabs of C3 / 8): fewer,
//          fatter tiles quantise them better and every streamed particle's shared-memory load and shuffles serve 4 pairs;
//   R = 3, 3 CTAs / SM (80 registers): blocks of 768, six warps per scheduler -- large problems: +1.9 % at C3
//          (profiles/r2_gravity_variants.md), where the FP64 pipe idles 20 % of the time with four warps per scheduler.
constexpr int SG_R_MAX = 4;
constexpr int sg_occ(int R) { return R == 3 ? 3 : 2; }
constexpr int SG_TILE = 128;         // streamed particles between block-wide reductions (4 chunks of 32)

struct SymCta {
	int rb, s0, s1;  // resident block, streamed blocks [s0, s1); a streamed block equal to rb is the diagonal tile
	int t_lo, t_hi;  // streamed particles [t_lo, t_hi) of each of those blocks (multiples of 32); the whole block is
	                 // [0, P).  Bodies of ~1e4 particles have too few tiles to fill the machine: their tiles are cut.
};

__device__ __forceinline__ vec4d load_or_dummy(const vec4d* __restrict__ posm, int idx, int lo, int hi) {
	if (idx >= lo && idx < hi) return posm[idx];
	// padding: massless and so far away that y^3 = r^-3 underflows to exactly 0 while r^2 (~3e300) stays finite, so a
	// dummy contributes exactly 0 to real particles with or without the mass factor; dummies are pairwise distinct
	vec4d p;
	p.x = 1.0e150; p.y = 1.0e150; p.z = 1.0e150 + 1.0e147 * (double) (idx & 0xffff); p.w = 0.0;
	return p;
}

// y^3 = r2^(-3/2) from the hardware seed y0 = MUFU.RSQ64H(r2) (relative error ~2^-22): with eps = 1 - r2 y0^2,
//   r2^(-3/2) = y0^3 (1 - eps)^(-3/2) = y0^3 (1 + eps (3/2 + 15/8 eps)) + O(eps^3)        [35/16 eps^3 ~ 2^-62]
// SIX FP64 instructions (common.cuh: rsqrt_cubed): the seed has 20 mantissa bits, so y0^2 is exact and eps is one FMA;
// y0^3 does not wait for the correction.
__device__ __forceinline__ double rsqrt_seed(double a) {
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
	return y;
}

// One streamed particle meets the thread's SG_R residents.  Written stage by stage across the residents (not resident
// by resident) so the SG_R dependent chains are interleaved in program order: FP64 results are needed several issue
// slots after they are produced and the warp does not sit in fixed-latency waits.
template <bool DIAG, bool UNI, int SG_R>
__device__ __forceinline__ void meet(const double2 xy, const double2 zm, int sidx, const double (&rx)[SG_R], const double (&ry)[SG_R],
		const double (&rz)[SG_R], const double (&rm)[SG_R], const int (&ridx)[SG_R], double (&rax)[SG_R], double (&ray)[SG_R],
		double (&raz)[SG_R], double& sax, double& say, double& saz, double soft2) {
	double dx[SG_R], dy[SG_R], dz[SG_R], r2[SG_R], y0[SG_R], e[SG_R], c[SG_R], y3[SG_R];
#pragma unroll
	for (int k = 0; k < SG_R; k++) { dx[k] = xy.x - rx[k]; dy[k] = xy.y - ry[k]; dz[k] = zm.x - rz[k]; }  // d = x_streamed - x_resident
#pragma unroll
	for (int k = 0; k < SG_R; k++) r2[k] = fma(dx[k], dx[k], soft2);
#pragma unroll
	for (int k = 0; k < SG_R; k++) r2[k] = fma(dy[k], dy[k], r2[k]);
#pragma unroll
	for (int k = 0; k < SG_R; k++) {
		r2[k] = fma(dz[k], dz[k], r2[k]);
		if (DIAG) r2[k] = (sidx == ridx[k]) ? 1.0 : r2[k];  // self pair: d == 0, contributes exactly 0
	}
#pragma unroll
	for (int k = 0; k < SG_R; k++) y0[k] = rsqrt_seed(r2[k]);
#pragma unroll
	for (int k = 0; k < SG_R; k++) c[k] = y0[k] * y0[k];           // exact: the seed has 20 mantissa bits
#pragma unroll
	for (int k = 0; k < SG_R; k++) e[k] = fma(-r2[k], c[k], 1.0);  // eps = 1 - r2 y0^2, one rounding
#pragma unroll
	for (int k = 0; k < SG_R; k++) c[k] = c[k] * y0[k];
#pragma unroll
	for (int k = 0; k < SG_R; k++) y3[k] = fma(1.875, e[k], 1.5);
#pragma unroll
	for (int k = 0; k < SG_R; k++) y3[k] = e[k] * y3[k];
#pragma unroll
	for (int k = 0; k < SG_R; k++) y3[k] = fma(c[k], y3[k], c[k]);
	if (UNI) {
		// equal masses: S_r = sum y^3 (x_s - x_r), S_s likewise; the reduction multiplies by G m
#pragma unroll
		for (int k = 0; k < SG_R; k++) {
			rax[k] = fma(y3[k], dx[k], rax[k]);
			ray[k] = fma(y3[k], dy[k], ray[k]);
			raz[k] = fma(y3[k], dz[k], raz[k]);
		}
		if (!DIAG) {
#pragma unroll
			for (int k = 0; k < SG_R; k++) {
				sax = fma(y3[k], dx[k], sax);
				say = fma(y3[k], dy[k], say);
				saz = fma(y3[k], dz[k], saz);
			}
		}
	} else {
#pragma unroll
		for (int k = 0; k < SG_R; k++) {
			const double ts = y3[k] * zm.y;  // streamed mass -> acts on the resident
			rax[k] = fma(ts, dx[k], rax[k]);  // S_r = sum m_s y^3 (x_s - x_r)  [a_r = +G S_r]
			ray[k] = fma(ts, dy[k], ray[k]);
			raz[k] = fma(ts, dz[k], raz[k]);
		}
		if (!DIAG) {
#pragma unroll
			for (int k = 0; k < SG_R; k++) {
				const double tr = y3[k] * rm[k];  // resident mass -> acts on the streamed particle
				sax = fma(tr, dx[k], sax);        // S_s = sum m_r y^3 (x_s - x_r)  [a_s = -G S_s]
				say = fma(tr, dy[k], say);
				saz = fma(tr, dz[k], saz);
			}
		}
	}
}

// Resident particles come from [rlo, rhi), streamed particles from [slo, shi).  rect == 0: the two ranges are the same
// range and the tiles are its lower triangle incl. the diagonal; rect != 0: the ranges are disjoint (two slabs of a
// sharded body) and the tiles are the full rectangle, nsb streamed blocks wide.
template <bool UNI, int SG_R>
__global__ void __launch_bounds__(SG_T, sg_occ(SG_R)) k_gravity_sym(const vec4d* __restrict__ posm, int rlo, int rhi, int slo, int shi, int rpad, int spad, int rect,
		int nsb, const SymCta* __restrict__ ctas, double soft2, double* __restrict__ res, size_t res_plane, double* __restrict__ str, size_t str_plane) {
	__shared__ double2 s_xy[SG_TILE], s_zm[SG_TILE];
	__shared__ double s_acc[SG_W][3][SG_TILE];
	constexpr int SG_P = SG_T * SG_R;  // particles per block
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const SymCta C = ctas[blockIdx.x];

	double rx[SG_R], ry[SG_R], rz[SG_R], rm[SG_R], rax[SG_R], ray[SG_R], raz[SG_R];
	int ridx[SG_R];
#pragma unroll
	for (int k = 0; k < SG_R; k++) {
		ridx[k] = rlo - rpad + C.rb * SG_P + tid + k * SG_T;  // blocks are counted from rlo - rpad: the padding leads block 0
		const vec4d p = load_or_dummy(posm, ridx[k], rlo, rhi);
		rx[k] = p.x; ry[k] = p.y; rz[k] = p.z; rm[k] = p.w;
		rax[k] = 0.0; ray[k] = 0.0; raz[k] = 0.0;
	}

	for (int sb = C.s0; sb < C.s1; sb++) {
		const bool diag = !rect && (sb == C.rb);
		const int sbase = slo - spad + sb * SG_P;
		const size_t tile_row = (rect ? (size_t) C.rb * (size_t) nsb + (size_t) sb                       // rectangle, row-major
		                              : (size_t) C.rb * (size_t) (C.rb - 1) / 2 + (size_t) sb) * (size_t) SG_P;  // strict lower triangle
		for (int t0 = C.t_lo; t0 < C.t_hi; t0 += SG_TILE) {
			const int cnt = min(SG_TILE, C.t_hi - t0);  // a piece may end inside a tile (multiples of 32)
			vec4d mine;
			if (tid < SG_TILE) mine = load_or_dummy(posm, sbase + t0 + tid, slo, shi);
			__syncthreads();  // the previous tile's fold has finished reading s_acc / nobody still reads s_xy
			if (tid < SG_TILE) {
				s_xy[tid] = make_double2(mine.x, mine.y);
				s_zm[tid] = make_double2(mine.z, mine.w);
			}
			__syncthreads();
			if (sbase + t0 >= shi || sbase + t0 + cnt <= slo) continue;  // a tile of pure padding (uniform across the CTA)
#pragma unroll 1
			for (int c0 = 0; c0 < cnt; c0 += 32) {
				double sax = 0.0, say = 0.0, saz = 0.0;
				if (diag) {
#pragma unroll 2
					for (int s = 0; s < 32; s++) {
						const int q = c0 + ((lane + s) & 31);
						meet<true, UNI, SG_R>(s_xy[q], s_zm[q], sbase + t0 + q, rx, ry, rz, rm, ridx, rax, ray, raz, sax, say, saz, soft2);
					}
				} else {
#pragma unroll 2
					for (int s = 0; s < 32; s++) {
						const int q = c0 + ((lane + s) & 31);
						meet<false, UNI, SG_R>(s_xy[q], s_zm[q], 0, rx, ry, rz, rm, ridx, rax, ray, raz, sax, say, saz, soft2);
						// the streamed particle moves on to the lane below; take over the one from the lane above
						sax = __shfl_sync(0xffffffffu, sax, (lane + 1) & 31);
						say = __shfl_sync(0xffffffffu, say, (lane + 1) & 31);
						saz = __shfl_sync(0xffffffffu, saz, (lane + 1) & 31);
					}
					// after 32 steps lane l holds the sums of streamed particle c0 + l again
					s_acc[warp][0][c0 + lane] = sax;
					s_acc[warp][1][c0 + lane] = say;
					s_acc[warp][2][c0 + lane] = saz;
				}
			}
			if (!diag) {
				__syncthreads();
				for (int v = tid; v < 3 * cnt; v += SG_T) {
					const int comp = v / cnt, q = v - comp * cnt;
					double a = s_acc[0][comp][q];
#pragma unroll
					for (int w = 1; w < SG_W; w++) a += s_acc[w][comp][q];
					str[(size_t) comp * str_plane + tile_row + (size_t) (t0 + q)] = a;
				}
			}
		}
	}
#pragma unroll
	for (int k = 0; k < SG_R; k++) {
		const size_t at = (size_t) blockIdx.x * SG_P + (size_t) (tid + k * SG_T);
		res[at] = rax[k];
		res[res_plane + at] = ray[k];
		res[2 * res_plane + at] = raz[k];
	}
}

// Triangle plans: a_i = G * (resident rows) - G * (streamed rows) for every particle of the range, rows added in a fixed
// order.  be[2b], be[2b+1] = the CTAs whose resident block is b.  uniform != 0: the rows hold unweighted sums and every
// particle has the mass of particle `lo`.  add != 0: out[i] += a_i instead of out[i] = a_i.
__global__ void k_gravity_sym_reduce(const double* __restrict__ res, size_t res_plane, const double* __restrict__ str, size_t str_plane,
		const int* __restrict__ be, int nb, int lo, int hi, int pad, int SG_P, double G, int uniform, const vec4d* __restrict__ posm, vec4d* __restrict__ out, int add) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	if (uniform) G *= posm[lo].w;
	const int b = (i - lo + pad) / SG_P, o = (i - lo + pad) - b * SG_P;
	double rs[3] = { 0.0, 0.0, 0.0 }, ss[3] = { 0.0, 0.0, 0.0 };
	for (int c = be[2 * b]; c < be[2 * b + 1]; c++) {
		const size_t at = (size_t) c * SG_P + o;
		rs[0] += res[at]; rs[1] += res[res_plane + at]; rs[2] += res[2 * res_plane + at];
	}
	for (int rb = b + 1; rb < nb; rb++) {
		const size_t at = ((size_t) rb * (size_t) (rb - 1) / 2 + (size_t) b) * SG_P + o;
		ss[0] += str[at]; ss[1] += str[str_plane + at]; ss[2] += str[2 * str_plane + at];
	}
	vec4d a;
	a.x = G * (rs[0] - ss[0]); a.y = G * (rs[1] - ss[1]); a.z = G * (rs[2] - ss[2]); a.w = 0.0;
	if (add) { const vec4d p = out[i]; a.x += p.x; a.y += p.y; a.z += p.z; }
	out[i] = a;
}

// Rectangle plans, resident side: a_i = +G * (rows of the CTAs whose resident block holds i), i in [rlo, rhi)
__global__ void k_gravity_rect_reduce_res(const double* __restrict__ res, size_t res_plane, const int* __restrict__ be, int rlo, int rhi, int rpad, int SG_P, double Gm,
		vec4d* __restrict__ out, int add) {
	const int i = rlo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= rhi) return;
	const int b = (i - rlo + rpad) / SG_P, o = (i - rlo + rpad) - b * SG_P;
	double rs[3] = { 0.0, 0.0, 0.0 };
	for (int c = be[2 * b]; c < be[2 * b + 1]; c++) {
		const size_t at = (size_t) c * SG_P + o;
		rs[0] += res[at]; rs[1] += res[res_plane + at]; rs[2] += res[2 * res_plane + at];
	}
	vec4d a;
	a.x = Gm * rs[0]; a.y = Gm * rs[1]; a.z = Gm * rs[2]; a.w = 0.0;
	if (add) { const vec4d p = out[i]; a.x += p.x; a.y += p.y; a.z += p.z; }
	out[i] = a;
}

// Rectangle plans, streamed side: a_j = -G * (rows of column block b over all resident blocks), j in [slo, shi)
__global__ void k_gravity_rect_reduce_str(const double* __restrict__ str, size_t str_plane, int nrb, int nsb, int slo, int shi, int spad, int SG_P, double Gm,
		vec4d* __restrict__ out, int add) {
	const int j = slo + blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= shi) return;
	const int b = (j - slo + spad) / SG_P, o = (j - slo + spad) - b * SG_P;
	double ss[3] = { 0.0, 0.0, 0.0 };
	for (int rb = 0; rb < nrb; rb++) {
		const size_t at = ((size_t) rb * (size_t) nsb + (size_t) b) * SG_P + o;
		ss[0] += str[at]; ss[1] += str[str_plane + at]; ss[2] += str[2 * str_plane + at];
	}
	vec4d a;
	a.x = -Gm * ss[0]; a.y = -Gm * ss[1]; a.z = -Gm * ss[2]; a.w = 0.0;
	if (add) { const vec4d p = out[j]; a.x += p.x; a.y += p.y; a.z += p.z; }
	out[j] = a;
}

// *differs = 1 if any particle of [lo, hi) has a mass whose bits differ from particle ref's
__global__ void k_mass_differs(const vec4d* __restrict__ posm, int lo, int hi, int ref, int* __restrict__ differs) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	if (__double_as_longlong(posm[i].w) != __double_as_longlong(posm[ref].w)) *differs = 1;
}

}  // namespace

struct SymPlan {
	int rlo = -1, rhi = -1, slo = -1, shi = -1;  // resident / streamed ranges (identical: triangle; disjoint: rectangle)
	bool rect = false;
	int nrb = 0, nsb = 0, n_cta = 0;
	int R = 4, P = 1024;        // residents per thread and particles per block of the kernel variant this plan runs
	int rpad = 0, spad = 0;     // dummies in front of block 0 of the resident / streamed range (blocks are counted from lo - pad)
	long long mass_epoch = -1;  // masses the `uniform` verdict was taken from (ass_sim::mass_epoch)
	int uniform = 0;
	double m_ref = 0.0;         // mass of particle rlo (the common mass when uniform)
	SymCta* d_ctas = nullptr;
	int* d_cta_begin = nullptr;
	double *d_res = nullptr, *d_str = nullptr;
	size_t res_plane = 0, str_plane = 0;
};

static void plan_free(SymPlan* p) {
	if (!p) return;
	cudaFree(p->d_ctas); cudaFree(p->d_cta_begin); cudaFree(p->d_res); cudaFree(p->d_str);
	delete p;
}

void ass_sym_plans_free(ass_sim* s) {
	for (SymPlan* p : s->sym_plans) plan_free(p);
	s->sym_plans.clear();
}

// Is the pair-once kernel worth it for this many particles?  It needs enough (possibly cut) tiles to fill the machine:
// a tile (blocks of 1024) can be cut into 8 pieces.
bool ass_gravity_sym_applicable(int count) {
	constexpr int SG_P = SG_T * SG_R_MAX;
	const long long nb = (count + SG_P - 1) / SG_P;
	return nb * (nb + 1) / 2 * (SG_P / SG_TILE) >= 2LL * ass_sm_count();  // (threshold as measured with pieces of 128)
}

static int check_masses(ass_sim* s, SymPlan* plan) {
	int* d_flag = (int*) s->d_red;  // the simulation's small reduction scratch (device), free between launches
	CU_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), s->stream));
	k_mass_differs<<<(plan->rhi - plan->rlo + 255) / 256, 256, 0, s->stream>>>(s->posm, plan->rlo, plan->rhi, plan->rlo, d_flag);
	LAUNCH_CHECK();
	if (plan->rect) {
		k_mass_differs<<<(plan->shi - plan->slo + 255) / 256, 256, 0, s->stream>>>(s->posm, plan->slo, plan->shi, plan->rlo, d_flag);
		LAUNCH_CHECK();
	}
	int differs = 1;
	CU_TRY(cudaMemcpyAsync(&differs, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaMemcpyAsync(&plan->m_ref, &s->posm[plan->rlo].w, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	const char* off = getenv("ASS_GRAVITY_NO_UNIFORM");
	plan->uniform = (!differs && !(off && *off && *off != '0')) ? 1 : 0;
	plan->mass_epoch = s->mass_epoch;
	return ASS_OK;
}

// The host half of a plan: which kernel variant, how the two ranges are cut into blocks, and the list of CTAs.  Pure host
// work (no device is touched: ass_selfcheck_gravity_plan drives it from the CPU test suite).
struct SymCut {
	int R = 4, P = 1024, nrb = 0, nsb = 0, rpad = 0, spad = 0, U = 1, cut = 1;
	std::vector<SymCta> ctas;
	std::vector<int> be;  // be[2b], be[2b+1]: the CTAs whose resident block is b
};

static void build_sym_cut(int n_res, int n_str, bool rect, int sm_count, SymCut& c) {
	// which kernel variant: three residents per thread (blocks of 768, three CTAs per SM) for large problems, four otherwise
	{
		const long long b4r = (n_res + 1023) / 1024, b4s = (n_str + 1023) / 1024;
		const long long tiles4 = rect ? b4r * b4s : b4r * (b4r + 1) / 2;
		c.R = tiles4 >= 2000 ? 3 : 4;
		if (const char* e = getenv("ASS_SYM_R")) { const int r = atoi(e); if (r == 3 || r == 4) c.R = r; }
		c.P = SG_T * c.R;
	}
	const int SG_P = c.P;
	const int nrb = (n_res + SG_P - 1) / SG_P, nsb = (n_str + SG_P - 1) / SG_P;
	c.nrb = nrb; c.nsb = nsb;
	// The range's partial block is block 0, its padding in FRONT: block 0 is the streamed side of all its off-diagonal tiles
	// (triangle: streamed block <= resident block), where tiles of pure padding are skipped 128 particles at a time, and
	// the resident side of only one.  With the partial block last (round 1) it was resident in nb tiles, all its padding
	// computed: a body of 9 327 particles (9.1 blocks) spent 10 of its 55 tiles on a block that is 11 % full.
	c.rpad = nrb * SG_P - n_res;
	c.spad = nsb * SG_P - n_str;
	const long long tiles = rect ? (long long) nrb * nsb : (long long) nrb * (nrb + 1) / 2;
	// tiles per CTA: enough CTAs (>= ~16 per resident slot) for the block scheduler to balance, few enough rows
	const long long slots = (long long) sg_occ(c.R) * sm_count;
	int U = (int) std::max(1LL, tiles / (slots * 16));
	// few tiles (N ~ 1e4): cut every tile's streamed side into `cut` pieces so that ~2 CTAs per slot exist
	// (measured and dropped: the coarsest cut with the shortest makespan in whole rounds of CTAs -- an SM with one
	// resident CTA does not run it twice as fast, so fewer, fatter CTAs lose: C2 gravity 100 -> 116 us)
	int cut = 1;
	while (U == 1 && tiles * cut < 2 * slots && cut < SG_P / SG_TILE) cut *= 2;
	// (Round 2, measured and left off: pieces of 64 streamed particles where they save a whole round of CTAs -- a body of
	// ~1e4 particles is 1.49 rounds of the 296 slots cut in 8 and 2.97 cut in 16.  On a GPU already under sustained FP64 load
	// (power-limited clocks) that gains 7 %: 120.5 -> 111.6 us; on a cool one, as a short benchmark finds it, it loses 10 %:
	// 97 -> 108 us, the per-CTA fixed cost -- resident block, scratch rows -- weighing more at full clocks; a 12 288-particle
	// slab loses either way.  The kernel accepts pieces that end inside a tile; ASS_SYM_CUT forces a cut for measurements.)
	if (const char* e = getenv("ASS_SYM_CUT")) {
		const int f = atoi(e);
		if (U == 1 && f >= 1 && f <= SG_P / 32 && (f & (f - 1)) == 0 && SG_P % f == 0 && (SG_P / f) % 32 == 0) cut = f;
	}
	c.U = U; c.cut = cut;
	const int piece = SG_P / cut;
	std::vector<std::vector<SymCta>> per_block(nrb);
	for (int rb = 0; rb < nrb; rb++) {
		const int s_end = rect ? nsb : rb + 1;  // rectangle: every streamed block; triangle: up to the diagonal
		for (int s0 = 0; s0 < s_end; s0 += U)
			for (int q = 0; q < cut; q++) {
				// a piece of streamed block 0 that lies in its leading padding has nothing to do
				if (U == 1 && s0 == 0 && (q + 1) * piece <= c.spad) continue;
				per_block[rb].push_back({ rb, s0, std::min(s0 + U, s_end), q * piece, (q + 1) * piece });
			}
	}
	// CTAs of one resident block stay contiguous (the reduction walks be[2b] .. be[2b+1]); big blocks first
	c.ctas.clear();
	c.be.assign((size_t) 2 * nrb, 0);
	for (int b = nrb - 1; b >= 0; b--) {
		c.be[2 * b] = (int) c.ctas.size();
		c.ctas.insert(c.ctas.end(), per_block[b].begin(), per_block[b].end());
		c.be[2 * b + 1] = (int) c.ctas.size();
	}
}

// Host-only self-check of the plan for n_res resident and n_str streamed particles (rect == 0: one range against itself, n_str
// ignored): every pair of a resident block and a REAL streamed particle must be covered by exactly one CTA (triangle: streamed
// block <= resident block), CTAs of a resident block contiguous in the list, pieces multiples of 32 inside their block.
// stats: [0] CTAs, [1] residents per thread, [2] cut, [3] tiles per CTA, [4] resident blocks, [5] streamed blocks,
// [6] violations (must be 0), [7] pieces skipped because they lie in the leading padding.
extern "C" int ass_selfcheck_gravity_plan(int n_res, int n_str, int rect, int sm_count, long long* stats) {
	REQUIRE(stats && n_res > 0 && (!rect || n_str > 0) && sm_count > 0, ASS_EINVAL, "bad argument");
	if (!rect) n_str = n_res;
	SymCut c;
	build_sym_cut(n_res, n_str, rect != 0, sm_count, c);
	long long viol = 0;
	const int P = c.P;
	// coverage counts per (resident block, streamed block, chunk of 32)
	std::vector<unsigned char> cover((size_t) c.nrb * c.nsb * (P / 32), 0);
	for (size_t i = 0; i < c.ctas.size(); i++) {
		const SymCta& q = c.ctas[i];
		if (q.rb < 0 || q.rb >= c.nrb || q.s0 < 0 || q.s1 <= q.s0 || q.s1 > c.nsb || q.t_lo < 0 || q.t_hi > P || q.t_lo >= q.t_hi || (q.t_lo & 31) || (q.t_hi & 31)) { viol++; continue; }
		if (!rect && q.s1 > q.rb + 1) viol++;
		if ((int) i < c.be[2 * q.rb] || (int) i >= c.be[2 * q.rb + 1]) viol++;
		for (int sb = q.s0; sb < q.s1; sb++)
			for (int t = q.t_lo; t < q.t_hi; t += 32) cover[((size_t) q.rb * c.nsb + sb) * (P / 32) + t / 32]++;
	}
	long long expected_ctas = 0, skipped = 0;
	for (int rb = 0; rb < c.nrb; rb++)
		for (int sb = 0; sb < c.nsb; sb++) {
			const bool in_domain = rect || sb <= rb;
			for (int t = 0; t < P; t += 32) {
				const int cnt = cover[((size_t) rb * c.nsb + sb) * (P / 32) + t / 32];
				const bool real = (long long) sb * P + t + 32 > c.spad;  // the chunk holds at least one real streamed particle
				if (!in_domain) { if (cnt) viol++; continue; }
				if (real ? cnt != 1 : cnt > 1) viol++;
			}
		}
	for (int b = 0; b < c.nrb; b++) {
		if (c.be[2 * b] > c.be[2 * b + 1]) viol++;
		expected_ctas += c.be[2 * b + 1] - c.be[2 * b];
	}
	if (expected_ctas != (long long) c.ctas.size()) viol++;
	if (c.rpad < 0 || c.rpad >= P || c.spad < 0 || c.spad >= P) viol++;
	{
		const long long full = (long long) (rect ? (long long) c.nrb * c.nsb : (long long) c.nrb * (c.nrb + 1) / 2);
		skipped = c.U == 1 ? full * c.cut - (long long) c.ctas.size() : 0;
	}
	stats[0] = (long long) c.ctas.size(); stats[1] = c.R; stats[2] = c.cut; stats[3] = c.U; stats[4] = c.nrb; stats[5] = c.nsb; stats[6] = viol; stats[7] = skipped;
	return ASS_OK;
}

static int get_plan(ass_sim* s, int rlo, int rhi, int slo, int shi, SymPlan** out) {
	for (SymPlan* q : s->sym_plans)
		if (q->rlo == rlo && q->rhi == rhi && q->slo == slo && q->shi == shi) {
			*out = q;
			return (q->mass_epoch != s->mass_epoch) ? check_masses(s, q) : ASS_OK;
		}
	SymPlan* plan = new SymPlan();
	plan->rlo = rlo; plan->rhi = rhi; plan->slo = slo; plan->shi = shi;
	plan->rect = !(rlo == slo && rhi == shi);
	SymCut cutp;
	build_sym_cut(rhi - rlo, shi - slo, plan->rect, ass_sm_count(), cutp);
	plan->R = cutp.R; plan->P = cutp.P; plan->nrb = cutp.nrb; plan->nsb = cutp.nsb; plan->rpad = cutp.rpad; plan->spad = cutp.spad;
	const int SG_P = plan->P;
	const int nrb = plan->nrb, nsb = plan->nsb;
	const std::vector<SymCta>& ctas = cutp.ctas;
	const std::vector<int>& be = cutp.be;
	plan->n_cta = (int) ctas.size();
	const size_t n_str_tiles = plan->rect ? (size_t) nrb * (size_t) nsb : (size_t) nrb * (size_t) (nrb - 1) / 2;
	plan->res_plane = (size_t) plan->n_cta * SG_P;
	plan->str_plane = std::max<size_t>(n_str_tiles, 1) * SG_P;
	if (cudaMalloc(&plan->d_ctas, sizeof(SymCta) * std::max<size_t>(ctas.size(), 1)) != cudaSuccess ||
	    cudaMalloc(&plan->d_cta_begin, sizeof(int) * std::max<size_t>(be.size(), 1)) != cudaSuccess ||
	    cudaMalloc(&plan->d_res, sizeof(double) * 3 * std::max<size_t>(plan->res_plane, 1)) != cudaSuccess ||
	    cudaMalloc(&plan->d_str, sizeof(double) * 3 * plan->str_plane) != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("pair-once gravity: out of device memory for %zu scratch bytes", sizeof(double) * 3 * (plan->res_plane + plan->str_plane));
		plan_free(plan);
		return ASS_ENOMEM;
	}
	CU_TRY(cudaMemcpyAsync(plan->d_ctas, ctas.data(), sizeof(SymCta) * ctas.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(plan->d_cta_begin, be.data(), sizeof(int) * be.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaStreamSynchronize(s->stream));
	s->sym_plans.push_back(plan);
	*out = plan;
	return check_masses(s, plan);
}

// Pair-once gravity between resident particles [rlo, rhi) and streamed particles [slo, shi) of s->posm.
//   identical ranges : every unordered pair of the range once; res_out[i] (=|+=) a_i for i in the range; str_out unused
//   disjoint ranges  : every (resident, streamed) pair once; res_out[i] (=|+=) the pull of the streamed range on resident i,
//                      str_out[j] (=|+=) the pull of the resident range on streamed j (Newton's third law)
int ass_launch_gravity_sym_pair(ass_sim* s, int rlo, int rhi, int slo, int shi, vec4d* res_out, bool res_add, vec4d* str_out, bool str_add) {
	if (rhi <= rlo || shi <= slo) return ASS_OK;
	ass_tic(s, ASS_K_GRAVITY);
	int rc = ass_launch_gravity_sym_main(s, rlo, rhi, slo, shi, s->stream);
	if (!rc) rc = ass_launch_gravity_sym_finish(s, rlo, rhi, slo, shi, res_out, res_add, str_out, str_add);
	if (rc) return rc;
	ass_toc(s, ASS_K_GRAVITY);
	return ASS_OK;
}

// The two halves of the above, for callers that run several range pairs at once (shard.cu: a rank's triangle, rectangles and
// antipodal half on separate streams, so that the last, partly filled round of CTAs of one is topped up with the next one's):
// the pair kernel into the plan's own scratch rows on ANY stream, and -- on the simulation's stream, after the caller has
// made it wait for that kernel -- the fixed-order reductions into the output arrays.
int ass_launch_gravity_sym_main(ass_sim* s, int rlo, int rhi, int slo, int shi, cudaStream_t st) {
	if (rhi <= rlo || shi <= slo) return ASS_OK;
	SymPlan* p = nullptr;
	int rc = get_plan(s, rlo, rhi, slo, shi, &p);
	if (rc) return rc;
	const double soft2 = s->prm.softening * s->prm.softening;
#define ASS_SYM_LAUNCH(U, R)                                                                                                                     \
	k_gravity_sym<U, R><<<p->n_cta, SG_T, 0, st>>>(s->posm, rlo, rhi, slo, shi, p->rpad, p->spad, p->rect ? 1 : 0, p->nsb, p->d_ctas, soft2, p->d_res, \
			p->res_plane, p->d_str, p->str_plane)
	if (p->R == 3) { if (p->uniform) ASS_SYM_LAUNCH(true, 3); else ASS_SYM_LAUNCH(false, 3); }
	else { if (p->uniform) ASS_SYM_LAUNCH(true, 4); else ASS_SYM_LAUNCH(false, 4); }
#undef ASS_SYM_LAUNCH
	LAUNCH_CHECK();
	return ASS_OK;
}

int ass_launch_gravity_sym_finish(ass_sim* s, int rlo, int rhi, int slo, int shi, vec4d* res_out, bool res_add, vec4d* str_out, bool str_add) {
	if (rhi <= rlo || shi <= slo) return ASS_OK;
	SymPlan* p = nullptr;
	int rc = get_plan(s, rlo, rhi, slo, shi, &p);
	if (rc) return rc;
	if (!p->rect) {
		k_gravity_sym_reduce<<<(rhi - rlo + 255) / 256, 256, 0, s->stream>>>(p->d_res, p->res_plane, p->d_str, p->str_plane, p->d_cta_begin, p->nrb, rlo, rhi,
				p->rpad, p->P, s->prm.G, p->uniform, s->posm, res_out, res_add ? 1 : 0);
		LAUNCH_CHECK();
	} else {
		// uniform: the rows are unweighted sums and every particle of both ranges has the mass of particle rlo
		const double Gm = p->uniform ? s->prm.G * p->m_ref : s->prm.G;
		k_gravity_rect_reduce_res<<<(rhi - rlo + 255) / 256, 256, 0, s->stream>>>(p->d_res, p->res_plane, p->d_cta_begin, rlo, rhi, p->rpad, p->P, Gm, res_out, res_add ? 1 : 0);
		LAUNCH_CHECK();
		k_gravity_rect_reduce_str<<<(shi - slo + 255) / 256, 256, 0, s->stream>>>(p->d_str, p->str_plane, p->nrb, p->nsb, slo, shi, p->spad, p->P, Gm, str_out, str_add ? 1 : 0);
		LAUNCH_CHECK();
	}
	return ASS_OK;
}

int ass_gravity_sym_prepare(ass_sim* s, int rlo, int rhi, int slo, int shi) {
	if (rhi <= rlo || shi <= slo) return ASS_OK;
	SymPlan* p = nullptr;
	return get_plan(s, rlo, rhi, slo, shi, &p);
}

// a = gravity among particles [lo, hi) (every particle both source and target), pair-once
int ass_launch_gravity_sym(ass_sim* s, int lo, int hi) { return ass_launch_gravity_sym_pair(s, lo, hi, lo, hi, s->acc, false, nullptr, false); }
