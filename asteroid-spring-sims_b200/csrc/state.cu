// state.cu -- simulation lifecycle, host<->device state transfer, spring CSR construction, timers.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <unordered_map>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------------
// errors / globals
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
long long g_launch_count = 0;
static int g_device = -1;
static int g_sm_count = 0;

void ass_set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char* ass_last_error(void) { return g_err; }
extern "C" int ass_api_version(void) { return ASS_API_VERSION; }
extern "C" long long ass_launch_count(void) { return g_launch_count; }
extern "C" void ass_free(void* p) { free(p); }

extern "C" int ass_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" int ass_init(int device) {
	int n = ass_device_count();
	if (n <= 0) {
		ass_set_error("ass_init: no CUDA device visible (this library has no CPU fallback)");
		return ASS_ENODEVICE;
	}
	REQUIRE(device >= 0 && device < n, ASS_EINVAL, "device index out of range");
	cudaDeviceProp prop;
	CU_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		ass_set_error("ass_init: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
		return ASS_ENODEVICE;
	}
	CU_TRY(cudaSetDevice(device));
	CU_TRY(cudaFree(0));
	g_device = device;
	g_sm_count = prop.multiProcessorCount;
	return ASS_OK;
}

int ass_ensure_device(void) {
	if (g_device >= 0) {
		// one process per GPU, but the caller may use several host threads (the reference's OPENGL build runs
		// reb_integrate on a pthread, src/rebound.c:676-689): bind the device on whichever thread calls in.
		cudaError_t e = cudaSetDevice(g_device);
		if (e != cudaSuccess) {
			ass_set_error("cudaSetDevice(%d) -> %s", g_device, cudaGetErrorString(e));
			return ASS_ECUDA;
		}
		return ASS_OK;
	}
	return ass_init(0);
}

int ass_sm_count(void) { return g_sm_count > 0 ? g_sm_count : 148; }

// does the current device launch kernels cooperatively (all CTAs resident, grid-wide barriers allowed)?
bool ass_cooperative_launch_ok(void) {
	static int ok[64];  // 0 unknown, 1 yes, 2 no
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
	if (!ok[dev]) {
		int v = 0;
		ok[dev] = (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && v) ? 1 : 2;
	}
	return ok[dev] == 1;
}

extern "C" int ass_device_info(int* sm_count, long long* l2_bytes, long long* mem_bytes, int* sm_clock_khz, int* cc) {
	int rc = ass_ensure_device();
	if (rc) return rc;
	cudaDeviceProp prop;
	CU_TRY(cudaGetDeviceProperties(&prop, g_device));
	if (sm_count) *sm_count = prop.multiProcessorCount;
	if (l2_bytes) *l2_bytes = prop.l2CacheSize;
	if (mem_bytes) *mem_bytes = (long long) prop.totalGlobalMem;
	if (sm_clock_khz) {
		int khz = 0;
		cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device);
		*sm_clock_khz = khz;
	}
	if (cc) *cc = prop.major * 10 + prop.minor;
	return ASS_OK;
}

int ass_buf_reserve(DevBuf& b, size_t bytes) {
	if (bytes <= b.bytes) return ASS_OK;
	if (b.p) cudaFree(b.p);
	b.p = nullptr;
	b.bytes = 0;
	size_t want = bytes + bytes / 4 + 256;
	cudaError_t e = cudaMalloc(&b.p, want);
	if (e != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
		return ASS_ENOMEM;
	}
	b.bytes = want;
	return ASS_OK;
}

// ---------------------------------------------------------------------------------------------------
// pinned host ranges
//
// `struct reb_particle[]` is ordinary malloc'd memory (src/particle.c:53-56).  A plain cudaMemcpy from it is staged by
// the driver through its own bounce buffer (measured on this pool: 0.68 ms for the 12.4 MB of C3, 11.2 ms for 128 MB),
// and a download has to be scattered into the caller's rows by a host loop (0.69 / 15.2 ms).  Once the caller has
// page-locked and mapped the range (ass_host_register for memory it already owns, ass_host_alloc for new memory) the
// copy engine reads the rows directly (0.25 / 2.4 ms = the PCIe Gen5 x16 link rate) and the download kernel writes
// the requested doubles straight into the caller's rows over the link (0.23 / 2.4 ms), no bounce and no host pass.
// tools/xfer_probe.cu holds the measurement that picked these two mechanisms.
// ---------------------------------------------------------------------------------------------------
#include <mutex>
namespace {
struct PinnedRange {
	unsigned char* base;
	size_t bytes;
	unsigned char* dev;  // device-side alias of `base`
	bool owned;          // allocated by ass_host_alloc (freed by ass_host_free) rather than registered
};
std::mutex g_pin_mu;
std::vector<PinnedRange> g_pinned;

// device alias of [p, p + bytes) if it lies inside a pinned range, else nullptr
unsigned char* pinned_alias(const void* p, size_t bytes) {
	std::lock_guard<std::mutex> lk(g_pin_mu);
	const unsigned char* q = (const unsigned char*) p;
	for (const PinnedRange& r : g_pinned)
		if (q >= r.base && q + bytes <= r.base + r.bytes) return r.dev + (q - r.base);
	return nullptr;
}
}  // namespace

extern "C" int ass_host_register(void* base, size_t bytes) {
	REQUIRE(base && bytes > 0, ASS_EINVAL, "null or empty range");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (pinned_alias(base, bytes)) return ASS_OK;  // already covered
	CU_TRY(cudaHostRegister(base, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
	void* dev = nullptr;
	cudaError_t e = cudaHostGetDevicePointer(&dev, base, 0);
	if (e != cudaSuccess) {
		cudaHostUnregister(base);
		ass_set_error("cudaHostGetDevicePointer -> %s", cudaGetErrorString(e));
		return ASS_ECUDA;
	}
	std::lock_guard<std::mutex> lk(g_pin_mu);
	g_pinned.push_back({ (unsigned char*) base, bytes, (unsigned char*) dev, false });
	return ASS_OK;
}

extern "C" int ass_host_unregister(void* base) {
	PinnedRange found{ nullptr, 0, nullptr, false };
	{
		std::lock_guard<std::mutex> lk(g_pin_mu);
		for (size_t i = 0; i < g_pinned.size(); i++)
			if (g_pinned[i].base == (unsigned char*) base && !g_pinned[i].owned) {
				found = g_pinned[i];
				g_pinned.erase(g_pinned.begin() + (long) i);
				break;
			}
	}
	REQUIRE(found.base, ASS_EINVAL, "range was not registered with ass_host_register");
	CU_TRY(cudaHostUnregister(base));
	return ASS_OK;
}

extern "C" int ass_host_alloc(void** out, size_t bytes) {
	REQUIRE(out && bytes > 0, ASS_EINVAL, "null result pointer or empty allocation");
	int rc = ass_ensure_device();
	if (rc) return rc;
	void *p = nullptr, *dev = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("cudaHostAlloc(%zu) failed", bytes);
		return ASS_ENOMEM;
	}
	CU_TRY(cudaHostGetDevicePointer(&dev, p, 0));
	std::lock_guard<std::mutex> lk(g_pin_mu);
	g_pinned.push_back({ (unsigned char*) p, bytes, (unsigned char*) dev, true });
	*out = p;
	return ASS_OK;
}

extern "C" int ass_host_free(void* p) {
	if (!p) return ASS_OK;
	bool found = false;
	{
		std::lock_guard<std::mutex> lk(g_pin_mu);
		for (size_t i = 0; i < g_pinned.size(); i++)
			if (g_pinned[i].base == (unsigned char*) p && g_pinned[i].owned) {
				g_pinned.erase(g_pinned.begin() + (long) i);
				found = true;
				break;
			}
	}
	REQUIRE(found, ASS_EINVAL, "pointer did not come from ass_host_alloc");
	CU_TRY(cudaFreeHost(p));
	return ASS_OK;
}

// ---------------------------------------------------------------------------------------------------
// timing
// ---------------------------------------------------------------------------------------------------
void ass_tic(ass_sim* s, int klass) {
	if (!s->timing) return;
	int idx = (int) s->ev_open.size();
	if (idx >= (int) s->ev_pool.size()) {
		cudaEvent_t a, b;
		cudaEventCreate(&a);
		cudaEventCreate(&b);
		s->ev_pool.push_back({ a, b });
	}
	cudaEventRecord(s->ev_pool[idx].first, s->stream);
	s->ev_open.push_back({ klass, idx });
}
void ass_toc(ass_sim* s, int klass) {
	(void) klass;
	if (!s->timing) return;
	cudaEventRecord(s->ev_pool[s->ev_open.back().second].second, s->stream);
}
static int timing_drain(ass_sim* s) {
	if (s->ev_open.empty()) return ASS_OK;
	CU_TRY(cudaStreamSynchronize(s->stream));
	for (auto& o : s->ev_open) {
		float ms = 0.f;
		// a bracket whose closing event was never recorded (early error return) is dropped
		if (cudaEventElapsedTime(&ms, s->ev_pool[o.second].first, s->ev_pool[o.second].second) != cudaSuccess) {
			cudaGetLastError();
			continue;
		}
		s->k_ms[o.first] += ms;
		s->k_launches[o.first]++;
	}
	s->ev_open.clear();
	return ASS_OK;
}
extern "C" int ass_timing_enable(ass_sim* s, int on) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	int rc = timing_drain(s);
	s->timing = on != 0;
	return rc;
}
extern "C" int ass_timing_reset(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	int rc = timing_drain(s);
	for (int i = 0; i < ASS_K_COUNT; i++) {
		s->k_ms[i] = 0;
		s->k_launches[i] = 0;
	}
	return rc;
}
extern "C" int ass_timing_get(ass_sim* s, double ms[ASS_K_COUNT], long long launches[ASS_K_COUNT]) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	int rc = timing_drain(s);
	if (rc) return rc;
	for (int i = 0; i < ASS_K_COUNT; i++) {
		if (ms) ms[i] = s->k_ms[i];
		if (launches) launches[i] = s->k_launches[i];
	}
	return ASS_OK;
}
extern "C" int ass_mark_begin(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	CU_TRY(cudaEventRecord(s->mark0, s->stream));
	return ASS_OK;
}
extern "C" int ass_mark_end(ass_sim* s, double* ms) {
	REQUIRE(s && ms, ASS_EINVAL, "null argument");
	CU_TRY(cudaEventRecord(s->mark1, s->stream));
	CU_TRY(cudaEventSynchronize(s->mark1));
	float f = 0.f;
	CU_TRY(cudaEventElapsedTime(&f, s->mark0, s->mark1));
	*ms = f;
	return ASS_OK;
}

// Lap timer: event pairs on the simulation's stream, recorded without waiting for anything, summed when collected.  For
// loops whose iterations must stay queued back to back (a sharded body's steps: a host that waited for every step
// would be in the loop it is supposed to stay out of) but whose timed region excludes something between the iterations
// (the L2 flush).
extern "C" int ass_lap_begin(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	if (s->lap_open >= (int) s->lap_ev.size()) {
		cudaEvent_t a, b;
		CU_TRY(cudaEventCreate(&a));
		CU_TRY(cudaEventCreate(&b));
		s->lap_ev.push_back({ a, b });
	}
	CU_TRY(cudaEventRecord(s->lap_ev[(size_t) s->lap_open].first, s->stream));
	return ASS_OK;
}
extern "C" int ass_lap_end(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->lap_open < (int) s->lap_ev.size(), ASS_ESTATE, "ass_lap_end without ass_lap_begin");
	CU_TRY(cudaEventRecord(s->lap_ev[(size_t) s->lap_open].second, s->stream));
	s->lap_open++;
	return ASS_OK;
}
extern "C" int ass_lap_collect(ass_sim* s, double* total_ms, int* laps) {
	REQUIRE(s && total_ms, ASS_EINVAL, "null argument");
	int rc = ass_sync(s);
	double tot = 0.0;
	for (int i = 0; i < s->lap_open; i++) {
		float ms = 0.f;
		CU_TRY(cudaEventElapsedTime(&ms, s->lap_ev[(size_t) i].first, s->lap_ev[(size_t) i].second));
		tot += ms;
	}
	*total_ms = tot;
	if (laps) *laps = s->lap_open;
	s->lap_open = 0;
	return rc;
}

__global__ void k_fill(double* p, size_t n, double v) {
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t) gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}
extern "C" int ass_l2_flush(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	const size_t bytes = (size_t) 256 << 20;  // > 126 MB L2
	int rc = ass_buf_reserve(s->flush, bytes);
	if (rc) return rc;
	k_fill<<<ass_sm_count() * 8, 256, 0, s->stream>>>((double*) s->flush.p, bytes / sizeof(double), 0.0);
	LAUNCH_CHECK();
	return ASS_OK;
}

// ---------------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------------
extern "C" int ass_create(ass_sim** out) {
	REQUIRE(out, ASS_EINVAL, "null out pointer");
	int rc = ass_ensure_device();
	if (rc) return rc;
	ass_sim* s = new ass_sim();
	memset(&s->prm, 0, sizeof(s->prm));
	s->prm.dt = 0.001;  // reb_init_simulation defaults, src/rebound.c:372-376
	s->prm.G = 1.0;
	s->prm.n_active = -1;
	s->prm.exact_finish_time = 1;
	s->prm.status = -1;
	s->prm.additional_forces = 1;
	for (int i = 0; i < ASS_K_COUNT; i++) {
		s->k_ms[i] = 0;
		s->k_launches[i] = 0;
	}
	if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreate(&s->mark0) != cudaSuccess || cudaEventCreate(&s->mark1) != cudaSuccess ||
	    cudaMalloc(&s->d_red, 64 * sizeof(double)) != cudaSuccess ||
	    cudaMallocHost(&s->h_red, 64 * sizeof(double)) != cudaSuccess) {
		ass_set_error("ass_create: %s", cudaGetErrorString(cudaGetLastError()));
		delete s;
		return ASS_ECUDA;
	}
	*out = s;
	return ASS_OK;
}

extern "C" int ass_destroy(ass_sim* s) {
	if (!s) return ASS_OK;
	ass_ensure_device();
	if (s->stream) cudaStreamSynchronize(s->stream);
	ass_shard_release(s);  // a sharded simulation's posm / vel alias its exported window
	ass_sym_plans_free(s);
	cudaFree(s->slotT); cudaFree(s->warp_off);
	cudaFree(s->heL); cudaFree(s->grpT); cudaFree(s->tile_tab); cudaFree(s->tile_desc); cudaFree(s->tile_sl_ord);
	cudaFree(s->order); cudaFree(s->rank); cudaFree(s->mir_block); cudaFree(s->mir_m);
	cudaFree(s->posm); cudaFree(s->vel); cudaFree(s->acc); cudaFree(s->posm_alt); cudaFree(s->vel_alt);
	cudaFree(s->he); cudaFree(s->row_ptr); cudaFree(s->slot); cudaFree(s->pal);
	cudaFree(s->stage.p); cudaFree(s->grav_partial.p); cudaFree(s->red_partial.p); cudaFree(s->flush.p);
	cudaFree(s->d_red); cudaFree(s->masks); cudaFree(s->d_step_bar);
	if (s->h_red) cudaFreeHost(s->h_red);
	if (s->h_pinned) cudaFreeHost(s->h_pinned);
	for (auto& e : s->ev_pool) {
		cudaEventDestroy(e.first);
		cudaEventDestroy(e.second);
	}
	for (auto& e : s->lap_ev) {
		cudaEventDestroy(e.first);
		cudaEventDestroy(e.second);
	}
	if (s->mark0) cudaEventDestroy(s->mark0);
	if (s->mark1) cudaEventDestroy(s->mark1);
	if (s->stream) cudaStreamDestroy(s->stream);
	cudaGetLastError();
	delete s;
	return ASS_OK;
}

extern "C" int ass_sync(ass_sim* s) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	if (s->shard.active) return ass_shard_sync(s);  // both streams, and the verdict of the device-side waits
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}

extern "C" int ass_set_params(ass_sim* s, const ass_params* p) {
	REQUIRE(s && p, ASS_EINVAL, "null argument");
	REQUIRE(p->gravity == 0 || p->gravity == 1, ASS_EINVAL, "gravity must be 0 (NONE) or 1 (BASIC)");
	REQUIRE(p->additional_forces >= 0 && p->additional_forces <= 2, ASS_EINVAL, "additional_forces must be 0, 1 or 2");
	REQUIRE(p->heartbeat == 0 || p->heartbeat == 1, ASS_EINVAL, "heartbeat must be 0 or 1");
	// positions that already carry the next step's half-drift (ass_shard_finish(pre_drift_next = 1)) were drifted with
	// the OLD dt: changing it now would silently mix two step sizes
	REQUIRE(!(s->predrifted && p->dt != s->prm.dt), ASS_ESTATE, "dt cannot change while positions are pre-drifted: finish a step without pre_drift_next first");
	s->prm = *p;
	return ASS_OK;
}
extern "C" int ass_get_params(ass_sim* s, ass_params* p) {
	REQUIRE(s && p, ASS_EINVAL, "null argument");
	*p = s->prm;
	return ASS_OK;
}
extern "C" int ass_num_particles(ass_sim* s) { return s ? s->n : ASS_EINVAL; }
extern "C" int ass_num_springs(ass_sim* s) { return s ? s->ns : ASS_EINVAL; }

// ---------------------------------------------------------------------------------------------------
// particles: host rows <-> device vectors
// ---------------------------------------------------------------------------------------------------
static int ensure_particles(ass_sim* s, int n) {
	if (n <= s->cap) return ASS_OK;
	int cap = std::max(n + n / 8, 1024);
	vec4d* nb[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
	for (int i = 0; i < 5; i++) {
		if (cudaMalloc(&nb[i], sizeof(vec4d) * (size_t) cap) != cudaSuccess) {
			cudaGetLastError();
			for (int j = 0; j < i; j++) cudaFree(nb[j]);
			ass_set_error("out of device memory for %d particles", cap);
			return ASS_ENOMEM;
		}
		cudaMemsetAsync(nb[i], 0, sizeof(vec4d) * (size_t) cap, s->stream);
	}
	vec4d** old[5] = { &s->posm, &s->vel, &s->acc, &s->posm_alt, &s->vel_alt };
	for (int i = 0; i < 5; i++) {
		if (*old[i] && s->n > 0 && i < 3)
			cudaMemcpyAsync(nb[i], *old[i], sizeof(vec4d) * (size_t) s->n, cudaMemcpyDeviceToDevice, s->stream);
	}
	cudaStreamSynchronize(s->stream);
	for (int i = 0; i < 5; i++) {
		cudaFree(*old[i]);
		*old[i] = nb[i];
	}
	s->cap = cap;
	return ASS_OK;
}

__global__ void k_unpack_rows(const unsigned char* __restrict__ rows, size_t stride, int n, unsigned fields,
		vec4d* __restrict__ posm, vec4d* __restrict__ vel, vec4d* __restrict__ acc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double* r = (const double*) (rows + (size_t) i * stride);
	if (fields & (ASS_F_POS | ASS_F_MASS)) {
		vec4d p = posm[i];
		if (fields & ASS_F_POS) { p.x = r[0]; p.y = r[1]; p.z = r[2]; }
		if (fields & ASS_F_MASS) p.w = r[9];
		posm[i] = p;
	}
	if (fields & ASS_F_VEL) { vec4d v; v.x = r[3]; v.y = r[4]; v.z = r[5]; v.w = 0.0; vel[i] = v; }
	if (fields & ASS_F_ACC) { vec4d a; a.x = r[6]; a.y = r[7]; a.z = r[8]; a.w = 0.0; acc[i] = a; }
}

__global__ void k_pack_rows(unsigned char* __restrict__ rows, size_t stride, int n, unsigned fields,
		const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const vec4d* __restrict__ acc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double* r = (double*) (rows + (size_t) i * stride);
	if (fields & (ASS_F_POS | ASS_F_MASS)) {
		vec4d p = posm[i];
		if (fields & ASS_F_POS) { r[0] = p.x; r[1] = p.y; r[2] = p.z; }
		if (fields & ASS_F_MASS) r[9] = p.w;
	}
	if (fields & ASS_F_VEL) { vec4d v = vel[i]; r[3] = v.x; r[4] = v.y; r[5] = v.z; }
	if (fields & ASS_F_ACC) { vec4d a = acc[i]; r[6] = a.x; r[7] = a.y; r[8] = a.z; }
}

// Download into page-locked, mapped host rows: 16 lanes per row, one double per lane, so a row's requested doubles
// leave the SM as one contiguous run (72 bytes for x..az) and cross the link as posted writes.  Untouched doubles of
// the row (r, lastcollision, pointers, hash) are never written.
__global__ void k_pack_rows_direct(unsigned char* __restrict__ rows, size_t stride, int n, unsigned fields,
		const vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const vec4d* __restrict__ acc) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	const int i = (int) (t >> 4), c = (int) (t & 15);
	if (i >= n || c > 9) return;
	const double* src;
	if (c < 3) { if (!(fields & ASS_F_POS)) return; src = (const double*) (posm + i) + c; }
	else if (c < 6) { if (!(fields & ASS_F_VEL)) return; src = (const double*) (vel + i) + (c - 3); }
	else if (c < 9) { if (!(fields & ASS_F_ACC)) return; src = (const double*) (acc + i) + (c - 6); }
	else { if (!(fields & ASS_F_MASS)) return; src = (const double*) (posm + i) + 3; }
	((double*) (rows + (size_t) i * stride))[c] = *src;
}

static int ensure_pinned(ass_sim* s, size_t bytes) {
	if (bytes <= s->h_pinned_bytes) return ASS_OK;
	if (s->h_pinned) cudaFreeHost(s->h_pinned);
	s->h_pinned = nullptr;
	s->h_pinned_bytes = 0;
	size_t want = bytes + bytes / 4 + 4096;
	if (cudaMallocHost(&s->h_pinned, want) != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("cudaMallocHost(%zu) failed", want);
		return ASS_ENOMEM;
	}
	s->h_pinned_bytes = want;
	return ASS_OK;
}

// rows [lo, hi) of the caller's array <-> particles [lo, hi) of the device state
static int upload_rows(ass_sim* s, const void* base, size_t stride, int lo, int hi, unsigned fields) {
	const int cnt = hi - lo;
	if (cnt <= 0) return ASS_OK;
	const size_t bytes = (size_t) cnt * stride;
	int rc = ass_buf_reserve(s->stage, bytes);
	if (rc) return rc;
	ass_tic(s, ASS_K_XFER);
	// One plain copy of the raw rows.  From a range pinned through ass_host_register / ass_host_alloc the copy engine
	// reads the caller's rows at link rate; from pageable memory (the default: realloc'd by reb_add,
	// src/particle.c:53-56) the driver stages it.
	CU_TRY(cudaMemcpyAsync(s->stage.p, (const unsigned char*) base + (size_t) lo * stride, bytes, cudaMemcpyHostToDevice, s->stream));
	k_unpack_rows<<<(cnt + 255) / 256, 256, 0, s->stream>>>((const unsigned char*) s->stage.p, stride, cnt, fields, s->posm + lo, s->vel + lo, s->acc + lo);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_XFER);
	// pageable source: the copy has been staged when cudaMemcpyAsync returns, but keep the contract simple
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}

static int download_rows(ass_sim* s, void* base, size_t stride, int lo, int hi, unsigned fields) {
	const int cnt = hi - lo;
	if (cnt <= 0) return ASS_OK;
	int rc;
	unsigned char* first = (unsigned char*) base + (size_t) lo * stride;
	// Only the requested doubles may change on the host (the rows also hold pointers, hash, r, ...).
	if (unsigned char* alias = pinned_alias(first, (size_t) (cnt - 1) * stride + 10 * sizeof(double))) {
		ass_tic(s, ASS_K_XFER);
		const long long threads = 16LL * cnt;
		k_pack_rows_direct<<<(unsigned) ((threads + 255) / 256), 256, 0, s->stream>>>(alias, stride, cnt, fields, s->posm + lo, s->vel + lo, s->acc + lo);
		LAUNCH_CHECK();
		ass_toc(s, ASS_K_XFER);
		CU_TRY(cudaStreamSynchronize(s->stream));
		return ASS_OK;
	}
	// pageable rows: pack the fields densely on the device, copy that, and scatter into the caller's rows here
	const size_t dense = (size_t) cnt * 10 * sizeof(double);
	if ((rc = ass_buf_reserve(s->stage, dense))) return rc;
	if ((rc = ensure_pinned(s, dense))) return rc;
	ass_tic(s, ASS_K_XFER);
	k_pack_rows<<<(cnt + 255) / 256, 256, 0, s->stream>>>((unsigned char*) s->stage.p, 10 * sizeof(double), cnt, fields, s->posm + lo, s->vel + lo, s->acc + lo);
	LAUNCH_CHECK();
	CU_TRY(cudaMemcpyAsync(s->h_pinned, s->stage.p, dense, cudaMemcpyDeviceToHost, s->stream));
	ass_toc(s, ASS_K_XFER);
	CU_TRY(cudaStreamSynchronize(s->stream));
	const double* src = (const double*) s->h_pinned;
	for (int i = 0; i < cnt; i++) {
		double* r = (double*) (first + (size_t) i * stride);
		const double* q = src + (size_t) i * 10;
		if (fields & ASS_F_POS) { r[0] = q[0]; r[1] = q[1]; r[2] = q[2]; }
		if (fields & ASS_F_VEL) { r[3] = q[3]; r[4] = q[4]; r[5] = q[5]; }
		if (fields & ASS_F_ACC) { r[6] = q[6]; r[7] = q[7]; r[8] = q[8]; }
		if (fields & ASS_F_MASS) r[9] = q[9];
	}
	return ASS_OK;
}

extern "C" int ass_upload_particles(ass_sim* s, const void* base, size_t stride, int n, unsigned fields) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(n >= 0, ASS_EINVAL, "negative particle count");
	REQUIRE(n == 0 || base, ASS_EINVAL, "null particle array");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "stride must hold the 10 leading doubles and be 8-byte aligned");
	REQUIRE((fields & ~ASS_F_ALL) == 0 && fields != 0, ASS_EINVAL, "bad field mask");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (n != s->n) {
		// a different particle count is a new body: everything must come from the host
		REQUIRE(fields == ASS_F_ALL, ASS_ESTATE, "particle count changed: upload ASS_F_ALL");
		REQUIRE(!s->shard.active, ASS_ESTATE, "a sharded simulation keeps its particle count");
		if ((rc = ensure_particles(s, n))) return rc;
		if (n != s->n) { s->ns = 0; s->n_he = 0; s->tiles_grid = 0; ass_sym_plans_free(s); }  // springs and gravity plans refer to the old body
		s->n = n;
	}
	// a pre-drifted body (ass_shard_step(..., pre_drift_last = 1)) may only be REPLACED: patching positions or velocities
	// into a half-drifted state would be applied to the wrong instant
	REQUIRE(!s->predrifted || (fields & (ASS_F_POS | ASS_F_VEL)) == (ASS_F_POS | ASS_F_VEL), ASS_ESTATE,
	        "positions are pre-drifted: upload positions and velocities together, or finish a step without pre_drift_last first");
	s->predrifted = false;
	if (fields & ASS_F_MASS) s->mass_epoch++;
	if ((rc = upload_rows(s, base, stride, 0, n, fields))) return rc;
	s->have_state = true;
	return ASS_OK;
}

// ass_upload_particles for rows [i_low, i_high) only (`base` is row 0 of the caller's array): a rank of a sharded body
// refreshing its own slab, or a module that moved a few particles on the host.
extern "C" int ass_upload_range(ass_sim* s, const void* base, size_t stride, int i_low, int i_high, unsigned fields) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->have_state, ASS_ESTATE, "upload the whole body first");
	REQUIRE(i_low >= 0 && i_high <= s->n && i_low <= i_high, ASS_EINVAL, "bad particle range");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "stride must hold the 10 leading doubles and be 8-byte aligned");
	REQUIRE((fields & ~ASS_F_ALL) == 0 && fields != 0, ASS_EINVAL, "bad field mask");
	REQUIRE(!s->predrifted, ASS_ESTATE, "positions are pre-drifted: finish a step without pre_drift_last first");
	if (i_low == i_high) return ASS_OK;
	REQUIRE(base, ASS_EINVAL, "null particle array");
	int rc = ass_ensure_device();
	if (rc) return rc;
	if (fields & ASS_F_MASS) s->mass_epoch++;
	return upload_rows(s, base, stride, i_low, i_high, fields);
}

extern "C" int ass_download_particles(ass_sim* s, void* base, size_t stride, int n, unsigned fields) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(n == s->n, ASS_EINVAL, "particle count does not match the device state");
	REQUIRE(n == 0 || base, ASS_EINVAL, "null particle array");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "stride must hold the 10 leading doubles and be 8-byte aligned");
	REQUIRE((fields & ~ASS_F_ALL) == 0 && fields != 0, ASS_EINVAL, "bad field mask");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	if (n == 0) return ASS_OK;
	int rc = ass_ensure_device();
	if (rc) return rc;
	return download_rows(s, base, stride, 0, n, fields);
}

// ass_download_particles for rows [i_low, i_high) only: `base` is row 0 of the caller's array (e.g. to read one point mass
// for output while the body stays resident, or a rank's own slab of a sharded body).
extern "C" int ass_download_range(ass_sim* s, void* base, size_t stride, int i_low, int i_high, unsigned fields) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(s->have_state, ASS_ESTATE, "no particle state on the device");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	REQUIRE(i_low >= 0 && i_high <= s->n && i_low <= i_high, ASS_EINVAL, "bad particle range");
	REQUIRE(stride >= 10 * sizeof(double) && stride % sizeof(double) == 0, ASS_EINVAL, "stride must hold the 10 leading doubles and be 8-byte aligned");
	REQUIRE((fields & ~ASS_F_ALL) == 0 && fields != 0, ASS_EINVAL, "bad field mask");
	if (i_low == i_high) return ASS_OK;
	REQUIRE(base, ASS_EINVAL, "null particle array");
	int rc = ass_ensure_device();
	if (rc) return rc;
	return download_rows(s, base, stride, i_low, i_high, fields);
}

// ---------------------------------------------------------------------------------------------------
// springs: caller's list -> node-sorted 16-byte half-edges + palette of distinct (k, gamma) pairs
// ---------------------------------------------------------------------------------------------------
namespace {
struct KG {
	uint64_t k, g;
	bool operator==(const KG& o) const { return k == o.k && g == o.g; }
};
struct KGHash {
	size_t operator()(const KG& v) const {
		uint64_t x = v.k * 0x9E3779B97F4A7C15ULL ^ (v.g + 0x7F4A7C15ULL + (v.k << 6) + (v.k >> 2));
		x ^= x >> 31;
		return (size_t) (x * 0xD6E8FEB86659FD93ULL);
	}
};
// bit patterns, not values: -0.0 and 0.0, or two NaNs, stay distinct entries and round-trip exactly
void build_palette(const ass_spring* sp, int ns, std::vector<double2>& pal, std::vector<int>& idx) {
	std::unordered_map<KG, int, KGHash> seen;
	idx.resize((size_t) ns);
	pal.clear();
	for (int q = 0; q < ns; q++) {
		KG key;
		memcpy(&key.k, &sp[q].k, 8);
		memcpy(&key.g, &sp[q].gamma, 8);
		auto it = seen.find(key);
		if (it == seen.end()) {
			it = seen.emplace(key, (int) pal.size()).first;
			pal.push_back(make_double2(sp[q].k, sp[q].gamma));
		}
		idx[(size_t) q] = it->second;
	}
}
// device palette: pal_cap entries (k, gamma) for the diagnostics, then pal_cap entries (-k, -max(gamma, 0)) for the
// force kernels (spring_math.cuh); the negations are exact, so the forces carry the bits of (-k)*(len - rs0)
int upload_palette(ass_sim* s, const std::vector<double2>& pal) {
	if (pal.size() > s->pal_cap) {
		cudaFree(s->pal);
		s->pal = nullptr; s->palf = nullptr; s->pal_cap = 0;
		const size_t cap = pal.size() + pal.size() / 4 + 16;
		if (cudaMalloc(&s->pal, sizeof(double2) * 2 * cap) != cudaSuccess) {
			cudaGetLastError();
			ass_set_error("out of device memory for a palette of %zu (k, gamma) pairs", cap);
			return ASS_ENOMEM;
		}
		s->pal_cap = cap;
		s->palf = s->pal + cap;
	}
	if (!pal.empty()) {
		std::vector<double2> palf(pal.size());
		for (size_t i = 0; i < pal.size(); i++) palf[i] = make_double2(-pal[i].x, pal[i].y > 0.0 ? -pal[i].y : -0.0);
		CU_TRY(cudaMemcpyAsync(s->pal, pal.data(), sizeof(double2) * pal.size(), cudaMemcpyHostToDevice, s->stream));
		CU_TRY(cudaMemcpyAsync(s->palf, palf.data(), sizeof(double2) * pal.size(), cudaMemcpyHostToDevice, s->stream));
		CU_TRY(cudaStreamSynchronize(s->stream));  // both vectors are pageable and die with their scopes
	}
	s->n_pal = (int) pal.size();
	return ASS_OK;
}
}  // namespace

namespace {
// interleave the low 21 bits of v with two zero bits between each
inline uint64_t spread21(uint64_t v) {
	v &= 0x1fffffULL;
	v = (v | v << 32) & 0x1f00000000ffffULL;
	v = (v | v << 16) & 0x1f0000ff0000ffULL;
	v = (v | v << 8) & 0x100f00f00f00f00fULL;
	v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
	v = (v | v << 2) & 0x1249249249249249ULL;
	return v;
}

// Rank the nodes along a Morton (Z-order) curve through their current positions: order[k] = node of rank k.
// Non-finite coordinates rank as the box corner; ties keep node order (stable), so the ranking is deterministic.
void morton_order(const vec4d* pos, size_t n, std::vector<int>& order) {
	double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
	for (size_t i = 0; i < n; i++) {
		const double c[3] = { pos[i].x, pos[i].y, pos[i].z };
		for (int a = 0; a < 3; a++)
			if (std::isfinite(c[a])) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); }
	}
	double span = 0.0;
	for (int a = 0; a < 3; a++)
		if (hi[a] >= lo[a]) span = std::max(span, hi[a] - lo[a]);
	const double scale = span > 0.0 ? 2097151.0 / span : 0.0;  // one cubic grid over the bounding box: cells stay cubes
	std::vector<std::pair<uint64_t, int>> keyed(n);
	for (size_t i = 0; i < n; i++) {
		const double c[3] = { pos[i].x, pos[i].y, pos[i].z };
		uint64_t key = 0;
		for (int a = 0; a < 3; a++) {
			double q = std::isfinite(c[a]) ? (c[a] - lo[a]) * scale : 0.0;
			q = std::min(std::max(q, 0.0), 2097151.0);
			key |= spread21((uint64_t) q) << a;
		}
		keyed[i] = { key, (int) i };
	}
	std::sort(keyed.begin(), keyed.end());
	order.resize(n);
	for (size_t k = 0; k < n; k++) order[k] = keyed[k].second;
}

int reserve_sorted(ass_sim* s, int n) {
	if ((size_t) n <= s->sorted_cap) return ASS_OK;
	cudaFree(s->order); cudaFree(s->rank); cudaFree(s->mir_block); cudaFree(s->mir_m);
	s->order = s->rank = nullptr;
	s->mir = SpringMirror(); s->mir_alt = SpringMirror(); s->mir_m = nullptr; s->mir_block = nullptr;
	s->sorted_cap = 0;
	const size_t cap = ((size_t) n + n / 8 + 64 + 7) & ~(size_t) 7;
	double2* block = nullptr;  // the mirror and its ping-pong twin: six arrays of 16-byte pairs
	if (cudaMalloc(&s->order, sizeof(int) * cap) != cudaSuccess || cudaMalloc(&s->rank, sizeof(int) * cap) != cudaSuccess ||
	    cudaMalloc(&block, sizeof(double2) * 6 * cap) != cudaSuccess || cudaMalloc(&s->mir_m, sizeof(double) * cap) != cudaSuccess) {
		cudaGetLastError();
		cudaFree(block);
		ass_set_error("out of device memory for the rank-ordered mirror of %zu nodes", cap);
		return ASS_ENOMEM;
	}
	s->mir_block = block;
	s->mir.xy = block; s->mir.zv = block + cap; s->mir.vxy = block + 2 * cap;
	s->mir_alt.xy = block + 3 * cap; s->mir_alt.zv = block + 4 * cap; s->mir_alt.vxy = block + 5 * cap;
	s->sorted_cap = cap;
	return ASS_OK;
}

template <class T>
int reserve_array(T*& ptr, size_t& cap, size_t need, const char* what) {
	if (need <= cap) return ASS_OK;
	cudaFree(ptr);
	ptr = nullptr; cap = 0;
	const size_t c = need + need / 8 + 64;
	if (cudaMalloc(&ptr, sizeof(T) * c) != cudaSuccess) {
		cudaGetLastError();
		ass_set_error("out of device memory for %zu %s", c, what);
		return ASS_ENOMEM;
	}
	cap = c;
	return ASS_OK;
}
}  // namespace

void ass_morton_order(const vec4d* pos, size_t n, std::vector<int>& order) { morton_order(pos, n, order); }

// Which of the eight 16-byte bank groups a node's state lands in is its rank mod 8 (mirror arrays, shared-memory tables).
// A row gathers fastest when its neighbours are spread evenly over the eight groups, so the Morton ranking is refined:
// ranks are permuted INSIDE every block of eight consecutive ranks (a node moves by at most seven places: locality is
// untouched) by local search -- swap two ranks of a block whenever that lowers  sum over rows, sum over groups of
// (neighbours of the row in the group)^2 -- for a few sweeps.  Deterministic.  Measured on a BASELINE C1 body: gather
// passes per quarter-warp step 1.56 -> 1.36 (with the row arrangement below: 1.32).
void refine_ranks_for_bank_groups(int n, const ass_spring* sp, int ns, std::vector<int>& order, std::vector<int>& rank) {
	if (n < 2 || ns <= 0) return;
	// adjacency by Morton rank (rows sorted: set differences by merging)
	std::vector<int> off((size_t) n + 1, 0);
	for (int q = 0; q < ns; q++) {
		off[(size_t) rank[sp[q].particle_1] + 1]++;
		off[(size_t) rank[sp[q].particle_2] + 1]++;
	}
	for (int k = 0; k < n; k++) off[(size_t) k + 1] += off[k];
	std::vector<int> adj((size_t) off[n]);
	{
		std::vector<int> fill(off.begin(), off.end() - 1);
		for (int q = 0; q < ns; q++) {
			const int a = rank[sp[q].particle_1], b = rank[sp[q].particle_2];
			adj[(size_t) fill[a]++] = b;
			adj[(size_t) fill[b]++] = a;
		}
	}
	for (int k = 0; k < n; k++) std::sort(adj.begin() + off[k], adj.begin() + off[(size_t) k + 1]);
	// cls[k]: bank group of the node whose MORTON rank is k (changes as ranks are swapped); cnt[i][c]: neighbours of i in c
	std::vector<unsigned char> cls((size_t) n);
	for (int k = 0; k < n; k++) cls[(size_t) k] = (unsigned char) (k & 7);
	std::vector<int> cnt((size_t) n * 8, 0);
	for (int k = 0; k < n; k++)
		for (int e = off[k]; e < off[(size_t) k + 1]; e++) cnt[(size_t) k * 8 + cls[(size_t) adj[(size_t) e]]]++;
	// change of the objective if nodes a and b (Morton ranks) exchange their groups; rows that hold both are unaffected
	auto walk = [&](int a, int b, bool apply) {
		const int ca = cls[(size_t) a], cb = cls[(size_t) b];
		long long d = 0;
		int ea = off[a], eb = off[b];
		const int ea1 = off[(size_t) a + 1], eb1 = off[(size_t) b + 1];
		while (ea < ea1 || eb < eb1) {
			const int ia = ea < ea1 ? adj[(size_t) ea] : 0x7fffffff, ib = eb < eb1 ? adj[(size_t) eb] : 0x7fffffff;
			if (ia == ib) { ea++; eb++; continue; }
			int* c = cnt.data() + (size_t) std::min(ia, ib) * 8;
			const int from = ia < ib ? ca : cb, to = ia < ib ? cb : ca;
			d += 2LL * (c[to] - c[from]) + 2;
			if (apply) { c[from]--; c[to]++; }
			if (ia < ib) ea++; else eb++;
		}
		return d;
	};
	for (int sweep = 0; sweep < 3; sweep++) {
		long long moves = 0;
		for (int b0 = 0; b0 < n; b0 += 8) {
			const int b1 = std::min(n, b0 + 8);
			for (int round = 0; round < 3; round++) {
				long long best = 0;
				int ba = -1, bb = -1;
				for (int a = b0; a < b1; a++)
					for (int b = a + 1; b < b1; b++) {
						const long long d = walk(a, b, false);
						if (d < best) { best = d; ba = a; bb = b; }
					}
				if (ba < 0) break;
				walk(ba, bb, true);
				std::swap(cls[(size_t) ba], cls[(size_t) bb]);
				moves++;
			}
		}
		if (moves == 0) break;
	}
	// new rank of the node with Morton rank k: same block of eight, its final group
	std::vector<int> new_order((size_t) n);
	for (int k = 0; k < n; k++) new_order[(size_t) ((k & ~7) | cls[(size_t) k])] = order[(size_t) k];
	order.swap(new_order);
	for (int k = 0; k < n; k++) rank[(size_t) order[(size_t) k]] = k;
}

// Host-side construction of everything the spring kernels read (common.cuh: SpringLayout): spatial ranking, CSR rows
// by rank (rows ordered bank-group-aware, see below), and the lane-transposed copy.  Shared by ass_set_springs and the
// ensemble (one layout per member), so a member sums in the same order as the single-body path and gets the same bits.
// pidx[q] = palette index of spring q; *bad_spring = first spring with an endpoint outside [0, n) or i == j, else -1.
int ass_build_spring_layout(const vec4d* pos, int n, const ass_spring* sp, int ns, const int* pidx, SpringLayout& out, int* bad_spring) {
	*bad_spring = -1;
	// ASS_LAYOUT_TIMING=1: the phases' host times on stderr (tuning)
	const char* tim_env = getenv("ASS_LAYOUT_TIMING");
	const bool timing = tim_env && *tim_env && *tim_env != '0';
	auto t_last = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!timing) return;
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[layout] %-28s %8.1f ms\n", what, 1e3 * std::chrono::duration<double>(now - t_last).count());
		t_last = now;
	};
	for (int q = 0; q < ns; q++) {
		const int i = sp[q].particle_1, j = sp[q].particle_2;
		if (i < 0 || j < 0 || i >= n || j >= n || i == j) {
			*bad_spring = q;
			return ASS_EINVAL;
		}
	}
	std::vector<int>& order = out.order;
	std::vector<int>& rank = out.rank;
	std::vector<int>& row = out.row;
	rank.assign((size_t) n, 0);
	morton_order(pos, (size_t) n, order);
	for (int k = 0; k < n; k++) rank[(size_t) order[k]] = k;
	lap("morton ranking");
	refine_ranks_for_bank_groups(n, sp, ns, order, rank);
	lap("rank refinement");
	// rows by rank
	row.assign((size_t) n + 1, 0);
	for (int q = 0; q < ns; q++) {
		row[(size_t) rank[sp[q].particle_1] + 1]++;
		row[(size_t) rank[sp[q].particle_2] + 1]++;
	}
	out.max_degree = 0;
	for (int k = 0; k < n; k++) {
		out.max_degree = std::max(out.max_degree, row[(size_t) k + 1]);
		row[(size_t) k + 1] += row[k];
	}
	const size_t nhe = (size_t) 2 * ns;
	// (neighbour rank, 2*spring + side) per half-edge, bucketed by owner row, then each row ordered by neighbour rank so
	// the lanes of a group gather ascending, mostly adjacent entries
	std::vector<std::pair<int, int>> ent(nhe);
	{
		std::vector<int> fill(row.begin(), row.end() - 1);
		for (int q = 0; q < ns; q++) {
			const int e[2] = { rank[sp[q].particle_1], rank[sp[q].particle_2] };
			for (int side = 0; side < 2; side++) ent[(size_t) fill[e[side]]++] = { e[1 - side], 2 * q + side };
		}
	}
	for (int k = 0; k < n; k++) std::sort(ent.begin() + row[k], ent.begin() + row[(size_t) k + 1]);
	lap("rows");
	// ---- the lane-transposed copy (common.cuh: ass_sim::heT) ----
	// nodes -> lane groups: inside every block of ASS_SORT_BLOCK ranks, by decreasing degree (ties: by rank)
	constexpr int L = SPRING_LANES, GPW = 32 / SPRING_LANES;
	const int n_swarps = (n + GPW - 1) / GPW;
	const size_t n_grp = (size_t) n_swarps * GPW;
	std::vector<int2>& grp = out.grp;
	grp.assign(n_grp, make_int2(-1, 0));
	for (int b0 = 0; b0 < n; b0 += ASS_SORT_BLOCK) {
		const int b1 = std::min(n, b0 + ASS_SORT_BLOCK);
		for (int k = b0; k < b1; k++) grp[(size_t) k] = make_int2(k, row[(size_t) k + 1] - row[k]);
		std::stable_sort(grp.begin() + b0, grp.begin() + b1, [](const int2& x, const int2& y) { return x.y > y.y; });
	}
	// ---- order inside the rows: spread every step's gathers over the eight 16-byte bank groups ----
	// Eight consecutive lanes (a quarter warp: QG = 8 / SPRING_LANES groups, i.e. QG nodes) are served together by the
	// L1 / shared-memory data array, 128 bytes per pass; their eight 16-byte gathers of a step take one pass only if the
	// eight neighbour ranks differ mod 8.  With rows simply sorted by neighbour rank a step costs ~2.1 passes (measured).
	// So the rows of the QG nodes that share a quarter warp are re-ordered TOGETHER, step by step: with two rows (four
	// lanes per node) the eight groups are split four and four in the way that places most entries on distinct groups;
	// otherwise each bank group goes to the row that has most entries left in it (greedy, fullest groups first); a row
	// short of distinct groups takes what it has left; inside a step entries stay in ascending rank.  The order of a row is the
	// order of its sum, in every kernel (spring_math.cuh): deterministic, a function of the spring list and positions.
	{
		constexpr int QG = 8 / SPRING_LANES;
		std::vector<std::pair<int, int>> bucket[QG][8], picked[QG];
		size_t taken[QG][8];
		for (size_t q0 = 0; q0 < n_grp; q0 += QG) {
			int left[QG], written[QG];
			bool any = false;
			for (int r = 0; r < QG; r++) {
				const int2 g = grp[q0 + r];
				left[r] = g.x >= 0 ? g.y : 0;
				written[r] = 0;
				for (int c = 0; c < 8; c++) { bucket[r][c].clear(); taken[r][c] = 0; }
				if (left[r] > 0) {
					any = true;
					for (int e = row[g.x]; e < row[(size_t) g.x + 1]; e++) bucket[r][ent[(size_t) e].first & 7].push_back(ent[(size_t) e]);
				}
			}
			if (!any) continue;
			auto remaining = [&](int r, int c) { return (int) (bucket[r][c].size() - taken[r][c]); };
			while (true) {
				int need[QG], total_need = 0;
				for (int r = 0; r < QG; r++) { need[r] = std::min(L, left[r]); total_need += need[r]; picked[r].clear(); }
				if (total_need == 0) break;
				if (QG == 2) {
					// two rows share the quarter warp: split the eight groups four and four so that as many entries as possible
					// land on distinct groups, drawing on the fullest groups first.  With value(c -> row) = 1000 + entries left
					// (0 if none) the best split gives row 0 the four groups with the largest value(c -> 0) - value(c -> 1).
					int diff[8], corder2[8];
					for (int c = 0; c < 8; c++) {
						const int v0 = remaining(0, c) > 0 ? 1000 + remaining(0, c) : 0, v1 = remaining(1, c) > 0 ? 1000 + remaining(1, c) : 0;
						diff[c] = v0 - v1;
						corder2[c] = c;
					}
					std::stable_sort(corder2, corder2 + 8, [&](int x, int y) { return diff[x] > diff[y]; });
					int best_mask = 0;
					for (int i = 0; i < 4; i++) best_mask |= 1 << corder2[i];
					for (int r = 0; r < 2; r++) {
						// the fullest `need` of the row's own groups (ties: the lower group)
						int cs[8], nc = 0;
						for (int c = 0; c < 8; c++)
							if ((((best_mask >> c) & 1) == (r == 0)) && remaining(r, c) > 0) cs[nc++] = c;
						std::stable_sort(cs, cs + nc, [&](int a, int b) { return remaining(r, a) > remaining(r, b); });
						for (int i = 0; i < nc && (int) picked[r].size() < need[r]; i++) picked[r].push_back(bucket[r][cs[i]][taken[r][cs[i]]++]);
					}
				}
				int corder[8];
				for (int c = 0; c < 8; c++) corder[c] = c;
				std::stable_sort(corder, corder + 8, [&](int a, int b) {
					int ma = 0, mb = 0;
					for (int r = 0; r < QG; r++) { ma = std::max(ma, remaining(r, a)); mb = std::max(mb, remaining(r, b)); }
					return ma > mb;
				});
				for (int ci = 0; ci < 8 && QG != 2; ci++) {
					const int c = corder[ci];
					int best = -1, best_left = 0;
					for (int r = 0; r < QG; r++)
						if ((int) picked[r].size() < need[r] && remaining(r, c) > best_left) { best = r; best_left = remaining(r, c); }
					if (best >= 0) picked[best].push_back(bucket[best][c][taken[best][c]++]);
				}
				for (int r = 0; r < QG; r++) {
					while ((int) picked[r].size() < need[r]) {  // not enough distinct groups left: take from the fullest
						int c_best = 0;
						for (int c = 1; c < 8; c++)
							if (remaining(r, c) > remaining(r, c_best)) c_best = c;
						picked[r].push_back(bucket[r][c_best][taken[r][c_best]++]);
					}
					std::sort(picked[r].begin(), picked[r].end());
					const int2 g = grp[q0 + r];
					for (auto& pe : picked[r]) ent[(size_t) row[g.x] + written[r]++] = pe;
					left[r] -= need[r];
				}
			}
		}
	}
	lap("groups + row arrangement");
	out.he.resize(nhe);
	out.slot.resize(nhe);
	for (size_t at = 0; at < nhe; at++) {
		const int q = ent[at].second >> 1;
		HalfEdge h;
		h.rs0 = sp[q].rs0;
		h.nbr = ent[at].first;
		h.pal = pidx[(size_t) q];
		out.he[at] = h;
		out.slot[(size_t) ent[at].second] = (int) at;
	}
	std::vector<int>& warp_off = out.warp_off;
	warp_off.assign((size_t) n_swarps + 1, 0);
	for (int w = 0; w < n_swarps; w++) {
		int steps = 0;
		for (int g = 0; g < GPW; g++) steps = std::max(steps, (grp[(size_t) w * GPW + g].y + L - 1) / L);
		steps = (steps + 1) & ~1;  // the kernels work on pairs of steps
		const long long next = (long long) warp_off[w] + 32LL * steps;
		if (next > 0x7fffffffLL) return ASS_ENOMEM;  // 32-bit record index
		warp_off[(size_t) w + 1] = (int) next;
	}
	const size_t n_heT = (size_t) warp_off[n_swarps];
	out.heT.resize(n_heT);
	out.slotT.resize(nhe);
	for (int w = 0; w < n_swarps; w++) {
		const int steps = (warp_off[(size_t) w + 1] - warp_off[w]) / 32;
		for (int g = 0; g < GPW; g++) {
			const int2 gi = grp[(size_t) w * GPW + g];
			HalfEdge padrec;  // loaded and computed, never accumulated: must gather from a valid entry
			padrec.rs0 = 0.0; padrec.nbr = gi.x >= 0 ? gi.x : 0; padrec.pal = 0;
			for (int t = 0; t < steps; t++)
				for (int l = 0; l < L; l++) {
					const int e = t * L + l;
					const size_t at_t = (size_t) warp_off[w] + (size_t) t * 32 + (size_t) g * L + l;
					if (e < gi.y) {
						const size_t at = (size_t) row[gi.x] + e;
						out.heT[at_t] = out.he[at];
						out.slotT[(size_t) ent[at].second] = (int) at_t;
					} else {
						out.heT[at_t] = padrec;
					}
				}
		}
	}
	lap("CSR + transposed copies");
	return ASS_OK;
}

extern "C" int ass_set_springs(ass_sim* s, const ass_spring* sp, int ns) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(ns >= 0, ASS_EINVAL, "negative spring count");
	REQUIRE(ns == 0 || sp, ASS_EINVAL, "null spring array");
	REQUIRE(s->have_state, ASS_ESTATE, "upload particles before springs");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	int rc = ass_ensure_device();
	if (rc) return rc;
	const int n = s->n;
	// spatial ranking of the nodes from their positions now (common.cuh: "springs")
	std::vector<vec4d> pos((size_t) n);
	if (n) {
		CU_TRY(cudaMemcpyAsync(pos.data(), s->posm, sizeof(vec4d) * (size_t) n, cudaMemcpyDeviceToHost, s->stream));
		CU_TRY(cudaStreamSynchronize(s->stream));
	}
	std::vector<double2> pal;
	std::vector<int> pidx;
	build_palette(sp, ns, pal, pidx);
	SpringLayout lay;
	int bad = -1;
	rc = ass_build_spring_layout(pos.data(), n, sp, ns, pidx.data(), lay, &bad);
	if (rc == ASS_EINVAL) {
		const int i = sp[bad].particle_1, j = sp[bad].particle_2;
		if (i == j) ass_set_error("ass_set_springs: spring %d connects particle %d to itself", bad, i);  // springs.cpp:57-59 throws
		else ass_set_error("ass_set_springs: spring %d joins (%d,%d) outside [0,%d)", bad, i, j, n);
		return rc;
	}
	if (rc) {
		ass_set_error("ass_set_springs: %d springs overflow the 32-bit record index of the lane layout", ns);
		return rc;
	}
	std::vector<vec4d>().swap(pos);
	const size_t nhe = lay.he.size();
	// from here on the old list's arrays may be freed and re-allocated: until the new list is complete the simulation
	// has NO springs (an allocation failure below must not leave kernels pointed at freed rows)
	s->ns = 0; s->n_he = 0; s->n_swarps = 0; s->tiles_grid = 0;
	s->spring_epoch++;
	if ((rc = reserve_sorted(s, n))) return rc;
	if (nhe > s->he_cap) {
		cudaFree(s->he); cudaFree(s->slot); cudaFree(s->slotT);
		s->he = nullptr; s->slot = nullptr; s->slotT = nullptr; s->he_cap = 0;
		size_t cap = nhe + nhe / 8 + 64;
		if (cudaMalloc(&s->he, sizeof(HalfEdge) * cap) != cudaSuccess || cudaMalloc(&s->slot, sizeof(int) * cap) != cudaSuccess ||
		    cudaMalloc(&s->slotT, sizeof(int) * cap) != cudaSuccess) {
			cudaGetLastError();
			ass_set_error("out of device memory for %zu half-edges", cap);
			return ASS_ENOMEM;
		}
		s->he_cap = cap;
	}
	if ((rc = reserve_array(s->warp_off, s->swarp_cap, lay.warp_off.size(), "warp offsets"))) return rc;
	if ((rc = reserve_array(s->row_ptr, s->row_cap, lay.row.size(), "row pointers"))) return rc;
	ass_tic(s, ASS_K_XFER);
	if ((rc = upload_palette(s, pal))) return rc;
	if (nhe) {
		CU_TRY(cudaMemcpyAsync(s->he, lay.he.data(), sizeof(HalfEdge) * nhe, cudaMemcpyHostToDevice, s->stream));
		CU_TRY(cudaMemcpyAsync(s->slot, lay.slot.data(), sizeof(int) * nhe, cudaMemcpyHostToDevice, s->stream));
		CU_TRY(cudaMemcpyAsync(s->slotT, lay.slotT.data(), sizeof(int) * nhe, cudaMemcpyHostToDevice, s->stream));
	}
	CU_TRY(cudaMemcpyAsync(s->row_ptr, lay.row.data(), sizeof(int) * lay.row.size(), cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync(s->warp_off, lay.warp_off.data(), sizeof(int) * lay.warp_off.size(), cudaMemcpyHostToDevice, s->stream));
	if (n) {
		CU_TRY(cudaMemcpyAsync(s->order, lay.order.data(), sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, s->stream));
		CU_TRY(cudaMemcpyAsync(s->rank, lay.rank.data(), sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, s->stream));
	}
	ass_toc(s, ASS_K_XFER);
	CU_TRY(cudaStreamSynchronize(s->stream));
	s->n_swarps = (int) lay.warp_off.size() - 1;
	if ((rc = ass_build_spring_tiles(s, lay))) return rc;  // the layout dies after this
	s->ns = ns;
	s->n_he = (int) nhe;
	s->max_degree = lay.max_degree;
	return ASS_OK;
}

// stage holds [ns x ass_spring | ns x int palette index]
__global__ void k_scatter_props(const ass_spring* __restrict__ sp, const int* __restrict__ pidx, const int* __restrict__ slot,
		const int* __restrict__ slotT, HalfEdge* __restrict__ he, HalfEdge* __restrict__ heL, int ns) {
	int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= ns) return;
	const double rs0 = sp[q].rs0;
	const int pi = pidx[q];
	for (int side = 0; side < 2; side++) {
		HalfEdge* h = he + slot[(size_t) 2 * q + side];
		h->rs0 = rs0;
		h->pal = pi;
		if (heL) {  // the tiled kernel's copy (absent when no tile cut exists: springs.cu, ass_build_spring_tiles)
			HalfEdge* hl = heL + slotT[(size_t) 2 * q + side];
			hl->rs0 = rs0;
			hl->pal = pi;
		}
	}
}

extern "C" int ass_update_spring_props(ass_sim* s, const ass_spring* sp, int ns) {
	REQUIRE(s, ASS_EINVAL, "null sim");
	REQUIRE(ns == s->ns, ASS_EINVAL, "spring count differs from the device list: call ass_set_springs");
	if (ns == 0) return ASS_OK;
	REQUIRE(sp, ASS_EINVAL, "null spring array");
	int rc = ass_ensure_device();
	if (rc) return rc;
	std::vector<double2> pal;
	std::vector<int> pidx;
	build_palette(sp, ns, pal, pidx);
	const size_t sp_bytes = sizeof(ass_spring) * (size_t) ns;
	if ((rc = ass_buf_reserve(s->stage, sp_bytes + sizeof(int) * (size_t) ns))) return rc;
	ass_tic(s, ASS_K_XFER);
	if ((rc = upload_palette(s, pal))) return rc;
	CU_TRY(cudaMemcpyAsync(s->stage.p, sp, sp_bytes, cudaMemcpyHostToDevice, s->stream));
	CU_TRY(cudaMemcpyAsync((char*) s->stage.p + sp_bytes, pidx.data(), sizeof(int) * (size_t) ns, cudaMemcpyHostToDevice, s->stream));
	k_scatter_props<<<(ns + 255) / 256, 256, 0, s->stream>>>((const ass_spring*) s->stage.p, (const int*) ((char*) s->stage.p + sp_bytes), s->slot, s->slotT, s->he, s->tiles_grid > 0 ? s->heL : nullptr, ns);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_XFER);
	CU_TRY(cudaStreamSynchronize(s->stream));
	return ASS_OK;
}
