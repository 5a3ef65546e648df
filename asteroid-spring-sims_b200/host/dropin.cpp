// dropin.cpp -- the reference's own symbols for the per-step path, re-implemented over the C-ABI of include/ass_b200.h.
//
// Compiled AGAINST the reference's headers (-I <reference>/src -I <reference>/src_spring), never with a copy of them:
// struct reb_simulation / reb_particle / spring / Vector are the reference's.  Linked in front of the reference's
// objects (whose same-named definitions are weakened) this makes an unmodified problem module run
// its step loop on the B200.  Residency rules are described in ass_dropin.h.
//
// Each function cites what it replaces.  Error convention: REBOUND-side failures print and
// exit(EXIT_FAILURE) through the reference's reb_exit(); spring-side failures `throw "literal"` exactly like
// src_spring/springs.cpp so EXPECT_ANY_THROW-style callers keep working; nothing throws across the C-ABI.
#ifdef __GNUC__
#define restrict __restrict__
#endif

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>
#include <mutex>
#include <unordered_map>
#include <vector>

extern "C" {
#include "rebound.h"
// exported by librebound but not declared in its public header
int reb_check_exit(struct reb_simulation* const r, const double tmax, double* last_full_dt);  // src/rebound.c:509
void reb_exit(const char* const msg);                                                           // src/rebound.c:147
}
#include "matrix_math.h"
#include "stress.h"
#include "springs.h"
#include "physics.h"
#include "shapes.h"
#include "orb.h"

#include "ass_b200.h"
#include "ass_dropin.h"

using std::vector;

// module-owned globals (e.g. modules/asteroid_spin/problem.cpp:37-47)
extern vector<spring> springs;
extern int num_springs;
extern int num_perts;
extern vector<stress_tensor> stresses;

static_assert(sizeof(spring) == sizeof(ass_spring), "struct spring must be 32 bytes {k, rs0, gamma, particle_1, particle_2}");

namespace {

struct Dropin {
	ass_sim* sim = nullptr;
	bool dev_fresh = false;    // device copy reflects the host arrays (or is newer)
	bool host_fresh = true;    // host arrays reflect the device copy (or are newer)
	unsigned pending = 0;      // fields written on the device inside a deferred sequence, not yet downloaded
	int defer = 0;             // >0: inside one of our own multi-kernel sequences
	bool in_run = false;       // inside reb_integrate
	// resident mode: calls of one reb_step deferred until reb_integrator_leapfrog_part2 can run them as fused kernels
	bool lazy_part1 = false, lazy_zero = false, lazy_springs = false;
	// spring cache validity
	bool topo_stale = true, prop_stale = true;
	size_t sp_size = 0;
	const spring* sp_ptr = nullptr;
	// r->particles page-locked for the length of a reb_integrate (uploads / downloads at PCIe link rate)
	void* pinned = nullptr;
	size_t pinned_bytes = 0;
	// diagnostics of one device state, shared by the functions write_resolved_with_E calls one after another
	uint64_t version = 1;        // bumped whenever the device state or the spring list may have changed
	uint64_t diag_version = 0;
	int diag_lo = -1, diag_hi = -1;
	bool diag_gpe = false;
	double diag[20];
	// fingerprints of the module's spring list: wiring (particle_1, particle_2) and properties (k, rs0, gamma) apart
	uint64_t sp_probe_topo = 0, sp_probe_prop = 0;
};

std::mutex g_mu;
std::unordered_map<reb_simulation*, Dropin> g_map;
int g_resident = -1;

bool resident() {
	if (g_resident < 0) {
		const char* e = getenv("ASS_RESIDENT");
		g_resident = (e && *e && strcmp(e, "0") != 0) ? 1 : 0;
	}
	return g_resident == 1;
}

[[noreturn]] void fatal(const char* where) {
	static char buf[1400];
	snprintf(buf, sizeof(buf), "%s: %s", where, ass_last_error());
	reb_exit(buf);  // prints "Fatal error! Exiting now." and exit(EXIT_FAILURE), src/rebound.c:147-153
	abort();
}
#define OK(call)                      \
	do {                              \
		if ((call) < 0) fatal(#call); \
	} while (0)

Dropin& get(reb_simulation* r) {
	std::lock_guard<std::mutex> lk(g_mu);
	Dropin& d = g_map[r];
	if (!d.sim) {
		if (ass_create(&d.sim) < 0) fatal("ass_create");
	}
	return d;
}

void push_params(reb_simulation* r, Dropin& d) {
	ass_params p;
	memset(&p, 0, sizeof(p));
	p.t = r->t; p.dt = r->dt; p.dt_last_done = r->dt_last_done;
	p.G = r->G; p.softening = r->softening;
	p.gravity = (r->gravity == reb_simulation::REB_GRAVITY_BASIC) ? 1 : 0;
	p.n_active = r->N_active;
	p.n_perts = num_perts;
	p.additional_forces = 0;  // callbacks are the module's and run on the host side of this layer
	p.heartbeat = 0;
	p.exact_finish_time = r->exact_finish_time;
	p.status = r->status;
	OK(ass_set_params(d.sim, &p));
}

// Inside reb_integrate the particle array is page-locked so transfers run at link rate.  reb_add may realloc it
// (src/particle.c:53-56): the registration follows the pointer; a failure to pin is not an error (transfers are staged).
void unpin_particles(Dropin& d) {
	if (d.pinned) ass_host_unregister(d.pinned);
	d.pinned = nullptr;
	d.pinned_bytes = 0;
}
void pin_particles(reb_simulation* r, Dropin& d) {
	const size_t bytes = sizeof(reb_particle) * (size_t) r->N;
	if (d.pinned == (void*) r->particles && bytes <= d.pinned_bytes) return;
	unpin_particles(d);
	const char* off = getenv("ASS_NO_PIN");
	if (!d.in_run || r->N <= 0 || (off && *off && strcmp(off, "0") != 0)) return;
	if (ass_host_register(r->particles, bytes) == ASS_OK) {
		d.pinned = r->particles;
		d.pinned_bytes = bytes;
	}
}

void push_params(reb_simulation* r, Dropin& d);
void ensure_springs(Dropin& d);

// Resident mode defers part1 / zero_accel / spring_forces of a reb_step (they return at once) so that
// reb_integrator_leapfrog_part2 can run the lot as the fused kernels of the step loop: two launches instead of five, no
// re-gather of the rank-ordered mirror.  Anything else that looks at the device state first SETTLES what is pending, in the
// order it was asked for, through the ordinary calls -- the result is the same bits either way.
// ASS_DROPIN_NO_FUSE=1: every call runs when it is made.
bool lazy_ok(const Dropin& d) {
	static int no_fuse = -1;
	if (no_fuse < 0) {
		const char* e = getenv("ASS_DROPIN_NO_FUSE");
		no_fuse = (e && *e && strcmp(e, "0") != 0) ? 1 : 0;
	}
	return resident() && d.in_run && !no_fuse;
}
void settle(reb_simulation* r, Dropin& d) {
	if (!(d.lazy_part1 || d.lazy_zero || d.lazy_springs)) return;
	const bool p1 = d.lazy_part1, z = d.lazy_zero, sp = d.lazy_springs;
	d.lazy_part1 = d.lazy_zero = d.lazy_springs = false;
	if (p1) {
		push_params(r, d);
		OK(ass_leapfrog_part1(d.sim));
	}
	if (z) OK(ass_zero_accel(d.sim));
	if (sp) OK(ass_spring_forces(d.sim));
}

void ensure_dev(reb_simulation* r, Dropin& d, bool settle_pending = true) {
	if (settle_pending) settle(r, d);
	if (d.dev_fresh && ass_num_particles(d.sim) == r->N) return;
	if (d.in_run) pin_particles(r, d);
	OK(ass_upload_particles(d.sim, r->particles, sizeof(reb_particle), r->N, ASS_F_ALL));
	d.version++;
	d.dev_fresh = true;
	d.host_fresh = true;
	d.pending = 0;
}

void flush_host(reb_simulation* r, Dropin& d) {
	settle(r, d);
	if (d.in_run && d.pending && r->N > 0) pin_particles(r, d);
	if (d.pending && r->N > 0) OK(ass_download_particles(d.sim, r->particles, sizeof(reb_particle), r->N, d.pending));
	d.pending = 0;
	d.host_fresh = true;
}

// a device operation wrote `fields`
void wrote(reb_simulation* r, Dropin& d, unsigned fields) {
	d.version++;
	d.pending |= fields;
	d.host_fresh = false;
	if (resident() && d.in_run) return;  // the device copy is the truth until someone asks for the host arrays
	if (d.defer) return;                 // end of the enclosing sequence flushes
	flush_host(r, d);
	d.dev_fresh = false;                 // coherent mode: anyone may edit the host arrays before the next call
}

struct Sequence {  // several kernels of ours back to back: transfer once at the end
	reb_simulation* r;
	Dropin& d;
	Sequence(reb_simulation* r_, Dropin& d_) : r(r_), d(d_) { d.defer++; }
	~Sequence() {
		if (--d.defer == 0 && !(resident() && d.in_run)) {
			flush_host(r, d);
			d.dev_fresh = false;
		}
	}
};

// cheap fingerprints of the module's list: they catch add / del / direct writes to springs[i] done by code this layer
// did not intercept.  Wiring and properties are hashed apart, so a direct `springs[i].k = ...` costs a property update
// (one kernel), not a rebuild of the layout (seconds of host time at 1e5 nodes).
void probe_springs(uint64_t* topo, uint64_t* prop) {
	const size_t n = springs.size();
	uint64_t ht = 1469598103934665603ULL ^ n, hp = 1469598103934665603ULL;
	const size_t step = n > 4096 ? n / 1024 : 1;
	auto mix = [&](const spring& sp) {
		uint64_t w[4];
		memcpy(w, &sp, sizeof(w));  // {k, rs0, gamma, (particle_1, particle_2)}
		for (int q = 0; q < 3; q++) { hp ^= w[q]; hp *= 1099511628211ULL; }
		ht ^= w[3]; ht *= 1099511628211ULL;
	};
	for (size_t i = 0; i < n; i += step) mix(springs[i]);
	if (n) mix(springs[n - 1]);
	*topo = ht;
	*prop = hp;
}

void ensure_springs(Dropin& d) {
	if (num_springs != (int) springs.size()) throw "Spring count and size of spring array don't match.\n";  // springs.cpp:38-40,99-101
	uint64_t pt, pp;
	probe_springs(&pt, &pp);
	if (springs.size() != d.sp_size || springs.data() != d.sp_ptr || pt != d.sp_probe_topo) d.topo_stale = true;
	else if (pp != d.sp_probe_prop) d.prop_stale = true;
	if (d.topo_stale) {
		OK(ass_set_springs(d.sim, (const ass_spring*) springs.data(), (int) springs.size()));
		d.version++;
	} else if (d.prop_stale) {
		OK(ass_update_spring_props(d.sim, (const ass_spring*) springs.data(), (int) springs.size()));
		d.version++;
	}
	d.topo_stale = d.prop_stale = false;
	d.sp_size = springs.size();
	d.sp_ptr = springs.data();
	d.sp_probe_topo = pt;
	d.sp_probe_prop = pp;
}

void props_changed() {
	std::lock_guard<std::mutex> lk(g_mu);
	for (auto& kv : g_map) kv.second.prop_stale = true;
}
void topo_changed() {
	std::lock_guard<std::mutex> lk(g_mu);
	for (auto& kv : g_map) kv.second.topo_stale = true;
}

void require_supported(reb_simulation* r) {
	if (r->integrator != reb_simulation::REB_INTEGRATOR_LEAPFROG)
		reb_exit("asteroid-spring-sims_b200: only REB_INTEGRATOR_LEAPFROG runs on the device (every reference module uses it).");
	if (r->gravity != reb_simulation::REB_GRAVITY_NONE && r->gravity != reb_simulation::REB_GRAVITY_BASIC)
		reb_exit("asteroid-spring-sims_b200: only REB_GRAVITY_NONE / REB_GRAVITY_BASIC run on the device.");
	if (r->boundary != reb_simulation::REB_BOUNDARY_NONE || r->collision != reb_simulation::REB_COLLISION_NONE || r->N_var != 0 ||
	    r->nghostx || r->nghosty || r->nghostz || r->testparticle_type)
		reb_exit("asteroid-spring-sims_b200: boundaries, collisions, ghost boxes, variational and test particles are not on the accelerated path.");
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// control surface
// ---------------------------------------------------------------------------------------------------
extern "C" void ass_dropin_set_resident(int on) { g_resident = on ? 1 : 0; }
extern "C" int ass_dropin_is_resident(void) { return resident() ? 1 : 0; }
extern "C" int ass_dropin_sync_host(struct reb_simulation* r) {
	Dropin& d = get(r);
	if (!d.host_fresh) flush_host(r, d);
	return 0;
}
extern "C" void ass_dropin_host_modified(struct reb_simulation* r) {
	Dropin& d = get(r);
	if (!d.host_fresh) flush_host(r, d);  // keep what the device computed for the fields the caller did not touch
	d.dev_fresh = false;
}
extern "C" void ass_dropin_springs_modified(void) { topo_changed(); }
extern "C" void ass_dropin_release(struct reb_simulation* r) {
	std::lock_guard<std::mutex> lk(g_mu);
	auto it = g_map.find(r);
	if (it == g_map.end()) return;
	unpin_particles(it->second);
	ass_destroy(it->second.sim);
	g_map.erase(it);
}

// ---------------------------------------------------------------------------------------------------
// C symbols of librebound
// ---------------------------------------------------------------------------------------------------
// src/integrator_leapfrog.c:39-51
extern "C" void reb_integrator_leapfrog_part1(struct reb_simulation* r) {
	r->gravity_ignore_terms = 0;
	Dropin& d = get(r);
	ensure_dev(r, d);
	if (lazy_ok(d)) {
		d.lazy_part1 = true;  // runs with what follows (settle / reb_integrator_leapfrog_part2)
	} else {
		push_params(r, d);
		OK(ass_leapfrog_part1(d.sim));
	}
	r->t += r->dt / 2.;
	wrote(r, d, ASS_F_POS);
}

// src/integrator_leapfrog.c:52-67
extern "C" void reb_integrator_leapfrog_part2(struct reb_simulation* r) {
	Dropin& d = get(r);
	if (lazy_ok(d) && d.lazy_springs) {
		// [part1] [zero_accel] spring_forces part2, deferred until here: the fused kernels
		ensure_dev(r, d, false);
		const int p1 = d.lazy_part1 ? 1 : 0, z = d.lazy_zero ? 1 : 0;
		d.lazy_part1 = d.lazy_zero = d.lazy_springs = false;
		push_params(r, d);
		OK(ass_substep_fused(d.sim, p1, z));
		r->t += r->dt / 2.;
		r->dt_last_done = r->dt;
		wrote(r, d, ASS_F_POS | ASS_F_VEL | ASS_F_ACC);
		return;
	}
	ensure_dev(r, d);
	push_params(r, d);
	OK(ass_leapfrog_part2(d.sim));
	r->t += r->dt / 2.;
	r->dt_last_done = r->dt;
	wrote(r, d, ASS_F_POS | ASS_F_VEL);
}

// src/gravity.c:57-125, branches NONE and BASIC
extern "C" void reb_calculate_acceleration(struct reb_simulation* r) {
	if (r->gravity == reb_simulation::REB_GRAVITY_NONE) return;  // "Do nothing."
	require_supported(r);
	Dropin& d = get(r);
	ensure_dev(r, d);
	push_params(r, d);
	OK(ass_gravity_basic(d.sim));
	wrote(r, d, ASS_F_ACC);
}

// src/rebound.c:69-144 for the configuration every spring module runs (leapfrog, no tree, no boundary, no collisions)
extern "C" void reb_step(struct reb_simulation* const r) {
	require_supported(r);
	Dropin& d = get(r);
	{
		Sequence seq(r, d);
		reb_integrator_leapfrog_part1(r);   // :72
		reb_calculate_acceleration(r);      // :109
	}
	if (r->additional_forces) r->additional_forces(r);  // :114 -- the module's callback; it calls back into this layer
	reb_integrator_leapfrog_part2(r);       // :119
	if (r->post_timestep_modifications) {   // :120-125 (NULL in every shipped module)
		reb_integrator_synchronize(r);
		r->post_timestep_modifications(r);
	}
	// reb_boundary_check / reb_collision_search are no-ops for BOUNDARY_NONE / COLLISION_NONE (:131-143)
}

// src/rebound.c:621-704, REB_VISUALIZATION_NONE branch
extern "C" enum REB_STATUS reb_integrate(struct reb_simulation* const r, double tmax) {
	if (r->visualization != reb_simulation::REB_VISUALIZATION_NONE) {
		reb_error(r, "asteroid-spring-sims_b200 was built headless (the reference says the same when compiled without OPENGL).");
		return REB_EXIT_ERROR;
	}
	require_supported(r);
	Dropin& d = get(r);
	double last_full_dt = r->dt;      // :629
	r->dt_last_done = 0.;             // :630
	r->status = REB_RUNNING;          // :632
	d.in_run = true;
	d.dev_fresh = false;              // whatever the module did during set-up is on the host
	// Set-up helpers (subtract_com, spin_body, ...) sum in the reference's order so positions, hence rs0, keep their
	// bits; the per-step heartbeat (center_sim every step in every module) uses the parallel tree unless told not to.
	const char* exact = getenv("ASS_EXACT_SUMS");
	ass_set_sum_order(d.sim, (exact && *exact && strcmp(exact, "0") != 0) ? 1 : 0);
	const bool host_checks = r->exit_max_distance != 0 || r->exit_min_distance != 0;
	auto heartbeat = [&]() {
		if (host_checks) ass_dropin_sync_host(r);  // reb_run_heartbeat reads r->particles for the exit tests (:576-607)
		reb_run_heartbeat(r);
	};
	heartbeat();                      // :633
	while (reb_check_exit(r, tmax, &last_full_dt) < 0) {  // :634
		reb_step(r);
		heartbeat();
	}
	reb_integrator_synchronize(r);    // :649
	if (r->exact_finish_time == 1) r->dt = last_full_dt;  // :653-655
	ass_set_sum_order(d.sim, 1);
	ass_dropin_sync_host(r);
	d.in_run = false;
	unpin_particles(d);
	d.dev_fresh = false;
	return r->status;
}

// src/output.c:47-61.  Modules gate every piece of host-side output on this test, which makes it the natural place
// to bring the host arrays up to date in resident mode.
extern "C" int reb_output_check(struct reb_simulation* r, double interval) {
	const double shift = r->t;
	int fire = 0;
	if (floor(shift / interval) != floor((shift - r->dt) / interval)) fire = 1;
	if (r->t == 0) fire = 1;
	if (fire) ass_dropin_sync_host(r);
	return fire;
}

// ---------------------------------------------------------------------------------------------------
// C++ symbols of src_spring
// ---------------------------------------------------------------------------------------------------
// src_spring/springs.cpp:288-346
void spring_forces(reb_simulation* const r) {
	Dropin& d = get(r);
	if (lazy_ok(d)) {
		ensure_dev(r, d, false);
		if (d.lazy_springs) settle(r, d);  // a second evaluation in one step: the first one runs now
		ensure_springs(d);                  // the module's list as it is NOW
		d.lazy_springs = true;
		wrote(r, d, ASS_F_ACC);
		return;
	}
	ensure_dev(r, d);
	ensure_springs(d);
	OK(ass_spring_forces(d.sim));
	wrote(r, d, ASS_F_ACC);
}

// src_spring/physics.cpp:557-564
void zero_accel(reb_simulation* const r) {
	Dropin& d = get(r);
	if (lazy_ok(d)) {
		ensure_dev(r, d, false);
		if (d.lazy_springs) settle(r, d);  // zeroing AFTER a deferred evaluation: keep the order
		d.lazy_zero = true;
		wrote(r, d, ASS_F_ACC);
		return;
	}
	ensure_dev(r, d);
	OK(ass_zero_accel(d.sim));
	wrote(r, d, ASS_F_ACC);
}

// src_spring/springs.cpp:96-106
void add_spring_helper(spring spr) {
	if (num_springs != (int) springs.size()) throw "Spring count and size of spring array don't match.\n";
	springs.push_back(spr);
	num_springs++;
	topo_changed();
}

// src_spring/springs.cpp:35-45
void del_spring(int spring_index) {
	if (num_springs != (int) springs.size()) throw "Spring count and size of spring array don't match.\n";
	springs.erase(springs.begin() + spring_index);
	num_springs--;
	topo_changed();
}

// src_spring/springs.cpp:110-143
void connect_springs_dist(reb_simulation* const r, double max_dist, int i_low, int i_high, spring spr) {
	if (i_high <= i_low) return;  // :116-117
	if (num_springs != (int) springs.size()) throw "Spring count and size of spring array don't match.\n";
	Dropin& d = get(r);
	ensure_dev(r, d);
	ass_spring* fresh = nullptr;
	int n_new = 0;
	OK(ass_connect_springs_dist(d.sim, max_dist, i_low, i_high, spr.k, spr.gamma, (const ass_spring*) springs.data(), (int) springs.size(), &fresh, &n_new));
	springs.reserve(springs.size() + (size_t) n_new);
	for (int q = 0; q < n_new; q++) {
		spring s = spr;  // keeps any field a future `struct spring` may add
		s.k = fresh[q].k; s.gamma = fresh[q].gamma; s.rs0 = fresh[q].rs0;
		s.particle_1 = fresh[q].particle_1; s.particle_2 = fresh[q].particle_2;
		springs.push_back(s);
	}
	num_springs = (int) springs.size();
	ass_free(fresh);
	topo_changed();
}

// src_spring/springs.cpp:146-154
void set_gamma(double new_gamma) {
	for (int i = 0; i < num_springs; i++) springs[i].gamma = new_gamma;
	std::cout.precision(2);
	std::cout << "Spring damping coefficients set to " << new_gamma << std::endl;
	props_changed();
}

// src_spring/springs.cpp:157-164
void divide_gamma(double gamma_fac) {
	for (int i = 0; i < num_springs; i++) springs[i].gamma /= gamma_fac;
	std::cout.precision(2);
	std::cout << "All gammas divided by " << gamma_fac << std::endl;
	props_changed();
}

// src_spring/springs.cpp:167-196
void adjust_spring_props(reb_simulation* const r, double new_k, double new_gamma, double r_min, double r_max) {
	const int i_low = 0, i_high = r->N - num_perts;
	Vector CoM = compute_com(r, i_low, i_high);
	ass_dropin_sync_host(r);  // spr_mid reads r->particles
	int NC = 0;
	for (int i = 0; i < num_springs; i++) {
		Vector x_mid = spr_mid(r, springs[i], CoM);
		const double r_mid = x_mid.len();
		if ((r_mid >= r_min) && (r_mid <= r_max)) {
			springs[i].k = new_k;
			springs[i].gamma = new_gamma;
			NC++;
		}
	}
	std::cout.precision(3);
	std::cout << "adjust_spring_props: \n\t Number of springs changed: " << NC << "\n\t Fraction of springs changed: "
	          << (double) NC / (double) num_springs << std::endl;
	props_changed();
}

// src_spring/springs.cpp:199-213
void kill_springs(reb_simulation* const r) {
	const int i_low = 0, i_high = r->N - num_perts;
	for (int i = i_low; i < i_high; i++) {
		if (stresses[i].failing) {
			const int s = stresses[i].max_force_spring;
			springs[s].k = 0.0;
			springs[s].gamma = 0.0;
		}
	}
	props_changed();
}

// src_spring/physics.cpp:41-57 -- parallel fixed-order sum; equals the reference's sequential sum to rounding
Vector compute_com(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	double c[3];
	OK(ass_compute_com(d.sim, i_low, i_high, c));
	return Vector({ c[0], c[1], c[2] });
}

// src_spring/physics.cpp:60-74
Vector compute_cov(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	double c[3];
	OK(ass_compute_cov(d.sim, i_low, i_high, c));
	return Vector({ c[0], c[1], c[2] });
}

// src_spring/physics.cpp:98-109
void center_sim(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	OK(ass_center_sim(d.sim, i_low, i_high));
	wrote(r, d, ASS_F_POS);
}

// src_spring/physics.cpp:77-84
void subtract_com(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	OK(ass_subtract_com(d.sim, i_low, i_high));
	wrote(r, d, ASS_F_POS);
}

// src_spring/physics.cpp:88-94
void subtract_cov(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	OK(ass_subtract_cov(d.sim, i_low, i_high));
	wrote(r, d, ASS_F_VEL);
}

// src_spring/physics.cpp:387-402
void move_resolved(reb_simulation* const r, Vector dx, Vector dv, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	const double a[3] = { dx.getX(), dx.getY(), dx.getZ() }, b[3] = { dv.getX(), dv.getY(), dv.getZ() };
	OK(ass_move_resolved(d.sim, a, b, i_low, i_high));
	wrote(r, d, ASS_F_POS | ASS_F_VEL);
}

// ---------------------------------------------------------------------------------------------------
// src_spring/orb.cpp: the per-step terms of the phobos / phobos_deimos modules
// ---------------------------------------------------------------------------------------------------
// orb.cpp:311-366 (called from additional_forces after spring_forces, modules/phobos_deimos/problem.cpp:380-387)
void quadrupole_accel(reb_simulation* const r, double J2_p, double R_p, double phi_p, double theta_p, int i_p) {
	if (i_p >= r->N) return;
	if (J2_p == 0.0) return;
	Dropin& d = get(r);
	ensure_dev(r, d);
	push_params(r, d);
	OK(ass_quadrupole_accel(d.sim, J2_p, R_p, phi_p, theta_p, i_p));
	wrote(r, d, ASS_F_ACC);
}

// orb.cpp:253-302 (heartbeat of the phobos* modules)
void drift_resolved(reb_simulation* const r, double timestep, double inv_tau_a, double inv_tau_e, int i_part, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	push_params(r, d);
	OK(ass_drift_resolved(d.sim, timestep, inv_tau_a, inv_tau_e, i_part, i_low, i_high));
	wrote(r, d, ASS_F_VEL);
}

// orb.cpp:198-250
void drift_bin(reb_simulation* const r, double timestep, double inv_tau_a, double inv_tau_e, int part_1, int part_2) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	push_params(r, d);
	OK(ass_drift_bin(d.sim, timestep, inv_tau_a, inv_tau_e, part_1, part_2));
	wrote(r, d, ASS_F_VEL);
}

// orb.cpp:438-449
double sum_mass(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	double m = 0.0;
	OK(ass_sum_mass(d.sim, i_low, i_high > i_low ? i_high : i_low, &m));
	return m;
}

// src_spring/stress.cpp:32-98.  The O(NS) part -- every node's sum of +-outer(F, L) and its strongest spring -- runs on the
// device and comes back with the reference's bits; the 3x3 eigenvalues stay the reference's own code, including its
// behaviour on tensors that are not exactly symmetric (eigenvalues() throws a `const char*` that the `catch (char*)`
// below, copied from the reference, does not match: the exception leaves update_stress, as it does there).
void update_stress(reb_simulation* const r) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	ensure_springs(d);
	const int N = r->N;
	if (!stresses.size()) stresses.resize(N);
	std::vector<double> T((size_t) 9 * N), F((size_t) N);
	std::vector<int> S((size_t) N);
	OK(ass_stress(d.sim, T.data(), F.data(), S.data()));
	for (int i = 0; i < N; i++) {
		const double* t = T.data() + (size_t) 9 * i;
		stresses[i].stress = Matrix({ { t[0], t[1], t[2] }, { t[3], t[4], t[5] }, { t[6], t[7], t[8] } });
		stresses[i].eigs[0] = 0.0; stresses[i].eigs[1] = 0.0; stresses[i].eigs[2] = 0.0;
		stresses[i].max_force = F[(size_t) i];
		stresses[i].max_force_spring = S[(size_t) i];
		stresses[i].failing = false;
	}
	for (int i = 0; i < N; i++) {
		try {
			eigenvalues(stresses[i].stress, stresses[i].eigs);
		} catch (char* str) {
			std::cerr << "Error: " << str << " Exiting." << std::endl;
			exit(1);
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// the rest of src_spring/physics.cpp that set-up code and write_resolved_with_E (output_spring.cpp:112-192) call
// ---------------------------------------------------------------------------------------------------
namespace {
// out[20] of ass_diagnostics for the current device state; one evaluation serves measure_L, compute_rot_kin,
// spring_potential_energy, dEdt_total and mom_inertia when they are called back to back on an unchanged state
const double* diagnostics(reb_simulation* r, int i_low, int i_high, bool want_gpe) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	if (num_springs > 0 || springs.size() > 0) ensure_springs(d);
	if (d.diag_version == d.version && d.diag_lo == i_low && d.diag_hi == i_high && (d.diag_gpe || !want_gpe)) return d.diag;
	push_params(r, d);
	OK(ass_diagnostics(d.sim, i_low, i_high, want_gpe ? 1 : 0, d.diag));
	d.diag_version = d.version; d.diag_lo = i_low; d.diag_hi = i_high; d.diag_gpe = want_gpe;
	return d.diag;
}
// spring sums do not depend on the particle range: reuse whatever range was evaluated last
const double* spring_diagnostics(reb_simulation* r) {
	Dropin& d = get(r);
	if (d.diag_version == d.version && d.diag_lo >= 0) {
		ensure_dev(r, d);
		ensure_springs(d);
		if (d.diag_version == d.version) return d.diag;
	}
	return diagnostics(r, 0, r->N - num_perts > 0 ? r->N - num_perts : r->N, false);
}
}  // namespace

// physics.cpp:184-210
void spin_body(reb_simulation* const r, int i_low, int i_high, Vector omega) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	const double w[3] = { omega.getX(), omega.getY(), omega.getZ() };
	OK(ass_spin_body(d.sim, i_low, i_high, w));
	wrote(r, d, ASS_F_VEL);
}

// physics.cpp:136-162
Vector measure_L(reb_simulation* const r, int i_low, int i_high) {
	if (i_high <= i_low) return Vector(zero_vec);
	const double* q = diagnostics(r, i_low, i_high, false);
	return Vector({ q[6], q[7], q[8] });
}

// physics.cpp:116-134
Matrix mom_inertia(reb_simulation* const r, int i_low, int i_high) {
	if (i_high <= i_low) return Matrix(zero_mat);
	const double* q = diagnostics(r, i_low, i_high, false);  // [13..18] xx yy zz xy xz yz
	return Matrix({ { q[13], q[16], q[17] }, { q[16], q[14], q[18] }, { q[17], q[18], q[15] } });
}

// physics.cpp:164-181
Vector body_spin(reb_simulation* const r, int i_low, int i_high) {
	Matrix inv_mom_inert = inverse(mom_inertia(r, i_low, i_high));
	return inv_mom_inert * measure_L(r, i_low, i_high);
}

// physics.cpp:508-528
double compute_rot_kin(reb_simulation* const r, int i_low, int i_high) {
	if (i_high <= i_low) return 0.0;
	return diagnostics(r, i_low, i_high, false)[9];
}

// physics.cpp:457-476
double spring_potential_energy(reb_simulation* const r) {
	if (num_springs <= 0 || r->N <= 0) return 0.0;
	return spring_diagnostics(r)[10];
}

// physics.cpp:409-454
double dEdt_total(reb_simulation* const r) {
	if (num_springs <= 0 || r->N <= 0) return 0.0;
	return spring_diagnostics(r)[12];
}

// physics.cpp:479-504 -- O(N^2): as expensive as a force evaluation; on the device it reuses the gravity tiling
double grav_potential_energy(reb_simulation* const r, int i_low, int i_high) {
	if (i_high <= i_low) return 0.0;
	return diagnostics(r, i_low, i_high, true)[11];
}

// physics.cpp:249-290
void rotate_body(reb_simulation* const r, int i_low, int i_high, double alpha, double beta, double gamma) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	// Rz2 * Rx * Rz1, composed by the reference's own Matrix arithmetic (left to right, as in :274)
	const Matrix R = getRotMatZ(gamma) * getRotMatX(beta) * getRotMatZ(alpha);
	const double m[9] = { R.getXX(), R.getXY(), R.getXZ(), R.getYX(), R.getYY(), R.getYZ(), R.getZX(), R.getZY(), R.getZZ() };
	OK(ass_transform_body(d.sim, i_low, i_high, m, 1));
	wrote(r, d, ASS_F_POS | ASS_F_VEL);
}

// physics.cpp:294-331
void rotate_origin(reb_simulation* const r, int i_low, int i_high, double alpha, double beta, double gamma) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	const Matrix R = getRotMatZ(gamma) * getRotMatX(beta) * getRotMatZ(alpha);
	const double m[9] = { R.getXX(), R.getXY(), R.getXZ(), R.getYX(), R.getYY(), R.getYZ(), R.getZX(), R.getZY(), R.getZZ() };
	OK(ass_transform_body(d.sim, i_low, i_high, m, 0));
	wrote(r, d, ASS_F_POS | ASS_F_VEL);
}

// physics.cpp:334-385: the 3x3 eigenproblem stays the reference's own code (matrix_math.cpp); the O(N) rotation runs on the device
void rotate_to_principal(reb_simulation* const r, int i_low, int i_high) {
	Matrix inertia_mat = mom_inertia(r, i_low, i_high);
	double eigs[3];
	eigenvalues(inertia_mat, eigs);
	Vector eigvec1 = eigenvector(inertia_mat, eigs[2]);  // note the order (:346-349)
	Vector eigvec2 = eigenvector(inertia_mat, eigs[1]);
	Vector eigvec3 = eigenvector(inertia_mat, eigs[0]);
	Matrix R = { eigvec1, eigvec2, eigvec3 };
	const double m[9] = { R.getXX(), R.getXY(), R.getXZ(), R.getYX(), R.getYY(), R.getYZ(), R.getZX(), R.getZY(), R.getZZ() };
	Dropin& d = get(r);
	ensure_dev(r, d);
	OK(ass_transform_body(d.sim, i_low, i_high, m, 0));
	wrote(r, d, ASS_F_POS | ASS_F_VEL);
}

// src_spring/shapes.cpp:814-833
double mindist(reb_simulation* const r, int i_low, int i_high) {
	Dropin& d = get(r);
	ensure_dev(r, d);
	double out = 0.0;
	OK(ass_mindist(d.sim, i_low, i_high, &out));
	return out;
}

namespace {
// append packed generator rows [n0, n_new) to the simulation, as the reference's reb_add loop does
void add_rows(reb_simulation* r, const double* rows, int n_from, int n_to) {
	for (int i = n_from; i < n_to; i++) {
		reb_particle pt;
		memset(&pt, 0, sizeof(pt));
		const double* q = rows + (size_t) i * 11;
		pt.x = q[0]; pt.y = q[1]; pt.z = q[2];
		pt.m = q[9]; pt.r = q[10];
		reb_add(r, pt);
	}
}
// existing particles as packed rows (rand_shape reads the shape model's vertices from the simulation)
double* rows_of(reb_simulation* r, int n, int* cap) {
	*cap = n > 0 ? n : 1;
	double* rows = (double*) calloc((size_t) *cap * 11, sizeof(double));
	for (int i = 0; i < n; i++) {
		const reb_particle& p = r->particles[i];
		double* q = rows + (size_t) i * 11;
		q[0] = p.x; q[1] = p.y; q[2] = p.z; q[3] = p.vx; q[4] = p.vy; q[5] = p.vz; q[9] = p.m; q[10] = p.r;
	}
	return rows;
}
}  // namespace

// src_spring/shapes.cpp:66-128
void rand_rectangle(reb_simulation* const r, double min_dist, double x, double y, double z, double total_mass) {
	const int n_part = 40.0 * x * y * z / pow(min_dist, 3.0);
	std::cout << "rand_rectangle: Trying to create " << n_part << " particles." << std::endl;
	ass_dropin_sync_host(r);
	double* rows = nullptr;
	int cap = 0;
	const int n_new = ass_rand_rectangle(&rows, &cap, 0, min_dist, x, y, z, total_mass);
	if (n_new < 0) fatal("ass_rand_rectangle");
	const int i_0 = r->N;
	add_rows(r, rows, 0, n_new);
	ass_free(rows);
	ass_dropin_host_modified(r);
	const double min_d = mindist(r, i_0, r->N);
	std::cout << "rand_rectangle: Created " << r->N - i_0 << " particles separated by minimum distance " << min_d << std::endl;
}

// src_spring/shapes.cpp:132-206
void rand_cone(reb_simulation* const r, double min_dist, double radius, double height, double total_mass) {
	const int n_part = 40 * pow(2.0 * radius / min_dist, 3.0);
	std::cout << "rand_cone: Trying to create " << n_part << " particles." << std::endl;
	ass_dropin_sync_host(r);
	double* rows = nullptr;
	int cap = 0;
	const int n_new = ass_rand_cone(&rows, &cap, 0, min_dist, radius, height, total_mass);
	if (n_new < 0) fatal("ass_rand_cone");
	const int i_0 = r->N;
	add_rows(r, rows, 0, n_new);
	ass_free(rows);
	ass_dropin_host_modified(r);
	const double min_d = mindist(r, i_0, r->N);
	std::cout << "rand_cone: Created " << r->N - i_0 << " particles separated by minimum distance " << min_d << std::endl;
}

// src_spring/shapes.cpp:435-516: the shape model's vertices are particles [0, N) already
void rand_shape(reb_simulation* const r, double min_dist, double total_mass) {
	ass_dropin_sync_host(r);
	const int N_shape = r->N;
	const double max_r_shape = max_radius(r, 0, N_shape);
	const int n_part = 40 * pow(2.0 * max_r_shape / min_dist, 3.0);
	std::cout << "rand_shape: Trying to create " << n_part << " particles." << std::endl;
	int cap = 0;
	double* rows = rows_of(r, N_shape, &cap);
	const int n_new = ass_rand_shape(&rows, &cap, N_shape, min_dist, total_mass);
	if (n_new < 0) fatal("ass_rand_shape");
	add_rows(r, rows, N_shape, n_new);
	free(rows);
	ass_dropin_host_modified(r);
	const double min_d = mindist(r, N_shape, r->N);
	std::cout << "rand_shape: Created " << r->N - N_shape << " particles separated by a minimum of " << min_d << std::endl;
}

// src_spring/shapes.cpp:282-352
void hcp_ellipsoid(reb_simulation* r, double min_dist, double x, double y, double z, double total_mass) {
	ass_dropin_sync_host(r);
	double* rows = nullptr;
	int cap = 0;
	const int n_new = ass_hcp_ellipsoid(&rows, &cap, 0, min_dist, x, y, z, total_mass);
	if (n_new < 0) fatal("ass_hcp_ellipsoid");
	const int i_0 = r->N;
	add_rows(r, rows, 0, n_new);
	const double min_d = n_new > 0 ? rows[10] * 4.0 : 0.0;  // r = min_d / 4 (:342)
	ass_free(rows);
	ass_dropin_host_modified(r);
	std::cout << "hcp_ellipsoid: Created " << r->N - i_0 << " particles with lattice spacing " << 1.08 * min_dist
	          << ", separated by minimum distance " << (n_new > 1 ? mindist(r, i_0, r->N) : min_d) << std::endl;
}

// src_spring/shapes.cpp:356-431
void cubic_ellipsoid(reb_simulation* r, double min_dist, double x, double y, double z, double total_mass) {
	ass_dropin_sync_host(r);
	double* rows = nullptr;
	int cap = 0;
	const int n_new = ass_cubic_ellipsoid(&rows, &cap, 0, min_dist, x, y, z, total_mass);
	if (n_new < 0) fatal("ass_cubic_ellipsoid");
	const int i_0 = r->N;
	add_rows(r, rows, 0, n_new);
	ass_free(rows);
	ass_dropin_host_modified(r);
	std::cout << "cubic_ellipsoid: Created " << r->N - i_0 << " particles, separated by minimum distance "
	          << (n_new > 1 ? mindist(r, i_0, r->N) : 0.0) << std::endl;
}

// src_spring/shapes.cpp:210-278
void rand_ellipsoid(reb_simulation* const r, double min_dist, double x, double y, double z, double total_mass) {
	const int n_part = 40 * pow(2.0 * x / min_dist, 3.0);
	std::cout << "rand_ellipsoid: Trying to create " << n_part << " particles." << std::endl;
	ass_dropin_sync_host(r);
	double* rows = nullptr;
	int cap = 0;
	const int n_new = ass_rand_ellipsoid(&rows, &cap, 0, min_dist, x, y, z, total_mass);
	if (n_new < 0) fatal("ass_rand_ellipsoid");
	const int i_0 = r->N;
	add_rows(r, rows, 0, n_new);
	ass_free(rows);
	ass_dropin_host_modified(r);
	const double min_d = mindist(r, i_0, r->N);
	std::cout << "rand_ellipsoid: Created " << r->N - i_0 << " particles separated by minimum distance " << min_d << std::endl;
}
