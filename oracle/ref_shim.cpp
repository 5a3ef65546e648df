// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI handle onto the UNMODIFIED reference (adebrecht661/asteroid-spring-sims) built
// from the sources where they lie under /root/reference (see oracle/Makefile; outputs go to
// oracle/_ref/ only, which is git-ignored).  Nothing here re-implements any arithmetic: every
// entry point forwards to the reference's own function.  Used by
//   * tests/  -- to pin the C restatement in oracle/ass_oracle.c (and the CUDA path) against the
//                real reference, and to generate the golden fixtures under tests/golden/;
//   * bench.py's cpu_baseline / --impl reference leg -- to time the reference's own CPU path.
// The product (asteroid-spring-sims_b200/) never links or loads this file.
//
// The shim plays the role of a reference "problem module": like modules/*/problem.cpp it owns the
// globals `springs`, `num_springs`, `num_perts` and the *_scale doubles that src_spring/*.cpp
// reads through `extern` (src_spring/springs.cpp:22-25, src_spring/input_spring.cpp).

// rebound.h uses C99 `restrict`; every reference .cpp opens with this same define (src_spring/springs.cpp:1-7).
#define restrict __restrict__

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <iostream>
extern "C" {
#include "rebound.h"
// Not in rebound.h's public section but exported by the C library (src/integrator_leapfrog.h).
void reb_integrator_leapfrog_part1(struct reb_simulation* r);
void reb_integrator_leapfrog_part2(struct reb_simulation* r);
void reb_calculate_acceleration(struct reb_simulation* r);   // src/gravity.h
}
#include "matrix_math.h"
#include "springs.h"
#include "shapes.h"
#include "physics.h"
#include "stress.h"
#include "orb.h"

using std::vector;

// ---- globals a reference module must define (modules/asteroid_spin/problem.cpp:37-47) ----
int num_springs = 0;
vector<spring> springs;
int num_perts = 0;
double mass_scale = 1, time_scale = 1, length_scale = 1, temp_scale = 1, omega_scale = 1,
       vel_scale = 1, p_scale = 1, L_scale = 1, a_scale = 1, F_scale = 1, E_scale = 1,
       dEdt_scale = 1, P_scale = 1;

namespace {
int g_zero_first = 0;   // additional_forces = zero_accel + spring_forces (asteroid_spin) or spring_forces only
int g_center_hb = 0;    // heartbeat = center_sim(0, N-num_perts) (asteroid_spin/problem.cpp:196)

void shim_additional_forces(reb_simulation* r) {
	if (g_zero_first) zero_accel(r);          // modules/asteroid_spin/problem.cpp:227
	spring_forces(r);                          // modules/asteroid_spin/problem.cpp:230
}
void shim_heartbeat(reb_simulation* r) {
	if (g_center_hb) center_sim(r, 0, r->N - num_perts);
}
}

extern "C" {

// The reference chats on std::cout (set_gamma, rand_ellipsoid, ...): mute it by failing the stream.
void ref_quiet(int on) {
	if (on) std::cout.setstate(std::ios_base::failbit);
	else std::cout.clear();
}

void* ref_new(void) {
	reb_simulation* r = reb_create_simulation();
	r->integrator = reb_simulation::REB_INTEGRATOR_LEAPFROG;
	r->gravity = reb_simulation::REB_GRAVITY_NONE;
	r->boundary = reb_simulation::REB_BOUNDARY_NONE;
	r->collision = reb_simulation::REB_COLLISION_NONE;
	r->visualization = reb_simulation::REB_VISUALIZATION_NONE;
	r->additional_forces = shim_additional_forces;
	r->heartbeat = shim_heartbeat;
	r->G = 1.0;
	return r;
}
void ref_free(void* h) {
	reb_free_simulation((reb_simulation*) h);
	vector<spring>().swap(springs);
	num_springs = 0;
	num_perts = 0;
}
// "Same seed" = srand AFTER reb_create_simulation (src/rebound.c:369 reseeds from the clock).
void ref_seed(unsigned s) { srand(s); }

void ref_config(void* h, int gravity_basic, double G, double softening, double dt,
		int zero_first, int center_heartbeat, int n_perts) {
	reb_simulation* r = (reb_simulation*) h;
	r->gravity = gravity_basic ? reb_simulation::REB_GRAVITY_BASIC : reb_simulation::REB_GRAVITY_NONE;
	r->G = G;
	r->softening = softening;
	r->dt = dt;
	g_zero_first = zero_first;
	g_center_hb = center_heartbeat;
	num_perts = n_perts;
}
// r->N_active: number of gravity sources (src/gravity.c:64,89); -1 = all.  Used to time a bounded sample of the O(N^2) sum.
void ref_set_n_active(void* h, int n_active) { ((reb_simulation*) h)->N_active = n_active; }
double ref_time(void* h) { return ((reb_simulation*) h)->t; }
void ref_set_time(void* h, double t) { ((reb_simulation*) h)->t = t; }
double ref_dt(void* h) { return ((reb_simulation*) h)->dt; }
int ref_n(void* h) { return ((reb_simulation*) h)->N; }

int ref_rand_ellipsoid(void* h, double min_dist, double a, double b, double c, double total_mass) {
	rand_ellipsoid((reb_simulation*) h, min_dist, a, b, c, total_mass);
	return ((reb_simulation*) h)->N;
}
int ref_hcp_ellipsoid(void* h, double min_dist, double a, double b, double c, double total_mass) {
	hcp_ellipsoid((reb_simulation*) h, min_dist, a, b, c, total_mass);
	return ((reb_simulation*) h)->N;
}
int ref_cubic_ellipsoid(void* h, double min_dist, double a, double b, double c, double total_mass) {
	cubic_ellipsoid((reb_simulation*) h, min_dist, a, b, c, total_mass);
	return ((reb_simulation*) h)->N;
}
int ref_rand_rectangle(void* h, double min_dist, double x, double y, double z, double total_mass) {
	rand_rectangle((reb_simulation*) h, min_dist, x, y, z, total_mass);   // shapes.cpp:66-128
	return ((reb_simulation*) h)->N;
}
int ref_rand_cone(void* h, double min_dist, double radius, double height, double total_mass) {
	rand_cone((reb_simulation*) h, min_dist, radius, height, total_mass);   // shapes.cpp:132-206
	return ((reb_simulation*) h)->N;
}
int ref_rand_shape(void* h, double min_dist, double total_mass) {
	rand_shape((reb_simulation*) h, min_dist, total_mass);   // shapes.cpp:435-516; the vertices are the particles already added
	return ((reb_simulation*) h)->N;
}
double ref_mindist(void* h, int lo, int hi) { return mindist((reb_simulation*) h, lo, hi); }

int ref_add_particle(void* h, const double* p /* x y z vx vy vz ax ay az m r */) {
	reb_simulation* r = (reb_simulation*) h;
	reb_particle pt;
	memset(&pt, 0, sizeof(pt));
	pt.x = p[0]; pt.y = p[1]; pt.z = p[2];
	pt.vx = p[3]; pt.vy = p[4]; pt.vz = p[5];
	pt.ax = p[6]; pt.ay = p[7]; pt.az = p[8];
	pt.m = p[9]; pt.r = p[10];
	reb_add(r, pt);
	return r->N;
}
// Packed rows of 11 doubles: x y z vx vy vz ax ay az m r  (reb_particle's first 11 fields, rebound.h:540-551)
void ref_get_particles(void* h, double* out) {
	reb_simulation* r = (reb_simulation*) h;
	for (int i = 0; i < r->N; i++) memcpy(out + 11 * (size_t) i, &r->particles[i], 11 * sizeof(double));
}
void ref_set_particles(void* h, const double* in) {
	reb_simulation* r = (reb_simulation*) h;
	for (int i = 0; i < r->N; i++) memcpy(&r->particles[i], in + 11 * (size_t) i, 11 * sizeof(double));
}
int ref_sizeof_particle(void) { return (int) sizeof(reb_particle); }
int ref_sizeof_spring(void) { return (int) sizeof(spring); }
int ref_sizeof_simulation(void) { return (int) sizeof(reb_simulation); }

// ---- springs ----
int ref_connect_springs_dist(void* h, double max_dist, int lo, int hi, double k, double gamma) {
	spring spr;
	memset(&spr, 0, sizeof(spr));
	spr.k = k; spr.gamma = gamma;
	connect_springs_dist((reb_simulation*) h, max_dist, lo, hi, spr);
	return num_springs;
}
int ref_add_spring(void* h, int i, int j, double k, double rs0, double gamma) {
	spring spr;
	memset(&spr, 0, sizeof(spr));
	spr.k = k; spr.rs0 = rs0; spr.gamma = gamma;
	try { return add_spring((reb_simulation*) h, i, j, spr); } catch (const char*) { return -1; }
}
int ref_num_springs(void) { return num_springs; }
void ref_get_springs(spring* out) { if (num_springs) memcpy(out, springs.data(), sizeof(spring) * (size_t) num_springs); }
void ref_set_springs(const spring* in, int n) {
	springs.assign(in, in + n);
	num_springs = n;
}
void ref_set_gamma(double g) { set_gamma(g); }
void ref_divide_gamma(double f) { divide_gamma(f); }
void ref_adjust_spring_props(void* h, double k, double g, double rmin, double rmax) {
	adjust_spring_props((reb_simulation*) h, k, g, rmin, rmax);
}
double ref_mean_spring_length(void) { return mean_spring_length(); }
double ref_strain(void* h, int s) { return strain((reb_simulation*) h, springs[s]); }
void ref_spring_i_force(void* h, int s, int damped, double* out3) {
	Vector f = damped ? spring_i_force((reb_simulation*) h, s) : spring_i_force_undamped((reb_simulation*) h, s);
	out3[0] = f.getX(); out3[1] = f.getY(); out3[2] = f.getZ();
}

// ---- the per-step path, piecewise and whole ----
void ref_zero_accel(void* h) { zero_accel((reb_simulation*) h); }
void ref_spring_forces(void* h) { spring_forces((reb_simulation*) h); }
void ref_gravity(void* h) { reb_calculate_acceleration((reb_simulation*) h); }
void ref_leapfrog_part1(void* h) { reb_integrator_leapfrog_part1((reb_simulation*) h); }
void ref_leapfrog_part2(void* h) { reb_integrator_leapfrog_part2((reb_simulation*) h); }
void ref_step(void* h, int n) {
	reb_simulation* r = (reb_simulation*) h;
	for (int s = 0; s < n; s++) reb_step(r);
}
// reb_step + heartbeat, as reb_integrate_raw does per iteration (src/rebound.c:640-641)
void ref_step_hb(void* h, int n) {
	reb_simulation* r = (reb_simulation*) h;
	for (int s = 0; s < n; s++) { reb_step(r); shim_heartbeat(r); }
}
int ref_integrate(void* h, double tmax) { return (int) reb_integrate((reb_simulation*) h, tmax); }

// ---- heartbeat helpers and diagnostics (src_spring/physics.cpp) ----
void ref_center_sim(void* h, int lo, int hi) { center_sim((reb_simulation*) h, lo, hi); }
void ref_subtract_com(void* h, int lo, int hi) { subtract_com((reb_simulation*) h, lo, hi); }
void ref_subtract_cov(void* h, int lo, int hi) { subtract_cov((reb_simulation*) h, lo, hi); }
void ref_move_resolved(void* h, const double* dx, const double* dv, int lo, int hi) {
	move_resolved((reb_simulation*) h, Vector({ dx[0], dx[1], dx[2] }), Vector({ dv[0], dv[1], dv[2] }), lo, hi);
}
void ref_spin_body(void* h, int lo, int hi, double ox, double oy, double oz) {
	spin_body((reb_simulation*) h, lo, hi, Vector({ ox, oy, oz }));
}
void ref_rotate_body(void* h, int lo, int hi, double alpha, double beta, double gamma) {
	rotate_body((reb_simulation*) h, lo, hi, alpha, beta, gamma);   // physics.cpp:249-290
}
void ref_rotate_origin(void* h, int lo, int hi, double alpha, double beta, double gamma) {
	rotate_origin((reb_simulation*) h, lo, hi, alpha, beta, gamma);   // physics.cpp:294-331
}
void ref_rotate_to_principal(void* h, int lo, int hi) { rotate_to_principal((reb_simulation*) h, lo, hi); }   // physics.cpp:334-385
// the matrix rotate_body / rotate_origin apply (Rz2 * Rx * Rz1, left to right), row-major
void ref_euler_matrix(double alpha, double beta, double gamma, double* out9) {
	Matrix R = getRotMatZ(gamma) * getRotMatX(beta) * getRotMatZ(alpha);
	out9[0] = R.getXX(); out9[1] = R.getXY(); out9[2] = R.getXZ();
	out9[3] = R.getYX(); out9[4] = R.getYY(); out9[5] = R.getYZ();
	out9[6] = R.getZX(); out9[7] = R.getZY(); out9[8] = R.getZZ();
}
// out[0..2] CoM, [3..5] CoV, [6..8] L (measure_L), [9] KE (compute_rot_kin), [10] SPE,
// [11] GPE (only if want_gpe, else 0), [12] dEdt_total, [13..18] I xx yy zz xy xz yz, [19] sum_mass
void ref_diagnostics(void* h, int lo, int hi, int want_gpe, double* out) {
	reb_simulation* r = (reb_simulation*) h;
	Vector c = compute_com(r, lo, hi), v = compute_cov(r, lo, hi), L = measure_L(r, lo, hi);
	out[0] = c.getX(); out[1] = c.getY(); out[2] = c.getZ();
	out[3] = v.getX(); out[4] = v.getY(); out[5] = v.getZ();
	out[6] = L.getX(); out[7] = L.getY(); out[8] = L.getZ();
	out[9] = compute_rot_kin(r, lo, hi);
	out[10] = spring_potential_energy(r);
	out[11] = want_gpe ? grav_potential_energy(r, lo, hi) : 0.0;
	out[12] = dEdt_total(r);
	Matrix I = mom_inertia(r, lo, hi);
	out[13] = I.getXX(); out[14] = I.getYY(); out[15] = I.getZZ();
	out[16] = I.getXY(); out[17] = I.getXZ(); out[18] = I.getYZ();
	out[19] = sum_mass(r, lo, hi);
}

// ---- the other per-step force / heartbeat terms of the phobos* modules (src_spring/orb.cpp) ----
void ref_quadrupole_accel(void* h, double J2, double R, double phi, double theta, int i_p) {
	quadrupole_accel((reb_simulation*) h, J2, R, phi, theta, i_p);   // orb.cpp:311-366
}
void ref_drift_resolved(void* h, double timestep, double inv_tau_a, double inv_tau_e, int i_part, int lo, int hi) {
	drift_resolved((reb_simulation*) h, timestep, inv_tau_a, inv_tau_e, i_part, lo, hi);   // orb.cpp:253-302
}
void ref_drift_bin(void* h, double timestep, double inv_tau_a, double inv_tau_e, int part_1, int part_2) {
	drift_bin((reb_simulation*) h, timestep, inv_tau_a, inv_tau_e, part_1, part_2);   // orb.cpp:198-250
}

// update_stress (stress.cpp:32-98).  Returns 0 on success, 1 if the reference threw (its eigenvalues() throws a
// `const char*` when the summed stress matrix is not EXACTLY symmetric, matrix_math.cpp:273-276, which update_stress's own
// `catch (char*)` does not match); out[i*4 + 0..2] = eigenvalues, out[i*4 + 3] = max_force.
int ref_update_stress(void* h, double* out) {
	reb_simulation* r = (reb_simulation*) h;
	extern std::vector<stress_tensor> stresses;
	stresses.clear();
	try {
		update_stress(r);
	} catch (const char*) {
		return 1;
	}
	for (int i = 0; i < r->N; i++) {
		out[4 * i + 0] = stresses[i].eigs[0]; out[4 * i + 1] = stresses[i].eigs[1]; out[4 * i + 2] = stresses[i].eigs[2];
		out[4 * i + 3] = stresses[i].max_force;
	}
	return 0;
}

// The per-node stress tensors (stress.cpp:32-85: sum of +-outer(F, L) over the node's springs, divided by the node
// volume) and the strongest spring per node, whether or not the eigenvalue stage that follows them threw.
// tensors: N x 9 row-major; returns 1 if update_stress threw.
int ref_stress(void* h, double* tensors, double* max_force, int* max_spring) {
	reb_simulation* r = (reb_simulation*) h;
	extern std::vector<stress_tensor> stresses;
	stresses.clear();
	int threw = 0;
	try {
		update_stress(r);
	} catch (const char*) {
		threw = 1;
	}
	for (int i = 0; i < r->N; i++) {
		const Matrix& m = stresses[i].stress;
		double* t = tensors + 9 * (size_t) i;
		t[0] = m.getXX(); t[1] = m.getXY(); t[2] = m.getXZ();
		t[3] = m.getYX(); t[4] = m.getYY(); t[5] = m.getYZ();
		t[6] = m.getZX(); t[7] = m.getZY(); t[8] = m.getZZ();
		max_force[i] = stresses[i].max_force;
		max_spring[i] = stresses[i].max_force_spring;
	}
	return threw;
}

}  // extern "C"
