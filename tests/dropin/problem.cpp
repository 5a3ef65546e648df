// tests/dropin/problem.cpp -- a problem module written against the REFERENCE's API only (rebound.h, springs.h,
// shapes.h, physics.h, libconfig), in the shape of modules/asteroid_spin/problem.cpp and
// modules/wobble_bennu2/problem.cpp: read problem.cfg, build a random ellipsoid, connect springs, spin it up, run
// reb_integrate with an additional_forces callback and a per-step heartbeat.
//
// It exists because no shipped module can run as shipped (missing particle files, cfg keys that do not
// match the lookups).  The same source is linked twice by tests/dropin/Makefile:
//   synth_ref    against the unmodified reference objects                      -> the CPU answer
//   synth_b200   against the same objects with the hot-path symbols replaced by asteroid-spring-sims_b200/host/dropin.cpp
// and tests/test_dropin.py compares what the two write.  This file contains no ASS_/ass_ symbol.
#ifdef __GNUC__
#define restrict __restrict__
#endif

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>
extern "C" {
#include "rebound.h"
}
#include "libconfig.h++"
#include "matrix_math.h"
#include "stress.h"
#include "springs.h"
#include "physics.h"
#include "shapes.h"
#include "kepcart.h"
#include "orb.h"
#include "input_spring.h"
#include "output_spring.h"

using namespace libconfig;
using std::string;
using std::vector;

// Globals every module defines (modules/asteroid_spin/problem.cpp:37-47)
int num_springs = 0;
vector<spring> springs;
int num_perts = 0;
double def_gamma, t_damp, print_interval;
string fileroot;
double mass_scale, time_scale, length_scale, temp_scale, omega_scale, vel_scale, p_scale, L_scale, a_scale, F_scale,
		E_scale, dEdt_scale, P_scale;
int use_gravity = 0;
// phobos / phobos_deimos shape (modules/phobos_deimos/problem.cpp): one central point mass with a J2, migration drift
int point_mass = 0;
int shape = 0;   // 0 rand_ellipsoid, 1 hcp_ellipsoid, 2 cubic_ellipsoid, 3 rand_rectangle, 4 rand_cone
int restart_index = -1;  // >= 0: read <restart_root>_<index>_particles.txt / _springs.txt (modules/asteroid_spin/problem.cpp:104-118)
string restart_root;
int stress = 0;  // 1: update_stress after the run, tensors written to <fileroot>_stress.txt
int rotate = 0;  // 1: rotate_body + rotate_to_principal between spin-up and the run (modules/wobble_bennu2/problem.cpp:300-312)
double J2_plus = 1.96e-3, R_plus = 0.5, theta_plus = 0.4, phi_plus = 0.3, inv_tau_a = 1e-2, inv_tau_e = 3e-2;

void additional_forces(reb_simulation *n_body_sim);
void heartbeat(reb_simulation *const n_body_sim);

int main(int argc, char *argv[]) {
	Config cfg;
	cfg.readFile("problem.cfg");
	read_scales(&cfg);

	double t_max, dt, max_spring_dist, gamma_fac, k, omega_in[3], min_dist, mass_ball, r_ball, ratio1, ratio2;
	int seed = 1;
	if (!(cfg.lookupValue("fileroot", fileroot) && cfg.lookupValue("dt", dt) && cfg.lookupValue("t_max", t_max)
			&& cfg.lookupValue("print_interval", print_interval) && cfg.lookupValue("max_spring_dist", max_spring_dist)
			&& cfg.lookupValue("k", k) && cfg.lookupValue("gamma", def_gamma) && cfg.lookupValue("damp_fac", gamma_fac)
			&& cfg.lookupValue("t_damp", t_damp) && cfg.lookupValue("omega_x", omega_in[0])
			&& cfg.lookupValue("omega_y", omega_in[1]) && cfg.lookupValue("omega_z", omega_in[2])
			&& cfg.lookupValue("min_particle_distance", min_dist) && cfg.lookupValue("mass_ball", mass_ball)
			&& cfg.lookupValue("radius_ball", r_ball) && cfg.lookupValue("ratio1", ratio1)
			&& cfg.lookupValue("ratio2", ratio2))) {
		throw "Failed to read in problem.cfg. Exiting.";
	}
	cfg.lookupValue("seed", seed);
	cfg.lookupValue("gravity", use_gravity);
	cfg.lookupValue("point_mass", point_mass);
	cfg.lookupValue("shape", shape);
	cfg.lookupValue("rotate", rotate);
	cfg.lookupValue("stress", stress);
	cfg.lookupValue("restart_index", restart_index);
	cfg.lookupValue("restart_root", restart_root);
	Vector omega(omega_in);

	reb_simulation *const n_body_sim = reb_create_simulation();
	srand(seed);  // reb_create_simulation reseeds from the clock (src/rebound.c:369): "same seed" = srand after it

	n_body_sim->integrator = reb_simulation::REB_INTEGRATOR_LEAPFROG;
	n_body_sim->gravity = use_gravity ? reb_simulation::REB_GRAVITY_BASIC : reb_simulation::REB_GRAVITY_NONE;
	n_body_sim->boundary = reb_simulation::REB_BOUNDARY_NONE;
	n_body_sim->visualization = reb_simulation::REB_VISUALIZATION_NONE;
	n_body_sim->G = 1.0;
	n_body_sim->additional_forces = additional_forces;
	n_body_sim->heartbeat = heartbeat;
	n_body_sim->dt = dt;
	n_body_sim->softening = min_dist / 100.0;  // modules/wobble_bennu2/problem.cpp:115

	spring default_spring;
	default_spring.gamma = gamma_fac * def_gamma;
	default_spring.k = k;

	if (restart_index >= 0) {
		// asteroid_spin's restart path: a body and its springs from the reference's own text dumps
		read_particles(n_body_sim, restart_root, restart_index);
		read_springs(restart_root, restart_index);
	} else
	switch (shape) {
	case 1: hcp_ellipsoid(n_body_sim, min_dist, r_ball, r_ball * ratio1, r_ball * ratio2, mass_ball); break;
	case 2: cubic_ellipsoid(n_body_sim, min_dist, r_ball, r_ball * ratio1, r_ball * ratio2, mass_ball); break;
	case 3: rand_rectangle(n_body_sim, min_dist, 2.0 * r_ball, 2.0 * r_ball * ratio1, 2.0 * r_ball * ratio2, mass_ball); break;  // modules/cube
	case 4: rand_cone(n_body_sim, min_dist, r_ball, 1.5 * r_ball, mass_ball); break;
	default: rand_ellipsoid(n_body_sim, min_dist, r_ball, r_ball * ratio1, r_ball * ratio2, mass_ball);
	}
	const int i_low = 0, i_high = n_body_sim->N;
	if (restart_index < 0) {
		subtract_com(n_body_sim, i_low, i_high);
		subtract_cov(n_body_sim, i_low, i_high);
		spin_body(n_body_sim, i_low, i_high, omega);
		connect_springs_dist(n_body_sim, max_spring_dist, i_low, i_high, default_spring);
	}
	if (rotate) {
		rotate_body(n_body_sim, i_low, i_high, 0.3, -0.2, 1.1);
		rotate_to_principal(n_body_sim, i_low, i_high);
	}
	if (point_mass) {
		// modules/phobos_deimos/problem.cpp:246-262: a Mars-like mass on a slightly eccentric, inclined orbit about the body
		OrbitalElements orb_el;
		orb_el.a = 7.0; orb_el.e = 0.02; orb_el.i = 0.1; orb_el.long_asc_node = 0.3; orb_el.arg_peri = 0.2; orb_el.mean_anom = 1.0;
		add_pt_mass_kep(n_body_sim, i_low, i_high, -1, 1.0, 0.5, orb_el);
		num_perts = 1;
	}

	std::ofstream runfile(fileroot + "_run.txt", std::ios::out | std::ios::trunc);
	runfile << std::hexfloat << "N " << n_body_sim->N << "\nNS " << num_springs << "\nmean_rs0 " << mean_spring_length() << "\n";
	if (restart_index < 0) {  // (the reference's mindist is O(N^2): not on a body of restart size)
		runfile << "mindist " << mindist(n_body_sim, i_low, i_high) << "\n";
		for (int i = 0; i < num_springs; i++)
			runfile << springs[i].particle_1 << " " << springs[i].particle_2 << " " << springs[i].rs0 << "\n";
	}
	runfile.close();
	std::ofstream(fileroot + "_ext.txt", std::ios::out | std::ios::trunc).close();

	if (t_max == 0.0) t_max = INFINITY;
	// a read-only helper before the clock starts: on the drop-in this is where the CUDA context comes up (0.3 - 1 s, once per
	// process), which is not part of what reb_integrate costs
	(void) compute_com(n_body_sim, i_low, i_high);
	const auto wall0 = std::chrono::steady_clock::now();
	reb_integrate(n_body_sim, t_max);
	const double wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
	std::ofstream(fileroot + "_timing.txt", std::ios::out | std::ios::trunc) << "integrate_s " << wall_s << "\nN " << n_body_sim->N << "\nt " << n_body_sim->t << "\n";

	std::ofstream fin(fileroot + "_final.txt", std::ios::out | std::ios::trunc);
	fin << std::hexfloat << "t " << n_body_sim->t << "\ndt " << n_body_sim->dt << "\nstatus " << n_body_sim->status << "\n";
	for (int i = 0; i < n_body_sim->N; i++) {
		const reb_particle &p = n_body_sim->particles[i];
		fin << p.x << " " << p.y << " " << p.z << " " << p.vx << " " << p.vy << " " << p.vz << " " << p.ax << " " << p.ay
				<< " " << p.az << " " << p.m << "\n";
	}
	if (stress) {
		// update_stress (stress.cpp:32-98) throws out of its eigenvalue stage on any real body (a tensor summed in floating
		// point is not EXACTLY symmetric, matrix_math.cpp:273-276); the tensors themselves are complete by then
		extern vector<stress_tensor> stresses;
		int threw = 0;
		try {
			update_stress(n_body_sim);
		} catch (const char *) {
			threw = 1;
		}
		std::ofstream sf(fileroot + "_stress.txt", std::ios::out | std::ios::trunc);
		sf << std::hexfloat << "threw " << threw << "\n";
		for (int i = 0; i < n_body_sim->N; i++) {
			const Matrix &m = stresses[i].stress;
			sf << m.getXX() << " " << m.getXY() << " " << m.getXZ() << " " << m.getYX() << " " << m.getYY() << " " << m.getYZ() << " "
					<< m.getZX() << " " << m.getZY() << " " << m.getZZ() << " " << stresses[i].max_force << " " << stresses[i].max_force_spring
					<< "\n";
		}
	}
	return 0;
}

// modules/asteroid_spin/problem.cpp:225-231 (GRAVITY_NONE needs zero_accel) / modules/wobble_bennu2 (springs only)
void additional_forces(reb_simulation *n_body_sim) {
	if (!use_gravity)
		zero_accel(n_body_sim);
	spring_forces(n_body_sim);
	if (point_mass)  // modules/phobos_deimos/problem.cpp:384-386
		quadrupole_accel(n_body_sim, J2_plus, R_plus, phi_plus, theta_plus, n_body_sim->N - num_perts);
}

// modules/asteroid_spin/problem.cpp:181-207
void heartbeat(reb_simulation *const n_body_sim) {
	static string extendedfile = fileroot + "_ext.txt";
	if (std::abs(n_body_sim->t - t_damp) < 0.9 * n_body_sim->dt)
		set_gamma(def_gamma);
	center_sim(n_body_sim, 0, n_body_sim->N - num_perts);
	if (point_mass) {  // modules/phobos_deimos/problem.cpp:322-344
		const double migfac = exp(-n_body_sim->t / 10.0);
		drift_resolved(n_body_sim, n_body_sim->dt, inv_tau_a * migfac, inv_tau_e * migfac, n_body_sim->N - num_perts, 0,
				n_body_sim->N - num_perts);
	}
	if (reb_output_check(n_body_sim, print_interval)) {
		const int i_low = 0, i_high = n_body_sim->N - num_perts;
		Vector L = measure_L(n_body_sim, i_low, i_high);
		const double KE = compute_rot_kin(n_body_sim, i_low, i_high);
		const double SPE = spring_potential_energy(n_body_sim);
		const double GPE = use_gravity ? grav_potential_energy(n_body_sim, i_low, i_high) : 0.0;
		std::ofstream out(extendedfile, std::ios::out | std::ios::app);
		out << std::setprecision(17) << n_body_sim->t << " " << L.getX() << " " << L.getY() << " " << L.getZ() << " " << KE
				<< " " << SPE << " " << GPE << " " << dEdt_total(n_body_sim) << "\n";
		// the module's own history file, through the reference's writer (output_spring.cpp:112-192): t, CoM, CoV, omega, L, I,
		// KErot, PEspr, PEgrav, Etot, dE/dt -- every number of it comes from a function the drop-in replaces
		write_resolved_with_E(n_body_sim, i_low, i_high, fileroot + "_res.txt", 0.0);
	}
}
