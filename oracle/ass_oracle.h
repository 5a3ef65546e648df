/* oracle/ass_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's per-step force + integration path and of its
 * build-time neighbour queries.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * it; the product (asteroid-spring-sims_b200/) never does.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (spring lists, generator output,
 * accelerations, leapfrog state) or to 1e-13 (diagnostic sums) against
 *   (1) the reference's own known-answer tests (modules/tests/tests.cpp SpringTest.* / PhysicsTest.*), and
 *   (2) outputs of the unmodified reference compiled here (oracle/_ref, see oracle/Makefile) -- live when
 *       oracle/_ref exists, and through the committed fixtures in tests/golden/ otherwise.
 *
 * Particle state is a packed array of rows of ORC_STRIDE doubles: x y z vx vy vz ax ay az m r, i.e. the first
 * eleven fields of `struct reb_particle` (reference src/rebound.h:540-551).
 */
#ifndef ASS_ORACLE_H
#define ASS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_STRIDE 11
enum { ORC_X = 0, ORC_Y, ORC_Z, ORC_VX, ORC_VY, ORC_VZ, ORC_AX, ORC_AY, ORC_AZ, ORC_M, ORC_R };

/* == struct spring, reference src_spring/springs.h:16-22 (32 bytes) */
typedef struct {
	double k, rs0, gamma;
	int particle_1, particle_2;
} orc_spring;

/* Simulation scalars that the path reads from struct reb_simulation (src/rebound.h:589-836). */
typedef struct {
	double t, dt, dt_last_done;
	double G, softening;
	int gravity_basic;     /* 1: REB_GRAVITY_BASIC, 0: REB_GRAVITY_NONE */
	int n_active;          /* -1 => all (src/gravity.c:64) */
	int zero_first;        /* additional_forces = zero_accel + spring_forces (modules/asteroid_spin/problem.cpp:225-231) */
	int center_heartbeat;  /* heartbeat = center_sim(0, N - n_perts)  (modules/asteroid_spin/problem.cpp:196) */
	int n_perts;
	int status;            /* enum REB_STATUS, src/rebound.h:358-368 */
	int exact_finish_time;
} orc_sim;

/* ---- per-step path ---- */
void orc_zero_accel(double* p, int n);                                            /* physics.cpp:557-564 */
void orc_spring_forces(double* p, const orc_spring* s, int ns);                   /* springs.cpp:288-346 */
void orc_gravity_basic(double* p, int n, int n_active, double G, double softening);/* gravity.c:57-125  */
/* same sum, restricted to targets i in [i_lo, i_hi) -- bounded CPU-baseline samples */
void orc_gravity_basic_slab(double* p, int n, int n_active, double G, double softening, int i_lo, int i_hi);
void orc_leapfrog_part1(double* p, int n, orc_sim* sim);                          /* integrator_leapfrog.c:39-51 */
void orc_leapfrog_part2(double* p, int n, orc_sim* sim);                          /* integrator_leapfrog.c:52-67 */
void orc_step(double* p, int n, const orc_spring* s, int ns, orc_sim* sim);       /* rebound.c:69-144 */
void orc_heartbeat(double* p, int n, orc_sim* sim);
int orc_integrate(double* p, int n, const orc_spring* s, int ns, orc_sim* sim, double tmax); /* rebound.c:509-704 */

/* ---- spring build ---- */
/* Literal restatement: O(N^2) pair scan + O(NS) duplicate scan per add (springs.cpp:52-143). */
int orc_connect_springs_dist_n2(const double* p, double max_dist, int i_low, int i_high, double k, double gamma,
		orc_spring** springs, int* ns, int* cap);
/* Results-identical uniform-grid version for N > 1e4. */
int orc_connect_springs_dist(const double* p, double max_dist, int i_low, int i_high, double k, double gamma,
		orc_spring** springs, int* ns, int* cap);
int orc_add_spring(const double* p, int i, int j, double k, double gamma, orc_spring** springs, int* ns, int* cap);
void orc_free(void* ptr);

/* ---- node generators (shapes.cpp) ---- */
/* Both return the new particle count; *p is realloc'd.  rand() stream = libc, seed with srand() first. */
int orc_rand_ellipsoid_n2(double** p, int n0, double min_dist, double a, double b, double c, double total_mass); /* shapes.cpp:210-278 literal */
int orc_rand_ellipsoid(double** p, int n0, double min_dist, double a, double b, double c, double total_mass);    /* cell-list */
int orc_hcp_ellipsoid(double** p, int n0, double min_dist, double a, double b, double c, double total_mass);     /* shapes.cpp:282-352 */
int orc_cubic_ellipsoid(double** p, int n0, double min_dist, double a, double b, double c, double total_mass);   /* shapes.cpp:356-431 */
double orc_mindist_n2(const double* p, int i_low, int i_high);                                                    /* shapes.cpp:814-833 */
double orc_mindist(const double* p, int i_low, int i_high);                                                       /* grid */
double orc_random_uniform(double lo, double hi);                                                                  /* tools.c:44-46 */

/* ---- heartbeat helpers / diagnostics (physics.cpp, orb.cpp) ---- */
double orc_sum_mass(const double* p, int i_low, int i_high);
void orc_compute_com(const double* p, int i_low, int i_high, double* out3);
void orc_compute_cov(const double* p, int i_low, int i_high, double* out3);
void orc_center_sim(double* p, int n, int i_low, int i_high);
void orc_move_resolved(double* p, const double* dx, const double* dv, int i_low, int i_high);
void orc_subtract_com(double* p, int i_low, int i_high);
void orc_subtract_cov(double* p, int i_low, int i_high);
void orc_spin_body(double* p, int i_low, int i_high, double ox, double oy, double oz);
void orc_set_gamma(orc_spring* s, int ns, double g);
/* out[20]: same layout as ref_diagnostics in oracle/ref_shim.cpp */
void orc_diagnostics(const double* p, int i_low, int i_high, const orc_spring* s, int ns, double G, int want_gpe, double* out);

/* ---- the other per-step terms of the phobos* modules (orb.cpp) ---- */
void orc_quadrupole_accel(double* p, int n, double G, double J2, double R, double phi, double theta, int i_p);     /* orb.cpp:311-366 */
void orc_drift_resolved(double* p, double G, double timestep, double inv_tau_a, double inv_tau_e, int i_part, int i_low, int i_high); /* orb.cpp:253-302 */
void orc_drift_bin(double* p, double G, double timestep, double inv_tau_a, double inv_tau_e, int part_1, int part_2);                 /* orb.cpp:198-250 */

/* ---- generator variants (shapes.cpp), literal O(candidates x N) restatements ---- */
int orc_rand_rectangle_n2(double** p, int n0, double min_dist, double x, double y, double z, double total_mass);  /* shapes.cpp:66-128 */
int orc_rand_cone_n2(double** p, int n0, double min_dist, double radius, double height, double total_mass);       /* shapes.cpp:132-206 */
int orc_rand_shape_n2(double** p, int n0, double min_dist, double total_mass);                                    /* shapes.cpp:435-516; rows [0, n0) = vertices */

/* ---- set-up rotations (physics.cpp:249-331) ---- */
void orc_euler_matrix(double alpha, double beta, double gamma, double* out9);  /* getRotMatZ(gamma) * getRotMatX(beta) * getRotMatZ(alpha) */
void orc_rotate_body(double* p, int i_low, int i_high, double alpha, double beta, double gamma);
void orc_rotate_origin(double* p, int i_low, int i_high, double alpha, double beta, double gamma);

/* ---- stress (stress.cpp:32-85): tensors n x 9, max_force n, max_spring n ---- */
void orc_stress(const double* p, int n, const orc_spring* s, int ns, double* tensors, double* max_force, int* max_spring);

/* ---- module-local per-step force terms (not library symbols): restated from the two modules that define them ---- */
/* apply_impact, modules/bennu2/problem.cpp:366-460 (force part); is_surf: n bytes; returns num_pushed */
int orc_apply_impact(double* p, int n_body, const unsigned char* is_surf, double lat, double longi, double dangle, double force, double envelope);
/* top_push / table_top, modules/cube/problem.cpp:205-264, with the faces already marked */
void orc_top_push(double* p, int n, const unsigned char* top, double pressure);
void orc_table_top(double* p, int n, const unsigned char* bottom);

#ifdef __cplusplus
}
#endif
#endif
