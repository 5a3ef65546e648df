/* include/ass_b200.h -- the C-ABI boundary of the B200-native hot path.
 *
 * One shared library, libass_b200.so (built from asteroid-spring-sims_b200/csrc/ for sm_100a), exports exactly the
 * symbols declared here.  Plain pointers and sizes only: no C++ types, no torch types, no reference headers.  Every
 * entry point names the reference function it replaces (paths relative to the reference tree).  Host code above this
 * line stays C++ (asteroid-spring-sims_b200/host/): it mirrors the reference's own symbols (reb_integrate, reb_step,
 * spring_forces, connect_springs_dist, ...) and forwards here.  See INTEGRATION.md.
 *
 * Conventions
 *   - All functions return 0 on success, a negative ASS_E* code on failure; ass_last_error() gives the text.
 *     Nothing throws across this boundary.  There is NO CPU fallback: without a CUDA device every compute call
 *     fails with ASS_ENODEVICE.
 *   - A simulation (`ass_sim`) owns device-resident SoA copies of the particle and spring state.  The host
 *     `struct reb_particle[]` / `std::vector<spring>` stay the API-visible truth; the device copy is a cache that the
 *     caller fills (ass_upload_*) and drains (ass_download_*) explicitly.
 *   - Host particle arrays are addressed as (base, stride_bytes): row i starts at base + i*stride and begins with the
 *     eleven doubles x y z vx vy vz ax ay az m r -- the leading fields of `struct reb_particle`
 *     (src/rebound.h:540-551; sizeof == 128).  A packed (N,11) double matrix is stride 88.
 *   - Work is queued on the simulation's own CUDA stream; calls that return data to the host synchronise it.
 */
#ifndef ASS_B200_H
#define ASS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASS_API_VERSION 1

/* ---- error codes ---- */
#define ASS_OK 0
#define ASS_EINVAL (-1)    /* bad argument (also: spring with particle_1 == particle_2, springs.cpp:57-59) */
#define ASS_ECUDA (-2)     /* CUDA runtime error; text in ass_last_error() */
#define ASS_ENODEVICE (-3) /* no usable sm_100 device */
#define ASS_ENOMEM (-4)
#define ASS_ESTATE (-5)    /* call order violated (e.g. forces before upload) */
#define ASS_ECOMM (-6)     /* multi-GPU exchange failed */

/* ---- field masks for upload / download ---- */
#define ASS_F_POS 1u
#define ASS_F_VEL 2u
#define ASS_F_ACC 4u
#define ASS_F_MASS 8u /* m and r */
#define ASS_F_ALL 15u

/* == `struct spring`, src_spring/springs.h:16-22: 32 bytes, same field order, so a std::vector<spring>::data()
 * pointer can be passed straight through. */
typedef struct ass_spring {
	double k;       /* spring constant */
	double rs0;     /* rest length */
	double gamma;   /* damping coefficient */
	int particle_1; /* lower node index  */
	int particle_2; /* higher node index */
} ass_spring;

/* The scalars of `struct reb_simulation` (src/rebound.h:589-836) that the path reads or writes. */
typedef struct ass_params {
	double t;            /* r->t */
	double dt;           /* r->dt */
	double dt_last_done; /* r->dt_last_done */
	double G;            /* r->G */
	double softening;    /* r->softening */
	int gravity;         /* 0 = REB_GRAVITY_NONE, 1 = REB_GRAVITY_BASIC (src/rebound.h:771-776) */
	int n_active;        /* r->N_active; -1 = all particles are sources (src/gravity.c:64) */
	int n_perts;         /* module global num_perts: trailing point masses that carry no springs */
	int additional_forces; /* what ass_step runs where reb_step calls r->additional_forces (src/rebound.c:114):
	                          0 nothing, 1 spring_forces, 2 zero_accel + spring_forces
	                          (modules/asteroid_spin/problem.cpp:225-231) */
	int heartbeat;       /* what ass_integrate runs where reb_integrate_raw calls the heartbeat (src/rebound.c:633,641):
	                          0 nothing, 1 center_sim(0, N - n_perts) (modules/asteroid_spin/problem.cpp:196) */
	int exact_finish_time; /* r->exact_finish_time (src/rebound.c:519) */
	int status;          /* enum REB_STATUS (src/rebound.h:358-368); written by ass_integrate */
} ass_params;

typedef struct ass_sim ass_sim; /* opaque */

/* ---- library / device ---- */
int ass_api_version(void);
const char* ass_last_error(void);
/* Select the CUDA device for this process (one process per GPU). Fails with ASS_ENODEVICE when there is none. */
int ass_init(int device);
int ass_device_count(void);
/* sm count, L2 bytes, total memory bytes, max SM clock (kHz), compute capability major*10+minor */
int ass_device_info(int* sm_count, long long* l2_bytes, long long* mem_bytes, int* sm_clock_khz, int* cc);

/* ---- lifecycle ---- */
int ass_create(ass_sim** out);
int ass_destroy(ass_sim* sim);
int ass_sync(ass_sim* sim);
int ass_set_params(ass_sim* sim, const ass_params* p);
int ass_get_params(ass_sim* sim, ass_params* p);

/* ---- state transfer (replaces nothing: it is the residency boundary, SURVEY section 8b "Ownership") ---- */
int ass_upload_particles(ass_sim* sim, const void* base, size_t stride_bytes, int n, unsigned fields);
int ass_download_particles(ass_sim* sim, void* base, size_t stride_bytes, int n, unsigned fields);
/* Page-lock and map a host range so the two calls above move rows at PCIe link rate: uploads are read by the copy
 * engine straight from the caller's rows, downloads are written by the pack kernel straight into them (no bounce buffer,
 * no host-side scatter).  ass_host_register pins memory the caller already owns -- e.g. r->particles (malloc'd by
 * reb_add, src/particle.c:53-56) for the length of a reb_integrate; unregister BEFORE the memory is freed or realloc'd.
 * ass_host_alloc / ass_host_free hand out pinned memory directly.  Unpinned arrays keep working (driver-staged). */
int ass_host_register(void* base, size_t bytes);
int ass_host_unregister(void* base);
int ass_host_alloc(void** out, size_t bytes);
int ass_host_free(void* p);
/* Replace the whole spring list (after connect_springs_dist / add_spring / del_spring). Builds the node-sorted
 * half-edge (CSR) layout used by ass_spring_forces.  Checks 0 <= particle_1 != particle_2 < n. */
int ass_set_springs(ass_sim* sim, const ass_spring* springs, int ns);
/* Same topology, new k / rs0 / gamma (set_gamma, divide_gamma, adjust_spring_props, kill_springs;
 * src_spring/springs.cpp:146-213). */
int ass_update_spring_props(ass_sim* sim, const ass_spring* springs, int ns);
int ass_num_particles(ass_sim* sim);
int ass_num_springs(ass_sim* sim);

/* ---- the per-step path, one reference function each ---- */
int ass_zero_accel(ass_sim* sim);     /* zero_accel, src_spring/physics.cpp:557-564 */
int ass_spring_forces(ass_sim* sim);  /* spring_forces, src_spring/springs.cpp:288-346: a += springs */
int ass_gravity_basic(ass_sim* sim);  /* reb_calculate_acceleration, REB_GRAVITY_BASIC, src/gravity.c:72-103: a = gravity.
                                         With params.gravity == 0 this is the NONE branch: a no-op. */
/* Which kernel evaluates the direct sum: 0 = choose by size (default), 1 = one-sided source tiles (17 FP64 instructions
 * per ordered pair), 2 = pair-once / Newton's third law (10.5 per ordered pair; needs N_active == N). Same result to
 * rounding; both deterministic. */
int ass_set_gravity_kernel(ass_sim* sim, int which);
int ass_leapfrog_part1(ass_sim* sim); /* reb_integrator_leapfrog_part1, src/integrator_leapfrog.c:39-51 */
int ass_leapfrog_part2(ass_sim* sim); /* reb_integrator_leapfrog_part2, src/integrator_leapfrog.c:52-67 */
/* [part1 ;] [zero_accel ;] spring_forces ; part2 of ONE step as the fused kernels run them (same bits as the separate calls):
 * for a caller that is handed the calls one by one and may defer them until part2 (the drop-in in resident mode). */
int ass_substep_fused(ass_sim* sim, int do_part1, int zero_first);
/* reb_step (src/rebound.c:69-144) x nsteps with the built-in additional_forces of params.additional_forces, fused
 * (gravity -> springs+kick+drift) and without returning to the host between steps.  with_heartbeat != 0 runs the
 * built-in heartbeat (params.heartbeat) after each step, as reb_integrate_raw does (src/rebound.c:640-641).
 * Where nothing but the spring forces separates the steps (params.gravity == 0, or additional_forces == 2, and no
 * heartbeat due) the whole batch runs as ONE cooperative launch with a grid barrier per step; same bits either way. */
int ass_step(ass_sim* sim, int nsteps, int with_heartbeat);
/* reb_integrate (src/rebound.c:621-704): heartbeat at t=0, then steps until tmax with reb_check_exit's
 * exact_finish_time handling (src/rebound.c:509-569).  Returns the REB_STATUS (>= 0) or a negative error. */
int ass_integrate(ass_sim* sim, double tmax);

/* ---- per-step heartbeat helpers (src_spring/physics.cpp) ---- */
/* How the mass-weighted sums behind compute_com / compute_cov (and everything built on them) are accumulated.  Until this
 * is called the helpers below sum sequentially and the BUILT-IN per-step heartbeat of ass_step / ass_integrate uses the
 * tree; a call fixes one order for both.
 * sequential != 0 (default for the helpers): in the reference's particle order, one rounding per operation -> bit-identical CoM / CoV,
 *                            hence bit-identical subtract_com / subtract_cov / spin_body / center_sim results.  One CTA,
 *                            ~8 cycles per particle: meant for set-up, where an ulp of position changes rs0 bits.
 * sequential == 0:           fixed-order parallel tree; equal to rounding; for per-step heartbeats at large N. */
int ass_set_sum_order(ass_sim* sim, int sequential);
int ass_center_sim(ass_sim* sim, int i_low, int i_high);      /* physics.cpp:98-109 */
int ass_subtract_com(ass_sim* sim, int i_low, int i_high);    /* physics.cpp:77-84  */
int ass_subtract_cov(ass_sim* sim, int i_low, int i_high);    /* physics.cpp:88-94  */
int ass_move_resolved(ass_sim* sim, const double dx[3], const double dv[3], int i_low, int i_high); /* physics.cpp:387-402 */
int ass_compute_com(ass_sim* sim, int i_low, int i_high, double out[3]); /* physics.cpp:41-57 */
int ass_compute_cov(ass_sim* sim, int i_low, int i_high, double out[3]); /* physics.cpp:60-74 */
int ass_spin_body(ass_sim* sim, int i_low, int i_high, const double omega[3]); /* physics.cpp:184-210 */
/* x <- R (x - c) + c, v <- R (v - u) + u for particles [i_low, i_high), R a row-major 3x3 matrix: about_com != 0 takes
 * c, u = CoM, CoV of the range (rotate_body, physics.cpp:249-290), about_com == 0 the origin (rotate_origin :294-331,
 * rotate_to_principal :334-385: the host composes R -- Euler matrices, eigenvectors -- the device applies it). */
int ass_transform_body(ass_sim* sim, int i_low, int i_high, const double R[9], int about_com);
/* out[20]: [0..2] CoM, [3..5] CoV, [6..8] L = measure_L (physics.cpp:136-162), [9] compute_rot_kin (:508-528),
 * [10] spring_potential_energy (:457-476), [11] grav_potential_energy (:479-504; 0 unless want_gpe),
 * [12] dEdt_total (:409-454), [13..18] mom_inertia xx yy zz xy xz yz (:116-134), [19] sum_mass (orb.cpp:438-449) */
int ass_diagnostics(ass_sim* sim, int i_low, int i_high, int want_gpe, double out[20]);

/* ---- stress (src_spring/stress.cpp) ---- */
/* The first half of update_stress, stress.cpp:32-85: per node the sum of +outer(F, L) (node is particle_1) / -outer(F, L)
 * (particle_2) over its springs in spring-index order, F = spring_i_force (springs.cpp:350-392), L = spring_r, divided
 * by the node volume 4 pi / (3 N); and the strongest spring per node (first index of the largest |F|; -1: none).
 * tensors: N x 9 row-major.  The eigenvalue stage (:88-97) is 3x3 host work and stays the caller's. */
int ass_stress(ass_sim* sim, double* tensors, double* max_force, int* max_force_spring);

/* ---- masked force terms of two modules' own additional_forces (module-local code, offered so that resident mode
 *      covers them): node masks 0..3 live on the device ---- */
int ass_set_node_mask(ass_sim* sim, int mask_id, const unsigned char* mask /* N bytes, non-zero = member */);
/* mask = particles of [i_low, i_high) with coordinate `axis` > threshold (greater != 0) or < threshold; *count = members
 * (modules/cube/problem.cpp:217-227, 247-257: the top / bottom faces, marked once) */
int ass_mask_from_plane(ass_sim* sim, int mask_id, int i_low, int i_high, int axis, double threshold, int greater, int* count);
/* apply_impact, modules/bennu2/problem.cpp:366-460 (force part): masked particles of [i_low, i_high) within angle
 * `dangle` of the direction (lat, longi) get  a -= force * x_hat / (m * num_pushed * |x|) * envelope, num_pushed
 * counted first; the caller passes envelope = 1 - cos(2 pi t / tau).  *num_pushed may be NULL (no host round trip). */
int ass_apply_radial_impulse(ass_sim* sim, int mask_id, int i_low, int i_high, double lat, double longi, double dangle, double force,
		double envelope, int* num_pushed);
/* top_push, modules/cube/problem.cpp:230-234: a[axis] -= per_node_force / m for masked particles (per_node_force =
 * pressure / N_surf);  table_top, :259-263: a[axis] = 0 for masked particles */
int ass_plane_push(ass_sim* sim, int mask_id, int axis, double per_node_force);
int ass_plane_constraint(ass_sim* sim, int mask_id, int axis);

/* ---- the other per-step terms of the phobos / phobos_deimos modules (src_spring/orb.cpp) ---- */
/* quadrupole_accel, orb.cpp:311-366: a += J2 acceleration from particle i_p (with params.G), reaction onto i_p.
 * No-op for i_p >= N or J2_p == 0, as in the reference. */
int ass_quadrupole_accel(ass_sim* sim, double J2_p, double R_p, double phi_p, double theta_p, int i_p);
/* drift_resolved, orb.cpp:253-302: orbital migration between point mass i_part and the resolved body [i_low, i_high):
 * velocities only, momentum of the pair conserved.  Entirely on the device (no host round trip). */
int ass_drift_resolved(ass_sim* sim, double timestep, double inv_tau_a, double inv_tau_e, int i_part, int i_low, int i_high);
/* drift_bin, orb.cpp:198-250: the same between two point masses */
int ass_drift_bin(ass_sim* sim, double timestep, double inv_tau_a, double inv_tau_e, int part_1, int part_2);
int ass_sum_mass(ass_sim* sim, int i_low, int i_high, double* out); /* orb.cpp:438-449 */
/* ass_download_particles for rows [i_low, i_high) only: `base` is row 0 of the caller's array (e.g. to read one point
 * mass for output while the body stays resident). */
int ass_download_range(ass_sim* sim, void* base, size_t stride_bytes, int i_low, int i_high, unsigned fields);

/* ---- spring build: connect_springs_dist, src_spring/springs.cpp:110-143 ---- */
/* Uniform-grid cell-list search on the device positions of particles [i_low, i_high).  Appends to the caller's list
 * exactly what the reference would: pairs (i<j) with sqrt(fma(dz,dz,fma(dx,dx,dy*dy))) < max_dist (strict), in
 * lexicographic (i,j) order, rs0 = that length bit for bit, k/gamma from the prototype, pairs already in
 * existing[0..n_existing) skipped (add_spring's duplicate rule, springs.cpp:70-79).
 * On return *out points at a malloc'd array of *n_out NEW springs (caller frees with ass_free). */
int ass_connect_springs_dist(ass_sim* sim, double max_dist, int i_low, int i_high, double k, double gamma,
		const ass_spring* existing, int n_existing, ass_spring** out, int* n_out);
/* mindist, src_spring/shapes.cpp:814-833, by the same grid */
int ass_mindist(ass_sim* sim, int i_low, int i_high, double* out);
void ass_free(void* p);

/* ---- node generators (host side of the build path; sequential by construction: one libc rand() stream and an
 *      accept test that depends on every earlier accept).  src_spring/shapes.cpp:210-278, 282-352, 356-431 ---- */
/* Appends to a packed (n,11) row array (realloc'd through *rows / *cap_rows); returns the new count or <0. */
int ass_rand_ellipsoid(double** rows, int* cap_rows, int n0, double min_dist, double a, double b, double c, double total_mass);
int ass_hcp_ellipsoid(double** rows, int* cap_rows, int n0, double min_dist, double a, double b, double c, double total_mass);
int ass_cubic_ellipsoid(double** rows, int* cap_rows, int n0, double min_dist, double a, double b, double c, double total_mass);
/* rand_rectangle (shapes.cpp:66-128), rand_cone (:132-206), rand_shape (:435-516; rows [0, n0) hold the shape model's
 * vertices).  They share one loop, also exported: n_candidates darts over the box [lo, hi] (x, y, z drawn in that order
 * from libc rand()), kept if `inside` (NULL: always) accepts them and no particle made by this call is within min_dist. */
typedef int (*ass_inside_fn)(void* ctx, double x, double y, double z);
int ass_rand_fill(double** rows, int* cap_rows, int n0, long long n_candidates, const double lo[3], const double hi[3],
		ass_inside_fn inside, void* ctx, double min_dist, double radius, double total_mass);
int ass_rand_rectangle(double** rows, int* cap_rows, int n0, double min_dist, double x, double y, double z, double total_mass);
int ass_rand_cone(double** rows, int* cap_rows, int n0, double min_dist, double radius, double height, double total_mass);
int ass_rand_shape(double** rows, int* cap_rows, int n0, double min_dist, double total_mass);

/* ---- ensembles: B independent small bodies, one CTA per body, no collective (BASELINE C5) ---- */
typedef struct ass_ensemble ass_ensemble;
/* Bodies are given concatenated: body b owns particle rows [node_off[b], node_off[b+1]) of `base` and springs
 * [spring_off[b], spring_off[b+1]) whose particle indices are LOCAL to the body (0 .. n_b-1).  params[b] per body
 * (dt, G, softening, gravity, additional_forces, heartbeat, n_perts, t).  A body must fit in shared memory (~2400 nodes);
 * larger bodies use ass_sim.  The per-step arithmetic is that of ass_step, member by member. */
int ass_ensemble_create(ass_ensemble** out, int n_bodies, const int* node_off, const int* spring_off,
		const void* base, size_t stride_bytes, const ass_spring* springs, const ass_params* params);
int ass_ensemble_set_params(ass_ensemble* e, const ass_params* params);
/* reb_step x nsteps for every member in ONE launch; *device_ms (may be NULL) receives the CUDA-event time (synchronises). */
int ass_ensemble_step(ass_ensemble* e, int nsteps, int with_heartbeat, double* device_ms);
int ass_ensemble_download(ass_ensemble* e, void* base, size_t stride_bytes, unsigned fields);
int ass_ensemble_get_params(ass_ensemble* e, ass_params* params);
int ass_ensemble_sync(ass_ensemble* e);
int ass_ensemble_destroy(ass_ensemble* e);

/* ---- multi-GPU: ONE large body, particles split into contiguous slabs, one rank (process, GPU) per slab (SURVEY 8e,
 *      BASELINE C4).  Every rank uploads the whole body and the whole spring list, then calls ass_shard_setup.  Rank g
 *      integrates, and sums the springs onto, particles [bounds[g], bounds[g+1]) only.  Gravity is divided by pairs of
 *      slabs so that every unordered particle pair is evaluated once and every rank does the same work; a rank leaves
 *      the partial accelerations it computed for other slabs in its exported window.  All transfers are pulls through
 *      peer-mapped memory (CUDA IPC over NVLink) by the copy engines: the peers' x and v while the SMs already work on
 *      the own slab (the analogue of one all-gather), the partials before the springs (one reduce-scatter).
 *      The host is not in the step: ass_shard_step queues whole steps and returns; the ranks meet on the device through
 *      64-bit epoch flags that producers store into their peers' windows and consumers poll locally (csrc/shard.cu).
 *      The launcher (asteroid-spring-sims_b200/launcher.py) only swaps the window handles once. ---- */
#define ASS_IPC_HANDLE_BYTES 64
int ass_shard_setup(ass_sim* sim, int rank, int nranks /* <= 16 */, const int* bounds /* nranks + 1 */);
int ass_shard_export(ass_sim* sim, unsigned char handle[ASS_IPC_HANDLE_BYTES]);
int ass_shard_import(ass_sim* sim, int peer_rank, const unsigned char handle[ASS_IPC_HANDLE_BYTES]);
/* same-process peers (several ass_sim in one process): share the pointer instead of an IPC handle */
int ass_shard_import_local(ass_sim* sim, int peer_rank, ass_sim* peer);
/* reb_step (src/rebound.c:69-144) x nsteps for the own slab, with the built-in additional_forces of
 * params.additional_forces.  Collective: every rank calls it with the same arguments.  Queued, not awaited: ass_sync,
 * a download or ass_get_params after ass_sync see the result.  pre_drift_last != 0 leaves the positions with the next
 * step's first half-drift (src/integrator_leapfrog.c:45-47) already applied: only more stepping may follow.
 * A peer that does not arrive within 30 s makes the next ass_sync / download fail with ASS_ECOMM instead of hanging. */
int ass_shard_step(ass_sim* sim, int nsteps, int pre_drift_last);
/* bring every slab's x, v home (before a full download).  Collective, queued; the next ass_shard_step waits on the
 * device until every peer has taken its copy. */
int ass_shard_gather(ass_sim* sim);
/* ass_upload_particles for rows [i_low, i_high) only: `base` is row 0 of the caller's array */
int ass_upload_range(ass_sim* sim, const void* base, size_t stride_bytes, int i_low, int i_high, unsigned fields);

/* ---- measurement ---- */
/* Kernel classes for the per-kernel CUDA-event timers */
enum { ASS_K_GRAVITY = 0, ASS_K_SPRINGS = 1, ASS_K_LEAPFROG = 2, ASS_K_REDUCE = 3, ASS_K_CONNECT = 4, ASS_K_XFER = 5, ASS_K_COUNT = 6 };
int ass_timing_enable(ass_sim* sim, int on);      /* record an event pair around every kernel launch (slower) */
int ass_timing_reset(ass_sim* sim);
/* total device ms and launch count per class since the last reset (synchronises) */
int ass_timing_get(ass_sim* sim, double ms[ASS_K_COUNT], long long launches[ASS_K_COUNT]);
/* stream-ordered stopwatch on the simulation's stream: returns device ms between the two marks (synchronises) */
int ass_mark_begin(ass_sim* sim);
int ass_mark_end(ass_sim* sim, double* ms);
/* lap timer on the simulation's stream: begin / end only record events (nothing waits), collect synchronises and
 * returns the sum of all laps since the last collect.  For loops that must stay queued (sharded steps). */
int ass_lap_begin(ass_sim* sim);
int ass_lap_end(ass_sim* sim);
int ass_lap_collect(ass_sim* sim, double* total_ms, int* laps);
int ass_l2_flush(ass_sim* sim);                    /* overwrite a buffer larger than L2 */
/* ass_step, measured: the same fused loop, with the L2 overwritten before every step and a lap around every step's own
 * kernels (the flush outside); *device_ms = sum of the laps.  The host queues the whole batch and waits once. */
int ass_step_timed(ass_sim* sim, int nsteps, int with_heartbeat, double* device_ms);
/* DFMA-only microbenchmark: the FP64 roofline denominator (MEASURED_PEAKS.json has none). iters ~ 1<<14. */
int ass_probe_fp64_peak(int iters, double* tflops, double* ms);
/* pair-once gravity between two particle ranges (identical: the triangle; disjoint: the rectangle), device-timed with the L2
 * flushed before each of `reps` repetitions: the per-rank pieces of a sharded body, measurable on one GPU (tools/) */
int ass_probe_gravity_pair(ass_sim* sim, int rlo, int rhi, int slo, int shi, int reps, double* ms_avg);
/* several range pairs per repetition (ranges: n_pieces x {rlo, rhi, slo, shi}, at most 4), one after the other or with the
 * pair kernels on streams of their own and the reductions afterwards: what a rank of a sharded body queues per step */
int ass_probe_gravity_pieces(ass_sim* sim, int n_pieces, const int* ranges, int concurrent, int reps, double* ms_avg);
/* device copy bandwidth (read+write bytes / time) for a buffer of `bytes` */
int ass_probe_copy_bw(long long bytes, double* gbs);
/* Self-check of the branch-free, correctly rounded square root inside the spring kernels (csrc/spring_math.cuh) against
 * CUDA's __dsqrt_rn on `samples` pseudo-random arguments with binary exponents in [exp_lo, exp_hi): the number of
 * arguments whose results differ in any bit, and the first such argument. */
int ass_selfcheck_sqrt(long long samples, int exp_lo, int exp_hi, long long* mismatches, double* first_bad);
/* Host-only self-check of what ass_set_springs builds for the spring kernels (no device is touched, so it runs in the CPU
 * test suite): the spatial ranking, the CSR and the lane-transposed copy of the half-edge list, and the tile cut for `grid`
 * persistent CTAs with shared-memory tables of at most `slots` entries.  pos_xyzm: n x {x, y, z, m}.  stats[0] slices,
 * [1] padding records, [2] tiles, [3] largest table, [4] gather passes per step summed over all quarter warps and
 * [5] their minimum (one each), [6] violated invariants (must be 0), [7] records in the lane-transposed copy. */
int ass_selfcheck_spring_layout(const double* pos_xyzm, int n, const ass_spring* springs, int ns, int grid, int slots, long long* stats);
/* Host-only self-check of how a sharded body's gravity is divided by pairs of slabs: over all ranks every unordered pair of body
 * particles is evaluated by exactly one piece, and the partial accelerations a rank adds are exactly those its peers compute.
 * stats[0] pieces with work, [1] segments, [2] violations (must be 0), [3] / [4] most / fewest pairs on a rank. */
int ass_selfcheck_shard_division(int nranks, const int* bounds, int n_body, long long* stats);
/* Host-only self-check of the pair-once gravity plan for n_res resident and n_str streamed particles on a device with sm_count SMs
 * (rect == 0: one range against itself): every (resident block, real streamed particle) pair covered by exactly one CTA.
 * stats[0] CTAs, [1] residents per thread, [2] cut, [3] tiles per CTA, [4] / [5] resident / streamed blocks, [6] violations
 * (must be 0), [7] pieces skipped because they lie in the leading padding. */
int ass_selfcheck_gravity_plan(int n_res, int n_str, int rect, int sm_count, long long* stats);
/* lanes that share one node's half-edge row in the spring kernels (a slice of the lane-transposed list = 32 / lanes nodes) */
int ass_spring_lanes(void);
/* total number of kernel launches issued by this library in this process */
long long ass_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ASS_B200_H */
