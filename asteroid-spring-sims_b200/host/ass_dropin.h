// ass_dropin.h -- optional control surface of the drop-in layer (asteroid-spring-sims_b200/host/dropin.cpp).
//
// The drop-in re-implements, over the C-ABI of include/ass_b200.h, the reference symbols that sit on the per-step
// path -- same names, same signatures, same error behaviour -- so a reference problem module compiles and links
// unchanged:
//
//   C   (librebound)   reb_integrate, reb_step, reb_calculate_acceleration, reb_integrator_leapfrog_part1/2,
//                      reb_output_check
//   C++ (src_spring)   spring_forces, zero_accel, connect_springs_dist, mindist, rand_ellipsoid,
//                      center_sim, subtract_com, subtract_cov, move_resolved, compute_com, compute_cov,
//                      set_gamma, divide_gamma, adjust_spring_props, kill_springs, add_spring_helper, del_spring,
//                      quadrupole_accel, drift_resolved, drift_bin, sum_mass
//
// Residency.  The host `r->particles[]` and the module-owned `std::vector<spring> springs` remain the API-visible
// truth (SURVEY 8b "Ownership"); the device copy is a cache.
//   coherent mode (default)  every entry point is self-contained: upload what it reads, run, download what it wrote.
//                            Always correct, whatever the module's callbacks do with r->particles.
//   resident mode            (ASS_RESIDENT=1 in the environment, or ass_dropin_set_resident(1)) inside reb_integrate
//                            the device copy is the truth: entry points launch kernels and return.  The host arrays
//                            are refreshed when reb_output_check() fires (modules gate all their output on it), at
//                            the end of reb_integrate, and on ass_dropin_sync_host().  Contract: callbacks touch
//                            particle data only through the functions listed above, or bracket direct access with
//                            ass_dropin_sync_host() / ass_dropin_host_modified().
#ifndef ASS_DROPIN_H
#define ASS_DROPIN_H

#ifdef __cplusplus
extern "C" {
#endif

struct reb_simulation;

void ass_dropin_set_resident(int on);
int ass_dropin_is_resident(void);
// make r->particles[] current (no-op when it already is). Returns 0 or a negative ASS_E* code.
int ass_dropin_sync_host(struct reb_simulation* r);
// r->particles[] was edited by the caller: re-upload before the next device operation.
void ass_dropin_host_modified(struct reb_simulation* r);
// `springs` was edited behind the library's back (direct writes to springs[i].k etc.).
void ass_dropin_springs_modified(void);
// drop the device state attached to r (also done by nothing else: call before reb_free_simulation).
void ass_dropin_release(struct reb_simulation* r);
// number of kernels launched so far / total device-side milliseconds are available through include/ass_b200.h

#ifdef __cplusplus
}
#endif
#endif
