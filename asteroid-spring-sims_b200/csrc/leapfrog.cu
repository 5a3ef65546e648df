// leapfrog.cu -- REBOUND's leapfrog drift/kick (reference src/integrator_leapfrog.c:39-67), zero_accel
// (src_spring/physics.cpp:557-564), and the step sequencing of reb_step / reb_integrate (src/rebound.c:69-144,
// 509-704).
//
// The reference's C files are built -std=c99, i.e. without FMA contraction: `x += 0.5*dt*vx` is two roundings of
// (0.5*dt)*vx followed by one of the add.  The kernels use __dmul_rn/__dadd_rn so positions and velocities are
// bit-identical to the reference given bit-identical accelerations.  HBM-bound: part1 moves 96 B/node, part2 160 B/node
// in this layout (algorithmic 72 / 120 B); in the fused loop both disappear into the spring kernel.
#include <math.h>

#include "common.cuh"

namespace {

__global__ void k_part1(vec4d* __restrict__ posm, const vec4d* __restrict__ vel, int lo, int hi, double dt) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const double h = __dmul_rn(0.5, dt);
	vec4d p = posm[i];
	const vec4d v = vel[i];
	p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
	p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
	p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
	posm[i] = p;
}

// part1 that also writes the rank-ordered mirror the spring kernel gathers from (common.cuh)
__global__ void k_part1_mirrored(vec4d* __restrict__ posm, const vec4d* __restrict__ vel, const int* __restrict__ rank, const SpringMirror mir,
		double* __restrict__ mir_m, int n, double dt) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double h = __dmul_rn(0.5, dt);
	vec4d p = posm[i];
	const vec4d v = vel[i];
	p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
	p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
	p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
	posm[i] = p;
	const int k = rank[i];
	mir.xy[k] = make_double2(p.x, p.y);
	mir.zv[k] = make_double2(p.z, v.z);
	mir.vxy[k] = make_double2(v.x, v.y);
	mir_m[k] = p.w;
}

__global__ void k_part2(vec4d* __restrict__ posm, vec4d* __restrict__ vel, const vec4d* __restrict__ acc, int lo, int hi, double dt, int pre_drift_next) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	const double h = __dmul_rn(0.5, dt);
	vec4d p = posm[i];
	vec4d v = vel[i];
	const vec4d a = acc[i];
	v.x = __dadd_rn(v.x, __dmul_rn(dt, a.x));
	v.y = __dadd_rn(v.y, __dmul_rn(dt, a.y));
	v.z = __dadd_rn(v.z, __dmul_rn(dt, a.z));
	p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
	p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
	p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
	if (pre_drift_next) {
		p.x = __dadd_rn(p.x, __dmul_rn(h, v.x));
		p.y = __dadd_rn(p.y, __dmul_rn(h, v.y));
		p.z = __dadd_rn(p.z, __dmul_rn(h, v.z));
	}
	vel[i] = v;
	posm[i] = p;
}

__global__ void k_zero_accel(vec4d* __restrict__ acc, int lo, int hi) {
	const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= hi) return;
	vec4d z; z.x = 0.0; z.y = 0.0; z.z = 0.0; z.w = 0.0;
	acc[i] = z;
}

}  // namespace

int ass_launch_part1_range(ass_sim* s, double dt, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	ass_tic(s, ASS_K_LEAPFROG);
	k_part1<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, lo, hi, dt);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_LEAPFROG);
	return ASS_OK;
}
int ass_launch_part1_mirrored(ass_sim* s, double dt) {
	if (s->n <= 0) return ASS_OK;
	ass_tic(s, ASS_K_LEAPFROG);
	k_part1_mirrored<<<(s->n + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, s->rank, s->mir, s->mir_m, s->n, dt);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_LEAPFROG);
	return ASS_OK;
}
int ass_launch_part2_range(ass_sim* s, double dt, bool pre_drift_next, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	ass_tic(s, ASS_K_LEAPFROG);
	k_part2<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->posm, s->vel, s->acc, lo, hi, dt, pre_drift_next ? 1 : 0);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_LEAPFROG);
	return ASS_OK;
}
int ass_launch_zero_accel_range(ass_sim* s, int lo, int hi) {
	if (hi <= lo) return ASS_OK;
	ass_tic(s, ASS_K_LEAPFROG);
	k_zero_accel<<<(hi - lo + 255) / 256, 256, 0, s->stream>>>(s->acc, lo, hi);
	LAUNCH_CHECK();
	ass_toc(s, ASS_K_LEAPFROG);
	return ASS_OK;
}
int ass_launch_part1(ass_sim* s, double dt) { return ass_launch_part1_range(s, dt, 0, s->n); }
static int launch_part2(ass_sim* s, double dt, bool pre_drift_next) { return ass_launch_part2_range(s, dt, pre_drift_next, 0, s->n); }
int ass_launch_part2(ass_sim* s, double dt) { return launch_part2(s, dt, false); }
int ass_launch_zero_accel(ass_sim* s) { return ass_launch_zero_accel_range(s, 0, s->n); }

#define ENTER(s)                                                              \
	REQUIRE(s, ASS_EINVAL, "null sim");                                       \
	REQUIRE((s)->have_state, ASS_ESTATE, "no particle state on the device");  \
	{                                                                         \
		int _rc = ass_ensure_device();                                        \
		if (_rc) return _rc;                                                  \
	}

extern "C" int ass_zero_accel(ass_sim* s) {
	ENTER(s);
	return ass_launch_zero_accel(s);
}

extern "C" int ass_leapfrog_part1(ass_sim* s) {
	ENTER(s);
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	int rc = ass_launch_part1(s, s->prm.dt);
	s->prm.t += s->prm.dt / 2.;  // integrator_leapfrog.c:50
	return rc;
}

extern "C" int ass_leapfrog_part2(ass_sim* s) {
	ENTER(s);
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	int rc = ass_launch_part2(s, s->prm.dt);
	s->prm.t += s->prm.dt / 2.;  // integrator_leapfrog.c:65-66
	s->prm.dt_last_done = s->prm.dt;
	return rc;
}

// [reb_integrator_leapfrog_part1] ; [zero_accel] ; spring_forces ; reb_integrator_leapfrog_part2 as the fused kernels of the step
// loop run them: for a caller that sees the four calls one by one but may defer them until the last (the drop-in in
// resident mode: host/dropin.cpp).  Same bits as the four calls.  Without springs it is [part1] ; [zero_accel] ; part2.
extern "C" int ass_substep_fused(ass_sim* s, int do_part1, int zero_first) {
	ENTER(s);
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	REQUIRE(!s->shard.active, ASS_ESTATE, "not for a sharded simulation");
	const double dt = s->prm.dt;
	const bool have_springs = s->ns > 0 && s->row_ptr != nullptr;
	int rc = ASS_OK;
	if (do_part1) {
		if ((rc = have_springs ? ass_launch_part1_mirrored(s, dt) : ass_launch_part1(s, dt))) return rc;
		s->prm.t += dt / 2.;
	}
	if (have_springs) {
		if ((rc = ass_launch_spring_kick_drift(s, zero_first != 0, dt, false, do_part1 != 0))) return rc;
	} else {
		if (zero_first && (rc = ass_launch_zero_accel(s))) return rc;
		if ((rc = ass_launch_part2(s, dt))) return rc;
	}
	s->prm.t += dt / 2.;
	s->prm.dt_last_done = dt;
	return ASS_OK;
}

// ---------------------------------------------------------------------------------------------------
// reb_step x nsteps without leaving the device
// ---------------------------------------------------------------------------------------------------
// One step of the reference (src/rebound.c:69-144) with the built-in callbacks is
//     part1 ; [gravity: a = g] ; [zero_accel] ; [spring_forces: a += f] ; part2 ; [heartbeat: center_sim]
// Fused schedule used here (positions "pre-drifted" = already carry the next part1):
//     gravity(x)                               -> acc            (only if gravity is on and its result survives)
//     springs + kick + drift [+ next part1]    -> posm', vel', acc   one kernel, ping-pong buffers
//     [heartbeat: CoM reduce ; shift + next part1]
// The built-in heartbeat (center_sim every step) sums as a parallel tree unless the caller asked for an order: the
// reference-order sum is one CTA walking all N particles (0.7 ms at N = 1e5, 25 x the spring kernel) and is meant for
// set-up helpers, where an ulp of position ends up in rs0.
struct HeartbeatSumOrder {
	ass_sim* s;
	bool saved;
	explicit HeartbeatSumOrder(ass_sim* s_) : s(s_), saved(s_->sum_sequential) {
		if (!s->sum_order_forced) s->sum_sequential = false;
	}
	~HeartbeatSumOrder() { s->sum_sequential = saved; }
};

static int step_loop(ass_sim* s, int nsteps, bool with_hb) {
	const ass_params& P = s->prm;
	HeartbeatSumOrder order_scope(s);
	const int af = P.additional_forces;
	const bool hb = with_hb && P.heartbeat == 1;
	const bool have_springs = s->ns > 0 && af != 0;
	// does gravity's result reach the kick?  (af == 2 zeroes it again, as the reference would)
	const bool grav = P.gravity == 1 && af != 2;
	int rc;
	// does the rank-ordered mirror hold exactly posm / vel?  Only the two kernels of this loop that write
	// both keep it so; anything else (heartbeat shift, a caller's own kernels between calls) leaves it stale and the
	// spring launcher re-gathers it.
	bool mirror_fresh = false;
	for (int step = 0; step < nsteps; step++) {
		const double dt = s->prm.dt;
		const bool last = (step == nsteps - 1);
		if (s->lap_flush) {  // measurement only (ass_step_timed): cold L2 for every step, the flush outside the lap
			if ((rc = ass_l2_flush(s)) || (rc = ass_lap_begin(s))) return rc;
		}
		if (!s->predrifted) {
			if ((rc = have_springs ? ass_launch_part1_mirrored(s, dt) : ass_launch_part1(s, dt))) return rc;
			mirror_fresh = have_springs;
		}
		// Nothing but springs between here and the end of the batch (REB_GRAVITY_NONE or zero_accel after gravity, no
		// heartbeat): the remaining steps run as ONE launch, the CTAs meeting at a grid barrier after every step
		// (springs.cu: k_spring_steps).  Same kernels' code, same bits as one launch per step (tests).
		if (have_springs && !grav && !hb && !s->lap_flush && nsteps - step >= 2 && mirror_fresh && ass_spring_steps_applicable(s)) {
			const int batch = nsteps - step;
			if ((rc = ass_launch_spring_steps(s, af == 2, dt, batch)) < 0) return rc;
			if (rc == 0) {
				for (int i = 0; i < batch; i++) {  // the clock advances as the reference's does: twice dt/2 per step
					s->prm.t += dt / 2.;
					s->prm.t += dt / 2.;
				}
				s->prm.dt_last_done = dt;
				s->predrifted = false;
				return ASS_OK;
			}
			// (refused by the device: one launch per step below)
		}
		s->predrifted = false;
		s->prm.t += dt / 2.;
		if (grav) {
			if ((rc = ass_launch_gravity(s))) return rc;
		}
		// pre-drifting folds the NEXT step's part1 into this step's last kernel; not when a heartbeat must see the
		// end-of-step positions first, and not on the last step of the batch (the caller sees end-of-step state).
		const bool pre = !last && !hb;
		if (have_springs) {
			// af == 1 with gravity NONE: accelerations accumulate across steps, exactly as in the reference
			const bool zero_first = (af == 2);
			if ((rc = ass_launch_spring_kick_drift(s, zero_first, dt, pre, mirror_fresh))) return rc;
			mirror_fresh = pre;
		} else {
			if (af == 2) {
				if ((rc = ass_launch_zero_accel(s))) return rc;
			}
			if ((rc = launch_part2(s, dt, pre))) return rc;
		}
		s->predrifted = pre;
		s->prm.t += dt / 2.;
		s->prm.dt_last_done = dt;
		if (hb) {
			const int hi = s->n - P.n_perts;
			if ((rc = ass_launch_com(s, 0, hi, false, s->d_red))) return rc;
			// x -= CoM for all N particles (physics.cpp:103-108) and, unless this is the last step, the next part1
			if ((rc = ass_launch_shift(s, 0, s->n, s->d_red, false, -1.0, last ? 0.0 : dt))) return rc;
			s->predrifted = !last;
			mirror_fresh = false;
		}
		if (s->lap_flush && (rc = ass_lap_end(s))) return rc;
	}
	return ASS_OK;
}

extern "C" int ass_step(ass_sim* s, int nsteps, int with_heartbeat) {
	ENTER(s);
	REQUIRE(nsteps >= 0, ASS_EINVAL, "negative step count");
	REQUIRE(!s->shard.active, ASS_ESTATE, "sharded simulations step through ass_leapfrog_part1 / ass_shard_publish / ... (host-driven)");
	return step_loop(s, nsteps, with_heartbeat != 0);
}

// ass_step as the product runs it -- the fused loop, positions pre-drifted from one step into the next, nothing between
// the steps on the host -- but measured step by step: before every step the L2 is overwritten, and a pair of events
// brackets the step's own kernels (the flush stays outside).  Returns the sum of the laps.  For bench.py: a body whose
// working set fits in L2 has to be flushed between timed iterations, and doing that from the host would put a
// synchronisation and the launch latency of a cold queue into every step.
extern "C" int ass_step_timed(ass_sim* s, int nsteps, int with_heartbeat, double* device_ms) {
	ENTER(s);
	REQUIRE(nsteps >= 0 && device_ms, ASS_EINVAL, "bad argument");
	REQUIRE(!s->shard.active, ASS_ESTATE, "sharded simulations are timed by the caller around ass_shard_step");
	double dummy;
	int laps = 0;
	int rc = ass_lap_collect(s, &dummy, &laps);  // drop laps somebody left open
	if (rc) return rc;
	s->lap_flush = true;
	rc = step_loop(s, nsteps, with_heartbeat != 0);
	s->lap_flush = false;
	if (rc) return rc;
	return ass_lap_collect(s, device_ms, &laps);
}

// reb_check_exit, src/rebound.c:509-569, for the leapfrog case (reb_integrator_synchronize is a no-op for leapfrog,
// src/integrator_leapfrog.c:69-71; no MPI, no paused visualisation)
static int check_exit(ass_params* r, int n, double tmax, double* last_full_dt) {
	const double dtsign = copysign(1., r->dt);
	if (r->status >= 0) {
		// exit now
	} else if (tmax != INFINITY) {
		if (r->exact_finish_time == 1) {
			if ((r->t + r->dt) * dtsign >= tmax * dtsign) {
				double tscale = 1e-12 * fabs(tmax);
				if (tscale < 1e-200) tscale = 1e-12;
				if (r->t == tmax) {
					r->status = 0;  // REB_EXIT_SUCCESS
				} else if (r->status == -2) {  // REB_RUNNING_LAST_STEP
					if (fabs(r->t - tmax) < tscale) r->status = 0;
					else r->dt = tmax - r->t;
				} else {
					r->status = -2;
					if (r->dt_last_done != 0.) *last_full_dt = r->dt_last_done;
					r->dt = tmax - r->t;
				}
			} else if (r->status == -2) {
				r->status = -1;
			}
		} else if (r->t * dtsign >= tmax * dtsign) {
			r->status = 0;
		}
	}
	if (n <= 0) r->status = 2;  // REB_EXIT_NOPARTICLES
	return r->status;
}

extern "C" int ass_integrate(ass_sim* s, double tmax) {
	ENTER(s);
	REQUIRE(!s->shard.active, ASS_ESTATE, "sharded simulations are stepped by the host launcher");
	REQUIRE(!s->predrifted, ASS_ESTATE, "internal: positions are mid-step");
	int rc;
	double last_full_dt = s->prm.dt;     // rebound.c:629
	s->prm.dt_last_done = 0.;            // :630
	s->prm.status = -1;                  // :632 REB_RUNNING
	if (s->prm.heartbeat == 1) {         // :633 heartbeat at t = 0
		HeartbeatSumOrder order_scope(s);
		const int hi = s->n - s->prm.n_perts;
		if ((rc = ass_launch_com(s, 0, hi, false, s->d_red))) return rc;
		if ((rc = ass_launch_shift(s, 0, s->n, s->d_red, false, -1.0, 0.0))) return rc;
	}
	if (tmax == INFINITY) {
		ass_set_error("ass_integrate: tmax = INFINITY never returns; drive ass_step from the host instead");
		return ASS_EINVAL;
	}
	// Steps with an unchanged dt are batched so the fused loop can run without host round trips; the step that
	// check_exit shortens (exact_finish_time) runs alone.
	while (check_exit(&s->prm, s->n, tmax, &last_full_dt) < 0) {
		int batch = 1;
		if (s->prm.status == -1 && s->prm.dt != 0.) {
			// how many more full steps fit strictly before the overshoot test fires?  Re-derive with the same
			// floating-point test as check_exit to stay bit-compatible: simulate t forward.
			ass_params probe = s->prm;
			double lfd = last_full_dt;
			batch = 0;
			while (batch < 4096) {
				probe.t += probe.dt / 2.;
				probe.t += probe.dt / 2.;
				probe.dt_last_done = probe.dt;
				batch++;
				if (check_exit(&probe, s->n, tmax, &lfd) != -1) break;
			}
		}
		if ((rc = step_loop(s, batch, true))) return rc;
	}
	if (s->prm.exact_finish_time == 1) s->prm.dt = last_full_dt;  // :653-655
	CU_TRY(cudaStreamSynchronize(s->stream));
	return s->prm.status;
}
