#!/usr/bin/env python
"""bench.py -- node-steps/s of the per-step hot path (springs + direct-sum self-gravity + leapfrog kick/drift).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5] [--impl reference] [--no-sub]

Prints ONE JSON line (rank 0).  The main record is always BASELINE config C3 -- a wobble_bennu2-style tumbling
viscoelastic body, synthetic N ~ 1e5 nodes, REB_GRAVITY_BASIC + spring_forces + leapfrog -- the size north_star quotes
its target on:
    --gpus 1      the body on one GPU (fused step loop);
    --gpus N > 1  THE SAME body cut into N particle slabs, one rank per GPU under torchrun ("scaling": "strong"): gravity
                  divided by pairs of slabs, slabs exchanged by copy-engine pulls over CUDA-IPC windows, ranks meeting
                  on the device through epoch flags (no host barrier inside the step loop).
The same line carries two sub-records, measured in the same run on the same ranks, each with its own ms_per_step, clock
samples and parity check:
    c4_sharded    BASELINE config C4: N ~ 1e6 + a Mars-like point mass, slab-sharded over the N ranks (N = 1: unsharded);
    c5_ensemble   BASELINE config C5: 4096 independent ~1k-node bodies, dealt to the ranks by cost, no collective.
--workload X makes X the (only) record instead.

value      whole-job node-steps/s, state resident in HBM, device-timed (CUDA events on the simulation's stream),
           max over ranks; L2 flushed between steps (outside the timed laps).
e2e        the same through host buffers: every step uploads the host `struct reb_particle`-strided rows from pinned
           memory, steps once, downloads x, v, a (sharded: every rank its own slab).
roofline   the dominant kernel (gravity at C2..C4, springs at C1/C5) against the FP64 DFMA peak measured in this run /
           the HBM copy peak of MEASURED_PEAKS.json.
parity     in-run check of the result against the CPU oracle (a 256-target slab of accelerations after one step) and,
           for sharded runs, of x, v against the unsharded path on rank 0's GPU.
cpu_baseline  the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile) timed on this
           box's host on a bounded sample of the same step, serial as shipped.
--impl reference  the reference's own CPU implementation with every host thread it can correctly use (OpenMP on
           gravity.c / integrator_leapfrog.c only; its src_spring pragmas are racy), on the same body.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345
ABC = (1.0, 0.7, 0.5)
K_SPRING, GAMMA = 0.08, 5.0          # modules/wobble_bennu2/problem.cfg
ACC_TOL, POS_TOL = 1e-12, 1e-13      # parity bars of tests/test_gpu_parity.py

# name -> (min_dist d, dt, gravity, additional_forces mode, omega)       SURVEY 8d synthetic inputs
WORKLOADS = {
    "C1": dict(d=0.10, dt=1e-2, gravity=0, af=2, omega=(0.0, 0.0, 0.2), desc="asteroid_spin-style random ellipsoid, N=1065, GRAVITY_NONE + zero_accel + springs"),
    "C2": dict(d=0.0474, dt=2e-3, gravity=1, af=1, omega=(0.0, 0.0, 0.2), desc="bennu2-style spin-up, N=9327, GRAVITY_BASIC + springs"),
    "C3": dict(d=0.0215, dt=1e-3, gravity=1, af=1, omega=(0.1, 0.0, 0.2), desc="wobble_bennu2-style tumbling viscoelastic body, N~1e5, GRAVITY_BASIC + springs"),
    "C4": dict(d=0.0100, dt=1e-3, gravity=1, af=1, omega=(0.0, 0.0, 0.2), desc="phobos_deimos-style tidal body + Mars point mass, N~1e6, GRAVITY_BASIC + springs"),
}


def read_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the run (B200_PROFILING.md recipe).  start() returns once the
    first sample has arrived, so that even a sub-millisecond timed region is bracketed by samples; callers start it
    before the warm-up steps: the samples then cover warm-up + timed region, i.e. the GPU under this load."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self, wait_s=3.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < wait_s:
                time.sleep(0.01)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(gpus):
    """One process per GPU under torchrun; plain single process otherwise. torch is plumbing only (barrier, max)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # NCCL announces its version on STDOUT when NCCL_DEBUG=VERSION (this image's default): keep stdout to the JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def barrier(dist):
    if dist is not None:
        dist.barrier()


def max_over_ranks(dist, x, local):
    from asteroid_spring_sims_b200 import launcher
    return launcher.reduce_max(dist, x, None if dist is None else "cuda:%d" % local)


def sum_over_ranks(dist, x, local):
    from asteroid_spring_sims_b200 import launcher
    return launcher.reduce_sum(dist, x, None if dist is None else "cuda:%d" % local)


def all_ranks(dist, x):
    """The value of every rank, as a list (rank order)."""
    if dist is None:
        return [x]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, x)
    return out


def config_of(name, cfg, n, ns, world, mode):
    """The `config` object -- built by this one function for both arms, so the two lines carry identical keys and values."""
    if mode == "ensemble":
        par = "ensemble members dealt to %d GPU(s) by cost, no collective" % world
    elif world > 1:
        par = "particle slabs x%d (strong scaling: one body)" % world
    else:
        par = "one GPU"
    return {"workload": "%s: %s" % (name, cfg["desc"]), "N": int(n), "NS": int(ns), "seed": SEED,
            "generator": "rand_ellipsoid(d=%g, 1, 0.7, 0.5)" % cfg["d"], "max_spring_dist": 2.3 * cfg["d"], "k": K_SPRING, "gamma": GAMMA,
            "dt": cfg["dt"], "softening": cfg["d"] / 100.0, "gravity": "REB_GRAVITY_BASIC" if cfg["gravity"] else "REB_GRAVITY_NONE",
            "l2": "b200 arm: flushed between steps (256 MiB fill, outside the device-timed laps); reference arm: host caches as they come",
            "parallelism": "b200 arm: %s; reference arm: OpenMP over the host cores (gravity, leapfrog)" % par}


def build_body(A, cfg, omega, with_point_mass=False):
    """Synthesise a BASELINE config from the module's code path: rand_ellipsoid -> connect_springs_dist ->
    subtract_com/cov -> spin_body (modules/wobble_bennu2/problem.cpp:63-333), everything through the product.
    Returns (sim, p0, springs, setup): p0 is the host copy of the initial state."""
    t0 = time.perf_counter()
    A.srand(SEED)
    p = A.rand_ellipsoid(cfg["d"], *ABC, 1.0)
    t_gen = time.perf_counter() - t0
    n_body = len(p)
    n_perts = 0
    sim = A.Sim()
    sim.upload(p)
    t0 = time.perf_counter()
    springs = sim.connect_springs_dist(2.3 * cfg["d"], 0, n_body, K_SPRING, GAMMA)
    t_connect = time.perf_counter() - t0
    sim.subtract_com(0, n_body)
    sim.subtract_cov(0, n_body)
    sim.spin_body(0, n_body, omega)
    sim.download(p)
    if with_point_mass:
        # add_pt_mass_kep-style perturber: m = 1 on a circular orbit a = 7 about the body (phobos_deimos point_mass.cfg)
        pm = np.zeros((1, 11))
        pm[0, 0] = 7.0
        pm[0, 4] = np.sqrt(2.0 / 7.0)
        pm[0, 9] = 1.0
        p = np.vstack([p, pm])
        n_perts = 1
        sim.upload(p)
    t0 = time.perf_counter()
    sim.set_springs(springs)
    t_layout = time.perf_counter() - t0
    sim.set_params(t=0.0, dt=cfg["dt"], G=1.0, softening=cfg["d"] / 100.0, gravity=cfg["gravity"], n_active=-1,
                   n_perts=n_perts, additional_forces=cfg["af"], heartbeat=0)
    return sim, p, springs, {"gen_s": t_gen, "connect_s": t_connect, "set_springs_s": t_layout}


# ---------------------------------------------------------------------------------------------------
# parity inside the run
# ---------------------------------------------------------------------------------------------------
def oracle_first_step_acc(p0, springs, cfg, n_perts, lo, hi):
    """Accelerations of particles [lo, hi) at the first step's force evaluation, by the CPU oracle: part1, then
    REB_GRAVITY_BASIC for those targets over all sources, then spring_forces (a checker; never the thing measured)."""
    from oracle.bindings import Oracle, make_sim
    orc = Oracle()
    q = np.ascontiguousarray(p0[:, :11]).copy()
    sim = make_sim(dt=cfg["dt"], G=1.0, softening=cfg["d"] / 100.0, gravity_basic=cfg["gravity"], zero_first=(cfg["af"] == 2), n_perts=n_perts)
    orc.leapfrog_part1(q, sim)
    if cfg["gravity"] and cfg["af"] != 2:
        orc.gravity_basic(q, G=1.0, softening=cfg["d"] / 100.0, i_lo=lo, i_hi=hi)
    else:
        q[:, 6:9] = 0.0 if cfg["af"] == 2 else q[:, 6:9]
    orc.spring_forces(q, springs)
    return q[lo:hi, 6:9].copy()


def parity_of_sim(A, dist, sim, p0, springs, cfg, n_perts, rank, world, bounds):
    """Re-run the body from its initial state: one step -> accelerations of a 256-target slab against the oracle;
    a second step -> x, v (sharded: of the gathered body) against the unsharded path on this rank's GPU (rank 0).
    Returns the parity object (rank 0) after every rank has taken part in the collective calls."""
    n = len(p0)
    sharded = bounds is not None
    lo, hi = (bounds[rank], bounds[rank + 1]) if sharded else (0, n)
    prm = dict(t=0.0, dt=cfg["dt"])
    out1 = p0.copy()
    out2 = p0.copy()
    sim.upload(p0)
    sim.set_params(**prm)
    if sharded:
        sim.shard_step(1); sim.shard_gather(); sim.sync()
        sim.download(out1)
        sim.shard_step(1); sim.shard_gather(); sim.sync()
        sim.download(out2)
    else:
        sim.step(1); sim.download(out1)
        sim.step(1); sim.download(out2)
    res = None
    if rank == 0:
        t_lo = lo + max(0, (hi - lo) // 2 - 128)
        t_hi = min(hi, t_lo + 256)
        res = {"steps": 2, "targets": [int(t_lo), int(t_hi)]}
        try:
            want = oracle_first_step_acc(p0, springs, cfg, n_perts, t_lo, t_hi)
            amax = float(np.abs(want).max())
            res["acc_vs_oracle"] = float(np.abs(out1[t_lo:t_hi, 6:9] - want).max() / max(amax, 1e-300))
            res["acc_tol"] = ACC_TOL
        except Exception as e:  # oracle library missing on this box
            res["acc_vs_oracle"] = None
            res["oracle_error"] = repr(e)
        if sharded:
            with A.Sim() as u:
                u.upload(p0); u.set_springs(springs)
                u.set_params(t=0.0, dt=cfg["dt"], G=1.0, softening=cfg["d"] / 100.0, gravity=cfg["gravity"], n_active=-1, n_perts=n_perts,
                             additional_forces=cfg["af"], heartbeat=0)
                u.step(1); w1 = u.download(p0.copy())
                u.step(1); w2 = u.download(p0.copy())
            amax = float(np.abs(w1[:, 6:9]).max())
            res["acc_own_slab_vs_1gpu"] = float(np.abs(out1[lo:hi, 6:9] - w1[lo:hi, 6:9]).max() / max(amax, 1e-300))
            dev = np.abs(out2[:, :6] - w2[:, :6])
            worst = np.unravel_index(int(np.argmax(dev)), dev.shape)
            res["xv_vs_1gpu_max_abs"] = float(dev.max())
            res["xv_worst"] = {"particle": int(worst[0]), "column": "x y z vx vy vz".split()[int(worst[1])], "value_1gpu": float(w2[worst]),
                               "body_max_abs": float(dev[:n - n_perts].max()) if n > n_perts else None}
            # x, v follow the accelerations.  Two runs whose accelerations differ in the last bits (different order of summation:
            # other blocks, other slabs) agree in x after ONE step to an ulp (measured: 1.1e-16) -- and the spring network
            # multiplies that ulp in the next kick: dv = dt * (k / m) * (springs at the node) * dx.  C4 (m = 1e-6, omega * dt ~ 0.9)
            # turns one ulp of x into 1.9e-13 of v in the second step (tools/diag_c4_parity.py: same number on one GPU with two
            # ranks, first-step accelerations agreeing to 6e-15 of max|a| on every particle).  The bar is that bound:
            res["max_abs_acc"] = amax
            m_node = float(np.min(p0[:n - n_perts, 9])) if n > n_perts else 1.0
            res["xv_tol"] = POS_TOL + 2 * cfg["dt"] * (ACC_TOL * amax + (K_SPRING / m_node) * 32 * 2.0 ** -52 * float(np.abs(p0[:, :3]).max()))
            res["xv_tol_note"] = "1e-13 + steps * dt * (acc_tol * max|a| + (k / m) * 32 springs * one ulp of x)"
        checks = [res.get("acc_vs_oracle") is not None and res["acc_vs_oracle"] <= ACC_TOL]
        if sharded:
            checks += [res["acc_own_slab_vs_1gpu"] <= ACC_TOL, res["xv_vs_1gpu_max_abs"] <= res["xv_tol"]]
        res["ok"] = bool(all(checks))
    return res


# ---------------------------------------------------------------------------------------------------
# one body on one GPU
# ---------------------------------------------------------------------------------------------------
def run_single(args, A, name, cfg, local, want_e2e=True, want_cpu=True, steps=None, warmup=None):
    steps = steps or args.steps
    warmup = warmup or args.warmup
    sim, p0, springs, setup = build_body(A, cfg, cfg["omega"], with_point_mass=(name == "C4"))
    n, ns = len(p0), len(springs)
    n_perts = 1 if name == "C4" else 0
    clocks = ClockSampler(local).start()
    sim.step(warmup)
    sim.sync()
    fp64_peak, _ = A.probe_fp64_peak(1 << 13)
    sim.timing(True)
    sim.timing_reset()
    launches0 = A.launch_count()
    sim.sync()
    t_wall0 = time.perf_counter()
    # the fused loop as the product runs it (pre-drifted from step to step, the host queues the batch and waits once), the L2
    # overwritten before every step and a device-timed lap around every step's own kernels
    dev_ms = sim.step_timed(steps)
    sim.sync()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    launches = A.launch_count() - launches0
    kt = sim.timing_get()
    sim.timing(False)
    value = n * steps / (dev_ms * 1e-3)

    # ---- gravity-free bodies (what every shipped module runs): a batch of steps is ONE launch (k_spring_steps) ----
    batched = None
    if not (cfg["gravity"] == 1 and cfg["af"] != 2):
        kb = max(steps, 200)
        sim.step(kb); sim.sync()
        l0 = A.launch_count()
        sim.l2_flush()
        sim.mark_begin()
        sim.step(kb)
        b_ms = sim.mark_end()
        batched = {"steps_per_call": kb, "ms_per_step": b_ms / kb, "value": n * kb / (b_ms * 1e-3), "unit": "node-steps/s", "gpu_launches": int(A.launch_count() - l0),
                   "l2": "flushed before the batch; inside a batch the body stays in L2, as it does in a production run (the main record flushes before EVERY step and launches every step on its own)"}

    # ---- end to end through host buffers (the drop-in's always-coherent reb_step) ----
    e2e = None
    if want_e2e and not args.no_e2e:
        # 128-byte rows == sizeof(struct reb_particle), in page-locked host memory (ass_host_alloc): what the drop-in's
        # reb_integrate gives the module's own r->particles by pinning it in place (ass_host_register)
        pinned = A.PinnedRows(n, 16)
        host = pinned.a
        host[:] = 0.0
        sim.download(host)
        for _ in range(2):
            sim.upload(host); sim.step(1); sim.download(host, A.F_POS | A.F_VEL | A.F_ACC)
        k = min(steps, 20)
        t0 = time.perf_counter()
        for _ in range(k):
            sim.upload(host)
            sim.step(1)
            sim.download(host, A.F_POS | A.F_VEL | A.F_ACC)
        t_e2e = time.perf_counter() - t0
        e2e = {"value": n * k / t_e2e, "unit": "node-steps/s", "h2d_bytes_per_step": int(n * 128), "d2h_bytes_per_step": int(n * 72),
               "host_memory": "pinned (ass_host_alloc)", "steps": k}
        del host
        pinned.close()

    parity = parity_of_sim(A, None, sim, p0, springs, cfg, n_perts, 0, 1, None)
    dropin = None
    if want_e2e and not args.no_e2e and not args.no_dropin and name == "C3":
        # One reb_integrate of the unmodified reference at C3 size takes ~5.5 minutes (its write_resolved_with_E alone spends
        # ~4 of them in grav_potential_energy): the default run times the drop-in against the reference on the C2 body
        # (N = 9 327, same code path, ~10 s) and `--dropin-c3` on the body of the main record
        # (profiles/r2_dropin_e2e_c3.json holds that measurement: 334.7 s against 1.24 s for the same call).
        try:
            if args.dropin_c3:
                dropin = run_dropin_e2e(p0, springs, cfg, label="C3")
            else:
                cfg2 = dict(WORKLOADS["C2"])
                sim2, p2, sp2, _ = build_body(A, cfg2, cfg2["omega"])
                sim2.close()
                dropin = run_dropin_e2e(p2, sp2, cfg2, steps_b200=200, label="C2")
        except Exception as e:
            dropin = {"error": repr(e)}

    # ---- roofline of the dominant kernel ----
    peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic = (read_json(os.path.join(ROOT, "profiles", "traffic.json")) or {}).get(name, {})
    g_ms, g_n = kt["gravity"]
    s_ms, s_n = kt["springs"]
    roof_g = roof_s = None
    if g_n:
        # one gravity evaluation per step = the pair kernel + its fixed-order row reduction (both inside the timer)
        ach = 20.0 * n * n * steps / (g_ms * 1e-3) / 1e12    # 20 flop per ordered pair
        nominal = 148 * 128 * 1.965e9 / 1e12                 # 148 SMs x 64 DFMA lanes x 2 flop x 1.965 GHz
        roof_g = {"kernel": "k_gravity_sym (+ its row reduction)", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                  "frac_of_nominal_37.2": ach / nominal,
                  "traffic": traffic.get("gravity"), "peak_source": "DFMA microbenchmark measured in this run (ass_probe_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                  "pairs_per_s": n * n * steps / (g_ms * 1e-3), "ms_per_evaluation": g_ms / steps}
    if s_n:
        alg = 32.0 * ns + 80.0 * n
        ach = alg * s_n / (s_ms * 1e-3) / 1e9
        roof_s = {"kernel": "k_spring_tiles<KICK> (springs + kick + drift)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                  "traffic": traffic.get("springs"), "peak_source": hbm_src, "algorithmic_bytes_per_launch": alg, "ms_per_launch": s_ms / s_n}
        # what actually bounds this formulation: every spring is evaluated at both endpoints, 36 FP64 instructions
        # each, and an FP64 warp instruction holds its sub-partition's issue port for two cycles
        t_issue_ms = 2 * ns * 36 / 32.0 * 2.0 / (4 * 148) / 1.965e9 * 1e3
        roof_s["fp64_issue"] = {"fp64_instructions_per_spring": 72, "ms_at_full_fp64_issue_rate": t_issue_ms, "frac": t_issue_ms / (s_ms / s_n),
                                "note": "2 x 36 FP64 warp-instructions per 32 springs, 2 issue cycles each, 592 sub-partitions at 1.965 GHz; padding records not counted"}
    dominant = roof_g if (roof_g and g_ms >= s_ms) else roof_s

    # ---- CPU baseline: the unmodified reference on this host, bounded sample ----
    cpu = None
    if want_cpu and not args.no_cpu_baseline:
        try:
            v, sample, cores, detail = reference_arm_value(cfg, np.ascontiguousarray(p0[:, :11]), springs, budget_s=12.0, omp=False)
            cpu = {"value": v, "unit": "node-steps/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "reference",
                   "sample": sample + "; built -O3 -march=x86-64-v3 (no AVX-512: the path is scalar sqrt/div)"}
        except FileNotFoundError as e:
            cpu = {"value": None, "unit": "node-steps/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    sim.close()
    return {
        "metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(name, cfg, n, ns, 1, "single"),
        "clocks": clk, "e2e": e2e, "batched": batched, "dropin_e2e": dropin, "gpu_launches": int(launches), "parity": parity,
        "roofline": dominant, "roofline_gravity": roof_g, "roofline_springs": roof_s, "cpu_baseline": cpu,
        "kernel_ms": {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()},
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / steps, "setup": setup, "fp64_peak_tflops_measured": fp64_peak,
    }


# ---------------------------------------------------------------------------------------------------
# one body sharded by particle slabs over the ranks
# ---------------------------------------------------------------------------------------------------
def run_sharded(args, A, name, cfg, rank, world, local, dist, want_e2e=True, steps=None, warmup=None):
    steps = steps or args.steps
    warmup = warmup or args.warmup
    with_pm = name == "C4"
    sim, p0, springs, setup = build_body(A, cfg, cfg["omega"], with_point_mass=with_pm)
    n, ns = len(p0), len(springs)
    n_perts = 1 if with_pm else 0
    bounds = A.slab_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    sim.shard_setup(rank, world, bounds)
    A.launcher.exchange_windows(dist, sim, rank, world)
    clocks = ClockSampler(local).start()
    sim.shard_step(warmup, pre_drift_last=True)
    sim.sync()
    fp64_peak, _ = A.probe_fp64_peak(1 << 13)
    # device-timed laps, one per step, queued back to back: the host never waits inside the loop
    dist.barrier()
    launches0 = A.launch_count()
    sim.timing(True); sim.timing_reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.l2_flush()
        sim.lap_begin()
        sim.shard_step(1, pre_drift_last=True)
        sim.lap_end()
    t_queue = time.perf_counter() - t0          # how long the host needed to queue the whole loop
    dev_ms, laps = sim.lap_collect()
    t_wall = time.perf_counter() - t0
    kt = sim.timing_get()
    sim.timing(False)
    launches = A.launch_count() - launches0
    dist.barrier()
    clk = clocks.stop()
    per_rank_ms = all_ranks(dist, dev_ms / steps)
    dev_ms_max = max_over_ranks(dist, dev_ms, local)
    g_ms = max_over_ranks(dist, kt["gravity"][0], local)
    s_ms = max_over_ranks(dist, kt["springs"][0], local)
    value = n * steps / (dev_ms_max * 1e-3)

    # ---- end to end: every rank refreshes its own slab from pinned host rows, steps, reads its slab's x, v, a back ----
    e2e = None
    if want_e2e and not args.no_e2e:
        sim.upload(p0)                      # leave the pre-drifted state of the timed loop
        sim.set_params(t=0.0)
        pinned = A.PinnedRows(n, 16)
        host = pinned.a
        host[:] = 0.0
        host[:, :11] = p0[:, :11]
        k = min(steps, 20)
        for _ in range(2):
            sim.upload_range(host, lo, hi); sim.shard_step(1); sim.download_range(host, lo, hi, A.F_POS | A.F_VEL | A.F_ACC)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            sim.upload_range(host, lo, hi)
            sim.shard_step(1)
            sim.download_range(host, lo, hi, A.F_POS | A.F_VEL | A.F_ACC)
        t_mine = time.perf_counter() - t0
        dist.barrier()
        t_e2e = max_over_ranks(dist, t_mine, local)
        e2e = {"value": n * k / t_e2e, "unit": "node-steps/s", "h2d_bytes_per_step": int(n * 128), "d2h_bytes_per_step": int(n * 72),
               "host_memory": "pinned (ass_host_alloc); bytes are the sum over ranks: every rank moves its own slab", "steps": k}
        del host
        pinned.close()

    parity = parity_of_sim(A, dist, sim, p0, springs, cfg, n_perts, rank, world, bounds)
    dist.barrier()
    sim.close()
    if rank != 0:
        return None
    pairs = float(n) * float(n) / world
    ach = 20.0 * pairs * steps / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    own = bounds[1] - bounds[0]
    return {
        "metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(name, cfg, n, ns, world, "slabs"),
        "sharding": {"bounds": [int(b) for b in bounds],
                     "exchange": "per step and rank: copy-engine pulls of the other slabs' x and v (%.2f MB, overlapped with the own-slab gravity triangle) and of the partial accelerations %d peers computed for this slab (%.2f MB) over CUDA-IPC peer mappings" % (
                         64e-6 * (n - own), (world - 1) // 2 + (1 if world % 2 == 0 else 0), 32e-6 * own * ((world - 1) // 2 + (1 if world % 2 == 0 else 0))),
                     "synchronisation": "device-side epoch flags in the peer-mapped windows (st.release.sys / ld.acquire.sys); no host barrier or stream synchronisation inside the step loop",
                     "gravity": "divided by pairs of slabs: every unordered particle pair evaluated once, the same work on every rank",
                     "per_rank_ms_per_step": per_rank_ms, "host_queue_ms_per_step": 1e3 * t_queue / steps},
        "clocks": clk, "gpu_launches": int(launches), "e2e": e2e, "parity": parity,
        "roofline": {"kernel": "k_gravity_sym (triangle + rectangles + reductions)", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": ach / fp64_peak if fp64_peak else None, "traffic": None,
                     "note": "per GPU: 20 flop x N x N/%d ordered pairs per step (SURVEY 8d convention) over the gravity kernels' device time (max over ranks)" % world,
                     "gravity_ms_per_step": g_ms / steps, "springs_ms_per_step": s_ms / steps},
        "cpu_baseline": None,
        "kernel_ms": {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()},
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / steps, "setup": setup, "fp64_peak_tflops_measured": fp64_peak,
    }


# ---------------------------------------------------------------------------------------------------
# the ensemble
# ---------------------------------------------------------------------------------------------------
def run_c5(args, A, rank, world, local, dist, steps=None, warmup=None):
    """BASELINE config 5: 4096 independent ~1k-node bodies (C1 generator), omega_z swept 0.1 -> 0.9, k in
    {0.02, 0.04, 0.08, 0.16}, GRAVITY_NONE + zero_accel + springs; members dealt to the ranks by cost, no collective."""
    steps = steps or args.steps
    warmup = warmup or args.warmup
    total = int(os.environ.get("ASS_C5_BODIES", "4096"))
    distinct = int(os.environ.get("ASS_C5_DISTINCT_SEEDS", "64"))   # bodies are regenerated per seed; 64 shapes x sweeps
    d = WORKLOADS["C1"]["d"]
    t0 = time.perf_counter()
    shapes = {}
    for sd in range(min(distinct, total)):
        A.srand(SEED + sd)
        p = A.rand_ellipsoid(d, *ABC, 1.0)
        with A.Sim() as s:
            s.upload(p)
            sp = s.connect_springs_dist(2.3 * d, 0, len(p), 1.0, GAMMA)
            s.subtract_com(0, len(p)); s.subtract_cov(0, len(p))
            s.download(p)
        shapes[sd] = (p, sp)
    # a member's cost is its spring count (the spring phase is the step): longest-processing-time-first dealing
    costs = [len(shapes[b % distinct][1]) for b in range(total)]
    mine = A.launcher.deal_members_by_cost(costs, world, rank)

    def member(b):
        p, sp = shapes[b % distinct]
        wz = 0.1 + 0.8 * b / max(total - 1, 1)
        k = (0.02, 0.04, 0.08, 0.16)[b % 4]
        q = p.copy()
        q[:, 3] = -wz * q[:, 1]; q[:, 4] = wz * q[:, 0]; q[:, 5] = 0.0      # spin_body about z for a centred body
        s2 = sp.copy(); s2["k"] = k
        return q, s2
    bodies, springs, params = [], [], []
    for b in mine:
        q, s2 = member(b)
        bodies.append(q); springs.append(s2)
        params.append(dict(dt=WORKLOADS["C1"]["dt"], gravity=0, additional_forces=2, heartbeat=0))
    setup_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ens = A.Ensemble(bodies, springs, params)
    create_s = time.perf_counter() - t0
    nodes = ens.total_nodes
    clocks = ClockSampler(local).start()
    ens.step(warmup)
    ens.sync()
    barrier(dist)
    launches0 = A.launch_count()
    ms = ens.step(steps, timed=True)
    barrier(dist)
    clk = clocks.stop()
    launches = A.launch_count() - launches0
    ms_max = max_over_ranks(dist, ms, local)
    per_rank_ms = all_ranks(dist, ms / steps)
    per_rank_members = all_ranks(dist, len(mine))
    total_nodes = sum_over_ranks(dist, float(nodes), local)
    total_springs = sum_over_ranks(dist, float(ens.total_springs), local)
    value = total_nodes * steps / (ms_max * 1e-3)
    # e2e: state is resident by design (that is the point of the ensemble kernel); every step's result is read back
    barrier(dist)
    k_e2e = min(steps, 20)
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        ens.step(1)
        ens.download(A.F_POS | A.F_VEL | A.F_ACC)
    t_mine = time.perf_counter() - t0
    barrier(dist)
    t_e2e = max_over_ranks(dist, t_mine, local)
    # parity: this rank's first member, re-run from its initial state -- bit for bit against the single-body path, and
    # against the oracle within the usual bars
    parity = None
    if rank == 0:
        parity = {"member": int(mine[0]), "steps": 3}
        try:
            q, s2 = member(mine[0])
            with A.Ensemble([q], [s2], [params[0]]) as e1:
                e1.step(3)
                got = e1.download()[0]
            with A.Sim() as s:
                s.upload(q); s.set_springs(s2); s.set_params(dt=WORKLOADS["C1"]["dt"], gravity=0, additional_forces=2, heartbeat=0)
                s.step(3)
                one = s.download(q.copy())
            parity["bitwise_equal_to_single_body_path"] = bool(np.array_equal(got[:, :10].view(np.uint64), one[:, :10].view(np.uint64)))
            from oracle.bindings import Oracle, make_sim
            ref = np.ascontiguousarray(q[:, :11]).copy()
            Oracle().step(ref, s2, make_sim(dt=WORKLOADS["C1"]["dt"], gravity_basic=0, zero_first=1), 3)
            parity["acc_vs_oracle"] = float(np.abs(got[:, 6:9] - ref[:, 6:9]).max() / np.abs(ref[:, 6:9]).max())
            parity["xv_vs_oracle_max_abs"] = float(np.abs(got[:, :6] - ref[:, :6]).max())
            parity["ok"] = bool(parity["bitwise_equal_to_single_body_path"] and parity["acc_vs_oracle"] <= 1e-11 and parity["xv_vs_oracle_max_abs"] <= POS_TOL)
        except Exception as e:
            parity["error"] = repr(e)
            parity["ok"] = False
    ens.close()
    if rank != 0:
        return None
    peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg = (32.0 * total_springs + 80.0 * total_nodes) / world              # per GPU per step
    ach = alg * steps / (ms_max * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cfg1 = dict(WORKLOADS["C1"])
            v, sample, cores, _ = reference_arm_value(cfg1, np.ascontiguousarray(bodies[0][:, :11]), springs[0], budget_s=5.0, omp=False)
            cpu = {"value": v, "unit": "node-steps/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "reference",
                   "sample": "one member of the ensemble (N=%d) stepped by the reference; the ensemble is that x %d" % (len(bodies[0]), total)}
        except FileNotFoundError as e:
            cpu = {"value": None, "unit": "node-steps/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    return {"metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5: ensemble spin-rate/material sweep, %d independent ~1k-node bodies (C1 generator, %d distinct seeds), GRAVITY_NONE + zero_accel + springs" % (total, distinct),
                       "bodies": total, "nodes_total": int(total_nodes), "springs_total": int(total_springs),
                       "l2": "inputs larger than L2: %.0f MB of half-edge records per GPU streamed every step" % (64.0 * total_springs / world / 1e6),
                       "parallelism": "ensemble members dealt to %d GPU(s) by cost (longest-processing-time first), no collective" % world},
            "dealing": {"members_per_gpu": per_rank_members, "per_gpu_ms_per_step": per_rank_ms},
            "clocks": clk,
            "e2e": {"value": total_nodes * k_e2e / t_e2e, "unit": "node-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(total_nodes * 80), "note": "state resident (the point of the ensemble kernel); every step's x, v, a read back", "steps": k_e2e},
            "gpu_launches": int(launches), "parity": parity,
            "roofline": {"kernel": "k_ensemble (spring phase)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": None, "algorithmic_bytes_per_step_per_gpu": alg},
            "cpu_baseline": cpu, "setup": {"build_bodies_s": setup_s, "ensemble_create_s": create_s}}


# ---------------------------------------------------------------------------------------------------
# the reference arm
# ---------------------------------------------------------------------------------------------------
def reference_arm_value(cfg, p, springs, budget_s, omp):
    """Time the UNMODIFIED reference's reb_step on this host.  Gravity is O(N^2): to stay inside the budget it is
    timed with N_active = N/f sources (the reference's own loop bound, gravity.c:64,89) and scaled by f; springs and
    leapfrog are timed in full.  Returns (node-steps/s, description, cores)."""
    from oracle.bindings import Ref
    n = len(p)
    r = Ref(omp=omp)
    try:
        r.add_particles(p)
        r.set_springs(springs)
        cores = (os.cpu_count() or 1) if omp else 1
        # non-gravity part of the step: part1 + additional_forces + part2 with gravity NONE
        r.config(gravity_basic=0, G=1.0, softening=cfg["d"] / 100.0, dt=cfg["dt"], zero_first=1 if cfg["af"] == 2 else 0)
        r.step(1)
        t0 = time.perf_counter()
        reps = 3
        r.step(reps)
        t_rest = (time.perf_counter() - t0) / reps
        t_grav = 0.0
        sample = "springs+leapfrog in full"
        if cfg["gravity"]:
            # calibrate the pair rate on a small source count, then size the sample to the budget
            r.config(gravity_basic=1, G=1.0, softening=cfg["d"] / 100.0, dt=cfg["dt"], zero_first=0)
            probe_src = max(1, min(n, int(2e8 / n)))
            r.lib.ref_set_n_active(r.h, probe_src)
            t0 = time.perf_counter()
            r.gravity()
            rate = n * probe_src / max(time.perf_counter() - t0, 1e-9)
            src = int(min(n, max(probe_src, budget_s * rate / n)))
            r.lib.ref_set_n_active(r.h, src)
            t0 = time.perf_counter()
            r.gravity()
            t_s = time.perf_counter() - t0
            t_grav = t_s * (n / src)
            sample = "gravity.c timed on N_active=%d of %d sources (%.3g of %.3g ordered pairs, %.1f s) and scaled x%.2f; %s" % (
                src, n, float(n) * src, float(n) * n, t_s, n / src, sample)
        return n / (t_rest + t_grav), sample, cores, {"t_rest_s": t_rest, "t_gravity_s": t_grav}
    finally:
        r.close()


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path, all host threads it can correctly use. Rank 0 only.  The body is
    the one the b200 arm steps (at every N: the b200 arm is strong-scaled), built by the CPU oracle's generator and
    connector -- nothing of libass_b200.so is loaded in this arm."""
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its children; the reference arm is meant to use every host thread it can
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    name = args.workload or "C3"
    if name == "C5":
        name = "C1"
    cfg = WORKLOADS[name]
    from oracle.bindings import Oracle
    orc = Oracle()
    orc.srand(SEED)
    p = orc.rand_ellipsoid(cfg["d"], *ABC, 1.0)
    n_body = len(p)
    springs = orc.connect_springs_dist(p, 2.3 * cfg["d"], 0, n_body, K_SPRING, GAMMA)
    orc.subtract_com(p, 0, n_body); orc.subtract_cov(p, 0, n_body); orc.spin_body(p, 0, n_body, cfg["omega"])
    if name == "C4":
        pm = np.zeros((1, p.shape[1]))
        pm[0, 0] = 7.0; pm[0, 4] = np.sqrt(2.0 / 7.0); pm[0, 9] = 1.0
        p = np.ascontiguousarray(np.vstack([p, pm]))
    total = args.steps + args.warmup
    budget = max(2.0, min(20.0, 150.0 / max(total, 1)))
    vals = []
    detail = None
    for i in range(total):
        v, sample, cores, detail = reference_arm_value(cfg, p, springs, budget, omp=True)
        if i >= args.warmup:
            vals.append(v)
    value = len(p) * len(vals) / sum(len(p) / v for v in vals)
    out = {"impl": "reference", "metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * len(p) / value, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_of(name, cfg, len(p), len(springs), max(args.gpus, 1), "slabs" if args.gpus > 1 else "single"),
           "cpu_baseline": {"value": value, "unit": "node-steps/s", "cores": cores, "kind": "reference",
                            "sample": sample + "; built -O3 -march=x86-64-v3 (no AVX-512: the path is scalar sqrt/div)"},
           "e2e": {"value": value, "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "detail": detail}
    emit(out)


# ---------------------------------------------------------------------------------------------------
# the drop-in itself: ONE module source linked against the reference and against the drop-in (tests/dropin)
# ---------------------------------------------------------------------------------------------------
def run_dropin_e2e(p0, springs, cfg, steps_b200=50, label=""):
    """Wall time of reb_integrate inside tests/dropin/_build/synth_ref (the unmodified reference, serial as shipped) and
    synth_b200 (the same module source on the drop-in) on the body of the main record, read from the reference's own
    restart files (modules/asteroid_spin's path: read_particles / read_springs).  One print at t = 0 (write_resolved_with_E
    with its O(N^2) potential energy) + the step loop.  The reference runs ONE step (it needs about a minute per step at
    this size); the drop-in runs `steps_b200` in coherent (every call self-contained: upload, run, download) and resident
    mode.  Returns None where the binaries are absent."""
    import shutil
    import tempfile
    build = os.path.join(ROOT, "tests", "dropin", "_build")
    bins = {k: os.path.join(build, k) for k in ("synth_ref", "synth_b200")}
    if not all(os.path.exists(b) for b in bins.values()):
        return {"unavailable": "tests/dropin/_build/synth_ref / synth_b200 not built (needs the reference tree at build time)"}
    n, ns = len(p0), len(springs)
    work = tempfile.mkdtemp(prefix="ass_dropin_")
    try:
        with open(os.path.join(work, "body_000000_particles.txt"), "w") as f:
            for r in p0:   # x y z vx vy vz r m (input_spring.cpp:108-136)
                f.write("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n" % (r[0], r[1], r[2], r[3], r[4], r[5], r[10] if p0.shape[1] > 10 else 0.0, r[9]))
        with open(os.path.join(work, "body_000000_springs.txt"), "w") as f:
            for q in springs:   # i j k rs0 gamma (input_spring.cpp:77-105)
                f.write("%d %d %.17g %.17g %.17g\n" % (q["particle_1"], q["particle_2"], q["k"], q["rs0"], q["gamma"]))
        base = open(os.path.join(ROOT, "tests", "dropin", "problem.cfg")).read()

        def run(binary, nsteps, env_extra):
            d = os.path.join(work, binary + "_" + env_extra.get("ASS_RESIDENT", "x"))
            os.makedirs(d, exist_ok=True)
            c = base
            fl = lambda x: repr(float(x))   # libconfig reads "5" as an int and refuses it to a double lookup: always "5.0"
            for key, val in (("dt", fl(cfg["dt"])), ("t_max", fl(nsteps * cfg["dt"])), ("print_interval", "1.0e9"),
                             ("min_particle_distance", fl(cfg["d"])), ("max_spring_dist", fl(2.3 * cfg["d"])),
                             ("k", fl(K_SPRING)), ("gamma", fl(GAMMA)), ("damp_fac", "1.0"), ("t_damp", "1.0e9"),
                             ("gravity", "%d" % cfg["gravity"]), ("restart_index", "0")):
                c = __import__("re").sub(r"(?m)^%s = [^;]*;" % key, "%s = %s;" % (key, val), c)
            c = c.replace('restart_root = "";', 'restart_root = "%s";' % os.path.join(work, "body"))
            open(os.path.join(d, "problem.cfg"), "w").write(c)
            env = dict(os.environ)
            env.update(env_extra)
            t0 = time.perf_counter()
            pr = subprocess.run([bins[binary]], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
            wall = time.perf_counter() - t0
            if pr.returncode != 0:
                return {"error": pr.stdout[-400:]}
            tm = dict(ln.split() for ln in open(os.path.join(d, "synth_timing.txt")).read().strip().split("\n"))
            return {"steps": nsteps, "integrate_s": float(tm["integrate_s"]), "process_wall_s": wall,
                    "res_row": open(os.path.join(d, "synth_res.txt")).read().rstrip("\n").split("\n")[1]}
        out = {"body": {"N": int(n), "NS": int(ns), "workload": label}, "what": "reb_integrate of tests/dropin/problem.cpp: heartbeat print at t = 0 (write_resolved_with_E incl. grav_potential_energy) + the steps"}
        out["reference"] = run("synth_ref", 1, {})
        out["b200_same_call"] = run("synth_b200", 1, {"ASS_RESIDENT": "1"})       # the reference's call, like for like
        # steady state per step: the difference of two runs of different length (the t = 0 print and the one-time set-up
        # inside the first spring_forces -- the spring layout -- cancel)
        for key, env in (("b200_coherent", {"ASS_RESIDENT": "0"}), ("b200_resident", {"ASS_RESIDENT": "1"})):
            short, long_ = run("synth_b200", steps_b200, env), run("synth_b200", 3 * steps_b200, env)
            out[key] = long_
            if "integrate_s" in short and "integrate_s" in long_:
                out[key]["ms_per_step"] = 1e3 * (long_["integrate_s"] - short["integrate_s"]) / (2 * steps_b200)
                out[key]["shorter_run"] = {"steps": short["steps"], "integrate_s": short["integrate_s"]}
        try:
            out["speedup_same_call"] = out["reference"]["integrate_s"] / out["b200_same_call"]["integrate_s"]
            # the t = 0 print must agree: Etot column (22) of the history file
            # (columns are 14 wide; the power column after Etot overflows its width and runs into it: cut by position)
            e_ref = float(out["reference"]["res_row"][14 * 22:14 * 23]); e_gpu = float(out["b200_same_call"]["res_row"][14 * 22:14 * 23])
            out["etot_t0_rel_diff"] = abs(e_gpu - e_ref) / max(abs(e_ref), 1e-300)
        except Exception as e:
            out["note"] = "incomplete: %r" % e
        for k in ("reference", "b200_same_call", "b200_coherent", "b200_resident"):
            if isinstance(out.get(k), dict):
                out[k].pop("res_row", None)
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything any library prints on fd 1 during the run (NCCL's version banner comes from C, past sys.stdout) goes
    to stderr; emit() writes the one JSON line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


def sub_record(fn, rank, dist, label):
    """A sub-record must never take the main record down with it: failures are reported inside the sub-record.
    Every rank takes part (the calls are collective); rank 0 returns the record."""
    try:
        rec = fn()
    except Exception as e:
        traceback.print_exc()
        rec = {"error": "%s failed on rank %d: %r" % (label, rank, e)}
        # the other ranks may be inside a collective: nothing sensible can be salvaged from this sub-record
    return rec


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=[None, "C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="main record only (skip the c4_sharded / c5_ensemble sub-records)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the dropin_e2e record")
    ap.add_argument("--dropin-c3", action="store_true", help="time the drop-in against the reference on the C3 body (the reference needs ~5.5 minutes)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    rank, world, local, dist = dist_setup(args.gpus)
    import asteroid_spring_sims_b200 as A
    A.init(local)
    name = args.workload or "C3"

    def body_record(nm, **kw):
        cfg = dict(WORKLOADS[nm])
        if nm == "C4" and os.environ.get("ASS_C4_D"):
            cfg["d"] = float(os.environ["ASS_C4_D"])      # a smaller body of the same kind, for quick checks
        if world > 1:
            return run_sharded(args, A, nm, cfg, rank, world, local, dist, **kw)
        return run_single(args, A, nm, cfg, local, **kw)

    if name == "C5":
        out = run_c5(args, A, rank, world, local, dist)
    else:
        out = body_record(name)
        if args.workload is None and not args.no_sub:
            c4 = sub_record(lambda: body_record("C4", want_e2e=False, **({} if world > 1 else {"want_cpu": False})), rank, dist, "c4_sharded")
            c5 = sub_record(lambda: run_c5(args, A, rank, world, local, dist), rank, dist, "c5_ensemble")
            if rank == 0:
                out["c4_sharded"] = c4
                out["c5_ensemble"] = c5
    if rank == 0:
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
