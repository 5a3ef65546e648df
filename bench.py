#!/usr/bin/env python
"""bench.py -- node-steps/s of the per-step hot path (springs + direct-sum self-gravity + leapfrog kick/drift).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5] [--impl reference]

Prints ONE JSON line (rank 0).  Default workload: BASELINE config C3 -- a wobble_bennu2-style tumbling viscoelastic
body, synthetic N ~ 1e5 nodes, REB_GRAVITY_BASIC + spring_forces + leapfrog -- the size north_star quotes its target on.
With --gpus N > 1 (launched by torchrun, one process per GPU) the default is N independent C3 ensemble members with a
swept spin rate, one per GPU, no data-path collective ("scaling": "weak"); --workload C4 runs the N = 1e6 body sharded
over the ranks instead ("strong"), --workload C5 the 4096 x 1k-node ensemble.

value      whole-job node-steps/s, state resident in HBM, device-timed (CUDA events on the simulation's stream),
           max over ranks; L2 flushed between steps.
e2e        the same through the host-facing call sequence a drop-in reb_step makes in its always-coherent mode:
           every step uploads the host `struct reb_particle`-strided array, steps once, downloads x, v, a.
roofline   the dominant kernel (gravity at C2..C4, springs at C1/C5) against the FP64 DFMA peak measured in this run /
           the HBM copy peak of MEASURED_PEAKS.json.
cpu_baseline  the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile) timed on this
           box's host on a bounded sample of the same step, serial as shipped.
--impl reference  the reference's own CPU implementation with every host thread it can correctly use (OpenMP on
           gravity.c / integrator_leapfrog.c only; its src_spring pragmas are racy).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345
ABC = (1.0, 0.7, 0.5)
K_SPRING, GAMMA = 0.08, 5.0          # modules/wobble_bennu2/problem.cfg

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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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


def build_body(A, cfg, omega, with_point_mass=False):
    """Synthesise a BASELINE config from the module's code path: rand_ellipsoid -> connect_springs_dist ->
    subtract_com/cov -> spin_body (modules/wobble_bennu2/problem.cpp:63-333), everything through the product."""
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
    sim.set_springs(springs)
    sim.set_params(t=0.0, dt=cfg["dt"], G=1.0, softening=cfg["d"] / 100.0, gravity=cfg["gravity"], n_active=-1,
                   n_perts=n_perts, additional_forces=cfg["af"], heartbeat=0)
    return sim, p, springs, {"gen_s": t_gen, "connect_s": t_connect}


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
    """--impl reference: the reference's own CPU path, all host threads it can correctly use. Rank 0 only."""
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its children; the reference arm is meant to use every host thread it can
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import asteroid_spring_sims_b200 as A  # host-side generator only (no GPU needed): builds the same body
    name = args.workload or "C3"
    cfg = WORKLOADS[name]
    A.srand(SEED)
    p = A.rand_ellipsoid(cfg["d"], *ABC, 1.0)
    from oracle.bindings import Oracle
    orc = Oracle()
    springs = orc.connect_springs_dist(p, 2.3 * cfg["d"], 0, len(p), K_SPRING, GAMMA)
    orc.subtract_com(p, 0, len(p)); orc.subtract_cov(p, 0, len(p)); orc.spin_body(p, 0, len(p), cfg["omega"])
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
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s: %s" % (name, cfg["desc"]), "N": len(p), "NS": len(springs)},
           "cpu_baseline": {"value": value, "unit": "node-steps/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": value, "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "detail": detail}
    emit(out)


def finish_line(out, dist):
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


def run_c5(args, A, rank, world, local, dist):
    """BASELINE config 5: 4096 independent ~1k-node bodies (C1 generator), omega_z swept 0.1 -> 0.9, k in
    {0.02, 0.04, 0.08, 0.16}, GRAVITY_NONE + zero_accel + springs; members dealt round-robin to the ranks, no collective."""
    total = int(os.environ.get("ASS_C5_BODIES", "4096"))
    distinct = int(os.environ.get("ASS_C5_DISTINCT_SEEDS", "64"))   # bodies are regenerated per seed; 64 shapes x sweeps
    d = WORKLOADS["C1"]["d"]
    mine = A.launcher.deal_members(total, world, rank)
    t0 = time.perf_counter()
    shapes = {}
    for sd in sorted({b % distinct for b in mine}):
        A.srand(SEED + sd)
        p = A.rand_ellipsoid(d, *ABC, 1.0)
        with A.Sim() as s:
            s.upload(p)
            sp = s.connect_springs_dist(2.3 * d, 0, len(p), 1.0, GAMMA)
            s.subtract_com(0, len(p)); s.subtract_cov(0, len(p))
            s.download(p)
        shapes[sd] = (p, sp)
    bodies, springs, params = [], [], []
    for b in mine:
        p, sp = shapes[b % distinct]
        wz = 0.1 + 0.8 * b / max(total - 1, 1)
        k = (0.02, 0.04, 0.08, 0.16)[b % 4]
        q = p.copy()
        q[:, 3] = -wz * q[:, 1]; q[:, 4] = wz * q[:, 0]; q[:, 5] = 0.0      # spin_body about z for a centred body
        s2 = sp.copy(); s2["k"] = k
        bodies.append(q); springs.append(s2)
        params.append(dict(dt=WORKLOADS["C1"]["dt"], gravity=0, additional_forces=2, heartbeat=0))
    setup_s = time.perf_counter() - t0
    ens = A.Ensemble(bodies, springs, params)
    nodes = ens.total_nodes
    ens.step(args.warmup)
    ens.sync()
    clocks = ClockSampler(local)
    barrier(dist)
    launches0 = A.launch_count()
    clocks.start()
    ms = ens.step(args.steps, timed=True)
    barrier(dist)
    clk = clocks.stop()
    launches = A.launch_count() - launches0
    ms_max = max_over_ranks(dist, ms, local)
    total_nodes = sum_over_ranks(dist, float(nodes), local)
    total_springs = sum_over_ranks(dist, float(ens.total_springs), local)
    value = total_nodes * args.steps / (ms_max * 1e-3)
    # e2e: state is resident by design (that is the point of the ensemble kernel); every step's result is read back
    barrier(dist)
    t0 = time.perf_counter()
    for _ in range(min(args.steps, 20)):
        ens.step(1)
        ens.download(A.F_POS | A.F_VEL | A.F_ACC)
    barrier(dist)
    t_e2e = max_over_ranks(dist, time.perf_counter() - t0, local)
    if rank != 0:
        return
    peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg = (32.0 * total_springs + 80.0 * total_nodes) / world              # per GPU per step
    ach = alg * args.steps / (ms_max * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cfg1 = dict(WORKLOADS["C1"])
            v, sample, cores, _ = reference_arm_value(cfg1, np.ascontiguousarray(bodies[0]), springs[0], budget_s=5.0, omp=False)
            cpu = {"value": v, "unit": "node-steps/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "reference",
                   "sample": "one member of the ensemble (N=%d) stepped by the reference; the ensemble is that x %d" % (len(bodies[0]), total)}
        except FileNotFoundError as e:
            cpu = {"value": None, "unit": "node-steps/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    out = {"metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "C5: ensemble spin-rate/material sweep, %d independent ~1k-node bodies (C1 generator, %d distinct seeds), GRAVITY_NONE + zero_accel + springs" % (total, distinct),
                      "bodies": total, "bodies_per_gpu": len(mine), "nodes_total": int(total_nodes), "springs_total": int(total_springs),
                      "l2": "inputs larger than L2: %.0f MB of half-edge records per GPU streamed every step" % (64.0 * total_springs / world / 1e6),
                      "parallelism": "ensemble x%d, no collective" % world},
           "clocks": clk,
           "e2e": {"value": total_nodes * min(args.steps, 20) / t_e2e, "unit": "node-steps/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": int(nodes * 80), "note": "state resident (the point of the ensemble kernel); every step's x, v, a read back"},
           "gpu_launches": int(launches),
           "roofline": {"kernel": "k_ensemble (spring phase)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": None, "algorithmic_bytes_per_step_per_gpu": alg},
           "cpu_baseline": cpu, "setup": {"build_bodies_s": setup_s}}
    finish_line(out, dist)


def run_c4_sharded(args, A, rank, world, local, dist):
    """BASELINE config 4: one N ~ 1e6 body (+ a Mars-like point mass) split into particle slabs over the ranks; the only
    exchange is the per-step pull of the other slabs' x and v through CUDA-IPC peer mappings (strong scaling)."""
    cfg = dict(WORKLOADS["C4"])
    if os.environ.get("ASS_C4_D"):
        cfg["d"] = float(os.environ["ASS_C4_D"])
    sim, p, springs, setup = build_body(A, cfg, cfg["omega"], with_point_mass=True)
    n, ns = len(p), len(springs)
    bounds = A.slab_bounds(n, world)
    sim.shard_setup(rank, world, bounds)
    A.launcher.exchange_windows(dist, sim, rank, world)

    def run(nsteps):
        A.launcher.sharded_steps(dist, sim, nsteps)

    run(args.warmup)
    fp64_peak, _ = A.probe_fp64_peak(1 << 13)
    sim.timing(True); sim.timing_reset()
    clocks = ClockSampler(local)
    dist.barrier()
    launches0 = A.launch_count()
    clocks.start()
    sim.sync()
    t0 = time.perf_counter()
    run(args.steps)
    sim.sync()
    dist.barrier()
    t_wall = max_over_ranks(dist, time.perf_counter() - t0, local)
    clk = clocks.stop()
    launches = A.launch_count() - launches0
    kt = sim.timing_get()
    sim.timing(False)
    g_ms = max_over_ranks(dist, kt["gravity"][0], local)
    value = n * args.steps / t_wall
    # e2e: the sharded body's natural host interaction is a gather + download at print cadence; per step that is
    # x, v of the own slab (the module's heartbeat) -- measured as one full download every step
    host = np.zeros((n, 16))
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(min(args.steps, 3)):
        A.launcher.sharded_steps(dist, sim, 1, pre_drift=False)
        sim.download(host, A.F_POS | A.F_VEL | A.F_ACC)
    dist.barrier()
    t_e2e = max_over_ranks(dist, time.perf_counter() - t0, local)
    if rank != 0:
        return
    pairs = float(n) * float(n) / world
    ach = 20.0 * pairs * args.steps / (g_ms * 1e-3) / 1e12
    out = {"metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "C4: %s, slabs over %d GPUs" % (cfg["desc"], world), "N": n, "NS": ns, "bounds": bounds,
                      "exchange": "per step: pull of the other slabs' x and v (%.1f MB, overlapped with the own-slab gravity triangle) and of the partial accelerations the peers computed for this slab (%.1f MB) over CUDA-IPC peer mappings, copy engines; 3 host barriers" % (
                          64e-6 * (n - (bounds[1] - bounds[0])), 32e-6 * (bounds[1] - bounds[0]) * (world // 2)),
                      "gravity": "divided by pairs of slabs: every unordered particle pair evaluated once, the same work on every rank",
                      "l2": "inputs larger than L2 (half-edges %.0f MB)" % (64e-6 * ns), "parallelism": "particle slabs x%d" % world},
           "clocks": clk, "gpu_launches": int(launches),
           "e2e": {"value": n * min(args.steps, 3) / t_e2e, "unit": "node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(n * 80)},
           "roofline": {"kernel": "k_gravity", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": None,
                        "note": "per GPU: 20 flop x N x N/%d ordered pairs per step (SURVEY 8d convention; the pair-once kernels execute 9.5 FP64 instructions per ordered pair) over the gravity kernels' device time (max over ranks)" % world},
           "cpu_baseline": None,
           "kernel_ms": {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()}, "setup": setup}
    finish_line(out, dist)


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
    if name == "C5":
        return run_c5(args, A, rank, world, local, dist)
    if name == "C4" and world > 1:
        return run_c4_sharded(args, A, rank, world, local, dist)
    cfg = dict(WORKLOADS[name])
    if name == "C4" and os.environ.get("ASS_C4_D"):
        cfg["d"] = float(os.environ["ASS_C4_D"])      # a smaller body of the same kind, for quick checks
    # ensemble member `rank`: spin rate swept across members (BASELINE config 5's sweep, applied to bodies of this size)
    omega = tuple(np.array(cfg["omega"]) * (1.0 + 0.1 * rank))
    sim, p, springs, setup = build_body(A, cfg, omega, with_point_mass=(name == "C4"))
    n, ns = len(p), len(springs)

    # ---- device-resident throughput ----
    sim.step(args.warmup)
    sim.sync()
    fp64_peak, _ = A.probe_fp64_peak(1 << 13)
    sim.timing(True)
    sim.timing_reset()
    clocks = ClockSampler(local)
    barrier(dist)
    launches0 = A.launch_count()
    clocks.start()
    sim.sync()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        sim.l2_flush()
        sim.mark_begin()
        sim.step(1)
        dev_ms += sim.mark_end()
    sim.sync()
    barrier(dist)
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    launches = A.launch_count() - launches0
    kt = sim.timing_get()
    sim.timing(False)
    dev_ms_max = max_over_ranks(dist, dev_ms, local)
    total_nodes = sum_over_ranks(dist, float(n), local)
    value = total_nodes * args.steps / (dev_ms_max * 1e-3)

    # ---- end to end through host buffers (the drop-in's always-coherent reb_step) ----
    e2e = None
    if not args.no_e2e:
        # 128-byte rows == sizeof(struct reb_particle), in page-locked host memory (ass_host_alloc): what the drop-in's
        # reb_integrate gives the module's own r->particles by pinning it in place (ass_host_register)
        pinned = A.PinnedRows(n, 16)
        host = pinned.a
        host[:] = 0.0
        sim.download(host)
        for _ in range(2):
            sim.upload(host); sim.step(1); sim.download(host, A.F_POS | A.F_VEL | A.F_ACC)
        barrier(dist)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sim.upload(host)
            sim.step(1)
            sim.download(host, A.F_POS | A.F_VEL | A.F_ACC)
        barrier(dist)
        t_e2e = max_over_ranks(dist, time.perf_counter() - t0, local)
        e2e = {"value": total_nodes * args.steps / t_e2e, "unit": "node-steps/s",
               "h2d_bytes_per_step": int(n * 128), "d2h_bytes_per_step": int(n * 72), "host_memory": "pinned (ass_host_alloc)"}
        del host
        pinned.close()

    if rank != 0:
        return

    # ---- roofline of the dominant kernel ----
    peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic = (read_json(os.path.join(ROOT, "profiles", "traffic.json")) or {}).get(name, {})
    g_ms, g_n = kt["gravity"]
    s_ms, s_n = kt["springs"]
    roof_g = roof_s = None
    if g_n:
        n_src = n
        # one gravity evaluation per step = the pair kernel + its fixed-order row reduction (both inside the timer)
        ach = 20.0 * n * n_src * args.steps / (g_ms * 1e-3) / 1e12    # 20 flop per ordered pair
        roof_g = {"kernel": "k_gravity (+k_gravity_reduce)", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                  "traffic": traffic.get("gravity"), "peak_source": "DFMA microbenchmark measured in this run (ass_probe_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                  "pairs_per_s": n * n_src * args.steps / (g_ms * 1e-3), "ms_per_evaluation": g_ms / args.steps}
    if s_n:
        alg = 32.0 * ns + 80.0 * n
        ach = alg * s_n / (s_ms * 1e-3) / 1e9
        roof_s = {"kernel": "k_spring_tiles<KICK> (springs + kick + drift)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                  "traffic": traffic.get("springs"), "peak_source": hbm_src, "algorithmic_bytes_per_launch": alg, "ms_per_launch": s_ms / s_n}
    dominant = roof_g if (roof_g and g_ms >= s_ms) else roof_s

    # ---- CPU baseline: the unmodified reference on this host, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, sample, cores, detail = reference_arm_value(cfg, np.ascontiguousarray(p), springs, budget_s=12.0, omp=False)
            cpu = {"value": v, "unit": "node-steps/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "reference", "sample": sample}
        except FileNotFoundError as e:
            cpu = {"value": None, "unit": "node-steps/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    out = {
        "metric": "node-steps/s", "value": value, "unit": "node-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s%s" % (name, cfg["desc"], "" if world == 1 else " x %d independent ensemble members, one per GPU (swept spin rate)" % world),
                   "N_per_gpu": n, "NS_per_gpu": ns, "seed": SEED, "generator": "rand_ellipsoid(d=%g, 1, 0.7, 0.5)" % cfg["d"],
                   "max_spring_dist": 2.3 * cfg["d"], "k": K_SPRING, "gamma": GAMMA, "dt": cfg["dt"], "softening": cfg["d"] / 100.0,
                   "l2": "flushed between steps (256 MiB fill, outside the device-timed bracket)", "parallelism": "ensemble x%d" % world},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": dominant, "roofline_gravity": roof_g, "roofline_springs": roof_s, "cpu_baseline": cpu,
        "kernel_ms": {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()},
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps, "setup": setup, "fp64_peak_tflops_measured": fp64_peak,
    }
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
