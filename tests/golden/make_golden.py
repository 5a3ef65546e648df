"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference (oracle/_ref/libref_oracle.so).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py [--c2]
The fixtures are committed; the GPU box has no /root/reference and reads only these files.
Every array stored here was produced by the reference's own functions (through oracle/ref_shim.cpp); nothing is
recomputed in Python.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.bindings import Ref  # noqa: E402

SEED = 12345
A, B, C_ = 1.0, 0.7, 0.5
K, GAMMA = 0.08, 5.0           # modules/wobble_bennu2/problem.cfg values
OMEGA = (0.1, 0.0, 0.2)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kat3():
    """modules/tests/tests.cpp fixtures SpringTest / PhysicsTest (:60-150) run through the reference itself."""
    out = {}
    P = np.zeros((3, 11))
    P[0, :3] = (0, 0, 0); P[1, :3] = (1, 2, 2); P[2, :3] = (2, 3, 6)
    P[:, 3:6] = P[:, :3]
    P[:, 9] = (1, 2, 3)
    with Ref() as r:
        r.add_particles(P)
        # SpringTest.ConnectSprings: cutoff 5 -> (0,1),(1,2)
        r.connect_springs_dist(5.0, 0, 3, 11.0, 35.0)
        out["connect5"] = r.springs()
    with Ref() as r:
        r.add_particles(P)
        r.connect_springs_dist(100.0, 0, 3, 11.0, 35.0)
        out["connect100"] = r.springs()
        # SpringTest.Force (:1043-1139)
        r.set_gamma(0.0)
        Q = r.get(); Q[0, :3] = (1, 1, 4); r.set(Q)
        out["force_in"] = r.get()
        out["f1_undamped"] = r.spring_i_force(1, damped=False)
        out["f1_damped_g0"] = r.spring_i_force(1, damped=True)
        r.zero_accel(); r.spring_forces()
        out["acc_g0"] = r.get()[:, 6:9]
        r.set_gamma(1.0)
        r.zero_accel(); r.spring_forces()
        out["acc_g1"] = r.get()[:, 6:9]
        out["f1_damped_g1"] = r.spring_i_force(1, damped=True)
        out["springs_g1"] = r.springs()
        out["diag_g1"] = r.diagnostics(0, 3, True)
    with Ref() as r:
        # PhysicsTest: accelerations = positions; CoM, CoV, KE, L, inertia, GPE with G = 1
        P2 = P.copy(); P2[:, 6:9] = P2[:, :3]
        r.add_particles(P2)
        r.config(G=1.0)
        out["phys_diag"] = r.diagnostics(0, 3, True)
        r.center_sim(0, 2)
        out["center_all"] = r.get()[:, :3]
    with Ref() as r:
        P2 = P.copy()
        r.add_particles(P2)
        r.subtract_com(0, 3)
        out["sub_com"] = r.get()[:, :6]
        r.set(P2)
        r.subtract_cov(0, 3)
        out["sub_cov"] = r.get()[:, :6]
        r.set(P2)
        r.spin_body(0, 3, (1.0, 12.0, 12.0))
        out["spin"] = r.get()[:, :6]
    # gravity + leapfrog have no test in the reference: pin them by executing it
    with Ref() as r:
        P3 = P.copy(); P3[:, 6:9] = 7.0
        r.add_particles(P3)
        r.config(gravity_basic=1, G=1.0, softening=0.01, dt=0.01)
        r.gravity()
        out["grav_acc"] = r.get()[:, 6:9]
        r.leapfrog_part1()
        out["lf1"] = r.get()
        r.leapfrog_part2()
        out["lf2"] = r.get()
        out["lf_t"] = np.array([r.t])
    np.savez_compressed(os.path.join(HERE, "kat3.npz"), **out)
    print("kat3.npz", sorted(out))


def body(name, d, nsteps_state, hist_steps, hist_every):
    """A random-ellipsoid body built and stepped by the reference."""
    out = {"d": np.array([d])}
    with Ref() as r:
        r.seed(SEED)
        n = r.rand_ellipsoid(d, A, B, C_, 1.0)
        out["gen"] = r.get()
        out["mindist"] = np.array([r.mindist(0, n)])
        ns = r.connect_springs_dist(2.3 * d, 0, n, K, GAMMA)
        out["springs"] = r.springs()
        r.subtract_com(0, n); r.subtract_cov(0, n); r.spin_body(0, n, OMEGA)
        start = r.get()
        out["start"] = start
        for tag, grav, zero_first, dt in (("none", 0, 1, 1e-2), ("basic", 1, 0, 2e-3)):
            r.set(start)
            r.set_time(0.0)
            r.set_springs(out["springs"])
            r.config(gravity_basic=grav, G=1.0, softening=d / 100.0, dt=dt, zero_first=zero_first)
            r.step(1)
            out["step1_" + tag] = r.get()
            r.step(nsteps_state - 1)
            out["step%d_%s" % (nsteps_state, tag)] = r.get()
            # history of the parity observables: E = KErot + PEspr (+ PEgrav), L
            r.set(start)
            r.set_time(0.0)
            r.config(gravity_basic=grav, G=1.0, softening=d / 100.0, dt=dt, zero_first=zero_first)
            hist = [r.diagnostics(0, n, bool(grav))]
            for _ in range(hist_steps // hist_every):
                r.step(hist_every)
                hist.append(r.diagnostics(0, n, bool(grav)))
            out["hist_" + tag] = np.array(hist)
        # reb_integrate with exact_finish_time and the center_sim heartbeat (asteroid_spin's callbacks)
        r.set(start)
        r.set_time(0.0)
        r.config(gravity_basic=0, G=1.0, softening=d / 100.0, dt=1e-2, zero_first=1, center_heartbeat=1)
        status = r.integrate(0.1234)
        out["integrate_state"] = r.get()
        out["integrate_meta"] = np.array([status, r.t, r.dt])
    out["hist_every"] = np.array([hist_every])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N", n, "NS", ns)


def lattices():
    out = {}
    with Ref() as r:
        n = r.hcp_ellipsoid(0.2, A, B, C_, 1.0)
        out["hcp"] = r.get()
    with Ref() as r:
        n = r.cubic_ellipsoid(0.2, A, B, C_, 1.0)
        out["cubic"] = r.get()
    np.savez_compressed(os.path.join(HERE, "lattices.npz"), **out)
    print("lattices", {k: v.shape for k, v in out.items()})


def orb():
    """src_spring/orb.cpp terms of the phobos* modules run by the reference: quadrupole_accel, drift_resolved, drift_bin.
    Fixture A is OrbTest's (modules/tests/tests.cpp:152-225: 3 body particles + a 400-mass perturber); fixture B is a
    330-node random ellipsoid with a Mars-like point mass and a second point mass."""
    out = {}
    P = np.zeros((4, 11))
    P[0, :3] = (0, 0, 0); P[1, :3] = (1, 2, 2); P[2, :3] = (2, 3, 6); P[3, :3] = (60, 3, -20)
    P[:, 3:6] = P[:, :3] * 0.1; P[3, 3:6] = (0.3, -1.0, 0.25)
    P[:, 9] = (1, 2, 3, 400); P[:, 6:9] = 0.5
    out["A_in"] = P
    for tag, (phi, theta) in (("pole", (0.0, 0.0)), ("tilt", (0.7, 1.1))):
        with Ref() as r:
            r.config(gravity_basic=0, G=0.7, n_perts=1); r.add_particles(P)
            r.quadrupole_accel(5.0, 50.0, phi, theta, 3)      # OrbTest.QuadAccel's J2, R_p
            out["A_quad_" + tag] = r.get()
    with Ref() as r:
        r.config(gravity_basic=0, G=0.7, n_perts=1); r.add_particles(P)
        r.drift_resolved(10.0, 0.2, 0.2, 3, 0, 3)             # OrbTest.DriftRes's arguments
        out["A_drift_resolved"] = r.get()
    with Ref() as r:
        r.config(gravity_basic=0, G=0.7, n_perts=1); r.add_particles(P)
        r.drift_bin(10.0, 0.2, 0.2, 3, 1)
        out["A_drift_bin"] = r.get()
    with Ref() as r:
        r.seed(SEED)
        n = r.rand_ellipsoid(0.15, A, B, C_, 1.0)
        r.subtract_com(0, n); r.subtract_cov(0, n); r.spin_body(0, n, OMEGA)
        body = r.get()
    pm = np.zeros((2, 11)); pm[0, :3] = (7.0, 0.3, -0.2); pm[0, 3:6] = (0.01, 0.38, 0.02); pm[0, 9] = 1.0
    pm[1, :3] = (-3.0, 9.0, 1.0); pm[1, 3:6] = (-0.2, -0.1, 0.0); pm[1, 9] = 0.05
    Q = np.vstack([body, pm]); Q[:, 6:9] = Q[:, :3] * 0.01
    out["B_in"] = Q
    with Ref() as r:
        r.config(gravity_basic=0, G=1.0, n_perts=2); r.add_particles(Q)
        r.quadrupole_accel(1.96e-3, 0.5, 0.3, 0.4, n)        # a Mars-like J2 seen from a body at a = 7
        out["B_quad"] = r.get()
        r.drift_resolved(1e-3, 1e-2, 3e-2, n, 0, n)
        r.drift_bin(1e-3, 1e-2, 3e-2, n, n + 1)
        out["B_drift"] = r.get()
    np.savez_compressed(os.path.join(HERE, "orb.npz"), **out)
    print("orb", {k: v.shape for k, v in out.items()})


def shape_vertices():
    """A deterministic closed surface of 162 vertices: a lumpy ellipsoid (what a .tab shape model is to rand_shape)."""
    v = []
    for i in range(9):
        th = np.pi * (i + 0.5) / 9
        for j in range(18):
            ph = 2 * np.pi * j / 18
            rr = 1.0 + 0.15 * np.sin(3 * th) * np.cos(2 * ph)
            v.append((rr * 1.0 * np.sin(th) * np.cos(ph), rr * 0.8 * np.sin(th) * np.sin(ph), rr * 0.6 * np.cos(th)))
    P = np.zeros((len(v), 11))
    P[:, :3] = np.array(v)
    P[:, 9] = 1.0
    return P


def extras():
    """Generator variants (shapes.cpp:66-206, 435-516) and the set-up rotations (physics.cpp:249-385), run by the reference."""
    out = {}
    with Ref() as r:
        r.seed(SEED)
        r.rand_rectangle(0.2, 2.0, 1.4, 1.0, 1.0)
        out["rect"] = r.get()
    with Ref() as r:
        r.seed(SEED)
        r.rand_cone(0.2, 1.0, 1.5, 1.0)
        out["cone"] = r.get()
    V = shape_vertices()
    out["shape_vertices"] = V
    with Ref() as r:
        r.add_particles(V)
        r.seed(SEED)
        r.rand_shape(0.2, 1.0)
        out["shape"] = r.get()
    g = np.load(os.path.join(HERE, "ellipsoid_d015.npz"))
    start = g["start"]
    n = len(start)
    with Ref() as r:
        r.add_particles(start)
        r.rotate_body(0, n, 0.3, -0.2, 1.1)
        out["rot_body"] = r.get()
        r.rotate_origin(0, n, -0.7, 0.4, 0.25)
        out["rot_origin"] = r.get()
        r.rotate_to_principal(0, n)
        out["rot_principal"] = r.get()
        out["euler"] = r.euler_matrix(0.3, -0.2, 1.1)
        out["euler2"] = r.euler_matrix(-0.7, 0.4, 0.25)
    # the stress tensor of every node (stress.cpp:32-85) on a real body: update_stress itself throws in its eigenvalue
    # stage there, the tensors and the strongest spring per node are complete by then
    with Ref() as r:
        r.add_particles(start)
        r.set_springs(g["springs"])
        T, F, S, threw = r.stress()
        out["stress"] = T; out["stress_max_force"] = F; out["stress_max_spring"] = S; out["stress_threw"] = np.array([threw])
    # StressTest's fixture (modules/tests/tests.cpp:1452-1644): 3 particles, zero velocities, particle 0 moved to (1,1,4)
    P = np.zeros((3, 11))
    P[0, :3] = (0, 0, 0); P[1, :3] = (1, 2, 2); P[2, :3] = (2, 3, 6)
    P[:, 9] = (1, 2, 3)
    with Ref() as r:
        r.add_particles(P)
        r.connect_springs_dist(100.0, 0, 3, 11.0, 35.0)
        Q = r.get(); Q[0, :3] = (1, 1, 4); r.set(Q)
        out["kat_stress_in"] = r.get()
        out["kat_stress_springs"] = r.springs()
        T, F, S, threw = r.stress()
        out["kat_stress"] = T; out["kat_stress_max_force"] = F; out["kat_stress_max_spring"] = S; out["kat_stress_threw"] = np.array([threw])
    np.savez_compressed(os.path.join(HERE, "extras.npz"), **out)
    print("extras", {k: v.shape for k, v in out.items()})


def c2():
    """BASELINE config C2: d = 0.0474 -> N = 9327, NS = 134162.  ~90 s of reference time (its O(N^2) paths)."""
    d = 0.0474
    with Ref() as r:
        r.seed(SEED)
        n = r.rand_ellipsoid(d, A, B, C_, 1.0)
        gen = r.get()
        ns = r.connect_springs_dist(2.3 * d, 0, n, K, GAMMA)
        sp = r.springs()
        md = r.mindist(0, n)
    meta = {"d": d, "seed": SEED, "N": int(n), "NS": int(ns), "mindist": md,
            "sha256_gen_rows": sha(gen), "sha256_springs": sha(sp),
            "first_springs": [[int(s["particle_1"]), int(s["particle_2"]), float(s["rs0"]).hex()] for s in sp[:5]],
            "last_springs": [[int(s["particle_1"]), int(s["particle_2"]), float(s["rs0"]).hex()] for s in sp[-5:]]}
    with open(os.path.join(HERE, "c2_d00474.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("c2", meta["N"], meta["NS"])


if __name__ == "__main__":
    kat3()
    lattices()
    body("ellipsoid_d015", 0.15, 20, 200, 10)
    body("c1_d010", 0.10, 20, 200, 20)
    orb()
    extras()
    if "--c2" in sys.argv:
        c2()
