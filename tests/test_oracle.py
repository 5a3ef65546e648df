"""Pin the CPU oracle (oracle/ass_oracle.c) before anything is allowed to trust it.

Three anchors (task section 3 / SURVEY 8c):
  1. the reference's own known-answer tests for this path (modules/tests/tests.cpp SpringTest.*, PhysicsTest.*),
     with the expected values as written there;
  2. golden vectors produced by the unmodified reference (tests/golden/*.npz, script tests/golden/make_golden.py);
  3. the unmodified reference itself, live, when oracle/_ref can be built here (skipped on the GPU box).
"""
import math

import numpy as np
import pytest

from conftest import kat_particles, same_bits
from oracle.bindings import SPRING_DTYPE, make_sim

A_, B_, C_ = 1.0, 0.7, 0.5
SEED = 12345
K, GAMMA = 0.08, 5.0
OMEGA = (0.1, 0.0, 0.2)
EMPTY = np.zeros(0, dtype=SPRING_DTYPE)


# ---------------------------------------------------------------------------------------------------
# 1. the reference's known-answer tests
# ---------------------------------------------------------------------------------------------------
def test_kat_connect_springs(oracle):
    """SpringTest.ConnectSprings, tests.cpp:987-994; AddSprings/DelSprings rs0, :931-973."""
    P = kat_particles()
    for n2 in (True, False):
        s = oracle.connect_springs_dist(P, 5.0, 0, 3, 11.0, 35.0, n2=n2)
        assert len(s) == 2
        assert (s["particle_1"][0], s["particle_2"][0]) == (0, 1)
        assert (s["particle_1"][1], s["particle_2"][1]) == (1, 2)
        assert s["rs0"][0] == 3.0  # SpringTest.Displacement: len == 3 exactly
        s = oracle.connect_springs_dist(P, 100.0, 0, 3, 11.0, 35.0, n2=n2)
        assert [(a, b) for a, b in zip(s["particle_1"], s["particle_2"])] == [(0, 1), (0, 2), (1, 2)]
        assert s["rs0"][1] == 7.0  # DelSprings: rs0 == 7 for (0,2)
        assert np.all(s["k"] == 11.0) and np.all(s["gamma"] == 35.0)


def test_kat_connect_duplicates_and_ranges(oracle):
    """add_spring returns the existing spring instead of adding (springs.cpp:70-79); i_high <= i_low is a no-op."""
    P = kat_particles()
    first = oracle.connect_springs_dist(P, 5.0, 0, 3, 11.0, 35.0)
    for n2 in (True, False):
        again = oracle.connect_springs_dist(P, 100.0, 0, 3, 2.0, 1.0, existing=first, n2=n2)
        assert len(again) == 3
        assert again[:2].tobytes() == first.tobytes()
        assert (again["particle_1"][2], again["particle_2"][2], again["k"][2]) == (0, 2, 2.0)
        assert len(oracle.connect_springs_dist(P, 100.0, 2, 2, 1.0, 1.0, n2=n2)) == 0
        assert len(oracle.connect_springs_dist(P, 100.0, 3, 1, 1.0, 1.0, n2=n2)) == 0
        assert len(oracle.connect_springs_dist(P, 100.0, 1, 3, 1.0, 1.0, n2=n2)) == 1


def test_kat_force(oracle):
    """SpringTest.Force, tests.cpp:1043-1139: 44*l^ undamped; damped term 0.75*(dv.l^)/len; 1e-6 relative."""
    P = kat_particles()
    s = oracle.connect_springs_dist(P, 100.0, 0, 3, 11.0, 35.0)
    s["gamma"] = 0.0
    P[0, :3] = (1, 1, 4)
    dx = P[0, :3] - P[2, :3]
    ln = np.linalg.norm(dx)
    hat = dx / ln
    ans_force = 44 * hat  # k=11, rs0=7, len=3: -11*(3-7) = 44
    # spring 0 = (0,1): len 3 -> rs0 3 ; after the move (0,1) has length |(0,-1,2)| = sqrt 5
    d01 = P[0, :3] - P[1, :3]
    f0 = -11 * (np.linalg.norm(d01) - 3.0) * d01 / np.linalg.norm(d01)
    oracle.zero_accel(P)
    oracle.spring_forces(P, s)
    assert np.allclose(P[0, 6:9], (ans_force + f0) / 1.0, rtol=1e-6)
    # particle 2: -F(0,2)/m2 - spring (1,2) which is at rest up to L_EPS
    assert np.allclose(P[2, 6:9], -ans_force / 3.0, rtol=1e-5)
    # damped: gamma = 1, m_red(0,2) = 0.75
    s["gamma"] = 1.0
    oracle.zero_accel(P)
    oracle.spring_forces(P, s)
    dv = P[0, 3:6] - P[2, 3:6]
    ans_force2 = 44 * hat - 0.75 * np.dot(dv, hat) / ln * hat
    dv01 = P[0, 3:6] - P[1, 3:6]
    h01 = d01 / np.linalg.norm(d01)
    f0d = f0 - (1 * 2 / 3.0) * np.dot(dv01, h01) / np.linalg.norm(d01) * h01
    assert np.allclose(P[0, 6:9], (ans_force2 + f0d) / 1.0, rtol=1e-5)


def test_kat_physics(oracle):
    """PhysicsTest.{ZeroAccel, CoM, CoV, CenterAll, NonTransKE, GPE, MeasureAngularMomentum, Inertia}, tests.cpp:1157-1446."""
    P = kat_particles()
    P[:, 6:9] = P[:, :3]
    d = oracle.diagnostics(P, EMPTY, 0, 3, G=1.0, want_gpe=True)
    assert tuple(d[0:3]) == (4. / 3., 13. / 6., 11. / 3.)          # CoM, EXPECT_EQ
    assert d[9] == 22.75                                           # NonTransKE, EXPECT_EQ
    # GPE: the reference's EXPECT_EQ holds for its G = 6.67e-8/F_scale...; with G = 1 the sequential sum is 1 ulp off
    assert abs(d[11] + (23. / 21. + math.sqrt(2.))) <= 2 * np.spacing(2.5)
    assert np.all(np.abs(d[6:9]) < 1e-15)                          # measure_L
    assert np.allclose(d[13:19], (253. / 6., 116. / 3., 61. / 6., -14. / 3., -32. / 3., -43. / 3.), rtol=1e-15)
    Q = P.copy(); Q[1, 3] = 2
    assert tuple(oracle.diagnostics(Q, EMPTY, 0, 3)[3:6]) == (5. / 3., 13. / 6., 11. / 3.)  # CoV
    Q = P.copy()
    oracle.center_sim(Q, 0, 2)                                     # CenterAll: CoM of first two, shift all
    assert tuple(Q[0, :3]) == (-2. / 3., -4. / 3., -4. / 3.)
    assert tuple(Q[2, :3]) == (2 - 2. / 3., 3 - 4. / 3., 6 - 4. / 3.)
    Q = P.copy()
    oracle.subtract_com(Q, 0, 3)
    assert tuple(Q[1, :3]) == (1 - 4. / 3., 2 - 13. / 6., 2 - 11. / 3.)
    assert np.all(Q[:, 3:6] == P[:, 3:6])
    oracle.zero_accel(Q)
    assert np.all(Q[:, 6:9] == 0.0)


def test_kat_spe_and_power(oracle):
    """PhysicsTest.SPE (= 88 + 3.20975 +- 1e-4) and SpringPower, tests.cpp:1395-1440."""
    P = kat_particles()
    s = oracle.connect_springs_dist(P, 100.0, 0, 3, 11.0, 35.0)
    P[0, :3] = (1, 1, 4)
    d = oracle.diagnostics(P, s, 0, 3)
    assert abs(d[10] - (88 + 3.20975)) < 1e-4
    s["gamma"] = 1.0
    dx = P[0, :3] - P[2, :3]
    ln = np.linalg.norm(dx)
    ans_pow = 0.75 * (np.dot(np.array([-2., -3., -6.]), dx / ln) / ln) ** 2
    one = s[1:2].copy()
    assert abs(oracle.diagnostics(P, one, 0, 3)[12] - ans_pow) < 1e-5


def test_kat_spin(oracle):
    """PhysicsTest.Spin, tests.cpp:1336-1352: spin_body then body_spin recovers omega; CoV unchanged."""
    P = kat_particles()
    cov0 = oracle.diagnostics(P, EMPTY, 0, 3)[3:6]
    oracle.spin_body(P, 0, 3, (1.0, 12.0, 12.0))
    d = oracle.diagnostics(P, EMPTY, 0, 3)
    I = np.array([[d[13], d[16], d[17]], [d[16], d[14], d[18]], [d[17], d[18], d[15]]])
    omega = np.linalg.solve(I, d[6:9])
    assert np.allclose(omega, (1, 12, 12), atol=1e-12)
    assert np.allclose(d[3:6], cov0, atol=1e-15)


# ---------------------------------------------------------------------------------------------------
# 2. golden vectors written by the unmodified reference
# ---------------------------------------------------------------------------------------------------
def test_golden_kat3(oracle, kat3):
    P = kat_particles()
    assert oracle.connect_springs_dist(P, 5.0, 0, 3, 11.0, 35.0).tobytes() == kat3["connect5"].tobytes()
    s = oracle.connect_springs_dist(P, 100.0, 0, 3, 11.0, 35.0)
    assert s.tobytes() == kat3["connect100"].tobytes()
    P = kat3["force_in"].copy()
    s["gamma"] = 0.0
    oracle.zero_accel(P); oracle.spring_forces(P, s)
    assert same_bits(P[:, 6:9], kat3["acc_g0"])
    s["gamma"] = 1.0
    oracle.zero_accel(P); oracle.spring_forces(P, s)
    assert same_bits(P[:, 6:9], kat3["acc_g1"])
    assert np.allclose(oracle.diagnostics(P, s, 0, 3, G=1.0, want_gpe=True), kat3["diag_g1"], rtol=1e-14, atol=1e-14)
    P = kat_particles(); P[:, 6:9] = P[:, :3]
    assert np.allclose(oracle.diagnostics(P, EMPTY, 0, 3, G=1.0, want_gpe=True), kat3["phys_diag"], rtol=1e-15, atol=1e-15)
    oracle.center_sim(P, 0, 2)
    assert same_bits(P[:, :3], kat3["center_all"])
    P = kat_particles(); oracle.subtract_com(P, 0, 3); assert same_bits(P[:, :6], kat3["sub_com"])
    P = kat_particles(); oracle.subtract_cov(P, 0, 3); assert same_bits(P[:, :6], kat3["sub_cov"])
    P = kat_particles(); oracle.spin_body(P, 0, 3, (1.0, 12.0, 12.0)); assert same_bits(P[:, :6], kat3["spin"])


def test_golden_gravity_leapfrog(oracle, kat3):
    """No test in the reference covers gravity.c or integrator_leapfrog.c: pinned by its own output."""
    P = kat_particles(); P[:, 6:9] = 7.0
    oracle.gravity_basic(P, G=1.0, softening=0.01)
    assert same_bits(P[:, 6:9], kat3["grav_acc"])
    sim = make_sim(dt=0.01, softening=0.01, gravity_basic=1)
    oracle.leapfrog_part1(P, sim)
    assert same_bits(P, kat3["lf1"])
    oracle.leapfrog_part2(P, sim)
    assert same_bits(P, kat3["lf2"])
    assert sim.t == kat3["lf_t"][0]


@pytest.mark.parametrize("fixture_name", ["body015", "body010"])
def test_golden_body(oracle, request, fixture_name):
    """Generator, spring list, setup and stepped state of a random-ellipsoid body, bit for bit."""
    g = request.getfixturevalue(fixture_name)
    d = float(g["d"][0])
    oracle.srand(SEED)
    p = oracle.rand_ellipsoid(d, A_, B_, C_, 1.0)
    assert same_bits(p, g["gen"])
    n = len(p)
    assert oracle.mindist(p) == g["mindist"][0]
    s = oracle.connect_springs_dist(p, 2.3 * d, 0, n, K, GAMMA)
    assert s.tobytes() == g["springs"].tobytes()
    oracle.subtract_com(p, 0, n); oracle.subtract_cov(p, 0, n); oracle.spin_body(p, 0, n, OMEGA)
    assert same_bits(p, g["start"])
    for tag, grav, zf, dt in (("none", 0, 1, 1e-2), ("basic", 1, 0, 2e-3)):
        q = g["start"].copy()
        sim = make_sim(dt=dt, softening=d / 100.0, gravity_basic=grav, zero_first=zf)
        oracle.step(q, s, sim, 1)
        assert same_bits(q, g["step1_" + tag])
        oracle.step(q, s, sim, 19)
        assert same_bits(q, g["step20_" + tag])
    # reb_integrate: center_sim heartbeat, last step shortened to land on tmax exactly
    q = g["start"].copy()
    sim = make_sim(dt=1e-2, softening=d / 100.0, gravity_basic=0, zero_first=1, center_heartbeat=1)
    status = oracle.integrate(q, s, sim, 0.1234)
    assert same_bits(q, g["integrate_state"])
    assert (status, sim.t, sim.dt) == tuple(g["integrate_meta"])
    assert sim.t == 0.1234


def test_golden_history(oracle, body015):
    """Energy and spin angular momentum history (SURVEY 8d parity observables) over 200 steps."""
    g = body015
    d = float(g["d"][0])
    s = g["springs"]
    n = len(g["start"])
    every = int(g["hist_every"][0])
    for tag, grav, zf, dt in (("none", 0, 1, 1e-2), ("basic", 1, 0, 2e-3)):
        q = g["start"].copy()
        sim = make_sim(dt=dt, softening=d / 100.0, gravity_basic=grav, zero_first=zf)
        ref_hist = g["hist_" + tag]
        for row in ref_hist:
            got = oracle.diagnostics(q, s, 0, n, G=1.0, want_gpe=bool(grav))
            assert np.allclose(got, row, rtol=1e-12, atol=1e-15)
            oracle.step(q, s, sim, every)


def test_golden_lattices(oracle, lattices):
    assert same_bits(oracle.hcp_ellipsoid(0.2, A_, B_, C_, 1.0), lattices["hcp"])
    assert same_bits(oracle.cubic_ellipsoid(0.2, A_, B_, C_, 1.0), lattices["cubic"])


def test_cell_list_equals_literal(oracle):
    """The large-N restatements (grid) must return what the literal O(N^2) restatements return."""
    for d in (0.3, 0.12):
        oracle.srand(7)
        a = oracle.rand_ellipsoid(d, A_, B_, C_, 1.0, n2=True)
        oracle.srand(7)
        b = oracle.rand_ellipsoid(d, A_, B_, C_, 1.0)
        assert same_bits(a, b)
        for cut in (1.5 * d, 2.3 * d, 4.0 * d):
            s1 = oracle.connect_springs_dist(a, cut, 0, len(a), 1.0, 0.5, n2=True)
            s2 = oracle.connect_springs_dist(a, cut, 0, len(a), 1.0, 0.5)
            assert s1.tobytes() == s2.tobytes()
        assert oracle.mindist(a) == oracle.mindist(a, n2=True)
    # repeated calls on sub-ranges with duplicates, as modules/cube/problem.cpp:128-137 does
    oracle.srand(11)
    a = oracle.rand_ellipsoid(0.12, A_, B_, C_, 1.0)
    n = len(a)
    for n2 in (True, False):
        s = oracle.connect_springs_dist(a, 0.2, 0, n // 2, 1.0, 0.0, n2=n2)
        s = oracle.connect_springs_dist(a, 0.3, n // 4, n, 2.0, 0.0, existing=s, n2=n2)
        s = oracle.connect_springs_dist(a, 0.25, 0, n, 3.0, 0.0, existing=s, n2=n2)
        if n2:
            want = s
        else:
            assert s.tobytes() == want.tobytes()
    # a big cloud goes through the grid mindist (count > 2048)
    oracle.srand(3)
    big = oracle.rand_ellipsoid(0.07, A_, B_, C_, 1.0)
    assert len(big) > 2048
    assert oracle.mindist(big) == oracle.mindist(big, n2=True)


def test_c2_hashes(oracle):
    """BASELINE config C2 (N = 9327, NS = 134162): generator and spring list hashed against the reference's."""
    import hashlib
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "c2_d00474.json")
    if not os.path.exists(path):
        pytest.skip("c2 fixture not generated")
    meta = json.load(open(path))
    oracle.srand(meta["seed"])
    p = oracle.rand_ellipsoid(meta["d"], A_, B_, C_, 1.0)
    assert len(p) == meta["N"]
    assert hashlib.sha256(p.tobytes()).hexdigest() == meta["sha256_gen_rows"]
    s = oracle.connect_springs_dist(p, 2.3 * meta["d"], 0, len(p), K, GAMMA)
    assert len(s) == meta["NS"]
    assert hashlib.sha256(s.tobytes()).hexdigest() == meta["sha256_springs"]
    assert oracle.mindist(p) == meta["mindist"]


# ---------------------------------------------------------------------------------------------------
# 3. the unmodified reference, live
# ---------------------------------------------------------------------------------------------------
def test_live_reference_layout(ref):
    """struct sizes the C-ABI relies on (128 B particle, 32 B spring, 1352 B simulation)."""
    assert ref.lib.ref_sizeof_particle() == 128
    assert ref.lib.ref_sizeof_spring() == 32
    assert ref.lib.ref_sizeof_simulation() == 1352


def test_live_reference_random_bodies(oracle, ref):
    """Fresh seeds and shapes, not in any fixture: generator, connect, and 10 steps, bit for bit."""
    for seed, d, abc in ((1, 0.2, (1.0, 1.0, 1.0)), (99, 0.13, (1.0, 0.6, 0.4))):
        ref.close()
        from oracle.bindings import Ref
        r = Ref()
        try:
            r.seed(seed)
            n = r.rand_ellipsoid(d, *abc, 2.5)
            oracle.srand(seed)
            p = oracle.rand_ellipsoid(d, *abc, 2.5)
            assert same_bits(p, r.get())
            r.connect_springs_dist(2.1 * d, 0, n, 0.3, 0.7)
            s = oracle.connect_springs_dist(p, 2.1 * d, 0, n, 0.3, 0.7)
            assert s.tobytes() == r.springs().tobytes()
            r.spin_body(0, n, (0.3, -0.2, 0.5)); oracle.spin_body(p, 0, n, (0.3, -0.2, 0.5))
            r.config(gravity_basic=1, G=0.7, softening=1e-3, dt=5e-3, zero_first=0)
            sim = make_sim(dt=5e-3, G=0.7, softening=1e-3, gravity_basic=1)
            r.step(10); oracle.step(p, s, sim, 10)
            assert same_bits(p, r.get())
            assert sim.t == r.t
        finally:
            r.close()


# ---------------------------------------------------------------------------------------------------
# src_spring/orb.cpp terms of the phobos* modules: quadrupole_accel, drift_resolved, drift_bin
# ---------------------------------------------------------------------------------------------------
def test_golden_orb_terms(oracle, orb_golden):
    g = orb_golden
    A = g["A_in"]
    for tag, (phi, theta) in (("pole", (0.0, 0.0)), ("tilt", (0.7, 1.1))):
        p = A.copy(); oracle.quadrupole_accel(p, 0.7, 5.0, 50.0, phi, theta, 3)
        assert same_bits(p, g["A_quad_" + tag]), tag
    p = A.copy(); oracle.drift_resolved(p, 0.7, 10.0, 0.2, 0.2, 3, 0, 3)
    assert same_bits(p, g["A_drift_resolved"])
    p = A.copy(); oracle.drift_bin(p, 0.7, 10.0, 0.2, 0.2, 3, 1)
    assert same_bits(p, g["A_drift_bin"])
    B = g["B_in"]; n = len(B) - 2
    p = B.copy(); oracle.quadrupole_accel(p, 1.0, 1.96e-3, 0.5, 0.3, 0.4, n)
    assert same_bits(p, g["B_quad"])
    oracle.drift_resolved(p, 1.0, 1e-3, 1e-2, 3e-2, n, 0, n)
    oracle.drift_bin(p, 1.0, 1e-3, 1e-2, 3e-2, n, n + 1)
    assert same_bits(p, g["B_drift"])
    # no-op rules of orb.cpp:315-320
    p = A.copy(); oracle.quadrupole_accel(p, 0.7, 0.0, 50.0, 0.0, 0.0, 3); assert same_bits(p, A)
    p = A.copy(); oracle.quadrupole_accel(p, 0.7, 5.0, 50.0, 0.0, 0.0, 4); assert same_bits(p, A)


def test_kat_quadrupole_closed_form(oracle, orb_golden):
    """OrbTest.QuadAccel (modules/tests/tests.cpp:1684-1733): pole along z, closed-form J2 field, 1e-5 relative --
    away from the reference's `+ 0.5` softening of r^5 (orb.cpp:340), i.e. with the perturber far away."""
    P = orb_golden["A_in"].copy(); P[:, 6:9] = 0.0
    G, J2s, Rp = 0.7, 5.0, 50.0
    want = np.zeros((4, 3))
    J2 = -J2s * G * P[3, 9] * Rp ** 2
    for i in range(3):
        x, y, z = P[3, :3] - P[i, :3]
        r7 = np.linalg.norm(P[3, :3] - P[i, :3]) ** 7
        want[i] = (J2 * x / r7 * (6 * z * z - 1.5 * (x * x + y * y)), J2 * y / r7 * (6 * z * z - 1.5 * (x * x + y * y)),
                   J2 * z / r7 * (3 * z * z - 4.5 * (x * x + y * y)))
        want[i] /= P[i, 9]
        want[3] -= want[i] * P[i, 9] / P[3, 9]
    oracle.quadrupole_accel(P, G, J2s, Rp, 0.0, 0.0, 3)
    assert np.all(np.abs(P[:, 6:9] - want) <= 1e-5 * np.abs(want))


def test_reference_update_stress_is_unusable_on_real_bodies(ref, body015):
    """Why SURVEY 8f's N4 (stress tensor) is not built: on any body that is not the unit test's integer-coordinate
    fixture the reference's own update_stress throws -- eigenvalues() demands an EXACTLY symmetric matrix
    (matrix_math.cpp:273-276) and the sum of outer(F, L) is symmetric only up to rounding -- and its `catch (char*)`
    (stress.cpp:88-93) does not match the `const char*` thrown, so a module calling it would terminate.  No shipped
    module calls it.  There is nothing to be in parity with."""
    import ctypes as C
    p = body015["start"].copy()
    ref.add_particles(p); ref.set_springs(body015["springs"])
    out = np.zeros((len(p), 4))
    assert ref.lib.ref_update_stress(ref.h, out.ctypes.data_as(C.c_void_p)) == 1
