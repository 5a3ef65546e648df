"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the reference's golden vectors.

Bars (task section 3, BASELINE.json north_star):
  * spring list: bit-exact (i, j, rs0 bits, order)
  * leapfrog drift/kick given the same accelerations: bit-exact
  * per-evaluation accelerations: max_i |a_gpu - a_ref| / max_i |a_ref| <= 1e-12   (ACC_TOL)
  * energy / spin angular momentum history over 200 steps: 1e-9 relative             (HIST_TOL)
"""
import numpy as np
import pytest

from conftest import kat_particles, same_bits
from oracle.bindings import SPRING_DTYPE, make_sim

pytestmark = pytest.mark.gpu

ACC_TOL = 1e-12
HIST_TOL = 1e-9
A_, B_, C_ = 1.0, 0.7, 0.5
K, GAMMA = 0.08, 5.0
OMEGA = (0.1, 0.0, 0.2)
EMPTY = np.zeros(0, dtype=SPRING_DTYPE)


def acc_err(got, want):
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)


def new_sim(gpu, p, springs=None, **params):
    s = gpu.Sim()
    s.upload(p)
    if springs is not None:
        s.set_springs(springs)
    if params:
        s.set_params(**params)
    return s


# ---------------------------------------------------------------------------------------------------
# the reference's 3-particle known-answer fixture
# ---------------------------------------------------------------------------------------------------
def test_kat_connect_and_force(gpu, kat3):
    P = kat_particles()
    with new_sim(gpu, P) as s:
        assert s.connect_springs_dist(5.0, 0, 3, 11.0, 35.0).tobytes() == kat3["connect5"].tobytes()
        sp = s.connect_springs_dist(100.0, 0, 3, 11.0, 35.0)
        assert sp.tobytes() == kat3["connect100"].tobytes()
        # duplicate rule (springs.cpp:70-79) and empty ranges (springs.cpp:116-117)
        again = s.connect_springs_dist(100.0, 0, 3, 2.0, 1.0, existing=kat3["connect5"])
        assert [(a, b) for a, b in zip(again["particle_1"], again["particle_2"])] == [(0, 2)]
        assert len(s.connect_springs_dist(100.0, 2, 2, 1.0, 1.0)) == 0
        assert len(s.connect_springs_dist(100.0, 3, 1, 1.0, 1.0)) == 0
        assert len(s.connect_springs_dist(-1.0, 0, 3, 1.0, 1.0)) == 0
    # SpringTest.Force: gamma = 0 then gamma = 1 after moving particle 0 to (1,1,4)
    P = kat3["force_in"].copy()
    sp = kat3["connect100"].copy()
    with new_sim(gpu, P, sp) as s:
        for gamma, key in ((0.0, "acc_g0"), (1.0, "acc_g1")):
            sp["gamma"] = gamma
            s.update_spring_props(sp)
            s.zero_accel()
            s.spring_forces()
            out = s.download(P.copy())
            assert acc_err(out[:, 6:9], kat3[key]) <= ACC_TOL
        d = s.diagnostics(0, 3, want_gpe=True)
        assert np.allclose(d, kat3["diag_g1"], rtol=1e-13, atol=1e-13)


def test_kat_physics_helpers(gpu, kat3):
    P = kat_particles(); P[:, 6:9] = P[:, :3]
    with new_sim(gpu, P, G=1.0) as s:
        d = s.diagnostics(0, 3, want_gpe=True)
        assert np.allclose(d, kat3["phys_diag"], rtol=1e-14, atol=1e-14)
        assert np.allclose(d[0:3], (4. / 3., 13. / 6., 11. / 3.), rtol=1e-15)   # PhysicsTest.CoM
        assert d[9] == 22.75                                                    # PhysicsTest.NonTransKE
        s.zero_accel()
        assert np.all(s.download(P.copy())[:, 6:9] == 0.0)                      # PhysicsTest.ZeroAccel
        s.center_sim(0, 2)
        assert same_bits(s.download(P.copy())[:, :3], kat3["center_all"])       # PhysicsTest.CenterAll (EXPECT_EQ)
    for op, key in (("subtract_com", "sub_com"), ("subtract_cov", "sub_cov")):
        P = kat_particles()
        with new_sim(gpu, P) as s:
            getattr(s, op)(0, 3)
            assert same_bits(s.download(P.copy())[:, :6], kat3[key])
    P = kat_particles()
    with new_sim(gpu, P) as s:
        s.spin_body(0, 3, (1.0, 12.0, 12.0))
        assert same_bits(s.download(P.copy())[:, :6], kat3["spin"])
        s.move_resolved((-1, -1, -1), (1, 1, 1), 0, 3)                          # PhysicsTest.MoveResolved
        out = s.download(P.copy())
        assert np.allclose(out[:, :3], kat3["spin"][:, :3] - 1, atol=1e-15)


def test_kat_gravity_and_leapfrog(gpu, kat3):
    P = kat_particles(); P[:, 6:9] = 7.0
    with new_sim(gpu, P, gravity=1, G=1.0, softening=0.01, dt=0.01) as s:
        s.gravity_basic()
        out = s.download(P.copy())
        assert acc_err(out[:, 6:9], kat3["grav_acc"]) <= ACC_TOL
        # leapfrog is bit-exact when fed the reference's own accelerations
        Q = P.copy(); Q[:, 6:9] = kat3["grav_acc"]
        s.upload(Q)
        s.leapfrog_part1()
        assert same_bits(s.download(Q.copy()), kat3["lf1"])
        s.leapfrog_part2()
        assert same_bits(s.download(Q.copy()), kat3["lf2"])
        assert s.params.t == kat3["lf_t"][0]
    # REB_GRAVITY_NONE is a no-op that leaves a untouched (gravity.c:68-70)
    with new_sim(gpu, P, gravity=0) as s:
        s.gravity_basic()
        assert np.all(s.download(P.copy())[:, 6:9] == 7.0)


# ---------------------------------------------------------------------------------------------------
# random-ellipsoid bodies: golden vectors written by the unmodified reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fixture_name", ["body015", "body010"])
def test_body_build(gpu, request, fixture_name):
    g = request.getfixturevalue(fixture_name)
    d = float(g["d"][0])
    gpu.srand(12345)
    p = gpu.rand_ellipsoid(d, A_, B_, C_, 1.0)
    assert same_bits(p, g["gen"])
    n = len(p)
    with new_sim(gpu, p) as s:
        assert s.mindist() == g["mindist"][0]
        sp = s.connect_springs_dist(2.3 * d, 0, n, K, GAMMA)
        assert sp.tobytes() == g["springs"].tobytes()           # bit-exact list: (i, j) order and rs0 bits
        # set-up helpers sum in the reference's particle order by default: bit-identical start state
        s.subtract_com(0, n); s.subtract_cov(0, n); s.spin_body(0, n, OMEGA)
        out = s.download(p.copy())
        assert same_bits(out[:, :6], g["start"][:, :6])
        # the parallel tree (per-step heartbeats at large N) agrees to rounding
        s.upload(p)
        s.set_sum_order(sequential=False)
        s.subtract_com(0, n); s.subtract_cov(0, n); s.spin_body(0, n, OMEGA)
        out = s.download(p.copy())
        assert np.allclose(out[:, :6], g["start"][:, :6], rtol=0, atol=1e-15)


@pytest.mark.parametrize("fixture_name", ["body015", "body010"])
def test_body_single_evaluations(gpu, oracle, request, fixture_name):
    """Each reference function on its own, on the reference's own state."""
    g = request.getfixturevalue(fixture_name)
    d = float(g["d"][0])
    p0 = g["step20_basic"].copy()       # a state with non-trivial strains and velocities
    sp = g["springs"]
    with new_sim(gpu, p0, sp, gravity=1, G=1.0, softening=d / 100.0, dt=2e-3) as s:
        # spring_forces: a += f
        want = p0.copy(); oracle.spring_forces(want, sp)
        s.spring_forces()
        got = s.download(p0.copy())
        assert acc_err(got[:, 6:9] - p0[:, 6:9], want[:, 6:9] - p0[:, 6:9]) <= ACC_TOL
        # gravity: a = g
        want = p0.copy(); oracle.gravity_basic(want, G=1.0, softening=d / 100.0)
        s.gravity_basic()
        got = s.download(p0.copy())
        assert acc_err(got[:, 6:9], want[:, 6:9]) <= ACC_TOL
        # softening = 0 exercises the explicit i == j skip
        s.set_params(softening=0.0)
        want = p0.copy(); oracle.gravity_basic(want, G=1.0, softening=0.0)
        s.gravity_basic()
        got = s.download(p0.copy())
        assert np.all(np.isfinite(got[:, 6:9]))
        assert acc_err(got[:, 6:9], want[:, 6:9]) <= ACC_TOL
        # N_active < N: only the first particles are sources (gravity.c:64,89)
        s.set_params(softening=d / 100.0, n_active=len(p0) // 3)
        want = p0.copy(); oracle.gravity_basic(want, G=1.0, softening=d / 100.0, n_active=len(p0) // 3)
        s.gravity_basic()
        got = s.download(p0.copy())
        assert acc_err(got[:, 6:9], want[:, 6:9]) <= ACC_TOL


@pytest.mark.parametrize("fixture_name", ["body015", "body010"])
def test_body_steps(gpu, request, fixture_name):
    """reb_step x 1 and x 20 against the reference, both shipped callback shapes."""
    g = request.getfixturevalue(fixture_name)
    d = float(g["d"][0])
    sp = g["springs"]
    for tag, grav, af, dt in (("none", 0, 2, 1e-2), ("basic", 1, 1, 2e-3)):
        start = g["start"].copy()
        with new_sim(gpu, start, sp, gravity=grav, G=1.0, softening=d / 100.0, dt=dt, additional_forces=af) as s:
            s.step(1)
            out = s.download(start.copy())
            want = g["step1_" + tag]
            assert acc_err(out[:, 6:9], want[:, 6:9]) <= ACC_TOL
            assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-14)
            s.step(19)
            out = s.download(start.copy())
            want = g["step20_" + tag]
            assert acc_err(out[:, 6:9], want[:, 6:9]) <= 1e-10      # 20 steps of accumulated rounding
            assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-12)
            assert abs(s.params.t - 20 * dt) < 1e-12
        # the unfused sequence of the same five calls gives the same state as the fused loop, bit for bit
        with new_sim(gpu, start, sp, gravity=grav, G=1.0, softening=d / 100.0, dt=dt, additional_forces=af) as a, \
                new_sim(gpu, start, sp, gravity=grav, G=1.0, softening=d / 100.0, dt=dt, additional_forces=af) as b:
            a.step(3)
            for _ in range(3):
                b.leapfrog_part1()
                b.gravity_basic()
                if af == 2:
                    b.zero_accel()
                b.spring_forces()
                b.leapfrog_part2()
            assert same_bits(a.download(start.copy()), b.download(start.copy()))
            assert a.params.t == b.params.t


def test_history_energy_and_angular_momentum(gpu, body015):
    """E = KErot + PEspr (+ PEgrav) and L over 200 steps within 1e-9 relative of the reference's history."""
    g = body015
    d = float(g["d"][0])
    sp = g["springs"]
    n = len(g["start"])
    every = int(g["hist_every"][0])
    for tag, grav, af, dt in (("none", 0, 2, 1e-2), ("basic", 1, 1, 2e-3)):
        ref_hist = g["hist_" + tag]
        with new_sim(gpu, g["start"].copy(), sp, gravity=grav, G=1.0, softening=d / 100.0, dt=dt, additional_forces=af) as s:
            for row in ref_hist:
                got = s.diagnostics(0, n, want_gpe=bool(grav))
                e_ref = row[9] + row[10] + row[11]
                e_got = got[9] + got[10] + got[11]
                assert abs(e_got - e_ref) <= HIST_TOL * abs(e_ref)
                assert np.abs(got[6:9] - row[6:9]).max() <= HIST_TOL * np.abs(row[6:9]).max()
                # damping power; at t = 0 the body rotates rigidly and the strain rates are pure rounding noise
                assert abs(got[12] - row[12]) <= 1e-8 * abs(row[12]) + 1e-20
                s.step(every)


def test_integrate_exact_finish_and_heartbeat(gpu, body015):
    """reb_integrate: center_sim heartbeat every step (and at t=0), last step shortened to land on tmax."""
    g = body015
    d = float(g["d"][0])
    start = g["start"].copy()
    with new_sim(gpu, start, g["springs"], gravity=0, G=1.0, softening=d / 100.0, dt=1e-2, additional_forces=2, heartbeat=1) as s:
        status = s.integrate(0.1234)
        out = s.download(start.copy())
        prm = s.params
        assert (status, prm.t, prm.dt) == tuple(g["integrate_meta"])
        assert prm.t == 0.1234
        assert np.allclose(out[:, :6], g["integrate_state"][:, :6], rtol=0, atol=1e-13)
        assert acc_err(out[:, 6:9], g["integrate_state"][:, 6:9]) <= 1e-10


# ---------------------------------------------------------------------------------------------------
# callback shapes, edge cases, error behaviour
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("uniform_mass", [False, True])
@pytest.mark.parametrize("softening", [1e-3, 0.0])
def test_gravity_kernels_agree(gpu, oracle, body010, softening, uniform_mass):
    """The one-sided tile kernel and the pair-once kernel against the oracle, on sizes that exercise padding (N not a
    multiple of the 1024-particle block), the diagonal tile's self-pair mask (softening = 0) and several blocks.
    uniform_mass: every particle carries the same mass (the pair-once kernel's 19-instruction specialisation, mass
    applied by the reduction) / a heavy point mass at the end (the general 21-instruction kernel)."""
    g = body010
    base = g["step20_basic"].copy()
    # 1065 particles -> 2 blocks; stack shifted copies for 4 blocks + a heavy point mass at the end
    p = np.vstack([base, base + np.r_[3.0, 0, 0, np.zeros(8)], base + np.r_[0, 3.0, 0, np.zeros(8)], base[:777] + np.r_[0, 0, 3.0, np.zeros(8)]])
    pm = np.zeros((1, 11)); pm[0, :3] = (9.0, 1.0, -2.0); pm[0, 9] = 25.0
    p = np.ascontiguousarray(p if uniform_mass else np.vstack([p, pm]))
    assert (len(set(p[:, 9])) == 1) == uniform_mass
    want = p.copy(); oracle.gravity_basic(want, G=0.7, softening=softening)
    for which in (1, 2):
        with new_sim(gpu, p, gravity=1, G=0.7, softening=softening) as s:
            s.set_gravity_kernel(which)
            s.gravity_basic()
            got = s.download(p.copy())
            assert np.all(np.isfinite(got[:, 6:9]))
            assert acc_err(got[:, 6:9], want[:, 6:9]) <= ACC_TOL, which
            s.gravity_basic()                                     # deterministic: same bits on a second evaluation
            assert same_bits(s.download(p.copy())[:, 6:9], got[:, 6:9])
            if which == 2 and uniform_mass:
                # a mass edit must be noticed: the plan re-checks the masses after an upload that carries them
                q = p.copy(); q[5, 9] *= 3.0
                want2 = q.copy(); oracle.gravity_basic(want2, G=0.7, softening=softening)
                s.upload(q)
                s.gravity_basic()
                assert acc_err(s.download(q.copy())[:, 6:9], want2[:, 6:9]) <= ACC_TOL


def test_accelerations_accumulate_without_zero_accel(gpu, oracle, body015):
    """with REB_GRAVITY_NONE and no zero_accel (bennu2, wobble_*) `a` accumulates across steps."""
    g = body015
    start = g["start"].copy(); start[:, 6:9] = 0.125
    sp = g["springs"]
    want = start.copy()
    sim = make_sim(dt=1e-3, gravity_basic=0, zero_first=0)
    oracle.step(want, sp, sim, 5)
    with new_sim(gpu, start, sp, gravity=0, dt=1e-3, additional_forces=1) as s:
        s.step(5)
        out = s.download(start.copy())
    assert acc_err(out[:, 6:9], want[:, 6:9]) <= 1e-11
    assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-13)


def test_no_springs_no_forces(gpu, oracle, body015):
    start = body015["start"].copy(); start[:, 6:9] = 0.5
    for grav, af in ((0, 0), (0, 1), (0, 2), (1, 0), (1, 2)):
        want = start.copy()
        sim = make_sim(dt=1e-3, softening=1e-3, gravity_basic=grav, zero_first=(af == 2))
        oracle.step(want, EMPTY, sim, 2)
        with new_sim(gpu, start, None, gravity=grav, softening=1e-3, dt=1e-3, additional_forces=af) as s:
            s.spring_forces()                    # no springs: a no-op
            s.step(2)
            out = s.download(start.copy())
        assert acc_err(out[:, 6:9], want[:, 6:9]) <= ACC_TOL or np.abs(want[:, 6:9]).max() == 0.0
        assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-14)


def test_point_mass_perturber(gpu, oracle, body015):
    """phobos_deimos shape: resolved body + one heavy point mass at the end of the array, num_perts = 1; the heartbeat
    recentres on the resolved body only but shifts everything (physics.cpp:98-109)."""
    g = body015
    body = g["start"].copy()
    n = len(body)
    pert = np.zeros((1, 11)); pert[0, :3] = (7.0, 0.0, 0.0); pert[0, 4] = 0.4; pert[0, 9] = 1.0
    p = np.vstack([body, pert])
    sp = g["springs"]
    want = p.copy()
    sim = make_sim(dt=2e-3, softening=1e-3, gravity_basic=1, zero_first=0, center_heartbeat=1, n_perts=1)
    oracle.step(want, sp, sim, 4, heartbeat=True)
    with new_sim(gpu, p, sp, gravity=1, softening=1e-3, dt=2e-3, additional_forces=1, heartbeat=1, n_perts=1) as s:
        s.step(4, heartbeat=True)
        out = s.download(p.copy())
    assert acc_err(out[:, 6:9], want[:, 6:9]) <= 1e-11
    assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-13)


def test_set_gamma_mid_run(gpu, oracle, body015):
    """set_gamma at t ~ t_damp (every module's heartbeat): host list changes, device copy is refreshed."""
    g = body015
    start = g["start"].copy()
    sp = g["springs"].copy()
    want = start.copy()
    sim = make_sim(dt=1e-2, gravity_basic=0, zero_first=1)
    oracle.step(want, sp, sim, 3)
    sp2 = sp.copy(); sp2["gamma"] = 1.0
    oracle.step(want, sp2, sim, 3)
    with new_sim(gpu, start, sp, gravity=0, dt=1e-2, additional_forces=2) as s:
        s.step(3)
        s.update_spring_props(sp2)
        s.step(3)
        out = s.download(start.copy())
    assert acc_err(out[:, 6:9], want[:, 6:9]) <= 1e-11
    assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-13)


def test_errors(gpu):
    P = kat_particles()
    s = gpu.Sim()
    with pytest.raises(gpu.AssError):            # nothing uploaded yet
        s.spring_forces()
    s.upload(P)
    bad = np.zeros(1, dtype=SPRING_DTYPE); bad["particle_1"] = 1; bad["particle_2"] = 1
    with pytest.raises(gpu.AssError) as e:       # "Cannot connect a particle to itself." (springs.cpp:57-59)
        s.set_springs(bad)
    assert e.value.code == -1 and "itself" in str(e.value)
    bad["particle_2"] = 9
    with pytest.raises(gpu.AssError):
        s.set_springs(bad)
    with pytest.raises(gpu.AssError):
        s.download(np.zeros((5, 11)))
    with pytest.raises(gpu.AssError):
        s.set_params(gravity=7)
    with pytest.raises(gpu.AssError):
        s.update_spring_props(np.zeros(4, dtype=SPRING_DTYPE))
    s.close()


def test_partial_upload_download_and_strides(gpu):
    """Rows may be `struct reb_particle` (128-byte stride): only the requested doubles move."""
    n = 100
    rng = np.random.default_rng(5)
    wide = rng.normal(size=(n, 16))                  # 128-byte rows; columns 11.. stand for pointers / hash
    p = wide[:, :]
    s = gpu.Sim()
    s.upload(p)
    back = np.full((n, 16), -7.0)
    s.download(back, fields=gpu.F_POS | gpu.F_ACC)
    assert np.array_equal(back[:, 0:3], wide[:, 0:3]) and np.array_equal(back[:, 6:9], wide[:, 6:9])
    assert np.all(back[:, 3:6] == -7.0) and np.all(back[:, 9:] == -7.0)
    newv = rng.normal(size=(n, 16))
    s.upload(newv, fields=gpu.F_VEL)
    full = np.zeros((n, 16))
    s.download(full)
    assert np.array_equal(full[:, 0:3], wide[:, 0:3]) and np.array_equal(full[:, 3:6], newv[:, 3:6])
    assert np.array_equal(full[:, 9], wide[:, 9]) and np.all(full[:, 10:] == 0.0)
    s.close()


def test_deterministic(gpu, body010):
    g = body010
    d = float(g["d"][0])
    outs = []
    for _ in range(2):
        with new_sim(gpu, g["start"].copy(), g["springs"], gravity=1, softening=d / 100.0, dt=2e-3, additional_forces=1) as s:
            s.step(10)
            outs.append(s.download(g["start"].copy()))
    assert same_bits(outs[0], outs[1])


# ---------------------------------------------------------------------------------------------------
# full-size checks (BASELINE C2 / C3 sizes): oracle on a slab + size-independent properties
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,expect_n", [(0.0474, 9327), (0.0215, None)])
def test_full_size(gpu, oracle, d, expect_n):
    gpu.srand(12345)
    p = gpu.rand_ellipsoid(d, A_, B_, C_, 1.0)
    n = len(p)
    if expect_n:
        assert n == expect_n
    else:
        assert 90_000 < n < 110_000
    soft = d / 100.0
    with new_sim(gpu, p, gravity=1, G=1.0, softening=soft, dt=1e-3, additional_forces=1) as s:
        # spring list: bit-exact against the oracle's cell-list restatement (itself pinned to the reference at C2)
        sp = s.connect_springs_dist(2.3 * d, 0, n, K, GAMMA)
        want_sp = oracle.connect_springs_dist(p, 2.3 * d, 0, n, K, GAMMA)
        assert sp.tobytes() == want_sp.tobytes()
        assert np.all(sp["particle_1"] < sp["particle_2"])
        key = sp["particle_1"].astype(np.int64) * n + sp["particle_2"]
        assert np.all(np.diff(key) > 0)                              # strictly lexicographic, no duplicates
        assert s.mindist() == oracle.mindist(p)
        s.set_springs(sp)
        s.subtract_com(0, n); s.subtract_cov(0, n); s.spin_body(0, n, OMEGA)
        s.step(2)                                                    # get off the rest state
        state = s.download(p.copy())
        # gravity on a slab of targets against the oracle (full sum over all N sources)
        lo, hi = n // 2 - 128, n // 2 + 128
        want = state.copy(); oracle.gravity_basic(want, G=1.0, softening=soft, i_lo=lo, i_hi=hi)
        s.gravity_basic()
        got = s.download(state.copy())
        assert acc_err(got[lo:hi, 6:9], want[lo:hi, 6:9]) <= ACC_TOL
        # Newton's third law: sum m a = 0 for gravity and for springs (size-independent)
        scale = np.abs(got[:, 9:10] * got[:, 6:9]).sum(axis=0)
        assert np.all(np.abs((got[:, 9:10] * got[:, 6:9]).sum(axis=0)) <= 1e-12 * scale)
        # springs: full comparison (the oracle is O(NS))
        want = state.copy(); want[:, 6:9] = 0.0; oracle.spring_forces(want, sp)
        s.zero_accel(); s.spring_forces()
        got = s.download(state.copy())
        assert acc_err(got[:, 6:9], want[:, 6:9]) <= ACC_TOL
        net = (got[:, 9:10] * got[:, 6:9]).sum(axis=0)
        assert np.all(np.abs(net) <= 1e-12 * np.abs(got[:, 9:10] * got[:, 6:9]).sum(axis=0))
        # angular momentum is conserved by the step (internal forces are central): 20 steps, 1e-11 relative
        L0 = s.diagnostics(0, n)[6:9]
        s.step(20)
        L1 = s.diagnostics(0, n)[6:9]
        assert np.abs(L1 - L0).max() <= 1e-11 * np.abs(L0).max()


# ---------------------------------------------------------------------------------------------------
# pinned host rows: the link-rate transfer path moves the same bits and touches only the requested doubles
# ---------------------------------------------------------------------------------------------------
def test_pinned_rows_transfer_matches_pageable(gpu):
    rng = np.random.default_rng(7)
    n = 5003
    P = rng.standard_normal((n, 16))                   # 128-byte rows, like struct reb_particle
    pin = gpu.PinnedRows(n, 16)
    try:
        pin.a[:] = P
        with gpu.Sim() as s:
            s.upload(pin.a)                            # copy engine reads the pinned rows directly
            out_pageable = np.full((n, 16), -7.0)
            s.download(out_pageable)
            pin.a[:] = -7.0
            s.download(pin.a)                          # pack kernel writes straight into the pinned rows
            assert same_bits(pin.a, out_pageable)
            assert same_bits(pin.a[:, :10], P[:, :10]) and np.all(pin.a[:, 10:] == -7.0)
            for fields, cols in ((gpu.F_POS, [0, 1, 2]), (gpu.F_VEL, [3, 4, 5]), (gpu.F_ACC, [6, 7, 8]), (gpu.F_MASS, [9]),
                                 (gpu.F_POS | gpu.F_ACC, [0, 1, 2, 6, 7, 8])):
                pin.a[:] = -7.0
                s.download(pin.a, fields)
                want = np.full((n, 16), -7.0)
                want[:, cols] = P[:, cols]
                assert same_bits(pin.a, want), fields
            # a sub-range of the pinned allocation (rows 10..) is still recognised as pinned
            with gpu.Sim() as t:
                pin.a[:] = P
                t.upload(pin.a[10:])
                got = np.zeros((n - 10, 16))
                t.download(got)
                assert same_bits(got[:, :10], P[10:, :10])
        # registering an existing array in place
        Q = P.copy()
        gpu.host_register(Q)
        try:
            with gpu.Sim() as s:
                s.upload(Q)
                Q[:, :10] = 0.0
                s.download(Q)
                assert same_bits(Q, P)
        finally:
            gpu.host_unregister(Q)
    finally:
        pin.close()


def test_branch_free_sqrt_matches_dsqrt_rn(gpu):
    """len = sqrt(...) must be the reference's correctly rounded value bit for bit (at rest one ulp of len is 1e-10 of the
    force).  The spring kernels use the fast path of __dsqrt_rn without its range check (csrc/spring_math.cuh)."""
    bad, first = gpu.selfcheck_sqrt(1 << 28, -200, 200)
    assert bad == 0, (bad, first)
    bad, first = gpu.selfcheck_sqrt(1 << 26, -40, 8)      # the range squared spring lengths live in
    assert bad == 0, (bad, first)


# ---------------------------------------------------------------------------------------------------
# src_spring/orb.cpp terms of the phobos* modules
# ---------------------------------------------------------------------------------------------------
def test_orb_terms_match_reference(gpu, oracle, orb_golden):
    g = orb_golden
    A = g["A_in"]
    with new_sim(gpu, A, G=0.7, n_perts=1) as s:
        for tag, (phi, theta) in (("pole", (0.0, 0.0)), ("tilt", (0.7, 1.1))):
            s.upload(A)
            s.quadrupole_accel(5.0, 50.0, phi, theta, 3)
            out = s.download(A.copy())
            assert acc_err(out[:, 6:9], g["A_quad_" + tag][:, 6:9]) <= ACC_TOL, tag
            assert same_bits(out[:, :6], A[:, :6])
        s.upload(A); s.quadrupole_accel(0.0, 50.0, 0.0, 0.0, 3); s.quadrupole_accel(5.0, 50.0, 0.0, 0.0, 4)   # orb.cpp:315-320
        assert same_bits(s.download(A.copy()), A)
        s.upload(A); s.drift_resolved(10.0, 0.2, 0.2, 3, 0, 3)
        out = s.download(A.copy())
        assert np.allclose(out[:, 3:6], g["A_drift_resolved"][:, 3:6], rtol=1e-13, atol=1e-15)
        assert same_bits(out[:, :3], A[:, :3])
        s.upload(A); s.drift_bin(10.0, 0.2, 0.2, 3, 1)
        out = s.download(A.copy())
        assert np.allclose(out[:, 3:6], g["A_drift_bin"][:, 3:6], rtol=1e-13, atol=1e-15)
        assert s.sum_mass(0, 3) == 6.0 and s.sum_mass(1, 1) == 0.0                                        # OrbTest.TotMass
    B = g["B_in"]; n = len(B) - 2
    for sequential in (True, False):
        with new_sim(gpu, B, G=1.0, n_perts=2) as s:
            s.set_sum_order(sequential)
            s.quadrupole_accel(1.96e-3, 0.5, 0.3, 0.4, n)
            out = s.download(B.copy())
            assert acc_err(out[:, 6:9], g["B_quad"][:, 6:9]) <= ACC_TOL
            s.drift_resolved(1e-3, 1e-2, 3e-2, n, 0, n)
            s.drift_bin(1e-3, 1e-2, 3e-2, n, n + 1)
            out = s.download(B.copy())
            assert np.allclose(out[:, 3:6], g["B_drift"][:, 3:6], rtol=1e-13, atol=1e-16)
            # a row range comes back alone (the point masses, for write_pt_mass while the body stays resident)
            part = np.full_like(B, -1.0)
            s.download_range(part, n, n + 2)
            assert same_bits(part[n:, :10], out[n:, :10]) and np.all(part[:n] == -1.0)
    with new_sim(gpu, A) as s:
        with pytest.raises(gpu.AssError):
            s.drift_resolved(1.0, 0.1, 0.1, 1, 0, 3)        # the point mass lies inside the body range
        with pytest.raises(gpu.AssError):
            s.drift_bin(1.0, 0.1, 0.1, 2, 2)


def _spring_kernel_mode(monkeypatch, mode):
    """tiles (default): lane-transposed records, endpoint state in shared-memory tables; csr: one group of lanes per node
    reading the CSR copy of the list through L1 (the kernel node slabs of a sharded body use)."""
    monkeypatch.delenv("ASS_SPRINGS_NO_STREAM", raising=False)
    if mode == "csr":
        monkeypatch.setenv("ASS_SPRINGS_NO_STREAM", "1")


def test_spring_kernels_agree_bit_for_bit(gpu, body015, monkeypatch):
    """The two spring kernels (springs.cu: tiled, CSR) run the same arithmetic in the same order: identical bits,
    with and without the fused kick, for uniform and non-uniform masses, for one and for several (k, gamma) pairs."""
    g = body015
    start = g["start"].copy(); start[:, 6:9] = 0.25
    sp = g["springs"]
    outs = []
    for mode in ("tiles", "csr"):
        _spring_kernel_mode(monkeypatch, mode)
        with new_sim(gpu, start, sp, gravity=0, dt=1e-3, additional_forces=1) as s:
            s.spring_forces()
            a1 = s.download(start.copy())
            s.step(3)
            outs.append((a1, s.download(start.copy())))
    for o in outs[1:]:
        assert same_bits(outs[0][0], o[0]) and same_bits(outs[0][1], o[1])
    # a mass edit on one sprung node switches the kernels to the general reduced-mass form; two (k, gamma) pairs to the
    # general palette path
    q = start.copy(); q[7, 9] *= 1.5
    sp2 = sp.copy(); sp2["k"][::3] *= 2.0; sp2["gamma"][::5] = 0.0
    for body, springs in ((q, sp), (start, sp2), (q, sp2)):
        res = []
        for mode in ("tiles", "csr"):
            _spring_kernel_mode(monkeypatch, mode)
            with new_sim(gpu, body, springs, gravity=0, dt=1e-3, additional_forces=2) as s:
                s.step(2)
                res.append(s.download(body.copy()))
        assert same_bits(res[0], res[1])


def test_tiled_spring_kernel_with_many_tiles_per_cta(gpu, body015, monkeypatch):
    """Large bodies need several tiles per CTA (a table holds ~2900 nodes); the library's test hooks shrink the grid and
    the table so that a small body takes that path: same bits as the CSR kernel, with the fused kick and across steps."""
    g = body015
    start = g["start"].copy(); start[:, 6:9] = 0.125
    sp = g["springs"]
    outs = []
    for grid, slots in ((None, None), ("2", "400"), ("3", "1000"), ("1", "64")):
        for k, v in (("ASS_TILE_GRID", grid), ("ASS_TILE_SLOTS", slots)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        with new_sim(gpu, start, sp, gravity=0, dt=1e-3, additional_forces=1) as s:
            s.spring_forces()
            a1 = s.download(start.copy())
            s.step(3)
            outs.append((a1, s.download(start.copy())))
    monkeypatch.delenv("ASS_TILE_GRID", raising=False); monkeypatch.delenv("ASS_TILE_SLOTS", raising=False)
    monkeypatch.setenv("ASS_SPRINGS_NO_STREAM", "1")
    with new_sim(gpu, start, sp, gravity=0, dt=1e-3, additional_forces=1) as s:
        s.spring_forces()
        a1 = s.download(start.copy())
        s.step(3)
        ref = (a1, s.download(start.copy()))
    for o in outs:
        assert same_bits(ref[0], o[0]) and same_bits(ref[1], o[1])


@pytest.mark.gpu
@pytest.mark.parametrize("af", [1, 2])
def test_batch_of_gravity_free_steps_in_one_launch(gpu, body015, monkeypatch, af):
    """Without gravity and heartbeat ass_step(n) runs the n steps as ONE cooperative launch (springs.cu: k_spring_steps, the
    CTAs meeting at a grid barrier after every step): the same bits as one launch per step, for a body that takes one tile per
    CTA and with the test hooks that give a CTA several, with accelerations accumulating (af = 1: no zero_accel)
    and not; the clock advances as the reference's does; a following call continues from the batch's end state."""
    g = body015
    start = g["start"].copy(); start[:, 6:9] = 0.2
    sp = g["springs"]
    runs = {}
    for name, env in (("per_step", {"ASS_SPRINGS_NO_MULTISTEP": "1"}), ("batch", {}), ("batch_tiles", {"ASS_TILE_GRID": "3", "ASS_TILE_SLOTS": "1000"})):
        for k in ("ASS_SPRINGS_NO_MULTISTEP", "ASS_TILE_GRID", "ASS_TILE_SLOTS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with new_sim(gpu, start, sp, gravity=0, dt=1e-3, additional_forces=af) as s:
            n0 = gpu.launch_count()
            s.step(7)
            launches = gpu.launch_count() - n0
            mid = s.download(start.copy())
            t_mid = s.params.t
            s.step(1); s.step(4)
            runs[name] = (mid, t_mid, s.download(start.copy()), s.params.t, launches)
    ref = runs["per_step"]
    for name in ("batch", "batch_tiles"):
        r = runs[name]
        assert same_bits(ref[0], r[0]) and same_bits(ref[2], r[2]), name
        assert r[1] == ref[1] and r[3] == ref[3]
        assert r[4] < ref[4]                       # part1 + one launch instead of part1 + seven
    assert np.isfinite(ref[2][:, :9]).all()


def test_hub_body_is_served_by_the_csr_kernel(gpu, oracle):
    """A node with more neighbours than a shared-memory table holds (~2 900) cannot be tiled: ass_set_springs then builds no
    tile cut and whole-body evaluations run on the CSR kernel -- including property updates -- with the usual accuracy."""
    rng = np.random.default_rng(11)
    n = 3200
    p = np.zeros((n, 11)); p[:, :3] = rng.uniform(-1.0, 1.0, (n, 3)); p[:, 3:6] = rng.normal(0.0, 0.01, (n, 3)); p[:, 9] = 1.0 / n
    pairs = [(0, j) for j in range(1, n)] + [(j, j + 1) for j in range(1, n - 1)]
    sp = np.zeros(len(pairs), dtype=gpu.SPRING_DTYPE)
    sp["particle_1"] = [a for a, _ in pairs]; sp["particle_2"] = [b for _, b in pairs]
    sp["k"] = 0.05; sp["gamma"] = 2.0
    sp["rs0"] = np.linalg.norm(p[sp["particle_1"], :3] - p[sp["particle_2"], :3], axis=1) * 0.98
    want = p.copy(); oracle.spring_forces(want, sp)
    sp2 = sp.copy(); sp2["k"] *= 1.5
    want2 = p.copy(); oracle.spring_forces(want2, sp2)
    with new_sim(gpu, p, sp) as s:
        s.spring_forces()
        out = s.download(p.copy())
        s.update_spring_props(sp2)
        s.zero_accel(); s.spring_forces()
        out2 = s.download(p.copy())
    assert acc_err(out[:, 6:9], want[:, 6:9]) <= ACC_TOL
    assert acc_err(out2[:, 6:9], want2[:, 6:9]) <= ACC_TOL


def test_spring_props_update_reaches_every_copy_of_the_list(gpu, oracle, body015, monkeypatch):
    """set_gamma / adjust_spring_props style edits go through ass_update_spring_props, which patches the CSR and the tile
    copies of the half-edge list in place: both kernels must see the new values."""
    g = body015
    p = g["start"].copy(); p[:, 6:9] = 0.0
    sp = g["springs"]
    sp2 = sp.copy(); sp2["gamma"] *= 0.25; sp2["k"][1::2] *= 1.5; sp2["rs0"] *= 1.001
    want = p.copy(); oracle.spring_forces(want, sp2)
    res = []
    for mode in ("tiles", "csr"):
        _spring_kernel_mode(monkeypatch, mode)
        with new_sim(gpu, p, sp) as s:
            s.spring_forces()
            s.update_spring_props(sp2)
            s.zero_accel()
            s.spring_forces()
            res.append(s.download(p.copy()))
    assert same_bits(res[0], res[1])
    assert acc_err(res[0][:, 6:9], want[:, 6:9]) <= ACC_TOL


def test_spring_forces_nonuniform_mass_matches_oracle(gpu, oracle, body015):
    g = body015
    rng = np.random.default_rng(3)
    p = g["start"].copy(); p[:, 9] *= rng.uniform(0.5, 2.0, len(p)); p[:, 6:9] = 0.0
    sp = g["springs"]
    want = p.copy(); oracle.spring_forces(want, sp)
    with new_sim(gpu, p, sp) as s:
        s.spring_forces()
        out = s.download(p.copy())
    assert acc_err(out[:, 6:9], want[:, 6:9]) <= ACC_TOL


# ---------------------------------------------------------------------------------------------------
# degenerate inputs the reference is well defined on (round-1 advisor findings)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["tiles", "csr"])
def test_zero_length_spring_gives_zero_force(gpu, oracle, body015, monkeypatch, mode):
    """Duplicate points get a spring with rs0 = 0 from connect_springs_dist; the reference gives it len = L_EPS,
    len_hat = 0 and a zero force (springs.cpp:302-312).  The branch-free sqrt must not turn that into NaN."""
    _spring_kernel_mode(monkeypatch, mode)
    g = body015
    p = g["start"].copy(); p[:, 6:9] = 0.0
    p[7, :3] = p[3, :3]                      # node 7 sits exactly on node 3 ...
    p[7, 3:6] = p[3, 3:6] + 0.01            # ... and moves relative to it
    with new_sim(gpu, p) as s:
        sp = s.connect_springs_dist(2.3 * 0.15, 0, len(p), K, GAMMA)
        want_sp = oracle.connect_springs_dist(p, 2.3 * 0.15, 0, len(p), K, GAMMA)
        assert sp.tobytes() == want_sp.tobytes()
        zero = sp[(sp["particle_1"] == 3) & (sp["particle_2"] == 7)]
        assert len(zero) == 1 and zero["rs0"][0] == 0.0
        s.set_springs(sp)
        s.spring_forces()
        out = s.download(p.copy())
        want = p.copy(); oracle.spring_forces(want, sp)
        assert np.all(np.isfinite(out[:, 6:9]))
        assert acc_err(out[:, 6:9], want[:, 6:9]) <= ACC_TOL
        s.set_params(dt=1e-3, gravity=0, additional_forces=2)
        s.step(3)
        assert np.all(np.isfinite(s.download(p.copy())[:, :9]))


def test_connect_duplicate_rule_matches_entries_as_stored(gpu, oracle):
    """add_spring only recognises an existing spring stored as (low, high) (springs.cpp:72-79): an entry written the
    other way round does not suppress the pair."""
    P = kat_particles()
    ex = np.zeros(2, dtype=SPRING_DTYPE)
    ex["particle_1"] = (1, 0); ex["particle_2"] = (0, 2)          # (1,0) is reversed, (0,2) is canonical
    ex["k"] = 1.0
    want = oracle.connect_springs_dist(P, 100.0, 0, 3, 2.0, 1.0, existing=ex, n2=True)
    with new_sim(gpu, P) as s:
        got = s.connect_springs_dist(100.0, 0, 3, 2.0, 1.0, existing=ex)
    assert got.tobytes() == want[len(ex):].tobytes()
    assert [(a, b) for a, b in zip(got["particle_1"], got["particle_2"])] == [(0, 1), (1, 2)]


def test_connect_and_mindist_survive_non_finite_positions(gpu, oracle, body015):
    """A blown-up particle (inf / NaN coordinate) pairs with nobody in the reference; the grid must neither hang nor
    lose the finite pairs."""
    g = body015
    p = g["start"].copy()
    p[5, 0] = np.inf
    p[11, 1] = np.nan
    p[17, 2] = -np.inf
    with new_sim(gpu, p) as s:
        got = s.connect_springs_dist(2.3 * 0.15, 0, len(p), K, GAMMA)
        md = s.mindist()
    want = oracle.connect_springs_dist(p, 2.3 * 0.15, 0, len(p), K, GAMMA, n2=True)
    assert got.tobytes() == want.tobytes()
    touched = set(got["particle_1"]) | set(got["particle_2"])
    assert not ({5, 11, 17} & touched)
    q = np.delete(p, [5, 11, 17], axis=0)
    assert md == oracle.mindist(q, n2=True)
    with new_sim(gpu, np.full((4, 11), np.nan)) as s:      # nothing finite at all
        assert len(s.connect_springs_dist(1.0, 0, 4, K, GAMMA)) == 0


def test_step_timed_is_the_same_loop(gpu, body015):
    """ass_step_timed (bench.py's timer: L2 overwritten before every step, a device-timed lap around every step's kernels)
    runs the very loop ass_step runs: same bits afterwards, with and without gravity; it returns the summed laps."""
    g = body015
    start = g["start"].copy(); start[:, 6:9] = 0.0
    for grav, af in ((1, 1), (0, 2)):
        outs = []
        for timed in (False, True):
            with new_sim(gpu, start, g["springs"], gravity=grav, dt=1e-3, additional_forces=af, softening=1e-3) as s:
                ms = s.step_timed(5) if timed else s.step(5)
                if timed:
                    assert 0.0 < ms < 1e3
                outs.append((s.download(start.copy()), s.params.t))
        assert same_bits(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


def test_gravity_pieces_on_concurrent_streams_give_the_same_bits(gpu, body010):
    """What a rank of a sharded body queues per step -- its own triangle and rectangles against other slabs -- as pair kernels on
    streams of their own with the reductions afterwards (ass_probe_gravity_pieces, concurrent = 1) leaves exactly the
    accelerations the same pieces leave one after the other: the kernels write disjoint scratch, the reductions keep their order."""
    p = np.vstack([body010["start"]] * 9)            # 9 585 particles: enough blocks for the pair-once tiles
    p[:, :3] += np.repeat(np.arange(9), len(body010["start"]))[:, None] * 3.0
    n = len(p)
    a, b, c = n // 4, n // 2, (3 * n) // 4
    pieces = [(0, a, 0, a), (0, a, a, c), (0, a // 2, c, n)]
    outs = []
    for conc in (0, 1):
        with new_sim(gpu, p, None, gravity=1, dt=1e-3, G=1.0, softening=1e-3) as s:
            ms = s.probe_gravity_pieces(pieces, conc, reps=2)
            assert ms > 0.0
            outs.append(s.download(p.copy())[:a, 6:9].copy())
    assert same_bits(outs[0], outs[1])
    assert np.isfinite(outs[0]).all() and np.abs(outs[0]).max() > 0.0
