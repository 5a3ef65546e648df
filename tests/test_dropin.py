"""The drop-in boundary, exercised the way a maintainer would use it: ONE problem-module source (tests/dropin/problem.cpp,
reference API only) linked (a) against the unmodified reference and (b) against the reference with its hot-path symbols
replaced by asteroid-spring-sims_b200/host/dropin.cpp.  Both binaries read the same libconfig problem.cfg; what they
write must agree (bit-exact spring list, accelerations/energy histories to the stated tolerances)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

DROPIN = os.path.join(ROOT, "tests", "dropin")
BUILD = os.path.join(DROPIN, "_build")
HAVE_REF = os.path.isdir("/root/reference/src_spring")


def build():
    if HAVE_REF:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"])
        import asteroid_spring_sims_b200 as A
        A.build.build_native()
        subprocess.check_call(["make", "-s", "-C", DROPIN, "-j8"])
    if not (os.path.exists(os.path.join(BUILD, "synth_ref")) and os.path.exists(os.path.join(BUILD, "synth_b200"))):
        pytest.skip("tests/dropin/_build not built and /root/reference absent")


def run(binary, workdir, gravity, env_extra=None, expect_ok=True, point_mass=0, shape=0, rotate=0, stress=0):
    os.makedirs(workdir, exist_ok=True)
    cfg = open(os.path.join(DROPIN, "problem.cfg")).read()
    cfg = re.sub(r"gravity = \d;", "gravity = %d;" % gravity, cfg)
    cfg = re.sub(r"point_mass = \d;", "point_mass = %d;" % point_mass, cfg)
    cfg = re.sub(r"shape = \d;", "shape = %d;" % shape, cfg)
    cfg = re.sub(r"rotate = \d;", "rotate = %d;" % rotate, cfg)
    cfg = re.sub(r"stress = \d;", "stress = %d;" % stress, cfg)
    open(os.path.join(workdir, "problem.cfg"), "w").write(cfg)
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([os.path.join(BUILD, binary)], cwd=workdir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    if expect_ok:
        assert p.returncode == 0, p.stdout[-2000:]
    return p


def parse(workdir):
    run_lines = open(os.path.join(workdir, "synth_run.txt")).read().split("\n")
    ext = np.loadtxt(os.path.join(workdir, "synth_ext.txt"), ndmin=2)
    fin = open(os.path.join(workdir, "synth_final.txt")).read().split("\n")
    head = {ln.split()[0]: ln.split()[1] for ln in fin[:3]}
    state = np.array([[float.fromhex(x) for x in ln.split()] for ln in fin[3:] if ln.strip()])
    return run_lines, ext, head, state


def parse_res(workdir):
    """<fileroot>_res.txt as written by the reference's write_resolved_with_E (output_spring.cpp:112-192): 25 columns of
    width 14.  The power column (10 significant digits + exponent = 15 characters) always runs into the one before it --
    in the reference's own output -- so the first 23 columns are cut by width and the rest by whitespace; whatever still
    does not parse is NaN."""
    rows = []
    for ln in open(os.path.join(workdir, "synth_res.txt")):
        if ln.startswith("#") or not ln.strip():
            continue
        ln = ln.rstrip("\n")
        vals = []
        for c in range(23):
            try:
                vals.append(float(ln[14 * c:14 * (c + 1)]))
            except ValueError:
                vals.append(np.nan)
        tail = ln[14 * 23:].split()
        for q in range(2):
            try:
                vals.append(float(tail[q]))
            except (ValueError, IndexError):
                vals.append(np.nan)
        rows.append(vals)
    return np.array(rows, ndmin=2)


def compare_res(g, r):
    """Printed with 4-10 significant digits: equal to within a unit of the last printed digit, quantities that are zero
    up to rounding (the centred body's CoM / CoV columns) to an absolute bar."""
    assert g.shape == r.shape and r.shape[1] == 25
    bad = np.isnan(g) | np.isnan(r)
    assert bad.mean() < 0.1, "too little of the history file could be parsed"
    g, r = np.where(bad, 0.0, g), np.where(bad, 0.0, r)
    assert np.array_equal(g[:, 0], r[:, 0])
    scale = np.maximum(np.abs(r), 1e-12)
    rel = np.abs(g - r) / scale
    big = np.abs(r) > 1e-9
    assert np.all(rel[big] <= 2e-4), rel[big].max()           # the coarsest columns carry 4 digits
    assert np.all(np.abs(g[~big]) <= 1e-9)
    # the energy columns carry 10 digits: KErot + PEspr + PEgrav == Etot as printed by both
    assert np.all(np.abs(g[:, 22] - r[:, 22]) <= 1e-8 * np.maximum(np.abs(r[:, 22]), 1e-30))


def test_module_links_both_ways_and_reference_runs(tmp_path):
    build()
    run("synth_ref", str(tmp_path / "ref"), 1)
    run_lines, ext, head, state = parse(str(tmp_path / "ref"))
    assert run_lines[0] == "N 330" and run_lines[1] == "NS 3789"
    assert float.fromhex(head["t"]) == 0.0811           # exact_finish_time landed on t_max
    assert len(ext) >= 4 and state.shape == (330, 10)
    # the drop-in build exports the reference's symbols (it defines them itself)
    syms = subprocess.run(["nm", "--defined-only", os.path.join(BUILD, "synth_b200")], stdout=subprocess.PIPE, text=True).stdout
    for s in ("reb_integrate", "reb_step", "_Z13spring_forcesP14reb_simulation", "_Z20connect_springs_distP14reb_simulationdii6spring",
              "_Z16quadrupole_accelP14reb_simulationddddi", "_Z14drift_resolvedP14reb_simulationdddiii", "_Z9drift_binP14reb_simulationdddii",
              "_Z8sum_massP14reb_simulationii"):
        assert re.search(r" T %s$" % re.escape(s), syms, re.M), s
    # ... and, demangled, everything set-up code and write_resolved_with_E reach for
    dem = subprocess.run(["nm", "-C", "--defined-only", os.path.join(BUILD, "dropin.o")], stdout=subprocess.PIPE, text=True).stdout
    for name in ("spin_body(", "measure_L(", "mom_inertia(", "body_spin(", "compute_rot_kin(", "spring_potential_energy(", "grav_potential_energy(",
                 "dEdt_total(", "rotate_body(", "rotate_origin(", "rotate_to_principal(", "hcp_ellipsoid(", "cubic_ellipsoid(", "rand_rectangle(",
                 "rand_cone(", "rand_shape(", "update_stress("):
        assert re.search(r" T %s" % re.escape(name), dem), name


def test_every_reference_module_links_against_the_dropin():
    """'Modules compile unchanged': each shipped modules/<name>/problem.cpp, compiled where it lies, links against the
    drop-in plus the reference objects whose same-named symbols were weakened."""
    if not HAVE_REF:
        pytest.skip("/root/reference absent")
    build()
    subprocess.check_call(["make", "-s", "-C", DROPIN, "-j8", "modules"])
    names = sorted(d for d in os.listdir("/root/reference/modules")
                   if d != "tests" and os.path.exists(os.path.join("/root/reference/modules", d, "problem.cpp")))
    assert len(names) == 12, names
    for nm_ in names:
        path = os.path.join(BUILD, "mod_%s_b200" % nm_)
        assert os.path.exists(path), nm_
        syms = subprocess.run(["nm", "--defined-only", path], stdout=subprocess.PIPE, text=True).stdout
        # the step loop and the force routine are the drop-in's strong definitions, the reference's are weak (W) or gone
        assert re.search(r" T reb_integrate$", syms, re.M) and re.search(r" T _Z13spring_forcesP14reb_simulation$", syms, re.M), nm_


def test_dropin_fails_loudly_without_a_gpu(tmp_path):
    build()
    import asteroid_spring_sims_b200 as A
    if A.device_count() > 0:
        pytest.skip("a GPU is present")
    p = run("synth_b200", str(tmp_path / "b200"), 1, expect_ok=False)
    assert p.returncode != 0
    assert "Fatal error" in p.stdout and "no CUDA device" in p.stdout   # reb_exit convention, no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("gravity", [0, 1])
@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_matches_reference(tmp_path, gravity, resident):
    build()
    ref_dir, gpu_dir = str(tmp_path / "ref"), str(tmp_path / "b200")
    run("synth_ref", ref_dir, gravity)
    run("synth_b200", gpu_dir, gravity, {"ASS_RESIDENT": resident})
    r_run, r_ext, r_head, r_state = parse(ref_dir)
    g_run, g_ext, g_head, g_state = parse(gpu_dir)
    # generator + spring list: bit-exact (hexfloat text)
    assert g_run == r_run
    assert g_head == r_head                                   # t, dt, status identical (exact finish time)
    # final state
    amax = np.abs(r_state[:, 6:9]).max()
    assert np.abs(g_state[:, 6:9] - r_state[:, 6:9]).max() <= 1e-10 * amax
    assert np.allclose(g_state[:, :6], r_state[:, :6], rtol=0, atol=1e-12)
    assert np.array_equal(g_state[:, 9], r_state[:, 9])
    # history written by the module's heartbeat through the reference's own diagnostics: t, L, KE, SPE, GPE, power
    assert g_ext.shape == r_ext.shape
    assert np.array_equal(g_ext[:, 0], r_ext[:, 0])
    e_r = r_ext[:, 4] + r_ext[:, 5] + r_ext[:, 6]
    e_g = g_ext[:, 4] + g_ext[:, 5] + g_ext[:, 6]
    assert np.all(np.abs(e_g - e_r) <= 1e-9 * np.abs(e_r))
    assert np.abs(g_ext[:, 1:4] - r_ext[:, 1:4]).max() <= 1e-9 * np.abs(r_ext[:, 1:4]).max()
    # the module's own history file, through the reference's writer on top of the drop-in's diagnostics
    compare_res(parse_res(gpu_dir), parse_res(ref_dir))


@pytest.mark.gpu
@pytest.mark.parametrize("gravity,point_mass", [(0, 0), (1, 0), (1, 1)])
def test_resident_mode_defers_and_fuses_with_the_same_bits(tmp_path, gravity, point_mass):
    """In resident mode part1 / zero_accel / spring_forces of a reb_step are deferred until reb_integrator_leapfrog_part2 runs them
    as the fused kernels (dropin.cpp: lazy_ok / settle; quadrupole_accel in the phobos-shaped callback settles them early).
    ASS_DROPIN_NO_FUSE=1 runs every call when it is made: the module's output files must be the same, byte for byte."""
    build()
    outs = []
    for fuse in ("0", "1"):
        d = str(tmp_path / ("fuse" + fuse))
        run("synth_b200", d, gravity, {"ASS_RESIDENT": "1", "ASS_DROPIN_NO_FUSE": "0" if fuse == "1" else "1"}, point_mass=point_mass)
        outs.append(parse(d))
    (a_run, a_ext, a_head, a_state), (b_run, b_ext, b_head, b_state) = outs
    assert a_run == b_run and a_head == b_head
    assert np.array_equal(a_state, b_state)
    assert np.array_equal(a_ext, b_ext)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [1, 2, 3, 4])
def test_dropin_other_shapes_rotations_and_stress(tmp_path, shape):
    """hcp / cubic lattices, the cube module's box, a cone; rotate_body + rotate_to_principal after the springs are
    connected; update_stress after the run.  Generator output and spring list bit-exact, the rest to the usual bars;
    the stress tensors are the reference's bits."""
    build()
    ref_dir, gpu_dir = str(tmp_path / "ref"), str(tmp_path / "b200")
    run("synth_ref", ref_dir, 1, shape=shape, rotate=1, stress=1)
    run("synth_b200", gpu_dir, 1, {"ASS_RESIDENT": "1"}, shape=shape, rotate=1, stress=1)
    r_run, r_ext, r_head, r_state = parse(ref_dir)
    g_run, g_ext, g_head, g_state = parse(gpu_dir)
    # N, NS, mean rest length and the spring list (connected before the rotations): bit-exact.  mindist is taken AFTER
    # rotate_to_principal, whose matrix comes from the inertia tensor -- a parallel sum on the device, equal to the
    # reference's sequential one to rounding -- so that one line is compared as a number.
    assert g_head == r_head and len(g_run) == len(r_run)
    for a, b in zip(g_run, r_run):
        if a.startswith("mindist"):
            assert abs(float.fromhex(a.split()[1]) - float.fromhex(b.split()[1])) <= 1e-13
        else:
            assert a == b
    amax = np.abs(r_state[:, 6:9]).max()
    assert np.abs(g_state[:, 6:9] - r_state[:, 6:9]).max() <= 1e-10 * amax
    assert np.allclose(g_state[:, :6], r_state[:, :6], rtol=0, atol=1e-12)
    compare_res(parse_res(gpu_dir), parse_res(ref_dir))
    # stress: both threw (or neither); tensors of the final state agree to rounding of that state
    rs = open(os.path.join(ref_dir, "synth_stress.txt")).read().split("\n")
    gs = open(os.path.join(gpu_dir, "synth_stress.txt")).read().split("\n")
    assert rs[0] == gs[0] and len(rs) == len(gs)
    R = np.array([[float.fromhex(x) for x in ln.split()[:10]] for ln in rs[1:] if ln.strip()])
    G = np.array([[float.fromhex(x) for x in ln.split()[:10]] for ln in gs[1:] if ln.strip()])
    assert np.abs(G - R).max() <= 1e-9 * np.abs(R).max()
    # the strongest spring of a node: the same wherever the top force is not a near-tie (on a lattice many springs of a
    # node carry equal forces up to rounding, and the two final states differ by ~1e-13)
    ri = np.array([int(ln.split()[10]) for ln in rs[1:] if ln.strip()])
    gi = np.array([int(ln.split()[10]) for ln in gs[1:] if ln.strip()])
    assert (ri == gi).mean() >= 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("resident", ["0", "1"])
def test_dropin_matches_reference_phobos_shape(tmp_path, resident):
    """modules/phobos_deimos's shape: a point mass added by the reference's own add_pt_mass_kep, REB_GRAVITY_BASIC,
    additional_forces = spring_forces + quadrupole_accel, heartbeat = center_sim + drift_resolved."""
    build()
    ref_dir, gpu_dir = str(tmp_path / "ref"), str(tmp_path / "b200")
    run("synth_ref", ref_dir, 1, point_mass=1)
    run("synth_b200", gpu_dir, 1, {"ASS_RESIDENT": resident}, point_mass=1)
    r_run, r_ext, r_head, r_state = parse(ref_dir)
    g_run, g_ext, g_head, g_state = parse(gpu_dir)
    assert g_run == r_run and g_head == r_head
    assert r_state.shape == (331, 10) and r_state[-1, 9] == 1.0            # the point mass is the last particle
    amax = np.abs(r_state[:, 6:9]).max()
    assert np.abs(g_state[:, 6:9] - r_state[:, 6:9]).max() <= 1e-10 * amax
    assert np.allclose(g_state[:, :6], r_state[:, :6], rtol=0, atol=1e-12)
    e_r = r_ext[:, 4] + r_ext[:, 5] + r_ext[:, 6]
    e_g = g_ext[:, 4] + g_ext[:, 5] + g_ext[:, 6]
    assert np.all(np.abs(e_g - e_r) <= 1e-9 * np.abs(e_r))
    assert np.abs(g_ext[:, 1:4] - r_ext[:, 1:4]).max() <= 1e-9 * np.abs(r_ext[:, 1:4]).max()
