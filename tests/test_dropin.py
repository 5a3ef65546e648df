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


def run(binary, workdir, gravity, env_extra=None, expect_ok=True, point_mass=0):
    os.makedirs(workdir, exist_ok=True)
    cfg = open(os.path.join(DROPIN, "problem.cfg")).read()
    cfg = re.sub(r"gravity = \d;", "gravity = %d;" % gravity, cfg)
    cfg = re.sub(r"point_mass = \d;", "point_mass = %d;" % point_mass, cfg)
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
