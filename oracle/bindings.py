"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two libraries:
  * ``libass_oracle.so``      oracle/ass_oracle.c, the plain-C restatement (class ``Oracle``)
  * ``_ref/libref_oracle.so`` the UNMODIFIED reference compiled by ``make -C oracle ref`` with the shim in
                              oracle/ref_shim.cpp (class ``Ref``); absent on machines without /root/reference
                              unless the prebuilt .so travelled with the snapshot.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Particle state is an (N, 11) float64 array: x y z vx vy vz ax ay az m r.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ASS_REFERENCE", "/root/reference")
STRIDE = 11
X, Y, Z, VX, VY, VZ, AX, AY, AZ, M, R = range(11)

SPRING_DTYPE = np.dtype([("k", "<f8"), ("rs0", "<f8"), ("gamma", "<f8"),
                         ("particle_1", "<i4"), ("particle_2", "<i4")], align=True)
assert SPRING_DTYPE.itemsize == 32

_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


class OrcSim(C.Structure):
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("dt_last_done", C.c_double),
                ("G", C.c_double), ("softening", C.c_double),
                ("gravity_basic", C.c_int), ("n_active", C.c_int), ("zero_first", C.c_int),
                ("center_heartbeat", C.c_int), ("n_perts", C.c_int), ("status", C.c_int),
                ("exact_finish_time", C.c_int)]


def make_sim(dt=1e-3, G=1.0, softening=0.0, gravity_basic=0, zero_first=0, center_heartbeat=0, n_perts=0, t=0.0):
    return OrcSim(t=t, dt=dt, dt_last_done=0.0, G=G, softening=softening, gravity_basic=int(gravity_basic),
                  n_active=-1, zero_first=int(zero_first), center_heartbeat=int(center_heartbeat),
                  n_perts=n_perts, status=-1, exact_finish_time=1)


def build_oracle(force=False):
    so = os.path.join(HERE, "libass_oracle.so")
    src = os.path.join(HERE, "ass_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libass_oracle.so"])
    return so


def build_ref(force=False):
    """Compile the unmodified reference into oracle/_ref/ when its sources are present. Returns path or None."""
    so = os.path.join(HERE, "_ref", "libref_oracle.so")
    if os.path.isdir(os.path.join(REFERENCE, "src_spring")):
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "ref_shim.cpp")):
            subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref", "REF=" + REFERENCE])
    return so if os.path.exists(so) else None


class Oracle:
    """oracle/ass_oracle.c through ctypes."""

    def __init__(self):
        self.lib = L = C.CDLL(build_oracle())
        L.orc_mindist.restype = C.c_double
        L.orc_mindist_n2.restype = C.c_double
        L.orc_sum_mass.restype = C.c_double
        L.orc_random_uniform.restype = C.c_double
        L.orc_random_uniform.argtypes = [C.c_double, C.c_double]
        self.libc = C.CDLL(None)

    @staticmethod
    def _p(p):
        assert p.dtype == np.float64 and p.ndim == 2 and p.shape[1] == STRIDE and p.flags.c_contiguous
        return _d(p)

    @staticmethod
    def _s(s):
        assert s.dtype == SPRING_DTYPE and s.flags.c_contiguous
        return s.ctypes.data_as(C.c_void_p)

    def srand(self, seed):
        self.libc.srand(C.c_uint(seed))

    # -- per-step path --
    def zero_accel(self, p):
        self.lib.orc_zero_accel(self._p(p), C.c_int(len(p)))

    def spring_forces(self, p, s):
        self.lib.orc_spring_forces(self._p(p), self._s(s), C.c_int(len(s)))

    def gravity_basic(self, p, G=1.0, softening=0.0, n_active=-1, i_lo=None, i_hi=None):
        if i_lo is None:
            self.lib.orc_gravity_basic(self._p(p), C.c_int(len(p)), C.c_int(n_active), C.c_double(G), C.c_double(softening))
        else:
            self.lib.orc_gravity_basic_slab(self._p(p), C.c_int(len(p)), C.c_int(n_active), C.c_double(G),
                                            C.c_double(softening), C.c_int(i_lo), C.c_int(i_hi))

    def leapfrog_part1(self, p, sim):
        self.lib.orc_leapfrog_part1(self._p(p), C.c_int(len(p)), C.byref(sim))

    def leapfrog_part2(self, p, sim):
        self.lib.orc_leapfrog_part2(self._p(p), C.c_int(len(p)), C.byref(sim))

    def step(self, p, s, sim, nsteps=1, heartbeat=False):
        for _ in range(nsteps):
            self.lib.orc_step(self._p(p), C.c_int(len(p)), self._s(s), C.c_int(len(s)), C.byref(sim))
            if heartbeat:
                self.lib.orc_heartbeat(self._p(p), C.c_int(len(p)), C.byref(sim))

    def integrate(self, p, s, sim, tmax):
        return self.lib.orc_integrate(self._p(p), C.c_int(len(p)), self._s(s), C.c_int(len(s)), C.byref(sim), C.c_double(tmax))

    # -- spring build --
    def connect_springs_dist(self, p, max_dist, i_low, i_high, k, gamma, existing=None, n2=False):
        fn = self.lib.orc_connect_springs_dist_n2 if n2 else self.lib.orc_connect_springs_dist
        ptr = C.c_void_p(None)
        ns = C.c_int(0)
        cap = C.c_int(0)
        if existing is not None and len(existing):
            # hand the C side a malloc'd copy it may realloc
            self.libc.malloc.restype = C.c_void_p
            nbytes = existing.nbytes
            ptr = C.c_void_p(self.libc.malloc(C.c_size_t(nbytes)))
            C.memmove(ptr, existing.ctypes.data, nbytes)
            ns = C.c_int(len(existing))
            cap = C.c_int(len(existing))
        rc = fn(self._p(p), C.c_double(max_dist), C.c_int(i_low), C.c_int(i_high), C.c_double(k), C.c_double(gamma),
                C.byref(ptr), C.byref(ns), C.byref(cap))
        assert rc >= 0
        out = np.empty(ns.value, dtype=SPRING_DTYPE)
        if ns.value:
            C.memmove(out.ctypes.data, ptr, out.nbytes)
        if ptr.value:
            self.lib.orc_free(ptr)
        return out

    # -- generators --
    def _gen(self, fn, min_dist, a, b, c, total_mass, existing=None):
        self.libc.malloc.restype = C.c_void_p
        n0 = 0 if existing is None else len(existing)
        ptr = C.c_void_p(None)
        if n0:
            ptr = C.c_void_p(self.libc.malloc(C.c_size_t(existing.nbytes)))
            C.memmove(ptr, existing.ctypes.data, existing.nbytes)
        n = fn(C.byref(ptr), C.c_int(n0), C.c_double(min_dist), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(total_mass))
        assert n >= 0
        out = np.empty((n, STRIDE), dtype=np.float64)
        if n:
            C.memmove(out.ctypes.data, ptr, out.nbytes)
        if ptr.value:
            self.lib.orc_free(ptr)
        return out

    def rand_ellipsoid(self, min_dist, a, b, c, total_mass, existing=None, n2=False):
        return self._gen(self.lib.orc_rand_ellipsoid_n2 if n2 else self.lib.orc_rand_ellipsoid, min_dist, a, b, c, total_mass, existing)

    def hcp_ellipsoid(self, min_dist, a, b, c, total_mass, existing=None):
        return self._gen(self.lib.orc_hcp_ellipsoid, min_dist, a, b, c, total_mass, existing)

    def cubic_ellipsoid(self, min_dist, a, b, c, total_mass, existing=None):
        return self._gen(self.lib.orc_cubic_ellipsoid, min_dist, a, b, c, total_mass, existing)

    def rand_rectangle(self, min_dist, x, y, z, total_mass, existing=None):
        return self._gen(self.lib.orc_rand_rectangle_n2, min_dist, x, y, z, total_mass, existing)

    def _gen_n(self, fn, args, existing=None):
        self.libc.malloc.restype = C.c_void_p
        n0 = 0 if existing is None else len(existing)
        ptr = C.c_void_p(None)
        if n0:
            existing = np.ascontiguousarray(existing, dtype=np.float64)
            ptr = C.c_void_p(self.libc.malloc(C.c_size_t(existing.nbytes)))
            C.memmove(ptr, existing.ctypes.data, existing.nbytes)
        n = fn(C.byref(ptr), C.c_int(n0), *[C.c_double(a) for a in args])
        assert n >= 0
        out = np.empty((n, STRIDE), dtype=np.float64)
        if n:
            C.memmove(out.ctypes.data, ptr, out.nbytes)
        if ptr.value:
            self.lib.orc_free(ptr)
        return out

    def rand_cone(self, min_dist, radius, height, total_mass):
        return self._gen_n(self.lib.orc_rand_cone_n2, (min_dist, radius, height, total_mass))

    def rand_shape(self, vertices, min_dist, total_mass):
        return self._gen_n(self.lib.orc_rand_shape_n2, (min_dist, total_mass), vertices)

    def euler_matrix(self, alpha, beta, gamma):
        out = (C.c_double * 9)()
        self.lib.orc_euler_matrix(C.c_double(alpha), C.c_double(beta), C.c_double(gamma), out)
        return np.array(out[:])

    def rotate_body(self, p, lo, hi, alpha, beta, gamma):
        self.lib.orc_rotate_body(self._p(p), C.c_int(lo), C.c_int(hi), C.c_double(alpha), C.c_double(beta), C.c_double(gamma))

    def rotate_origin(self, p, lo, hi, alpha, beta, gamma):
        self.lib.orc_rotate_origin(self._p(p), C.c_int(lo), C.c_int(hi), C.c_double(alpha), C.c_double(beta), C.c_double(gamma))

    def stress(self, p, s):
        n = len(p)
        T = np.zeros((n, 9)); F = np.zeros(n); S = np.zeros(n, dtype=np.int32)
        self.lib.orc_stress(self._p(p), C.c_int(n), self._s(s), C.c_int(len(s)), T.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p))
        return T.reshape(n, 3, 3), F, S

    def apply_impact(self, p, n_body, is_surf, lat, longi, dangle, force, envelope):
        m = np.ascontiguousarray(is_surf, dtype=np.uint8)
        return self.lib.orc_apply_impact(self._p(p), C.c_int(n_body), m.ctypes.data_as(C.c_void_p), C.c_double(lat), C.c_double(longi), C.c_double(dangle),
                                         C.c_double(force), C.c_double(envelope))

    def top_push(self, p, top, pressure):
        m = np.ascontiguousarray(top, dtype=np.uint8)
        self.lib.orc_top_push(self._p(p), C.c_int(len(p)), m.ctypes.data_as(C.c_void_p), C.c_double(pressure))

    def table_top(self, p, bottom):
        m = np.ascontiguousarray(bottom, dtype=np.uint8)
        self.lib.orc_table_top(self._p(p), C.c_int(len(p)), m.ctypes.data_as(C.c_void_p))

    def mindist(self, p, lo=0, hi=None, n2=False):
        hi = len(p) if hi is None else hi
        fn = self.lib.orc_mindist_n2 if n2 else self.lib.orc_mindist
        return fn(self._p(p), C.c_int(lo), C.c_int(hi))

    # -- helpers / diagnostics --
    def center_sim(self, p, lo, hi):
        self.lib.orc_center_sim(self._p(p), C.c_int(len(p)), C.c_int(lo), C.c_int(hi))

    def subtract_com(self, p, lo, hi):
        self.lib.orc_subtract_com(self._p(p), C.c_int(lo), C.c_int(hi))

    def subtract_cov(self, p, lo, hi):
        self.lib.orc_subtract_cov(self._p(p), C.c_int(lo), C.c_int(hi))

    def spin_body(self, p, lo, hi, omega):
        self.lib.orc_spin_body(self._p(p), C.c_int(lo), C.c_int(hi), C.c_double(omega[0]), C.c_double(omega[1]), C.c_double(omega[2]))

    def diagnostics(self, p, s, lo=0, hi=None, G=1.0, want_gpe=False):
        hi = len(p) if hi is None else hi
        out = np.zeros(20)
        self.lib.orc_diagnostics(self._p(p), C.c_int(lo), C.c_int(hi), self._s(s), C.c_int(len(s)), C.c_double(G), C.c_int(int(want_gpe)), _d(out))
        return out


    def quadrupole_accel(self, p, G, J2, R, phi, theta, i_p):
        self.lib.orc_quadrupole_accel(self._p(p), C.c_int(len(p)), C.c_double(G), C.c_double(J2), C.c_double(R), C.c_double(phi),
                                      C.c_double(theta), C.c_int(i_p))

    def drift_resolved(self, p, G, timestep, inv_tau_a, inv_tau_e, i_part, lo, hi):
        self.lib.orc_drift_resolved(self._p(p), C.c_double(G), C.c_double(timestep), C.c_double(inv_tau_a), C.c_double(inv_tau_e),
                                    C.c_int(i_part), C.c_int(lo), C.c_int(hi))

    def drift_bin(self, p, G, timestep, inv_tau_a, inv_tau_e, part_1, part_2):
        self.lib.orc_drift_bin(self._p(p), C.c_double(G), C.c_double(timestep), C.c_double(inv_tau_a), C.c_double(inv_tau_e),
                               C.c_int(part_1), C.c_int(part_2))


class Ref:
    """The unmodified reference (oracle/_ref/libref_oracle*.so). One live simulation at a time (its spring list is a
    process-wide global, as in the reference's modules)."""

    def __init__(self, omp=False):
        so = build_ref()
        if so is None:
            raise FileNotFoundError("oracle/_ref/libref_oracle.so not built and /root/reference absent")
        if omp:
            so = so.replace("libref_oracle.so", "libref_oracle_omp.so")
        self.lib = L = C.CDLL(so)
        L.ref_new.restype = C.c_void_p
        for name in ("ref_time", "ref_dt", "ref_mindist", "ref_mean_spring_length", "ref_strain"):
            getattr(L, name).restype = C.c_double
        L.ref_quiet(1)
        self.h = C.c_void_p(L.ref_new())

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def seed(self, s):
        self.lib.ref_seed(C.c_uint(s))

    def config(self, gravity_basic=0, G=1.0, softening=0.0, dt=1e-3, zero_first=0, center_heartbeat=0, n_perts=0):
        self.lib.ref_config(self.h, C.c_int(int(gravity_basic)), C.c_double(G), C.c_double(softening), C.c_double(dt),
                            C.c_int(int(zero_first)), C.c_int(int(center_heartbeat)), C.c_int(n_perts))

    @property
    def n(self):
        return self.lib.ref_n(self.h)

    @property
    def t(self):
        return self.lib.ref_time(self.h)

    @property
    def dt(self):
        return self.lib.ref_dt(self.h)

    def set_time(self, t):
        self.lib.ref_set_time(self.h, C.c_double(t))

    def rand_ellipsoid(self, d, a, b, c, M):
        return self.lib.ref_rand_ellipsoid(self.h, C.c_double(d), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(M))

    def hcp_ellipsoid(self, d, a, b, c, M):
        return self.lib.ref_hcp_ellipsoid(self.h, C.c_double(d), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(M))

    def cubic_ellipsoid(self, d, a, b, c, M):
        return self.lib.ref_cubic_ellipsoid(self.h, C.c_double(d), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(M))

    def stress(self):
        """(N x 3 x 3 tensors, max_force, max_force_spring, threw) of update_stress (stress.cpp:32-98)."""
        n = self.n
        T = np.zeros((n, 9)); F = np.zeros(n); S = np.zeros(n, dtype=np.int32)
        threw = self.lib.ref_stress(self.h, T.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p))
        return T.reshape(n, 3, 3), F, S, int(threw)

    def rand_rectangle(self, d, x, y, z, M):
        return self.lib.ref_rand_rectangle(self.h, C.c_double(d), C.c_double(x), C.c_double(y), C.c_double(z), C.c_double(M))

    def rand_cone(self, d, radius, height, M):
        return self.lib.ref_rand_cone(self.h, C.c_double(d), C.c_double(radius), C.c_double(height), C.c_double(M))

    def rand_shape(self, d, M):
        return self.lib.ref_rand_shape(self.h, C.c_double(d), C.c_double(M))

    def rotate_body(self, lo, hi, alpha, beta, gamma):
        self.lib.ref_rotate_body(self.h, C.c_int(lo), C.c_int(hi), C.c_double(alpha), C.c_double(beta), C.c_double(gamma))

    def rotate_origin(self, lo, hi, alpha, beta, gamma):
        self.lib.ref_rotate_origin(self.h, C.c_int(lo), C.c_int(hi), C.c_double(alpha), C.c_double(beta), C.c_double(gamma))

    def rotate_to_principal(self, lo, hi):
        self.lib.ref_rotate_to_principal(self.h, C.c_int(lo), C.c_int(hi))

    def euler_matrix(self, alpha, beta, gamma):
        out = (C.c_double * 9)()
        self.lib.ref_euler_matrix(C.c_double(alpha), C.c_double(beta), C.c_double(gamma), out)
        return np.array(out[:])

    def mindist(self, lo, hi):
        return self.lib.ref_mindist(self.h, C.c_int(lo), C.c_int(hi))

    def add_particles(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        for row in p:
            self.lib.ref_add_particle(self.h, _d(np.ascontiguousarray(row)))

    def get(self):
        out = np.empty((self.n, STRIDE))
        self.lib.ref_get_particles(self.h, _d(out))
        return out

    def set(self, p):
        assert p.shape == (self.n, STRIDE)
        self.lib.ref_set_particles(self.h, _d(np.ascontiguousarray(p)))

    def connect_springs_dist(self, max_dist, lo, hi, k, gamma):
        return self.lib.ref_connect_springs_dist(self.h, C.c_double(max_dist), C.c_int(lo), C.c_int(hi), C.c_double(k), C.c_double(gamma))

    def add_spring(self, i, j, k, rs0, gamma):
        return self.lib.ref_add_spring(self.h, C.c_int(i), C.c_int(j), C.c_double(k), C.c_double(rs0), C.c_double(gamma))

    def springs(self):
        n = self.lib.ref_num_springs()
        out = np.zeros(n, dtype=SPRING_DTYPE)
        if n:
            self.lib.ref_get_springs(out.ctypes.data_as(C.c_void_p))
        return out

    def set_springs(self, s):
        s = np.ascontiguousarray(s)
        self.lib.ref_set_springs(s.ctypes.data_as(C.c_void_p), C.c_int(len(s)))

    def set_gamma(self, g):
        self.lib.ref_set_gamma(C.c_double(g))

    def spring_i_force(self, s, damped=True):
        out = np.zeros(3)
        self.lib.ref_spring_i_force(self.h, C.c_int(s), C.c_int(int(damped)), _d(out))
        return out

    def zero_accel(self):
        self.lib.ref_zero_accel(self.h)

    def spring_forces(self):
        self.lib.ref_spring_forces(self.h)

    def gravity(self):
        self.lib.ref_gravity(self.h)

    def leapfrog_part1(self):
        self.lib.ref_leapfrog_part1(self.h)

    def leapfrog_part2(self):
        self.lib.ref_leapfrog_part2(self.h)

    def step(self, n=1, heartbeat=False):
        (self.lib.ref_step_hb if heartbeat else self.lib.ref_step)(self.h, C.c_int(n))

    def integrate(self, tmax):
        return self.lib.ref_integrate(self.h, C.c_double(tmax))

    def center_sim(self, lo, hi):
        self.lib.ref_center_sim(self.h, C.c_int(lo), C.c_int(hi))

    def subtract_com(self, lo, hi):
        self.lib.ref_subtract_com(self.h, C.c_int(lo), C.c_int(hi))

    def subtract_cov(self, lo, hi):
        self.lib.ref_subtract_cov(self.h, C.c_int(lo), C.c_int(hi))

    def spin_body(self, lo, hi, omega):
        self.lib.ref_spin_body(self.h, C.c_int(lo), C.c_int(hi), C.c_double(omega[0]), C.c_double(omega[1]), C.c_double(omega[2]))

    def diagnostics(self, lo, hi, want_gpe=False):
        out = np.zeros(20)
        self.lib.ref_diagnostics(self.h, C.c_int(lo), C.c_int(hi), C.c_int(int(want_gpe)), _d(out))
        return out

    def quadrupole_accel(self, J2, R, phi, theta, i_p):
        self.lib.ref_quadrupole_accel(self.h, C.c_double(J2), C.c_double(R), C.c_double(phi), C.c_double(theta), C.c_int(i_p))

    def drift_resolved(self, timestep, inv_tau_a, inv_tau_e, i_part, lo, hi):
        self.lib.ref_drift_resolved(self.h, C.c_double(timestep), C.c_double(inv_tau_a), C.c_double(inv_tau_e), C.c_int(i_part),
                                    C.c_int(lo), C.c_int(hi))

    def drift_bin(self, timestep, inv_tau_a, inv_tau_e, part_1, part_2):
        self.lib.ref_drift_bin(self.h, C.c_double(timestep), C.c_double(inv_tau_a), C.c_double(inv_tau_e), C.c_int(part_1), C.c_int(part_2))
