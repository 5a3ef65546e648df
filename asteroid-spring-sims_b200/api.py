"""ctypes binding over the C-ABI in include/ass_b200.h.

This is measurement / test plumbing, not a second implementation: every method is one C call into libass_b200.so.
If the library is missing or there is no B200, construction raises -- there is no CPU fallback.

Particle state on the host is an (N, 11) float64 array: x y z vx vy vz ax ay az m r (the leading fields of
``struct reb_particle``, reference src/rebound.h:540-551) or any array of rows with that prefix and a byte stride.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

F_POS, F_VEL, F_ACC, F_MASS, F_ALL = 1, 2, 4, 8, 15
K_NAMES = ("gravity", "springs", "leapfrog", "reduce", "connect", "xfer")

SPRING_DTYPE = np.dtype([("k", "<f8"), ("rs0", "<f8"), ("gamma", "<f8"),
                         ("particle_1", "<i4"), ("particle_2", "<i4")], align=True)
assert SPRING_DTYPE.itemsize == 32


class Params(C.Structure):
    """struct ass_params"""
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("dt_last_done", C.c_double),
                ("G", C.c_double), ("softening", C.c_double),
                ("gravity", C.c_int), ("n_active", C.c_int), ("n_perts", C.c_int),
                ("additional_forces", C.c_int), ("heartbeat", C.c_int),
                ("exact_finish_time", C.c_int), ("status", C.c_int)]


class AssError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libass_b200: error %d: %s" % (code, text))
        self.code = code


_lib = None


def lib():
    """Load (building if needed) libass_b200.so. Raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build_native()
    L = C.CDLL(path)
    L.ass_last_error.restype = C.c_char_p
    L.ass_launch_count.restype = C.c_longlong
    _lib = L
    return L


def _chk(rc):
    if rc < 0:
        raise AssError(rc, lib().ass_last_error().decode())
    return rc


def device_count():
    return lib().ass_device_count()


def init(device=0):
    _chk(lib().ass_init(C.c_int(device)))


def device_info():
    sm, cc, khz = C.c_int(), C.c_int(), C.c_int()
    l2, mem = C.c_longlong(), C.c_longlong()
    _chk(lib().ass_device_info(C.byref(sm), C.byref(l2), C.byref(mem), C.byref(khz), C.byref(cc)))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "mem_bytes": mem.value, "sm_clock_khz": khz.value, "cc": cc.value}


def launch_count():
    return lib().ass_launch_count()


def probe_fp64_peak(iters=1 << 14):
    tf, ms = C.c_double(), C.c_double()
    _chk(lib().ass_probe_fp64_peak(C.c_int(iters), C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def probe_copy_bw(nbytes=1 << 30):
    g = C.c_double()
    _chk(lib().ass_probe_copy_bw(C.c_longlong(nbytes), C.byref(g)))
    return g.value


class PinnedRows:
    """An (n, cols) float64 row array in page-locked, device-mapped host memory (``ass_host_alloc``): the host side of
    an upload / download at PCIe link rate.  ``.a`` is the numpy view; keep this object alive while the view is used."""

    def __init__(self, n, cols=16):
        self._p = C.c_void_p()
        self.nbytes = max(int(n) * int(cols) * 8, 8)
        _chk(lib().ass_host_alloc(C.byref(self._p), C.c_size_t(self.nbytes)))
        buf = (C.c_double * (int(n) * int(cols))).from_address(self._p.value)
        self.a = np.frombuffer(buf, dtype=np.float64).reshape(int(n), int(cols))

    def close(self):
        if self._p:
            self.a = None
            lib().ass_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_register(arr):
    """Pin an existing numpy array in place (``ass_host_register``); pair with host_unregister before it is freed."""
    _chk(lib().ass_host_register(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes)))


def host_unregister(arr):
    _chk(lib().ass_host_unregister(C.c_void_p(arr.ctypes.data)))


def selfcheck_sqrt(samples=1 << 28, exp_lo=-200, exp_hi=200):
    """(mismatches, first differing argument) of the spring kernels' branch-free sqrt against __dsqrt_rn."""
    bad, first = C.c_longlong(), C.c_double()
    _chk(lib().ass_selfcheck_sqrt(C.c_longlong(samples), C.c_int(exp_lo), C.c_int(exp_hi), C.byref(bad), C.byref(first)))
    return bad.value, first.value


def selfcheck_shard_division(bounds, n_body):
    """Host-only check of the division of a sharded body's gravity by pairs of slabs; needs no GPU.  'violations' must be 0."""
    b = (C.c_int * len(bounds))(*[int(x) for x in bounds])
    stats = (C.c_longlong * 8)()
    _chk(lib().ass_selfcheck_shard_division(C.c_int(len(bounds) - 1), b, C.c_int(n_body), stats))
    keys = ("pieces", "segments", "violations", "most_pairs", "fewest_pairs")
    return dict(zip(keys, [int(v) for v in stats]))


def selfcheck_gravity_plan(n_res, n_str=0, rect=False, sm_count=148):
    """Host-only check of the pair-once gravity plan (block cut, CTA list); needs no GPU.  'violations' must be 0."""
    stats = (C.c_longlong * 8)()
    _chk(lib().ass_selfcheck_gravity_plan(C.c_int(n_res), C.c_int(n_str), C.c_int(int(rect)), C.c_int(sm_count), stats))
    keys = ("ctas", "residents", "cut", "tiles_per_cta", "resident_blocks", "streamed_blocks", "violations", "skipped")
    return dict(zip(keys, [int(v) for v in stats]))


def spring_lanes():
    """Lanes sharing one node's row in the spring kernels; a slice of the lane-transposed list holds 32 // lanes nodes."""
    return int(lib().ass_spring_lanes())


def selfcheck_spring_layout(p, springs, grid=148, slots=1 << 20):
    """Host-only check of what ass_set_springs builds (ranking, CSR and lane-transposed half-edge copies, tile cut for
    `grid` CTAs with tables of at most `slots` entries); needs no GPU.  Returns a dict of counts; 'violations' must be 0."""
    pos = np.ascontiguousarray(np.column_stack([p[:, 0], p[:, 1], p[:, 2], p[:, 9]]), dtype=np.float64)
    s, ptr, n = _springs(springs)
    stats = (C.c_longlong * 8)()
    _chk(lib().ass_selfcheck_spring_layout(pos.ctypes.data_as(C.c_void_p), C.c_int(len(pos)), ptr, n, C.c_int(grid), C.c_int(slots), stats))
    keys = ("slices", "padding", "tiles", "max_slots", "gather_passes", "gather_passes_min", "violations", "records")
    return dict(zip(keys, [int(v) for v in stats]))


def _rows(p):
    assert isinstance(p, np.ndarray) and p.dtype == np.float64 and p.ndim == 2 and p.shape[1] >= 10
    assert p.strides[1] == 8, "rows must be contiguous doubles"
    return p.ctypes.data_as(C.c_void_p), C.c_size_t(p.strides[0]), C.c_int(p.shape[0])


def _springs(s):
    s = np.ascontiguousarray(s, dtype=SPRING_DTYPE)
    return s, s.ctypes.data_as(C.c_void_p), C.c_int(len(s))


class Sim:
    """One simulated body resident on the GPU (``ass_sim``)."""

    def __init__(self):
        self._h = C.c_void_p()
        _chk(lib().ass_create(C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ass_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters --
    @property
    def params(self):
        p = Params()
        _chk(lib().ass_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = self.params
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        _chk(lib().ass_set_params(self._h, C.byref(p)))

    @property
    def n(self):
        return lib().ass_num_particles(self._h)

    @property
    def ns(self):
        return lib().ass_num_springs(self._h)

    # -- state --
    def upload(self, p, fields=F_ALL):
        base, stride, n = _rows(p)
        _chk(lib().ass_upload_particles(self._h, base, stride, n, C.c_uint(fields)))

    def download(self, p, fields=F_ALL):
        base, stride, n = _rows(p)
        _chk(lib().ass_download_particles(self._h, base, stride, n, C.c_uint(fields)))
        return p

    def set_springs(self, s):
        s, ptr, n = _springs(s)
        _chk(lib().ass_set_springs(self._h, ptr, n))

    def update_spring_props(self, s):
        s, ptr, n = _springs(s)
        _chk(lib().ass_update_spring_props(self._h, ptr, n))

    def sync(self):
        _chk(lib().ass_sync(self._h))

    # -- the path --
    def zero_accel(self):
        _chk(lib().ass_zero_accel(self._h))

    def spring_forces(self):
        _chk(lib().ass_spring_forces(self._h))

    def gravity_basic(self):
        _chk(lib().ass_gravity_basic(self._h))

    def set_gravity_kernel(self, which):
        """0 = by size (default), 1 = one-sided source tiles, 2 = pair-once (Newton's third law)."""
        _chk(lib().ass_set_gravity_kernel(self._h, C.c_int(int(which))))

    def probe_gravity_pair(self, rlo, rhi, slo, shi, reps=5):
        """Device time (ms) of pair-once gravity between two particle ranges (identical or disjoint), L2 flushed per repetition."""
        ms = C.c_double()
        _chk(lib().ass_probe_gravity_pair(self._h, C.c_int(rlo), C.c_int(rhi), C.c_int(slo), C.c_int(shi), C.c_int(reps), C.byref(ms)))
        return ms.value

    def probe_gravity_pieces(self, ranges, concurrent, reps=5):
        """Device time (ms) of several pair-once gravity pieces [(rlo, rhi, slo, shi), ...] per repetition, run one after the
        other or concurrently (pair kernels on their own streams, reductions afterwards)."""
        flat = (C.c_int * (4 * len(ranges)))(*[int(v) for r in ranges for v in r])
        ms = C.c_double()
        _chk(lib().ass_probe_gravity_pieces(self._h, C.c_int(len(ranges)), flat, C.c_int(int(concurrent)), C.c_int(reps), C.byref(ms)))
        return ms.value

    def leapfrog_part1(self):
        _chk(lib().ass_leapfrog_part1(self._h))

    def leapfrog_part2(self):
        _chk(lib().ass_leapfrog_part2(self._h))

    def step(self, nsteps=1, heartbeat=False):
        _chk(lib().ass_step(self._h, C.c_int(nsteps), C.c_int(int(heartbeat))))

    def step_timed(self, nsteps=1, heartbeat=False):
        """ass_step with an L2 flush before and a device-timed lap around every step; returns the sum of the laps in ms."""
        ms = C.c_double()
        _chk(lib().ass_step_timed(self._h, C.c_int(nsteps), C.c_int(int(heartbeat)), C.byref(ms)))
        return ms.value

    def integrate(self, tmax):
        return _chk(lib().ass_integrate(self._h, C.c_double(tmax)))

    # -- helpers --
    def set_sum_order(self, sequential=True):
        """CoM / CoV sums in the reference's particle order (bit-exact, default) or as a parallel tree."""
        _chk(lib().ass_set_sum_order(self._h, C.c_int(int(sequential))))

    def center_sim(self, lo, hi):
        _chk(lib().ass_center_sim(self._h, C.c_int(lo), C.c_int(hi)))

    def subtract_com(self, lo, hi):
        _chk(lib().ass_subtract_com(self._h, C.c_int(lo), C.c_int(hi)))

    def subtract_cov(self, lo, hi):
        _chk(lib().ass_subtract_cov(self._h, C.c_int(lo), C.c_int(hi)))

    def move_resolved(self, dx, dv, lo, hi):
        a = (C.c_double * 3)(*dx)
        b = (C.c_double * 3)(*dv)
        _chk(lib().ass_move_resolved(self._h, a, b, C.c_int(lo), C.c_int(hi)))

    def compute_com(self, lo, hi):
        out = (C.c_double * 3)()
        _chk(lib().ass_compute_com(self._h, C.c_int(lo), C.c_int(hi), out))
        return np.array(out[:])

    def compute_cov(self, lo, hi):
        out = (C.c_double * 3)()
        _chk(lib().ass_compute_cov(self._h, C.c_int(lo), C.c_int(hi), out))
        return np.array(out[:])

    def spin_body(self, lo, hi, omega):
        w = (C.c_double * 3)(*omega)
        _chk(lib().ass_spin_body(self._h, C.c_int(lo), C.c_int(hi), w))

    def transform_body(self, lo, hi, R, about_com=True):
        """x <- R (x - c) + c, v <- R (v - u) + u: rotate_body (about the range's CoM / CoV) or rotate_origin / rotate_to_principal."""
        m = (C.c_double * 9)(*[float(v) for v in np.asarray(R, dtype=np.float64).reshape(9)])
        _chk(lib().ass_transform_body(self._h, C.c_int(lo), C.c_int(hi), m, C.c_int(int(about_com))))

    def diagnostics(self, lo=0, hi=None, want_gpe=False):
        hi = self.n if hi is None else hi
        out = (C.c_double * 20)()
        _chk(lib().ass_diagnostics(self._h, C.c_int(lo), C.c_int(hi), C.c_int(int(want_gpe)), out))
        return np.array(out[:])

    # -- stress (stress.cpp) and the masked force terms of the bennu2 / cube modules --
    def stress(self):
        """(N x 3 x 3 stress tensors, max_force, max_force_spring): the first half of update_stress."""
        n = self.n
        T = np.zeros((n, 9)); F = np.zeros(n); S = np.zeros(n, dtype=np.int32)
        _chk(lib().ass_stress(self._h, T.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p)))
        return T.reshape(n, 3, 3), F, S

    def set_node_mask(self, mask_id, mask):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        assert len(m) == self.n
        _chk(lib().ass_set_node_mask(self._h, C.c_int(mask_id), m.ctypes.data_as(C.c_void_p)))

    def mask_from_plane(self, mask_id, lo, hi, axis, threshold, greater=True):
        c = C.c_int()
        _chk(lib().ass_mask_from_plane(self._h, C.c_int(mask_id), C.c_int(lo), C.c_int(hi), C.c_int(axis), C.c_double(threshold), C.c_int(int(greater)), C.byref(c)))
        return c.value

    def apply_radial_impulse(self, mask_id, lo, hi, lat, longi, dangle, force, envelope, want_count=True):
        c = C.c_int()
        _chk(lib().ass_apply_radial_impulse(self._h, C.c_int(mask_id), C.c_int(lo), C.c_int(hi), C.c_double(lat), C.c_double(longi), C.c_double(dangle),
                                            C.c_double(force), C.c_double(envelope), C.byref(c) if want_count else None))
        return c.value

    def plane_push(self, mask_id, axis, per_node_force):
        _chk(lib().ass_plane_push(self._h, C.c_int(mask_id), C.c_int(axis), C.c_double(per_node_force)))

    def plane_constraint(self, mask_id, axis):
        _chk(lib().ass_plane_constraint(self._h, C.c_int(mask_id), C.c_int(axis)))

    # -- phobos* terms (orb.cpp) --
    def quadrupole_accel(self, J2, R, phi, theta, i_p):
        _chk(lib().ass_quadrupole_accel(self._h, C.c_double(J2), C.c_double(R), C.c_double(phi), C.c_double(theta), C.c_int(i_p)))

    def drift_resolved(self, timestep, inv_tau_a, inv_tau_e, i_part, lo, hi):
        _chk(lib().ass_drift_resolved(self._h, C.c_double(timestep), C.c_double(inv_tau_a), C.c_double(inv_tau_e), C.c_int(i_part),
                                      C.c_int(lo), C.c_int(hi)))

    def drift_bin(self, timestep, inv_tau_a, inv_tau_e, part_1, part_2):
        _chk(lib().ass_drift_bin(self._h, C.c_double(timestep), C.c_double(inv_tau_a), C.c_double(inv_tau_e), C.c_int(part_1), C.c_int(part_2)))

    def sum_mass(self, lo, hi):
        out = C.c_double()
        _chk(lib().ass_sum_mass(self._h, C.c_int(lo), C.c_int(hi), C.byref(out)))
        return out.value

    def upload_range(self, p, lo, hi, fields=F_ALL):
        """Rows [lo, hi) of the caller's full array `p` -> particles [lo, hi)."""
        base, stride, n = _rows(p)
        _chk(lib().ass_upload_range(self._h, base, stride, C.c_int(lo), C.c_int(hi), C.c_uint(fields)))

    def download_range(self, p, lo, hi, fields=F_ALL):
        base, stride, n = _rows(p)
        _chk(lib().ass_download_range(self._h, base, stride, C.c_int(lo), C.c_int(hi), C.c_uint(fields)))
        return p

    # -- spring build --
    def connect_springs_dist(self, max_dist, lo, hi, k, gamma, existing=None):
        ptr = C.c_void_p()
        n_out = C.c_int()
        if existing is not None and len(existing):
            ex, exp, exn = _springs(existing)
        else:
            ex, exp, exn = None, C.c_void_p(), C.c_int(0)
        _chk(lib().ass_connect_springs_dist(self._h, C.c_double(max_dist), C.c_int(lo), C.c_int(hi), C.c_double(k),
                                            C.c_double(gamma), exp, exn, C.byref(ptr), C.byref(n_out)))
        out = np.empty(n_out.value, dtype=SPRING_DTYPE)
        if n_out.value:
            C.memmove(out.ctypes.data, ptr, out.nbytes)
        if ptr.value:
            lib().ass_free(ptr)
        return out

    def mindist(self, lo=0, hi=None):
        hi = self.n if hi is None else hi
        out = C.c_double()
        _chk(lib().ass_mindist(self._h, C.c_int(lo), C.c_int(hi), C.byref(out)))
        return out.value

    # -- measurement --
    def timing(self, on=True):
        _chk(lib().ass_timing_enable(self._h, C.c_int(int(on))))

    def timing_reset(self):
        _chk(lib().ass_timing_reset(self._h))

    def timing_get(self):
        ms = (C.c_double * len(K_NAMES))()
        nl = (C.c_longlong * len(K_NAMES))()
        _chk(lib().ass_timing_get(self._h, ms, nl))
        return {k: (ms[i], nl[i]) for i, k in enumerate(K_NAMES)}

    def mark_begin(self):
        _chk(lib().ass_mark_begin(self._h))

    def mark_end(self):
        ms = C.c_double()
        _chk(lib().ass_mark_end(self._h, C.byref(ms)))
        return ms.value

    def lap_begin(self):
        _chk(lib().ass_lap_begin(self._h))

    def lap_end(self):
        _chk(lib().ass_lap_end(self._h))

    def lap_collect(self):
        """(sum of the laps in ms, number of laps) since the last collect; synchronises."""
        ms, n = C.c_double(), C.c_int()
        _chk(lib().ass_lap_collect(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def l2_flush(self):
        _chk(lib().ass_l2_flush(self._h))

    # -- one large body sharded by particle slab over ranks (include/ass_b200.h, "multi-GPU") --
    def shard_setup(self, rank, nranks, bounds):
        b = (C.c_int * (nranks + 1))(*[int(x) for x in bounds])
        _chk(lib().ass_shard_setup(self._h, C.c_int(rank), C.c_int(nranks), b))

    def shard_export(self):
        buf = (C.c_ubyte * 64)()
        _chk(lib().ass_shard_export(self._h, buf))
        return bytes(buf)

    def shard_import(self, peer_rank, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        _chk(lib().ass_shard_import(self._h, C.c_int(peer_rank), buf))

    def shard_import_local(self, peer_rank, peer):
        _chk(lib().ass_shard_import_local(self._h, C.c_int(peer_rank), peer._h))

    def shard_step(self, nsteps=1, pre_drift_last=False):
        """reb_step x nsteps for the own slab; collective, queued (sync() or a download waits for it)."""
        _chk(lib().ass_shard_step(self._h, C.c_int(nsteps), C.c_int(int(pre_drift_last))))

    def shard_gather(self):
        _chk(lib().ass_shard_gather(self._h))


from .launcher import slab_bounds  # noqa: E402,F401  (kept here for callers of the binding)


class Ensemble:
    """B independent small bodies stepped by one kernel, one CTA per body (``ass_ensemble``)."""

    def __init__(self, bodies, springs, params):
        """bodies: list of (n_b, >=10) float64 row arrays; springs: list of SPRING_DTYPE arrays with LOCAL indices;
        params: list of dicts of Params fields."""
        assert len(bodies) == len(springs) == len(params) and len(bodies) > 0
        self.nb = len(bodies)
        self.node_off = np.zeros(self.nb + 1, dtype=np.int32)
        self.spring_off = np.zeros(self.nb + 1, dtype=np.int32)
        self.node_off[1:] = np.cumsum([len(b) for b in bodies])
        self.spring_off[1:] = np.cumsum([len(s) for s in springs])
        rows = np.ascontiguousarray(np.vstack([np.asarray(b, dtype=np.float64)[:, :11] for b in bodies]))
        sp = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=SPRING_DTYPE) for s in springs]))
        P = (Params * self.nb)()
        for i, kw in enumerate(params):
            P[i].dt = 1e-3; P[i].G = 1.0; P[i].n_active = -1; P[i].exact_finish_time = 1; P[i].status = -1; P[i].additional_forces = 1
            for k, v in kw.items():
                setattr(P[i], k, v)
        self._h = C.c_void_p()
        _chk(lib().ass_ensemble_create(C.byref(self._h), C.c_int(self.nb), self.node_off.ctypes.data_as(C.c_void_p),
                                       self.spring_off.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.c_void_p),
                                       C.c_size_t(rows.strides[0]), sp.ctypes.data_as(C.c_void_p), P))
        self.total_nodes = int(self.node_off[-1])
        self.total_springs = int(self.spring_off[-1])

    def close(self):
        if self._h:
            lib().ass_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, nsteps=1, heartbeat=False, timed=False):
        ms = C.c_double()
        _chk(lib().ass_ensemble_step(self._h, C.c_int(nsteps), C.c_int(int(heartbeat)), C.byref(ms) if timed else None))
        return ms.value

    def download(self, fields=F_ALL):
        """Returns the list of per-body (n_b, 11) arrays (column 10, r, is not transferred)."""
        out = np.zeros((self.total_nodes, 11))
        _chk(lib().ass_ensemble_download(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.strides[0]), C.c_uint(fields)))
        return [out[self.node_off[b]:self.node_off[b + 1]] for b in range(self.nb)]

    def params(self):
        P = (Params * self.nb)()
        _chk(lib().ass_ensemble_get_params(self._h, P))
        return list(P)

    def sync(self):
        _chk(lib().ass_ensemble_sync(self._h))


def _gen(fn, min_dist, a, b, c, total_mass, existing):
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    n0 = 0 if existing is None else len(existing)
    ptr = C.c_void_p(None)
    cap = C.c_int(n0)
    if n0:
        existing = np.ascontiguousarray(existing, dtype=np.float64)
        assert existing.shape[1] == 11
        ptr = C.c_void_p(libc.malloc(C.c_size_t(existing.nbytes)))
        C.memmove(ptr, existing.ctypes.data, existing.nbytes)
    n = _chk(fn(C.byref(ptr), C.byref(cap), C.c_int(n0), C.c_double(min_dist), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(total_mass)))
    out = np.empty((n, 11), dtype=np.float64)
    if n:
        C.memmove(out.ctypes.data, ptr, out.nbytes)
    if ptr.value:
        lib().ass_free(ptr)
    return out


def srand(seed):
    """Seed libc's rand(): 'same seed' in the reference means srand() after reb_create_simulation."""
    C.CDLL(None).srand(C.c_uint(seed))


def rand_ellipsoid(min_dist, a, b, c, total_mass, existing=None):
    return _gen(lib().ass_rand_ellipsoid, min_dist, a, b, c, total_mass, existing)


def hcp_ellipsoid(min_dist, a, b, c, total_mass, existing=None):
    return _gen(lib().ass_hcp_ellipsoid, min_dist, a, b, c, total_mass, existing)


def cubic_ellipsoid(min_dist, a, b, c, total_mass, existing=None):
    return _gen(lib().ass_cubic_ellipsoid, min_dist, a, b, c, total_mass, existing)


def rand_rectangle(min_dist, x, y, z, total_mass, existing=None):
    return _gen(lib().ass_rand_rectangle, min_dist, x, y, z, total_mass, existing)


def _gen2(fn, args, existing):
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    n0 = 0 if existing is None else len(existing)
    ptr = C.c_void_p(None)
    cap = C.c_int(n0)
    if n0:
        existing = np.ascontiguousarray(existing, dtype=np.float64)
        assert existing.shape[1] == 11
        ptr = C.c_void_p(libc.malloc(C.c_size_t(existing.nbytes)))
        C.memmove(ptr, existing.ctypes.data, existing.nbytes)
    n = _chk(fn(C.byref(ptr), C.byref(cap), C.c_int(n0), *[C.c_double(a) for a in args]))
    out = np.empty((n, 11), dtype=np.float64)
    if n:
        C.memmove(out.ctypes.data, ptr, out.nbytes)
    if ptr.value:
        lib().ass_free(ptr)
    return out


def rand_cone(min_dist, radius, height, total_mass, existing=None):
    return _gen2(lib().ass_rand_cone, (min_dist, radius, height, total_mass), existing)


def rand_shape(vertices, min_dist, total_mass):
    """vertices: (V, 11) rows of the shape model; returns vertices followed by the new particles (as the reference does)."""
    return _gen2(lib().ass_rand_shape, (min_dist, total_mass), vertices)
