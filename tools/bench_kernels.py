#!/usr/bin/env python
"""Kernel-level timings used while tuning (not the graded bench): spring_forces and gravity at C3 size, cold and warm L2,
and the ensemble kernel on a few hundred C1 bodies.  Prints one line per measurement."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asteroid_spring_sims_b200 as A  # noqa: E402


def body(d, seed=12345, k=0.08, gamma=5.0, omega=(0.1, 0.0, 0.2)):
    A.srand(seed)
    p = A.rand_ellipsoid(d, 1.0, 0.7, 0.5, 1.0)
    n = len(p)
    s = A.Sim()
    s.upload(p)
    sp = s.connect_springs_dist(2.3 * d, 0, n, k, gamma)
    s.subtract_com(0, n); s.subtract_cov(0, n); s.spin_body(0, n, omega)
    s.download(p)
    s.set_springs(sp)
    return s, p, sp


def timeit(s, fn, reps, flush):
    fn(); s.sync()
    tot = 0.0
    for _ in range(reps):
        if flush:
            s.l2_flush()
        s.mark_begin()
        fn()
        tot += s.mark_end()
    return tot / reps


def main():
    A.init(0)
    what = sys.argv[1:] or ["springs", "gravity", "ensemble"]
    if "springs" in what or "gravity" in what:
        d = float(os.environ.get("D", "0.0215"))
        s, p, sp = body(d)
        n, ns = len(p), len(sp)
        s.set_params(dt=1e-3, G=1.0, softening=d / 100, gravity=1, additional_forces=1)
        alg = 32.0 * ns + 80.0 * n
        if "springs" in what:
            for flush in (True, False):
                ms = timeit(s, s.spring_forces, 20, flush)
                print("springs N=%d NS=%d %s L2: %.2f us  -> %.0f GB/s algorithmic (%.1f%% of 6538)" % (n, ns, "cold" if flush else "warm", ms * 1e3, alg / ms / 1e6, alg / ms / 1e6 / 65.38))
        if "gravity" in what:
            for which in (1, 2):
                s.set_gravity_kernel(which)
                ms = timeit(s, s.gravity_basic, 5, True)
                print("gravity kernel %d N=%d: %.3f ms -> %.3e pairs/s, %.2f TFLOP/s by the 20-flop convention" % (which, n, ms, n * n / ms * 1e3, 20.0 * n * n / ms / 1e9))
        s.close()
    if "ensemble" in what:
        nb = int(os.environ.get("BODIES", "592"))
        s, p, sp = body(0.10, omega=(0, 0, 0.3))
        s.close()
        ens = A.Ensemble([p] * nb, [sp] * nb, [dict(dt=1e-2, gravity=0, additional_forces=2)] * nb)
        ens.step(20)
        ms = ens.step(200, timed=True) / 200
        alg = (32.0 * len(sp) + 80.0 * len(p)) * nb
        print("ensemble %d bodies x N=%d NS=%d: %.1f us/step -> %.3e node-steps/s, %.0f GB/s algorithmic" % (nb, len(p), len(sp), ms * 1e3, nb * len(p) / ms * 1e3, alg / ms / 1e6))


if __name__ == "__main__":
    main()
