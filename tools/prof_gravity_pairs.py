#!/usr/bin/env python
"""Pair-once gravity on the pieces a rank of a sharded C3 body launches per step (8 ranks: own triangle, the rectangle against
the run of the next three slabs, half of the antipodal rectangle), timed on ONE GPU against the DFMA peak measured in the same
run; and the whole body / a C2-size body for comparison.  Prints one line per piece."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import asteroid_spring_sims_b200 as A  # noqa: E402


def main():
    A.init(0)
    peak, _ = A.probe_fp64_peak(1 << 13)
    print("DFMA peak measured: %.2f TFLOP/s" % peak)
    for d in (0.0215, 0.0474):
        A.srand(12345)
        p = A.rand_ellipsoid(d, 1.0, 0.7, 0.5, 1.0)
        n = len(p)
        s = A.Sim()
        s.upload(p)
        s.set_params(dt=1e-3, G=1.0, softening=d / 100, gravity=1, additional_forces=0)

        def piece(name, rlo, rhi, slo, shi, pairs):
            ms = s.probe_gravity_pair(rlo, rhi, slo, shi, reps=5)
            tf = 40.0 * pairs / (ms * 1e-3) / 1e12   # 20 flop per ORDERED pair = 40 per unordered pair evaluated once
            print("N=%6d %-34s %9.1f us  %6.2f TFLOP/s = %5.1f %% of the DFMA peak" % (n, name, ms * 1e3, tf, 100 * tf / peak))
            return ms

        piece("whole body (triangle)", 0, n, 0, n, n * (n - 1) / 2)
        if n > 50000:
            w = 12288
            t = piece("slab 0 triangle (12288)", 0, w, 0, w, w * (w - 1) / 2)
            r = piece("slab 0 x slabs 1-3 (run of 36864)", 0, w, w, 4 * w, w * 3.0 * w)
            a = piece("half of slab 0 x slab 4", 0, w // 2, 4 * w, 5 * w, (w // 2) * 1.0 * w)
            tot_pairs = w * (w - 1) / 2 + 3.0 * w * w + 0.5 * w * w
            print("  a rank's gravity per step: %.1f us = %.1f %% of the DFMA peak" % ((t + r + a) * 1e3, 100 * 40.0 * tot_pairs / ((t + r + a) * 1e-3) / 1e12 / peak))
            pieces = [(0, w, 0, w), (0, w, w, 4 * w), (0, w // 2, 4 * w, 5 * w)]
            for conc in (0, 1):
                ms = s.probe_gravity_pieces(pieces, conc, reps=5)
                print("  the three pieces in one timed region, %s: %.1f us = %.1f %% of the DFMA peak" % ("pair kernels on three streams" if conc else "one after the other", ms * 1e3, 100 * 40.0 * tot_pairs / (ms * 1e-3) / 1e12 / peak))
            for name, order in (("rectangle first", (1, 0, 2)), ("rectangle, antipode, triangle", (1, 2, 0))):
                ms = s.probe_gravity_pieces([pieces[i] for i in order], 1, reps=5)
                print("  three streams, %s: %.1f us = %.1f %%" % (name, ms * 1e3, 100 * 40.0 * tot_pairs / (ms * 1e-3) / 1e12 / peak))
        s.close()


if __name__ == "__main__":
    main()
