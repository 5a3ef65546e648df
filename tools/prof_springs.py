#!/usr/bin/env python
"""Gravity-free steps of a C3-size body (REB_GRAVITY_NONE + zero_accel + spring_forces: what every shipped module runs):
the fused springs + kick + drift kernel IS the step.  Prints the device time per step (L2 flushed before every step);
under ncu the same script is the capture target (tools/ncu_springs.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import asteroid_spring_sims_b200 as A  # noqa: E402
from bench_kernels import body  # noqa: E402


def main():
    A.init(0)
    d = float(os.environ.get("D", "0.0215"))
    steps = int(os.environ.get("STEPS", "20"))
    s, p, sp = body(d)
    n, ns = len(p), len(sp)
    s.set_params(dt=1e-3, G=1.0, softening=d / 100, gravity=0, additional_forces=2)
    s.step(3)
    s.sync()
    alg = 32.0 * ns + 80.0 * n
    if os.environ.get("UNTIMED"):
        s.step(steps)
        s.sync()
        print("N=%d NS=%d: %d untimed steps" % (n, ns, steps))
    else:
        ms = s.step_timed(steps) / steps
        print("gravity-free step N=%d NS=%d (cold L2): %.2f us/step -> %.0f GB/s algorithmic (%.1f%% of 6538)" % (n, ns, ms * 1e3, alg / ms / 1e6, alg / ms / 1e6 / 65.38))
        import time
        s.sync(); t0 = time.perf_counter(); s.step(200); s.sync(); t1 = time.perf_counter()
        print("  200 back-to-back steps, no flush (wall): %.2f us/step" % ((t1 - t0) / 200 * 1e6))
    s.close()


if __name__ == "__main__":
    main()
