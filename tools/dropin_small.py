#!/usr/bin/env python
"""The drop-in on a SMALL gravity-free body (BASELINE C1: asteroid_spin as shipped, ~1k nodes): tests/dropin/problem.cpp linked
against the reference and against the drop-in, per-step wall time of reb_integrate in coherent and resident mode.
At this size a step is launch- and call-bound, which is what the numbers show."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import asteroid_spring_sims_b200 as A  # noqa: E402


def main():
    A.init(0)
    name = sys.argv[1] if len(sys.argv) > 1 else "C1"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    cfg = dict(bench.WORKLOADS[name])
    sim, p, sp, _ = bench.build_body(A, cfg, cfg["omega"])
    sim.close()
    out = bench.run_dropin_e2e(p, sp, cfg, steps_b200=steps, label=name)
    for k in ("reference", "b200_same_call", "b200_coherent", "b200_resident"):
        if isinstance(out.get(k), dict):
            out[k].pop("res_row", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
