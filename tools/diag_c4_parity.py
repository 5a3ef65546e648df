#!/usr/bin/env python
"""Diagnosis of the C4 sub-record's x, v difference between the sharded and the unsharded path (1.9e-13 after two steps, the
same number in every run of the round): both paths on ONE GPU (two ranks in one process), the first step's accelerations of
every particle compared, the worst ones checked against the CPU oracle."""
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
import asteroid_spring_sims_b200 as A  # noqa: E402


def main():
    A.init(0)
    name = os.environ.get("W", "C4")
    nranks = int(os.environ.get("RANKS", "2"))
    cfg = dict(bench.WORKLOADS[name])
    sim, p0, springs, _ = bench.build_body(A, cfg, cfg["omega"], with_point_mass=(name == "C4"))
    n = len(p0)
    n_perts = 1 if name == "C4" else 0
    prm = dict(t=0.0, dt=cfg["dt"], G=1.0, softening=cfg["d"] / 100.0, gravity=cfg["gravity"], n_active=-1, n_perts=n_perts, additional_forces=cfg["af"], heartbeat=0)
    sim.upload(p0); sim.set_params(**prm)
    sim.step(1); w1 = sim.download(p0.copy())
    sim.step(1); w2 = sim.download(p0.copy())
    sim.close()
    bounds = A.slab_bounds(n, nranks)
    sims = []
    for r in range(nranks):
        s = A.Sim(); s.upload(p0); s.set_springs(springs); s.set_params(**prm); s.shard_setup(r, nranks, bounds)
        sims.append(s)
    for r in range(nranks):
        for q in range(nranks):
            if q != r:
                sims[r].shard_import_local(q, sims[q])
    for s in sims: s.shard_step(1)
    for s in sims: s.shard_gather()
    for s in sims: s.sync()
    o1 = [s.download(p0.copy()) for s in sims]
    for s in sims: s.shard_step(1)
    for s in sims: s.shard_gather()
    for s in sims: s.sync()
    o2 = [s.download(p0.copy()) for s in sims]
    acc = np.zeros((n, 3))
    for r in range(nranks):
        acc[bounds[r]:bounds[r + 1]] = o1[r][bounds[r]:bounds[r + 1], 6:9]
    amax = np.abs(w1[:, 6:9]).max()
    da = np.abs(acc - w1[:, 6:9]).max(axis=1)
    worst = np.argsort(da)[-5:][::-1]
    print("max |a| %.4g; first-step acceleration, sharded vs unsharded: max %.3e (%.3e of max |a|) at particle %d" % (amax, da.max(), da.max() / amax, worst[0]))
    dv1 = np.abs(o1[0][:, :6] - w1[:, :6]).max(axis=1)
    dv2 = np.abs(o2[0][:, :6] - w2[:, :6]).max(axis=1)
    print("x, v after step 1: max %.3e at %d; after step 2: max %.3e at %d" % (dv1.max(), dv1.argmax(), dv2.max(), dv2.argmax()))
    for i in list(worst[:3]) + [int(dv2.argmax())]:
        want = bench.oracle_first_step_acc(p0, springs, cfg, n_perts, int(i), int(i) + 1)[0]
        print("particle %d: oracle a = %s\n   unsharded - oracle = %s\n   sharded   - oracle = %s   degree-ish: x,v diff step1 %.3e step2 %.3e" % (
            i, want, w1[i, 6:9] - want, acc[i] - want, dv1[i], dv2[i]))
    for s in sims: s.close()


if __name__ == "__main__":
    main()
