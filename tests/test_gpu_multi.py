"""Ensemble (independent bodies, one CTA each) and slab sharding of one large body, checked on one GPU.

Ensemble members must reproduce what the single-body path gives the same body (bit for bit without gravity: same
arithmetic, same lane order), which the other tests already tie to the reference.  The sharded path is run as a
P-rank protocol inside one process (P ass_sim objects exchanging through same-process peer pointers, all on one GPU):
every rank's steps are queued up front by the one host thread and the ranks then meet through the same device-side
epoch flags as real ranks do -- the host cannot order anything, so a missing wait shows up as a wrong result.  The
multi-process CUDA-IPC path is exercised by test_ipc_two_ranks when two GPUs are visible and by bench.py --gpus N under
torchrun (which checks parity inside the run).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, same_bits
from oracle.bindings import make_sim

pytestmark = pytest.mark.gpu

A_, B_, C_ = 1.0, 0.7, 0.5


def make_body(gpu, d, seed, omega, k=0.08, gamma=5.0):
    gpu.srand(seed)
    p = gpu.rand_ellipsoid(d, A_, B_, C_, 1.0)
    n = len(p)
    with gpu.Sim() as s:
        s.upload(p)
        sp = s.connect_springs_dist(2.3 * d, 0, n, k, gamma)
        s.subtract_com(0, n); s.subtract_cov(0, n); s.spin_body(0, n, omega)
        s.download(p)
    return p, sp


@pytest.mark.parametrize("item_steps", [None, "1", "3"])
def test_ensemble_matches_single_body_path(gpu, oracle, monkeypatch, item_steps):
    """item_steps: how many consecutive steps of a member one work item covers (k_ensemble's persistent CTAs pull
    (member, batch) items; None = the library's choice).  A member's state crosses global memory between batches,
    possibly from one SM to another: the result must not depend on it."""
    if item_steps is None:
        monkeypatch.delenv("ASS_ENS_ITEM_STEPS", raising=False)
    else:
        monkeypatch.setenv("ASS_ENS_ITEM_STEPS", item_steps)
    bodies, springs, params = [], [], []
    for b, (d, seed, wz, k) in enumerate([(0.15, 12345, 0.2, 0.08), (0.2, 7, 0.5, 0.02), (0.15, 99, 0.9, 0.16), (0.12, 5, 0.1, 0.04)]):
        p, sp = make_body(gpu, d, seed, (0.0, 0.0, wz), k=k)
        bodies.append(p); springs.append(sp)
        params.append(dict(dt=1e-2, gravity=0, additional_forces=2, heartbeat=0))
    # no gravity, zero_accel + springs (asteroid_spin callbacks): bitwise equal to ass_step on each body
    with gpu.Ensemble(bodies, springs, params) as e:
        e.step(7)
        got = e.download()
        prm = e.params()
    for b in range(len(bodies)):
        with gpu.Sim() as s:
            s.upload(bodies[b]); s.set_springs(springs[b])
            s.set_params(dt=1e-2, gravity=0, additional_forces=2, heartbeat=0)
            s.step(7)
            want = s.download(bodies[b].copy())
            assert prm[b].t == s.params.t
        assert same_bits(got[b][:, :10], want[:, :10])
        # and therefore equal to the reference within the usual bars
        ref = bodies[b].copy()
        sim = make_sim(dt=1e-2, gravity_basic=0, zero_first=1)
        oracle.step(ref, springs[b], sim, 7)
        assert np.abs(got[b][:, 6:9] - ref[:, 6:9]).max() <= 1e-11 * np.abs(ref[:, 6:9]).max()
        assert np.allclose(got[b][:, :6], ref[:, :6], rtol=0, atol=1e-13)


@pytest.mark.parametrize("item_steps", [None, "2"])
def test_ensemble_gravity_heartbeat_and_accumulating_modes(gpu, oracle, monkeypatch, item_steps):
    if item_steps is None:
        monkeypatch.delenv("ASS_ENS_ITEM_STEPS", raising=False)
    else:
        monkeypatch.setenv("ASS_ENS_ITEM_STEPS", item_steps)
    p, sp = make_body(gpu, 0.15, 12345, (0.1, 0.0, 0.2))
    cases = [dict(dt=2e-3, gravity=1, additional_forces=1, heartbeat=1, softening=1.5e-3),
             dict(dt=1e-3, gravity=0, additional_forces=1, heartbeat=0),      # accelerations accumulate
             dict(dt=2e-3, gravity=1, additional_forces=2, heartbeat=0, softening=1.5e-3),
             dict(dt=2e-3, gravity=0, additional_forces=0, heartbeat=1)]
    start = p.copy(); start[:, 6:9] = 0.03125
    with gpu.Ensemble([start] * len(cases), [sp] * len(cases), cases) as e:
        e.step(5, heartbeat=True)
        got = e.download()
    for b, c in enumerate(cases):
        ref = start.copy()
        sim = make_sim(dt=c["dt"], softening=c.get("softening", 0.0), gravity_basic=c["gravity"], zero_first=(c["additional_forces"] == 2),
                       center_heartbeat=c["heartbeat"])
        springs_b = sp if c["additional_forces"] else sp[:0]
        oracle.step(ref, springs_b, sim, 5, heartbeat=True)
        amax = max(np.abs(ref[:, 6:9]).max(), 1e-300)
        assert np.abs(got[b][:, 6:9] - ref[:, 6:9]).max() <= 1e-11 * amax, c
        assert np.allclose(got[b][:, :6], ref[:, :6], rtol=0, atol=1e-13), c


def test_ensemble_errors(gpu):
    p, sp = make_body(gpu, 0.2, 3, (0, 0, 0.2))
    bad = sp.copy(); bad["particle_2"][0] = len(p) + 5
    with pytest.raises(gpu.AssError):
        gpu.Ensemble([p], [bad], [dict()])
    big = np.zeros((6000, 11)); big[:, 9] = 1.0   # 56 B of shared memory per node: the limit is ~4000
    with pytest.raises(gpu.AssError) as e:
        gpu.Ensemble([big], [sp[:0]], [dict()])
    assert "shared memory" in str(e.value)


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 5, 8])
def test_sharded_body_in_process(gpu, oracle, body010, nranks):
    """C4's protocol at small size: every 'rank' is an ass_sim of this process; result = the unsharded path's."""
    g = body010
    d = float(g["d"][0])
    body = g["start"].copy()
    n_body = len(body)
    pert = np.zeros((1, 11)); pert[0, :3] = (7.0, 0.0, 0.0); pert[0, 4] = 0.4; pert[0, 9] = 1.0   # Mars-like point mass
    p = np.vstack([body, pert])
    n = len(p)
    sp = g["springs"]
    prm = dict(t=0.0, dt=2e-3, G=1.0, softening=d / 100.0, gravity=1, additional_forces=1, n_perts=1)
    bounds = gpu.slab_bounds(n, nranks, align=256)
    sims = []
    for r in range(nranks):
        s = gpu.Sim(); s.upload(p); s.set_springs(sp); s.set_params(**prm)
        s.shard_setup(r, nranks, bounds)
        sims.append(s)
    for r in range(nranks):
        for q in range(nranks):
            if q != r:
                sims[r].shard_import_local(q, sims[q])
    nsteps = 4
    # two batches, the first one ending pre-drifted: everything is queued before any rank has run a single kernel
    for s in sims: s.shard_step(3, pre_drift_last=True)
    for s in sims: s.shard_step(nsteps - 3)
    for s in sims: s.shard_gather()
    for s in sims: s.sync()
    # reference: the oracle, and the unsharded GPU path
    ref = p.copy()
    sim = make_sim(dt=2e-3, softening=d / 100.0, gravity_basic=1, zero_first=0, n_perts=1)
    oracle.step(ref, sp, sim, nsteps)
    with gpu.Sim() as u:
        u.upload(p); u.set_springs(sp); u.set_params(**prm)
        u.step(nsteps)
        unsharded = u.download(p.copy())
    for r, s in enumerate(sims):
        out = s.download(p.copy())
        lo, hi = bounds[r], bounds[r + 1]
        # every rank sees every particle's x, v after the gather; accelerations only for its own slab
        assert np.allclose(out[:, :6], ref[:, :6], rtol=0, atol=1e-13)
        assert np.allclose(out[:, :6], unsharded[:, :6], rtol=0, atol=1e-14)
        if hi > lo:
            amax = np.abs(ref[:, 6:9]).max()
            assert np.abs(out[lo:hi, 6:9] - ref[lo:hi, 6:9]).max() <= 1e-11 * amax
        assert abs(s.params.t - nsteps * 2e-3) < 1e-15
        s.close()


def test_sharded_gravity_streams_and_serial_order_agree(gpu, body010, monkeypatch):
    """The rectangle / antipodal pair kernels of a rank run on streams of their own next to the triangle (shard.cu);
    ASS_SHARD_SERIAL_GRAVITY=1 queues them one after the other on the main stream.  Same scratch, same fixed-order reductions:
    the gathered body must come out bit for bit the same (four ranks in this process, with a point mass)."""
    g = body010
    d = float(g["d"][0])
    pert = np.zeros((1, 11)); pert[0, :3] = (7.0, 0.0, 0.0); pert[0, 4] = 0.4; pert[0, 9] = 1.0
    p = np.vstack([g["start"].copy(), pert])
    n = len(p)
    prm = dict(t=0.0, dt=2e-3, G=1.0, softening=d / 100.0, gravity=1, additional_forces=1, n_perts=1)
    nranks = 4
    bounds = gpu.slab_bounds(n, nranks, align=256)
    res = []
    for serial in ("1", None):
        if serial is None:
            monkeypatch.delenv("ASS_SHARD_SERIAL_GRAVITY", raising=False)
        else:
            monkeypatch.setenv("ASS_SHARD_SERIAL_GRAVITY", serial)
        sims = []
        for r in range(nranks):
            s = gpu.Sim(); s.upload(p); s.set_springs(g["springs"]); s.set_params(**prm); s.shard_setup(r, nranks, bounds)
            sims.append(s)
        for r in range(nranks):
            for q in range(nranks):
                if q != r:
                    sims[r].shard_import_local(q, sims[q])
        for s in sims: s.shard_step(3)
        for s in sims: s.shard_gather()
        for s in sims: s.sync()
        outs = [s.download(p.copy()) for s in sims]
        for s in sims: s.close()
        res.append(outs)
    for r in range(nranks):
        lo, hi = bounds[r], bounds[r + 1]
        assert same_bits(res[0][r][:, :6], res[1][r][:, :6])
        assert same_bits(res[0][r][lo:hi, 6:9], res[1][r][lo:hi, 6:9])


IPC_WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import asteroid_spring_sims_b200 as A
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
one_gpu = os.environ.get("ASS_TEST_ONE_GPU") == "1"    # both ranks on GPU 0: two processes, two contexts, time-sliced
dev = 0 if one_gpu else rank
torch.cuda.set_device(dev)
if one_gpu:
    dist.init_process_group("gloo")                    # (NCCL refuses two ranks on one device; only handles are swapped)
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
A.init(dev)
g = np.load(os.path.join(%r, "tests", "golden", "c1_d010.npz"))
p = g["start"].copy(); sp = g["springs"]; n = len(p); d = float(g["d"][0])
prm = dict(t=0.0, dt=2e-3, G=1.0, softening=d / 100.0, gravity=1, additional_forces=1)
bounds = A.slab_bounds(n, world)
s = A.Sim(); s.upload(p); s.set_springs(sp); s.set_params(**prm); s.shard_setup(rank, world, bounds)
handles = [None] * world
dist.all_gather_object(handles, s.shard_export())
for q in range(world):
    if q != rank: s.shard_import(q, handles[q])
if rank == 1:
    for _ in range(40): s.l2_flush()      # one rank starts late: the others must wait on the device, not run ahead
s.shard_step(2, pre_drift_last=True)
if rank == 0:
    for _ in range(40): s.l2_flush()      # ... and the other one falls behind between batches
s.shard_step(2)
s.shard_gather(); s.sync()
out = s.download(p.copy())
u = A.Sim(); u.upload(p); u.set_springs(sp); u.set_params(**prm); u.step(4); want = u.download(p.copy())
lo, hi = bounds[rank], bounds[rank + 1]
assert np.allclose(out[:, :6], want[:, :6], rtol=0, atol=1e-14)
assert np.abs(out[lo:hi, 6:9] - want[lo:hi, 6:9]).max() <= 1e-12 * np.abs(want[:, 6:9]).max()
print("rank", rank, "ok")
dist.barrier(); s.close(); dist.destroy_process_group()
"""


def test_ipc_two_ranks(gpu, tmp_path):
    """The real thing: two processes, two GPUs, CUDA IPC windows, device-side flags; the host never synchronises the ranks
    inside the run (torch.distributed only swaps the window handles)."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(IPC_WORKER % (ROOT, ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("ok") >= 2, r.stdout[-3000:]   # both ranks print "rank <r> ok" (their lines may interleave)


def test_ipc_two_ranks_on_one_gpu(gpu, tmp_path):
    """The same two processes on ONE GPU (what a single-GPU test box can run): separate CUDA contexts, the windows mapped
    through CUDA IPC, the ranks meeting through the device-side flags while the driver time-slices the two contexts -- a
    missing flag, a stale window or a wrong epoch shows here exactly as it would across GPUs (only slower: a waiting
    kernel spins through its time slice)."""
    script = tmp_path / "worker.py"
    script.write_text(IPC_WORKER % (ROOT, ROOT))
    env = dict(os.environ, ASS_TEST_ONE_GPU="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29543", str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    if r.returncode != 0 and ("busy or unavailable" in r.stdout or "exclusive" in r.stdout.lower()):
        pytest.skip("this GPU does not admit a second process (exclusive compute mode)")
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("ok") >= 2, r.stdout[-3000:]


def test_ensemble_more_members_than_sms(gpu):
    """More members than SMs: items of ten steps keep every SM busy; members must still come out identical to the same
    member stepped alone (here: 200 copies of one body => every copy bit-identical to copy 0, and to a 1-member run)."""
    p, sp = make_body(gpu, 0.2, 3, (0.0, 0.0, 0.3))
    prm = dict(dt=1e-2, gravity=0, additional_forces=2, heartbeat=0)
    with gpu.Ensemble([p], [sp], [prm]) as e:
        e.step(23)
        want = e.download()[0]
    nb = 200
    with gpu.Ensemble([p] * nb, [sp] * nb, [prm] * nb) as e:
        e.step(23)
        got = e.download()
    for b in range(nb):
        assert same_bits(got[b][:, :10], want[:, :10]), b
