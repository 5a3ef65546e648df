"""The N > 1 host logic on CPU: two processes, torch.distributed with the gloo backend (what NCCL does on the GPU box).

Covers asteroid-spring-sims_b200/launcher.py -- member dealing, slab bounds, the window-handle exchange and the
two-barrier ordering of the sharded step -- with a stand-in for the simulation that records when each phase runs.
No kernel is involved (the CUDA side of the same protocol is tests/test_gpu_multi.py)."""
import json
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

WORKER = textwrap.dedent('''
    import json, os, sys, time
    sys.path.insert(0, %(root)r)
    import torch.distributed as dist
    from asteroid_spring_sims_b200 import launcher

    rank, world, local = launcher.dist_env()
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)

    class FakeSim:
        """Records (phase, step, t_start, t_end); `shard_exchange` checks the peers' handles arrived intact."""
        def __init__(self):
            self.log, self.step, self.peers = [], 0, {}
        def _run(self, phase, seconds):
            t0 = time.monotonic_ns(); time.sleep(seconds); self.log.append((phase, self.step, t0, time.monotonic_ns()))
        def shard_export(self):
            return bytes([rank]) * 64
        def shard_import(self, q, handle):
            assert handle == bytes([q]) * 64
            self.peers[q] = handle
        def shard_begin(self):
            self._run("begin", 0.02 * (rank + 1))          # ranks finish their drift at different times
        def shard_exchange(self):
            assert sorted(self.peers) == [q for q in range(world) if q != rank]
            self._run("exchange", 0.03 * (world - rank))
        def shard_gravity(self):
            self._run("gravity", 0.015 * (rank + 1))
        def shard_finish(self, pre_drift_next=False):
            self._run("finish:%%d" %% int(pre_drift_next), 0.01)
            self.step += 1

    sim = FakeSim()
    launcher.exchange_windows(dist, sim, rank, world)
    launcher.sharded_steps(dist, sim, 3)
    out = {
        "rank": rank, "world": world,
        "members": launcher.deal_members(11, world, rank),
        "max": launcher.reduce_max(dist, 1.5 + rank),
        "sum": launcher.reduce_sum(dist, 10.0 * (rank + 1)),
        "log": sim.log,
    }
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("RESULT " + json.dumps(gathered))
    dist.destroy_process_group()
''')


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_launcher_protocol_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(free_port()), str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    ranks = json.loads(line[len("RESULT "):])
    assert [r["rank"] for r in ranks] == [0, 1] and all(r["world"] == 2 for r in ranks)
    # ensemble members: every member exactly once, round-robin
    assert ranks[0]["members"] == [0, 2, 4, 6, 8, 10] and ranks[1]["members"] == [1, 3, 5, 7, 9]
    # reductions reach every rank
    assert all(r["max"] == 2.5 and r["sum"] == 30.0 for r in ranks)
    # ordering: in every step, nobody starts pulling before every slab has drifted, nobody writes partials before every
    # pull landed, and nobody collects partials before every rank finished writing them
    for step in range(3):
        ev = {(r["rank"], e[0].split(":")[0]): e for r in ranks for e in r["log"] if e[1] == step}
        begin_end = max(ev[(q, "begin")][3] for q in (0, 1))
        exch_start = min(ev[(q, "exchange")][2] for q in (0, 1))
        exch_end = max(ev[(q, "exchange")][3] for q in (0, 1))
        grav_start = min(ev[(q, "gravity")][2] for q in (0, 1))
        grav_end = max(ev[(q, "gravity")][3] for q in (0, 1))
        fin_start = min(ev[(q, "finish")][2] for q in (0, 1))
        assert begin_end <= exch_start and exch_end <= grav_start and grav_end <= fin_start, step
    # the last step of a batch does not pre-drift (the caller sees end-of-step positions)
    assert [e[0] for e in ranks[0]["log"] if e[0].startswith("finish")] == ["finish:1", "finish:1", "finish:0"]


def test_slab_bounds_and_dealing():
    import asteroid_spring_sims_b200 as A
    b = A.launcher.slab_bounds(945595, 8)
    assert b[0] == 0 and b[-1] == 945595 and all(x <= y for x, y in zip(b, b[1:])) and all(x % 256 == 0 for x in b[:-1])
    assert A.launcher.slab_bounds(10, 3, align=256) == [0, 10, 10, 10]
    assert sorted(sum((A.launcher.deal_members(4096, 8, r) for r in range(8)), [])) == list(range(4096))
    with pytest.raises(ValueError):
        A.launcher.deal_members(4, 2, 2)
