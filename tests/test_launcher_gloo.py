"""The N > 1 host logic on CPU: two processes, torch.distributed with the gloo backend (what NCCL does on the GPU box).

Covers asteroid-spring-sims_b200/launcher.py -- member dealing (round-robin and by cost), slab bounds, the window-handle
exchange, the reductions bench.py reports with, and that a sharded batch is ONE queued call per rank with no host
barrier in it -- with a stand-in for the simulation.  No kernel is involved (the device side of the protocol, the epoch
flags, is tests/test_gpu_multi.py)."""
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
        """Records the calls the launcher makes; `shard_step` checks the peers' handles arrived intact."""
        def __init__(self):
            self.calls, self.peers = [], {}
        def shard_export(self):
            return bytes([rank]) * 64
        def shard_import(self, q, handle):
            assert handle == bytes([q]) * 64
            self.peers[q] = handle
        def shard_step(self, nsteps, pre_drift_last=False):
            assert sorted(self.peers) == [q for q in range(world) if q != rank]
            self.calls.append(("step", nsteps, bool(pre_drift_last)))
        def shard_gather(self):
            self.calls.append(("gather",))

    class CountingDist:
        """torch.distributed with its collectives counted: the stepping calls must not touch it."""
        def __init__(self):
            self.n = 0
        def __getattr__(self, name):
            attr = getattr(dist, name)
            if name in ("barrier", "all_reduce", "all_gather_object", "broadcast"):
                def wrapped(*a, **k):
                    self.n += 1
                    return attr(*a, **k)
                return wrapped
            return attr

    cd = CountingDist()
    sim = FakeSim()
    launcher.exchange_windows(cd, sim, rank, world)
    before = cd.n
    time.sleep(0.05 * rank)                       # ranks arrive at different times: nothing on the host waits for that
    launcher.sharded_steps(cd, sim, 3, pre_drift=True)
    launcher.sharded_steps(cd, sim, 2)
    launcher.sharded_gather(cd, sim)
    collectives_in_stepping = cd.n - before
    costs = [13908 + 50 * (b %% 7) for b in range(41)]
    out = {
        "rank": rank, "world": world,
        "members": launcher.deal_members(11, world, rank),
        "by_cost": launcher.deal_members_by_cost(costs, world, rank),
        "max": launcher.reduce_max(dist, 1.5 + rank),
        "sum": launcher.reduce_sum(dist, 10.0 * (rank + 1)),
        "calls": sim.calls, "collectives_in_stepping": collectives_in_stepping,
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
    # by cost: a partition, and both ranks computed it from the same table
    assert sorted(ranks[0]["by_cost"] + ranks[1]["by_cost"]) == list(range(41))
    costs = [13908 + 50 * (b % 7) for b in range(41)]
    loads = [sum(costs[b] for b in r["by_cost"]) for r in ranks]
    assert abs(loads[0] - loads[1]) <= max(costs)
    # reductions reach every rank
    assert all(r["max"] == 2.5 and r["sum"] == 30.0 for r in ranks)
    # a batch of steps is one queued call per rank; the launcher puts no collective inside the stepping
    for r in ranks:
        assert r["calls"] == [["step", 3, True], ["step", 2, False], ["gather"]]
        assert r["collectives_in_stepping"] == 0


def test_slab_bounds_and_dealing():
    import asteroid_spring_sims_b200 as A
    b = A.launcher.slab_bounds(945595, 8)
    assert b[0] == 0 and b[-1] == 945595 and all(x <= y for x, y in zip(b, b[1:])) and all(x % 256 == 0 for x in b[:-1])
    assert A.launcher.slab_bounds(10, 3, align=256) == [0, 10, 10, 10]
    assert sorted(sum((A.launcher.deal_members(4096, 8, r) for r in range(8)), [])) == list(range(4096))
    with pytest.raises(ValueError):
        A.launcher.deal_members(4, 2, 2)
    # longest-processing-time first: a partition; equal costs come out as balanced counts; a heavy member is offset
    table = A.launcher.deal_members_by_cost([13908] * 4096, 8)
    assert sorted(sum(table, [])) == list(range(4096)) and all(len(t) == 512 for t in table)
    table = A.launcher.deal_members_by_cost([9, 5, 5, 5, 1, 1, 1, 1, 2], 3)
    assert sorted(sum(table, [])) == list(range(9))
    assert [sum([9, 5, 5, 5, 1, 1, 1, 1, 2][b] for b in t) for t in table] == [10, 10, 10]
    assert A.launcher.deal_members_by_cost([3, 1], 2, 1) == [1]
    with pytest.raises(ValueError):
        A.launcher.deal_members_by_cost([1, 2], 2, 5)
