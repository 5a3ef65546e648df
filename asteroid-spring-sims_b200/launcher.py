"""Host-side launcher logic for more than one GPU: one process per GPU, torch.distributed for the plumbing only.

Two ways the path shards (include/ass_b200.h "multi-GPU"):
  * ensemble members are independent: they are dealt to the ranks by cost (longest-processing-time first), no
    data-path collective;
  * one large body is cut into contiguous particle slabs.  The host's part is set-up only: swap the CUDA-IPC window
    handles once (all_gather_object) and map them.  A step is `sim.shard_step(n)`: it queues whole steps on the rank's
    streams and the ranks meet on the device through epoch flags in the peer-mapped windows (csrc/shard.cu) -- no
    host barrier, no stream synchronisation inside the loop.
Everything here is backend-agnostic (NCCL on the GPU box, gloo in the CPU tests: tests/test_launcher_gloo.py) and
touches the simulation only through the methods named above, so the host logic can be tested without a GPU.
"""
import os


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) for a plain single process."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def deal_members(total, world, rank):
    """Indices of the ensemble members rank `rank` owns: round-robin, so sweeps that vary slowly with the member index
    spread evenly over the GPUs.  Right when all members cost the same; see deal_members_by_cost otherwise."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, total, world))


def deal_members_by_cost(costs, world, rank=None):
    """Longest-processing-time-first assignment of ensemble members to ranks: members sorted by decreasing cost (ties: by
    index), each given to the rank with the least work so far (ties: fewest members, then lowest rank).  A member's cost
    is its spring count (the spring phase is the step).  Deterministic, so every rank computes the same table.
    Returns rank `rank`'s member indices in ascending order, or the whole table (list per rank) when rank is None."""
    if world < 1 or (rank is not None and not (0 <= rank < world)):
        raise ValueError("rank %r outside world of %d" % (rank, world))
    import heapq
    heap = [(0.0, 0, r) for r in range(world)]
    table = [[] for _ in range(world)]
    for b in sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i)):
        load, cnt, r = heapq.heappop(heap)
        table[r].append(b)
        heapq.heappush(heap, (load + float(costs[b]), cnt + 1, r))
    for t in table:
        t.sort()
    return table if rank is None else table[rank]


def slab_bounds(n, nranks, align=256):
    """Contiguous particle slabs of near-equal size, boundaries aligned to the gravity tile (256 sources)."""
    per = -(-n // nranks)
    per = -(-per // align) * align
    return [min(n, r * per) for r in range(nranks)] + [n]


def _tensor(x, device):
    import torch
    return torch.tensor([x], dtype=torch.float64, device=device)


def reduce_max(dist, x, device=None):
    """max over ranks of a Python float (every rank gets it); dist=None: single process."""
    if dist is None:
        return x
    t = _tensor(x, device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, x, device=None):
    if dist is None:
        return x
    t = _tensor(x, device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def exchange_windows(dist, sim, rank, world):
    """Swap the slab windows' IPC handles and map every peer's window (ass_shard_export / ass_shard_import)."""
    handles = [None] * world
    dist.all_gather_object(handles, sim.shard_export())
    for q in range(world):
        if q != rank:
            sim.shard_import(q, handles[q])
    dist.barrier()
    return handles


def sharded_steps(dist, sim, nsteps, pre_drift=False):
    """reb_step x nsteps of a slab-sharded body: ONE queued call per rank, no host barrier (`dist` is unused and kept
    for callers that pass it).  pre_drift leaves the positions half a drift ahead for a following batch."""
    sim.shard_step(nsteps, pre_drift_last=pre_drift)


def sharded_gather(dist, sim):
    """Every slab's x, v on every rank (before a full download); collective, device-synchronised."""
    sim.shard_gather()
