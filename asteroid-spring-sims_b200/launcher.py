"""Host-side launcher logic for more than one GPU: one process per GPU, torch.distributed for the plumbing only.

Two ways the path shards (include/ass_b200.h "multi-GPU"):
  * ensemble members are independent: they are dealt round-robin to the ranks, no data-path collective;
  * one large body is cut into contiguous particle slabs; per step the ranks run
        shard_begin -> barrier -> shard_exchange -> barrier -> shard_gravity -> barrier -> shard_finish
    where the exchanges are pulls through CUDA-IPC windows (handles swapped once with all_gather_object): the peers'
    slabs of x and v, and the partial accelerations the peers computed for this rank's slab.
Everything here is backend-agnostic (NCCL on the GPU box, gloo in the CPU tests: tests/test_launcher_gloo.py) and
touches the simulation only through the methods named above, so the ordering rules can be tested without a GPU.
"""
import os


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) for a plain single process."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def deal_members(total, world, rank):
    """Indices of the ensemble members rank `rank` owns: round-robin, so sweeps that vary slowly with the member index
    spread evenly over the GPUs."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, total, world))


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


def sharded_steps(dist, sim, nsteps, pre_drift=True):
    """reb_step x nsteps of a slab-sharded body.  Barrier 1: every slab carries its drifted positions before anyone
    pulls.  Barrier 2: every pull has landed before anyone overwrites anything a peer reads.  Barrier 3: every rank's
    partial accelerations for the other slabs are complete before their owners pull them."""
    for s in range(nsteps):
        sim.shard_begin()
        dist.barrier()
        sim.shard_exchange()
        dist.barrier()
        sim.shard_gravity()
        dist.barrier()
        sim.shard_finish(pre_drift_next=(pre_drift and s < nsteps - 1))
