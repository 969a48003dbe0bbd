"""Particle sharding of one ABC generation over the GPUs of a box.

One process per GPU (torchrun); particles of a chunk are split into contiguous ranges of
the GLOBAL particle index, every rank scores its range with the fused kernel (the Philox
streams are keyed on the global index, so results do not depend on the number of ranks),
and ONE all-gather of the per-particle distances per chunk lets every rank apply the same
accept / sequential-stop bookkeeping.  No other data-path collective exists: proposals are
drawn redundantly on every rank from the same host RNG seed.

The reference has no distributed layer; this module is new.

On GPUs the exchange happens INSIDE the library (itjl_api_multi.cu): `attach_library_comm` joins this rank's
context to an NCCL communicator (the 128-byte unique id is the only thing that travels through
`torch.distributed`, once), after which `itjl_abc_generation` / `itjl_pmc_weights` shard over the ranks and
all-gather device to device themselves - `gpu_sharded_scorer` is then nothing but the model's own `generation`.
A single process driving several GPUs needs none of this: `Context(devices=[0, 1, ...])`.
`make_sharded_scorer` (host all-gather through `torch.distributed`) remains for the gloo CPU tests of the
host logic.
"""
import numpy as np


def shard_bounds(n, world_size, rank):
    """Contiguous range [lo, hi) of `rank` when n items are split over world_size ranks."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_distances(local, counts, group=None, device=None):
    """All-gather variable-length float64 shards (padded to the largest shard)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = max(counts)
    dev = device if device is not None else "cpu"
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=dev)
    if local.size:
        buf[: local.size] = torch.as_tensor(local, dtype=torch.float64).to(dev)
    out = torch.empty(world * width, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, width)
    return np.concatenate([out[r, : counts[r]] for r in range(world)])


def stop_and_accept(dist_all, epsilon, needed):
    """Host restatement of K3 for the gathered vector (used by the gloo CPU tests and as the
    cross-check of itjl_abc_accept): returns (isaccepted, n_total, n_accepted)."""
    acc = dist_all <= epsilon
    if needed > 0:
        c = np.cumsum(acc)
        hit = np.nonzero(c == needed)[0]
        if hit.size and acc[hit[0]]:
            n_total = int(hit[0]) + 1
            return acc, n_total, int(needed)
    return acc, int(dist_all.size), int(acc.sum())


def make_sharded_scorer(local_scorer, accept_fn=None, group=None, device=None):
    """Build the `scorer` callback of abc.basic_abc for a multi-rank run.

    local_scorer(theta_slice, generation, particle_offset) -> distances of the slice
    accept_fn(dist_all, epsilon, needed) -> (isaccepted, n_total, n_accepted)
        (default: stop_and_accept; on GPUs pass the itjl_abc_accept wrapper)
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    accept = accept_fn or stop_and_accept

    def scorer(theta, epsilon, needed, generation, particle_offset):
        n = theta.shape[0]
        bounds = [shard_bounds(n, world, r) for r in range(world)]
        lo, hi = bounds[rank]
        local = (local_scorer(theta[lo:hi], generation, particle_offset + lo)
                 if hi > lo else np.zeros(0))
        d = all_gather_distances(np.asarray(local, dtype=np.float64),
                                 [b[1] - b[0] for b in bounds], group=group, device=device)
        isacc, n_total, n_acc = accept(d, epsilon, needed)
        return d, isacc, n_total, n_acc

    return scorer


def attach_library_comm(ctx, group=None):
    """Make `ctx` (this rank's single-GPU context) a member of an in-library NCCL communicator spanning the
    ranks of `group`: rank 0 creates the unique id, `torch.distributed` broadcasts its 128 bytes, every rank
    calls itjl_ctx_attach_comm (collective).  Idempotent."""
    import torch.distributed as dist

    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if ctx.world()[1] == world and world > 1:
        return ctx
    box = [_lib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.attach_comm(rank, world, box[0])
    return ctx


def gpu_sharded_scorer(model, ctx=None, seed=0, group=None):
    """`scorer` callback of abc.basic_abc for a multi-rank GPU run: the library shards the chunk over the
    communicator of `ctx`, all-gathers the distances over NVLink and applies accept / stop on the device;
    every rank gets the full result (COLLECTIVE: every rank must call it with the same arguments)."""
    from . import _lib
    ctx = ctx or _lib.default_context()
    attach_library_comm(ctx, group)
    gm = model.gpu_model(ctx)

    def scorer(theta, epsilon, needed, generation, particle_offset):
        return gm.generation(theta, epsilon, needed, seed=seed, generation=generation,
                             particle_offset=particle_offset)

    return scorer
