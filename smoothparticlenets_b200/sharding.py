"""Multi-GPU host logic: the path shards over INDEPENDENT SCENES.

One process per GPU (torch.distributed); every kernel of the path is independent per batch element
(`b*N` strides everywhere in the reference, e.g. common_funcs.h:470-473), so the forward and the
data/position gradients need no communication at all.  The only cross-rank quantity is the
gradient of SHARED parameters (ConvSP / ConvSDF weight and bias when with_params=True): it is the
sum over all scenes, i.e. one all-reduce(sum) per parameter.  The fluid-simulation layers have
with_params=False and communicate nothing.
"""
import torch
import torch.distributed as dist


def scene_shard(num_scenes, world_size, rank):
    """Contiguous, balanced slice of scene indices owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, rem = divmod(num_scenes, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def allreduce_parameter_grads(modules, group=None):
    """Sum the gradients of all parameters of `modules` over the process group, in place.  Call after
    backward when a batch of scenes was split over ranks; a no-op without an initialised group."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    grads = [p.grad for m in modules for p in m.parameters() if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])  # one bucket: these tensors are tiny (O*C*ncells)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return len(grads)
