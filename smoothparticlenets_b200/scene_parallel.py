"""One scene over several GPUs: replicated positions, sharded queries (first stage of SURVEY.md 8(e)).

Every rank holds the positions of the whole scene (12 bytes per particle: 200 MB at 2^24 particles)
and runs the bounds + hash + stable sort itself, which is deterministic, so all ranks agree on the
cell-sorted order bit for bit without any communication.  Rank r then OWNS the contiguous slice
[start, end) of that order (a slab of cell layers along dimension 0, because the cell hash is
row-major with dimension 0 most significant, common_funcs.h:114-118): it builds the neighbour rows of
its own particles only (ParticleCollision(..., query_range=...)) and evaluates ConvSP for them.

What has to cross NVLink is the per-particle FEATURE data of the layers: a rank needs data[j] for
every neighbour j of its particles.  This module exchanges it with one all-gather per layer input
(forward) and one reduce-scatter per layer input gradient (backward) over NCCL -- exact, simple, and
already cheap next to the compute (67 MB per channel at 2^24 particles against ~2.5 ms of ConvSP
per rank); restricting the exchange to the one boundary cell layer per neighbouring slab (the halo
exchange of SURVEY.md 8(e)) is the follow-up optimisation and does not change any result.
"""
import torch
import torch.distributed as dist


def owned_range(N, world_size, rank):
    """[start, end) of the cell-sorted particles owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(N, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class _AllGatherRows(torch.autograd.Function):
    """[B, n_r, C] per rank -> [B, N, C] on every rank; backward = reduce-scatter(sum) of the gradient."""

    @staticmethod
    def forward(ctx, x, sizes, group):
        ctx.sizes, ctx.group = sizes, group
        ctx.rank = dist.get_rank(group)
        parts = [x.new_empty(x.shape[0], n, x.shape[2]) for n in sizes]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, 1)

    @staticmethod
    def backward(ctx, g):
        chunks = [c.contiguous() for c in torch.split(g, ctx.sizes, 1)]
        out = torch.empty_like(chunks[ctx.rank])
        dist.reduce_scatter(out, chunks, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None, None


def gather_particle_rows(x_local, N, group=None):
    """All-gather the rows owned by every rank into the full [B, N, C] tensor (differentiable)."""
    world = dist.get_world_size(group)
    sizes = [owned_range(N, world, r)[1] - owned_range(N, world, r)[0] for r in range(world)]
    return _AllGatherRows.apply(x_local, sizes, group)


class ShardedScene(object):
    """Neighbour search and ConvSP for the slice of one scene owned by this rank.

        scene = ShardedScene(coll)                       # coll: ParticleCollision
        locs_sorted, idxs, nbrs = scene.collide(locs)    # locs: the WHOLE scene, same on every rank
        out_mine = scene.convsp(conv, data_mine)         # data_mine: rows [start, end) of the sorted order
    """

    def __init__(self, coll, group=None):
        self.coll, self.group = coll, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def collide(self, locs):
        N = locs.shape[1]
        self.N = N
        self.start, self.end = owned_range(N, self.world, self.rank)
        self.locs, self.idxs, self.neighbors = self.coll(locs, query_range=(self.start, self.end))
        return self.locs, self.idxs, self.neighbors

    def local_rows(self, x_sorted_full):
        return x_sorted_full[:, self.start:self.end]

    def convsp(self, conv, data_local, locs=None):
        """conv(locs, data, neighbors) for the owned particles; data_local are the owned rows."""
        locs = self.locs if locs is None else locs
        data_full = gather_particle_rows(data_local, self.N, self.group)
        return conv(locs, data_full, self.neighbors, qlocs=locs[:, self.start:self.end])
