"""One scene over several GPUs: replicated positions, sharded queries (first stage of SURVEY.md 8(e)).

Every rank holds the positions of the whole scene (12 bytes per particle: 200 MB at 2^24 particles)
and runs the bounds + hash + stable sort itself, which is deterministic, so all ranks agree on the
cell-sorted order bit for bit without any communication.  Rank r then OWNS the contiguous slice
[start, end) of that order (a slab of cell layers along dimension 0, because the cell hash is
row-major with dimension 0 most significant, common_funcs.h:114-118): it builds the neighbour rows of
its own particles only (ParticleCollision(..., query_range=...)) and evaluates ConvSP for them.

What has to cross NVLink is the per-particle FEATURE data of the layers: a rank needs data[j] for
every neighbour j of its particles.  Two exchange modes, same results:

  exchange="halo" (default): the neighbours of a slab lie in the slab itself plus the boundary cell
    layers of the adjacent slabs, i.e. -- the order being cell-sorted -- in ONE contiguous index range
    [lo, hi) around [start, end).  collide() reads lo/hi off the neighbour rows it just built, the ranks
    swap their four numbers, and every layer call then moves only the overlap rows with NCCL
    send/recv pairs (forward: rows to whoever references them; backward: their gradients back to the
    owner, who adds them).  Traffic ~ N^(2/3) per rank instead of N.
  exchange="allgather": one all-gather per layer input (forward) and one reduce-scatter per layer
    input gradient (backward); kept as the simple reference for the halo mode.
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


class HaloPlan(object):
    """Who needs which rows: sends / recvs are lists of (peer rank, first row, end row) in GLOBAL sorted
    indices; sends are rows this rank owns, recvs rows it references but does not own."""

    def __init__(self, N, start, end, sends, recvs, group):
        self.N, self.start, self.end, self.sends, self.recvs, self.group = N, start, end, sends, recvs, group

    def halo_rows(self):
        return sum(b - a for _, a, b in self.recvs)


def make_halo_plan(neighbors, N, start, end, group=None, keys=None, layer_cells=None):
    """The rows this rank must borrow / lend.

    With `keys` ([B, N] sorted integer cell keys, identical on every rank because the positions are
    replicated) and `layer_cells` (cells per layer of the leading grid dimension) every rank derives the
    referenced range of EVERY rank from the cell structure -- a slab's own cell layers plus one layer on each
    side, two binary searches per rank -- so the plan needs no communication.  Otherwise the range is read
    off the neighbour rows (a full pass over them) and the ranks swap their numbers with one tiny
    all-gather."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = neighbors.device
    if keys is not None and layer_cells is not None:
        B = keys.shape[0]
        owned = [owned_range(N, world, q) for q in range(world)]
        firsts = torch.tensor([min(s_q, N - 1) for s_q, _ in owned], device=dev)
        lasts = torch.tensor([max(e_q - 1, 0) for _, e_q in owned], device=dev)
        first_layer = torch.div(keys[:, firsts], layer_cells, rounding_mode="floor").min(0).values
        last_layer = torch.div(keys[:, lasts], layer_cells, rounding_mode="floor").max(0).values
        lo_key = ((first_layer - 1).clamp(min=0) * layer_cells).to(keys.dtype).expand(B, world).contiguous()
        hi_key = ((last_layer + 2) * layer_cells).to(keys.dtype).expand(B, world).contiguous()
        los = torch.searchsorted(keys, lo_key).min(0).values.tolist()   # the one host sync of the plan
        his = torch.searchsorted(keys, hi_key).max(0).values.tolist()
        table = [[min(int(lo), s_q), max(int(hi), e_q), s_q, e_q] if e_q > s_q else [s_q, s_q, s_q, e_q]
                 for lo, hi, (s_q, e_q) in zip(los, his, owned)]
    else:
        valid = neighbors >= 0
        big = torch.where(valid, neighbors, torch.full_like(neighbors, float(N)))
        lo = torch.minimum(big.min(), torch.tensor(float(start), device=dev))
        hi = torch.maximum(neighbors.max() + 1, torch.tensor(float(end), device=dev))
        mine = torch.stack([lo, hi, lo.new_tensor(float(start)), lo.new_tensor(float(end))]).to(torch.int64)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        table = [[int(v) for v in t.tolist()] for t in every]  # (lo, hi, start, end) per rank; one host sync
    lo_r, hi_r = table[rank][0], table[rank][1]
    sends, recvs = [], []
    for q in range(world):
        if q == rank:
            continue
        lo_q, hi_q, s_q, e_q = table[q]
        a, b = max(lo_r, s_q), min(hi_r, e_q)      # rows of q that I reference
        if a < b:
            recvs.append((q, a, b))
        a, b = max(lo_q, start), min(hi_q, end)    # rows of mine that q references
        if a < b:
            sends.append((q, a, b))
    return HaloPlan(N, start, end, sends, recvs, group)


def _swap(out_ops, in_ops, group):
    """out_ops: [(peer, tensor)] to send; in_ops: [(peer, tensor)] to fill.  Lower ranks' messages first on
    both sides, so the pairwise NCCL send/recv calls match up."""
    ops = []
    for peer, t in sorted(out_ops + in_ops, key=lambda pt: pt[0]):
        is_send = any(t is u for _, u in out_ops)
        ops.append(dist.P2POp(dist.isend if is_send else dist.irecv, t, peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class _HaloRows(torch.autograd.Function):
    """[B, n_r, C] owned rows -> [B, N, C] with the owned and the halo rows filled (other rows are never
    referenced by this rank's neighbour lists and stay uninitialised); backward returns the gradient of the
    owned rows, including what the ranks that used them send back."""

    @staticmethod
    def forward(ctx, x, plan):
        ctx.plan = plan
        B, n, C = x.shape
        full = x.new_empty(B, plan.N, C)
        full[:, plan.start:plan.end] = x
        outs = [(q, x[:, a - plan.start:b - plan.start].contiguous()) for q, a, b in plan.sends]
        ins = [(q, x.new_empty(B, b - a, C)) for q, a, b in plan.recvs]
        _swap(outs, ins, plan.group)
        for (q, a, b), (_, t) in zip(plan.recvs, ins):
            full[:, a:b] = t
        return full

    @staticmethod
    def backward(ctx, g):
        plan = ctx.plan
        B, _, C = g.shape
        own = g[:, plan.start:plan.end].clone()
        outs = [(q, g[:, a:b].contiguous()) for q, a, b in plan.recvs]        # gradients of rows I borrowed
        ins = [(q, g.new_empty(B, b - a, C)) for q, a, b in plan.sends]       # gradients of rows I lent
        _swap(outs, ins, plan.group)
        for (q, a, b), (_, t) in zip(plan.sends, ins):
            own[:, a - plan.start:b - plan.start] += t
        return own, None


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

    def __init__(self, coll, group=None, exchange="halo"):
        if exchange not in ("halo", "allgather"):
            raise ValueError("exchange must be 'halo' or 'allgather'")
        self.coll, self.group, self.exchange = coll, group, exchange
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.plan = None

    def collide(self, locs):
        N = locs.shape[1]
        self.N = N
        self.start, self.end = owned_range(N, self.world, self.rank)
        self.locs, self.idxs, self.neighbors = self.coll(locs, query_range=(self.start, self.end))
        self.plan = None
        if self.exchange == "halo":
            # sorted cell keys and grid of the call above (module scratch): a layer of the leading dimension
            # is prod(grid_dims[1:]) cells, and the hash is row-major with that dimension most significant
            B = locs.shape[0]
            keys = self.coll.cellIDs[:B].reshape(B, -1)[:, :N].view(torch.int32)
            gd = self.coll.last_grid_dims
            layer_cells = int(gd[:, 1:].prod(1).max().item()) if gd.shape[1] > 1 else 1
            same_grid = bool((gd == gd[:1]).all())
            self.plan = make_halo_plan(self.neighbors, N, self.start, self.end, self.group,
                                       keys=keys if same_grid else None, layer_cells=max(layer_cells, 1))
        return self.locs, self.idxs, self.neighbors

    def local_rows(self, x_sorted_full):
        return x_sorted_full[:, self.start:self.end]

    def convsp(self, conv, data_local, locs=None):
        """conv(locs, data, neighbors) for the owned particles; data_local are the owned rows."""
        locs = self.locs if locs is None else locs
        if self.plan is not None:
            data_full = _HaloRows.apply(data_local, self.plan)
        else:
            data_full = gather_particle_rows(data_local, self.N, self.group)
        return conv(locs, data_full, self.neighbors, qlocs=locs[:, self.start:self.end])
