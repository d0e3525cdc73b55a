"""ParticleCollision and ReorderData on libspnb (sm_100a).

Drop-in for python/SmoothParticleNets/ParticleCollision.py of the reference: same constructors,
registered buffer names (cellIDs, cellStarts, cellEnds, cuda_buffer), return tuples and autograd
semantics (only the reorder pass-through carries gradients, ParticleCollision.py:238-243,
277-283, 303-314).  The whole forward is stream-ordered: bounds, cell hash, stable sort, reorder,
cell table and neighbour lists run without a host synchronisation.
"""
import numbers  # noqa: F401
import os
import weakref

import torch

from . import _native as nat
from . import error_checking as ec
from . import sidecar
from .convsp import nat_max_dim


class ReorderData(torch.nn.Module):
    """reverse=False: ret = input[idxs];  reverse=True: ret[idxs] = input
    (reference ParticleCollision.py:13-58)."""

    def __init__(self, reverse=False):
        super(ReorderData, self).__init__()
        self.reverse = (1 if reverse else 0)

    def forward(self, idxs, locs, data=None):
        batch_size = locs.size()[0]
        N = locs.size()[1]
        ec.check_tensor_dims(locs, "locs", (batch_size, N, -1))
        ec.check_tensor_dims(idxs, "idxs", (batch_size, N))
        if data is not None:
            ec.check_tensor_dims(data, "data", (batch_size, N, -1))
            data = data.contiguous()
        locs = locs.contiguous()
        idxs = idxs.contiguous()
        nlocs, ndata, _ = _ReorderDataFunction.apply(idxs, locs, data, self.reverse)
        if data is None:
            return nlocs
        return nlocs, ndata


def _reorder(idxs, locs, data, reverse, want_pos4=False):
    nat.require_cuda_f32(idxs, "idxs")
    nat.require_cuda_f32(locs, "locs")
    B, N, D = locs.shape
    nlocs = torch.empty_like(locs)
    ndata = None
    C = 0
    if data is not None:
        nat.require_cuda_f32(data, "data")
        C = data.shape[2]
        ndata = torch.empty_like(data)
    pos4 = None
    with torch.cuda.device(locs.device):
        if want_pos4:
            pos4 = torch.empty(B, N, 4, device=locs.device, dtype=torch.float32)
            nat.check(nat.lib().spnb_reorder_data_pos4(nat.ptr(locs), nat.ptr(data), nat.ptr(idxs), nat.ptr(nlocs),
                                                       nat.ptr(ndata), nat.ptr(pos4), B, N, D, C, nat.stream()),
                      "spnb_reorder_data_pos4")
        else:
            nat.check(nat.lib().spnb_reorder_data(nat.ptr(locs), nat.ptr(data), nat.ptr(idxs),
                                                  nat.ptr(nlocs), nat.ptr(ndata), B, N, D, C,
                                                  int(reverse), nat.stream()), "spnb_reorder_data")
    return nlocs, ndata, pos4


class _ReorderDataFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, idxs, locs, data, reverse, want_pos4=False):
        ctx.save_for_backward(idxs)
        ctx.reverse = reverse
        ctx.has_data = data is not None
        nlocs, ndata, pos4 = _reorder(idxs, locs, data, reverse, want_pos4)
        if ndata is None:
            ndata = locs.new_empty(0)
            ctx.mark_non_differentiable(ndata)
        if pos4 is None:
            pos4 = locs.new_empty(0)
        ctx.mark_non_differentiable(pos4)
        return nlocs, ndata, pos4

    @staticmethod
    def backward(ctx, grad_locs, grad_data, _grad_pos4):
        idxs, = ctx.saved_tensors
        gd = grad_data.contiguous() if ctx.has_data else None
        glocs, gdata, _ = _reorder(idxs, grad_locs.contiguous(), gd, 1 - ctx.reverse)
        return None, glocs, gdata, None, None


def tile_lists_of(neighbors):
    """The compact tile lists (csrc/tile_lists.cuh) of a neighbour tensor returned by ParticleCollision, or
    None: lists built for separate query locations, ndim > 3, tile_lists switched off, or the tensor was edited
    in place after ParticleCollision returned it (sidecar.py)."""
    sc = sidecar.lookup(neighbors)
    if sc is None:
        return None
    return sc.tiles


def sym_flag_of(neighbors):
    """Device int32[1] left by ParticleCollision (0 = every list is complete and no query lies beyond a clamped
    grid, so the neighbour relation is symmetric), or None for tensors of unknown origin / edited in place."""
    sc = sidecar.lookup(neighbors)
    return None if sc is None else sc.sym_flag


class ParticleCollision(torch.nn.Module):
    """Hash-grid neighbour search (reference ParticleCollision.py:61-203)."""

    def __init__(self, ndim, radius, max_grid_dim=96, max_collisions=128, include_self=True):
        super(ParticleCollision, self).__init__()
        self.ndim = ec.check_conditions(ndim, "ndim", "%s > 0", "%s < " + str(nat_max_dim()),
                                        "isinstance(%s, numbers.Integral)")
        self.radius = ec.check_conditions(radius, "radius", "%s >= 0",
                                          "isinstance(%s, numbers.Real)")
        self.max_grid_dim = ec.check_conditions(max_grid_dim, "max_grid_dim", "%s > 0",
                                                "isinstance(%s, numbers.Integral)")
        self.max_collisions = ec.check_conditions(max_collisions, "max_collisions", "%s > 0",
                                                  "isinstance(%s, numbers.Integral)")
        self.include_self = 1 if include_self else 0
        # extension (not in the reference): the compact tile lists the ConvSP tile kernels consume.  True
        # (default): the float rows AND the tile lists come out of one kernel (spnb_compute_collisions_tiled)
        # whenever the particles are their own queries and ndim <= 3; False: float rows only (the general
        # kernel).  The float neighbour tensor that is returned is bit-identical either way.
        self.tile_lists = os.environ.get("SPNB_TILE_LISTS", "1") != "0"
        self._generation = 0
        self.radixsort_buffer_size = -1
        # Same buffer names as the reference (ParticleCollision.py:97-100) so state_dicts load.
        # cellStarts/cellEnds are allocated lazily at [B, max_grid_dim**ndim] on first use.
        self.register_buffer("cellIDs", torch.zeros(1, 1))
        self.register_buffer("cellStarts", torch.zeros(1, 1))
        self.register_buffer("cellEnds", torch.zeros(1, 1))
        self.register_buffer("cuda_buffer", torch.zeros(1,))
        self.reorder = ReorderData(reverse=False)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # The four buffers are scratch: accept whatever shape a checkpoint carries.
        for name in ("cellIDs", "cellStarts", "cellEnds", "cuda_buffer"):
            key = prefix + name
            if key in state_dict:
                getattr(self, name).resize_(state_dict[key].shape)
        super(ParticleCollision, self)._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _scratch(self, name, shape, device):
        buf = getattr(self, name)
        if buf.device != device or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, device=device, dtype=torch.float32)
            setattr(self, name, buf)
        return buf

    def forward(self, locs, data=None, qlocs=None, query_range=None, bounds=None):
        """Returns (locs, [data], idxs, neighbors) exactly as the reference
        (ParticleCollision.py:104-203): locs/data reordered by hash-grid cell, idxs[b,i] = original
        index of the particle now at i, neighbors BxMxK float lists terminated by -1.

        query_range=(start, end) is an extension for splitting ONE scene over several GPUs
        (scene_parallel.py): the queries are the reordered particles start..end-1 only, so neighbors
        is Bx(end-start)xK -- the rows start..end-1 of what the call without it returns.
        bounds=(lower_bounds, grid_dims), both BxD, replaces the grid computed from `locs` (slab_parallel.py:
        a rank that holds a slab of the scene hashes it on the WHOLE scene's grid)."""
        batch_size = locs.size()[0]
        N = locs.size()[1]
        ec.check_tensor_dims(locs, "locs", (batch_size, N, self.ndim))
        has_data = data is not None
        if has_data:
            ec.check_tensor_dims(data, "data", (batch_size, N, -1))
            data = data.contiguous()
        if qlocs is not None:
            ec.check_tensor_dims(qlocs, "qlocs", (batch_size, -1, self.ndim))
            qlocs = qlocs.contiguous()
        locs = locs.contiguous()
        nat.require_cuda_f32(locs, "locs")
        dev = locs.device
        L = nat.lib()
        D, G = self.ndim, self.max_grid_dim
        # the scratch buffers are about to be overwritten: lazy tile builders of earlier calls must refuse
        self._generation += 1
        ncells = G ** D

        ws_bytes = L.spnb_hashgrid_workspace_bytes(batch_size, N, D, G)
        self.radixsort_buffer_size = ws_bytes
        ws = self._scratch("cuda_buffer", ((ws_bytes + 3) // 4,), dev)
        cellIDs = self._scratch("cellIDs", (batch_size + 2, N, 1), dev)
        cellStarts = self._scratch("cellStarts", (batch_size, ncells), dev)
        cellEnds = self._scratch("cellEnds", (batch_size, ncells), dev)

        with torch.no_grad(), torch.cuda.device(dev):
            st = nat.stream()
            ld = locs.detach()
            if bounds is not None:
                lower_bounds = nat.require_cuda_f32(bounds[0].contiguous(), "bounds[0]")
                grid_dims = nat.require_cuda_f32(bounds[1].contiguous(), "bounds[1]")
                ec.check_tensor_dims(lower_bounds, "bounds[0]", (batch_size, D))
                ec.check_tensor_dims(grid_dims, "bounds[1]", (batch_size, D))
            else:
                lower_bounds = torch.empty(batch_size, D, device=dev, dtype=torch.float32)
                grid_dims = torch.empty(batch_size, D, device=dev, dtype=torch.float32)
                nat.check(L.spnb_grid_bounds(nat.ptr(ld), batch_size, N, D, float(self.radius), G,
                                             nat.ptr(lower_bounds), nat.ptr(grid_dims), nat.ptr(ws),
                                             ws_bytes, st), "spnb_grid_bounds")
            idxs = torch.empty(batch_size, N, device=dev, dtype=torch.float32)
            nat.check(L.spnb_hashgrid_order(nat.ptr(ld), nat.ptr(lower_bounds), nat.ptr(grid_dims),
                                            nat.ptr(cellIDs), nat.ptr(idxs), nat.ptr(ws), ws_bytes,
                                            batch_size, N, D, float(self.radius), G, st),
                      "spnb_hashgrid_order")

        K = self.max_collisions
        tile_bytes = (L.spnb_tile_lists_bytes(batch_size, N, D, K)
                      if (qlocs is None and query_range is None and self.tile_lists) else 0)
        # Reorder locs (and data) -- the only differentiable step.  On the tiled path the reordered positions
        # are also written as a float4 plane, which the list kernel stages its candidates from.
        locs, data_r, pos4 = _ReorderDataFunction.apply(idxs, locs, data if has_data else None, 0, tile_bytes > 0)
        if has_data:
            data = data_r

        if query_range is not None:
            if qlocs is not None:
                raise ValueError("query_range and qlocs are mutually exclusive")
            qs, qe = int(query_range[0]), int(query_range[1])
            if not (0 <= qs < qe <= N):
                raise ValueError("query_range must satisfy 0 <= start < end <= N")
            qlocs = locs.detach()[:, qs:qe].contiguous()
        with torch.no_grad(), torch.cuda.device(dev):
            q = locs.detach() if qlocs is None else qlocs.detach()
            nat.require_cuda_f32(q, "qlocs")
            M = q.shape[1]
            neighbors = torch.empty(batch_size, M, K, device=dev, dtype=torch.float32)
            trunc = torch.zeros(1, device=dev, dtype=torch.int32)
            tiles = None
            if tile_bytes > 0:
                tiles = torch.empty(tile_bytes, device=dev, dtype=torch.uint8)
                nat.check(L.spnb_compute_collisions_tiled(
                    nat.ptr(pos4), nat.ptr(locs.detach()), nat.ptr(lower_bounds), nat.ptr(grid_dims),
                    nat.ptr(cellIDs), nat.ptr(cellStarts), nat.ptr(cellEnds), nat.ptr(neighbors), batch_size, N, D,
                    K, ncells, float(self.radius), float(self.radius), self.include_self, nat.ptr(trunc),
                    nat.ptr(tiles), tile_bytes, nat.stream()), "spnb_compute_collisions_tiled")
            else:
                nat.check(L.spnb_compute_collisions(
                    nat.ptr(q), nat.ptr(locs.detach()), nat.ptr(lower_bounds), nat.ptr(grid_dims),
                    nat.ptr(cellIDs), nat.ptr(cellStarts), nat.ptr(cellEnds), nat.ptr(neighbors),
                    batch_size, M, N, D, K, ncells, float(self.radius),
                    float(self.radius), self.include_self, nat.ptr(trunc), nat.stream()),
                    "spnb_compute_collisions")
        if qlocs is None:
            # Lists built with the particles as their own queries are symmetric unless one was cut at
            # max_collisions or a query lies beyond a clamped grid; ConvSP's backward uses this to avoid atomics.
            sidecar.attach(neighbors, sidecar.Sidecar(sym_flag=trunc, tiles=tiles))
            if tile_bytes > 0:
                sidecar.attach(locs, sidecar.Sidecar(pos4=pos4))
        self.last_lower_bounds = lower_bounds
        self.last_grid_dims = grid_dims
        # device int32: non-zero when a row of THIS call was cut at max_collisions or a query lay beyond a clamped grid
        # (the slab decomposition combines the flags of all ranks before it relies on symmetric lists)
        self.last_trunc_flag = trunc
        if has_data:
            return locs, data, idxs, neighbors
        return locs, idxs, neighbors


def grid_bounds(locs, radius, max_grid_dim):
    """(lower_bounds, grid_dims), both BxD, of ParticleCollision's hash grid for `locs` BxNxD -- the fp32 torch
    CPU arithmetic of the reference (ParticleCollision.py:174-181), bit for bit, on the device."""
    nat.require_cuda_f32(locs, "locs")
    locs = locs.contiguous()
    B, N, D = locs.shape
    L = nat.lib()
    wsb = L.spnb_hashgrid_workspace_bytes(B, N, D, max_grid_dim)
    with torch.no_grad(), torch.cuda.device(locs.device):
        ws = torch.empty((wsb + 3) // 4, device=locs.device, dtype=torch.float32)
        low = torch.empty(B, D, device=locs.device, dtype=torch.float32)
        gd = torch.empty(B, D, device=locs.device, dtype=torch.float32)
        nat.check(L.spnb_grid_bounds(nat.ptr(locs), B, N, D, float(radius), int(max_grid_dim), nat.ptr(low),
                                     nat.ptr(gd), nat.ptr(ws), wsb, nat.stream()), "spnb_grid_bounds")
    return low, gd
