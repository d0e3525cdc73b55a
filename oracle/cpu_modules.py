"""The reference's module API on CPU tensors, backed by the reference's own native CPU functions.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  The reference's Python modules
cannot run on a current PyTorch (legacy instance-style autograd Functions), so this
file re-hosts the same three layers -- ConvSP, ParticleCollision, ReorderData -- as new-style
Functions that call the reference CPU extension (oracle/_ref, `kind = "reference"`) or, where that
was never built, the plain-C restatement (`kind = "port"`), with the reference's allocation
conventions (convsp.py:155-203, ParticleCollision.py:174-314).  bench.py times these as the
`cpu_baseline` and as `--impl reference`; tests use them as a second opinion on the fluid step.
"""
import numpy as np
import torch

from . import spn_oracle as so

KERNEL_NAMES = so.KERNEL_NAMES

_backend = None
# False: the reference CPU path's own (unstable) selection sort, cpu_layer_funcs.cpp:263-287 -- what the CPU
# baseline times.  True: the stable order of the reference's GPU path (and of the product), so that a test can
# compare the two steps with the SAME permutation and neighbour order (set by tests only; needs the C port).
STABLE_ORDER = False


def backend():
    """RefOracle when oracle/_ref exists, else COracle."""
    global _backend
    if _backend is None:
        try:
            _backend = so.RefOracle()
        except Exception:
            _backend = so.COracle()
    return _backend


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32, copy=False)


def _kernel_table():
    import math
    pi = math.pi
    # python/SmoothParticleNets/kernels.py:88,98 -- only what fluid_sim's rest-density needs
    return {
        "spiky": lambda d, H: 15.0 / (pi * H ** 3) * (1.0 - d / H) ** 2,
        "dspiky": lambda d, H: -15.0 / (pi * H ** 3) * 2.0 * (1.0 - d / H) / H,
    }


KERNEL_FN = _kernel_table()


class _ConvSPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qlocs, locs, data, neighbors, weight, bias, cfg):
        radius, ksize, dil, dis_norm, fn = cfg
        a = [_np(t) for t in (qlocs, locs, data, neighbors, weight, bias)]
        ctx.save_for_backward(qlocs, locs, data, neighbors, weight, bias)
        ctx.cfg = cfg
        return torch.from_numpy(backend().convsp_forward(*a, radius, ksize, dil, dis_norm, fn))

    @staticmethod
    def backward(ctx, go):
        radius, ksize, dil, dis_norm, fn = ctx.cfg
        a = [_np(t) for t in ctx.saved_tensors]
        dq, dl, dd, dw, db = backend().convsp_backward(*a, radius, ksize, dil, dis_norm, fn, _np(go))
        f = torch.from_numpy
        return f(dq), f(dl), f(dd), None, f(dw), f(np.ascontiguousarray(db)), None


class ConvSP(torch.nn.Module):
    def __init__(self, in_channels, out_channels, ndim, kernel_size, dilation, radius, dis_norm=False,
                 kernel_fn='default', with_params=True):
        super(ConvSP, self).__init__()
        ks = [kernel_size] * ndim if np.isscalar(kernel_size) else list(kernel_size)
        dl = [dilation] * ndim if np.isscalar(dilation) else list(dilation)
        self.ksize = np.array(ks, np.float32)
        self.dil = np.array(dl, np.float32)
        self.radius, self.dis_norm = radius, int(bool(dis_norm))
        self.kernel_fn = KERNEL_NAMES.index(kernel_fn)
        w = torch.zeros(out_channels, in_channels, int(np.prod(ks)))
        b = torch.zeros(out_channels)
        if with_params:
            self.weight, self.bias = torch.nn.Parameter(w), torch.nn.Parameter(b)
        else:
            self.register_buffer("weight", w)
            self.register_buffer("bias", b)

    def forward(self, locs, data, neighbors, qlocs=None):
        cfg = (self.radius, self.ksize, self.dil, self.dis_norm, self.kernel_fn)
        return _ConvSPFn.apply(locs if qlocs is None else qlocs, locs, data, neighbors, self.weight,
                               self.bias, cfg)


class _ReorderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idxs, locs, data, reverse):
        ctx.save_for_backward(idxs)
        ctx.reverse, ctx.has_data = reverse, data is not None
        nl, nd = backend().reorder_data(_np(locs), None if data is None else _np(data), _np(idxs), reverse)
        nd_t = torch.from_numpy(nd) if nd is not None else locs.new_empty(0)
        return torch.from_numpy(nl), nd_t

    @staticmethod
    def backward(ctx, gl, gd):
        idxs, = ctx.saved_tensors
        nl, nd = backend().reorder_data(_np(gl), _np(gd) if ctx.has_data else None, _np(idxs),
                                        1 - ctx.reverse)
        return None, torch.from_numpy(nl), (torch.from_numpy(nd) if nd is not None else None), None


class ReorderData(torch.nn.Module):
    def __init__(self, reverse=False):
        super(ReorderData, self).__init__()
        self.reverse = 1 if reverse else 0

    def forward(self, idxs, locs, data=None):
        nl, nd = _ReorderFn.apply(idxs, locs, data, self.reverse)
        return nl if data is None else (nl, nd)


class ParticleCollision(torch.nn.Module):
    def __init__(self, ndim, radius, max_grid_dim=96, max_collisions=128, include_self=True):
        super(ParticleCollision, self).__init__()
        self.ndim, self.radius, self.max_grid_dim = ndim, radius, max_grid_dim
        self.max_collisions, self.include_self = max_collisions, 1 if include_self else 0
        self.reorder = ReorderData(reverse=False)

    def forward(self, locs, data=None, qlocs=None, query_range=None, bounds=None):
        """query_range / bounds: the two extensions of the product's ParticleCollision that the multi-GPU host
        logic uses (tests/test_sharding_gloo.py drives that logic with these CPU modules)."""
        be = backend()
        ln = _np(locs)
        if bounds is not None:
            low, gd = _np(bounds[0]), _np(bounds[1])
        else:
            low, gd = so.grid_bounds_torch(ln, self.radius, self.max_grid_dim)  # the reference's torch ops
        if STABLE_ORDER:
            ids, idxs = so.COracle().hashgrid_order(ln, low, gd, self.radius, stable=True)
        else:
            ids, idxs = be.hashgrid_order(ln, low, gd, self.radius, stable=False)  # its selection sort
        idxs_t = torch.from_numpy(idxs)
        if data is not None:
            locs, data = self.reorder(idxs_t, locs, data)
        else:
            locs = self.reorder(idxs_t, locs)
        q = _np(locs) if qlocs is None else _np(qlocs)
        if query_range is not None:
            q = np.ascontiguousarray(_np(locs)[:, int(query_range[0]):int(query_range[1])])
        nb, _, _ = be.compute_collisions(q, _np(locs), low, gd, ids, self.radius, self.radius,
                                         self.max_collisions, self.include_self,
                                         self.max_grid_dim ** self.ndim)
        nb_t = torch.from_numpy(nb)
        if data is not None:
            return locs, data, idxs_t, nb_t
        return locs, idxs_t, nb_t
