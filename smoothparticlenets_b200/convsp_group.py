"""ConvSPGroup -- several ConvSP layers evaluated in one pass over shared neighbour lists.

An opt-in extension, not part of the reference API: the solver iteration of
the reference's fluid simulation evaluates 9 ConvSP layers on the same ``(locs, neighbors)``
(examples/fluid_sim.py:367-397); each layer call re-reads the lists, re-gathers the neighbour
positions and recomputes every distance.  ``ConvSPGroup([layer_a, layer_b, ...])(locs, [data_a,
data_b, ...], neighbors)`` returns exactly what ``[layer_a(locs, data_a, neighbors), ...]`` returns
(same values within fp32 rounding, same gradients), but walks the lists once per group.

Requirements for the fused path: every layer has kernel_size 1, the same ndim and radius, no query
locations (the particles are their own queries), and the
layer list is one of the compiled-in signatures (csrc/convsp_group_inst_*.cu: the groups of the fluid
step, and ANY single layer with up to 4 input channels).  Anything else silently runs the ordinary
per-layer path, so the module is always safe to use.  A data entry may be None: a one-channel layer
whose data is all ones (density, neighbour count) then reads no data at all.

Trainable weights: d(weight) of a kernel_size-1 layer is go^T T with T[i, c] = sum_j W * norm * data[j, c], the
quantity the forward accumulates before it applies the weights.  The backward obtains T from one more pass of
the same fused forward with identity weights and contracts it with grad_output (a [O x BN] x [BN x C] product),
so layers with ``with_params=True`` fuse as well (reference formula: common_funcs.h:542-547).
"""
import ctypes

import torch

from . import _native as nat
from .convsp import ConvSP
from .particlecollision import tile_lists_of, sym_flag_of


class ConvSPGroup(torch.nn.Module):

    def __init__(self, layers):
        super(ConvSPGroup, self).__init__()
        layers = list(layers)
        if not layers or not all(isinstance(l, ConvSP) for l in layers):
            raise ValueError("ConvSPGroup needs a non-empty list of ConvSP layers")
        self.layers = torch.nn.ModuleList(layers)
        l0 = layers[0]
        self._fusable = (len(layers) <= 6 and
                         all(l.ncells == 1 and l.ndim == l0.ndim and float(l.radius) == float(l0.radius)
                             for l in layers))

    def forward(self, locs, datas, neighbors):
        """locs BxNxD, datas: one BxNxC_l tensor per layer -- or None for a one-channel layer whose data is all
        ones (a density / neighbour count: nothing is read for it) --, neighbors BxNxK.  Returns a tuple with one
        BxNxO_l tensor per layer."""
        layers = list(self.layers)
        if len(datas) != len(layers):
            raise ValueError("ConvSPGroup: expected %d data tensors, got %d" % (len(layers), len(datas)))
        for l, d in zip(layers, datas):
            if d is None and l.nchannels != 1:
                raise ValueError("ConvSPGroup: data=None (ones) needs a layer with one input channel")
        out = group_apply(layers, locs, datas, neighbors) if self._fusable else None
        if out is not None:
            return out
        B, N = locs.shape[0], locs.shape[1]
        ones = None
        res = []
        for l, d in zip(layers, datas):
            if d is None:
                if ones is None:
                    ones = torch.ones(B, N, 1, device=locs.device, dtype=locs.dtype)
                d = ones
            res.append(l(locs, d, neighbors))
        return tuple(res)


def group_apply(layers, locs, datas, neighbors):
    """The fused evaluation of `layers` (all kernel_size 1, same ndim and radius) on (locs, neighbors), or None when
    it does not apply: CPU tensors, separate query locations, or a channel layout that is not compiled in."""
    if not (locs.is_cuda and neighbors.dim() == 3 and neighbors.shape[1] == locs.shape[1]):
        return None
    locs_c = locs.contiguous()
    datas_c = [None if d is None else d.contiguous() for d in datas]
    cfg = tuple((l.kernel_fn, l.dis_norm, l.nchannels, l.nkernels) for l in layers)
    if not _supported(locs_c, datas_c, layers, cfg):
        return None
    sym_flag = sym_flag_of(neighbors)
    tiles = tile_lists_of(neighbors)
    flat = list(datas_c) + [l.weight for l in layers] + [l.bias for l in layers]
    return _ConvSPGroupFunction.apply(locs_c, neighbors.contiguous(), (sym_flag, tiles), float(layers[0].radius), cfg,
                                      *flat)


def _layer_array(locs, datas, weights, biases, cfg, outs=None, gos=None, ddatas=None):
    n = len(cfg)
    arr = (nat.GroupLayer * n)()
    for i, (fn, dn, C, O) in enumerate(cfg):
        arr[i].data = nat.ptr(datas[i])
        arr[i].weight = nat.ptr(weights[i])
        arr[i].bias = nat.ptr(biases[i]) if biases is not None else None
        arr[i].out = nat.ptr(outs[i]) if outs is not None else None
        arr[i].grad_out = nat.ptr(gos[i]) if gos is not None else None
        arr[i].ddata = nat.ptr(ddatas[i]) if ddatas is not None else None
        arr[i].nchannels, arr[i].nkernels, arr[i].kernel_fn, arr[i].dis_norm = C, O, fn, dn
    return arr


def _supported(locs, datas, layers, cfg):
    B, N, D = locs.shape
    arr = _layer_array(locs, datas, [l.weight for l in layers], None, cfg)
    return nat.lib().spnb_convsp_group_workspace_bytes(nat.ptr(locs), B, N, D, float(layers[0].radius),
                                                       len(cfg), arr, 0) > 0


class _ConvSPGroupFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, locs, neighbors, flags, radius, cfg, *flat):
        sym_flag, tiles = flags
        if tiles is not None and tiles.numel() != nat.lib().spnb_tile_lists_bytes(
                locs.shape[0], locs.shape[1], locs.shape[2], neighbors.shape[2]):
            tiles = None  # not the sidecar of a tensor of this shape
        n = len(cfg)
        datas, weights, biases = flat[:n], flat[n:2 * n], flat[2 * n:3 * n]
        for t in (locs, neighbors) + tuple(flat):
            if t is not None:
                nat.require_cuda_f32(t, "ConvSPGroup operand")
        B, N, D = locs.shape
        K = neighbors.shape[2]
        L = nat.lib()
        outs = [torch.empty(B, N, c[3], device=locs.device, dtype=torch.float32) for c in cfg]
        arr = _layer_array(locs, datas, weights, biases, cfg, outs=outs)
        wsb = L.spnb_convsp_group_workspace_bytes(nat.ptr(locs), B, N, D, radius, n, arr, 0)
        ws = torch.empty((wsb + 3) // 4, device=locs.device, dtype=torch.float32)
        with torch.cuda.device(locs.device):
            nat.check(L.spnb_convsp_group_forward(nat.ptr(locs), nat.ptr(neighbors), B, N, D, K, radius, n,
                                                  arr, nat.ptr(ws), wsb, nat.ptr(tiles), nat.stream()),
                      "spnb_convsp_group_forward")
        ctx.has_data = [d is not None for d in datas]
        ctx.save_for_backward(locs, neighbors, *[d for d in datas if d is not None], *weights)
        ctx.cfg, ctx.radius, ctx.sym_flag, ctx.tiles = cfg, radius, sym_flag, tiles
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        cfg, radius = ctx.cfg, ctx.radius
        n = len(cfg)
        saved = ctx.saved_tensors
        locs, neighbors = saved[0], saved[1]
        nd = sum(ctx.has_data)
        it = iter(saved[2:2 + nd])
        datas = [next(it) if h else None for h in ctx.has_data]
        weights = saved[2 + nd:2 + nd + n]
        B, N, D = locs.shape
        K = neighbors.shape[2]
        dev = locs.device
        gos = [g.contiguous() if g is not None else torch.zeros(B, N, cfg[i][3], device=dev)
               for i, g in enumerate(grad_outs)]
        need_locs = ctx.needs_input_grad[0]
        dlocs = torch.empty(B, N, D, device=dev, dtype=torch.float32)
        # every data TENSOR gets its gradient buffer (the compiled signatures do not depend on requires_grad)
        ddatas = [torch.empty_like(d) if d is not None else None for d in datas]
        L = nat.lib()
        arr = _layer_array(locs, datas, weights, None, cfg, gos=gos, ddatas=ddatas)
        wsb = L.spnb_convsp_group_workspace_bytes(nat.ptr(locs), B, N, D, radius, n, arr, 1)
        ws = torch.empty((wsb + 3) // 4, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            nat.check(L.spnb_convsp_group_backward(nat.ptr(locs), nat.ptr(neighbors), B, N, D, K, radius, n,
                                                   arr, nat.ptr(dlocs), nat.ptr(ctx.sym_flag), nat.ptr(ws),
                                                   wsb, nat.ptr(ctx.tiles), nat.stream()),
                      "spnb_convsp_group_backward")
        need_data = ctx.needs_input_grad[5:5 + n]
        ddatas = [g if need_data[i] else None for i, g in enumerate(ddatas)]
        need_w = ctx.needs_input_grad[5 + n:5 + 2 * n]
        dweights = [None] * n
        if any(need_w):
            # T_l[i, c] = sum_j W * norm * data_l[j, c]: the fused forward once more, with identity weights
            eyes = [torch.eye(c[2], device=dev, dtype=torch.float32).reshape(c[2], c[2], 1).contiguous() for c in cfg]
            zeros = [torch.zeros(c[2], device=dev, dtype=torch.float32) for c in cfg]
            cfg_t = tuple((fn, dn, C, C) for fn, dn, C, O in cfg)
            ts = [torch.empty(B, N, c[2], device=dev, dtype=torch.float32) for c in cfg]
            arr_t = _layer_array(locs, datas, eyes, zeros, cfg_t, outs=ts)
            wsf = L.spnb_convsp_group_workspace_bytes(nat.ptr(locs), B, N, D, radius, n, arr_t, 0)
            wst = torch.empty((wsf + 3) // 4, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                nat.check(L.spnb_convsp_group_forward(nat.ptr(locs), nat.ptr(neighbors), B, N, D, K, radius, n,
                                                      arr_t, nat.ptr(wst), wsf, nat.ptr(ctx.tiles), nat.stream()),
                          "spnb_convsp_group_forward (d(weight) pass)")
            for i in range(n):
                if need_w[i]:
                    dweights[i] = torch.einsum("bno,bnc->oc", gos[i], ts[i]).unsqueeze(-1)
        dbias = [gos[i].sum(1).sum(0) if ctx.needs_input_grad[5 + 2 * n + i] else None for i in range(n)]
        return ((dlocs if need_locs else None, None, None, None, None) + tuple(ddatas) + tuple(dweights) +
                tuple(dbias))
