"""ConvSP -- the Smooth Particle Convolution layer, on libspnb (sm_100a).

Drop-in for python/SmoothParticleNets/convsp.py of the reference: same constructor, buffers,
``forward(locs, data, neighbors, qlocs=None)`` and autograd semantics (gradients for qlocs, locs,
data, weight, bias; none for neighbors, convsp.py:198-203).  The legacy instance-style autograd
Function of the reference (convsp.py:140-203) is a static new-style Function here.
"""
import numbers
import os  # noqa: F401

import numpy as np
import torch

from . import _native as nat
from . import error_checking as ec
from . import sidecar
from .kernels import KERNEL_NAMES


class ConvSP(torch.nn.Module):
    """Smooth Particle Convolution (reference convsp.py:14-132)."""

    def __init__(self, in_channels, out_channels, ndim, kernel_size, dilation, radius,
                 dis_norm=False, kernel_fn='default', with_params=True):
        super(ConvSP, self).__init__()
        self.nchannels = ec.check_conditions(in_channels, "in_channels",
                                             "%s > 0", "isinstance(%s, numbers.Integral)")
        self.nkernels = ec.check_conditions(out_channels, "out_channels",
                                            "%s > 0", "isinstance(%s, numbers.Integral)")
        self.ndim = ec.check_conditions(ndim, "ndim", "%s > 0",
                                        "%s < " + str(nat_max_dim()),
                                        "isinstance(%s, numbers.Integral)")
        self._kernel_size = ec.make_list(kernel_size, ndim, "kernel_size", "%s >= 0",
                                         "%s %% 2 == 1 # Must be odd",
                                         "isinstance(%s, numbers.Integral)")
        self._dilation = ec.make_list(dilation, ndim, "dilation", "%s >= 0",
                                      "isinstance(%s, numbers.Real)")
        self.radius = ec.check_conditions(radius, "radius", "%s >= 0",
                                          "isinstance(%s, numbers.Real)")
        self.kernel_fn = ec.check_conditions(kernel_fn, "kernel_fn", "%s in " + str(KERNEL_NAMES))
        self.kernel_fn = KERNEL_NAMES.index(self.kernel_fn)
        self.dis_norm = (1 if dis_norm else 0)
        self.ncells = int(np.prod(self._kernel_size))

        # Uninitialised storage, as in the reference (convsp.py:71-80): the caller fills it.
        weight = torch.empty(self.nkernels, self.nchannels, self.ncells)
        bias = torch.empty(self.nkernels)
        if with_params:
            self.register_parameter("weight", torch.nn.Parameter(weight))
            self.register_parameter("bias", torch.nn.Parameter(bias))
        else:
            self.register_buffer("weight", weight)
            self.register_buffer("bias", bias)
        self.register_buffer("kernel_size", ec.list2tensor(self._kernel_size))
        self.register_buffer("dilation", ec.list2tensor(self._dilation))
        # Kept for attribute compatibility (convsp.py:85-86); unused.
        self.nshared_device_mem = -1
        self.device_id = -1
        # True (default; environment SPNB_FAST_PATH=0 turns it off) routes a kernel_size-1 layer with up to 4 input
        # channels through the single-layer signature of the tile kernels (pack pre-pass + k_tile_fwd / k_tile_bwd) when
        # the neighbour tensor carries tile lists: same values within fp32 rounding; d(weight) = go^T T comes from one
        # more fused forward pass with identity weights (convsp_group.py).
        # Measured on B200 (profiles/README.md): the per-layer fluid step 6.86 -> 6.53 ms; a tile kernel costs the
        # same for one layer as for six, so ConvSPGroup is where layers sharing (locs, neighbors) really win.
        self.fast_path = os.environ.get("SPNB_FAST_PATH", "1") != "0"

    def forward(self, locs, data, neighbors, qlocs=None):
        """locs BxNxD, data BxNxC, neighbors BxMxK (float indices, -1 terminated), qlocs BxMxD or
        None.  Returns BxMxO evaluated at qlocs (or locs)."""
        batch_size = locs.size()[0]
        N = locs.size()[1]
        ec.check_tensor_dims(locs, "locs", (batch_size, N, self.ndim))
        ec.check_tensor_dims(data, "data", (batch_size, N, self.nchannels))
        locs = locs.contiguous()
        data = data.contiguous()
        if qlocs is not None:
            ec.check_tensor_dims(qlocs, "qlocs", (batch_size, -1, self.ndim))
            qlocs = qlocs.contiguous()
        # Symmetry tag left by ParticleCollision on the neighbour tensor it returned (device int32:
        # 0 = every list is complete, so the relation is symmetric).
        sc = sidecar.lookup(neighbors) if qlocs is None else None
        sym_flag = None if sc is None else sc.sym_flag
        if (sc is not None and sc.tiles is not None and self.ncells == 1 and self.nchannels <= 4
                and self.fast_path):
            # kernel_size 1 on lists that carry tile lists: the single-layer signature of the tile kernels
            # (csrc/convsp_group.cuh) -- same values within fp32 rounding
            from .convsp_group import group_apply
            out = group_apply([self], locs, [data], neighbors)
            if out is not None:
                return out[0]
        neighbors = neighbors.contiguous() if not neighbors.is_contiguous() else neighbors
        return _ConvSPFunction.apply(qlocs, locs, data, neighbors, self.weight, self.bias,
                                     float(self.radius), self.kernel_size, self.dilation,
                                     self.dis_norm, self.kernel_fn, self.ncells, sym_flag)


def nat_max_dim():
    # MAX_CARTESIAN_DIM of the reference (src/constants.h:7); does not need the GPU.
    return 20


class _ConvSPFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, qlocs, locs, data, neighbors, weight, bias, radius, kernel_size, dilation,
                dis_norm, kernel_fn, ncells, sym_flag):
        for t, n in ((locs, "locs"), (data, "data"), (neighbors, "neighbors"), (weight, "weight"),
                     (bias, "bias"), (kernel_size, "kernel_size"), (dilation, "dilation")):
            nat.require_cuda_f32(t, n)
        if qlocs is not None:
            nat.require_cuda_f32(qlocs, "qlocs")
        q = locs if qlocs is None else qlocs
        B, N, D = locs.shape
        M = q.shape[1]
        C = data.shape[2]
        K = neighbors.shape[2]
        O = weight.shape[0]
        if neighbors.shape[0] != B or neighbors.shape[1] != M:
            raise ValueError("neighbors must be %dx%dxK, not %s" % (B, M, tuple(neighbors.shape)))
        out = torch.empty(B, M, O, device=locs.device, dtype=torch.float32)
        L = nat.lib()
        # wide channel counts: factored gather + dense contraction (csrc/convsp_wide.cu)
        wide_bytes = L.spnb_convsp_forward_wide_workspace_bytes(O, C, D, ncells) if C * O * ncells >= 4096 else 0
        with torch.cuda.device(locs.device):
            if wide_bytes:
                ws = torch.empty((wide_bytes + 3) // 4, device=locs.device, dtype=torch.float32)
                nat.check(L.spnb_convsp_forward_wide(
                    nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(neighbors), nat.ptr(weight),
                    nat.ptr(bias), B, M, N, C, D, K, O, ncells, radius, nat.ptr(kernel_size),
                    nat.ptr(dilation), dis_norm, kernel_fn, nat.ptr(out), nat.ptr(ws), wide_bytes,
                    nat.stream()), "spnb_convsp_forward_wide")
            else:
                nat.check(L.spnb_convsp_forward(
                    nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(neighbors), nat.ptr(weight),
                    nat.ptr(bias), B, M, N, C, D, K, O, ncells, radius, nat.ptr(kernel_size),
                    nat.ptr(dilation), dis_norm, kernel_fn, nat.ptr(out), nat.stream()),
                    "spnb_convsp_forward")
        ctx.save_for_backward(qlocs, locs, data, neighbors, weight, kernel_size, dilation)
        ctx.cfg = (radius, dis_norm, kernel_fn, ncells)
        ctx.sym_flag = sym_flag
        return out

    @staticmethod
    def backward(ctx, grad_output):
        qlocs, locs, data, neighbors, weight, kernel_size, dilation = ctx.saved_tensors
        radius, dis_norm, kernel_fn, ncells = ctx.cfg
        need_q, need_l, need_d, _, need_w, need_b = ctx.needs_input_grad[:6]
        grad_output = grad_output.contiguous()
        q = locs if qlocs is None else qlocs
        B, N, D = locs.shape
        M = q.shape[1]
        C = data.shape[2]
        K = neighbors.shape[2]
        O = weight.shape[0]
        dev = locs.device
        dq = dl = dd = dw = None
        if qlocs is None:
            # locs plays both roles: one buffer receives d/dqlocs + d/dlocs.
            if need_l:
                dq = dl = torch.empty(B, N, D, device=dev, dtype=torch.float32)
        else:
            if need_q:
                dq = torch.empty(B, M, D, device=dev, dtype=torch.float32)
            if need_l:
                dl = torch.empty(B, N, D, device=dev, dtype=torch.float32)
        if need_d:
            dd = torch.empty(B, N, C, device=dev, dtype=torch.float32)
        if need_w:
            dw = torch.empty_like(weight)
        L = nat.lib()
        wide_bytes = (L.spnb_convsp_backward_wide_workspace_bytes(O, C, D, ncells)
                      if C * O * ncells >= 4096 else 0)
        if wide_bytes and (dq is not None or dl is not None or dd is not None or dw is not None):
            # wide channel counts: two tensor-core contractions + one list walk (csrc/convsp_wide_bwd.cu)
            ws = torch.empty((wide_bytes + 3) // 4, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                nat.check(L.spnb_convsp_backward_wide(
                    nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(neighbors), nat.ptr(weight),
                    B, M, N, C, D, K, O, ncells, radius, nat.ptr(kernel_size), nat.ptr(dilation),
                    dis_norm, kernel_fn, nat.ptr(grad_output), nat.ptr(dq), nat.ptr(dl),
                    nat.ptr(dd), nat.ptr(dw), nat.ptr(ws), wide_bytes, nat.stream()),
                    "spnb_convsp_backward_wide")
        elif dq is not None or dl is not None or dd is not None or dw is not None:
            with torch.cuda.device(dev):
                nat.check(L.spnb_convsp_backward(
                    nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(neighbors), nat.ptr(weight),
                    B, M, N, C, D, K, O, ncells, radius, nat.ptr(kernel_size), nat.ptr(dilation),
                    dis_norm, kernel_fn, nat.ptr(grad_output), nat.ptr(dq), nat.ptr(dl),
                    nat.ptr(dd), nat.ptr(dw), nat.ptr(ctx.sym_flag), None, nat.stream()),
                    "spnb_convsp_backward")
        db = grad_output.sum(1).sum(0) if need_b else None  # convsp.py:203
        return (dq if qlocs is not None else None, dl, dd, None, dw, db,
                None, None, None, None, None, None, None)
