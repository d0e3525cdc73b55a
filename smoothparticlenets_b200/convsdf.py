"""ConvSDF -- convolution of signed-distance-field objects at particle locations, on libspnb.

Drop-in for python/SmoothParticleNets/convsdf.py of the reference: same constructor, ``SetSDFs``
packing (flat atlas + exclusive-cumsum offsets + [dims..., cell_size] shape rows, convsdf.py:82-98),
``forward(locs, idxs, poses, scales)`` and autograd semantics (gradients for locs, weight, bias and
-- iff compute_pose_grads -- poses; none for idxs / scales, convsdf.py:226-233).  As in the
reference, the rotation components of the pose gradient are forward finite differences with
eps = 1e-3 evaluated in Python (convsdf.py:211-224); the translation components are analytic.

Extension: ``compute_pose_grads="analytic"`` computes the rotation components in the backward
kernel as well (exact derivative of the forward formula with respect to the 2-D angle / the four
quaternion components as independent variables), which removes the M * R extra forward passes.
"""
import numbers  # noqa: F401

import numpy as np
import torch

from . import _native as nat
from . import error_checking as ec
from .convsp import nat_max_dim


class ConvSDF(torch.nn.Module):

    def __init__(self, sdfs, sdf_sizes, out_channels, ndim, kernel_size, dilation,
                 max_distance, with_params=True, compute_pose_grads=False):
        super(ConvSDF, self).__init__()
        self.nkernels = ec.check_conditions(out_channels, "out_channels", "%s > 0",
                                            "isinstance(%s, numbers.Integral)")
        self.ndim = ec.check_conditions(ndim, "ndim", "%s > 0", "%s < " + str(nat_max_dim()),
                                        "%s in [1, 2, 3] # Only 1-, 2-, and 3-D are suported",
                                        "isinstance(%s, numbers.Integral)")
        self.max_distance = ec.check_conditions(max_distance, "max_distance", "%s >= 0",
                                                "isinstance(%s, numbers.Real)")
        self._kernel_size = ec.make_list(kernel_size, ndim, "kernel_size", "%s >= 0",
                                         "%s %% 2 == 1 # Must be odd",
                                         "isinstance(%s, numbers.Integral)")
        self._dilation = ec.make_list(dilation, ndim, "dilation", "%s >= 0",
                                      "isinstance(%s, numbers.Real)")
        self.register_buffer("sdfs", torch.zeros(1))
        self.register_buffer("sdf_shapes", torch.zeros(1))
        self.register_buffer("sdf_offsets", torch.zeros(1))
        self.SetSDFs(sdfs, sdf_sizes)

        self.ncells = int(np.prod(self._kernel_size))
        weight = torch.empty(self.nkernels, self.ncells)
        bias = torch.empty(self.nkernels)
        if with_params:
            self.register_parameter("weight", torch.nn.Parameter(weight))
            self.register_parameter("bias", torch.nn.Parameter(bias))
        else:
            self.register_buffer("weight", weight)
            self.register_buffer("bias", bias)
        # True: the reference's recipe (finite-difference rotation columns); "analytic": in-kernel.
        self.compute_pose_grads = ("analytic" if compute_pose_grads == "analytic"
                                   else bool(compute_pose_grads))
        self._kernel_size = ec.list2tensor(self._kernel_size)
        self._dilation = ec.list2tensor(self._dilation)
        self.register_buffer("kernel_size", self._kernel_size)
        self.register_buffer("dilation", self._dilation)

    def SetSDFs(self, sdfs, sdf_sizes):
        cell_sizes = [ec.check_conditions(x, "sdf_sizes[%d]" % i, "%s > 0",
                                          "isinstance(%s, numbers.Real)")
                      for i, x in enumerate(sdf_sizes)]
        _sdfs = [ec.check_conditions(sdf, "sdfs[%d]" % i, "isinstance(%s, torch.Tensor)",
                                     "len(%s.size()) == " + str(self.ndim))
                 for i, sdf in enumerate(sdfs)]
        dev = self.sdfs.device
        shapes = ec.list2tensor([list(x.size()) + [cell_sizes[i]] for i, x in enumerate(_sdfs)])
        flat = [x.contiguous().view(-1).to(dtype=torch.float32) for x in _sdfs]
        offsets = ec.list2tensor([0] + np.cumsum([x.size()[0] for x in flat])[:-1].tolist())
        self.sdfs = torch.cat(flat).to(dev)
        self.sdf_shapes = shapes.to(dev)
        self.sdf_offsets = offsets.to(dev)

    def forward(self, locs, idxs, poses, scales):
        """locs BxNxD, idxs BxM (float SDF indices, -1 = unused), poses BxMx(D+R) (R = 0/1/4;
        quaternions xyzw), scales BxM.  Returns BxNxO."""
        batch_size = locs.size()[0]
        N = locs.size()[1]
        M = idxs.size()[1]
        R = {1: 0, 2: 1, 3: 4}[self.ndim]
        ec.check_tensor_dims(locs, "locs", (batch_size, N, self.ndim))
        ec.check_tensor_dims(idxs, "idxs", (batch_size, M,))
        ec.check_tensor_dims(poses, "poses", (batch_size, M, self.ndim + R))
        ec.check_tensor_dims(scales, "scales", (batch_size, M,))
        return _ConvSDFFunction.apply(locs.contiguous(), idxs.contiguous(), poses.contiguous(),
                                      scales.contiguous(), self.weight, self.bias, self.sdfs,
                                      self.sdf_offsets, self.sdf_shapes, self.kernel_size,
                                      self.dilation, float(self.max_distance),
                                      self.compute_pose_grads)


def _sdf_forward(locs, idxs, poses, scales, weight, bias, sdfs, offs, shapes, ksize, dil, max_distance):
    B, N, D = locs.shape
    S = idxs.shape[1]
    O, ncells = weight.shape
    out = torch.empty(B, N, O, device=locs.device, dtype=torch.float32)
    with torch.cuda.device(locs.device):
        nat.check(nat.lib().spnb_convsdf_forward(
            nat.ptr(locs), B, N, D, nat.ptr(idxs), nat.ptr(poses), nat.ptr(scales), S,
            poses.shape[2], nat.ptr(sdfs), sdfs.numel(), nat.ptr(offs), nat.ptr(shapes),
            shapes.shape[0], nat.ptr(weight), nat.ptr(bias), O, ncells, nat.ptr(ksize), nat.ptr(dil),
            max_distance, nat.ptr(out), nat.stream()), "spnb_convsdf_forward")
    return out


class _ConvSDFFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, locs, idxs, poses, scales, weight, bias, sdfs, offs, shapes, ksize, dil,
                max_distance, compute_pose_grads):
        for t, n in ((locs, "locs"), (idxs, "idxs"), (poses, "poses"), (scales, "scales"),
                     (weight, "weight"), (bias, "bias"), (sdfs, "sdfs"), (offs, "sdf_offsets"),
                     (shapes, "sdf_shapes"), (ksize, "kernel_size"), (dil, "dilation")):
            nat.require_cuda_f32(t, n)
        ctx.save_for_backward(locs, idxs, poses, scales, weight, bias, sdfs, offs, shapes, ksize, dil)
        ctx.max_distance = max_distance
        ctx.compute_pose_grads = compute_pose_grads
        return _sdf_forward(locs, idxs, poses, scales, weight, bias, sdfs, offs, shapes, ksize, dil,
                            max_distance)

    @staticmethod
    def backward(ctx, grad_output):
        locs, idxs, poses, scales, weight, bias, sdfs, offs, shapes, ksize, dil = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        need_l, _, need_p, _, need_w, need_b = ctx.needs_input_grad[:6]
        need_p = need_p and ctx.compute_pose_grads
        B, N, D = locs.shape
        S = idxs.shape[1]
        O, ncells = weight.shape
        dev = locs.device
        dl = torch.empty_like(locs) if need_l else None
        dw = torch.empty_like(weight) if need_w else None
        dp = torch.empty_like(poses) if need_p else None
        analytic = ctx.compute_pose_grads == "analytic"
        if need_l or need_w or need_p:
            fn = "spnb_convsdf_backward_analytic" if analytic else "spnb_convsdf_backward"
            with torch.cuda.device(dev):
                nat.check(getattr(nat.lib(), fn)(
                    nat.ptr(locs), B, N, D, nat.ptr(idxs), nat.ptr(poses), nat.ptr(scales), S,
                    poses.shape[2], nat.ptr(sdfs), sdfs.numel(), nat.ptr(offs), nat.ptr(shapes),
                    shapes.shape[0], nat.ptr(weight), O, ncells, nat.ptr(ksize), nat.ptr(dil),
                    ctx.max_distance, nat.ptr(grad_output), nat.ptr(dl), nat.ptr(dw), nat.ptr(dp),
                    nat.stream()), fn)
        if need_p and poses.shape[2] > D and not analytic:
            # Rotation components by forward differences, as the reference does
            # (convsdf.py:211-224): eps = 1e-3, one extra forward per (object, component).
            args = (scales, weight, bias, sdfs, offs, shapes, ksize, dil, ctx.max_distance)
            baseline = _sdf_forward(locs, idxs, poses, *args)
            pp = poses.clone()
            for m in range(S):
                for i in range(D, poses.shape[2]):
                    pp[:, m, i] += 1e-3
                    nn = _sdf_forward(locs, idxs, pp, *args)
                    pp[:, m, i] = poses[:, m, i]
                    gg = (nn - baseline) / 1e-3
                    dp[:, m, i] = torch.sum(torch.sum(gg * grad_output, 1), 1)
        db = grad_output.sum(1).sum(0) if need_b else None
        return (dl, None, dp, None, dw, db, None, None, None, None, None, None, None)
