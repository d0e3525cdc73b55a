"""ParticleProjection -- splat particles as Gaussians into a camera image, on libspnb.

Drop-in for python/SmoothParticleNets/ParticleProjection.py of the reference: same constructor
(camera_fl, camera_size, filter_std, filter_scale), ``forward(locs, camera_pose, camera_rot, depth_mask=None)``
returning a BxHxW image, gradients for locs (and, through the torch camera transform, for camera_pose); none for
camera_rot (the reference builds the rotation matrix from ``.data``, ParticleProjection.py:52-84) and a zero
gradient for depth_mask (ParticleProjection.py:188-204).  3-D particles only.
"""
import numbers  # noqa: F401

import numpy as np
import torch

from . import _native as nat
from . import error_checking as ec

MAX_FLOAT = float(np.finfo(np.float32).max)


def rotation_matrix_from_quaternion(quat):
    """Bx4 xyzw quaternions -> Bx3x3, the transposed layout the reference multiplies particles with
    (ParticleProjection.py:52-84); detached from autograd like the reference's."""
    quat = quat.detach()
    qx, qy, qz, qw = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    ret = quat.new_empty(quat.shape[0], 3, 3)
    ret[:, 0, 0] = 1 - 2 * qy * qy - 2 * qz * qz
    ret[:, 1, 0] = 2 * qx * qy - 2 * qz * qw
    ret[:, 2, 0] = 2 * qx * qz + 2 * qy * qw
    ret[:, 0, 1] = 2 * qx * qy + 2 * qz * qw
    ret[:, 1, 1] = 1 - 2 * qx * qx - 2 * qz * qz
    ret[:, 2, 1] = 2 * qy * qz - 2 * qx * qw
    ret[:, 0, 2] = 2 * qx * qz - 2 * qy * qw
    ret[:, 1, 2] = 2 * qy * qz + 2 * qx * qw
    ret[:, 2, 2] = 1 - 2 * qx * qx - 2 * qy * qy
    return ret


def to_camera_space(locs, camera_pose, camera_rot):
    """World-space particles -> camera space (ParticleProjection.py:125-147): translate, normalise the quaternion,
    invert it, rotate."""
    locs = locs - camera_pose.unsqueeze(1)
    camera_rot = camera_rot / torch.sqrt(torch.sum(camera_rot ** 2, 1, keepdim=True))
    inv = camera_rot.new_tensor([[-1.0, -1.0, -1.0, 1.0]])
    rot = rotation_matrix_from_quaternion(camera_rot * inv)
    return torch.bmm(locs, rot).contiguous()


class ParticleProjection(torch.nn.Module):

    def __init__(self, camera_fl, camera_size, filter_std, filter_scale):
        super(ParticleProjection, self).__init__()
        self.camera_size = ec.make_list(camera_size, 2, "camera_size", "%s > 0",
                                        "isinstance(%s, numbers.Integral)")
        self.camera_fl = ec.check_conditions(camera_fl, "camera_fl", "%s > 0", "isinstance(%s, numbers.Real)")
        self.filter_std = ec.check_conditions(filter_std, "filter_std", "%s > 0", "isinstance(%s, numbers.Real)")
        self.filter_scale = ec.check_conditions(filter_scale, "filter_scale", "%s > 0",
                                                "isinstance(%s, numbers.Real)")
        self.register_buffer("empty_depth_mask",
                             torch.ones(1, self.camera_size[1], self.camera_size[0]) * MAX_FLOAT)

    def forward(self, locs, camera_pose, camera_rot, depth_mask=None):
        """locs BxNx3, camera_pose Bx3, camera_rot Bx4 (xyzw), depth_mask BxHxW or None.  Returns BxHxW."""
        batch_size = locs.size()[0]
        N = locs.size()[1]
        ec.check_tensor_dims(locs, "locs", (batch_size, N, 3))
        ec.check_tensor_dims(camera_pose, "camera_pose", (batch_size, 3))
        ec.check_tensor_dims(camera_rot, "camera_rot", (batch_size, 4))
        if depth_mask is not None:
            ec.check_tensor_dims(depth_mask, "depth_mask", (batch_size, self.camera_size[1], self.camera_size[0]))
            depth_mask = depth_mask.contiguous()
        else:
            if self.empty_depth_mask.size()[0] != batch_size:
                self.empty_depth_mask = self.empty_depth_mask.new_full(
                    (batch_size, self.camera_size[1], self.camera_size[0]), MAX_FLOAT)
            depth_mask = self.empty_depth_mask.to(locs.device)
        cam = to_camera_space(locs, camera_pose, camera_rot)
        return _ParticleProjectionFunction.apply(cam, depth_mask, float(self.camera_fl), float(self.filter_std),
                                                 float(self.filter_scale))


class _ParticleProjectionFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, locs, depth_mask, camera_fl, filter_std, filter_scale):
        nat.require_cuda_f32(locs, "locs")
        nat.require_cuda_f32(depth_mask, "depth_mask")
        ctx.save_for_backward(locs, depth_mask)
        ctx.cfg = (camera_fl, filter_std, filter_scale)
        B, N, _ = locs.shape
        H, W = depth_mask.shape[1], depth_mask.shape[2]
        out = torch.empty(B, H, W, device=locs.device, dtype=torch.float32)
        with torch.cuda.device(locs.device):
            nat.check(nat.lib().spnb_particleprojection_forward(
                nat.ptr(locs), B, N, camera_fl, W, H, filter_std, filter_scale, nat.ptr(depth_mask), nat.ptr(out),
                nat.stream()), "spnb_particleprojection_forward")
        return out

    @staticmethod
    def backward(ctx, grad_output):
        locs, depth_mask = ctx.saved_tensors
        camera_fl, filter_std, filter_scale = ctx.cfg
        B, N, _ = locs.shape
        H, W = depth_mask.shape[1], depth_mask.shape[2]
        dl = None
        if ctx.needs_input_grad[0]:
            grad_output = grad_output.contiguous()
            dl = torch.empty_like(locs)
            with torch.cuda.device(locs.device):
                nat.check(nat.lib().spnb_particleprojection_backward(
                    nat.ptr(locs), B, N, camera_fl, W, H, filter_std, filter_scale, nat.ptr(depth_mask),
                    nat.ptr(grad_output), nat.ptr(dl), nat.stream()), "spnb_particleprojection_backward")
        dm = torch.zeros_like(depth_mask) if ctx.needs_input_grad[1] else None  # ParticleProjection.py:190-191
        return dl, dm, None, None, None
