"""ImageProjection -- bilinear sample of image features at the particles' pixel positions, on libspnb.

Drop-in for python/SmoothParticleNets/ImageProjection.py of the reference: ``ImageProjection(camera_fl)``,
``forward(locs, image, camera_pose, camera_rot, depth_mask=None)`` returning BxNxC, gradients for locs and image
(zero for depth_mask, none for camera_rot; ImageProjection.py:186-209).  NaN checks as in the reference
(ImageProjection.py:101-112, 133-141).  3-D particles only.
"""
import numbers  # noqa: F401

import torch

from . import _native as nat
from . import error_checking as ec
from .particleprojection import MAX_FLOAT, to_camera_space


class ImageProjection(torch.nn.Module):

    def __init__(self, camera_fl):
        super(ImageProjection, self).__init__()
        self.camera_fl = ec.check_conditions(camera_fl, "camera_fl", "%s > 0", "isinstance(%s, numbers.Real)")
        self.register_buffer("empty_depth_mask", torch.ones(1, 1, 1) * MAX_FLOAT)

    def forward(self, locs, image, camera_pose, camera_rot, depth_mask=None):
        """locs BxNx3, image BxCxHxW, camera_pose Bx3, camera_rot Bx4 (xyzw), depth_mask BxHxW or None."""
        batch_size = locs.size()[0]
        N = locs.size()[1]
        width, height, channels = image.size()[3], image.size()[2], image.size()[1]
        ec.check_tensor_dims(locs, "locs", (batch_size, N, 3))
        ec.check_tensor_dims(image, "image", (batch_size, channels, height, width))
        ec.check_tensor_dims(camera_pose, "camera_pose", (batch_size, 3))
        ec.check_tensor_dims(camera_rot, "camera_rot", (batch_size, 4))
        ec.check_nans(locs, "locs")
        ec.check_nans(image, "image")
        ec.check_nans(camera_pose, "camera_pose")
        ec.check_nans(camera_rot, "camera_rot")
        if depth_mask is not None:
            ec.check_tensor_dims(depth_mask, "depth_mask", (batch_size, height, width))
            ec.check_nans(depth_mask, "depth_mask")
            depth_mask = depth_mask.contiguous()
        else:
            if tuple(self.empty_depth_mask.shape) != (batch_size, height, width):
                self.empty_depth_mask = self.empty_depth_mask.new_full((batch_size, height, width), MAX_FLOAT)
            depth_mask = self.empty_depth_mask.to(locs.device)
        cam = to_camera_space(locs, camera_pose, camera_rot)
        if bool((cam != cam).any()):
            raise ValueError("Rotating locs by rotation matrix resulted in NaNs.")
        return _ImageProjectionFunction.apply(cam, image.contiguous(), depth_mask, float(self.camera_fl))


class _ImageProjectionFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, locs, image, depth_mask, camera_fl):
        for t, n in ((locs, "locs"), (image, "image"), (depth_mask, "depth_mask")):
            nat.require_cuda_f32(t, n)
        ctx.save_for_backward(locs, image, depth_mask)
        ctx.camera_fl = camera_fl
        B, N, _ = locs.shape
        C, H, W = image.shape[1:]
        out = torch.empty(B, N, C, device=locs.device, dtype=torch.float32)
        with torch.cuda.device(locs.device):
            nat.check(nat.lib().spnb_imageprojection_forward(
                nat.ptr(locs), nat.ptr(image), B, N, camera_fl, W, H, C, nat.ptr(depth_mask), nat.ptr(out),
                nat.stream()), "spnb_imageprojection_forward")
        return out

    @staticmethod
    def backward(ctx, grad_output):
        locs, image, depth_mask = ctx.saved_tensors
        B, N, _ = locs.shape
        C, H, W = image.shape[1:]
        need_l, need_i, need_m = ctx.needs_input_grad[:3]
        dl = torch.empty_like(locs) if need_l else None
        di = torch.empty_like(image) if need_i else None
        if need_l or need_i:
            grad_output = grad_output.contiguous()
            with torch.cuda.device(locs.device):
                nat.check(nat.lib().spnb_imageprojection_backward(
                    nat.ptr(locs), nat.ptr(image), B, N, ctx.camera_fl, W, H, C, nat.ptr(depth_mask),
                    nat.ptr(grad_output), nat.ptr(dl), nat.ptr(di), nat.stream()),
                    "spnb_imageprojection_backward")
        dm = torch.zeros_like(depth_mask) if need_m else None
        return dl, di, dm, None
