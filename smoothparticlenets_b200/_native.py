"""ctypes binding of libspnb.so (C ABI: include/spnb.h).

This is the only place the Python layer touches native code: it passes ``tensor.data_ptr()``
values, plain sizes and the current CUDA stream handle.  There is NO CPU fallback -- if the library
is missing or an operand is not a contiguous float32 CUDA tensor, the call fails loudly.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPNB_LIB") or os.path.join(_HERE, "libspnb.so")

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

_SIGNATURES = {
    "spnb_version": (ctypes.c_int, []),
    "spnb_last_error": (ctypes.c_char_p, []),
    "spnb_max_cartesian_dim": (ctypes.c_int, []),
    "spnb_launch_count": (ctypes.c_ulonglong, []),
    "spnb_hashgrid_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "spnb_grid_bounds": (_i, [_vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    "spnb_hashgrid_order": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _i, _vp]),
    "spnb_compute_collisions": (_i, [_vp] * 8 + [_i] * 6 + [_f, _f, _i, _vp, _vp]),
    "spnb_tile_lists_bytes": (_sz, [_i, _i, _i, _i]),
    "spnb_compute_collisions_tiled": (_i, [_vp] * 8 + [_i] * 5 + [_f, _f, _i, _vp, _vp, _sz, _vp]),
    "spnb_reorder_data": (_i, [_vp] * 5 + [_i] * 5 + [_vp]),
    "spnb_reorder_data_pos4": (_i, [_vp] * 6 + [_i] * 4 + [_vp]),
    "spnb_convsp_forward": (_i, [_vp] * 6 + [_i] * 8 + [_f, _vp, _vp, _i, _i, _vp, _vp]),
    "spnb_convsp_forward_wide_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "spnb_convsp_backward_block": (_i, [_vp] * 4 + [_i] * 8 + [_f, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "spnb_convsp_backward_wide_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "spnb_convsp_backward_wide": (_i, [_vp] * 5 + [_i] * 8 + [_f, _vp, _vp, _i, _i] + [_vp] * 6 + [_sz, _vp]),
    "spnb_convsp_forward_wide": (_i, [_vp] * 6 + [_i] * 8 + [_f, _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "spnb_convsp_backward_workspace_bytes": (_sz, [_i, _i, _i]),
    "spnb_convsp_backward": (_i, [_vp] * 5 + [_i] * 8 + [_f, _vp, _vp, _i, _i] + [_vp] * 8),
    "spnb_convsp_group_workspace_bytes": (_sz, [_vp, _i, _i, _i, _f, _i, _vp, _i]),
    "spnb_convsp_group_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _sz, _vp, _vp]),
    "spnb_convsp_group_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "spnb_pbf_stage1_forward": (_i, [_vp] * 7 + [ctypes.c_longlong, _i, _f, _f, _vp]),
    "spnb_pbf_stage1_backward": (_i, [_vp] * 10 + [ctypes.c_longlong, _i, _f, _f, _vp]),
    "spnb_pbf_stage2_forward": (_i, [_vp] * 9 + [ctypes.c_longlong, _i, _f, _f, _f, _f, _f, _vp]),
    "spnb_pbf_stage2_backward": (_i, [_vp] * 14 + [ctypes.c_longlong, _i, _f, _f, _f, _f, _f, _vp]),
    "spnb_pbf_stage3_forward": (_i, [_vp] * 6 + [ctypes.c_longlong, _i, _f, _f, _vp]),
    "spnb_pbf_stage3_backward": (_i, [_vp] * 8 + [ctypes.c_longlong, _i, _f, _f, _vp]),
    "spnb_sum_n": (_i, [ctypes.POINTER(_vp), _i, _vp, ctypes.c_longlong, _vp]),
    "spnb_pbf_integrate_forward": (_i, [_vp] * 4 + [ctypes.c_longlong, _i, ctypes.POINTER(_f), _f, _f, _vp]),
    "spnb_pbf_integrate_backward": (_i, [_vp] * 4 + [ctypes.c_longlong, _i, ctypes.POINTER(_f), _f, _f, _vp]),
    "spnb_pbf_velocity": (_i, [_vp] * 4 + [ctypes.c_longlong, _f, _i, _vp]),
    "spnb_pbf_viscosity_forward": (_i, [_vp] * 4 + [ctypes.c_longlong, _i, _f, _vp]),
    "spnb_pbf_viscosity_backward": (_i, [_vp] * 6 + [ctypes.c_longlong, _i, _f, _vp]),
    "spnb_convsdf_forward": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp, _i, _vp,
                                  _vp, _i, _i, _vp, _vp, _f, _vp, _vp]),
    "spnb_convsdf_backward": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp, _i, _vp,
                                   _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp]),
    "spnb_particleprojection_forward": (_i, [_vp, _i, _i, _f, _i, _i, _f, _f, _vp, _vp, _vp]),
    "spnb_particleprojection_backward": (_i, [_vp, _i, _i, _f, _i, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "spnb_imageprojection_forward": (_i, [_vp, _vp, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp]),
    "spnb_imageprojection_backward": (_i, [_vp, _vp, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "spnb_convsdf_backward_analytic": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp,
                                            _vp, _i, _vp, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp,
                                            _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class GroupLayer(ctypes.Structure):
    """struct SpnbGroupLayer (include/spnb.h)."""
    _fields_ = [("data", _vp), ("weight", _vp), ("bias", _vp), ("out", _vp), ("grad_out", _vp),
                ("ddata", _vp), ("nchannels", _i), ("nkernels", _i), ("kernel_fn", _i), ("dis_norm", _i)]

_lib = None


class NativeError(Exception):
    """A libspnb call returned failure (the reference raises a bare Exception("Cuda error"))."""


def lib():
    """The loaded library.  Raises if it has not been built: the product has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libspnb.so is not built (%s). Run `python -m smoothparticlenets_b200.build`; "
                "smoothparticlenets_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is not a CUDA tensor: smoothparticlenets_b200 runs only on the GPU "
                           "(no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, not %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def check(status, what):
    if not status:
        raise NativeError("%s failed: %s" % (what, lib().spnb_last_error().decode()))
