"""Python face of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two interchangeable back ends with the same numpy-in / numpy-out methods:

* ``COracle``   -- oracle/spn_oracle.c (the plain-C restatement), loaded with ctypes.
* ``RefOracle`` -- oracle/_ref/_ext*.so, the UNMODIFIED reference CPU extension
  (/root/reference/src/cpu_layer_funcs.cpp compiled in place by oracle/build_ref.py), called with
  the allocation / pre-fill conventions of the reference's autograd Functions
  (convsp.py:155-203, ParticleCollision.py:222-314, convsdf.py:166-233).

tests/test_oracle_pinned.py requires both to agree bit-for-bit; the GPU parity tests then use
``COracle`` (which also travels to the GPU box as source and is compiled there by build()).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "libspn_oracle.so")

KERNEL_NAMES = ["cohesion", "constant", "ddefault", "ddefault2", "default", "dpressure",
                "dpressure2", "dspiky", "indirect", "pressure", "sigmoid", "spiky"]

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build_c_oracle(force=False):
    src = os.path.join(HERE, "spn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off",
                               "-w", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def kernel_id(kernel_fn):
    return kernel_fn if isinstance(kernel_fn, int) else KERNEL_NAMES.index(kernel_fn)


class COracle(object):
    """ctypes binding of oracle/spn_oracle.c."""
    kind = "port"

    def __init__(self):
        self.lib = ctypes.CDLL(build_c_oracle())
        L = self.lib
        L.spno_kernel_w.restype = ctypes.c_float
        L.spno_kernel_w.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
        L.spno_kernel_dw.restype = ctypes.c_float
        L.spno_kernel_dw.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
        for name in ("spno_grid_bounds", "spno_cell_keys", "spno_hashgrid_order_selection",
                     "spno_hashgrid_order_stable", "spno_reorder_data", "spno_cell_table",
                     "spno_compute_collisions", "spno_convsp", "spno_convsdf", "spno_particleprojection",
                     "spno_imageprojection"):
            getattr(L, name).restype = None

    # ---- SPH kernels -------------------------------------------------------------------------
    def kernel_w(self, d, H, fn):
        return self.lib.spno_kernel_w(ctypes.c_float(d), ctypes.c_float(H), kernel_id(fn))

    def kernel_dw(self, d, H, fn):
        return self.lib.spno_kernel_dw(ctypes.c_float(d), ctypes.c_float(H), kernel_id(fn))

    # ---- hash grid ---------------------------------------------------------------------------
    def grid_bounds(self, locs, radius, max_grid_dim):
        locs = _f(locs)
        B, N, D = locs.shape
        low = np.empty((B, D), np.float32)
        dims = np.empty((B, D), np.float32)
        self.lib.spno_grid_bounds(_p(locs), B, N, D, ctypes.c_float(radius), int(max_grid_dim),
                                  _p(low), _p(dims))
        return low, dims

    def cell_keys(self, locs, low, dims, edge):
        locs, low, dims = _f(locs), _f(low), _f(dims)
        B, N, D = locs.shape
        keys = np.empty((B, N), np.int32)
        self.lib.spno_cell_keys(_p(locs), _p(low), _p(dims), B, N, D, ctypes.c_float(edge),
                                keys.ctypes.data_as(_i32p))
        return keys

    def hashgrid_order(self, locs, low, dims, edge, stable=True):
        """Returns (sorted cellIDs [B,N], idxs [B,N]) as float32.  stable=False is the CPU
        reference's selection sort; stable=True is the reference GPU path's contract."""
        locs, low, dims = _f(locs), _f(low), _f(dims)
        B, N, D = locs.shape
        ids = np.empty((B, N), np.float32)
        idxs = np.empty((B, N), np.float32)
        fn = self.lib.spno_hashgrid_order_stable if stable else self.lib.spno_hashgrid_order_selection
        fn(_p(locs), _p(low), _p(dims), B, N, D, ctypes.c_float(edge), _p(ids), _p(idxs))
        return ids, idxs

    def reorder_data(self, locs, data, idxs, reverse=0):
        locs, idxs = _f(locs), _f(idxs)
        B, N, D = locs.shape
        nlocs = np.empty_like(locs)
        C = 0
        ndata = None
        if data is not None:
            data = _f(data)
            C = data.shape[2]
            ndata = np.empty_like(data)
        self.lib.spno_reorder_data(_p(locs), _p(data), _p(idxs), _p(nlocs), _p(ndata), B, N, D, C,
                                   int(reverse))
        return nlocs, ndata

    def compute_collisions(self, qlocs, locs, low, dims, sorted_ids, edge, radius, max_collisions,
                           include_self, ncells):
        """Returns (neighbors [B,M,K], cellStarts [B,ncells], cellEnds [B,ncells])."""
        qlocs, locs, low, dims, sorted_ids = _f(qlocs), _f(locs), _f(low), _f(dims), _f(sorted_ids)
        B, N, D = locs.shape
        M = qlocs.shape[1]
        starts = np.zeros((B, ncells), np.float32)
        ends = np.zeros((B, ncells), np.float32)
        self.lib.spno_cell_table(_p(sorted_ids), B, N, int(ncells), _p(starts), _p(ends))
        coll = np.full((B, M, max_collisions), -1, np.float32)
        self.lib.spno_compute_collisions(_p(qlocs), _p(locs), _p(low), _p(dims), _p(starts), _p(ends),
                                         B, M, N, D, int(ncells), ctypes.c_float(edge),
                                         ctypes.c_float(radius), _p(coll), int(max_collisions),
                                         int(include_self))
        return coll, starts, ends

    # ---- ConvSP ------------------------------------------------------------------------------
    def _convsp(self, qlocs, locs, data, neighbors, weight, radius, ksize, dil, dis_norm, kernel_fn,
                out, dq, dl, dd, dw):
        B, N, D = locs.shape
        M = qlocs.shape[1]
        C = data.shape[2]
        K = neighbors.shape[2]
        O, _, ncells = weight.shape
        self.lib.spno_convsp(_p(qlocs), _p(locs), _p(data), _p(neighbors), _p(weight), B, M, N, C, D,
                             K, O, ncells, ctypes.c_float(radius), _p(ksize), _p(dil), int(dis_norm),
                             kernel_id(kernel_fn), _p(out), _p(dq), _p(dl), _p(dd), _p(dw))

    def convsp_forward(self, qlocs, locs, data, neighbors, weight, bias, radius, ksize, dil,
                       dis_norm, kernel_fn):
        qlocs, locs, data, neighbors, weight = map(_f, (qlocs, locs, data, neighbors, weight))
        ksize, dil = _f(ksize), _f(dil)
        out = np.zeros((locs.shape[0], qlocs.shape[1], weight.shape[0]), np.float32)
        self._convsp(qlocs, locs, data, neighbors, weight, radius, ksize, dil, dis_norm, kernel_fn,
                     out, None, None, None, None)
        out += _f(bias).reshape(1, 1, -1)  # convsp.py:172
        return out

    def convsp_backward(self, qlocs, locs, data, neighbors, weight, bias, radius, ksize, dil,
                        dis_norm, kernel_fn, grad_out):
        """Returns (dqlocs, dlocs, ddata, dweight, dbias)."""
        qlocs, locs, data, neighbors, weight = map(_f, (qlocs, locs, data, neighbors, weight))
        ksize, dil, go = _f(ksize), _f(dil), _f(grad_out)
        dq, dl, dd, dw = (np.zeros_like(qlocs), np.zeros_like(locs), np.zeros_like(data),
                          np.zeros_like(weight))
        self._convsp(qlocs, locs, data, neighbors, weight, radius, ksize, dil, dis_norm, kernel_fn,
                     go, dq, dl, dd, dw)
        return dq, dl, dd, dw, go.sum(1).sum(0)  # convsp.py:203

    # ---- ConvSDF -----------------------------------------------------------------------------
    def _convsdf(self, locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize, dil,
                 max_distance, out, dl, dw, dp):
        B, N, D = locs.shape
        S = idxs.shape[1]
        pose_len = poses.shape[2]
        O, ncells = weight.shape
        self.lib.spno_convsdf(_p(locs), B, N, D, _p(idxs), _p(poses), _p(scales), S, pose_len,
                              _p(sdfs), _p(offs), _p(shapes), _p(weight), _p(bias), O, ncells,
                              _p(ksize), _p(dil), ctypes.c_float(max_distance), _p(out), _p(dl),
                              _p(dw), _p(dp))

    def convsdf_forward(self, locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize,
                        dil, max_distance):
        args = list(map(_f, (locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize, dil)))
        out = np.zeros((args[0].shape[0], args[0].shape[1], args[7].shape[0]), np.float32)
        self._convsdf(*args, max_distance, out, None, None, None)
        return out

    def convsdf_backward(self, locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize,
                         dil, max_distance, grad_out, pose_grads=False):
        """Returns (dlocs, dweight, dposes-or-None, dbias).  dposes holds the analytic translation
        columns only (rotation columns are finite differences in the Python layer)."""
        args = list(map(_f, (locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize, dil)))
        go = _f(grad_out).copy()
        dl = np.zeros_like(args[0])
        dw = np.zeros_like(args[7])
        dp = np.zeros_like(args[2]) if pose_grads else None
        self._convsdf(*args, max_distance, go, dl, dw, dp)
        return dl, dw, dp, go.sum(1).sum(0)


    # ---- ParticleProjection / ImageProjection (camera-space particles) ------------------------------
    def particleprojection_forward(self, locs, camera_fl, filter_std, filter_scale, depth_mask):
        locs, dm = _f(locs), _f(depth_mask)
        B, N, _ = locs.shape
        H, W = dm.shape[1], dm.shape[2]
        out = np.zeros((B, H, W), np.float32)
        self.lib.spno_particleprojection(_p(locs), B, N, ctypes.c_float(camera_fl), W, H,
                                         ctypes.c_float(filter_std), ctypes.c_float(filter_scale), _p(dm),
                                         _p(out), None, None)
        return out

    def particleprojection_backward(self, locs, camera_fl, filter_std, filter_scale, depth_mask, grad_out):
        locs, dm, go = _f(locs), _f(depth_mask), _f(grad_out)
        B, N, _ = locs.shape
        H, W = dm.shape[1], dm.shape[2]
        dl = np.zeros_like(locs)
        self.lib.spno_particleprojection(_p(locs), B, N, ctypes.c_float(camera_fl), W, H,
                                         ctypes.c_float(filter_std), ctypes.c_float(filter_scale), _p(dm),
                                         None, _p(go), _p(dl))
        return dl

    def imageprojection_forward(self, locs, image, camera_fl, depth_mask):
        locs, image, dm = _f(locs), _f(image), _f(depth_mask)
        B, N, _ = locs.shape
        C, H, W = image.shape[1:]
        out = np.zeros((B, N, C), np.float32)
        self.lib.spno_imageprojection(_p(locs), _p(image), B, N, ctypes.c_float(camera_fl), W, H, C, _p(dm),
                                      _p(out), None, None, None)
        return out

    def imageprojection_backward(self, locs, image, camera_fl, depth_mask, grad_out):
        locs, image, dm, go = _f(locs), _f(image), _f(depth_mask), _f(grad_out)
        B, N, _ = locs.shape
        C, H, W = image.shape[1:]
        dl, di = np.zeros_like(locs), np.zeros_like(image)
        self.lib.spno_imageprojection(_p(locs), _p(image), B, N, ctypes.c_float(camera_fl), W, H, C, _p(dm),
                                      None, _p(go), _p(dl), _p(di))
        return dl, di


class RefOracle(object):
    """The unmodified reference CPU extension (oracle/_ref), same methods as COracle."""
    kind = "reference"

    def __init__(self):
        from . import build_ref
        self.ext = build_ref.load()
        if self.ext is None:
            raise RuntimeError("oracle/_ref is not built (run `python oracle/build_ref.py` where "
                               "/root/reference exists)")
        import torch
        self.torch = torch

    def _t(self, a):
        return self.torch.from_numpy(_f(a).copy())

    def hashgrid_order(self, locs, low, dims, edge, stable=False):
        if stable:
            raise ValueError("the reference CPU path only has the selection sort")
        t = self.torch
        locs_t, low_t, dims_t = self._t(locs), self._t(low), self._t(dims)
        B, N, _ = locs_t.shape
        ids = t.zeros(B, N)
        idxs = t.zeros(B, N)
        self.ext.spn_hashgrid_order(locs_t, low_t, dims_t, ids, idxs, float(edge))
        return ids.numpy(), idxs.numpy()

    def reorder_data(self, locs, data, idxs, reverse=0):
        t = self.torch
        locs_t, idxs_t = self._t(locs), self._t(idxs)
        nlocs = t.zeros_like(locs_t)
        if data is not None:
            data_t = self._t(data)
            ndata = t.zeros_like(data_t)
        else:
            data_t, ndata = t.Tensor(), t.Tensor()
        self.ext.spn_reorder_data(locs_t, data_t, idxs_t, nlocs, ndata, int(reverse))
        return nlocs.numpy(), (ndata.numpy() if data is not None else None)

    def compute_collisions(self, qlocs, locs, low, dims, sorted_ids, edge, radius, max_collisions,
                           include_self, ncells):
        t = self.torch
        q, l, lo, gd, ids = map(self._t, (qlocs, locs, low, dims, sorted_ids))
        B, M = q.shape[0], q.shape[1]
        starts = t.zeros(B, ncells)
        ends = t.zeros(B, ncells)
        coll = t.full((B, M, max_collisions), -1.0)
        self.ext.spn_compute_collisions(q, l, lo, gd, ids, starts, ends, coll, float(edge),
                                        float(radius), int(include_self))
        return coll.numpy(), starts.numpy(), ends.numpy()

    def convsp_forward(self, qlocs, locs, data, neighbors, weight, bias, radius, ksize, dil,
                       dis_norm, kernel_fn):
        t = self.torch
        a = list(map(self._t, (qlocs, locs, data, neighbors, weight, bias)))
        out = t.zeros(a[1].shape[0], a[0].shape[1], a[4].shape[0])
        self.ext.spn_convsp_forward(*a, float(radius), self._t(ksize), self._t(dil), int(dis_norm),
                                    kernel_id(kernel_fn), out)
        out += a[5].view(1, 1, -1)
        return out.numpy()

    def convsp_backward(self, qlocs, locs, data, neighbors, weight, bias, radius, ksize, dil,
                        dis_norm, kernel_fn, grad_out):
        t = self.torch
        a = list(map(self._t, (qlocs, locs, data, neighbors, weight, bias)))
        go = self._t(grad_out)
        dq, dl, dd, dw = t.zeros_like(a[0]), t.zeros_like(a[1]), t.zeros_like(a[2]), t.zeros_like(a[4])
        self.ext.spn_convsp_backward(*a, float(radius), self._t(ksize), self._t(dil), int(dis_norm),
                                     kernel_id(kernel_fn), go, dq, dl, dd, dw)
        return dq.numpy(), dl.numpy(), dd.numpy(), dw.numpy(), go.sum(1).sum(0).numpy()

    def convsdf_forward(self, locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize,
                        dil, max_distance):
        t = self.torch
        a = list(map(self._t, (locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize, dil)))
        out = t.zeros(a[0].shape[0], a[0].shape[1], a[7].shape[0])
        self.ext.spn_convsdf_forward(*a, float(max_distance), out)
        return out.numpy()

    def convsdf_backward(self, locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize,
                         dil, max_distance, grad_out, pose_grads=False):
        t = self.torch
        a = list(map(self._t, (locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize, dil)))
        go = self._t(grad_out)
        dl, dw = t.zeros_like(a[0]), t.zeros_like(a[7])
        if pose_grads:
            # The reference writes dposes row "smallest_m = -1" (common_funcs.h:805-820) when no SDF
            # is closer than max_distance: one row BEFORE the buffer for b = 0.  It only ever adds
            # +-0 there, but that is still an out-of-bounds write, so give it a spare leading row.
            B_, S_, P_ = a[2].shape
            dp_store = t.zeros(B_ * S_ + 1, P_)
            dp = dp_store[1:].view(B_, S_, P_)
        else:
            dp = t.zeros(a[2].shape[0] + 1)  # "disabled" marker, convsdf.py:192-196
        self.ext.spn_convsdf_backward(*a, float(max_distance), go, dl, dw, dp)
        return dl.numpy(), dw.numpy(), (dp.numpy() if pose_grads else None), go.sum(1).sum(0).numpy()


    def particleprojection_forward(self, locs, camera_fl, filter_std, filter_scale, depth_mask):
        l, dm = self._t(locs), self._t(depth_mask)
        out = self.torch.zeros(dm.shape)
        self.ext.spn_particleprojection_forward(l, float(camera_fl), float(filter_std), float(filter_scale), dm, out)
        return out.numpy()

    def particleprojection_backward(self, locs, camera_fl, filter_std, filter_scale, depth_mask, grad_out):
        l, dm, go = self._t(locs), self._t(depth_mask), self._t(grad_out)
        dl = self.torch.zeros(l.shape)
        self.ext.spn_particleprojection_backward(l, float(camera_fl), float(filter_std), float(filter_scale), dm,
                                                 go, dl)
        return dl.numpy()

    def imageprojection_forward(self, locs, image, camera_fl, depth_mask):
        l, im, dm = self._t(locs), self._t(image), self._t(depth_mask)
        out = self.torch.zeros(l.shape[0], l.shape[1], im.shape[1])
        self.ext.spn_imageprojection_forward(l, im, float(camera_fl), dm, out)
        return out.numpy()

    def imageprojection_backward(self, locs, image, camera_fl, depth_mask, grad_out):
        l, im, dm, go = self._t(locs), self._t(image), self._t(depth_mask), self._t(grad_out)
        dl, di = self.torch.zeros(l.shape), self.torch.zeros(im.shape)
        self.ext.spn_imageprojection_backward(l, im, float(camera_fl), dm, go, dl, di)
        return dl.numpy(), di.numpy()


def grid_bounds_torch(locs, radius, max_grid_dim):
    """The reference's own bounds code (ParticleCollision.py:174-181) on float32 CPU tensors."""
    import torch
    locs = torch.from_numpy(_f(locs))
    lower_bounds, _ = locs.min(1)
    upper_bounds, _ = locs.max(1)
    grid_dims = torch.ceil(torch.clamp((upper_bounds - lower_bounds) / radius, 0, max_grid_dim))
    center = (lower_bounds + upper_bounds) / 2
    lower_bounds = center - grid_dims * radius / 2
    return lower_bounds.contiguous().numpy(), grid_dims.contiguous().numpy()
