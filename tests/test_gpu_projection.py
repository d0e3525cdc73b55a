"""GPU parity: ParticleProjection / ImageProjection against the oracle, and the reference's
own tests for the two layers (tests/test_particleprojection.py:59-143, tests/test_imageprojection.py:59-157)
restated: same scene recipe, an independent float64 python ground truth, decimal=3 on the values, gradcheck with
the reference's tolerances."""
import numpy as np
import pytest
import torch

import cases
import gpu_util as gu
from gradcheck_compat import gradcheck
from smoothparticlenets_b200 import _native as nat

pytestmark = pytest.mark.gpu


def _pp(c, std, scale, go=None):
    L = nat.lib()
    locs, dm = gu.dev(c["locs"]), gu.dev(c["depth_mask"])
    B, N, _ = c["locs"].shape
    H, W = c["depth_mask"].shape[1:]
    if go is None:
        out = torch.full((B, H, W), 7.0, device="cuda")  # the entry point zero-fills
        nat.check(L.spnb_particleprojection_forward(nat.ptr(locs), B, N, float(c["fl"]), W, H, std, scale, nat.ptr(dm),
                                                    nat.ptr(out), nat.stream()), "pp fwd")
        return gu.host(out)
    dl = torch.full((B, N, 3), 7.0, device="cuda")
    g = gu.dev(go)
    nat.check(L.spnb_particleprojection_backward(nat.ptr(locs), B, N, float(c["fl"]), W, H, std, scale, nat.ptr(dm),
                                                 nat.ptr(g), nat.ptr(dl), nat.stream()), "pp bwd")
    return gu.host(dl)


@pytest.mark.parametrize("std,scale,W,H", [(2.5, 3.0, 64, 48), (0.8, 1.0 / 0.06, 33, 57), (5.0, 1.0, 120, 90),
                                           (17.0, 2.0, 50, 40)])
def test_particleprojection_vs_oracle(oracle, std, scale, W, H):
    """std = 17: windows wider than the 64 pixels a warp enumerates (single-lane path)."""
    c = cases.projection_case(3, B=3, N=300 if std < 10 else 20, W=W, H=H, fl=W * 0.6)
    want = oracle.particleprojection_forward(c["locs"], float(c["fl"]), std, scale, c["depth_mask"])
    got = _pp(c, std, scale)
    assert (want != 0).sum() > 50
    assert np.array_equal(got != 0, want != 0), "the same pixels are touched"
    # expf differs by ulps between CUDA and glibc, sums of up to N terms in another order
    gu.assert_close(got, want, 1e-5, 1e-6 * float(want.max()), "pp fwd")
    go = cases.rng(9).rand(*want.shape).astype(np.float32)
    wdl = oracle.particleprojection_backward(c["locs"], float(c["fl"]), std, scale, c["depth_mask"], go)
    gdl = _pp(c, std, scale, go)
    gu.assert_close(gdl, wdl, 1e-5, 2e-6 * float(np.abs(wdl).max()), "pp dlocs")
    assert np.array_equal(gdl[c["locs"][..., 2] <= 0], np.zeros_like(gdl[c["locs"][..., 2] <= 0]))


@pytest.mark.parametrize("C,W,H", [(3, 64, 48), (1, 9, 7), (70, 31, 40)])
def test_imageprojection_vs_oracle(oracle, C, W, H):
    c = cases.projection_case(5, B=2, N=400, W=W, H=H, C=C, fl=W * 0.5)
    L = nat.lib()
    locs, im, dm = gu.dev(c["locs"]), gu.dev(c["image"]), gu.dev(c["depth_mask"])
    B, N, _ = c["locs"].shape
    out = torch.full((B, N, C), 7.0, device="cuda")
    nat.check(L.spnb_imageprojection_forward(nat.ptr(locs), nat.ptr(im), B, N, float(c["fl"]), W, H, C, nat.ptr(dm),
                                             nat.ptr(out), nat.stream()), "ip fwd")
    want = oracle.imageprojection_forward(c["locs"], c["image"], float(c["fl"]), c["depth_mask"])
    assert (want != 0).sum() > 20 * C
    assert np.array_equal(gu.host(out), want), "same formula, same order: bit-exact"
    go = cases.rng(10).rand(B, N, C).astype(np.float32)
    wdl, wdi = oracle.imageprojection_backward(c["locs"], c["image"], float(c["fl"]), c["depth_mask"], go)
    dl = torch.full((B, N, 3), 7.0, device="cuda")
    di = torch.full(im.shape, 7.0, device="cuda")
    g = gu.dev(go)
    nat.check(L.spnb_imageprojection_backward(nat.ptr(locs), nat.ptr(im), B, N, float(c["fl"]), W, H, C, nat.ptr(dm),
                                              nat.ptr(g), nat.ptr(dl), nat.ptr(di), nat.stream()), "ip bwd")
    gu.assert_close(gu.host(dl), wdl, 1e-5, 2e-6 * float(np.abs(wdl).max()), "ip dlocs")
    gu.assert_close(gu.host(di), wdi, 1e-5, 2e-6 * float(np.abs(wdi).max()), "ip dimage")
    # only one of the two gradients
    di2 = torch.full(im.shape, 7.0, device="cuda")
    nat.check(L.spnb_imageprojection_backward(nat.ptr(locs), nat.ptr(im), B, N, float(c["fl"]), W, H, C, nat.ptr(dm),
                                              nat.ptr(g), None, nat.ptr(di2), nat.stream()), "ip bwd (image only)")
    gu.assert_close(gu.host(di2), wdi, 1e-5, 2e-6 * float(np.abs(wdi).max()), "ip dimage only")


# ---- the reference's module tests, restated -------------------------------------------------------------------
def _look_at(pose):
    """Quaternion (xyzw) of a camera at `pose` looking at the origin: +Z out of the camera, +Y down, +X right."""
    z = -pose / np.linalg.norm(pose)
    x = np.cross(np.array([0.0, -1.0, 0.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.array([x, y, z]).T
    w = np.sqrt(max(0.0, 1.0 + m[0, 0] + m[1, 1] + m[2, 2])) / 2.0
    if w > 1e-6:
        q = [(m[2, 1] - m[1, 2]) / (4 * w), (m[0, 2] - m[2, 0]) / (4 * w), (m[1, 0] - m[0, 1]) / (4 * w), w]
    else:  # half-turn: take the largest diagonal element
        i = int(np.argmax(np.diag(m)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
        q = [0.0, 0.0, 0.0, (m[k, j] - m[j, k]) / s]
        q[i], q[j], q[k] = s / 4, (m[j, i] + m[i, j]) / s, (m[k, i] + m[i, k]) / s
    return np.array(q, np.float32)


def _camera_space(locs, pose, rot, dtype):
    """conj(q) * (p - pose) * q, written out (float64 unless dtype says otherwise)."""
    out = np.zeros(locs.shape, dtype)
    for b in range(locs.shape[0]):
        x, y, z, w = [dtype(v) for v in rot[b] / np.sqrt((rot[b].astype(np.float64) ** 2).sum())]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype)
        out[b] = (locs[b].astype(dtype) - pose[b].astype(dtype)) @ R  # R^T applied to column vectors
    return out


def _scene(cam_size, N=5, B=2):
    np.random.seed(1)
    fl = cam_size[0] / 2 / (45.0 / 180.0 * np.pi / 2.0)
    pose = 5.0 * (np.random.rand(B, 3).astype(np.float32) - 0.5)
    rot = np.stack([_look_at(pose[b].astype(np.float64)) for b in range(B)])
    locs = 2.0 * (np.random.rand(B, N, 3).astype(np.float32) - 0.5)
    W, H = cam_size
    dm = np.full((B, H, W), np.finfo(np.float32).max, np.float32)
    i0, i1 = int(W / 2 - W * 0.2), int(W / 2 + W * 0.2) + 1
    j0, j1 = int(H / 2 - H * 0.2), int(H / 2 + H * 0.2) + 1
    for i in range(i0, i1):
        for j in range(j0, j1):
            u, v = (i - i0) / (i1 - i0), (j - j0) / (j1 - j0)
            dm[0, j, i] = (0.0 * (1 - v) + 3.5 * v) * (1 - u) + (5.0 * (1 - v) + 10.0 * v) * u
    return fl, pose, rot, locs, dm


def _splat(fl, cam_size, std, scale, locs, pose, rot, dm=None, dtype=np.float64):
    W, H = cam_size
    B, N, _ = locs.shape
    out = np.zeros((B, H, W), dtype)
    cam = _camera_space(locs, pose, rot, dtype)
    s = np.ceil(std * 2)
    f = scale / (std * np.sqrt(2 * np.pi))
    for b in range(B):
        for n in range(N):
            t = cam[b, n]
            if t[2] <= 0:
                continue
            px, py = t[0] * fl / t[2] + W / 2.0, t[1] * fl / t[2] + H / 2.0
            for i in np.arange(max(0, px - s), min(W, px + s + 1), 1):
                for j in np.arange(max(0, py - s), min(H, py + s + 1), 1):
                    if dm is not None and dm[b, int(j), int(i)] < t[2]:
                        continue
                    d2 = (int(i) + 0.5 - px) ** 2 + (int(j) + 0.5 - py) ** 2
                    if d2 <= s * s:
                        out[b, int(j), int(i)] += f * np.exp(-d2 / (2.0 * std * std))
    return out


def test_particleprojection_reference_test(spn):
    cam_size, std, scale = (120, 90), 5, 1.0 / 0.06
    fl, pose, rot, locs, dm = _scene(cam_size)
    layer = spn.ParticleProjection(fl, cam_size, std, scale).cuda()
    lt = gu.dev(locs).requires_grad_(True)
    pt, rt = gu.dev(pose), gu.dev(rot)
    pred = layer(lt, pt, rt, gu.dev(dm))
    truth = _splat(fl, cam_size, std, scale, locs, pose, rot, dm)
    assert truth.max() > 1.0
    np.testing.assert_array_almost_equal(gu.host(pred), truth, decimal=3)

    def func_numerical(l):
        return torch.from_numpy(_splat(fl, cam_size, std, scale, l.detach().cpu().numpy(), pose, rot)).cuda()

    def func_analytical(l):
        return layer(l, pt, rt)
    assert gradcheck(func_analytical, (lt,), eps=1e-6, atol=1e-3, rtol=1e-2, func_numerical=func_numerical,
                     use_double=True)
    # the camera translation gets its gradient through the torch transform: d/dpose = -sum_n d/dlocs
    p2 = gu.dev(pose).requires_grad_(True)
    l2 = gu.dev(locs).requires_grad_(True)
    layer(l2, p2, rt).sum().backward()
    gu.assert_close(gu.host(p2.grad), -gu.host(l2.grad).sum(1), 1e-4, 1e-4 * float(l2.grad.abs().max()), "dpose")


def _sample(locs, image, fl, pose, rot, dm=None, dtype=np.float64):
    B, N, _ = locs.shape
    C, H, W = image.shape[1:]
    out = np.zeros((B, N, C), dtype)
    cam = _camera_space(locs, pose, rot, dtype)
    for b in range(B):
        for n in range(N):
            t = cam[b, n]
            if t[2] <= 0:
                continue
            px, py = t[0] * fl / t[2] + W / 2.0, t[1] * fl / t[2] + H / 2.0
            if px <= 0.5 or px >= W - 0.5 or py <= 0.5 or py >= H - 0.5:
                continue
            if dm is not None and 0 < dm[b, int(py), int(px)] < t[2]:
                continue
            li, lj = int(px - 0.5), int(py - 0.5)
            di, dj = px - 0.5 - li, py - 0.5 - lj
            im = image[b].astype(dtype)
            out[b, n] = (im[:, lj, li] * (1 - di) * (1 - dj) + im[:, lj + 1, li] * (1 - di) * dj +
                         im[:, lj, li + 1] * di * (1 - dj) + im[:, lj + 1, li + 1] * di * dj)
    return out


def test_imageprojection_reference_test(spn):
    cam_size, C = (30, 30), 2
    fl, pose, rot, locs, dm = _scene(cam_size)
    image = np.random.rand(2, C, cam_size[1], cam_size[0]).astype(np.float32)
    layer = spn.ImageProjection(fl).cuda()
    lt = gu.dev(locs).requires_grad_(True)
    it = gu.dev(image).requires_grad_(True)
    pt, rt = gu.dev(pose), gu.dev(rot)
    pred = layer(lt, it, pt, rt, gu.dev(dm))
    truth = _sample(locs, image, fl, pose, rot, dm)
    assert (truth != 0).any()
    np.testing.assert_array_almost_equal(gu.host(pred), truth, decimal=3)

    def func_numerical(l, i):
        return torch.from_numpy(_sample(l.detach().cpu().numpy(), i.detach().cpu().numpy(), fl, pose, rot)).cuda()

    def func_analytical(l, i):
        return layer(l, i, pt, rt)
    assert gradcheck(func_analytical, (lt, it), eps=1e-6, atol=1e-3, rtol=1e-2, func_numerical=func_numerical,
                     use_double=True)
    with pytest.raises(ValueError):
        layer(lt * float("nan"), it, pt, rt)


def test_projection_launches_and_errors(spn):
    layer = spn.ParticleProjection(30.0, (32, 24), 1.5, 1.0).cuda()
    locs = torch.rand(2, 100, 3, device="cuda") + torch.tensor([-0.5, -0.5, 0.5], device="cuda")
    pose, rot = torch.zeros(2, 3, device="cuda"), torch.tensor([[0.0, 0, 0, 1]] * 2, device="cuda")
    n0 = nat.lib().spnb_launch_count()
    out = layer(locs, pose, rot)
    assert nat.lib().spnb_launch_count() - n0 == 1 and out.shape == (2, 24, 32) and float(out.sum()) > 0
    with pytest.raises(ValueError):
        layer(locs[:, :, :2], pose, rot)
    with pytest.raises((ValueError, TypeError)):
        spn.ParticleProjection(-1.0, (32, 24), 1.5, 1.0)
    with pytest.raises(Exception):
        layer(locs.cpu(), pose.cpu(), rot.cpu())  # no CPU fallback
