"""GPU parity: ConvSDF forward / backward against the oracle.

Forward sums run in the reference's order, so 3-D / 1-D outputs are expected to agree to the last
bit or two; the bar is the north_star tolerance (1e-5 rel / 1e-6 abs).  2-D goes through
atan2f/sinf/cosf whose CUDA and glibc implementations differ by ulps.
"""
import numpy as np
import pytest
import torch

import cases
import gpu_util as gu
from smoothparticlenets_b200 import _native as nat

pytestmark = pytest.mark.gpu


def native_sdf(c, dil, max_distance, go=None, pose_grads=False):
    L = nat.lib()
    t = {k: gu.dev(v) for k, v in c.items()}
    dilt = gu.dev(dil)
    B, N, D = c["locs"].shape
    S, P = c["idxs"].shape[1], c["poses"].shape[2]
    O, nc = c["weight"].shape
    if go is None:
        out = torch.empty(B, N, O, device="cuda")
        nat.check(L.spnb_convsdf_forward(
            nat.ptr(t["locs"]), B, N, D, nat.ptr(t["idxs"]), nat.ptr(t["poses"]), nat.ptr(t["scales"]),
            S, P, nat.ptr(t["sdfs"]), t["sdfs"].numel(), nat.ptr(t["offs"]), nat.ptr(t["shapes"]),
            c["shapes"].shape[0], nat.ptr(t["weight"]), nat.ptr(t["bias"]), O, nc, nat.ptr(t["ksize"]),
            nat.ptr(dilt), float(max_distance), nat.ptr(out), nat.stream()), "convsdf fwd")
        return gu.host(out)
    dl = torch.full((B, N, D), float("nan"), device="cuda")
    dw = torch.full((O, nc), float("nan"), device="cuda")
    dp = torch.full((B, S, P), float("nan"), device="cuda") if pose_grads else None
    nat.check(L.spnb_convsdf_backward(
        nat.ptr(t["locs"]), B, N, D, nat.ptr(t["idxs"]), nat.ptr(t["poses"]), nat.ptr(t["scales"]),
        S, P, nat.ptr(t["sdfs"]), t["sdfs"].numel(), nat.ptr(t["offs"]), nat.ptr(t["shapes"]),
        c["shapes"].shape[0], nat.ptr(t["weight"]), O, nc, nat.ptr(t["ksize"]), nat.ptr(dilt),
        float(max_distance), nat.ptr(gu.dev(go)), nat.ptr(dl), nat.ptr(dw), nat.ptr(dp), nat.stream()),
        "convsdf bwd")
    return gu.host(dl), gu.host(dw), (gu.host(dp) if pose_grads else None)


@pytest.mark.parametrize("D,ks", [(3, (3, 1, 3)), (3, (1, 1, 1)), (3, (3, 5, 3)), (2, (3, 3)), (1, (3,))])
@pytest.mark.parametrize("max_distance", [0.5, 0.05])
def test_convsdf_vs_oracle(spn, oracle, D, ks, max_distance):
    c = cases.convsdf_case(1, B=2, N=500, D=D, S=4, O=3, ksize=ks)
    dil = np.full(D, 0.01, np.float32)
    a = (c["locs"], c["idxs"], c["poses"], c["scales"], c["sdfs"], c["offs"], c["shapes"], c["weight"],
         c["bias"], c["ksize"], dil, max_distance)
    want = oracle.convsdf_forward(*a)
    got = native_sdf(c, dil, max_distance)
    tol = 1e-6 if D != 2 else 2e-5
    gu.assert_close(got, want, 1e-5, tol, "convsdf fwd D=%d" % D)
    assert (np.abs(want - want.max()) > 1e-6).mean() > 0.1, "case must exercise SDF lookups"

    go = np.random.RandomState(3).rand(*want.shape).astype(np.float32)
    odl, odw, odp, _ = oracle.convsdf_backward(*a, go, pose_grads=True)
    dl, dw, dp = native_sdf(c, dil, max_distance, go, pose_grads=True)
    s = max(1.0, float(np.abs(odl).max()))
    gu.assert_close(dl, odl, 1e-5, (1e-6 if D != 2 else 1e-4) * s, "convsdf dlocs")
    gu.assert_close(dw, odw, 1e-5, 1e-6 * max(1.0, float(np.abs(odw).max())) * 8, "convsdf dweight")
    sp = max(1.0, float(np.abs(odp[..., :D]).max()))
    gu.assert_close(dp[..., :D], odp[..., :D], 1e-5, (1e-6 if D != 2 else 1e-4) * sp * 8,
                    "convsdf dposes (translation)")
    assert np.all(dp[..., D:] == 0)


def test_convsdf_module_and_pose_finite_differences(spn, oracle):
    """Module API incl. SetSDFs packing and the Python finite-difference rotation gradients
    (convsdf.py:211-224), compared with the same recipe run on the oracle."""
    D, ks, md = 3, (1, 1, 1), 0.5
    c = cases.convsdf_case(4, B=2, N=200, D=D, S=3, O=1, ksize=ks)
    # rebuild the individual SDF tensors from the packed atlas to go through SetSDFs
    sdfs, sizes = [], []
    for i in range(c["shapes"].shape[0]):
        shp = c["shapes"][i, :D].astype(int)
        off = int(c["offs"][i])
        sdfs.append(torch.from_numpy(c["sdfs"][off:off + int(np.prod(shp))].reshape(shp).copy()))
        sizes.append(float(c["shapes"][i, D]))
    layer = spn.ConvSDF(sdfs, sizes, 1, D, 1, 0.01, md, with_params=False, compute_pose_grads=True).cuda()
    assert np.array_equal(gu.host(layer.sdf_shapes), c["shapes"])
    assert np.array_equal(gu.host(layer.sdf_offsets), c["offs"])
    assert np.array_equal(gu.host(layer.sdfs), c["sdfs"])
    layer.weight.data.copy_(gu.dev(c["weight"]))
    layer.bias.data.copy_(gu.dev(c["bias"]))
    lt = gu.dev(c["locs"]).requires_grad_(True)
    pt = gu.dev(c["poses"]).requires_grad_(True)
    out = layer(lt, gu.dev(c["idxs"]), pt, gu.dev(c["scales"]))
    go = torch.rand_like(out)
    out.backward(go)
    dil = np.full(D, 0.01, np.float32)

    def ofwd(poses):
        return oracle.convsdf_forward(c["locs"], c["idxs"], poses, c["scales"], c["sdfs"], c["offs"],
                                      c["shapes"], c["weight"], c["bias"], c["ksize"], dil, md)
    base = ofwd(c["poses"])
    gu.assert_close(gu.host(out), base, 1e-5, 1e-6, "module fwd")
    gon = gu.host(go)
    odl, _, odp, _ = oracle.convsdf_backward(c["locs"], c["idxs"], c["poses"], c["scales"], c["sdfs"],
                                             c["offs"], c["shapes"], c["weight"], c["bias"], c["ksize"],
                                             dil, md, gon, pose_grads=True)
    gu.assert_close(gu.host(lt.grad), odl, 1e-5, 1e-6 * max(1, np.abs(odl).max()), "locs.grad")
    got_p = gu.host(pt.grad)
    gu.assert_close(got_p[..., :D], odp[..., :D], 1e-5, 1e-5 * max(1, np.abs(odp).max()), "pose t grad")
    for m in range(c["poses"].shape[1]):
        for i in range(D, c["poses"].shape[2]):
            pp = torch.from_numpy(c["poses"].copy())
            pp[:, m, i] += 1e-3  # float32 perturbation like the reference's in-place add
            gg = (ofwd(pp.numpy()) - base) / np.float32(1e-3)
            want = (gg * gon).sum(axis=(1, 2))
            # finite differences of fp32 outputs: identical recipe, ulp-level differences amplified
            # by 1/eps -> compare with a correspondingly loose absolute tolerance
            np.testing.assert_allclose(got_p[:, m, i], want, rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("D,ks", [(3, (3, 1, 3)), (3, (1, 1, 1)), (2, (3, 3))])
def test_analytic_rotation_pose_grads(spn, D, ks):
    """compute_pose_grads="analytic": rotation columns of poses.grad from the backward kernel
    against the float64 evaluation of the same forward formula -- both its autograd derivative and
    central finite differences (eps 1e-6), which float32 forward differences cannot resolve.
    The finite-difference recipe of the reference (convsdf.py:211-224) stays the default."""
    from sdf_float64 import convsdf_float64
    md = 0.5
    c = cases.convsdf_case(7, B=2, N=400, D=D, S=3, O=2, ksize=ks)
    dil = np.full(D, 0.02, np.float32)
    sdfs, sizes = [], []
    for i in range(c["shapes"].shape[0]):
        shp = c["shapes"][i, :D].astype(int)
        off = int(c["offs"][i])
        sdfs.append(torch.from_numpy(c["sdfs"][off:off + int(np.prod(shp))].reshape(*shp).copy()))
        sizes.append(float(c["shapes"][i, D]))
    layer = spn.ConvSDF(sdfs, sizes, 2, D, list(ks), 0.02, md, compute_pose_grads="analytic").cuda()
    layer.weight.data.copy_(gu.dev(c["weight"]))
    layer.bias.data.copy_(gu.dev(c["bias"]))
    lt = gu.dev(c["locs"]).requires_grad_(True)
    pt = gu.dev(c["poses"]).requires_grad_(True)
    out = layer(lt, gu.dev(c["idxs"]), pt, gu.dev(c["scales"]))
    go = torch.rand_like(out)
    launches = nat.lib().spnb_launch_count()
    out.backward(go)
    assert nat.lib().spnb_launch_count() - launches == 1, "analytic pose gradients need no extra forward passes"
    got = gu.host(pt.grad)

    def f64(poses):
        return convsdf_float64(c["locs"], c["idxs"], poses, c["scales"], c["sdfs"], c["offs"],
                               c["shapes"], c["weight"], c["bias"], c["ksize"], dil, md)
    p64 = torch.from_numpy(c["poses"].astype(np.float64)).requires_grad_(True)
    ref = f64(p64)
    gu.assert_close(gu.host(out), ref.detach().numpy(), 1e-5, 1e-5, "fwd vs float64")
    go64 = go.detach().cpu().double()
    (ref * go64).sum().backward()
    want = p64.grad.numpy()
    scale = float(np.abs(want).max())
    assert scale > 0.1, "degenerate case"
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5 * scale)
    # central differences in float64 on a few rotation components
    eps = 1e-6
    with torch.no_grad():
        for b, m, i in [(0, 0, D), (1, 1, c["poses"].shape[2] - 1), (0, 2, D)]:
            if c["idxs"][b, m] < 0:
                continue
            hi = p64.detach().clone()
            lo = p64.detach().clone()
            hi[b, m, i] += eps
            lo[b, m, i] -= eps
            fd = float((((f64(hi) - f64(lo)) / (2 * eps)) * go64).sum())
            assert abs(fd - want[b, m, i]) <= 1e-5 * scale + 1e-5 * abs(fd), (b, m, i, fd, want[b, m, i])
            assert abs(fd - got[b, m, i]) <= 2e-5 * scale + 1e-4 * abs(fd), (b, m, i, fd, got[b, m, i])


def test_2d_loc_grads(spn):
    """tests/test_convsdf.py:207-229: 2-D 2x2 SDF on a 49x49 lattice, analytic d/dlocs against
    central differences of the layer itself (eps 1e-3, atol 1e-3)."""
    sdfs = [torch.from_numpy(np.array([[0, 0.5], [0.5, 1]], dtype=np.float32))]
    layer = spn.ConvSDF(sdfs, [1], 1, 2, 1, 1, max_distance=1, with_params=False).cuda()
    layer.weight.data.fill_(1)
    layer.bias.data.fill_(0)
    pts = [[x, y] for x in np.arange(0.51, 1.49, 0.02) for y in np.arange(0.51, 1.49, 0.02)]
    locs = torch.tensor([pts], dtype=torch.float32, device="cuda", requires_grad=True)
    idxs = torch.zeros(1, 1, device="cuda")
    poses = torch.zeros(1, 1, 3, device="cuda")
    scales = torch.ones(1, 1, device="cuda")
    out = layer(locs, idxs, poses, scales)
    out.sum().backward()
    eps = 1e-3
    for k in range(2):
        d = torch.zeros_like(locs)
        d[..., k] = eps
        num = (layer(locs.detach() + d, idxs, poses, scales) -
               layer(locs.detach() - d, idxs, poses, scales)) / (2 * eps)
        assert torch.allclose(locs.grad[..., k], num[..., 0], atol=1e-3)
