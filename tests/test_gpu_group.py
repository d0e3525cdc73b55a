"""GPU parity of the fused ConvSPGroup path against the per-layer modules and
the oracle: same outputs and gradients within the north_star tolerance, in both backward modes."""
import numpy as np
import pytest
import torch

import cases
import fluidstep
import gpu_util as gu
from smoothparticlenets_b200 import _native as nat

pytestmark = pytest.mark.gpu


def make_layers(spn, D, specs, r):
    layers = []
    for kernel, C, normed in specs:
        c = spn.ConvSP(C, C, D, 1, 1, 0.1, dis_norm=normed, with_params=False, kernel_fn=kernel).cuda()
        c.weight.copy_(gu.dev(r.rand(C, C, 1).astype(np.float32)))
        c.bias.copy_(gu.dev(r.rand(C).astype(np.float32)))
        layers.append(c)
    return layers


def close(a, b, what, k=4):
    b = gu.host(b)
    gu.assert_close(gu.host(a), b, 1e-5, 1e-6 * k * max(1.0, float(np.abs(b).max())), what)


@pytest.mark.parametrize("D", [3, 2])
@pytest.mark.parametrize("mode", ["sym", "atomic"])
def test_group_matches_per_layer(spn, D, mode):
    B, N = 2, 700
    r = cases.rng(3)
    locs, vel, L = cases.fluid_cloud(5, B, N, D=D, density=7640.0 if D == 3 else 600.0)
    coll = spn.ParticleCollision(D, 0.1, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    if mode == "atomic":
        nb = nb.clone()  # drops the symmetry tag -> scatter path
    ones = torch.ones(B, N, 1, device="cuda")
    press = gu.dev(r.rand(B, N, 1).astype(np.float32))
    groups = {
        "A": ([("spiky", 1, False), ("dspiky", D, True), ("dspiky", 1, True), ("cohesion", D, True),
               ("cohesion", 1, True), ("constant", 1, False)], lambda l: [ones, l, ones, l, ones, ones]),
        "B": ([("dspiky", D, True), ("dspiky", 1, True)], lambda l: [l * press, press]),
        "V": ([("spiky", D, False), ("spiky", 1, False)], lambda l: [sv, ones]),
        "C": ([("constant", D, False)], lambda l: [sv]),
    }
    for name, (specs, mk) in groups.items():
        layers = make_layers(spn, D, specs, r)
        group = spn.ConvSPGroup(layers)
        gos = None
        res = {}
        for which in ("fused", "ref"):
            l = sl.detach().clone().requires_grad_(True)
            datas = mk(l)
            n0 = nat.lib().spnb_launch_count()
            outs = group(l, datas, nb) if which == "fused" else tuple(
                lay(l, d, nb) for lay, d in zip(layers, datas))
            if which == "fused":  # pack + one walk, not one launch per layer
                assert nat.lib().spnb_launch_count() - n0 == 2, "group %s did not take the fused path" % name
            if gos is None:
                gos = [torch.rand_like(o) for o in outs]
            torch.autograd.backward(outs, gos)
            res[which] = ([o.detach() for o in outs], l.grad.detach().clone())
        for i, (a, b) in enumerate(zip(res["fused"][0], res["ref"][0])):
            close(a, b, "group %s layer %d forward (%s)" % (name, i, mode))
        close(res["fused"][1], res["ref"][1], "group %s locs.grad (%s)" % (name, mode), k=16)


def test_group_falls_back_when_unsupported(spn):
    """kernel_size 3, trainable weights or an unknown channel layout run the per-layer path."""
    B, N, D = 1, 300, 3
    locs, vel, _ = cases.fluid_cloud(6, B, N)
    coll = spn.ParticleCollision(D, 0.15).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    a = spn.ConvSP(3, 2, D, 3, 0.05, 0.1, kernel_fn="spiky").cuda()
    b = spn.ConvSP(3, 5, D, 1, 1, 0.1, kernel_fn="spiky").cuda()
    for m in (a, b):
        torch.nn.init.uniform_(m.weight)
        torch.nn.init.uniform_(m.bias)
    g = spn.ConvSPGroup([a, b])
    o = g(sl, [sv, sv], nb)
    assert torch.equal(o[0], a(sl, sv, nb)) and torch.equal(o[1], b(sl, sv, nb))
    (o[0].sum() + o[1].sum()).backward()
    assert a.weight.grad is not None and b.weight.grad is not None


def test_fused_fluid_step_matches_layerwise(spn):
    B, N = 2, 4096
    locs, vel, _ = cases.fluid_cloud(9, B, N)
    out = {}
    for fused in (False, True):
        model = fluidstep.FluidStep(spn, fused=fused).cuda()
        l = gu.dev(locs).requires_grad_(True)
        v = gu.dev(vel).requires_grad_(True)
        ol, ov = model(l, v)
        g = torch.Generator(device="cuda").manual_seed(1)
        gl, gv = torch.rand(ol.shape, device="cuda", generator=g), torch.rand(ov.shape, device="cuda", generator=g)
        torch.autograd.backward([ol, ov], [gl, gv])
        out[fused] = (ol.detach(), ov.detach(), l.grad, v.grad)
    for a, b, nm in zip(out[True], out[False], ("locs", "vel", "dlocs", "dvel")):
        # 3 solver iterations chain ~30 fp32 reductions: compare at 1e-4 of the tensor's scale
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-4 * scale, (nm, float((a - b).abs().max()), scale)
