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
    b = gu.host(b) if isinstance(b, torch.Tensor) else np.asarray(b)
    gu.assert_close(gu.host(a), b, 1e-5, 1e-6 * k * max(1.0, float(np.abs(b).max())), what)


@pytest.mark.parametrize("D", [3, 2])
@pytest.mark.parametrize("mode", ["tile", "sym", "atomic"])
def test_group_matches_per_layer(spn, D, mode):
    B, N = 2, 700
    r = cases.rng(3)
    locs, vel, L = cases.fluid_cloud(5, B, N, D=D, density=7640.0 if D == 3 else 600.0)
    coll = spn.ParticleCollision(D, 0.1, include_self=False).cuda()
    coll.tile_lists = mode == "tile"  # tile: compact tile lists; sym/atomic: walk of the float lists
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    if mode == "tile":
        assert int(spn.tile_lists_of(nb)[:4].view(torch.int32).item()) == 0, "tile lists usable"
    if mode == "atomic":
        nb = nb.clone()  # drops the symmetry tag -> scatter path
    ones = torch.ones(B, N, 1, device="cuda")
    press = gu.dev(r.rand(B, N, 1).astype(np.float32))
    groups = {
        "A": ([("spiky", 1, False), ("dspiky", D, True), ("dspiky", 1, True), ("cohesion", D, True),
               ("cohesion", 1, True), ("constant", 1, False)], lambda l: [None, l, None, l, None, None]),
        "B": ([("dspiky", D, True), ("dspiky", 1, True)], lambda l: [l * press, press]),
        "V": ([("spiky", D, False), ("spiky", 1, False)], lambda l: [sv, None]),
        "C": ([("constant", D, False)], lambda l: [sv]),
        # any single layer takes the fused path too (run-time kernel id), here two that are not in the fluid step
        "S1": ([("pressure", 2, False)], lambda l: [sv[..., :2].contiguous()]),
        "S2": ([("sigmoid", D, True)], lambda l: [l]),
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
                lay(l, ones if d is None else d, nb) for lay, d in zip(layers, datas))
            if which == "fused":  # pack + one walk, not one launch per layer
                # pack + one kernel (tile kernel, or the list walk when there are no tile lists)
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
        err = float((a - b).abs().max())
        assert err <= 1e-4 * scale, "%s: max |fused - layerwise| = %g, scale %g" % (nm, err, scale)


def test_group_tile_flag_falls_back_on_device(spn):
    """Lists cut at max_collisions raise the tile flag; the group kernels then run from the float lists,
    decided on the device, with the same results."""
    B, N, D = 1, 1500, 3
    r = cases.rng(4)
    locs = (r.rand(B, N, D) * 0.25).astype(np.float32)  # 1500 particles within 2-3 cells per axis
    vel = r.rand(B, N, D).astype(np.float32)
    layers = make_layers(spn, D, [("spiky", D, False), ("spiky", 1, False)], r)
    group = spn.ConvSPGroup(layers)
    res = {}
    for tiles_on in (True, False):
        coll = spn.ParticleCollision(D, 0.1, max_collisions=128, include_self=False).cuda()
        coll.tile_lists = tiles_on
        sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
        if tiles_on:
            assert int(spn.tile_lists_of(nb)[:4].view(torch.int32).item()) != 0
        l = sl.detach().clone().requires_grad_(True)
        n0 = nat.lib().spnb_launch_count()
        outs = group(l, [sv, None], nb)
        assert nat.lib().spnb_launch_count() - n0 == 2, "fused path"
        torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
        res[tiles_on] = ([o.detach() for o in outs], l.grad.clone())
    for a, b in zip(res[True][0], res[False][0]):
        assert torch.equal(a, b), "the same float-list walk with and without (unusable) tile lists"
    # lists are cut at K here, so the backward is the atomics mode in both runs: order-dependent rounding
    close(res[True][1], res[False][1], "locs.grad", k=64)


@pytest.mark.parametrize("N,K,want_flag", [(3000, 256, 0)])
def test_group_oversized_tiles(spn, oracle, N, K, want_flag):
    """Blocks whose candidate ranges exceed the staged tile (dense cloud, lists not cut).  Up to 4095 candidates
    the tile kernels stage the block's records in several chunks and walk its lists once per chunk; beyond that the
    list kernel computes the rows with the general routine, raises bit 1 of the tile flag and the group kernels
    run the float-list walk.  Rows stay bit-exact, results equal to the float-list walk."""
    from smoothparticlenets_b200 import tile_lists as tl
    B, D = 1, 3
    r = cases.rng(8)
    locs = (r.rand(B, N, D) * 0.5).astype(np.float32)
    vel = r.rand(B, N, D).astype(np.float32)
    layers = make_layers(spn, D, [("spiky", D, False), ("spiky", 1, False)], r)
    group = spn.ConvSPGroup(layers)
    res = {}
    for tiles_on in (True, False):
        coll = spn.ParticleCollision(D, 0.1, max_collisions=K, include_self=False).cuda()
        coll.tile_lists = tiles_on
        sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
        assert int(spn.sym_flag_of(nb).item()) == 0
        if tiles_on:
            flag, counts, dec, max_total = tl.decode(spn.tile_lists_of(nb), B, N, K)
            assert flag == want_flag and max_total + 1 > tl.TILE_CAP
            assert (max_total + 1 > tl.MAX_SLOTS) == bool(want_flag)
            low, gd = oracle.grid_bounds(locs, 0.1, 96)
            ids, oi = oracle.hashgrid_order(locs, low, gd, 0.1, stable=True)
            nl, _ = oracle.reorder_data(locs, None, oi)
            onb, _, _ = oracle.compute_collisions(nl, nl, low, gd, ids, 0.1, 0.1, K, 0, 96 ** 3)
            gu.assert_bit_equal(gu.host(nb), onb, "rows of oversized blocks")
            have = counts >= 0
            assert np.array_equal(dec[have], onb.astype(np.int64)[have])
        l = sl.detach().clone().requires_grad_(True)
        outs = group(l, [sv, None], nb)
        torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
        res[tiles_on] = ([o.detach() for o in outs], l.grad.clone())
    for i, (a, b) in enumerate(zip(res[True][0], res[False][0])):
        close(a, b, "layer %d forward" % i)
    close(res[True][1], res[False][1], "locs.grad", k=16)


def test_pbf_stages_match_torch_expressions(spn):
    """The fused elementwise solver stages (csrc/fluid_glue.cu) against the torch expressions of
    examples/fluid_sim.py:367-397 they replace: forward values and every input gradient."""
    B, N, D = 2, 1000, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    rnd = lambda *s: torch.rand(*s, device="cuda", generator=g)
    k, rho0, coh, rad, st, cs, relax, damp = 0.37, 0.45, 0.1, 0.1, 0.3, 7.0, 1.0, 1.0

    def run(fused):
        x, nj, njp, nj_c, cd = (rnd(B, N, D).requires_grad_(True) for _ in range(5))
        density, ni_s, nip_s, ni_cs = (rnd(B, N, 1).requires_grad_(True) for _ in range(4))
        ncount = (rnd(B, N, 1) * 6).requires_grad_(True)
        ins = [x, nj, njp, nj_c, cd, density, ni_s, nip_s, ni_cs, ncount]
        if fused:
            p, xp, nij = spn.pbf_stage1(x, density, nj, ni_s, k, rho0)
            d0, nrm = spn.pbf_stage2(x, p, nij, njp, nip_s, nj_c, ni_cs, coh, rad, st, rho0, cs)
            xn = spn.pbf_stage3(x, d0, cd, nrm, ncount, relax, damp)
        else:
            relu = torch.nn.functional.relu
            p = k * relu(density - rho0)
            xp = x * p
            nij = x * ni_s - nj
            nijp = x * nip_s - njp
            d0 = -(p * nij + nijp)
            nij2 = x * ni_cs - nj_c
            d0 = d0 + -coh * nij2 * rad
            nrm = nij2 * st / rho0 / cs
            delta = d0 + (cd - nrm * ncount)
            scale = relu(ncount / (1.0 + relax) - damp) + damp
            xn = x + delta / scale
        outs = [p, xp, nij, d0, nrm, xn]
        return ins, outs

    res = {}
    for fused in (True, False):
        g.manual_seed(5)
        ins, outs = run(fused)
        gg = torch.Generator(device="cuda").manual_seed(9)
        gos = [torch.rand(o.shape, device="cuda", generator=gg) for o in outs]
        torch.autograd.backward(outs, gos)
        res[fused] = ([o.detach() for o in outs], [i.grad for i in ins])
    for i, (a, b) in enumerate(zip(res[True][0], res[False][0])):
        close(a, b, "pbf output %d" % i)
    for i, (a, b) in enumerate(zip(res[True][1], res[False][1])):
        close(a, b, "pbf input gradient %d" % i, k=16)


def test_pbf_step_ends_match_torch_expressions(spn):
    """pbf_integrate / pbf_velocity / pbf_viscosity against the torch expressions of fluid_sim.py:355-365 and
    412-424 (forward and input gradients); speeds straddle the cap so both branches of the clamp are hit."""
    B, N, D = 2, 1500, 3
    dt, cap, c = 1.0 / 60, 3.0, 0.37
    grav = [0.0, -9.8, 0.0]
    relu = torch.nn.functional.relu

    def run(fused):
        g = torch.Generator(device="cuda").manual_seed(11)
        rnd = lambda *s: torch.rand(*s, device="cuda", generator=g)
        x, xs, vj = (rnd(B, N, D).requires_grad_(True) for _ in range(3))
        v = ((rnd(B, N, D) - 0.5) * 8.0).requires_grad_(True)
        vi_s = rnd(B, N, 1).requires_grad_(True)
        ins = [x, v, xs, vj, vi_s]
        if fused:
            v2, x1 = spn.pbf_integrate(x, v, grav, dt, cap)
            w0 = spn.pbf_velocity(x1, xs, dt)
            w1 = spn.pbf_viscosity(w0, vj, vi_s, c)
        else:
            gt = torch.tensor(grav, device="cuda").view(1, 1, -1)
            v1 = v + gt * dt
            vv = torch.norm(v1, 2, v1.dim() - 1, keepdim=True)
            vv = cap / (vv + 0.0001)
            vv = -(relu(-vv + 1.0) - 1.0)
            v2 = v1 * vv
            x1 = x + v2 * dt
            w0 = (x1 - xs) / dt
            w1 = w0 + c * (vj - w0 * vi_s)
        outs = [v2, x1, w0, w1]
        gg = torch.Generator(device="cuda").manual_seed(13)
        gos = [torch.rand(o.shape, device="cuda", generator=gg) for o in outs]
        torch.autograd.backward(outs, gos)
        return [o.detach() for o in outs], [i.grad for i in ins]

    a, b = run(True), run(False)
    for i, (p, q) in enumerate(zip(a[0], b[0])):
        close(p, q, "output %d" % i)
    for i, (p, q) in enumerate(zip(a[1], b[1])):
        close(p, q, "input gradient %d" % i, k=64)


@pytest.mark.parametrize("n,shape", [(2, (2, 1000, 3)), (8, (1, 777, 3)), (11, (3, 5, 1)), (1, (1, 10, 3))])
def test_fanout_adds_all_gradients(spn, n, shape):
    """fanout(x, n): n aliases of x whose gradients come back added in one pass (two launches above 8)."""
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.rand(*shape, device="cuda", generator=g).requires_grad_(True)
    ws = [torch.rand(*shape, device="cuda", generator=g) for _ in range(n)]
    n0 = nat.lib().spnb_launch_count()
    outs = spn.fanout(x, n)
    assert len(outs) == n and all(torch.equal(o, x) and o.data_ptr() == x.data_ptr() for o in outs)
    sum((o * w).sum() for o, w in zip(outs, ws)).backward()
    assert nat.lib().spnb_launch_count() - n0 == (0 if n == 1 else 1 if n <= 8 else 2)
    want = ws[0].clone()
    for w in ws[1:]:
        want = want + w
    close(x.grad, want, "fanout gradient")


def test_single_layer_fast_path_option(spn, oracle):
    """ConvSP.fast_path = True: a kernel_size-1 layer on lists with tile lists runs as pack + tile kernel (the
    single-layer signature, kernel id at run time) and matches the oracle like the default float-list kernels."""
    B, N, D, R = 2, 1500, 3, 0.1
    locs, vel, _ = cases.fluid_cloud(31, B, N)
    coll = spn.ParticleCollision(D, R, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    nl, nv, nbh = gu.host(sl), gu.host(sv), gu.host(nb)
    one3 = np.ones(3, np.float32)
    for kernel, normed, C, O in (("default", False, 3, 5), ("indirect", True, 3, 2), ("ddefault2", False, 2, 2)):
        conv = spn.ConvSP(C, O, D, 1, 1, R, dis_norm=normed, kernel_fn=kernel, with_params=False).cuda()
        r = cases.rng(1)
        w = r.rand(O, C, 1).astype(np.float32)
        b = r.rand(O).astype(np.float32)
        conv.weight.copy_(gu.dev(w))
        conv.bias.copy_(gu.dev(b))
        data = nv[..., :C].copy()
        go = r.rand(B, N, O).astype(np.float32)
        want = oracle.convsp_forward(nl, nl, data, nbh, w, b, R, one3, one3, int(normed), kernel)
        dq, dl, dd, _, _ = oracle.convsp_backward(nl, nl, data, nbh, w, b, R, one3, one3, int(normed), kernel, go)
        for fast in (True, False):
            conv.fast_path = fast
            lt = sl.detach().clone().requires_grad_(True)
            dt = gu.dev(data).requires_grad_(True)
            n0 = nat.lib().spnb_launch_count()
            out = conv(lt, dt, nb)
            assert nat.lib().spnb_launch_count() - n0 == (2 if fast else 1)
            close(out, want, "%s fast=%s fwd" % (kernel, fast))
            out.backward(gu.dev(go))
            close(lt.grad, dq.astype(np.float64) + dl, "%s fast=%s dlocs" % (kernel, fast), k=4)
            close(dt.grad, dd, "%s fast=%s ddata" % (kernel, fast), k=4)


def test_trainable_weights_fuse(spn, oracle):
    """Layers with with_params=True (the ConvSP default) take the fused path too: d(weight) = go^T T, with T from one
    more pass of the fused forward with identity weights (common_funcs.h:542-547 is the reference's per-term
    formula).  Checked for a single layer (the per-layer fast path) and for a two-layer group that is not one of
    the fluid signatures' weight shapes (3 -> 5 and 1 -> 2 outputs), against the oracle and float64."""
    from test_gpu_parity_configs import closer_than_reference, convsp_float64
    B, N, D, R = 2, 1200, 3, 0.1
    locs, vel, _ = cases.fluid_cloud(41, B, N)
    coll = spn.ParticleCollision(D, R, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    nl, nv, nbh = gu.host(sl), gu.host(sv), gu.host(nb)
    one3 = np.ones(3, np.float32)
    r = cases.rng(2)
    # --- single layer, default construction (trainable weight and bias)
    conv = spn.ConvSP(3, 5, D, 1, 1, R, dis_norm=False, kernel_fn="spiky").cuda()
    w, b = (r.rand(5, 3, 1) - 0.3).astype(np.float32), r.rand(5).astype(np.float32)
    conv.weight.data.copy_(gu.dev(w))
    conv.bias.data.copy_(gu.dev(b))
    go = r.rand(B, N, 5).astype(np.float32)
    n0 = nat.lib().spnb_launch_count()
    out = conv(sl, sv, nb)
    assert nat.lib().spnb_launch_count() - n0 == 2, "pack + tile kernel"
    n0 = nat.lib().spnb_launch_count()
    out.backward(gu.dev(go))
    assert nat.lib().spnb_launch_count() - n0 == 4, "backward pack + kernel, identity-weight forward pack + kernel"
    close(out, oracle.convsp_forward(nl, nl, nv, nbh, w, b, R, one3, one3, 0, "spiky"), "fwd")
    _, _, _, dw, db = oracle.convsp_backward(nl, nl, nv, nbh, w, b, R, one3, one3, 0, "spiky", go)
    _, w64 = convsp_float64(spn, nl, nl, nv, nbh, w, b, go, R, (1, 1, 1), [1.0] * 3, 0, "spiky")
    closer_than_reference(gu.host(conv.weight.grad), dw, w64, "single-layer dweight")
    close(conv.bias.grad, db, "dbias", k=4)
    # --- a group: 3 -> 5 on the velocities and 1 -> 2 on implicit ones, both trainable
    dens = spn.ConvSP(1, 2, D, 1, 1, R, dis_norm=False, kernel_fn="spiky").cuda()
    w2, b2 = r.rand(2, 1, 1).astype(np.float32), r.rand(2).astype(np.float32)
    dens.weight.data.copy_(gu.dev(w2))
    dens.bias.data.copy_(gu.dev(b2))
    conv.zero_grad()
    group = spn.ConvSPGroup([conv, dens])
    go2 = r.rand(B, N, 2).astype(np.float32)
    lt = sl.detach().clone().requires_grad_(True)
    n0 = nat.lib().spnb_launch_count()
    o1, o2 = group(lt, [sv, None], nb)
    assert nat.lib().spnb_launch_count() - n0 == 2, "one fused pass for both layers (compiled two-layer signature)"
    (o1 * gu.dev(go)).sum().add((o2 * gu.dev(go2)).sum()).backward()
    ones = np.ones((B, N, 1), np.float32)
    _, _, _, dwd, dbd = oracle.convsp_backward(nl, nl, ones, nbh, w2, b2, R, one3, one3, 0, "spiky", go2)
    _, wd64 = convsp_float64(spn, nl, nl, ones, nbh, w2, b2, go2, R, (1, 1, 1), [1.0] * 3, 0, "spiky")
    closer_than_reference(gu.host(conv.weight.grad), dw, w64, "group dweight (layer 1)")
    closer_than_reference(gu.host(dens.weight.grad), dwd, wd64, "group dweight (layer 2)")
    close(dens.bias.grad, dbd, "group dbias", k=4)
    dq1, dl1, _, _, _ = oracle.convsp_backward(nl, nl, nv, nbh, w, b, R, one3, one3, 0, "spiky", go)
    dq2, dl2, _, _, _ = oracle.convsp_backward(nl, nl, ones, nbh, w2, b2, R, one3, one3, 0, "spiky", go2)
    close(lt.grad, dq1.astype(np.float64) + dl1 + dq2 + dl2, "group dlocs", k=8)
