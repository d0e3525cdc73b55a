"""GPU parity at the shapes BASELINE.json names (c1..c4) and of the benchmarked object itself.

  * c2: one full 65 536-particle scene (K = 128): permutation, reordered tensors and neighbour rows bit-exact
    against the oracle fed with OUR permutation; then every fluid layer type forward + backward, through the
    per-layer modules AND through the fused ConvSPGroup path, against oracle.convsp_*;
  * the fluid step that bench.py times -- FluidStep(spn, fused=True) and fused=False -- against the same step
    run on the reference's CPU functions (oracle/cpu_modules.py), outputs and input gradients;
  * c4: ConvSDF with 16 objects of 64^3 cells;  c3: 64 -> 64 channels, kernel_size 5, n-bar ~ 64 on a query subset;
  * d(weight): compared with a float64 accumulation of the same terms -- the CUDA result must be at least as
    close to it as the reference's own sequential fp32 sum is.

Tolerance everywhere: 1e-5 relative + 1e-6 * max|reference| absolute (north_star), unless a test says otherwise and
why.  The oracle is the checker only (oracle/__init__.py).
"""
import numpy as np
import pytest
import torch

import cases
import fluidstep
import gpu_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(got, want, what, k=1.0, rtol=RTOL):
    want = np.asarray(want)
    gu.assert_close(gu.host(got) if isinstance(got, torch.Tensor) else got, want, rtol,
                    1e-6 * k * max(1.0, float(np.abs(want).max())), what)


def identity_layer(spn, kernel, C, D, normed, radius=0.1):
    conv = spn.ConvSP(C, C, D, 1, 1, radius, dis_norm=normed, with_params=False, kernel_fn=kernel).cuda()
    conv.weight.copy_(torch.eye(C, device="cuda").view(C, C, 1))
    conv.bias.zero_()
    return conv


# ---------------------------------------------------------------------------------------------------------
# c2 at full size
# ---------------------------------------------------------------------------------------------------------
def test_c2_full_scene_vs_oracle(spn, oracle):
    B, N, D, R, K, G = 1, 65536, 3, 0.1, 128, 96
    locs, vel, _ = cases.fluid_cloud(4242, B, N)
    coll = spn.ParticleCollision(D, R, max_collisions=K, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))

    low, gd = oracle.grid_bounds(locs, R, G)
    ids, oi = oracle.hashgrid_order(locs, low, gd, R, stable=True)
    assert np.array_equal(gu.host(idxs), oi), "permutation == stable order of the oracle's keys"
    nl, nv = oracle.reorder_data(locs, vel, gu.host(idxs))
    gu.assert_bit_equal(gu.host(sl), nl, "reordered locs")
    gu.assert_bit_equal(gu.host(sv), nv, "reordered data")
    onb, _, _ = oracle.compute_collisions(nl, nl, low, gd, ids, R, R, K, 0, G ** D)
    gu.assert_bit_equal(gu.host(nb), onb, "neighbour rows (order, truncation, -1 padding)")
    nbar = float((onb >= 0).sum()) / (B * N)
    assert 25 < nbar < 36

    r = cases.rng(7)
    scal = np.ones((B, N, 1), np.float32)  # the one-channel layers of the fluid step all run on ones
    one3 = np.ones(D, np.float32)
    want = {}
    layers = {}
    for kernel, dim, normed in fluidstep.LAYER_TYPES:
        C = D if dim == 'D' else 1
        data = nv if C == D else scal
        w = np.eye(C, dtype=np.float32).reshape(C, C, 1)
        b0 = np.zeros(C, np.float32)
        go = cases.rng(11 + C).rand(B, N, C).astype(np.float32)
        fw = oracle.convsp_forward(nl, nl, data, onb, w, b0, R, one3, one3, int(normed), kernel)
        dq, dl, dd, _, _ = oracle.convsp_backward(nl, nl, data, onb, w, b0, R, one3, one3, int(normed), kernel, go)
        want[(kernel, C, normed)] = (fw, dq.astype(np.float64) + dl, dd, go, data)
        conv = identity_layer(spn, kernel, C, D, normed, R)
        layers[(kernel, C, normed)] = conv
        # ---- per-layer drop-in module
        lt = sl.detach().clone().requires_grad_(True)
        dt = gu.dev(data).requires_grad_(True)
        out = conv(lt, dt, nb)
        close(out, fw, "fwd %s C=%d" % (kernel, C))
        out.backward(gu.dev(go))
        close(lt.grad, want[(kernel, C, normed)][1], "dlocs %s C=%d" % (kernel, C), k=4)
        close(dt.grad, dd, "ddata %s C=%d" % (kernel, C), k=4)

    # ---- the fused groups of the fluid step on the same scene (tile lists, TMA-staged tiles), with the data
    # pattern fluidstep.FluidStep feeds them: None = ones, the position tensor itself, or a data tensor
    from smoothparticlenets_b200 import _native as nat
    tiles = spn.tile_lists_of(nb)
    assert tiles is not None and int(tiles[:4].view(torch.int32).item()) == 0, "tile lists usable at c2"
    press = r.rand(B, N, 1).astype(np.float32)
    xp = (nl * press).astype(np.float32)
    LOCS = "locs"
    groups = [
        [(("spiky", 1, False), None), (("dspiky", D, True), LOCS), (("dspiky", 1, True), None),
         (("cohesion", D, True), LOCS), (("cohesion", 1, True), None), (("constant", 1, False), None)],
        [(("dspiky", D, True), xp), (("dspiky", 1, True), press)],
        [(("constant", D, False), nv)],
        [(("spiky", D, False), nv), (("spiky", 1, False), None)]]
    for gi, specs in enumerate(groups):
        group = spn.ConvSPGroup([layers[s] for s, _ in specs])
        lt = sl.detach().clone().requires_grad_(True)
        datas, refs = [], []
        for (kernel, C, normed), d in specs:
            dn = scal if d is None else (nl if d is LOCS else d)
            datas.append(None if d is None else (lt if d is LOCS else gu.dev(d).requires_grad_(True)))
            w = np.eye(C, dtype=np.float32).reshape(C, C, 1)
            go = cases.rng(50 + 7 * gi + C).rand(B, N, C).astype(np.float32)
            fw = oracle.convsp_forward(nl, nl, dn, onb, w, np.zeros(C, np.float32), R, one3, one3, int(normed), kernel)
            dq, dl, dd, _, _ = oracle.convsp_backward(nl, nl, dn, onb, w, np.zeros(C, np.float32), R, one3, one3,
                                                      int(normed), kernel, go)
            refs.append((fw, dq.astype(np.float64) + dl, dd, go))
        n0 = nat.lib().spnb_launch_count()
        outs = group(lt, datas, nb)
        assert nat.lib().spnb_launch_count() - n0 == 2, "pack + one tile kernel"
        for (sp_, d), o, ref in zip(specs, outs, refs):
            close(o, ref[0], "group %d fwd %s C=%d" % (gi, sp_[0], sp_[1]))
        torch.autograd.backward(outs, [gu.dev(ref[3]) for ref in refs])
        # locs.grad: the geometry of every layer + the data gradient of the layers whose data IS the positions
        want_l = sum(ref[1] for ref in refs) + sum(ref[2].astype(np.float64) for (sp_, d), ref in zip(specs, refs)
                                                    if d is LOCS)
        close(lt.grad, want_l, "group %d dlocs" % gi, k=4 * len(specs))
        for (sp_, d), t, ref in zip(specs, datas, refs):
            if t is not None and d is not LOCS:
                close(t.grad, ref[2], "group %d ddata %s C=%d" % (gi, sp_[0], sp_[1]), k=4)


# ---------------------------------------------------------------------------------------------------------
# the benchmarked step against the reference's CPU step
# ---------------------------------------------------------------------------------------------------------
STEP_TOL = 1e-4  # of the tensor's scale; the achieved figures are printed and recorded in DESIGN.md


@pytest.mark.parametrize("fused", [True, False])
def test_fluid_step_vs_reference_cpu_step(spn, fused, capsys):
    """bench.py's step (examples/fluid_sim.py:355-424 data flow) on the reference's own CPU functions vs on
    libspnb: new positions, new velocities and the gradients wrt both inputs.  The CPU side is given the stable
    permutation (the contract of the reference's GPU sort) so that both sides sum their neighbours in one order."""
    from oracle import cpu_modules as cm
    B, N = 1, 8192
    locs, vel, _ = cases.fluid_cloud(77, B, N)
    gl = cases.rng(5).rand(B, N, 3).astype(np.float32)
    gv = cases.rng(6).rand(B, N, 3).astype(np.float32)

    cm.STABLE_ORDER = True
    try:
        ref = fluidstep.FluidStep(cm, radius=0.1, max_collisions=128)
        lt = torch.from_numpy(locs).requires_grad_(True)
        vt = torch.from_numpy(vel).requires_grad_(True)
        ol, ov = ref(lt, vt)
        torch.autograd.backward([ol, ov], [torch.from_numpy(gl), torch.from_numpy(gv)])
    finally:
        cm.STABLE_ORDER = False
    want = [ol.detach().numpy(), ov.detach().numpy(), lt.grad.numpy(), vt.grad.numpy()]

    model = fluidstep.FluidStep(spn, radius=0.1, max_collisions=128, fused=fused).cuda()
    lg = gu.dev(locs).requires_grad_(True)
    vg = gu.dev(vel).requires_grad_(True)
    pl, pv = model(lg, vg)
    torch.autograd.backward([pl, pv], [gu.dev(gl), gu.dev(gv)])
    got = [gu.host(t) for t in (pl, pv, lg.grad, vg.grad)]
    report = []
    for g, w, nm in zip(got, want, ("new_locs", "new_vel", "d/dlocs", "d/dvel")):
        scale = float(np.abs(w).max())
        err = float(np.abs(g.astype(np.float64) - w).max())
        report.append("%s %.2e" % (nm, err / scale))
        assert err <= STEP_TOL * scale, "%s: max |cuda - cpu reference| = %g at scale %g" % (nm, err, scale)
    with capsys.disabled():
        print("\n[fluid step vs reference CPU step, fused=%s] max err / scale: %s" % (fused, ", ".join(report)))


# ---------------------------------------------------------------------------------------------------------
# c4: ConvSDF, 16 objects of 64^3 cells
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ks,O", [((1, 1, 1), 1), ((3, 1, 1), 1), ((3, 3, 3), 4)])
def test_c4_convsdf_shape(spn, oracle, ks, O):
    B, N, D, S, n = 2, 16384, 3, 16, 64
    r = cases.rng(31)
    sdfs, cells = [], []
    for i in range(S):
        cell = 1.0 / n
        if i % 2:
            lo = 0.25 + 0.1 * r.rand(3)
            sdfs.append(cases.box_sdf(n, cell, lo.tolist(), (lo + 0.3 + 0.1 * r.rand(3)).tolist()))
        else:
            sdfs.append(cases.sphere_sdf(n, cell, (0.4 + 0.2 * r.rand(3)).tolist(), 0.15 + 0.15 * r.rand()))
        cells.append(cell)
    flat, offs, shapes = cases.pack_sdfs(sdfs, cells)
    locs = (r.rand(B, N, D) * 2.4 - 0.7).astype(np.float32)
    idxs = r.permutation(S).astype(np.float32)[None].repeat(B, 0)
    idxs[1, 3] = -1
    poses = np.zeros((B, S, 7), np.float32)
    poses[..., :3] = r.rand(B, S, 3) * 1.0 - 0.5
    poses[..., 3:] = cases.random_quats(r, (B, S))
    scales = (r.rand(B, S) + 0.5).astype(np.float32)
    ncells = int(np.prod(ks))
    weight = r.rand(O, ncells).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    ksz = np.array(ks, np.float32)
    dil = np.full(3, 0.01, np.float32)
    maxd = 0.3
    want = oracle.convsdf_forward(locs, idxs, poses, scales, flat, offs, shapes, weight, bias, ksz, dil, maxd)
    go = r.rand(B, N, O).astype(np.float32)
    wdl, wdw, _, wdb = oracle.convsdf_backward(locs, idxs, poses, scales, flat, offs, shapes, weight, bias, ksz, dil,
                                               maxd, go)
    conv = spn.ConvSDF([torch.from_numpy(s) for s in sdfs], cells, O, D, ks, 0.01, maxd)
    conv.weight = torch.nn.Parameter(torch.from_numpy(weight.copy()))
    conv.bias = torch.nn.Parameter(torch.from_numpy(bias.copy()))
    conv = conv.cuda()
    lt = gu.dev(locs).requires_grad_(True)
    out = conv(lt, gu.dev(idxs), gu.dev(poses), gu.dev(scales))
    # values are distances of O(0.1): 1e-6 absolute here is 1e-5 relative of the typical value
    gu.assert_close(gu.host(out), want, RTOL, 1e-6 * max(1.0, float(np.abs(want).max())), "convsdf fwd")
    out.backward(gu.dev(go))
    close(lt.grad, wdl, "convsdf dlocs", k=4)
    # d(weight) sums B*N = 32768 terms per entry: allow for the reference's own sequential fp32 accumulation
    close(conv.weight.grad, wdw, "convsdf dweight", k=4, rtol=1e-5 + 5e-7 * np.sqrt(B * N))
    close(conv.bias.grad, wdb, "convsdf dbias", k=4, rtol=1e-5 + 5e-7 * np.sqrt(B * N))


# ---------------------------------------------------------------------------------------------------------
# c3: 64 -> 64 channels, kernel_size 5, n-bar ~ 64, through the module on a query subset
# ---------------------------------------------------------------------------------------------------------
def test_c3_shape_query_subset(spn, oracle, capsys):
    """The CPU oracle needs ~30 ms per query forward and ~160 ms backward at this shape (64*64*125 weights per
    in-radius pair and cell), so the forward is compared on 256 queries and the gradients on the first 64."""
    B, N, M, MB, D, C, O, KS = 1, 24000, 256, 64, 3, 64, 64, 5
    R, DIL = 0.1, 0.025
    coll_r = R + DIL * 2  # the reference test's list radius R + DIL*(k-1)/2
    r = cases.rng(64)
    dens = 64.0 / (4.0 / 3.0 * np.pi * coll_r ** 3)  # ~64 particles per list ball
    L = (N / dens) ** (1.0 / 3)
    locs = (r.rand(B, N, D) * L).astype(np.float32)
    qsel = r.permutation(N)[:M]
    data = (r.rand(B, N, C) - 0.5).astype(np.float32)
    weight = ((r.rand(O, C, KS ** 3) - 0.5) / 8).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    coll = spn.ParticleCollision(D, coll_r, max_collisions=128, include_self=True).cuda()
    qlocs = locs[:, qsel] + (r.rand(B, M, D).astype(np.float32) - 0.5) * 0.01
    conv = spn.ConvSP(C, O, D, KS, DIL, R, kernel_fn="spiky")
    conv.weight = torch.nn.Parameter(torch.from_numpy(weight.copy()))
    conv.bias = torch.nn.Parameter(torch.from_numpy(bias.copy()))
    conv = conv.cuda()
    ksz, dl = np.full(3, KS, np.float32), np.full(3, DIL, np.float32)

    sl, sd, idxs, nb = coll(gu.dev(locs), gu.dev(data), gu.dev(qlocs))
    nbar = float((nb >= 0).sum().item()) / (B * M)
    nl, nd, nbh = gu.host(sl), gu.host(sd), gu.host(nb)
    out = conv(sl, sd, nb, gu.dev(qlocs))
    want = oracle.convsp_forward(qlocs, nl, nd, nbh, weight, bias, R, ksz, dl, 0, "spiky")
    # every output sums ~10^5 signed terms of magnitude ~10^3: the reference's sequential fp32 sum is itself
    # only good to ~1e-5 of the result here, so both sides are measured against a float64 evaluation
    f64, _ = convsp_float64(spn, qlocs, nl, nd, nbh, weight, bias, None, R, (KS,) * 3, [DIL] * 3, 0, "spiky")
    e_fwd = closer_than_reference(gu.host(out), want, f64, "c3 forward")

    qb, nbb = qlocs[:, :MB].copy(), nbh[:, :MB].copy()
    lt = sl.detach().clone().requires_grad_(True)
    dt = sd.detach().clone().requires_grad_(True)
    qt = gu.dev(qb).requires_grad_(True)
    outb = conv(lt, dt, gu.dev(nbb), qt)
    closer_than_reference(gu.host(outb), want[:, :MB], f64[:, :MB], "c3 forward (gradient subset)")
    go = r.rand(B, MB, O).astype(np.float32)
    outb.backward(gu.dev(go))
    wq, wl, wd, ww, wb = oracle.convsp_backward(qb, nl, nd, nbb, weight, bias, R, ksz, dl, 0, "spiky", go)
    # gradients sum up to ~10^7 signed terms each: both sides against float64 again
    _, w64, q64, l64, d64 = convsp_float64(spn, qb, nl, nd, nbb, weight, bias, go, R, (KS,) * 3, [DIL] * 3, 0, "spiky",
                                           grads=True)
    e_bwd = [closer_than_reference(gu.host(g), w32, w_64, nm) for g, w32, w_64, nm in (
        (qt.grad, wq, q64, "c3 dqlocs"), (lt.grad, wl, l64, "c3 dlocs"), (dt.grad, wd, d64, "c3 ddata"),
        (conv.weight.grad, ww, w64, "c3 dweight"))]
    close(conv.bias.grad, wb, "c3 dbias", k=4)
    with capsys.disabled():
        print("\n[c3 shape] n-bar %.1f, %d queries forward, %d with gradients; forward vs float64: cuda %.2e, "
              "reference fp32 %.2e (of max |out|); dqlocs/dlocs/ddata/dweight: cuda %s, reference fp32 %s" % (
                  nbar, M, MB, e_fwd[0], e_fwd[1], " ".join("%.1e" % e[0] for e in e_bwd),
                  " ".join("%.1e" % e[1] for e in e_bwd)))


# ---------------------------------------------------------------------------------------------------------
# d(weight): closer to a float64 accumulation than the reference's own fp32 sum
# ---------------------------------------------------------------------------------------------------------
def convsp_float64(spn, qlocs, locs, data, nb, weight, bias, go, radius, ks, dil, dis_norm, fn, grads=False):
    """ConvSP forward and gradients with every sum in float64, vectorised over the listed pairs; which pairs and
    kernel cells take part is decided on the fp32 geometry exactly as the reference does (common_funcs.h:476-566).
    Returns (out [B,M,O], dweight [O,C,ncells] or None when go is None) and, with grads=True, additionally
    (dqlocs, dlocs, ddata) -- the reference's formulas, including its omission of d(1/d)/dx under dis_norm."""
    import itertools
    wfn, dwfn = spn.KERNEL_FN[fn], spn.DKERNEL_FN[fn]
    B, M, K = nb.shape
    N, D = locs.shape[1], locs.shape[2]
    valid = np.cumprod(nb >= 0, axis=2).astype(bool)  # the list ends at the first negative entry
    j = np.where(valid, nb, 0).astype(np.int64)
    bidx = np.arange(B)[:, None, None]
    xj = locs.astype(np.float64)[bidx, j]             # B M K D
    dj = data.astype(np.float64)[bidx, j]             # B M K C
    q = qlocs.astype(np.float64)[:, :, None, :]
    half = [k // 2 for k in ks]
    ncells = int(np.prod(ks))
    O, C = weight.shape[0], data.shape[2]
    out = np.zeros((B, M, O)) + bias.astype(np.float64)[None, None]
    dw = np.zeros((O, C, ncells)) if go is not None else None
    dq = np.zeros((B, M, D)); dl = np.zeros((B, N, D)); dd = np.zeros((B, N, C))
    qf, xf = qlocs[:, :, None, :].astype(np.float32), locs[bidx, j].astype(np.float32)
    root = np.float32(1.73205 if D == 3 else (1.41421 if D == 2 else 1.0))
    cull = np.float32(np.float32(radius) + np.float32(max(ks) // 2) * np.float32(max(dil)) * root)
    d0 = np.zeros(nb.shape, np.float32)
    for k in range(D):
        d0 = d0 + (qf[..., k] - xf[..., k]) * (qf[..., k] - xf[..., k])
    valid = valid & ~(d0 > cull * cull)
    g64 = go.astype(np.float64) if go is not None else None
    bflat = np.broadcast_to(bidx, j.shape)
    for cell, idx in enumerate(itertools.product(*[range(s) for s in ks[::-1]])):
        off = np.array([np.float32(idx[::-1][k] - half[k]) * np.float32(dil[k]) for k in range(D)], np.float32)
        d2f = np.zeros(nb.shape, np.float32)
        for k in range(D):
            nr = qf[..., k] + off[k] - xf[..., k]
            d2f = d2f + nr * nr
        inr = valid & (d2f < np.float32(radius) * np.float32(radius))
        disp = (q + off.astype(np.float64)) - xj      # B M K D
        d = np.sqrt((disp ** 2).sum(-1))
        norm = np.where(d > 0, 1.0 / np.where(d > 0, d, 1.0), 1.0) if dis_norm else np.ones_like(d)
        wv = np.where(inr, wfn(d, float(radius)), 0.0) * norm
        T = (wv[..., None] * dj).sum(2)               # B M C
        Wc = weight[:, :, cell].astype(np.float64)    # O C
        out += T @ Wc.T
        if dw is not None:
            dw[:, :, cell] = np.einsum("bmo,bmc->oc", g64, T)
        if grads:
            u = g64 @ Wc                               # B M C  = sum_o go[o] w[o,c]
            np.add.at(dd, (bflat, j), wv[..., None] * u[:, :, None, :])
            a = (dj * u[:, :, None, :]).sum(-1)        # B M K
            dwv = np.where(inr & (d > 0), dwfn(np.where(d > 0, d, 1.0), float(radius)) / np.where(d > 0, d, 1.0), 0.0) * norm
            t = (a * dwv)[..., None] * disp            # B M K D
            dq += t.sum(2)
            np.add.at(dl, (bflat, j), -t)
    if grads:
        return out, dw, dq, dl, dd
    return out, dw


def closer_than_reference(got, ref32, ref64, what, capsys=None):
    """|got - float64| must not exceed the larger of |reference fp32 - float64| and the north_star tolerance
    (1e-5 relative + 1e-6 * max absolute, against the float64 values); returns the two max-norm errors."""
    got, ref32 = np.asarray(got, np.float64), np.asarray(ref32, np.float64)
    scale = float(np.abs(ref64).max())
    e_got, e_ref = float(np.abs(got - ref64).max()), float(np.abs(ref32 - ref64).max())
    tol = 1e-6 * scale + 1e-5 * np.abs(ref64)
    ok = (np.abs(got - ref64) <= np.maximum(tol, e_ref)).all()
    assert ok, "%s: max |cuda - float64| = %.3g, reference's own fp32 error %.3g, scale %.3g" % (what, e_got, e_ref, scale)
    return e_got / scale, e_ref / scale


def test_dweight_against_float64(spn, oracle, capsys):
    """c1 shape (B4 N1024 D3 4 -> 8, kernel_size 3, dilation 0.05, radius 0.1, spiky).  The fp32 tolerance of
    test_gpu_convsp.py for d(weight) is wider than 1e-5 because the reference adds ~10^5 terms per entry
    sequentially in fp32; this shows which side that error is on."""
    B, N, D, C, O, R, DIL = 4, 1024, 3, 4, 8, 0.1, 0.05
    ks = (3, 3, 3)
    r = cases.rng(101)
    locs = (r.rand(B, N, D) * 0.55).astype(np.float32)
    data = r.rand(B, N, C).astype(np.float32)
    weight = r.rand(O, C, 27).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    coll = spn.ParticleCollision(D, R + DIL, max_collisions=128).cuda()
    sl, sd, idxs, nb = coll(gu.dev(locs), gu.dev(data))
    nl, nd, nbh = gu.host(sl), gu.host(sd), gu.host(nb)
    go = r.rand(B, N, O).astype(np.float32)
    ksz, dl = np.array(ks, np.float32), np.full(3, DIL, np.float32)
    _, _, _, w32, _ = oracle.convsp_backward(nl, nl, nd, nbh, weight, bias, R, ksz, dl, 0, "spiky", go)
    _, w64 = convsp_float64(spn, nl, nl, nd, nbh, weight, bias, go, R, ks, [DIL] * 3, 0, "spiky")
    _, _, _, dw = gu.convsp_backward(sl, sl, sd, nb, gu.dev(weight), R, gu.dev(ksz), gu.dev(dl), 0,
                                     cases.KERNEL_NAMES.index("spiky"), gu.dev(go), same=True)
    got = gu.host(dw).astype(np.float64)
    scale = np.abs(w64).max()
    e_gpu, e_ref = np.abs(got - w64).max() / scale, np.abs(w32 - w64).max() / scale
    with capsys.disabled():
        print("\n[dweight vs float64] cuda %.2e, reference fp32 sum %.2e (relative to max |dweight|)" % (e_gpu, e_ref))
    assert e_gpu <= max(e_ref, 1e-6), "the CUDA d(weight) must be at least as close to float64 as the reference's"
    gu.assert_close(got, w64, 1e-5, 1e-6 * scale, "dweight vs float64 at the north_star tolerance")


# ---------------------------------------------------------------------------------------------------------
# sidecars: an in-place edit of the neighbour tensor must not be ignored
# ---------------------------------------------------------------------------------------------------------
def test_neighbors_edited_in_place(spn, oracle):
    B, N, D, R = 2, 3000, 3, 0.1
    locs, vel, _ = cases.fluid_cloud(15, B, N)
    coll = spn.ParticleCollision(D, R, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    assert spn.tile_lists_of(nb) is not None and spn.sym_flag_of(nb) is not None
    with torch.no_grad():
        nb[:, ::2, 6:] = -1          # mask: every other particle keeps at most 6 neighbours
    assert spn.tile_lists_of(nb) is None and spn.sym_flag_of(nb) is None, "stale sidecars must be dropped"
    nl, nv, nbh = gu.host(sl), gu.host(sv), gu.host(nb)
    one3 = np.ones(3, np.float32)
    layers = [identity_layer(spn, "dspiky", 3, D, True), identity_layer(spn, "dspiky", 1, D, True)]
    press = cases.rng(2).rand(B, N, 1).astype(np.float32)
    lt = sl.detach().clone().requires_grad_(True)
    outs = spn.ConvSPGroup(layers)(lt, [sv, gu.dev(press)], nb)
    gos = [cases.rng(3).rand(B, N, 3).astype(np.float32), cases.rng(4).rand(B, N, 1).astype(np.float32)]
    torch.autograd.backward(outs, [gu.dev(g) for g in gos])
    tot = 0
    for conv, data, go, out, C in zip(layers, (nv, press), gos, outs, (3, 1)):
        w = np.eye(C, dtype=np.float32).reshape(C, C, 1)
        want = oracle.convsp_forward(nl, nl, data, nbh, w, np.zeros(C, np.float32), R, one3, one3, 1, "dspiky")
        close(out, want, "edited lists fwd C=%d" % C)
        dq, dl, _, _, _ = oracle.convsp_backward(nl, nl, data, nbh, w, np.zeros(C, np.float32), R, one3, one3, 1,
                                                 "dspiky", go)
        tot = tot + dq.astype(np.float64) + dl
        single = conv(sl, gu.dev(data), nb)
        close(single, want, "edited lists, per-layer fwd C=%d" % C)
    close(lt.grad, tot, "edited lists dlocs (asymmetric relation -> scatter path)", k=8)
