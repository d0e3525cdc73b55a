"""GPU parity: ConvSP forward / backward against the oracle.

Bar (north_star): 1e-5 relative / 1e-6 absolute in fp32.  Individual terms are bit-identical to the
reference's (same association, same float/double promotions); only the summation order differs, so
the absolute part of the tolerance is scaled by the magnitude of the summed terms (the SPH kernels
reach 1e4..1e7 for radius 0.1), i.e. atol = 1e-6 * max|result|.
"""
import itertools

import numpy as np
import pytest
import torch

import cases
import gpu_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def atol_for(ref):
    return 1e-6 * max(1.0, float(np.abs(ref).max()))


def build_lists(oracle, locs, data, qlocs, coll_radius, K, include_self=1, G=96):
    D = locs.shape[2]
    low, gd = oracle.grid_bounds(locs, coll_radius, G)
    ids, idxs = oracle.hashgrid_order(locs, low, gd, coll_radius, stable=True)
    nl, nd = oracle.reorder_data(locs, data, idxs)
    q = nl if qlocs is None else qlocs
    nb, _, _ = oracle.compute_collisions(q, nl, low, gd, ids, coll_radius, coll_radius, K, include_self,
                                         G ** D)
    return nl, nd, q, nb


def run_case(oracle, locs, data, qlocs, weight, bias, radius, ksize, dil, dis_norm, fn, nb,
             sym_possible):
    D = locs.shape[2]
    ks_np = np.array(ksize, np.float32)
    dil_np = np.full(D, dil, np.float32) if np.isscalar(dil) else np.array(dil, np.float32)
    fid = cases.KERNEL_NAMES.index(fn)
    q_np = locs if qlocs is None else qlocs
    lt, dt, nbt, wt, bt = map(gu.dev, (locs, data, nb, weight, bias))
    qt = lt if qlocs is None else gu.dev(qlocs)
    kst, dilt = gu.dev(ks_np), gu.dev(dil_np)

    want = oracle.convsp_forward(q_np, locs, data, nb, weight, bias, radius, ks_np, dil_np, dis_norm, fn)
    got = gu.host(gu.convsp_forward(qt, lt, dt, nbt, wt, bt, radius, kst, dilt, dis_norm, fid))
    gu.assert_close(got, want, RTOL, atol_for(want), "fwd %s norm=%d" % (fn, dis_norm))

    go = np.random.RandomState(5).rand(*want.shape).astype(np.float32)
    odq, odl, odd, odw, _ = oracle.convsp_backward(q_np, locs, data, nb, weight, bias, radius, ks_np,
                                                   dil_np, dis_norm, fn, go)
    got_ = gu.convsp_backward(qt, lt, dt, nbt, wt, radius, kst, dilt, dis_norm, fid, gu.dev(go))
    # dweight sums one term per (listed pair, kernel cell) over the WHOLE batch; the reference adds
    # them sequentially in fp32, whose own rounding error grows like eps*sqrt(n)*sum.  The GPU sum is
    # hierarchical (more accurate), so the comparison must allow for the oracle's accumulation error.
    nterms = float((nb >= 0).sum()) * weight.shape[2]
    rtol_w = RTOL + 5e-7 * np.sqrt(nterms)
    for g, w, nm in zip(got_, (odq, odl, odd, odw), ("dqlocs", "dlocs", "ddata", "dweight")):
        gu.assert_close(gu.host(g), w, rtol_w if nm == "dweight" else RTOL, atol_for(w) * 4,
                        "%s %s norm=%d" % (nm, fn, dis_norm))
    if qlocs is None:
        # one buffer for d/dqlocs + d/dlocs, atomic and (if the lists allow) symmetric-gather modes
        want_sum = odq.astype(np.float64) + odl
        modes = [None]
        if sym_possible:
            modes.append(torch.zeros(1, device="cuda", dtype=torch.int32))
            modes.append(torch.ones(1, device="cuda", dtype=torch.int32))  # flag set -> atomic path
        for flag in modes:
            dq, dl, dd, dw = gu.convsp_backward(qt, lt, dt, nbt, wt, radius, kst, dilt, dis_norm, fid,
                                                gu.dev(go), sym_flag=flag, same=True)
            tag = "sym" if (flag is not None and int(flag.item()) == 0) else "atomic"
            gu.assert_close(gu.host(dq), want_sum, RTOL, atol_for(want_sum) * 4, "dq+dl %s %s" % (tag, fn))
            gu.assert_close(gu.host(dd), odd, RTOL, atol_for(odd) * 4, "ddata %s %s" % (tag, fn))
            gu.assert_close(gu.host(dw), odw, rtol_w, atol_for(odw) * 4, "dweight %s %s" % (tag, fn))


@pytest.mark.parametrize("use_qlocs", [True, False])
def test_reference_test_shape_all_kernels(spn, oracle, use_qlocs):
    """tests/test_convsp.py:65-115: B2 N5 M3 D2 ks(3,1) R1.0 dil .05 C2 O3, all 12 kernels."""
    locs, qlocs, data, weight, bias = cases.convsp_case(0)
    ks, R, dil = (3, 1), 1.0, 0.05
    nl, nd, q, nb = build_lists(oracle, locs, data, qlocs if use_qlocs else None,
                                R + dil * max((k - 1) / 2 for k in ks), K=128)
    for fn, dn in itertools.product(cases.KERNEL_NAMES, (0, 1)):
        run_case(oracle, nl, nd, q if use_qlocs else None, weight, bias, R, ks, dil, dn, fn, nb, False)


@pytest.mark.parametrize("C,O,D", [(1, 1, 3), (3, 3, 3), (1, 1, 2), (2, 2, 2)])
def test_fluid_layers_small_path(spn, oracle, C, O, D):
    """The ncells == 1 register path used by the fluid layers (fluid_sim.py:156-175): every kernel,
    dis_norm on/off, symmetric-gather and atomic backward modes."""
    B, N, R = 2, 600, 0.1
    locs, vel, L = cases.fluid_cloud(2, B, N, D=D, density=7640.0 if D == 3 else 600.0)
    r = cases.rng(9)
    data = r.rand(B, N, C).astype(np.float32)
    weight = r.rand(O, C, 1).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, None, R, K=128, include_self=0)
    assert (nb[..., -1] < 0).all()
    for fn, dn in itertools.product(cases.KERNEL_NAMES, (0, 1)):
        run_case(oracle, nl, nd, None, weight, bias, R, (1,) * D, 1.0, dn, fn, nb, True)
    # separate query set through the same kernels
    qlocs = (r.rand(B, 77, D) * L).astype(np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, qlocs, R, K=128)
    for fn in ("spiky", "cohesion"):
        run_case(oracle, nl, nd, q, weight, bias, R, (1,) * D, 1.0, 1, fn, nb, False)


def test_config1_shape(spn, oracle):
    """BASELINE.json config 1: B4 N1024 D3 4->8 ks3 dil .05 r .1 spiky (collision radius .15)."""
    B, N, D, C, O = 4, 1024, 3, 4, 8
    r = cases.rng(0)
    locs = r.rand(B, N, D).astype(np.float32)
    data = r.rand(B, N, C).astype(np.float32)
    weight = r.rand(O, C, 27).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, None, 0.15, K=128)
    run_case(oracle, nl, nd, None, weight, bias, 0.1, (3, 3, 3), 0.05, 0, "spiky", nb, False)


@pytest.mark.parametrize("D,ks,C,O", [(1, (5,), 3, 2), (3, (1, 3, 1), 5, 11), (4, (1, 1, 3, 1), 2, 2)])
def test_generic_shapes(spn, oracle, D, ks, C, O):
    B, N = 2, 300
    r = cases.rng(4)
    locs = r.rand(B, N, D).astype(np.float32)
    qlocs = r.rand(B, 41, D).astype(np.float32)
    data = r.rand(B, N, C).astype(np.float32)
    weight = r.rand(O, C, int(np.prod(ks))).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    R = {1: 0.02, 3: 0.2, 4: 0.4}[D]
    dil = 0.3 * R
    nl, nd, q, nb = build_lists(oracle, locs, data, qlocs, R + dil * max((k - 1) / 2 for k in ks), K=32,
                                G=16 if D == 4 else 96)
    for fn, dn in (("default", 0), ("dspiky", 1), ("sigmoid", 0)):
        run_case(oracle, nl, nd, q, weight, bias, R, ks, dil, dn, fn, nb, False)


def test_truncated_and_garbage_after_terminator(spn, oracle):
    """Lists are consumed up to the FIRST negative entry (common_funcs.h:476), full rows have no
    terminator, and entries after the terminator are ignored."""
    B, N, D, C, O = 1, 400, 3, 3, 3
    locs, vel, L = cases.fluid_cloud(7, B, N)
    r = cases.rng(1)
    data = r.rand(B, N, C).astype(np.float32)
    weight = r.rand(O, C, 1).astype(np.float32)
    bias = np.zeros(O, np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, None, 0.1, K=8, include_self=1)
    assert (nb[..., -1] >= 0).any(), "some rows must be full for this test"
    junk = nb.copy()
    for row in junk.reshape(-1, 8):
        neg = np.where(row < 0)[0]
        if len(neg) and neg[0] + 1 < 8:
            row[neg[0] + 1:] = 3.0  # garbage behind the terminator
    run_case(oracle, nl, nd, None, weight, bias, 0.1, (1, 1, 1), 1.0, 0, "spiky", junk, False)


def test_module_autograd_matches_oracle(spn, oracle):
    """ParticleCollision -> ConvSP modules with autograd: loss gradients wrt locs (both roles
    summed), data, weight and bias equal the oracle's (convsp.py:176-203)."""
    B, N, D, C, O, R = 2, 500, 3, 3, 3, 0.1
    locs, vel, L = cases.fluid_cloud(11, B, N)
    r = cases.rng(2)
    coll = spn.ParticleCollision(D, R, include_self=False).cuda()
    conv = spn.ConvSP(C, O, D, 1, 1, R, dis_norm=True, kernel_fn="dspiky").cuda()
    weight = r.rand(O, C, 1).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    conv.weight.data.copy_(gu.dev(weight))
    conv.bias.data.copy_(gu.dev(bias))
    lt = gu.dev(locs)
    sl, sv, idxs, nb = coll(lt, gu.dev(vel))
    sl = sl.detach().requires_grad_(True)
    sv = sv.detach().requires_grad_(True)
    out = conv(sl, sv, nb)
    go = torch.rand_like(out)
    out.backward(go)
    nl, nd, nbn = gu.host(sl), gu.host(sv), gu.host(nb)
    ks, dil = np.ones(D, np.float32), np.ones(D, np.float32)
    want = oracle.convsp_forward(nl, nl, nd, nbn, weight, bias, R, ks, dil, 1, "dspiky")
    gu.assert_close(gu.host(out), want, RTOL, atol_for(want), "module fwd")
    odq, odl, odd, odw, odb = oracle.convsp_backward(nl, nl, nd, nbn, weight, bias, R, ks, dil, 1,
                                                     "dspiky", gu.host(go))
    s = odq.astype(np.float64) + odl
    gu.assert_close(gu.host(sl.grad), s, RTOL, atol_for(s) * 4, "locs.grad")
    gu.assert_close(gu.host(sv.grad), odd, RTOL, atol_for(odd) * 4, "data.grad")
    rtol_w = RTOL + 5e-7 * np.sqrt(float((nbn >= 0).sum()))
    gu.assert_close(gu.host(conv.weight.grad), odw, rtol_w, atol_for(odw) * 4, "weight.grad")
    gu.assert_close(gu.host(conv.bias.grad), odb, 1e-5, atol_for(odb), "bias.grad")


def test_gradcheck_double_numeric(spn):
    """The reference's gradcheck (tests/test_convsp.py:134-156): analytic fp32 gradients against a
    float64 central-difference Jacobian of an independent numpy implementation, eps 1e-4,
    atol 1e-3, rtol 1e-1, for the shape of the reference test."""
    locs, qlocs, data, weight, bias = cases.convsp_case(0)
    ks, R, dil, fn = (3, 1), 1.0, 0.05, "default"
    w_fn = spn.KERNEL_FN[fn]

    def pyconv(q, l, d, w, b):
        B, M, N = l.shape[0], q.shape[1], l.shape[1]
        out = np.zeros((B, M, w.shape[0]))
        centers = (np.array(ks) - 1) / 2
        for bb, i, j in itertools.product(range(B), range(M), range(N)):
            for k, idx in enumerate(itertools.product(*[range(x) for x in ks[::-1]])):
                dd = np.square(q[bb, i] + (np.array(idx[::-1]) - centers) * dil - l[bb, j]).sum()
                if dd > R * R:
                    continue
                out[bb, i] += w[:, :, k].dot(w_fn(np.sqrt(dd), R) * d[bb, j])
        return out + b[None, None]

    coll = spn.ParticleCollision(2, R + dil).cuda()
    conv = spn.ConvSP(2, 3, 2, ks, dil, R, kernel_fn=fn).cuda()
    sl, sd, idxs, nb = coll(gu.dev(locs), gu.dev(data), gu.dev(qlocs))
    args = [sl.detach(), sd.detach(), gu.dev(weight), gu.dev(bias), gu.dev(qlocs)]
    args = [a.clone().requires_grad_(True) for a in args]

    lt, dt, wt, bt, qt = args
    conv.weight, conv.bias = torch.nn.Parameter(wt.detach()), torch.nn.Parameter(bt.detach())
    out = conv(lt, dt, nb, qt)
    go = torch.rand_like(out)
    out.backward(go)
    grads = [lt.grad, dt.grad, conv.weight.grad, conv.bias.grad, qt.grad]
    base = [gu.host(a).astype(np.float64) for a in (lt, dt, wt, bt, qt)]
    gon = gu.host(go).astype(np.float64)
    eps = 1e-4
    for ai, g in enumerate(grads):
        num = np.zeros_like(base[ai])
        it = np.nditer(base[ai], flags=["multi_index"])
        for _ in it:
            mi = it.multi_index
            hi = [x.copy() for x in base]
            lo = [x.copy() for x in base]
            hi[ai][mi] += eps
            lo[ai][mi] -= eps
            f = lambda v: (pyconv(v[4], v[0], v[1], v[2], v[3]) * gon).sum()
            num[mi] = (f(hi) - f(lo)) / (2 * eps)
        a = gu.host(g).astype(np.float64)
        assert np.all(np.abs(a - num) <= 1e-3 + 1e-1 * np.abs(num)), ("gradcheck arg %d" % ai)


@pytest.mark.parametrize("D,ks,C,O,fn,dn", [(3, (3, 3, 3), 64, 64, "spiky", 0), (3, (5, 5, 5), 64, 64, "spiky", 0),
                                           (3, (3, 1, 3), 32, 8, "dspiky", 1), (2, (3, 3), 96, 70, "default", 0),
                                           (1, (5,), 36, 3, "cohesion", 1)])
def test_wide_channel_forward(spn, oracle, D, ks, C, O, fn, dn):
    """BASELINE.json config 3 shape (64 -> 64, kernel_size 5) and relatives through the factored wide-channel
    kernels -- gather + tcgen05 3xTF32 contraction (csrc/convsp_wide_mma.cu) for C in {32, 64}, the CUDA-core
    kernel (csrc/convsp_wide.cu) otherwise -- against the oracle on a query subset."""
    B, N, M = 2, 400, 37
    r = cases.rng(8)
    R = {1: 0.01, 2: 0.08, 3: 0.2}[D]
    dil = 0.4 * R
    locs = r.rand(B, N, D).astype(np.float32)
    qlocs = r.rand(B, M, D).astype(np.float32)
    data = r.randn(B, N, C).astype(np.float32)
    weight = (r.randn(O, C, int(np.prod(ks))) / np.sqrt(C * np.prod(ks))).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, qlocs, R + dil * max((k - 1) / 2 for k in ks), K=64)
    assert (nb >= 0).sum() > 4 * B * M
    ks_np, dil_np = np.array(ks, np.float32), np.full(D, dil, np.float32)
    want = oracle.convsp_forward(q, nl, nd, nb, weight, bias, R, ks_np, dil_np, dn, fn)
    got = gu.convsp_forward_wide(gu.dev(q), gu.dev(nl), gu.dev(nd), gu.dev(nb), gu.dev(weight), gu.dev(bias), R,
                                 gu.dev(ks_np), gu.dev(dil_np), dn, cases.KERNEL_NAMES.index(fn))
    # signed data and weights: terms cancel, so the absolute tolerance follows the magnitude of the terms
    terms = oracle.convsp_forward(q, nl, np.abs(nd), nb, np.abs(weight), np.abs(bias), R, ks_np, dil_np, dn, fn)
    gu.assert_close(gu.host(got), want, RTOL, 1e-6 * max(1.0, float(np.abs(terms).max())), "wide fwd")
    # the module picks the wide kernel by itself for these shapes (transpose + kernel = 2 launches)
    from smoothparticlenets_b200 import _native as nat
    conv = spn.ConvSP(C, O, D, ks, dil, R, dis_norm=bool(dn), kernel_fn=fn).cuda()
    conv.weight.data.copy_(gu.dev(weight))
    conv.bias.data.copy_(gu.dev(bias))
    n0 = nat.lib().spnb_launch_count()
    with torch.no_grad():
        out = conv(gu.dev(nl), gu.dev(nd), gu.dev(nb), gu.dev(q))
    if C * O * int(np.prod(ks)) >= 4096:
        # C in {32, 64}: weight images + gather + tcgen05 GEMM (convsp_wide_mma.cu); else transpose + CUDA-core kernel
        assert nat.lib().spnb_launch_count() - n0 == (3 if C in (32, 64) else 2)
        assert torch.equal(out, got)
    else:  # small weight tensors stay on the generic kernel
        gu.assert_close(gu.host(out), want, RTOL, 1e-6 * max(1.0, float(np.abs(terms).max())), "module fwd")


@pytest.mark.parametrize("D,ks,C,O,fn,dn,alias,K", [(3, (3, 3, 3), 64, 64, "spiky", 0, False, 64),
                                                    (3, (3, 3, 3), 64, 40, "default", 1, True, 64),
                                                    (2, (5, 5), 32, 64, "cohesion", 0, False, 160),
                                                    (3, (5, 1, 3), 32, 16, "dspiky", 1, True, 48)])
def test_wide_channel_backward(spn, oracle, D, ks, C, O, fn, dn, alias, K):
    """All four gradients of the wide-channel shapes through the tensor-core backward (csrc/convsp_wide_bwd.cu:
    dG and dweight as tcgen05 3xTF32 contractions, one list walk for ddata / dlocs / dqlocs) against the oracle;
    both sides are measured against a float64 evaluation because every entry sums 10^4..10^6 signed terms.
    alias: qlocs is None, i.e. one gradient buffer receives d/dqlocs + d/dlocs.  K = 160: lists longer than the
    128-entry staging round."""
    from test_gpu_parity_configs import closer_than_reference, convsp_float64
    from smoothparticlenets_b200 import _native as nat
    B, N = 2, 300
    M = N if alias else 45
    r = cases.rng(21)
    R = {2: 0.09, 3: 0.2}[D]
    if K > 128:
        R = 0.25  # long lists
    dil = 0.4 * R
    locs = r.rand(B, N, D).astype(np.float32)
    qlocs = locs if alias else r.rand(B, M, D).astype(np.float32)
    data = r.randn(B, N, C).astype(np.float32)
    weight = (r.randn(O, C, int(np.prod(ks))) / np.sqrt(C * np.prod(ks))).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, qlocs, R + dil * max((k - 1) / 2 for k in ks), K=K)
    if alias:
        q = nl  # the particles, in their sorted order, are their own queries
    if K > 128:
        assert (nb >= 0).sum(-1).max() > 128
    ks_np, dil_np = np.array(ks, np.float32), np.full(D, dil, np.float32)
    go = r.randn(B, M, O).astype(np.float32)
    assert nat.lib().spnb_convsp_backward_wide_workspace_bytes(O, C, D, int(np.prod(ks))) > 0
    conv = spn.ConvSP(C, O, D, ks, dil, R, dis_norm=bool(dn), kernel_fn=fn).cuda()
    conv.weight.data.copy_(gu.dev(weight))
    conv.bias.data.copy_(gu.dev(bias))
    lt = gu.dev(nl).requires_grad_(True)
    dt = gu.dev(nd).requires_grad_(True)
    qt = None if alias else gu.dev(q).requires_grad_(True)
    out = conv(lt, dt, gu.dev(nb), qt)
    n0 = nat.lib().spnb_launch_count()
    out.backward(gu.dev(go))
    # transposed weights, go images, dG GEMM, list walk, transposed gather, dweight GEMM
    assert nat.lib().spnb_launch_count() - n0 == 6
    wq, wl, wd, ww, wb = oracle.convsp_backward(q, nl, nd, nb, weight, bias, R, ks_np, dil_np, dn, fn, go)
    _, w64, q64, l64, d64 = convsp_float64(spn, q, nl, nd, nb, weight, bias, go, R, ks, [dil] * D, dn, fn, grads=True)
    if alias:
        closer_than_reference(gu.host(lt.grad), wq + wl, q64 + l64, "wide dlocs (+dqlocs)")
    else:
        closer_than_reference(gu.host(qt.grad), wq, q64, "wide dqlocs")
        closer_than_reference(gu.host(lt.grad), wl, l64, "wide dlocs")
    closer_than_reference(gu.host(dt.grad), wd, d64, "wide ddata")
    closer_than_reference(gu.host(conv.weight.grad), ww, w64, "wide dweight")
    gu.assert_close(gu.host(conv.bias.grad), wb, 1e-5, 1e-5 * float(np.abs(wb).max()), "wide dbias")
    # only some gradients requested
    dt2 = gu.dev(nd).requires_grad_(True)
    for p_ in conv.parameters():
        p_.requires_grad_(False)
    out2 = conv(gu.dev(nl), dt2, gu.dev(nb), None if alias else gu.dev(q))
    out2.backward(gu.dev(go))
    # (float atomics: the order of the additions differs from run to run)
    assert float((dt2.grad - dt.grad).abs().max()) <= 1e-5 * float(dt.grad.abs().max())


def test_wide_gather_on_tensor_cores(spn, oracle, monkeypatch):
    """SPNB_WIDE_TC=1: the gather itself as per-query tcgen05 GEMMs (k_wide_gather_tc: S[cell, neighbour] x
    features, accumulators of 8 queries in tensor memory) -- an opt-in alternative to the CUDA-core gather; same
    forward and the same four gradients."""
    monkeypatch.setenv("SPNB_WIDE_TC", "1")
    test_wide_channel_forward(spn, oracle, 3, (3, 3, 3), 64, 64, "spiky", 0)
    test_wide_channel_backward(spn, oracle, 3, (3, 3, 3), 64, 64, "spiky", 0, False, 64)
    test_wide_channel_backward(spn, oracle, 3, (5, 1, 3), 32, 16, "dspiky", 1, True, 48)


def test_full_size_properties(spn):
    """BASELINE.json config 2 size (8 x 65536 particles): size-independent properties of ConvSP --
    the `constant` kernel with unit data counts neighbours, outputs are linear in the data, the fused
    group path equals the per-layer path, and the two backward modes (symmetric gather / atomics) agree."""
    B, N, R = 8, 65536, 0.1
    locs, vel, _ = cases.fluid_cloud(0, B, N)
    coll = spn.ParticleCollision(3, R, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    ones = torch.ones(B, N, 1, device="cuda")

    def layer(kernel, C, normed):
        c = spn.ConvSP(C, C, 3, 1, 1, R, dis_norm=normed, with_params=False, kernel_fn=kernel).cuda()
        c.weight.zero_()
        c.bias.zero_()
        for i in range(C):
            c.weight[i, i, 0] = 1
        return c
    count = layer("constant", 1, False)(sl, ones, nb)
    assert torch.equal(count[..., 0], (nb >= 0).sum(2).float()), "constant kernel counts the list entries"
    spiky3 = layer("spiky", 3, False)
    a, b_ = spiky3(sl, sv, nb), spiky3(sl, 2.0 * sv + 1.0, nb)
    dens = layer("spiky", 1, False)(sl, ones, nb)
    lin = 2.0 * a + dens  # W-weighted sum of (2 v + 1) = 2 * sum(W v) + sum(W)
    assert float((b_ - lin).abs().max()) <= 2e-5 * float(lin.abs().max()), "linearity in the data"
    group = spn.ConvSPGroup([spiky3, layer("spiky", 1, False)])
    g3, g1 = group(sl, [sv, ones], nb)
    assert float((g3 - a).abs().max()) <= 1e-5 * float(a.abs().max())
    assert float((g1 - dens).abs().max()) <= 1e-5 * float(dens.abs().max())
    # backward: symmetric gather (tagged lists) vs atomics (untagged copy of the same lists)
    go = torch.rand(B, N, 3, device="cuda")
    dsp = layer("dspiky", 3, True)
    grads = []
    for lists in (nb, nb.clone()):
        l = sl.detach().clone().requires_grad_(True)
        d = sv.detach().clone().requires_grad_(True)
        dsp(l, d, lists).backward(go)
        grads.append((l.grad, d.grad))
    for x, y in zip(*grads):
        assert float((x - y).abs().max()) <= 2e-5 * float(y.abs().max()), "gather and atomic backward agree"


@pytest.mark.parametrize("D,C,fn,dn", [(3, 3, "dspiky", 1), (3, 1, "spiky", 0), (2, 2, "cohesion", 1)])
def test_backward_for_a_query_block(spn, oracle, D, C, fn, dn):
    """spnb_convsp_backward_block (the slab decomposition's backward): the queries are a block of the particles,
    grad_output is known for all particles, and the block's gradients are gathered over its own lists -- equal to the
    block's rows of the reference's scatter over ALL queries."""
    from smoothparticlenets_b200 import _native as nat
    B, N, K, R = 1, 900, 64, 0.12 if D == 3 else 0.08
    r = cases.rng(17)
    locs = r.rand(B, N, D).astype(np.float32)
    data = r.randn(B, N, C).astype(np.float32)
    w = (r.rand(C, C, 1) - 0.4).astype(np.float32)
    nl, nd, q, nb = build_lists(oracle, locs, data, None, R, K=K, include_self=0)
    assert (nb >= 0).sum(-1).max() < K, "no list is cut: the relation is symmetric"
    go = r.randn(B, N, C).astype(np.float32)
    one = np.ones(D, np.float32)
    dq, dl, dd, _, _ = oracle.convsp_backward(nl, nl, nd, nb, w, np.zeros(C, np.float32), R, one, one, dn, fn, go)
    want_l, want_d = dq.astype(np.float64) + dl, dd
    a, b = 211, 640
    tl, td, tn, tw, tg = gu.dev(nl), gu.dev(nd), gu.dev(nb[:, a:b].copy()), gu.dev(w), gu.dev(go)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    gl = torch.full((B, b - a, D), 7.0, device="cuda")
    gd = torch.full((B, b - a, C), 7.0, device="cuda")
    nat.check(nat.lib().spnb_convsp_backward_block(
        nat.ptr(tl), nat.ptr(td), nat.ptr(tn), nat.ptr(tw), B, b - a, N, C, D, K, C, 1, R, dn,
        cases.KERNEL_NAMES.index(fn), nat.ptr(tg), a, nat.ptr(flag), nat.ptr(gl), nat.ptr(gd), nat.stream()), "block bwd")
    gu.assert_close(gu.host(gl), want_l[:, a:b], 1e-5, 4e-6 * float(np.abs(want_l).max()), "block dlocs")
    gu.assert_close(gu.host(gd), want_d[:, a:b], 1e-5, 4e-6 * float(np.abs(want_d).max()), "block ddata")
    # a raised flag: the kernel leaves the block buffers alone instead of scattering into rows they do not have
    flag.fill_(1)
    gl.fill_(7.0)
    nat.check(nat.lib().spnb_convsp_backward_block(
        nat.ptr(tl), nat.ptr(td), nat.ptr(tn), nat.ptr(tw), B, b - a, N, C, D, K, C, 1, R, dn,
        cases.KERNEL_NAMES.index(fn), nat.ptr(tg), a, nat.ptr(flag), nat.ptr(gl), nat.ptr(gd), nat.stream()), "block bwd")
    assert float(gl.min()) == 7.0 and float(gl.max()) == 7.0
