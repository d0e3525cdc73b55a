"""CPU tests that PIN the oracle (oracle/spn_oracle.c) to the reference.

1. against the committed golden vectors (tests/golden/*.npz, generated from the unmodified reference
   CPU extension by tests/golden/make_golden.py) -- runs everywhere, incl. the GPU box;
2. against the reference CPU extension itself (oracle/_ref) on fresh seeded inputs -- runs where
   oracle/_ref exists (it is built from /root/reference in the build container).
Everything native is compared bit for bit; only dbias (a torch/numpy reduction in Python) is not.
"""
import os

import numpy as np
import pytest

import cases
from oracle.spn_oracle import grid_bounds_torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def bit_equal(a, b, what):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, what
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), what


@pytest.mark.parametrize("name", ["hashgrid_2d", "hashgrid_3d", "hashgrid_1d_clamped"])
def test_golden_hashgrid(oracle, name):
    g = load(name)
    radius, G = float(g["radius"]), int(g["G"])
    locs, D = g["locs"], g["locs"].shape[2]
    low, gd = oracle.grid_bounds(locs, radius, G)
    bit_equal(low, g["low"], "lower_bounds")
    bit_equal(gd, g["grid_dims"], "grid_dims")
    assert np.array_equal(oracle.cell_keys(locs, low, gd, radius), g["keys"].astype(np.int32))
    ids, idxs = oracle.hashgrid_order(locs, low, gd, radius, stable=False)
    bit_equal(ids, g["ref_ids"], "selection-sort ids")
    bit_equal(idxs, g["ref_idxs"], "selection-sort idxs")
    sids, sidx = oracle.hashgrid_order(locs, low, gd, radius, stable=True)
    bit_equal(sids, g["stable_ids"], "stable ids")
    bit_equal(sidx, g["stable_idxs"], "stable idxs")
    nl, nd = oracle.reorder_data(locs, g["data"], sidx, 0)
    bit_equal(nl, g["sorted_locs"], "sorted locs")
    bit_equal(nd, g["sorted_data"], "sorted data")
    bl, bd = oracle.reorder_data(nl, nd, sidx, 1)
    bit_equal(bl, locs, "round trip")
    for inc in (0, 1):
        for K in (4, 64):
            for tag, q in (("q", g["qlocs"]), ("self", nl)):
                co, _, _ = oracle.compute_collisions(q, nl, low, gd, sids, radius, radius, K, inc, G ** D)
                bit_equal(co, g["coll_%s_K%d_s%d" % (tag, K, inc)], "neighbour rows")


@pytest.mark.parametrize("name", ["convsp_ref_test_shape", "convsp_3d_k3", "convsp_3d_k1"])
def test_golden_convsp(oracle, name):
    g = load(name)
    R = float(g["radius"])
    for tag in ("q", "self"):
        q = g["qlocs"] if tag == "q" else g["locs"]
        nb, go = g["nb_" + tag], g["go_" + tag]
        for fn in cases.KERNEL_NAMES:
            for dn in (0, 1):
                key = "%s_%s_n%d" % (tag, fn, dn)
                a = (q, g["locs"], g["data"], nb, g["weight"], g["bias"], R, g["ksize"], g["dil"], dn, fn)
                bit_equal(oracle.convsp_forward(*a), g["fwd_" + key], "fwd " + key)
                dq, dl, dd, dw, _ = oracle.convsp_backward(*a, go)
                for got, nm in ((dq, "dq_"), (dl, "dl_"), (dd, "dd_"), (dw, "dw_")):
                    bit_equal(got, g[nm + key], nm + key)


@pytest.mark.parametrize("name", ["convsdf_3d", "convsdf_2d", "convsdf_1d"])
def test_golden_convsdf(oracle, name):
    g = load(name)
    D = g["locs"].shape[2]
    for md in (0.5, 0.05):
        t = "md%g" % md
        a = (g["locs"], g["idxs"], g["poses"], g["scales"], g["sdfs"], g["offs"], g["shapes"], g["weight"],
             g["bias"], g["ksize"], g["dil"], md)
        bit_equal(oracle.convsdf_forward(*a), g["fwd_" + t], "convsdf fwd")
        dl, dw, dp, _ = oracle.convsdf_backward(*a, g["go_" + t], pose_grads=True)
        bit_equal(dl, g["dl_" + t], "convsdf dlocs")
        bit_equal(dw, g["dw_" + t], "convsdf dweight")
        bit_equal(dp[..., :D], g["dpt_" + t], "convsdf dposes translation")


@pytest.mark.parametrize("name", ["projection_a", "projection_b"])
def test_golden_projections(oracle, name):
    """ParticleProjection / ImageProjection (common_funcs.h:979-1183) against vectors from the reference CPU build."""
    g = load(name)
    fl, std, scale = float(g["fl"]), float(g["std"]), float(g["scale"])
    bit_equal(oracle.particleprojection_forward(g["locs"], fl, std, scale, g["depth_mask"]), g["pp_fwd"], "pp fwd")
    bit_equal(oracle.particleprojection_backward(g["locs"], fl, std, scale, g["depth_mask"], g["pp_go"]), g["pp_dl"],
              "pp dlocs")
    bit_equal(oracle.imageprojection_forward(g["locs"], g["image"], fl, g["depth_mask"]), g["ip_fwd"], "ip fwd")
    dl, di = oracle.imageprojection_backward(g["locs"], g["image"], fl, g["depth_mask"], g["ip_go"])
    bit_equal(dl, g["ip_dl"], "ip dlocs")
    bit_equal(di, g["ip_di"], "ip dimage")
    assert (g["pp_fwd"] != 0).sum() > 100 and (g["ip_fwd"] != 0).sum() > 50


def test_golden_kernel_table(oracle):
    """Kernel ids and formulas: the reference's KERNEL_NAMES / KERNEL_FN (kernels.py:123-131)
    against (a) the oracle's C table, (b) the product's Python table."""
    import smoothparticlenets_b200 as spn
    g = load("kernel_fn")
    names = [str(n) for n in g["names"]]
    assert names == spn.KERNEL_NAMES == cases.KERNEL_NAMES
    for i, n in enumerate(names):
        for j, H in enumerate(g["H"]):
            for k, f in enumerate(g["dfrac"]):
                want = g["values"][i, j, k]
                d = f * H
                scale = np.abs(g["values"][i, j]).max()
                assert np.isclose(spn.KERNEL_FN[n](d, H), want, rtol=1e-12, atol=1e-12 * scale), (n, d, H)
                got = oracle.kernel_w(np.float32(d), np.float32(H), n)
                # python evaluates at double (d, H); the C table at their float32 roundings
                want32 = spn.KERNEL_FN[n](float(np.float32(d)), float(np.float32(H)))
                # fp32 evaluation of polynomials that cancel at d = H: tolerance relative to the
                # largest value of this kernel over its support
                if float(np.float32(d)) <= float(np.float32(H)):
                    assert np.isclose(got, want32, rtol=2e-5, atol=4e-6 * scale), (n, d, H)
    # derivative table against central differences of the value table
    for n in names:
        for H in (0.1, 1.0):
            for f in (0.2, 0.5, 0.8):
                d, e = f * H, 1e-6 * H
                num = (spn.KERNEL_FN[n](d + e, H) - spn.KERNEL_FN[n](d - e, H)) / (2 * e)
                assert np.isclose(spn.DKERNEL_FN[n](d, H), num, rtol=1e-5, atol=1e-6 * abs(num) + 1e-9), n
                got = oracle.kernel_dw(np.float32(d), np.float32(H), n)
                want32 = spn.DKERNEL_FN[n](float(np.float32(d)), float(np.float32(H)))
                assert np.isclose(got, want32, rtol=2e-5, atol=1e-30), (n, d, H)
    assert oracle.kernel_w(2.0, 1.0, "spiky") == 0.0      # beyond the support, common_funcs.h:60
    assert oracle.kernel_w(0.5, 1.0, 99) == -1.0          # unknown id, common_funcs.h:61-65


def test_bounds_match_torch_cpu(oracle):
    """Row a1: the oracle's bounds equal the reference's float32 torch CPU ops
    (ParticleCollision.py:174-181) bit for bit, incl. clamped and degenerate extents."""
    for seed in range(4):
        for D in (1, 2, 3):
            for radius, G, ext in ((0.1, 96, 1.0), (0.037, 96, 2.3), (0.01, 16, 5.0), (0.2, 96, 1e-3)):
                locs, _, _ = cases.collision_case(seed, B=3, N=97, M=1, D=D, C=1, extent=ext)
                locs[0, :, 0] = 0.25  # one degenerate dimension
                low, gd = oracle.grid_bounds(locs, radius, G)
                tl, tg = grid_bounds_torch(locs, radius, G)
                bit_equal(low, tl, "low")
                bit_equal(gd, tg, "grid_dims")


def test_live_reference_agreement(oracle, ref_oracle):
    """Fresh seeds, every stage, C restatement vs the compiled reference, bit for bit."""
    C, R = oracle, ref_oracle
    for seed in (11, 12):
        for D in (1, 2, 3):
            locs, q, data = cases.collision_case(seed, B=2, N=150, M=30, D=D, C=3, extent=1.7)
            radius = 0.11
            lo, gd = C.grid_bounds(locs, radius, 96)
            a, b = C.hashgrid_order(locs, lo, gd, radius, stable=False), R.hashgrid_order(locs, lo, gd, radius)
            bit_equal(a[0], b[0], "ids"), bit_equal(a[1], b[1], "idxs")
            sids, sidx = C.hashgrid_order(locs, lo, gd, radius, stable=True)
            (nl, nd), (nl2, nd2) = C.reorder_data(locs, data, sidx), R.reorder_data(locs, data, sidx)
            bit_equal(nl, nl2, "nlocs"), bit_equal(nd, nd2, "ndata")
            for inc, K in ((0, 8), (1, 64)):
                x = C.compute_collisions(q, nl, lo, gd, sids, radius, radius, K, inc, 96 ** D)
                y = R.compute_collisions(q, nl, lo, gd, sids, radius, radius, K, inc, 96 ** D)
                for u, v in zip(x, y):
                    bit_equal(u, v, "collisions")
    for D, ks, dil in ((2, (3, 1), 0.05), (3, (3, 3, 3), 0.04), (3, (1, 1, 1), 1.0)):
        locs, qlocs, data, weight, bias = cases.convsp_case(21, B=2, N=50, M=9, D=D, C=2, O=3, ksize=ks)
        radius = 0.3
        cr = radius + dil * max((k - 1) / 2 for k in ks)
        lo, gd = C.grid_bounds(locs, cr, 96)
        sids, sidx = C.hashgrid_order(locs, lo, gd, cr)
        nl, nd = C.reorder_data(locs, data, sidx)
        nb, _, _ = C.compute_collisions(qlocs, nl, lo, gd, sids, cr, cr, 32, 1, 96 ** D)
        go = cases.rng(3).rand(2, 9, 3).astype(np.float32)
        ksz, dl = np.array(ks, np.float32), np.full(D, dil, np.float32)
        for fn in cases.KERNEL_NAMES:
            a = (qlocs, nl, nd, nb, weight, bias, radius, ksz, dl, 1, fn)
            bit_equal(C.convsp_forward(*a), R.convsp_forward(*a), "fwd " + fn)
            for u, v in list(zip(C.convsp_backward(*a, go), R.convsp_backward(*a, go)))[:4]:
                bit_equal(u, v, "bwd " + fn)
    for D, ks in ((3, (3, 1, 3)), (2, (3, 3)), (1, (3,))):
        c = cases.convsdf_case(31, B=2, N=60, D=D, S=4, O=2, ksize=ks)
        dil = np.full(D, 0.01, np.float32)
        a = (c["locs"], c["idxs"], c["poses"], c["scales"], c["sdfs"], c["offs"], c["shapes"], c["weight"],
             c["bias"], c["ksize"], dil, 0.3)
        f1, f2 = C.convsdf_forward(*a), R.convsdf_forward(*a)
        bit_equal(f1, f2, "convsdf fwd")
        go = cases.rng(5).rand(*f1.shape).astype(np.float32)
        g1, g2 = C.convsdf_backward(*a, go, pose_grads=True), R.convsdf_backward(*a, go, pose_grads=True)
        bit_equal(g1[0], g2[0], "dlocs"), bit_equal(g1[1], g2[1], "dweight")
        bit_equal(g1[2][..., :D], g2[2][..., :D], "dposes translation")


def test_live_reference_projections(oracle, ref_oracle):
    """ParticleProjection / ImageProjection, fresh seeds and sizes, C restatement vs the compiled reference."""
    for seed, (W, H), std in ((31, (40, 30), 1.3), (32, (33, 57), 3.0), (33, (8, 8), 5.0)):
        c = cases.projection_case(seed, B=3, N=120, W=W, H=H, C=2, fl=25.0)
        a = (c["locs"], 25.0, std, 2.0, c["depth_mask"])
        fwd = oracle.particleprojection_forward(*a)
        bit_equal(fwd, ref_oracle.particleprojection_forward(*a), "pp fwd")
        go = cases.rng(seed).rand(*fwd.shape).astype(np.float32)
        bit_equal(oracle.particleprojection_backward(*a, go), ref_oracle.particleprojection_backward(*a, go), "pp dl")
        b = (c["locs"], c["image"], 25.0, c["depth_mask"])
        f2 = oracle.imageprojection_forward(*b)
        bit_equal(f2, ref_oracle.imageprojection_forward(*b), "ip fwd")
        go = cases.rng(seed + 1).rand(*f2.shape).astype(np.float32)
        for u, v in zip(oracle.imageprojection_backward(*b, go), ref_oracle.imageprojection_backward(*b, go)):
            bit_equal(u, v, "ip grads")


def test_selection_sort_is_a_per_cell_permutation_of_stable(oracle):
    """the CPU reference's order and the stable contract agree on the sorted keys
    and, cell by cell, on the SET of particles."""
    locs, _, _ = cases.collision_case(5, B=2, N=400, M=1, D=3, C=1)
    lo, gd = oracle.grid_bounds(locs, 0.2, 96)
    ids_u, idx_u = oracle.hashgrid_order(locs, lo, gd, 0.2, stable=False)
    ids_s, idx_s = oracle.hashgrid_order(locs, lo, gd, 0.2, stable=True)
    assert np.array_equal(ids_u, ids_s)
    assert not np.array_equal(idx_u, idx_s), "the selection sort is expected to be unstable here"
    for b in range(2):
        for c in np.unique(ids_s[b]):
            m = ids_s[b] == c
            assert set(idx_u[b][m]) == set(idx_s[b][m])
