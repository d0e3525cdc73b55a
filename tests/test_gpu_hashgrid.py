"""GPU parity: hash-grid neighbour search through the C ABI.

Bar: bit-exact.  Grid bounds, cell keys, the stable (key, index) order, reordered rows, and the
neighbour rows incl. order, truncation at K and -1 padding must equal the oracle's on the same
inputs.  The CPU reference's own selection sort is unstable, so -- as prescribed
there -- the permutation is compared with the stable contract of the reference GPU path
(oracle.hashgrid_order(stable=True)) and the downstream stages are checked by feeding OUR
permutation to the oracle.
"""
import numpy as np
import pytest
import torch

import cases
import gpu_util as gu

pytestmark = pytest.mark.gpu

CASES = [
    # B, N, M, D, C, extent, radius, G
    (2, 100, 77, 2, 2, 1.0, 0.2, 96),      # tests/test_particlecollision.py shape
    (3, 257, 50, 1, 1, 2.3, 0.037, 96),
    (3, 1000, 333, 3, 3, 1.0, 0.1, 96),
    (2, 5000, 100, 3, 4, 1.0, 0.05, 96),   # 20^3 cells
    (1, 4099, 64, 3, 2, 4.0, 0.03, 16),    # extent clamps: many particles in border cells
    (2, 3000, 10, 2, 2, 1.0, 0.01, 96),    # 3 sort passes' worth of bits in 2-D (96^2 cells)
    (1, 1, 1, 3, 1, 1.0, 0.1, 96),         # single particle: degenerate grid (grid_dims == 0)
    (2, 40, 7, 4, 2, 1.0, 0.3, 8),         # runtime-ndims path
]


@pytest.mark.parametrize("B,N,M,D,C,extent,radius,G", CASES)
def test_search_pipeline_bit_exact(spn, oracle, B, N, M, D, C, extent, radius, G):
    locs, qlocs, data = cases.collision_case(1, B=B, N=N, M=M, D=D, C=C, extent=extent)
    lt, qt, dt = gu.dev(locs), gu.dev(qlocs), gu.dev(data)

    # a1 bounds
    low, gd = gu.grid_bounds(lt, radius, G)
    o_low, o_gd = oracle.grid_bounds(locs, radius, G)
    gu.assert_bit_equal(gu.host(gd), o_gd, "grid_dims")
    gu.assert_bit_equal(gu.host(low), o_low, "lower_bounds")

    # a2 + a3 keys and stable order
    ids, idxs = gu.hashgrid_order(lt, low, gd, radius, G)
    o_ids, o_idxs = oracle.hashgrid_order(locs, o_low, o_gd, radius, stable=True)
    degenerate = bool((o_gd == 0).any())
    got_ids = gu.host(ids).view(np.uint32).astype(np.int64)
    if not degenerate:
        assert np.array_equal(got_ids, o_ids.astype(np.int64)), "sorted cell keys"
        gu.assert_bit_equal(gu.host(idxs), o_idxs, "idxs")
    else:
        # every key is "outside all cells": order must stay the identity (stable)
        assert np.array_equal(gu.host(idxs), np.tile(np.arange(N, dtype=np.float32), (B, 1)))
    my_idxs = gu.host(idxs)
    for b in range(B):
        assert np.array_equal(np.sort(my_idxs[b]), np.arange(N)), "idxs is a permutation"

    # a4 reorder, both directions
    nl, nd = gu.reorder(lt, dt, idxs, 0)
    o_nl, o_nd = oracle.reorder_data(locs, data, my_idxs, 0)
    gu.assert_bit_equal(gu.host(nl), o_nl, "reordered locs")
    gu.assert_bit_equal(gu.host(nd), o_nd, "reordered data")
    rl, rd = gu.reorder(nl, nd, idxs, 1)
    gu.assert_bit_equal(gu.host(rl), locs, "reverse reorder restores locs")
    gu.assert_bit_equal(gu.host(rd), data, "reverse reorder restores data")
    only_l, none = gu.reorder(lt, None, idxs, 0)
    gu.assert_bit_equal(gu.host(only_l), o_nl, "reorder without data")

    if degenerate:
        coll, _ = gu.collisions(qt, nl, low, gd, ids, radius, radius, 8, 1, G)
        assert np.all(gu.host(coll) == -1), "degenerate grid finds nothing (oracle behaviour)"
        return

    # a5 + a6 neighbour rows
    sorted_ids_f = got_ids.astype(np.float32)
    for include_self in (0, 1):
        for K in (4, 128):
            for q_np, q_t in ((qlocs, qt), (o_nl, nl)):
                coll, flag = gu.collisions(q_t, nl, low, gd, ids, radius, radius, K, include_self, G)
                o_coll, _, _ = oracle.compute_collisions(q_np, o_nl, o_low, o_gd, sorted_ids_f, radius,
                                                         radius, K, include_self, G ** D)
                # oracle leaves entries after the terminator at their -1 pre-fill: full rows compare
                gu.assert_bit_equal(gu.host(coll), o_coll, "neighbour rows K=%d self=%d" % (K, include_self))
                # the flag says "the relation may not be symmetric": a row was cut at K, or a query lies two or
                # more cells beyond the upper border of a clamped grid (it then sees no cell at all while the
                # border cell it was hashed into is still scanned by its neighbours)
                full = bool((o_coll[..., K - 1] >= 0).any())
                gq = np.trunc((q_np - o_low[:, None, :]) / np.float32(radius))
                beyond = bool(((o_gd[:, None, :] > 0) & (gq >= o_gd[:, None, :] + 1)).any())
                assert bool(flag.item()) == (full or beyond), "asymmetry flag"


def test_module_api_reference_test_shape(spn, oracle):
    """tests/test_particlecollision.py:36-122 re-run against the drop-in modules."""
    B, N, M, D, R, C = 2, 100, 77, 2, 0.2, 2
    locs, qlocs, data = cases.collision_case(0, B=B, N=N, M=M, D=D, C=C)
    gt = np.ones((B, M, N), dtype=int) * -1
    for b in range(B):
        for i in range(M):
            d = np.square(qlocs[b, i][None] - locs[b]).sum(1)
            js = np.where(d <= R * R)[0]
            gt[b, i, :len(js)] = js
    lt, qt, dt = gu.dev(locs), gu.dev(qlocs), gu.dev(data)
    coll = spn.ParticleCollision(D, R, max_collisions=N).cuda()
    vlocs, vdata, vidxs, vneighbors = coll(lt, dt, qt)
    idxs = gu.host(vidxs).astype(int)
    nb = gu.host(vneighbors).astype(int)
    for b in range(B):
        assert sorted(idxs[b]) == list(range(N))
        assert np.array_equal(locs[b][idxs[b]], gu.host(vlocs)[b])
        assert np.array_equal(data[b][idxs[b]], gu.host(vdata)[b])
    assert np.array_equal(gu.host(lt), locs) and np.array_equal(gu.host(dt), data)
    for b in range(B):
        for i in range(M):
            mine = set(idxs[b][j] for j in nb[b, i] if j >= 0)
            want = set(j for j in gt[b, i] if j >= 0)
            assert mine == want
    reorder = spn.ReorderData(reverse=True).cuda()
    rl, rd = reorder(vidxs, vlocs, vdata)
    assert np.array_equal(gu.host(rl), locs) and np.array_equal(gu.host(rd), data)
    # no-data / no-qlocs return arities
    out = coll(lt)
    assert len(out) == 3 and out[2].shape == (B, N, N)
    assert spn.sym_flag_of(out[2]) is not None


def test_reorder_gradient_is_inverse_permutation(spn):
    """_ReorderDataFunction.backward (ParticleCollision.py:303-314): grads flow back through the
    inverse permutation; collisions/hash order contribute nothing."""
    B, N, D, C = 2, 64, 3, 2
    locs, _, data = cases.collision_case(3, B=B, N=N, M=1, D=D, C=C)
    lt = gu.dev(locs).requires_grad_(True)
    dt = gu.dev(data).requires_grad_(True)
    coll = spn.ParticleCollision(D, 0.3).cuda()
    sl, sd, idxs, nb = coll(lt, dt)
    gl = torch.rand_like(sl)
    gdd = torch.rand_like(sd)
    (sl * gl).sum().backward(retain_graph=True)
    (sd * gdd).sum().backward()
    ii = gu.host(idxs).astype(int)
    want_l = np.zeros_like(locs)
    want_d = np.zeros_like(data)
    for b in range(B):
        want_l[b][ii[b]] = gu.host(gl)[b]
        want_d[b][ii[b]] = gu.host(gdd)[b]
    gu.assert_bit_equal(gu.host(lt.grad), want_l, "dlocs")
    gu.assert_bit_equal(gu.host(dt.grad), want_d, "ddata")


def test_full_size_properties(spn):
    """BASELINE.json config 2 size (8 x 65536, D=3, r=0.1): size-independent properties --
    sortedness, permutation validity, reorder round trip, neighbour symmetry and distances."""
    B, N, R, K = 8, 65536, 0.1, 128
    locs, vel, L = cases.fluid_cloud(0, B, N)
    lt, vt = gu.dev(locs), gu.dev(vel)
    coll = spn.ParticleCollision(3, R, max_collisions=K, include_self=False).cuda()
    sl, sv, idxs, nb = coll(lt, vt)
    keys = coll.cellIDs[:B].view(torch.int32).view(B, N)
    assert bool((keys[:, 1:] >= keys[:, :-1]).all()), "keys sorted"
    assert bool((idxs.sort(1).values == torch.arange(N, device="cuda", dtype=torch.float32)).all())
    rl, rv = spn.ReorderData(reverse=True)(idxs, sl, sv)
    assert torch.equal(rl, lt) and torch.equal(rv, vt)
    # stable: within equal keys the original indices ascend
    same = keys[:, 1:] == keys[:, :-1]
    assert bool((idxs[:, 1:][same] > idxs[:, :-1][same]).all()), "ties by ascending original index"
    assert int(spn.sym_flag_of(nb).item()) == 0
    cnt = (nb >= 0).sum(2)
    assert 20 < float(cnt.float().mean()) < 40, "n-bar ~ 30 at this density"
    # every listed neighbour is within the radius, none is the particle itself
    b = 3
    j = nb[b].long().clamp(min=0)
    d2 = ((sl[b][:, None, :] - sl[b][j]) ** 2).sum(2)
    valid = nb[b] >= 0
    assert bool((d2[valid] < R * R).all()) and bool((d2[valid] > 0).all())
    # symmetry of the relation: i in list(j) <=> j in list(i), checked through degree sums
    deg_out = valid.sum(1)
    deg_in = torch.zeros(N, device="cuda", dtype=torch.long).index_add_(
        0, j[valid], torch.ones(int(valid.sum()), device="cuda", dtype=torch.long))
    assert torch.equal(deg_in, deg_out)


TILE_CASES = [
    # B, N, D, extent, radius, K, include_self, density
    (2, 1000, 3, 1.0, 0.1, 128, 0),
    (3, 777, 2, 1.0, 0.07, 64, 1),
    (2, 300, 1, 2.0, 0.02, 32, 0),
    (1, 5000, 3, 1.0, 0.05, 128, 1),
    (2, 64, 3, 0.3, 0.1, 16, 1),      # K reached: lists cut, flag bit 0
    (1, 4099, 3, 4.0, 0.03, 128, 0),  # border cells hold many particles (G = 16 clamps)
]


@pytest.mark.parametrize("B,N,D,extent,radius,K,include_self", TILE_CASES)
def test_tile_lists_describe_the_same_lists(spn, B, N, D, extent, radius, K, include_self):
    """The compact sidecar (csrc/tile_lists.cuh) must hold exactly the float lists: same entries, same
    order, same truncation; every entry must resolve inside its tile's staged ranges."""
    from smoothparticlenets_b200 import tile_lists as tl
    locs, _, _ = cases.collision_case(11, B=B, N=N, M=1, D=D, C=1, extent=extent)
    G = 16 if extent > 2.0 and D == 3 else 96
    coll = spn.ParticleCollision(D, radius, max_grid_dim=G, max_collisions=K, include_self=bool(include_self)).cuda()
    sl, idxs, nb = coll(gu.dev(locs))
    tiles = spn.tile_lists_of(nb)
    assert tiles is not None and tiles.dtype == torch.uint8
    flag, counts, dec, max_total = tl.decode(tiles, B, N, K)
    nbh = gu.host(nb).astype(np.int64)
    want_cnt = (nbh >= 0).sum(2)
    have = counts >= 0  # blocks with more than 4096 candidates (clamped border cells) carry no tile rows
    assert bool(flag & 2) == (not have.all()), "capacity bit of the tile flag"
    assert np.array_equal(counts[have], want_cnt[have]), "list lengths"
    truncated = bool((nbh[..., K - 1] >= 0).any())
    low, gd, slh = gu.host(coll.last_lower_bounds), gu.host(coll.last_grid_dims), gu.host(sl)
    gq = np.trunc((slh - low[:, None, :]) / np.float32(radius))
    beyond = bool(((gd[:, None, :] > 0) & (gq >= gd[:, None, :] + 1)).any())
    assert bool(flag & 1) == (truncated or beyond), "asymmetry bit of the tile flag"
    assert bool(spn.sym_flag_of(nb).item()) == (truncated or beyond)
    assert not (flag & ~3), "no inconsistency"
    assert have.all() or N == 4099, "only the clamped-grid case has blocks beyond the format's capacity"
    assert np.array_equal(dec[have], nbh[have]), "decoded tile lists == float lists"
    # the plain entry point writes the same float rows
    coll2 = spn.ParticleCollision(D, radius, max_grid_dim=G, max_collisions=K, include_self=bool(include_self)).cuda()
    coll2.tile_lists = False
    sl2, idxs2, nb2 = coll2(gu.dev(locs))
    assert spn.tile_lists_of(nb2) is None
    assert torch.equal(nb, nb2) and torch.equal(idxs, idxs2)


def test_tile_lists_full_size(spn):
    """c2 size: decode a sample of tile blocks and compare with the float rows; flag must be clear."""
    from smoothparticlenets_b200 import tile_lists as tl
    B, N, R, K = 8, 65536, 0.1, 128
    locs, vel, L = cases.fluid_cloud(0, B, N)
    coll = spn.ParticleCollision(3, R, max_collisions=K, include_self=False).cuda()
    sl, sv, idxs, nb = coll(gu.dev(locs), gu.dev(vel))
    tiles = spn.tile_lists_of(nb)
    assert int(tiles[:4].view(torch.int32).item()) == 0
    raw = tiles.cpu().numpy()
    desc, goff, maxcnt, sumcnt = tl.descriptors(raw, B, N, K)
    totals = desc[:, :, 1]
    assert totals.min() > 0 and totals.max() + 1 <= tl.MAX_SLOTS
    assert (totals + 1 > tl.TILE_CAP).mean() < 0.005, "blocks that need more than one staged chunk are rare"
    cnt = gu.host((nb >= 0).sum(2)).reshape(B, -1, tl.TILE_Q)
    assert np.array_equal(sumcnt, cnt.sum(2)) and np.array_equal(maxcnt, cnt.max(2))
    # rows actually stored per block vs the entries they hold: the cost of rank-ordered octiles
    slots = goff[:, :, tl.OCTILES].astype(np.int64).sum() * 32
    assert slots < 1.25 * sumcnt.sum(), "octile padding stays below 25 %% (%.3f)" % (slots / sumcnt.sum())
    blocks = [(5, tb) for tb in range(24)] + [(7, lay) for lay in (1000, 1023)]
    flag, c1, dec, _ = tl.decode(raw, B, N, K, blocks=blocks)
    nbh = gu.host(nb).astype(np.int64)
    for b, tb in blocks:
        sl_ = slice(tb * tl.TILE_Q, (tb + 1) * tl.TILE_Q)
        assert np.array_equal(dec[b, sl_], nbh[b, sl_])


def test_tile_lists_belong_to_their_call(spn):
    """The sidecar is written together with the rows and owns its buffer: a later forward() of the same module
    must not change what an earlier neighbour tensor's sidecar says."""
    from smoothparticlenets_b200 import tile_lists as tl
    B, N = 2, 500
    a, _, _ = cases.collision_case(21, B=B, N=N, M=1, D=3, C=1)
    b, _, _ = cases.collision_case(22, B=B, N=N, M=1, D=3, C=1)
    coll = spn.ParticleCollision(3, 0.1).cuda()
    assert coll.tile_lists is True
    _, _, nb1 = coll(gu.dev(a))
    t1 = spn.tile_lists_of(nb1)
    before = t1.clone()
    _, _, nb2 = coll(gu.dev(b))
    assert spn.tile_lists_of(nb1) is t1 and torch.equal(t1, before)
    assert np.array_equal(tl.decode(t1, B, N, 128)[2], gu.host(nb1).astype(np.int64))
    assert np.array_equal(tl.decode(spn.tile_lists_of(nb2), B, N, 128)[2], gu.host(nb2).astype(np.int64))


@pytest.mark.parametrize("N", [1000, 4096 + 17])
def test_reorder_rows_of_any_width(spn, N):
    """spnb_reorder_data for rows of 1..12 floats (thread-per-row kernel with the shared-memory transpose up to 8,
    piece-per-thread kernel beyond), both directions, against numpy indexing -- bit-exact (a pure copy)."""
    from smoothparticlenets_b200 import _native as nat
    L = nat.lib()
    r = cases.rng(23)
    B = 2
    perm = np.stack([r.permutation(N) for _ in range(B)]).astype(np.float32)
    idx = gu.dev(perm)
    for W in range(1, 13):
        x = r.rand(B, N, W).astype(np.float32)
        xt = gu.dev(x)
        for reverse in (0, 1):
            out = torch.full_like(xt, -1.0)
            nat.check(L.spnb_reorder_data(nat.ptr(xt), None, nat.ptr(idx), nat.ptr(out), None, B, N, W, 0, reverse,
                                          nat.stream()), "spnb_reorder_data")
            want = np.empty_like(x)
            ii = perm.astype(int)
            for b in range(B):
                if reverse:
                    want[b][ii[b]] = x[b]
                else:
                    want[b] = x[b][ii[b]]
            gu.assert_bit_equal(gu.host(out), want, "W=%d reverse=%d" % (W, reverse))
        # locs and data of different widths in one call
        if W <= 6:
            y = r.rand(B, N, W + 2).astype(np.float32)
            yt, oy, ox = gu.dev(y), torch.empty(B, N, W + 2, device="cuda"), torch.empty_like(xt)
            nat.check(L.spnb_reorder_data(nat.ptr(xt), nat.ptr(yt), nat.ptr(idx), nat.ptr(ox), nat.ptr(oy), B, N, W, W + 2,
                                          0, nat.stream()), "spnb_reorder_data")
            for b in range(B):
                gu.assert_bit_equal(gu.host(ox)[b], x[b][perm[b].astype(int)], "locs part")
                gu.assert_bit_equal(gu.host(oy)[b], y[b][perm[b].astype(int)], "data part")
