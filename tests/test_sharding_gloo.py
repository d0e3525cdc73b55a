"""world_size-2 gloo test of the multi-GPU host logic (CPU): scenes are split over ranks, every rank
computes its shard independently (here with the oracle-backed CPU modules standing in for the GPU
kernels), parameter gradients are all-reduced, and the result equals the single-process run on the
whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(seed):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from oracle import cpu_modules as cm
    B, N, D, C, O = 4, 120, 3, 2, 3
    r = cases.rng(seed)
    locs = (r.rand(B, N, D) * 0.5).astype(np.float32)
    data = r.rand(B, N, C).astype(np.float32)
    coll = cm.ParticleCollision(D, 0.15)
    conv = cm.ConvSP(C, O, D, 3, 0.05, 0.1, kernel_fn="spiky", with_params=True)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(r.rand(O, C, 27).astype(np.float32)))
        conv.bias.copy_(torch.from_numpy(r.rand(O).astype(np.float32)))
    go = torch.from_numpy(r.rand(B, N, O).astype(np.float32))
    return torch.from_numpy(locs), torch.from_numpy(data), coll, conv, go


def _step(locs, data, coll, conv, go):
    sl, sd, idxs, nb = coll(locs, data)
    sl = sl.detach().requires_grad_(True)
    out = conv(sl, sd, nb)
    out.backward(go)
    return out.detach(), sl.grad.detach()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smoothparticlenets_b200.sharding import scene_shard, allreduce_parameter_grads
    locs, data, coll, conv, go = _build(0)
    mine = scene_shard(locs.shape[0], world, rank)
    sel = torch.tensor(list(mine))
    out, dl = _step(locs[sel], data[sel], coll, conv, go[sel])
    n = allreduce_parameter_grads([conv])
    assert n == 2
    ret[rank] = (list(mine), out.numpy(), dl.numpy(), conv.weight.grad.numpy().copy(), conv.bias.grad.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_scene_shard_partition():
    sys.path.insert(0, ROOT)
    from smoothparticlenets_b200.sharding import scene_shard
    for n in (0, 1, 7, 8, 32):
        for w in (1, 2, 3, 8):
            parts = [list(scene_shard(n, w, r)) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_two_ranks_match_single_process():
    sys.path.insert(0, ROOT)
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    locs, data, coll, conv, go = _build(0)
    out, dl = _step(locs, data, coll, conv, go)
    for rank in range(world):
        scenes, o, d, dw, db = res[rank]
        # per-scene results need no communication and are bit-identical to the full-batch run
        assert np.array_equal(o, out[scenes].numpy())
        assert np.array_equal(d, dl[scenes].numpy())
        # shared-parameter gradients: all-reduced sum == full-batch gradient (fp32 summation order differs)
        np.testing.assert_allclose(dw, conv.weight.grad.numpy(), rtol=2e-5, atol=1e-6 * np.abs(dw).max())
        np.testing.assert_allclose(db, conv.bias.grad.numpy(), rtol=2e-5)
    assert np.array_equal(res[0][3], res[1][3])


def _halo_rows_of(rank, world, N, K, reach):
    """Synthetic banded neighbour rows of the particles owned by `rank`: indices within +-reach."""
    from smoothparticlenets_b200 import scene_parallel as sp
    start, end = sp.owned_range(N, world, rank)
    nb = -torch.ones(1, end - start, K)
    for i in range(start, end):
        js = [j for j in range(i - reach, i + reach + 1, 2) if 0 <= j < N][:K]
        nb[0, i - start, :len(js)] = torch.tensor(js, dtype=torch.float32)
    return nb, start, end


def _halo_worker(rank, world, port, ret):
    """One scene over `world` ranks on CPU: the halo exchange (scene_parallel._HaloRows) must hand every rank
    the rows its lists reference and return to every owner the gradient contributions of ALL ranks -- checked
    against a single-process computation of the same thing (no collectives)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from smoothparticlenets_b200 import scene_parallel as sp
    N, C, K, reach = 101, 3, 8, 7
    g = torch.Generator().manual_seed(5)
    data = torch.rand(1, N, C, generator=g)
    nb, start, end = _halo_rows_of(rank, world, N, K, reach)
    plan = sp.make_halo_plan(nb, N, start, end)
    assert 0 < plan.halo_rows() <= 2 * reach, (plan.recvs, plan.sends)
    layer = lambda full, rows, q: (full[0, rows[rows >= 0].long()] * (q + 1.5)).pow(2).sum()
    x = data[:, start:end].clone().requires_grad_(True)
    full = sp._HaloRows.apply(x, plan)
    used = nb[nb >= 0].long()
    got_rows = full[0, used].detach().clone()
    layer(full, nb, rank).backward()
    # reference: the whole scene in this process, the losses of all ranks summed
    ref = data.clone().requires_grad_(True)
    sum(layer(ref, _halo_rows_of(q, world, N, K, reach)[0], q) for q in range(world)).backward()
    ret[rank] = (bool(torch.equal(got_rows, data[0, used])),
                 bool(torch.allclose(x.grad, ref.grad[:, start:end], rtol=1e-6, atol=1e-6)),
                 plan.halo_rows())
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_matches_single_process_on_three_ranks():
    world = 3
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_halo_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    for rank in range(world):
        same_rows, same_grads, halo = res[rank]
        assert same_rows and same_grads, (rank, res[rank])


# ---- spatial slabs: positions sharded (smoothparticlenets_b200/slab_parallel.py) ------------------------------
def _slab_case():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    N, D, C, O, R = 900, 3, 2, 3, 0.1
    r = cases.rng(12)
    locs = (r.rand(1, N, D) * np.array([1.3, 0.5, 0.5])).astype(np.float32)   # ~13 cell layers along dim 0
    data = r.rand(1, N, C).astype(np.float32)
    w = r.rand(O, C, 1).astype(np.float32)
    go = r.rand(1, N, O).astype(np.float32)
    return N, D, C, O, R, locs, data, w, go


def _slab_modules(R, C, O, D, w):
    from oracle import cpu_modules as cm
    from oracle import spn_oracle as so
    coll = cm.ParticleCollision(D, R, max_collisions=64, include_self=False)
    conv = cm.ConvSP(C, O, D, 1, 1, R, kernel_fn="spiky", with_params=False)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(w))
    bounds_fn = lambda mm: tuple(torch.from_numpy(a) for a in so.grid_bounds_torch(mm.numpy(), R, coll.max_grid_dim))
    return cm, coll, conv, bounds_fn


def _slab_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import cpu_modules as cm
    cm.STABLE_ORDER = True   # the product's stable order (ties by ascending index), which the decomposition relies on
    from smoothparticlenets_b200.slab_parallel import SlabScene
    N, D, C, O, R, locs, data, w, go = _slab_case()
    _, coll, conv, bounds_fn = _slab_modules(R, C, O, D, w)
    # an arbitrary initial distribution: particle i starts on rank (7 i) % world
    mine = np.array([i for i in range(N) if (7 * i) % world == rank])
    lt = torch.from_numpy(locs[:, mine]).requires_grad_(True)
    dt = torch.from_numpy(data[:, mine]).requires_grad_(True)
    scene = SlabScene(coll, bounds_fn)
    own_locs, own_data, own_gid, nbrs = scene.collide(lt, torch.from_numpy(mine), dt)
    out_own = scene.convsp(conv, own_data)
    out_back = scene.to_origin(out_own)
    out_back.backward(torch.from_numpy(go[:, mine]))
    ret[rank] = dict(mine=mine, gid=own_gid.numpy(), nbrs=nbrs.numpy(), nl=scene.nl, m=scene.m, cuts=scene.cuts,
                     out=out_back.detach().numpy(), dlocs=lt.grad.numpy(), ddata=dt.grad.numpy(),
                     own_locs=own_locs.detach().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_slab_decomposition_matches_single_process_on_three_ranks():
    """Positions sharded over 3 ranks (bucket exchange, halo layers, local search on the global grid): the own
    neighbour rows equal the single-process rows up to the block's index shift, outputs and the gradients that
    come back to the particles' original owners equal the single-process ones."""
    world = 3
    port = 33500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_slab_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    from oracle import cpu_modules as cm
    cm.STABLE_ORDER = True
    try:
        N, D, C, O, R, locs, data, w, go = _slab_case()
        _, coll, conv, _ = _slab_modules(R, C, O, D, w)
        lt = torch.from_numpy(locs).requires_grad_(True)
        dt = torch.from_numpy(data).requires_grad_(True)
        sl, sd, idxs, nb = coll(lt, dt)
        out = conv(sl, sd, nb)
        back = cm.ReorderData(reverse=True)(idxs, out)
        back.backward(torch.from_numpy(go))
    finally:
        cm.STABLE_ORDER = False
    gidx = idxs[0].numpy().astype(np.int64)          # sorted position -> global id
    nbh = nb[0].numpy()
    start = 0
    assert sum(res[r]["m"] for r in range(world)) == N
    for r in range(world):
        d = res[r]
        m, nl = d["m"], d["nl"]
        assert m > 0 and d["cuts"] == res[0]["cuts"]
        # the own block is the next m particles of the global cell-sorted order ...
        assert np.array_equal(d["gid"], gidx[start:start + m])
        assert np.array_equal(d["own_locs"][0], sl[0].detach().numpy()[start:start + m])
        # ... and its rows are the global rows shifted by (start - nl)
        want = nbh[start:start + m]
        shifted = np.where(want >= 0, want - (start - nl), -1)
        assert np.array_equal(d["nbrs"][0], shifted)
        np.testing.assert_allclose(d["out"][0], back[0].detach().numpy()[d["mine"]], rtol=1e-6, atol=1e-6)
        # gradients of rows lent to a neighbour rank come back over the halo link and are added last: the same
        # terms in another order (fp32)
        gl, gdt = lt.grad[0].numpy(), dt.grad[0].numpy()
        np.testing.assert_allclose(d["dlocs"][0], gl[d["mine"]], rtol=1e-4, atol=1e-6 * np.abs(gl).max())
        np.testing.assert_allclose(d["ddata"][0], gdt[d["mine"]], rtol=1e-4, atol=1e-6 * np.abs(gdt).max())
        start += m
