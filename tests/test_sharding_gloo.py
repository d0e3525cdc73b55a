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
