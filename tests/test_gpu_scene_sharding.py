"""Two-GPU test (NCCL) of the single-scene sharding: sharded neighbour rows and ConvSP outputs /
gradients equal the single-GPU results.  Skipped on boxes with fewer than two GPUs."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import cases
    import smoothparticlenets_b200 as spn
    from smoothparticlenets_b200.scene_parallel import ShardedScene
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    N, D, C = 20000, 3, 3
    locs_h, vel_h, _ = cases.fluid_cloud(3, 1, N)
    locs, vel = torch.from_numpy(locs_h).cuda(), torch.from_numpy(vel_h).cuda()
    coll = spn.ParticleCollision(D, 0.1, include_self=False).cuda()
    conv = spn.ConvSP(C, C, D, 1, 1, 0.1, dis_norm=True, kernel_fn="dspiky").cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.eye(C).view(C, C, 1) + 0.1)
        conv.bias.fill_(0.5)
    go_full = torch.from_numpy(cases.rng(4).rand(1, N, C).astype(np.float32)).cuda()
    # single-GPU truth (every rank computes it itself)
    sl, sv, idxs, nb = coll(locs, vel)
    d_full = sv.detach().clone().requires_grad_(True)
    out_full = conv(sl, d_full, nb)
    out_full.backward(go_full)
    # sharded: halo exchange (default) and the all-gather reference
    oks = []
    for exchange in ("halo", "allgather"):
        scene = ShardedScene(coll, exchange=exchange)
        sl2, idxs2, nb2 = scene.collide(locs)
        assert torch.equal(sl2, sl) and torch.equal(idxs2, idxs)
        assert torch.equal(nb2, nb[:, scene.start:scene.end])
        if exchange == "halo":
            # the neighbours of a slab of the cell-sorted order are a thin band around it
            assert 0 < scene.plan.halo_rows() < N // 4, scene.plan.recvs
        d_loc = scene.local_rows(sv).detach().clone().requires_grad_(True)
        out_loc = scene.convsp(conv, d_loc)
        out_loc.backward(scene.local_rows(go_full).contiguous())
        ok_out = torch.allclose(out_loc, out_full[:, scene.start:scene.end], rtol=1e-5,
                                atol=1e-6 * float(out_full.abs().max()))
        want = d_full.grad[:, scene.start:scene.end]
        ok_grad = torch.allclose(d_loc.grad, want, rtol=1e-5, atol=4e-6 * float(want.abs().max()))
        oks += [bool(ok_out), bool(ok_grad)]
    ret[rank] = (all(oks[0::2]), all(oks[1::2]), scene.start, scene.end)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_scene_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 1000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        res = dict(ret)
    assert res[0][:2] == (True, True) and res[1][:2] == (True, True), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 20000
