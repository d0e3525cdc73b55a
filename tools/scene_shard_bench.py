"""Single scene of 2^24 particles split over the GPUs of one box (torchrun), SURVEY.md config c5.

    python -m torch.distributed.run --nproc-per-node N tools/scene_shard_bench.py [--particles 16777216]

Prints, per step of {sharded ParticleCollision, ConvSP 1->1 and 3->3 forward+backward on the owned
slice incl. the NCCL halo exchange (or all-gather / reduce-scatter) of the features}, the max-over-ranks device time."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
import smoothparticlenets_b200 as spn  # noqa: E402
from smoothparticlenets_b200.scene_parallel import ShardedScene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1 << 24)
    ap.add_argument("--exchange", default="halo", choices=["halo", "allgather"])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, D = args.particles, 3
    r = cases.rng(3)
    L = (N / 7640.0) ** (1 / 3.0)
    locs = torch.from_numpy((r.rand(1, N, D) * L).astype(np.float32)).cuda()
    vel = torch.from_numpy(r.rand(1, N, D).astype(np.float32)).cuda()
    coll = spn.ParticleCollision(D, 0.1, max_grid_dim=160, include_self=False).cuda()
    scene = ShardedScene(coll, exchange=args.exchange)

    def timed(fn, iters=3):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    t_coll = timed(lambda: scene.collide(locs))
    sl, idxs, nb = scene.collide(locs)
    sv = spn.ReorderData()(idxs, locs, vel)[1]
    layers = {}
    for name, C, fn, dn in (("1->1 spiky", 1, "spiky", False), ("3->3 dspiky*", 3, "dspiky", True)):
        conv = spn.ConvSP(C, C, D, 1, 1, 0.1, dis_norm=dn, kernel_fn=fn, with_params=False).cuda()
        conv.weight.zero_()
        conv.bias.zero_()
        for i in range(C):
            conv.weight[i, i, 0] = 1
        full = torch.ones(1, N, 1, device="cuda") if C == 1 else sv
        d_loc = scene.local_rows(full).contiguous().requires_grad_(True)
        go = torch.rand(1, scene.end - scene.start, C, device="cuda")

        def fb():
            out = scene.convsp(conv, d_loc)
            torch.autograd.grad(out, [d_loc], go)
        layers[name] = timed(fb)
    if rank == 0:
        halo = scene.plan.halo_rows() if scene.plan is not None else None
        print("c5 sharded scene (%s%s): N=%d over %d GPU(s): ParticleCollision (own rows) %.2f ms; %s" % (
            args.exchange, "" if halo is None else ", %d halo rows on rank 0" % halo, N, world, t_coll,
            "; ".join("ConvSP %s fwd+bwd %.2f ms" % kv for kv in layers.items())))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
