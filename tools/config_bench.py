"""Timing of the non-headline BASELINE.json configs on one GPU (CUDA events, resident inputs).

    python tools/config_bench.py [c1] [c3] [c4]

c1: ConvSP fwd+bwd, B4 N1024 D3 4->8 k3 dil .05 r .1 spiky (tests/test_convsp.py-style shape)
c3: ConvSP 64->64 kernel_size 5 at 1M particles (generic CUDA-core path; no tensor-core path yet)
c4: ConvSDF with 16 SDF objects + ParticleCollision on 256k particles, 4 scenes (one GPU's share)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
import smoothparticlenets_b200 as spn  # noqa: E402


def ev(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def c1():
    B, N, D, C, O = 4, 1024, 3, 4, 8
    r = cases.rng(0)
    locs = torch.from_numpy(r.rand(B, N, D).astype(np.float32)).cuda()
    data = torch.from_numpy(r.rand(B, N, C).astype(np.float32)).cuda()
    coll = spn.ParticleCollision(D, 0.15).cuda()
    conv = spn.ConvSP(C, O, D, 3, 0.05, 0.1, kernel_fn="spiky").cuda()
    torch.nn.init.uniform_(conv.weight)
    torch.nn.init.uniform_(conv.bias)
    sl, sd, idxs, nb = coll(locs, data)
    sl = sl.detach().requires_grad_(True)
    sd = sd.detach().requires_grad_(True)
    go = torch.rand(B, N, O, device="cuda")

    def fb():
        out = conv(sl, sd, nb)
        torch.autograd.grad(out, [sl, sd, conv.weight, conv.bias], go)
    t_coll = ev(lambda: coll(locs, data))
    t = ev(fb)
    print("c1  ParticleCollision %.3f ms; ConvSP fwd+bwd %.3f ms  -> %.2f M particles/s (fwd+bwd), n-bar %.1f" % (
        t_coll, t, B * N / t / 1e3, float((nb >= 0).sum()) / (B * N)))


def c3(N=1 << 20, M=1 << 14):
    D, C, O = 3, 64, 64
    r = cases.rng(1)
    L = (N / 1910.0) ** (1 / 3.0)
    locs = torch.from_numpy((r.rand(1, N, D) * L).astype(np.float32)).cuda()
    data = torch.randn(1, N, C, device="cuda")
    coll = spn.ParticleCollision(D, 0.2, max_grid_dim=96, max_collisions=128).cuda()
    conv = spn.ConvSP(C, O, D, 5, 0.05, 0.1, kernel_fn="spiky").cuda()
    with torch.no_grad():
        conv.weight.normal_(0, 1.0 / np.sqrt(C * 125))
        conv.bias.zero_()
    t_coll = ev(lambda: coll(locs), iters=3, warm=1)
    qlocs = locs[:, :M].contiguous()
    sl, idxs, nb = coll(locs, qlocs=qlocs)
    sdata = spn.ReorderData()(idxs, locs, data)[1]
    with torch.no_grad():
        t = ev(lambda: conv(sl, sdata, nb, qlocs), iters=2, warm=1)
    nbar = float((nb >= 0).sum()) / M
    flops = 2.0 * C * O * 125 * M
    print("c3  ParticleCollision(1M) %.2f ms; ConvSP 64->64 k5 forward on a %d-query subset %.1f ms "
          "(n-bar %.1f) -> %.3f M queries/s, %.2f TFLOP/s of the dense-contraction count" % (
              t_coll, M, t, nbar, M / t / 1e3, flops / (t * 1e-3) / 1e12))


def c4(B=4, N=262144, S=16):
    D = 3
    r = cases.rng(2)
    locs = torch.from_numpy((r.rand(B, N, D) * 3.25).astype(np.float32)).cuda()
    n = 64
    sdfs = []
    for i in range(S):
        if i % 2:
            sdfs.append(torch.from_numpy(cases.box_sdf(n, 1.0 / n, [0.25] * 3, [0.75] * 3)))
        else:
            sdfs.append(torch.from_numpy(cases.sphere_sdf(n, 1.0 / n, [0.5] * 3, 0.3)))
    idxs = torch.arange(S, dtype=torch.float32).repeat(B, 1).cuda()
    poses = torch.zeros(B, S, 7)
    poses[..., :3] = torch.from_numpy(r.rand(B, S, 3).astype(np.float32)) * 3
    poses[..., 3:] = torch.from_numpy(cases.random_quats(r, (B, S)).astype(np.float32))
    poses = poses.cuda()
    scales = torch.from_numpy((r.rand(B, S) + 0.5).astype(np.float32)).cuda()
    depth = spn.ConvSDF(sdfs, [1.0 / n] * S, 1, D, 1, 1, max_distance=0.5, with_params=False).cuda()
    depth.weight.fill_(-1)
    depth.bias.fill_(0)
    grad = spn.ConvSDF(sdfs, [1.0 / n] * S, 1, D, (3, 1, 1), 5e-4, max_distance=0.5, with_params=False).cuda()
    grad.weight.zero_()
    grad.weight[0, 0], grad.weight[0, 2] = -1000.0, 1000.0
    grad.bias.fill_(0)
    coll = spn.ParticleCollision(D, 0.1).cuda()
    lg = locs.clone().requires_grad_(True)
    go = torch.rand(B, N, 1, device="cuda")

    def sdf_fb(layer):
        def run():
            out = layer(lg, idxs, poses, scales)
            torch.autograd.grad(out, [lg], go)
        return run
    with torch.no_grad():
        t_d = ev(lambda: depth(locs, idxs, poses, scales))
        t_g = ev(lambda: grad(locs, idxs, poses, scales))
    t_dfb, t_gfb = ev(sdf_fb(depth)), ev(sdf_fb(grad))
    t_coll = ev(lambda: coll(locs), iters=3, warm=1)
    inside = float((depth(locs, idxs, poses, scales) > -0.5).float().mean())
    P = B * N
    print("c4  ConvSDF S=16: depth fwd %.3f ms (%.0f M loc/s), fwd+bwd %.3f ms; gradient(k=3x1x1) fwd %.3f ms, "
          "fwd+bwd %.3f ms; ParticleCollision %.3f ms (%.0f M particles/s); %.0f%% of locations within "
          "max_distance of an object" % (t_d, P / t_d / 1e3, t_dfb, t_g, t_gfb, t_coll, P / t_coll / 1e3,
                                        100 * inside))


def c5(N=1 << 24):
    """Single scene of 2^24 particles (the float32 index limit of the API) on ONE GPU."""
    D = 3
    r = cases.rng(3)
    L = (N / 7640.0) ** (1 / 3.0)
    locs = torch.from_numpy((r.rand(1, N, D) * L).astype(np.float32)).cuda()
    vel = torch.rand(1, N, D, device="cuda")
    coll = spn.ParticleCollision(D, 0.1, max_grid_dim=160, include_self=False).cuda()
    t_coll = ev(lambda: coll(locs, vel), iters=2, warm=1)
    sl, sv, idxs, nb = coll(locs, vel)
    keys = coll.cellIDs[:1].view(torch.int32).view(1, N)
    ok_sorted = bool((keys[:, 1:] >= keys[:, :-1]).all())
    ok_perm = bool((idxs.sort(1).values == torch.arange(N, device="cuda", dtype=torch.float32)).all())
    nbar = float((nb >= 0).sum()) / N
    ones = torch.ones(1, N, 1, device="cuda")
    c1 = spn.ConvSP(1, 1, D, 1, 1, 0.1, kernel_fn="spiky", with_params=False).cuda()
    c3_ = spn.ConvSP(3, 3, D, 1, 1, 0.1, dis_norm=True, kernel_fn="dspiky", with_params=False).cuda()
    for c in (c1, c3_):
        c.weight.zero_()
        c.bias.zero_()
        for i in range(c.nchannels):
            c.weight[i, i, 0] = 1
    l = sl.detach().requires_grad_(True)
    go1, go3 = torch.rand(1, N, 1, device="cuda"), torch.rand(1, N, 3, device="cuda")

    def fb(conv, data, go):
        def run():
            out = conv(l, data, nb)
            torch.autograd.grad(out, [l], go)
        return run
    t1, t3 = ev(fb(c1, ones, go1), iters=3, warm=1), ev(fb(c3_, sv, go3), iters=3, warm=1)
    print("c5  N=2^24 one scene, grid %s: ParticleCollision %.1f ms (%.0f M particles/s), sorted=%s perm=%s "
          "n-bar %.1f; ConvSP 1->1 fwd+bwd %.1f ms, 3->3 fwd+bwd %.1f ms; peak memory %.1f GB" % (
              coll.last_grid_dims.tolist(), t_coll, N / t_coll / 1e3, ok_sorted, ok_perm, nbar, t1, t3,
              torch.cuda.max_memory_allocated() / 2 ** 30))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c4"]
    for w in which:
        t0 = time.time()
        {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[w]()
        print("    (%s took %.1f s wall)" % (w, time.time() - t0))
