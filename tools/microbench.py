"""Micro-benchmark of individual libspnb launches on the c2 workload (8 x 65536 particles).

    python tools/microbench.py [--only fwd1,fwd3,bwd1,bwd3,collide,sort,reorder] [--iters 20]

Prints ms per launch, algorithmic GB/s (SURVEY.md 8(d) byte counts) and fraction of the measured HBM
peak.  Meant to be wrapped in ncu for a single kernel:
    ncu --set full --clock-control none --import-source on -k regex:k_convsp_fwd_small -s 5 -c 1 \
        -o gpurun_out/prof python tools/microbench.py --only fwd1 --iters 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import cases  # noqa: E402
import fluidstep  # noqa: E402
import smoothparticlenets_b200 as spn  # noqa: E402
from smoothparticlenets_b200 import _native as nat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="fwd1,fwd3,bwd1,bwd3,collide,search,reorder,gA_fwd,gA_fb,gB_fwd,gB_fb,gV_fwd,gC_fwd,gC_fb")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--particles", type=int, default=65536)
    ap.add_argument("--kernel", default="dspiky")
    ap.add_argument("--nosym", action="store_true")
    ap.add_argument("--graph", action="store_true", help="time a CUDA-graph replay (no CPU launch overhead)")
    args = ap.parse_args()
    only = set(args.only.split(","))
    B, N, D, R, K = args.scenes, args.particles, 3, 0.1, 128
    locs_h, vel_h, _ = cases.fluid_cloud(1000, B, N)
    locs, vel = torch.from_numpy(locs_h).cuda(), torch.from_numpy(vel_h).cuda()
    model = fluidstep.FluidStep(spn, radius=R, max_collisions=K).cuda()
    L = nat.lib()
    with torch.no_grad():
        sl, sv, idxs, nb = model.coll(locs, vel)
    nbar = float((nb >= 0).sum().item()) / (B * N)
    flag = None if args.nosym else spn.sym_flag_of(nb)
    peak = 6550.1
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = json.load(open(pp))["hbm_gbs"]
    ones = torch.ones(B, N, 1, device="cuda")
    go1, go3 = torch.rand(B, N, 1, device="cuda"), torch.rand(B, N, 3, device="cuda")
    l1 = getattr(model, args.kernel + "1" + ("normd" if args.kernel in ("dspiky", "cohesion") else ""))
    l3 = getattr(model, args.kernel + "D" + ("normd" if args.kernel in ("dspiky", "cohesion") else ""))

    def fwd(layer, data, O):
        out = torch.empty(B, N, O, device="cuda")
        return lambda: L.spnb_convsp_forward(
            nat.ptr(sl), nat.ptr(sl), nat.ptr(data), nat.ptr(nb), nat.ptr(layer.weight), nat.ptr(layer.bias),
            B, N, N, data.shape[2], D, K, O, 1, R, nat.ptr(layer.kernel_size), nat.ptr(layer.dilation),
            layer.dis_norm, layer.kernel_fn, nat.ptr(out), nat.stream())

    def bwd(layer, data, go):
        dl, dd = torch.empty_like(sl), torch.empty_like(data)
        return lambda: L.spnb_convsp_backward(
            nat.ptr(sl), nat.ptr(sl), nat.ptr(data), nat.ptr(nb), nat.ptr(layer.weight), B, N, N,
            data.shape[2], D, K, go.shape[2], 1, R, nat.ptr(layer.kernel_size), nat.ptr(layer.dilation),
            layer.dis_norm, layer.kernel_fn, nat.ptr(go), nat.ptr(dl), nat.ptr(dl), nat.ptr(dd), None,
            nat.ptr(flag), None, nat.stream())

    def search():
        with torch.no_grad():
            model.coll(locs, vel)

    low, gd = model.coll.last_lower_bounds, model.coll.last_grid_dims
    coll_out = torch.empty_like(nb)
    tflag = torch.zeros(1, device="cuda", dtype=torch.int32)

    def collide():
        L.spnb_compute_collisions(nat.ptr(sl), nat.ptr(sl), nat.ptr(low), nat.ptr(gd), nat.ptr(model.coll.cellIDs),
                                  nat.ptr(model.coll.cellStarts), nat.ptr(model.coll.cellEnds), nat.ptr(coll_out),
                                  B, N, N, D, K, 96 ** 3, R, R, 0, nat.ptr(tflag), nat.stream())

    ro_l, ro_v = torch.empty_like(locs), torch.empty_like(vel)

    def reorder():
        L.spnb_reorder_data(nat.ptr(locs), nat.ptr(vel), nat.ptr(idxs), nat.ptr(ro_l), nat.ptr(ro_v), B, N, D, 3, 0, nat.stream())

    # fused groups (ConvSPGroup) of the fluid step: forward and forward+backward through autograd
    model_f = fluidstep.FluidStep(spn, radius=R, max_collisions=K, fused=True).cuda()
    press = torch.rand(B, N, 1, device="cuda")

    def group_fwd(group, mk):
        def run():
            with torch.no_grad():
                group(sl, mk(sl), nb)
        return run

    def group_fb(group, mk):
        l = sl.detach().clone().requires_grad_(True)
        outs0 = group(l, mk(l), nb)
        gos = [torch.rand_like(o) for o in outs0]

        def run():
            outs = group(l, mk(l), nb)
            torch.autograd.grad(outs, [l], gos)
        return run

    mkA = lambda l: [ones, l, ones, l, ones, ones]
    mkB = lambda l: [l * press, press]
    mkV = lambda l: [sv, ones]
    mkC = lambda l: [sv]

    P = B * N
    fb = lambda C, O: 4 * D + 4 * C + 4 * (nbar + 1) + 4 * O
    bb = lambda C, O: fb(C, O) + 8 * D + 4 * C
    table = {
        "fwd1": (fwd(l1, ones, 1), P * fb(1, 1)), "fwd3": (fwd(l3, sl, 3), P * fb(3, 3)),
        "bwd1": (bwd(l1, ones, go1), P * bb(1, 1)), "bwd3": (bwd(l3, sl, go3), P * bb(3, 3)),
        "collide": (collide, P * (4 * D + 4 + 4 * K)),
        "search": (search, P * (12 + 20 + 52 + 528)),
        "reorder": (reorder, P * 52),
        "gA_fwd": (group_fwd(model_f.group_a, mkA), P * (2 * fb(1, 1) * 2 + 2 * fb(3, 3))),
        "gA_fb": (group_fb(model_f.group_a, mkA), P * (4 * (fb(1, 1) + bb(1, 1)) + 2 * (fb(3, 3) + bb(3, 3)))),
        "gB_fwd": (group_fwd(model_f.group_b, mkB), P * (fb(1, 1) + fb(3, 3))),
        "gB_fb": (group_fb(model_f.group_b, mkB), P * (fb(1, 1) + bb(1, 1) + fb(3, 3) + bb(3, 3))),
        "gV_fwd": (group_fwd(model_f.group_v, mkV), P * (fb(1, 1) + fb(3, 3))),
        "gC_fwd": (group_fwd(model_f.group_c, mkC), P * fb(3, 3)),
        "gC_fb": (group_fb(model_f.group_c, mkC), P * (fb(3, 3) + bb(3, 3))),
    }
    # the group ops called directly through the C ABI (pack + tile kernel + fallback), no autograd around them
    from smoothparticlenets_b200 import convsp_group as cg
    tiles = spn.tile_lists_of(nb)
    if os.environ.get("SPNB_MB_NOTILES"):
        tiles = None

    def direct(group, datas):
        layers = list(group.layers)
        cfg = tuple((l.kernel_fn, l.dis_norm, l.nchannels, l.nkernels) for l in layers)
        ws_ = [l.weight for l in layers]
        outs = [torch.empty(B, N, c[3], device="cuda") for c in cfg]
        gos = [torch.rand(B, N, c[3], device="cuda") for c in cfg]
        dds = [torch.empty_like(d) if d is not ones else None for d in datas]
        dl = torch.empty_like(sl)
        afw = cg._layer_array(sl, datas, ws_, [l.bias for l in layers], cfg, outs=outs)
        abw = cg._layer_array(sl, datas, ws_, None, cfg, gos=gos, ddatas=dds)
        wf = L.spnb_convsp_group_workspace_bytes(nat.ptr(sl), B, N, D, float(R), len(cfg), afw, 0)
        wb = L.spnb_convsp_group_workspace_bytes(nat.ptr(sl), B, N, D, float(R), len(cfg), abw, 1)
        wsf, wsb = torch.empty(wf // 4 + 1, device="cuda"), torch.empty(wb // 4 + 1, device="cuda")
        keep = (outs, gos, dds, dl, afw, abw, wsf, wsb, datas)
        f = lambda: (keep, L.spnb_convsp_group_forward(nat.ptr(sl), nat.ptr(nb), B, N, D, K, float(R), len(cfg), afw,
                                                       nat.ptr(wsf), wf, nat.ptr(tiles), nat.stream()))
        g = lambda: (keep, L.spnb_convsp_group_backward(nat.ptr(sl), nat.ptr(nb), B, N, D, K, float(R), len(cfg), abw,
                                                        nat.ptr(dl), nat.ptr(flag), nat.ptr(wsb), wb, nat.ptr(tiles),
                                                        nat.stream()))
        eq_f = P * sum(fb(c[2], c[3]) for c in cfg)
        eq_b = P * sum(bb(c[2], c[3]) for c in cfg)
        return (f, eq_f), (g, eq_b)

    for nm, grp, datas in (("A", model_f.group_a, mkA(sl)), ("B", model_f.group_b, mkB(sl)),
                           ("C", model_f.group_c, mkC(sl)), ("V", model_f.group_v, mkV(sl))):
        table["k%s_f" % nm], table["k%s_b" % nm] = direct(grp, datas)
    if "wide" in only:
        # config 3: 64 -> 64, kernel_size 5, collision radius 0.2 (n-bar ~ 63), 16384 queries of a 1M scene
        import numpy as np
        Nw, Mw = 1 << 20, 1 << 14
        rr = cases.rng(1)
        Lw = (Nw / 1910.0) ** (1 / 3.0)
        wl = torch.from_numpy((rr.rand(1, Nw, 3) * Lw).astype(np.float32)).cuda()
        wd = torch.randn(1, Nw, 64, device="cuda")
        wcoll = spn.ParticleCollision(3, 0.2).cuda()
        wconv = spn.ConvSP(64, 64, 3, 5, 0.05, 0.1, kernel_fn="spiky").cuda()
        with torch.no_grad():
            wconv.weight.normal_(0, 0.01)
            wconv.bias.zero_()
            wq = wl[:, :Mw].contiguous()
            wsl, widx, wnb = wcoll(wl, qlocs=wq)
            wsd = spn.ReorderData()(widx, wl, wd)[1]

        def wide():
            with torch.no_grad():
                wconv(wsl, wsd, wnb, wq)
        table["wide"] = (wide, Mw * 4 * (3 + 64 + 64 + 64))
    print("nbar %.2f  peak %.0f GB/s" % (nbar, peak))
    for name, (fn, byts) in table.items():
        if name not in only:
            continue
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if args.graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(args.iters):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0.record()
            gr.replay()
            e1.record()
        else:
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        gbs = byts / (ms * 1e-3) / 1e9
        print("%-8s %8.4f ms  %8.1f GB/s algorithmic  %.3f of peak" % (name, ms, gbs, gbs / peak))


if __name__ == "__main__":
    main()
