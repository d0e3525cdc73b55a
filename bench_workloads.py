"""Additional bench.py workloads (BASELINE.json configs 3 and 5) and the multi-rank parity leg.

    python bench.py --workload c3 [--particles N]            # ConvSP 64 -> 64, kernel_size 5 (tcgen05 contraction)
    torchrun ... bench.py --workload c5 --gpus N              # one scene of 2^24 particles as spatial slabs
    torchrun ... bench.py --workload c5 --gpus N --check      # slab decomposition vs the single-GPU run: parity_ok

Same JSON contract as the c2 line (bench.py): device-timed steps, max over ranks, clocks, roofline.
"""
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), ("measured (MEASURED_PEAKS.json)" if d else "fallback")


def _time_steps(torch, fn, steps, warmup, barrier):
    for _ in range(max(3, warmup)):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


# ------------------------------------------------------------------------------------------------------------
# c3: ConvSP wide-channel forward and backward (tcgen05 contractions), 64 -> 64, kernel_size 5
# ------------------------------------------------------------------------------------------------------------
def run_c3(args, ClockSampler):
    import torch
    from smoothparticlenets_b200 import build as spn_build
    spn_build.build_library()
    import smoothparticlenets_b200 as spn
    from smoothparticlenets_b200 import _native as nat
    torch.cuda.set_device(0)
    N = args.particles if args.particles != 65536 else (1 << 20)
    M = min(N, args.queries)
    D, C, O, KS, R, DIL = 3, 64, 64, 5, 0.1, 0.025
    coll_r = R + DIL * 2
    g = torch.Generator(device="cuda").manual_seed(3)
    dens = 40.0 / (4.0 / 3.0 * np.pi * coll_r ** 3)       # ~40 particles per list ball
    L = (N / dens) ** (1.0 / 3)
    locs = torch.rand(1, N, D, device="cuda", generator=g) * L
    data = torch.rand(1, N, C, device="cuda", generator=g) - 0.5
    coll = spn.ParticleCollision(D, coll_r, max_grid_dim=160, max_collisions=128).cuda()
    conv = spn.ConvSP(C, O, D, KS, DIL, R, kernel_fn="spiky", with_params=False).cuda()
    conv.weight.copy_((torch.rand(O, C, KS ** 3, device="cuda", generator=g) - 0.5) / 8)
    conv.bias.copy_(torch.rand(O, device="cuda", generator=g))
    with torch.no_grad():
        sl, sd, idxs, nb_all = coll(locs, data)
        # the timed op evaluates M queries (a contiguous block of the sorted particles) against all N particles
        q = sl[:, :M].contiguous()
        nb = nb_all[:, :M].contiguous()
        nbar = float((nb >= 0).sum().item()) / M
        out = conv(sl, sd, nb, q)
    clocks = ClockSampler(0)
    clocks.start()

    def fwd():
        with torch.no_grad():
            conv(sl, sd, nb, q)
    n0 = nat.lib().spnb_launch_count()
    ms = _time_steps(torch, fwd, args.steps, args.warmup, torch.cuda.synchronize)
    launches = (nat.lib().spnb_launch_count() - n0) // (args.steps + max(3, args.warmup)) * args.steps
    clk = clocks.stop()
    # the contraction kernels alone (library knob: skips the gather, multiplies whatever the image buffer holds)
    os.environ["SPNB_WIDE_GEMM_ONLY"] = "1"
    ms_gemm = _time_steps(torch, fwd, args.steps, 1, torch.cuda.synchronize) / args.steps
    del os.environ["SPNB_WIDE_GEMM_ONLY"]
    # backward (all four gradients) on the same queries: dG / dweight contractions + list walk + transposed gather
    conv_g = spn.ConvSP(C, O, D, KS, DIL, R, kernel_fn="spiky").cuda()
    with torch.no_grad():
        conv_g.weight.copy_(conv.weight)
        conv_g.bias.copy_(conv.bias)
    slg, sdg, qg = (t.detach().clone().requires_grad_(True) for t in (sl, sd, q))
    go = torch.rand(out.shape, device="cuda", generator=g)
    bwd_state = {}

    def fwd_g():
        bwd_state["out"] = conv_g(slg, sdg, nb, qg)

    def fwd_bwd():
        fwd_g()
        for t in (slg, sdg, qg, conv_g.weight, conv_g.bias):
            t.grad = None
        bwd_state["out"].backward(go)
    bsteps = max(2, args.steps // 2)
    ms_fwd_g = _time_steps(torch, fwd_g, bsteps, 2, torch.cuda.synchronize) / bsteps
    ms_fb = _time_steps(torch, fwd_bwd, bsteps, 2, torch.cuda.synchronize) / bsteps
    ms_bwd = ms_fb - ms_fwd_g
    del bwd_state["out"]
    # e2e: host positions / features / lists in, output out, every step
    hq, hl, hd, hn = (t.cpu().pin_memory() for t in (q, sl, sd, nb))
    ho = torch.empty(out.shape).pin_memory()

    def e2e():
        a, b_, c, d = (t.cuda(non_blocking=True) for t in (hq, hl, hd, hn))
        with torch.no_grad():
            ho.copy_(conv(b_, c, d, a), non_blocking=True)
    ms_e2e = _time_steps(torch, e2e, max(2, args.steps // 4), 1, torch.cuda.synchronize) / max(2, args.steps // 4)
    hbm, tf, src = _peaks()
    ncells = KS ** 3
    flops = 2.0 * M * ncells * C * O                     # the dense contraction
    # in-radius (neighbour, cell) pairs each cost C multiply-adds in the gather phase
    t_ms = ms / args.steps
    tf32_peak = tf / 2.0                                 # dense TF32 is half the bf16 rate; 3xTF32 a third of that
    line = {
        "metric": "ConvSP wide-channel forward queries/sec (64->64, kernel_size 5)", "value": M / (t_ms * 1e-3),
        "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": t_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 contraction)",
        "data": "synthetic",
        "config": {"workload": "c3 ConvSP 64->64 kernel_size 5, %d particles/scene, %d queries per step, n-bar %.1f, "
                               "radius %.2g dilation %.3g" % (N, M, nbar, R, DIL),
                   "l2": "weights 2 MB (L2 resident by design), features %d MB" % (N * C * 4 >> 20)},
        "clocks": clk,
        "e2e": {"value": M / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": sum(t.numel() * 4 for t in (hq, hl, hd, hn)), "d2h_bytes_per_step": ho.numel() * 4},
        "gpu_launches": int(launches),
        "backward": {"ms_per_step": ms_bwd, "queries_per_s": M / (ms_bwd * 1e-3), "vs_forward": ms_bwd / t_ms,
                     "gradients": "dqlocs, dlocs, ddata, dweight, dbias",
                     "contraction_tflops": 2.0 * flops / (ms_bwd * 1e-3) / 1e12,
                     "note": "backward time = (forward + backward) - forward, CUDA events; the two contractions "
                             "(dG = go*W, dweight = go^T*G) are 2x the forward's dense FLOPs"},
        # the contraction (k_wide_gemm + weight images): 3 TF32 MMAs per product; it streams the A operand images
        # (2 x 32 KB per 128 queries and kernel cell) from HBM, which is what bounds it
        "roofline": {"bound": "tensor", "kernel": "k_wide_gemm", "achieved": flops / (ms_gemm * 1e-3) / 1e12,
                     "peak": tf32_peak / 3.0, "unit": "TFLOP/s", "frac": flops / (ms_gemm * 1e-3) / 1e12 / (tf32_peak / 3.0),
                     "traffic": None, "peak_source": src + ": bf16 dense / 2 (TF32) / 3 (3xTF32 products)",
                     "ms_per_launch_set": ms_gemm, "share_of_step": ms_gemm / t_ms,
                     "hbm_gbs_of_operand_stream": (M * ncells * C * 8.0) / (ms_gemm * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                     "note": "dense-contraction FLOPs (2*ncells*C*O per query) over the time of the contraction kernels "
                             "alone; the step is dominated by the CUDA-core gather (k_wide_gather) that builds the A "
                             "operand: %.1f %% of the step" % (100.0 * (1.0 - ms_gemm / t_ms))},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# c5: one scene as spatial slabs
# ------------------------------------------------------------------------------------------------------------
def _c5_setup(torch, spn, rank, world, N, R, G, K):
    from smoothparticlenets_b200.slab_parallel import SlabScene
    D = 3
    n_local = N // world + (1 if rank < N % world else 0)
    first = rank * (N // world) + min(rank, N % world)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    L = (N / 7640.0) ** (1.0 / 3)
    locs = torch.rand(1, n_local, D, device="cuda", generator=g) * L
    gid = torch.arange(first, first + n_local, device="cuda", dtype=torch.int64)
    coll = spn.ParticleCollision(D, R, max_grid_dim=G, max_collisions=K, include_self=False).cuda()
    bounds_fn = lambda mm: spn.grid_bounds(mm, R, G)
    scene = SlabScene(coll, bounds_fn)
    c1 = spn.ConvSP(1, 1, D, 1, 1, R, kernel_fn="spiky", with_params=False).cuda()
    c3 = spn.ConvSP(3, 3, D, 1, 1, R, dis_norm=True, kernel_fn="dspiky", with_params=False).cuda()
    c1.weight.fill_(1.0)
    c1.bias.zero_()
    c3.weight.copy_(torch.eye(3, device="cuda").view(3, 3, 1))
    c3.bias.zero_()
    return locs, gid, scene, c1, c3


def _c5_step(torch, scene, c1, c3, locs, gid, vel):
    lt = locs.detach().requires_grad_(True)
    vt = vel.detach().requires_grad_(True)
    own_locs, own_vel, own_gid, nb = scene.collide(lt, gid, vt)
    ones = torch.ones(1, scene.m, 1, device=locs.device)
    dens = scene.convsp(c1, ones)
    vsm = scene.convsp(c3, own_vel)
    out = scene.to_origin(torch.cat([dens, vsm], 2))
    out.backward(torch.ones_like(out))
    return out, lt.grad, vt.grad, own_gid, nb


def run_c5(args, ClockSampler):
    import torch
    import torch.distributed as dist
    from smoothparticlenets_b200 import build as spn_build
    spn_build.build_library()
    import smoothparticlenets_b200 as spn
    from smoothparticlenets_b200 import _native as nat
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29512")
    # NCCL logs, incl. the version banner of NCCL_DEBUG=VERSION/WARN, go to stdout by default, and NCCL honours
    # NCCL_DEBUG_FILE only above the VERSION level: keep stdout to the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    if "NCCL_DEBUG_FILE" not in os.environ:
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    N = args.particles if args.particles != 65536 else (1 << 24)
    if args.check:
        N = min(N, 1 << 18)
    R, G, K = 0.1, 160, 128
    locs, gid, scene, c1, c3 = _c5_setup(torch, spn, rank, world, N, R, G, K)
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    vel = torch.rand(1, locs.shape[1], 3, device="cuda", generator=g)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    if args.check:
        # the same scene on ONE GPU (rank 0 gathers everything) vs the slab decomposition over `world` ranks
        out, dl, dv, own_gid, nb = _c5_step(torch, scene, c1, c3, locs, gid, vel)
        sizes = [N // world + (1 if r < N % world else 0) for r in range(world)]
        def gather(t, width):
            parts = [torch.empty(1, n, width, device="cuda") for n in sizes]
            dist.all_gather(parts, t.contiguous())
            return torch.cat(parts, 1)
        all_locs, all_vel = gather(locs, 3), gather(vel, 3)
        all_out, all_dl, all_dv = gather(out.detach(), 4), gather(dl, 3), gather(dv, 3)
        ok, detail = True, {}
        if rank == 0:
            coll = spn.ParticleCollision(3, R, max_grid_dim=G, max_collisions=K, include_self=False).cuda()
            lt = all_locs.clone().requires_grad_(True)
            vt = all_vel.clone().requires_grad_(True)
            sl, sv, idxs, nbr = coll(lt, vt)
            ones = torch.ones(1, N, 1, device="cuda")
            c1.fast_path = c3.fast_path = False
            ref = torch.cat([c1(sl, ones, nbr, qlocs=sl), c3(sl, sv, nbr, qlocs=sl)], 2)
            back = spn.ReorderData(reverse=True)(idxs, ref)
            back.backward(torch.ones_like(back))
            def rel(a, b):
                return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
            detail = {"out": rel(all_out, back.detach()), "dlocs": rel(all_dl, lt.grad), "dvel": rel(all_dv, vt.grad)}
            # rows of rank 0's block: the first m particles of the global order, indices shifted by nl (= 0)
            m0 = scene.m
            rows_ok = bool(torch.equal(nb[0], nbr[0, :m0])) and bool(torch.equal(own_gid, idxs[0, :m0].long()))
            detail["rows_rank0_bit_exact"] = rows_ok
            ok = rows_ok and all(v <= 1e-5 for k, v in detail.items() if k != "rows_rank0_bit_exact")
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, 0)
        if rank == 0:
            print(json.dumps({"metric": "c5 slab decomposition parity", "parity_ok": bool(flag.item()), "n_gpus": world,
                              "particles": N, "max_rel_err": detail, "cuts": scene.cuts,
                              "tolerance": "rows bit-exact; outputs and gradients 1e-5 of the tensor's maximum"}))
        dist.barrier()
        dist.destroy_process_group()
        return

    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.reset_peak_memory_stats()
    n0 = nat.lib().spnb_launch_count()
    step = lambda: _c5_step(torch, scene, c1, c3, locs, gid, vel)
    ms = _time_steps(torch, step, args.steps, args.warmup, barrier)
    launches = (nat.lib().spnb_launch_count() - n0) // (args.steps + max(3, args.warmup)) * args.steps
    clk = clocks.stop()
    peak_mem = torch.cuda.max_memory_allocated()
    halo_rows = scene.nl + scene.nr
    t = torch.tensor([ms, float(peak_mem), float(halo_rows), float(scene.m)], device="cuda", dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax[0])
    # e2e: this rank's share from pinned host memory, results back to the host, every step
    hl, hv = locs.cpu().pin_memory(), vel.cpu().pin_memory()

    def e2e():
        l, v = hl.cuda(non_blocking=True), hv.cuda(non_blocking=True)
        out, dl, dv, _, _ = _c5_step(torch, scene, c1, c3, l, gid, v)
        return out.cpu(), dl.cpu(), dv.cpu()
    ks = max(2, args.steps // 4)
    ms_e2e = _time_steps(torch, e2e, ks, 1, barrier)
    te = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    ms_e2e = float(te[0]) / ks
    if rank == 0:
        hbm, tf, src = _peaks()
        t_ms = ms / args.steps
        nbar = 30.0
        # algorithmic bytes per particle of the step (SURVEY.md 8(d) formulas): search chain + 2 ConvSP fwd+bwd
        D = 3
        fb = lambda C, O: 4 * D + 4 * C + 4 * (nbar + 1) + 4 * O
        bb = lambda C, O: fb(C, O) + 8 * D + 4 * C
        per = 12 + 20 + (4 + 8 * 6) + (4 * D + 4 + 4 * K) + fb(1, 1) + bb(1, 1) + fb(3, 3) + bb(3, 3)
        line = {
            "metric": "single-scene step particles/sec (slab decomposition: collision + 2 ConvSP fwd+bwd)",
            "value": N / (t_ms * 1e-3), "unit": "particles/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": t_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5 one scene of %d particles over %d GPU(s) as dim-0 slabs: bounds all-reduce, layer "
                                   "histogram all-reduce, all_to_all bucket exchange, halo layers, local search, ConvSP "
                                   "1->1 + 3->3 fwd+bwd with halo exchange of features and gradients" % (N, world),
                       "particles_per_rank_max": int(tmax[3]), "halo_rows_per_rank_max": int(tmax[2]),
                       "halo_bytes_per_layer_call": int(tmax[2]) * 4 * 3, "peak_memory_bytes_per_rank_max": int(tmax[1]),
                       "l2": "inputs larger than L2"},
            "clocks": clk,
            "e2e": {"value": N / (ms_e2e * 1e-3), "unit": "particles/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(hl.numel() + hv.numel()) * 4 * world,
                    "d2h_bytes_per_step": int(N * (4 + 3 + 3) * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": per * N / world / (t_ms * 1e-3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": per * N / world / (t_ms * 1e-3) / 1e9 / hbm,
                         "traffic": None, "peak_source": src,
                         "note": "algorithmic bytes per particle (SURVEY.md 8(d)) x particles per rank / step time"},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
