#!/usr/bin/env python
"""bench.py -- fluid-step particles/sec (ConvSP + collision, forward + backward) on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (libspnb, CUDA)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path, host cores

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "c2"): the SPH part of examples/fluid_sim.py --
ParticleCollision + 3 ReorderData + 29 ConvSP, forward and backward -- on 8 scenes x 65536 uniform
random particles per GPU (density 7640 / unit^3, radius 0.1, n-bar ~ 30), fp32, synthetic.
A "step" is one forward + backward pass of that pipeline over the batch.  Scenes are independent, so
N GPUs run N x 8 scenes with no data-path collective ("weak" scaling).

`value`  : whole-job particles/s with inputs resident in HBM, the step replayed as one CUDA graph.
`e2e`    : same step through the public module API with HOST (pinned) inputs and outputs: H2D of
           locs+vel, graph replay, D2H of new locs/vel and their input gradients, every step, run
           through smoothparticlenets_b200.graph.PipelinedStep (the copies of step i+1 / i-1 overlap
           the compute of step i on separate copy streams; every step's data still crosses PCIe).
`roofline`: dominant kernel, algorithmic bytes / CUDA-event time, vs the measured
           HBM copy peak in MEASURED_PEAKS.json.
`cpu_baseline`: the reference's CPU implementation (oracle/_ref, else the C port) on a bounded
           sample, rank 0, N = 1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fluid-step particles/sec (ConvSP+collision fwd+bwd)"
UNIT = "particles/s"
SCENES, PARTICLES, RADIUS, DENSITY, K_NEIGH = 8, 65536, 0.1, 7640.0, 128
CPU_SAMPLE_PARTICLES = 16384


def workload_name(scenes, particles):
    return ("c2 fluid_sim SPH step: ParticleCollision + 3 ReorderData + 29 ConvSP fwd+bwd, "
            "%d scenes x %d particles per GPU, radius %.2g, n-bar~30, K=%d" % (scenes, particles, RADIUS,
                                                                            K_NEIGH))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nme, v in zip(names, p[2:6]):
                if v == "Active":
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU implementation behind the same module API
# ------------------------------------------------------------------------------------------------
def _cpu_scene_step(seed_n):
    seed, n = seed_n
    import torch
    import cases
    import fluidstep
    from oracle import cpu_modules as cm
    torch.set_num_threads(1)
    locs, vel, _ = cases.fluid_cloud(seed, 1, n, density=DENSITY)
    model = fluidstep.FluidStep(cm, radius=RADIUS, max_collisions=K_NEIGH)
    lt = torch.from_numpy(locs).requires_grad_(True)
    vt = torch.from_numpy(vel).requires_grad_(True)
    t0 = time.perf_counter()
    ol, ov = model(lt, vt)
    torch.autograd.grad([ol, ov], [lt, vt], [torch.ones_like(ol), torch.ones_like(ov)])
    return time.perf_counter() - t0, cm.backend().kind


def cpu_baseline_single(n=CPU_SAMPLE_PARTICLES):
    dt, kind = _cpu_scene_step((0, n))
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "1 scene x %d particles (same density/radius/K as the GPU workload), one fluid "
                      "step fwd+bwd, 1 thread, %.1f s" % (n, dt)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n = args.cpu_particles
    # load the reference's compiled CPU extension (oracle/_ref) in THIS process too: the pool workers are forked
    # from it, and the driver's record of loaded native libraries then shows what the arm actually ran
    from oracle import cpu_modules as cm
    parent_kind = cm.backend().kind  # (no autograd here: its threads would not survive the fork below)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_scene_step, [(100 + w * cores + i, n) for i in range(cores)])
        t0 = time.perf_counter()
        kinds = []
        for s in range(args.steps):
            res = pool.map(_cpu_scene_step, [(s * cores + i, n) for i in range(cores)])
            kinds.append(res[0][1])
        dt = time.perf_counter() - t0
    value = cores * n * args.steps / dt
    sample = ("%d independent scenes x %d particles per step (one per host core, same density/radius/K "
              "as the GPU workload), fluid step fwd+bwd" % (cores, n))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": workload_name(SCENES, PARTICLES), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kinds[0], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# kernel-level timing for the roofline entry
# ------------------------------------------------------------------------------------------------
def time_kernels(torch, spn, model, locs, vel, iters=10):
    """Times the individual libspnb launches of the step, on resident data, with CUDA events on the
    launching (current) stream.  Returns {name: (ms_per_launch, launches_per_step, algorithmic bytes
    per launch)} using the per-particle byte counts of SURVEY.md 8(d)."""
    from smoothparticlenets_b200 import _native as nat
    L = nat.lib()
    B, N, D = locs.shape
    with torch.no_grad():
        sl, sv, idxs, nb = model.coll(locs, vel)
    nbar = float((nb >= 0).sum().item()) / (B * N)
    flag = spn.sym_flag_of(nb)
    tiles = spn.tile_lists_of(nb)
    ones = torch.ones(B, N, 1, device="cuda")
    go1, go3 = torch.rand(B, N, 1, device="cuda"), torch.rand(B, N, 3, device="cuda")
    def ev_time(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def fwd(layer, data, O):
        out = torch.empty(B, N, O, device="cuda")
        return lambda: L.spnb_convsp_forward(
            nat.ptr(sl), nat.ptr(sl), nat.ptr(data), nat.ptr(nb), nat.ptr(layer.weight), nat.ptr(layer.bias),
            B, N, N, data.shape[2], D, nb.shape[2], O, 1, float(RADIUS), nat.ptr(layer.kernel_size),
            nat.ptr(layer.dilation), layer.dis_norm, layer.kernel_fn, nat.ptr(out), nat.stream())

    def bwd(layer, data, go):
        dl, dd = torch.empty_like(sl), torch.empty_like(data)
        return lambda: L.spnb_convsp_backward(
            nat.ptr(sl), nat.ptr(sl), nat.ptr(data), nat.ptr(nb), nat.ptr(layer.weight), B, N, N,
            data.shape[2], D, nb.shape[2], go.shape[2], 1, float(RADIUS), nat.ptr(layer.kernel_size),
            nat.ptr(layer.dilation), layer.dis_norm, layer.kernel_fn, nat.ptr(go), nat.ptr(dl), nat.ptr(dl),
            nat.ptr(dd), None, nat.ptr(flag), None, nat.stream())

    P = B * N
    fb = lambda C, O: 4 * D + 4 * C + 4 * (nbar + 1) + 4 * O
    bb = lambda C, O: fb(C, O) + 8 * D + 4 * C
    res = {
        "convsp_fwd_1to1": (ev_time(fwd(model.dspiky1normd, ones, 1)), 16, P * fb(1, 1)),
        "convsp_fwd_3to3": (ev_time(fwd(model.dspikyDnormd, sl, 3)), 13, P * fb(3, 3)),
        "convsp_bwd_1to1": (ev_time(bwd(model.dspiky1normd, ones, go1)), 16, P * bb(1, 1)),
        "convsp_bwd_3to3": (ev_time(bwd(model.dspikyDnormd, sl, go3)), 13, P * bb(3, 3)),
    }

    def collide():
        with torch.no_grad():
            model.coll(locs, vel)
    # whole neighbour-search chain (bounds, sort, reorder, table, lists): 12+20+52+528 B/particle
    res["particle_collision"] = (ev_time(collide), 1, P * (4 * D + (4 * D + 8) + (4 + 8 * (D + D)) + (4 * D + 4 + 4 * K_NEIGH)))
    # the list-building launch alone (cell table + k_collide), 528 B/particle
    low, gd = model.coll.last_lower_bounds, model.coll.last_grid_dims
    coll_out = torch.empty_like(nb)
    tflag = torch.zeros(1, device="cuda", dtype=torch.int32)
    from smoothparticlenets_b200 import sidecar
    pos4 = sidecar.lookup(sl).pos4
    tbytes = L.spnb_tile_lists_bytes(B, N, D, K_NEIGH)
    tbuf = torch.empty(tbytes, device="cuda", dtype=torch.uint8)
    # cell table + k_collide_tiles: float rows and tile lists from one kernel; bytes: positions and keys read,
    # 4K bytes of rows + ~2(n-bar+...) bytes of tile rows written
    res["collide"] = (ev_time(lambda: L.spnb_compute_collisions_tiled(
        nat.ptr(pos4), nat.ptr(sl), nat.ptr(low), nat.ptr(gd), nat.ptr(model.coll.cellIDs),
        nat.ptr(model.coll.cellStarts), nat.ptr(model.coll.cellEnds), nat.ptr(coll_out), B, N, D, K_NEIGH,
        model.coll.max_grid_dim ** D, float(RADIUS), float(RADIUS), 0, nat.ptr(tflag), nat.ptr(tbuf), tbytes,
        nat.stream())), 1, P * (4 * D + 4 + 4 * K_NEIGH))
    if not getattr(model, "fused", False):
        return res, nbar, {}

    # ---- fused groups: one op = pack pre-pass + one walk over the lists (C ABI called directly)
    from smoothparticlenets_b200 import convsp_group as cg
    press = torch.rand(B, N, 1, device="cuda")
    groups = {
        "A": (model.group_a, [None, sl, None, sl, None, None], 3),
        "B": (model.group_b, [sl * press, press], 3),
        "C": (model.group_c, [sv], 3),
        "V": (model.group_v, [sv, None], 1),
    }
    fused, reduced = {}, {}
    for name, (grp, datas, per_step) in groups.items():
        layers = list(grp.layers)
        cfg = tuple((l.kernel_fn, l.dis_norm, l.nchannels, l.nkernels) for l in layers)
        ws_ = [l.weight for l in layers]
        bs_ = [l.bias for l in layers]
        outs = [torch.empty(B, N, c[3], device="cuda") for c in cfg]
        gos = [torch.rand(B, N, c[3], device="cuda") for c in cfg]
        dds = [torch.empty_like(d) if d is not None else None for d in datas]
        dl = torch.empty_like(sl)
        afw = cg._layer_array(sl, datas, ws_, bs_, cfg, outs=outs)
        abw = cg._layer_array(sl, datas, ws_, None, cfg, gos=gos, ddatas=dds)
        wf = L.spnb_convsp_group_workspace_bytes(nat.ptr(sl), B, N, D, float(RADIUS), len(cfg), afw, 0)
        wb = L.spnb_convsp_group_workspace_bytes(nat.ptr(sl), B, N, D, float(RADIUS), len(cfg), abw, 1)
        wsf = torch.empty(wf // 4 + 1, device="cuda")
        wsb = torch.empty(wb // 4 + 1, device="cuda")
        f_fn = lambda afw=afw, wsf=wsf, wf=wf, n=len(cfg): L.spnb_convsp_group_forward(
            nat.ptr(sl), nat.ptr(nb), B, N, D, nb.shape[2], float(RADIUS), n, afw, nat.ptr(wsf), wf, nat.ptr(tiles),
            nat.stream())
        b_fn = lambda abw=abw, wsb=wsb, wb=wb, n=len(cfg), dl=dl: L.spnb_convsp_group_backward(
            nat.ptr(sl), nat.ptr(nb), B, N, D, nb.shape[2], float(RADIUS), n, abw, nat.ptr(dl), nat.ptr(flag),
            nat.ptr(wsb), wb, nat.ptr(tiles), nat.stream())
        # SURVEY 8(d) bytes of the layers this op replaces, and the op's own compulsory bytes
        eq_f = sum(fb(c[2], c[3]) for c in cfg)
        eq_b = sum(bb(c[2], c[3]) for c in cfg)
        distinct = {id(d): d.shape[2] for d in datas if d is not None and d is not sl}
        ch_out = sum(c[3] for c in cfg)
        # the op's own compulsory bytes: positions + distinct data + lists read once, outputs written once
        # (backward: + grad_out of every layer read, d/dlocs and the data gradients written)
        own_f = 4 * D + 4 * (nbar + 1) + 4 * sum(distinct.values()) + 4 * ch_out
        own_b = own_f + 4 * D + sum(4 * d.shape[2] for d, g_ in zip(datas, dds) if g_ is not None)
        fused["group%s_fwd" % name] = (ev_time(f_fn), per_step, P * eq_f)
        fused["group%s_bwd" % name] = (ev_time(b_fn), per_step, P * eq_b)
        reduced["group%s_fwd" % name] = P * own_f
        reduced["group%s_bwd" % name] = P * own_b
    fused["particle_collision"] = res["particle_collision"]
    fused["collide"] = res["collide"]
    return fused, nbar, reduced


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import cases
    import fluidstep
    from smoothparticlenets_b200 import build as spn_build
    spn_build.build_library()
    import smoothparticlenets_b200 as spn
    from smoothparticlenets_b200 import _native as nat
    from smoothparticlenets_b200.graph import GraphedStep, PipelinedStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; smoothparticlenets_b200 has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL logs, incl. the version banner of NCCL_DEBUG=VERSION/WARN, go to stdout by default, and NCCL honours
        # NCCL_DEBUG_FILE only above the VERSION level: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        if "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, N = args.scenes, args.particles
    locs_h, vel_h, _ = cases.fluid_cloud(1000 + rank, B, N, density=DENSITY)
    locs_pin = torch.from_numpy(locs_h).pin_memory()
    vel_pin = torch.from_numpy(vel_h).pin_memory()
    locs, vel = locs_pin.cuda(), vel_pin.cuda()
    fused = not args.layerwise
    model = fluidstep.FluidStep(spn, radius=RADIUS, max_collisions=K_NEIGH, fused=fused).cuda()
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    grad_outs = [torch.rand(B, N, 3, device="cuda", generator=g) for _ in range(2)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- eager warm-up, then capture the whole fwd+bwd step in one CUDA graph
    launches0 = nat.lib().spnb_launch_count()
    step = GraphedStep(lambda l, v: model(l, v), [locs, vel], grad_outs, warmup=max(3, args.warmup))
    torch.cuda.synchronize()
    per_step_launches = (nat.lib().spnb_launch_count() - launches0) // (max(3, args.warmup) + 1)
    for _ in range(2):
        step.replay()
    finite = all(bool(torch.isfinite(t).all()) for t in step.outputs + step.grads)

    # ---- timed region: K graph replays, device time, max over ranks
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step.replay()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # ---- end to end: host inputs -> device -> step -> host outputs, every step
    outs_pin = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in step.outputs + step.grads]
    h2d = sum(t.numel() * 4 for t in (locs_pin, vel_pin))
    d2h = sum(t.numel() * 4 for t in outs_pin)

    # PipelinedStep (smoothparticlenets_b200/graph.py): every step's inputs start in pinned host memory
    # and its results end in pinned host memory; the copies of neighbouring steps overlap the compute.
    pipe = PipelinedStep(step, depth=2)
    outs_pin2 = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in step.outputs + step.grads]
                 for _ in range(2)]

    def e2e_run(k):
        for i in range(k):
            pipe.submit([locs_pin, vel_pin], outs_pin2[pipe.next_slot()])
        pipe.wait()

    ms_e2e = float("nan")
    if not args.lite:
        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
        for a, b_ in zip(outs_pin2[(pipe.count - 1) % 2], step.outputs + step.grads):
            finite = finite and bool(torch.equal(a, b_.cpu()))  # the host copy is the step's result
    clk = clocks.stop()
    if args.lite:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * B * N * args.steps / (ms * 1e-3), "unit": UNIT,
                              "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
                              "lite": True, "note": "profiling run, not a bench value"}))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- eager (no graph) step time, for the record
    def eager():
        l = locs.detach().requires_grad_(True)
        v = vel.detach().requires_grad_(True)
        o = model(l, v)
        torch.autograd.grad(o, [l, v], grad_outs)
    eager()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eager()
    torch.cuda.synchronize()
    ms_eager = 1e3 * (time.perf_counter() - t0) / 3

    # ---- the same step through the per-layer drop-in modules (reported beside the headline)
    ms_alt = None
    if fused:
        alt = fluidstep.FluidStep(spn, radius=RADIUS, max_collisions=K_NEIGH, fused=False).cuda()
        alt_step = GraphedStep(lambda l, v: alt(l, v), [locs, vel], grad_outs, warmup=3)
        alt_step.replay()
        barrier()
        e0.record()
        for _ in range(args.steps):
            alt_step.replay()
        e1.record()
        barrier()
        ms_alt = e0.elapsed_time(e1)

    if dist is not None:
        t = torch.tensor([ms, ms_e2e, ms_alt or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        ms_alt = float(t[2]) if ms_alt is not None else None

    total_particles = world * B * N
    value = total_particles * args.steps / (ms * 1e-3)
    e2e_value = total_particles * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        kern, nbar, reduced = time_kernels(torch, spn, model, locs, vel)
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            peaks = json.load(open(ppath))
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        # "particle_collision" is the whole search chain (9 kernels); the roofline entry is for ONE op, so
        # the chain is listed under "kernels" but its list-building launch ("collide") is the candidate
        shares = {k: v[0] * v[1] for k, v in kern.items() if k != "particle_collision"}
        top = max(shares, key=shares.get)
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, from profiles/
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top, {}).get("dram_bytes_per_launch")
        t_ms, cnt, byts = kern[top]
        own = reduced.get(top, byts)          # the op's OWN compulsory bytes (fused ops: inputs / outputs of the
        achieved = own / (t_ms * 1e-3) / 1e9  # fused op itself, not of the layers it replaces)
        step_bytes = fluidstep.algorithmic_bytes_per_particle_step(nbar) * B * N
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B, N), "scenes_per_gpu": B, "particles_per_scene": N,
                       "nbar": round(nbar, 2), "execution": "one CUDA graph per fwd+bwd step",
                       "convsp_path": ("ConvSPGroup: layers sharing (locs, neighbors) fused per dependency phase"
                                       if fused else "per-layer drop-in modules"),
                       "l2": "inputs larger than L2 (neighbour lists alone are %d MB per GPU)" % (B * N * K_NEIGH * 4 >> 20),
                       "outputs_finite": finite},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(per_step_launches * args.steps),
            # dominant op of the step (largest share of its device time), on its OWN compulsory bytes
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": own, "ms_per_launch": t_ms,
                         "share_of_step": shares[top] / sum(shares.values()),
                         "bytes_definition": ("the fused op's own compulsory bytes: positions, distinct data tensors and "
                                              "the lists read once, grad_out read once, every output written once "
                                              "(pack pre-pass + tile kernel timed together)" if fused else
                                              "SURVEY.md 8(d) per-layer bytes"),
                         # the same time against the SURVEY.md 8(d) bytes of the per-layer calls the op replaces
                         "frac_vs_unfused": (byts / (t_ms * 1e-3) / 1e9 / peak) if top in reduced else None},
            # the whole step against the sum of the SURVEY.md 8(d) bytes of its 29 ConvSP fwd+bwd, search and reorders
            "step_roofline": {"algorithmic_bytes_per_step": step_bytes,
                              "achieved_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                              "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
            "kernels": {k: {"ms": round(v[0], 4), "per_step": v[1],
                            "gbs_own": round(reduced.get(k, v[2]) / (v[0] * 1e-3) / 1e9, 1),
                            "frac_own": round(reduced.get(k, v[2]) / (v[0] * 1e-3) / 1e9 / peak, 4)}
                        for k, v in kern.items()},
            "eager_ms_per_step": ms_eager,
        }
        if ms_alt is not None:
            line["per_layer"] = {"value": total_particles * args.steps / (ms_alt * 1e-3), "unit": UNIT,
                                 "ms_per_step": ms_alt / args.steps,
                                 "note": "same step through the drop-in per-layer ConvSP modules (no ConvSPGroup)"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_particles)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=SCENES)
    ap.add_argument("--particles", type=int, default=PARTICLES)
    ap.add_argument("--cpu-particles", type=int, default=CPU_SAMPLE_PARTICLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lite", action="store_true",
                    help="only warm-up + the timed graph replays (for ncu launch lists): no e2e, no "
                         "per-layer comparison, no kernel timing, no CPU baseline")
    ap.add_argument("--layerwise", action="store_true",
                    help="headline through the per-layer drop-in modules instead of ConvSPGroup")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5"],
                    help="c2 (default, BASELINE.json configs[1], the headline); c3: ConvSP 64->64 kernel_size 5 "
                         "(tcgen05 contraction); c5: one scene of 2^24 particles as spatial slabs over --gpus ranks")
    ap.add_argument("--check", action="store_true",
                    help="c5 only: run the slab decomposition against the single-GPU computation of the same scene "
                         "and print parity_ok instead of timing")
    ap.add_argument("--queries", type=int, default=1 << 17, help="c3: queries evaluated per step")
    args = ap.parse_args()
    if args.workload != "c2" and args.impl == "ours":
        import bench_workloads
        return (bench_workloads.run_c3 if args.workload == "c3" else bench_workloads.run_c5)(args, ClockSampler)
    if args.impl == "reference":
        # bounded sample: each step is one 8192-particle scene per host core (~2.5 s)
        if args.cpu_particles == CPU_SAMPLE_PARTICLES:
            args.cpu_particles = 8192
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
