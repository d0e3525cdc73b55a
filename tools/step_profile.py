"""Per-kernel device time of ONE fluid step (CUDA-graph replay) from torch.profiler (CUPTI).
    python tools/step_profile.py [--layerwise] [--top 25]
"""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import cases  # noqa: E402
import fluidstep  # noqa: E402
import smoothparticlenets_b200 as spn  # noqa: E402
from smoothparticlenets_b200.graph import GraphedStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layerwise", action="store_true")
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()
    B, N = 8, 65536
    locs_h, vel_h, _ = cases.fluid_cloud(1000, B, N)
    locs, vel = torch.from_numpy(locs_h).cuda(), torch.from_numpy(vel_h).cuda()
    model = fluidstep.FluidStep(spn, fused=not args.layerwise).cuda()
    gos = [torch.rand(B, N, 3, device="cuda") for _ in range(2)]
    step = GraphedStep(lambda l, v: model(l, v), [locs, vel], gos)
    for _ in range(3):
        step.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step.replay()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"\(.*", "", ev.name)
            name = re.sub(r"<unnamed>::", "", name)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "spnb" in k or "k_" in k)
    print("one step: %d kernels, %.1f us device time, libspnb %.1f us (%.1f%%)" % (
        sum(a[0] for a in agg.values()), tot, ours, 100 * ours / max(tot, 1e-9)))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("%-86s %4d %9.1f us %7.2f avg %5.1f%%" % (k[:86], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == "__main__":
    main()
