"""Timeline probe of PipelinedStep: prints when the memcpys and the first/last kernel of each step ran."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import ProfilerActivity, profile
import cases, fluidstep
import smoothparticlenets_b200 as spn
from smoothparticlenets_b200.graph import GraphedStep, PipelinedStep

B, N = 8, 65536
locs_h, vel_h, _ = cases.fluid_cloud(1000, B, N)
lp, vp = torch.from_numpy(locs_h).pin_memory(), torch.from_numpy(vel_h).pin_memory()
model = fluidstep.FluidStep(spn, fused=True).cuda()
gos = [torch.rand(B, N, 3, device="cuda") for _ in range(2)]
step = GraphedStep(lambda l, v: model(l, v), [lp.cuda(), vp.cuda()], gos)
pipe = PipelinedStep(step)
outs = [[torch.empty(t.shape).pin_memory() for t in step.outputs + step.grads] for _ in range(2)]
def run(k):
    for i in range(k):
        pipe.submit([lp, vp], outs[pipe.next_slot()])
    pipe.wait()
run(3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(5)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in ev)
rows = []
for e in ev:
    nm = e.name
    if "Memcpy" in nm or "memcpy" in nm:
        rows.append((e.time_range.start - t0, e.time_range.end - t0, nm[:40]))
ks = sorted((e.time_range.start - t0, e.time_range.end - t0) for e in ev if "emcpy" not in e.name and "emset" not in e.name)
print("kernels: first start %.0f us, last end %.0f us, count %d" % (ks[0][0], max(k[1] for k in ks), len(ks)))
# gaps > 100us between consecutive kernels = step boundaries
prev = ks[0][0]
start = ks[0][0]
for a, b in ks:
    if a - prev > 100:
        print("  compute burst %.0f .. %.0f us" % (start, prev)); start = a
    prev = max(prev, b)
print("  compute burst %.0f .. %.0f us" % (start, prev))
for r in sorted(rows):
    if r[1] - r[0] > 20:
        print("  memcpy %.0f .. %.0f us (%s)" % r)
