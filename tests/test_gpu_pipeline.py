"""GraphedStep / PipelinedStep: a captured fluid step fed from pinned host memory must deliver, for every
submitted step, exactly the results of that step's inputs (bit for bit: same graph, same kernels), no matter
how the copies overlap the compute of neighbouring steps."""
import numpy as np
import pytest
import torch

import cases
import fluidstep
import gpu_util as gu
from smoothparticlenets_b200.graph import GraphedStep, PipelinedStep

pytestmark = pytest.mark.gpu


def test_pipelined_steps_match_direct_replays(spn):
    B, N, steps = 2, 4096, 5
    clouds = [cases.fluid_cloud(20 + k, B, N)[:2] for k in range(steps)]
    pins = [[torch.from_numpy(l).pin_memory(), torch.from_numpy(v).pin_memory()] for l, v in clouds]
    model = fluidstep.FluidStep(spn, fused=True).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    gos = [torch.rand(B, N, 3, device="cuda", generator=g) for _ in range(2)]
    step = GraphedStep(lambda l, v: model(l, v), [pins[0][0].cuda(), pins[0][1].cuda()], gos)

    want = []
    for l, v in pins:
        step.inputs[0].detach().copy_(l)
        step.inputs[1].detach().copy_(v)
        step.replay()
        torch.cuda.synchronize()
        want.append([t.detach().cpu().clone() for t in step.outputs + step.grads])

    pipe = PipelinedStep(step, depth=2)
    outs = [[torch.empty(t.shape).pin_memory() for t in step.outputs + step.grads] for _ in range(steps)]
    for k in range(steps):
        pipe.submit(pins[k], outs[k])
    pipe.wait()
    for k in range(steps):
        for a, b in zip(outs[k], want[k]):
            assert torch.equal(a, b), "step %d" % k
    # different inputs really give different results (the comparison above is not vacuous)
    assert not torch.equal(want[0][0], want[1][0])
