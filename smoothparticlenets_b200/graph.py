"""CUDA-graph capture of a forward+backward step built from the layers of this package.

All libspnb entry points are stream-ordered, allocation-free and free of host synchronisation, so a
whole differentiable step (dozens of ConvSP launches plus the torch elementwise glue between them)
can be captured once and replayed with a single launch -- the B200-native replacement for the
reference's launch-and-synchronise-per-kernel execution (gpu_kernels.cu:124,233,335,546).
"""
import torch


class GraphedStep(object):
    """Captures ``outs = fn(*inputs); grads = d(sum(outs*grad_outs))/d(inputs)`` into one CUDA graph.

    ``inputs`` / ``grad_outputs`` become static device buffers (``self.inputs``, ``self.grad_outputs``):
    copy new data into them, call ``replay()``, read ``self.outputs`` and ``self.grads``.
    """

    def __init__(self, fn, inputs, grad_outputs=None, warmup=3):
        self.inputs = [t.detach().clone().requires_grad_(True) for t in inputs]
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            for _ in range(max(1, warmup)):  # lets every layer size its scratch before capture
                outs = fn(*self.inputs)
                outs = outs if isinstance(outs, (tuple, list)) else (outs,)
                if grad_outputs is None:
                    grad_outputs = [torch.ones_like(o) for o in outs]
                torch.autograd.grad(outs, self.inputs, grad_outputs)
        torch.cuda.current_stream().wait_stream(stream)
        self.grad_outputs = [g.detach().clone() for g in grad_outputs]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            outs = fn(*self.inputs)
            outs = outs if isinstance(outs, (tuple, list)) else (outs,)
            grads = torch.autograd.grad(outs, self.inputs, self.grad_outputs)
        self.outputs = [o.detach() for o in outs]
        self.grads = list(grads)

    def replay(self):
        self.graph.replay()


class PipelinedStep(object):
    """Runs a GraphedStep on a stream of HOST inputs with the PCIe copies overlapped with compute.

    ``submit(host_inputs, host_outputs)`` enqueues one step: the pinned host inputs are copied to a device
    staging slot on a copy stream, the compute stream moves them into the graph's static inputs (a
    device-to-device copy of a few microseconds), replays the graph and parks the results in an output
    staging slot, and a second copy stream writes that slot to the pinned host outputs.  With
    ``depth`` slots the host->device copy of step i+1 and the device->host copy of step i-1 run while step
    i computes (the two directions use separate copy engines).  Every step still consumes inputs that
    were in host memory and delivers its results to host memory; ``wait()`` returns when all submitted
    steps have done so.  Host buffers must be pinned for the copies to be asynchronous.
    """

    def __init__(self, step, depth=2):
        self.step = step
        self.depth = depth
        self.compute = torch.cuda.current_stream()
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
        results = step.outputs + step.grads
        self.in_slots = [[torch.empty_like(t.detach()) for t in step.inputs] for _ in range(depth)]
        self.out_slots = [[torch.empty_like(t) for t in results] for _ in range(depth)]
        ev = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.in_ready, self.in_free, self.out_ready, self.out_free = ev(), ev(), ev(), ev()
        self.count = 0

    def next_slot(self):
        """Index (0..depth-1) of the slot the next submit() uses; callers that keep one set of host output
        buffers per slot pick it with this."""
        return self.count % self.depth

    def submit(self, host_inputs, host_outputs):
        s = self.count % self.depth
        first = self.count < self.depth
        if not first:
            # the host outputs handed to this slot's previous submit are complete (bounds the run-ahead of
            # the host to `depth` steps)
            self.out_free[s].synchronize()
        with torch.cuda.stream(self.h2d):
            if not first:
                self.h2d.wait_event(self.in_free[s])      # compute has consumed this slot's previous contents
            for dst, src in zip(self.in_slots[s], host_inputs):
                dst.copy_(src, non_blocking=True)
            self.in_ready[s].record(self.h2d)
        c = self.compute
        c.wait_event(self.in_ready[s])
        for dst, src in zip(self.step.inputs, self.in_slots[s]):
            dst.detach().copy_(src, non_blocking=True)
        self.in_free[s].record(c)
        self.step.replay()
        if not first:
            c.wait_event(self.out_free[s])                # the previous results of this slot are on the host
        for dst, src in zip(self.out_slots[s], self.step.outputs + self.step.grads):
            dst.copy_(src, non_blocking=True)
        self.out_ready[s].record(c)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.out_ready[s])
            for dst, src in zip(host_outputs, self.out_slots[s]):
                dst.copy_(src, non_blocking=True)
            self.out_free[s].record(self.d2h)
        self.count += 1

    def wait(self):
        self.d2h.synchronize()
        self.compute.synchronize()
