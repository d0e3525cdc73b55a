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
