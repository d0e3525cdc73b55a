"""Finite-difference gradient checker with the call signature and pass criteria of the reference's
tests/gradcheck.py (a fork of PyTorch 0.4's gradcheck that cannot be imported on Python 3.12 /
torch 2: ``collections.Iterable``, ``Variable``, ``add_(-1, x)``).  Written from its semantics, not
copied:

* analytic Jacobian: one backward pass per output element with a one-hot ``grad_output``, done TWICE;
  the two Jacobians may differ by at most ``retol`` (re-entrancy, gradcheck.py:104-131);
* numeric Jacobian: finite differences of ``func_numerical`` (default: ``func``), evaluated on double
  copies of the inputs when ``use_double`` (gradcheck.py:67-102).  The reference perturbs through a
  0-dim view that aliases the storage, which makes its difference one-sided; central differences are
  used here (same limit, smaller truncation error);
* pass criterion ``|a - n| <= atol + rtol * |n|`` for every Jacobian entry (gradcheck.py:203);
* a backward pass with an all-zero ``grad_output`` must produce all-zero gradients
  ("backward not multiplied by grad_output", gradcheck.py:209-222).
"""
import numpy as np
import torch


def _as_tuple(x):
    if isinstance(x, tuple):
        return x
    if isinstance(x, list):
        return tuple(x)
    return (x,)


def _diff_inputs(inputs):
    return [t for t in inputs if isinstance(t, torch.Tensor) and t.requires_grad]


def _analytic(inputs, output):
    """Two Jacobians d(output)/d(input) [numel(input), numel(output)] per differentiable input."""
    xs = _diff_inputs(inputs)
    n_out = output.numel()
    jac = [[np.zeros((x.numel(), n_out)) for x in xs] for _ in range(2)]
    sizes_ok = True
    go = torch.zeros_like(output)
    flat = go.view(-1)
    for i in range(n_out):
        flat.zero_()
        flat[i] = 1
        for rep in range(2):
            grads = torch.autograd.grad(output, xs, go, retain_graph=True, allow_unused=True)
            for j, (g, x) in enumerate(zip(grads, xs)):
                if g is None:
                    continue
                if g.size() != x.size():
                    sizes_ok = False
                jac[rep][j][:, i] = g.detach().double().cpu().reshape(-1).numpy()
    reentrant = max([float(np.abs(a - b).max()) for a, b in zip(jac[0], jac[1]) if a.size] + [0.0])
    return jac[0], reentrant, sizes_ok


def _numeric(fn, inputs, eps, use_double):
    work = []
    for t in inputs:
        if isinstance(t, torch.Tensor):
            c = t.detach().clone()
            if use_double:
                c = c.double()
            if isinstance(t, torch.nn.Parameter):  # functions that assign their arguments to a module need this
                c = torch.nn.Parameter(c, requires_grad=t.requires_grad)
            else:
                c.requires_grad_(t.requires_grad)
            work.append(c)
        else:
            work.append(t)
    ev = lambda: fn(work).detach().double().cpu().reshape(-1).numpy().copy()
    n_out = ev().size
    jac = []
    for x in work:
        if not (isinstance(x, torch.Tensor) and x.requires_grad):
            continue
        j = np.zeros((x.numel(), n_out))
        flat = x.data.view(-1)
        for i in range(flat.numel()):
            orig = float(flat[i])
            flat[i] = orig - eps
            a = ev()
            flat[i] = orig + eps
            b = ev()
            flat[i] = orig
            j[i] = (b - a) / (2 * eps)
        jac.append(j)
    return jac


def gradcheck(func, inputs, eps=1e-6, atol=1e-5, rtol=1e-3, retol=1e-4, raise_exception=True,
              func_numerical=None, use_double=False):
    inputs = _as_tuple(inputs)

    def fail(msg):
        if raise_exception:
            raise RuntimeError(msg)
        return False

    outputs = tuple(o for o in _as_tuple(func(*inputs)) if o.requires_grad)
    if func_numerical is None:
        func_numerical = func
    for i, o in enumerate(outputs):
        analytic, reentrant, sizes_ok = _analytic(inputs, o)
        numeric = _numeric(lambda inp: _as_tuple(func_numerical(*inp))[i], inputs, eps, use_double)
        if reentrant > retol:
            return fail("not reentrant, %g exceeded reentrance tolerance of %g." % (reentrant, retol))
        for j, (a, n) in enumerate(zip(analytic, numeric)):
            if a.size == 0 and n.size == 0:
                continue
            bad = ~(np.abs(a - n) <= atol + rtol * np.abs(n))
            if bad.any():
                k = np.unravel_index(np.argmax(np.abs(a - n) * bad), a.shape)
                return fail("for input no. %d: analytical %g vs numerical %g at %s (%d of %d entries off)" % (
                    j, a[k], n[k], k, int(bad.sum()), a.size))
        if not sizes_ok:
            return fail("not correct_grad_sizes")
    # a zero grad_output must give zero gradients
    outputs = tuple(o for o in _as_tuple(func(*inputs)) if o.requires_grad)
    xs = _diff_inputs(inputs)
    if outputs and xs:
        grads = torch.autograd.grad(outputs, xs, [torch.zeros_like(o) for o in outputs], allow_unused=True)
        for g in grads:
            if g is not None and not bool((g == 0).all()):
                return fail("backward not multiplied by grad_output")
    return True
