"""Fused per-particle stages of the Position-Based-Fluids solver iteration (csrc/fluid_glue.cu).

An opt-in extension like ConvSPGroup, not part of the reference API: the
elementwise arithmetic that examples/fluid_sim.py:367-397 writes as ~35 tiny torch ops between the
ConvSP layers of one solver iteration, as three differentiable ops with hand-derived backward
kernels.  ``pbf_stage1/2/3`` return exactly what the torch expressions in their docstrings return
(same operation order per element).  CUDA float32 tensors only; no fallback.
"""
import ctypes

import torch

from . import _native as nat


def _c(t):
    return nat.require_cuda_f32(t.contiguous(), "pbf operand")


def _z(g, like):
    """Gradient or None -> contiguous tensor or None (the kernels treat NULL as zero)."""
    return None if g is None else g.contiguous()


class _Stage1(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, density, nj, ni_s, stiffness, rho0):
        x, density, nj, ni_s = _c(x), _c(density), _c(nj), _c(ni_s)
        BN, D = x.numel() // x.shape[-1], x.shape[-1]
        p, xp, nij = torch.empty_like(density), torch.empty_like(x), torch.empty_like(x)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_stage1_forward(nat.ptr(x), nat.ptr(density), nat.ptr(nj), nat.ptr(ni_s),
                                                        nat.ptr(p), nat.ptr(xp), nat.ptr(nij), BN, D,
                                                        stiffness, rho0, nat.stream()), "spnb_pbf_stage1_forward")
        ctx.save_for_backward(x, density, ni_s)
        ctx.consts = (stiffness, rho0)
        return p, xp, nij

    @staticmethod
    def backward(ctx, g_p, g_xp, g_nij):
        x, density, ni_s = ctx.saved_tensors
        BN, D = x.numel() // x.shape[-1], x.shape[-1]
        g_p, g_xp, g_nij = _z(g_p, density), _z(g_xp, x), _z(g_nij, x)
        g_x, g_nj = torch.empty_like(x), torch.empty_like(x)
        g_density, g_ni_s = torch.empty_like(density), torch.empty_like(density)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_stage1_backward(
                nat.ptr(x), nat.ptr(density), nat.ptr(ni_s), nat.ptr(g_p), nat.ptr(g_xp), nat.ptr(g_nij),
                nat.ptr(g_x), nat.ptr(g_density), nat.ptr(g_nj), nat.ptr(g_ni_s), BN, D, ctx.consts[0],
                ctx.consts[1], nat.stream()), "spnb_pbf_stage1_backward")
        return g_x, g_density, g_nj, g_ni_s, None, None


class _Stage2(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, p, nij, njp, nip_s, nj_c, ni_cs, cohesion, radius, surface_tension, rho0, cscale):
        x, p, nij, njp, nip_s, nj_c, ni_cs = [_c(t) for t in (x, p, nij, njp, nip_s, nj_c, ni_cs)]
        BN, D = x.numel() // x.shape[-1], x.shape[-1]
        d0, nrm = torch.empty_like(x), torch.empty_like(x)
        consts = (cohesion, radius, surface_tension, rho0, cscale)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_stage2_forward(
                nat.ptr(x), nat.ptr(p), nat.ptr(nij), nat.ptr(njp), nat.ptr(nip_s), nat.ptr(nj_c), nat.ptr(ni_cs),
                nat.ptr(d0), nat.ptr(nrm), BN, D, *consts, nat.stream()), "spnb_pbf_stage2_forward")
        ctx.save_for_backward(x, p, nij, nip_s, ni_cs)
        ctx.consts = consts
        return d0, nrm

    @staticmethod
    def backward(ctx, g_d0, g_nrm):
        x, p, nij, nip_s, ni_cs = ctx.saved_tensors
        BN, D = x.numel() // x.shape[-1], x.shape[-1]
        g_d0, g_nrm = _z(g_d0, x), _z(g_nrm, x)
        g_x, g_nij, g_njp, g_nj_c = (torch.empty_like(x) for _ in range(4))
        g_p, g_nip_s, g_ni_cs = (torch.empty_like(p) for _ in range(3))
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_stage2_backward(
                nat.ptr(x), nat.ptr(p), nat.ptr(nij), nat.ptr(nip_s), nat.ptr(ni_cs), nat.ptr(g_d0), nat.ptr(g_nrm),
                nat.ptr(g_x), nat.ptr(g_p), nat.ptr(g_nij), nat.ptr(g_njp), nat.ptr(g_nip_s), nat.ptr(g_nj_c),
                nat.ptr(g_ni_cs), BN, D, *ctx.consts, nat.stream()), "spnb_pbf_stage2_backward")
        return g_x, g_p, g_nij, g_njp, g_nip_s, g_nj_c, g_ni_cs, None, None, None, None, None


class _Stage3(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, d0, cd, nrm, ncount, relaxation, damp):
        x, d0, cd, nrm, ncount = [_c(t) for t in (x, d0, cd, nrm, ncount)]
        BN, D = x.numel() // x.shape[-1], x.shape[-1]
        xnew = torch.empty_like(x)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_stage3_forward(
                nat.ptr(x), nat.ptr(d0), nat.ptr(cd), nat.ptr(nrm), nat.ptr(ncount), nat.ptr(xnew), BN, D,
                relaxation, damp, nat.stream()), "spnb_pbf_stage3_forward")
        ctx.save_for_backward(d0, cd, nrm, ncount)
        ctx.consts = (relaxation, damp)
        return xnew

    @staticmethod
    def backward(ctx, g):
        d0, cd, nrm, ncount = ctx.saved_tensors
        BN, D = d0.numel() // d0.shape[-1], d0.shape[-1]
        g = g.contiguous()
        g_d0, g_nrm, g_ncount = torch.empty_like(d0), torch.empty_like(d0), torch.empty_like(ncount)
        with torch.cuda.device(d0.device):
            nat.check(nat.lib().spnb_pbf_stage3_backward(
                nat.ptr(d0), nat.ptr(cd), nat.ptr(nrm), nat.ptr(ncount), nat.ptr(g), nat.ptr(g_d0), nat.ptr(g_nrm),
                nat.ptr(g_ncount), BN, D, ctx.consts[0], ctx.consts[1], nat.stream()), "spnb_pbf_stage3_backward")
        # d(xnew)/dx = 1 and d/d(cd) = d/d(d0)
        return g, g_d0, g_d0, g_nrm, g_ncount, None, None


class _Integrate(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, v, gravity, dt, cap):
        x, v = _c(x), _c(v)
        BN, D = x.numel() // x.shape[-1], x.shape[-1]
        g = (ctypes.c_float * 3)(*([float(t) for t in gravity] + [0.0] * 3)[:3])
        v2, x1 = torch.empty_like(v), torch.empty_like(x)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_integrate_forward(nat.ptr(x), nat.ptr(v), nat.ptr(v2), nat.ptr(x1), BN, D,
                                                           g, dt, cap, nat.stream()), "spnb_pbf_integrate_forward")
        ctx.save_for_backward(v)
        ctx.consts = (g, dt, cap)
        return v2, x1

    @staticmethod
    def backward(ctx, g_v2, g_x1):
        v, = ctx.saved_tensors
        BN, D = v.numel() // v.shape[-1], v.shape[-1]
        g_v2, g_x1 = _z(g_v2, v), _z(g_x1, v)
        g_v = torch.empty_like(v)
        with torch.cuda.device(v.device):
            nat.check(nat.lib().spnb_pbf_integrate_backward(nat.ptr(v), nat.ptr(g_v2), nat.ptr(g_x1), nat.ptr(g_v), BN,
                                                            D, ctx.consts[0], ctx.consts[1], ctx.consts[2],
                                                            nat.stream()), "spnb_pbf_integrate_backward")
        return g_x1, g_v, None, None, None


class _Velocity(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, xs, dt):
        x, xs = _c(x), _c(xs)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().spnb_pbf_velocity(nat.ptr(x), nat.ptr(xs), nat.ptr(out), None, x.numel(), dt, 0,
                                                  nat.stream()), "spnb_pbf_velocity")
        ctx.dt = dt
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        ga = torch.empty_like(g)
        gb = torch.empty_like(g) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(g.device):
            nat.check(nat.lib().spnb_pbf_velocity(nat.ptr(g), None, nat.ptr(ga), nat.ptr(gb), g.numel(), ctx.dt, 1,
                                                  nat.stream()), "spnb_pbf_velocity")
        return ga, gb, None


class _Viscosity(torch.autograd.Function):

    @staticmethod
    def forward(ctx, w0, vj, vi_s, c):
        w0, vj, vi_s = _c(w0), _c(vj), _c(vi_s)
        BN, D = w0.numel() // w0.shape[-1], w0.shape[-1]
        w1 = torch.empty_like(w0)
        with torch.cuda.device(w0.device):
            nat.check(nat.lib().spnb_pbf_viscosity_forward(nat.ptr(w0), nat.ptr(vj), nat.ptr(vi_s), nat.ptr(w1), BN, D,
                                                           c, nat.stream()), "spnb_pbf_viscosity_forward")
        ctx.save_for_backward(w0, vi_s)
        ctx.c = c
        return w1

    @staticmethod
    def backward(ctx, g):
        w0, vi_s = ctx.saved_tensors
        BN, D = w0.numel() // w0.shape[-1], w0.shape[-1]
        g = g.contiguous()
        g_w0, g_vj, g_vi_s = torch.empty_like(w0), torch.empty_like(w0), torch.empty_like(vi_s)
        with torch.cuda.device(w0.device):
            nat.check(nat.lib().spnb_pbf_viscosity_backward(nat.ptr(w0), nat.ptr(vi_s), nat.ptr(g), nat.ptr(g_w0),
                                                            nat.ptr(g_vj), nat.ptr(g_vi_s), BN, D, ctx.c,
                                                            nat.stream()), "spnb_pbf_viscosity_backward")
        return g_w0, g_vj, g_vi_s, None


class _Fanout(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, n):
        ctx.n = n
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        gs = [g.contiguous() for g in grads if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        out = torch.empty_like(gs[0])
        with torch.cuda.device(out.device):
            first = True
            while gs:
                # up to 8 inputs per launch; further launches add onto the running sum (element-wise in place)
                chunk, gs = (gs[:8], gs[8:]) if first else ([out] + gs[:7], gs[7:])
                first = False
                ptrs = (ctypes.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
                nat.check(nat.lib().spnb_sum_n(ptrs, len(chunk), nat.ptr(out), out.numel(), nat.stream()),
                          "spnb_sum_n")
        return out, None


def fanout(x, n):
    """n aliases of x (views, no copy) for n consumers; the n gradients that come back are added in ONE
    kernel instead of autograd's n-1 pairwise accumulation launches.  Values are untouched."""
    nat.require_cuda_f32(x, "fanout operand")
    return _Fanout.apply(x, int(n))


def pbf_integrate(x, v, gravity, dt, max_speed):
    """v1 = v + gravity * dt;  v2 = v1 * -(relu(-max_speed / (|v1| + 1e-4) + 1) - 1)  (fluid_sim.py:240-245);
    returns (v2, x + v2 * dt)."""
    return _Integrate.apply(x, v, tuple(float(t) for t in gravity), float(dt), float(max_speed))


def pbf_velocity(x, xs, dt):
    """(x - xs) / dt."""
    return _Velocity.apply(x, xs, float(dt))


def pbf_viscosity(w0, vj, vi_s, c):
    """w0 + c * (vj - w0 * vi_s)."""
    return _Viscosity.apply(w0, vj, vi_s, float(c))


def pbf_stage1(x, density, nj, ni_s, stiffness, rest_density):
    """p = stiffness * relu(density - rest_density);  returns (p, x * p, x * ni_s - nj)."""
    return _Stage1.apply(x, density, nj, ni_s, float(stiffness), float(rest_density))


def pbf_stage2(x, p, nij, njp, nip_s, nj_c, ni_cs, cohesion, radius, surface_tension, rest_density,
               constraint_scale):
    """nijp = x * nip_s - njp;  nij2 = x * ni_cs - nj_c;
    returns (-(p * nij + nijp) + -cohesion * nij2 * radius,
             nij2 * surface_tension / rest_density / constraint_scale)."""
    return _Stage2.apply(x, p, nij, njp, nip_s, nj_c, ni_cs, float(cohesion), float(radius),
                         float(surface_tension), float(rest_density), float(constraint_scale))


def pbf_stage3(x, d0, cd, normals, ncount, relaxation, damp):
    """delta = d0 + (cd - normals * ncount);  scale = relu(ncount / (1 + relaxation) - damp) + damp;
    returns x + delta / scale."""
    return _Stage3.apply(x, d0, cd, normals, ncount, float(relaxation), float(damp))
