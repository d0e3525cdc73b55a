"""The SPH part of the reference's Position-Based-Fluids step, as a benchmark / test driver.

Re-creates the data flow of examples/fluid_sim.py:355-424 of the reference (FluidSim.forward minus
the ConvSDF static-collision passes): gravity + velocity cap, ONE ParticleCollision carrying the
velocities, numIterations = 3 solver iterations of 9 ConvSP layers each, the velocity update through
ReorderData, 2 viscosity ConvSP layers and the final ReorderData(reverse) -- 29 ConvSP (16 with
C=O=1, 13 with C=O=3, all kernel_size 1), 1 collision, 3 reorders per step.

It is written against a module namespace `ns` that provides ConvSP / ParticleCollision /
ReorderData with the reference's signatures, so the same code drives the product
(smoothparticlenets_b200, CUDA) and -- for the CPU baseline only -- the oracle-backed CPU modules
(oracle/cpu_modules.py).  This file is a driver, not part of the product package.
"""
import numpy as np
import torch
import torch.nn as nn

# constants of examples/fluid_sim.py:18-36
DT = 1.0 / 60
GRAVITY = [0, -9.8, 0]
COHESION = 0.1
VISCOSITY = 60.0
SURFACE_TENSION = 0.0
SURFACE_CONSTRAINT_SCALE = 168628.0
NUM_ITERATIONS = 3
RELAXATION = 1.0
DAMP = 1.0
FLUID_REST_DISTANCE = 0.55

# (kernel, channels = 1 or ndim, dis_norm) -- fluid_sim.py:156-165
LAYER_TYPES = [
    ('dspiky', 1, True), ('dspiky', 'D', True), ('constant', 1, False), ('constant', 'D', False),
    ('spiky', 1, False), ('spiky', 'D', False), ('cohesion', 1, True), ('cohesion', 'D', True),
]

CONVSP_PER_STEP = 29
CONVSP_1TO1_PER_STEP = 16
CONVSP_DTOD_PER_STEP = 13


def tight_pack3d(radius, separation, max_points=2048):
    """fluid_sim.py:207-228: hexagonal close packing around the origin, origin excluded."""
    dim = int(np.ceil(1.0 * radius / separation))
    pts = []
    for z in range(-dim, dim + 1):
        for y in range(-dim, dim + 1):
            for x in range(-dim, dim + 1):
                xp = x * separation + (separation * 0.5 if ((y + z) & 1) else 0.0)
                yp = y * np.sqrt(0.75) * separation
                zp = z * np.sqrt(0.75) * separation
                r2 = xp ** 2 + yp ** 2 + zp ** 2
                if r2 == 0.0:
                    continue
                if len(pts) < max_points and np.sqrt(r2) <= radius:
                    pts.append([xp, yp, zp])
    return np.array(pts)


def rest_density(kernel_fn, radius):
    """fluid_sim.py:195-205: rest density and stiffness from a tight packing."""
    d = np.sqrt((tight_pack3d(radius, FLUID_REST_DISTANCE * radius) ** 2).sum(1))
    rho = sum(kernel_fn["spiky"](x, radius) for x in d)
    rhoderiv = sum(kernel_fn["dspiky"](x, radius) ** 2 for x in d)
    return float(rho), float(1.0 / rhoderiv)


class FluidStep(nn.Module):
    def __init__(self, ns, radius=0.1, ndim=3, max_collisions=128, kernel_fn_table=None, fused=False):
        """fused=True evaluates the layers that share (locs, neighbors) through ns.ConvSPGroup (one walk
        over the neighbour lists per dependency phase) instead of one call per layer; results are the
        same within fp32 rounding."""
        super(FluidStep, self).__init__()
        self.radius, self.ndim = radius, ndim
        self.fused = bool(fused)
        kf = kernel_fn_table if kernel_fn_table is not None else ns.KERNEL_FN
        self.density_rest, self.stiffness = rest_density(kf, radius)
        self.max_speed = 0.5 * 0.1 / DT
        self.coll = ns.ParticleCollision(ndim, radius, max_collisions=max_collisions, include_self=False)
        self.reorder_un2sort = ns.ReorderData(reverse=False)
        self.reorder_sort2un = ns.ReorderData(reverse=True)
        for kernel, dim, normed in LAYER_TYPES:
            c = ndim if dim == 'D' else 1
            conv = ns.ConvSP(c, c, ndim, kernel_size=1, dilation=1, radius=radius, dis_norm=normed,
                             with_params=False, kernel_fn=kernel)
            conv.bias.data.fill_(0)
            conv.weight.data.fill_(0)
            for i in range(c):
                conv.weight.data[i, i, 0] = 1
            setattr(self, "%s%s%s" % (kernel, "D" if dim == 'D' else "1", "normd" if normed else ""), conv)
        self.register_buffer("gravity", torch.tensor(GRAVITY[:ndim], dtype=torch.float32).view(1, 1, -1))
        self.relu = nn.ReLU()
        # fused=True also uses the namespace's fused elementwise solver stages when it has them (pbf.py)
        self.pbf = ns if (self.fused and hasattr(ns, "pbf_stage1")) else None
        if self.fused:
            self.group_a = ns.ConvSPGroup([self.spiky1, self.dspikyDnormd, self.dspiky1normd,
                                           self.cohesionDnormd, self.cohesion1normd, self.constant1])
            self.group_b = ns.ConvSPGroup([self.dspikyDnormd, self.dspiky1normd])
            self.group_c = ns.ConvSPGroup([self.constantD])
            self.group_v = ns.ConvSPGroup([self.spikyD, self.spiky1])

    def _cap_magnitude(self, A, cap):  # fluid_sim.py:240-245
        vv = torch.norm(A, 2, A.dim() - 1, keepdim=True)
        vv = cap / (vv + 0.0001)
        vv = -(self.relu(-vv + 1.0) - 1.0)
        return A * vv

    def forward(self, locs, vel):
        dt = DT
        ones = torch.ones(locs.shape[:-1] + (1,), device=locs.device, dtype=locs.dtype)
        if self.pbf is not None and hasattr(self.pbf, "pbf_integrate"):
            vel, new_locs = self.pbf.pbf_integrate(locs, vel, GRAVITY[:self.ndim], dt, self.max_speed)
        else:
            vel = vel + self.gravity * dt
            vel = self._cap_magnitude(vel, self.max_speed)
            new_locs = locs + vel * dt
        new_locs, vel, pidxs, neighbors = self.coll(new_locs, vel)
        for _ in range(NUM_ITERATIONS if self.pbf is not None else 0):
            # same data flow as below; the elementwise arithmetic between the groups in three fused stages
            pbf = self.pbf
            # new_locs has 8 consumers per iteration: give each its own alias so that the 8 gradients are
            # added in one kernel (pbf.fanout) instead of 7 pairwise accumulations
            fan = pbf.fanout(new_locs, 8) if hasattr(pbf, "fanout") else (new_locs,) * 8
            xa, xa1, xa2, x1, xb, x2, xc, x3 = fan
            density, nj, ni_s, nj_c, ni_cs, ncount = self.group_a(
                xa, [None, xa1, None, xa2, None, None], neighbors)  # None: data of ones, nothing is read for it
            pressure, xp, nij = pbf.pbf_stage1(x1, density, nj, ni_s, self.stiffness, self.density_rest)
            njp, nip_s = self.group_b(xb, [xp, pressure], neighbors)
            delta0, normals = pbf.pbf_stage2(x2, pressure, nij, njp, nip_s, nj_c, ni_cs, COHESION,
                                             self.radius, SURFACE_TENSION, self.density_rest,
                                             SURFACE_CONSTRAINT_SCALE)
            cd, = self.group_c(xc, [normals], neighbors)
            new_locs = pbf.pbf_stage3(x3, delta0, cd, normals, ncount, RELAXATION, DAMP)
        for _ in range(NUM_ITERATIONS if (self.fused and self.pbf is None) else 0):
            # same data flow as below; layers sharing (new_locs, neighbors) grouped by dependency
            density, nj, ni_s, nj_c, ni_cs, ncount = self.group_a(
                new_locs, [None, new_locs, None, new_locs, None, None], neighbors)
            nij = new_locs * ni_s - nj
            pressure = self.stiffness * self.relu(density - self.density_rest)
            njp, nip_s = self.group_b(new_locs, [new_locs * pressure, pressure], neighbors)
            nijp = new_locs * nip_s - njp
            delta = -(pressure * nij + nijp)
            nij = new_locs * ni_cs - nj_c
            delta = delta + -COHESION * nij * self.radius
            normals = nij * SURFACE_TENSION / self.density_rest / SURFACE_CONSTRAINT_SCALE
            cd, = self.group_c(new_locs, [normals], neighbors)
            delta = delta + (cd - normals * ncount)
            scale = ncount / (1.0 + RELAXATION)
            scale = self.relu(scale - DAMP) + DAMP
            delta = delta / scale
            new_locs = new_locs + delta
        for _ in range(0 if self.fused else NUM_ITERATIONS):
            density = self.spiky1(new_locs, ones, neighbors)
            nj = self.dspikyDnormd(new_locs, new_locs, neighbors)
            ni = new_locs * self.dspiky1normd(new_locs, ones, neighbors)
            nij = ni - nj
            pressure = self.stiffness * self.relu(density - self.density_rest)
            njp = self.dspikyDnormd(new_locs, new_locs * pressure, neighbors)
            nip = new_locs * self.dspiky1normd(new_locs, pressure, neighbors)
            nijp = nip - njp
            delta = -(pressure * nij + nijp)
            nj = self.cohesionDnormd(new_locs, new_locs, neighbors)
            ni = new_locs * self.cohesion1normd(new_locs, ones, neighbors)
            nij = ni - nj
            delta = delta + -COHESION * nij * self.radius
            normals = nij * SURFACE_TENSION / self.density_rest / SURFACE_CONSTRAINT_SCALE
            ncount = self.constant1(new_locs, ones, neighbors)
            delta = delta + (self.constantD(new_locs, normals, neighbors) - normals * ncount)
            scale = ncount / (1.0 + RELAXATION)
            scale = self.relu(scale - DAMP) + DAMP
            delta = delta / scale
            new_locs = new_locs + delta
        if self.pbf is not None and hasattr(self.pbf, "pbf_velocity"):
            vel = self.pbf.pbf_velocity(new_locs, self.reorder_un2sort(pidxs, locs), dt)
            vj, vi_s = self.group_v(new_locs, [vel, None], neighbors)
            vel = self.pbf.pbf_viscosity(vel, vj, vi_s, dt * VISCOSITY / self.density_rest)
        else:
            vel = (new_locs - self.reorder_un2sort(pidxs, locs)) / dt
            if self.fused:
                vj, vi_s = self.group_v(new_locs, [vel, None], neighbors)
                vi = vel * vi_s
            else:
                vj = self.spikyD(new_locs, vel, neighbors)
                vi = vel * self.spiky1(new_locs, ones, neighbors)
            vel = vel + dt * VISCOSITY / self.density_rest * (vj - vi)
        new_locs, vel = self.reorder_sort2un(pidxs, new_locs, vel)
        return new_locs, vel


def algorithmic_bytes_per_particle_step(nbar, ndim=3, K=128):
    """compulsory HBM bytes of one fluid step (forward + backward) per particle --
    every API-visible input read once and every output written once."""
    D = ndim
    bounds = 4 * D
    sort = 4 * D + 4 + 4
    reorder = lambda C: 4 + 8 * (D + C)
    collide = 4 * D + 4 + 4 * K
    fwd = lambda C, O: 4 * D + 4 * C + 4 * (nbar + 1) + 4 * O
    bwd = lambda C, O: 4 * D + 4 * C + 4 * (nbar + 1) + 4 * O + 8 * D + 4 * C
    total = bounds + sort + 2 * reorder(D) + collide
    total += CONVSP_1TO1_PER_STEP * (fwd(1, 1) + bwd(1, 1))
    total += CONVSP_DTOD_PER_STEP * (fwd(D, D) + bwd(D, D))
    total += 2 * reorder(0) + 2 * reorder(D)
    return total
