"""float64 evaluation of the ConvSDF forward formula, for checking derivatives.

Test infrastructure only.  A plain torch restatement of what ConvSDF computes (src/common_funcs.h:
203-235 rotate_point, 317-384 n-linear interpolation, 647-837 kernel cells and min over objects),
in double precision and without the reference's conservative pre-cull (which does not change the
result for distance fields).  Used for central finite differences and autograd derivatives that
float32 finite differences cannot resolve.
"""
import itertools

import numpy as np
import torch


def _inverse_rotate(v, rot, D):
    if D == 1:
        return v
    if D == 2:
        m = torch.sqrt((v ** 2).sum(-1))
        th = torch.atan2(v[..., 1], v[..., 0]) - rot[0]
        return torch.stack([m * torch.cos(th), m * torch.sin(th)], -1)
    x, y, z, w = rot[0], rot[1], rot[2], rot[3]
    px, py, pz = v[..., 0], v[..., 1], v[..., 2]
    aw = -px * x - py * y - pz * z
    ax = px * w + py * z - pz * y
    ay = py * w + pz * x - px * z
    az = pz * w + px * y - py * x
    return torch.stack([w * ax - x * aw - y * az + z * ay,
                        w * ay - y * aw - z * ax + x * az,
                        w * az - z * aw - x * ay + y * ax], -1)


def convsdf_float64(locs, idxs, poses, scales, sdfs, offs, shapes, weight, bias, ksize, dil,
                    max_distance):
    """All arguments numpy / torch; poses may be a float64 torch tensor that requires grad.
    Returns a float64 torch tensor B x N x O."""
    t64 = lambda a: a.double() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, np.float64))
    locs, poses, scales, sdfs, weight, bias = map(t64, (locs, poses, scales, sdfs, weight, bias))
    idxs = np.asarray(idxs)
    shapes = np.asarray(shapes, np.float64)
    offs = np.asarray(offs)
    B, N, D = locs.shape
    S = idxs.shape[1]
    ks = [int(k) for k in np.asarray(ksize)]
    dil = [float(d) for d in np.asarray(dil)]
    cells = []  # dimension 0 varies fastest (common_funcs.h:760-775)
    for rev in itertools.product(*[range(k) for k in ks[::-1]]):
        kidx = rev[::-1]
        cells.append([(kidx[i] - ks[i] // 2) * dil[i] for i in range(D)])
    cells = torch.tensor(cells, dtype=torch.float64)  # ncells x D
    out = []
    for b in range(B):
        pts = locs[b][:, None, :] + cells[None]  # N x ncells x D
        best = torch.full(pts.shape[:2], float(max_distance), dtype=torch.float64)
        for m in range(S):
            mm = int(idxs[b, m])
            if mm < 0:
                continue
            shp = shapes[mm, :D].astype(int)
            cell = float(shapes[mm, D]) * scales[b, m]
            grid = sdfs[int(offs[mm]):int(offs[mm]) + int(np.prod(shp))].reshape(*shp)
            p = _inverse_rotate(pts - poses[b, m, :D], poses[b, m, D:], D)
            inside = torch.ones(pts.shape[:2], dtype=torch.bool)
            for i in range(D):
                inside &= (p[..., i] >= 0.5 * cell) & (p[..., i] <= (shp[i] - 0.5) * cell)
            u = p / cell - 0.5
            low = torch.floor(u.detach()).long()
            frac = u - low
            val = torch.zeros(pts.shape[:2], dtype=torch.float64)
            for corner in itertools.product((0, 1), repeat=D):
                wgt = torch.ones_like(val)
                index = []
                for i in range(D):
                    wgt = wgt * (frac[..., i] if corner[i] else 1 - frac[..., i])
                    index.append((low[..., i] + corner[i]).clamp(0, shp[i] - 1))
                val = val + wgt * grid[tuple(index)]
            val = val * scales[b, m]
            best = torch.where(inside & (val < best), val, best)
        out.append(best @ weight.t() + bias)
    return torch.stack(out)
