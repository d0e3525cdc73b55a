"""Seeded input generators shared by the oracle tests, the golden-vector generator and the GPU
parity tests.  Shapes follow the reference's own tests (tests/test_convsp.py:65-75,
tests/test_particlecollision.py:36-46, tests/test_convsdf.py:78-124) and BASELINE.json's configs."""
import numpy as np

KERNEL_NAMES = ["cohesion", "constant", "ddefault", "ddefault2", "default", "dpressure",
                "dpressure2", "dspiky", "indirect", "pressure", "sigmoid", "spiky"]


def rng(seed):
    return np.random.RandomState(seed)


def collision_case(seed=0, B=2, N=100, M=77, D=2, C=2, extent=1.0):
    """tests/test_particlecollision.py:37-48 (seed 0, B2 N100 M77 D2 R0.2 C2)."""
    r = rng(seed)
    locs = (r.rand(B, N, D) * extent).astype(np.float32)
    qlocs = (r.rand(B, M, D) * extent).astype(np.float32)
    data = r.rand(B, N, C).astype(np.float32)
    return locs, qlocs, data


def convsp_case(seed=0, B=2, N=5, M=3, D=2, C=2, O=3, ksize=(3, 1), extent=1.0):
    """tests/test_convsp.py:66-82 (seed 0; B2 N5 M3 D2 ks(3,1) C2 O3)."""
    r = rng(seed)
    locs = (r.rand(B, N, D) * extent).astype(np.float32)
    qlocs = (r.rand(B, M, D) * extent).astype(np.float32)
    data = r.rand(B, N, C).astype(np.float32)
    weight = r.rand(O, C, int(np.prod(ksize))).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    return locs, qlocs, data, weight, bias


def sphere_sdf(n, cell, centre, radius):
    """Signed distance to a sphere sampled at cell centres of an n^D grid."""
    D = len(centre)
    ax = [(np.arange(n) + 0.5) * cell for _ in range(D)]
    g = np.meshgrid(*ax, indexing="ij")
    d = np.sqrt(sum((g[k] - centre[k]) ** 2 for k in range(D))) - radius
    return d.astype(np.float32)


def box_sdf(n, cell, lo, hi):
    D = len(lo)
    ax = [(np.arange(n) + 0.5) * cell for _ in range(D)]
    g = np.meshgrid(*ax, indexing="ij")
    q = [np.maximum(lo[k] - g[k], g[k] - hi[k]) for k in range(D)]
    outside = np.sqrt(sum(np.maximum(qk, 0) ** 2 for qk in q))
    inside = np.minimum(np.max(np.stack(q), axis=0), 0)
    return (outside + inside).astype(np.float32)


def pack_sdfs(sdfs, cell_sizes):
    """convsdf.py:82-98: flat atlas, exclusive-cumsum offsets, shape rows [dims..., cell_size]."""
    shapes = np.array([list(s.shape) + [cs] for s, cs in zip(sdfs, cell_sizes)], np.float32)
    flat = [np.ascontiguousarray(s, np.float32).reshape(-1) for s in sdfs]
    offs = np.array([0] + np.cumsum([f.size for f in flat])[:-1].tolist(), np.float32)
    return np.concatenate(flat), offs, shapes


def random_quats(r, shape):
    q = r.randn(*shape, 4)
    q /= np.sqrt((q ** 2).sum(-1, keepdims=True))
    return q


def convsdf_case(seed=0, B=2, N=64, D=3, S=3, O=2, ksize=(3, 1, 3), nsdf=3, grid=12):
    """A posed multi-object scene in the spirit of tests/test_convsdf.py:78-124 (one idx = -1)."""
    r = rng(seed)
    sdfs, cells = [], []
    for i in range(nsdf):
        n = grid + 2 * i
        cell = 1.0 / n
        if D == 3 and i % 2 == 1:
            sdfs.append(box_sdf(n, cell, [0.3] * D, [0.7] * D))
        else:
            sdfs.append(sphere_sdf(n, cell, [0.5] * D, 0.25 + 0.05 * i))
        cells.append(cell)
    flat, offs, shapes = pack_sdfs(sdfs, cells)
    locs = (r.rand(B, N, D) * 1.6 - 0.3).astype(np.float32)
    idxs = r.randint(0, nsdf, size=(B, S)).astype(np.float32)
    idxs[-1, -1] = -1
    R = {1: 0, 2: 1, 3: 4}[D]
    poses = np.zeros((B, S, D + R), np.float32)
    poses[..., :D] = r.rand(B, S, D) * 0.6 - 0.3
    if D == 3:
        poses[..., D:] = random_quats(r, (B, S))
    elif D == 2:
        poses[..., D] = r.rand(B, S) * 2 * np.pi
    scales = (r.rand(B, S) + 0.5).astype(np.float32)
    weight = r.rand(O, int(np.prod(ksize))).astype(np.float32)
    bias = r.rand(O).astype(np.float32)
    return dict(locs=locs, idxs=idxs, poses=poses, scales=scales, sdfs=flat, offs=offs,
                shapes=shapes, weight=weight, bias=bias,
                ksize=np.array(ksize, np.float32))


def fluid_cloud(seed, B, N, density=7640.0, D=3):
    """SURVEY.md 8(d) c2 inputs: uniform cloud of side L with rho = N / L^D particles per unit
    volume (n-bar ~ 30 at radius 0.1 for rho = 7640), velocities ~ U[0,1)."""
    r = rng(seed)
    L = (N / density) ** (1.0 / D)
    locs = (r.rand(B, N, D) * L).astype(np.float32)
    vel = r.rand(B, N, D).astype(np.float32)
    return locs, vel, L


def projection_case(seed=0, B=2, N=200, W=64, H=48, C=3, fl=40.0):
    """Camera-space particles in front of, beside and behind a W x H camera, a depth mask that hides part of the
    image (in the spirit of tests/test_particleprojection.py:84-116) and an image to sample."""
    r = rng(seed)
    locs = (r.rand(B, N, 3).astype(np.float32) - 0.5) * np.array([4, 3, 0], np.float32)
    locs[..., 2] = r.rand(B, N).astype(np.float32) * 4 - 0.5  # some behind the camera
    dm = np.full((B, H, W), np.finfo(np.float32).max, np.float32)
    dm[0, H // 5:H * 3 // 5, W // 3:W * 3 // 4] = r.rand(H * 3 // 5 - H // 5, W * 3 // 4 - W // 3).astype(np.float32) * 3
    image = r.rand(B, C, H, W).astype(np.float32)
    return dict(locs=locs, depth_mask=dm, image=image, fl=np.float32(fl))
