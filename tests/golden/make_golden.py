"""Generates tests/golden/*.npz from the UNMODIFIED reference CPU extension (oracle/_ref).

Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Each fixture stores seeded inputs (tests/cases.py) together with the outputs of the reference's own
spn_* functions called with the allocation conventions of its autograd Functions (see
oracle/spn_oracle.py:RefOracle).  The fixtures pin oracle/spn_oracle.c where the reference itself
cannot travel (the GPU box has no /root/reference), and kernels.py's formulas via the reference's
KERNEL_FN lambdas (python/SmoothParticleNets/kernels.py:126-131).
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle.spn_oracle import RefOracle, grid_bounds_torch  # noqa: E402


def hashgrid_fixture(R, seed, **kw):
    radius = kw.pop("radius")
    G = kw.pop("G", 96)
    locs, qlocs, data = cases.collision_case(seed, **kw)
    low, gd = grid_bounds_torch(locs, radius, G)
    ids, idxs = R.hashgrid_order(locs, low, gd, radius)  # the CPU reference's (unstable) order
    keys = None
    # per-particle keys, recovered by undoing the permutation
    keys = np.zeros_like(ids)
    for b in range(ids.shape[0]):
        keys[b, idxs[b].astype(int)] = ids[b]
    # the contract order: stable by key
    sidx = np.stack([np.argsort(keys[b], kind="stable") for b in range(keys.shape[0])]).astype(np.float32)
    sids = np.take_along_axis(keys, sidx.astype(int), 1)
    nl, nd = R.reorder_data(locs, data, sidx, 0)
    out = dict(locs=locs, qlocs=qlocs, data=data, radius=np.float32(radius), G=np.int32(G), low=low,
               grid_dims=gd, ref_ids=ids, ref_idxs=idxs, keys=keys, stable_idxs=sidx, stable_ids=sids,
               sorted_locs=nl, sorted_data=nd)
    D = locs.shape[2]
    for inc in (0, 1):
        for K in (4, 64):
            for tag, q in (("q", qlocs), ("self", nl)):
                co, _, _ = R.compute_collisions(q, nl, low, gd, sids, radius, radius, K, inc, G ** D)
                out["coll_%s_K%d_s%d" % (tag, K, inc)] = co
    return out


def convsp_fixture(R, seed, D, ks, dil, radius, C, O, N, M, extent=1.0):
    locs, qlocs, data, weight, bias = cases.convsp_case(seed, B=2, N=N, M=M, D=D, C=C, O=O, ksize=ks,
                                                        extent=extent)
    cr = radius + dil * max((k - 1) / 2 for k in ks)
    low, gd = grid_bounds_torch(locs, cr, 96)
    ids, idxs = R.hashgrid_order(locs, low, gd, cr)
    nl, nd = R.reorder_data(locs, data, idxs, 0)
    ksz = np.array(ks, np.float32)
    dl = np.full(D, dil, np.float32)
    out = dict(locs=nl, data=nd, qlocs=qlocs, weight=weight, bias=bias, radius=np.float32(radius),
               ksize=ksz, dil=dl)
    go = cases.rng(seed + 100)
    for tag, q in (("q", qlocs), ("self", nl)):
        nb, _, _ = R.compute_collisions(q, nl, low, gd, ids, cr, cr, 32, 1, 96 ** D)
        out["nb_" + tag] = nb
        g = go.rand(2, q.shape[1], O).astype(np.float32)
        out["go_" + tag] = g
        for fn in cases.KERNEL_NAMES:
            for dn in (0, 1):
                key = "%s_%s_n%d" % (tag, fn, dn)
                out["fwd_" + key] = R.convsp_forward(q, nl, nd, nb, weight, bias, radius, ksz, dl, dn, fn)
                dq, dloc, dd, dw, db = R.convsp_backward(q, nl, nd, nb, weight, bias, radius, ksz, dl, dn,
                                                         fn, g)
                out["dq_" + key], out["dl_" + key], out["dd_" + key], out["dw_" + key] = dq, dloc, dd, dw
    return out


def convsdf_fixture(R, seed, D, ks):
    c = cases.convsdf_case(seed, B=2, N=40, D=D, S=3, O=2, ksize=ks)
    dil = np.full(D, 0.01, np.float32)
    out = dict(c)
    out["dil"] = dil
    for md in (0.5, 0.05):
        a = (c["locs"], c["idxs"], c["poses"], c["scales"], c["sdfs"], c["offs"], c["shapes"], c["weight"],
             c["bias"], c["ksize"], dil, md)
        fwd = R.convsdf_forward(*a)
        go = cases.rng(seed + 7).rand(*fwd.shape).astype(np.float32)
        dl, dw, dp, _ = R.convsdf_backward(*a, go, pose_grads=True)
        t = "md%g" % md
        out["fwd_" + t], out["go_" + t], out["dl_" + t], out["dw_" + t] = fwd, go, dl, dw
        out["dpt_" + t] = dp[..., :D]
    return out


def projection_fixture(R, seed, std, scale):
    c = cases.projection_case(seed)
    fl = float(c["fl"])
    out = dict(c)
    out["std"], out["scale"] = np.float32(std), np.float32(scale)
    out["pp_fwd"] = R.particleprojection_forward(c["locs"], fl, std, scale, c["depth_mask"])
    out["pp_go"] = cases.rng(seed + 3).rand(*out["pp_fwd"].shape).astype(np.float32)
    out["pp_dl"] = R.particleprojection_backward(c["locs"], fl, std, scale, c["depth_mask"], out["pp_go"])
    out["ip_fwd"] = R.imageprojection_forward(c["locs"], c["image"], fl, c["depth_mask"])
    out["ip_go"] = cases.rng(seed + 4).rand(*out["ip_fwd"].shape).astype(np.float32)
    out["ip_dl"], out["ip_di"] = R.imageprojection_backward(c["locs"], c["image"], fl, c["depth_mask"], out["ip_go"])
    return out


def kernel_fixture():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location(
            "_ref_kernels", os.path.join(build_ref.REF_ROOT, "python", "SmoothParticleNets", "kernels.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    names = list(mod.KERNEL_NAMES)
    H = np.array([0.1, 1.0, 0.37])
    d = np.linspace(0.0, 1.0, 21)
    vals = np.zeros((len(names), len(H), len(d)))
    for i, n in enumerate(names):
        for j, h in enumerate(H):
            for k, x in enumerate(d):
                vals[i, j, k] = mod.KERNEL_FN[n](x * h, h)
    return dict(names=np.array(names), H=H, dfrac=d, values=vals)


def main():
    assert build_ref.build() is not None, "needs /root/reference"
    R = RefOracle()
    fx = {
        "hashgrid_2d": hashgrid_fixture(R, 0, B=2, N=100, M=77, D=2, C=2, radius=0.2),
        "hashgrid_3d": hashgrid_fixture(R, 1, B=2, N=300, M=40, D=3, C=3, radius=0.15),
        "hashgrid_1d_clamped": hashgrid_fixture(R, 2, B=1, N=200, M=20, D=1, C=1, extent=4.0, radius=0.01, G=16),
        "convsp_ref_test_shape": convsp_fixture(R, 0, 2, (3, 1), 0.05, 1.0, 2, 3, 5, 3),
        "convsp_3d_k3": convsp_fixture(R, 1, 3, (3, 3, 3), 0.05, 0.3, 3, 4, 60, 17),
        "convsp_3d_k1": convsp_fixture(R, 2, 3, (1, 1, 1), 1.0, 0.3, 3, 3, 60, 17),
        "convsdf_3d": convsdf_fixture(R, 0, 3, (3, 1, 3)),
        "convsdf_2d": convsdf_fixture(R, 1, 2, (3, 3)),
        "convsdf_1d": convsdf_fixture(R, 2, 1, (3,)),
        "projection_a": projection_fixture(R, 0, 2.5, 3.0),
        "projection_b": projection_fixture(R, 1, 0.8, 1.0 / 0.06),
        "kernel_fn": kernel_fixture(),
    }
    for name, d in fx.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-28s %7.1f KiB" % (name, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
