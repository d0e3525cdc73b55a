"""Helpers for the GPU parity tests: raw C-ABI calls on torch CUDA tensors (the tests go through
the same ctypes binding the product uses) and comparison utilities."""
import numpy as np
import torch

from smoothparticlenets_b200 import _native as nat


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


def host(t):
    return t.detach().cpu().numpy()


def bits(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        raise AssertionError("%s: %d of %d elements differ bitwise; first at %s: %r vs %r" % (
            what, len(bad), a.size, bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=""):
    """|a-b| <= atol + rtol*|b| (north_star tolerance: 1e-5 relative / 1e-6 absolute)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), a.shape)
        raise AssertionError("%s: max violation at %s: got %r want %r (err %.3g, tol %.3g)" % (
            what, i, a[i], b[i], err[i], tol[i]))


def grid_bounds(locs_t, radius, G):
    L = nat.lib()
    B, N, D = locs_t.shape
    wsb = L.spnb_hashgrid_workspace_bytes(B, N, D, G)
    ws = torch.empty((wsb + 3) // 4, device="cuda", dtype=torch.float32)
    low = torch.empty(B, D, device="cuda")
    gd = torch.empty(B, D, device="cuda")
    nat.check(L.spnb_grid_bounds(nat.ptr(locs_t), B, N, D, float(radius), G, nat.ptr(low), nat.ptr(gd),
                                 nat.ptr(ws), wsb, nat.stream()), "grid_bounds")
    return low, gd


def hashgrid_order(locs_t, low, gd, edge, G):
    L = nat.lib()
    B, N, D = locs_t.shape
    wsb = L.spnb_hashgrid_workspace_bytes(B, N, D, G)
    ws = torch.empty((wsb + 3) // 4, device="cuda", dtype=torch.float32)
    ids = torch.zeros(B, N, device="cuda")
    idxs = torch.zeros(B, N, device="cuda")
    nat.check(L.spnb_hashgrid_order(nat.ptr(locs_t), nat.ptr(low), nat.ptr(gd), nat.ptr(ids),
                                    nat.ptr(idxs), nat.ptr(ws), wsb, B, N, D, float(edge), G,
                                    nat.stream()), "hashgrid_order")
    return ids, idxs


def reorder(locs_t, data_t, idxs_t, reverse=0):
    L = nat.lib()
    B, N, D = locs_t.shape
    nl = torch.empty_like(locs_t)
    nd = torch.empty_like(data_t) if data_t is not None else None
    C = data_t.shape[2] if data_t is not None else 0
    nat.check(L.spnb_reorder_data(nat.ptr(locs_t), nat.ptr(data_t), nat.ptr(idxs_t), nat.ptr(nl),
                                  nat.ptr(nd), B, N, D, C, reverse, nat.stream()), "reorder")
    return nl, nd


def collisions(q_t, locs_t, low, gd, ids, edge, radius, K, include_self, G):
    L = nat.lib()
    B, N, D = locs_t.shape
    M = q_t.shape[1]
    ncells = G ** D
    st = torch.empty(B, ncells, device="cuda")
    en = torch.empty(B, ncells, device="cuda")
    coll = torch.empty(B, M, K, device="cuda")
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    nat.check(L.spnb_compute_collisions(nat.ptr(q_t), nat.ptr(locs_t), nat.ptr(low), nat.ptr(gd),
                                        nat.ptr(ids), nat.ptr(st), nat.ptr(en), nat.ptr(coll), B, M, N,
                                        D, K, ncells, float(edge), float(radius), int(include_self),
                                        nat.ptr(flag), nat.stream()), "collisions")
    return coll, flag


def convsp_forward(q, locs, data, nb, w, bias, radius, ksize, dil, dis_norm, fn):
    L = nat.lib()
    B, N, D = locs.shape
    M, C, K, O, nc = q.shape[1], data.shape[2], nb.shape[2], w.shape[0], w.shape[2]
    out = torch.empty(B, M, O, device="cuda")
    nat.check(L.spnb_convsp_forward(nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(nb), nat.ptr(w),
                                    nat.ptr(bias), B, M, N, C, D, K, O, nc, float(radius),
                                    nat.ptr(ksize), nat.ptr(dil), int(dis_norm), int(fn),
                                    nat.ptr(out), nat.stream()), "convsp_forward")
    return out


def convsp_backward(q, locs, data, nb, w, radius, ksize, dil, dis_norm, fn, go, sym_flag=None,
                    same=False, want=(True, True, True, True)):
    """Returns (dq, dl, dd, dw).  same=True passes one buffer for dq and dl (qlocs is locs)."""
    L = nat.lib()
    B, N, D = locs.shape
    M, C, K, O, nc = q.shape[1], data.shape[2], nb.shape[2], w.shape[0], w.shape[2]
    dq = torch.full((B, M, D), float("nan"), device="cuda") if want[0] else None
    dl = (dq if same else torch.full((B, N, D), float("nan"), device="cuda")) if want[1] else None
    dd = torch.full((B, N, C), float("nan"), device="cuda") if want[2] else None
    dw = torch.full_like(w, float("nan")) if want[3] else None
    nat.check(L.spnb_convsp_backward(nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(nb), nat.ptr(w),
                                     B, M, N, C, D, K, O, nc, float(radius), nat.ptr(ksize),
                                     nat.ptr(dil), int(dis_norm), int(fn), nat.ptr(go), nat.ptr(dq),
                                     nat.ptr(dl), nat.ptr(dd), nat.ptr(dw), nat.ptr(sym_flag), None,
                                     nat.stream()), "convsp_backward")
    return dq, dl, dd, dw


def convsp_forward_wide(q, locs, data, nb, w, bias, radius, ksize, dil, dis_norm, fn):
    L = nat.lib()
    B, N, D = locs.shape
    M, C, K, O, nc = q.shape[1], data.shape[2], nb.shape[2], w.shape[0], w.shape[2]
    wsb = L.spnb_convsp_forward_wide_workspace_bytes(O, C, D, nc)
    assert wsb > 0, "shape not supported by the wide path"
    ws = torch.empty(wsb // 4 + 1, device="cuda")
    out = torch.empty(B, M, O, device="cuda")
    nat.check(L.spnb_convsp_forward_wide(nat.ptr(q), nat.ptr(locs), nat.ptr(data), nat.ptr(nb), nat.ptr(w),
                                         nat.ptr(bias), B, M, N, C, D, K, O, nc, float(radius),
                                         nat.ptr(ksize), nat.ptr(dil), int(dis_norm), int(fn),
                                         nat.ptr(out), nat.ptr(ws), wsb, nat.stream()), "convsp_forward_wide")
    return out
