"""Host-side mirror of the compact tile-list layout (csrc/tile_lists.cuh).

ParticleCollision keeps the sidecar of the neighbour tensor it returns in the sidecar registry
(``tile_lists_of(neighbors)``, a uint8 CUDA tensor).  ``decode`` turns it back into the API's [B,N,K] index
lists so tests (and users debugging a pipeline) can check that both forms describe the same lists; nothing on
the product path calls it.
"""
import numpy as np

TILE_Q = 64
TILE_CAP = 1016
DESC_INTS = 32
MAX_RANGES = 9
OCTILES = 8
HEADER = 64
ROW = 64
MAX_K = 512
MAX_SLOTS = 4096


def layout(B, N, K):
    ntb = (N + TILE_Q - 1) // TILE_Q
    desc_off = 128
    list_off = desc_off + 4 * DESC_INTS * B * ntb
    stride = (HEADER + OCTILES * ((K + 3) // 4) * ROW + 255) // 256 * 256
    total = list_off + B * ntb * stride
    return dict(ntb=ntb, desc_off=desc_off, list_off=list_off, stride=stride, total=total)


def descriptors(raw, B, N, K):
    """(desc int32 [B,ntb,32], goff uint16 [B,ntb,9], maxcnt [B,ntb], sumcnt [B,ntb])."""
    lay = layout(B, N, K)
    d8 = raw[lay["desc_off"]:lay["list_off"]].reshape(B, lay["ntb"], 4 * DESC_INTS)
    desc = d8.view(np.int32).reshape(B, lay["ntb"], DESC_INTS)
    goff = np.ascontiguousarray(d8[:, :, 84:102]).view(np.uint16).reshape(B, lay["ntb"], OCTILES + 1)
    maxcnt = np.ascontiguousarray(d8[:, :, 102:104]).view(np.uint16).reshape(B, lay["ntb"])
    return desc, goff, maxcnt, desc[:, :, 26]


def decode(tiles, B, N, K, blocks=None):
    """tiles: uint8 tensor/array.  Returns (flag, counts [B,N] int, neighbors [B,N,K] int64 with -1 padding, max
    staged records of any tile).  Entries past a list's end inside its octile's rows must be the sentinel.
    Queries of blocks without tile rows (more than MAX_SLOTS candidates) get count -1.
    `blocks`: optional iterable of (b, tb) to decode only those tile blocks (the rest stays -1 / 0)."""
    raw = tiles.detach().cpu().numpy() if hasattr(tiles, "detach") else np.asarray(tiles)
    lay = layout(B, N, K)
    assert raw.size == lay["total"], (raw.size, lay["total"])
    ntb = lay["ntb"]
    flag = int(raw[:4].view(np.int32)[0])
    desc, goff, maxcnt, sumcnt = descriptors(raw, B, N, K)
    out = -np.ones((B, N, K), dtype=np.int64)
    counts = np.zeros((B, N), dtype=np.int64)
    todo = blocks if blocks is not None else [(b, tb) for b in range(B) for tb in range(ntb)]
    for b, tb in todo:
        d = desc[b, tb]
        nr, total = int(d[0]), int(d[1])
        if total + 1 > MAX_SLOTS:
            # more candidates than 16-bit entries address: no rows are written for such a block (flag bit 1)
            counts[b, tb * TILE_Q:(tb + 1) * TILE_Q] = -1
            continue
        start, prefix = d[2:2 + MAX_RANGES], d[2 + MAX_RANGES:2 + 2 * MAX_RANGES + 1]
        slot2idx = np.full(max(total, 0) + 1, -1, dtype=np.int64)  # slot (1-based) -> sorted particle index
        for r in range(nr):
            ln = int(prefix[r + 1] - prefix[r])
            slot2idx[1 + prefix[r]:1 + prefix[r] + ln] = start[r] + np.arange(ln)
        o = lay["list_off"] + (b * ntb + tb) * lay["stride"]
        perm = raw[o:o + HEADER]
        nq = min(TILE_Q, N - tb * TILE_Q)
        assert sorted(perm.tolist()) == list(range(TILE_Q)), "perm is a permutation of the block's queries"
        g = goff[b, tb].astype(np.int64)
        rows = raw[o + HEADER:o + HEADER + int(g[OCTILES]) * ROW].view(np.uint16).reshape(-1, 8, 4)
        lens = []
        for rank in range(TILE_Q):
            oc, r = rank // 8, rank % 8
            ent = rows[g[oc]:g[oc + 1], r, :].reshape(-1)      # entry 4s+e of this rank
            assert ((ent & 15) == 0).all(), "entries are slot * 16"
            ent = ent >> 4
            c = int((ent != 0).sum())
            assert (ent[:c] != 0).all() and (ent[c:] == 0).all(), "padding after the list's end is the sentinel"
            lens.append(c)
            ql = int(perm[rank])
            if ql < nq:
                n = tb * TILE_Q + ql
                counts[b, n] = c
                out[b, n, :c] = slot2idx[ent[:c]]
            else:
                assert c == 0
        assert lens == sorted(lens, reverse=True), "queries are ranked by list length, longest first"
        assert int(maxcnt[b, tb]) == lens[0] and int(sumcnt[b, tb]) == sum(lens)
        for oc in range(OCTILES):
            assert g[oc + 1] - g[oc] == (lens[8 * oc] + 3) // 4, "rows of an octile cover its longest list"
    return flag, counts, out, int(desc[:, :, 1].max())
