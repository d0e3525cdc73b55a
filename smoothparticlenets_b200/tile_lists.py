"""Host-side mirror of the compact tile-list layout (csrc/tile_lists.cuh).

ParticleCollision attaches the sidecar to the neighbour tensor it returns (``neighbors._spnb_tiles``, a
uint8 CUDA tensor).  ``decode`` turns it back into the API's [B,N,K] index lists so tests (and users
debugging a pipeline) can check that both forms describe the same lists; nothing on the product path
calls it.
"""
import numpy as np

TILE_Q = 64
TILE_CAP = 1024
TILE_UNIT = 16
DESC_INTS = 32
MAX_RANGES = 9


def layout(B, N, K):
    ntb = (N + TILE_Q - 1) // TILE_Q
    desc_off = 128
    cnt_off = desc_off + 4 * DESC_INTS * B * ntb
    list_off = cnt_off + (4 * B * N + 255) // 256 * 256
    total = list_off + B * ntb * TILE_Q * K * 2
    return dict(ntb=ntb, desc_off=desc_off, cnt_off=cnt_off, list_off=list_off, total=total)


def decode(tiles, B, N, K):
    """tiles: uint8 tensor/array.  Returns (flag, counts [B,N] int, neighbors [B,N,K] int64 with -1
    padding, max staged records of any tile)."""
    raw = tiles.detach().cpu().numpy() if hasattr(tiles, "detach") else np.asarray(tiles)
    lay = layout(B, N, K)
    assert raw.size == lay["total"], (raw.size, lay["total"])
    ntb = lay["ntb"]
    flag = int(raw[:4].view(np.int32)[0])
    desc = raw[lay["desc_off"]:lay["cnt_off"]].view(np.int32).reshape(B, ntb, DESC_INTS)
    counts = raw[lay["cnt_off"]:lay["cnt_off"] + 4 * B * N].view(np.int32).reshape(B, N)
    # lists[b, tb, rowgroup, unit, slot-in-rowgroup, entry]
    lists = raw[lay["list_off"]:].view(np.uint16).reshape(B, ntb, TILE_Q // 8, K // TILE_UNIT, 8, TILE_UNIT)
    out = -np.ones((B, N, K), dtype=np.int64)
    for b in range(B):
        for tb in range(ntb):
            d = desc[b, tb]
            nr, total = int(d[0]), int(d[1])
            start, prefix = d[2:2 + MAX_RANGES], d[2 + MAX_RANGES:2 + 2 * MAX_RANGES + 1]
            # slot (1-based) -> sorted particle index
            slot2idx = np.full(max(total, 0) + 1, -1, dtype=np.int64)
            for r in range(nr):
                ln = int(prefix[r + 1] - prefix[r])
                slot2idx[1 + prefix[r]:1 + prefix[r] + ln] = start[r] + np.arange(ln)
            for ql in range(min(TILE_Q, N - tb * TILE_Q)):
                n = tb * TILE_Q + ql
                c = int(counts[b, n])
                ent = lists[b, tb, ql // 8, :, ql % 8, :]
                ent = ent.reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1)  # units are stored 4x4-transposed
                cpad = (c + TILE_UNIT - 1) // TILE_UNIT * TILE_UNIT
                assert ((ent[:cpad] & 15) == 0).all(), "entries are slot * 16"
                ent = ent >> 4
                assert (ent[c:cpad] == 0).all(), "tail of the last unit must hold the sentinel"
                out[b, n, :c] = slot2idx[ent[:c]]
    return flag, counts, out, int(desc[:, :, 1].max())
