// Compact "tile lists": an opt-in sidecar of the neighbour lists.
//
// The API-visible neighbour tensor stays [B,M,K] float32, -1 terminated (ParticleCollision.py:129-135).
// Next to it ParticleCollision can emit the SAME lists (same entries, same order, same truncation)
// in a form the ConvSP kernels can consume faster (spnb_build_tile_lists: k_tile_ranges + k_tile_build
// in hashgrid.cu, run on the float rows k_collide wrote):
//
//  * the cell-sorted particles of a scene are cut into TILE BLOCKS of kTileQ = 64 consecutive
//    queries.  Because the cell hash is row-major (common_funcs.h:114-118), every neighbour of such
//    a block lies in at most 3^(D-1) CONTIGUOUS ranges of the sorted order (one per offset of the
//    leading D-1 grid dimensions; the +-1 cells of the last dimension are adjacent in memory).  The
//    ranges are merged into disjoint ascending ones and stored in a TileDesc; a ConvSP kernel stages
//    them into shared memory with a handful of TMA bulk copies (cp.async.bulk) and then only gathers
//    from shared memory;
//  * a list entry is the 16-bit position of the neighbour inside that staged tile, stored times 16
//    (the byte offset of its record quarter in a shared-memory plane; 0 = sentinel, a record placed
//    far outside every radius), 2 bytes instead of 4;
//  * entries are stored in 32-byte UNITS of 16 entries, interleaved over the 8 queries of a row group
//    so that a warp reads whole 256-byte lines whether it spends 1, 2 or 4 lanes per query:
//        addr(b, tb, ql, k) = lists + (((b*ntb + tb)*8 + ql/8) * (K/16) + k/16) * 256 + (ql%8)*32 + (k%16)*2
//    Inside a unit the 16 entries are stored 4x4-transposed (entry e at position (e%4)*4 + e/4), so
//    that whatever the number of lanes per query, every lane's EVEN 32-bit words hold entries 0..7 of
//    the unit and its ODD words entries 8..15: a warp skips the odd words of a unit when none of its
//    queries has more than 8 entries left (most lists end in the first half of their last unit).
//    The tail of a query's last unit is filled with the sentinel; units past it are never read
//    (counts[b][n] holds the list length).
//
// A tile with more than kTileCap - 1 records is not staged: its block resolves slots back to sorted
// indices through the TileDesc and gathers from global memory (rare: dense clumps).
//
// Measured alternative (profiles/README.md): staging only the ~400 records a block's lists actually
// reference (an index list per tile, cp.async gathers, lists renumbered and sorted) halves the staged
// bytes but gave the same kernel times -- the consumers are bound by the shared-memory pipe (random
// LDS.128 gathers conflict ~2x), not by staging -- while its builder cost 4x more; not kept.
//
// flag (first int of the buffer) != 0 marks the sidecar unusable for this call: bit 0 = a list is full,
// i.e. may have been cut at K (the symmetric backward is then invalid, see convsp_group.cu), bit 1 = a
// neighbour lies outside the block's ranges or a tile has more than 4095 records.  Consumers test it on
// the DEVICE and run the ordinary list walk instead (inside the same kernel), so nothing depends on a
// host synchronisation.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace spnb {

#ifndef SPNB_TILE_Q
#define SPNB_TILE_Q 64
#endif
constexpr int kTileQ = SPNB_TILE_Q;  // queries per tile block (a multiple of 64)
#ifndef SPNB_TILE_CAP
#define SPNB_TILE_CAP 1024
#endif
constexpr int kTileCap = SPNB_TILE_CAP;  // staged records per tile, including the sentinel at slot 0
static_assert(kTileCap <= 4096, "entries are slot * 16 in 16 bits");
constexpr int kTileMaxRanges = 9;   // 3^(D-1) for D <= 3
constexpr int kTileUnit = 16;       // entries per 32-byte unit
constexpr int kTileMaxNdim = 3;

struct TileDesc {                   // 128 bytes
    int nr;                         // number of disjoint ranges
    int total;                      // staged records (without the sentinel)
    int start[kTileMaxRanges];      // first sorted particle of range r
    int prefix[kTileMaxRanges + 1]; // records staged before range r (prefix[nr] == total)
    int pad[11];
};
static_assert(sizeof(TileDesc) == 128, "TileDesc layout");

struct TileLayout {
    size_t desc_off, cnt_off, list_off, total;
    int ntb;                        // tile blocks per scene
};

__host__ __device__ inline TileLayout tile_layout(int B, int N, int K)
{
    TileLayout t;
    t.ntb = (N + kTileQ - 1) / kTileQ;
    size_t off = 128;               // header: int flag
    t.desc_off = off;
    off += sizeof(TileDesc) * (size_t)B * t.ntb;
    t.cnt_off = off;
    off += ((sizeof(int) * (size_t)B * N + 255) / 256) * 256;
    t.list_off = off;
    off += (size_t)B * t.ntb * kTileQ * K * 2;
    t.total = off;
    return t;
}

__host__ __device__ inline bool tile_lists_supported(int N, int D, int K)
{
    return D >= 1 && D <= kTileMaxNdim && K >= kTileUnit && (K % kTileUnit) == 0 && N >= 1;
}

// byte offset of entry k of query ql (0..kTileQ-1) of tile block (b, tb), relative to the list area
__host__ __device__ inline size_t tile_entry_off(int ntb, int K, int b, int tb, int ql, int k)
{
    const int e = k % kTileUnit;
    return ((((size_t)b * ntb + tb) * (kTileQ / 8) + (ql >> 3)) * (size_t)(K / kTileUnit) + (k / kTileUnit)) * 256 +
           (size_t)(ql & 7) * 32 + (size_t)((e & 3) * 4 + (e >> 2)) * 2;
}

}  // namespace spnb
