// Compact "tile lists": an opt-in sidecar of the neighbour lists.
//
// The API-visible neighbour tensor stays [B,M,K] float32, -1 terminated (ParticleCollision.py:129-135).
// Next to it ParticleCollision emits the SAME lists (same entries, same order, same truncation) in a
// form the ConvSP kernels consume faster.  Both are written by one kernel, k_collide_tiles (hashgrid.cu),
// from the same hits.
//
//  * TILES.  The cell-sorted particles of a scene are cut into tile blocks of kTileQ = 64 consecutive
//    queries.  Because the cell hash is row-major (common_funcs.h:114-118), every neighbour of such a block
//    lies in at most 3^(D-1) CONTIGUOUS ranges of the sorted order (one per offset of the leading D-1 grid
//    dimensions; the +-1 cells of the last dimension are adjacent in memory).  The ranges are merged into
//    disjoint ascending ones and stored in a TileDesc; a kernel stages them into shared memory with a
//    handful of TMA bulk copies (cp.async.bulk) and then only gathers from shared memory.  Slot s >= 1 of a
//    tile is the s-th particle of its ranges; slot 0 is a sentinel record placed far outside every radius.
//
//  * RANKS.  The 64 queries of a block are ranked by list length, longest first (perm[rank] = query within
//    the block), and consumed in rank order: the 8 queries of an OCTILE (ranks 8g .. 8g+7) have nearly
//    equal lengths, so the lanes of a warp that walk them in lock step waste few slots.
//
//  * ENTRIES.  A list entry is the 16-bit slot of the neighbour times 16 (the byte offset of its record in
//    a shared-memory plane of 16-byte record quarters).  Entries are stored in ROWS of 32 (64 bytes):
//    row (goff[g] + s) holds, for octile g and step s, entry 4s+e of rank 8g+r at position 4r+e.  Whether a
//    consumer spends 4, 2 or 1 lanes per query, a step of a warp reads whole rows (one u16, one u32 or one
//    u64 per lane).  An octile has S_g = ceil(longest list of the octile / 4) rows; shorter lists are padded
//    with the sentinel.  The rows of a block follow a 64-byte header (perm) in a blob of fixed stride.
//
// Consumers (convsp_group.cu) stream the rows of their octiles through shared memory with cp.async, so the
// inner loop touches shared memory only.
//
// flag (first int of the buffer) != 0 marks the sidecar unusable for this call: bit 0 = the neighbour
// relation may not be symmetric (a list was cut at K, or a query lies beyond a clamped grid: the tile
// kernels' backward only has the symmetric gather mode), bit 1 = a tile has more than kTileMaxSlots
// records, bit 2 = internal inconsistency (a cell outside the block's ranges; never expected).  Consumers
// test it on the DEVICE and run the ordinary float-list walk instead (inside the same kernel), so nothing
// depends on a host synchronisation.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace spnb {

#ifndef SPNB_TILE_Q
#define SPNB_TILE_Q 64
#endif
constexpr int kTileQ = SPNB_TILE_Q;  // queries per tile block
static_assert(kTileQ == 64, "the rank / octile layout assumes 64 queries per block");
// 1016, not 1024: a backward kernel that stages three float4 planes (48 KB) plus its 8 KB row stage then fits FOUR times
// into an SM's 228 KB of shared memory (1 KB per CTA is reserved): 3 -> 4 CTAs per SM, -12 % on those kernels (measured)
#ifndef SPNB_TILE_CAP
#define SPNB_TILE_CAP 1016
#endif
constexpr int kTileCap = SPNB_TILE_CAP;  // staged records per tile, including the sentinel at slot 0
constexpr int kTileMaxSlots = 4096;      // entries are slot * 16 in 16 bits
static_assert(kTileCap <= kTileMaxSlots, "entries are slot * 16 in 16 bits");
constexpr int kTileMaxRanges = 9;   // 3^(D-1) for D <= 3
constexpr int kTileMaxNdim = 3;
constexpr int kTileOctiles = kTileQ / 8;
constexpr int kTileMaxK = 512;      // longest list the format is built for (rows per octile fit 8 bits)
constexpr int kTileRowBytes = 64;   // 32 entries
constexpr int kTileHeaderBytes = 64;  // perm[64]

struct TileDesc {                   // 128 bytes
    int nr;                         // number of disjoint ranges
    int total;                      // staged records (without the sentinel)
    int start[kTileMaxRanges];      // first sorted particle of range r
    int prefix[kTileMaxRanges + 1]; // records staged before range r (prefix[nr] == total)
    unsigned short goff[kTileOctiles + 1];  // first row of octile g; goff[8] = rows of the block
    unsigned short maxcnt;          // longest list of the block
    int sumcnt;                     // entries of the block (without padding)
    int pad[5];
};
static_assert(sizeof(TileDesc) == 128, "TileDesc layout");

struct TileLayout {
    size_t desc_off, list_off, blob_stride, total;
    int ntb;                        // tile blocks per scene
};

__host__ __device__ inline TileLayout tile_layout(int B, int N, int K)
{
    TileLayout t;
    t.ntb = (N + kTileQ - 1) / kTileQ;
    size_t off = 128;               // header: int flag
    t.desc_off = off;
    off += sizeof(TileDesc) * (size_t)B * t.ntb;
    t.list_off = off;
    // header + 8 octiles of at most ceil(K/4) rows, rounded to 256 bytes
    t.blob_stride = ((size_t)kTileHeaderBytes + (size_t)kTileOctiles * ((K + 3) / 4) * kTileRowBytes + 255) / 256 * 256;
    off += (size_t)B * t.ntb * t.blob_stride;
    t.total = off;
    return t;
}

__host__ __device__ inline bool tile_lists_supported(int N, int D, int K)
{
    return D >= 1 && D <= kTileMaxNdim && K >= 1 && K <= kTileMaxK && N >= 1;
}

}  // namespace spnb
