// Neighbour-list walker shared by the ConvSP kernels (convsp_small.cu, convsp_group.cu).
//
// G lanes cooperate on one query (G = 1, 2, 4 or 8; R = 32/G queries per warp).  The rows of the
// warp's queries are staged 32 entries (128 bytes) at a time in shared memory and the row's
// terminator (the list ends at the first negative entry, common_funcs.h:476) is located once per
// chunk, so the inner loop needs no per-entry termination logic.  Lane `sub` of a group then
// consumes entries sub, sub+G, ... of its row, U at a
// time, which gives U independent gathers in flight per lane; the G lanes of a group touch G
// consecutive list entries, i.e. (the lists being in cell order) mostly consecutive particles.
//
// G trades instruction overhead against cache footprint (measured, profiles/README.md): G = 8 keeps
// few queries in flight per SM (L1 hit rate > 80 %) but spends more than half of its issued
// instructions on per-group bookkeeping and reductions; G = 1 has almost no overhead but 8x more
// queries in flight, and its gathers become latency-bound on L1 misses.
#pragma once
#include <cuda_runtime.h>

namespace spnb {

// Row staging: 0 = cp.async (LDGSTS, 16 bytes per lane), 1 = one cp.async.bulk (TMA 1-D bulk copy,
// UBLKCP) per 128-byte row chunk completing on a per-warp mbarrier.  Both are kept; the A/B on one B200
// (tools/ab_test.sh, profiles/README.md) has LDGSTS 5-10 % faster for these 128-byte transfers (the
// mbarrier arm/try_wait round trip per chunk costs more than it saves), so it is the default.
#ifndef SPNB_WALK_TMA
#define SPNB_WALK_TMA 0
#endif

template <int G>
struct WalkSmem {
    static constexpr int R = 32 / G;       // queries per warp
    static constexpr int STRIDE = 32 + G;  // floats per staged row: group g starts at bank g*G
    alignas(16) float nb[R * STRIDE];
    int cnt[R];
    alignas(8) unsigned long long bar;     // mbarrier of this warp's row copies (TMA staging)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier.
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// warp_rows: row of the warp's first query (rows are K floats apart).  nrows: how many of the warp's
// R queries exist.  body(const int* j, const bool* valid) is called with U list entries of this
// lane's query (valid[u] == false: no entry).
//
// Staging: when rows are 16-byte aligned and the chunk lies inside the row, every group copies ITS OWN
// row chunk with cp.async (LDGSTS, 16 bytes per lane and instruction, no register round trip); the
// terminator is then found by the group itself: each lane scans the 32/G entries it is going to
// consume anyway and the group takes the minimum with log2(G) shuffles.  Otherwise (ragged K,
// unaligned base) the warp stages row by row with one ballot per row.
template <int G, int U, typename Body>
__device__ __forceinline__ void walk_rows(const float* __restrict__ warp_rows, int K, int nrows,
                                          WalkSmem<G>& sm, Body body)
{
    constexpr int R = WalkSmem<G>::R, STRIDE = WalkSmem<G>::STRIDE, EPL = 32 / G;
    const int lane = threadIdx.x & 31;
    const int g = lane / G, sub = lane % G;
    const bool aligned = ((K & 3) == 0) && ((reinterpret_cast<size_t>(warp_rows) & 15) == 0);
    bool mine = g < nrows;  // my group's row is still being walked
#if SPNB_WALK_TMA
    unsigned parity = 0;
    if (lane == 0) mbar_init(&sm.bar, 1);
    __syncwarp();
#endif
    for (int base = 0; base < K; base += 32) {
        if (!__any_sync(0xffffffffu, mine)) break;
        __syncwarp();  // the previous chunk has been consumed by every lane
        int cnt = 0;
        if (aligned && base + 32 <= K) {
#if SPNB_WALK_TMA
            // one 128-byte TMA bulk copy per live row, issued by the first lane of the row's group;
            // lane 0 arms the warp's mbarrier with the byte count, everybody waits on its phase
            const unsigned live = __ballot_sync(0xffffffffu, mine && sub == 0);
            if (lane == 0) mbar_expect_tx(&sm.bar, 128u * __popc(live));
            __syncwarp();
            if (mine && sub == 0)
                bulk_copy_g2s(&sm.nb[g * STRIDE], warp_rows + (size_t)g * K + base, 128u, &sm.bar);
            mbar_wait(&sm.bar, parity);
            parity ^= 1u;
#else
            if (mine) {
                const float* src = warp_rows + (size_t)g * K + base;
#pragma unroll
                for (int i = 0; i < 8 / G; ++i)  // 128 bytes per row chunk, 16 bytes per lane and copy
                    cp_async16(&sm.nb[g * STRIDE + (sub + i * G) * 4], src + (sub + i * G) * 4);
            }
            cp_async_wait_all();
            __syncwarp();
#endif
            // first negative entry among mine (positions sub, sub+G, ...), then over the group
            int first = 32;
            if (mine) {
#pragma unroll
                for (int i = EPL - 1; i >= 0; --i)
                    if (!(sm.nb[g * STRIDE + sub + i * G] >= 0.0f)) first = sub + i * G;
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            cnt = mine ? first : 0;
        } else {
            const unsigned live = __ballot_sync(0xffffffffu, mine && sub == 0);
            const int p = base + lane;
            for (unsigned m = live; m; m &= m - 1) {
                const int r = (__ffs(m) - 1) / G;
                const float f = p < K ? warp_rows[(size_t)r * K + p] : -1.0f;
                sm.nb[r * STRIDE + lane] = f;
                const unsigned neg = __ballot_sync(0xffffffffu, !(f >= 0.0f));
                if (lane == 0) sm.cnt[r] = neg ? __ffs(neg) - 1 : 32;
            }
            __syncwarp();
            cnt = mine ? sm.cnt[g] : 0;
        }
        if (cnt < 32) mine = false;  // terminator seen (or no row): nothing after this chunk
        for (int t0 = sub; t0 < 32; t0 += G * U) {
            if (!__any_sync(0xffffffffu, t0 < cnt)) break;
            int j[U];
            bool valid[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u * G;
                valid[u] = t < cnt;
                j[u] = valid[u] ? (int)sm.nb[g * STRIDE + t] : 0;
            }
            body(j, valid);
        }
    }
}

// L2 prefetch of the neighbour rows that a LATER block will walk.  The walk is latency-bound on the
// first touch of each row (measured: a walk with no gathers and no math already costs ~60 us at c2,
// 4x the DRAM time of the bytes it reads), because a warp alternates between waiting for its rows
// and consuming them.  Blocks are dispatched in linear order, so a block asks L2 for the rows of the
// block that will run roughly one "wave of resident blocks" later; by then they are L2 hits.
// rows_per_block queries per block, `lines` 128-byte lines per row are prefetched.
#ifndef SPNB_PREFETCH_BLOCKS
#define SPNB_PREFETCH_BLOCKS (148 * 16)
#endif
// (bx, by) of (nbx, gridDim.y-equivalent) is the logical block; by default the launch's own block index.
__device__ __forceinline__ void prefetch_rows_ahead(const float* __restrict__ neighbors, int M, int K,
                                                    int rows_per_block, int lines, int bx, int by, int nbx, int nby)
{
    const long long lb = (long long)by * nbx + bx + SPNB_PREFETCH_BLOCKS;
    if (lb >= (long long)nbx * nby) return;
    const int b2 = (int)(lb / nbx), x2 = (int)(lb % nbx);
    const int t = threadIdx.x;
    const int row = t / lines, line = t % lines;
    if (row >= rows_per_block) return;
    const int m = x2 * rows_per_block + row;
    if (m >= M || line * 32 >= K) return;
    const float* p = neighbors + ((size_t)b2 * M + m) * K + line * 32;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_rows_ahead(const float* __restrict__ neighbors, int M, int K,
                                                    int rows_per_block, int lines)
{
    prefetch_rows_ahead(neighbors, M, K, rows_per_block, lines, blockIdx.x, blockIdx.y, gridDim.x, gridDim.y);
}

// Sum over the G lanes of a group (all lanes receive the total).
template <int G>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 1/sqrt(x) as a single MUFU.RSQ (about 1 ulp); callers guard x > 0.
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace spnb
