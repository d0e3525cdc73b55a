// Neighbour-list walker shared by the ConvSP kernels (convsp_small.cu, convsp_group.cu).
//
// G lanes cooperate on one query (G = 1, 2, 4 or 8; R = 32/G queries per warp).  The rows of the
// warp's queries are staged 32 entries at a time through shared memory: for each row the warp reads
// one coalesced 128-byte segment, and the same pass finds the row's terminator with ONE ballot (the
// list ends at the first negative entry, common_funcs.h:476), so the inner loop needs no per-entry
// termination logic.  Lane `sub` of a group then consumes entries sub, sub+G, ... of its row, U at a
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

template <int G>
struct WalkSmem {
    static constexpr int R = 32 / G;       // queries per warp
    static constexpr int STRIDE = 32 + G;  // floats per staged row: group g starts at bank g*G
    float nb[R * STRIDE];
    int cnt[R];
};

// warp_rows: row of the warp's first query (rows are K floats apart).  nrows: how many of the warp's
// R queries exist.  body(const int* j, const bool* valid) is called with U list entries of this
// lane's query (valid[u] == false: no entry).
template <int G, int U, typename Body>
__device__ __forceinline__ void walk_rows(const float* __restrict__ warp_rows, int K, int nrows,
                                          WalkSmem<G>& sm, Body body)
{
    constexpr int R = WalkSmem<G>::R, STRIDE = WalkSmem<G>::STRIDE;
    const int lane = threadIdx.x & 31;
    const int g = lane / G, sub = lane % G;
    unsigned live = nrows >= 32 ? 0xffffffffu : ((1u << nrows) - 1u);  // rows still walking
    for (int base = 0; base < K && live; base += 32) {
        const unsigned start_live = live;
        const int p = base + lane;
        __syncwarp();  // the previous chunk has been consumed by every lane
        if (start_live == (R >= 32 ? 0xffffffffu : ((1u << R) - 1u))) {
            // all rows: batches of 8 independent loads
#pragma unroll
            for (int r0 = 0; r0 < R; r0 += 8) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (r0 + i < R) f[i] = p < K ? warp_rows[(size_t)(r0 + i) * K + p] : -1.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (r0 + i < R) {
                        sm.nb[(r0 + i) * STRIDE + lane] = f[i];
                        const unsigned neg = __ballot_sync(0xffffffffu, !(f[i] >= 0.0f));
                        if (lane == 0) sm.cnt[r0 + i] = neg ? __ffs(neg) - 1 : 32;
                        if (neg) live &= ~(1u << (r0 + i));
                    }
            }
        } else {
            for (unsigned m = start_live; m; m &= m - 1) {
                const int r = __ffs(m) - 1;
                const float f = p < K ? warp_rows[(size_t)r * K + p] : -1.0f;
                sm.nb[r * STRIDE + lane] = f;
                const unsigned neg = __ballot_sync(0xffffffffu, !(f >= 0.0f));
                if (lane == 0) sm.cnt[r] = neg ? __ffs(neg) - 1 : 32;
                if (neg) live &= ~(1u << r);
            }
        }
        __syncwarp();
        const int cnt = ((start_live >> g) & 1u) ? sm.cnt[g] : 0;
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
