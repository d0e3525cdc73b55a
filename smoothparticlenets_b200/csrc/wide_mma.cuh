// Shared pieces of the wide-channel ConvSP kernels (convsp_wide_mma.cu: forward, convsp_wide_bwd.cu: backward):
// tcgen05 / tensor-memory PTX wrappers, the shared-memory operand descriptors, and the CUDA-core gather that builds
// the G operand images.
#pragma once
#include <stdlib.h>

#include "list_walk.cuh"
#include "spnb_common.cuh"

namespace spnb {

namespace wide {

constexpr int kMQ = 128;            // queries per GEMM CTA = UMMA M
constexpr int kTQ = 8;              // queries per gather CTA (one warp each) = one 8-row group of a tile
constexpr int kGThreads = kTQ * 32;
constexpr int kGemmThreads = 5 * 32;  // warp 0: TMA + MMA issue; warps 1-4: accumulator flush + epilogue

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, unsigned ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 128 x N x 8, TF32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                          unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* v)
{
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor of a K-major operand without swizzle (cute::UMMA::SmemDescriptor, sm_100):
// core matrices of 8 rows x 16 bytes stored as 128 contiguous bytes; `lbo` = byte distance between the two core
// matrices an instruction's K = 8 spans, `sbo` = byte distance between 8-row groups.
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo, unsigned sbo)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fffu);
    d |= (unsigned long long)((lbo >> 4) & 0x3fffu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;  // descriptor version of sm_100
    return d;         // base offset 0, layout type 0 = no swizzle
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A and B TF32, both K-major, M x N
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// hi + lo split without the (slow) cvt.rna.tf32: hi = x rounded to 10 mantissa bits with integer arithmetic (round half
// up in magnitude), lo = x - hi exactly; lo is handed to the tensor core as it is (kind::tf32 ignores the low 13
// mantissa bits: 2^-21 of x at most, the order of the dropped lo * lo term)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo)
{
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// ---- gather: G tiles as A operand images --------------------------------------------------------------------
// gimg[tile][cell][part][ki][mi][8][4] floats: tile = 128 consecutive queries of the chunk, mi = 8-row group,
// element (row = 8 mi + r, channel = 4 ki + e).  One CTA = 8 queries = one mi of one tile, one warp per query.
//
// A warp first compacts its query's neighbour list: the entries that survive the reference's cull test
// (common_funcs.h:497-505) go to shared memory as (x, y, z, index), in list order.  The kernel cells are then
// handled 32 at a time, ONE CELL PER LANE with the query's G row of that cell -- all C channels -- in the lane's
// registers: per surviving neighbour the lane evaluates its cell's in-radius predicate and kernel weight, and the
// neighbour's feature row (the same address for all lanes: a broadcast load) is multiplied in with C predicated
// FMAs.  No shared-memory read-modify-write, no shuffles; the sums run in list order like the reference's.
constexpr int kGStage = 128;  // neighbours staged per round (lists longer than this are staged once per cell pass)

template <int D, int C>
struct GatherLayout {
    static constexpr int CS = C + 4;        // floats per (query, cell): lanes 16 bytes apart mod 128 -> float4 stores
    static constexpr int QS = 32 * CS + 4;  // floats per query: = 4 (mod 32), the transposed read is conflict-free
    static constexpr int NB = kTQ * QS;              // float offset of the staged neighbours (float4 each)
    static constexpr int ROWS = NB + kTQ * kGStage * 4;  // float offset of the per-warp feature-row rings
    static constexpr int SW = ROWS + kTQ * 2 * C;    // float offset of the per-warp cell weights (backward)
    static constexpr size_t bytes = sizeof(float) * ((size_t)SW + (size_t)kTQ * 32);
};

// Feature row of a staged neighbour -> this warp's ring slot, C/32 floats per lane (cp.async, no registers).
template <int C>
__device__ __forceinline__ void row_prefetch(float* ring_slot, const float* __restrict__ sd, float nb_index, int lane,
                                             bool issue)
{
    if (issue) {
        const float* src = sd + (size_t)(int)nb_index * C + lane * (C / 32);
        const unsigned dst = smem_u32(ring_slot + lane * (C / 32));
        if (C == 64) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// One round of the list: entries [j0, j0 + kGStage) -> survivors in s_nb (list order).  Returns the number of
// survivors; `ended` is set when the list's terminator was seen.
template <int D>
__device__ __forceinline__ int gather_stage(const float* __restrict__ row, int K, int j0, const float* __restrict__ sl,
                                            const float* x, float cull2, float4* s_nb, int lane, bool& ended)
{
    int n = 0;
    for (int r0 = j0; r0 < K && r0 < j0 + kGStage && !ended; r0 += 32) {
        const int jj = r0 + lane;
        const float nb = jj < K ? row[jj] : -1.0f;
        const unsigned neg = __ballot_sync(0xffffffffu, !(nb >= 0.0f));
        const unsigned before = neg ? ((1u << (__ffs(neg) - 1)) - 1u) : 0xffffffffu;  // entries ahead of the terminator
        if (neg) ended = true;
        bool keep = (before >> lane) & 1u;
        float4 rec = make_float4(0.0f, 0.0f, 0.0f, nb);
        if (keep) {
            const float* y = sl + (size_t)(int)nb * D;
            float d0 = 0.0f;
            float yy[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < D; ++k) {
                yy[k] = y[k];
                d0 += (x[k] - yy[k]) * (x[k] - yy[k]);
            }
            rec.x = yy[0]; rec.y = yy[1]; rec.z = yy[2];
            keep = !(d0 > cull2);
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) s_nb[n + __popc(km & lanemask_lt())] = rec;
        n += __popc(km);
    }
    __syncwarp();
    return n;
}

// T == false: gimg as above.  T == true (dweight): the transposed image gimg[tile][cell group][part][kq][mi][8][4]
// with rows (cell within the group of 128/C cells, channel) and K = the tile's queries: element
// (row = 8 mi + r, query = 4 kq + e).
template <int D, int C, bool T = false>
__global__ void __launch_bounds__(kGThreads, 2)
k_wide_gather(const float* __restrict__ qlocs, const float* __restrict__ locs, const float* __restrict__ data,
              const float* __restrict__ neighbors, int q_first, int M, int N, int K, int ncells,
              float radius, const float* __restrict__ ksize, const float* __restrict__ dilation, int dis_norm,
              SphParams sp, float* __restrict__ gimg)
{
    using L = GatherLayout<D, C>;
    extern __shared__ __align__(16) float s_G[];  // [kTQ][QS] G rows of one cell pass, then the staged neighbours
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int m = q_first + blockIdx.x * kTQ + warp;  // query within the scene
    const bool active = m < M;
    const size_t q = (size_t)b * M + (active ? m : 0);
    float* Gq = s_G + (size_t)warp * L::QS;
    float4* s_nb = reinterpret_cast<float4*>(s_G + L::NB) + (size_t)warp * kGStage;
    float* s_row = s_G + L::ROWS + (size_t)warp * 2 * C;  // two feature rows in flight

    int ks[D], half[D];
    float dil[D], x[D];
    float maxdil = dilation[0], maxks = ksize[0];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        ks[k] = (int)ksize[k];
        half[k] = ((int)ksize[k]) / 2;
        dil[k] = dilation[k];
        if (dilation[k] > maxdil) maxdil = dilation[k];
        if (ksize[k] > maxks) maxks = ksize[k];
        x[k] = qlocs[q * D + k];
    }
    const float nr = radius + ((int)maxks / 2) * maxdil * fast_root_dim(D);
    const float cull2 = nr * nr, rad2 = radius * radius;
    const float* row = neighbors + q * K;
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    const SphF sf = {sp.H, 1.0f / sp.H, sp.H * sp.H};
    const float wcoef = (float)(sp.w_expr == E_DSPIKY ? sp.w_coef / (double)sp.H : sp.w_coef);  // sph_fast's convention
    // this CTA's place in the image buffer
    const int tile = blockIdx.x / (kMQ / kTQ), mi = blockIdx.x % (kMQ / kTQ);
    const size_t img_cell = (size_t)2 * kMQ * C;  // floats per (tile, cell): hi + lo
    float* gtile = gimg + ((size_t)b * gridDim.x / (kMQ / kTQ) + tile) * ncells * img_cell;

    const bool one_round = K <= kGStage;
    int n_staged = 0;
    bool ended = false;
    if (active && one_round) n_staged = gather_stage<D>(row, K, 0, sl, x, cull2, s_nb, lane, ended);

    for (int cell0 = 0; cell0 < ncells; cell0 += 32) {
        const int ncs = min(32, ncells - cell0);
        const bool valid = lane < ncs;
        // this lane's kernel cell: query position + cell offset, dimension 0 fastest (common_funcs.h:494,575-580)
        float xo[D];
        {
            int rem = valid ? cell0 + lane : 0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int ik = rem % ks[k];
                rem /= ks[k];
                xo[k] = x[k] + (ik - half[k]) * dil[k];
            }
        }
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.0f;
        if (active) {
            bool fin = false;
            for (int j0 = 0; j0 < K && !fin; j0 += kGStage) {
                if (!one_round) {
                    __syncwarp();
                    if (j0 == 0) ended = false;
                    n_staged = gather_stage<D>(row, K, j0, sl, x, cull2, s_nb, lane, ended);
                }
                fin = one_round || ended;
                if (n_staged > 0) row_prefetch<C>(s_row, sd, s_nb[0].w, lane, true);
                for (int i = 0; i < n_staged; ++i) {
                    const float4 rec = s_nb[i];  // broadcast
                    float d = 0.0f;
                    {
                        const float yy[3] = {rec.x, rec.y, rec.z};
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float t = xo[k] - yy[k];
                            d += t * t;
                        }
                    }
                    const bool hit = valid && d < rad2;
                    const bool any = __any_sync(0xffffffffu, hit);  // also: every lane is done with row i-1
                    asm volatile("cp.async.wait_group 0;" ::: "memory");  // my part of row i has landed
                    row_prefetch<C>(s_row + ((i + 1) & 1) * C, sd, i + 1 < n_staged ? s_nb[i + 1].w : 0.0f, lane,
                                    i + 1 < n_staged);
                    if (!any) continue;
                    float s = 0.0f;
                    if (hit) {
                        const float dist = sqrtf(d);  // exact: decides the d > H guard like the reference
                        s = dist > sp.H ? 0.0f : sph_fast(sp.w_expr, dist, d, wcoef, sf);
                        if (dis_norm && dist > 0.0f) s *= fast_rsqrt(d);
                    }
                    __syncwarp();  // ... and everybody else's
                    const float4* dj = reinterpret_cast<const float4*>(s_row + (i & 1) * C);
#pragma unroll
                    for (int c4 = 0; c4 < C / 4; ++c4) {
                        const float4 v = dj[c4];  // same address in every lane: one broadcast wavefront
                        if (hit) {
                            acc[4 * c4 + 0] = fmaf(s, v.x, acc[4 * c4 + 0]);
                            acc[4 * c4 + 1] = fmaf(s, v.y, acc[4 * c4 + 1]);
                            acc[4 * c4 + 2] = fmaf(s, v.z, acc[4 * c4 + 2]);
                            acc[4 * c4 + 3] = fmaf(s, v.w, acc[4 * c4 + 3]);
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
        }
        // registers -> shared G rows of this pass
        {
            float4* g4 = reinterpret_cast<float4*>(Gq + (size_t)lane * L::CS);
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4)
                g4[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
        }
        __syncthreads();
        if (!T) {
            // the pass as core matrices: (cell, ki) -> 8 queries x 4 channels = 128 contiguous bytes, hi and lo;
            // a lane moves one query's 4 channels, a quarter warp one core matrix
            const int qq = lane & 7, kk = lane >> 3;
            for (int t = warp; t < ncs * (C / 16); t += kTQ) {
                const int cl = t / (C / 16), ki = 4 * (t % (C / 16)) + kk;
                const float4 v = *reinterpret_cast<const float4*>(s_G + (size_t)qq * L::QS + cl * L::CS + 4 * ki);
                float4 hi, lo;
                split_tf32(v.x, hi.x, lo.x);
                split_tf32(v.y, hi.y, lo.y);
                split_tf32(v.z, hi.z, lo.z);
                split_tf32(v.w, hi.w, lo.w);
                float* dst = gtile + (size_t)(cell0 + cl) * img_cell + ((size_t)ki * (kMQ / 8) + mi) * 32 + qq * 4;
                *reinterpret_cast<float4*>(dst) = hi;
                *reinterpret_cast<float4*>(dst + (size_t)kMQ * C) = lo;
            }
        } else {
            constexpr int CPG = kMQ / C;
            const int ncg = (ncells + CPG - 1) / CPG;
            float* ttile = gimg + ((size_t)b * gridDim.x / (kMQ / kTQ) + tile) * ncg * ((size_t)2 * kMQ * kMQ);
            // core matrix (cell, ci, query half): 8 channels x 4 queries
            for (int t = warp; t < ncs * (C / 8) * 2; t += kTQ) {
                const int qh = t & 1, ci = (t >> 1) % (C / 8), cl = (t >> 1) / (C / 8);
                const float v = s_G[(size_t)(4 * qh + (lane & 3)) * L::QS + cl * L::CS + 8 * ci + (lane >> 2)];
                float hi, lo;
                split_tf32(v, hi, lo);
                const int cell = cell0 + cl, cg = cell / CPG, rowgrp = (cell % CPG) * (C / 8) + ci, kq = 2 * mi + qh;
                float* dst = ttile + (size_t)cg * ((size_t)2 * kMQ * kMQ) + ((size_t)kq * (kMQ / 8) + rowgrp) * 32 + lane;
                dst[0] = hi;
                dst[(size_t)kMQ * kMQ] = lo;
            }
        }
        __syncthreads();
    }
}


// ---- the gather on the tensor cores -----------------------------------------------------------------------------
// Per query the gather IS a small GEMM: G[cell, c] = sum_j S[cell, j] * data[j, c] with S[cell, j] = W * norm of the
// in-radius (cell, neighbour) pairs and 0 elsewhere: M = 128 kernel cells, N = C channels, K = the query's
// (compacted) neighbours.  One CTA owns 8 queries (one 8-row group of an image tile):
//   * each warp compacts one query's list (gather_stage);
//   * per query (and 64 neighbours) all warps build the two operands in shared memory as the tensor core's K-major
//     core-matrix images, split hi + lo TF32: A = S (a lane evaluates one (cell, neighbour) pair: exact fp32
//     predicate, sph_fast weight; 32 lanes = one 128-byte core matrix, conflict-free stores), B = the neighbours'
//     feature rows transposed; fence.proxy.async, then one thread issues the 3xTF32 tcgen05.mma chain into the
//     query's own TMEM accumulator (8 x C columns = all of tensor memory at C = 64); the operands are double
//     buffered, so the MMAs of query r run under the operand build of query r + 1;
//   * epilogue: a thread = one kernel cell pulls its 8 queries' values out of TMEM and writes whole 128-byte core
//     matrices of the operand image (T == false: rows = queries, K = channels; T == true: rows = (cell, channel),
//     K = queries).
// The sums differ from the reference's only in order (and by the 2^-22 relative lo*lo terms).
constexpr int kTcK = 64;  // neighbours per operand build
#ifndef SPNB_TC_THREADS
#define SPNB_TC_THREADS 512
#endif
constexpr int kTcThreads = SPNB_TC_THREADS;  // one CTA per SM (all of tensor memory): the warps that hide the operand build's latency
constexpr int kTcWarps = kTcThreads / 32;

struct GatherTcSmem {
    unsigned long long free_[2], done;
    unsigned tmem_base;
    int n[kTQ];
    float4 qpos[kTQ];
};

template <int C>
struct GatherTcLayout {
    static constexpr int A_PART = kMQ * kTcK * 4;          // bytes of one part (hi or lo) of S
    static constexpr int B_PART = C * kTcK * 4;
    static constexpr int A_OFF = 0;                        // two buffers of [hi | lo]
    static constexpr int B_OFF = 2 * 2 * A_PART;
    static constexpr int NB_OFF = B_OFF + 2 * 2 * B_PART;  // staged neighbours: kTQ x kGStage float4
    static constexpr int OFF_OFF = NB_OFF + kTQ * kGStage * 16;  // kernel-cell offsets: 128 float4
    static constexpr int CTL_OFF = OFF_OFF + kMQ * 16;
    static constexpr size_t bytes = CTL_OFF + sizeof(GatherTcSmem) + 1024;  // + alignment slack
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4_nowait(unsigned taddr, float* v)
{
    unsigned r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld8_nowait(unsigned taddr, float* v)
{
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int D, int C, bool T>
__global__ void __launch_bounds__(kTcThreads, 1)
k_wide_gather_tc(const float* __restrict__ qlocs, const float* __restrict__ locs, const float* __restrict__ data,
                 const float* __restrict__ neighbors, int q_first, int M, int N, int K, int ncells, float radius,
                 const float* __restrict__ ksize, const float* __restrict__ dilation, int dis_norm, SphParams sp,
                 float* __restrict__ gimg)
{
    using L = GatherTcLayout<C>;
    extern __shared__ unsigned char s_dyn[];
    unsigned char* s_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<size_t>(s_dyn) + 1023) & ~(size_t)1023);
    float4* s_nb_all = reinterpret_cast<float4*>(s_raw + L::NB_OFF);
    float4* s_off = reinterpret_cast<float4*>(s_raw + L::OFF_OFF);
    GatherTcSmem* sm = reinterpret_cast<GatherTcSmem*>(s_raw + L::CTL_OFF);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    constexpr unsigned A_LBO = (kMQ / 8) * 128, A_SBO = 128, B_LBO = (C / 8) * 128, B_SBO = 128;
    constexpr unsigned tmem_cols = kTQ * C;  // 512 or 256

    // kernel shape, cull radius (common_funcs.h:481-485), the cells' offsets
    int ks[D], half[D];
    float dil[D];
    float maxdil = dilation[0], maxks = ksize[0];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        ks[k] = (int)ksize[k];
        half[k] = ((int)ksize[k]) / 2;
        dil[k] = dilation[k];
        if (dilation[k] > maxdil) maxdil = dilation[k];
        if (ksize[k] > maxks) maxks = ksize[k];
    }
    const float nr = radius + ((int)maxks / 2) * maxdil * fast_root_dim(D);
    const float cull2 = nr * nr, rad2 = radius * radius;
    const SphF sf = {sp.H, 1.0f / sp.H, sp.H * sp.H};
    const float wcoef = (float)(sp.w_expr == E_DSPIKY ? sp.w_coef / (double)sp.H : sp.w_coef);
    if (tid < kMQ) {
        float o[3] = {0.0f, 0.0f, 0.0f};
        int rem = tid < ncells ? tid : 0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int ik = rem % ks[k];
            rem /= ks[k];
            o[k] = (ik - half[k]) * dil[k];
        }
        s_off[tid] = make_float4(o[0], o[1], o[2], tid < ncells ? 1.0f : 0.0f);
    }
    if (tid == 0) {
        mbar_init(&sm->free_[0], 1);
        mbar_init(&sm->free_[1], 1);
        mbar_init(&sm->done, 1);
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, tmem_cols);

    // ---- warp r compacts the list of query r
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    if (warp < kTQ) {
        const int m = q_first + blockIdx.x * kTQ + warp;
        int n = 0;
        float x[3] = {0.0f, 0.0f, 0.0f};
        if (m < M) {
            const size_t q = (size_t)b * M + m;
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = qlocs[q * D + k];
            bool ended = false;
            n = gather_stage<D>(neighbors + q * K, K, 0, sl, x, cull2, s_nb_all + warp * kGStage, lane, ended);
        }
        if (lane == 0) {
            sm->n[warp] = n;
            sm->qpos[warp] = make_float4(x[0], x[1], x[2], 0.0f);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = sm->tmem_base;

    // ---- units of work: (query r, 64 neighbours)
    int unit = 0;
    for (int r = 0; r < kTQ; ++r) {
        const int n_r = sm->n[r];
        const float4 xq = sm->qpos[r];
        const float4* nb = s_nb_all + r * kGStage;
        for (int j0 = 0; j0 < n_r; j0 += kTcK, ++unit) {
            const int buf = unit & 1;
            const int cnt = min(kTcK, n_r - j0);
            const int nkq = ((cnt + 7) >> 3) << 1;  // 4-neighbour groups, whole K steps of 8
            if (unit >= 2) mbar_wait(&sm->free_[buf], (unsigned)((unit - 2) >> 1) & 1u);
            float* a_hi = reinterpret_cast<float*>(s_raw + L::A_OFF + buf * 2 * L::A_PART);
            float* a_lo = a_hi + L::A_PART / 4;
            float* b_hi = reinterpret_cast<float*>(s_raw + L::B_OFF + buf * 2 * L::B_PART);
            float* b_lo = b_hi + L::B_PART / 4;
            // B = the neighbours' features: core matrix (kq, ci) = 8 channels x 4 neighbours; the global loads go first,
            // their latency passes under the A build
            constexpr int kBIter = (2 * (kTcK / 8) * (C / 8) + kTcWarps - 1) / kTcWarps;
            float bv[kBIter];
#pragma unroll
            for (int u = 0; u < kBIter; ++u) {
                const int it = warp + u * kTcWarps;
                const int ci = it % (C / 8), kq = it / (C / 8);
                const int c = 8 * ci + (lane >> 2), j = j0 + 4 * kq + (lane & 3);
                bv[u] = (it < nkq * (C / 8) && j < n_r) ? __ldg(sd + (size_t)(int)nb[j].w * C + c) : 0.0f;
            }
            // A = S: core matrix (kq, mi) = 8 cells x 4 neighbours, lane = (cell & 7) * 4 + (j & 3)
            for (int it = warp; it < nkq * (kMQ / 8); it += kTcWarps) {
                const int mi = it & (kMQ / 8 - 1), kq = it / (kMQ / 8);
                const int cell = 8 * mi + (lane >> 2), j = j0 + 4 * kq + (lane & 3);
                const float4 off = s_off[cell];
                float s = 0.0f;
                if (j < n_r && off.w != 0.0f) {
                    const float4 y = nb[j];
                    const float xo[3] = {xq.x + off.x, xq.y + off.y, xq.z + off.z};
                    const float yy[3] = {y.x, y.y, y.z};
                    float d = 0.0f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float t = xo[k] - yy[k];
                        d += t * t;
                    }
                    if (d < rad2) {
                        const float dist = sqrtf(d);  // exact: decides the d > H guard like the reference
                        s = dist > sp.H ? 0.0f : sph_fast(sp.w_expr, dist, d, wcoef, sf);
                        if (dis_norm && dist > 0.0f) s *= fast_rsqrt(d);
                    }
                }
                float hi, lo;
                split_tf32(s, hi, lo);
                a_hi[it * 32 + lane] = hi;
                a_lo[it * 32 + lane] = lo;
            }
            // B: the loads were issued before the A build
#pragma unroll
            for (int u = 0; u < kBIter; ++u) {
                const int it = warp + u * kTcWarps;
                if (it < nkq * (C / 8)) {
                    float hi, lo;
                    split_tf32(bv[u], hi, lo);
                    b_hi[it * 32 + lane] = hi;
                    b_lo[it * 32 + lane] = lo;
                }
            }
            fence_proxy_async();  // the tensor core reads shared memory through the async proxy
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const unsigned idesc = umma_idesc_tf32(kMQ, C);
                const unsigned acc = tmem + (unsigned)(r * C);
                const unsigned ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
                for (int kk = 0; kk < (nkq >> 1); ++kk) {
                    const unsigned ao = (unsigned)kk * 2u * A_LBO, bo = (unsigned)kk * 2u * B_LBO;
                    const unsigned long long dah = umma_desc(ah + ao, A_LBO, A_SBO), dal = umma_desc(al + ao, A_LBO, A_SBO);
                    const unsigned long long dbh = umma_desc(bh + bo, B_LBO, B_SBO), dbl = umma_desc(bl + bo, B_LBO, B_SBO);
                    umma_tf32(acc, dal, dbh, idesc, (j0 > 0 || kk > 0) ? 1u : 0u);
                    umma_tf32(acc, dah, dbl, idesc, 1u);
                    umma_tf32(acc, dah, dbh, idesc, 1u);
                }
                umma_commit(&sm->free_[buf]);
            }
        }
    }
    if (tid == 0) umma_commit(&sm->done);  // arrives when every MMA above is complete
    mbar_wait(&sm->done, 0);
    tc_fence_after();

    // ---- epilogue: thread = kernel cell (TMEM lane); warps w, w + 4, ... share a lane quarter
    const int quarter = warp & 3, hsel = warp >> 2;
    const int cell = quarter * 32 + lane;
    const unsigned trow = tmem + ((unsigned)(quarter * 32) << 16);
    const int tile = blockIdx.x / (kMQ / kTQ), mi = blockIdx.x % (kMQ / kTQ);
    if (!T) {
        const size_t img_cell = (size_t)2 * kMQ * C;
        float* gtile = gimg + ((size_t)b * gridDim.x / (kMQ / kTQ) + tile) * ncells * img_cell;
        for (int ki = hsel; ki < C / 4; ki += kTcWarps / 4) {
            float v[kTQ][4];
#pragma unroll
            for (int q = 0; q < kTQ; ++q) {
                if (sm->n[q] > 0) tmem_ld4_nowait(trow + (unsigned)(q * C + 4 * ki), v[q]);
                else v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.0f;
            }
            tmem_wait_ld();
            if (cell < ncells) {
                float4* dst = reinterpret_cast<float4*>(gtile + (size_t)cell * img_cell + ((size_t)ki * (kMQ / 8) + mi) * 32);
                float4* dst_lo = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + (size_t)kMQ * C);
#pragma unroll
                for (int q = 0; q < kTQ; ++q) {
                    float4 hi, lo;
                    split_tf32(v[q][0], hi.x, lo.x);
                    split_tf32(v[q][1], hi.y, lo.y);
                    split_tf32(v[q][2], hi.z, lo.z);
                    split_tf32(v[q][3], hi.w, lo.w);
                    dst[q] = hi;
                    dst_lo[q] = lo;
                }
            }
        }
    } else {
        constexpr int CPG = kMQ / C;
        const int ncg = (ncells + CPG - 1) / CPG;
        float* ttile = gimg + ((size_t)b * gridDim.x / (kMQ / kTQ) + tile) * ncg * ((size_t)2 * kMQ * kMQ);
        const int cg = cell / CPG;
        // core matrix (ci, query half qh): 8 channels x 4 queries
        for (int ci = hsel; ci < C / 8; ci += kTcWarps / 4) {
#pragma unroll
            for (int qh = 0; qh < 2; ++qh) {
                float v[4][8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int q = 4 * qh + e;
                    if (sm->n[q] > 0) tmem_ld8_nowait(trow + (unsigned)(q * C + 8 * ci), v[e]);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[e][i] = 0.0f;
                    }
                }
                tmem_wait_ld();
                if (cell < ncells) {
                    const int rowgrp = (cell % CPG) * (C / 8) + ci, kq = 2 * mi + qh;
                    float4* dst = reinterpret_cast<float4*>(ttile + (size_t)cg * ((size_t)2 * kMQ * kMQ) +
                                                            ((size_t)kq * (kMQ / 8) + rowgrp) * 32);
                    float4* dst_lo = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + (size_t)kMQ * kMQ);
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        float4 hi, lo;
                        split_tf32(v[0][rr], hi.x, lo.x);
                        split_tf32(v[1][rr], hi.y, lo.y);
                        split_tf32(v[2][rr], hi.z, lo.z);
                        split_tf32(v[3][rr], hi.w, lo.w);
                        dst[rr] = hi;
                        dst_lo[rr] = lo;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, tmem_cols);
    }
}

}  // namespace wide
}  // namespace spnb
