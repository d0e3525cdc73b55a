// ConvSP backward for WIDE channel counts on the 5th-generation tensor cores (BASELINE.json config 3: 64 -> 64
// channels, kernel_size 5).
//
// The reference's backward (compute_kernel_cells with non-NULL gradients, src/common_funcs.h:512-572) spends O*C
// multiply-adds AND O*C float atomics on d(weight) per (neighbour, kernel cell) pair.  With the factored form of
// the forward (convsp_wide_mma.cu)
//
//     G[q, cell, c] = sum_j S(q, cell, j) * data[j, c],      S = W(d) * norm for |q + off_cell - x_j| < r
//     out[q, o]     = bias[o] + sum_{cell, c} weight[o, c, cell] * G[q, cell, c]
//
// the gradients are two dense contractions plus one walk over the neighbour lists:
//
//     dG[q, cell, c]       = sum_o go[q, o] * weight[o, c, cell]                  tcgen05 GEMM, K = o
//     dweight[o, c, cell]  = sum_q go[q, o] * G[q, cell, c]                       tcgen05 GEMM, K = q
//     ddata[j, c]         += sum_cell S(q, cell, j) * dG[q, cell, c]
//     t(q, cell, j)        = (sum_c data[j, c] * dG[q, cell, c]) * norm * W'(d)/d * (q + off_cell - x_j)
//     dqlocs[q] += t,  dlocs[j] -= t                                              (d > 0 only, as the reference)
//
// Per chunk of 128-query tiles:
//   k_wide_go_images   go -> the A operand of the dG GEMM (rows q, K = o) and the B operand of the dweight GEMM
//                      (rows o, K = q), both as hi + lo TF32 terms in the tensor core's shared-memory image;
//   k_wide_dg_gemm     per tile: go image resident, the transposed weight images streamed cell by cell with TMA
//                      bulk copies, 3xTF32 tcgen05.mma into alternating TMEM accumulators, epilogue warps transpose 32 x C blocks
//                      through shared memory and store whole rows of dG[tile][cell][q][c];
//   k_wide_scatter     one warp per query, one kernel cell per lane with its dG row in registers: the channel dot
//                      product for the position gradients is C FMAs against the neighbour's broadcast feature
//                      row; ddata goes through the lanes-as-channels view of the same dG tile in shared memory
//                      and one coalesced float atomic per (neighbour, channel) and cell pass;
//   k_wide_gather<T>   the forward's gather, writing G transposed (rows (cell, c), K = q);
//   k_wide_dw_gemm     one CTA per group of 128/C kernel cells and K split: streams the G^T and go images of its
//                      tiles, accumulates per 64 queries in TMEM, flushes into fp32 registers (round to nearest;
//                      the tensor core's own accumulation truncates), adds to dweight at the end.
#include "wide_mma.cuh"

namespace spnb {

using namespace wide;

namespace {

constexpr int kBwdChunkTiles = 296;

// ---- weights as the B operand of dG = go * W: rows c, K = o -----------------------------------------------------
// img[cell][part][ko][ci][8][4]: element (c = 8 ci + r, o = 4 ko + e) of weight[o][c][cell]; o >= O is zero.
__global__ void __launch_bounds__(256)
k_wide_prep_weights_t(const float* __restrict__ w, float* __restrict__ img, int O, int Opad, int C, int ncells)
{
    const long long per = (long long)Opad * C;
    const long long n = per * ncells;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int cell = (int)(i / per);
        const int t = (int)(i % per);
        const int e = t & 3, r = (t >> 2) & 7, rest = t >> 5;
        const int ci = rest % (C / 8), ko = rest / (C / 8);
        const int c = 8 * ci + r, o = 4 * ko + e;
        const float v = o < O ? w[((size_t)o * C + c) * ncells + cell] : 0.0f;
        const float hi = to_tf32(v);
        img[((size_t)cell * 2) * per + t] = hi;
        img[((size_t)cell * 2 + 1) * per + t] = to_tf32(v - hi);
    }
}

// ---- go as operand images ----------------------------------------------------------------------------------------
// goA[bt][part][ko][mi][8][4]: element (q = 8 mi + r, o = 4 ko + e)     (rows q, K = o; bt = scene * tiles + tile)
// goB[bt][part][kq][oi][8][4]: element (o = 8 oi + r, q = 4 kq + e)     (rows o, K = q)
__global__ void __launch_bounds__(256)
k_wide_go_images(const float* __restrict__ go, int q_first, int M, int O, int Opad, int ntiles,
                 float* __restrict__ goA, float* __restrict__ goB)
{
    const int per = kMQ * Opad;  // floats per (bt, part)
    const int b = blockIdx.y;
    const long long n = (long long)ntiles * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int tile = (int)(i / per);
        const int t = (int)(i % per);
        const size_t bt = (size_t)b * ntiles + tile;
        {
            const int e = t & 3, r = (t >> 2) & 7, rest = t >> 5;
            const int mi = rest % (kMQ / 8), ko = rest / (kMQ / 8);
            const int m = q_first + tile * kMQ + 8 * mi + r, o = 4 * ko + e;
            const float v = (m < M && o < O) ? go[((size_t)b * M + m) * O + o] : 0.0f;
            const float hi = to_tf32(v);
            goA[(bt * 2) * per + t] = hi;
            goA[(bt * 2 + 1) * per + t] = to_tf32(v - hi);
        }
        {
            const int e = t & 3, r = (t >> 2) & 7, rest = t >> 5;
            const int oi = rest % (Opad / 8), kq = rest / (Opad / 8);
            const int m = q_first + tile * kMQ + 4 * kq + e, o = 8 * oi + r;
            const float v = (m < M && o < O) ? go[((size_t)b * M + m) * O + o] : 0.0f;
            const float hi = to_tf32(v);
            goB[(bt * 2) * per + t] = hi;
            goB[(bt * 2 + 1) * per + t] = to_tf32(v - hi);
        }
    }
}

// ---- dG[128 q x C] per kernel cell = go[128 x Opad] * Wt_cell[C x Opad]^T ------------------------------------------------
struct DgSmem {
    unsigned long long a_full, full[2], free_[2], acc_free[2];
    unsigned tmem_base;
};

template <int C>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_wide_dg_gemm(const float* __restrict__ goA, const float* __restrict__ wimg_t, int Opad, int ncells,
               float* __restrict__ dgbuf)
{
    constexpr int B_ROWS = C;
    constexpr unsigned A_LBO = (kMQ / 8) * 128, A_SBO = 128;
    constexpr unsigned B_LBO = (B_ROWS / 8) * 128, B_SBO = 128;
    extern __shared__ __align__(1024) unsigned char s_raw[];
    const int A_BYTES = kMQ * Opad * 4;     // one part of the go image
    const int B_BYTES = B_ROWS * Opad * 4;  // one part of a weight image
    const int STAGE = 2 * B_BYTES;
    unsigned char* s_a = s_raw;                    // [A hi | A lo]
    unsigned char* s_b = s_raw + 2 * (size_t)A_BYTES;  // two stages of [B hi | B lo]
    DgSmem* sm = reinterpret_cast<DgSmem*>(s_b + 2 * (size_t)STAGE + (size_t)4 * 32 * (C + 4) * 4);  // after the epilogue's staging tiles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t bt = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    constexpr unsigned tmem_cols = 2 * C <= 64 ? 64u : 128u;

    if (tid == 0) {
        mbar_init(&sm->a_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm->full[i], 1);
            mbar_init(&sm->free_[i], 1);
            mbar_init(&sm->acc_free[i], 4);
        }
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = sm->tmem_base;

    if (warp == 0) {
        const unsigned idesc = umma_idesc_tf32(kMQ, C);
        if (lane == 0) {
            mbar_expect_tx(&sm->a_full, 2u * (unsigned)A_BYTES);
            bulk_copy_g2s(s_a, goA + bt * 2 * (size_t)(kMQ * Opad), 2u * (unsigned)A_BYTES, &sm->a_full);
        }
        auto load = [&](int cell) {
            const int buf = cell & 1;
            if (lane == 0) {
                mbar_expect_tx(&sm->full[buf], 2u * (unsigned)B_BYTES);
                bulk_copy_g2s(s_b + (size_t)buf * STAGE, wimg_t + (size_t)cell * 2 * (size_t)(B_ROWS * Opad),
                              2u * (unsigned)B_BYTES, &sm->full[buf]);
            }
        };
        load(0);
        mbar_wait(&sm->a_full, 0);
        for (int cell = 0; cell < ncells; ++cell) {
            const int buf = cell & 1;
            const unsigned ph = (unsigned)(cell >> 1) & 1u;
            if (cell + 1 < ncells) {
                if (cell >= 1) mbar_wait(&sm->free_[buf ^ 1], (unsigned)((cell - 1) >> 1) & 1u);
                load(cell + 1);
            }
            mbar_wait(&sm->full[buf], ph);
            if (cell >= 2) mbar_wait(&sm->acc_free[buf], (unsigned)((cell - 2) >> 1) & 1u);
            tc_fence_after();
            const unsigned acc = tmem + (unsigned)(buf * C);
            if (lane == 0) {
                const unsigned a_hi = smem_u32(s_a), a_lo = a_hi + (unsigned)A_BYTES;
                const unsigned b_hi = smem_u32(s_b + (size_t)buf * STAGE), b_lo = b_hi + (unsigned)B_BYTES;
#pragma unroll 1
                for (int kk = 0; kk < Opad / 8; ++kk) {
                    const unsigned ao = (unsigned)kk * 2u * A_LBO, bo = (unsigned)kk * 2u * B_LBO;
                    const unsigned long long dah = umma_desc(a_hi + ao, A_LBO, A_SBO), dal = umma_desc(a_lo + ao, A_LBO, A_SBO);
                    const unsigned long long dbh = umma_desc(b_hi + bo, B_LBO, B_SBO), dbl = umma_desc(b_lo + bo, B_LBO, B_SBO);
                    umma_tf32(acc, dal, dbh, idesc, kk != 0 ? 1u : 0u);
                    umma_tf32(acc, dah, dbl, idesc, 1u);
                    umma_tf32(acc, dah, dbh, idesc, 1u);
                }
                umma_commit(&sm->free_[buf]);
            }
            __syncwarp();
        }
    } else {
        // accumulator rows (thread = query) -> this warp's staging tile in shared memory (row stride C + 4 floats:
        // conflict-free 128-bit stores), read back with the lanes along the rows, so that every global store of the
        // warp covers two whole 256-byte rows of dG[tile][cell][query][c]
        const int quarter = warp & 3;
        constexpr int RS = C + 4;
        float* stage = reinterpret_cast<float*>(s_b + 2 * (size_t)STAGE) + (size_t)quarter * 32 * RS;
        float* gdst = dgbuf + (bt * ncells * kMQ + (size_t)quarter * 32) * C;  // + cell * kMQ * C
        for (int cell = 0; cell < ncells; ++cell) {
            const int fb = cell & 1;
            mbar_wait(&sm->free_[fb], (unsigned)(cell >> 1) & 1u);
            tc_fence_after();
            __syncwarp();  // the previous cell's read-back is complete
#pragma unroll
            for (int g = 0; g < C / 32; ++g) {
                float v[32];
                tmem_ld32(tmem + ((unsigned)(quarter * 32) << 16) + (unsigned)(fb * C + g * 32), v);
                float4* row = reinterpret_cast<float4*>(stage + (size_t)lane * RS + g * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) row[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_(&sm->acc_free[fb]);
            constexpr int PPR = C / 4;  // 16-byte pieces per row
            float4* out4 = reinterpret_cast<float4*>(gdst + (size_t)cell * kMQ * C);
#pragma unroll
            for (int t = 0; t < PPR; ++t) {
                const int idx = t * 32 + lane, r = idx / PPR, pc = idx % PPR;
                out4[idx] = *reinterpret_cast<const float4*>(stage + (size_t)r * RS + 4 * pc);
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, tmem_cols);
    }
}

// ---- the list walk of the backward ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int D, int C>
__global__ void __launch_bounds__(kGThreads, 2)
k_wide_scatter(const float* __restrict__ qlocs, const float* __restrict__ locs, const float* __restrict__ data,
               const float* __restrict__ neighbors, int q_first, int M, int N, int K, int ncells, float radius,
               const float* __restrict__ ksize, const float* __restrict__ dilation, int dis_norm, SphParams sp,
               const float* __restrict__ dgbuf, float* dq, float* dl, float* dd, int same_q_l)
{
    using L = GatherLayout<D, C>;
    extern __shared__ __align__(16) float s_G[];  // [kTQ][QS] dG rows of one cell pass, staged neighbours, row rings
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int m = q_first + blockIdx.x * kTQ + warp;
    if (m >= M) return;  // warps are independent here (no block-wide barrier below)
    const size_t q = (size_t)b * M + m;
    float* Gq = s_G + (size_t)warp * L::QS;
    float4* s_nb = reinterpret_cast<float4*>(s_G + L::NB) + (size_t)warp * kGStage;
    float* s_row = s_G + L::ROWS + (size_t)warp * 2 * C;
    float* s_sw = s_G + L::SW + (size_t)warp * 32;

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
    // sph_fast's convention: the 1/H of dspiky is part of the coefficient
    const float wcoef = (float)(sp.w_expr == E_DSPIKY ? sp.w_coef / (double)sp.H : sp.w_coef);
    const float dwcoef = (float)(sp.dw_expr == E_DSPIKY ? sp.dw_coef / (double)sp.H : sp.dw_coef);
    float* dlb = dl ? dl + (size_t)b * N * D : nullptr;
    float* ddb = dd ? dd + (size_t)b * N * C : nullptr;
    const bool want_locs = dq != nullptr || dl != nullptr;
    // dG of this query: dgbuf[tile][cell][query in tile][c]
    const int qt = blockIdx.x * kTQ + warp;  // query within the chunk
    const float* dgq = dgbuf + (((size_t)b * (gridDim.x * kTQ / kMQ) + qt / kMQ) * ncells * kMQ + (qt % kMQ)) * C;

    const bool one_round = K <= kGStage;
    int n_staged = 0;
    bool ended = false;
    if (one_round) n_staged = gather_stage<D>(row, K, 0, sl, x, cull2, s_nb, lane, ended);

    float aq[D];
#pragma unroll
    for (int k = 0; k < D; ++k) aq[k] = 0.0f;

    for (int cell0 = 0; cell0 < ncells; cell0 += 32) {
        const int ncs = min(32, ncells - cell0);
        const bool valid = lane < ncs;
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
        // this lane's cell: dG row -> registers (the dot products) and shared memory (the channel view)
        float dg[C];
        __syncwarp();
        {
            const float4* src = reinterpret_cast<const float4*>(dgq + (size_t)(cell0 + (valid ? lane : 0)) * kMQ * C);
            float4* g4 = reinterpret_cast<float4*>(Gq + (size_t)lane * L::CS);
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                float4 v = valid ? __ldg(src + c4) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                dg[4 * c4] = v.x; dg[4 * c4 + 1] = v.y; dg[4 * c4 + 2] = v.z; dg[4 * c4 + 3] = v.w;
                g4[c4] = v;
            }
        }
        __syncwarp();
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
                const float4 rec = s_nb[i];
                float disp[D];
                float d = 0.0f;
                {
                    const float yy[3] = {rec.x, rec.y, rec.z};
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        disp[k] = xo[k] - yy[k];
                        d += disp[k] * disp[k];
                    }
                }
                const bool hit = valid && d < rad2;
                const unsigned hits = __ballot_sync(0xffffffffu, hit);  // also: every lane is done with row i-1
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                row_prefetch<C>(s_row + ((i + 1) & 1) * C, sd, i + 1 < n_staged ? s_nb[i + 1].w : 0.0f, lane,
                                i + 1 < n_staged);
                if (!hits) continue;
                float s = 0.0f, gscale = 0.0f;
                if (hit) {
                    const float dist = sqrtf(d);  // exact: decides the d > H guard like the reference
                    const bool in = !(dist > sp.H);
                    const float inv = dist > 0.0f ? fast_rsqrt(d) : 0.0f;
                    const float norm = (dis_norm && dist > 0.0f) ? inv : 1.0f;
                    s = (in ? sph_fast(sp.w_expr, dist, d, wcoef, sf) : 0.0f) * norm;
                    gscale = norm * (in ? sph_fast(sp.dw_expr, dist, d, dwcoef, sf) : 0.0f) * inv;  // 0 at d == 0
                }
                __syncwarp();
                const int j = (int)rec.w;
                if (want_locs) {
                    const float4* dj = reinterpret_cast<const float4*>(s_row + (i & 1) * C);
                    float dot0 = 0.0f, dot1 = 0.0f;
#pragma unroll
                    for (int c4 = 0; c4 < C / 4; ++c4) {
                        const float4 v = dj[c4];  // broadcast
                        dot0 = fmaf(v.x, dg[4 * c4 + 0], dot0);
                        dot1 = fmaf(v.y, dg[4 * c4 + 1], dot1);
                        dot0 = fmaf(v.z, dg[4 * c4 + 2], dot0);
                        dot1 = fmaf(v.w, dg[4 * c4 + 3], dot1);
                    }
                    const float tt = (dot0 + dot1) * gscale;  // 0 for lanes without a hit or with d == 0
                    float t[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        t[k] = hit ? tt * disp[k] : 0.0f;
                        aq[k] += t[k];
                    }
                    if (dlb) {
                        float mine = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float r = warp_sum(t[k]);
                            if (lane == k) mine = r;
                        }
                        if (lane < D) atomicAdd(dlb + (size_t)j * D + lane, -mine);
                    }
                }
                if (ddb) {
                    // lanes as channels (C = 64: the pair 2 lane, 2 lane + 1): sum over the cells of S * dG[cell][c];
                    // the cell weights go through shared memory, four per broadcast load
                    s_sw[lane] = s;  // 0 where the cell is not hit
                    __syncwarp();
                    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        if (!((hits >> (4 * c4)) & 0xFu)) continue;
                        const float4 s4 = *reinterpret_cast<const float4*>(s_sw + 4 * c4);
                        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float* g = Gq + (size_t)(4 * c4 + e) * L::CS;
                            if (C == 64) {
                                const float2 v = *reinterpret_cast<const float2*>(g + 2 * lane);
                                a0 = fmaf(sv[e], v.x, a0);
                                a1 = fmaf(sv[e], v.y, a1);
                            } else {
                                a0 = fmaf(sv[e], g[lane], a0);
                            }
                        }
                    }
                    if (C == 64) atomicAdd(reinterpret_cast<float2*>(ddb + (size_t)j * C) + lane, make_float2(a0, a1));
                    else atomicAdd(ddb + (size_t)j * C + lane, a0);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
    }
    if (dq) {
        float mine = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float r = warp_sum(aq[k]);
            if (lane == k) mine = r;
        }
        if (lane < D) {
            if (same_q_l) atomicAdd(dq + q * D + lane, mine);
            else dq[q * D + lane] = mine;
        }
    }
}

// ---- dweight: D[(cell, c) x Opad] += G^T[(cell, c) x q] * go[Opad x q]^T over the queries --------------------------
struct DwSmem {
    unsigned long long full[2], free_[2], acc_free[2];
    unsigned tmem_base;
};

template <int C>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_wide_dw_gemm(const float* __restrict__ gimg_t, const float* __restrict__ goB, int n_bt, int O, int Opad, int ncells,
               int ncg, float* dweight)
{
    constexpr int CPG = kMQ / C;  // kernel cells per 128-row group
    constexpr int HQ = 64;        // queries per stage
    constexpr int A_HALF = kMQ * HQ * 4;  // one part, half a tile
    constexpr unsigned A_LBO = (kMQ / 8) * 128, A_SBO = 128;
    extern __shared__ __align__(1024) unsigned char s_raw[];
    const int B_HALF = Opad * HQ * 4;
    const int STAGE = 2 * A_HALF + 2 * B_HALF;  // [A hi | A lo | B hi | B lo]
    const unsigned B_LBO = (unsigned)(Opad / 8) * 128, B_SBO = 128;
    DwSmem* sm = reinterpret_cast<DwSmem*>(s_raw + 2 * (size_t)STAGE);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = blockIdx.x;
    // this CTA's share of the tiles (K split)
    const int per_split = (n_bt + gridDim.y - 1) / gridDim.y;
    const int bt0 = blockIdx.y * per_split;
    const int bt1 = min(n_bt, bt0 + per_split);
    const int nstages = 2 * max(0, bt1 - bt0);
    const unsigned tmem_cols = 2 * Opad <= 32 ? 32u : (2 * Opad <= 64 ? 64u : 128u);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm->full[i], 1);
            mbar_init(&sm->free_[i], 1);
            mbar_init(&sm->acc_free[i], 4);
        }
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = sm->tmem_base;
    const size_t a_part = (size_t)kMQ * kMQ;  // floats per part of a G^T image
    const size_t b_part = (size_t)kMQ * Opad;

    if (warp == 0) {
        const unsigned idesc = umma_idesc_tf32(kMQ, Opad);
        auto load = [&](int st) {
            const int buf = st & 1;
            if (lane == 0) {
                const size_t bt = (size_t)(bt0 + (st >> 1));
                const int h = st & 1;
                unsigned char* dst = s_raw + (size_t)buf * STAGE;
                const float* a = gimg_t + (bt * ncg + cg) * 2 * a_part + (size_t)h * (A_HALF / 4);
                const float* bsrc = goB + bt * 2 * b_part + (size_t)h * (B_HALF / 4);
                mbar_expect_tx(&sm->full[buf], 2u * A_HALF + 2u * (unsigned)B_HALF);
                bulk_copy_g2s(dst, a, A_HALF, &sm->full[buf]);
                bulk_copy_g2s(dst + A_HALF, a + a_part, A_HALF, &sm->full[buf]);
                bulk_copy_g2s(dst + 2 * A_HALF, bsrc, (unsigned)B_HALF, &sm->full[buf]);
                bulk_copy_g2s(dst + 2 * A_HALF + B_HALF, bsrc + b_part, (unsigned)B_HALF, &sm->full[buf]);
            }
        };
        if (nstages > 0) load(0);
        for (int st = 0; st < nstages; ++st) {
            const int buf = st & 1;
            const unsigned ph = (unsigned)(st >> 1) & 1u;
            if (st + 1 < nstages) {
                if (st >= 1) mbar_wait(&sm->free_[buf ^ 1], (unsigned)((st - 1) >> 1) & 1u);
                load(st + 1);
            }
            mbar_wait(&sm->full[buf], ph);
            if (st >= 2) mbar_wait(&sm->acc_free[buf], (unsigned)((st - 2) >> 1) & 1u);
            tc_fence_after();
            const unsigned acc = tmem + (unsigned)(buf * Opad);
            if (lane == 0) {
                const unsigned a_hi = smem_u32(s_raw + (size_t)buf * STAGE), a_lo = a_hi + A_HALF;
                const unsigned b_hi = a_hi + 2 * A_HALF, b_lo = b_hi + (unsigned)B_HALF;
#pragma unroll 1
                for (int kk = 0; kk < HQ / 8; ++kk) {
                    const unsigned ao = (unsigned)kk * 2u * A_LBO, bo = (unsigned)kk * 2u * B_LBO;
                    const unsigned long long dah = umma_desc(a_hi + ao, A_LBO, A_SBO), dal = umma_desc(a_lo + ao, A_LBO, A_SBO);
                    const unsigned long long dbh = umma_desc(b_hi + bo, B_LBO, B_SBO), dbl = umma_desc(b_lo + bo, B_LBO, B_SBO);
                    umma_tf32(acc, dal, dbh, idesc, kk != 0 ? 1u : 0u);
                    umma_tf32(acc, dah, dbl, idesc, 1u);
                    umma_tf32(acc, dah, dbh, idesc, 1u);
                }
                umma_commit(&sm->free_[buf]);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;
        float res[2][32];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int i = 0; i < 32; ++i) res[g][i] = 0.0f;
        for (int st = 0; st < nstages; ++st) {
            const int fb = st & 1;
            mbar_wait(&sm->free_[fb], (unsigned)(st >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (g * 32 < Opad) {
                    float v[32];
                    tmem_ld32(tmem + ((unsigned)(quarter * 32) << 16) + (unsigned)(fb * Opad + g * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) res[g][i] += v[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_(&sm->acc_free[fb]);
        }
        const int rr = quarter * 32 + lane;
        const int cell = cg * CPG + rr / C, c = rr % C;
        if (cell < ncells && nstages > 0) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (g * 32 + i < O) atomicAdd(dweight + ((size_t)(g * 32 + i) * C + c) * ncells + cell, res[g][i]);
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, tmem_cols);
    }
}

static size_t dg_smem_bytes(int C, int Opad)
{
    return (size_t)2 * kMQ * Opad * 4 + (size_t)2 * 2 * C * Opad * 4 + (size_t)4 * 32 * (C + 4) * 4 + sizeof(DgSmem) + 64;
}
static size_t dw_smem_bytes(int Opad)
{
    return (size_t)2 * (2 * kMQ * 64 * 4 + 2 * Opad * 64 * 4) + sizeof(DwSmem) + 64;
}

struct BwdPlan {
    int Opad, ncg;
    size_t wimg_t, go_a, go_b, big;  // floats
};
static BwdPlan bwd_plan(int O, int C, int ncells)
{
    BwdPlan p;
    p.Opad = (O + 15) / 16 * 16;
    p.ncg = (ncells + kMQ / C - 1) / (kMQ / C);
    p.wimg_t = (size_t)2 * p.Opad * C * ncells;
    p.go_a = p.go_b = (size_t)kBwdChunkTiles * 2 * kMQ * p.Opad;
    const size_t dg = (size_t)kBwdChunkTiles * kMQ * ncells * C;
    const size_t gt = (size_t)kBwdChunkTiles * p.ncg * 2 * kMQ * kMQ;
    p.big = dg > gt ? dg : gt;
    return p;
}

}  // namespace

bool convsp_wide_bwd_supported(int O, int C, int D)
{
    const int Opad = (O + 15) / 16 * 16;
    return D >= 1 && D <= 3 && (C == 32 || C == 64) && O >= 1 && Opad <= 64 && dg_smem_bytes(C, Opad) <= 225 * 1024 &&
           dw_smem_bytes(Opad) <= 225 * 1024;
}

size_t convsp_wide_bwd_workspace_bytes(int O, int C, int ncells)
{
    const BwdPlan p = bwd_plan(O, C, ncells);
    return sizeof(float) * (p.wimg_t + p.go_a + p.go_b + p.big);
}

// Returns the number of launches, -1 on failure.  dq/dl/dd/dw must be zero-filled by the caller where they are
// accumulated into (dl, dd, dw always; dq when it aliases dl).
int launch_convsp_wide_bwd(const float* qlocs, const float* locs, const float* data, const float* neighbors,
                           const float* weight, int B, int M, int N, int C, int D, int K, int O, int ncells, float radius,
                           const float* kernel_size, const float* dilation, int dis_norm, int kernel_fn, const float* go,
                           float* dq, float* dl, float* dd, float* dw, void* workspace, cudaStream_t stream)
{
    const BwdPlan p = bwd_plan(O, C, ncells);
    const int Opad = p.Opad;
    const SphParams sp = make_sph_params(kernel_fn, radius);
    float* wimg_t = (float*)workspace;
    float* go_a = wimg_t + p.wimg_t;
    float* go_b = go_a + p.go_a;
    float* big = go_b + p.go_b;
    const bool need_x = dq || dl || dd;
    const bool use_tc = K <= kGStage && ncells <= kMQ && getenv("SPNB_WIDE_TC") != nullptr;
    const int same = (dq != nullptr && dq == dl) ? 1 : 0;
    int launches = 0;
    if (need_x) {
        k_wide_prep_weights_t<<<148 * 4, 256, 0, stream>>>(weight, wimg_t, O, Opad, C, ncells);
        ++launches;
    }
    const size_t gsmem = C == 64 ? GatherLayout<3, 64>::bytes : GatherLayout<3, 32>::bytes;
    const size_t dgsmem = dg_smem_bytes(C, Opad), dwsmem = dw_smem_bytes(Opad);
    const int tiles_total = cdiv(M, kMQ);
    const int chunk_tiles = B > 1 ? kBwdChunkTiles / B : kBwdChunkTiles;
    if (chunk_tiles == 0) {
        set_error("spnb_convsp_backward_wide: batch size %d exceeds the tile chunk", B);
        return -1;
    }
    for (int t0 = 0; t0 < tiles_total; t0 += chunk_tiles) {
        const int nt = tiles_total - t0 < chunk_tiles ? tiles_total - t0 : chunk_tiles;
        const int q_first = t0 * kMQ;
        const dim3 ggrid(nt * (kMQ / kTQ), B), mgrid(nt, B);
        k_wide_go_images<<<dim3(148 * 2, B), 256, 0, stream>>>(go, q_first, M, O, Opad, nt, go_a, go_b);
        ++launches;
#define SETATTR(KERNEL, BYTES)                                                                                    \
    if (cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)) != cudaSuccess) { \
        set_error("spnb_convsp_backward_wide: shared memory not available");                                     \
        return -1;                                                                                                \
    }
#define LAUNCH(DD, CC)                                                                                            \
    do {                                                                                                          \
        if (need_x) {                                                                                             \
            SETATTR(k_wide_dg_gemm<CC>, dgsmem);                                                                  \
            SETATTR((k_wide_scatter<DD, CC>), gsmem);                                                             \
            k_wide_dg_gemm<CC><<<mgrid, kGemmThreads, dgsmem, stream>>>(go_a, wimg_t, Opad, ncells, big);         \
            k_wide_scatter<DD, CC><<<ggrid, kGThreads, gsmem, stream>>>(qlocs, locs, data, neighbors, q_first, M, N, K, \
                                                                       ncells, radius, kernel_size, dilation,     \
                                                                       dis_norm, sp, big, dq, dl, dd, same);      \
            launches += 2;                                                                                        \
        }                                                                                                         \
        if (dw) {                                                                                                 \
            SETATTR(k_wide_dw_gemm<CC>, dwsmem);                                                                  \
            if (use_tc) {                                                                                         \
                SETATTR((k_wide_gather_tc<DD, CC, true>), GatherTcLayout<CC>::bytes);                             \
                k_wide_gather_tc<DD, CC, true><<<ggrid, kTcThreads, GatherTcLayout<CC>::bytes, stream>>>(          \
                    qlocs, locs, data, neighbors, q_first, M, N, K, ncells, radius, kernel_size, dilation, dis_norm, \
                    sp, big);                                                                                     \
            } else {                                                                                              \
                SETATTR((k_wide_gather<DD, CC, true>), gsmem);                                                    \
                k_wide_gather<DD, CC, true><<<ggrid, kGThreads, gsmem, stream>>>(qlocs, locs, data, neighbors,    \
                                                                                q_first, M, N, K, ncells, radius, \
                                                                                kernel_size, dilation, dis_norm,  \
                                                                                sp, big);                         \
            }                                                                                                     \
            const int n_bt = nt * B;                                                                              \
            int ksplit = (2 * 148) / p.ncg;                                                                       \
            if (ksplit < 1) ksplit = 1;                                                                           \
            if (ksplit > n_bt) ksplit = n_bt;                                                                     \
            k_wide_dw_gemm<CC><<<dim3(p.ncg, ksplit), kGemmThreads, dwsmem, stream>>>(big, go_b, n_bt, O, Opad,   \
                                                                                     ncells, p.ncg, dw);          \
            launches += 2;                                                                                        \
        }                                                                                                         \
    } while (0)
        if (C == 64) {
            if (D == 1) LAUNCH(1, 64);
            else if (D == 2) LAUNCH(2, 64);
            else LAUNCH(3, 64);
        } else {
            if (D == 1) LAUNCH(1, 32);
            else if (D == 2) LAUNCH(2, 32);
            else LAUNCH(3, 32);
        }
#undef LAUNCH
#undef SETATTR
    }
    return launches;
}

}  // namespace spnb

using namespace spnb;

extern "C" {

size_t spnb_convsp_backward_wide_workspace_bytes(int nkernels, int nchannels, int ndims, int ncells)
{
    if (ncells < 1 || !convsp_wide_bwd_supported(nkernels, nchannels, ndims)) return 0;
    return convsp_wide_bwd_workspace_bytes(nkernels, nchannels, ncells);
}

int spnb_convsp_backward_wide(const float* qlocs, const float* locs, const float* data, const float* neighbors,
                              const float* weight, int B, int M, int N, int C, int D, int K, int O, int ncells,
                              float radius, const float* kernel_size, const float* dilation, int dis_norm,
                              int kernel_fn, const float* grad_out, float* dqlocs, float* dlocs, float* ddata,
                              float* dweight, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t need = spnb_convsp_backward_wide_workspace_bytes(O, C, D, ncells);
    if (need == 0) {
        set_error("spnb_convsp_backward_wide: unsupported shape (C=%d O=%d ndims=%d)", C, O, D);
        return 0;
    }
    if (!qlocs || !locs || !data || !neighbors || !weight || !kernel_size || !dilation || !grad_out || !workspace ||
        workspace_bytes < need || B <= 0 || M <= 0 || N <= 0 || K <= 0 || kernel_fn < 0 ||
        kernel_fn >= SPNB_NUM_KERNEL_FNS) {
        set_error("spnb_convsp_backward_wide: bad arguments (workspace %zu of %zu bytes)", workspace_bytes, need);
        return 0;
    }
    if (dqlocs != nullptr && dqlocs == dlocs && M != N) {
        set_error("spnb_convsp_backward_wide: dqlocs == dlocs requires M == N");
        return 0;
    }
    if (dlocs) cudaMemsetAsync(dlocs, 0, sizeof(float) * (size_t)B * N * D, stream);
    if (ddata) cudaMemsetAsync(ddata, 0, sizeof(float) * (size_t)B * N * C, stream);
    if (dweight) cudaMemsetAsync(dweight, 0, sizeof(float) * (size_t)O * C * ncells, stream);
    const int nl = launch_convsp_wide_bwd(qlocs, locs, data, neighbors, weight, B, M, N, C, D, K, O, ncells, radius,
                                          kernel_size, dilation, dis_norm, kernel_fn, grad_out, dqlocs, dlocs, ddata,
                                          dweight, workspace, stream);
    if (nl < 0) return 0;
    count_launches(nl);
    return check_launch("spnb_convsp_backward_wide") ? 1 : 0;
}

}  // extern "C"
