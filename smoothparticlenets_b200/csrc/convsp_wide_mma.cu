// ConvSP forward for WIDE channel counts on the 5th-generation tensor cores (BASELINE.json config 3: 64 -> 64
// channels, kernel_size 5, 1 M particles).
//
// The reference (compute_kernel_cells, src/common_funcs.h:512-572) spends O*C multiply-adds per (neighbour, kernel
// cell) pair.  Factored:
//
//     G[q, cell, c] = sum_{j in nbr(q), |q + off_cell - x_j| < r}  W(d) * norm * data[j, c]      (gather)
//     out[q, o]     = bias[o] + sum_{cell, c} weight[o, c, cell] * G[q, cell, c]                    (contraction)
//
// the contraction is a dense [queries x (ncells*C)] x [(ncells*C) x O] GEMM -- 1.02 MFLOP per query at c3 -- and runs
// here as tcgen05.mma with fp32 accumulators in tensor memory.  Two kernels per chunk of queries:
//
//  * k_wide_gather (CUDA cores): one warp per query walks the neighbour list; per neighbour the 32 lanes test 32
//    kernel cells at once (exact fp32 in-radius predicate, the reference's float/double kernel evaluation), the hits
//    are enumerated with ballot/ffs and the lanes add W*norm*data[j, c] for their channels into the query's G row
//    in shared memory.  The finished G slab is written to global memory ALREADY as the A operand of the GEMM:
//    per 128-query tile and kernel cell the exact shared-memory image the tensor core reads -- K-major core
//    matrices of 8 rows x 16 bytes, no swizzle -- and split into two TF32 terms hi + lo (hi = x rounded to TF32,
//    lo = x - hi);
//  * k_wide_gemm (tcgen05): one CTA per 128-query tile = the M dimension of a 128 x O x 8 UMMA (cta_group::1,
//    kind::tf32).  Warp 0 streams, per kernel cell, the A image and the B image (the weights, pre-arranged once
//    per call by k_wide_prep_weights) into a 2-stage shared-memory ring with TMA bulk copies (cp.async.bulk ->
//    mbarrier) and issues, per 8-channel K step, the three products hi*hi + lo*hi + hi*lo ("3xTF32": the dropped
//    lo*lo term is 2^-22 relative), committing each cell to an mbarrier (tcgen05.commit) that frees the stage;
//  * the tensor core adds into its accumulator with truncation, which over the ~3000 accumulating MMAs of a
//    whole output would leave a bias of ~5e-5 of the result (measured).  The accumulator therefore holds ONE
//    cell's partial product only: two TMEM accumulators alternate, and while cell k+1 is multiplied four
//    epilogue warps pull cell k's 128 x O partial out with tcgen05.ld and add it, round-to-nearest, into fp32
//    registers; at the end registers + bias -> out.
//
// (A first version built the A operand inside the GEMM kernel, cell by cell; with 128 queries per CTA that forces
// the gather to revisit every neighbour list once per kernel cell and ran 2.5x slower than the CUDA-core kernel --
// profiles/README.md.)  Shapes outside C in {32, 64}, O <= 128, ndims <= 3 keep using convsp_wide.cu / convsp.cu.
#include "wide_mma.cuh"

namespace spnb {

using namespace wide;

namespace {

// ---- weights: B operand images ---------------------------------------------------------------------------
// img[cell][part][ki][ni][8][4]: part 0 = hi, 1 = lo; element (o = 8 ni + r, c = 4 ki + e) of weight[o][c][cell];
// rows o >= O are zero.  One image (Opad * C floats) is exactly what the MMA reads from shared memory.
__global__ void __launch_bounds__(256)
k_wide_prep_weights(const float* __restrict__ w, float* __restrict__ img, int O, int Opad, int C, int ncells)
{
    const long long per = (long long)Opad * C;
    const long long n = per * ncells;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int cell = (int)(i / per);
        const int t = (int)(i % per);
        const int e = t & 3, r = (t >> 2) & 7, rest = t >> 5;
        const int ni = rest % (Opad / 8), ki = rest / (Opad / 8);
        const int o = 8 * ni + r, c = 4 * ki + e;
        const float v = o < O ? w[((size_t)o * C + c) * ncells + cell] : 0.0f;
        const float hi = to_tf32(v);
        img[((size_t)cell * 2) * per + t] = hi;
        img[((size_t)cell * 2 + 1) * per + t] = to_tf32(v - hi);
    }
}

// ---- GEMM: out[128 x O] = sum over cells of A_cell[128 x C] * B_cell[O x C]^T ------------------------------------
struct GemmSmem {
    unsigned long long full[2], free_[2], acc_free[2];
    unsigned tmem_base;
};

template <int C>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_wide_gemm(const float* __restrict__ gimg, const float* __restrict__ wimg, const float* __restrict__ bias,
            int q_first, int M, int O, int Opad, int ncells, float* __restrict__ out)
{
    constexpr int A_BYTES = kMQ * C * 4;  // one A operand (hi or lo)
    constexpr unsigned A_LBO = (kMQ / 8) * 128, A_SBO = 128;
    extern __shared__ __align__(1024) unsigned char s_raw[];
    const int B_BYTES = Opad * C * 4;
    const int STAGE = 2 * A_BYTES + 2 * B_BYTES;  // [A hi | A lo | B hi | B lo]
    GemmSmem* sm = reinterpret_cast<GemmSmem*>(s_raw + 2 * (size_t)STAGE);
    const unsigned B_LBO = (unsigned)(Opad / 8) * 128, B_SBO = 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, b = blockIdx.y;
    const unsigned tmem_cols = 2 * Opad <= 32 ? 32u : (2 * Opad <= 64 ? 64u : (2 * Opad <= 128 ? 128u : 256u));

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

    if (warp == 0) {
        const unsigned idesc = umma_idesc_tf32(kMQ, Opad);
        const float* atile = gimg + ((size_t)b * gridDim.x + tile) * ncells * (size_t)(2 * kMQ * C);
        auto load = [&](int cell) {
            const int buf = cell & 1;
            if (lane == 0) {
                unsigned char* st = s_raw + (size_t)buf * STAGE;
                mbar_expect_tx(&sm->full[buf], 2u * A_BYTES + 2u * (unsigned)B_BYTES);
                bulk_copy_g2s(st, atile + (size_t)cell * (2 * kMQ * C), 2u * A_BYTES, &sm->full[buf]);
                bulk_copy_g2s(st + 2 * A_BYTES, wimg + (size_t)cell * (2 * (size_t)Opad * C), 2u * (unsigned)B_BYTES,
                              &sm->full[buf]);
            }
        };
        load(0);
        for (int cell = 0; cell < ncells; ++cell) {
            const int buf = cell & 1;
            const unsigned ph = (unsigned)(cell >> 1) & 1u;
            if (cell + 1 < ncells) {
                // the other stage is free once the MMAs of cell-1 are done
                if (cell >= 1) mbar_wait(&sm->free_[buf ^ 1], (unsigned)((cell - 1) >> 1) & 1u);
                load(cell + 1);
            }
            mbar_wait(&sm->full[buf], ph);
            // the accumulator of this parity was last used by cell-2: its partial has been pulled out
            if (cell >= 2) mbar_wait(&sm->acc_free[buf], (unsigned)((cell - 2) >> 1) & 1u);
            tc_fence_after();
            const unsigned acc = tmem + (unsigned)(buf * Opad);
            if (lane == 0) {
                const unsigned a_hi = smem_u32(s_raw + (size_t)buf * STAGE), a_lo = a_hi + A_BYTES;
                const unsigned b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + (unsigned)B_BYTES;
#pragma unroll 1
                for (int kk = 0; kk < C / 8; ++kk) {
                    const unsigned ao = (unsigned)kk * 2u * A_LBO, bo = (unsigned)kk * 2u * B_LBO;
                    const unsigned long long dah = umma_desc(a_hi + ao, A_LBO, A_SBO), dal = umma_desc(a_lo + ao, A_LBO, A_SBO);
                    const unsigned long long dbh = umma_desc(b_hi + bo, B_LBO, B_SBO), dbl = umma_desc(b_lo + bo, B_LBO, B_SBO);
                    umma_tf32(acc, dal, dbh, idesc, kk != 0 ? 1u : 0u);  // small terms first
                    umma_tf32(acc, dah, dbl, idesc, 1u);
                    umma_tf32(acc, dah, dbh, idesc, 1u);
                }
                // arrives when the MMAs of this cell are done: its stage may be overwritten, its accumulator read
                umma_commit(&sm->free_[buf]);
            }
            __syncwarp();
        }
    } else {
        // ===================== accumulator flush + epilogue (warps 1..4 = TMEM lane quarters 1, 2, 3, 0) ==========
        const int quarter = warp & 3;
        float res[4][32];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < 32; ++i) res[g][i] = 0.0f;
        for (int cell = 0; cell < ncells; ++cell) {
            const int fb = cell & 1;
            mbar_wait(&sm->free_[fb], (unsigned)(cell >> 1) & 1u);  // MMAs of `cell` complete
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
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
        const int m = q_first + tile * kMQ + quarter * 32 + lane;
        if (m < M) {
            float* orow = out + ((size_t)b * M + m) * O;
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (g * 32 + i < O) orow[g * 32 + i] = res[g][i] + (bias ? bias[g * 32 + i] : 0.0f);
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, tmem_cols);
    }
}

constexpr int kChunkTiles = 296;  // 128-query tiles per pass (two per SM): bounds the image buffer

static size_t gemm_smem_bytes(int C, int Opad)
{
    return (size_t)2 * (2 * kMQ * C * 4 + 2 * Opad * C * 4) + sizeof(GemmSmem) + 64;
}

}  // namespace

bool convsp_wide_mma_supported(int O, int C, int D)
{
    const int Opad = (O + 15) / 16 * 16;
    return D >= 1 && D <= 3 && (C == 32 || C == 64) && O >= 1 && Opad <= 128 && gemm_smem_bytes(C, Opad) <= 225 * 1024;
}

// weight images + the G images of one chunk of query tiles
size_t convsp_wide_mma_workspace_bytes(int O, int C, int ncells)
{
    const int Opad = (O + 15) / 16 * 16;
    return sizeof(float) * 2 * (size_t)Opad * C * ncells + sizeof(float) * 2 * (size_t)kMQ * C * ncells * kChunkTiles;
}

// Returns the number of launches, -1 on failure.
int launch_convsp_wide_mma(const float* qlocs, const float* locs, const float* data, const float* neighbors,
                           const float* weight, const float* bias, int B, int M, int N, int C, int D, int K, int O,
                           int ncells, float radius, const float* kernel_size, const float* dilation, int dis_norm,
                           int kernel_fn, float* out, void* workspace, cudaStream_t stream)
{
    const int Opad = (O + 15) / 16 * 16;
    const SphParams sp = make_sph_params(kernel_fn, radius);
    float* wimg = (float*)workspace;
    float* gimg = wimg + 2 * (size_t)Opad * C * ncells;
    k_wide_prep_weights<<<148 * 4, 256, 0, stream>>>(weight, wimg, O, Opad, C, ncells);
    int launches = 1;
    // measurement knob (bench.py): time the contraction alone, on whatever the image buffer holds
    const bool gemm_only = getenv("SPNB_WIDE_GEMM_ONLY") != nullptr;
    // opt-in (environment SPNB_WIDE_TC): the gather as per-query tcgen05 GEMMs (k_wide_gather_tc), possible when a list
    // fits one staging round and the kernel cells fit the 128 rows of an MMA.  Parity-tested, but measured 30 % slower
    // than the CUDA-core gather on B200 (profiles/README.md), which therefore stays the default
    const bool use_tc = K <= kGStage && ncells <= kMQ && getenv("SPNB_WIDE_TC") != nullptr;
    const size_t gsmem = C == 64 ? GatherLayout<3, 64>::bytes : GatherLayout<3, 32>::bytes;
    const size_t msmem = gemm_smem_bytes(C, Opad);
    // scenes are processed one after the other when B * tiles exceeds the chunk (the image buffer is per chunk)
    const int tiles_total = cdiv(M, kMQ);
    const int chunk_tiles = B > 1 ? (kChunkTiles / B > 0 ? kChunkTiles / B : 0) : kChunkTiles;
    if (chunk_tiles == 0) {
        set_error("spnb_convsp_forward_wide: batch size %d exceeds the tile chunk", B);
        return -1;
    }
    for (int t0 = 0; t0 < tiles_total; t0 += chunk_tiles) {
        const int nt = tiles_total - t0 < chunk_tiles ? tiles_total - t0 : chunk_tiles;
        const int q_first = t0 * kMQ;
        const dim3 ggrid(nt * (kMQ / kTQ), B), mgrid(nt, B);
#define LAUNCH(DD, CC)                                                                                            \
    do {                                                                                                          \
        if (cudaFuncSetAttribute(k_wide_gather<DD, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem) != \
                cudaSuccess ||                                                                                    \
            cudaFuncSetAttribute(k_wide_gemm<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem) !=     \
                cudaSuccess) {                                                                                    \
            set_error("spnb_convsp_forward_wide: shared memory not available");                                  \
            return -1;                                                                                            \
        }                                                                                                         \
        if (!gemm_only && use_tc) {                                                                               \
            if (cudaFuncSetAttribute((k_wide_gather_tc<DD, CC, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)GatherTcLayout<CC>::bytes) != cudaSuccess) {                            \
                set_error("spnb_convsp_forward_wide: shared memory not available");                              \
                return -1;                                                                                        \
            }                                                                                                     \
            k_wide_gather_tc<DD, CC, false><<<ggrid, kTcThreads, GatherTcLayout<CC>::bytes, stream>>>(             \
                qlocs, locs, data, neighbors, q_first, M, N, K, ncells, radius, kernel_size, dilation, dis_norm, sp, \
                gimg);                                                                                            \
        } else if (!gemm_only) {                                                                                  \
            k_wide_gather<DD, CC><<<ggrid, kGThreads, gsmem, stream>>>(qlocs, locs, data, neighbors, q_first, M, N, K, \
                                                                      ncells, radius, kernel_size, dilation,      \
                                                                      dis_norm, sp, gimg);                        \
        }                                                                                                         \
        k_wide_gemm<CC><<<mgrid, kGemmThreads, msmem, stream>>>(gimg, wimg, bias, q_first, M, O, Opad, ncells, out); \
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
        launches += 2;
    }
    return launches;
}

}  // namespace spnb
