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
// here as tcgen05.mma with fp32 accumulators in tensor memory:
//
//  * one CTA owns 128 consecutive queries = the M dimension of a 128 x O x 8 UMMA (cta_group::1, kind::tf32);
//  * per kernel cell the 8 gather warps build G_cell[128 x C] on the CUDA cores (exact fp32 in-radius predicate and
//    the reference's float/double kernel evaluation; lanes = neighbours for the tests, lanes = channels for the
//    accumulation, sums in registers) and write it into shared memory as the A operand -- K-major core matrices
//    of 8 rows x 16 bytes, no swizzle -- SPLIT into two TF32 terms hi + lo (hi = x rounded to TF32, lo = x - hi);
//  * the weights of the cell, pre-arranged once per call by k_wide_prep_weights as the B operand's exact
//    shared-memory image (hi and lo), arrive by one TMA bulk copy each (cp.async.bulk -> mbarrier);
//  * an MMA warp (one elected lane) issues, per 8-channel K step, the three products hi*hi + lo*hi + hi*lo
//    ("3xTF32": the dropped lo*lo term is 2^-22 relative) accumulating into TMEM, and commits the batch to an
//    mbarrier that hands the A / B buffers back (tcgen05.commit); buffers are double-buffered, so the tensor
//    core works on cell k while the gather warps build cell k+1;
//  * the tensor core adds into its accumulator with truncation, which over the ~3000 accumulating MMAs of a
//    whole output would leave a bias of ~5e-5 of the result (measured).  The accumulator therefore holds ONE
//    cell's partial product only: two TMEM accumulators alternate, and while cell k+1 is multiplied the gather
//    warps pull cell k's 128 x O partial out with tcgen05.ld and add it, round-to-nearest, into fp32 registers;
//  * epilogue: registers + bias -> out (each thread owns 32 or 64 outputs of one query).
//
// With 128 queries per CTA the 2 MB of weights are streamed from L2 once per 128 queries (the CUDA-core kernel
// in convsp_wide.cu streams them once per 8).  Shapes outside C in {32, 64}, O <= 128, ndims <= 3 keep using
// convsp_wide.cu / convsp.cu.
#include "list_walk.cuh"
#include "spnb_common.cuh"

namespace spnb {

namespace {

constexpr int kMQ = 128;            // queries per CTA = UMMA M
constexpr int kGatherWarps = 8;
constexpr int kMmaThreads = kGatherWarps * 32 + 32;
constexpr int kStageStride = 68;    // floats per staged row (64 channels + 4: conflict-free transposed reads)

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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

struct MmaSmem {
    unsigned long long b_full[2], a_full[2], free_[2], acc_free[2];
    unsigned tmem_base;
};

template <int D, int C>
__global__ void __launch_bounds__(kMmaThreads, 1)
k_convsp_wide_mma_fwd(const float* __restrict__ qlocs, const float* __restrict__ locs, const float* __restrict__ data,
                      const float* __restrict__ neighbors, const float* __restrict__ wimg,
                      const float* __restrict__ bias, int M, int N, int K, int O, int Opad, int ncells, float radius,
                      const float* __restrict__ ksize, const float* __restrict__ dilation, int dis_norm, SphParams sp,
                      float* __restrict__ out)
{
    constexpr int A_BYTES = kMQ * C * 4;                 // one A operand (hi or lo)
    constexpr unsigned A_LBO = (kMQ / 8) * 128, A_SBO = 128;
    extern __shared__ __align__(1024) unsigned char s_raw[];
    // [A: buf][part] | [B: buf][part] | staging [warp][8][kStageStride] | barriers
    const int B_BYTES = Opad * C * 4;
    unsigned char* s_A = s_raw;
    unsigned char* s_B = s_raw + 4 * A_BYTES;
    float* s_stage = reinterpret_cast<float*>(s_B + 4 * B_BYTES);
    MmaSmem* sm = reinterpret_cast<MmaSmem*>(s_stage + kGatherWarps * 8 * kStageStride);
    const unsigned B_LBO = (unsigned)(Opad / 8) * 128, B_SBO = 128;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, m0 = blockIdx.x * kMQ;
    const unsigned tmem_cols = 2 * Opad <= 32 ? 32u : (2 * Opad <= 64 ? 64u : (2 * Opad <= 128 ? 128u : 256u));

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm->b_full[i], 1);
            mbar_init(&sm->a_full[i], kGatherWarps);
            mbar_init(&sm->free_[i], 1);
            mbar_init(&sm->acc_free[i], kGatherWarps);
        }
    }
    if (warp == kGatherWarps) tmem_alloc(&sm->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = sm->tmem_base;

    if (warp == kGatherWarps) {
        // ===================== MMA + weight-TMA warp =====================
        const unsigned idesc = umma_idesc_tf32(kMQ, Opad);
        const size_t img_stride = (size_t)2 * Opad * C;  // floats per cell (hi + lo)
        auto load_b = [&](int cell) {
            const int buf = cell & 1;
            if (lane == 0) {
                mbar_expect_tx(&sm->b_full[buf], 2u * (unsigned)B_BYTES);
                bulk_copy_g2s(s_B + (size_t)(2 * buf) * B_BYTES, wimg + (size_t)cell * img_stride, 2u * (unsigned)B_BYTES,
                              &sm->b_full[buf]);
            }
        };
        load_b(0);
        for (int cell = 0; cell < ncells; ++cell) {
            const int buf = cell & 1;
            const unsigned ph = (unsigned)(cell >> 1) & 1u;
            // weights of the next cell into the other buffer, once the MMAs of cell-1 have released it
            if (cell + 1 < ncells) {
                if (cell >= 1) mbar_wait(&sm->free_[buf ^ 1], (unsigned)((cell - 1) >> 1) & 1u);
                load_b(cell + 1);
            }
            mbar_wait(&sm->b_full[buf], ph);
            mbar_wait(&sm->a_full[buf], ph);
            // the accumulator of this parity was last used by cell-2: its partial has been pulled out
            if (cell >= 2) mbar_wait(&sm->acc_free[buf], (unsigned)((cell - 2) >> 1) & 1u);
            tc_fence_after();
            const unsigned acc = tmem + (unsigned)(buf * Opad);
            if (lane == 0) {
                const unsigned a_hi = smem_u32(s_A + (size_t)(2 * buf) * A_BYTES), a_lo = a_hi + A_BYTES;
                const unsigned b_hi = smem_u32(s_B + (size_t)(2 * buf) * B_BYTES), b_lo = b_hi + (unsigned)B_BYTES;
#pragma unroll 1
                for (int kk = 0; kk < C / 8; ++kk) {
                    const unsigned ao = (unsigned)kk * 2u * A_LBO, bo = (unsigned)kk * 2u * B_LBO;
                    const unsigned long long dah = umma_desc(a_hi + ao, A_LBO, A_SBO), dal = umma_desc(a_lo + ao, A_LBO, A_SBO);
                    const unsigned long long dbh = umma_desc(b_hi + bo, B_LBO, B_SBO), dbl = umma_desc(b_lo + bo, B_LBO, B_SBO);
                    umma_tf32(acc, dal, dbh, idesc, kk != 0 ? 1u : 0u);  // small terms first
                    umma_tf32(acc, dah, dbl, idesc, 1u);
                    umma_tf32(acc, dah, dbh, idesc, 1u);
                }
                // arrives when the MMAs of this cell are done: its A / B buffers may be overwritten and its
                // accumulator may be read
                umma_commit(&sm->free_[buf]);
            }
            __syncwarp();
        }
    } else {
        // ===================== gather warps: G_cell for 16 queries each =====================
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
        const float* sl = locs + (size_t)b * N * D;
        const float* sd = data + (size_t)b * N * C;
        float* stage = s_stage + warp * 8 * kStageStride;
        // this thread's share of the output: query row (TMEM lane) 32*(warp&3) + lane, columns of its column group
        constexpr int NCG = 4;  // up to 4 column groups of 32 per thread (Opad <= 128 with two warps per lane quarter: 2)
        float res[NCG / 2][32];
#pragma unroll
        for (int g = 0; g < NCG / 2; ++g)
#pragma unroll
            for (int i = 0; i < 32; ++i) res[g][i] = 0.0f;
        const int quarter = warp & 3;
        auto flush = [&](int cell) {
            const int fb = cell & 1;
            mbar_wait(&sm->free_[fb], (unsigned)(cell >> 1) & 1u);  // MMAs of `cell` complete
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < NCG / 2; ++g) {
                const int c0 = (warp >> 2) * 32 + g * 64;
                if (c0 < Opad) {
                    float v[32];
                    tmem_ld32(tmem + ((unsigned)(quarter * 32) << 16) + (unsigned)(fb * Opad + c0), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) res[g][i] += v[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_(&sm->acc_free[fb]);
        };
        for (int cell = 0; cell < ncells; ++cell) {
            const int buf = cell & 1;
            if (cell >= 2) mbar_wait(&sm->free_[buf], (unsigned)((cell - 2) >> 1) & 1u);
            float off[D];
            {
                int rem = cell;  // kernel cell index, dimension 0 fastest (common_funcs.h:494,575-580)
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int ik = rem % ks[k];
                    rem /= ks[k];
                    off[k] = (ik - half[k]) * dil[k];
                }
            }
            float* a_hi = reinterpret_cast<float*>(s_A + (size_t)(2 * buf) * A_BYTES);
            float* a_lo = reinterpret_cast<float*>(s_A + (size_t)(2 * buf + 1) * A_BYTES);
#pragma unroll 1
            for (int grp = 0; grp < 2; ++grp) {
                const int row0 = warp * 16 + grp * 8;  // first of 8 query rows of the 128-row tile
#pragma unroll 1
                for (int r = 0; r < 8; ++r) {
                    const int m = m0 + row0 + r;
                    float acc[C / 32];
#pragma unroll
                    for (int i = 0; i < C / 32; ++i) acc[i] = 0.0f;
                    if (m < M) {
                        const size_t q = (size_t)b * M + m;
                        float x[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) x[k] = qlocs[q * D + k];
                        const float* row = neighbors + q * K;
                        for (int base = 0; base < K; base += 32) {
                            // lanes = neighbours: list entry, culling, in-radius test and kernel value for this cell
                            const float nb = base + lane < K ? row[base + lane] : -1.0f;
                            const unsigned neg = __ballot_sync(0xffffffffu, !(nb >= 0.0f));
                            const int cnt = neg ? __ffs(neg) - 1 : 32;  // the list ends at its first negative entry
                            float s = 0.0f;
                            int j = 0;
                            bool hit = false;
                            if (lane < cnt) {
                                j = (int)nb;
                                float y[D];
                                float d0 = 0.0f;
#pragma unroll
                                for (int k = 0; k < D; ++k) {
                                    y[k] = sl[(size_t)j * D + k];
                                    d0 += (x[k] - y[k]) * (x[k] - y[k]);
                                }
                                if (!(d0 > cull2)) {
                                    float d = 0.0f;
#pragma unroll
                                    for (int k = 0; k < D; ++k) {
                                        const float t = x[k] + off[k] - y[k];
                                        d += t * t;
                                    }
                                    if (d < rad2) {
                                        d = sqrtf(d);
                                        float norm = 1.0f;
                                        if (dis_norm && d > 0.0f) norm /= d;
                                        s = (d > sp.H ? 0.0f : sph_eval(sp.w_expr, d, sp.H, sp.w_coef)) * norm;
                                        hit = true;
                                    }
                                }
                            }
                            // lanes = channels: accumulate the hits in list order
                            unsigned mask = __ballot_sync(0xffffffffu, hit);
                            while (mask) {
                                const int src = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const float sc = __shfl_sync(0xffffffffu, s, src);
                                const int jj = __shfl_sync(0xffffffffu, j, src);
                                const float* dj = sd + (size_t)jj * C;
#pragma unroll
                                for (int i = 0; i < C / 32; ++i) acc[i] = fmaf(sc, dj[lane + 32 * i], acc[i]);
                            }
                            if (cnt < 32) break;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < C / 32; ++i) stage[r * kStageStride + lane + 32 * i] = acc[i];
                }
                __syncwarp();
                // 8 rows x C channels -> core matrices (8 rows x 4 channels = 128 contiguous bytes), split hi + lo
#pragma unroll
                for (int ki = 0; ki < C / 4; ++ki) {
                    const float v = stage[(lane >> 2) * kStageStride + 4 * ki + (lane & 3)];
                    const float hi = to_tf32(v);
                    const int o = (ki * (kMQ / 8) + (row0 >> 3)) * 32 + lane;
                    a_hi[o] = hi;
                    a_lo[o] = to_tf32(v - hi);
                }
                __syncwarp();
            }
            fence_async_smem();  // generic-proxy writes of the A tile -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive_(&sm->a_full[buf]);
            if (cell >= 1) flush(cell - 1);  // the previous cell's partial product, while this one is multiplied
        }
        flush(ncells - 1);
        // ===================== epilogue: + bias -> out =====================
        const int m = m0 + quarter * 32 + lane;
        if (m < M) {
            float* orow = out + ((size_t)b * M + m) * O;
#pragma unroll
            for (int g = 0; g < NCG / 2; ++g) {
                const int c0 = (warp >> 2) * 32 + g * 64;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < O) orow[c0 + i] = res[g][i] + (bias ? bias[c0 + i] : 0.0f);
            }
        }
    }
    __syncthreads();
    if (warp == kGatherWarps) {
        tc_fence_after();
        tmem_dealloc(tmem, tmem_cols);
    }
}

static size_t mma_smem_bytes(int C, int Opad)
{
    return (size_t)4 * kMQ * C * 4 + (size_t)4 * Opad * C * 4 + sizeof(float) * kGatherWarps * 8 * kStageStride +
           sizeof(MmaSmem) + 64;
}

}  // namespace

bool convsp_wide_mma_supported(int O, int C, int D)
{
    const int Opad = (O + 15) / 16 * 16;
    return D >= 1 && D <= 3 && (C == 32 || C == 64) && O >= 1 && Opad <= 128 && mma_smem_bytes(C, Opad) <= 225 * 1024;
}

size_t convsp_wide_mma_workspace_bytes(int O, int C, int ncells)
{
    const int Opad = (O + 15) / 16 * 16;
    return sizeof(float) * 2 * (size_t)Opad * C * ncells;
}

// workspace: the weight images (convsp_wide_mma_workspace_bytes).  Returns the number of launches, -1 on failure.
int launch_convsp_wide_mma(const float* qlocs, const float* locs, const float* data, const float* neighbors,
                           const float* weight, const float* bias, int B, int M, int N, int C, int D, int K, int O,
                           int ncells, float radius, const float* kernel_size, const float* dilation, int dis_norm,
                           int kernel_fn, float* out, void* workspace, cudaStream_t stream)
{
    const int Opad = (O + 15) / 16 * 16;
    const SphParams sp = make_sph_params(kernel_fn, radius);
    float* img = (float*)workspace;
    k_wide_prep_weights<<<148 * 4, 256, 0, stream>>>(weight, img, O, Opad, C, ncells);
    const size_t smem = mma_smem_bytes(C, Opad);
    const dim3 grid(cdiv(M, kMQ), B);
#define LAUNCH(DD, CC)                                                                                            \
    do {                                                                                                          \
        if (cudaFuncSetAttribute(k_convsp_wide_mma_fwd<DD, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 (int)smem) != cudaSuccess) {                                                     \
            set_error("spnb_convsp_forward_wide: %zu bytes of shared memory not available", smem);               \
            return -1;                                                                                            \
        }                                                                                                         \
        k_convsp_wide_mma_fwd<DD, CC><<<grid, kMmaThreads, smem, stream>>>(                                       \
            qlocs, locs, data, neighbors, img, bias, M, N, K, O, Opad, ncells, radius, kernel_size, dilation,    \
            dis_norm, sp, out);                                                                                   \
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
    return 2;
}

}  // namespace spnb
