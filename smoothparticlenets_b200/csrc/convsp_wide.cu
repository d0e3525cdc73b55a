// ConvSP forward for WIDE channel counts (BASELINE.json config 3: 64 -> 64 channels, kernel_size 5).
//
// Replaces kernel_convsp (reference src/gpu_kernels.cu:50-126) for these shapes.  The reference (and
// the generic kernel in convsp.cu) spend O*C multiply-adds per (neighbour, kernel cell) pair:
// ~135 k MACs per neighbour at 64x64x33 in-radius cells.  Here the sum is FACTORED:
//
//     G[q, cell, c] = sum_{j in nbr(q), |q + off_cell - x_j| < r}  W(d) * norm * data[j, c]
//     out[q, o]     = bias[o] + sum_{cell, c} weight[o, c, cell] * G[q, cell, c]
//
// so a neighbour costs C MACs per in-radius cell and the weights enter once per query, as a dense
// [queries x (ncells*C)] x [(ncells*C) x O] contraction (1.02 MFLOP per query at c3).
//
// One CTA owns kTQ = 8 queries (one warp each).  The kernel cells are processed in slabs whose G
// tile fits in shared memory twice per SM (8 queries x 42 cells x 64 channels x 4 B = 86 KB at c3):
//   phase 1 (gather): each warp walks its query's neighbour list; for a neighbour the 32 lanes test
//     32 kernel cells at once (exact fp32 predicate, the reference's float/double W evaluation), the
//     hits are enumerated with ballot/ffs, and for each hit cell the lanes add W*norm*data[j, c] for
//     their channels c = lane, lane+32, ... into the query's G row (no conflicts, no atomics);
//   phase 2 (contraction): the CTA multiplies its [8 x slab*C] G tile with the matching rows of the
//     TRANSPOSED weights Wt[cell][c][o] (a pre-pass, so that the o-values of one (cell, c) are
//     contiguous) on the CUDA cores, 2 outputs per thread, 4 k-values per step.
// The contraction is the part that belongs on the tensor cores (tcgen05, 3xTF32 to hold 1e-5); that
// swap is the next step for this kernel -- the gather phase and the weight streaming are written so
// that it can be replaced in place (G tile and Wt rows are already K-contiguous in shared/global).
#include <stdlib.h>

#include "spnb_common.cuh"

namespace spnb {

// convsp_wide_mma.cu: the tcgen05 version (C in {32, 64}, O <= 128)
bool convsp_wide_mma_supported(int O, int C, int D);
size_t convsp_wide_mma_workspace_bytes(int O, int C, int ncells);
int launch_convsp_wide_mma(const float* qlocs, const float* locs, const float* data, const float* neighbors,
                           const float* weight, const float* bias, int B, int M, int N, int C, int D, int K, int O,
                           int ncells, float radius, const float* kernel_size, const float* dilation, int dis_norm,
                           int kernel_fn, float* out, void* workspace, cudaStream_t stream);

namespace {

constexpr int kTQ = 8;         // queries per CTA (one warp each)
constexpr int kWThreads = 256;
constexpr int kMaxGBytes = 108 * 1024;  // G tile per CTA: two CTAs per SM (227 KB of shared memory)

// Wt[cell][c][o] = weight[o][c][cell]
__global__ void __launch_bounds__(256)
k_wide_transpose(const float* __restrict__ w, float* __restrict__ wt, int O, int C, int ncells)
{
    const long long n = (long long)O * C * ncells;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i % O);
        const long long r = i / O;
        const int c = (int)(r % C);
        const int cell = (int)(r / C);
        wt[i] = w[((size_t)o * C + c) * ncells + cell];
    }
}

template <int D>
__global__ void __launch_bounds__(kWThreads)
k_convsp_wide_fwd(const float* __restrict__ qlocs, const float* __restrict__ locs,
                  const float* __restrict__ data, const float* __restrict__ neighbors,
                  const float* __restrict__ wt, const float* __restrict__ bias, int M, int N, int C, int K,
                  int O, int ncells, int slab_cells, float radius, const float* __restrict__ ksize,
                  const float* __restrict__ dilation, int dis_norm, SphParams sp, float* __restrict__ out)
{
    extern __shared__ __align__(16) float s_G[];  // [kTQ][slab_cells*C] then [slab_cells][D] cell offsets
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int m = blockIdx.x * kTQ + warp;
    const bool active = m < M;
    const size_t q = (size_t)b * M + (active ? m : 0);
    const int SK = slab_cells * C;  // K-columns per slab
    float* Gq = s_G + (size_t)warp * SK;
    float* s_off = s_G + (size_t)kTQ * SK;  // (idx_k - ks_k/2) * dilation_k of every cell of the slab

    // kernel shape and the cull radius (common_funcs.h:481-485)
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

    // phase-2 ownership: thread -> output column o_t (per 64-wide tile) and queries qh, qh+4
    const int o_t = threadIdx.x & 63, qh = threadIdx.x >> 6;
    const int n_otiles = (O + 63) / 64;
    // accumulators for up to 4 o-tiles (O <= 256); wider O loops the whole kernel body per tile group
    float acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0f;

    for (int cell0 = 0; cell0 < ncells; cell0 += slab_cells) {
        const int ncs = min(slab_cells, ncells - cell0);
        // ---- phase 1: G tile of this slab
        for (int i = lane; i < SK; i += 32) Gq[i] = 0.0f;
        for (int cl = threadIdx.x; cl < ncs; cl += kWThreads) {
            int rem = cell0 + cl;  // kernel cell index, dimension 0 fastest (common_funcs.h:494,575-580)
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int ik = rem % ks[k];
                rem /= ks[k];
                s_off[cl * D + k] = (ik - half[k]) * dil[k];
            }
        }
        __syncthreads();
        if (active) {
            for (int jj = 0; jj < K; ++jj) {
                const float nb = row[jj];
                if (!(nb >= 0.0f)) break;  // warp-uniform: every lane reads the same entry
                const int j = (int)nb;
                float y[D];
                float d0 = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    y[k] = sl[(size_t)j * D + k];
                    d0 += (x[k] - y[k]) * (x[k] - y[k]);
                }
                if (d0 > cull2) continue;
                const float* dj = sd + (size_t)j * C;
                float djr[4];  // this lane's channels lane, lane+32, lane+64, lane+96 of the neighbour
#pragma unroll
                for (int i = 0; i < 4; ++i) djr[i] = lane + 32 * i < C ? dj[lane + 32 * i] : 0.0f;
                for (int r0 = 0; r0 < ncs; r0 += 32) {
                    const int cl = r0 + lane;  // cell within the slab
                    float s = 0.0f;
                    bool hit = false;
                    if (cl < ncs) {
                        float d = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float t = x[k] + s_off[cl * D + k] - y[k];
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
                    unsigned mask = __ballot_sync(0xffffffffu, hit);
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float sc = __shfl_sync(0xffffffffu, s, src);
                        float* g = Gq + (size_t)(r0 + src) * C;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (lane + 32 * i < C) g[lane + 32 * i] = fmaf(sc, djr[i], g[lane + 32 * i]);
                        for (int c = lane + 128; c < C; c += 32) g[c] = fmaf(sc, dj[c], g[c]);
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase 2: acc[q][o] += sum_k G[q][k] * Wt[(cell0*C + k)][o]
        const float* wslab = wt + (size_t)cell0 * C * O;
        const int nk = ncs * C;
        const float* G0 = s_G + (size_t)qh * SK;
        const float* G1 = s_G + (size_t)(qh + 4) * SK;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (t < n_otiles) {
                const int o = t * 64 + o_t;
                if (o < O) {
                    float a0 = acc[t][0], a1 = acc[t][1];
                    const float* wp = wslab + o;
                    int k = 0;
                    for (; k + 4 <= nk; k += 4, wp += 4 * O) {
                        const float4 g0 = *reinterpret_cast<const float4*>(G0 + k);
                        const float4 g1 = *reinterpret_cast<const float4*>(G1 + k);
                        const float w0 = wp[0], w1 = wp[O], w2 = wp[2 * O], w3 = wp[3 * O];
                        a0 = fmaf(g0.x, w0, a0); a1 = fmaf(g1.x, w0, a1);
                        a0 = fmaf(g0.y, w1, a0); a1 = fmaf(g1.y, w1, a1);
                        a0 = fmaf(g0.z, w2, a0); a1 = fmaf(g1.z, w2, a1);
                        a0 = fmaf(g0.w, w3, a0); a1 = fmaf(g1.w, w3, a1);
                    }
                    for (; k < nk; ++k, wp += O) {
                        const float w0 = wp[0];
                        a0 = fmaf(G0[k], w0, a0);
                        a1 = fmaf(G1[k], w0, a1);
                    }
                    acc[t][0] = a0;
                    acc[t][1] = a1;
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (t < n_otiles) {
            const int o = t * 64 + o_t;
            if (o < O) {
                const float bo = bias ? bias[o] : 0.0f;
                const int m0 = blockIdx.x * kTQ + qh, m1 = m0 + 4;
                if (m0 < M) out[((size_t)b * M + m0) * O + o] = acc[t][0] + bo;
                if (m1 < M) out[((size_t)b * M + m1) * O + o] = acc[t][1] + bo;
            }
        }
    }
}

}  // namespace
}  // namespace spnb

using namespace spnb;

extern "C" {

size_t spnb_convsp_forward_wide_workspace_bytes(int nkernels, int nchannels, int ndims, int ncells)
{
    // supported: ndims <= 3, 32 <= C (so that the lanes of a warp have channels to own), O <= 256,
    // a G tile of at least one cell per slab
    if (ndims < 1 || ndims > 3 || nchannels < 32 || nkernels < 1 || nkernels > 256 || ncells < 1) return 0;
    if ((size_t)kTQ * nchannels * sizeof(float) > (size_t)kMaxGBytes || (nchannels & 3)) return 0;
    if (convsp_wide_mma_supported(nkernels, nchannels, ndims) && !getenv("SPNB_WIDE_NO_MMA"))
        return convsp_wide_mma_workspace_bytes(nkernels, nchannels, ncells);
    return sizeof(float) * (size_t)nkernels * nchannels * ncells;
}

int spnb_convsp_forward_wide(const float* qlocs, const float* locs, const float* data,
                             const float* neighbors, const float* weight, const float* bias, int B, int M,
                             int N, int C, int D, int K, int O, int ncells, float radius,
                             const float* kernel_size, const float* dilation, int dis_norm, int kernel_fn,
                             float* out, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t need = spnb_convsp_forward_wide_workspace_bytes(O, C, D, ncells);
    if (need == 0) {
        set_error("spnb_convsp_forward_wide: unsupported shape (C=%d O=%d ndims=%d)", C, O, D);
        return 0;
    }
    if (!qlocs || !locs || !data || !neighbors || !weight || !kernel_size || !dilation || !out ||
        !workspace || workspace_bytes < need || B <= 0 || M <= 0 || N <= 0 || K <= 0 ||
        kernel_fn < 0 || kernel_fn >= SPNB_NUM_KERNEL_FNS) {
        set_error("spnb_convsp_forward_wide: bad arguments (workspace %zu of %zu bytes)", workspace_bytes, need);
        return 0;
    }
    if (convsp_wide_mma_supported(O, C, D) && !getenv("SPNB_WIDE_NO_MMA")) {
        // tensor-core contraction (tcgen05, 3xTF32, accumulators in tensor memory): convsp_wide_mma.cu
        const int nl = launch_convsp_wide_mma(qlocs, locs, data, neighbors, weight, bias, B, M, N, C, D, K, O, ncells,
                                              radius, kernel_size, dilation, dis_norm, kernel_fn, out, workspace, stream);
        if (nl < 0) return 0;
        count_launches(nl);
        return check_launch("spnb_convsp_forward_wide") ? 1 : 0;
    }
    const SphParams sp = make_sph_params(kernel_fn, radius);
    float* wt = (float*)workspace;
    k_wide_transpose<<<148 * 4, 256, 0, stream>>>(weight, wt, O, C, ncells);
    int slab_cells = kMaxGBytes / (kTQ * C * (int)sizeof(float));
    if (slab_cells > ncells) slab_cells = ncells;
    // balance the slabs (e.g. 125 cells -> 42, 42, 41 rather than 52, 52, 21)
    const int nslabs = cdiv(ncells, slab_cells);
    slab_cells = cdiv(ncells, nslabs);
    const size_t smem = sizeof(float) * ((size_t)kTQ * slab_cells * C + (size_t)slab_cells * D);
    const dim3 grid(cdiv(M, kTQ), B);
#define LAUNCH(DD)                                                                                 \
    do {                                                                                           \
        cudaFuncSetAttribute(k_convsp_wide_fwd<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)smem);                                                           \
        k_convsp_wide_fwd<DD><<<grid, kWThreads, smem, stream>>>(                                  \
            qlocs, locs, data, neighbors, wt, bias, M, N, C, K, O, ncells, slab_cells, radius,     \
            kernel_size, dilation, dis_norm, sp, out);                                             \
    } while (0)
    if (D == 1) LAUNCH(1);
    else if (D == 2) LAUNCH(2);
    else LAUNCH(3);
#undef LAUNCH
    count_launches(2);
    return check_launch("spnb_convsp_forward_wide") ? 1 : 0;
}

}  // extern "C"
