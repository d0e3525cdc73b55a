// ConvSP forward / backward for sm_100a.
//
// Replaces kernel_convsp + cuda_convsp (reference src/gpu_kernels.cu:50-126) and the per-query math
// of compute_kernel_cells (src/common_funcs.h:439-583).
//
// Work decomposition: a GROUP of G lanes (G = 8) cooperates on one query; lane `sub` takes the
// neighbour-list entries sub, sub+G, ... so that the G lanes read one 32-byte sector of the list per
// step, the four groups of a warp read four rows, and the per-query sums are finished with log2(G)
// shuffles.  The list is consumed up to the first negative entry, as the reference does.
//
//  * convsp_small.cu: ncells == 1 (kernel_size all 1) with compile-time channel counts -- the
//    fluid-simulation layers (examples/fluid_sim.py:156-175).  Its backward has an atomics-free
//    mode: when the neighbour relation is symmetric (qlocs is locs, lists not truncated),
//    d(loss)/d(locs[j]) and d(loss)/d(data[j]) are GATHERED over j's own list instead of scattered
//    with atomics.
//  * this file: k_convsp_fwd_generic / k_convsp_bwd_generic for any ndims / channels /
//    kernel_size, and the C ABI.  Here per-term arithmetic keeps the reference's association
//    ((w*data)*W)*norm and its float/double promotions, so individual terms are bit-identical to
//    the CPU reference and only the summation order differs (-fmad=false; see spnb_common.cuh).
#include "convsp_small.cuh"
#include "spnb_common.cuh"

namespace spnb {

constexpr int kG = 8;  // lanes per query
constexpr int kThreads = 256;

// First negative entry among the G lanes of my group (G if none) from a warp ballot of "negative".
__device__ __forceinline__ int group_first_neg(unsigned neg, int lane, int sub)
{
    const unsigned g = (neg >> (lane - sub)) & ((1u << kG) - 1u);
    return g ? __ffs(g) - 1 : kG;
}

template <typename T>
__device__ __forceinline__ T group_sum(T v)
{
#pragma unroll
    for (int o = kG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// =================================================================================================
// generic path: any ndims / channels / kernel_size
// =================================================================================================
struct KernelShape {
    float dil[SPNB_MAXD];
    int ks[SPNB_MAXD];
    int half[SPNB_MAXD];
};

// Reads kernel_size / dilation (device float arrays) and the cull radius of common_funcs.h:481-485.
template <int DT>
__device__ __forceinline__ void load_shape(const float* __restrict__ ksize,
                                           const float* __restrict__ dil, int D, float radius,
                                           int* ks, int* half, float* dl, float& cull2)
{
    float maxdil = dil[0], maxks = ksize[0];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        ks[k] = (int)ksize[k];
        half[k] = ((int)ksize[k]) / 2;
        dl[k] = dil[k];
        if (dil[k] > maxdil) maxdil = dil[k];
        if (ksize[k] > maxks) maxks = ksize[k];
    }
    const float nr = radius + ((int)maxks / 2) * maxdil * fast_root_dim(D);
    cull2 = nr * nr;
}

constexpr int kOChunk = 8;

// weights are staged in shared memory when they fit (w_in_smem), else read through L1/L2.
template <int DT>
__global__ void __launch_bounds__(kThreads)
k_convsp_fwd_generic(const float* __restrict__ qlocs, const float* __restrict__ locs,
                     const float* __restrict__ data, const float* __restrict__ neighbors,
                     const float* __restrict__ weight, const float* __restrict__ bias, long long BM,
                     int M, int N, int C, int ndims, int K, int O, int ncells, float radius,
                     const float* __restrict__ ksize, const float* __restrict__ dilation,
                     int dis_norm, SphParams sp, float* __restrict__ out, int w_in_smem)
{
    constexpr int MD = DT > 0 ? DT : SPNB_MAXD;
    const int D = DT > 0 ? DT : ndims;
    extern __shared__ float s_w[];
    if (w_in_smem) {
        for (int i = threadIdx.x; i < O * C * ncells; i += kThreads) s_w[i] = weight[i];
        __syncthreads();
    }
    const float* wp = w_in_smem ? s_w : weight;
    const int lane = threadIdx.x & 31, sub = lane & (kG - 1);
    const long long q = ((long long)blockIdx.x * kThreads + threadIdx.x) / kG;
    const bool active = q < BM;
    const long long qq = active ? q : 0;
    const int b = (int)(qq / M);
    int ks[MD], half[MD];
    float dl[MD], x[MD], cull2;
    load_shape<DT>(ksize, dilation, D, radius, ks, half, dl, cull2);
    const float rad2 = radius * radius;
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[qq * D + k];
    const float* row = neighbors + qq * K;
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;

    for (int o0 = 0; o0 < O; o0 += kOChunk) {
        const int on = min(kOChunk, O - o0);
        float acc[kOChunk];
#pragma unroll
        for (int o = 0; o < kOChunk; ++o) acc[o] = 0.0f;
        bool done = !active;  // my group's row has ended (the list stops at its first negative entry)
        for (int jj0 = 0; jj0 < K; jj0 += kG) {
            const int jj = jj0 + sub;
            const float nb = (!done && jj < K) ? row[jj] : -1.0f;
            const unsigned neg = __ballot_sync(0xffffffffu, !(nb >= 0.0f));
            const int fneg = group_first_neg(neg, lane, sub);
            if (fneg < kG) done = true;
            if (sub < fneg) {
                const int j = (int)nb;
                const float* y = sl + (size_t)j * D;
                const float* dj = sd + (size_t)j * C;
                float yy[MD];
                float d = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    yy[k] = y[k];
                    d += (x[k] - yy[k]) * (x[k] - yy[k]);
                }
                if (!(d > cull2)) {
                    int kidx[MD];
#pragma unroll
                    for (int k = 0; k < D; ++k) kidx[k] = 0;
                    for (int cell = 0; cell < ncells; ++cell) {
                        d = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float nr = x[k] + (kidx[k] - half[k]) * dl[k] - yy[k];
                            d += nr * nr;
                        }
                        if (d < rad2) {
                            d = sqrtf(d);
                            float norm = 1.0f;
                            if (dis_norm && d > 0.0f) norm /= d;
                            const float kw = d > sp.H ? 0.0f : sph_eval(sp.w_expr, d, sp.H, sp.w_coef);
                            for (int c = 0; c < C; ++c) {
                                const float dv = dj[c];
#pragma unroll
                                for (int o = 0; o < kOChunk; ++o)
                                    if (o < on)
                                        acc[o] += wp[((size_t)(o0 + o) * C + c) * ncells + cell] * dv * kw * norm;
                            }
                        }
                        ++kidx[0];
#pragma unroll
                        for (int k = 0; k < D - 1; ++k)
                            if (kidx[k] >= ks[k]) {
                                kidx[k] = 0;
                                ++kidx[k + 1];
                            }
                    }
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
#pragma unroll
        for (int o = 0; o < kOChunk; ++o) acc[o] = group_sum(acc[o]);
        if (active && sub == 0)
            for (int o = 0; o < on; ++o) out[q * O + o0 + o] = acc[o] + (bias ? bias[o0 + o] : 0.0f);
    }
}

// Generic backward: dq by group reduction, everything neighbour-side by float atomics into
// zero-filled buffers (the launcher zero-fills); d(weight) is k_convsp_dweight_generic below.
template <int DT>
__global__ void __launch_bounds__(kThreads)
k_convsp_bwd_generic(const float* __restrict__ qlocs, const float* __restrict__ locs,
                     const float* __restrict__ data, const float* __restrict__ neighbors,
                     const float* __restrict__ weight, const float* __restrict__ go, long long BM,
                     int M, int N, int C, int ndims, int K, int O, int ncells, float radius,
                     const float* __restrict__ ksize, const float* __restrict__ dilation,
                     int dis_norm, SphParams sp, float* dq, float* dl, float* dd, float* dw,
                     int w_in_smem, int same_q_l)
{
    constexpr int MD = DT > 0 ? DT : SPNB_MAXD;
    const int D = DT > 0 ? DT : ndims;
    extern __shared__ float s_w[];  // weights when w_in_smem
    const int nw = O * C * ncells;
    if (w_in_smem) {
        for (int i = threadIdx.x; i < nw; i += kThreads) s_w[i] = weight[i];
        __syncthreads();
    }
    const float* wp = w_in_smem ? s_w : weight;
    const int lane = threadIdx.x & 31, sub = lane & (kG - 1);
    const long long q = ((long long)blockIdx.x * kThreads + threadIdx.x) / kG;
    const bool active = q < BM;
    const long long qq = active ? q : 0;
    const int b = (int)(qq / M);
    int ks[MD], half[MD];
    float dil[MD], x[MD], cull2;
    load_shape<DT>(ksize, dilation, D, radius, ks, half, dil, cull2);
    const float rad2 = radius * radius;
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[qq * D + k];
    const float* row = neighbors + qq * K;
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    const float* gi = go + qq * O;
    float a_dq[MD];
#pragma unroll
    for (int k = 0; k < D; ++k) a_dq[k] = 0.0f;

    bool done = !active;  // my group's row has ended (the list stops at its first negative entry)
    for (int jj0 = 0; jj0 < K; jj0 += kG) {
        const int jj = jj0 + sub;
        const float nb = (!done && jj < K) ? row[jj] : -1.0f;
        const unsigned neg = __ballot_sync(0xffffffffu, !(nb >= 0.0f));
        const int fneg = group_first_neg(neg, lane, sub);
        if (fneg < kG) done = true;
        if (sub < fneg) {
            const int j = (int)nb;
            const float* y = sl + (size_t)j * D;
            const float* dj = sd + (size_t)j * C;
            float yy[MD];
            float d = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                yy[k] = y[k];
                d += (x[k] - yy[k]) * (x[k] - yy[k]);
            }
            if (!(d > cull2)) {
                int kidx[MD];
#pragma unroll
                for (int k = 0; k < D; ++k) kidx[k] = 0;
                for (int cell = 0; cell < ncells; ++cell) {
                    float disp[MD];
                    d = 0.0f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        disp[k] = x[k] + (kidx[k] - half[k]) * dil[k] - yy[k];
                        d += disp[k] * disp[k];
                    }
                    if (d < rad2) {
                        d = sqrtf(d);
                        float norm = 1.0f;
                        if (dis_norm && d > 0.0f) norm /= d;
                        const bool in = !(d > sp.H);
                        const float kw = in ? sph_eval(sp.w_expr, d, sp.H, sp.w_coef) : 0.0f;
                        const float kdw = (in ? sph_eval(sp.dw_expr, d, sp.H, sp.dw_coef) : 0.0f) / d;
                        float l_dl[MD];
#pragma unroll
                        for (int k = 0; k < D; ++k) l_dl[k] = 0.0f;
                        for (int c = 0; c < C; ++c) {
                            const float dv = dj[c];
                            float v_dd = 0.0f;
                            for (int o = 0; o < O; ++o) {
                                const size_t wi = ((size_t)o * C + c) * ncells + cell;
                                const float wv = wp[wi];
                                const float g = gi[o];
                                v_dd += g * wv * kw * norm;
                                if (d > 0.0f) {
#pragma unroll
                                    for (int k = 0; k < D; ++k) {
                                        const float t = wv * dv * norm * (kdw * disp[k]) * g;
                                        a_dq[k] += t;
                                        l_dl[k] -= t;
                                    }
                                }
                            }
                            if (dd) atomicAdd(dd + ((size_t)b * N + j) * C + c, v_dd);
                        }
                        if (dl && d > 0.0f) {
#pragma unroll
                            for (int k = 0; k < D; ++k) atomicAdd(dl + ((size_t)b * N + j) * D + k, l_dl[k]);
                        }
                    }
                    ++kidx[0];
#pragma unroll
                    for (int k = 0; k < D - 1; ++k)
                        if (kidx[k] >= ks[k]) {
                            kidx[k] = 0;
                            ++kidx[k + 1];
                        }
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) a_dq[k] = group_sum(a_dq[k]);
    if (active && sub == 0 && dq) {
        for (int k = 0; k < D; ++k) {
            if (same_q_l) atomicAdd(dq + q * D + k, a_dq[k]);
            else dq[q * D + k] = a_dq[k];
        }
    }
}

// d(weight) for the generic path, in a kernel of its own so that the whole warp stays convergent:
// every lane walks the kernel cells in lockstep, per cell the data[j,c]*W*norm terms of a query's
// lanes are summed with shuffles and one lane per query adds go[q,o] times that sum to the block's
// shared accumulator (the reference does one global atomicAdd per term, all threads on the same
// nkernels*nchannels*ncells addresses, common_funcs.h:542-547).
template <int DT>
__global__ void __launch_bounds__(kThreads)
k_convsp_dweight_generic(const float* __restrict__ qlocs, const float* __restrict__ locs,
                         const float* __restrict__ data, const float* __restrict__ neighbors,
                         const float* __restrict__ go, long long BM, int M, int N, int C, int ndims, int K,
                         int O, int ncells, float radius, const float* __restrict__ ksize,
                         const float* __restrict__ dilation, int dis_norm, SphParams sp, float* dw,
                         int acc_in_smem)
{
    constexpr int MD = DT > 0 ? DT : SPNB_MAXD;
    const int D = DT > 0 ? DT : ndims;
    extern __shared__ float s_acc[];
    const int nw = O * C * ncells;
    if (acc_in_smem) {
        for (int i = threadIdx.x; i < nw; i += kThreads) s_acc[i] = 0.0f;
        __syncthreads();
    }
    float* acc = acc_in_smem ? s_acc : dw;
    const int lane = threadIdx.x & 31, sub = lane & (kG - 1);
    const long long q = ((long long)blockIdx.x * kThreads + threadIdx.x) / kG;
    const bool active = q < BM;
    const long long qq = active ? q : 0;
    const int b = (int)(qq / M);
    int ks[MD], half[MD];
    float dil[MD], x[MD], cull2;
    load_shape<DT>(ksize, dilation, D, radius, ks, half, dil, cull2);
    const float rad2 = radius * radius;
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[qq * D + k];
    const float* row = neighbors + qq * K;
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    const float* gi = go + qq * O;

    bool done = !active;  // my group's row has ended (the list stops at its first negative entry)
    for (int jj0 = 0; jj0 < K; jj0 += kG) {
        const int jj = jj0 + sub;
        const float nb = (!done && jj < K) ? row[jj] : -1.0f;
        const unsigned neg = __ballot_sync(0xffffffffu, !(nb >= 0.0f));
        const int fneg = group_first_neg(neg, lane, sub);
        if (fneg < kG) done = true;
        bool live = sub < fneg;
        const int j = live ? (int)nb : 0;
        const float* dj = sd + (size_t)j * C;
        float yy[MD];
        float d0 = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            yy[k] = sl[(size_t)j * D + k];
            d0 += (x[k] - yy[k]) * (x[k] - yy[k]);
        }
        if (d0 > cull2) live = false;
        if (__any_sync(0xffffffffu, live)) {
            int kidx[MD];
#pragma unroll
            for (int k = 0; k < D; ++k) kidx[k] = 0;
            for (int cell = 0; cell < ncells; ++cell) {
                float d = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float nr = x[k] + (kidx[k] - half[k]) * dil[k] - yy[k];
                    d += nr * nr;
                }
                float s = 0.0f;
                if (live && d < rad2) {
                    d = sqrtf(d);
                    float norm = 1.0f;
                    if (dis_norm && d > 0.0f) norm /= d;
                    s = (d > sp.H ? 0.0f : sph_eval(sp.w_expr, d, sp.H, sp.w_coef)) * norm;
                }
                if (__any_sync(0xffffffffu, s != 0.0f)) {
                    // the lanes of a group share the query, hence go[q,:]: reduce data*W*norm over the
                    // group first, then one lane per group applies go and adds O values per channel
                    for (int c = 0; c < C; ++c) {
                        float tc = s != 0.0f ? dj[c] * s : 0.0f;
                        tc = group_sum(tc);
                        if (sub == 0 && tc != 0.0f)
                            for (int o = 0; o < O; ++o)
                                atomicAdd(acc + ((size_t)o * C + c) * ncells + cell, gi[o] * tc);
                    }
                }
                ++kidx[0];
#pragma unroll
                for (int k = 0; k < D - 1; ++k)
                    if (kidx[k] >= ks[k]) {
                        kidx[k] = 0;
                        ++kidx[k + 1];
                    }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
    if (acc_in_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nw; i += kThreads) {
            const float v = s_acc[i];
            if (v != 0.0f) atomicAdd(dw + i, v);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------
static bool validate(const char* fn, int B, int M, int N, int C, int D, int K, int O, int ncells,
                     int kernel_fn)
{
    if (B <= 0 || M <= 0 || N <= 0 || C <= 0 || D <= 0 || K <= 0 || O <= 0 || ncells <= 0) {
        set_error("%s: non-positive size", fn);
        return false;
    }
    if (D > SPNB_MAX_NDIM) {
        set_error("%s: ndims=%d > %d", fn, D, SPNB_MAX_NDIM);
        return false;
    }
    if (kernel_fn < 0 || kernel_fn >= SPNB_NUM_KERNEL_FNS) {
        set_error("%s: unknown kernel function id %d", fn, kernel_fn);
        return false;
    }
    return true;
}

constexpr int kMaxSmemWeights = 40 * 1024;  // bytes of weights staged per block (x2 in backward)

}  // namespace spnb

using namespace spnb;

extern "C" {

int spnb_convsp_forward(const float* qlocs, const float* locs, const float* data,
                        const float* neighbors, const float* weight, const float* bias, int B, int M,
                        int N, int C, int D, int K, int O, int ncells, float radius,
                        const float* kernel_size, const float* dilation, int dis_norm, int kernel_fn,
                        float* out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate("spnb_convsp_forward", B, M, N, C, D, K, O, ncells, kernel_fn)) return 0;
    if (!qlocs || !locs || !data || !neighbors || !weight || !kernel_size || !dilation || !out) {
        set_error("spnb_convsp_forward: null pointer");
        return 0;
    }
    const SphParams sp = make_sph_params(kernel_fn, radius);
    const long long BM = (long long)B * M;
    const int blocks = cdiv(BM * kG, kThreads);
    bool done = false;
    if (convsp_small_supported(D, C, O, ncells)) {
        launch_convsp_fwd_small(qlocs, locs, data, neighbors, weight, bias, B, M, N, C, D, K, O, radius,
                                dis_norm, kernel_fn, out, stream);
        done = true;
    }
    if (!done) {
        const size_t wbytes = sizeof(float) * (size_t)O * C * ncells;
        const int ws = wbytes <= (size_t)kMaxSmemWeights;
        const size_t smem = ws ? wbytes : 0;
#define LAUNCH(DT)                                                                                 \
    k_convsp_fwd_generic<DT><<<blocks, kThreads, smem, stream>>>(                                  \
        qlocs, locs, data, neighbors, weight, bias, BM, M, N, C, D, K, O, ncells, radius,          \
        kernel_size, dilation, dis_norm, sp, out, ws)
        switch (D) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        default: LAUNCH(0); break;
        }
#undef LAUNCH
    }
    count_launches(1);
    return check_launch("spnb_convsp_forward") ? 1 : 0;
}

size_t spnb_convsp_backward_workspace_bytes(int nkernels, int nchannels, int ncells)
{
    (void)nkernels; (void)nchannels; (void)ncells;
    return 0;
}

int spnb_convsp_backward(const float* qlocs, const float* locs, const float* data,
                         const float* neighbors, const float* weight, int B, int M, int N, int C,
                         int D, int K, int O, int ncells, float radius, const float* kernel_size,
                         const float* dilation, int dis_norm, int kernel_fn, const float* grad_out,
                         float* dqlocs, float* dlocs, float* ddata, float* dweight,
                         const int* sym_flag, void* workspace, void* stream_)
{
    (void)workspace;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate("spnb_convsp_backward", B, M, N, C, D, K, O, ncells, kernel_fn)) return 0;
    if (!qlocs || !locs || !data || !neighbors || !weight || !kernel_size || !dilation || !grad_out) {
        set_error("spnb_convsp_backward: null pointer");
        return 0;
    }
    const SphParams sp = make_sph_params(kernel_fn, radius);
    const long long BM = (long long)B * M;
    const int blocks = cdiv(BM * kG, kThreads);
    const int same = (dqlocs != nullptr && dqlocs == dlocs) ? 1 : 0;
    if (same && M != N) {
        set_error("spnb_convsp_backward: dqlocs == dlocs requires M == N");
        return 0;
    }
    // symmetric gather needs the query set to BE the particle set
    if (sym_flag && (qlocs != locs || M != N)) sym_flag = nullptr;
    if (dweight) cudaMemsetAsync(dweight, 0, sizeof(float) * (size_t)O * C * ncells, stream);

    const bool small = convsp_small_supported(D, C, O, ncells);
    // Scatter targets must start from zero whenever the atomic path can run.  (The symmetric
    // gather overwrites, so the fill is redundant there, but whether it runs is only known on the
    // device.)
    if (dlocs) cudaMemsetAsync(dlocs, 0, sizeof(float) * (size_t)B * N * D, stream);
    if (ddata) cudaMemsetAsync(ddata, 0, sizeof(float) * (size_t)B * N * C, stream);

    if (small) {
        launch_convsp_bwd_small(qlocs, locs, data, neighbors, weight, B, M, N, C, D, K, O, radius,
                                dis_norm, kernel_fn, grad_out, dqlocs, dlocs, ddata, dweight, sym_flag,
                                same, stream);
    } else {
        const size_t wbytes = sizeof(float) * (size_t)O * C * ncells;
        const int ws = wbytes <= (size_t)kMaxSmemWeights;
        const size_t smem = ws ? wbytes : 0;
#define LAUNCH(DT)                                                                                 \
    do {                                                                                           \
        if (smem > 48 * 1024)                                                                      \
            cudaFuncSetAttribute(k_convsp_bwd_generic<DT>,                                         \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        k_convsp_bwd_generic<DT><<<blocks, kThreads, smem, stream>>>(                              \
            qlocs, locs, data, neighbors, weight, grad_out, BM, M, N, C, D, K, O, ncells, radius,  \
            kernel_size, dilation, dis_norm, sp, dqlocs, dlocs, ddata, dweight, ws, same);         \
    } while (0)
        switch (D) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        default: LAUNCH(0); break;
        }
#undef LAUNCH
        if (dweight) {
            const int as = wbytes <= (size_t)(96 * 1024);
            const size_t asmem = as ? wbytes : 0;
#define LAUNCHW(DT)                                                                                \
    do {                                                                                           \
        if (asmem > 48 * 1024)                                                                     \
            cudaFuncSetAttribute(k_convsp_dweight_generic<DT>,                                     \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem);         \
        k_convsp_dweight_generic<DT><<<blocks, kThreads, asmem, stream>>>(                         \
            qlocs, locs, data, neighbors, grad_out, BM, M, N, C, D, K, O, ncells, radius,          \
            kernel_size, dilation, dis_norm, sp, dweight, as);                                     \
    } while (0)
            switch (D) {
            case 1: LAUNCHW(1); break;
            case 2: LAUNCHW(2); break;
            case 3: LAUNCHW(3); break;
            default: LAUNCHW(0); break;
            }
#undef LAUNCHW
            count_launches(1);
        }
    }
    count_launches(1);
    return check_launch("spnb_convsp_backward") ? 1 : 0;
}

int spnb_convsp_backward_block(const float* locs, const float* data, const float* neighbors, const float* weight,
                               int B, int M, int N, int C, int D, int K, int O, int ncells, float radius,
                               int dis_norm, int kernel_fn, const float* grad_out_all, int query_offset,
                               const int* sym_flag, float* dlocs_block, float* ddata_block, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate("spnb_convsp_backward_block", B, M, N, C, D, K, O, ncells, kernel_fn)) return 0;
    if (B != 1) {
        set_error("spnb_convsp_backward_block: one scene per call (batch_size 1)");
        return 0;
    }
    if (!locs || !data || !neighbors || !weight || !grad_out_all || !sym_flag || query_offset < 0 ||
        query_offset + M > N) {
        set_error("spnb_convsp_backward_block: null pointer or query block [%d, %d) outside 0..%d", query_offset,
                  query_offset + M, N);
        return 0;
    }
    if (!convsp_small_supported(D, C, O, ncells)) {
        set_error("spnb_convsp_backward_block: shape not supported (kernel_size 1, up to 4 channels)");
        return 0;
    }
    // the symmetric gather of k_convsp_bwd_small with the queries = particles query_offset .. query_offset + M - 1:
    // d/dqlocs + d/dlocs of the block's particles in one buffer, nothing is scattered
    launch_convsp_bwd_small(locs + (size_t)query_offset * D, locs, data, neighbors, weight, B, M, N, C, D, K, O, radius,
                            dis_norm, kernel_fn, grad_out_all, dlocs_block, dlocs_block, ddata_block, nullptr, sym_flag,
                            1, stream, query_offset, N);
    count_launches(1);
    return check_launch("spnb_convsp_backward_block") ? 1 : 0;
}

}  // extern "C"
