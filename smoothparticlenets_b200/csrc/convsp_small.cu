// ConvSP fast path for sm_100a: kernel_size == 1 in every dimension (ncells == 1) and small
// compile-time channel counts -- the layers of the fluid simulation (examples/fluid_sim.py:156-175:
// C = O in {1, ndim}, spiky / dspiky / constant / cohesion kernels).
//
// Replaces kernel_convsp (reference src/gpu_kernels.cu:50-126) for these shapes; the math is
// compute_kernel_cells (src/common_funcs.h:439-583) with ncells = 1.
//
// Design (DESIGN.md "ConvSP list path"):
//  * a group of 8 lanes owns one query; per step the group reads a 128-byte chunk of the neighbour
//    row (32 entries in flight: lane `sub` takes entries sub, sub+8, sub+16, sub+24 of the chunk, so
//    every load instruction of the group covers one 32-byte sector), and each lane then has up to 4
//    independent gathers (locs / data / grad rows) outstanding -- the kernel is latency bound, and
//    memory-level parallelism per lane is what buys throughput.  The strided assignment packs a
//    short tail into slot 0 of the lanes, so the empty slots 1..3 are skipped warp-uniformly;
//  * the next chunk of the row is requested before the current one is processed, but only once it
//    is known to be needed (no terminator yet), so short lists cost one 128-byte read;
//  * blockIdx.y is the scene, so no 64-bit division is needed to find it;
//  * the in-radius predicate d2 < r*r is evaluated with separately rounded fp32 operations in the
//    reference's order (this file is compiled with -fmad=false), so membership is bit-identical to
//    the CPU reference; everything AFTER the predicate (distance, 1/d, W(d), the channel
//    contraction) uses fast fp32: rsqrtf, reciprocal multiplies and host-precomputed coefficients
//    -- a few ulp per term, far inside the 1e-5 tolerance (the generic kernels in convsp.cu keep the
//    reference's exact float/double evaluation);
//  * backward: symmetric-gather mode without atomics (see convsp.cu header) or scatter with
//    red.global.add.f32; d(weight) only in the WDW instantiation, reduced warp -> block -> one
//    atomic per element per block.
#include "convsp_small.cuh"

namespace spnb {

namespace {

constexpr int kG = 8;          // lanes per query
constexpr int kThreads = 256;  // 32 queries per block
constexpr int kEPL = 4;        // list entries per lane per chunk (one float4)
constexpr int kChunk = kG * kEPL;
#ifndef SPNB_BWD_MIN_BLOCKS
#define SPNB_BWD_MIN_BLOCKS 3
#endif

struct SphFast {
    int w_expr, dw_expr;
    float H, invH, H2;
    float wc, dwc;  // float roundings of the double coefficient prefixes
};

// Fast fp32 evaluation of expression `e` (ids of spnb_common.cuh) at distance d (d2 = d*d).
__device__ __forceinline__ float sph_fast(int e, float d, float d2, float c, const SphFast& p)
{
    switch (e) {
    case E_DEFAULT:   { const float q = p.H2 - d2; return c * q * q * q; }
    case E_DDEFAULT:  { const float q = p.H2 - d2; return c * q * q * d; }
    case E_DDEFAULT2: return c * (p.H2 * p.H2 + d2 * (5.0f * d2 - 6.0f * p.H2));
    case E_D_DDEFAULT2: return c * d * (20.0f * d2 - 12.0f * p.H2);
    case E_PRESSURE:  { const float q = p.H - d; return c * q * q * q; }
    case E_DPRESSURE: { const float q = p.H - d; return c * q * q; }
    case E_DPRESSURE2: return c * (p.H - d);
    case E_D_DPRESSURE2: return c;
    case E_INDIRECT:  return p.H - d;
    case E_D_INDIRECT: return -1.0f;
    case E_CONSTANT:  return 1.0f;
    case E_D_CONSTANT: return 0.0f;
    case E_SPIKY:     { const float q = 1.0f - d * p.invH; return c * q * q; }
    case E_DSPIKY:    return c * (1.0f - d * p.invH);   // c = -15/(pi H^3) * 2 / H
    case E_D_DSPIKY:  return c;
    case E_COHESION:  { const float t = d * p.invH; return (7.0f - 6.0f * t) * t * t - 1.0f; }
    case E_D_COHESION: return 2.0f * d * (7.0f * p.H - 9.0f * d) * (p.invH * p.invH * p.invH);
    case E_SIGMOID:   return 1.0f / (1.0f + expf((d - 0.2f * p.H) * 20.0f * p.invH));
    case E_D_SIGMOID: { const float ex = expf((d - 0.2f * p.H) * 20.0f * p.invH);
                        return -20.0f * ex * p.invH / ((ex + 1.0f) * (ex + 1.0f)); }
    default: return 0.0f;
    }
}

// Entries base+sub, base+sub+8, base+sub+16, base+sub+24 of a row (-1 beyond K).
__device__ __forceinline__ void load_entries(const float* __restrict__ row, int p, int K, float* e)
{
    if (p - (p & (kG - 1)) + kChunk <= K) {  // whole chunk inside the row (uniform per group)
#pragma unroll
        for (int i = 0; i < kEPL; ++i) e[i] = row[p + i * kG];
    } else {
#pragma unroll
        for (int i = 0; i < kEPL; ++i) e[i] = p + i * kG < K ? row[p + i * kG] : -1.0f;
    }
}

// Per-lane number of valid slots in the current chunk and whether the group saw a terminator.
// The list ends at the first negative entry of the row (common_funcs.h:476).
__device__ __forceinline__ int chunk_valid(const float* e, int lane, int sub, bool& ended)
{
    int first = kChunk;  // position of the first negative entry within the chunk
#pragma unroll
    for (int i = kEPL - 1; i >= 0; --i) {
        const unsigned negb = __ballot_sync(0xffffffffu, !(e[i] >= 0.0f));
        const unsigned g = (negb >> (lane - sub)) & ((1u << kG) - 1u);
        if (g) first = i * kG + __ffs(g) - 1;
    }
    ended = first < kChunk;
    // slots i with i*kG + sub < first are valid
    return first > sub ? (first - sub + kG - 1) / kG : 0;
}

// 1/sqrt(x) as a single MUFU.RSQ (about 1 ulp); callers guard x > 0.
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <typename T>
__device__ __forceinline__ T group_sum(T v)
{
#pragma unroll
    for (int o = kG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- forward -------------------------------------------------------------------------------------
template <int D, int C, int O, int FN>
__global__ void __launch_bounds__(kThreads)
k_convsp_fwd_small(const float* __restrict__ qlocs, const float* __restrict__ locs,
                   const float* __restrict__ data, const float* __restrict__ neighbors,
                   const float* __restrict__ weight, const float* __restrict__ bias, long long BM,
                   int M, int N, int K, float rad2, int dis_norm, SphFast sp, int vec,
                   float* __restrict__ out)
{
    const int lane = threadIdx.x & 31, sub = lane & (kG - 1);
    const int b = blockIdx.y;
    const int m = (blockIdx.x * kThreads + threadIdx.x) / kG;  // query within the scene
    const bool active = m < M;
    const long long q = (long long)b * M + m;
    const long long qq = active ? q : (long long)b * M;
    const int we = FN >= 0 ? FN : sp.w_expr;
    float w[O * C];
#pragma unroll
    for (int i = 0; i < O * C; ++i) w[i] = weight[i];
    float x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[qq * D + k];
    const float* row = neighbors + qq * K;
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    float acc[O];
#pragma unroll
    for (int o = 0; o < O; ++o) acc[o] = 0.0f;

    float e[kEPL], nxt[kEPL];
#pragma unroll
    for (int i = 0; i < kEPL; ++i) e[i] = nxt[i] = -1.0f;
    if (active) load_entries(row, sub, K, e);
    for (int base = 0; base < K; base += kChunk) {
        bool ended;
        const int nvalid = chunk_valid(e, lane, sub, ended);
        if (!ended && base + kChunk < K) load_entries(row, base + kChunk + sub, K, nxt);
        const int maxvalid = __reduce_max_sync(0xffffffffu, nvalid);

        float y[kEPL][D], dj[kEPL][C];
#pragma unroll
        for (int i = 0; i < kEPL; ++i) {
            if (i < maxvalid) {
                const int j = i < nvalid ? (int)e[i] : 0;
#pragma unroll
                for (int k = 0; k < D; ++k) y[i][k] = sl[(unsigned)j * (unsigned)D + k];
#pragma unroll
                for (int c = 0; c < C; ++c) dj[i][c] = sd[(unsigned)j * (unsigned)C + c];
            }
        }
#pragma unroll
        for (int i = 0; i < kEPL; ++i) {
            if (i < maxvalid) {
                float d2 = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float nr = x[k] - y[i][k];
                    d2 += nr * nr;
                }
                if (i < nvalid && d2 < rad2) {
                    const float inv = fast_rsqrt(d2);
                    const float d = d2 > 0.0f ? d2 * inv : 0.0f;
                    float s = sph_fast(we, d, d2, sp.wc, sp);
                    if (dis_norm && d2 > 0.0f) s *= inv;
#pragma unroll
                    for (int o = 0; o < O; ++o)
#pragma unroll
                        for (int c = 0; c < C; ++c) acc[o] = fmaf(w[o * C + c] * dj[i][c], s, acc[o]);
                }
            }
        }
        if (__all_sync(0xffffffffu, ended)) break;
#pragma unroll
        for (int i = 0; i < kEPL; ++i) {
            e[i] = nxt[i];
            nxt[i] = -1.0f;
        }
    }
#pragma unroll
    for (int o = 0; o < O; ++o) acc[o] = group_sum(acc[o]);
    if (active && sub == 0) {
#pragma unroll
        for (int o = 0; o < O; ++o) out[q * O + o] = acc[o] + (bias ? bias[o] : 0.0f);
    }
}

// ---- backward ------------------------------------------------------------------------------------
// `go` = grad_output [B,M,O].  See spnb_convsp_backward (include/spnb.h) for the buffer contract.
template <int D, int C, int O, int FN, bool WDW>
__global__ void __launch_bounds__(kThreads, SPNB_BWD_MIN_BLOCKS)
k_convsp_bwd_small(const float* __restrict__ qlocs, const float* __restrict__ locs,
                   const float* __restrict__ data, const float* __restrict__ neighbors,
                   const float* __restrict__ weight, const float* __restrict__ go, long long BM,
                   int M, int N, int K, float rad2, int dis_norm, SphFast sp, int vec, float* dq,
                   float* dl, float* dd, float* dw, const int* sym_flag, int same_q_l)
{
    __shared__ float s_dw[WDW ? O * C : 1];
    const bool sym = sym_flag != nullptr && *sym_flag == 0;
    const int lane = threadIdx.x & 31, sub = lane & (kG - 1);
    const int b = blockIdx.y;
    const int m = (blockIdx.x * kThreads + threadIdx.x) / kG;  // query within the scene
    const bool active = m < M;
    const long long q = (long long)b * M + m;
    const long long qq = active ? q : (long long)b * M;
    const int we = FN >= 0 ? FN : sp.w_expr;
    const int dwe = FN >= 0 ? (FN == E_SPIKY ? E_DSPIKY : FN == E_DSPIKY ? E_D_DSPIKY
                              : FN == E_CONSTANT ? E_D_CONSTANT : FN == E_COHESION ? E_D_COHESION
                              : sp.dw_expr) : sp.dw_expr;
    if (WDW) {
        if (threadIdx.x < O * C) s_dw[threadIdx.x] = 0.0f;
        __syncthreads();
    }
    float w[O * C];
#pragma unroll
    for (int i = 0; i < O * C; ++i) w[i] = weight[i];
    float x[D], gi[O], di[C];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[qq * D + k];
#pragma unroll
    for (int o = 0; o < O; ++o) gi[o] = go[qq * O + o];
    const float* row = neighbors + qq * K;
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    const float* sg = go + (size_t)b * M * O;  // symmetric mode only (M == N)
#pragma unroll
    for (int c = 0; c < C; ++c) di[c] = sym ? sd[(qq - (long long)b * M) * C + c] : 0.0f;
    float a_dq[D], a_dl[D], a_dd[C], a_dw[WDW ? O * C : 1];
#pragma unroll
    for (int k = 0; k < D; ++k) a_dq[k] = a_dl[k] = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) a_dd[c] = 0.0f;
#pragma unroll
    for (int i = 0; i < (WDW ? O * C : 1); ++i) a_dw[i] = 0.0f;

    float e[kEPL], nxt[kEPL];
#pragma unroll
    for (int i = 0; i < kEPL; ++i) e[i] = nxt[i] = -1.0f;
    if (active) load_entries(row, sub, K, e);
    for (int base = 0; base < K; base += kChunk) {
        bool ended;
        const int nvalid = chunk_valid(e, lane, sub, ended);
        if (!ended && base + kChunk < K) load_entries(row, base + kChunk + sub, K, nxt);
        const int maxvalid = __reduce_max_sync(0xffffffffu, nvalid);

        int jn[kEPL];
        float y[kEPL][D], dj[kEPL][C], gj[kEPL][O];
#pragma unroll
        for (int i = 0; i < kEPL; ++i) {
            if (i >= maxvalid) continue;
            const int j = i < nvalid ? (int)e[i] : 0;
            jn[i] = j;
#pragma unroll
            for (int k = 0; k < D; ++k) y[i][k] = sl[(unsigned)j * (unsigned)D + k];
#pragma unroll
            for (int c = 0; c < C; ++c) dj[i][c] = sd[(unsigned)j * (unsigned)C + c];
            if (sym) {
#pragma unroll
                for (int o = 0; o < O; ++o) gj[i][o] = sg[(unsigned)j * (unsigned)O + o];
            }
        }
#pragma unroll
        for (int i = 0; i < kEPL; ++i) {
            if (i >= maxvalid) continue;
            float disp[D];
            float d2 = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                disp[k] = x[k] - y[i][k];
                d2 += disp[k] * disp[k];
            }
            if (i < nvalid && d2 < rad2) {
                const float inv = fast_rsqrt(d2);
                const bool pos = d2 > 0.0f;
                const float d = pos ? d2 * inv : 0.0f;
                const float norm = (dis_norm && pos) ? inv : 1.0f;
                const float s = sph_fast(we, d, d2, sp.wc, sp) * norm;  // W(d)*norm
                // (dW/dd)/d * norm; gradients wrt positions vanish for coincident points (d == 0)
                const float t = pos ? sph_fast(dwe, d, d2, sp.dwc, sp) * inv * norm : 0.0f;
                // A = sum_{o,c} go[q,o] w[o,c] data[j,c] : pair (q, j), this row as the query
                float A = 0.0f;
#pragma unroll
                for (int o = 0; o < O; ++o)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float gd = gi[o] * dj[i][c];
                        A = fmaf(w[o * C + c], gd, A);
                        if (WDW) a_dw[o * C + c] = fmaf(gd, s, a_dw[o * C + c]);
                    }
                const float At = A * t;
#pragma unroll
                for (int k = 0; k < D; ++k) a_dq[k] = fmaf(At, disp[k], a_dq[k]);
                if (sym) {
                    // pair (j, q): this particle as the neighbour of query j -- same distance,
                    // displacement negated; gathers what the reference scatters with atomics
                    float Bt = 0.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float v = 0.0f;
#pragma unroll
                        for (int o = 0; o < O; ++o) v = fmaf(gj[i][o], w[o * C + c], v);
                        a_dd[c] = fmaf(v, s, a_dd[c]);
                        Bt = fmaf(v, di[c], Bt);
                    }
                    Bt *= t;
#pragma unroll
                    for (int k = 0; k < D; ++k) a_dl[k] = fmaf(Bt, disp[k], a_dl[k]);
                } else {
                    const size_t jo = (size_t)b * N + jn[i];
                    if (dd) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            float v = 0.0f;
#pragma unroll
                            for (int o = 0; o < O; ++o) v = fmaf(gi[o], w[o * C + c], v);
                            atomicAdd(dd + jo * C + c, v * s);
                        }
                    }
                    if (dl && pos) {
#pragma unroll
                        for (int k = 0; k < D; ++k) atomicAdd(dl + jo * D + k, -At * disp[k]);
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, ended)) break;
#pragma unroll
        for (int i = 0; i < kEPL; ++i) {
            e[i] = nxt[i];
            nxt[i] = -1.0f;
        }
    }
    if (sym && same_q_l) {
        // only the sum d/dqlocs + d/dlocs is wanted: reduce once
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a_dq[k] += a_dl[k];
            a_dl[k] = 0.0f;
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        a_dq[k] = group_sum(a_dq[k]);
        if (sym && !same_q_l) a_dl[k] = group_sum(a_dl[k]);
    }
    if (sym) {
#pragma unroll
        for (int c = 0; c < C; ++c) a_dd[c] = group_sum(a_dd[c]);
    }
    if (active && sub == 0) {
        if (sym) {
            if (same_q_l) {
                if (dq)
#pragma unroll
                    for (int k = 0; k < D; ++k) dq[q * D + k] = a_dq[k] + a_dl[k];
            } else {
                if (dq)
#pragma unroll
                    for (int k = 0; k < D; ++k) dq[q * D + k] = a_dq[k];
                if (dl)
#pragma unroll
                    for (int k = 0; k < D; ++k) dl[q * D + k] = a_dl[k];
            }
            if (dd)
#pragma unroll
                for (int c = 0; c < C; ++c) dd[q * C + c] = a_dd[c];
        } else if (dq) {
            if (same_q_l) {
#pragma unroll
                for (int k = 0; k < D; ++k) atomicAdd(dq + q * D + k, a_dq[k]);
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) dq[q * D + k] = a_dq[k];
            }
        }
    }
    if (WDW) {
#pragma unroll
        for (int i = 0; i < O * C; ++i) {
            float v = a_dw[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(&s_dw[i], v);
        }
        __syncthreads();
        if (threadIdx.x < O * C) atomicAdd(dw + threadIdx.x, s_dw[threadIdx.x]);
    }
}

SphFast make_fast(const SphParams& p)
{
    SphFast f;
    f.w_expr = p.w_expr;
    f.dw_expr = p.dw_expr;
    f.H = p.H;
    f.invH = 1.0f / p.H;
    f.H2 = p.H * p.H;
    double wc = p.w_coef, dwc = p.dw_coef;
    // dspiky's expression divides by H once more after the prefix (kernels.py:98)
    if (p.w_expr == E_DSPIKY) wc = wc / (double)p.H;
    if (p.dw_expr == E_DSPIKY) dwc = dwc / (double)p.H;
    f.wc = (float)wc;
    f.dwc = (float)dwc;
    return f;
}

bool rows_vectorizable(const float* neighbors, int K)
{
    return (K % 4 == 0) && (((uintptr_t)neighbors & 15u) == 0);
}

}  // namespace

#define SPNB_SMALL_SHAPES(X) X(3, 1, 1) X(3, 3, 3) X(2, 1, 1) X(2, 2, 2)
#define SPNB_SMALL_FNS(X, DD, CC, OO) \
    X(DD, CC, OO, E_SPIKY) X(DD, CC, OO, E_DSPIKY) X(DD, CC, OO, E_CONSTANT) X(DD, CC, OO, E_COHESION)

bool convsp_small_supported(int D, int C, int O, int ncells)
{
    if (ncells != 1) return false;
#define X(DD, CC, OO) if (D == DD && C == CC && O == OO) return true;
    SPNB_SMALL_SHAPES(X)
#undef X
    return false;
}

void launch_convsp_fwd_small(const float* qlocs, const float* locs, const float* data,
                             const float* neighbors, const float* weight, const float* bias, int B,
                             int M, int N, int C, int D, int K, int O, float radius, int dis_norm,
                             int kernel_fn, float* out, cudaStream_t stream)
{
    const SphFast sp = make_fast(make_sph_params(kernel_fn, radius));
    const long long BM = (long long)B * M;
    const dim3 blocks(cdiv((long long)M * kG, kThreads), B);
    const int vec = rows_vectorizable(neighbors, K) ? 1 : 0;
    const float rad2 = radius * radius;
    bool done = false;
#define LAUNCH(DD, CC, OO, FN)                                                                     \
    k_convsp_fwd_small<DD, CC, OO, FN><<<blocks, kThreads, 0, stream>>>(                           \
        qlocs, locs, data, neighbors, weight, bias, BM, M, N, K, rad2, dis_norm, sp, vec, out)
#define XF(DD, CC, OO, FN)                                                                         \
    if (!done && kernel_fn == FN) { LAUNCH(DD, CC, OO, FN); done = true; }
#define X(DD, CC, OO)                                                                              \
    if (!done && D == DD && C == CC && O == OO) {                                                  \
        SPNB_SMALL_FNS(XF, DD, CC, OO)                                                             \
        if (!done) { LAUNCH(DD, CC, OO, -1); done = true; }                                        \
    }
    SPNB_SMALL_SHAPES(X)
#undef X
#undef XF
#undef LAUNCH
}

void launch_convsp_bwd_small(const float* qlocs, const float* locs, const float* data,
                             const float* neighbors, const float* weight, int B, int M, int N, int C,
                             int D, int K, int O, float radius, int dis_norm, int kernel_fn,
                             const float* grad_out, float* dqlocs, float* dlocs, float* ddata,
                             float* dweight, const int* sym_flag, int same, cudaStream_t stream)
{
    const SphFast sp = make_fast(make_sph_params(kernel_fn, radius));
    const long long BM = (long long)B * M;
    const dim3 blocks(cdiv((long long)M * kG, kThreads), B);
    const int vec = rows_vectorizable(neighbors, K) ? 1 : 0;
    const float rad2 = radius * radius;
    bool done = false;
#define LAUNCH(DD, CC, OO, FN, WD)                                                                 \
    k_convsp_bwd_small<DD, CC, OO, FN, WD><<<blocks, kThreads, 0, stream>>>(                       \
        qlocs, locs, data, neighbors, weight, grad_out, BM, M, N, K, rad2, dis_norm, sp, vec,      \
        dqlocs, dlocs, ddata, dweight, sym_flag, same)
#define XF(DD, CC, OO, FN)                                                                         \
    if (!done && kernel_fn == FN && !dweight) { LAUNCH(DD, CC, OO, FN, false); done = true; }
#define X(DD, CC, OO)                                                                              \
    if (!done && D == DD && C == CC && O == OO) {                                                  \
        SPNB_SMALL_FNS(XF, DD, CC, OO)                                                             \
        if (!done) {                                                                               \
            if (dweight) LAUNCH(DD, CC, OO, -1, true);                                             \
            else LAUNCH(DD, CC, OO, -1, false);                                                    \
            done = true;                                                                           \
        }                                                                                          \
    }
    SPNB_SMALL_SHAPES(X)
#undef X
#undef XF
#undef LAUNCH
}

}  // namespace spnb
