// ConvSP fast path for sm_100a: kernel_size == 1 in every dimension (ncells == 1) and small
// compile-time channel counts -- the layers of the fluid simulation (examples/fluid_sim.py:156-175:
// C = O in {1, ndim}, spiky / dspiky / constant / cohesion kernels).
//
// Replaces kernel_convsp (reference src/gpu_kernels.cu:50-126) for these shapes; the math is
// compute_kernel_cells (src/common_funcs.h:439-583) with ncells = 1.
//
// Design (DESIGN.md "ConvSP list path"):
//  * the neighbour rows are walked with list_walk.cuh: G = 4 lanes per query, rows staged through
//    shared memory 32 entries at a time with one ballot per row to find the terminator, a few
//    independent gathers in flight per lane;
//  * blockIdx.y is the scene, so no 64-bit division is needed to find it;
//  * the in-radius predicate d2 < r*r is evaluated with separately rounded fp32 operations in the
//    reference's order (this file is compiled with -fmad=false), so membership is bit-identical to
//    the CPU reference; everything AFTER the predicate (distance, 1/d, W(d), the channel
//    contraction) uses fast fp32: rsqrt.approx, reciprocal multiplies and host-precomputed
//    coefficients -- a few ulp per term, far inside the 1e-5 tolerance (the generic kernels in
//    convsp.cu keep the reference's exact float/double evaluation);
//  * the weights are applied once per query, not once per pair: forward accumulates
//    sum_j W*norm*data[j,c] and multiplies by w[o,c] in the epilogue; backward folds them into
//    u_i[c] = sum_o go[i,o] w[o,c];
//  * backward: symmetric-gather mode without atomics (see convsp.cu header) or scatter with
//    red.global.add.f32; d(weight) only in the WDW instantiation, reduced warp -> block -> one
//    atomic per element per block.
#include "convsp_small.cuh"
#include "list_walk.cuh"

namespace spnb {

namespace {

// lanes per query / list entries in flight per lane (list_walk.cuh); tunable at build time
#ifndef SPNB_SMALL_G
#define SPNB_SMALL_G 4
#endif
#ifndef SPNB_SMALL_FWD_U
#define SPNB_SMALL_FWD_U 2
#endif
#ifndef SPNB_SMALL_BWD_U
#define SPNB_SMALL_BWD_U 1
#endif
constexpr int kG = SPNB_SMALL_G, kFwdU = SPNB_SMALL_FWD_U, kBwdU = SPNB_SMALL_BWD_U;
constexpr int kThreads = 128;

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

// ---- forward -------------------------------------------------------------------------------------
template <int D, int C, int O, int FN>
__global__ void __launch_bounds__(kThreads)
k_convsp_fwd_small(const float* __restrict__ qlocs, const float* __restrict__ locs,
                   const float* __restrict__ data, const float* __restrict__ neighbors,
                   const float* __restrict__ weight, const float* __restrict__ bias, int M, int N, int K,
                   float rad2, int dis_norm, SphFast sp, float* __restrict__ out)
{
    constexpr int G = kG, U = kFwdU, QPB = kThreads / G, R = 32 / G;
    __shared__ WalkSmem<G> s_walk[kThreads / 32];
    const int warp = threadIdx.x >> 5, sub = threadIdx.x % G;
    const int b = blockIdx.y;
    const int m = blockIdx.x * QPB + threadIdx.x / G;  // query within the scene
    const bool active = m < M;
    const size_t q = (size_t)b * M + (active ? m : 0);
    const int m0 = blockIdx.x * QPB + warp * R;
    const int nrows = min(R, max(0, M - m0));
    const int we = FN >= 0 ? FN : sp.w_expr;
    float x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[q * D + k];
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
    float acc[C];  // sum_j W*norm*data[j,c]; the weights are applied once per query below
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.0f;
    const float* warp_rows = neighbors + ((size_t)b * M + min(m0, M - 1)) * K;
    prefetch_rows_ahead(neighbors, M, K, QPB, 2);

    walk_rows<G, U>(warp_rows, K, nrows, s_walk[warp], [&](const int* j, const bool* valid) {
        float y[U][D], dj[U][C];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < D; ++k) y[u][k] = sl[(unsigned)j[u] * (unsigned)D + k];
#pragma unroll
            for (int c = 0; c < C; ++c) dj[u][c] = sd[(unsigned)j[u] * (unsigned)C + c];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float d2 = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float nr = x[k] - y[u][k];
                d2 += nr * nr;
            }
            if (valid[u] && d2 < rad2) {
                const float inv = fast_rsqrt(d2);
                const float d = d2 > 0.0f ? d2 * inv : 0.0f;
                float s = sph_fast(we, d, d2, sp.wc, sp);
                if (dis_norm && d2 > 0.0f) s *= inv;
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(s, dj[u][c], acc[c]);
            }
        }
    });
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = group_sum<G>(acc[c]);
    if (active) {
        for (int o = sub; o < O; o += G) {
            float v = bias ? bias[o] : 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) v = fmaf(weight[o * C + c], acc[c], v);
            out[q * O + o] = v;
        }
    }
}

// ---- backward ------------------------------------------------------------------------------------
// `go` = grad_output [B,M,O].  See spnb_convsp_backward (include/spnb.h) for the buffer contract.
template <int D, int C, int O, int FN, bool WDW>
__global__ void __launch_bounds__(kThreads)
k_convsp_bwd_small(const float* __restrict__ qlocs, const float* __restrict__ locs,
                   const float* __restrict__ data, const float* __restrict__ neighbors,
                   const float* __restrict__ weight, const float* __restrict__ go, int M, int N, int K,
                   float rad2, int dis_norm, SphFast sp, float* dq, float* dl, float* dd, float* dw,
                   const int* sym_flag, int same_q_l, int q_off, int go_rows)
{
    constexpr int G = kG, U = kBwdU, QPB = kThreads / G, R = 32 / G;
    __shared__ WalkSmem<G> s_walk[kThreads / 32];
    __shared__ float s_dw[WDW ? O * C : 1];
    const bool sym = sym_flag != nullptr && *sym_flag == 0;
    // query-block call (spnb_convsp_backward_block): the outputs have M rows, so only the gather is possible
    if (!sym && (q_off != 0 || go_rows != M)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = threadIdx.x % G;
    const int b = blockIdx.y;
    const int m = blockIdx.x * QPB + threadIdx.x / G;
    const bool active = m < M;
    const size_t q = (size_t)b * M + (active ? m : 0);
    const int m0 = blockIdx.x * QPB + warp * R;
    const int nrows = min(R, max(0, M - m0));
    const int we = FN >= 0 ? FN : sp.w_expr;
    const int dwe = FN >= 0 ? (FN == E_SPIKY ? E_DSPIKY : FN == E_DSPIKY ? E_D_DSPIKY
                              : FN == E_CONSTANT ? E_D_CONSTANT : FN == E_COHESION ? E_D_COHESION
                              : sp.dw_expr) : sp.dw_expr;
    if (WDW) {
        if (threadIdx.x < O * C) s_dw[threadIdx.x] = 0.0f;
        __syncthreads();
    }
    float x[D], gi[O], ui[C], di[C];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = qlocs[q * D + k];
    // grad_output rows per scene: M in the ordinary call (row = query), N in the query-block call (row = particle,
    // the queries are the particles q_off .. q_off + M - 1)
    const float* sg = go + (size_t)b * go_rows * O;
    const size_t me = (size_t)q_off + (active ? m : 0);  // this query as a particle (symmetric mode)
#pragma unroll
    for (int o = 0; o < O; ++o) gi[o] = sg[me * O + o];
    // u_i[c] = sum_o go[i,o] w[o,c]: the weights fold into one vector per particle
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float v = 0.0f;
#pragma unroll
        for (int o = 0; o < O; ++o) v = fmaf(gi[o], weight[o * C + c], v);
        ui[c] = v;
    }
    const float* sl = locs + (size_t)b * N * D;
    const float* sd = data + (size_t)b * N * C;
#pragma unroll
    for (int c = 0; c < C; ++c) di[c] = sym ? sd[me * C + c] : 0.0f;
    float wreg[O * C];
#pragma unroll
    for (int i = 0; i < O * C; ++i) wreg[i] = weight[i];
    float a_dq[D], a_dl[D], a_dd[C], a_dw[WDW ? O * C : 1];
#pragma unroll
    for (int k = 0; k < D; ++k) a_dq[k] = a_dl[k] = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) a_dd[c] = 0.0f;
#pragma unroll
    for (int i = 0; i < (WDW ? O * C : 1); ++i) a_dw[i] = 0.0f;
    const float* warp_rows = neighbors + ((size_t)b * M + min(m0, M - 1)) * K;
    prefetch_rows_ahead(neighbors, M, K, QPB, 2);

    walk_rows<G, U>(warp_rows, K, nrows, s_walk[warp], [&](const int* j, const bool* valid) {
        float y[U][D], dj[U][C], gj[U][O];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < D; ++k) y[u][k] = sl[(unsigned)j[u] * (unsigned)D + k];
#pragma unroll
            for (int c = 0; c < C; ++c) dj[u][c] = sd[(unsigned)j[u] * (unsigned)C + c];
            if (sym) {
#pragma unroll
                for (int o = 0; o < O; ++o) gj[u][o] = sg[(unsigned)j[u] * (unsigned)O + o];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float disp[D];
            float d2 = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                disp[k] = x[k] - y[u][k];
                d2 += disp[k] * disp[k];
            }
            if (valid[u] && d2 < rad2) {
                const float inv = fast_rsqrt(d2);
                const bool pos = d2 > 0.0f;
                const float d = pos ? d2 * inv : 0.0f;
                const float norm = (dis_norm && pos) ? inv : 1.0f;
                const float s = sph_fast(we, d, d2, sp.wc, sp) * norm;  // W(d)*norm
                // (dW/dd)/d * norm; gradients wrt positions vanish for coincident points (d == 0)
                const float t = pos ? sph_fast(dwe, d, d2, sp.dwc, sp) * inv * norm : 0.0f;
                // pair (q, j), this row as the query: A = u_i . data[j]
                float A = 0.0f;
#pragma unroll
                for (int c = 0; c < C; ++c) A = fmaf(ui[c], dj[u][c], A);
                if (WDW) {
#pragma unroll
                    for (int o = 0; o < O; ++o)
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            a_dw[o * C + c] = fmaf(gi[o] * dj[u][c], s, a_dw[o * C + c]);
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
                        for (int o = 0; o < O; ++o) v = fmaf(gj[u][o], wreg[o * C + c], v);
                        a_dd[c] = fmaf(v, s, a_dd[c]);
                        Bt = fmaf(v, di[c], Bt);
                    }
                    Bt *= t;
#pragma unroll
                    for (int k = 0; k < D; ++k) a_dl[k] = fmaf(Bt, disp[k], a_dl[k]);
                } else {
                    const size_t jo = (size_t)b * N + j[u];
                    if (dd) {
#pragma unroll
                        for (int c = 0; c < C; ++c) atomicAdd(dd + jo * C + c, ui[c] * s);
                    }
                    if (dl && pos) {
#pragma unroll
                        for (int k = 0; k < D; ++k) atomicAdd(dl + jo * D + k, -At * disp[k]);
                    }
                }
            }
        }
    });
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
        a_dq[k] = group_sum<G>(a_dq[k]);
        if (sym && !same_q_l) a_dl[k] = group_sum<G>(a_dl[k]);
    }
    if (sym) {
#pragma unroll
        for (int c = 0; c < C; ++c) a_dd[c] = group_sum<G>(a_dd[c]);
    }
    if (active && sub == 0) {
        if (sym) {
            if (same_q_l) {
                if (dq)
#pragma unroll
                    for (int k = 0; k < D; ++k) dq[q * D + k] = a_dq[k];
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
    const dim3 blocks(cdiv((long long)M * kG, kThreads), B);
    const float rad2 = radius * radius;
    bool done = false;
#define LAUNCH(DD, CC, OO, FN)                                                                     \
    k_convsp_fwd_small<DD, CC, OO, FN><<<blocks, kThreads, 0, stream>>>(                           \
        qlocs, locs, data, neighbors, weight, bias, M, N, K, rad2, dis_norm, sp, out)
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
                             float* dweight, const int* sym_flag, int same, cudaStream_t stream, int q_off,
                             int go_rows)
{
    if (go_rows <= 0) go_rows = M;
    const SphFast sp = make_fast(make_sph_params(kernel_fn, radius));
    const dim3 blocks(cdiv((long long)M * kG, kThreads), B);
    const float rad2 = radius * radius;
    bool done = false;
#define LAUNCH(DD, CC, OO, FN, WD)                                                                 \
    k_convsp_bwd_small<DD, CC, OO, FN, WD><<<blocks, kThreads, 0, stream>>>(                       \
        qlocs, locs, data, neighbors, weight, grad_out, M, N, K, rad2, dis_norm, sp, dqlocs, dlocs, \
        ddata, dweight, sym_flag, same, q_off, go_rows)
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
