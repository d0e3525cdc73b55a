// Fused ConvSP "group": several ConvSP layers that share (locs, neighbors, radius) -- and have
// kernel_size 1 -- evaluated in ONE walk over the neighbour lists.
//
// This is SURVEY.md section 8(f) rank 1: the solver iteration of the reference's fluid simulation calls
// 9 ConvSP layers on the same particle set (examples/fluid_sim.py:367-397); each per-layer kernel
// re-reads the neighbour list, re-gathers the neighbour positions and recomputes the distance.  The
// math per layer is unchanged (compute_kernel_cells, src/common_funcs.h:439-583, ncells = 1):
//     out_l[i,o] = bias_l[o] + sum_c w_l[o,c] * sum_j W_l(d_ij) * norm_l(d_ij) * data_l[j,c]
//
// Design:
//  * a PACK pre-pass writes one 16-byte aligned record per particle: position, the distinct data
//    tensors of the group (a data tensor that IS the position tensor is not duplicated) and, for the
//    backward pass, U_l[n,c] = sum_o grad_out_l[n,o] * w_l[o,c].  The main kernel then gathers each
//    neighbour with a few LDG.128 instead of many scalar loads from separate tensors;
//  * the main kernels walk the rows with list_walk.cuh (one thread per query, rows staged through
//    shared memory, exact in-radius predicate, fast fp32 after it); per pair the geometry is computed once,
//    W / dW once per distinct (kernel, dis_norm), and only C_l FMAs per layer are spent on channels
//    because the weights are applied once per query in the epilogue (forward) or folded into U_l
//    (backward);
//  * backward: symmetric-gather (no atomics) when the device flag allows, else scatter with float
//    atomics -- same rule as the per-layer kernels.  d(weight) is not produced here; groups whose
//    layers need it fall back to the per-layer kernels in the Python layer.
//
// The channel layout of a group is a compile-time signature (template parameters), so every record
// field and accumulator lives in a register.  Signatures used by the fluid step are instantiated at
// the bottom; anything else reports "unsupported" and the caller uses the per-layer path.
#include <string.h>

#include "convsp_small.cuh"
#include "list_walk.cuh"
#include "tile_lists.cuh"

namespace spnb {

namespace {

constexpr int kThreads = 128;
// lanes per query / list entries in flight per lane (see list_walk.cuh); tunable at build time.
// Defaults from the sweep in profiles/README.md (tools/tune_group.sh): G = 4 with few entries in flight.
#ifndef SPNB_GROUP_FWD_G
#define SPNB_GROUP_FWD_G 4
#endif
#ifndef SPNB_GROUP_FWD_U
#define SPNB_GROUP_FWD_U 2
#endif
#ifndef SPNB_GROUP_BWD_G
#define SPNB_GROUP_BWD_G 4
#endif
#ifndef SPNB_GROUP_BWD_U
#define SPNB_GROUP_BWD_U 1
#endif
// tile-list kernels (tile_lists.cuh): lanes per query
// (measured, tools/tune_tile.sh: one lane per query when the record is a single float4, two when the
// gather needs two LDS.128; four for the backward records)
#ifndef SPNB_TILE_FWD_G
#define SPNB_TILE_FWD_G 0  // 0 = by record width
#endif
#ifndef SPNB_TILE_BWD_G
#define SPNB_TILE_BWD_G 4
#endif
constexpr int kMaxLayers = 6;
constexpr unsigned kSrcLocs = 15;  // "data is the position tensor"

// ---- compile-time group signature -------------------------------------------------------------------
// CS: 4 bits per layer = in-channels C_l;  SS: 4 bits per layer = index of the distinct data tensor
// feeding the layer, or kSrcLocs;  FS: 4 bits per layer = kernel id;  NS: 1 bit per layer = dis_norm.
// Kernel ids are compile-time so that W / dW are straight-line code (a run-time switch per layer and
// pair costs more than the arithmetic itself and blows the instruction cache -- measured).
__host__ __device__ constexpr int deriv_expr_ct(int fn)
{
    return fn == E_DEFAULT ? E_DDEFAULT : fn == E_DDEFAULT ? E_DDEFAULT2 : fn == E_DDEFAULT2 ? E_D_DDEFAULT2
         : fn == E_PRESSURE ? E_DPRESSURE : fn == E_DPRESSURE ? E_DPRESSURE2 : fn == E_DPRESSURE2 ? E_D_DPRESSURE2
         : fn == E_INDIRECT ? E_D_INDIRECT : fn == E_CONSTANT ? E_D_CONSTANT : fn == E_SPIKY ? E_DSPIKY
         : fn == E_DSPIKY ? E_D_DSPIKY : fn == E_COHESION ? E_D_COHESION : fn == E_SIGMOID ? E_D_SIGMOID
         : E_D_CONSTANT;
}

template <int D_, int NL_, unsigned CS_, unsigned SS_, unsigned FS_, unsigned NS_>
struct Sig {
    static constexpr int D = D_, NL = NL_;
    static __host__ __device__ constexpr int FN(int l) { return (FS_ >> (4 * l)) & 15; }
    static __host__ __device__ constexpr int DFN(int l) { return deriv_expr_ct(FN(l)); }
    static __host__ __device__ constexpr int NORM(int l) { return (NS_ >> l) & 1; }
    static __host__ __device__ constexpr int SAL(int l)  // first layer with the same (kernel, dis_norm)
    {
        for (int m = 0; m < l; ++m)
            if (FN(m) == FN(l) && NORM(m) == NORM(l)) return m;
        return l;
    }
    static __host__ __device__ constexpr int C(int l) { return (CS_ >> (4 * l)) & 15; }
    static __host__ __device__ constexpr int S(int l) { return (SS_ >> (4 * l)) & 15; }
    static __host__ __device__ constexpr int nsrc()
    {
        int n = 0;
        for (int l = 0; l < NL_; ++l)
            if (S(l) != (int)kSrcLocs && S(l) + 1 > n) n = S(l) + 1;
        return n;
    }
    static __host__ __device__ constexpr int src_channels(int s)
    {
        for (int l = 0; l < NL_; ++l)
            if (S(l) == s) return C(l);
        return 0;
    }
    static __host__ __device__ constexpr int src_off(int s)  // record offset of distinct data tensor s
    {
        int o = D_;
        for (int t = 0; t < s; ++t) o += src_channels(t);
        return o;
    }
    static __host__ __device__ constexpr int data_off(int l) { return S(l) == (int)kSrcLocs ? 0 : src_off(S(l)); }
    static __host__ __device__ constexpr int ctot()
    {
        int n = 0;
        for (int l = 0; l < NL_; ++l) n += C(l);
        return n;
    }
    static __host__ __device__ constexpr int chan_off(int l)  // offset of layer l in the concatenated channel space
    {
        int n = 0;
        for (int t = 0; t < l; ++t) n += C(t);
        return n;
    }
    static __host__ __device__ constexpr int fwd_floats() { return src_off(nsrc()); }
    static __host__ __device__ constexpr int u_off(int l) { return fwd_floats() + chan_off(l); }
    static __host__ __device__ constexpr int bwd_floats() { return fwd_floats() + ctot(); }
    static __host__ __device__ constexpr int fwd_vec() { return (fwd_floats() + 3) / 4; }
    static __host__ __device__ constexpr int bwd_vec() { return (bwd_floats() + 3) / 4; }
};

struct LayerArgs {
    const float* data;      // [B,N,C]
    const float* weight;    // [O,C]
    const float* bias;      // [O] or NULL
    float* out;             // fwd: [B,N,O]
    const float* grad_out;  // bwd: [B,N,O]
    float* ddata;           // bwd: [B,N,C] or NULL
    int C, O;
    int w_expr, dw_expr, dis_norm;
    int salias;             // first layer with the same (kernel, dis_norm)
    float wc, dwc;
};
struct GroupArgs {
    LayerArgs l[kMaxLayers];
    const float* src[kMaxLayers];  // distinct data tensors
    float H, invH, H2, rad2;
};

struct SphF { float H, invH, H2; };

__device__ __forceinline__ float sph_fast(int e, float d, float d2, float c, const SphF& p)
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
    case E_SPIKY:     { const float q = fmaf(-d, p.invH, 1.0f); return c * q * q; }
    case E_DSPIKY:    return c * fmaf(-d, p.invH, 1.0f);
    case E_D_DSPIKY:  return c;
    case E_COHESION:  { const float t = d * p.invH; return fmaf(fmaf(-6.0f, t, 7.0f) * t, t, -1.0f); }
    case E_D_COHESION: return 2.0f * d * (7.0f * p.H - 9.0f * d) * (p.invH * p.invH * p.invH);
    case E_SIGMOID:   return 1.0f / (1.0f + expf((d - 0.2f * p.H) * 20.0f * p.invH));
    case E_D_SIGMOID: { const float ex = expf((d - 0.2f * p.H) * 20.0f * p.invH);
                        return -20.0f * ex * p.invH / ((ex + 1.0f) * (ex + 1.0f)); }
    default: return 0.0f;
    }
}

// ---- pack pre-pass ----------------------------------------------------------------------------------
// rec[n] = [ locs(D) | distinct data ... | (BWD) U_l(C_l) for every layer ], padded to float4s.
template <typename SG, bool BWD>
__global__ void __launch_bounds__(256)
k_group_pack(const float* __restrict__ locs, GroupArgs ga, long long BN, float* __restrict__ rec,
             const int* __restrict__ tile_flag, float* __restrict__ dlocs, const int* __restrict__ sym_flag)
{
    // planar (one float4 array per record quarter) for the tile kernels, record-major for the list walk
    const bool planar = tile_flag != nullptr && *tile_flag == 0 && !(BWD && sym_flag != nullptr && *sym_flag != 0);
    constexpr int V = BWD ? SG::bwd_vec() : SG::fwd_vec();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    float r[V * 4];
#pragma unroll
    for (int i = 0; i < V * 4; ++i) r[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < SG::D; ++k) r[k] = locs[n * SG::D + k];
#pragma unroll
    for (int s = 0; s < SG::nsrc(); ++s)
#pragma unroll
        for (int c = 0; c < SG::src_channels(s); ++c)
            r[SG::src_off(s) + c] = ga.src[s][n * SG::src_channels(s) + c];
    if (BWD) {
        if (!(sym_flag != nullptr && *sym_flag == 0)) {
            // the scatter mode of k_group_bwd accumulates with atomics: its targets start from zero
#pragma unroll
            for (int k = 0; k < SG::D; ++k) dlocs[n * SG::D + k] = 0.0f;
#pragma unroll
            for (int l = 0; l < SG::NL; ++l)
                if (ga.l[l].ddata) {
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c) ga.l[l].ddata[n * SG::C(l) + c] = 0.0f;
                }
        }
#pragma unroll
        for (int l = 0; l < SG::NL; ++l) {
            const LayerArgs& L = ga.l[l];
            for (int o = 0; o < L.O; ++o) {
                const float g = L.grad_out[n * L.O + o];
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c)
                    r[SG::u_off(l) + c] = fmaf(g, L.weight[o * SG::C(l) + c], r[SG::u_off(l) + c]);
            }
        }
    }
    float4* dst = reinterpret_cast<float4*>(rec) + (planar ? n : n * V);
    const long long vs = planar ? BN : 1;
#pragma unroll
    for (int v = 0; v < V; ++v) dst[v * vs] = make_float4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
}

// Per-layer kernel coefficients copied out of the parameter block once per thread (the tile kernels keep
// them in registers: re-reading the constant bank per pair costs ~3 issue slots per pair).
template <int NL>
struct LayerCoef {
    struct { float wc, dwc; } l[NL];
};
template <typename SG>
__device__ __forceinline__ LayerCoef<SG::NL> load_coef(const GroupArgs& ga)
{
    LayerCoef<SG::NL> c;
#pragma unroll
    for (int l = 0; l < SG::NL; ++l) {
        c.l[l].wc = ga.l[l].wc;
        c.l[l].dwc = ga.l[l].dwc;
    }
    return c;
}

// s_l = W_l(d) * norm_l (and t_l = dW_l/dd / d * norm_l) for every layer, evaluated once per distinct
// (kernel, dis_norm); everything about the layer list is a compile-time constant.  GA: GroupArgs or
// LayerCoef (anything with .l[l].wc / .dwc).
template <typename SG, bool WITH_T, typename GA>
__device__ __forceinline__ void layer_scales(const GA& ga, const SphF& sp, float d, float d2,
                                             float inv, bool pos, float* s, float* t)
{
#pragma unroll
    for (int l = 0; l < SG::NL; ++l) {
        if (SG::SAL(l) == l) {
            const float norm = (SG::NORM(l) && pos) ? inv : 1.0f;
            s[l] = sph_fast(SG::FN(l), d, d2, ga.l[l].wc, sp) * norm;
            if (WITH_T) t[l] = pos ? sph_fast(SG::DFN(l), d, d2, ga.l[l].dwc, sp) * inv * norm : 0.0f;
        } else {
            s[l] = s[SG::SAL(l)];
            if (WITH_T) t[l] = t[SG::SAL(l)];
        }
    }
}

// ---- forward ----------------------------------------------------------------------------------------
#ifndef SPNB_GROUP_FWD_MINB
#define SPNB_GROUP_FWD_MINB 1
#endif
template <typename SG, int THREADS>
__device__ __forceinline__ void group_fwd_block(const float* __restrict__ rec, const float* __restrict__ neighbors,
                                                const GroupArgs& ga, int N, int K, int bx, int b, int nbx, int nby,
                                                WalkSmem<SPNB_GROUP_FWD_G>* s_walk)  // one per warp of the block
{
    constexpr int D = SG::D, V = SG::fwd_vec(), CT = SG::ctot();
    constexpr int G = SPNB_GROUP_FWD_G, kU = SPNB_GROUP_FWD_U, QPB = THREADS / G, R = 32 / G;
    const int warp = threadIdx.x >> 5, sub = threadIdx.x % G;
    const int m = bx * QPB + threadIdx.x / G;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    const int m0 = bx * QPB + warp * R;  // first query of this warp
    const int nrows = min(R, max(0, N - m0));
    const SphF sp = {ga.H, ga.invH, ga.H2};
    const float4* srec = reinterpret_cast<const float4*>(rec) + (size_t)b * N * V;
    float x[D];
    {
        const float4 r0 = srec[(size_t)(active ? m : 0) * V];
        const float t[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = t[k];
    }
    float G_[CT];
#pragma unroll
    for (int i = 0; i < CT; ++i) G_[i] = 0.0f;
    const float* warp_rows = neighbors + ((size_t)b * N + min(m0, N - 1)) * K;
    prefetch_rows_ahead(neighbors, N, K, QPB, 2, bx, b, nbx, nby);

    walk_rows<G, kU>(warp_rows, K, nrows, s_walk[warp], [&](const int* j, const bool* valid) {
#if defined(SPNB_DEBUG_WALK_ONLY)
        // experiment: list walk without gathers or math (measures the row-staging floor)
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (valid[u]) G_[0] += (float)j[u];
        return;
#endif
        float r[kU][V * 4];
#pragma unroll
        for (int u = 0; u < kU; ++u)
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = srec[(unsigned)j[u] * (unsigned)V + v];
                r[u][4 * v] = t.x; r[u][4 * v + 1] = t.y; r[u][4 * v + 2] = t.z; r[u][4 * v + 3] = t.w;
            }
#if defined(SPNB_DEBUG_NO_MATH)
        // experiment: gathers without the pair math (measures the gather floor)
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (valid[u]) G_[0] += r[u][0] + r[u][V * 4 - 1];
        return;
#endif
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            float d2 = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float nr = x[k] - r[u][k];
                d2 += nr * nr;
            }
            if (valid[u] && d2 < ga.rad2) {
                const bool pos = d2 > 0.0f;
                const float inv = fast_rsqrt(d2);
                const float d = pos ? d2 * inv : 0.0f;
                float s[SG::NL];
                layer_scales<SG, false>(ga, sp, d, d2, inv, pos, s, nullptr);
#pragma unroll
                for (int l = 0; l < SG::NL; ++l)
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c)
                        G_[SG::chan_off(l) + c] = fmaf(s[l], r[u][SG::data_off(l) + c], G_[SG::chan_off(l) + c]);
            }
        }
    });
    // epilogue: apply the weights once per query (the G lanes of a group split the outputs)
    if (G > 1) {
#pragma unroll
        for (int i = 0; i < CT; ++i) G_[i] = group_sum<G>(G_[i]);
    }
    if (active) {
#pragma unroll
        for (int l = 0; l < SG::NL; ++l) {
            const LayerArgs& L = ga.l[l];
            for (int o = sub; o < L.O; o += G) {
                float v = L.bias ? L.bias[o] : 0.0f;
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c) v = fmaf(L.weight[o * SG::C(l) + c], G_[SG::chan_off(l) + c], v);
                L.out[q * L.O + o] = v;
            }
        }
    }
}

// The list walk as a kernel of its own (calls without tile lists).  With tile lists the same body is the
// device-side fallback inside the tile kernels (k_tile_fwd / k_tile_bwd), taken when the tile flag is set.
template <typename SG>
__global__ void __launch_bounds__(kThreads, SPNB_GROUP_FWD_MINB)
k_group_fwd(const float* __restrict__ rec, const float* __restrict__ neighbors, GroupArgs ga, int N, int K)
{
    __shared__ WalkSmem<SPNB_GROUP_FWD_G> s_walk[kThreads / 32];
    group_fwd_block<SG, kThreads>(rec, neighbors, ga, N, K, blockIdx.x, blockIdx.y, gridDim.x, gridDim.y, s_walk);
}

// ---- backward ---------------------------------------------------------------------------------------
// dlocs [B,N,D]: d(sum_l loss_l)/d(locs) through the geometry (query role + neighbour role).
// ddata_l [B,N,C_l] (may be NULL).  sym: gather; else scatter with atomics into zero-filled buffers.
#ifndef SPNB_GROUP_BWD_MINB
#define SPNB_GROUP_BWD_MINB 1
#endif
template <typename SG, int THREADS>
__device__ __forceinline__ void group_bwd_block(const float* __restrict__ rec, const float* __restrict__ neighbors,
                                                const GroupArgs& ga, int N, int K, float* dlocs, const int* sym_flag,
                                                int bx, int b, int nbx, int nby,
                                                WalkSmem<SPNB_GROUP_BWD_G>* s_walk)  // one per warp of the block
{
    constexpr int D = SG::D, V = SG::bwd_vec(), CT = SG::ctot();
    constexpr int G = SPNB_GROUP_BWD_G, UB = SPNB_GROUP_BWD_U, QPB = THREADS / G, R = 32 / G;
    const bool sym = sym_flag != nullptr && *sym_flag == 0;
    const int warp = threadIdx.x >> 5, sub = threadIdx.x % G;
    const int m = bx * QPB + threadIdx.x / G;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    const int m0 = bx * QPB + warp * R;
    const int nrows = min(R, max(0, N - m0));
    const SphF sp = {ga.H, ga.invH, ga.H2};
    const float4* srec = reinterpret_cast<const float4*>(rec) + (size_t)b * N * V;
    float me[V * 4];  // my own record: position, data_l[i], U_l[i]
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float4 t = srec[(size_t)(active ? m : 0) * V + v];
        me[4 * v] = t.x; me[4 * v + 1] = t.y; me[4 * v + 2] = t.z; me[4 * v + 3] = t.w;
    }
    float a_dl[D], a_dd[CT];
#pragma unroll
    for (int k = 0; k < D; ++k) a_dl[k] = 0.0f;
#pragma unroll
    for (int i = 0; i < CT; ++i) a_dd[i] = 0.0f;
    const float* warp_rows = neighbors + ((size_t)b * N + min(m0, N - 1)) * K;
    prefetch_rows_ahead(neighbors, N, K, QPB, 2, bx, b, nbx, nby);

    walk_rows<G, UB>(warp_rows, K, nrows, s_walk[warp], [&](const int* j, const bool* valid) {
        float r[UB][V * 4];
#pragma unroll
        for (int u = 0; u < UB; ++u)
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = srec[(unsigned)j[u] * (unsigned)V + v];
                r[u][4 * v] = t.x; r[u][4 * v + 1] = t.y; r[u][4 * v + 2] = t.z; r[u][4 * v + 3] = t.w;
            }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            float disp[D];
            float d2 = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                disp[k] = me[k] - r[u][k];
                d2 += disp[k] * disp[k];
            }
            if (valid[u] && d2 < ga.rad2) {
                const bool pos = d2 > 0.0f;
                const float inv = fast_rsqrt(d2);
                const float d = pos ? d2 * inv : 0.0f;
                float s[SG::NL], t[SG::NL];
                layer_scales<SG, true>(ga, sp, d, d2, inv, pos, s, t);
                float TA = 0.0f, TB = 0.0f;  // position-gradient coefficients of the two pair roles
#pragma unroll
                for (int l = 0; l < SG::NL; ++l) {
                    float A = 0.0f, Bv = 0.0f;
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c) {
                        // pair (i, j): U_l[i] . data_l[j]      pair (j, i): U_l[j] . data_l[i]
                        A = fmaf(me[SG::u_off(l) + c], r[u][SG::data_off(l) + c], A);
                        Bv = fmaf(r[u][SG::u_off(l) + c], me[SG::data_off(l) + c], Bv);
                        if (sym)
                            a_dd[SG::chan_off(l) + c] = fmaf(s[l], r[u][SG::u_off(l) + c], a_dd[SG::chan_off(l) + c]);
                    }
                    TA = fmaf(A, t[l], TA);
                    TB = fmaf(Bv, t[l], TB);
                }
                if (sym) {
                    const float T = TA + TB;
#pragma unroll
                    for (int k = 0; k < D; ++k) a_dl[k] = fmaf(T, disp[k], a_dl[k]);
                } else {
                    const size_t jo = (size_t)b * N + j[u];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        a_dl[k] = fmaf(TA, disp[k], a_dl[k]);
                        if (pos) atomicAdd(dlocs + jo * D + k, -TA * disp[k]);
                    }
#pragma unroll
                    for (int l = 0; l < SG::NL; ++l) {
                        if (ga.l[l].ddata) {
#pragma unroll
                            for (int c = 0; c < SG::C(l); ++c)
                                atomicAdd(ga.l[l].ddata + jo * SG::C(l) + c, s[l] * me[SG::u_off(l) + c]);
                        }
                    }
                }
            }
        }
    });
    if (G > 1) {
#pragma unroll
        for (int k = 0; k < D; ++k) a_dl[k] = group_sum<G>(a_dl[k]);
        if (sym) {
#pragma unroll
            for (int i = 0; i < CT; ++i) a_dd[i] = group_sum<G>(a_dd[i]);
        }
    }
    if (active && sub == 0) {
        if (sym) {
#pragma unroll
            for (int k = 0; k < D; ++k) dlocs[q * D + k] = a_dl[k];
#pragma unroll
            for (int l = 0; l < SG::NL; ++l) {
                if (ga.l[l].ddata) {
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c) ga.l[l].ddata[q * SG::C(l) + c] = a_dd[SG::chan_off(l) + c];
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) atomicAdd(dlocs + q * D + k, a_dl[k]);
        }
    }
}


template <typename SG>
__global__ void __launch_bounds__(kThreads, SPNB_GROUP_BWD_MINB)
k_group_bwd(const float* __restrict__ rec, const float* __restrict__ neighbors, GroupArgs ga, int N, int K,
            float* dlocs, const int* sym_flag)
{
    __shared__ WalkSmem<SPNB_GROUP_BWD_G> s_walk[kThreads / 32];
    group_bwd_block<SG, kThreads>(rec, neighbors, ga, N, K, dlocs, sym_flag, blockIdx.x, blockIdx.y, gridDim.x,
                                  gridDim.y, s_walk);
}

// ---- tile-list kernels --------------------------------------------------------------------------------
// Same math as k_group_fwd / k_group_bwd (symmetric mode), but driven by the compact tile lists of
// tile_lists.cuh: one block per tile block of 64 queries; the planar records of the block's candidate
// ranges are staged in shared memory by TMA bulk copies (one cp.async.bulk per range and record quarter,
// completion on an mbarrier), the 16-bit lists are read in coalesced 32-byte units, and every neighbour
// gather is an LDS.128 instead of an L1/L2 round trip.
struct TileArgs {
    const int* flag;
    const TileDesc* descs;
    const int* counts;
    const unsigned char* lists;
    int ntb;
};

template <int V>
__device__ __forceinline__ void tile_stage(const float4* __restrict__ planes, size_t plane_stride, size_t scene_off,
                                           const TileDesc& d, float4* s_rec, unsigned long long* bar)
{
    // warp 0: lane 0 arms the barrier with the byte count, then the lanes issue the copies
    const int lane = threadIdx.x;
    const bool fits = d.total + 1 <= kTileCap;  // else: nothing is staged, the block gathers from global memory
    if (lane == 0) mbar_expect_tx(bar, fits ? (unsigned)d.total * 16u * V : 0u);
    __syncwarp();
    const int ncopies = fits ? d.nr * V : 0;
    for (int i = lane; i < ncopies; i += 32) {
        const int r = i / V, v = i % V;
        const int len = d.prefix[r + 1] - d.prefix[r];
        if (len > 0)
            bulk_copy_g2s(s_rec + (size_t)v * kTileCap + 1 + d.prefix[r],
                          planes + (size_t)v * plane_stride + scene_off + d.start[r], (unsigned)len * 16u, bar);
    }
}

// sorted particle index of staged slot `slot` (>= 1) -- only the rare blocks whose tile does not fit
// kTileCap use this (they gather from global memory instead of the staged copy)
__device__ __forceinline__ int tile_slot_to_index(const TileDesc& d, unsigned slot)
{
    const int s = (int)slot - 1;
    int idx = 0;
#pragma unroll
    for (int r = 0; r < kTileMaxRanges; ++r)
        if (r < d.nr && s >= d.prefix[r]) idx = d.start[r] + s - d.prefix[r];
    return idx;
}

template <int G>
struct TileUnit {
    static constexpr int WORDS = 8 / G;  // 32-bit words (2 entries each) per lane and unit
    unsigned w[WORDS];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < WORDS; ++i) w[i] = 0u;
    }
    __device__ __forceinline__ void load(const unsigned char* p)
    {
        if (WORDS == 8) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4 % WORDS] = b.x; w[5 % WORDS] = b.y; w[6 % WORDS] = b.z; w[7 % WORDS] = b.w;
        } else if (WORDS == 4) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
            w[0] = a.x; w[1] = a.y; w[2 % WORDS] = a.z; w[3 % WORDS] = a.w;
        } else {
            const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
            w[0] = a.x; w[1 % WORDS] = a.y;
        }
    }
};

// Calls body(slot * 16) -- list entries are stored as byte offsets into a record plane -- for every list
// entry of this lane's share of its query's list (sentinel slots included: they fail every radius
// test).  Units are prefetched SPNB_TILE_PF ahead.
#ifndef SPNB_TILE_PF
#define SPNB_TILE_PF 1  // list units fetched ahead of the one being consumed
#endif
// 1: no branch around the pair math (an out-of-radius / sentinel pair contributes through zeroed kernel
// values); lets the compiler overlap the shared-memory gathers of later entries with the math of earlier ones
#ifndef SPNB_TILE_BRANCHFREE
#define SPNB_TILE_BRANCHFREE 0
#endif
template <int G, typename Body>
__device__ __forceinline__ void tile_walk(const unsigned char* __restrict__ my_units, int cnt, Body body)
{
    constexpr int PF = SPNB_TILE_PF;
    const int nunits = (cnt + kTileUnit - 1) / kTileUnit;
    const int wmax = __reduce_max_sync(0xffffffffu, nunits);
    TileUnit<G> q[PF + 1];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
        q[i].clear();
        if (i < nunits) q[i].load(my_units + (size_t)i * 256);
    }
    for (int u = 0; u < wmax; ++u) {
        q[PF].clear();
        if (u + PF < nunits) q[PF].load(my_units + (size_t)(u + PF) * 256);
        // even words: entries 0..7 of the unit (4x4-transposed storage, tile_lists.cuh)
#pragma unroll
        for (int i = 0; i < TileUnit<G>::WORDS; i += 2) {
            body(q[0].w[i] & 0xffffu);
            body(q[0].w[i] >> 16);
        }
        // odd words: entries 8..15 -- skipped when no query of the warp has that many left
        if (__any_sync(0xffffffffu, cnt - u * kTileUnit > kTileUnit / 2)) {
#pragma unroll
            for (int i = 1; i < TileUnit<G>::WORDS; i += 2) {
                body(q[0].w[i] & 0xffffu);
                body(q[0].w[i] >> 16);
            }
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) q[i] = q[i + 1];
    }
}

template <typename SG, int G>
__global__ void __launch_bounds__(kTileQ * G)
k_tile_fwd(const float* __restrict__ rec, TileArgs ta, GroupArgs ga, int N, int K, long long BN,
           const float* __restrict__ neighbors)
{
    constexpr int D = SG::D, V = SG::fwd_vec(), CT = SG::ctot();
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_rec = reinterpret_cast<float4*>(s_raw);
    __shared__ TileDesc s_desc;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, tb = blockIdx.x, b = blockIdx.y;
    // the descriptor load is issued together with the flag load (both are inputs of this call's
    // predecessors only), so the flag test does not add a global-memory latency to the prologue
    int desc_word = 0;
    if (tid < 32) desc_word = __ldg(reinterpret_cast<const int*>(ta.descs + (size_t)b * ta.ntb + tb) + tid);
    if (*ta.flag != 0) {
        // tile lists unusable for this call (a list reached K, ...): the float-list walk, strided over the
        // grid, with the (unused) tile buffer as its row-staging scratch; records are record-major then
        constexpr int THREADS = kTileQ * G;
        const int nbx = (int)(((long long)N * SPNB_GROUP_FWD_G + THREADS - 1) / THREADS), B = gridDim.y;
        for (int t = blockIdx.y * gridDim.x + blockIdx.x; t < nbx * B; t += gridDim.x * gridDim.y) {
            group_fwd_block<SG, THREADS>(rec, neighbors, ga, N, K, t % nbx, t / nbx, nbx, B,
                                         reinterpret_cast<WalkSmem<SPNB_GROUP_FWD_G>*>(s_raw));
            __syncthreads();
        }
        return;
    }
    if (tid < 32) reinterpret_cast<int*>(&s_desc)[tid] = desc_word;
    if (tid == 0) mbar_init(&s_bar, 1);
    if (tid < V) s_rec[tid * kTileCap] = tid == 0 ? make_float4(1e18f, 1e18f, 1e18f, 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();
    const float4* planes = reinterpret_cast<const float4*>(rec);
    if (tid < 32) tile_stage<V>(planes, (size_t)BN, (size_t)b * N, s_desc, s_rec, &s_bar);

    const int ql = tid / G, sub = tid % G;
    const int m = tb * kTileQ + ql;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    const SphF sp = {ga.H, ga.invH, ga.H2};
    float x[D];
    {
        const float4 r0 = planes[q];
        const float t[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = t[k];
    }
    const int cnt = active ? ta.counts[q] : 0;
    const unsigned char* my_units = ta.lists + tile_entry_off(ta.ntb, K, b, tb, ql, 0) + sub * (32 / G);
    float G_[CT];
#pragma unroll
    for (int i = 0; i < CT; ++i) G_[i] = 0.0f;
    mbar_wait(&s_bar, 0);

    const LayerCoef<SG::NL> co = load_coef<SG>(ga);
    const float rad2 = ga.rad2;
    auto pair = [&](const float* r) {
        float d2 = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float nr = x[k] - r[k];
            d2 += nr * nr;
        }
        const bool in = d2 < rad2;
        if (SPNB_TILE_BRANCHFREE || in) {
            const bool pos = d2 > 0.0f;
            const float dd = SPNB_TILE_BRANCHFREE ? fminf(d2, rad2) : d2;  // keeps the sentinel's 1e36 out of W
            const float inv = fast_rsqrt(dd);
            const float d = pos ? dd * inv : 0.0f;
            float s[SG::NL];
            layer_scales<SG, false>(co, sp, d, dd, inv, pos, s, nullptr);
#pragma unroll
            for (int l = 0; l < SG::NL; ++l) {
                if (SPNB_TILE_BRANCHFREE && SG::SAL(l) == l) s[l] = in ? s[l] : 0.0f;
                if (SPNB_TILE_BRANCHFREE && SG::SAL(l) != l) s[l] = s[SG::SAL(l)];
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c)
                    G_[SG::chan_off(l) + c] = fmaf(s[l], r[SG::data_off(l) + c], G_[SG::chan_off(l) + c]);
            }
        }
    };
    if (s_desc.total + 1 <= kTileCap) {
        tile_walk<G>(my_units, cnt, [&](unsigned slot) {
            float r[V * 4];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(s_raw + v * (kTileCap * 16) + slot);
                r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
            }
            pair(r);
        });
    } else {
        tile_walk<G>(my_units, cnt, [&](unsigned slot) {
            if (slot == 0) return;
            const size_t j = (size_t)b * N + tile_slot_to_index(s_desc, slot >> 4);
            float r[V * 4];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = planes[(size_t)v * BN + j];
                r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
            }
            pair(r);
        });
    }
    if (G > 1) {
#pragma unroll
        for (int i = 0; i < CT; ++i) G_[i] = group_sum<G>(G_[i]);
    }
    if (active) {
#pragma unroll
        for (int l = 0; l < SG::NL; ++l) {
            const LayerArgs& L = ga.l[l];
            for (int o = sub; o < L.O; o += G) {
                float v = L.bias ? L.bias[o] : 0.0f;
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c) v = fmaf(L.weight[o * SG::C(l) + c], G_[SG::chan_off(l) + c], v);
                L.out[q * L.O + o] = v;
            }
        }
    }
}

// (min blocks: the staged tile bounds the residency at 3 CTAs per SM for four-plane records and 4 for
//  three-plane ones; keep the registers of the merged tile + list-walk code within that)
template <typename SG, int G>
__global__ void __launch_bounds__(kTileQ * G, (kTileQ * G >= 256 ? (SG::bwd_vec() <= 3 ? 4 : 3) : 1))
k_tile_bwd(const float* __restrict__ rec, TileArgs ta, GroupArgs ga, int N, int K, long long BN, float* dlocs,
           const float* __restrict__ neighbors, const int* sym_flag)
{
    constexpr int D = SG::D, V = SG::bwd_vec(), CT = SG::ctot();
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_rec = reinterpret_cast<float4*>(s_raw);
    __shared__ TileDesc s_desc;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, tb = blockIdx.x, b = blockIdx.y;
    // the descriptor load is issued together with the flag load (both are inputs of this call's
    // predecessors only), so the flag test does not add a global-memory latency to the prologue
    int desc_word = 0;
    if (tid < 32) desc_word = __ldg(reinterpret_cast<const int*>(ta.descs + (size_t)b * ta.ntb + tb) + tid);
    if (*ta.flag != 0 || (sym_flag != nullptr && *sym_flag != 0)) {
        // tile lists unusable for this call, or the relation is not symmetric (the tile path only has the
        // gather mode): the float-list walk (gather or atomics mode by sym_flag)
        constexpr int THREADS = kTileQ * G;
        const int nbx = (int)(((long long)N * SPNB_GROUP_BWD_G + THREADS - 1) / THREADS), B = gridDim.y;
        for (int t = blockIdx.y * gridDim.x + blockIdx.x; t < nbx * B; t += gridDim.x * gridDim.y) {
            group_bwd_block<SG, THREADS>(rec, neighbors, ga, N, K, dlocs, sym_flag, t % nbx, t / nbx, nbx, B,
                                         reinterpret_cast<WalkSmem<SPNB_GROUP_BWD_G>*>(s_raw));
            __syncthreads();
        }
        return;
    }
    if (tid < 32) reinterpret_cast<int*>(&s_desc)[tid] = desc_word;
    if (tid == 0) mbar_init(&s_bar, 1);
    if (tid < V) s_rec[tid * kTileCap] = tid == 0 ? make_float4(1e18f, 1e18f, 1e18f, 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();
    const float4* planes = reinterpret_cast<const float4*>(rec);
    if (tid < 32) tile_stage<V>(planes, (size_t)BN, (size_t)b * N, s_desc, s_rec, &s_bar);

    const int ql = tid / G, sub = tid % G;
    const int m = tb * kTileQ + ql;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    const SphF sp = {ga.H, ga.invH, ga.H2};
    float me[V * 4];  // my own record: position, data_l[i], U_l[i]
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float4 t = planes[(size_t)v * BN + q];
        me[4 * v] = t.x; me[4 * v + 1] = t.y; me[4 * v + 2] = t.z; me[4 * v + 3] = t.w;
    }
    const int cnt = active ? ta.counts[q] : 0;
    const unsigned char* my_units = ta.lists + tile_entry_off(ta.ntb, K, b, tb, ql, 0) + sub * (32 / G);
    float a_dl[D], a_dd[CT];
#pragma unroll
    for (int k = 0; k < D; ++k) a_dl[k] = 0.0f;
#pragma unroll
    for (int i = 0; i < CT; ++i) a_dd[i] = 0.0f;
    mbar_wait(&s_bar, 0);

    const LayerCoef<SG::NL> co = load_coef<SG>(ga);
    const float rad2 = ga.rad2;
    auto pair = [&](const float* r) {
        float disp[D];
        float d2 = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            disp[k] = me[k] - r[k];
            d2 += disp[k] * disp[k];
        }
        const bool in = d2 < rad2;
        if (SPNB_TILE_BRANCHFREE || in) {
            const bool pos = d2 > 0.0f;
            const float dd = SPNB_TILE_BRANCHFREE ? fminf(d2, rad2) : d2;  // keeps the sentinel's 1e36 out of W
            const float inv = fast_rsqrt(dd);
            const float d = pos ? dd * inv : 0.0f;
            float s[SG::NL], t[SG::NL];
            layer_scales<SG, true>(co, sp, d, dd, inv, pos, s, t);
            if (SPNB_TILE_BRANCHFREE) {
#pragma unroll
                for (int l = 0; l < SG::NL; ++l) {
                    s[l] = SG::SAL(l) == l ? (in ? s[l] : 0.0f) : s[SG::SAL(l)];
                    t[l] = SG::SAL(l) == l ? (in ? t[l] : 0.0f) : t[SG::SAL(l)];
                }
            }
            float T = 0.0f;
#pragma unroll
            for (int l = 0; l < SG::NL; ++l) {
                float AB = 0.0f;
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c) {
                    // pair (i, j): U_l[i] . data_l[j]   +   pair (j, i): U_l[j] . data_l[i]
                    AB = fmaf(me[SG::u_off(l) + c], r[SG::data_off(l) + c], AB);
                    AB = fmaf(r[SG::u_off(l) + c], me[SG::data_off(l) + c], AB);
                    a_dd[SG::chan_off(l) + c] = fmaf(s[l], r[SG::u_off(l) + c], a_dd[SG::chan_off(l) + c]);
                }
                T = fmaf(AB, t[l], T);
            }
#pragma unroll
            for (int k = 0; k < D; ++k) a_dl[k] = fmaf(T, disp[k], a_dl[k]);
        }
    };
    if (s_desc.total + 1 <= kTileCap) {
        tile_walk<G>(my_units, cnt, [&](unsigned slot) {
            float r[V * 4];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(s_raw + v * (kTileCap * 16) + slot);
                r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
            }
            pair(r);
        });
    } else {
        tile_walk<G>(my_units, cnt, [&](unsigned slot) {
            if (slot == 0) return;
            const size_t j = (size_t)b * N + tile_slot_to_index(s_desc, slot >> 4);
            float r[V * 4];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = planes[(size_t)v * BN + j];
                r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
            }
            pair(r);
        });
    }
    if (G > 1) {
#pragma unroll
        for (int k = 0; k < D; ++k) a_dl[k] = group_sum<G>(a_dl[k]);
#pragma unroll
        for (int i = 0; i < CT; ++i) a_dd[i] = group_sum<G>(a_dd[i]);
    }
    if (active && sub == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) dlocs[q * D + k] = a_dl[k];
#pragma unroll
        for (int l = 0; l < SG::NL; ++l) {
            if (ga.l[l].ddata) {
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c) ga.l[l].ddata[q * SG::C(l) + c] = a_dd[SG::chan_off(l) + c];
            }
        }
    }
}

// ---- host -------------------------------------------------------------------------------------------
struct Signature {
    int D, NL;
    unsigned CS, SS, FS, NS;
};

// Distinct-data numbering and (kernel, dis_norm) aliases of a host-side layer list.
static bool make_signature(const float* locs, int D, int nl, const SpnbGroupLayer* layers, Signature& sg,
                           GroupArgs& ga, float radius)
{
    if (nl < 1 || nl > kMaxLayers) return false;
    sg.D = D;
    sg.NL = nl;
    sg.CS = sg.SS = sg.FS = sg.NS = 0;
    int nsrc = 0;
    for (int l = 0; l < nl; ++l) {
        const SpnbGroupLayer& L = layers[l];
        if (L.nchannels < 1 || L.nchannels > 4 || L.nkernels < 1) return false;
        unsigned s = kSrcLocs;
        if (L.data != locs || L.nchannels != D) {
            int found = -1;
            for (int t = 0; t < nsrc; ++t)
                if (ga.src[t] == L.data) found = t;
            if (found < 0) {
                // a distinct tensor must have the same channel count for every layer that uses it
                found = nsrc;
                ga.src[nsrc++] = L.data;
            }
            s = (unsigned)found;
        }
        sg.CS |= (unsigned)L.nchannels << (4 * l);
        sg.SS |= s << (4 * l);
        sg.FS |= (unsigned)L.kernel_fn << (4 * l);
        sg.NS |= (L.dis_norm ? 1u : 0u) << l;
        const SphParams p = make_sph_params(L.kernel_fn, radius);
        LayerArgs& A = ga.l[l];
        A.data = L.data; A.weight = L.weight; A.bias = L.bias; A.out = L.out;
        A.grad_out = L.grad_out; A.ddata = L.ddata;
        A.C = L.nchannels; A.O = L.nkernels;
        A.w_expr = p.w_expr; A.dw_expr = p.dw_expr; A.dis_norm = L.dis_norm ? 1 : 0;
        double wc = p.w_coef, dwc = p.dw_coef;
        if (p.w_expr == E_DSPIKY) wc /= (double)radius;
        if (p.dw_expr == E_DSPIKY) dwc /= (double)radius;
        A.wc = (float)wc; A.dwc = (float)dwc;
        A.salias = l;
        for (int m = 0; m < l; ++m)
            if (ga.l[m].w_expr == A.w_expr && ga.l[m].dis_norm == A.dis_norm) { A.salias = m; break; }
    }
    // same tensor used with different channel counts -> not representable
    for (int l = 0; l < nl; ++l)
        for (int m = 0; m < l; ++m)
            if (((sg.SS >> (4 * l)) & 15) == ((sg.SS >> (4 * m)) & 15) && ((sg.SS >> (4 * l)) & 15) != kSrcLocs &&
                layers[l].nchannels != layers[m].nchannels)
                return false;
    ga.H = radius; ga.invH = 1.0f / radius; ga.H2 = radius * radius; ga.rad2 = radius * radius;
    return true;
}

// The instantiated signatures: X(D, NL, CS, SS, FS, NS); layer 0 is the lowest nibble / bit.  Kernel
// ids: cohesion 0, constant 1, dspiky 7, spiky 0xB (kernels.py:123).  These are the layer groups of
// the fluid step (fluid_sim.py:367-397,419-420), for ndim 3 and 2:
//   A: spiky1(ones) dspikyD*(locs) dspiky1*(ones) cohesionD*(locs) cohesion1*(ones) constant1(ones)
//   B: dspikyD*(locs*pressure) dspiky1*(pressure)      V: spikyD(vel) spiky1(ones)
//   C: constantD(normals)                               (* = dis_norm)
#define SPNB_GROUP_SIGS(X)                                  \
    X(3, 6, 0x113131u, 0x00F0F0u, 0x10077Bu, 0x1Eu)         \
    X(3, 2, 0x13u, 0x10u, 0x77u, 0x3u)                      \
    X(3, 2, 0x13u, 0x10u, 0xBBu, 0x0u)                      \
    X(3, 1, 0x3u, 0x0u, 0x1u, 0x0u)                         \
    X(2, 6, 0x112121u, 0x00F0F0u, 0x10077Bu, 0x1Eu)         \
    X(2, 2, 0x12u, 0x10u, 0x77u, 0x3u)                      \
    X(2, 2, 0x12u, 0x10u, 0xBBu, 0x0u)                      \
    X(2, 1, 0x2u, 0x0u, 0x1u, 0x0u)

static bool make_tile_args(const void* tiles, int B, int N, int K, TileArgs& ta)
{
    if (!tiles) return false;
    const TileLayout tl = tile_layout(B, N, K);
    const unsigned char* base = (const unsigned char*)tiles;
    ta.flag = (const int*)base;
    ta.descs = (const TileDesc*)(base + tl.desc_off);
    ta.counts = (const int*)(base + tl.cnt_off);
    ta.lists = base + tl.list_off;
    ta.ntb = tl.ntb;
    return true;
}

static bool launched(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return true;
    set_error("convsp group: launch of %s failed: %s", what, cudaGetErrorString(e));
    return false;
}

template <typename KernelT>
static bool allow_smem(KernelT* kernel, size_t bytes)
{
    // static + dynamic shared memory above 48 KB needs the opt-in
    if (bytes + 1024 > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(max dynamic smem %zu): %s", bytes, cudaGetErrorString(e));
            return false;
        }
    }
    return true;
}

// With tile lists: pack (layout chosen on the device by the tile flag) + the tile kernel, which runs the
// float-list walk itself when the flag is set.  Without: pack + list walk.
template <typename SG>
static int run_fwd(const float* locs, const float* neighbors, const GroupArgs& ga, int B, int N, int K,
                   float* rec, const void* tiles, cudaStream_t stream)
{
    const long long BN = (long long)B * N;
    TileArgs ta;
    const bool tiled = make_tile_args(tiles, B, N, K, ta);
    k_group_pack<SG, false><<<cdiv(BN, 256), 256, 0, stream>>>(locs, ga, BN, rec, tiled ? ta.flag : nullptr, nullptr, nullptr);
    if (!launched("k_group_pack")) return -1;
    if (tiled) {
        constexpr int G = SPNB_TILE_FWD_G > 0 ? SPNB_TILE_FWD_G : (SG::fwd_vec() == 1 ? 1 : 2);
        static_assert(sizeof(WalkSmem<SPNB_GROUP_FWD_G>) * (kTileQ * G / 32) <= kTileCap * sizeof(float4), "walk scratch fits the tile buffer");
        const size_t smem = (size_t)SG::fwd_vec() * kTileCap * sizeof(float4);
        if (!allow_smem(k_tile_fwd<SG, G>, smem)) return -1;
        k_tile_fwd<SG, G><<<dim3(ta.ntb, B), kTileQ * G, smem, stream>>>(rec, ta, ga, N, K, BN, neighbors);
        if (!launched("k_tile_fwd")) return -1;
    } else {
        k_group_fwd<SG><<<dim3(cdiv((long long)N * SPNB_GROUP_FWD_G, kThreads), B), kThreads, 0, stream>>>(rec, neighbors, ga, N, K);
    }
    return 2;
}
template <typename SG>
static int run_bwd(const float* locs, const float* neighbors, const GroupArgs& ga, int B, int N, int K,
                   float* rec, float* dlocs, const int* sym_flag, const void* tiles, cudaStream_t stream)
{
    const long long BN = (long long)B * N;
    TileArgs ta;
    const bool tiled = make_tile_args(tiles, B, N, K, ta);
    k_group_pack<SG, true><<<cdiv(BN, 256), 256, 0, stream>>>(locs, ga, BN, rec, tiled ? ta.flag : nullptr, dlocs, sym_flag);
    if (!launched("k_group_pack")) return -1;
    if (tiled) {
        constexpr int G = SPNB_TILE_BWD_G;
        static_assert(sizeof(WalkSmem<SPNB_GROUP_BWD_G>) * (kTileQ * G / 32) <= kTileCap * sizeof(float4), "walk scratch fits the tile buffer");
        const size_t smem = (size_t)SG::bwd_vec() * kTileCap * sizeof(float4);
        if (!allow_smem(k_tile_bwd<SG, G>, smem)) return -1;
        k_tile_bwd<SG, G><<<dim3(ta.ntb, B), kTileQ * G, smem, stream>>>(rec, ta, ga, N, K, BN, dlocs, neighbors, sym_flag);
        if (!launched("k_tile_bwd")) return -1;
    } else {
        k_group_bwd<SG><<<dim3(cdiv((long long)N * SPNB_GROUP_BWD_G, kThreads), B), kThreads, 0, stream>>>(rec, neighbors, ga, N, K, dlocs, sym_flag);
    }
    return 2;
}

static size_t record_floats(const Signature& sg, bool bwd)
{
    size_t n = 0;
#define X(DD, NN, CC, SS_, FF, NO)                                                                 \
    if (sg.D == DD && sg.NL == NN && sg.CS == CC && sg.SS == SS_ && sg.FS == FF && sg.NS == NO)    \
        n = 4 * (size_t)(bwd ? Sig<DD, NN, CC, SS_, FF, NO>::bwd_vec() : Sig<DD, NN, CC, SS_, FF, NO>::fwd_vec());
    SPNB_GROUP_SIGS(X)
#undef X
    return n;
}

}  // namespace
}  // namespace spnb

using namespace spnb;

extern "C" {

size_t spnb_convsp_group_workspace_bytes(const float* locs, int batch_size, int N, int ndims, float radius,
                                         int nlayers, const SpnbGroupLayer* layers, int backward)
{
    Signature sg;
    GroupArgs ga;
    memset(&ga, 0, sizeof(ga));
    if (!layers || !make_signature(locs, ndims, nlayers, layers, sg, ga, radius)) return 0;
    return sizeof(float) * record_floats(sg, backward != 0) * (size_t)batch_size * N;
}

int spnb_convsp_group_forward(const float* locs, const float* neighbors, int B, int N, int D, int K,
                              float radius, int nlayers, const SpnbGroupLayer* layers, void* workspace,
                              size_t workspace_bytes, const void* tile_lists, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Signature sg;
    GroupArgs ga;
    memset(&ga, 0, sizeof(ga));
    if (!locs || !neighbors || !layers || B <= 0 || N <= 0 || K <= 0) {
        set_error("spnb_convsp_group_forward: bad arguments");
        return 0;
    }
    for (int l = 0; l < nlayers; ++l)
        if (!layers[l].data || !layers[l].weight || !layers[l].out || layers[l].kernel_fn < 0 ||
            layers[l].kernel_fn >= SPNB_NUM_KERNEL_FNS) {
            set_error("spnb_convsp_group_forward: layer %d: null pointer or bad kernel id", l);
            return 0;
        }
    const size_t need = make_signature(locs, D, nlayers, layers, sg, ga, radius)
                            ? sizeof(float) * record_floats(sg, false) * (size_t)B * N : 0;
    if (need == 0) {
        set_error("spnb_convsp_group_forward: unsupported group signature");
        return 0;
    }
    if (!workspace || workspace_bytes < need) {
        set_error("spnb_convsp_group_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return 0;
    }
#define X(DD, NN, CC, SS_, FF, NO)                                                                 \
    if (sg.D == DD && sg.NL == NN && sg.CS == CC && sg.SS == SS_ && sg.FS == FF && sg.NS == NO)    \
        nl = run_fwd<Sig<DD, NN, CC, SS_, FF, NO>>(locs, neighbors, ga, B, N, K, (float*)workspace, tile_lists, stream);
    int nl = 0;
    SPNB_GROUP_SIGS(X)
#undef X
    if (nl < 0) return 0;
    count_launches(nl);
    return check_launch("spnb_convsp_group_forward") ? 1 : 0;
}

int spnb_convsp_group_backward(const float* locs, const float* neighbors, int B, int N, int D, int K,
                               float radius, int nlayers, const SpnbGroupLayer* layers, float* dlocs,
                               const int* sym_flag, void* workspace, size_t workspace_bytes,
                               const void* tile_lists, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Signature sg;
    GroupArgs ga;
    memset(&ga, 0, sizeof(ga));
    if (!locs || !neighbors || !layers || !dlocs || B <= 0 || N <= 0 || K <= 0) {
        set_error("spnb_convsp_group_backward: bad arguments");
        return 0;
    }
    for (int l = 0; l < nlayers; ++l)
        if (!layers[l].data || !layers[l].weight || !layers[l].grad_out || layers[l].kernel_fn < 0 ||
            layers[l].kernel_fn >= SPNB_NUM_KERNEL_FNS) {
            set_error("spnb_convsp_group_backward: layer %d: null pointer or bad kernel id", l);
            return 0;
        }
    const size_t need = make_signature(locs, D, nlayers, layers, sg, ga, radius)
                            ? sizeof(float) * record_floats(sg, true) * (size_t)B * N : 0;
    if (need == 0) {
        set_error("spnb_convsp_group_backward: unsupported group signature");
        return 0;
    }
    if (!workspace || workspace_bytes < need) {
        set_error("spnb_convsp_group_backward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return 0;
    }
    // (the pack pre-pass zero-fills the scatter targets when the atomics mode is going to run)
#define X(DD, NN, CC, SS_, FF, NO)                                                                 \
    if (sg.D == DD && sg.NL == NN && sg.CS == CC && sg.SS == SS_ && sg.FS == FF && sg.NS == NO)    \
        nl = run_bwd<Sig<DD, NN, CC, SS_, FF, NO>>(locs, neighbors, ga, B, N, K, (float*)workspace, dlocs, sym_flag, tile_lists, stream);
    int nl = 0;
    SPNB_GROUP_SIGS(X)
#undef X
    if (nl < 0) return 0;
    count_launches(nl);
    return check_launch("spnb_convsp_group_backward") ? 1 : 0;
}

}  // extern "C"
