// Fused ConvSP "group": several ConvSP layers that share (locs, neighbors, radius) -- and have
// kernel_size 1 -- evaluated in ONE walk over the neighbour lists.  Also the fast path of a single such layer.
//
// This is SURVEY.md section 8(f) rank 1: the solver iteration of the reference's fluid simulation calls
// 9 ConvSP layers on the same particle set (examples/fluid_sim.py:367-397); each per-layer kernel
// re-reads the neighbour list, re-gathers the neighbour positions and recomputes the distance.  The
// math per layer is unchanged (compute_kernel_cells, src/common_funcs.h:439-583, ncells = 1):
//     out_l[i,o] = bias_l[o] + sum_c w_l[o,c] * sum_j W_l(d_ij) * norm_l(d_ij) * data_l[j,c]
//
// Design:
//  * a PACK pre-pass writes one record per particle as planes of float4: position, the distinct data
//    tensors of the group (a data tensor that IS the position tensor is not duplicated, a data tensor of ones
//    is not stored at all) and, for the backward pass, U_l[n,c] = sum_o grad_out_l[n,o] * w_l[o,c] of every
//    layer that takes part in it.  What a layer needs is known at compile time: a layer whose kernel has
//    dW/dd == 0 ("constant") contributes nothing to d/dlocs, a layer whose data needs no gradient has no
//    ddata -- so e.g. the 6-layer group of the fluid solver packs 12 floats (3 planes) instead of 14 (4);
//  * TILE kernels (tile_lists.cuh): one CTA per block of 64 queries; the planes of the block's candidate
//    ranges are staged in shared memory by TMA bulk copies, the block's 16-bit lists stream through shared
//    memory with cp.async, every neighbour gather is an LDS.128; queries are walked in list-length rank order;
//  * when the tile lists are unusable (flag set on the device: a list was cut at K, ...) or absent, the same
//    kernel walks the float lists with list_walk.cuh instead -- decided on the device, no host round trip;
//  * per pair the geometry is computed once, W / dW once per distinct (kernel, dis_norm), and only C_l FMAs
//    per layer are spent on channels because the weights are applied once per query in the epilogue
//    (forward) or folded into U_l (backward);
//  * backward: symmetric-gather (no atomics) when the device flag allows, else scatter with float atomics --
//    same rule as the per-layer kernels.  d(weight) is not produced here.
//
// The layer list of a group is a compile-time signature (template parameters), so every record field and
// accumulator lives in a register.  convsp_group_inst*.cu instantiate the signatures (the fluid step's groups
// and the single-layer family); anything else reports "unsupported" and the caller uses the per-layer path.
#pragma once
#include <string.h>

#include "list_walk.cuh"
#include "spnb_common.cuh"
#include "tile_lists.cuh"

namespace spnb {
namespace grp {

// lanes per query / list entries in flight per lane of the float-list walk (list_walk.cuh)
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
// tile kernels: lanes per query (0 = by record width: 1 plane -> 2 lanes, else 4)
#ifndef SPNB_TILE_FWD_G
#define SPNB_TILE_FWD_G 0
#endif
#ifndef SPNB_TILE_BWD_G
#define SPNB_TILE_BWD_G 4
#endif
constexpr int kMaxLayers = 6;
constexpr unsigned kSrcLocs = 15;  // "data is the position tensor"
constexpr unsigned kSrcOnes = 14;  // "data is all ones" (one channel, not stored)
constexpr int kFnRuntime = 15;     // kernel id not compiled in: evaluated through a switch per pair

// ---- compile-time group signature -------------------------------------------------------------------
// CS: 4 bits per layer = in-channels C_l;  SS: 4 bits per layer = index of the distinct data tensor feeding
// the layer, kSrcLocs or kSrcOnes;  FS: 4 bits per layer = kernel id (kFnRuntime: run-time id);  NS: 1 bit per
// layer = dis_norm;  DM: 1 bit per layer = the layer's data needs a gradient.
__host__ __device__ constexpr int deriv_expr_ct(int fn)
{
    return fn == E_DEFAULT ? E_DDEFAULT : fn == E_DDEFAULT ? E_DDEFAULT2 : fn == E_DDEFAULT2 ? E_D_DDEFAULT2
         : fn == E_PRESSURE ? E_DPRESSURE : fn == E_DPRESSURE ? E_DPRESSURE2 : fn == E_DPRESSURE2 ? E_D_DPRESSURE2
         : fn == E_INDIRECT ? E_D_INDIRECT : fn == E_CONSTANT ? E_D_CONSTANT : fn == E_SPIKY ? E_DSPIKY
         : fn == E_DSPIKY ? E_D_DSPIKY : fn == E_COHESION ? E_D_COHESION : fn == E_SIGMOID ? E_D_SIGMOID
         : E_D_CONSTANT;
}

template <int D_, int NL_, unsigned CS_, unsigned SS_, unsigned FS_, unsigned NS_, unsigned DM_>
struct Sig {
    static constexpr int D = D_, NL = NL_;
    static constexpr unsigned CS = CS_, SS = SS_, FS = FS_, NS = NS_, DM = DM_;
    static __host__ __device__ constexpr int FN(int l) { return (FS_ >> (4 * l)) & 15; }
    static __host__ __device__ constexpr bool RT(int l) { return FN(l) == kFnRuntime; }
    static __host__ __device__ constexpr int DFN(int l) { return deriv_expr_ct(FN(l)); }
    static __host__ __device__ constexpr int NORM(int l) { return (NS_ >> l) & 1; }
    static __host__ __device__ constexpr int SAL(int l)  // first layer with the same (kernel, dis_norm)
    {
        if (RT(l)) return l;
        for (int m = 0; m < l; ++m)
            if (FN(m) == FN(l) && NORM(m) == NORM(l)) return m;
        return l;
    }
    static __host__ __device__ constexpr int C(int l) { return (CS_ >> (4 * l)) & 15; }
    static __host__ __device__ constexpr int S(int l) { return (SS_ >> (4 * l)) & 15; }
    static __host__ __device__ constexpr bool ONES(int l) { return S(l) == (int)kSrcOnes; }
    static __host__ __device__ constexpr bool LOCS(int l) { return S(l) == (int)kSrcLocs; }
    static __host__ __device__ constexpr bool DD(int l) { return (DM_ >> l) & 1; }
    // dW/dd of the layer's kernel is not identically zero: the layer contributes to d/dlocs
    static __host__ __device__ constexpr bool HAS_T(int l) { return RT(l) || DFN(l) != E_D_CONSTANT; }
    static __host__ __device__ constexpr bool HAS_U(int l) { return HAS_T(l) || DD(l); }  // takes part in the backward
    static __host__ __device__ constexpr int nsrc()
    {
        int n = 0;
        for (int l = 0; l < NL_; ++l)
            if (S(l) < (int)kSrcOnes && S(l) + 1 > n) n = S(l) + 1;
        return n;
    }
    static __host__ __device__ constexpr int src_channels(int s)
    {
        for (int l = 0; l < NL_; ++l)
            if (S(l) == s) return C(l);
        return 0;
    }
    // ---- forward record: [ locs(D) | distinct data tensors ]
    static __host__ __device__ constexpr int src_off(int s)
    {
        int o = D_;
        for (int t = 0; t < s; ++t) o += src_channels(t);
        return o;
    }
    static __host__ __device__ constexpr int data_off(int l) { return LOCS(l) ? 0 : src_off(S(l)); }  // not ONES
    static __host__ __device__ constexpr int fwd_floats() { return src_off(nsrc()); }
    // ---- backward record: [ locs(D) | data tensors of layers with HAS_T | U_l of layers with HAS_U ]
    static __host__ __device__ constexpr bool src_in_bwd(int s)
    {
        for (int l = 0; l < NL_; ++l)
            if (S(l) == s && HAS_T(l)) return true;
        return false;
    }
    static __host__ __device__ constexpr int bsrc_off(int s)
    {
        int o = D_;
        for (int t = 0; t < s; ++t)
            if (src_in_bwd(t)) o += src_channels(t);
        return o;
    }
    static __host__ __device__ constexpr int bdata_off(int l) { return LOCS(l) ? 0 : bsrc_off(S(l)); }  // HAS_T, not ONES
    static __host__ __device__ constexpr int u_off(int l)
    {
        int o = bsrc_off(nsrc());
        for (int t = 0; t < l; ++t)
            if (HAS_U(t)) o += C(t);
        return o;
    }
    static __host__ __device__ constexpr int bwd_floats() { return u_off(NL_); }
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
    static __host__ __device__ constexpr int fwd_vec() { return (fwd_floats() + 3) / 4; }
    static __host__ __device__ constexpr int bwd_vec() { return (bwd_floats() + 3) / 4; }
};

struct LayerArgs {
    const float* data;      // [B,N,C] (NULL: ones)
    const float* weight;    // [O,C]
    const float* bias;      // [O] or NULL
    float* out;             // fwd: [B,N,O]
    const float* grad_out;  // bwd: [B,N,O]
    float* ddata;           // bwd: [B,N,C] or NULL
    int C, O;
    int w_expr, dw_expr, dis_norm;
    float wc, dwc;
};
struct GroupArgs {
    LayerArgs l[kMaxLayers];
    const float* src[kMaxLayers];  // distinct data tensors
    float H, invH, H2, rad2;
};

// SphF / sph_fast: spnb_common.cuh

// ---- pack pre-pass ----------------------------------------------------------------------------------
// Planar (one float4 array per record quarter) for the tile kernels, record-major for the list walk; which one
// is decided by the same device flags the main kernel tests.
__device__ __forceinline__ bool tiles_usable(const int* tile_flag, const int* sym_flag, bool bwd)
{
    return tile_flag != nullptr && *tile_flag == 0 && !(bwd && sym_flag != nullptr && *sym_flag != 0);
}

template <typename SG, bool BWD>
__global__ void __launch_bounds__(256)
k_group_pack(const float* __restrict__ locs, GroupArgs ga, long long BN, float* __restrict__ rec,
             const int* __restrict__ tile_flag, float* __restrict__ dlocs, const int* __restrict__ sym_flag)
{
    const bool planar = tiles_usable(tile_flag, sym_flag, BWD);
    constexpr int V = BWD ? SG::bwd_vec() : SG::fwd_vec();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    float r[V * 4];
#pragma unroll
    for (int i = 0; i < V * 4; ++i) r[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < SG::D; ++k) r[k] = locs[n * SG::D + k];
#pragma unroll
    for (int s = 0; s < SG::nsrc(); ++s) {
        if (!BWD || SG::src_in_bwd(s)) {
#pragma unroll
            for (int c = 0; c < SG::src_channels(s); ++c)
                r[(BWD ? SG::bsrc_off(s) : SG::src_off(s)) + c] = ga.src[s][n * SG::src_channels(s) + c];
        }
    }
    if (BWD) {
        if (!(sym_flag != nullptr && *sym_flag == 0)) {
            // the scatter mode of the list walk accumulates with atomics: its targets start from zero
#pragma unroll
            for (int k = 0; k < SG::D; ++k) dlocs[n * SG::D + k] = 0.0f;
#pragma unroll
            for (int l = 0; l < SG::NL; ++l)
                if (SG::DD(l)) {
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c) ga.l[l].ddata[n * SG::C(l) + c] = 0.0f;
                }
        }
#pragma unroll
        for (int l = 0; l < SG::NL; ++l) {
            if (!SG::HAS_U(l)) continue;
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

// Per-layer kernel coefficients copied out of the parameter block once per thread (re-reading the constant bank
// per pair costs ~3 issue slots per pair).
template <int NL>
struct LayerCoef {
    struct { float wc, dwc; int we, dwe, norm; } l[NL];
};
template <typename SG>
__device__ __forceinline__ LayerCoef<SG::NL> load_coef(const GroupArgs& ga)
{
    LayerCoef<SG::NL> c;
#pragma unroll
    for (int l = 0; l < SG::NL; ++l) {
        c.l[l].wc = ga.l[l].wc;
        c.l[l].dwc = ga.l[l].dwc;
        c.l[l].we = ga.l[l].w_expr;
        c.l[l].dwe = ga.l[l].dw_expr;
        c.l[l].norm = ga.l[l].dis_norm;
    }
    return c;
}

// s_l = W_l(d) * norm_l (and t_l = dW_l/dd / d * norm_l) for every layer, evaluated once per distinct
// (kernel, dis_norm); everything about the layer list is a compile-time constant (except run-time kernel ids).
template <typename SG, bool WITH_T>
__device__ __forceinline__ void layer_scales(const LayerCoef<SG::NL>& co, const SphF& sp, float d, float d2,
                                             float inv, bool pos, float* s, float* t)
{
#pragma unroll
    for (int l = 0; l < SG::NL; ++l) {
        if (SG::SAL(l) == l) {
            const float norm = ((SG::RT(l) ? co.l[l].norm != 0 : SG::NORM(l) != 0) && pos) ? inv : 1.0f;
            s[l] = sph_fast(SG::RT(l) ? co.l[l].we : SG::FN(l), d, d2, co.l[l].wc, sp) * norm;
            if (WITH_T)
                t[l] = (SG::HAS_T(l) && pos)
                           ? sph_fast(SG::RT(l) ? co.l[l].dwe : SG::DFN(l), d, d2, co.l[l].dwc, sp) * inv * norm
                           : 0.0f;
        } else {
            s[l] = s[SG::SAL(l)];
            if (WITH_T) t[l] = t[SG::SAL(l)];
        }
    }
}

// data_l[c] of a record (ones are not stored)
template <typename SG, bool BWD>
__device__ __forceinline__ float rec_data(const float* r, int l, int c)
{
    return SG::ONES(l) ? 1.0f : r[(BWD ? SG::bdata_off(l) : SG::data_off(l)) + c];
}

// ---- pair math -----------------------------------------------------------------------------------------
template <typename SG>
struct FwdPair {
    float x[SG::D];
    float G_[SG::ctot()];
    LayerCoef<SG::NL> co;
    SphF sp;
    float rad2;
    __device__ __forceinline__ void init(const GroupArgs& ga)
    {
        co = load_coef<SG>(ga);
        sp.H = ga.H; sp.invH = ga.invH; sp.H2 = ga.H2;
        rad2 = ga.rad2;
#pragma unroll
        for (int i = 0; i < SG::ctot(); ++i) G_[i] = 0.0f;
    }
    // r: record of the neighbour; `ok`: the entry exists
    __device__ __forceinline__ void operator()(const float* r, bool ok)
    {
        float d2 = 0.0f;
#pragma unroll
        for (int k = 0; k < SG::D; ++k) {
            const float nr = x[k] - r[k];
            d2 += nr * nr;
        }
        if (ok && d2 < rad2) {
            const bool pos = d2 > 0.0f;
            const float inv = fast_rsqrt(d2);
            const float d = pos ? d2 * inv : 0.0f;
            float s[SG::NL];
            layer_scales<SG, false>(co, sp, d, d2, inv, pos, s, nullptr);
#pragma unroll
            for (int l = 0; l < SG::NL; ++l)
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c)
                    G_[SG::chan_off(l) + c] = fmaf(s[l], rec_data<SG, false>(r, l, c), G_[SG::chan_off(l) + c]);
        }
    }
    // epilogue: apply the weights once per query (the G lanes of a group split the outputs)
    template <int G>
    __device__ __forceinline__ void finish(const GroupArgs& ga, size_t q, bool active, int sub)
    {
        if (G > 1) {
#pragma unroll
            for (int i = 0; i < SG::ctot(); ++i) G_[i] = group_sum<G>(G_[i]);
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
};

// Backward in the symmetric (gather) mode: everything particle i receives -- through its query role and its
// neighbour role -- is summed over i's own list (same distances, negated displacement).
template <typename SG>
struct BwdPairSym {
    float me[SG::bwd_vec() * 4];  // my own record: position, data_l[i], U_l[i]
    float a_dl[SG::D], a_dd[SG::ctot()];
    LayerCoef<SG::NL> co;
    SphF sp;
    float rad2;
    __device__ __forceinline__ void init(const GroupArgs& ga)
    {
        co = load_coef<SG>(ga);
        sp.H = ga.H; sp.invH = ga.invH; sp.H2 = ga.H2;
        rad2 = ga.rad2;
#pragma unroll
        for (int k = 0; k < SG::D; ++k) a_dl[k] = 0.0f;
#pragma unroll
        for (int i = 0; i < SG::ctot(); ++i) a_dd[i] = 0.0f;
    }
    __device__ __forceinline__ void operator()(const float* r, bool ok)
    {
        float disp[SG::D];
        float d2 = 0.0f;
#pragma unroll
        for (int k = 0; k < SG::D; ++k) {
            disp[k] = me[k] - r[k];
            d2 += disp[k] * disp[k];
        }
        if (ok && d2 < rad2) {
            const bool pos = d2 > 0.0f;
            const float inv = fast_rsqrt(d2);
            const float d = pos ? d2 * inv : 0.0f;
            float s[SG::NL], t[SG::NL];
            layer_scales<SG, true>(co, sp, d, d2, inv, pos, s, t);
            float T = 0.0f;
#pragma unroll
            for (int l = 0; l < SG::NL; ++l) {
                if (!SG::HAS_U(l)) continue;
                float AB = 0.0f;
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c) {
                    if (SG::HAS_T(l)) {
                        // pair (i, j): U_l[i] . data_l[j]   +   pair (j, i): U_l[j] . data_l[i]
                        AB = fmaf(me[SG::u_off(l) + c], rec_data<SG, true>(r, l, c), AB);
                        AB = fmaf(r[SG::u_off(l) + c], rec_data<SG, true>(me, l, c), AB);
                    }
                    if (SG::DD(l))
                        a_dd[SG::chan_off(l) + c] = fmaf(s[l], r[SG::u_off(l) + c], a_dd[SG::chan_off(l) + c]);
                }
                if (SG::HAS_T(l)) T = fmaf(AB, t[l], T);
            }
#pragma unroll
            for (int k = 0; k < SG::D; ++k) a_dl[k] = fmaf(T, disp[k], a_dl[k]);
        }
    }
    template <int G>
    __device__ __forceinline__ void finish(const GroupArgs& ga, float* dlocs, size_t q, bool active, int sub)
    {
        if (G > 1) {
#pragma unroll
            for (int k = 0; k < SG::D; ++k) a_dl[k] = group_sum<G>(a_dl[k]);
#pragma unroll
            for (int l = 0; l < SG::NL; ++l)
                if (SG::DD(l)) {
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c)
                        a_dd[SG::chan_off(l) + c] = group_sum<G>(a_dd[SG::chan_off(l) + c]);
                }
        }
        if (active && sub == 0) {
#pragma unroll
            for (int k = 0; k < SG::D; ++k) dlocs[q * SG::D + k] = a_dl[k];
#pragma unroll
            for (int l = 0; l < SG::NL; ++l)
                if (SG::DD(l)) {
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c) ga.l[l].ddata[q * SG::C(l) + c] = a_dd[SG::chan_off(l) + c];
                }
        }
    }
};

// ---- float-list walk (records record-major) -----------------------------------------------------------
template <typename SG, int THREADS>
__device__ __forceinline__ void group_fwd_block(const float* __restrict__ rec, const float* __restrict__ neighbors,
                                                const GroupArgs& ga, int N, int K, int bx, int b, int nbx, int nby,
                                                WalkSmem<SPNB_GROUP_FWD_G>* s_walk)  // one per warp of the block
{
    constexpr int D = SG::D, V = SG::fwd_vec();
    constexpr int G = SPNB_GROUP_FWD_G, kU = SPNB_GROUP_FWD_U, QPB = THREADS / G, R = 32 / G;
    const int warp = threadIdx.x >> 5, sub = threadIdx.x % G;
    const int m = bx * QPB + threadIdx.x / G;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    const int m0 = bx * QPB + warp * R;  // first query of this warp
    const int nrows = min(R, max(0, N - m0));
    const float4* srec = reinterpret_cast<const float4*>(rec) + (size_t)b * N * V;
    FwdPair<SG> P;
    P.init(ga);
    {
        const float4 r0 = srec[(size_t)(active ? m : 0) * V];
        const float t[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
        for (int k = 0; k < D; ++k) P.x[k] = t[k];
    }
    const float* warp_rows = neighbors + ((size_t)b * N + min(m0, N - 1)) * K;
    prefetch_rows_ahead(neighbors, N, K, QPB, 2, bx, b, nbx, nby);
    walk_rows<G, kU>(warp_rows, K, nrows, s_walk[warp], [&](const int* j, const bool* valid) {
        float r[kU][V * 4];
#pragma unroll
        for (int u = 0; u < kU; ++u)
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 t = srec[(unsigned)j[u] * (unsigned)V + v];
                r[u][4 * v] = t.x; r[u][4 * v + 1] = t.y; r[u][4 * v + 2] = t.z; r[u][4 * v + 3] = t.w;
            }
#pragma unroll
        for (int u = 0; u < kU; ++u) P(r[u], valid[u]);
    });
    P.template finish<G>(ga, q, active, sub);
}

// dlocs [B,N,D]: d(sum_l loss_l)/d(locs) through the geometry (query role + neighbour role).
// ddata_l [B,N,C_l] for the layers of the signature's DM mask.  sym: gather; else scatter with atomics into
// buffers the pack pre-pass zero-filled.
template <typename SG, int THREADS>
__device__ __forceinline__ void group_bwd_block(const float* __restrict__ rec, const float* __restrict__ neighbors,
                                                const GroupArgs& ga, int N, int K, float* dlocs, const int* sym_flag,
                                                int bx, int b, int nbx, int nby,
                                                WalkSmem<SPNB_GROUP_BWD_G>* s_walk)  // one per warp of the block
{
    constexpr int D = SG::D, V = SG::bwd_vec();
    constexpr int G = SPNB_GROUP_BWD_G, UB = SPNB_GROUP_BWD_U, QPB = THREADS / G, R = 32 / G;
    const bool sym = sym_flag != nullptr && *sym_flag == 0;
    const int warp = threadIdx.x >> 5, sub = threadIdx.x % G;
    const int m = bx * QPB + threadIdx.x / G;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    const int m0 = bx * QPB + warp * R;
    const int nrows = min(R, max(0, N - m0));
    const float4* srec = reinterpret_cast<const float4*>(rec) + (size_t)b * N * V;
    BwdPairSym<SG> P;
    P.init(ga);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float4 t = srec[(size_t)(active ? m : 0) * V + v];
        P.me[4 * v] = t.x; P.me[4 * v + 1] = t.y; P.me[4 * v + 2] = t.z; P.me[4 * v + 3] = t.w;
    }
    const float* warp_rows = neighbors + ((size_t)b * N + min(m0, N - 1)) * K;
    prefetch_rows_ahead(neighbors, N, K, QPB, 2, bx, b, nbx, nby);
    if (sym) {
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
            for (int u = 0; u < UB; ++u) P(r[u], valid[u]);
        });
        P.template finish<G>(ga, dlocs, q, active, sub);
        return;
    }
    // ---- scatter mode: pair (i, j) only; what j receives goes through float atomics
    float a_dl[D];
#pragma unroll
    for (int k = 0; k < D; ++k) a_dl[k] = 0.0f;
    walk_rows<G, 1>(warp_rows, K, nrows, s_walk[warp], [&](const int* j, const bool* valid) {
        float r[V * 4];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 t = srec[(unsigned)j[0] * (unsigned)V + v];
            r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
        }
        float disp[D];
        float d2 = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            disp[k] = P.me[k] - r[k];
            d2 += disp[k] * disp[k];
        }
        if (valid[0] && d2 < P.rad2) {
            const bool pos = d2 > 0.0f;
            const float inv = fast_rsqrt(d2);
            const float d = pos ? d2 * inv : 0.0f;
            float s[SG::NL], t[SG::NL];
            layer_scales<SG, true>(P.co, P.sp, d, d2, inv, pos, s, t);
            float TA = 0.0f;
#pragma unroll
            for (int l = 0; l < SG::NL; ++l) {
                if (!SG::HAS_T(l)) continue;
                float A = 0.0f;
#pragma unroll
                for (int c = 0; c < SG::C(l); ++c) A = fmaf(P.me[SG::u_off(l) + c], rec_data<SG, true>(r, l, c), A);
                TA = fmaf(A, t[l], TA);
            }
            const size_t jo = (size_t)b * N + j[0];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                a_dl[k] = fmaf(TA, disp[k], a_dl[k]);
                if (pos) atomicAdd(dlocs + jo * D + k, -TA * disp[k]);
            }
#pragma unroll
            for (int l = 0; l < SG::NL; ++l) {
                if (SG::DD(l)) {
#pragma unroll
                    for (int c = 0; c < SG::C(l); ++c)
                        atomicAdd(ga.l[l].ddata + jo * SG::C(l) + c, s[l] * P.me[SG::u_off(l) + c]);
                }
            }
        }
    });
    if (G > 1) {
#pragma unroll
        for (int k = 0; k < D; ++k) a_dl[k] = group_sum<G>(a_dl[k]);
    }
    if (active && sub == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) atomicAdd(dlocs + q * D + k, a_dl[k]);
    }
}

// ---- tile-list kernels --------------------------------------------------------------------------------
struct TileArgs {
    const int* flag;             // NULL: no tile lists for this call
    const TileDesc* descs;
    const unsigned char* blobs;  // per block: perm[64] + entry rows (tile_lists.cuh)
    size_t blob_stride;
    int ntb;
};

// Stages chunk `chunk` of the block's records: the slots 1 + chunk*(kTileCap-1) ... of the concatenated ranges become
// the local slots 1 .. kTileCap-1.  Almost every block has one chunk; a block whose ranges hold more records than
// the staged tile is processed in several passes over its lists (entries outside the staged chunk are skipped).
template <int V>
__device__ __forceinline__ void tile_stage(const float4* __restrict__ planes, size_t plane_stride, size_t scene_off,
                                           const TileDesc& d, int chunk, float4* s_rec, unsigned long long* bar)
{
    // warp 0: lane 0 arms the barrier with the byte count, then the lanes issue the copies
    const int lane = threadIdx.x;
    const int c0 = chunk * (kTileCap - 1), c1 = min(d.total, c0 + kTileCap - 1);
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(c1 - c0) * 16u * V);
    __syncwarp();
    const int ncopies = d.nr * V;
    for (int i = lane; i < ncopies; i += 32) {
        const int r = i / V, v = i % V;
        const int a = max(d.prefix[r], c0), z = min(d.prefix[r + 1], c1);
        if (z > a)
            bulk_copy_g2s(s_rec + (size_t)v * kTileCap + 1 + (a - c0),
                          planes + (size_t)v * plane_stride + scene_off + d.start[r] + (a - d.prefix[r]),
                          (unsigned)(z - a) * 16u, bar);
    }
}

// Walks the entry rows of this warp's octiles (tile_lists.cuh).  A warp of a kernel with G lanes per query owns
// NO = 4/G octiles (32/G queries in rank order); lane -> (octile o, query r of the octile, share `sub` of its
// entries).  The rows stream from global memory through a double-buffered shared-memory stage of 8 rows (512
// bytes, one 16-byte cp.async per lane) per octile, so the loop itself only touches shared memory; per step a
// lane reads its 8/G bytes of the row and calls body(slot * 16) for each of its 4/G entries (padding entries
// are slot 0, the sentinel record: they fail every radius test).
constexpr int kTileStageBytes = 8 * 1024;  // per CTA, whatever G: warps * NO * 2 buffers * 512 bytes

template <int G>
struct TileWalk {
    static constexpr int NO = 4 / G, LPO = 32 / NO;
    int S[NO], R0[NO];
    int smax_all, my_s, nch;
    unsigned char* stage;
    const unsigned char* rows;

    __device__ __forceinline__ void issue(int c) const
    {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int oo = 0; oo < NO; ++oo) {
            const int nrows = min(8, S[oo] - 8 * c);
            if (lane * 16 < nrows * kTileRowBytes)
                cp_async16(stage + (oo * 2 + (c & 1)) * 512 + lane * 16,
                           rows + (size_t)(R0[oo] + 8 * c) * kTileRowBytes + lane * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // reads the octile row counts and requests the first 8 rows of every octile of this warp
    __device__ __forceinline__ void begin(const unsigned char* __restrict__ rows_, const TileDesc& d,
                                          unsigned char* __restrict__ stage_cta)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int o = lane / LPO;
        rows = rows_;
        stage = stage_cta + warp * (NO * 1024);
        smax_all = 0;
        my_s = 0;
#pragma unroll
        for (int oo = 0; oo < NO; ++oo) {
            R0[oo] = d.goff[warp * NO + oo];
            S[oo] = d.goff[warp * NO + oo + 1] - R0[oo];
            smax_all = max(smax_all, S[oo]);
            if (oo == o) my_s = S[oo];
        }
        nch = (smax_all + 7) >> 3;
        if (nch > 0) issue(0);
    }
    // `first`: chunk 0 is already in flight (begin() requested it)
    template <typename Body>
    __device__ __forceinline__ void run(bool first, Body body) const
    {
        const int lane = threadIdx.x & 31;
        const int o = lane / LPO, li = lane % LPO;
        if (!first && nch > 0) {
            __syncwarp();
            issue(0);
        }
        for (int c = 0; c < nch; ++c) {
            __syncwarp();  // every lane is done with the buffer the next copy overwrites
            if (c + 1 < nch) {
                issue(c + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            const unsigned char* buf = stage + (o * 2 + (c & 1)) * 512 + li * (8 / G);
            const int ns = min(8, my_s - 8 * c);  // steps of MY octile in this chunk (<= 0: none)
            if (G == 4) {
                // one entry per lane and step: two steps per iteration, so that the second gather is issued
                // before the first pair's arithmetic
#pragma unroll 1
                for (int s_ = 0; s_ < ns; s_ += 2) {
                    const unsigned e0 = *reinterpret_cast<const unsigned short*>(buf + s_ * kTileRowBytes);
                    const unsigned e1 = s_ + 1 < ns ? *reinterpret_cast<const unsigned short*>(buf + (s_ + 1) * kTileRowBytes) : 0u;
                    body(e0, e1);
                }
            } else if (G == 2) {
#pragma unroll 1
                for (int s_ = 0; s_ < ns; ++s_) {
                    const unsigned w = *reinterpret_cast<const unsigned*>(buf + s_ * kTileRowBytes);
                    body(w & 0xffffu, w >> 16);
                }
            } else {
#pragma unroll 1
                for (int s_ = 0; s_ < ns; ++s_) {
                    const uint2 w = *reinterpret_cast<const uint2*>(buf + s_ * kTileRowBytes);
                    body(w.x & 0xffffu, w.x >> 16);
                    body(w.y & 0xffffu, w.y >> 16);
                }
            }
        }
    }
};

// Query of this thread: rank order within the block (tile_lists.cuh).  Returns the query's index within the block.
template <int G>
__device__ __forceinline__ int tile_my_query(const unsigned char* __restrict__ blob)
{
    constexpr int NO = 4 / G, LPO = 32 / NO;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = 8 * (warp * NO + lane / LPO) + (lane % LPO) / G;
    return (int)__ldg(blob + rank);
}

// The common frame of the two tile kernels: descriptor, records staged chunk by chunk, the walk.  `pair2(ra, rb)`
// consumes two gathered records.
template <int V, int G, typename Pair>
__device__ __forceinline__ void tile_run(const float4* __restrict__ planes, long long BN, size_t scene_off,
                                         const TileDesc& s_desc, const TileWalk<G>& W, unsigned char* s_raw,
                                         unsigned long long* s_bar, Pair& P)
{
    const int tid = threadIdx.x;
    float4* s_rec = reinterpret_cast<float4*>(s_raw);
    const int nchunks = max(1, (s_desc.total + kTileCap - 2) / (kTileCap - 1));
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        if (chunk > 0) {
            __syncthreads();  // every warp is done with the previous chunk's records
            if (tid < 32) tile_stage<V>(planes, (size_t)BN, scene_off, s_desc, chunk, s_rec, s_bar);
        }
        // one warp waits on the mbarrier (a polling loop), the others at the block barrier (no issue slots)
        if (tid < 32) mbar_wait(s_bar, chunk & 1);
        __syncthreads();
        // entry (slot * 16) -> byte offset of the record in the staged chunk, 0 (the sentinel) when outside it
        const unsigned cb = (unsigned)chunk * (kTileCap - 1) * 16u + 16u;
        auto local = [&](unsigned e) {
            const unsigned t = e - cb;
            return t < (kTileCap - 1) * 16u ? t + 16u : 0u;
        };
        W.run(chunk == 0, [&](unsigned ea, unsigned eb) {
            const unsigned sa = nchunks > 1 ? local(ea) : ea, sb = nchunks > 1 ? local(eb) : eb;
            float ra[V * 4], rb[V * 4];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 f = *reinterpret_cast<const float4*>(s_raw + v * (kTileCap * 16) + sa);
                ra[4 * v] = f.x; ra[4 * v + 1] = f.y; ra[4 * v + 2] = f.z; ra[4 * v + 3] = f.w;
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 f = *reinterpret_cast<const float4*>(s_raw + v * (kTileCap * 16) + sb);
                rb[4 * v] = f.x; rb[4 * v + 1] = f.y; rb[4 * v + 2] = f.z; rb[4 * v + 3] = f.w;
            }
            P(ra, true);
            P(rb, true);
        });
    }
}

template <typename SG, int G>
__global__ void __launch_bounds__(kTileQ * G)
k_tile_fwd(const float* __restrict__ rec, TileArgs ta, GroupArgs ga, int N, int K, long long BN,
           const float* __restrict__ neighbors)
{
    constexpr int D = SG::D, V = SG::fwd_vec();
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_rec = reinterpret_cast<float4*>(s_raw);
    __shared__ TileDesc s_desc;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, tb = blockIdx.x, b = blockIdx.y;
    // the descriptor load is issued together with the flag load (both are inputs of this call's
    // predecessors only), so the flag test does not add a global-memory latency to the prologue
    int desc_word = 0, ql = 0;
    const unsigned char* blob = ta.blobs + ((size_t)b * ta.ntb + tb) * ta.blob_stride;
    if (ta.flag != nullptr) {
        if (tid < 32) desc_word = __ldg(reinterpret_cast<const int*>(ta.descs + (size_t)b * ta.ntb + tb) + tid);
        ql = tile_my_query<G>(blob);  // my query (rank order); issued with the descriptor load
    }
    if (!tiles_usable(ta.flag, nullptr, false)) {
        // no usable tile lists for this call (a list reached K, ...): the float-list walk, strided over the
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
    if (tid < 32) tile_stage<V>(planes, (size_t)BN, (size_t)b * N, s_desc, 0, s_rec, &s_bar);
    TileWalk<G> W;
    W.begin(blob + kTileHeaderBytes, s_desc, s_raw + V * (kTileCap * 16));  // first list rows in flight

    const int sub = tid % G;
    const int m = tb * kTileQ + ql;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    FwdPair<SG> P;
    P.init(ga);
    {
        const float4 r0 = planes[q];
        const float t[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
        for (int k = 0; k < D; ++k) P.x[k] = t[k];
    }
    tile_run<V, G>(planes, BN, (size_t)b * N, s_desc, W, s_raw, &s_bar, P);
    P.template finish<G>(ga, q, active, sub);
}

template <typename SG, int G>
__global__ void __launch_bounds__(kTileQ * G, (kTileQ * G >= 256 ? (SG::bwd_vec() <= 3 ? 4 : 3) : 1))
k_tile_bwd(const float* __restrict__ rec, TileArgs ta, GroupArgs ga, int N, int K, long long BN, float* dlocs,
           const float* __restrict__ neighbors, const int* sym_flag)
{
    constexpr int V = SG::bwd_vec();
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_rec = reinterpret_cast<float4*>(s_raw);
    __shared__ TileDesc s_desc;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, tb = blockIdx.x, b = blockIdx.y;
    int desc_word = 0, ql = 0;
    const unsigned char* blob = ta.blobs + ((size_t)b * ta.ntb + tb) * ta.blob_stride;
    if (ta.flag != nullptr) {
        if (tid < 32) desc_word = __ldg(reinterpret_cast<const int*>(ta.descs + (size_t)b * ta.ntb + tb) + tid);
        ql = tile_my_query<G>(blob);  // my query (rank order); issued with the descriptor load
    }
    if (!tiles_usable(ta.flag, sym_flag, true)) {
        // no usable tile lists for this call, or the relation is not symmetric (the tile path only has the
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
    if (tid < 32) tile_stage<V>(planes, (size_t)BN, (size_t)b * N, s_desc, 0, s_rec, &s_bar);
    TileWalk<G> W;
    W.begin(blob + kTileHeaderBytes, s_desc, s_raw + V * (kTileCap * 16));  // first list rows in flight

    const int sub = tid % G;
    const int m = tb * kTileQ + ql;
    const bool active = m < N;
    const size_t q = (size_t)b * N + (active ? m : 0);
    BwdPairSym<SG> P;
    P.init(ga);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float4 t = planes[(size_t)v * BN + q];
        P.me[4 * v] = t.x; P.me[4 * v + 1] = t.y; P.me[4 * v + 2] = t.z; P.me[4 * v + 3] = t.w;
    }
    tile_run<V, G>(planes, BN, (size_t)b * N, s_desc, W, s_raw, &s_bar, P);
    P.template finish<G>(ga, dlocs, q, active, sub);
}

// ---- host: launch of one signature --------------------------------------------------------------------
struct Signature {
    int D, NL;
    unsigned CS, SS, FS, NS, DM;
};

struct RunArgs {
    const float* locs;
    const float* neighbors;
    GroupArgs ga;
    int B, N, K;
    float* rec;
    float* dlocs;
    const int* sym_flag;
    const void* tiles;
    cudaStream_t stream;
};

inline bool make_tile_args(const void* tiles, int B, int N, int K, TileArgs& ta)
{
    ta.flag = nullptr;
    ta.descs = nullptr;
    ta.blobs = nullptr;
    ta.blob_stride = 0;
    const TileLayout tl = tile_layout(B, N, K);
    ta.ntb = tl.ntb;
    if (!tiles) return false;
    const unsigned char* base = (const unsigned char*)tiles;
    ta.flag = (const int*)base;
    ta.descs = (const TileDesc*)(base + tl.desc_off);
    ta.blobs = base + tl.list_off;
    ta.blob_stride = tl.blob_stride;
    return true;
}

inline bool launched(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return true;
    set_error("convsp group: launch of %s failed: %s", what, cudaGetErrorString(e));
    return false;
}

template <typename KernelT>
inline bool allow_smem(KernelT* kernel, size_t bytes)
{
    // as many CTAs per SM as the shared memory allows: ask for the largest carve-out (the driver's default choice left
    // the 3-plane backward kernels at 3 CTAs per SM where 4 fit)
    cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
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

// pack (layout chosen on the device by the flags) + the tile kernel, which walks the float lists itself when
// the tile lists are absent or unusable.  Returns the number of launches, -1 on failure.
template <typename SG>
int run_fwd(const RunArgs& a)
{
    const long long BN = (long long)a.B * a.N;
    TileArgs ta;
    make_tile_args(a.tiles, a.B, a.N, a.K, ta);
    k_group_pack<SG, false><<<cdiv(BN, 256), 256, 0, a.stream>>>(a.locs, a.ga, BN, a.rec, ta.flag, nullptr, nullptr);
    if (!launched("k_group_pack")) return -1;
    constexpr int G = SPNB_TILE_FWD_G > 0 ? SPNB_TILE_FWD_G : (SG::fwd_vec() == 1 ? 2 : 4);  // measured (profiles/README.md)
    static_assert(sizeof(WalkSmem<SPNB_GROUP_FWD_G>) * (kTileQ * G / 32) <= kTileCap * sizeof(float4), "walk scratch fits the tile buffer");
    const size_t smem = (size_t)SG::fwd_vec() * kTileCap * sizeof(float4) + kTileStageBytes;
    if (!allow_smem(k_tile_fwd<SG, G>, smem)) return -1;
    k_tile_fwd<SG, G><<<dim3(ta.ntb, a.B), kTileQ * G, smem, a.stream>>>(a.rec, ta, a.ga, a.N, a.K, BN, a.neighbors);
    if (!launched("k_tile_fwd")) return -1;
    return 2;
}

template <typename SG>
int run_bwd(const RunArgs& a)
{
    const long long BN = (long long)a.B * a.N;
    TileArgs ta;
    make_tile_args(a.tiles, a.B, a.N, a.K, ta);
    k_group_pack<SG, true><<<cdiv(BN, 256), 256, 0, a.stream>>>(a.locs, a.ga, BN, a.rec, ta.flag, a.dlocs, a.sym_flag);
    if (!launched("k_group_pack")) return -1;
    constexpr int G = SPNB_TILE_BWD_G;
    static_assert(sizeof(WalkSmem<SPNB_GROUP_BWD_G>) * (kTileQ * G / 32) <= kTileCap * sizeof(float4), "walk scratch fits the tile buffer");
    const size_t smem = (size_t)SG::bwd_vec() * kTileCap * sizeof(float4) + kTileStageBytes;
    if (!allow_smem(k_tile_bwd<SG, G>, smem)) return -1;
    k_tile_bwd<SG, G><<<dim3(ta.ntb, a.B), kTileQ * G, smem, a.stream>>>(a.rec, ta, a.ga, a.N, a.K, BN, a.dlocs, a.neighbors, a.sym_flag);
    if (!launched("k_tile_bwd")) return -1;
    return 2;
}

// One table row per instantiated signature (filled by the convsp_group_inst*.cu translation units).
struct SigEntry {
    Signature sg;
    int fwd_vec, bwd_vec;
    int (*fwd)(const RunArgs&);
    int (*bwd)(const RunArgs&);
};

template <typename SG>
constexpr SigEntry sig_entry()
{
    return SigEntry{{SG::D, SG::NL, SG::CS, SG::SS, SG::FS, SG::NS, SG::DM}, SG::fwd_vec(), SG::bwd_vec(),
                    &run_fwd<SG>, &run_bwd<SG>};
}

// tables defined in the instantiation units
extern const SigEntry kSigsFluid3[];
extern const int kNumSigsFluid3;
extern const SigEntry kSigsFluid2[];
extern const int kNumSigsFluid2;
extern const SigEntry kSigsSingle3[];
extern const int kNumSigsSingle3;
extern const SigEntry kSigsSingle2[];
extern const int kNumSigsSingle2;

}  // namespace grp
}  // namespace spnb
