// Fused per-particle stages of the Position-Based-Fluids solver iteration.
//
// The reference's fluid step (examples/fluid_sim.py:367-397) strings ~35 tiny elementwise torch ops
// between the ConvSP layers of one solver iteration, and autograd doubles that in the backward pass;
// at 8 x 65 536 particles every one of them is a 2-3 us launch that moves a few MB.  The three stages
// below evaluate the same arithmetic (same operation order per element) in one pass each, with
// hand-derived backward kernels, so that an iteration is 3 ConvSP group calls + 3 stage launches.
//
//   stage 1 (after group A):  p    = k * relu(density - rho0)                  fluid_sim.py:373
//                             xp   = x * p                                      :374 (data of dspikyD)
//                             nij  = x * ni_s - nj                              :370-372
//   stage 2 (after group B):  nijp = x * nip_s - njp                            :374-376
//                             nij2 = x * ni_cs - nj_c                           :380-382
//                             d0   = -(p * nij + nijp) + (-cohesion * nij2 * radius)      :377,383
//                             nrm  = nij2 * surface_tension / rho0 / constraint_scale     :386
//   stage 3 (after group C):  delta = d0 + (cd - nrm * ncount)                  :390
//                             scale = relu(ncount / (1 + relaxation) - damp) + damp       :392-393
//                             xnew  = x + delta / scale                         :394-395
//
// All tensors are contiguous float32: vectors [BN, D], scalars [BN, 1].  Thread per particle; the
// kernels are plain streaming passes (HBM-bound, 60-130 bytes per particle).
#include "spnb_common.cuh"

namespace spnb {
namespace {

constexpr int kGlueThreads = 256;

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf1_fwd(const float* __restrict__ x, const float* __restrict__ density, const float* __restrict__ nj,
           const float* __restrict__ ni_s, float* __restrict__ p, float* __restrict__ xp,
           float* __restrict__ nij, long long BN, float k, float rho0)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float t = density[n] - rho0;
    const float pv = k * (t > 0.0f ? t : 0.0f);
    const float s = ni_s[n];
    p[n] = pv;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float xv = x[n * D + c];
        xp[n * D + c] = xv * pv;
        nij[n * D + c] = xv * s - nj[n * D + c];
    }
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf1_bwd(const float* __restrict__ x, const float* __restrict__ density, const float* __restrict__ ni_s,
           const float* __restrict__ g_p, const float* __restrict__ g_xp, const float* __restrict__ g_nij,
           float* __restrict__ g_x, float* __restrict__ g_density, float* __restrict__ g_nj,
           float* __restrict__ g_ni_s, long long BN, float k, float rho0)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float t = density[n] - rho0;
    const float pv = k * (t > 0.0f ? t : 0.0f);
    const float s = ni_s[n];
    float gp = g_p ? g_p[n] : 0.0f, gs = 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float xv = x[n * D + c];
        const float a = g_xp ? g_xp[n * D + c] : 0.0f;
        const float bq = g_nij ? g_nij[n * D + c] : 0.0f;
        g_x[n * D + c] = a * pv + bq * s;
        gp += a * xv;
        gs += bq * xv;
        g_nj[n * D + c] = -bq;
    }
    g_density[n] = t > 0.0f ? gp * k : 0.0f;
    g_ni_s[n] = gs;
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf2_fwd(const float* __restrict__ x, const float* __restrict__ p, const float* __restrict__ nij,
           const float* __restrict__ njp, const float* __restrict__ nip_s, const float* __restrict__ nj_c,
           const float* __restrict__ ni_cs, float* __restrict__ d0, float* __restrict__ nrm, long long BN,
           float coh, float radius, float st, float rho0, float cscale)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float pv = p[n], a = nip_s[n], bq = ni_cs[n];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float xv = x[n * D + c];
        const float nijp = xv * a - njp[n * D + c];
        const float nij2 = xv * bq - nj_c[n * D + c];
        d0[n * D + c] = -(pv * nij[n * D + c] + nijp) + (-coh * nij2 * radius);
        nrm[n * D + c] = nij2 * st / rho0 / cscale;
    }
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf2_bwd(const float* __restrict__ x, const float* __restrict__ p, const float* __restrict__ nij,
           const float* __restrict__ nip_s, const float* __restrict__ ni_cs, const float* __restrict__ g_d0,
           const float* __restrict__ g_nrm, float* __restrict__ g_x, float* __restrict__ g_p,
           float* __restrict__ g_nij, float* __restrict__ g_njp, float* __restrict__ g_nip_s,
           float* __restrict__ g_nj_c, float* __restrict__ g_ni_cs, long long BN, float coh, float radius,
           float st, float rho0, float cscale)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float pv = p[n], a = nip_s[n], bq = ni_cs[n];
    float gp = 0.0f, ga = 0.0f, gb = 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float xv = x[n * D + c];
        const float gd = g_d0 ? g_d0[n * D + c] : 0.0f;
        const float gn = g_nrm ? g_nrm[n * D + c] : 0.0f;
        const float g2 = -coh * radius * gd + gn * st / rho0 / cscale;  // d/d(nij2)
        const float gq = -gd;                                           // d/d(nijp)
        gp += gq * nij[n * D + c];
        g_nij[n * D + c] = gq * pv;
        g_x[n * D + c] = gq * a + g2 * bq;
        g_njp[n * D + c] = gd;
        ga += gq * xv;
        g_nj_c[n * D + c] = -g2;
        gb += g2 * xv;
    }
    g_p[n] = gp;
    g_nip_s[n] = ga;
    g_ni_cs[n] = gb;
}

__device__ __forceinline__ float pbf_scale(float ncount, float relax, float damp, bool* active)
{
    const float t = ncount / (1.0f + relax) - damp;
    *active = t > 0.0f;
    return (t > 0.0f ? t : 0.0f) + damp;
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf3_fwd(const float* __restrict__ x, const float* __restrict__ d0, const float* __restrict__ cd,
           const float* __restrict__ nrm, const float* __restrict__ ncount, float* __restrict__ xnew,
           long long BN, float relax, float damp)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float nc = ncount[n];
    bool act;
    const float scale = pbf_scale(nc, relax, damp, &act);
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float delta = d0[n * D + c] + (cd[n * D + c] - nrm[n * D + c] * nc);
        xnew[n * D + c] = x[n * D + c] + delta / scale;
    }
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf3_bwd(const float* __restrict__ d0, const float* __restrict__ cd, const float* __restrict__ nrm,
           const float* __restrict__ ncount, const float* __restrict__ g, float* __restrict__ g_d0,
           float* __restrict__ g_nrm, float* __restrict__ g_ncount, long long BN, float relax, float damp)
{
    // d/dx = g and d/d(cd) = d/d(d0): the caller aliases those tensors instead of copying them
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float nc = ncount[n];
    bool act;
    const float scale = pbf_scale(nc, relax, damp, &act);
    const float inv = 1.0f / scale;
    float gnc = 0.0f, gsc = 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float gv = g[n * D + c];
        const float gd = gv * inv;  // d/d(delta)
        const float nv = nrm[n * D + c];
        const float delta = d0[n * D + c] + (cd[n * D + c] - nv * nc);
        g_d0[n * D + c] = gd;
        g_nrm[n * D + c] = -gd * nc;
        gnc -= gd * nv;
        gsc -= gv * delta * inv * inv;
    }
    g_ncount[n] = gnc + (act ? gsc / (1.0f + relax) : 0.0f);
}

}  // namespace
}  // namespace spnb

using namespace spnb;

#define SPNB_GLUE_LAUNCH(KERNEL, ...)                                                          \
    do {                                                                                       \
        const int blocks = cdiv(BN, kGlueThreads);                                             \
        if (D == 3) KERNEL<3><<<blocks, kGlueThreads, 0, stream>>>(__VA_ARGS__);               \
        else if (D == 2) KERNEL<2><<<blocks, kGlueThreads, 0, stream>>>(__VA_ARGS__);          \
        else if (D == 1) KERNEL<1><<<blocks, kGlueThreads, 0, stream>>>(__VA_ARGS__);          \
        else {                                                                                 \
            set_error("spnb_pbf_*: ndims must be 1, 2 or 3");                                  \
            return 0;                                                                          \
        }                                                                                      \
        count_launches(1);                                                                     \
    } while (0)

extern "C" {

int spnb_pbf_stage1_forward(const float* x, const float* density, const float* nj, const float* ni_s, float* p,
                            float* xp, float* nij, long long BN, int D, float stiffness, float rho0,
                            void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !density || !nj || !ni_s || !p || !xp || !nij || BN <= 0) {
        set_error("spnb_pbf_stage1_forward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf1_fwd, x, density, nj, ni_s, p, xp, nij, BN, stiffness, rho0);
    return check_launch("spnb_pbf_stage1_forward") ? 1 : 0;
}

int spnb_pbf_stage1_backward(const float* x, const float* density, const float* ni_s, const float* g_p,
                             const float* g_xp, const float* g_nij, float* g_x, float* g_density, float* g_nj,
                             float* g_ni_s, long long BN, int D, float stiffness, float rho0, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !density || !ni_s || !g_x || !g_density || !g_nj || !g_ni_s || BN <= 0) {
        set_error("spnb_pbf_stage1_backward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf1_bwd, x, density, ni_s, g_p, g_xp, g_nij, g_x, g_density, g_nj, g_ni_s, BN, stiffness,
                     rho0);
    return check_launch("spnb_pbf_stage1_backward") ? 1 : 0;
}

int spnb_pbf_stage2_forward(const float* x, const float* p, const float* nij, const float* njp,
                            const float* nip_s, const float* nj_c, const float* ni_cs, float* d0, float* nrm,
                            long long BN, int D, float cohesion, float radius, float surface_tension, float rho0,
                            float constraint_scale, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !p || !nij || !njp || !nip_s || !nj_c || !ni_cs || !d0 || !nrm || BN <= 0) {
        set_error("spnb_pbf_stage2_forward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf2_fwd, x, p, nij, njp, nip_s, nj_c, ni_cs, d0, nrm, BN, cohesion, radius,
                     surface_tension, rho0, constraint_scale);
    return check_launch("spnb_pbf_stage2_forward") ? 1 : 0;
}

int spnb_pbf_stage2_backward(const float* x, const float* p, const float* nij, const float* nip_s,
                             const float* ni_cs, const float* g_d0, const float* g_nrm, float* g_x, float* g_p,
                             float* g_nij, float* g_njp, float* g_nip_s, float* g_nj_c, float* g_ni_cs,
                             long long BN, int D, float cohesion, float radius, float surface_tension, float rho0,
                             float constraint_scale, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !p || !nij || !nip_s || !ni_cs || !g_x || !g_p || !g_nij || !g_njp || !g_nip_s || !g_nj_c ||
        !g_ni_cs || BN <= 0) {
        set_error("spnb_pbf_stage2_backward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf2_bwd, x, p, nij, nip_s, ni_cs, g_d0, g_nrm, g_x, g_p, g_nij, g_njp, g_nip_s, g_nj_c,
                     g_ni_cs, BN, cohesion, radius, surface_tension, rho0, constraint_scale);
    return check_launch("spnb_pbf_stage2_backward") ? 1 : 0;
}

int spnb_pbf_stage3_forward(const float* x, const float* d0, const float* cd, const float* nrm,
                            const float* ncount, float* xnew, long long BN, int D, float relaxation, float damp,
                            void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !d0 || !cd || !nrm || !ncount || !xnew || BN <= 0) {
        set_error("spnb_pbf_stage3_forward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf3_fwd, x, d0, cd, nrm, ncount, xnew, BN, relaxation, damp);
    return check_launch("spnb_pbf_stage3_forward") ? 1 : 0;
}

int spnb_pbf_stage3_backward(const float* d0, const float* cd, const float* nrm, const float* ncount,
                             const float* g, float* g_d0, float* g_nrm, float* g_ncount, long long BN, int D,
                             float relaxation, float damp, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d0 || !cd || !nrm || !ncount || !g || !g_d0 || !g_nrm || !g_ncount || BN <= 0) {
        set_error("spnb_pbf_stage3_backward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf3_bwd, d0, cd, nrm, ncount, g, g_d0, g_nrm, g_ncount, BN, relaxation, damp);
    return check_launch("spnb_pbf_stage3_backward") ? 1 : 0;
}

}  // extern "C"
