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

// ---- the ends of the step (fluid_sim.py:355-365, 412-424) ----------------------------------------------
//   integrate:  v1 = v + gravity * dt;  m = -(relu(-cap / (|v1| + 1e-4) + 1) - 1);  v2 = v1 * m;  x1 = x + v2 * dt
//   velocity :  w0 = (x - xs) / dt                       (new velocity from the position change)
//   viscosity:  w1 = w0 + c * (vj - w0 * vi_s)           (c = dt * viscosity / rest density)
struct GVec { float v[3]; };

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf_integrate_fwd(const float* __restrict__ x, const float* __restrict__ v, float* __restrict__ v2,
                    float* __restrict__ x1, long long BN, GVec g, float dt, float cap)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    float v1[D], ss = 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        v1[c] = v[n * D + c] + g.v[c] * dt;
        ss += v1[c] * v1[c];
    }
    const float s = cap / (sqrtf(ss) + 0.0001f);
    const float t = -s + 1.0f;
    const float m = -((t > 0.0f ? t : 0.0f) - 1.0f);
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float o = v1[c] * m;
        v2[n * D + c] = o;
        x1[n * D + c] = x[n * D + c] + o * dt;
    }
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf_integrate_bwd(const float* __restrict__ v, const float* __restrict__ g_v2, const float* __restrict__ g_x1,
                    float* __restrict__ g_v, long long BN, GVec g, float dt, float cap)
{
    // d/dx = g_x1: the caller aliases it
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    float v1[D], gt[D], ss = 0.0f, dot = 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        v1[c] = v[n * D + c] + g.v[c] * dt;
        ss += v1[c] * v1[c];
        gt[c] = (g_v2 ? g_v2[n * D + c] : 0.0f) + (g_x1 ? g_x1[n * D + c] * dt : 0.0f);
        dot += gt[c] * v1[c];
    }
    const float nn = sqrtf(ss);
    const float den = nn + 0.0001f;
    const float s = cap / den;
    const float t = -s + 1.0f;
    const float m = -((t > 0.0f ? t : 0.0f) - 1.0f);
    // m = s where s < 1 (capped), else 1;  ds/d|v1| = -cap / (|v1| + eps)^2;  d|v1|/dv1 = v1 / |v1|
    const float k = (t > 0.0f && nn > 0.0f) ? dot * (-cap / (den * den)) / nn : 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) g_v[n * D + c] = gt[c] * m + k * v1[c];
}

__global__ void __launch_bounds__(kGlueThreads)
k_pbf_velocity(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, float* __restrict__ o2,
               long long n_floats, float dt, int backward)
{
    // forward: o = (a - b) / dt;  backward (a = grad): o = a / dt, o2 = -(a / dt)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_floats) return;
    if (!backward) {
        o[i] = (a[i] - b[i]) / dt;
    } else {
        const float gq = a[i] / dt;
        o[i] = gq;
        if (o2) o2[i] = -gq;
    }
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf_viscosity_fwd(const float* __restrict__ w0, const float* __restrict__ vj, const float* __restrict__ vi_s,
                    float* __restrict__ w1, long long BN, float c)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float s = vi_s[n];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const float w = w0[n * D + k];
        w1[n * D + k] = w + c * (vj[n * D + k] - w * s);
    }
}

template <int D>
__global__ void __launch_bounds__(kGlueThreads)
k_pbf_viscosity_bwd(const float* __restrict__ w0, const float* __restrict__ vi_s, const float* __restrict__ g,
                    float* __restrict__ g_w0, float* __restrict__ g_vj, float* __restrict__ g_vi_s, long long BN,
                    float c)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= BN) return;
    const float s = vi_s[n];
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const float gv = g[n * D + k];
        const float gc = gv * c;  // d/d(vj - w0 * vi_s)
        g_vj[n * D + k] = gc;
        g_w0[n * D + k] = gv + -gc * s;
        acc += -gc * w0[n * D + k];
    }
    g_vi_s[n] = acc;
}

// ---- n-ary sum (backward of a fan-out) --------------------------------------------------------------------
// A tensor with k consumers receives k gradients, which autograd adds pairwise: k-1 launches that each
// re-read the running sum.  The fan-out op (pbf.py: fanout) hands every consumer its own alias of the tensor
// and adds all gradients here in one pass.
struct SumPtrs { const float* p[8]; };

__global__ void __launch_bounds__(kGlueThreads)
k_sum_n(SumPtrs in, int n, float* __restrict__ out, long long nfloat4, long long nfloats)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nfloat4) {
        float4 acc = reinterpret_cast<const float4*>(in.p[0])[i];
        for (int k = 1; k < n; ++k) {
            const float4 v = reinterpret_cast<const float4*>(in.p[k])[i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    } else {
        const long long j = nfloat4 * 4 + (i - nfloat4);  // tail elements
        if (j < nfloats) {
            float acc = in.p[0][j];
            for (int k = 1; k < n; ++k) acc += in.p[k][j];
            out[j] = acc;
        }
    }
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

int spnb_sum_n(const float* const* inputs_host, int n, float* out, long long nfloats, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!inputs_host || !out || n < 1 || n > 8 || nfloats <= 0) {
        set_error("spnb_sum_n: bad arguments (1 <= n <= 8)");
        return 0;
    }
    SumPtrs in;
    bool aligned = (reinterpret_cast<size_t>(out) & 15) == 0;
    for (int k = 0; k < 8; ++k) {
        in.p[k] = k < n ? inputs_host[k] : nullptr;
        if (k < n && !in.p[k]) {
            set_error("spnb_sum_n: null input");
            return 0;
        }
        if (k < n && (reinterpret_cast<size_t>(in.p[k]) & 15) != 0) aligned = false;
    }
    const long long n4 = aligned ? nfloats / 4 : 0;
    const long long threads = n4 + (nfloats - n4 * 4);
    k_sum_n<<<cdiv(threads, kGlueThreads), kGlueThreads, 0, stream>>>(in, n, out, n4, nfloats);
    count_launches(1);
    return check_launch("spnb_sum_n") ? 1 : 0;
}

int spnb_pbf_integrate_forward(const float* x, const float* v, float* v2, float* x1, long long BN, int D,
                               const float* gravity_host, float dt, float cap, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !v || !v2 || !x1 || !gravity_host || BN <= 0) {
        set_error("spnb_pbf_integrate_forward: bad arguments");
        return 0;
    }
    GVec g = {{0.0f, 0.0f, 0.0f}};
    for (int c = 0; c < D && c < 3; ++c) g.v[c] = gravity_host[c];
    SPNB_GLUE_LAUNCH(k_pbf_integrate_fwd, x, v, v2, x1, BN, g, dt, cap);
    return check_launch("spnb_pbf_integrate_forward") ? 1 : 0;
}

int spnb_pbf_integrate_backward(const float* v, const float* g_v2, const float* g_x1, float* g_v, long long BN,
                                int D, const float* gravity_host, float dt, float cap, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!v || !g_v || !gravity_host || BN <= 0) {
        set_error("spnb_pbf_integrate_backward: bad arguments");
        return 0;
    }
    GVec g = {{0.0f, 0.0f, 0.0f}};
    for (int c = 0; c < D && c < 3; ++c) g.v[c] = gravity_host[c];
    SPNB_GLUE_LAUNCH(k_pbf_integrate_bwd, v, g_v2, g_x1, g_v, BN, g, dt, cap);
    return check_launch("spnb_pbf_integrate_backward") ? 1 : 0;
}

int spnb_pbf_velocity(const float* a, const float* b, float* o, float* o2, long long n_floats, float dt,
                      int backward, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || !o || (!backward && !b) || n_floats <= 0) {
        set_error("spnb_pbf_velocity: bad arguments");
        return 0;
    }
    k_pbf_velocity<<<cdiv(n_floats, kGlueThreads), kGlueThreads, 0, stream>>>(a, b, o, o2, n_floats, dt, backward);
    count_launches(1);
    return check_launch("spnb_pbf_velocity") ? 1 : 0;
}

int spnb_pbf_viscosity_forward(const float* w0, const float* vj, const float* vi_s, float* w1, long long BN, int D,
                               float c, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!w0 || !vj || !vi_s || !w1 || BN <= 0) {
        set_error("spnb_pbf_viscosity_forward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf_viscosity_fwd, w0, vj, vi_s, w1, BN, c);
    return check_launch("spnb_pbf_viscosity_forward") ? 1 : 0;
}

int spnb_pbf_viscosity_backward(const float* w0, const float* vi_s, const float* g, float* g_w0, float* g_vj,
                                float* g_vi_s, long long BN, int D, float c, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!w0 || !vi_s || !g || !g_w0 || !g_vj || !g_vi_s || BN <= 0) {
        set_error("spnb_pbf_viscosity_backward: bad arguments");
        return 0;
    }
    SPNB_GLUE_LAUNCH(k_pbf_viscosity_bwd, w0, vi_s, g, g_w0, g_vj, g_vi_s, BN, c);
    return check_launch("spnb_pbf_viscosity_backward") ? 1 : 0;
}

}  // extern "C"
