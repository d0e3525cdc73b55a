// Shared helpers for libspnb (sm_100a).  Compiled with -fmad=false: every float expression below is
// evaluated with separately rounded IEEE operations in source order, which is what the reference's
// CPU build (g++ without FMA) does -- predicates such as d < r*r must not flip at the boundary
// (SURVEY.md 7.2-2).  Where a fused multiply-add is wanted for speed it is spelled fmaf().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spnb.h"

#define SPNB_MAXD 20  // array bound for the runtime-ndims path (ndim <= SPNB_MAX_NDIM)

namespace spnb {

void set_error(const char* fmt, ...);
bool check_launch(const char* what);
void count_launches(int n);  // kernels (not memsets) enqueued by this process, for bench accounting

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- SPH kernel table -----------------------------------------------------------------------
// Expression ids: 0..11 are the kernels of kernels.py in alphabetical order (the public kernel_fn
// ids, kernels.py:123); 12.. are derivative expressions that are not themselves kernels.
enum ExprId {
    E_COHESION = 0, E_CONSTANT, E_DDEFAULT, E_DDEFAULT2, E_DEFAULT, E_DPRESSURE, E_DPRESSURE2,
    E_DSPIKY, E_INDIRECT, E_PRESSURE, E_SIGMOID, E_SPIKY,
    E_D_DDEFAULT2, E_D_DPRESSURE2, E_D_DSPIKY, E_D_COHESION, E_D_SIGMOID, E_D_INDIRECT, E_D_CONSTANT
};

// Host: expression id of dW/dd for kernel `fn` (DKERNELS table, kernels.py), and the
// d-independent double-precision prefix of an expression, evaluated with the association of the C
// expression strings so it is bit-identical to what the reference computes at run time.
int deriv_expr_of(int fn);
double expr_coef(int expr, float H);

struct SphParams {
    float H;      // support radius
    int w_expr;   // expression id of W
    int dw_expr;  // expression id of dW/dd
    double w_coef;
    double dw_coef;
};
SphParams make_sph_params(int kernel_fn, float radius);

// Value of expression `e` at distance d (caller has applied the d > H guard, common_funcs.h:57-79).
// Float/double promotions follow the C expression strings in kernels.py (M_PI is a double).
__device__ __forceinline__ float sph_eval(int e, float d, float H, double coef)
{
    switch (e) {
    case E_DEFAULT:   { float q = H * H - d * d; return (float)(coef * q * q * q); }
    case E_DDEFAULT:  { float q = H * H - d * d; return (float)(coef * q * q * d); }
    case E_DDEFAULT2: { float q = H * H * H * H - 6 * H * H * d * d + 5 * d * d * d * d;
                        return (float)(coef * q); }
    case E_D_DDEFAULT2: { float q = 20 * d * d * d - 12 * H * H * d; return (float)(coef * q); }
    case E_PRESSURE:  { float q = H - d; return (float)(coef * q * q * q); }
    case E_DPRESSURE: { float q = H - d; return (float)(coef * q * q); }
    case E_DPRESSURE2:{ float q = H - d; return (float)(coef * q); }
    case E_D_DPRESSURE2: return (float)coef;
    case E_INDIRECT:  return H - d;
    case E_D_INDIRECT: return -1.0f;
    case E_CONSTANT:  return 1.0f;
    case E_D_CONSTANT: return 0.0f;
    case E_SPIKY:     { float q = 1.0f - d / H; return (float)(coef * q * q); }
    case E_DSPIKY:    { float q = 1.0f - d / H; return (float)(coef * q / H); }
    case E_D_DSPIKY:  return (float)coef;
    case E_COHESION:  { float t = d / H; return -6.0f * t * t * t + 7 * t * t - 1; }
    case E_D_COHESION: return 2.0f * d * (7.0f * H - 9.0f * d) / (H * H * H);
    case E_SIGMOID:   return 1.0f / (1.0f + expf((d - 0.2f * H) * 20.0f / H));
    case E_D_SIGMOID: { float ex = expf((d - 0.2f * H) * 20.0f / H);
                        return -20.0f * ex / (H * (ex + 1.0f) * (ex + 1.0f)); }
    default: return 0.0f;
    }
}

// All-float evaluation of the same expressions (coefficient rounded to float, 1/H precomputed): within an ulp or
// two of sph_eval; used where the sum order already differs from the reference's (group and wide kernels).
// Convention: for E_DSPIKY the caller folds the expression's 1/H into the coefficient c.
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

// ---- hash-grid coordinate helpers (common_funcs.h:96-119) --------------------------------------
__device__ __forceinline__ int grid_coord_of(float x, float low, float edge)
{
    int g = __float2int_rz((x - low) / edge);
    return g >= 0 ? g : 0;
}
// int*float products truncated back to int, exactly like the reference's `c *= grid_dims[dd]`.
__device__ __forceinline__ int hash_term(int g, const float* gd, int dim, int D)
{
    if ((float)g >= gd[dim]) g = __float2int_rz(gd[dim] - 1);
    else if (g < 0) g = 0;
    int c = g;
    for (int dd = dim + 1; dd < D; ++dd) c = __float2int_rz((float)c * gd[dd]);
    return c;
}

__device__ __forceinline__ float fast_root_dim(int D)  // common_funcs.h:132-143
{
    if (D == 1) return 1.0f;
    if (D == 2) return 1.41421f;
    if (D == 3) return 1.73205f;
    return sqrtf((float)D);
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace spnb
