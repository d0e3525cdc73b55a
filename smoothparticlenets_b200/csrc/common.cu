// Host-side common code of libspnb: error reporting and the SPH kernel coefficient table.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "spnb_common.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace spnb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;

void count_launches(int n) { g_launches += (unsigned long long)n; }

bool check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return false;
    }
    return true;
}

// dW/dd expression of each kernel: the DKERNELS table of kernels.py.
int deriv_expr_of(int fn)
{
    switch (fn) {
    case E_DEFAULT: return E_DDEFAULT;
    case E_DDEFAULT: return E_DDEFAULT2;
    case E_DDEFAULT2: return E_D_DDEFAULT2;
    case E_PRESSURE: return E_DPRESSURE;
    case E_DPRESSURE: return E_DPRESSURE2;
    case E_DPRESSURE2: return E_D_DPRESSURE2;
    case E_INDIRECT: return E_D_INDIRECT;
    case E_CONSTANT: return E_D_CONSTANT;
    case E_SPIKY: return E_DSPIKY;
    case E_DSPIKY: return E_D_DSPIKY;
    case E_COHESION: return E_D_COHESION;
    case E_SIGMOID: return E_D_SIGMOID;
    default: return E_D_CONSTANT;
    }
}

// M_PI*H*H*...*H with p factors of H, left to right in double (H promoted from float), optionally
// led by a float literal: the denominators of the expression strings in kernels.py.
static double pi_h(double lead, float H, int p)
{
    double v = lead * M_PI;
    for (int i = 0; i < p; ++i) v = v * H;
    return v;
}
static double pi_h(float H, int p)
{
    double v = M_PI;
    for (int i = 0; i < p; ++i) v = v * H;
    return v;
}

double expr_coef(int e, float H)
{
    switch (e) {
    case E_DEFAULT: return 315.0f / pi_h(64.0f, H, 9);
    case E_DDEFAULT:
    case E_DDEFAULT2:
    case E_D_DDEFAULT2: return -945.0f / pi_h(32.0f, H, 9);
    case E_PRESSURE: return 15.0f / pi_h(H, 6);
    case E_DPRESSURE: return -45.0f / pi_h(H, 6);
    case E_DPRESSURE2: return 90.0f / pi_h(H, 6);
    case E_D_DPRESSURE2: return -90.0f / pi_h(H, 6);
    case E_SPIKY: return 15.0f / pi_h(H, 3);
    case E_DSPIKY: return -15.0f / pi_h(H, 3) * 2.0f;
    case E_D_DSPIKY: return -15.0f / pi_h(H, 3) * 2.0f * (-1.0f / H) / H;
    default: return 0.0;
    }
}

SphParams make_sph_params(int kernel_fn, float radius)
{
    SphParams p;
    p.H = radius;
    p.w_expr = kernel_fn;
    p.dw_expr = deriv_expr_of(kernel_fn);
    p.w_coef = expr_coef(p.w_expr, radius);
    p.dw_coef = expr_coef(p.dw_expr, radius);
    return p;
}

}  // namespace spnb

extern "C" {

int spnb_version(void) { return 100; }
const char* spnb_last_error(void) { return spnb::g_err; }
int spnb_max_cartesian_dim(void) { return SPNB_MAX_NDIM + 1; }
unsigned long long spnb_launch_count(void) { return spnb::g_launches; }

}  // extern "C"
