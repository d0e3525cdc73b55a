/*
 * spn_oracle.c -- CPU restatement of the SmoothParticleNets particle-interaction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (smoothparticlenets_b200/) may link, load
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and there only as the checker / CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function below bit-for-bit
 * against the unmodified reference CPU extension (oracle/_ref, built by oracle/build_ref.py from
 * /root/reference/src/cpu_layer_funcs.cpp) when it is present, and against golden vectors generated
 * from it (tests/golden/, generator tests/golden/make_golden.py) when it is not.
 *
 * Each function cites the reference lines it restates.  The arithmetic is single precision with the
 * reference's evaluation order and its float/double promotions (the SPH kernel table mixes in
 * M_PI, a double); compile with -ffp-contract=off and without -march=native so no FMA is formed.
 * All tensors are float32, including indices, exactly as in the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SPNO_MAX_DIM 20 /* reference src/constants.h:7 */

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int spno_max_cartesian_dim(void) { return SPNO_MAX_DIM; }

/* ------------------------------------------------------------------------------------------------
 * SPH kernel table.  Restates python/SmoothParticleNets/kernels.py:13-121 as compiled through
 * setup.py:60-79 into KERNEL_W / KERNEL_DW, with ids = alphabetical order (kernels.py:123):
 *   0 cohesion 1 constant 2 ddefault 3 ddefault2 4 default 5 dpressure 6 dpressure2 7 dspiky
 *   8 indirect 9 pressure 10 sigmoid 11 spiky
 * The double-typed sub-expressions are the ones that contain M_PI in the C expression strings.
 * ---------------------------------------------------------------------------------------------- */
enum { K_COHESION, K_CONSTANT, K_DDEFAULT, K_DDEFAULT2, K_DEFAULT, K_DPRESSURE, K_DPRESSURE2,
       K_DSPIKY, K_INDIRECT, K_PRESSURE, K_SIGMOID, K_SPIKY, K_COUNT };

static double pi_times_pow(double lead, float H, int p)
{
    /* lead*M_PI*H*H*...*H (p factors), left to right, in double. */
    double v = lead * M_PI;
    int i;
    for (i = 0; i < p; ++i) v = v * H;
    return v;
}
static double pi_pow(float H, int p)
{
    /* M_PI*H*H*...*H (p factors), left to right, in double. */
    double v = M_PI;
    int i;
    for (i = 0; i < p; ++i) v = v * H;
    return v;
}

/* value of kernel expression `fn` (no d>H guard) */
static float sph_expr(int fn, float d, float H)
{
    switch (fn) {
    case K_DEFAULT: {
        float q = H * H - d * d;
        return (float)((315.0f / pi_times_pow(64.0f, H, 9)) * q * q * q);
    }
    case K_DDEFAULT: {
        float q = H * H - d * d;
        return (float)((-945.0f / pi_times_pow(32.0f, H, 9)) * q * q * d);
    }
    case K_DDEFAULT2: {
        float q = H * H * H * H - 6 * H * H * d * d + 5 * d * d * d * d;
        return (float)((-945.0f / pi_times_pow(32.0f, H, 9)) * q);
    }
    case K_PRESSURE: {
        float q = H - d;
        return (float)((15.0f / pi_pow(H, 6)) * q * q * q);
    }
    case K_DPRESSURE: {
        float q = H - d;
        return (float)((-45.0f / pi_pow(H, 6)) * q * q);
    }
    case K_DPRESSURE2: {
        float q = H - d;
        return (float)((90.0f / pi_pow(H, 6)) * q);
    }
    case K_INDIRECT:
        return H - d;
    case K_CONSTANT:
        return 1.0f;
    case K_SPIKY: {
        float q = 1.0f - d / H;
        return (float)(15.0f / pi_pow(H, 3) * q * q);
    }
    case K_DSPIKY: {
        float q = 1.0f - d / H;
        return (float)(-15.0f / pi_pow(H, 3) * 2.0f * q / H);
    }
    case K_COHESION: {
        float t = d / H;
        return -6.0f * t * t * t + 7 * t * t - 1;
    }
    case K_SIGMOID:
        return 1.0f / (1.0f + expf((d - 0.2f * H) * 20.0f / H));
    default:
        return 0.0f;
    }
}

/* value of the derivative expression of kernel `fn` (DKERNELS table, kernels.py) */
static float sph_dexpr(int fn, float d, float H)
{
    switch (fn) {
    case K_DEFAULT:   return sph_expr(K_DDEFAULT, d, H);
    case K_DDEFAULT:  return sph_expr(K_DDEFAULT2, d, H);
    case K_DDEFAULT2: {
        float q = 20 * d * d * d - 12 * H * H * d;
        return (float)((-945.0f / pi_times_pow(32.0f, H, 9)) * q);
    }
    case K_PRESSURE:   return sph_expr(K_DPRESSURE, d, H);
    case K_DPRESSURE:  return sph_expr(K_DPRESSURE2, d, H);
    case K_DPRESSURE2: return (float)(-90.0f / pi_pow(H, 6));
    case K_INDIRECT:   return -1.0f;
    case K_CONSTANT:   return 0.0f;
    case K_SPIKY:      return sph_expr(K_DSPIKY, d, H);
    case K_DSPIKY:     return (float)(-15.0f / pi_pow(H, 3) * 2.0f * (-1.0f / H) / H);
    case K_COHESION:   return 2.0f * d * (7.0f * H - 9.0f * d) / (H * H * H);
    case K_SIGMOID: {
        float e = expf((d - 0.2f * H) * 20.0f / H);
        return -20.0f * e / (H * (e + 1.0f) * (e + 1.0f));
    }
    default:
        return 0.0f;
    }
}

/* kernel_w / kernel_dw: common_funcs.h:57-79 (0 beyond the support, -1 for an unknown id) */
float spno_kernel_w(float d, float H, int fn)
{
    if (d > H) return 0.0f;
    if (fn < 0 || fn >= K_COUNT) return -1.0f;
    return sph_expr(fn, d, H);
}
float spno_kernel_dw(float d, float H, int fn)
{
    if (d > H) return 0.0f;
    if (fn < 0 || fn >= K_COUNT) return -1.0f;
    return sph_dexpr(fn, d, H);
}

/* ------------------------------------------------------------------------------------------------
 * Hash grid: bounds, cell keys, ordering, cell table, neighbour lists, reorder.
 * ---------------------------------------------------------------------------------------------- */

/* Grid bounds as ParticleCollision.forward computes them with float32 torch CPU ops
 * (python/SmoothParticleNets/ParticleCollision.py:174-181):
 *   grid_dims = ceil(clamp((upper-lower)/radius, 0, max_grid_dim))
 *   lower_bounds = (lower+upper)/2 - grid_dims*radius/2
 * `radius` arrives as the float32 rounding of the Python double, as it does in torch's scalar ops. */
void spno_grid_bounds(const float* locs, int B, int N, int D, float radius, int max_grid_dim,
                      float* low_out, float* dims_out)
{
    int b, i, k;
    for (b = 0; b < B; ++b) {
        for (k = 0; k < D; ++k) {
            float lo = locs[(size_t)b * N * D + k], hi = lo;
            for (i = 1; i < N; ++i) {
                float v = locs[((size_t)b * N + i) * D + k];
                if (v < lo) lo = v;
                if (v > hi) hi = v;
            }
            float ext = (hi - lo) / radius;
            if (ext < 0.0f) ext = 0.0f;
            if (ext > (float)max_grid_dim) ext = (float)max_grid_dim;
            float gd = ceilf(ext);
            float center = (lo + hi) / 2;
            dims_out[b * D + k] = gd;
            low_out[b * D + k] = center - gd * radius / 2;
        }
    }
}

/* loc2grid: common_funcs.h:96-104 */
static int grid_coord_of(float x, float low, float edge)
{
    int g = (int)((x - low) / edge);
    return g >= 0 ? g : 0;
}

/* partial_grid_hash: common_funcs.h:107-119.  Note the int*float products truncated back to int. */
static int hash_term(int g, const float* gdims, int dim, int D)
{
    int dd, c;
    if (g >= gdims[dim])
        g = (int)(gdims[dim] - 1);
    else if (g < 0)
        g = 0;
    c = g;
    for (dd = dim + 1; dd < D; ++dd) c = (int)(c * gdims[dd]);
    return c;
}

/* Per-particle cell keys: first loop of spn_hashgrid_order, cpu_layer_funcs.cpp:248-261. */
void spno_cell_keys(const float* locs, const float* low, const float* gdims, int B, int N, int D,
                    float edge, int32_t* keys)
{
    int b, i, k;
    for (b = 0; b < B; ++b)
        for (i = 0; i < N; ++i) {
            int h = 0;
            for (k = 0; k < D; ++k)
                h += hash_term(grid_coord_of(locs[((size_t)b * N + i) * D + k], low[b * D + k], edge),
                               gdims + b * D, k, D);
            keys[(size_t)b * N + i] = h;
        }
}

/* spn_hashgrid_order exactly as the CPU reference runs it (cpu_layer_funcs.cpp:230-290):
 * keys as floats, then an in-place SELECTION sort with swaps -- unstable within a cell. */
void spno_hashgrid_order_selection(const float* locs, const float* low, const float* gdims, int B,
                                   int N, int D, float edge, float* cellIDs, float* idxs)
{
    int b, i, j;
    int32_t* keys = (int32_t*)malloc(sizeof(int32_t) * (size_t)B * N);
    spno_cell_keys(locs, low, gdims, B, N, D, edge, keys);
    for (i = 0; i < B * N; ++i) {
        cellIDs[i] = (float)keys[i];
        idxs[i] = (float)(i % N);
    }
    free(keys);
    for (b = 0; b < B; ++b) {
        float* ck = cellIDs + (size_t)b * N;
        float* ci = idxs + (size_t)b * N;
        for (i = 0; i < N; ++i) {
            int best = (int)ck[i], at = i;
            for (j = i + 1; j < N; ++j)
                if (ck[j] < best) {
                    best = (int)ck[j];
                    at = j;
                }
            if (at != i) {
                float t = ck[i]; ck[i] = ck[at]; ck[at] = t;
                t = ci[i]; ci[i] = ci[at]; ci[at] = t;
            }
        }
    }
}

/* The ordering contract of the reference's GPU path (gpu_kernels.cu:308-327): a STABLE radix sort
 * of (cellID, original index) pairs, i.e. ties broken by ascending original index.  Implemented
 * here as a counting sort on the keys of spno_cell_keys.  This is what the B200 kernels must
 * reproduce bit-exactly. */
void spno_hashgrid_order_stable(const float* locs, const float* low, const float* gdims, int B,
                                int N, int D, float edge, float* cellIDs, float* idxs)
{
    int b, i;
    int32_t* keys = (int32_t*)malloc(sizeof(int32_t) * (size_t)B * N);
    spno_cell_keys(locs, low, gdims, B, N, D, edge, keys);
    for (b = 0; b < B; ++b) {
        const int32_t* kb = keys + (size_t)b * N;
        int32_t kmin = kb[0], kmax = kb[0];
        for (i = 1; i < N; ++i) {
            if (kb[i] < kmin) kmin = kb[i];
            if (kb[i] > kmax) kmax = kb[i];
        }
        size_t span = (size_t)((int64_t)kmax - kmin + 1);
        int32_t* cnt = (int32_t*)calloc(span + 1, sizeof(int32_t));
        for (i = 0; i < N; ++i) cnt[kb[i] - kmin + 1]++;
        for (size_t c = 0; c < span; ++c) cnt[c + 1] += cnt[c];
        for (i = 0; i < N; ++i) {
            int32_t p = cnt[kb[i] - kmin]++;
            cellIDs[(size_t)b * N + p] = (float)kb[i];
            idxs[(size_t)b * N + p] = (float)i;
        }
        free(cnt);
    }
    free(keys);
}

/* spn_reorder_data: cpu_layer_funcs.cpp:378-425.  data may be NULL (C = 0). */
void spno_reorder_data(const float* locs, const float* data, const float* idxs, float* nlocs,
                       float* ndata, int B, int N, int D, int C, int reverse)
{
    int b, i, k;
    for (b = 0; b < B; ++b)
        for (i = 0; i < N; ++i) {
            int from = (int)idxs[(size_t)b * N + i], to = i;
            if (reverse) {
                to = from;
                from = i;
            }
            for (k = 0; k < D; ++k)
                nlocs[((size_t)b * N + to) * D + k] = locs[((size_t)b * N + from) * D + k];
            if (data)
                for (k = 0; k < C; ++k)
                    ndata[((size_t)b * N + to) * C + k] = data[((size_t)b * N + from) * C + k];
        }
}

/* cellStart/cellEnd table from sorted keys: cpu_layer_funcs.cpp:320-346.  Tables are
 * [B, ncells] floats that the caller has zero-filled (ParticleCollision.py:264-265). */
void spno_cell_table(const float* sorted_ids, int B, int N, int ncells, float* starts, float* ends)
{
    int b, i;
    for (b = 0; b < B; ++b)
        for (i = 0; i < N; ++i) {
            int c = (int)sorted_ids[(size_t)b * N + i];
            if (i == 0)
                starts[(size_t)b * ncells + c] = (float)i;
            else {
                int p = (int)sorted_ids[(size_t)b * N + i - 1];
                if (c != p) {
                    starts[(size_t)b * ncells + c] = (float)i;
                    ends[(size_t)b * ncells + p] = (float)i;
                }
            }
            if (i == N - 1) ends[(size_t)b * ncells + c] = (float)(i + 1);
        }
}

/* compute_collisions for every query: common_funcs.h:875-948 driven by cpu_layer_funcs.cpp:348-373.
 * `coll` is [B, M, K], pre-filled with -1 by the caller (ParticleCollision.py:262-263). */
void spno_compute_collisions(const float* qlocs, const float* locs, const float* low,
                             const float* gdims, const float* starts, const float* ends, int B,
                             int M, int N, int D, int ncells, float edge, float radius, float* coll,
                             int K, int include_self)
{
    const float r2 = radius * radius;
    int b, q, k;
    for (b = 0; b < B; ++b)
        for (q = 0; q < M; ++q) {
            const float* x = qlocs + ((size_t)b * M + q) * D;
            float* row = coll + ((size_t)b * M + q) * K;
            int gc[SPNO_MAX_DIM], off[SPNO_MAX_DIM], found = 0;
            for (k = 0; k < D; ++k) {
                gc[k] = grid_coord_of(x[k], low[b * D + k], edge);
                off[k] = -1;
            }
            while (off[D - 1] <= 1 && found < K) {
                int ok = 1, cell = 0;
                for (k = 0; k < D && ok; ++k) {
                    int c = gc[k] + off[k];
                    if (c < 0 || c >= gdims[b * D + k])
                        ok = 0;
                    else
                        cell += hash_term(c, gdims + b * D, k, D);
                }
                if (ok) {
                    int i;
                    for (i = (int)starts[(size_t)b * ncells + cell];
                         i < ends[(size_t)b * ncells + cell] && found < K; ++i) {
                        const float* y = locs + ((size_t)b * N + i) * D;
                        float d = 0.0f;
                        for (k = 0; k < D; ++k) {
                            float t = x[k] - y[k];
                            d += t * t;
                        }
                        if (d < r2 && (d > 0 || include_self)) row[found++] = (float)i;
                    }
                }
                /* odometer over {-1,0,1}^D, dimension 0 fastest */
                ++off[0];
                for (k = 0; k < D - 1 && off[k] > 1; ++k) {
                    off[k] = -1;
                    ++off[k + 1];
                }
            }
            if (found < K) row[found] = -1.0f;
        }
}

/* ------------------------------------------------------------------------------------------------
 * ConvSP: compute_kernel_cells (common_funcs.h:439-583) driven by cpu_convsp
 * (cpu_layer_funcs.cpp:102-121).  Forward accumulates into `out` (pre-zeroed; bias is added by the
 * Python caller, convsp.py:172).  Backward: `out` holds grad_output and the four gradient arrays
 * (pre-zeroed) are accumulated in the reference's loop order; any of them may be NULL.
 * ---------------------------------------------------------------------------------------------- */
static float fast_root(float x) /* common_funcs.h:132-143 */
{
    if (x == 1.0f) return 1.0f;
    if (x == 2.0f) return 1.41421f;
    if (x == 3.0f) return 1.73205f;
    return sqrtf(x);
}
static float max_of(const float* v, int n) /* common_funcs.h:145-156 */
{
    float m = v[0];
    int i;
    for (i = 1; i < n; ++i)
        if (v[i] > m) m = v[i];
    return m;
}

void spno_convsp(const float* qlocs, const float* locs, const float* data, const float* neighbors,
                 const float* weight, int B, int M, int N, int C, int D, int K, int O, int ncells,
                 float radius, const float* ksize, const float* dil, int dis_norm, int kernel_fn,
                 float* out, float* dqlocs, float* dlocs, float* ddata, float* dweight)
{
    const int bwd = (dqlocs || dlocs || ddata || dweight);
    /* cull radius, common_funcs.h:481-484 */
    const float cull = radius + ((int)max_of(ksize, D) / 2) * max_of(dil, D) * fast_root((float)D);
    const float cull2 = cull * cull;
    const float rad2 = radius * radius;
    int b, q, jj, k, o, c;
    for (b = 0; b < B; ++b)
        for (q = 0; q < M; ++q) {
            const float* x = qlocs + ((size_t)b * M + q) * D;
            const float* nb = neighbors + ((size_t)b * M + q) * K;
            float* orow = out + ((size_t)b * M + q) * O;
            for (jj = 0; jj < K && nb[jj] >= 0; ++jj) {
                const int j = (int)nb[jj];
                const float* y = locs + ((size_t)b * N + j) * D;
                const float* dj = data + ((size_t)b * N + j) * C;
                float d = 0.0f;
                for (k = 0; k < D; ++k) d += (x[k] - y[k]) * (x[k] - y[k]);
                if (d > cull2) continue;

                int kidx[SPNO_MAX_DIM], cell;
                float disp[SPNO_MAX_DIM], dkw[SPNO_MAX_DIM];
                for (k = 0; k < D; ++k) kidx[k] = 0;
                for (cell = 0; kidx[D - 1] < ksize[D - 1]; ++cell) {
                    d = 0.0f;
                    for (k = 0; k < D; ++k) {
                        disp[k] = x[k] + (kidx[k] - ((int)ksize[k]) / 2) * dil[k] - y[k];
                        d += disp[k] * disp[k];
                    }
                    if (d < rad2) {
                        d = sqrtf(d);
                        float norm = 1.0f;
                        if (dis_norm && d > 0.0f) norm /= d;
                        const float kw = spno_kernel_w(d, radius, kernel_fn);
                        if (bwd)
                            for (k = 0; k < D; ++k)
                                dkw[k] = spno_kernel_dw(d, radius, kernel_fn) / d * disp[k];
                        for (o = 0; o < O; ++o)
                            for (c = 0; c < C; ++c) {
                                const float w = weight[((size_t)o * C + c) * ncells + cell];
                                if (!bwd) {
                                    orow[o] += w * dj[c] * kw * norm;
                                    continue;
                                }
                                if (ddata) ddata[((size_t)b * N + j) * C + c] += orow[o] * w * kw * norm;
                                if (dweight)
                                    dweight[((size_t)o * C + c) * ncells + cell] += orow[o] * dj[c] * kw * norm;
                                if (dqlocs && d > 0)
                                    for (k = 0; k < D; ++k)
                                        dqlocs[((size_t)b * M + q) * D + k] += w * dj[c] * norm * dkw[k] * orow[o];
                                if (dlocs && d > 0)
                                    for (k = 0; k < D; ++k)
                                        dlocs[((size_t)b * N + j) * D + k] += -w * dj[c] * norm * dkw[k] * orow[o];
                            }
                    }
                    /* odometer over kernel cells, dimension 0 fastest */
                    ++kidx[0];
                    for (k = 0; k < D - 1 && kidx[k] >= ksize[k]; ++k) {
                        kidx[k] = 0;
                        ++kidx[k + 1];
                    }
                }
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * ConvSDF: compute_sdf_kernel_cells (common_funcs.h:647-837) with point_in_coordinate_frame /
 * rotate_point (:203-235,317-325) and nlinear_interp (:327-384), driven by cpu_convsdf
 * (cpu_layer_funcs.cpp:201-227).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z, w; } quat;

static quat q_mul(quat a, quat b) /* common_funcs.h:181-190 */
{
    quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
static quat q_conj(quat q) /* common_funcs.h:171-179; the -1.0 literals make these double products */
{
    quat r = q;
    r.x = (float)(r.x * -1.0);
    r.y = (float)(r.y * -1.0);
    r.z = (float)(r.z * -1.0);
    return r;
}

/* rotate_point: common_funcs.h:203-235.  inverse != 0 applies the inverse rotation. */
static void rotate_vec(float* p, int D, const float* rot, int inverse)
{
    if (D == 2) {
        int sgn = inverse ? -1 : 1;
        float m = sqrtf(p[0] * p[0] + p[1] * p[1]);
        float th = atan2f(p[1], p[0]) + sgn * rot[0];
        p[0] = m * cosf(th);
        p[1] = m * sinf(th);
    } else if (D == 3) {
        quat r = { rot[0], rot[1], rot[2], rot[3] };
        quat v = { p[0], p[1], p[2], 0.0f };
        if (inverse)
            v = q_mul(q_conj(r), q_mul(v, r));
        else
            v = q_mul(r, q_mul(v, q_conj(r)));
        p[0] = v.x;
        p[1] = v.y;
        p[2] = v.z;
    }
}

/* rec_nlinear_interp: common_funcs.h:327-362.  The innermost lerp runs over the last dimension. */
static float lerp_rec(const float* grid, const float* gshape, int D, const float* frac, int* low,
                      int dim, float* grad)
{
    if (dim == D) {
        const float* p = grid;
        int i, j;
        for (i = 0; i < D; ++i) {
            int s = low[i];
            for (j = i + 1; j < D; ++j) s = (int)(s * gshape[j]);
            p += s;
        }
        return *p;
    }
    float g1[SPNO_MAX_DIM], g2[SPNO_MAX_DIM];
    float a = lerp_rec(grid, gshape, D, frac, low, dim + 1, grad ? g1 : NULL);
    low[dim] += 1;
    float b = lerp_rec(grid, gshape, D, frac, low, dim + 1, grad ? g2 : NULL);
    low[dim] -= 1;
    if (grad) {
        int i;
        grad[dim] = -a + b;
        for (i = dim + 1; i < D; ++i) grad[i] = (1 - frac[dim]) * g1[i] + frac[dim] * g2[i];
    }
    return (1 - frac[dim]) * a + frac[dim] * b;
}

/* nlinear_interp: common_funcs.h:364-384 (cell-centred samples: x/cell - 0.5, the 0.5 a double) */
static float sdf_sample(const float* grid, const float* gshape, int D, float cell, const float* p,
                        float* grad)
{
    int low[SPNO_MAX_DIM], i;
    float frac[SPNO_MAX_DIM];
    for (i = 0; i < D; ++i) {
        float u = (float)(p[i] / cell - 0.5);
        low[i] = (int)u;
        frac[i] = u - floorf(u);
    }
    float v = lerp_rec(grid, gshape, D, frac, low, 0, grad);
    if (grad)
        for (i = 0; i < D; ++i) grad[i] /= cell;
    return v;
}

static int sdf_contains(const float* p, const float* gshape, int D, float cell)
{
    int i;
    for (i = 0; i < D; ++i)
        if (p[i] < 0.5 * cell || p[i] > (gshape[i] - 0.5) * cell) return 0;
    return 1;
}

void spno_convsdf(const float* locs, int B, int N, int D, const float* idxs, const float* poses,
                  const float* scales, int S, int pose_len, const float* sdfs,
                  const float* sdf_offsets, const float* sdf_shapes, const float* weight,
                  const float* bias, int O, int ncells, const float* ksize, const float* dil,
                  float max_distance, float* out, float* dlocs, float* dweight, float* dposes)
{
    const int bwd = (dweight || dlocs || dposes);
    int* live = (int*)malloc(sizeof(int) * (S > 0 ? S : 1));
    float* centre_v = (float*)malloc(sizeof(float) * (S > 0 ? S : 1));
    /* reach of the kernel footprint, common_funcs.h:689-692 */
    const float reach = ((int)max_of(ksize, D) / 2) * max_of(dil, D) * fast_root((float)D);
    int b, n, o, m, i, k;
    for (b = 0; b < B; ++b)
        for (n = 0; n < N; ++n)
            for (o = 0; o < O; ++o) {
                const float* x = locs + ((size_t)b * N + n) * D;
                float p[SPNO_MAX_DIM];
                /* pre-cull with the kernel centre, common_funcs.h:699-738 */
                for (m = 0; m < S; ++m) {
                    live[m] = 1;
                    centre_v[m] = 0;
                }
                for (m = 0; m < S; ++m) {
                    const int mm = (int)idxs[b * S + m];
                    if (mm < 0) {
                        live[m] = 0;
                        continue;
                    }
                    const float* shp = sdf_shapes + (size_t)mm * (D + 1);
                    const float* pose = poses + ((size_t)b * S + m) * pose_len;
                    const float cell = shp[D] * scales[b * S + m];
                    for (i = 0; i < D; ++i) p[i] = x[i] - pose[i];
                    rotate_vec(p, D, pose + D, 1);
                    int inside = 1;
                    for (i = 0; i < D && live[m]; ++i) {
                        if (p[i] + reach < 0.5 * cell || p[i] - reach > (shp[i] - 0.5) * cell) live[m] = 0;
                        if (p[i] < 0.5 * cell || p[i] > (shp[i] - 0.5) * cell) inside = 0;
                    }
                    if (!live[m] || !inside) continue;
                    centre_v[m] = sdf_sample(sdfs + (int)sdf_offsets[mm], shp, D, cell, p, NULL) *
                                  scales[b * S + m];
                }
                for (m = 0; m < S; ++m)
                    if (live[m] && centre_v[m] - reach > max_distance) live[m] = 0;

                float* optr = out + ((size_t)b * N + n) * O + o;
                if (!bwd) *optr = 0;
                int kidx[SPNO_MAX_DIM], cell_i;
                for (k = 0; k < D; ++k) kidx[k] = 0;
                for (cell_i = 0; kidx[D - 1] < ksize[D - 1]; ++cell_i) {
                    float pt[SPNO_MAX_DIM], best_g[SPNO_MAX_DIM], g[SPNO_MAX_DIM];
                    for (i = 0; i < D; ++i) {
                        pt[i] = x[i] + (kidx[i] - ((int)ksize[i] / 2)) * dil[i];
                        best_g[i] = 0.0f;
                    }
                    float best = max_distance;
                    int best_m = -1;
                    for (m = 0; m < S; ++m) {
                        if (!live[m]) continue;
                        const int mm = (int)idxs[b * S + m];
                        const float* shp = sdf_shapes + (size_t)mm * (D + 1);
                        const float* pose = poses + ((size_t)b * S + m) * pose_len;
                        const float cell = shp[D] * scales[b * S + m];
                        for (i = 0; i < D; ++i) p[i] = pt[i] - pose[i];
                        rotate_vec(p, D, pose + D, 1);
                        if (!sdf_contains(p, shp, D, cell)) continue;
                        float v = sdf_sample(sdfs + (int)sdf_offsets[mm], shp, D, cell, p, bwd ? g : NULL) *
                                  scales[b * S + m];
                        if (v < best) {
                            best = v;
                            best_m = m;
                            if (bwd) {
                                for (i = 0; i < D; ++i) g[i] *= scales[b * S + m];
                                rotate_vec(g, D, pose + D, 0);
                                for (i = 0; i < D; ++i) best_g[i] = g[i];
                            }
                        }
                    }
                    const float w = weight[(size_t)o * ncells + cell_i];
                    if (!bwd)
                        *optr += w * best;
                    else {
                        if (dweight) dweight[(size_t)o * ncells + cell_i] += best * (*optr);
                        if (dlocs)
                            for (i = 0; i < D; ++i)
                                dlocs[((size_t)b * N + n) * D + i] += best_g[i] * (*optr) * w;
                        /* Translation part only: the reference's rotation entries
                         * (common_funcs.h:810-820) are overwritten by finite differences in
                         * convsdf.py:211-224, so they are not part of the contract. */
                        if (dposes && best_m >= 0)
                            for (i = 0; i < D; ++i)
                                dposes[((size_t)b * S + best_m) * pose_len + i] += -best_g[i] * (*optr) * w;
                    }
                    ++kidx[0];
                    for (k = 0; k < D - 1 && kidx[k] >= ksize[k]; ++k) {
                        kidx[k] = 0;
                        ++kidx[k + 1];
                    }
                }
                if (!bwd) *optr += bias[o];
            }
    free(live);
    free(centre_v);
}

/* ------------------------------------------------------------------------------------------------
 * ParticleProjection / ImageProjection (3-D only).  Restates compute_particle_projection and
 * compute_image_projection (src/common_funcs.h:979-1052 and 1087-1183) and their CPU drivers
 * (src/cpu_layer_funcs.cpp:428-486, 489-556): particles are already in camera space.
 * The pixel loops run on FLOAT counters exactly as the reference writes them.
 * ---------------------------------------------------------------------------------------------- */
/* go == NULL: forward, adds the Gaussians into out [B,H,W] (caller zero-fills).
 * go != NULL: backward, go = dL/dout [B,H,W]; adds into dlocs [B,N,3] (caller zero-fills). */
void spno_particleprojection(const float* locs, int B, int N, float camera_fl, int width, int height,
                             float filter_std, float filter_scale, const float* depth_mask, float* out,
                             const float* go, float* dlocs)
{
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) {
            const float* rr = locs + ((size_t)b * N + n) * 3;
            const float rx = rr[0], ry = rr[1], rz = rr[2];
            if (rz <= 0) continue;
            const float px = rx * camera_fl / rz + width / 2;
            const float py = ry * camera_fl / rz + height / 2;
            const int s = ceilf(filter_std * 2);
            const float s2 = s * s;
            const float f = filter_scale / (filter_std * sqrtf(2 * M_PI));
            const float std2 = filter_std * filter_std;
            float i, j;
            for (i = (px - s > 0 ? px - s : 0); i < width && i < px + s + 1; i += 1)
                for (j = (py - s > 0 ? py - s : 0); j < height && j < py + s + 1; j += 1) {
                    const int ii = i, jj = j;
                    const size_t pix = (size_t)b * width * height + (size_t)jj * width + ii;
                    const float depth_val = depth_mask[pix];
                    if (depth_val > 0.0f && depth_val < rz) continue;
                    const float xi = ii + 0.5f, yj = jj + 0.5f;
                    const float d2 = (xi - px) * (xi - px) + (yj - py) * (yj - py);
                    if (d2 > s2) continue;
                    const float v = f * expf(-d2 / (2.0f * std2));
                    if (!go) {
                        out[pix] += v;
                    } else {
                        const float g = go[pix];
                        float* dl = dlocs + ((size_t)b * N + n) * 3;
                        dl[0] += g * (xi - px) * v / std2 * camera_fl / rz;
                        dl[1] += g * (yj - py) * v / std2 * camera_fl / rz;
                        dl[2] += g * v / std2 * camera_fl / (rz * rz) * ((xi - px) * -rx + (yj - py) * -ry);
                    }
                }
        }
}

/* go == NULL: forward into out [B,N,C] (caller zero-fills).  go != NULL: backward, go = dL/dout;
 * adds into dlocs [B,N,3] and dimage [B,C,H,W] (caller zero-fills both). */
void spno_imageprojection(const float* locs, const float* image, int B, int N, float camera_fl, int width,
                          int height, int channels, const float* depth_mask, float* out, const float* go,
                          float* dlocs, float* dimage)
{
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) {
            const float* rr = locs + ((size_t)b * N + n) * 3;
            const float rx = rr[0], ry = rr[1], rz = rr[2];
            if (rz <= 0) continue;
            const float px = rx * camera_fl / rz + width / 2;
            const float py = ry * camera_fl / rz + height / 2;
            if (px <= 0.5 || px >= width - 0.5 || py <= 0.5 || py >= height - 0.5) continue;
            const int ii = px, jj = py;
            const float depth_val = depth_mask[(size_t)b * width * height + (size_t)jj * width + ii];
            if (depth_val > 0.0f && depth_val < rz) continue;
            for (int c = 0; c < channels; ++c) {
                const int lowi = px - 0.5, highi = px + 0.5, lowj = py - 0.5, highj = py + 0.5;
                const float di = px - 0.5 - lowi, dj = py - 0.5 - lowj;
                const size_t plane = ((size_t)b * channels + c) * width * height;
                const float* ip = image + plane;
                const float vll = ip[lowj * width + lowi], vlh = ip[highj * width + lowi];
                const float vhl = ip[lowj * width + highi], vhh = ip[highj * width + highi];
                const float v = vll * (1 - di) * (1 - dj) + vlh * (1 - di) * dj + vhl * di * (1 - dj) + vhh * di * dj;
                if (!go) {
                    out[((size_t)b * N + n) * channels + c] += v;
                } else {
                    const float g = go[((size_t)b * N + n) * channels + c];
                    const float doutpx = -vll * (1 - dj) + -vlh * dj + vhl * (1 - dj) + vhh * dj;
                    const float doutpy = -vll * (1 - di) + vlh * (1 - di) + -vhl * di + vhh * di;
                    float* dl = dlocs + ((size_t)b * N + n) * 3;
                    dl[0] += g * camera_fl / rz * doutpx;
                    dl[1] += g * camera_fl / rz * doutpy;
                    dl[2] += g * -rx * camera_fl / (rz * rz) * doutpx + g * -ry * camera_fl / (rz * rz) * doutpy;
                    float* dp = dimage + plane;
                    dp[lowj * width + lowi] += g * (1 - di) * (1 - dj);
                    dp[highj * width + lowi] += g * (1 - di) * dj;
                    dp[lowj * width + highi] += g * di * (1 - dj);
                    dp[highj * width + highi] += g * di * dj;
                }
            }
        }
}
