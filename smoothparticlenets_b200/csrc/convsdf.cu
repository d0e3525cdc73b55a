// ConvSDF forward / backward for sm_100a.
//
// Replaces kernel_convsdf + cuda_convsdf (reference src/gpu_kernels.cu:129-235) and the math of
// compute_sdf_kernel_cells / point_in_coordinate_frame / rotate_point / nlinear_interp
// (src/common_funcs.h:203-235, 317-384, 647-837).
//
// One thread per query location (b, n); the reference uses one thread per (b, n, out-channel) and
// therefore repeats every SDF lookup nkernels times.  Per-object pose records are staged once per
// block in shared memory (a block never straddles two scenes), the n-linear interpolation is
// unrolled at compile time (no recursion, no device stack limit), and the min-over-objects value
// of a kernel cell is computed once and contracted with all output channels.  The float/double
// promotions of the reference's bounds tests and of `x/cell - 0.5` are kept (see the oracle).
#include "spnb_common.cuh"

namespace spnb {

constexpr int kSdfThreads = 128;
constexpr int kMaxObjects = 256;  // SDF objects per scene supported by the pose staging
constexpr int kSdfOChunk = 8;

template <int D>
struct ObjRec {
    float t[D];
    float rot[D == 3 ? 4 : 1];
    float scale;
    float cell;     // cell size * scale
    float gshape[D];
    double hi[D];   // (shape - 0.5) * cell, in double as the reference evaluates it
    long long off;  // offset of this SDF in the atlas, -1 if the slot is unused (idx < 0)
};

// rotate_point (common_funcs.h:203-235); inverse applies the inverse rotation.
template <int D>
__device__ __forceinline__ void rotate_vec(float* p, const float* rot, bool inverse)
{
    if (D == 2) {
        const int sgn = inverse ? -1 : 1;
        const float m = sqrtf(p[0] * p[0] + p[1] * p[1]);
        const float th = atan2f(p[1], p[0]) + sgn * rot[0];
        p[0] = m * cosf(th);
        p[1] = m * sinf(th);
    } else if (D == 3) {
        // quaternion products written out for q = (rx,ry,rz,rw), v = (px,py,pz,0)
        const float rx = rot[0], ry = rot[1], rz = rot[2], rw = rot[3];
        float ax, ay, az, aw;  // first product
        float cx, cy, cz;      // conjugate components
        cx = -rx; cy = -ry; cz = -rz;
        if (inverse) {
            // a = v * r ; result = conj(r) * a
            aw = 0.0f * rw - p[0] * rx - p[1] * ry - p[2] * rz;
            ax = 0.0f * rx + p[0] * rw + p[1] * rz - p[2] * ry;
            ay = 0.0f * ry + p[1] * rw + p[2] * rx - p[0] * rz;
            az = 0.0f * rz + p[2] * rw + p[0] * ry - p[1] * rx;
            p[0] = rw * ax + cx * aw + cy * az - cz * ay;
            p[1] = rw * ay + cy * aw + cz * ax - cx * az;
            p[2] = rw * az + cz * aw + cx * ay - cy * ax;
        } else {
            // a = v * conj(r) ; result = r * a
            aw = 0.0f * rw - p[0] * cx - p[1] * cy - p[2] * cz;
            ax = 0.0f * cx + p[0] * rw + p[1] * cz - p[2] * cy;
            ay = 0.0f * cy + p[1] * rw + p[2] * cx - p[0] * cz;
            az = 0.0f * cz + p[2] * rw + p[0] * cy - p[1] * cx;
            p[0] = rw * ax + rx * aw + ry * az - rz * ay;
            p[1] = rw * ay + ry * aw + rz * ax - rx * az;
            p[2] = rw * az + rz * aw + rx * ay - ry * ax;
        }
    }
}

// Vector part of conj(l) * (0,v) * r for quaternions stored xyzw.  rotate_point's inverse rotation
// is B(r, r, v); it is bilinear in (l, r), which is what the analytic pose gradient differentiates.
__device__ __forceinline__ void quat_sandwich(const float* l, const float* r, const float* v, float* o)
{
    const float aw = -v[0] * r[0] - v[1] * r[1] - v[2] * r[2];
    const float ax = v[0] * r[3] + v[1] * r[2] - v[2] * r[1];
    const float ay = v[1] * r[3] + v[2] * r[0] - v[0] * r[2];
    const float az = v[2] * r[3] + v[0] * r[1] - v[1] * r[0];
    o[0] = l[3] * ax - l[0] * aw - l[1] * az + l[2] * ay;
    o[1] = l[3] * ay - l[1] * aw - l[2] * ax + l[0] * az;
    o[2] = l[3] * az - l[2] * aw - l[0] * ay + l[1] * ax;
}

// d value / d rotation parameters of the object that owns a kernel cell, given the gradient `gl` of
// the value in the object's frame, the frame point `pl` and the world offset `d` = point - t.
// The parameters are differentiated as they enter rotate_point (common_funcs.h:203-235): the 2-D
// angle, or the four quaternion components as independent variables (no renormalisation).
template <int D>
__device__ __forceinline__ void rotation_grad(const float* rot, const float* gl, const float* pl,
                                              const float* d, float* out)
{
    if (D == 2) {
        out[0] = gl[0] * pl[1] - gl[1] * pl[0];
    } else if (D == 3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float e[4] = {0.0f, 0.0f, 0.0f, 0.0f}, u[3], w[3];
            e[k] = 1.0f;
            quat_sandwich(e, rot, d, u);
            quat_sandwich(rot, e, d, w);
            out[k] = gl[0] * (u[0] + w[0]) + gl[1] * (u[1] + w[1]) + gl[2] * (u[2] + w[2]);
        }
    }
}

// rec_nlinear_interp (common_funcs.h:327-362) unrolled at compile time.  `grad` (if GRAD) receives
// d value / d frac per dimension.
template <int D, int DIM, bool GRAD>
struct Lerp {
    __device__ __forceinline__ static float run(const float* __restrict__ atlas, long long off,
                                                long long last, const float* gshape,
                                                const float* frac, int* low, float* grad)
    {
        float g1[D], g2[D];
        const float a = Lerp<D, DIM + 1, GRAD>::run(atlas, off, last, gshape, frac, low, g1);
        low[DIM] += 1;
        const float b = Lerp<D, DIM + 1, GRAD>::run(atlas, off, last, gshape, frac, low, g2);
        low[DIM] -= 1;
        if (GRAD) {
            grad[DIM] = -a + b;
#pragma unroll
            for (int i = DIM + 1; i < D; ++i) grad[i] = (1 - frac[DIM]) * g1[i] + frac[DIM] * g2[i];
        }
        return (1 - frac[DIM]) * a + frac[DIM] * b;
    }
};
template <int D, bool GRAD>
struct Lerp<D, D, GRAD> {
    __device__ __forceinline__ static float run(const float* __restrict__ atlas, long long off,
                                                long long last, const float* gshape, const float*,
                                                int* low, float*)
    {
        long long idx = off;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            int s = low[i];
#pragma unroll
            for (int j = i + 1; j < D; ++j) s = __float2int_rz((float)s * gshape[j]);
            idx += s;
        }
        // The reference reads one element past a row / the atlas at the exact upper bound
        // (weight 0); keep the same flat indexing but never leave the atlas.
        idx = idx < 0 ? 0 : (idx > last ? last : idx);
        return atlas[idx];
    }
};

// nlinear_interp (common_funcs.h:364-384)
template <int D, bool GRAD>
__device__ __forceinline__ float sdf_sample(const float* __restrict__ atlas, long long off,
                                            long long last, const float* gshape, float cell,
                                            const float* p, float* grad)
{
    int low[D];
    float frac[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float u = (float)((double)(p[i] / cell) - 0.5);
        low[i] = __float2int_rz(u);
        frac[i] = u - floorf(u);
    }
    const float v = Lerp<D, 0, GRAD>::run(atlas, off, last, gshape, frac, low, grad);
    if (GRAD) {
#pragma unroll
        for (int i = 0; i < D; ++i) grad[i] /= cell;
    }
    return v;
}

template <int D>
__device__ __forceinline__ bool sdf_contains(const ObjRec<D>& r, const float* p)
{
#pragma unroll
    for (int i = 0; i < D; ++i)
        if ((double)p[i] < 0.5 * (double)r.cell || (double)p[i] > r.hi[i]) return false;
    return true;
}

template <int D>
__device__ __forceinline__ void to_frame(const ObjRec<D>& r, const float* pt, float* p)
{
#pragma unroll
    for (int i = 0; i < D; ++i) p[i] = pt[i] - r.t[i];
    rotate_vec<D>(p, r.rot, true);
}

// BWD == false: out[b,n,o] = bias[o] + sum_cell w[o,cell] * min(max_distance, min_m sdf_m(cell))
// BWD == true : `go` = grad_output; dlocs overwritten per thread, dweight / dposes accumulated via
//               shared memory + one global atomic per element per block (zero-filled by launcher).
template <int D, bool BWD>
__global__ void __launch_bounds__(kSdfThreads)
k_convsdf(const float* __restrict__ locs, int N, const float* __restrict__ idxs,
          const float* __restrict__ poses, const float* __restrict__ scales, int S, int pose_len,
          const float* __restrict__ atlas, long long atlas_len,
          const float* __restrict__ sdf_offsets, const float* __restrict__ sdf_shapes, int nsdfs,
          const float* __restrict__ weight, const float* __restrict__ bias, int O, int ncells,
          const float* __restrict__ ksize, const float* __restrict__ dilation, float max_distance,
          float* __restrict__ out, const float* __restrict__ go, float* __restrict__ dlocs,
          float* dweight, float* dposes, int dw_in_smem, int rot_grads)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    ObjRec<D>* recs = reinterpret_cast<ObjRec<D>*>(s_raw);
    float* s_dw = reinterpret_cast<float*>(recs + S);  // [O*ncells] if dw_in_smem
    float* s_dp = s_dw + (dw_in_smem ? O * ncells : 0); // [S*pose_len] if dposes
    const int b = blockIdx.y;
    const int n = blockIdx.x * kSdfThreads + threadIdx.x;

    for (int m = threadIdx.x; m < S; m += kSdfThreads) {
        ObjRec<D> r;
        const int mm = (int)idxs[b * S + m];
        const float* pose = poses + ((size_t)b * S + m) * pose_len;
#pragma unroll
        for (int i = 0; i < D; ++i) r.t[i] = pose[i];
#pragma unroll
        for (int i = 0; i < (D == 3 ? 4 : 1); ++i) r.rot[i] = (D == 1) ? 0.0f : pose[D + i];
        r.scale = scales[b * S + m];
        r.off = -1;
        r.cell = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) { r.gshape[i] = 0.0f; r.hi[i] = 0.0; }
        if (mm >= 0 && mm < nsdfs) {
            const float* shp = sdf_shapes + (size_t)mm * (D + 1);
            r.cell = shp[D] * r.scale;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                r.gshape[i] = shp[i];
                r.hi[i] = ((double)shp[i] - 0.5) * (double)r.cell;
            }
            r.off = (long long)(int)sdf_offsets[mm];
        }
        recs[m] = r;
    }
    if (BWD) {
        if (dw_in_smem)
            for (int i = threadIdx.x; i < O * ncells; i += kSdfThreads) s_dw[i] = 0.0f;
        if (dposes)
            for (int i = threadIdx.x; i < S * pose_len; i += kSdfThreads) s_dp[i] = 0.0f;
    }
    __syncthreads();

    if (n < N) {
        float x[D], dil[D];
        int ks[D], half[D];
        float maxdil = dilation[0], maxks = ksize[0];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            x[i] = locs[((size_t)b * N + n) * D + i];
            dil[i] = dilation[i];
            ks[i] = (int)ksize[i];
            half[i] = (int)ksize[i] / 2;
            if (dilation[i] > maxdil) maxdil = dilation[i];
            if (ksize[i] > maxks) maxks = ksize[i];
        }
        const float reach = ((int)maxks / 2) * maxdil * fast_root_dim(D);
        const long long last = atlas_len - 1;

        // pre-cull with the kernel centre (common_funcs.h:699-738)
        uint32_t live[kMaxObjects / 32];
#pragma unroll
        for (int i = 0; i < kMaxObjects / 32; ++i) live[i] = 0;
        for (int m = 0; m < S; ++m) {
            const ObjRec<D>& r = recs[m];
            if (r.off < 0) continue;
            float p[D];
            to_frame<D>(r, x, p);
            bool keep = true, inside = true;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                if (keep) {
                    if ((double)(p[i] + reach) < 0.5 * (double)r.cell || (double)(p[i] - reach) > r.hi[i])
                        keep = false;
                    if ((double)p[i] < 0.5 * (double)r.cell || (double)p[i] > r.hi[i]) inside = false;
                }
            }
            if (!keep) continue;
            float cv = 0.0f;
            if (inside) cv = sdf_sample<D, false>(atlas, r.off, last, r.gshape, r.cell, p, nullptr) * r.scale;
            if (cv - reach > max_distance) continue;
            live[m >> 5] |= 1u << (m & 31);
        }

        float a_dl[D];
#pragma unroll
        for (int i = 0; i < D; ++i) a_dl[i] = 0.0f;

        for (int o0 = 0; o0 < O; o0 += kSdfOChunk) {
            const int on = min(kSdfOChunk, O - o0);
            float acc[kSdfOChunk], g_o[kSdfOChunk];
#pragma unroll
            for (int o = 0; o < kSdfOChunk; ++o) {
                acc[o] = 0.0f;
                g_o[o] = (BWD && o < on) ? go[((size_t)b * N + n) * O + o0 + o] : 0.0f;
            }
            int kidx[D];
#pragma unroll
            for (int i = 0; i < D; ++i) kidx[i] = 0;
            for (int cell = 0; cell < ncells; ++cell) {
                float pt[D];
#pragma unroll
                for (int i = 0; i < D; ++i) pt[i] = x[i] + (kidx[i] - half[i]) * dil[i];
                float best = max_distance, coef = 0.0f;
                int best_m = -1;
                float best_g[D], best_gl[D], best_pl[D];
#pragma unroll
                for (int i = 0; i < D; ++i) best_g[i] = best_gl[i] = best_pl[i] = 0.0f;
                for (int m = 0; m < S; ++m) {
                    if (!((live[m >> 5] >> (m & 31)) & 1u)) continue;
                    const ObjRec<D>& r = recs[m];
                    float p[D], g[D];
                    to_frame<D>(r, pt, p);
                    if (!sdf_contains<D>(r, p)) continue;
                    const float v = sdf_sample<D, BWD>(atlas, r.off, last, r.gshape, r.cell, p, g) * r.scale;
                    if (v < best) {
                        best = v;
                        best_m = m;
                        if (BWD) {
#pragma unroll
                            for (int i = 0; i < D; ++i) {
                                g[i] *= r.scale;
                                best_gl[i] = g[i];
                                best_pl[i] = p[i];
                            }
                            rotate_vec<D>(g, r.rot, false);
#pragma unroll
                            for (int i = 0; i < D; ++i) best_g[i] = g[i];
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < kSdfOChunk; ++o) {
                    if (o < on) {
                        const float w = weight[(size_t)(o0 + o) * ncells + cell];
                        if (!BWD) {
                            acc[o] += w * best;
                        } else {
                            if (dweight) {
                                const float v = best * g_o[o];
                                if (dw_in_smem) atomicAdd(&s_dw[(o0 + o) * ncells + cell], v);
                                else atomicAdd(dweight + (size_t)(o0 + o) * ncells + cell, v);
                            }
#pragma unroll
                            for (int i = 0; i < D; ++i) {
                                a_dl[i] += best_g[i] * g_o[o] * w;
                                if (dposes && best_m >= 0)
                                    atomicAdd(&s_dp[best_m * pose_len + i], -best_g[i] * g_o[o] * w);
                            }
                            coef += g_o[o] * w;
                        }
                    }
                }
                if (BWD && D > 1 && rot_grads && dposes && best_m >= 0) {
                    const ObjRec<D>& r = recs[best_m];
                    float d[D], dr[D == 3 ? 4 : 1];
#pragma unroll
                    for (int i = 0; i < D; ++i) d[i] = pt[i] - r.t[i];
                    rotation_grad<D>(r.rot, best_gl, best_pl, d, dr);
#pragma unroll
                    for (int i = 0; i < (D == 3 ? 4 : 1); ++i)
                        atomicAdd(&s_dp[best_m * pose_len + D + i], dr[i] * coef);
                }
                ++kidx[0];
#pragma unroll
                for (int i = 0; i < D - 1; ++i)
                    if (kidx[i] >= ks[i]) {
                        kidx[i] = 0;
                        ++kidx[i + 1];
                    }
            }
            if (!BWD)
                for (int o = 0; o < on; ++o) out[((size_t)b * N + n) * O + o0 + o] = acc[o] + bias[o0 + o];
        }
        if (BWD && dlocs) {
#pragma unroll
            for (int i = 0; i < D; ++i) dlocs[((size_t)b * N + n) * D + i] = a_dl[i];
        }
    }
    if (BWD) {
        __syncthreads();
        if (dweight && dw_in_smem)
            for (int i = threadIdx.x; i < O * ncells; i += kSdfThreads) {
                const float v = s_dw[i];
                if (v != 0.0f) atomicAdd(dweight + i, v);
            }
        if (dposes)
            for (int i = threadIdx.x; i < S * pose_len; i += kSdfThreads) {
                const float v = s_dp[i];
                if (v != 0.0f) atomicAdd(dposes + (size_t)b * S * pose_len + i, v);
            }
    }
}

template <int D, bool BWD>
static int launch_convsdf(const float* locs, int B, int N, const float* idxs, const float* poses,
                          const float* scales, int S, int pose_len, const float* sdfs,
                          size_t sdfs_len, const float* sdf_offsets, const float* sdf_shapes,
                          int nsdfs, const float* weight, const float* bias, int O, int ncells,
                          const float* ksize, const float* dil, float max_distance, float* out,
                          const float* go, float* dlocs, float* dweight, float* dposes,
                          int rot_grads, cudaStream_t stream)
{
    size_t smem = sizeof(ObjRec<D>) * (size_t)S;
    int dw_in_smem = 0;
    if (BWD) {
        if (dweight && sizeof(float) * (size_t)O * ncells <= 32 * 1024) {
            dw_in_smem = 1;
            smem += sizeof(float) * (size_t)O * ncells;
        }
        if (dposes) smem += sizeof(float) * (size_t)S * pose_len;
        if (dweight) cudaMemsetAsync(dweight, 0, sizeof(float) * (size_t)O * ncells, stream);
        if (dposes) cudaMemsetAsync(dposes, 0, sizeof(float) * (size_t)B * S * pose_len, stream);
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_convsdf<D, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const dim3 grid(cdiv(N, kSdfThreads), B);
    k_convsdf<D, BWD><<<grid, kSdfThreads, smem, stream>>>(
        locs, N, idxs, poses, scales, S, pose_len, sdfs, (long long)sdfs_len, sdf_offsets,
        sdf_shapes, nsdfs, weight, bias, O, ncells, ksize, dil, max_distance, out, go, dlocs,
        dweight, dposes, dw_in_smem, rot_grads);
    count_launches(1);
    return check_launch(BWD ? "spnb_convsdf_backward" : "spnb_convsdf_forward") ? 1 : 0;
}

static bool validate_sdf(const char* fn, int B, int N, int D, int S, int pose_len, int O, int ncells,
                         int nsdfs, size_t sdfs_len)
{
    if (B <= 0 || N <= 0 || S <= 0 || O <= 0 || ncells <= 0 || nsdfs <= 0 || sdfs_len == 0) {
        set_error("%s: non-positive size", fn);
        return false;
    }
    if (D < 1 || D > 3) {
        set_error("%s: ndims=%d, only 1-, 2- and 3-D are supported (convsdf.py:43-47)", fn, D);
        return false;
    }
    const int need = D + (D == 3 ? 4 : (D == 2 ? 1 : 0));
    if (pose_len != need) {
        set_error("%s: pose_len=%d, expected %d for ndims=%d", fn, pose_len, need, D);
        return false;
    }
    if (S > kMaxObjects) {
        set_error("%s: %d SDF objects per scene > supported maximum %d", fn, S, kMaxObjects);
        return false;
    }
    return true;
}

}  // namespace spnb

using namespace spnb;

extern "C" {

int spnb_convsdf_forward(const float* locs, int B, int N, int D, const float* idxs,
                         const float* poses, const float* scales, int S, int pose_len,
                         const float* sdfs, size_t sdfs_len, const float* sdf_offsets,
                         const float* sdf_shapes, int nsdfs, const float* weight, const float* bias,
                         int O, int ncells, const float* kernel_size, const float* dilation,
                         float max_distance, float* out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate_sdf("spnb_convsdf_forward", B, N, D, S, pose_len, O, ncells, nsdfs, sdfs_len)) return 0;
    if (!locs || !idxs || !poses || !scales || !sdfs || !sdf_offsets || !sdf_shapes || !weight ||
        !bias || !kernel_size || !dilation || !out) {
        set_error("spnb_convsdf_forward: null pointer");
        return 0;
    }
#define GO(DD)                                                                                     \
    return launch_convsdf<DD, false>(locs, B, N, idxs, poses, scales, S, pose_len, sdfs, sdfs_len, \
                                     sdf_offsets, sdf_shapes, nsdfs, weight, bias, O, ncells,      \
                                     kernel_size, dilation, max_distance, out, nullptr, nullptr,   \
                                     nullptr, nullptr, 0, stream)
    if (D == 1) GO(1);
    if (D == 2) GO(2);
    GO(3);
#undef GO
}

static int convsdf_backward_impl(const char* fn, int rot_grads, const float* locs, int B, int N, int D, const float* idxs,
                          const float* poses, const float* scales, int S, int pose_len,
                          const float* sdfs, size_t sdfs_len, const float* sdf_offsets,
                          const float* sdf_shapes, int nsdfs, const float* weight, int O, int ncells,
                          const float* kernel_size, const float* dilation, float max_distance,
                          const float* grad_out, float* dlocs, float* dweight, float* dposes,
                          void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate_sdf(fn, B, N, D, S, pose_len, O, ncells, nsdfs, sdfs_len)) return 0;
    if (!locs || !idxs || !poses || !scales || !sdfs || !sdf_offsets || !sdf_shapes || !weight ||
        !kernel_size || !dilation || !grad_out) {
        set_error("%s: null pointer", fn);
        return 0;
    }
#define GO(DD)                                                                                     \
    return launch_convsdf<DD, true>(locs, B, N, idxs, poses, scales, S, pose_len, sdfs, sdfs_len,  \
                                    sdf_offsets, sdf_shapes, nsdfs, weight, nullptr, O, ncells,    \
                                    kernel_size, dilation, max_distance, nullptr, grad_out, dlocs, \
                                    dweight, dposes, rot_grads, stream)
    if (D == 1) GO(1);
    if (D == 2) GO(2);
    GO(3);
#undef GO
}

#define SDF_BWD_ARGS                                                                              \
    locs, B, N, D, idxs, poses, scales, S, pose_len, sdfs, sdfs_len, sdf_offsets, sdf_shapes,     \
        nsdfs, weight, O, ncells, kernel_size, dilation, max_distance, grad_out, dlocs, dweight,  \
        dposes, stream_

int spnb_convsdf_backward(const float* locs, int B, int N, int D, const float* idxs,
                          const float* poses, const float* scales, int S, int pose_len,
                          const float* sdfs, size_t sdfs_len, const float* sdf_offsets,
                          const float* sdf_shapes, int nsdfs, const float* weight, int O, int ncells,
                          const float* kernel_size, const float* dilation, float max_distance,
                          const float* grad_out, float* dlocs, float* dweight, float* dposes,
                          void* stream_)
{
    return convsdf_backward_impl("spnb_convsdf_backward", 0, SDF_BWD_ARGS);
}

int spnb_convsdf_backward_analytic(const float* locs, int B, int N, int D, const float* idxs,
                                   const float* poses, const float* scales, int S, int pose_len,
                                   const float* sdfs, size_t sdfs_len, const float* sdf_offsets,
                                   const float* sdf_shapes, int nsdfs, const float* weight, int O,
                                   int ncells, const float* kernel_size, const float* dilation,
                                   float max_distance, const float* grad_out, float* dlocs,
                                   float* dweight, float* dposes, void* stream_)
{
    return convsdf_backward_impl("spnb_convsdf_backward_analytic", 1, SDF_BWD_ARGS);
}
#undef SDF_BWD_ARGS

}  // extern "C"
