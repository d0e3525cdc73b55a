// ParticleProjection / ImageProjection for sm_100a (3-D particles already in camera space).
//
// Replaces kernel_particleprojection + cuda_particleprojection and kernel_imageprojection +
// cuda_imageprojection (reference src/gpu_kernels.cu:561-712, gpu_kernels.h:111-140) and the math of
// compute_particle_projection / compute_image_projection (src/common_funcs.h:979-1052, 1087-1183).
//
// The reference runs one thread per particle, which walks the whole (2s+2)^2 pixel window of the Gaussian
// serially and, in the backward pass, adds every pixel's term to dlocs with a float atomic.  Here one WARP owns a
// particle: the window's pixel columns and rows are enumerated once, with the reference's float loop counters
// (so exactly the same pixels are visited), and the lanes share the pixels with the image x coordinate fastest --
// forward atomics of a warp fall on consecutive addresses -- and the backward reduces its three sums with
// shuffles and writes dlocs once, without atomics.  ImageProjection: one warp per particle, lanes over channels.
#include "spnb_common.cuh"

namespace spnb {

namespace {

constexpr int kProjThreads = 256;
constexpr int kProjWarps = kProjThreads / 32;
constexpr int kMaxWin = 64;  // pixel columns / rows of a window enumerated per warp (filter_std <= 15.5)

__device__ __forceinline__ float warp_sum_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The reference's loop `for (i = MAX(p - s, 0); i < limit && i < p + s + 1; i += 1)` on a float counter: writes the
// visited integer coordinates (int)i to `dst` (at most kMaxWin; returns the count, or -1 if there are more).
__device__ __forceinline__ int enumerate_axis(float p, int s, int limit, int* dst, int lane)
{
    const float first = fmaxf(p - s, 0.0f);
    const float stop = p + s + 1;
    int count = 0;
    for (int k0 = 0; k0 < kMaxWin + 32; k0 += 32) {
        float i = first;
        for (int t = 0; t < k0 + lane; ++t) i += 1;  // the same sequence of roundings as the reference's counter
        const bool ok = i < limit && i < stop;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok && k0 + lane < kMaxWin) dst[k0 + lane] = (int)i;
        count += __popc(m);
        if (m != 0xffffffffu) break;
        if (k0 + 32 >= kMaxWin + 32) return -1;
    }
    return count > kMaxWin ? -1 : count;
}

// BWD == false: out[b, y, x] += Gaussian of every particle (out zero-filled by the launcher).
// BWD == true : dlocs[b, n, :] = sum over the window of go * d(value)/d(locs)   (written, not accumulated)
template <bool BWD>
__global__ void __launch_bounds__(kProjThreads)
k_particle_projection(const float* __restrict__ locs, long long BN, int N, float camera_fl, int width, int height,
                      float filter_std, float filter_scale, const float* __restrict__ depth_mask,
                      float* out, const float* __restrict__ go, float* __restrict__ dlocs)
{
    __shared__ int s_cols[kProjWarps][kMaxWin], s_rows[kProjWarps][kMaxWin];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p = (long long)blockIdx.x * kProjWarps + warp;
    if (p >= BN) return;
    const int b = (int)(p / N);
    const float rx = locs[p * 3 + 0], ry = locs[p * 3 + 1], rz = locs[p * 3 + 2];
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    if (rz > 0) {  // behind the camera: no contribution
        const float px = rx * camera_fl / rz + width / 2;
        const float py = ry * camera_fl / rz + height / 2;
        const int s = (int)ceilf(filter_std * 2);
        const float s2 = (float)(s * s);
        const float f = filter_scale / (filter_std * sqrtf((float)(2 * 3.14159265358979323846)));
        const float std2 = filter_std * filter_std;
        const size_t plane = (size_t)b * width * height;
        const int ni = enumerate_axis(px, s, width, s_cols[warp], lane);
        const int nj = enumerate_axis(py, s, height, s_rows[warp], lane);
        __syncwarp();
        if (ni >= 0 && nj >= 0) {
            for (int idx = lane; idx < ni * nj; idx += 32) {
                const int ii = s_cols[warp][idx % ni], jj = s_rows[warp][idx / ni];
                const size_t pix = plane + (size_t)jj * width + ii;
                const float depth_val = depth_mask[pix];
                if (depth_val > 0.0f && depth_val < rz) continue;
                const float xi = ii + 0.5f, yj = jj + 0.5f;
                const float d2 = (xi - px) * (xi - px) + (yj - py) * (yj - py);
                if (d2 > s2) continue;
                const float v = f * expf(-d2 / (2.0f * std2));
                if (!BWD) {
                    atomicAdd(out + pix, v);
                } else {
                    const float g = go[pix];
                    a0 += g * (xi - px) * v / std2 * camera_fl / rz;
                    a1 += g * (yj - py) * v / std2 * camera_fl / rz;
                    a2 += g * v / std2 * camera_fl / (rz * rz) * ((xi - px) * -rx + (yj - py) * -ry);
                }
            }
        } else if (lane == 0) {
            // windows wider than kMaxWin pixels: the reference's loops as they are, on one lane
            float i, j;
            for (i = fmaxf(px - s, 0.0f); i < width && i < px + s + 1; i += 1)
                for (j = fmaxf(py - s, 0.0f); j < height && j < py + s + 1; j += 1) {
                    const int ii = (int)i, jj = (int)j;
                    const size_t pix = plane + (size_t)jj * width + ii;
                    const float depth_val = depth_mask[pix];
                    if (depth_val > 0.0f && depth_val < rz) continue;
                    const float xi = ii + 0.5f, yj = jj + 0.5f;
                    const float d2 = (xi - px) * (xi - px) + (yj - py) * (yj - py);
                    if (d2 > s2) continue;
                    const float v = f * expf(-d2 / (2.0f * std2));
                    if (!BWD) {
                        atomicAdd(out + pix, v);
                    } else {
                        const float g = go[pix];
                        a0 += g * (xi - px) * v / std2 * camera_fl / rz;
                        a1 += g * (yj - py) * v / std2 * camera_fl / rz;
                        a2 += g * v / std2 * camera_fl / (rz * rz) * ((xi - px) * -rx + (yj - py) * -ry);
                    }
                }
        }
    }
    if (BWD) {
        a0 = warp_sum_f(a0);
        a1 = warp_sum_f(a1);
        a2 = warp_sum_f(a2);
        if (lane == 0) {
            dlocs[p * 3 + 0] = a0;
            dlocs[p * 3 + 1] = a1;
            dlocs[p * 3 + 2] = a2;
        }
    }
}

// BWD == false: out[b, n, c] = bilinear sample of image[b, c] at the particle's pixel position (0 if not visible).
// BWD == true : dlocs[b, n, :] written; dimage accumulated with float atomics (zero-filled by the launcher).
template <bool BWD>
__global__ void __launch_bounds__(kProjThreads)
k_image_projection(const float* __restrict__ locs, const float* __restrict__ image, long long BN, int N,
                   float camera_fl, int width, int height, int channels, const float* __restrict__ depth_mask,
                   float* __restrict__ out, const float* __restrict__ go, float* __restrict__ dlocs, float* dimage)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p = (long long)blockIdx.x * kProjWarps + warp;
    if (p >= BN) return;
    const int b = (int)(p / N);
    const float rx = locs[p * 3 + 0], ry = locs[p * 3 + 1], rz = locs[p * 3 + 2];
    bool visible = rz > 0;
    float px = 0.0f, py = 0.0f;
    if (visible) {
        px = rx * camera_fl / rz + width / 2;
        py = ry * camera_fl / rz + height / 2;
        // the reference compares against double literals (0.5, width - 0.5)
        visible = !((double)px <= 0.5 || (double)px >= width - 0.5 || (double)py <= 0.5 || (double)py >= height - 0.5);
    }
    if (visible) {
        const float depth_val = depth_mask[(size_t)b * width * height + (size_t)(int)py * width + (int)px];
        if (depth_val > 0.0f && depth_val < rz) visible = false;
    }
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    if (visible) {
        const int lowi = (int)((double)px - 0.5), highi = (int)((double)px + 0.5);
        const int lowj = (int)((double)py - 0.5), highj = (int)((double)py + 0.5);
        const float di = (float)((double)px - 0.5 - lowi), dj = (float)((double)py - 0.5 - lowj);
        for (int c = lane; c < channels; c += 32) {
            const size_t plane = ((size_t)b * channels + c) * width * height;
            const float* ip = image + plane;
            const float vll = ip[lowj * width + lowi], vlh = ip[highj * width + lowi];
            const float vhl = ip[lowj * width + highi], vhh = ip[highj * width + highi];
            if (!BWD) {
                out[p * channels + c] = vll * (1 - di) * (1 - dj) + vlh * (1 - di) * dj + vhl * di * (1 - dj) + vhh * di * dj;
            } else {
                const float g = go[p * channels + c];
                if (dlocs) {
                    const float doutpx = -vll * (1 - dj) + -vlh * dj + vhl * (1 - dj) + vhh * dj;
                    const float doutpy = -vll * (1 - di) + vlh * (1 - di) + -vhl * di + vhh * di;
                    a0 += g * camera_fl / rz * doutpx;
                    a1 += g * camera_fl / rz * doutpy;
                    a2 += g * -rx * camera_fl / (rz * rz) * doutpx + g * -ry * camera_fl / (rz * rz) * doutpy;
                }
                if (dimage) {
                    float* dp = dimage + plane;
                    atomicAdd(dp + lowj * width + lowi, g * (1 - di) * (1 - dj));
                    atomicAdd(dp + highj * width + lowi, g * (1 - di) * dj);
                    atomicAdd(dp + lowj * width + highi, g * di * (1 - dj));
                    atomicAdd(dp + highj * width + highi, g * di * dj);
                }
            }
        }
    } else if (!BWD) {
        for (int c = lane; c < channels; c += 32) out[p * channels + c] = 0.0f;
    }
    if (BWD && dlocs) {
        a0 = warp_sum_f(a0);
        a1 = warp_sum_f(a1);
        a2 = warp_sum_f(a2);
        if (lane == 0) {
            dlocs[p * 3 + 0] = a0;
            dlocs[p * 3 + 1] = a1;
            dlocs[p * 3 + 2] = a2;
        }
    }
}

bool validate_proj(const char* fn, int B, int N, int width, int height, float camera_fl)
{
    if (B <= 0 || N <= 0 || width <= 0 || height <= 0) {
        set_error("%s: non-positive size", fn);
        return false;
    }
    if (!(camera_fl > 0)) {
        set_error("%s: camera_fl must be positive", fn);
        return false;
    }
    return true;
}

}  // namespace
}  // namespace spnb

using namespace spnb;

extern "C" {

int spnb_particleprojection_forward(const float* locs, int B, int N, float camera_fl, int width, int height,
                                    float filter_std, float filter_scale, const float* depth_mask, float* out,
                                    void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate_proj("spnb_particleprojection_forward", B, N, width, height, camera_fl)) return 0;
    if (!locs || !depth_mask || !out || !(filter_std > 0)) {
        set_error("spnb_particleprojection_forward: null pointer or non-positive filter_std");
        return 0;
    }
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * width * height, stream);
    const long long BN = (long long)B * N;
    k_particle_projection<false><<<(unsigned)cdiv(BN, kProjWarps), kProjThreads, 0, stream>>>(
        locs, BN, N, camera_fl, width, height, filter_std, filter_scale, depth_mask, out, nullptr, nullptr);
    count_launches(1);
    return check_launch("spnb_particleprojection_forward") ? 1 : 0;
}

int spnb_particleprojection_backward(const float* locs, int B, int N, float camera_fl, int width, int height,
                                     float filter_std, float filter_scale, const float* depth_mask,
                                     const float* grad_out, float* dlocs, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate_proj("spnb_particleprojection_backward", B, N, width, height, camera_fl)) return 0;
    if (!locs || !depth_mask || !grad_out || !dlocs || !(filter_std > 0)) {
        set_error("spnb_particleprojection_backward: null pointer or non-positive filter_std");
        return 0;
    }
    const long long BN = (long long)B * N;
    k_particle_projection<true><<<(unsigned)cdiv(BN, kProjWarps), kProjThreads, 0, stream>>>(
        locs, BN, N, camera_fl, width, height, filter_std, filter_scale, depth_mask, nullptr, grad_out, dlocs);
    count_launches(1);
    return check_launch("spnb_particleprojection_backward") ? 1 : 0;
}

int spnb_imageprojection_forward(const float* locs, const float* image, int B, int N, float camera_fl, int width,
                                 int height, int channels, const float* depth_mask, float* out, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate_proj("spnb_imageprojection_forward", B, N, width, height, camera_fl)) return 0;
    if (!locs || !image || !depth_mask || !out || channels <= 0) {
        set_error("spnb_imageprojection_forward: null pointer or no channels");
        return 0;
    }
    const long long BN = (long long)B * N;
    k_image_projection<false><<<(unsigned)cdiv(BN, kProjWarps), kProjThreads, 0, stream>>>(
        locs, image, BN, N, camera_fl, width, height, channels, depth_mask, out, nullptr, nullptr, nullptr);
    count_launches(1);
    return check_launch("spnb_imageprojection_forward") ? 1 : 0;
}

int spnb_imageprojection_backward(const float* locs, const float* image, int B, int N, float camera_fl, int width,
                                  int height, int channels, const float* depth_mask, const float* grad_out,
                                  float* dlocs, float* dimage, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!validate_proj("spnb_imageprojection_backward", B, N, width, height, camera_fl)) return 0;
    if (!locs || !image || !depth_mask || !grad_out || channels <= 0) {
        set_error("spnb_imageprojection_backward: null pointer or no channels");
        return 0;
    }
    if (!dlocs && !dimage) return 1;
    if (dimage) cudaMemsetAsync(dimage, 0, sizeof(float) * (size_t)B * channels * width * height, stream);
    const long long BN = (long long)B * N;
    k_image_projection<true><<<(unsigned)cdiv(BN, kProjWarps), kProjThreads, 0, stream>>>(
        locs, image, BN, N, camera_fl, width, height, channels, depth_mask, nullptr, grad_out, dlocs, dimage);
    count_launches(1);
    return check_launch("spnb_imageprojection_backward") ? 1 : 0;
}

}  // extern "C"
