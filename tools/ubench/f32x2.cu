// Micro-benchmark: packed fma.rn.f32x2 (FFMA2) vs scalar FFMA issue throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int MODE>
__global__ void k(float* out, int iters, float s)
{
    float a[8], b = s, c = s * 0.5f;
    unsigned long long pa[8], pb, pc;
    for (int i = 0; i < 8; ++i) {
        a[i] = threadIdx.x + i;
        pa[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f);
    }
    pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);
            else pa[i] = fma2(pa[i], pb, pc);
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += MODE == 0 ? a[i] : __uint_as_float((unsigned)(pa[i] >> 32)) + __uint_as_float((unsigned)pa[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main()
{
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.0001f);
            else k<1><<<148 * 8, 256>>>(out, iters, 1.0001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double instr = (double)148 * 8 * 256 / 32 * iters * 8;
            printf("mode %d (%s): %.3f ms, %.1f G warp-instr/s, %.2f TFLOP/s\n", mode, mode ? "fma.rn.f32x2" : "fma.rn.f32",
                   ms, instr / ms / 1e6, instr * 32 * (mode ? 4 : 2) / ms / 1e9);
        }
    }
    return 0;
}
