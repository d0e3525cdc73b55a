// Launchers of the ConvSP ncells == 1 fast path (convsp_small.cu), used by the C ABI in convsp.cu.
#pragma once
#include "spnb_common.cuh"

namespace spnb {

bool convsp_small_supported(int D, int C, int O, int ncells);

void launch_convsp_fwd_small(const float* qlocs, const float* locs, const float* data,
                             const float* neighbors, const float* weight, const float* bias, int B,
                             int M, int N, int C, int D, int K, int O, float radius, int dis_norm,
                             int kernel_fn, float* out, cudaStream_t stream);

void launch_convsp_bwd_small(const float* qlocs, const float* locs, const float* data,
                             const float* neighbors, const float* weight, int B, int M, int N, int C,
                             int D, int K, int O, float radius, int dis_norm, int kernel_fn,
                             const float* grad_out, float* dqlocs, float* dlocs, float* ddata,
                             float* dweight, const int* sym_flag, int same, cudaStream_t stream, int q_off = 0,
                             int go_rows = 0);

}  // namespace spnb
