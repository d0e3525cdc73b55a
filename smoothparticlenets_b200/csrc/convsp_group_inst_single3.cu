// The single-layer family, ndim 3: any of the 12 kernels with or without dis_norm (evaluated through a switch per
// pair: kernel id 0xF = run time), 1..4 input channels from a data tensor, the position tensor itself, or ones.
// This is the fast path of a per-layer spn.ConvSP call with kernel_size 1 on lists that carry tile lists.
#include "convsp_group.cuh"

namespace spnb {
namespace grp {
const SigEntry kSigsSingle3[] = {
    sig_entry<Sig<3, 1, 0x1u, 0x0u, 0xFu, 0x0u, 0x1u>>(), sig_entry<Sig<3, 1, 0x2u, 0x0u, 0xFu, 0x0u, 0x1u>>(),
    sig_entry<Sig<3, 1, 0x3u, 0x0u, 0xFu, 0x0u, 0x1u>>(), sig_entry<Sig<3, 1, 0x4u, 0x0u, 0xFu, 0x0u, 0x1u>>(),
    sig_entry<Sig<3, 1, 0x3u, 0xFu, 0xFu, 0x0u, 0x1u>>(), sig_entry<Sig<3, 1, 0x1u, 0xEu, 0xFu, 0x0u, 0x0u>>(),
};
const int kNumSigsSingle3 = sizeof(kSigsSingle3) / sizeof(kSigsSingle3[0]);
}  // namespace grp
}  // namespace spnb
