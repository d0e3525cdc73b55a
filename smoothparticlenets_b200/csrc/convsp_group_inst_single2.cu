// The single-layer family, ndim 2 (see convsp_group_inst_single3.cu).
#include "convsp_group.cuh"

namespace spnb {
namespace grp {
const SigEntry kSigsSingle2[] = {
    sig_entry<Sig<2, 1, 0x1u, 0x0u, 0xFu, 0x0u, 0x1u>>(), sig_entry<Sig<2, 1, 0x2u, 0x0u, 0xFu, 0x0u, 0x1u>>(),
    sig_entry<Sig<2, 1, 0x3u, 0x0u, 0xFu, 0x0u, 0x1u>>(), sig_entry<Sig<2, 1, 0x4u, 0x0u, 0xFu, 0x0u, 0x1u>>(),
    sig_entry<Sig<2, 1, 0x2u, 0xFu, 0xFu, 0x0u, 0x1u>>(), sig_entry<Sig<2, 1, 0x1u, 0xEu, 0xFu, 0x0u, 0x0u>>(),
};
const int kNumSigsSingle2 = sizeof(kSigsSingle2) / sizeof(kSigsSingle2[0]);
}  // namespace grp
}  // namespace spnb
