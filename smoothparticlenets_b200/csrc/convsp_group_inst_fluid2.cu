// Group signatures of the fluid step for ndim 2 (see convsp_group_inst_fluid3.cu).
#include "convsp_group.cuh"

namespace spnb {
namespace grp {
const SigEntry kSigsFluid2[] = {
    sig_entry<Sig<2, 6, 0x112121u, 0xEEFEFEu, 0x10077Bu, 0x1Eu, 0x0Au>>(),
    sig_entry<Sig<2, 2, 0x12u, 0x10u, 0x77u, 0x3u, 0x3u>>(),
    sig_entry<Sig<2, 2, 0x12u, 0xE0u, 0xBBu, 0x0u, 0x1u>>(),
    sig_entry<Sig<2, 1, 0x2u, 0x0u, 0x1u, 0x0u, 0x1u>>(),
};
const int kNumSigsFluid2 = sizeof(kSigsFluid2) / sizeof(kSigsFluid2[0]);
}  // namespace grp
}  // namespace spnb
