// Group signatures of the fluid step (examples/fluid_sim.py:367-397, 419-420), ndim 3.  Kernel ids: cohesion 0,
// constant 1, dspiky 7, spiky 0xB (kernels.py:123); layer 0 is the lowest nibble / bit; data source E = ones,
// F = the position tensor, else the index of the distinct data tensor; * = dis_norm:
//   A: spiky1(ones) dspikyD*(locs) dspiky1*(ones) cohesionD*(locs) cohesion1*(ones) constant1(ones)
//   B: dspikyD*(locs*pressure) dspiky1*(pressure)      V: spikyD(vel) spiky1(ones)      C: constantD(normals)
#include "convsp_group.cuh"

namespace spnb {
namespace grp {
const SigEntry kSigsFluid3[] = {
    sig_entry<Sig<3, 6, 0x113131u, 0xEEFEFEu, 0x10077Bu, 0x1Eu, 0x0Au>>(),
    sig_entry<Sig<3, 2, 0x13u, 0x10u, 0x77u, 0x3u, 0x3u>>(),
    sig_entry<Sig<3, 2, 0x13u, 0xE0u, 0xBBu, 0x0u, 0x1u>>(),
    sig_entry<Sig<3, 1, 0x3u, 0x0u, 0x1u, 0x0u, 0x1u>>(),
};
const int kNumSigsFluid3 = sizeof(kSigsFluid3) / sizeof(kSigsFluid3[0]);
}  // namespace grp
}  // namespace spnb
