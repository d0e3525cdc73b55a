"""smoothparticlenets_b200 -- B200 (sm_100a) implementation of the SmoothParticleNets
particle-interaction hot path, behind the reference's own nn.Module API.

    import smoothparticlenets_b200 as spn
    coll = spn.ParticleCollision(3, 0.1).cuda()
    conv = spn.ConvSP(4, 8, 3, 3, 0.05, 0.1, kernel_fn='spiky').cuda()

Exports the same names as ``import SmoothParticleNets as spn``: ConvSP, ConvSDF, ParticleCollision,
ReorderData, ParticleProjection, ImageProjection, KERNEL_NAMES, KERNEL_FN (reference python/SmoothParticleNets/
__init__.py:5-10).  Native code: libspnb.so (C ABI in include/spnb.h), built by
``python -m smoothparticlenets_b200.build``.  There is no CPU fallback.
"""
from .kernels import KERNEL_NAMES, KERNEL_FN, DKERNEL_FN, KERNELS, DKERNELS  # noqa: F401
from .convsp import ConvSP  # noqa: F401
from .convsp_group import ConvSPGroup  # noqa: F401
from .particlecollision import ParticleCollision, ReorderData, tile_lists_of, sym_flag_of, grid_bounds  # noqa: F401
from .convsdf import ConvSDF  # noqa: F401
from .particleprojection import ParticleProjection  # noqa: F401
from .imageprojection import ImageProjection  # noqa: F401
from .pbf import (pbf_stage1, pbf_stage2, pbf_stage3, pbf_integrate, pbf_velocity,  # noqa: F401
                  pbf_viscosity, fanout)
from . import error_checking  # noqa: F401

__all__ = ["ConvSP", "ConvSPGroup", "ConvSDF", "ParticleProjection", "ImageProjection", "ParticleCollision", "ReorderData", "KERNEL_NAMES", "KERNEL_FN",
           "DKERNEL_FN", "KERNELS", "DKERNELS", "tile_lists_of", "sym_flag_of", "grid_bounds", "pbf_stage1", "pbf_stage2", "pbf_stage3", "pbf_integrate",
           "pbf_velocity", "pbf_viscosity", "fanout"]
