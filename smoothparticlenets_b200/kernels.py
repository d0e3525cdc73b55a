"""SPH kernel table.

Mirrors python/SmoothParticleNets/kernels.py of the reference: the integer id of a kernel is its
position in the alphabetically sorted name list (kernels.py:123) and is what the native layer
receives as ``kernel_fn``.  ``KERNEL_FN`` gives double-precision Python callables ``w(d, H)`` like
the reference's table of lambdas (kernels.py:126-131); ``DKERNEL_FN`` the derivatives dW/dd.

The device implementation of the same table is ``sph_eval`` in csrc/spnb_common.cuh.
"""
import math

_PI = math.pi

KERNEL_FN = {
    "default": lambda d, H: (315.0 / (64.0 * _PI * H ** 9)) * (H * H - d * d) ** 3,
    "ddefault": lambda d, H: (-945.0 / (32.0 * _PI * H ** 9)) * (H * H - d * d) ** 2 * d,
    "ddefault2": lambda d, H: (-945.0 / (32.0 * _PI * H ** 9)) * (H ** 4 - 6 * H * H * d * d + 5 * d ** 4),
    "pressure": lambda d, H: (15.0 / (_PI * H ** 6)) * (H - d) ** 3,
    "dpressure": lambda d, H: (-45.0 / (_PI * H ** 6)) * (H - d) ** 2,
    "dpressure2": lambda d, H: (90.0 / (_PI * H ** 6)) * (H - d),
    "indirect": lambda d, H: H - d,
    "constant": lambda d, H: 1.0,
    "spiky": lambda d, H: 15.0 / (_PI * H ** 3) * (1.0 - d / H) ** 2,
    "dspiky": lambda d, H: -15.0 / (_PI * H ** 3) * 2.0 * (1.0 - d / H) / H,
    "cohesion": lambda d, H: -6.0 * (d / H) ** 3 + 7 * (d / H) ** 2 - 1,
    "sigmoid": lambda d, H: 1.0 / (1.0 + math.exp((d - 0.2 * H) * 20.0 / H)),
}

DKERNEL_FN = {
    "default": KERNEL_FN["ddefault"],
    "ddefault": KERNEL_FN["ddefault2"],
    "ddefault2": lambda d, H: (-945.0 / (32.0 * _PI * H ** 9)) * (20 * d ** 3 - 12 * H * H * d),
    "pressure": KERNEL_FN["dpressure"],
    "dpressure": KERNEL_FN["dpressure2"],
    "dpressure2": lambda d, H: -90.0 / (_PI * H ** 6),
    "indirect": lambda d, H: -1.0,
    "constant": lambda d, H: 0.0,
    "spiky": KERNEL_FN["dspiky"],
    "dspiky": lambda d, H: -15.0 / (_PI * H ** 3) * 2.0 * (-1.0 / H) / H,
    "cohesion": lambda d, H: 2.0 * d * (7.0 * H - 9.0 * d) / (H ** 3),
    "sigmoid": lambda d, H: (-20.0 * math.exp((d - 0.2 * H) * 20.0 / H) /
                             (H * (math.exp((d - 0.2 * H) * 20.0 / H) + 1.0) ** 2)),
}

KERNEL_NAMES = sorted(KERNEL_FN.keys())
# Kept for API compatibility with code that only looks at the keys (convsp.py:10).
KERNELS = {k: k for k in KERNEL_NAMES}
DKERNELS = {k: k for k in KERNEL_NAMES}
