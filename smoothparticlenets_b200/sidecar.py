"""Side information that travels with a tensor ParticleCollision returned -- without living ON the tensor.

ParticleCollision knows more about its outputs than their float values say: whether the neighbour relation
is symmetric (device flag), the compact tile lists (csrc/tile_lists.cuh) and, for the reordered positions,
their float4 plane.  The consumers (ConvSP, ConvSPGroup) look that up here.  An entry is keyed on the
tensor OBJECT and validated against ``(data_ptr, _version, shape)``: an in-place edit of the tensor (masking
neighbour entries is a normal thing to do) bumps ``_version`` and silently drops the entry, so a consumer can
never compute from lists that no longer describe the tensor.  Nothing is stored in the tensor's ``__dict__``
(``torch.save`` / pickling of the tensor keeps working) and entries die with their tensor (weak references).
"""
import weakref

_REG = {}  # id(tensor) -> (weakref to the tensor, Sidecar)


class Sidecar(object):
    __slots__ = ("key", "sym_flag", "tiles", "builder", "pos4")

    def __init__(self, sym_flag=None, tiles=None, builder=None, pos4=None):
        self.key = None
        self.sym_flag = sym_flag  # device int32[1]: 0 = the neighbour relation is symmetric
        self.tiles = tiles        # uint8 buffer: tile lists
        self.builder = builder    # callable(tensor) -> tiles or None (lazy construction)
        self.pos4 = pos4          # float4 plane [B, N, 4] of a position tensor


def _key(t):
    return (t.data_ptr(), t._version, tuple(t.shape))


def attach(t, sc):
    i = id(t)
    sc.key = _key(t)
    _REG[i] = (weakref.ref(t, lambda _r, i=i: _REG.pop(i, None)), sc)
    return sc


def lookup(t):
    """The Sidecar of tensor object `t`, or None (never attached, tensor edited in place since, resized)."""
    e = _REG.get(id(t))
    if e is None:
        return None
    if e[0]() is not t:
        return None
    if e[1].key != _key(t):
        _REG.pop(id(t), None)
        return None
    return e[1]


def share(src, dst):
    """Let `dst` (an alias of `src`: same storage, same version counter) find the same Sidecar."""
    sc = lookup(src)
    if sc is not None and dst.data_ptr() == src.data_ptr() and tuple(dst.shape) == tuple(src.shape):
        i = id(dst)
        _REG[i] = (weakref.ref(dst, lambda _r, i=i: _REG.pop(i, None)), sc)
    return dst
