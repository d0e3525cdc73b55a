"""Argument validation with the reference's messages and exception types
(python/SmoothParticleNets/error_checking.py:1-46): every failure is a ValueError."""
import numbers  # noqa: F401  (referenced from condition strings)

import numpy as np
import torch  # noqa: F401  (referenced from condition strings)


def check_nans(v, name):
    if (v != v).any():
        raise ValueError("Found NaNs in %s" % name)


def check_conditions(v, name, *conditions):
    """Each condition is a template such as "%s > 0" evaluated on the value."""
    for condition in conditions:
        if not eval(condition % "v"):
            raise ValueError(("%s must meet the following condition: " + condition) % (name, name))
    return v


def make_list(l, length, name, *conditions):
    try:
        l = list(l)
    except TypeError:
        l = [l] * length
    if len(l) != length:
        raise ValueError("%s must be a list of length %d." % (name, length))
    return [check_conditions(x, name, *conditions) for x in l]


def check_tensor_dims(t, name, dims):
    s = t.size()
    if len(s) != len(dims):
        raise ValueError("%s must be a %d-dimensional tensor." % (name, len(dims)))
    for i, want in enumerate(dims):
        if want >= 0 and s[i] != want:
            raise ValueError("The %dth dimension of %s must have size %d, not %d." % (i, name, want, s[i]))


def list2tensor(l):
    return torch.from_numpy(np.array(l, dtype=np.float32))
