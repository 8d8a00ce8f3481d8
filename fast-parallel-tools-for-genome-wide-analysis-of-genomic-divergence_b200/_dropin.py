"""Shared argument checking for the four drop-in functions.

The reference's Cython signatures type every array as ``np.ndarray[np.float64_t, ndim=1]`` /
``np.ndarray[np.int32_t, ndim=1]`` (statistics/fisher/fisher_cython_parallel.pyx:14,
statistics/css/css_cython_parallel.pyx:14) and then hand the raw ``.data`` pointer to C, so a wrong
dtype or rank raises ``ValueError`` at the call and a non-contiguous array would be read as garbage.
Here the first behaviour is kept and the second becomes a ``ValueError`` too.
"""
import numpy as np


def _arr(name, a, dtype, writable=False):
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch for '%s', expected '%s' but got '%s'" % (name, np.dtype(dtype).name, a.dtype.name))
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions for '%s' (expected 1, got %d)" % (name, a.ndim))
    if not a.flags.c_contiguous:
        raise ValueError("'%s' must be C-contiguous" % name)
    if writable and not a.flags.writeable:
        raise ValueError("'%s' must be writable" % name)
    return a.ctypes.data


def check_inputs(avals, bvals, apos, bpos, alen, blen):
    pa, pb = _arr("avals", avals, np.float64), _arr("bvals", bvals, np.float64)
    qa, qb = _arr("apos", apos, np.int32), _arr("bpos", bpos, np.int32)
    if alen > avals.size or alen > apos.size or blen > bvals.size or blen > bpos.size:
        raise ValueError("alen/blen exceed the array lengths")
    return pa, pb, qa, qb


def check_outputs(regend, wstep, **outs):
    need = regend // wstep if wstep > 0 else 0
    ptrs = []
    for name, a in outs.items():
        ptrs.append(_arr(name, a, np.float64, writable=True))
        if a.size < need:
            raise ValueError("'%s' holds %d entries but regend/wstep = %d windows may be written" % (name, a.size, need))
    return ptrs
