"""
PSWF-derived window tables of SwiFTly (host side, init only).

Mirrors ``SwiftlyCore._calculate_pswf / _calculate_Fb / _calculate_Fn``
(reference ``fourier_transform/core.py:104-150``): the same
``scipy.special.pro_ang1`` call on the same coordinates, so the tables handed to
the CUDA plan are bit-identical to the ones the reference's numpy backend uses.
This is O(yN) work done once per configuration; it is not on the hot path.
"""

import numpy
import scipy.special


def facet_coordinates(n):
    """``coordinates(n)`` of the reference (fourier_algorithm.py:125-138)."""
    return (numpy.arange(n) - n // 2) / n


def pswf_samples(W, yN_size):
    """Prolate-spheroidal wave function at padded-facet resolution (core.py:119-150)."""
    pswf = numpy.empty(yN_size, dtype=float)
    arg = 2 * facet_coordinates(yN_size)
    chunk = 500  # the reference evaluates in chunks of 500 (scipy work-around)
    for lo in range(1, yN_size, chunk):
        pswf[lo : lo + chunk] = scipy.special.pro_ang1(
            0, 0, numpy.pi * W / 2, arg[lo : lo + chunk]
        )[0]
    pswf[0] = 0.0  # the x = -1 sample is NaN
    return pswf


def window_tables(W, N, xM_size, yN_size):
    """Return ``(Fb, Fn)``: grid-correction (len yN-1) and gridding (len m) tables.

    ``Fb = 1 / pswf[1:]`` (core.py:104-108);
    ``Fn = pswf[(yN//2) % (N/xM) :: N/xM]`` (core.py:110-117).
    """
    pswf = pswf_samples(W, yN_size)
    step = N // xM_size
    Fb = 1.0 / pswf[1:]
    Fn = pswf[(yN_size // 2) % step :: step]
    return numpy.ascontiguousarray(Fb), numpy.ascontiguousarray(Fn)
