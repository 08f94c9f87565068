"""
Host-side data generators of the SwiFTly API (not on the hot path).

The reference exports ``make_facet_from_sources`` / ``make_subgrid_from_sources``
(``fourier_transform/fourier_algorithm.py:218-315``) so that drivers and tests can
build synthetic facets from point-source lists and check subgrids against a direct
DFT.  They are plain numpy utilities there and stay host utilities here: they
produce *inputs* and *ground truth*, they are not an implementation of the
transform.
"""

import numpy

__all__ = ["make_facet_from_sources", "make_subgrid_from_sources"]


def _stretch(vec, dims, axis):
    shape = [1] * dims
    shape[axis] = -1
    return numpy.reshape(numpy.asarray(vec), shape)


def make_facet_from_sources(sources, image_size, facet_size, facet_offsets, facet_masks=None):
    """Facet with the given point sources painted in (wrapping modulo the image).

    :param sources: list of ``(intensity, *coords)``, integer pixel coordinates
        relative to the image centre
    :param image_size: image size ``N`` (coordinates are modulo this)
    :param facet_size: size of the facet to generate
    :param facet_offsets: facet mid-point offsets; their number sets the dimension
    :param facet_masks: optional per-axis masks
    """
    dims = len(facet_offsets)
    facet = numpy.zeros([facet_size] * dims, dtype=complex)
    corner = numpy.asarray(facet_offsets, dtype=int) - facet_size // 2
    for intensity, *pos in sources:
        pix = (numpy.asarray(pos, dtype=int) - corner) % image_size
        if (pix < facet_size).all():
            facet[tuple(pix)] += intensity
    for axis, mask in enumerate(facet_masks or []):
        if mask is not None:
            facet *= _stretch(mask, dims, axis)
    return facet


def make_subgrid_from_sources(sources, image_size, subgrid_size, subgrid_offsets,
                              subgrid_masks=None):
    """Subgrid of the given point sources by direct Fourier transform.

    ``sum_s I_s / N^dims * exp(2 pi i (u . x_s) / N)`` for grid coordinates ``u``
    from ``off - size//2`` to ``off + (size + 1)//2 - 1`` on every axis.
    """
    dims = len(subgrid_offsets)
    coords = [numpy.arange(off - subgrid_size // 2, off + (subgrid_size + 1) // 2)
              for off in subgrid_offsets]
    subgrid = numpy.zeros([subgrid_size] * dims, dtype=complex)
    for intensity, *pos in sources:
        dot = numpy.zeros([subgrid_size] * dims)
        for axis in range(dims):
            dot = dot + _stretch(coords[axis] * pos[axis], dims, axis)
        subgrid += (intensity / image_size**dims) * numpy.exp(2j * numpy.pi / image_size * dot)
    for axis, mask in enumerate(subgrid_masks or []):
        if mask is not None:
            subgrid *= _stretch(mask, dims, axis)
    return subgrid
