"""
Helper functions of the SwiFTly API: cover / mask generation, synthetic data and
error metrics (host side), and the per-task glue that says which core primitive
runs on which axis in which order.

Mirrors the public names of the reference's ``api_helper.py`` so that driver
scripts keep working: ``make_facet``, ``make_subgrid``, ``check_facet``,
``check_subgrid``, ``check_residual`` (:15-70), ``sum_and_finish_subgrid`` (:73-112),
``prepare_and_split_subgrid`` (:115-139), ``accumulate_column`` (:142-152),
``accumulate_facet`` (:155-179), ``finish_facet`` (:182-197), ``extract_column``
(:200-210), ``make_full_cover_config`` (:213-240), ``make_mask_from_slice`` (:243-253).

The task bodies take any object with the eight-primitive core interface; with
``SwiftlyCoreB200`` they run on the GPU for numpy arrays and CUDA tensors alike.
``SwiftlyForward`` does not call them -- it uses the fused kernels -- but they are
the unfused definition of what those kernels compute and the tests compare both.
"""

import numpy

from .fourier_algorithm import make_facet_from_sources, make_subgrid_from_sources


def _to_host(a):
    """numpy view of a result (CUDA tensors are copied to the host)."""
    if hasattr(a, "detach") and hasattr(a, "cpu"):
        return a.detach().cpu().numpy()
    return numpy.asarray(a)


# ------------------------------------------------------------------ synthetic data / checks
def make_subgrid(image_size, sg_config, sources):
    """Ground-truth subgrid for a source list (direct DFT)."""
    return make_subgrid_from_sources(
        sources, image_size, sg_config.size, [sg_config.off0, sg_config.off1],
        [sg_config.mask0, sg_config.mask1],
    )


def make_facet(image_size, facet_config, sources):
    """Facet data for a source list."""
    return make_facet_from_sources(
        sources, image_size, facet_config.size, [facet_config.off0, facet_config.off1],
        [facet_config.mask0, facet_config.mask1],
    )


def make_facet_device(image_size, facet_config, sources, device, out=None):
    """:func:`make_facet` painted directly on the GPU (no host array, no H2D copy).

    Same pixel rule as ``make_facet_from_sources`` (fourier_algorithm.py:218-264): a source at
    image coordinate ``x`` lands on pixel ``(x - (off - size // 2)) mod N`` if that is inside
    the facet; masks multiply.  ``out``: optional ``(size, size)`` complex128 device tensor.
    """
    import torch  # pylint: disable=import-outside-toplevel

    size = facet_config.size
    if out is None:
        out = torch.zeros((size, size), dtype=torch.complex128, device=device)
    else:
        out.zero_()
    masks = [facet_config.mask0, facet_config.mask1]
    corner = (facet_config.off0 - size // 2, facet_config.off1 - size // 2)
    rows, cols, vals = [], [], []
    for intensity, x0, x1 in sources:
        p0 = (int(x0) - corner[0]) % image_size
        p1 = (int(x1) - corner[1]) % image_size
        if p0 >= size or p1 >= size:
            continue
        w = 1.0
        if masks[0] is not None:
            w *= float(masks[0][p0])
        if masks[1] is not None:
            w *= float(masks[1][p1])
        rows.append(p0)
        cols.append(p1)
        vals.append(complex(intensity) * w)
    if rows:
        idx = (torch.tensor(rows, device=out.device), torch.tensor(cols, device=out.device))
        out.index_put_(idx, torch.tensor(vals, dtype=torch.complex128, device=out.device),
                       accumulate=True)
    return out


def _rms(diff):
    return numpy.sqrt(numpy.average(numpy.abs(diff) ** 2))


def check_facet(image_size, facet_config, approx_facet, sources):
    """RMS error of a facet against the source list."""
    truth = make_facet(image_size, facet_config, sources)
    return _rms(truth - _to_host(approx_facet))


def check_residual(residual_facet):
    """RMS of a residual image."""
    return _rms(_to_host(residual_facet))


def check_subgrid(image_size, sg_config, approx_subgrid, sources):
    """RMS error of a subgrid against the direct DFT of the source list."""
    approx = _to_host(approx_subgrid)
    truth = make_subgrid_from_sources(
        sources, image_size, approx.shape[0], [sg_config.off0, sg_config.off1],
        [sg_config.mask0, sg_config.mask1],
    )
    return _rms(truth - approx)


# ------------------------------------------------------------------ covers and masks
def make_mask_from_slice(slice_list, mask_size):
    """0/1 mask of length ``mask_size`` that is one on the given slices."""
    mask = numpy.zeros((mask_size,))
    for sl in slice_list:
        mask[sl] = 1
    return mask


def make_full_cover_config(N, chunk_size, class_name):
    """Configs of a full cover of the ``N x N`` plane with ``chunk_size`` chunks.

    Chunks sit at offsets ``chunk_size * k``; each owns the pixels up to half way
    to its neighbours (mask borders), off0 varies slowest.
    """
    offsets = chunk_size * numpy.arange(int(numpy.ceil(N / chunk_size)))
    mids = (offsets + numpy.append(offsets[1:], N + offsets[0])) // 2
    spans = []
    for i, off in enumerate(offsets):
        lo = (mids[i - 1] - off + chunk_size // 2) % N
        hi = mids[i] - off + chunk_size // 2
        spans.append([[slice(lo, hi)], chunk_size])
    return [
        class_name(off0, off1, chunk_size, spans[i0], spans[i1])
        for i0, off0 in enumerate(offsets)
        for i1, off1 in enumerate(offsets)
    ]


# ------------------------------------------------------------------ task bodies (unfused definition)
def _scale_rows(a, mask):
    """``a *= mask[:, None]`` for numpy arrays and CUDA tensors."""
    if mask is None:
        return a
    if hasattr(a, "detach"):
        import torch  # pylint: disable=import-outside-toplevel

        a *= torch.as_tensor(numpy.asarray(mask, dtype=float), device=a.device)[:, None]
    else:
        a *= numpy.asarray(mask)[:, numpy.newaxis]
    return a


def _scale_cols(a, mask):
    if mask is None:
        return a
    if hasattr(a, "detach"):
        import torch  # pylint: disable=import-outside-toplevel

        a *= torch.as_tensor(numpy.asarray(mask, dtype=float), device=a.device)[None, :]
    else:
        a *= numpy.asarray(mask)[numpy.newaxis, :]
    return a


def extract_column(distriFFT, BF_F, subgrid_off0, facet_off1):
    """Contribution window of a subgrid column along axis 0, prepared along axis 1."""
    rows = distriFFT.extract_from_facet(BF_F, subgrid_off0, axis=0)
    return distriFFT.prepare_facet(rows, facet_off1, axis=1)


def sum_and_finish_subgrid(distributedFFT, NMBF_NMBF_tasks, facets_config_list, subgrid_config):
    """Sum the facet contributions of one subgrid and finish it.

    Facets with equal ``off1`` are first combined along axis 0, the combined
    columns then along axis 1 (deterministic order: sorted ``off1``).
    """
    total = None
    for off1 in sorted({cfg.off1 for cfg in facets_config_list}):
        column = None
        for cfg, contrib in zip(facets_config_list, NMBF_NMBF_tasks):
            if cfg.off1 == off1:
                column = distributedFFT.add_to_subgrid(contrib, cfg.off0, axis=0, out=column)
        total = distributedFFT.add_to_subgrid(column, off1, axis=1, out=total)
    result = distributedFFT.finish_subgrid(
        total, [subgrid_config.off0, subgrid_config.off1], subgrid_config.size
    )
    _scale_rows(result, subgrid_config.mask0)
    _scale_cols(result, subgrid_config.mask1)
    return result


def prepare_and_split_subgrid(distributedFFT, subgrid, subgrid_offs, facets_config_list):
    """Prepare a subgrid and extract its contribution to every facet."""
    prepared = distributedFFT.prepare_subgrid(subgrid, tuple(subgrid_offs))
    by_off0 = {}
    pieces = []
    for cfg in facets_config_list:
        if cfg.off0 not in by_off0:
            by_off0[cfg.off0] = distributedFFT.extract_from_subgrid(prepared, cfg.off0, axis=0)
        pieces.append(distributedFFT.extract_from_subgrid(by_off0[cfg.off0], cfg.off1, axis=1))
    return pieces


def accumulate_column(distributedFFT, NAF_NAF, NAF_MNAF, subgrid_off1):
    """Add one subgrid's contribution to a facet's column accumulator (in place)."""
    return distributedFFT.add_to_facet(NAF_NAF, subgrid_off1, axis=1, out=NAF_MNAF)


def accumulate_facet(distributedFFT, NAF_MNAF, MNAF_BMNAF, facet_config, sg_off0):
    """Fold a finished subgrid column into a facet accumulator (in place)."""
    part = distributedFFT.finish_facet(NAF_MNAF, facet_config.off1, facet_config.size, axis=1)
    _scale_cols(part, facet_config.mask1)
    return distributedFFT.add_to_facet(part, sg_off0, axis=0, out=MNAF_BMNAF)


def finish_facet(distriFFT, MNAF_BMNAF, facet_config):
    """Finish a facet accumulator along axis 0 (zeros if nothing was accumulated)."""
    if MNAF_BMNAF is None:
        return numpy.zeros((facet_config.size, facet_config.size), dtype=complex)
    facet = distriFFT.finish_facet(MNAF_BMNAF, facet_config.off0, facet_config.size, axis=0)
    _scale_rows(facet, facet_config.mask0)
    return facet
