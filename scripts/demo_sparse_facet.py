#!/usr/bin/env python
"""
Sparse facet cover demo on the GPU: only the facets that intersect a circular field of view
are transformed (the capability of the reference's ``scripts/demo_sparse_facet.py``); every
subgrid is checked against the direct DFT of the sources that fall inside the covered area.

    python scripts/demo_sparse_facet.py --swift_config "4k[1]-n2k-512" --fov_facets 2.12
"""

import argparse
import logging
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ska_sdp_distributed_fourier_transform_b200 import (  # noqa: E402
    SWIFT_CONFIGS, FacetConfig, SwiftlyConfig, SwiftlyForward, check_subgrid, make_facet,
    make_full_subgrid_cover)

log = logging.getLogger("fourier-logger")


def disc_cover_offsets(N, facet_size, fov_pixels):
    """Mid-point offsets of the facets of a centred row/column layout that touch a disc of
    diameter ``fov_pixels`` (rows of facets centred on the image centre, each row holding as
    many facets as the chord of the disc at that row needs)."""
    n_rows = int(numpy.ceil(fov_pixels / facet_size))
    row_offs = (numpy.arange(n_rows) - (n_rows - 1) / 2) * facet_size
    offsets = []
    for off1 in row_offs:
        inner = max(abs(off1) - facet_size / 2, 0.0)
        chord = 2 * numpy.sqrt(max((fov_pixels / 2) ** 2 - inner**2, 0.0))
        n = max(1, int(numpy.ceil(chord / facet_size)))
        for off0 in (numpy.arange(n) - (n - 1) / 2) * facet_size:
            offsets.append((int(round(off0)) % N, int(round(off1)) % N))
    return offsets


def demo(params, fov_facets, n_sources, queue_size):
    cfg = SwiftlyConfig(**params)
    N, yB = cfg.image_size, cfg.max_facet_size
    step = cfg.facet_off_step
    offsets = disc_cover_offsets(N, yB, fov_facets * yB)
    for off0, off1 in offsets:
        if off0 % step or off1 % step:
            raise ValueError("facet offsets must be multiples of facet_off_step")
    facet_cfgs = [FacetConfig(o0, o1, yB) for o0, o1 in offsets]
    log.info("%d facets cover the field of view (full cover would need %d)",
             len(facet_cfgs), int(numpy.ceil(N / yB)) ** 2)
    # sources inside the first facet so that the sparse cover holds all the flux
    o0, o1 = offsets[0]
    sources = [(1, (o0 + i + 1 + N // 2) % N - N // 2, (o1 + i + N // 2) % N - N // 2)
               for i in range(n_sources)]
    fwd = SwiftlyForward(cfg, [(fc, make_facet(N, fc, sources)) for fc in facet_cfgs],
                         1, queue_size)
    worst = 0.0
    for sg in make_full_subgrid_cover(cfg):
        worst = max(worst, check_subgrid(N, sg, fwd.get_subgrid_task(sg).tensor, sources))
    log.info("max subgrid RMS error over the full grid: %e", worst)
    return worst


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--swift_config", default="4k[1]-n2k-512")
    ap.add_argument("--fov_facets", type=float, default=2.12,
                    help="field-of-view diameter in units of the facet size")
    ap.add_argument("--source_number", type=int, default=10)
    ap.add_argument("--queue_size", type=int, default=20)
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    demo(SWIFT_CONFIGS[args.swift_config], args.fov_facets, args.source_number, args.queue_size)


if __name__ == "__main__":
    main()
