#!/usr/bin/env python
"""
Round-trip demo on the GPU: facets -> subgrids -> facets for one or more named parameter
sets (same command line as the reference's ``scripts/demo_api.py``; no Dask, no cluster).

    python scripts/demo_api.py --swift_config "1k[1]-n512-256" --queue_size 20
"""

import argparse
import logging
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ska_sdp_distributed_fourier_transform_b200 import (  # noqa: E402
    SWIFT_CONFIGS, SwiftlyBackward, SwiftlyConfig, SwiftlyForward, check_facet, check_subgrid,
    make_facet, make_full_facet_cover, make_full_subgrid_cover)

log = logging.getLogger("fourier-logger")


def demo_api(params, queue_size, lru_forward, lru_backward, source_number, check_subgrids):
    """Forward + backward transform of random point sources; returns the facet RMS errors."""
    cfg = SwiftlyConfig(**params)
    N = cfg.image_size
    rng = numpy.random.default_rng(123456789)
    if source_number <= 1:
        sources = [(1, 1, 0)]
    else:
        sources = [(float(rng.random()), int(rng.integers(-N // 2, N // 2)),
                    int(rng.integers(-N // 2, N // 2))) for _ in range(source_number)]
    facet_cfgs = make_full_facet_cover(cfg)
    sg_cfgs = make_full_subgrid_cover(cfg)
    fwd = SwiftlyForward(cfg, [(fc, make_facet(N, fc, sources)) for fc in facet_cfgs],
                         lru_forward, queue_size)
    bwd = SwiftlyBackward(cfg, facet_cfgs, lru_backward, queue_size)
    t0 = time.perf_counter()
    for sg in sg_cfgs:
        task = fwd.get_subgrid_task(sg)
        if check_subgrids:
            log.info("subgrid %d/%d error %e", sg.off0, sg.off1,
                     check_subgrid(N, sg, task.tensor, sources))
        bwd.add_new_subgrid_task(sg, task)
    facets = bwd.finish()
    dt = time.perf_counter() - t0
    errors = [check_facet(N, fc, t.result(), sources) for fc, t in zip(facet_cfgs, facets)]
    for fc, err in zip(facet_cfgs, errors):
        log.info("error facet, off0/off1:%d/%d: %e", fc.off0, fc.off1, err)
    log.info("%d facets x %d subgrids forward+backward in %.3f s", len(facet_cfgs),
             len(sg_cfgs), dt)
    return errors


def main():
    ap = argparse.ArgumentParser(description=__doc__, fromfile_prefix_chars="@")
    ap.add_argument("--swift_config", default="1k[1]-n512-256",
                    help="comma separated SWIFT_CONFIGS keys")
    ap.add_argument("--queue_size", type=int, default=20)
    ap.add_argument("--lru_forward", type=int, default=1)
    ap.add_argument("--lru_backward", type=int, default=1)
    ap.add_argument("--source_number", type=int, default=10)
    ap.add_argument("--check_subgrid", action="store_true")
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    for key in args.swift_config.split(","):
        if key not in SWIFT_CONFIGS:
            raise KeyError(f"{key} is not a SWIFT_CONFIGS key")
        log.info("Running for swift-config: %s", key)
        errors = demo_api(SWIFT_CONFIGS[key], args.queue_size, args.lru_forward,
                          args.lru_backward, args.source_number, args.check_subgrid)
        log.info("max facet RMS error %e", max(errors))


if __name__ == "__main__":
    main()
