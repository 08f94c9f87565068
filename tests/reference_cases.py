"""
The reference's own known-answer tests for the eight primitives (reference
tests/test_core.py:93-428), re-expressed over a core factory so that they can be run
against the CUDA backend (tests/test_gpu_reference_cases.py) and -- thinned out -- against
the host-emulated kernels.  Assertions and tolerances are the reference's
(decimal=15 / 8 / 13 / 11); ground truth is analytic (direct DFT of point sources).
"""

import itertools

import numpy

from ska_sdp_distributed_fourier_transform_b200 import (
    make_facet_from_sources,
    make_subgrid_from_sources,
)

TEST_PARAMS = {"W": 13.5625, "N": 1024, "yB_size": 416, "yN_size": 512, "xA_size": 228,
               "xM_size": 256}


def facet_to_subgrid_basic(dft, xA_size, yB_size, thin=1):
    """tests/test_core.py:93-136: single pixel at the image centre => constant subgrid."""
    N = TEST_PARAMS["N"]
    Nx, Ny = dft.subgrid_off_step, dft.facet_off_step
    combos = list(itertools.product([0, 1, 0.1], numpy.arange(-5 * Ny, 5 * Ny // 2, Ny)))
    for val, facet_off in combos[::thin]:
        facet = numpy.zeros(yB_size)
        facet[yB_size // 2 - facet_off] = val
        prepped = dft.prepare_facet(facet, int(facet_off), axis=0)
        for sg_off in numpy.arange(0, 10 * Nx, Nx)[::thin]:
            contrib = dft.extract_from_facet(prepped, int(sg_off), axis=0)
            acc = dft.add_to_subgrid(contrib, int(facet_off), axis=0)
            subgrid = dft.finish_subgrid(acc, int(sg_off), xA_size)
            numpy.testing.assert_array_almost_equal(subgrid, val / N, decimal=15)


def facet_to_subgrid_dft_1d(dft, xA_size, yB_size, thin=1):
    """tests/test_core.py:139-199: against the direct DFT, decimal=8."""
    N = TEST_PARAMS["N"]
    Nx, Ny = dft.subgrid_off_step, dft.facet_off_step
    source_lists = [
        [(1, 0)], [(2, 1)], [(1, -3)], [(-0.1, 5)],
        [(1 / 8, 20), (2 / 8, 5), (3 / 8, -4)],
        [(1, -yB_size)], [(1, yB_size)],
        [(1 / 16, i) for i in range(-10, 10)],
    ]
    combos = list(itertools.product(source_lists, numpy.arange(-100 * Ny, 100 * Ny, 10 * Ny)))
    for sources, facet_off in combos[::thin]:
        facet_off = int(facet_off)
        min_x = -(yB_size - 1) // 2 + facet_off
        max_x = min_x + yB_size - 1
        sources = [(i, min(max(x, min_x), max_x)) for i, x in sources]
        facet = make_facet_from_sources(sources, N, yB_size, [facet_off])
        assert numpy.sum(facet) == sum(src[0] for src in sources)
        prepped = dft.prepare_facet(facet, facet_off, axis=0)
        for sg_off in [0, Nx, -Nx, N]:
            contrib = dft.extract_from_facet(prepped, sg_off, axis=0)
            acc = dft.add_to_subgrid(contrib, facet_off, axis=0)
            subgrid = dft.finish_subgrid(acc, sg_off, xA_size)
            expected = make_subgrid_from_sources(sources, N, xA_size, [sg_off])
            numpy.testing.assert_array_almost_equal(subgrid, expected, decimal=8,
                                                    err_msg=str(sources))


def facet_to_subgrid_dft_2d(dft):
    """tests/test_core.py:202-254."""
    N, xA_size, yB_size = TEST_PARAMS["N"], TEST_PARAMS["xA_size"], TEST_PARAMS["yB_size"]
    Nx, Ny = dft.subgrid_off_step, dft.facet_off_step
    for sources, facet_offs in itertools.product(
        [[(1, 1, 2)], [(1 / 8, 20, 4), (2 / 8, 2, 5), (3 / 8, -5, -4)]],
        [[0, 0], [Ny, Ny], [-Ny, Ny], [0, -Ny]],
    ):
        facet = make_facet_from_sources(sources, N, yB_size, facet_offs)
        assert numpy.sum(facet) == sum(src[0] for src in sources)
        prepped0 = dft.prepare_facet(facet, facet_offs[0], axis=0)
        prepped = dft.prepare_facet(prepped0, facet_offs[1], axis=1)
        for sg_offs in [[0, 0], [0, Nx], [Nx, 0], [-Nx, -Nx]]:
            c0 = dft.extract_from_facet(prepped, sg_offs[0], axis=0)
            c = dft.extract_from_facet(c0, sg_offs[1], axis=1)
            a0 = dft.add_to_subgrid(c, facet_offs[0], axis=0)
            a = dft.add_to_subgrid(a0, facet_offs[1], axis=1)
            subgrid = dft.finish_subgrid(a, sg_offs, xA_size)
            expected = make_subgrid_from_sources(sources, N, xA_size, sg_offs)
            numpy.testing.assert_array_almost_equal(subgrid, expected, decimal=8)


def subgrid_to_facet_basic(dft, xA_size, yB_size, thin=1):
    """tests/test_core.py:257-293: constant subgrid => pixel value at the image centre."""
    Nx, Ny = dft.subgrid_off_step, dft.facet_off_step
    sg_offs = Nx * numpy.arange(-9, 8)
    facet_offs = Ny * numpy.arange(-9, 8)
    combos = list(itertools.product([0, 1, 0.1], sg_offs))
    for val, sg_off in combos[::thin]:
        prepped = dft.prepare_subgrid((val / xA_size) * numpy.ones(xA_size), int(sg_off))
        for facet_off in facet_offs[::thin]:
            extracted = dft.extract_from_subgrid(prepped, int(facet_off), axis=0)
            accumulated = dft.add_to_facet(extracted, int(sg_off), axis=0)
            facet = dft.finish_facet(accumulated, int(facet_off), yB_size, axis=0)
            numpy.testing.assert_array_almost_equal(
                facet[yB_size // 2 - facet_off], val, decimal=13)


def subgrid_to_facet_dft(dft, xA_size, yB_size, thin=1):
    """tests/test_core.py:296-361."""
    N = TEST_PARAMS["N"]
    Nx, Ny = dft.subgrid_off_step, dft.facet_off_step
    source_lists = [[(1, 0)], [(2, 1)], [(1, -3)], [(-0.1, 5)]]
    sg_offs = Nx * numpy.arange(-9, 8)
    facet_offs = Ny * numpy.arange(-9, 8)
    combos = list(itertools.product(source_lists, sg_offs))
    for sources, sg_off in combos[::thin]:
        sg_off = int(sg_off)
        subgrid = make_subgrid_from_sources(sources, N, xA_size, [sg_off]) / xA_size * N
        prepped = dft.prepare_subgrid(subgrid, sg_off)
        for facet_off in facet_offs[::thin]:
            facet_off = int(facet_off)
            extracted = dft.extract_from_subgrid(prepped, facet_off, axis=0)
            accumulated = dft.add_to_facet(extracted, sg_off, axis=0)
            facet = dft.finish_facet(accumulated, facet_off, yB_size, axis=0)
            expected = make_facet_from_sources(sources, N, yB_size, [facet_off])
            numpy.testing.assert_array_almost_equal(
                facet[expected != 0], expected[expected != 0], decimal=11)
            if sources[0][0] > 0:
                numpy.testing.assert_array_less(facet[expected == 0].real,
                                                numpy.max(expected.real))
            else:
                numpy.testing.assert_array_less(-facet[expected == 0].real,
                                                numpy.max(-expected.real))


def subgrid_to_facet_dft_2d(dft):
    """tests/test_core.py:364-428."""
    N, xA_size, yB_size = TEST_PARAMS["N"], TEST_PARAMS["xA_size"], TEST_PARAMS["yB_size"]
    Nx, Ny = dft.subgrid_off_step, dft.facet_off_step
    for sources, sg_off in itertools.product(
        [[(1, 0, 0)], [(1, 20, 4)], [(3, -5, 4)]],
        [[0, 0], [0, Nx], [Nx, 0], [-Nx, -Nx]],
    ):
        subgrid = (make_subgrid_from_sources(sources, N, xA_size, sg_off)
                   / xA_size / xA_size * N * N)
        prepped = dft.prepare_subgrid(subgrid, tuple(sg_off))
        for facet_off in [[0, 0], [Ny, Ny], [-Ny, Ny], [0, -Ny]]:
            e0 = dft.extract_from_subgrid(prepped, facet_off[0], axis=0)
            e1 = dft.extract_from_subgrid(e0, facet_off[1], axis=1)
            a0 = dft.add_to_facet(e1, sg_off[0], axis=0)
            a1 = dft.add_to_facet(a0, sg_off[1], axis=1)
            f0 = dft.finish_facet(a1, facet_off[0], yB_size, axis=0)
            f1 = dft.finish_facet(f0, facet_off[1], yB_size, axis=1)
            expected = make_facet_from_sources(sources, N, yB_size, facet_off)
            numpy.testing.assert_array_almost_equal(
                f1[expected != 0], expected[expected != 0], decimal=11)
            numpy.testing.assert_array_less(f1[expected == 0].real, numpy.max(expected.real))
