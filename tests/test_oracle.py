"""
Pin the CPU oracle: bit-for-bit against fixtures produced by the real
reference (tests/golden/make_golden.py) and against the reference's analytic
known-answer tests (direct DFT of point sources; reference tests/test_core.py).
"""

import itertools

import numpy
import pytest

from oracle.swiftly_oracle import (
    OracleCore,
    backward_reference_order,
    centred_fft,
    centred_ifft,
    cover_mask,
    extract_mid,
    facet_from_sources,
    forward_reference_order,
    full_cover_offsets,
    pad_mid,
    subgrid_from_sources,
)

TEST_PARAMS = dict(W=13.5625, N=1024, yB_size=416, yN_size=512, xA_size=228, xM_size=256)


def make_core(p=TEST_PARAMS):
    return OracleCore(p["W"], p["N"], p["xM_size"], p["yN_size"])


def same(a, b):
    """bit-for-bit equality (NaN-free data)."""
    return a.shape == b.shape and numpy.array_equal(a, b)


# ------------------------------------------------------------------ helpers
def test_pad_extract_known_answers():
    # reference tests/test_fourier_algorithm.py:26-242 behaviour
    assert list(pad_mid(numpy.arange(1, 6), 8, 0)) == [0, 0, 1, 2, 3, 4, 5, 0]
    assert list(pad_mid(numpy.array([1, 2, 3, 4]), 8, 0)) == [0, 0, 1, 2, 3, 4, 0, 0]
    assert list(pad_mid(numpy.array([1, 2, 3]), 6, 0)) == [0, 0, 1, 2, 3, 0]
    assert list(extract_mid(numpy.arange(8), 4, 0)) == [2, 3, 4, 5]
    assert list(extract_mid(numpy.arange(8), 3, 0)) == [3, 4, 5]
    assert list(extract_mid(numpy.arange(7), 3, 0)) == [2, 3, 4]
    assert list(extract_mid(numpy.arange(7), 4, 0)) == [1, 2, 3, 4]
    for n0, n in itertools.product(range(1, 9), range(1, 17)):
        if n0 <= n:
            x = numpy.arange(1, n0 + 1)
            assert same(extract_mid(pad_mid(x, n, 0), n0, 0), x)


def test_centred_fft_delta_and_roundtrip():
    n = 16
    d = numpy.zeros(n)
    d[n // 2] = 1
    numpy.testing.assert_allclose(centred_fft(d, 0), numpy.ones(n), atol=1e-15)
    numpy.testing.assert_allclose(centred_ifft(d, 0), numpy.ones(n) / n, atol=1e-15)
    rng = numpy.random.default_rng(1)
    x = rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))
    numpy.testing.assert_allclose(centred_ifft(centred_fft(x, 1), 1), x, atol=1e-14)


def test_params_and_derived():
    core = make_core()
    assert core.xM_yN_size == 128
    assert core.subgrid_off_step == 2 and core.facet_off_step == 4
    bad = dict(TEST_PARAMS, N=1050)
    with pytest.raises(ValueError):
        make_core(bad)


# ------------------------------------------------------------------ golden
def test_windows_match_reference(golden_1d, golden_windows):
    core = make_core()
    assert same(core._Fb, golden_1d["Fb"])
    assert same(core._Fn, golden_1d["Fn"])
    c2 = OracleCore(13.5625, 8192, 2048, 4096)
    assert same(c2._Fb[::16], golden_windows["cfg2_Fb_s"])
    assert same(c2._Fn[::16], golden_windows["cfg2_Fn_s"])
    numpy.testing.assert_allclose(
        [c2._Fb.sum(), c2._Fn.sum(), (c2._Fb**2).sum(), (c2._Fn**2).sum()],
        golden_windows["cfg2_sums"], rtol=1e-14)


def test_1d_primitives_bit_exact(golden_1d):
    g = golden_1d
    core = make_core()
    for idx, (yB, xA, f_off, s_off) in enumerate(g["cases"]):
        yB, xA, f_off, s_off = int(yB), int(xA), int(f_off), int(s_off)
        k = lambda name: g[f"c{idx}_{name}"]  # noqa: E731
        prep = core.prepare_facet(k("facet"), f_off, axis=0)
        assert same(prep, k("prep"))
        contrib = core.extract_from_facet(prep, s_off, axis=0)
        assert same(contrib, k("contrib"))
        acc = core.add_to_subgrid(contrib, f_off, axis=0)
        assert same(acc, k("acc"))
        assert same(core.finish_subgrid(acc, s_off, xA), k("sg"))
        psg = core.prepare_subgrid(k("subgrid"), s_off)
        assert same(psg, k("psg"))
        ext = core.extract_from_subgrid(psg, f_off, axis=0)
        assert same(ext, k("ext"))
        accf = core.add_to_facet(ext, s_off, axis=0)
        assert same(accf, k("accf"))
        assert same(core.finish_facet(accf, f_off, yB, axis=0), k("fin"))


def _core_2d(g):
    W, N, xM, yN, yB, xA = g["params"]
    return OracleCore(float(W), int(N), int(xM), int(yN)), int(yB), int(xA)


def test_2d_primitives_bit_exact(golden_2d):
    g = golden_2d
    core, yB, xA = _core_2d(g)
    f_off, s_off = (int(v) for v in g["prim_offs"])
    for axis in (0, 1):
        k = lambda name: g[f"ax{axis}_{name}"]  # noqa: E731
        prep = core.prepare_facet(k("facet"), f_off, axis=axis)
        assert same(prep, k("prep"))
        contrib = core.extract_from_facet(prep, s_off, axis=axis)
        assert same(contrib, k("contrib"))
        acc = core.add_to_subgrid(contrib, f_off, axis=axis, out=k("acc0").copy())
        assert same(acc, k("acc"))
        assert same(core.finish_facet(prep, f_off, yB - 1, axis=axis), k("fin"))
        ext = core.extract_from_subgrid(acc, f_off, axis=axis)
        assert same(ext, k("ext"))
        accf = core.add_to_facet(ext, s_off, axis=axis, out=k("accf0").copy())
        assert same(accf, k("accf"))
    Nx = core.subgrid_off_step
    assert same(core.finish_subgrid(g["fs_in"], [2 * Nx, -Nx], xA - 1), g["fs_out"])
    assert same(core.prepare_subgrid(g["ps_in"], (2 * Nx, -Nx)), g["ps_out"])


def test_2d_full_forward_backward_vs_reference(golden_2d):
    g = golden_2d
    core, yB, xA = _core_2d(g)
    facet_offs = [tuple(int(v) for v in o) for o in g["facet_offs"]]
    sg_offs = [tuple(int(v) for v in o) for o in g["sg_offs"]]
    sg_masks = list(zip(g["sg_mask0"], g["sg_mask1"]))
    keep = {}
    subgrids = forward_reference_order(
        core, list(g["facets"]), facet_offs, sg_offs, xA, subgrid_masks=sg_masks, keep=keep
    )
    ref = g["subgrids"]
    scale = numpy.abs(ref).max()
    # addition order over facet columns may differ (python set order in the
    # reference): allow rounding-level differences only
    for a, b in zip(subgrids, ref):
        assert numpy.abs(a - b).max() <= 1e-13 * scale
    assert same(keep["BF_F"][0], g["BF_F0"])
    assert same(numpy.array(keep["NMBF_BF"]), g["NMBF_BF_last"])
    f_masks = list(zip(g["facet_mask0"], g["facet_mask1"]))
    back = backward_reference_order(core, list(ref), sg_offs, facet_offs, yB, facet_masks=f_masks)
    bscale = numpy.abs(g["back_facets"]).max()
    for a, b in zip(back, g["back_facets"]):
        assert numpy.abs(a - b).max() <= 1e-13 * bscale


def test_cover_matches_reference(golden_2d):
    g = golden_2d
    _, N, _, _, yB, xA = (int(v) if i else v for i, v in enumerate(g["params"]))
    offs = full_cover_offsets(N, yB)
    assert [tuple(o) for o in g["facet_offs"]] == [(a, b) for a in offs for b in offs]
    nf = len(offs)
    for i in range(nf):
        assert same(cover_mask(N, yB, i), g["facet_mask0"][i * nf])
    soffs = full_cover_offsets(N, xA)
    ns = len(soffs)
    for i in range(ns):
        assert same(cover_mask(N, xA, i), g["sg_mask1"][i])


# ------------------------------------------------------------------ analytic KATs
@pytest.mark.parametrize("xA_size", [228, 227])
@pytest.mark.parametrize("yB_size", [416, 415])
def test_facet_to_subgrid_constant(xA_size, yB_size):
    """reference tests/test_core.py:93-136 (single pixel => constant subgrid val/N)."""
    N = TEST_PARAMS["N"]
    core = make_core()
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    for val, facet_off in itertools.product([0, 1, 0.1], numpy.arange(-5 * Ny, 5 * Ny // 2, Ny)):
        facet = numpy.zeros(yB_size)
        facet[yB_size // 2 - facet_off] = val
        prepped = core.prepare_facet(facet, facet_off, axis=0)
        for sg_off in numpy.arange(0, 10 * Nx, Nx):
            c = core.extract_from_facet(prepped, sg_off, axis=0)
            acc = core.add_to_subgrid(c, facet_off, axis=0)
            sg = core.finish_subgrid(acc, int(sg_off), xA_size)
            numpy.testing.assert_array_almost_equal(sg, val / N, decimal=15)


@pytest.mark.parametrize("xA_size", [228, 227])
@pytest.mark.parametrize("yB_size", [416, 415])
def test_facet_to_subgrid_dft_1d(xA_size, yB_size):
    """reference tests/test_core.py:139-199 (vs direct DFT, decimal=8)."""
    N = TEST_PARAMS["N"]
    core = make_core()
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    source_lists = [
        [(1, 0)], [(2, 1)], [(1, -3)], [(-0.1, 5)],
        [(1 / 8, 20), (2 / 8, 5), (3 / 8, -4)],
        [(1, -yB_size)], [(1, yB_size)],
        [(1 / 16, i) for i in range(-10, 10)],
    ]
    for sources, facet_off in itertools.product(source_lists, numpy.arange(-100 * Ny, 100 * Ny, 10 * Ny)):
        facet_off = int(facet_off)
        min_x = -(yB_size - 1) // 2 + facet_off
        max_x = min_x + yB_size - 1
        sources = [(i, min(max(x, min_x), max_x)) for i, x in sources]
        facet = facet_from_sources(sources, N, yB_size, [facet_off])
        assert numpy.sum(facet) == sum(src[0] for src in sources)
        prepped = core.prepare_facet(facet, facet_off, axis=0)
        for sg_off in [0, Nx, -Nx, N]:
            c = core.extract_from_facet(prepped, sg_off, axis=0)
            acc = core.add_to_subgrid(c, facet_off, axis=0)
            sg = core.finish_subgrid(acc, sg_off, xA_size)
            expected = subgrid_from_sources(sources, N, xA_size, [sg_off])
            numpy.testing.assert_array_almost_equal(sg, expected, decimal=8)


def test_subgrid_to_facet_dft_2d():
    """reference tests/test_core.py:364-428."""
    N, xA_size, yB_size = 1024, 228, 416
    core = make_core()
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    for sources, sg_off in itertools.product(
        [[(1, 0, 0)], [(1, 20, 4)], [(3, -5, 4)]],
        [[0, 0], [0, Nx], [Nx, 0], [-Nx, -Nx]],
    ):
        subgrid = subgrid_from_sources(sources, N, xA_size, sg_off) / xA_size / xA_size * N * N
        prepped = core.prepare_subgrid(subgrid, tuple(sg_off))
        for facet_off in [[0, 0], [Ny, Ny], [-Ny, Ny], [0, -Ny]]:
            e0 = core.extract_from_subgrid(prepped, facet_off[0], axis=0)
            e1 = core.extract_from_subgrid(e0, facet_off[1], axis=1)
            a0 = core.add_to_facet(e1, sg_off[0], axis=0)
            a1 = core.add_to_facet(a0, sg_off[1], axis=1)
            f0 = core.finish_facet(a1, facet_off[0], yB_size, axis=0)
            f1 = core.finish_facet(f0, facet_off[1], yB_size, axis=1)
            expected = facet_from_sources(sources, N, yB_size, facet_off)
            numpy.testing.assert_array_almost_equal(
                f1[expected != 0], expected[expected != 0], decimal=11)
            numpy.testing.assert_array_less(f1[expected == 0].real, numpy.max(expected.real))
