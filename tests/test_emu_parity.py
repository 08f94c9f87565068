"""
Kernel index algebra without a GPU: the CUDA kernel bodies (csrc/*.cuh) are
compiled for the host with every CUDA thread running as a fibre
(tests/emu/emu_runtime.h) and driven through the same C ABI and the same Python
wrapper as the product.  TEST TOOLING ONLY -- the product never loads the
emulated library; the real parity gate is tests/test_gpu_*.py on the B200.
"""

import numpy
import pytest

from tests import parity_cases as pc


@pytest.fixture(scope="module")
def emu_core_cls():
    from tests.emu_support import emu_core_class

    return emu_core_class()


SMALL = dict(W=13.5625, N=256, xM=64, yN=128)     # m = 32
TESTP = dict(W=13.5625, N=1024, xM=256, yN=512)   # m = 128 (reference TEST_PARAMS)
MID = dict(W=13.5625, N=4096, xM=1024, yN=2048)   # m = 512: radix 16,16,8 / 16,16,4 / 16,16,2


@pytest.mark.parametrize("p,yB,xA", [(SMALL, 96, 52), (SMALL, 95, 51), (TESTP, 416, 228),
                                     (TESTP, 415, 227), (MID, 1500, 700)])
def test_emu_1d_chain(emu_core_cls, p, yB, xA):
    core, oracle = pc.make_pair(emu_core_cls, **p)
    rng = numpy.random.default_rng(7)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    for f_off, s_off in [(0, 0), (3 * Ny, 5 * Nx), (-7 * Ny, -2 * Nx), (p["N"], p["N"] + Nx)]:
        pc.check_1d_chain(core, oracle, yB, xA, f_off, s_off, rng)


@pytest.mark.parametrize("axis", [0, 1])
def test_emu_2d_axes(emu_core_cls, axis):
    core, oracle = pc.make_pair(emu_core_cls, **SMALL)
    rng = numpy.random.default_rng(8)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    pc.check_2d_axis(core, oracle, 95, axis, 19, 5 * Ny, -3 * Nx, rng)
    pc.check_2d_axis(core, oracle, 96, axis, 37, -Ny, 4 * Nx, rng)


def test_emu_2d_subgrid_ops(emu_core_cls):
    core, oracle = pc.make_pair(emu_core_cls, **SMALL)
    rng = numpy.random.default_rng(9)
    Nx = core.subgrid_off_step
    pc.check_2d_subgrid_ops(core, oracle, 51, (2 * Nx, -Nx), rng)
    pc.check_2d_subgrid_ops(core, oracle, 52, (0, 7 * Nx), rng)


def test_emu_split_path(emu_core_cls):
    """yN lines done as 2 x yN/2 (the path yN = 16384 takes on the GPU)."""
    for p, yB in ((SMALL, 63), (TESTP, 250), (TESTP, 416)):
        core, oracle = pc.make_pair(emu_core_cls, force_split=True, **p)
        rng = numpy.random.default_rng(10)
        Ny = core.facet_off_step
        for f_off in (0, 3 * Ny, -9 * Ny):
            facet = pc.rand_c(rng, 5, yB)
            pc.close(core.prepare_facet(facet, f_off, axis=1),
                     oracle.prepare_facet(facet, f_off, axis=1), what="split prepare_facet")
            acc = pc.rand_c(rng, 5, p["yN"])
            pc.close(core.finish_facet(acc, f_off, yB, axis=1),
                     oracle.finish_facet(acc, f_off, yB, axis=1), what="split finish_facet")


def test_emu_errors(emu_core_cls):
    core, _ = pc.make_pair(emu_core_cls, **SMALL)
    pc.check_errors(core)
    with pytest.raises(ValueError):
        emu_core_cls(13.5625, 1050, 256, 512)


def test_emu_golden_1d(emu_core_cls, golden_1d):
    g = golden_1d
    core, _ = pc.make_pair(emu_core_cls, **TESTP)
    for idx, (yB, xA, f_off, s_off) in enumerate(g["cases"][::5]):
        idx = idx * 5
        yB, xA, f_off, s_off = int(yB), int(xA), int(f_off), int(s_off)
        k = lambda name: g[f"c{idx}_{name}"]  # noqa: E731
        pc.close(core.prepare_facet(k("facet"), f_off, axis=0), k("prep"))
        assert numpy.array_equal(core.extract_from_facet(k("prep"), s_off, axis=0), k("contrib"))
        pc.close(core.add_to_subgrid(k("contrib"), f_off, axis=0), k("acc"))
        pc.close(core.finish_subgrid(k("acc"), s_off, xA), k("sg"))
        pc.close(core.prepare_subgrid(k("subgrid"), s_off), k("psg"))
        pc.close(core.extract_from_subgrid(k("psg"), f_off, axis=0), k("ext"))
        assert numpy.array_equal(core.add_to_facet(k("ext"), s_off, axis=0), k("accf"))
        pc.close(core.finish_facet(k("accf"), f_off, yB, axis=0), k("fin"))


@pytest.mark.parametrize("p,yB,cols", [(TESTP, 415, 37), (MID, 1500, 16), (TESTP, 416, 64)])
def test_emu_prepare_facet_axis0_two_pass(emu_core_cls, p, yB, cols):
    """Strided axis with >= 16 adjacent lines takes the two-pass (four-step) kernels."""
    core, oracle = pc.make_pair(emu_core_cls, **p)
    rng = numpy.random.default_rng(12)
    Ny = core.facet_off_step
    for f_off in (0, 5 * Ny, -11 * Ny):
        facet = pc.rand_c(rng, yB, cols)
        pc.close(core.prepare_facet(facet, f_off, axis=0),
                 oracle.prepare_facet(facet, f_off, axis=0), what="two-pass prepare_facet")


NONPOW2 = [
    dict(W=11.0, N=1536, xM=512, yN=768),    # catalogue 1536[1]-n768-512: yN = 3 * 256, m = 256
    dict(W=9.25, N=1792, xM=256, yN=1792),   # catalogue 1792[1]-n1792-256: yN = 7 * 256
    dict(W=11.0, N=1280, xM=320, yN=640),    # xM = 5 * 64, yN = 5 * 128, m = 160 = 5 * 32
    dict(W=11.0, N=2304, xM=576, yN=1152),   # factors of 9: 9 * 64, 9 * 128, m = 288 = 9 * 32
]


@pytest.mark.parametrize("p", NONPOW2)
def test_emu_non_power_of_two_lengths(emu_core_cls, p):
    """FFT lengths F * 2^k (F = 3, 5, 7, 9) go through the generic split-F kernel."""
    core, oracle = pc.make_pair(emu_core_cls, **p)
    rng = numpy.random.default_rng(13)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    yB = (p["yN"] * 11 // 16) | 1
    xA = (p["xM"] * 7 // 8) & ~1
    pc.check_1d_chain(core, oracle, yB, xA, 3 * Ny, -5 * Nx, rng)
    pc.check_1d_chain(core, oracle, yB - 1, xA - 1, -2 * Ny, 4 * Nx, rng)
    pc.check_2d_axis(core, oracle, yB, 0, 5, Ny, -Nx, rng)
    pc.check_2d_axis(core, oracle, yB, 1, 5, Ny, -Nx, rng)


def test_emu_prepare_facet_column_tiles(emu_core_cls):
    """Two-pass prepare_facet along the strided axis in several column tiles (the debug hook
    shrinks the tile to 32 columns: 3 full tiles and a ragged one for 110 columns)."""
    import ctypes

    core, oracle = pc.make_pair(emu_core_cls, W=13.5625, N=1024, xM=256, yN=512)
    core._lib.swiftly_b200_debug_sg_variant.argtypes = [ctypes.c_void_p, ctypes.c_int]
    core._lib.swiftly_b200_debug_sg_variant(core._plan, 10)
    rng = numpy.random.default_rng(10)
    facet = pc.rand_c(rng, 256, 110)
    got = core.prepare_facet(facet, 3 * core.facet_off_step, axis=0)
    pc.close(got, oracle.prepare_facet(facet, 3 * core.facet_off_step, axis=0), what="tiled")
