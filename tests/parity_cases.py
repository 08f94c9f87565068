"""
Shared parity checks: a core object exposing the eight SwiFTly primitives is
compared with the CPU oracle on seeded inputs.  Used by the GPU parity tests
(CUDA library) and, for kernel index algebra, by the host-emulated build.

Tolerance: ``max|a - ref| <= RTOL * max|ref|`` per array with RTOL = 1e-9
(north-star "rtol = 1e-9", taken in the max norm); the observed
differences are ~1e-15 (fp64 re-association only).
"""

import numpy

from oracle.swiftly_oracle import OracleCore

RTOL = 1e-9
TIGHT = 1e-12  # what we actually expect from fp64 kernels


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def close(a, ref, rtol=TIGHT, what=""):
    a = numpy.asarray(a)
    assert a.shape == ref.shape, f"{what}: shape {a.shape} vs {ref.shape}"
    scale = max(numpy.abs(ref).max(), 1e-300)
    err = numpy.abs(a - ref).max()
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def check_1d_chain(core, oracle, yB, xA, facet_off, sg_off, rng, rtol=TIGHT):
    facet = rand_c(rng, yB)
    prep = core.prepare_facet(facet, facet_off, axis=0)
    oprep = oracle.prepare_facet(facet, facet_off, axis=0)
    close(prep, oprep, rtol=rtol, what="prepare_facet")
    contrib = core.extract_from_facet(oprep, sg_off, axis=0)
    ocontrib = oracle.extract_from_facet(oprep, sg_off, axis=0)
    assert numpy.array_equal(contrib, ocontrib), "extract_from_facet must be exact"
    acc = core.add_to_subgrid(ocontrib, facet_off, axis=0)
    oacc = oracle.add_to_subgrid(ocontrib, facet_off, axis=0)
    close(acc, oacc, rtol=rtol, what="add_to_subgrid")
    sg = core.finish_subgrid(oacc, sg_off, xA)
    close(sg, oracle.finish_subgrid(oacc, sg_off, xA), rtol=rtol, what="finish_subgrid")
    subgrid = rand_c(rng, xA)
    psg = core.prepare_subgrid(subgrid, sg_off)
    opsg = oracle.prepare_subgrid(subgrid, sg_off)
    close(psg, opsg, rtol=rtol, what="prepare_subgrid")
    ext = core.extract_from_subgrid(opsg, facet_off, axis=0)
    oext = oracle.extract_from_subgrid(opsg, facet_off, axis=0)
    close(ext, oext, rtol=rtol, what="extract_from_subgrid")
    accf = core.add_to_facet(oext, sg_off, axis=0)
    oaccf = oracle.add_to_facet(oext, sg_off, axis=0)
    assert numpy.array_equal(accf, oaccf), "add_to_facet must be exact"
    fin = core.finish_facet(oaccf, facet_off, yB, axis=0)
    close(fin, oracle.finish_facet(oaccf, facet_off, yB, axis=0), rtol=rtol, what="finish_facet")


def check_2d_axis(core, oracle, yB, axis, other, facet_off, sg_off, rng, rtol=TIGHT):
    """Every primitive along ``axis`` of a 2-D array with ``other`` lines."""
    yN, xM, m = oracle.yN_size, oracle.xM_size, oracle.xM_yN_size

    def shp(n):
        s = [other, other]
        s[axis] = n
        return s

    facet = rand_c(rng, *shp(yB))
    oprep = oracle.prepare_facet(facet, facet_off, axis=axis)
    close(core.prepare_facet(facet, facet_off, axis=axis), oprep, rtol=rtol, what=f"prepare_facet ax{axis}")
    ocontrib = oracle.extract_from_facet(oprep, sg_off, axis=axis)
    assert numpy.array_equal(core.extract_from_facet(oprep, sg_off, axis=axis), ocontrib)
    acc0 = rand_c(rng, *shp(xM))
    oacc = oracle.add_to_subgrid(ocontrib, facet_off, axis=axis, out=acc0.copy())
    acc = core.add_to_subgrid(ocontrib, facet_off, axis=axis, out=acc0.copy())
    close(acc, oacc, rtol=rtol, what=f"add_to_subgrid ax{axis} (accumulate)")
    close(core.add_to_subgrid(ocontrib, facet_off, axis=axis),
          oracle.add_to_subgrid(ocontrib, facet_off, axis=axis), rtol=rtol, what="add_to_subgrid (fresh)")
    close(core.finish_facet(oprep, facet_off, yB, axis=axis),
          oracle.finish_facet(oprep, facet_off, yB, axis=axis), rtol=rtol, what=f"finish_facet ax{axis}")
    oext = oracle.extract_from_subgrid(oacc, facet_off, axis=axis)
    close(core.extract_from_subgrid(oacc, facet_off, axis=axis), oext,
          what=f"extract_from_subgrid ax{axis}")
    accf0 = rand_c(rng, *shp(yN))
    oaccf = oracle.add_to_facet(oext, sg_off, axis=axis, out=accf0.copy())
    accf = core.add_to_facet(oext, sg_off, axis=axis, out=accf0.copy())
    close(accf, oaccf, rtol=rtol, what=f"add_to_facet ax{axis}")
    assert m == ocontrib.shape[axis]


def check_2d_subgrid_ops(core, oracle, xA, sg_offs, rng):
    xM = oracle.xM_size
    summed = rand_c(rng, xM, xM)
    close(core.finish_subgrid(summed, list(sg_offs), xA),
          oracle.finish_subgrid(summed, list(sg_offs), xA), what="finish_subgrid 2d")
    sg = rand_c(rng, xA, xA)
    close(core.prepare_subgrid(sg, tuple(sg_offs)),
          oracle.prepare_subgrid(sg, tuple(sg_offs)), what="prepare_subgrid 2d")


def check_errors(core):
    import pytest

    yN, xM, m = core.yN_size, core.xM_size, core.xM_yN_size
    with pytest.raises(ValueError):
        core.prepare_facet(numpy.zeros(yN // 2), 0, axis=0, out=numpy.zeros(yN + 1, dtype=complex))
    with pytest.raises(ValueError):
        core.extract_from_facet(numpy.zeros(yN - 1, dtype=complex), 0, axis=0)
    with pytest.raises(ValueError):
        core.add_to_subgrid(numpy.zeros(m + 1, dtype=complex), 0, axis=0)
    with pytest.raises(ValueError):
        core.finish_subgrid(numpy.zeros((xM, xM), dtype=complex), 0, xM // 2)
    with pytest.raises(ValueError):
        core.prepare_subgrid(numpy.zeros((8, 8), dtype=complex), (0,))
    with pytest.raises(ValueError):
        core.prepare_facet(numpy.zeros((4, 4, 4)), 0, axis=0)
    with pytest.raises(ValueError):
        core.prepare_facet(numpy.zeros((yN // 2, 4)), 0, axis=2)


def make_pair(cls, W, N, xM, yN, **kw):
    return cls(W, N, xM, yN, **kw), OracleCore(W, N, xM, yN)
