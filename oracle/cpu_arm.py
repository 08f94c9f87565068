"""
TEST / BENCH INFRASTRUCTURE -- the CPU arm of ``bench.py``: the reference's algorithm
actually RUNNING on the host cores, timed end to end on a bounded slice of the workload.

Imported only by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg); never by
the product package.

What is timed ("one subgrid column of the transform"): for a workload with ``F`` facets and
``ns x ns`` subgrids the complete forward transform consists of ``ns`` subgrid columns.  The
slice does everything ONE column needs, with real arrays of the real shapes, in the
reference's own task order (``api.py:238-324``, ``api_helper.py:73-112,200-210``):

  phase A, one task per facet (like the reference's per-facet Dask tasks), all cores:
    * ``prepare_facet(axis 0)`` on a ``yB / ns``-column slab of the facet: stage 1 is shared
      by all ``ns`` columns, its lines (columns) are independent, so ``1 / ns`` of its lines
      is exactly this column's share of that stage;
    * ``extract_column`` = ``extract_from_facet(axis 0)`` + ``prepare_facet(axis 1)`` at full
      size -> ``NMBF_BF`` (``m x yN``), kept in shared memory;
  phase B, one task per subgrid of the column, all cores:
    * per facet ``extract_from_facet(axis 1)`` (the contribution), then
      ``sum_and_finish_subgrid``: ``add_to_subgrid(axis 0)`` summed per facet column,
      ``add_to_subgrid(axis 1)``, ``finish_subgrid``, masks.

``F x ns`` facet->subgrid contributions are produced; the rate is that count divided by the
wall time of phases A + B.  No unit-cost model, no task counting: the number is a timed run.

``kind``: ``"reference"`` when the unmodified reference package can be imported from
``/root/reference/src`` (build container; stubs for dask / ska_sdp_func as in SURVEY.md
Appendix A) -- then the reference's own ``SwiftlyCore`` and ``api_helper`` functions run;
``"port"`` otherwise (GPU box): the oracle restatement, which is pinned bit-for-bit to the
reference (``tests/test_oracle.py``), with the same numpy / pocketfft calls.
"""

import multiprocessing as mp
import os
import sys
import time
import types

import numpy

REF_SRC = "/root/reference/src"
_G = {}


def _import_reference():
    """The real reference (numpy backend) or None."""
    if not os.path.isdir(REF_SRC):
        return None

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _Unavailable:  # pylint: disable=too-few-public-methods
        def __init__(self, *a, **k):
            raise ImportError("stub: not installed")

    def _delayed(*a, **k):
        raise ImportError("stub: dask not installed")

    try:
        func = mod("ska_sdp_func")
        ft = mod("ska_sdp_func.fourier_transforms")
        sw = mod("ska_sdp_func.fourier_transforms.swiftly", Swiftly=_Unavailable)
        func.fourier_transforms = ft
        ft.swiftly = sw
        dask = mod("dask", delayed=_delayed)
        dask.array = mod("dask.array")
        dask.distributed = mod("dask.distributed")
        mod("distributed", Client=_Unavailable)
        if REF_SRC not in sys.path:
            sys.path.insert(0, REF_SRC)
        from ska_sdp_exec_swiftly import api, api_helper  # pylint: disable=import-outside-toplevel
        from ska_sdp_exec_swiftly.fourier_transform import core  # pylint: disable=import-outside-toplevel

        return api, api_helper, core
    except Exception:  # pylint: disable=broad-except
        return None


class _Impl:
    """The functions the slice calls, bound either to the reference or to the oracle port."""

    def __init__(self, params):
        p = params
        ref = _import_reference()
        if ref is not None:
            api, api_helper, core = ref
            self.kind = "reference"
            self.core = core.SwiftlyCore(p["W"], p["N"], p["xM_size"], p["yN_size"])
            self.extract_column = lambda bf, off0, off1: api_helper.extract_column(
                self.core, bf, off0, off1)
            self.FacetConfig = api.FacetConfig
            self.SubgridConfig = api.SubgridConfig
            self.sum_and_finish = lambda contribs, fcs, sg: api_helper.sum_and_finish_subgrid(
                self.core, contribs, fcs, sg)
            self.cover = api_helper.make_full_cover_config
        else:
            from oracle.swiftly_oracle import OracleCore  # pylint: disable=import-outside-toplevel

            self.kind = "port"
            self.core = OracleCore(p["W"], p["N"], p["xM_size"], p["yN_size"])
            core = self.core

            def extract_column(bf, off0, off1):  # api_helper.py:200-210
                return core.prepare_facet(core.extract_from_facet(bf, off0, axis=0), off1, axis=1)

            def sum_and_finish(contribs, fcs, sg):  # api_helper.py:73-112
                acc = None
                for off1 in sorted({fc.off1 for fc in fcs}):
                    col = None
                    for c, fc in zip(contribs, fcs):
                        if fc.off1 == off1:
                            col = core.add_to_subgrid(c, fc.off0, axis=0, out=col)
                    acc = core.add_to_subgrid(col, off1, axis=1, out=acc)
                out = core.finish_subgrid(acc, [sg.off0, sg.off1], sg.size)
                if sg.mask0 is not None:
                    out = out * numpy.asarray(sg.mask0)[:, None]
                if sg.mask1 is not None:
                    out = out * numpy.asarray(sg.mask1)[None, :]
                return out

            class _Cfg:  # pylint: disable=too-few-public-methods
                def __init__(self, off0, off1, size, mask0=None, mask1=None):
                    self.off0, self.off1, self.size = off0, off1, size
                    self.mask0, self.mask1 = mask0, mask1

            def cover(N, size, cls):
                offs = size * numpy.arange(int(numpy.ceil(N / size)))
                return [cls(int(a), int(b), size) for a in offs for b in offs]

            self.extract_column = extract_column
            self.sum_and_finish = sum_and_finish
            self.FacetConfig = _Cfg
            self.SubgridConfig = _Cfg
            self.cover = cover


def _shared(shape):
    """complex128 array in anonymous shared memory (inherited by forked workers)."""
    n = int(numpy.prod(shape))
    buf = mp.RawArray("d", 2 * n)
    return numpy.frombuffer(buf, dtype=numpy.complex128, count=n).reshape(shape)


def _phase_a(task):
    """One facet's share of the column -- or, when there are fewer facets than cores, one of
    ``chunks`` row blocks of it (the lines of both stages are independent; the reference's
    task is the whole facet, smaller tasks only help the CPU)."""
    j, c = task
    g = _G
    impl, fc = g["impl"], g["facet_cfgs"][j]
    chunks = g["chunks"]
    rng = numpy.random.default_rng(123456789 + j * 64 + c)
    yB = g["yB"]
    slab = max(1, g["slab"] // chunks)
    facet_slab = rng.standard_normal((yB, slab)) + 1j * rng.standard_normal((yB, slab))
    impl.core.prepare_facet(facet_slab, fc.off0, axis=0)  # this column's share of stage 1
    if chunks == 1:
        g["nmbf"][j][:] = impl.extract_column(g["bf_f"], g["sg_off0"], fc.off1)
        return j
    m = g["nmbf"][j].shape[0]
    r0, r1 = c * m // chunks, (c + 1) * m // chunks
    rows = impl.core.extract_from_facet(g["bf_f"], g["sg_off0"], axis=0)[r0:r1]
    g["nmbf"][j][r0:r1] = impl.core.prepare_facet(rows, fc.off1, axis=1)
    return j


def _phase_b(i):
    g = _G
    impl, sg = g["impl"], g["sg_cfgs"][i]
    contribs = [impl.core.extract_from_facet(g["nmbf"][j], sg.off1, axis=1)
                for j in range(len(g["facet_cfgs"]))]
    out = impl.sum_and_finish(contribs, g["facet_cfgs"], sg)
    return float(numpy.abs(out).max())


def run_column_slice(params, cores, max_facets=None, chunks=1):
    """Time one subgrid column of the forward transform on ``cores`` worker processes.

    Returns a dict with ``rate`` (contributions / s), ``contributions``, ``wall_s``, the phase
    times, ``kind`` and a description of the sample.
    """
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    impl = _Impl(params)
    N, yB, yN = params["N"], params["yB_size"], params["yN_size"]
    xA = params["xA_size"]
    m = impl.core.xM_yN_size
    facet_cfgs = impl.cover(N, yB, impl.FacetConfig)
    if max_facets is not None:
        facet_cfgs = facet_cfgs[:max_facets]
    all_sgs = impl.cover(N, xA, impl.SubgridConfig)
    ns = int(numpy.ceil(N / xA))
    column = ns // 2
    sg_cfgs = [sg for sg in all_sgs if sg.off0 == all_sgs[column * ns].off0]
    F = len(facet_cfgs)
    slab = max(1, yB // ns)
    rng = numpy.random.default_rng(7)
    bf_f = _shared((yN, yB))  # stands for a prepared facet (stage 2 input); values immaterial
    bf_f.real[:] = rng.standard_normal((yN, yB))
    bf_f.imag[:] = 0.5
    nmbf = [_shared((m, yN)) for _ in range(F)]
    chunks = max(1, int(chunks))
    _G.update(impl=impl, facet_cfgs=facet_cfgs, sg_cfgs=sg_cfgs, yB=yB, slab=slab, bf_f=bf_f,
              nmbf=nmbf, sg_off0=sg_cfgs[0].off0, chunks=chunks)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_phase_b, [])  # workers are up before the clock starts
        t0 = time.perf_counter()
        pool.map(_phase_a, [(j, c) for j in range(F) for c in range(chunks)], chunksize=1)
        t1 = time.perf_counter()
        peaks = pool.map(_phase_b, range(len(sg_cfgs)), chunksize=1)
        t2 = time.perf_counter()
    _G.clear()
    count = F * len(sg_cfgs)
    wall = t2 - t0
    return {
        "rate": count / wall, "contributions": count, "wall_s": wall,
        "phase_a_s": t1 - t0, "phase_b_s": t2 - t1, "kind": impl.kind, "cores": cores,
        "checksum": float(numpy.sum(peaks)),
        "sample": (f"one subgrid column ({len(sg_cfgs)} of {len(all_sgs)} subgrids) of the "
                   f"forward transform for {F} facets, real shapes, reference task order, on "
                   f"{cores} worker processes: per facet prepare_facet(axis 0) on its "
                   f"{slab}-column share of the facet + extract_column at full size"
                   + (f" (in {chunks} row blocks)" if chunks > 1 else "") + ", then per "
                   f"subgrid {F} contributions + sum_and_finish_subgrid; {count} contributions "
                   f"in {wall:.1f} s wall"),
    }
