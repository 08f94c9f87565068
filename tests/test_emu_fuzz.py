"""
Randomised parity (hypothesis) of the eight primitives on the host-emulated kernels:
parameter sets with power-of-two and F * 2^k lengths, odd / even facet and subgrid sizes,
offsets anywhere in [-3N, 3N] (multiples of the offset steps), 1-D and 2-D along both axes.
"""

import numpy
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle.swiftly_oracle import OracleCore
from tests import parity_cases as pc
from tests.emu_support import emu_core_class

# (N, xM, yN): m = xM * yN / N must be an integer >= 16 with supported lengths
PARAM_SETS = [
    (256, 64, 128), (512, 128, 128), (512, 64, 256), (1024, 256, 512), (2048, 256, 1024),
    (768, 192, 384), (1280, 320, 640), (1536, 512, 768), (1792, 256, 1792), (2304, 576, 1152),
]
_cores = {}


def cores(idx):
    if idx not in _cores:
        N, xM, yN = PARAM_SETS[idx]
        _cores[idx] = (emu_core_class()(11.0, N, xM, yN), OracleCore(11.0, N, xM, yN))
    return _cores[idx]


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck))
@given(idx=st.integers(0, len(PARAM_SETS) - 1), fo=st.integers(-3, 3), so=st.integers(-3, 3),
       fk=st.integers(0, 40), sk=st.integers(0, 40), yb_frac=st.floats(0.2, 0.85),
       xa_frac=st.floats(0.2, 1.0), seed=st.integers(0, 2**31 - 1), axis=st.integers(0, 1),
       other=st.integers(1, 9))
def test_fuzz_primitives(idx, fo, so, fk, sk, yb_frac, xa_frac, seed, axis, other):
    core, oracle = cores(idx)
    N, xM, yN = PARAM_SETS[idx]
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    facet_off = fo * N + fk * Ny
    sg_off = so * N + sk * Nx
    yB = max(1, min(yN - 1, int(yb_frac * yN)))
    xA = max(1, min(xM, int(xa_frac * xM)))
    rng = numpy.random.default_rng(seed)
    # Fb grows steeply towards yB -> yN (1 / PSWF): the rounding of the yN-point transform is
    # amplified by max(Fb) in finish_facet, hence 1e-10 here instead of the 1e-12 of the fixed cases
    pc.check_1d_chain(core, oracle, yB, xA, facet_off, sg_off, rng, rtol=1e-10)
    pc.check_2d_axis(core, oracle, yB, axis, other, facet_off, sg_off, rng, rtol=1e-10)


@pytest.mark.parametrize("idx", [0, 5])
def test_fuzz_degenerate_sizes(idx):
    """size-1 facets / subgrids and the largest legal facet (yN - 1)."""
    core, oracle = cores(idx)
    N, xM, yN = PARAM_SETS[idx]
    rng = numpy.random.default_rng(3)
    pc.check_1d_chain(core, oracle, 1, 1, 0, 0, rng)
    pc.check_1d_chain(core, oracle, yN - 1, xM, core.facet_off_step, -core.subgrid_off_step, rng,
                      rtol=1e-7)


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck))
@given(n_src=st.integers(1, 9), seed=st.integers(0, 2**31 - 1), axis=st.integers(0, 1),
       sk=st.integers(-40, 40), use_mask=st.booleans(), contrib_sized=st.booleans())
def test_fuzz_fused_sum_finish(n_src, seed, axis, sk, use_mask, contrib_sized):
    """Fused sum-and-finish kernel with arbitrary source sets: random facet offsets (any
    overlap pattern -> round scheduling, tiling shortcut on/off), both axes, masks."""
    import torch

    core, oracle = cores(0)  # N=256, xM=64, yN=128, m=32
    N, xM, yN = PARAM_SETS[0]
    m = core.xM_yN_size
    rng = numpy.random.default_rng(seed)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    sg_off = sk * Nx
    offs = [int(rng.integers(-64, 64)) * Ny for _ in range(n_src)]
    size = m if contrib_sized else yN
    lines, sz = 7, int(rng.integers(1, xM + 1))
    shape = (lines, size) if axis == 1 else (size, lines)
    srcs = [pc.rand_c(rng, *shape) for _ in range(n_src)]
    mask = (rng.random(sz) > 0.3).astype(float) if use_mask else None
    out = torch.empty((lines, sz) if axis == 1 else (sz, lines), dtype=torch.complex128)
    core.sum_finish_axis([(torch.from_numpy(s.copy()), o) for s, o in zip(srcs, offs)], out,
                         axis=axis, subgrid_off=sg_off,
                         mask=None if mask is None else torch.from_numpy(mask))
    acc = None
    for s, o in zip(srcs, offs):
        c = s if contrib_sized else oracle.extract_from_facet(s, sg_off, axis=axis)
        acc = oracle.add_to_subgrid(c, o, axis=axis, out=acc)
    fin = [oracle.finish_subgrid(line, sg_off, sz) for line in (acc if axis == 1 else acc.T)]
    ref = numpy.array(fin)
    if mask is not None:
        ref = ref * mask[None, :]
    if axis == 0:
        ref = ref.T
    pc.close(out.numpy(), ref, rtol=1e-11, what="fused sum_finish_axis")


def test_core_pickles_by_parameters():
    """Like SwiftlyCoreFunc (core.py:513-525) the core pickles by constructor arguments."""
    core, _ = cores(0)
    state = core.__getstate__()
    assert state == {"W": 11.0, "N": 256, "xM_size": 64, "yN_size": 128, "device": 0}
    assert repr(core).endswith("(W=11.0, N=256, xM_size=64, yN_size=128)")


_k2_cores = {}


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(form=st.sampled_from([2, 3, 4, 5, 6]), n_facets=st.integers(1, 3),
       sizes=st.lists(st.integers(8, 500), min_size=3, max_size=3),
       whole_chunks=st.booleans(), fk=st.lists(st.integers(-40, 40), min_size=3, max_size=3),
       sk=st.integers(-64, 64), blocks=st.integers(1, 9), seed=st.integers(0, 2**31 - 1))
def test_fuzz_k2_split_forms(form, n_facets, sizes, whole_chunks, fk, sk, blocks, seed):
    """``extract_columns`` through the split K2 kernels that serve yN = 16384 on the GPU, forced
    at yN = 512 (2 = 4 x Q with the L2 scratch, 3 = DIF / DIT with the L2 scratch, 4 / 5 / 6 = the
    tensor-memory kernels: DIF with pair stores, DIT, DIT with the store phases half a line
    apart): random facet counts and row lengths (whole 128-byte chunks -> swizzled tensor
    loads, otherwise linear bulk copies; longer and shorter than yN / 2), offsets, and grid
    sizes from one CTA walking every line to one line per CTA."""
    import ctypes

    import torch

    W, N, xM, yN = 13.5625, 1024, 256, 512
    if form not in _k2_cores:
        _k2_cores[form] = (emu_core_class()(W, N, xM, yN, force_split=form), OracleCore(W, N, xM, yN))
    core, oracle = _k2_cores[form]
    core._lib.swiftly_b200_debug_max_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int]
    core._lib.swiftly_b200_debug_max_blocks(core._plan, blocks * 16 if blocks == 9 else blocks)
    rng = numpy.random.default_rng(seed)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    fss = [(fs // 8) * 8 if whole_chunks else fs for fs in sizes[:n_facets]]
    fss = [max(8, fs) for fs in fss]
    offs = [k * Ny for k in fk[:n_facets]]
    sg_off0 = sk * Nx
    bfs = [pc.rand_c(rng, yN, fs) for fs in fss]
    refs = [oracle.prepare_facet(oracle.extract_from_facet(bf, sg_off0, axis=0), off1, axis=1)
            for bf, off1 in zip(bfs, offs)]
    outs = core.extract_columns([torch.from_numpy(bf.copy()) for bf in bfs], sg_off0, offs)
    for o, r in zip(outs, refs):
        assert numpy.abs(o.numpy() - r).max() <= 1e-11 * numpy.abs(r).max()
