"""
Shared API-level checks (SwiftlyForward / SwiftlyBackward / fused kernels) run by
tests/test_gpu_api.py on the B200 and by tests/test_emu_api.py on the host-emulated
kernels.  ``make_config(W, N, yB, yN, xA, xM)`` returns a SwiftlyConfig.
"""

import numpy
import torch

from oracle.swiftly_oracle import OracleCore, forward_reference_order
from ska_sdp_distributed_fourier_transform_b200 import (
    FacetConfig,
    SwiftlyBackward,
    SwiftlyForward,
    check_facet,
    check_subgrid,
    make_facet,
    make_full_facet_cover,
    make_full_subgrid_cover,
)
from ska_sdp_distributed_fourier_transform_b200.api import _device_of
from tests import parity_cases as pc


def devof(cfg):
    return _device_of(cfg.core)


def dev(cfg, arr):
    # clone: on the emulated (CPU) device .to() would alias the numpy array
    return torch.from_numpy(numpy.ascontiguousarray(arr)).clone().to(devof(cfg))


def case_fused_ops_vs_oracle(make_config):
    cfg = make_config(13.5625, 256, 96, 128, 52, 64)
    core = cfg.core
    oracle = OracleCore(13.5625, 256, 64, 128)
    assert core.fused_forward_supported()
    rng = numpy.random.default_rng(3)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    bf = pc.rand_c(rng, 128, 95)
    got = core.extract_column(dev(cfg, bf), 7 * Nx, -5 * Ny).cpu().numpy()
    ref = oracle.prepare_facet(oracle.extract_from_facet(bf, 7 * Nx, axis=0), -5 * Ny, axis=1)
    pc.close(got, ref, what="extract_column")
    # sum_finish along axis 1: three prepared-facet sources with overlapping windows
    srcs = [pc.rand_c(rng, 32, 128) for _ in range(3)]
    offs = [0, 24 * Ny, -24 * Ny]
    mask = (rng.random(51) > 0.3).astype(float)
    out = torch.empty((32, 51), dtype=torch.complex128, device=devof(cfg))
    core.sum_finish_axis([(dev(cfg, s), o) for s, o in zip(srcs, offs)], out, axis=1,
                         subgrid_off=-3 * Nx, mask=dev(cfg, mask))
    acc = None
    for s, o in zip(srcs, offs):
        acc = oracle.add_to_subgrid(oracle.extract_from_facet(s, -3 * Nx, axis=1), o, axis=1, out=acc)
    # finish along axis 1 only: use the 1-D finish on every row
    ref = numpy.array([oracle.finish_subgrid(row, -3 * Nx, 51) for row in acc]) * mask[None, :]
    pc.close(out.cpu().numpy(), ref, what="sum_finish_axis axis1")
    # axis 0 with contribution-sized sources (strips)
    strips = [pc.rand_c(rng, 32, 40) for _ in range(2)]
    out0 = torch.empty((52, 40), dtype=torch.complex128, device=devof(cfg))
    core.sum_finish_axis([(dev(cfg, s), o) for s, o in zip(strips, [0, 24 * Ny])], out0,
                         axis=0, subgrid_off=5 * Nx)
    acc = None
    for s, o in zip(strips, [0, 24 * Ny]):
        acc = oracle.add_to_subgrid(s, o, axis=0, out=acc)
    ref = numpy.array([oracle.finish_subgrid(col, 5 * Nx, 52) for col in acc.T]).T
    pc.close(out0.cpu().numpy(), ref, what="sum_finish_axis axis0")


def case_forward_backward_vs_reference_golden(make_config, golden_2d):
    g = golden_2d
    W, N, xM, yN, yB, xA = g["params"]
    cfg = make_config(float(W), int(N), int(yB), int(yN), int(xA), int(xM))
    facet_cfgs = make_full_facet_cover(cfg)
    sg_cfgs = make_full_subgrid_cover(cfg)
    assert [(c.off0, c.off1) for c in facet_cfgs] == [tuple(o) for o in g["facet_offs"]]
    assert [(c.off0, c.off1) for c in sg_cfgs] == [tuple(o) for o in g["sg_offs"]]
    fwd = SwiftlyForward(cfg, list(zip(facet_cfgs, g["facets"])), lru_forward=1, queue_size=5)
    bwd = SwiftlyBackward(cfg, facet_cfgs, lru_backward=1, queue_size=5)
    scale = numpy.abs(g["subgrids"]).max()
    for i, sg in enumerate(sg_cfgs):
        task = fwd.get_subgrid_task(sg)
        got = task.result()
        assert numpy.abs(got - g["subgrids"][i]).max() <= 1e-12 * scale, f"subgrid {i}"
        bwd.add_new_subgrid_task(sg, task)
    facets = [t.result() for t in bwd.finish()]
    bscale = numpy.abs(g["back_facets"]).max()
    for got, ref in zip(facets, g["back_facets"]):
        assert numpy.abs(got - ref).max() <= 1e-11 * bscale


def case_sparse_facets_shuffled_subgrids(make_config):
    """Arbitrary facet list (sparse cover, ragged rows) and shuffled subgrid order."""
    W, N, yB, yN, xA, xM = 13.5625, 256, 96, 128, 52, 64
    cfg = make_config(W, N, yB, yN, xA, xM)
    oracle = OracleCore(W, N, xM, yN)
    rng = numpy.random.default_rng(5)
    offs = [(0, 0), (0, 96), (96, 0), (-96, 192), (192, 192)]
    facet_cfgs = [FacetConfig(o0, o1, yB) for o0, o1 in offs]
    facets = [pc.rand_c(rng, yB, yB) for _ in offs]
    sg_cfgs = make_full_subgrid_cover(cfg)
    order = rng.permutation(len(sg_cfgs))[:7]
    fwd = SwiftlyForward(cfg, list(zip(facet_cfgs, facets)), lru_forward=2, queue_size=3)
    got = [fwd.get_subgrid_task(sg_cfgs[i]).result() for i in order]
    ref = forward_reference_order(
        oracle, facets, offs, [(sg_cfgs[i].off0, sg_cfgs[i].off1) for i in order], xA,
        subgrid_masks=[(sg_cfgs[i].mask0, sg_cfgs[i].mask1) for i in order],
    )
    scale = max(numpy.abs(r).max() for r in ref)
    for a, b in zip(got, ref):
        assert numpy.abs(a - b).max() <= 1e-12 * scale


def case_api_round_trip(make_config, lru_forward, lru_backward, shuffle):
    """reference tests/test_api.py:56-125 (round trip, unit source, facet RMSE < 3e-10)."""
    import random

    p = dict(W=13.5625, N=1024, yB=416, yN=512, xA=228, xM=256)
    cfg = make_config(**p)
    sources = [(1, 1, 0)]
    sg_cfgs = make_full_subgrid_cover(cfg)
    facet_cfgs = make_full_facet_cover(cfg)
    facet_tasks = [(fc, make_facet(cfg.image_size, fc, sources)) for fc in facet_cfgs]
    fwd = SwiftlyForward(cfg, facet_tasks, lru_forward, 100)
    bwd = SwiftlyBackward(cfg, facet_cfgs, lru_backward, 100)
    if shuffle:
        random.Random(1).shuffle(sg_cfgs)
    worst_sg = 0.0
    for sg in sg_cfgs:
        task = fwd.get_subgrid_task(sg)
        worst_sg = max(worst_sg, check_subgrid(cfg.image_size, sg, task.tensor, sources))
        bwd.add_new_subgrid_task(sg, task)
    assert worst_sg < 1e-13
    for fc, task in zip(facet_cfgs, bwd.finish()):
        assert check_facet(cfg.image_size, fc, task.result(), sources) < 3e-10


def case_fused_backward_ops_vs_oracle(make_config, W=13.5625, N=256, yB=96, yN=128, xA=52, xM=64,
                                      **kw):
    """subgrid_to_facets / fold_column against the unfused oracle chain."""
    cfg = make_config(W, N, yB, yN, xA, xM, **kw)
    core = cfg.core
    oracle = OracleCore(W, N, xM, yN)
    assert core.fused_backward_supported()
    m = core.xM_yN_size
    rng = numpy.random.default_rng(21)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    off1s = [0, 24 * Ny, -9 * Ny]
    blocks = [pc.rand_c(rng, m, xM) for _ in range(2)]
    use = [0, 1, 0]  # facet -> block (facets 0 and 2 share a row group)
    acc0 = [pc.rand_c(rng, m, yN) for _ in off1s]
    accs = [dev(cfg, a) for a in acc0]
    sg_off1 = 7 * Nx
    core.subgrid_to_facets([dev(cfg, blocks[u]) for u in use], accs, off1s, sg_off1)
    for j, (u, o1) in enumerate(zip(use, off1s)):
        ref = oracle.add_to_facet(oracle.extract_from_subgrid(blocks[u], o1, axis=1), sg_off1,
                                  axis=1, out=acc0[j].copy())
        pc.close(accs[j].cpu().numpy(), ref, what=f"subgrid_to_facets facet {j}")
    # fold: finish axis 1, mask, add axis 0
    sizes = [yB, yB - 1, yB]
    f0 = [pc.rand_c(rng, yN, s) for s in sizes]
    faccs = [dev(cfg, a) for a in f0]
    masks = [None, (rng.random(yB - 1) > 0.2).astype(float), None]
    sg_off0 = -5 * Nx
    core.fold_column(accs, faccs, off1s,
                     [None if mk is None else dev(cfg, mk) for mk in masks], sg_off0)
    for j in range(3):
        col = accs[j].cpu().numpy()
        part = oracle.finish_facet(col, off1s[j], sizes[j], axis=1)
        if masks[j] is not None:
            part = part * masks[j][None, :]
        ref = oracle.add_to_facet(part, sg_off0, axis=0, out=f0[j].copy())
        pc.close(faccs[j].cpu().numpy(), ref, what=f"fold_column facet {j}")


def case_many_sources(make_config):
    """More facets than one launch of the fused kernel carries (64 source slots, 16 groups):
    the C side cuts the job into several launches (api_helper.py:73-112 sums any list)."""
    # (1) one group, 40 sources with clashing windows: 40 rounds -> pieces ADD to the output
    cfg = make_config(13.5625, 256, 96, 128, 52, 64)
    core = cfg.core
    oracle = OracleCore(13.5625, 256, 64, 128)
    rng = numpy.random.default_rng(77)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step
    srcs = [pc.rand_c(rng, 8, 128) for _ in range(40)]
    offs = [int(rng.integers(-40, 40)) * Ny for _ in range(40)]
    mask = (rng.random(52) > 0.3).astype(float)
    out = torch.full((8, 52), 1e30 + 0j, dtype=torch.complex128, device=devof(cfg))
    core.sum_finish_axis([(dev(cfg, s), o) for s, o in zip(srcs, offs)], out, axis=1,
                         subgrid_off=5 * Nx, mask=dev(cfg, mask))
    acc = None
    for s, o in zip(srcs, offs):
        acc = oracle.add_to_subgrid(oracle.extract_from_facet(s, 5 * Nx, axis=1), o, axis=1, out=acc)
    ref = numpy.array([oracle.finish_subgrid(row, 5 * Nx, 52) for row in acc]) * mask[None, :]
    pc.close(out.cpu().numpy(), ref, what="sum_finish_axis, 40 sources")
    # (2) 20 groups in one grouped call
    groups = [[(dev(cfg, srcs[(3 * g + k) % 40]), offs[(3 * g + k) % 40]) for k in range(3)]
              for g in range(20)]
    outg = torch.empty((20, 8, 52), dtype=torch.complex128, device=devof(cfg))
    core.sum_finish_axis_grouped(groups, outg, axis=1, subgrid_off=-3 * Nx)
    for g in range(20):
        acc = None
        for k in range(3):
            i = (3 * g + k) % 40
            acc = oracle.add_to_subgrid(oracle.extract_from_facet(srcs[i], -3 * Nx, axis=1),
                                        offs[i], axis=1, out=acc)
        ref = numpy.array([oracle.finish_subgrid(row, -3 * Nx, 52) for row in acc])
        pc.close(outg[g].cpu().numpy(), ref, what=f"grouped, group {g}")
    # (3) the advisor's reproduction: 10 x 10 facet cover through SwiftlyForward
    W, N, yB, yN, xA, xM = 13.5625, 1024, 112, 512, 228, 256
    cfg = make_config(W, N, yB, yN, xA, xM)
    sources = [(1.0, 1, 0), (0.5, -300, 17), (0.25, 411, -222)]
    facet_cfgs = make_full_facet_cover(cfg)
    assert len(facet_cfgs) == 100
    fwd = SwiftlyForward(cfg, [(fc, make_facet(N, fc, sources)) for fc in facet_cfgs])
    sg_cfgs = make_full_subgrid_cover(cfg)
    for sg in (sg_cfgs[0], sg_cfgs[7], sg_cfgs[-1]):
        err = check_subgrid(N, sg, fwd.get_subgrid_task(sg).tensor, sources)
        assert err < 1e-12, err


def case_forward_backward_vs_oracle(make_config, W=13.5625, N=2048, yB=512, yN=1024, xA=256, xM=512,
                                    sg_variant=None):
    """Sparse facet list, forward + backward through the fused kernels vs the oracle in the
    reference's call order, at a geometry with xA / xM = 0.5 like the BASELINE configs: the
    ping-pong subgrid kernel stages finished lines and hands them to the TMA engine (tensor
    store; transposed strips)."""
    from oracle.swiftly_oracle import backward_reference_order

    cfg = make_config(W, N, yB, yN, xA, xM)
    if sg_variant is not None:
        import ctypes

        cfg.core._lib.swiftly_b200_debug_sg_variant.argtypes = [ctypes.c_void_p, ctypes.c_int]
        cfg.core._lib.swiftly_b200_debug_sg_variant(cfg.core._plan, sg_variant)
    oracle = OracleCore(W, N, xM, yN)
    rng = numpy.random.default_rng(1)
    offs = [(0, 0), (0, yB), (yB, 0), (-yB, 2 * yB), (2 * yB, 2 * yB)]
    facet_cfgs = [FacetConfig(a, b, yB) for a, b in offs]
    facets = [pc.rand_c(rng, yB, yB) for _ in offs]
    sgs = make_full_subgrid_cover(cfg)
    sgs = [sgs[0], sgs[1], sgs[9], sgs[-1]]
    fwd = SwiftlyForward(cfg, list(zip(facet_cfgs, facets)), lru_forward=1, queue_size=4)
    bwd = SwiftlyBackward(cfg, facet_cfgs, lru_backward=1, queue_size=4)
    got = []
    for sg in sgs:
        task = fwd.get_subgrid_task(sg)
        got.append(task.result())
        bwd.add_new_subgrid_task(sg, task)
    back = [t.result() for t in bwd.finish()]
    sg_offs = [(s.off0, s.off1) for s in sgs]
    ref = forward_reference_order(oracle, facets, offs, sg_offs, xA,
                                  subgrid_masks=[(s.mask0, s.mask1) for s in sgs])
    scale = max(numpy.abs(b).max() for b in ref)
    for a, b in zip(got, ref):
        assert numpy.abs(a - b).max() <= 1e-12 * scale
    bref = backward_reference_order(oracle, ref, sg_offs, offs, yB)
    bscale = max(numpy.abs(b).max() for b in bref)
    for a, b in zip(back, bref):
        assert numpy.abs(a - b).max() <= 1e-11 * bscale
