"""
N > 1 path on CPU: two ``gloo`` ranks, facets sharded by ``partition_facets``, strips
exchanged with ``all_to_all``, owners finish their subgrids.  The kernels run on the
host-emulated library (test tooling); what is under test is the sharding, the strip
layout, the batching / double buffering and the ownership logic of
``SwiftlyForwardSharded`` -- checked against the single-process oracle.
"""

import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import parity_cases as pc

W, N, yB, yN, xA, xM = 13.5625, 256, 96, 128, 52, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, sparse, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.swiftly_oracle import OracleCore, forward_reference_order
        from ska_sdp_distributed_fourier_transform_b200 import (
            FacetConfig, SwiftlyConfig, make_full_facet_cover, make_full_subgrid_cover)
        from ska_sdp_distributed_fourier_transform_b200.distributed import (
            SwiftlyForwardSharded, partition_facets)
        from tests.emu_support import emu_core_class

        core = emu_core_class()(W, N, xM, yN)
        cfg = SwiftlyConfig(W=W, fov=1.0, N=N, yB_size=yB, yN_size=yN, xA_size=xA,
                            xM_size=xM, core=core)
        if sparse:
            offs = [(0, 0), (0, 96), (96, 0), (-96, 192), (192, 192)]
            facet_cfgs = [FacetConfig(a, b, yB) for a, b in offs]
        else:
            facet_cfgs = make_full_facet_cover(cfg)
        rng = numpy.random.default_rng(42)
        facets = [pc.rand_c(rng, yB, yB) for _ in facet_cfgs]  # same on every rank
        owner = partition_facets(facet_cfgs, world)
        assert sorted(set(owner)) == list(range(world))
        local = {i: facets[i] for i, o in enumerate(owner) if o == rank}
        fwd = SwiftlyForwardSharded(cfg, facet_cfgs, local, lru_forward=1)
        sgs = make_full_subgrid_cover(cfg)[:7]  # 7 subgrids: last batch is ragged
        tasks = fwd.get_subgrid_tasks(sgs)
        assert sorted(tasks) == [i for i in range(len(sgs)) if i % world == rank]
        oracle = OracleCore(W, N, xM, yN)
        ref = forward_reference_order(
            oracle, facets, [(c.off0, c.off1) for c in facet_cfgs],
            [(s.off0, s.off1) for s in sgs], xA,
            subgrid_masks=[(s.mask0, s.mask1) for s in sgs])
        scale = max(numpy.abs(r).max() for r in ref)
        worst = 0.0
        for i, t in tasks.items():
            worst = max(worst, numpy.abs(t.result() - ref[i]).max() / scale)
        # backward, sharded: every rank supplies the subgrids it owns, facets stay local
        from oracle.swiftly_oracle import backward_reference_order
        from ska_sdp_distributed_fourier_transform_b200.distributed import SwiftlyBackwardSharded

        bwd = SwiftlyBackwardSharded(cfg, facet_cfgs, lru_backward=1)
        supply = [tasks.get(i) for i in range(len(sgs))]
        bwd.add_subgrid_tasks(sgs, supply)
        mine = bwd.finish()
        assert sorted(mine) == [i for i, o in enumerate(owner) if o == rank]
        back_ref = backward_reference_order(
            oracle, ref, [(s.off0, s.off1) for s in sgs],
            [(c.off0, c.off1) for c in facet_cfgs], yB,
            facet_masks=[(c.mask0, c.mask1) for c in facet_cfgs])
        bscale = max(numpy.abs(b).max() for b in back_ref)
        for i, t in mine.items():
            worst = max(worst, numpy.abs(t.result() - back_ref[i]).max() / bscale)
        q.put((rank, worst, len(tasks)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sparse", [False, True])
def test_sharded_forward_two_ranks_gloo(sparse):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sparse, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert [g[0] for g in got] == [0, 1]
    assert sum(g[2] for g in got) == 7
    for _, worst, _ in got:
        assert worst <= 1e-11


def test_partition_facets_full_and_sparse():
    from ska_sdp_distributed_fourier_transform_b200 import FacetConfig
    from ska_sdp_distributed_fourier_transform_b200.distributed import partition_facets

    full = [FacetConfig(a * 8192, b * 8192, 8192) for a in range(8) for b in range(8)]
    for world in (1, 2, 4, 8):
        owner = partition_facets(full, world)
        # whole facet rows per rank
        for r in range(world):
            rows = {full[i].off0 for i, o in enumerate(owner) if o == r}
            assert len(rows) == 8 // world
        assert [owner.count(r) for r in range(world)] == [64 // world] * world
    sparse = [FacetConfig(a, b, 8192) for a in (0, 8192, 49152, 57344) for b in (0, 8192, 49152, 57344)]
    owner = partition_facets(sparse, 8)
    assert [owner.count(r) for r in range(8)] == [2] * 8
