"""
Multi-GPU correctness check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29501 tests/multi_gpu_check.py

Facets sharded over the ranks (NCCL all_to_all of the strips), every owned subgrid is
compared with the single-process CPU oracle.  Exit code 0 = parity.
"""

import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from oracle.swiftly_oracle import OracleCore, forward_reference_order
    from ska_sdp_distributed_fourier_transform_b200 import (
        FacetConfig, SwiftlyConfig, make_full_facet_cover, make_full_subgrid_cover)
    from ska_sdp_distributed_fourier_transform_b200.distributed import (
        SwiftlyForwardSharded, partition_facets)

    worst_all = 0.0
    for name, (W, N, yB, yN, xA, xM), sparse, exchange in (
        ("cfg1", (13.5625, 1024, 416, 512, 228, 256), False, "nccl"),
        ("cfg1-p2p", (13.5625, 1024, 416, 512, 228, 256), False, "p2p"),
        ("n2048", (13.5625, 2048, 512, 1024, 256, 512), False, "nccl"),
        ("n2048-p2p", (13.5625, 2048, 512, 1024, 256, 512), False, "p2p"),
        ("n2048-sparse-p2p", (13.5625, 2048, 512, 1024, 256, 512), True, "p2p"),
        ("n2048-sparse", (13.5625, 2048, 512, 1024, 256, 512), True, "nccl"),
        ("n2048-copy", (13.5625, 2048, 512, 1024, 256, 512), False, "copy"),
        ("n2048-sparse-copy", (13.5625, 2048, 512, 1024, 256, 512), True, "copy"),
    ):
        cfg = SwiftlyConfig(W=W, fov=1.0, N=N, yB_size=yB, yN_size=yN, xA_size=xA, xM_size=xM,
                            device=local)
        facet_cfgs = make_full_facet_cover(cfg)
        if sparse:
            facet_cfgs = [FacetConfig(a, b, yB) for a, b in
                          ((0, 0), (0, 512), (512, 0), (-512, 1024), (1024, 1024))]
        rng = numpy.random.default_rng(2024)
        facets = [rng.standard_normal((yB, yB)) + 1j * rng.standard_normal((yB, yB))
                  for _ in facet_cfgs]
        owner = partition_facets(facet_cfgs, world)
        local_facets = {i: facets[i] for i, o in enumerate(owner) if o == rank}
        fwd = SwiftlyForwardSharded(cfg, facet_cfgs, local_facets, lru_forward=1,
                                    exchange=exchange)
        sgs = make_full_subgrid_cover(cfg)
        sgs = sgs[:2 * world + 1] + sgs[-3:]
        tasks = fwd.get_subgrid_tasks(sgs)
        oracle = OracleCore(W, N, xM, yN)
        mine = sorted(tasks)
        ref = forward_reference_order(
            oracle, facets, [(c.off0, c.off1) for c in facet_cfgs],
            [(sgs[i].off0, sgs[i].off1) for i in mine], xA,
            subgrid_masks=[(sgs[i].mask0, sgs[i].mask1) for i in mine])
        worst = 0.0
        for r, i in zip(ref, mine):
            worst = max(worst, float(numpy.abs(tasks[i].result() - r).max() / numpy.abs(r).max()))
        # sharded backward on the same data: every rank supplies its own subgrids
        from oracle.swiftly_oracle import backward_reference_order
        from ska_sdp_distributed_fourier_transform_b200.distributed import SwiftlyBackwardSharded

        # (only at yB/yN = 0.5: at the reference's test parameters the backward chain amplifies
        # fp64 rounding by ~4e6, see tests/test_gpu_parity.py)
        back = {}
        if name.startswith("n2048"):
            bwd = SwiftlyBackwardSharded(cfg, facet_cfgs, lru_backward=1)
            bwd.add_subgrid_tasks(sgs, [tasks.get(i) for i in range(len(sgs))])
            back = bwd.finish()
        if back:
            full_ref = forward_reference_order(
                oracle, facets, [(c.off0, c.off1) for c in facet_cfgs],
                [(s.off0, s.off1) for s in sgs], xA,
                subgrid_masks=[(s.mask0, s.mask1) for s in sgs])
            back_ref = backward_reference_order(
                oracle, full_ref, [(s.off0, s.off1) for s in sgs],
                [(c.off0, c.off1) for c in facet_cfgs], yB,
                facet_masks=[(c.mask0, c.mask1) for c in facet_cfgs])
            bscale = max(float(numpy.abs(b).max()) for b in back_ref)
            for i, task in back.items():
                worst = max(worst, float(numpy.abs(task.result() - back_ref[i]).max() / bscale))
        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name}: world={world} exchange={fwd.exchange} max rel err {t.item():.2e}")
        worst_all = max(worst_all, t.item())
    # BASELINE cfg2 geometry (N=8192, m=1024, xM=2048: the ping-pong kernels, TMA tensor stores
    # into the owners' buffers, four-slot pipeline over more batches than slots): full facet
    # cover painted with point sources ON THE DEVICE, every owned subgrid vs the analytic DFT
    from ska_sdp_distributed_fourier_transform_b200 import make_facet_device
    from ska_sdp_distributed_fourier_transform_b200.fourier_algorithm import (
        make_subgrid_from_sources)

    W, N, yB, yN, xA, xM = 13.5625, 8192, 2048, 4096, 1024, 2048
    for exchange in ("nccl", "p2p", "copy"):
        cfg = SwiftlyConfig(W=W, fov=1.0, N=N, yB_size=yB, yN_size=yN, xA_size=xA, xM_size=xM,
                            device=local)
        facet_cfgs = make_full_facet_cover(cfg)
        rng = numpy.random.default_rng(8192)
        sources = [(float(rng.random()) + 0.5, int(rng.integers(-N // 2, N // 2)),
                    int(rng.integers(-N // 2, N // 2))) for _ in range(12)]
        owner = partition_facets(facet_cfgs, world)
        dev = torch.device("cuda", local)
        local_facets = {i: make_facet_device(N, facet_cfgs[i], sources, dev)
                        for i, o in enumerate(owner) if o == rank}
        fwd = SwiftlyForwardSharded(cfg, facet_cfgs, local_facets, lru_forward=1,
                                    exchange=exchange)
        sgs = make_full_subgrid_cover(cfg)
        sgs = sgs[:5 * world + 1] + sgs[-2:]
        tasks = fwd.get_subgrid_tasks(sgs)
        worst = 0.0
        for i, task in tasks.items():
            sg = sgs[i]
            truth = make_subgrid_from_sources(sources, N, xA, [sg.off0, sg.off1],
                                              [sg.mask0, sg.mask1])
            worst = max(worst, float(numpy.abs(task.result() - truth).max()
                                     / numpy.abs(truth).max()))
        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"cfg2: world={world} exchange={fwd.exchange} max rel err {t.item():.2e}")
        worst_all = max(worst_all, t.item())
        del fwd, tasks, local_facets
    dist.destroy_process_group()
    assert worst_all <= 1e-9, worst_all
    if rank == 0:
        print("MULTI-GPU PARITY OK")


if __name__ == "__main__":
    main()
