"""
The peer-memory (p2p) driver of the sharded forward transform on the host-emulated kernels:
single process, the symmetric-memory module replaced by plain tensors, so that the software
pipeline (four slots, signal / wait flags, scattered per-group outputs of the axis-1 kernel)
is exercised without GPUs.  TEST TOOLING -- the real thing runs in tests/multi_gpu_check.py.
"""

import numpy
import pytest
import torch

from oracle.swiftly_oracle import OracleCore, forward_reference_order
from ska_sdp_distributed_fourier_transform_b200 import (
    FacetConfig,
    SwiftlyConfig,
    make_full_subgrid_cover,
)
from ska_sdp_distributed_fourier_transform_b200.distributed import SwiftlyForwardSharded
from tests import parity_cases as pc
from tests.emu_support import emu_core_class


class _FakeHandle:
    def __init__(self, buf):
        self.buf = buf

    def get_buffer(self, rank, shape, dtype, offset):  # pylint: disable=unused-argument
        n = int(numpy.prod(shape))
        return self.buf.view(dtype)[:n].view(shape)

    def barrier(self, channel=0):  # pylint: disable=unused-argument
        pass


class _FakeSymm:
    @staticmethod
    def empty(n, dtype, device):  # pylint: disable=unused-argument
        return torch.empty(n, dtype=dtype)

    @staticmethod
    def rendezvous(buf, group):  # pylint: disable=unused-argument
        return _FakeHandle(buf)


@pytest.mark.parametrize("mode", ["p2p", "copy"])
def test_emu_p2p_pipeline_single_rank(mode):
    W, N, yB, yN, xA, xM = 13.5625, 2048, 512, 1024, 256, 512
    core = emu_core_class()(W, N, xM, yN)
    cfg = SwiftlyConfig(W=W, fov=1.0, N=N, yB_size=yB, yN_size=yN, xA_size=xA, xM_size=xM,
                        core=core)
    rng = numpy.random.default_rng(11)
    offs = [(0, 0), (0, yB), (yB, 0), (-yB, 2 * yB)]
    facet_cfgs = [FacetConfig(a, b, yB) for a, b in offs]
    facets = [pc.rand_c(rng, yB, yB) for _ in offs]
    fwd = SwiftlyForwardSharded(cfg, facet_cfgs, dict(enumerate(facets)), exchange="nccl")
    fwd._symm = {"mod": _FakeSymm, "group": None, "slots": {}}  # pylint: disable=protected-access
    fwd.exchange = mode
    sgs = make_full_subgrid_cover(cfg)
    sgs = sgs[:6] + sgs[-1:]  # seven batches of one: every slot is reused
    tasks = fwd.get_subgrid_tasks(sgs)
    oracle = OracleCore(W, N, xM, yN)
    ref = forward_reference_order(oracle, facets, offs, [(s.off0, s.off1) for s in sgs], xA,
                                  subgrid_masks=[(s.mask0, s.mask1) for s in sgs])
    scale = max(numpy.abs(r).max() for r in ref)
    assert sorted(tasks) == list(range(len(sgs)))
    for i, r in enumerate(ref):
        assert numpy.abs(tasks[i].result() - r).max() <= 1e-12 * scale
    # a second call keeps counting the flags upwards
    tasks = fwd.get_subgrid_tasks(sgs[:2])
    for i in range(2):
        assert numpy.abs(tasks[i].result() - ref[i]).max() <= 1e-12 * scale
