"""Time one full backward (subgrid -> facet) transform at a BASELINE geometry (dev tool)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from ska_sdp_distributed_fourier_transform_b200 import (  # noqa: E402
    SWIFT_CONFIGS, SwiftlyBackward, SwiftlyConfig, make_full_facet_cover, make_full_subgrid_cover)

name = {"cfg4": "64k[1]-n16k-4k", "cfg3": "32k[1]-n8k-4k", "cfg2": "8k[1]-n4k-2k"}[
    sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
cfg = SwiftlyConfig(**SWIFT_CONFIGS[name])
facet_cfgs = make_full_facet_cover(cfg)
sg_cfgs = make_full_subgrid_cover(cfg)
xA = cfg.max_subgrid_size
dev = torch.device("cuda")
subgrids = [torch.randn(xA, xA, dtype=torch.complex128, device=dev) for _ in range(4)]
for rep in range(2):
    bwd = SwiftlyBackward(cfg, facet_cfgs, lru_backward=1, queue_size=8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i, sg in enumerate(sg_cfgs):
        bwd.add_new_subgrid_task(sg, subgrids[i % 4])
    e1.record()
    tasks = bwd.finish()
    e2.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n = len(facet_cfgs) * len(sg_cfgs)
    print(f"{name} rep {rep}: subgrids {e0.elapsed_time(e1):.1f} ms, finish {e1.elapsed_time(e2):.1f} ms, "
          f"wall {wall*1e3:.1f} ms -> {n / wall:.0f} subgrid->facet contributions/s "
          f"(mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)", flush=True)
    del bwd, tasks
    torch.cuda.empty_cache()
