"""Run ONE kernel of the forward path a few times (for ncu): prof_one.py <f1|f2|f3|f4> [cfg]."""
import sys

import torch

sys.path.insert(0, ".")
from ska_sdp_distributed_fourier_transform_b200 import SwiftlyCoreB200  # noqa: E402

which = sys.argv[1]
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg4"
W = 13.5625
N, yB, yN, xA, xM = {"cfg4": (65536, 8192, 16384, 2048, 4096), "cfg3": (32768, 4096, 8192, 2048, 4096),
                     "cfg2": (8192, 2048, 4096, 1024, 2048)}[cfg]
core = SwiftlyCoreB200(W, N, xM, yN)
m = core.xM_yN_size
dev = torch.device("cuda")
nf = 8
reps = 3
if which == "f1":
    facet = torch.randn(yB, yB, dtype=torch.complex128, device=dev)
    bf = torch.empty(yN, yB, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        core.prepare_facet(facet, 0, axis=0, out=bf, window_lines=True)
elif which == "f2":
    bfs = [torch.randn(yN, yB, dtype=torch.complex128, device=dev) for _ in range(nf)]
    outs = [torch.empty(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)]
    for _ in range(reps):
        core.extract_columns(bfs, 4096, [yB * i for i in range(nf)], outs=outs, prewindowed=True)
elif which in ("f3", "f4"):
    # distinct prepared facets per facet row, as in a real step (no L2 reuse across groups)
    nrows = nf if which == "f3" else 1
    nmbf = [[torch.randn(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)]
            for _ in range(nrows)]
    # strips transposed (contribution index contiguous), as the product stores them
    strips = torch.randn(nf, xA, m, dtype=torch.complex128, device=dev).transpose(1, 2)
    out = torch.empty(xA, xA, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        if which == "f3":
            core.sum_finish_axis_grouped([[(nmbf[g][i], i * yB) for i in range(nf)] for g in range(nf)],
                                         strips, axis=1, subgrid_off=2048)
        else:
            core.sum_finish_axis([(strips[i], i * yB) for i in range(nf)], out, axis=0, subgrid_off=4096)
torch.cuda.synchronize()
