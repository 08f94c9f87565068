"""Quick per-kernel timing at the cfg4 geometry on a reduced facet set (dev tool)."""
import sys
import time

import numpy
import torch

sys.path.insert(0, ".")
from ska_sdp_distributed_fourier_transform_b200 import SwiftlyCoreB200  # noqa: E402

W, N, yB, yN, xA, xM = 13.5625, 65536, 8192, 16384, 2048, 4096
if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
    N, yB, yN, xA, xM = 32768, 4096, 8192, 2048, 4096
if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
    N, yB, yN, xA, xM = 8192, 2048, 4096, 1024, 2048
core = SwiftlyCoreB200(W, N, xM, yN)
m = core.xM_yN_size
dev = torch.device("cuda")
nf = 8
HBM = 6584.5e9


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return min(ts), sum(ts) / len(ts)


facet = torch.randn(yB, yB, dtype=torch.complex128, device=dev)
bf = torch.empty(yN, yB, dtype=torch.complex128, device=dev)
t, ta = timeit(lambda: core.prepare_facet(facet, 0, axis=0, out=bf), 3)
by = 16 * (yB * yB + yN * yB)
print(f"F1 prepare_facet ax0: {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
del facet
nmbf = [torch.empty(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)]
t, ta = timeit(lambda: core.extract_column(bf, 4096, 8192, out=nmbf[0]))
by = 16 * (m * yB + m * yN)
print(f"F2 extract_column:   {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
for i in range(1, nf):
    core.extract_column(bf, 4096, 8192 * i, out=nmbf[i])
strips = torch.empty(nf, m, xA, dtype=torch.complex128, device=dev)
srcs = [(nmbf[i], i * yB) for i in range(nf)]
t, ta = timeit(lambda: core.sum_finish_axis(srcs, strips[0], axis=1, subgrid_off=2048))
by = 16 * (nf * m * m + m * xA)
print(f"F3 sum_finish ax1 ({nf} src): {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
for i in range(1, nf):
    core.sum_finish_axis(srcs, strips[i], axis=1, subgrid_off=2048)
out = torch.empty(xA, xA, dtype=torch.complex128, device=dev)
srcs0 = [(strips[i], i * yB) for i in range(nf)]
t, ta = timeit(lambda: core.sum_finish_axis(srcs0, out, axis=0, subgrid_off=4096))
by = 16 * (nf * m * xA + xA * xA)
print(f"F4 sum_finish ax0 ({nf} src): {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
# plain copy for reference
a = torch.empty(1 << 27, dtype=torch.complex128, device=dev)
b = torch.empty_like(a)
t, ta = timeit(lambda: b.copy_(a))
print(f"copy 2 GiB: {t:.3f} ms -> {2*a.numel()*16/t*1e3/1e9:.0f} GB/s")
