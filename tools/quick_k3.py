"""Fused subgrid kernel (axis 1: 8 facet rows x 8 facets, transposed strips; axis 0) at the cfg4
geometry for a list of sg_variant values (dev tool): python tools/quick_k3.py 0 21 22 23"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from ska_sdp_distributed_fourier_transform_b200 import SwiftlyCoreB200  # noqa: E402

W, N, yB, yN, xA, xM = 13.5625, 65536, 8192, 16384, 2048, 4096
core = SwiftlyCoreB200(W, N, xM, yN)
m = core.xM_yN_size
dev = torch.device("cuda")
nf = 8
HBM = 6584.5e9
core._lib.swiftly_b200_debug_sg_variant.argtypes = [ctypes.c_void_p, ctypes.c_int]
variants = [int(v) for v in sys.argv[1:]] or [0]


def timeit(fn, reps=10):
    fn()
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


big = [[torch.randn(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)] for _ in range(nf)]
groups = [[(big[g][i], i * yB) for i in range(nf)] for g in range(nf)]
strips_t = torch.empty(nf, xA, m, dtype=torch.complex128, device=dev).transpose(1, 2)
out = torch.empty(xA, xA, dtype=torch.complex128, device=dev)
ref3 = ref4 = None
for v in variants:
    core._lib.swiftly_b200_debug_sg_variant(core._plan, v)
    t, ta = timeit(lambda: core.sum_finish_axis_grouped(groups, strips_t, axis=1, subgrid_off=2048))
    by = 16 * nf * (nf * m * m + m * xA)
    msg = f"K3 variant {v}: {t:.4f} ms (avg {ta:.4f})  frac {by/t*1e3/HBM:.3f}"
    if ref3 is None:
        ref3 = strips_t.clone()
    else:
        msg += f"  max|diff| {(strips_t - ref3).abs().max().item():.2e}"
    print(msg, flush=True)
    srcs0t = [(strips_t[i], i * yB) for i in range(nf)]
    t, ta = timeit(lambda: core.sum_finish_axis(srcs0t, out, axis=0, subgrid_off=4096))
    by = 16 * (nf * m * xA + xA * xA)
    msg = f"K4 variant {v}: {t:.4f} ms (avg {ta:.4f})  frac {by/t*1e3/HBM:.3f}"
    if ref4 is None:
        ref4 = out.clone()
    else:
        msg += f"  max|diff| {(out - ref4).abs().max().item():.2e}"
    print(msg, flush=True)
    # K3 then K4 back to back, as in a step
    def both():
        core.sum_finish_axis_grouped(groups, strips_t, axis=1, subgrid_off=2048)
        core.sum_finish_axis(srcs0t, out, axis=0, subgrid_off=4096)
    t, ta = timeit(both)
    print(f"K3+K4 variant {v}: {t:.4f} ms (avg {ta:.4f})", flush=True)
core._lib.swiftly_b200_debug_sg_variant(core._plan, 0)
