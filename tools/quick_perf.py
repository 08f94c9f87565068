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


import ctypes  # noqa: E402

core._lib.swiftly_b200_debug_sg_variant.argtypes = [ctypes.c_void_p, ctypes.c_int]
facet = torch.randn(yB, yB, dtype=torch.complex128, device=dev)
bf = torch.empty(yN, yB, dtype=torch.complex128, device=dev)
by = 16 * (yB * yB + yN * yB)
for variant, name in ((8, "one tile (round 1)"), (9, "128 MiB column tiles"), (0, "64 MiB column tiles")):
    core._lib.swiftly_b200_debug_sg_variant(core._plan, variant)
    t, ta = timeit(lambda: core.prepare_facet(facet, 0, axis=0, out=bf), 3)
    print(f"F1 prepare_facet ax0 [{name}]: {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
del facet
nmbf = [torch.empty(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)]
t, ta = timeit(lambda: core.extract_column(bf, 4096, 8192, out=nmbf[0]))
by = 16 * (m * yB + m * yN)
print(f"F2 extract_column:   {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
for i in range(1, nf):
    core.extract_column(bf, 4096, 8192 * i, out=nmbf[i])
bfs = [bf] * nf
offs = [yB * i for i in range(nf)]
by = 16 * (m * yB + m * yN) * nf
keep = None
for variant, name in ((4, "round-1 kernel"), (7, "TMA rows, 2 x (yN/2), swizzled"),
                      (15, "TMA rows, two independent groups (DIF across / DIT within)"),
                      (20, "TMA rows, 4 x 4096, two groups + CTA-wide combine, L2 scratch"),
                      (0, "TMA rows, default (DIT / DIT, TMEM parking + swap, skewed stores)")):
    core._lib.swiftly_b200_debug_sg_variant(core._plan, variant)
    t, ta = timeit(lambda: core.extract_columns(bfs, 4096, offs, outs=nmbf))
    print(f"F2 extract_columns x{nf} [{name}]: {t:.3f} ms (avg {ta:.3f})  frac {by/t*1e3/HBM:.3f}")
    if keep is None:
        keep = nmbf[3].clone()
    else:
        print(f"   max |diff|: {(nmbf[3] - keep).abs().max().item():.3e} (max |ref| {keep.abs().max().item():.3e})")
core._lib.swiftly_b200_debug_sg_variant(core._plan, 0)
t, ta = timeit(lambda: core.extract_columns(bfs, 4096, offs, outs=nmbf, prewindowed=True))
print(f"F2 extract_columns x{nf} [default kernel, PRE-WINDOWED rows (no Fb fetch)]: {t:.3f} ms (avg {ta:.3f})  frac {by/t*1e3/HBM:.3f}")
core.extract_columns(bfs, 4096, offs, outs=nmbf)
strips = torch.empty(nf, m, xA, dtype=torch.complex128, device=dev)
srcs = [(nmbf[i], i * yB) for i in range(nf)]
ref = None
for variant, name in ((1, "round-1 kernel"), (2, "two groups + LSU token"), (0, "two groups (default)")):
    core._lib.swiftly_b200_debug_sg_variant(core._plan, variant)
    t, ta = timeit(lambda: core.sum_finish_axis(srcs, strips[0], axis=1, subgrid_off=2048))
    by = 16 * (nf * m * m + m * xA)
    print(f"F3 sum_finish ax1 ({nf} src) [{name}]: {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
    # as the step launches it: all 8 facet rows of a subgrid in one grouped launch
    t, ta = timeit(lambda: core.sum_finish_axis_grouped([srcs] * nf, strips, axis=1, subgrid_off=2048))
    by = 16 * nf * (nf * m * m + m * xA)
    print(f"F3 grouped x{nf} [{name}]: {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
    if ref is None:
        ref = strips.clone()
    else:
        print(f"   max |diff| vs round-1 kernel: {(strips - ref).abs().max().item():.3e} (max |ref| {ref.abs().max().item():.3e})")
for i in range(1, nf):
    core.sum_finish_axis(srcs, strips[i], axis=1, subgrid_off=2048)
out = torch.empty(xA, xA, dtype=torch.complex128, device=dev)
srcs0 = [(strips[i], i * yB) for i in range(nf)]
t, ta = timeit(lambda: core.sum_finish_axis(srcs0, out, axis=0, subgrid_off=4096))
by = 16 * (nf * m * xA + xA * xA)
print(f"F4 sum_finish ax0 ({nf} src): {t:.3f} ms (avg {ta:.3f})  {by/t*1e3/1e9:.0f} GB/s  frac {by/t*1e3/HBM:.3f}")
# transposed strips (what the product uses): finished lines leave through the TMA engine.
# Distinct prepared facets per group as in a real step (64 x 256 MiB: no L2 reuse across groups).
ref0 = out.clone()
strips_t = torch.empty(nf, xA, m, dtype=torch.complex128, device=dev).transpose(1, 2)
big = [[torch.randn(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)] for _ in range(nf)]
groups = [[(big[g][i], i * yB) for i in range(nf)] for g in range(nf)]
for variant, name in ((1, "round-1 kernel"), (11, "two groups, no L2 prefetch"),
                      (12, "two groups, per-thread prefetch"), (13, "two groups, group 1 starts 5 us late"),
                      (16, "two groups, complex exchange through the accumulator in round 1"), (0, "two groups, bulk prefetch (default)")):
    core._lib.swiftly_b200_debug_sg_variant(core._plan, variant)
    t, ta = timeit(lambda: core.sum_finish_axis_grouped(groups, strips_t, axis=1, subgrid_off=2048))
    by = 16 * nf * (nf * m * m + m * xA)
    print(f"F3 grouped x{nf}, distinct facets -> TRANSPOSED strips [{name}]: {t:.3f} ms (avg {ta:.3f})  frac {by/t*1e3/HBM:.3f}")
del big, groups
for i in range(nf):
    core.sum_finish_axis(srcs, strips_t[i], axis=1, subgrid_off=2048)
for variant, name in ((5, "two groups, direct 16-byte stores"), (0, "two groups, TMA tensor stores (default)")):
    core._lib.swiftly_b200_debug_sg_variant(core._plan, variant)
    srcs0t = [(strips_t[i], i * yB) for i in range(nf)]
    t, ta = timeit(lambda: core.sum_finish_axis(srcs0t, out, axis=0, subgrid_off=4096))
    by = 16 * (nf * m * xA + xA * xA)
    print(f"F4 sum_finish ax0 from TRANSPOSED strips [{name}]: {t:.3f} ms (avg {ta:.3f})  frac {by/t*1e3/HBM:.3f}")
    print(f"   max |diff| vs row-major path: {(out - ref0).abs().max().item():.3e} (max |ref| {ref0.abs().max().item():.3e})")
core._lib.swiftly_b200_debug_sg_variant(core._plan, 0)
# plain copy for reference
a = torch.empty(1 << 27, dtype=torch.complex128, device=dev)
b = torch.empty_like(a)
t, ta = timeit(lambda: b.copy_(a))
print(f"copy 2 GiB: {t:.3f} ms -> {2*a.numel()*16/t*1e3/1e9:.0f} GB/s")
