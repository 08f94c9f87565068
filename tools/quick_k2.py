"""K2 (extract_columns) variants at the cfg4 geometry: time per 8 facets and difference to the
default kernel on identical data (dev tool; run under `timeout`)."""
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


def timeit(fn, reps=8):
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


# distinct prepared facets (8 x 2 GiB): no L2 reuse between facets, as in a step
bfs = [torch.randn(yN, yB, dtype=torch.complex128, device=dev) for _ in range(nf)]
nmbf = [torch.empty(m, yN, dtype=torch.complex128, device=dev) for _ in range(nf)]
offs = [yB * i for i in range(nf)]
by = 16 * (m * yB + m * yN) * nf
VARIANTS = ((20, "4 x 4096, two groups, CTA-wide combine, L2 scratch (round-2 default)"),
            (18, "DIT across / DIT within, TMEM parking + swap, unit-stride stores"),
            (0, "DIT / DIT, TMEM parking + swap, group 1 stores half a line later (default)"),
            (15, "DIF across / DIT within, L2 scratch, 16-byte stores at 32-byte stride"),
            (17, "DIF across / DIT within, TMEM parking + swap, 32-byte pair stores"),
            )
for pre in (False, True):
    keep = None
    for variant, name in VARIANTS:
        core._lib.swiftly_b200_debug_sg_variant(core._plan, variant)
        for o in nmbf:
            o.zero_()
        t, ta = timeit(lambda: core.extract_columns(bfs, 4096 + 2048, offs, outs=nmbf, prewindowed=pre))
        print(f"K2 x{nf} prewindowed={int(pre)} [{variant}: {name}]: {t:.3f} ms (avg {ta:.3f})  "
              f"frac {by/t*1e3/HBM:.3f}", flush=True)
        if keep is None:
            keep = [o.clone() for o in (nmbf[0], nmbf[3], nmbf[7])]
        else:
            d = max((a - b).abs().max().item() for a, b in zip((nmbf[0], nmbf[3], nmbf[7]), keep))
            print(f"   max |diff| vs the L2-scratch kernel: {d:.3e} (max |ref| {keep[1].abs().max().item():.3e})", flush=True)
core._lib.swiftly_b200_debug_sg_variant(core._plan, 0)
