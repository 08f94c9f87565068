#!/usr/bin/env python
"""
bench.py -- facet->subgrid contributions per second of the B200 SwiFTly hot path.

    python bench.py --gpus N --steps K --warmup W            (this repo, CUDA)
    python bench.py --impl reference --gpus N ...            (reference algorithm on host cores)

One "step" is one COMPLETE forward transform (stages 1-6 of SURVEY.md section 3.2: prepare all
facets, every subgrid column, every subgrid, masks; at N > 1 including the strip exchange)
of the workload, default BASELINE cfg4: N=65536, 8x8 facets of 8192^2 -> 32x32 subgrids of
2048^2, complex128, dense synthetic facets.  metric = (#facets x #subgrids) / step time.

Timing: every step is bracketed by barrier + synchronize and timed with CUDA events on the
compute stream; the step time is the max over ranks; ms_per_step is the mean over the K
timed steps.  Stage 1 consumes the facets (on one GPU the prepared facets reuse their
storage: 64 GiB + 128 GiB do not fit 180 GB otherwise), so facets are regenerated on the
device between steps, outside the timed region.  All inputs are far larger than L2.
Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (swift config key, description)
    "cfg1": "1k[1]-n512-256",
    "cfg2": "8k[1]-n4k-2k",
    "cfg3": "32k[1]-n8k-4k",
    "cfg4": "64k[1]-n16k-4k",
}
METRIC = "facet->subgrid contributions/sec"
UNIT = "contributions/s"


def note(msg):
    """Progress line on stderr (stdout carries only the JSON line)."""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
        sys.stderr.flush()


def workload_params(name):
    from ska_sdp_distributed_fourier_transform_b200.swift_configs import SWIFT_CONFIGS

    return dict(SWIFT_CONFIGS[WORKLOADS[name]])


def host_cores():
    """CPU cores this process may actually use: min(affinity, cgroup quota, cpu_count).

    The GPU boxes report 128 CPUs but run in a container limited to 16 CPUs / 200 GiB;
    starting 128 numpy workers there exhausts the memory limit.
    """
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def measured_traffic(kernel_name, world):
    """DRAM bytes per launch of the named kernel from the committed ncu capture (N=1 only)."""
    if world != 1:
        return None
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                table = json.load(f)
        except (OSError, ValueError):
            continue
        for key, val in table.items():
            if not key.startswith("_") and key in kernel_name:
                return val["dram_bytes_per_launch"]
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ====================================================================== clocks sampler
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks, maxc, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    clocks.append(float(parts[1]))
                    maxc.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if clocks:
            # samples under load: the busy half of the distribution
            out["sm_mhz"] = float(numpy.median(clocks))
            out["sm_max_mhz"] = float(max(maxc))
            out["samples"] = len(clocks)
        out["reasons"] = sorted(reasons)
        return out


# ====================================================================== CPU arm (oracle)
def _cpu_stage_sample(args):
    """Time the per-stage unit costs of the reference algorithm (oracle port) once.

    Runs in a worker process; returns per-unit seconds for the stage model of
    ``cpu_model_rate``.  Stage 1 is timed on a slab of ``cols1`` facet columns (lines are
    independent) and scaled to the full facet.
    """
    params, cols1, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle.swiftly_oracle import OracleCore

    p = params
    core = _cpu_stage_sample.cache.get("core")
    if core is None:
        core = OracleCore(p["W"], p["N"], p["xM_size"], p["yN_size"])
        _cpu_stage_sample.cache["core"] = core
    yB, yN, xA, xM, m = p["yB_size"], p["yN_size"], p["xA_size"], p["xM_size"], core.xM_yN_size
    rng = numpy.random.default_rng(seed)
    Nx, Ny = core.subgrid_off_step, core.facet_off_step

    def rc(*s):
        return rng.standard_normal(s) + 1j * rng.standard_normal(s)

    t = {}
    slab = rc(yB, cols1)
    t0 = time.perf_counter()
    bf = core.prepare_facet(slab, 3 * Ny, axis=0)
    t["stage1_prepare_facet_ax0"] = (time.perf_counter() - t0) * (yB / cols1)
    bf_full = rc(yN, min(yB, 4 * cols1))  # stands for BF_F; width scales stage 2a / 2b
    scale2 = yB / bf_full.shape[1]
    t0 = time.perf_counter()
    rows = core.extract_from_facet(bf_full, 5 * Nx, axis=0)
    t["stage2a_extract_ax0"] = (time.perf_counter() - t0) * scale2
    rows_full = rc(m, yB)
    t0 = time.perf_counter()
    nmbf = core.prepare_facet(rows_full, -2 * Ny, axis=1)
    t["stage2b_prepare_facet_ax1"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    contrib = core.extract_from_facet(nmbf, 7 * Nx, axis=1)
    t["stage3_extract_ax1"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    col = core.add_to_subgrid(contrib, 3 * Ny, axis=0)
    t["stage4_add_to_subgrid_ax0"] = time.perf_counter() - t0
    acc = numpy.zeros((xM, xM), dtype=complex)
    t0 = time.perf_counter()
    acc = core.add_to_subgrid(col, -2 * Ny, axis=1, out=acc)
    t["stage5_add_to_subgrid_ax1"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    core.finish_subgrid(acc, [5 * Nx, 7 * Nx], xA)
    t["stage6_finish_subgrid"] = time.perf_counter() - t0
    del bf, rows
    return t


_cpu_stage_sample.cache = {}


def cpu_model_rate(params, unit_times, cores):
    """contributions/s of the reference task graph on ``cores`` independent workers.

    T_full = F t1 + ns F (t2a + t2b) + S F (t3 + t4) + S nf1 t5 + S t6   (SURVEY.md 3.2),
    every task single threaded (numpy FFT), tasks embarrassingly parallel across workers
    like the reference's Dask graph; unit times measured with all workers busy.
    """
    N, yB, xA = params["N"], params["yB_size"], params["xA_size"]
    nf1 = -(-N // yB)
    F = nf1 * nf1
    ns = -(-N // xA)
    S = ns * ns
    u = unit_times
    total = (F * u["stage1_prepare_facet_ax0"]
             + ns * F * (u["stage2a_extract_ax0"] + u["stage2b_prepare_facet_ax1"])
             + S * F * (u["stage3_extract_ax1"] + u["stage4_add_to_subgrid_ax0"])
             + S * nf1 * u["stage5_add_to_subgrid_ax1"]
             + S * u["stage6_finish_subgrid"])
    return F * S / (total / cores), total


def run_cpu_sample(params, cores, repeats=1, cols1=None):
    """Run the stage sample on ``cores`` worker processes simultaneously."""
    import multiprocessing as mp

    if cols1 is None:
        cols1 = max(16, min(params["yB_size"], (1 << 22) // params["yN_size"]))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        results = []
        for rep in range(repeats):
            results += pool.map(_cpu_stage_sample, [(params, cols1, 1000 * rep + i) for i in range(cores)])
    wall = time.perf_counter() - t0
    keys = results[0].keys()
    unit = {k: float(numpy.mean([r[k] for r in results])) for k in keys}
    return unit, wall, cols1


def cpu_baseline_entry(params, cores, repeats=1):
    unit, wall, cols1 = run_cpu_sample(params, cores, repeats)
    rate, total = cpu_model_rate(params, unit, cores)
    sample = (f"oracle (numpy port of the reference SwiftlyCore + task order) on {cores} worker "
              f"processes, each timing every stage unit once per repeat (stage 1 on a {cols1}-column "
              f"slab scaled to the facet); rate from the reference task-count model; "
              f"sample wall {wall:.1f} s")
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "model_full_transform_core_seconds": total, "stage_unit_seconds": unit}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = workload_params(args.workload)
    cores = args.cpu_cores or host_cores()
    for _ in range(max(0, min(args.warmup, 1))):
        run_cpu_sample(params, cores, 1)
    t0 = time.perf_counter()
    entry = cpu_baseline_entry(params, cores, repeats=max(1, args.steps))
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": entry["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)",
        "data": "synthetic", "config": workload_config(args.workload, params, args.gpus),
        "cpu_baseline": entry,
        "e2e": {"value": entry["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(name, params, gpus):
    N, yB, xA = params["N"], params["yB_size"], params["xA_size"]
    nf = -(-N // yB)
    ns = -(-N // xA)
    return {
        "workload": f"{name}: 2D N={N}, {nf}x{nf} facets of {yB}^2 -> {ns}x{ns} subgrids of "
                    f"{xA}^2, yN={params['yN_size']}, xM={params['xM_size']}, W={params['W']}, "
                    f"complex128, dense standard-normal facets",
        "contributions_per_step": nf * nf * ns * ns,
        "parallelism": "1 GPU" if gpus == 1 else f"facet rows sharded over {gpus} GPUs, "
                       "strips exchanged per subgrid batch",
        "l2": "inputs (>= 64 GiB/step) far larger than the 126 MB L2; no flush needed",
        "timing": "per-step CUDA events, facets regenerated on device between steps (untimed)",
    }


# ====================================================================== GPU arm
def main_gpu(args):
    # Everything but the final JSON line goes to stderr -- also what libraries print on the C
    # level (NCCL writes its version banner to stdout when the first communicator is created).
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _main_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line))
        sys.stdout.flush()


def _main_gpu(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner on STDOUT (NCCL_DEBUG=VERSION in this image); stdout
        # must carry the JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    from ska_sdp_distributed_fourier_transform_b200 import bench_support as bs

    params = workload_params(args.workload)
    note(f"setting up {args.workload} on {world} GPU(s)")
    runner = bs.ForwardBenchRunner(params, dev, rank, world, exchange=args.exchange)
    hbm_gbs, peak_src = measured_peaks()

    # ---- device-resident runs (value) ----------------------------------------------
    for i in range(args.warmup):
        runner.step(timed=False)
        note(f"warm-up step {i + 1}/{args.warmup} done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times = []
    for i in range(args.steps):
        times.append(runner.step(timed=True))
        note(f"timed step {i + 1}/{args.steps}: {times[-1]:.1f} ms")
    clocks = sampler.stop() if rank == 0 else None
    ms = float(numpy.mean(times))
    contributions = runner.contributions_per_step
    value = contributions / (ms * 1e-3)

    parity = None
    if not args.no_selfcheck:
        parity = runner.selfcheck()  # raises above 1e-9; collective at N > 1
        note(f"self-check: max rel err {parity['parity_max_rel_err']:.2e} "
             f"({parity['subgrids_checked']} subgrids)")
    extra = {}
    if rank == 0 or world > 1:
        extra = runner.kernel_rooflines(hbm_gbs, step_ms=ms) if not args.no_roofline else {}
    note("kernel rooflines done")
    e2e = None
    if not args.no_e2e:
        e2e = runner.e2e(steps=args.e2e_steps, progress=note)
        note(f"e2e done: {e2e['ms_per_step']:.1f} ms/step")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_entry(params, args.cpu_cores or host_cores())
        note("cpu baseline done")
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)",
        "data": "synthetic", "config": workload_config(args.workload, params, world),
        "step_ms": times, "gpu_launches": runner.launches_per_step,
        "clocks": clocks,
    }
    if parity is not None:
        line["parity_max_rel_err"] = parity["parity_max_rel_err"]
        line["parity_check"] = parity
    if world > 1:
        line["config"]["exchange"] = runner.exchange_used
    if extra:
        line["roofline"] = extra["dominant"]
        line["roofline"]["peak_source"] = peak_src
        if args.workload == "cfg4":
            line["roofline"]["traffic"] = measured_traffic(line["roofline"]["kernel"], world)
        line["kernel_rooflines"] = extra["kernels"]
        line["bmin_roofline"] = extra["bmin"]
    if e2e is not None:
        line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = cpu
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true",
                    help="skip the output parity check that follows the timed steps")
    ap.add_argument("--selfcheck", action="store_true", help="(default) kept for symmetry")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU strip exchange: peer-memory stores or NCCL all_to_all")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
