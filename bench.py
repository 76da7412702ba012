#!/usr/bin/env python
"""bench.py -- cell-steps/s of the AFV build + force + step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells-per-gpu M]

Our arm: each rank owns one periodic x-slab holding ``cells-per-gpu`` cells (default 10^6: at N=1 this is
BASELINE config[3] "N=10^6 periodic active AFV on 1xB200, dense monolayer"; weak scaling for N>1, and
``--cells-per-gpu 12500000 --gpus 8`` is config[4], N=10^8).  A step = bin + sort + geometry + force + active
Langevin update of every owned cell (plus the halo exchange / migration when sharded).  ``value`` is timed with
CUDA events on the stream the kernels run on, inputs resident in HBM; ``e2e`` goes through the public
``FiniteVoronoiSimulator`` API with pinned host buffers and host<->device copies inside the timed region.

Reference arm (``--impl reference``): the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py)
running its own periodic recipe (tile_pbc + ParallelFiniteVoronoiSimulator over all host cores + numpy Euler
step) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

# physical / integration parameters of the workload (examples/active_FV.py:11-14)
STEP = dict(dt=0.01, mu=1.0, v0=2.4, Dr=0.3)
DENSITY = 1.0          # cells per l^2 (tests/test_bench_build.py:19-20 convention: L = sqrt(N))
ALG_BYTES_PER_CELL_STEP = 100.0   # read x,y,theta,A0 + write x,y,theta (56 B) + F,A,P,L,coord (44 B)
ALG_FLOP_PER_CELL_STEP = 1.6e3    # SURVEY 8(d) lower-bound FP64 flop count, dense random, rho l^2 = 1


def measured_peaks() -> tuple[dict, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return json.loads(f.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the warm-up and the timed region run."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_slab(rank: int, world: int, cells: int, seed: int):
    """Uniform random cells of rank's slab [rank*W, (rank+1)*W) x [0, L), L = sqrt(world*cells/DENSITY)."""
    n_total = world * cells
    L = float(np.sqrt(n_total / DENSITY))
    W = L / world
    rng = np.random.default_rng(seed + 1000 * rank)
    pts = np.empty((cells, 2))
    pts[:, 0] = (rank + rng.random(cells)) * W
    pts[:, 1] = rng.random(cells) * L
    theta = 2 * np.pi * rng.random(cells) - np.pi
    return pts, theta, L


# ------------------------------------------------------------------------------------------------------
def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist
    import pyafv_b200 as afv
    from pyafv_b200 import _lib

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the AFV path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner out of the one-JSON-line stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    cells = args.cells_per_gpu
    pts, theta, L = make_slab(rank, world, cells, seed=42)
    phys = afv.PhysicalParams()
    # a dedicated stream: the handle launches on it, the CUDA events are recorded on it, NCCL orders against it
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        sim = afv.FiniteVoronoiSimulator(pts, phys, periodic=L, device=local, stream=stream.cuda_stream)
        sim.step(0.0, theta=theta, n_steps=0)
        h = sim._h
        def one_step(i):
            sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=1)
    else:
        from pyafv_b200.sharded import SlabShardedSimulator
        # the late-cell overlap pays while the exchange is a visible share of the step: measured +7 % at 10^6 cells/GPU on
        # 8 GPUs, -4 % at 1.25x10^7 cells/GPU
        overlap = (not args.no_overlap) and (args.overlap or cells <= 4_000_000)
        sim = SlabShardedSimulator(pts, theta, phys, L, rank=rank, world=world, device=local, stream=stream.cuda_stream,
                                   overlap=overlap)
        h = sim._h
        def one_step(i):
            sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        one_step(i)
    # ---- device-timed region ----
    h.call("afv_set_profiling", 1)
    ms4, l4 = np.zeros(6), np.zeros(6, dtype=np.int64)
    h.call("afv_get_stage_times", ms4.ctypes.data, l4.ctypes.data, 1)
    launches0 = int(h.lib.afv_launch_count(h._h))
    ov0 = getattr(sim, "overlapped_steps", 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if world == 1:
        # the K timed steps are issued by one call: one host synchronisation at the end instead of one per step
        sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=args.steps)
    else:
        for i in range(args.steps):
            one_step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(h.lib.afv_launch_count(h._h)) - launches0
    ov_steps = getattr(sim, "overlapped_steps", 0) - ov0
    if world > 1 and rank == 0:
        print(f"[bench] overlapped steps in the timed region: {ov_steps} of {args.steps}; {sim.exchange_stats()}", file=sys.stderr)
    h.call("afv_get_stage_times", ms4.ctypes.data, l4.ctypes.data, 1)
    h.call("afv_set_profiling", 0)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nown = torch.tensor([float(sim.n_owned)], device="cuda", dtype=torch.float64)
        dist.all_reduce(nown, op=dist.ReduceOp.SUM)
        total_cells = int(nown.item())
    else:
        total_cells = cells
    value = total_cells * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (rank 0's stage timers; events on the launching stream) ----
    peaks, peak_src = measured_peaks()
    stage_names = ["bin+sort+reorder+bin_start", "k_hull", "k_ring", "k_force(+update)", "export", "k_geometry_slow"]
    dom = int(np.argmax(ms4[1:4])) + 1
    dom_ms = float(ms4[dom] / max(int(l4[dom]), 1))
    n_local = int(h.n)
    fp64 = _lib.C.c_double()
    h.call("afv_measure_fp64_peak", _lib.C.byref(fp64))
    ach_gbs = ALG_BYTES_PER_CELL_STEP * n_local / (dom_ms * 1e-3) / 1e9
    ach_tf = ALG_FLOP_PER_CELL_STEP * n_local / (dom_ms * 1e-3) / 1e12
    roofline = {"kernel": stage_names[dom], "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach_gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_cell_step": ALG_BYTES_PER_CELL_STEP, "kernel_ms": dom_ms,
                "whole_step": {"achieved": ALG_BYTES_PER_CELL_STEP * total_cells / world / (ms / args.steps * 1e-3) / 1e9,
                               "frac": ALG_BYTES_PER_CELL_STEP * total_cells / world / (ms / args.steps * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "stage_ms_per_step": {nm: float(ms4[i] / max(args.steps, 1)) for i, nm in enumerate(stage_names)}}
    roofline_fp64 = {"kernel": stage_names[dom], "bound": "fp64", "achieved": ach_tf, "peak": float(fp64.value),
                     "unit": "TFLOP/s", "frac": ach_tf / float(fp64.value) if fp64.value else None,
                     "alg_flop_per_cell_step": ALG_FLOP_PER_CELL_STEP, "peak_source": "DFMA micro-benchmark, this run",
                     "whole_step": {"achieved": ALG_FLOP_PER_CELL_STEP * total_cells / world / (ms / args.steps * 1e-3) / 1e12,
                                    "frac": (ALG_FLOP_PER_CELL_STEP * total_cells / world / (ms / args.steps * 1e-3) / 1e12 / float(fp64.value)) if fp64.value else None}}
    traffic_file = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
            if tr.get("cells") == n_local:
                roofline["traffic"] = tr.get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- end to end through the public API with HOST buffers (pinned): every step uploads the positions, advances one
    #      fused step on the device and reads the new positions back; the literal reference loop (build + host Euler
    #      update, examples/active_FV.py:43-54) is timed as well ----
    e2e = None
    if world == 1:
        pin = _lib.PinnedArray((cells, 2))
        pin.array[:] = sim.pts
        k_e2e = max(3, min(args.steps, 10))
        for _ in range(2):
            sim.update_positions(pin.array)
            sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=1)
            sim.positions_into(pin.array)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            sim.update_positions(pin.array)                                              # H2D 16 B / cell
            sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=1)  # bin, sort, geometry, force, update
            sim.positions_into(pin.array)                                                # D2H 16 B / cell
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        e2e = {"value": cells * k_e2e / dt_e2e, "unit": "cell-steps/s", "h2d_bytes_per_step": 16 * cells,
               "d2h_bytes_per_step": 16 * cells, "steps": k_e2e,
               "api": "FiniteVoronoiSimulator.update_positions(pinned) + step(n_steps=1) + positions_into(pinned)"}
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            sim.update_positions(pin.array)                 # H2D 16 B / cell
            d = sim.build(connect=False, rings=False)       # D2H forces, areas, perimeters, arclens, coord_nums (56 B / cell)
            pin.array += STEP["mu"] * d["forces"] * STEP["dt"]
            np.mod(pin.array, L, out=pin.array)
        torch.cuda.synchronize()
        dt_ref_loop = time.perf_counter() - t0
        e2e["reference_style_loop"] = {"value": cells * k_e2e / dt_ref_loop, "unit": "cell-steps/s", "h2d_bytes_per_step": 16 * cells,
                                       "d2h_bytes_per_step": 56 * cells,
                                       "api": "update_positions + build(connect=False, rings=False) + numpy Euler update on the host"}
        pin.free()
    else:
        e2e = sim.e2e_measure(max(3, min(args.steps, 10)), STEP)

    out = None
    if rank == 0:
        cpu = cpu_baseline_sample()
        out = {"metric": "cell-steps/s (build+force+step)", "value": value, "unit": "cell-steps/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"N={total_cells} periodic active AFV, uniform random rho*l^2=1 "
                                      f"(BASELINE config[{3 if world == 1 else 4}] shape), {cells} cells/GPU, x-slabs",
                          "cells_per_gpu": cells, "box_L": L, **STEP,
                          "exchange": ("none (1 GPU)" if world == 1 else "per step, NCCL point-to-point with both ring neighbours, " +
                                       ("overlapped with the sort and the interior geometry of the step (late cells)" if getattr(sim, "overlap", False) else "then step")),
                          "exchange_overlapped_steps": ov_steps,
                          "l2_policy": "per-step working set (ring records "
                          f"{n_local * 1024 / 1e6:.0f} MB + state) exceeds the 126 MB L2; no explicit flush"},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_fp64": roofline_fp64,
               "cpu_baseline": cpu}
        _emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------------
def reference_step_time(n_cells: int, workers: int, steps: int, warmup: int):
    """Seconds per (tile_pbc + parallel build + Euler) step of the unmodified reference on n_cells cells."""
    import build_ref
    afv = build_ref.import_reference()
    L = float(np.sqrt(n_cells / DENSITY))
    rng = np.random.default_rng(42)
    pts = rng.random((n_cells, 2)) * L
    theta = 2 * np.pi * rng.random(n_cells) - np.pi
    phys = afv.PhysicalParams()
    gx = int(np.ceil(np.sqrt(workers)))
    while workers % gx:
        gx += 1
    grid = (gx, workers // gx)

    def step(sim):
        nonlocal pts, theta
        tiled, idx = afv.tile_pbc(pts, L, phys.r)
        sim.update_positions(tiled)
        F = sim.build(connect=False)["forces"][:n_cells]
        pts = np.mod(pts + (STEP["mu"] * F + STEP["v0"] * np.column_stack((np.cos(theta), np.sin(theta)))) * STEP["dt"], L)
        theta = theta + np.sqrt(2 * STEP["Dr"] * STEP["dt"]) * rng.standard_normal(n_cells)

    tiled, _ = afv.tile_pbc(pts, L, phys.r)
    if workers > 1:
        sim = afv.ParallelFiniteVoronoiSimulator(tiled, phys, grid_shape=grid, n_workers=workers, decomposition_method="binned")
        with sim:
            for _ in range(warmup):
                step(sim)
            t0 = time.perf_counter()
            for _ in range(steps):
                step(sim)
            dt = time.perf_counter() - t0
    else:
        sim = afv.FiniteVoronoiSimulator(tiled, phys)
        for _ in range(warmup):
            step(sim)
        t0 = time.perf_counter()
        for _ in range(steps):
            step(sim)
        dt = time.perf_counter() - t0
    return dt / steps, grid


def cpu_baseline_sample() -> dict:
    """Serial Cython reference (1 core) on a bounded sample of the workload: ~10-20 s of CPU work."""
    try:
        n = 100_000
        t, _ = reference_step_time(n, 1, steps=1, warmup=1)
        return {"value": n / t, "unit": "cell-steps/s", "cores": 1, "kind": "reference",
                "sample": f"unmodified reference v0.4.16 (Cython), tile_pbc + FiniteVoronoiSimulator.build(connect=False) + numpy "
                          f"Euler step, N={n} periodic dense sample of the workload, 1 warm-up + 1 timed step"}
    except Exception as e:  # reference binaries missing on this box
        try:
            from afv_oracle import oracle_build
            n = 300_000
            L = float(np.sqrt(n))
            pts = np.random.default_rng(42).random((n, 2)) * L
            oracle_build(pts[:1000] * 0 + pts[:1000], periodic=None)
            t0 = time.perf_counter()
            oracle_build(pts, periodic=L)
            t = time.perf_counter() - t0
            return {"value": n / t, "unit": "cell-steps/s", "cores": 1, "kind": "port",
                    "sample": f"oracle/afv_oracle.c build on N={n} periodic dense (reference binaries unavailable: {e})"}
        except Exception as e2:
            return {"value": None, "unit": "cell-steps/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e2}"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    n = args.reference_cells
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    # keep the whole run within a few minutes: the serial reference does ~1.5e4 cells/s/core
    est = n / (1.2e4 * max(1, min(workers, 16)) * 0.6) * (steps + warmup)
    if est > 240:
        steps = max(1, int(steps * 240 / est))
        warmup = 1
    t, grid = reference_step_time(n, workers, steps, warmup)
    value = n / t
    world = int(os.environ.get("WORLD_SIZE", 1))
    cells = args.cells_per_gpu
    out = {"impl": "reference", "metric": "cell-steps/s (build+force+step)", "value": value, "unit": "cell-steps/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"N={world * cells} periodic active AFV, uniform random rho*l^2=1 "
                                  f"(BASELINE config[{3 if world == 1 else 4}] shape), {cells} cells/GPU, x-slabs",
                      "cells_per_gpu": cells, **STEP},
           "cpu_baseline": {"value": value, "unit": "cell-steps/s", "cores": workers, "kind": "reference",
                            "sample": f"unmodified reference v0.4.16: tile_pbc + ParallelFiniteVoronoiSimulator(grid={grid}, "
                                      f"n_workers={workers}, binned, persistent pool).build(connect=False) + numpy Euler step on "
                                      f"an N={n} periodic dense sample of the workload"},
           "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(json.dumps(out))


def _emit(line: str) -> None:
    """The one JSON line goes to the real stdout; everything else any library prints was diverted to stderr."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)       # NCCL / torch banners must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--reference-cells", type=int, default=200_000)
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange, then step (default up to 4x10^6 cells/GPU: the exchange of "
                    "a step travels while the step sorts its cells and builds its interior geometry -- late cells, DESIGN.md section 5)")
    ap.add_argument("--overlap", action="store_true", help="N > 1: the late-cell overlap at any size")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
