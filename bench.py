#!/usr/bin/env python
"""bench.py -- cell-steps/s of the AFV build + force + step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells-per-gpu M]

Our arm.  Each rank owns one periodic x-slab holding ``cells-per-gpu`` cells (default 10^6: at N=1 this is BASELINE
config[3] "N=10^6 periodic active AFV on 1xB200, dense monolayer"; weak scaling for N>1).  A step = bin + sort +
neighbour lists + clipping/ring/measures + force gather + active Langevin update of every owned cell (plus the halo
exchange / migration when sharded).  ``value`` is timed with CUDA events on the stream the kernels run on, inputs
resident in HBM, the library's per-stage timers off (they put two event records around every kernel: 0.03 ms per step at
N = 10^6); the stage times behind ``roofline`` come from a repeat of the same K steps with the timers on; ``e2e`` goes through the public ``FiniteVoronoiSimulator`` API with pinned host buffers and
host<->device copies inside the timed region.  The same JSON line also carries

* ``config4_n1e8``: BASELINE config[4], N = 10^8 cells in total on the N GPUs of this run (10^8/N per GPU,
  exchange-then-step), its own ``value`` / ``ms_per_step`` / ``roofline``;
* ``regimes`` (N=1): the jammed tissue (hex lattice + jitter, per-cell A0) and the fragmenting rho l^2 = 0.1 regime at
  N = 10^6;
* ``verify`` (N>1): a 2x10^5-cell system stepped 30 times both sharded over the real ranks and as one handle on
  rank 0 (over the transport the run uses: peer memory or NCCL) -- the trajectories must be bit-identical -- plus cell conservation, zero ring look-up misses and zero net force.

Reference arm (``--impl reference``): the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py) running its
own periodic recipe (tile_pbc + ParallelFiniteVoronoiSimulator over all host cores + numpy Euler step) on the N = 10^6
workload of the 1-GPU config (fewer timed steps than asked when they would not fit a few minutes; the line says so).
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
GRAPH_SETTLE = 8     # untimed steps in one call before a single-GPU timing: afv_step captures its loop as a CUDA graph on the first call of >= 6 steps
STAGES = ["bin+sort+reorder+bin_start", "k_nbrs", "k_clip", "k_force(+update)", "export", "k_geometry_slow"]
N_CONFIG4 = 100_000_000


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


def make_slab(rank: int, world: int, cells: int, seed: int, density: float = DENSITY):
    """Uniform random cells of rank's slab [rank*W, (rank+1)*W) x [0, L), L = sqrt(world*cells/density)."""
    n_total = world * cells
    L = float(np.sqrt(n_total / density))
    W = L / world
    rng = np.random.default_rng(seed + 1000 * rank)
    pts = np.empty((cells, 2))
    pts[:, 0] = np.minimum((rank + rng.random(cells)) * W, np.nextafter((rank + 1) * W, 0.0))
    pts[:, 1] = rng.random(cells) * L
    theta = 2 * np.pi * rng.random(cells) - np.pi
    return pts, theta, L


def jammed_tissue(n_target: int, seed: int = 3):
    """SURVEY 8(d) C3 at size n_target: hex lattice (rho l^2 = 0.5, a = 1.5197 l) in an exactly square periodic box,
    Gaussian jitter 0.05 l, per-cell A0 = 2.0 (1 + 0.1 N(0,1)) clipped to >= 0.4."""
    a = 1.5197
    nx = int(round(np.sqrt(n_target * np.sqrt(3) / 2)))
    ny = int(round(nx * 2 / np.sqrt(3) / 2)) * 2          # even, so that the rows close periodically
    L = nx * a
    sy = L / (ny * a * np.sqrt(3) / 2)                      # rescale y so that the box is exactly square
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    pts = np.column_stack(((ix + 0.5 * (iy % 2)).ravel() * a, iy.ravel() * a * np.sqrt(3) / 2 * sy))
    rng = np.random.default_rng(seed)
    pts = np.mod(pts + 0.05 * rng.standard_normal(pts.shape), L)
    A0 = np.maximum(2.0 * (1 + 0.1 * rng.standard_normal(len(pts))), 0.4)
    return pts, A0, float(L)


class Timed:
    """K device-timed steps of a simulator (CUDA events on its stream, barrier + synchronize on both sides, max over ranks)."""

    def __init__(self, torch, dist, world, stream):
        self.torch, self.dist, self.world, self.stream = torch, dist, world, stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, h, step_k, steps: int, stages: bool = True):
        """stages: per-stage CUDA-event timers on (two event records around every kernel: they cost ~0.03 ms per step at
        N = 10^6, so the headline `value` is timed with stages=False and the stage times come from the repeat after it)."""
        torch = self.torch
        h.call("afv_set_profiling", 1 if stages else 0)
        ms6, l6 = np.zeros(6), np.zeros(6, dtype=np.int64)
        h.call("afv_get_stage_times", ms6.ctypes.data, l6.ctypes.data, 1)
        launches0 = int(h.lib.afv_launch_count(h._h))
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        step_k(steps)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = int(h.lib.afv_launch_count(h._h)) - launches0
        h.call("afv_get_stage_times", ms6.ctypes.data, l6.ctypes.data, 1)
        h.call("afv_set_profiling", 0)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ms6, launches


def rooflines(ms6, steps: int, n_local: int, cells_per_rank_step: float, ms_step: float, fp64_peak: float):
    """Roofline of the dominant kernel (stage time / steps: a stage may be launched more than once per step) and of the
    whole step, against the measured HBM copy bandwidth and the DFMA peak measured in this run."""
    peaks, peak_src = measured_peaks()
    dom = int(np.argmax(ms6[1:4])) + 1
    dom_ms = float(ms6[dom] / max(steps, 1))
    ach_gbs = ALG_BYTES_PER_CELL_STEP * n_local / (dom_ms * 1e-3) / 1e9
    ach_tf = ALG_FLOP_PER_CELL_STEP * n_local / (dom_ms * 1e-3) / 1e12
    ws_gbs = ALG_BYTES_PER_CELL_STEP * cells_per_rank_step / (ms_step * 1e-3) / 1e9
    ws_tf = ALG_FLOP_PER_CELL_STEP * cells_per_rank_step / (ms_step * 1e-3) / 1e12
    rl = {"kernel": STAGES[dom], "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
          "frac": ach_gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
          "alg_bytes_per_cell_step": ALG_BYTES_PER_CELL_STEP, "kernel_ms": dom_ms,
          "whole_step": {"achieved": ws_gbs, "frac": ws_gbs / peaks["hbm_gbs"]},
          "stage_ms_per_step": {nm: float(ms6[i] / max(steps, 1)) for i, nm in enumerate(STAGES)}}
    rl64 = {"kernel": STAGES[dom], "bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ach_tf / fp64_peak if fp64_peak else None, "alg_flop_per_cell_step": ALG_FLOP_PER_CELL_STEP,
            "peak_source": "DFMA micro-benchmark, this run",
            "whole_step": {"achieved": ws_tf, "frac": ws_tf / fp64_peak if fp64_peak else None}}
    return rl, rl64


# ------------------------------------------------------------------------------------------------------
def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist
    import pyafv_b200 as afv
    from pyafv_b200 import _lib
    from pyafv_b200.sharded import SlabShardedSimulator

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
    phys = afv.PhysicalParams()
    # a dedicated stream: the handles launch on it, the CUDA events are recorded on it, NCCL orders against it
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    timed = Timed(torch, dist, world, stream)

    def make_sim(cells_, overlap, density=DENSITY, seed=42):
        pts, theta, L = make_slab(rank, world, cells_, seed=seed, density=density)
        if world == 1:
            sim = afv.FiniteVoronoiSimulator(pts, phys, periodic=L, device=local, stream=stream.cuda_stream)
            sim.step(0.0, theta=theta, n_steps=0)
            step_k = lambda k: sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=k)   # one call, one host sync
        else:
            sim = SlabShardedSimulator(pts, theta, phys, L, rank=rank, world=world, device=local, stream=stream.cuda_stream,
                                       overlap=overlap)
            step_k = lambda k: sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=7, n_steps=k)
        return sim, step_k, L

    def owned_total(sim, cells_):
        if world == 1:
            return cells_
        nown = torch.tensor([float(sim.n_owned)], device="cuda", dtype=torch.float64)
        dist.all_reduce(nown, op=dist.ReduceOp.SUM)
        return int(nown.item())

    # ---- verification across the real ranks (N > 1), before anything is timed ----
    verify = verify_sharded(torch, dist, afv, SlabShardedSimulator, phys, rank, world, local, stream) if world > 1 else None

    # ---- the headline workload: `cells` per GPU ----
    # exchange-then-step by default: with this round's kernels the late-cell overlap (face pass, late sort, event waits) costs
    # more than the ~0.1 ms of exchange it hides (measured on 2 GPUs: 0.947 against 0.993 ms/step)
    overlap = world > 1 and args.overlap and not args.no_overlap
    sim, step_k, L = make_sim(cells, overlap)
    transport = transport_name(sim)
    h = sim._h
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_k(args.warmup)
    if world == 1:
        step_k(GRAPH_SETTLE)     # untimed, on top of the W warm-up steps: the first long call captures the time loop as a CUDA graph
    ov0 = getattr(sim, "overlapped_steps", 0)
    ms, _, launches = timed.run(h, step_k, args.steps, stages=False)
    clocks = sampler.stop() if rank == 0 else None
    ov_steps = getattr(sim, "overlapped_steps", 0) - ov0
    if world > 1 and rank == 0:
        print(f"[bench] overlapped steps in the timed region: {ov_steps} of {args.steps}; {sim.exchange_stats()}", file=sys.stderr)
    total_cells = owned_total(sim, cells)
    if world > 1:
        assert total_cells == world * cells, f"cells lost or duplicated by the exchange: {total_cells} != {world * cells}"
    value = total_cells * args.steps / (ms * 1e-3)
    # spread: three more repeats of the same K steps (not part of `value`)
    rep_ms = []
    for _ in range(3):
        m_, _, _ = timed.run(h, step_k, args.steps, stages=False)
        rep_ms.append(m_ / args.steps)
    # the same K steps once more with the per-stage event timers on: the kernels' durations for the rooflines
    ms_staged, ms6, _ = timed.run(h, step_k, args.steps, stages=True)
    n_local = int(h.n)
    fp64 = _lib.C.c_double()
    h.call("afv_measure_fp64_peak", _lib.C.byref(fp64))
    roofline, roofline_fp64 = rooflines(ms6, args.steps, n_local, total_cells / world, ms / args.steps, float(fp64.value))
    traffic_file = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
            if tr.get("cells") == n_local and tr.get("kernel") == roofline["kernel"]:
                roofline["traffic"] = tr.get("dram_bytes_per_launch")
                # what actually binds the kernel: warp instructions issued (counted by ncu for this launch shape) per second
                # of the live launch time, against 4 schedulers x 1 instruction per clock per SM
                wi = (tr.get("warp_instructions_per_launch") or {}).get(roofline["kernel"])
                if wi:
                    props = torch.cuda.get_device_properties(local)
                    peak = 4.0 * props.multi_processor_count * props.clock_rate * 1e3 / 1e9     # Gwarp-inst/s at the boost clock
                    ach = wi / (roofline["kernel_ms"] * 1e-3) / 1e9
                    roofline["issue_slots"] = {"achieved": ach, "peak": peak, "unit": "Gwarp-inst/s", "frac": ach / peak,
                                               "warp_instructions_per_launch": wi,
                                               "under_ncu_pct_of_peak": (tr.get("under_ncu_pct_of_peak") or {}).get(roofline["kernel"])}
        except Exception:
            pass

    # ---- end to end through the public API with HOST buffers (pinned): every step uploads the positions, advances one
    #      fused step on the device and reads the new positions back; the literal reference loop (build + host Euler
    #      update, examples/active_FV.py:43-54) is timed as well ----
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
    diag = np.zeros(4, dtype=np.int64)
    h.call("afv_get_diagnostics", diag.ctypes.data)
    overlap_used = bool(getattr(sim, "overlap", False))
    sim.close()
    del sim

    # ---- the other regimes of BASELINE config[3] (N = 1): jammed tissue with per-cell A0, fragmenting rho l^2 = 0.1 ----
    regimes = None
    if world == 1 and not args.skip_regimes:
        regimes = {}
        pts, A0, Lj = jammed_tissue(1_000_000)
        sj = afv.FiniteVoronoiSimulator(pts, phys, periodic=Lj, device=local, stream=stream.cuda_stream)
        sj.update_preferred_areas(A0)
        relax = lambda k: sj.step(STEP["dt"], STEP["mu"], 0.0, 0.0, n_steps=k)
        relax(300)
        k_ = max(args.steps, 10)
        m_, _, _ = timed.run(sj._h, relax, k_, stages=False)
        _, s6, _ = timed.run(sj._h, relax, k_, stages=True)
        regimes["jammed"] = {"value": len(pts) * k_ / (m_ * 1e-3), "ms_per_step": m_ / k_, "cells": len(pts),
                             "workload": "hex lattice rho*l^2=0.5 + jitter 0.05 l, per-cell A0 = 2(1+0.1 N(0,1)), 300 relaxation steps, "
                                         "then timed passive relaxation steps (v0 = 0)",
                             "stage_ms_per_step": {nm: float(s6[i] / k_) for i, nm in enumerate(STAGES)},
                             "diagnostics": sj.diagnostics()}
        sj.close()
        sf, stepf, _ = make_sim(1_000_000, False, density=0.1, seed=43)
        stepf(args.warmup)
        stepf(GRAPH_SETTLE)
        m_, _, _ = timed.run(sf._h, stepf, k_, stages=False)
        _, s6, _ = timed.run(sf._h, stepf, k_, stages=True)
        regimes["rho0.1"] = {"value": 1_000_000 * k_ / (m_ * 1e-3), "ms_per_step": m_ / k_, "cells": 1_000_000,
                             "workload": "uniform random, rho*l^2=0.1 (fragmenting regime, 28 % isolated cells), active",
                             "stage_ms_per_step": {nm: float(s6[i] / k_) for i, nm in enumerate(STAGES)},
                             "diagnostics": sf.diagnostics()}
        sf.close()

    # ---- BASELINE config[4]: N = 10^8 in total on the GPUs of this run, exchange-then-step ----
    config4 = None
    if not args.skip_config4:
        c4 = N_CONFIG4 // world
        try:
            s4, step4, L4 = make_sim(c4, False, seed=42 + N_CONFIG4)
            k4 = max(10, min(args.steps, 20))
            step4(3)
            if world == 1:
                step4(GRAPH_SETTLE)
            m4, _, _ = timed.run(s4._h, step4, k4, stages=False)
            _, s6, _ = timed.run(s4._h, step4, k4, stages=True)
            tot4 = owned_total(s4, c4)
            if world > 1:
                assert tot4 == world * c4, f"cells lost or duplicated by the exchange: {tot4} != {world * c4}"
            rl4, rl4_64 = rooflines(s6, k4, int(s4._h.n), tot4 / world, m4 / k4, float(fp64.value))
            config4 = {"value": tot4 * k4 / (m4 * 1e-3), "unit": "cell-steps/s", "ms_per_step": m4 / k4, "steps": k4, "warmup": 3,
                       "cells_total": tot4, "cells_per_gpu": c4, "box_L": L4, "exchange": "none (1 GPU)" if world == 1 else f"exchange over {transport_name(s4)}, then step",
                       "roofline": rl4, "roofline_fp64": rl4_64}
            s4.close()
        except Exception as e:   # e.g. not enough memory on a smaller device: say so instead of dying
            config4 = {"value": None, "error": repr(e)[:300], "cells_per_gpu": c4}

    out = None
    if rank == 0:
        cpu = cpu_baseline_sample()
        out = {"metric": "cell-steps/s (build+force+step)", "value": value, "unit": "cell-steps/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"N={total_cells} periodic active AFV, uniform random rho*l^2=1 "
                                      f"(BASELINE config[{3 if world == 1 else 4}] shape), {cells} cells/GPU" + (", x-slabs" if world > 1 else ""),
                          "cells_per_gpu": cells, "box_L": L, **STEP,
                          "exchange": ("none (1 GPU)" if world == 1 else f"per step, {transport} with both ring neighbours, " +
                                       ("overlapped with the sort and the interior geometry of the step (late cells)" if overlap_used else "then step")),
                          "exchange_overlapped_steps": ov_steps,
                          "l2_policy": f"per-step working set (~{n_local * 0.45 / 1e3:.0f} MB of rings, neighbour lists and state) exceeds the 126 MB L2; "
                                       "no explicit flush"},
               "ms_per_step_repeats": rep_ms,
               "untimed_steps_before_timing": args.warmup + (GRAPH_SETTLE if world == 1 else 0),
               "stage_timing": {"ms_per_step_with_stage_timers": ms_staged / args.steps,
                                "note": "`value`, `ms_per_step` and the repeats are timed without the per-stage CUDA-event timers (two event records around "
                                        "every kernel); roofline.stage_ms_per_step / kernel_ms come from the repeat after them, on the same stream, timers on"},
               "diagnostics": {"slow_path_cells": int(diag[0]), "lookup_misses": int(diag[1]), "max_ring_len": int(diag[2])},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_fp64": roofline_fp64,
               "config4_n1e8": config4, "regimes": regimes, "verify": verify, "cpu_baseline": cpu}
        _emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def transport_name(sim) -> str:
    """How the ranks exchange their messages (pyafv_b200/sharded.py: NvlinkPeerComm or TorchDistComm)."""
    name = type(getattr(sim, "comm", None)).__name__
    return {"NvlinkPeerComm": "peer memory over NVLink (the pack kernel writes into the neighbours' inboxes, flag words for the hand-over)",
            "TorchDistComm": "NCCL point-to-point"}.get(name, name)


def verify_sharded(torch, dist, afv, SlabShardedSimulator, phys, rank, world, local, stream, n_total=200_000, steps=30) -> dict:
    """The real multi-rank path against one handle: n_total cells stepped `steps` times (a) sharded over the ranks of this
    run, exchanging over NCCL (both step forms), and (b) as a single periodic handle on rank 0.  Positions must
    agree bit for bit (results do not depend on the number of GPUs); also checked: no cell lost or duplicated, no ring
    look-up miss, zero net force."""
    L = float(np.sqrt(n_total / DENSITY))
    rng = np.random.default_rng(2024)
    pts = rng.random((n_total, 2)) * L
    theta = 2 * np.pi * rng.random(n_total) - np.pi
    W = L / world
    owner = np.minimum((pts[:, 0] // W).astype(int), world - 1)
    mine = owner == rank
    gids = np.flatnonzero(mine)
    # half of the steps exchange-then-step (the form the bench times), half with the late-cell overlap
    sim = SlabShardedSimulator(pts[mine], theta[mine], phys, L, rank=rank, world=world, device=local, stream=stream.cuda_stream,
                               gids=gids, overlap=False)
    sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=11, n_steps=steps // 2)
    sim._refresh_halo()
    sim.overlap = True
    sim._xs = torch.cuda.Stream(sim.device, priority=-1)
    sim.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], seed=11, n_steps=steps - steps // 2)
    res = sim.build()
    diag = np.zeros(4, dtype=np.int64)
    sim._h.call("afv_get_diagnostics", diag.ctypes.data)
    # gather (gid, x, y, fx, fy) of every owned cell on rank 0
    mine_rows = torch.from_numpy(np.column_stack((res["gids"].astype(np.float64), res["pts"], res["forces"]))).to("cuda")
    cnt = torch.tensor([mine_rows.shape[0]], device="cuda", dtype=torch.int64)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    pad = torch.zeros((max(cnts), 5), device="cuda", dtype=torch.float64)
    pad[: mine_rows.shape[0]] = mine_rows
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    misses = torch.tensor([float(diag[1])], device="cuda", dtype=torch.float64)
    dist.all_reduce(misses)
    transport = transport_name(sim)
    sim.close()
    out = None
    if rank == 0:
        rows = np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, cnts)])
        order = np.argsort(rows[:, 0])
        rows = rows[order]
        conserved = len(rows) == n_total and np.array_equal(rows[:, 0].astype(np.int64), np.arange(n_total))
        ref = afv.FiniteVoronoiSimulator(pts, phys, periodic=L, device=local, stream=stream.cuda_stream)
        ref.step(STEP["dt"], STEP["mu"], STEP["v0"], STEP["Dr"], theta=theta, seed=11, n_steps=steps)
        rp = ref.pts
        rf = ref.build(connect=False, rings=False)["forces"]
        ref.close()
        if conserved:
            d = rows[:, 1:3] - rp
            d -= L * np.round(d / L)
            max_dx = float(np.abs(d).max())
            max_df = float(np.abs(rows[:, 3:5] - rf).max())
            net = float(np.abs(rows[:, 3:5].sum(axis=0)).max())
        else:
            max_dx = max_df = net = float("nan")
        fscale = float(np.abs(rf).max())
        out = {"cells": n_total, "steps": steps, "ranks": world, "transport": transport + "; half of the steps exchange-then-step, half with the late-cell overlap", "cells_conserved": bool(conserved),
               "max_dx": max_dx, "max_dF": max_df, "bit_identical": bool(conserved and max_dx == 0.0 and max_df == 0.0),
               "lookup_misses": int(misses.item()), "net_force_over_max_force": net / fscale if fscale else None,
               "ok": bool(conserved and max_dx <= 1e-9 and int(misses.item()) == 0 and net <= 1e-9 * max(1.0, fscale) * np.sqrt(n_total))}
        print(f"[bench] verify: {out}", file=sys.stderr)
    dist.barrier()
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
    world = int(os.environ.get("WORLD_SIZE", 1))
    # the reference runs on the host cores of ONE box whatever --gpus says: the N = 10^6 workload of the 1-GPU config
    # (N = 10^8 is out of its reach: ~1 h and ~600 GB per serial build); the line names the N it ran
    n = args.reference_cells if args.reference_cells else 1_000_000
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    # keep the whole run within a few minutes: ~4e5 cell-steps/s on 16 cores
    per_step = n / (2.5e4 * max(1, min(workers, 32)))
    budget = 200.0
    if per_step * (steps + warmup) > budget:
        warmup = 1
        steps = max(2, int(budget / per_step) - warmup)
    t, grid = reference_step_time(n, workers, steps, warmup)
    value = n / t
    out = {"impl": "reference", "metric": "cell-steps/s (build+force+step)", "value": value, "unit": "cell-steps/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"N={n} periodic active AFV, uniform random rho*l^2=1 (BASELINE config[3] shape), {n} cells on the host "
                                  f"cores of one box" + ("" if world == 1 else f" (the GPU arm of this run holds {world * args.cells_per_gpu} cells; "
                                                                              "the reference cannot run more than ~10^6 in minutes)"),
                      "cells_per_gpu": n, "box_L": float(np.sqrt(n / DENSITY)), **STEP},
           "cpu_baseline": {"value": value, "unit": "cell-steps/s", "cores": workers, "kind": "reference",
                            "sample": f"unmodified reference v0.4.16: tile_pbc + ParallelFiniteVoronoiSimulator(grid={grid}, "
                                      f"n_workers={workers}, binned, persistent pool).build(connect=False) + numpy Euler step on "
                                      f"the full N={n} periodic dense workload, {warmup} warm-up + {steps} timed steps"},
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
    ap.add_argument("--reference-cells", type=int, default=0, help="reference arm: cells (default: the full N = 10^6 workload)")
    ap.add_argument("--skip-config4", action="store_true", help="do not time the N = 10^8 block")
    ap.add_argument("--skip-regimes", action="store_true", help="N = 1: do not time the jammed / rho l^2 = 0.1 regimes")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange, then step (the default)")
    ap.add_argument("--overlap", action="store_true", help="N > 1: the exchange of a step travels while the step sorts its cells and builds "
                    "its interior geometry (late cells, DESIGN.md section 5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
