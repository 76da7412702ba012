#!/usr/bin/env python
"""Copy the evidence of a scripts/gpu_final.sh run (gpurun_out/<tag>_*) into profiles/r2_*: python scripts/refresh_profiles.py r2i"""
import csv, io, json, os, shutil, subprocess, sys
tag = sys.argv[1]
G, P = "gpurun_out/", "profiles/"
KERNELS = ("k_clip", "k_force", "k_nbrs")
for k in KERNELS:
    subprocess.run([sys.executable, "scripts/ncu_summary.py", f"{G}{tag}_{k}.ncu-rep", f"{P}r2_{k}.txt"], stdout=subprocess.DEVNULL)
shutil.copy(f"{G}{tag}_launches.csv", f"{P}r2_bench_launches.csv")
shutil.copy(f"{G}{tag}_bench.json", f"{P}r2_bench_1gpu.json")
shutil.copy(f"{G}{tag}_bench_reference.json", f"{P}r2_bench_reference.json")
rows = [r for r in csv.DictReader(l for l in open(f"{P}r2_bench_launches.csv") if not l.startswith("=="))]
agg = {}
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:60]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
with open(f"{P}r2_bench_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --skip-config4 --skip-regimes\n"
            "# (cold-cache, serialised: compare SHARES with roofline.stage_ms_per_step of the bench line, not absolute times)\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:62s} launches {n:4d}  total {t/1e3:10.1f} us  share {100*t/tot:5.1f}%\n")
def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    val = lambda name: float(v[h.index(name)]) * scale.get(u[h.index(name)], 1)
    return {"dram": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "warp_inst": val("smsp__inst_executed.sum"),
            "issue_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1_wavefront_pct": val("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
            "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio")}
m = {k: raw_metrics(f"{G}{tag}_{k}.ncu-rep") for k in KERNELS}
t = {k: m[k]["dram"] for k in KERNELS}
json.dump({"cells": 1000000, "source": "ncu --set full --clock-control none, scripts/prof_step.py 1000000 4 (same workload as bench.py at 1 GPU)",
           "dram_bytes_per_launch": t["k_clip"], "kernel": "k_clip", "all_kernels": t, "step_total_of_these": sum(t.values()),
           "warp_instructions_per_launch": {k: m[k]["warp_inst"] for k in KERNELS},
           "under_ncu_pct_of_peak": {k: {q: m[k][q] for q in ("issue_pct", "l1_wavefront_pct", "fp64_pipe_pct", "lanes_per_inst")} for k in KERNELS}},
          open(f"{P}dominant_kernel_traffic.json", "w"), indent=1)
d = json.loads(open(f"{P}r2_bench_1gpu.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()})
r = json.loads(open(f"{P}r2_bench_reference.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], "ratio e2e", d["e2e"]["value"] / r["value"])
print({k: round(v / 1e6, 1) for k, v in t.items()}, "MB; launch shares:"); print(open(f"{P}r2_bench_launch_shares.txt").read())
for n in (2, 4, 8):
    src = f"{G}{tag}_bench{n}.json"
    if os.path.exists(src):
        shutil.copy(src, f"{P}r2_bench_{n}gpu.json")
        d = json.loads(open(src).read().strip().splitlines()[-1])
        print(f"{n} GPUs: {d['value']:.4g} cell-steps/s {d['ms_per_step']:.3f} ms/step; config4 {d['config4_n1e8']['value']:.4g}; verify {d['verify']['ok']} bit_identical {d['verify']['bit_identical']}")
