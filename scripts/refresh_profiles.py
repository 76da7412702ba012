#!/usr/bin/env python
"""Copy the evidence of a scripts/gpu_final.sh run (gpurun_out/<tag>_*) into profiles/: python scripts/refresh_profiles.py r25"""
import csv, io, json, shutil, subprocess, sys
tag = sys.argv[1]
G, P = "gpurun_out/", "profiles/"
for k in ("k_hull", "k_ring", "k_force"):
    subprocess.run([sys.executable, "scripts/ncu_summary.py", f"{G}{tag}_{k}.ncu-rep", f"{P}r1_final_{k}.txt"], stdout=subprocess.DEVNULL)
shutil.copy(f"{G}{tag}_launches.csv", f"{P}r1_final_bench_launches.csv")
shutil.copy(f"{G}{tag}_bench.json", f"{P}r1_final_bench_1gpu.json")
shutil.copy(f"{G}{tag}_bench_reference.json", f"{P}r1_final_bench_reference.json")
shutil.copy(f"{G}{tag}_regimes.txt", f"{P}r1_final_regimes.txt")
rows = [r for r in csv.DictReader(l for l in open(f"{P}r1_final_bench_launches.csv") if not l.startswith("=="))]
agg = {}
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:60]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
with open(f"{P}r1_final_bench_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 (cold-cache, serialised: compare shares)\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:62s} launches {n:4d}  total {t/1e3:10.1f} us  share {100*t/tot:5.1f}%\n")
def traffic(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
    val = lambda name: float(v[h.index(name)]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[h.index(name)]]
    return val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
t = {k: traffic(f"{G}{tag}_{k}.ncu-rep") for k in ("k_hull", "k_ring", "k_force")}
json.dump({"cells": 1000000, "source": "ncu --set full --clock-control none, scripts/prof_step.py 1000000 4 (same workload as bench.py at 1 GPU)",
           "dram_bytes_per_launch": t["k_hull"], "kernel": "k_hull", "all_kernels": t}, open(f"{P}dominant_kernel_traffic.json", "w"), indent=1)
d = json.loads(open(f"{P}r1_final_bench_1gpu.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()}))
print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "fp64", d["roofline_fp64"]["frac"], d["roofline_fp64"]["whole_step"], "clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"])
r = json.loads(open(f"{P}r1_final_bench_reference.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
print(t)
# multi-GPU lines of scripts/gpu_scaling.sh, when the tag has them
import os
for n in (1, 2, 4, 8):
    for suffix, dst in (("", f"r1_final_bench_{n}gpu.json"), ("_1e8", f"r1_final_bench_{n}gpu_N1e8.json")):
        src = f"{G}{tag}_bench{n}{suffix}.json"
        if os.path.exists(src) and not (n == 1 and suffix == ""):      # the 1-GPU default line is the gpu_final.sh one
            shutil.copy(src, P + dst)
            d = json.loads(open(src).read().strip().splitlines()[-1])
            print(f"{n} GPUs{suffix}: {d['value']:.4g} cell-steps/s  {d['ms_per_step']:.3f} ms/step")
