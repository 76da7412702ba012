#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line: python scripts/bench_summary.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
r = lambda x: None if x is None else float(f"{x:.4g}")
print("value", r(d["value"]), "ms/step", r(d["ms_per_step"]), "repeats", [r(x) for x in d.get("ms_per_step_repeats", [])], "launches", d["gpu_launches"])
print("stages", {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()})
print("roofline", d["roofline"]["kernel"], "frac", r(d["roofline"]["frac"]), "whole-step", r(d["roofline"]["whole_step"]["frac"]),
      "| fp64", r(d["roofline_fp64"]["frac"]), "whole-step", r(d["roofline_fp64"]["whole_step"]["frac"]), "peak", r(d["roofline_fp64"]["peak"]))
print("e2e", r(d["e2e"]["value"]), "ref-style", r(d["e2e"].get("reference_style_loop", {}).get("value")), "| cpu", r(d["cpu_baseline"]["value"]), d["cpu_baseline"]["kind"])
c4 = d.get("config4_n1e8")
if c4:
    print("config4", r(c4.get("value")), "ms/step", r(c4.get("ms_per_step")), "cells/gpu", c4.get("cells_per_gpu"), c4.get("error", ""),
          "whole-step fp64", r((c4.get("roofline_fp64") or {}).get("whole_step", {}).get("frac")))
for k, v in (d.get("regimes") or {}).items():
    print("regime", k, r(v["value"]), "ms/step", r(v["ms_per_step"]), v["diagnostics"])
print("verify", d.get("verify"), "| diag", d.get("diagnostics"), "| clocks", d.get("clocks"))
