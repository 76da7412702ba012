#!/bin/bash
# usage: scripts/gpu_check.sh TAG [ncu-kernel-regex]   (run on the GPU box through gpurun)
TAG=$1; KRE=${2:-}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g  ms/step %.3f  stages %s  e2e %.4g  cpu %.4g (%s)  fp64frac %.3f" % (d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}, d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["roofline_fp64"]["frac"]))
PY
tail -3 gpurun_out/${TAG}_bench.err
if [ -n "$KRE" ]; then
  python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 1 -o gpurun_out/${TAG}_prof python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
fi
