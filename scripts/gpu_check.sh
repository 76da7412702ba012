#!/bin/bash
# usage: scripts/gpu_check.sh TAG [bench args]   (run on the GPU box through gpurun): GPU tests + one bench line, summarised
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
SECONDS=0; timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python scripts/bench_summary.py gpurun_out/${TAG}_bench.json
tail -4 gpurun_out/${TAG}_bench.err
