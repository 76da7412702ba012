#!/bin/bash
# usage: scripts/gpu_final.sh TAG   (one GPU, through gpurun): the evidence of a round -- GPU tests, both bench arms, the ncu
# launch list of the bench command and one --set full capture per hot kernel (each after its own plain run exited 0)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"; cat gpurun_out/${TAG}_bench_reference.json | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 3 --skip-config4 --skip-regimes > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err \
  && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --skip-config4 --skip-regimes > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_launch.log | cut -c1-200
python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_plain.log 2>&1 || { cat gpurun_out/${TAG}_plain.log; exit 1; }
for K in k_clip k_force k_nbrs; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/${TAG}_$K python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$K.log
done
