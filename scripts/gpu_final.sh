#!/bin/bash
# Final evidence run on the GPU box: tests, bench, ncu launch list of the bench command, ncu --set full of the three hot kernels.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_plain.log 2>&1 && for k in k_hull k_ring k_force; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 2 -c 1 -o gpurun_out/${TAG}_$k python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_ncu_$k.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$k.log; done
for args in "1000000 0.1" "1000000 0.3183" "1000000 1.0 30 jammed" "4000000 1.0 10" "16000000 1.0 5"; do python scripts/ab_stage.py $args; done > gpurun_out/${TAG}_regimes.txt 2>&1; cat gpurun_out/${TAG}_regimes.txt
