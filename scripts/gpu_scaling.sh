#!/bin/bash
# Multi-GPU bench lines: bash scripts/gpu_scaling.sh TAG N [small]   (run under `gpurun --gpus N`; "small": skip N = 10^8).
# Writes gpurun_out/TAG_bench<N>.json (10^6 cells per GPU, the driver's scaling run) and TAG_bench<N>_1e8.json (N = 10^8 in total).
TAG=${1:-scale}; N=${2:-2}
mkdir -p gpurun_out
run() {  # out, extra args
  if [ "$N" = 1 ]; then timeout 900 python bench.py --gpus 1 "${@:2}" > gpurun_out/$1.json 2> gpurun_out/$1.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "${@:2}" > gpurun_out/$1.json 2> gpurun_out/$1.err; fi
  echo "$1 rc=$?"; tail -1 gpurun_out/$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], d['config']['workload'][:40])"
}
run ${TAG}_bench${N} --steps 100 --warmup 10
[ "$3" = small ] || run ${TAG}_bench${N}_1e8 --cells-per-gpu $((100000000 / N)) --steps 10 --warmup 3
