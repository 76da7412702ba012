#!/bin/bash
# usage: scripts/gpu_scaling.sh TAG N   (through gpurun --gpus N): one bench line on N GPUs, one process per GPU
TAG=$1; N=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/${TAG}_bench${N}.json 2> gpurun_out/${TAG}_bench${N}.err; echo "rc=$?"
python scripts/bench_summary.py gpurun_out/${TAG}_bench${N}.json | grep -v "^regime"
