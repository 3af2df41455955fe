#!/bin/bash
# round 2: 8 GPUs -- C4 strong scaling (bench.py as the driver launches it) and the C5 batch-size sweep
tag=${1:-r2mg8}; N=${2:-8}
mkdir -p gpurun_out
bash benchmarks/r2_mg.sh $tag $N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  benchmarks/sweep.py --scale 24 --max-batch 100000000 --reps 4 > gpurun_out/${tag}_sweep_n$N.jsonl 2> gpurun_out/${tag}_sweep_n$N.err; echo "sweep exit $?"
cut -c1-330 gpurun_out/${tag}_sweep_n$N.jsonl
tail -3 gpurun_out/${tag}_sweep_n$N.err | cut -c1-300
