#!/bin/bash
# Run on the GPU box (via gpurun): launch list of the engine's own kernels + one --set full capture of the hot ones.
#   profiles/capture.sh <tag> [bench args...]
set -u
tag=${1:-cap}; shift || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 800 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" \
    > gpurun_out/${tag}_launches.log 2>&1
# core load = first batch; skip its launches, capture the warm-up batch of the bench workload
ncu --set full --clock-control none --import-source on \
    -k 'regex:^(k_rebalance_p|k_rebalance|k_rebalance_small|k_os_pass|k_locate)$' --launch-skip ${SKIP:-7} -c ${COUNT:-7} \
    -f -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" \
    > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out/
