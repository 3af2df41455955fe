# one --set full capture of the rebalance kernel of the warm-up batch (C2 by default); usage: run_ncu.sh <tag> <kernel regex> [bench args]
set -u
tag=$1; shift
kre=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:^(${kre})\$" --launch-skip ${SKIP:-1} -c ${COUNT:-1} \
    -f -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${tag}_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${tag}_full.log; ls -la gpurun_out/${tag}_full.ncu-rep
