set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:^(k_locate|k_gather_inserts)$' --launch-skip 2 -c 2 \
    -f -o gpurun_out/r2ac_full python bench.py --steps 1 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2ac_full.log 2>&1
echo "ncu exit $?"
