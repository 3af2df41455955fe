set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:^(k_pagerank_push_leaves|k_locate)$' --launch-skip 1 -c 3 \
    -f -o gpurun_out/r2u_full python bench.py --steps 1 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2u_full.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2u_full.ncu-rep
