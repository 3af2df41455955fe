# final profiles of the round: launch list of the C4 pipeline + one --set full capture of the sort pass, locate, gather and key builder
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 400 --csv \
    --log-file gpurun_out/r2fin_launches.csv python bench.py --steps 2 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r2fin_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^(k_os_pass|k_locate|k_build_keys|k_gather_inserts)$' --launch-skip 12 -c 10 \
    -f -o gpurun_out/r2fin_full python bench.py --steps 1 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2fin_full.log 2>&1
ls -la gpurun_out/r2fin*
