# C3/C5 quick check + launch list and full captures of the sort pass / locate kernels on C4
set -u
mkdir -p gpurun_out
REPS=1 C4=0 C3=1 C5=1 bash benchmarks/run_ab2.sh r2q cur:PPCSR_X=1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 400 --csv \
    --log-file gpurun_out/r2q_launches.csv python bench.py --steps 1 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r2q_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^(k_os_pass|k_locate|k_build_keys|k_gather_inserts)$' --launch-skip 12 -c 10 \
    -f -o gpurun_out/r2q_full python bench.py --steps 1 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2q_full.log 2>&1
ls -la gpurun_out/r2q*
