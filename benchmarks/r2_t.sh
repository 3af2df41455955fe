# ncu captures of the rank-dense rebalance kernel on C4 / C2 / C3 / C5 (one launch each) + launch list of the C4 pipeline
set -u
mkdir -p gpurun_out
for cfg in C4 C2 C3 C5; do
  ncu --set full --clock-control none --import-source on -k 'regex:^k_rebalance_m$' --launch-skip 1 -c 1 \
      -f -o gpurun_out/r2t_${cfg}_full python bench.py --config $cfg --steps 1 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 \
      > gpurun_out/r2t_${cfg}_full.log 2>&1
  echo "$cfg ncu exit $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 400 --csv \
    --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 1 --only-headline --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r2t_launches.log 2>&1
ls -la gpurun_out/r2t*
